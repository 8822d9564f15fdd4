mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_preprocess.py tests/test_gpu_metrics.py -m gpu -q -x > gpurun_out/t9_pre_met.log 2>&1; echo "rc=$?" >> gpurun_out/t9_pre_met.log
timeout 600 python tools/bench_pre_metrics.py > gpurun_out/pre_metrics_tma.jsonl 2> gpurun_out/pre_metrics_tma.err
GPRAS_B200_NO_TMA=1 timeout 600 python tools/bench_pre_metrics.py > gpurun_out/pre_metrics_notma.jsonl 2> gpurun_out/pre_metrics_notma.err
