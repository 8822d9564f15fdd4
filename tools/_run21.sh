mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_metrics.py tests/test_gpu_preprocess.py tests/test_gpu_parity_full.py -m gpu -q -x > gpurun_out/t21.log 2>&1; echo "rc=$?" >> gpurun_out/t21.log
