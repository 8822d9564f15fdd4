mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -rP > gpurun_out/t7_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t7_gpu.log
timeout 900 python bench.py > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "rc=$?" >> gpurun_out/bench7.err
