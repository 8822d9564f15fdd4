mkdir -p gpurun_out/final2
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/final2/gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/final2/gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final2/smoke.log 2>&1; echo "rc=$?" >> gpurun_out/final2/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/final2/bench_1gpu.json 2> gpurun_out/final2/bench_1gpu.err; echo "rc=$?" >> gpurun_out/final2/bench_1gpu.err
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/final2/bench_reference.json 2>> gpurun_out/final2/bench_1gpu.err
