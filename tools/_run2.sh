mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/t2_parity.log 2>&1; echo "rc=$?" >> gpurun_out/t2_parity.log
timeout 300 ./tools/microbench/chain_timing > gpurun_out/chain_default.txt 2>&1
for g in 1 2 3 4; do for m in 20 28 36 48; do
  echo "== G=$g MIN_REM=$m" >> gpurun_out/chain_sweep.txt
  GPRAS_B200_PANEL_GROUP=$g GPRAS_B200_PAIR_MIN_REM=$m timeout 120 ./tools/microbench/chain_timing 2>&1 | grep "potrf n=" >> gpurun_out/chain_sweep.txt
done; done
timeout 300 python tools/microbench/lib_bars.py > gpurun_out/lib_bars.json 2>&1
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_small.py gp > gpurun_out/memcheck_gp.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py gp > gpurun_out/racecheck_gp.log 2>&1
timeout 600 python bench.py > gpurun_out/bench1.json 2> gpurun_out/bench1.err
