mkdir -p gpurun_out
for g in 2 3 4 6; do for m in 12 20 28; do
  echo "== G=$g MIN_REM=$m" >> gpurun_out/chain_sweep6.txt
  CHAIN_ONLY_POTRF=1 GPRAS_B200_PANEL_GROUP=$g GPRAS_B200_PAIR_MIN_REM=$m timeout 60 ./tools/microbench/chain_timing 2>&1 | grep "potrf n=" >> gpurun_out/chain_sweep6.txt
done; done
CHAIN_TIMELINE=1 CHAIN_ONLY_POTRF=1 timeout 120 ./tools/microbench/chain_timing > gpurun_out/timeline6.txt 2>&1
timeout 100 ./tools/microbench/chain_timing 2>&1 | grep -E "potrf" > gpurun_out/chain6_default.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench6.json 2> gpurun_out/bench6.err
