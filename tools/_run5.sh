mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "potrf or golden or matches_oracle or sgpr or concurrent" > gpurun_out/t5_parity.log 2>&1; echo "rc=$?" >> gpurun_out/t5_parity.log
timeout 100 ./tools/microbench/chain_timing 2>&1 | grep -E "trsm|potrf" > gpurun_out/chain5_default.txt
for g in 3 4; do for m in 16 24 32; do for w in 0 32 99; do
  echo "== G=$g MIN_REM=$m WIDE=$w" >> gpurun_out/chain_sweep5.txt
  CHAIN_ONLY_POTRF=1 GPRAS_B200_PANEL_GROUP=$g GPRAS_B200_PAIR_MIN_REM=$m GPRAS_B200_WIDE_COL_REM=$w timeout 60 ./tools/microbench/chain_timing 2>&1 | grep "potrf n=" >> gpurun_out/chain_sweep5.txt
done; done; done
CHAIN_TIMELINE=1 CHAIN_ONLY_POTRF=1 timeout 120 ./tools/microbench/chain_timing > gpurun_out/timeline5.txt 2>&1
