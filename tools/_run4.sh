mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "potrf or golden or matches_oracle or sgpr or concurrent" > gpurun_out/t4_parity.log 2>&1; echo "rc=$?" >> gpurun_out/t4_parity.log
for g in 1 2 3 4 6; do for m in 16 24 32; do for t in 1 2 4; do
  echo "== G=$g MIN_REM=$m TAIL=$t" >> gpurun_out/chain_sweep4.txt
  CHAIN_ONLY_POTRF=1 GPRAS_B200_PANEL_GROUP=$g GPRAS_B200_PAIR_MIN_REM=$m GPRAS_B200_TAIL_GROUP=$t timeout 60 ./tools/microbench/chain_timing 2>&1 | grep "potrf n=" >> gpurun_out/chain_sweep4.txt
done; done; done
CHAIN_TIMELINE=1 CHAIN_ONLY_POTRF=1 GPRAS_B200_PANEL_GROUP=4 GPRAS_B200_PAIR_MIN_REM=24 timeout 120 ./tools/microbench/chain_timing > gpurun_out/timeline4_g4.txt 2>&1
