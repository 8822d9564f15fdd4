mkdir -p gpurun_out
CHAIN_TIMELINE=1 CHAIN_ONLY_POTRF=1 timeout 120 ./tools/microbench/chain_timing > gpurun_out/timeline_g2.txt 2>&1
CHAIN_TIMELINE=1 CHAIN_ONLY_POTRF=1 GPRAS_B200_PANEL_GROUP=1 timeout 120 ./tools/microbench/chain_timing > gpurun_out/timeline_g1.txt 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
CHAIN_ONLY_POTRF=1 GPRAS_B200_PANEL_GROUP=1 timeout 600 ncu --metrics $M --clock-control none --launch-skip 260 -c 260 --csv --log-file gpurun_out/ncu_potrf_g1.csv ./tools/microbench/chain_timing > gpurun_out/ncu_potrf_g1.log 2>&1
CHAIN_ONLY_POTRF=1 GPRAS_B200_PANEL_GROUP=4 GPRAS_B200_PAIR_MIN_REM=20 timeout 600 ncu --metrics $M --clock-control none --launch-skip 260 -c 260 --csv --log-file gpurun_out/ncu_potrf_g4.csv ./tools/microbench/chain_timing > gpurun_out/ncu_potrf_g4.log 2>&1
