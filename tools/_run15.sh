mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/t15_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t15_gpu.log
timeout 300 python tools/bench_pre_metrics.py --transform-only 2>/dev/null | grep transform > gpurun_out/proj_after_pitch_fix.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:'project_tma' --csv --log-file gpurun_out/ncu_project_after_fix.csv python tools/profile_tma.py > /dev/null 2>&1
for c in 1 2 3 4 6; do
  timeout 300 python bench.py --quick --no-cpu-baseline --steps 24 --concurrent $c > gpurun_out/bench_conc$c.json 2>/dev/null
done
