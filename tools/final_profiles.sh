#!/bin/bash
# Round-end evidence run on the GPU box (one gpurun call): bench lines, ncu launch list of the bench command, per-launch
# traffic of one whole evaluation, and `--set full` captures of the main kernels.  Outputs -> gpurun_out/final/.
set -u
O=gpurun_out/final; mkdir -p $O
python bench.py --steps 10 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err || echo "bench failed"
python bench.py --impl reference --steps 10 --warmup 3 > $O/bench_reference.json 2>> $O/bench_1gpu.err || echo "ref failed"
python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline > $O/bench_short_plain.json 2>> $O/bench_1gpu.err || echo "short bench failed"
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $O/ncu_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --quick --no-cpu-baseline > $O/ncu_bench.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
python tools/profile_eval.py 2 > $O/eval_plain.log 2>&1
# launches of ONE evaluation = gpu_launches / steps of the bench line; skip the first (eager) evaluation, capture the second
L=$(python -c "import json; d=json.loads(open('$O/bench_short_plain.json').read().strip().splitlines()[-1]); print(d['gpu_launches'] // d['steps'])")
ncu --metrics $M --clock-control none --launch-skip $L -c $L --csv --log-file $O/ncu_eval_traffic.csv python tools/profile_eval.py 2 > $O/ncu_eval.log 2>&1
# the TMA-streamed kernels and the general-variance expansion: per-launch time and DRAM throughput
python tools/profile_tma.py > $O/tma_plain.log 2>&1
ncu --metrics $M --clock-control none -k regex:'colstats|project_|metrics_plain|metrics_stream|cells_general|center_weight' --csv \
    --log-file $O/ncu_tma_kernels.csv python tools/profile_tma.py > $O/ncu_tma.log 2>&1
full() { # name regex skip count script
  ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:$2" --launch-skip $3 -c $4 -o $O/full_$1 -f \
      python ${5:-tools/profile_eval.py 2} > $O/ncu_full_$1.log 2>&1
  # the summaries are made from the raw page; the reports themselves are too big to bring back (64 MiB limit), keep two
  ncu -i $O/full_$1.ncu-rep --page raw --csv > $O/full_$1.csv 2>/dev/null
  case $1 in lauum|rest0) ;; *) rm -f $O/full_$1.ncu-rep ;; esac
}
[ -n "${SKIP_FULL:-}" ] && { ls -la $O; exit 0; }
full lauum 'Li3ELi1EEELb1ELb1' 1 1
full rest0 'Li16ELi3ELi2EEELb0ELb0' 100 1
full trtri_top 'Li3ELi1EEELb0ELb1' 22 2
full leaf_potrf 'leaf_potrf_kernel' 64 1
full trsm 'trsm_panel_kernel' 63 1
full cov 'cov_kernel' 1 1
full grad 'grad_kernel' 1 1
full colstats_tma 'colstats_tma' 1 1 tools/profile_tma.py
full project_tma16 'project_tma_kernelILi16' 0 1 tools/profile_tma.py
full metrics_plain_tma 'metrics_plain_tma' 0 1 tools/profile_tma.py
full cells_general 'cells_general' 0 1 tools/profile_tma.py
ls -la $O
