mkdir -p gpurun_out
for w in 1 2 4 8; do
  echo "== waves $w" >> gpurun_out/proj_waves.txt
  GPRAS_B200_PROJ_WAVES=$w timeout 300 python tools/bench_pre_metrics.py --transform-only 2>/dev/null | grep transform >> gpurun_out/proj_waves.txt
done
echo "== waves 4 no TMA" >> gpurun_out/proj_waves.txt
GPRAS_B200_NO_TMA=1 timeout 300 python tools/bench_pre_metrics.py --transform-only 2>/dev/null | grep transform >> gpurun_out/proj_waves.txt
