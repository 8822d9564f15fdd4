mkdir -p gpurun_out/final
O=gpurun_out/final
L=$(python - <<'PY'
import sys
sys.path.insert(0, ".")
from gpras_b200.engine import ExactGP
from gpras_b200.synth import make_gp_data, fixed_theta
n, d, p = 8192, 32, 32
data = make_gp_data(n, d, p, 0, seed=0)
v, s, ls = fixed_theta(d, True)
gp = ExactGP("Matern52", n, d, p); gp.set_data(data.x, data.y)
th = gp.theta_vector(v, s, ls); gp.lml_grad(th); gp.lml_grad(th)
print(gp.last_launches())
PY
)
echo "launches per eval: $L" > $O/eval_launches.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed
timeout 900 ncu --metrics $M --clock-control none --launch-skip $L -c $L --csv --log-file $O/ncu_eval_traffic.csv python tools/profile_eval.py 2 > $O/ncu_eval.log 2>&1
