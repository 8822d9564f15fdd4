mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -rP > gpurun_out/t12_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/t12_gpu.log
python - > gpurun_out/kmeans_timing.txt 2>&1 <<'PY'
import sys, time
sys.path.insert(0, ".")
import numpy as np
from sklearn.cluster import KMeans, kmeans_plusplus
from gpras_b200.engine import kmeans_lloyd
rng = np.random.default_rng(0)
for n, d, m in [(5000, 10, 50), (8192, 32, 300)]:
    x = rng.normal(size=(n, d))
    t0 = time.perf_counter(); km = KMeans(n_clusters=m, random_state=0, n_init="auto").fit(x); t_sk = time.perf_counter() - t0
    mu = x.mean(0)
    t0 = time.perf_counter(); seeds, _ = kmeans_plusplus(x - mu, m, random_state=0); t_seed = time.perf_counter() - t0
    kmeans_lloyd(x - mu, seeds)
    t0 = time.perf_counter(); c, lab, inertia, it = kmeans_lloyd(x - mu, seeds); t_dev = time.perf_counter() - t0
    print(f"N={n} D={d} M={m}: sklearn KMeans.fit {t_sk*1e3:.1f} ms ({km.n_iter_} its); k-means++ seeding {t_seed*1e3:.1f} ms + device Lloyd {t_dev*1e3:.1f} ms ({it} its); inertia rel diff {abs(inertia-km.inertia_)/km.inertia_:.2e}")
PY
