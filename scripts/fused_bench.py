"""Differential-feature workflow on a C2-shaped model (regions x cells x topics): dense path as the reference does it
(impute -> host matrix -> markers per contrast) against the factored path (nothing materialised), plus the reference's
own CPU arithmetic (scipy.stats.ranksums per region) timed on a sample.  usage: fused_bench.py [R C K groups]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
from pycistopic_b200 import diff_features as pdf

R = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 50
G = int(sys.argv[4]) if len(sys.argv) > 4 else 10
rng = np.random.default_rng(1)
a = rng.dirichlet(np.full(R, 0.1), size=K).T.astype(np.float32)
labels = rng.integers(0, G, C)
conc = np.full((G, K), 0.2)
for g in range(G):
    conc[g, (g * K // G):((g + 1) * K // G)] = 4.0
b = np.stack([rng.dirichlet(conc[g]) for g in labels]).T.astype(np.float32)
cells = [f"c{j}" for j in range(C)]
regions = [f"chr1:{j * 500}-{j * 500 + 500}" for j in range(R)]
names = np.array([f"type{g:02d}" for g in range(G)])[labels]

class Obj:
    cell_data = pd.DataFrame({"celltype": names}, index=cells)

import logging; logging.disable(logging.INFO)
devices = [int(d) for d in os.environ.get("CGS_DEVICES", "0").split(",")]   # regions sharded over these GPUs
print("devices:", devices, flush=True)
fact = pdf.FactoredImputedFeatures(a, b, regions, cells, 10 ** 6, 20000, "bench", device=devices)
fact.ranksum([([0], [1])])  # warm-up (library load, pool)
t = time.perf_counter(); res_f = pdf.find_diff_features(Obj(), fact, "celltype"); t_f = time.perf_counter() - t
print(f"factored: find_diff_features {G} one-vs-rest contrasts on {R} x {C} x K={K}: {t_f:.2f} s "
      f"({G * R / t_f:.3e} region-tests/s); markers per contrast: {[len(v) for v in res_f.values()]}", flush=True)
t = time.perf_counter(); sel = pdf.find_highly_variable_features(pdf.normalize_scores(fact, 10 ** 4, materialize=False), plot=False); t_v = time.perf_counter() - t
print(f"factored: normalize_scores(materialize=False) -> find_highly_variable_features: {t_v:.2f} s, {len(sel)} features", flush=True)
t = time.perf_counter(); rk = fact.make_rankings(seed=123); t_r = time.perf_counter() - t
print(f"factored: make_rankings, all {C} cells: {t_r:.2f} s ({rk.mtx.nbytes / 1e9:.1f} GB of int32 rankings)", flush=True)
for c in (0, C // 2, C - 1):   # every column a permutation
    assert np.array_equal(np.sort(rk.mtx[:, c]), np.arange(rk.mtx.shape[0]))
del rk
if os.environ.get("CGS_FUSED_ONLY"):
    sys.exit(0)
t = time.perf_counter(); dense = fact.to_dense(); t_i = time.perf_counter() - t
t = time.perf_counter(); res_d = pdf.find_diff_features(Obj(), dense, "celltype"); t_d = time.perf_counter() - t
print(f"dense:    impute_accessibility to host {t_i:.2f} s ({dense.mtx.nbytes / 1e9:.1f} GB) + find_diff_features {t_d:.2f} s", flush=True)
same = all(res_f[k].index.tolist() == res_d[k].index.tolist() for k in res_f)
print("identical marker tables:", same, flush=True)
# the reference's CPU arithmetic on a sample of regions of the first contrast
from scipy.stats import ranksums
fg = np.flatnonzero(names == "type00"); bg = np.flatnonzero(names != "type00")
n = 200
t = time.perf_counter()
for i in range(n):
    ranksums(dense.mtx[i, fg], dense.mtx[i, bg])
t_c = (time.perf_counter() - t) / n
print(f"scipy.stats.ranksums on one host core: {t_c * 1e3:.2f} ms per region -> {G * R * t_c:.0f} s for the same {G} x {R} tests "
      f"(x{G * R * t_c / t_f:.0f})", flush=True)
if R * C <= 2.5e9:
    t = time.perf_counter(); nf = fact.normalized(10 ** 4); t_n = time.perf_counter() - t
    print(f"factored normalize_scores: {t_n:.2f} s for {nf.mtx.nbytes / 1e9:.1f} GB of float64 output", flush=True)
