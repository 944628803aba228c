"""The reference's own (dense) workflow on a C2-shaped model through the drop-in API, host arrays in and out:
impute_accessibility -> normalize_scores -> find_highly_variable_features -> find_diff_features -> make_rankings (a block
of cells).  usage: dense_paths_bench.py [R C K groups]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
import logging; logging.disable(logging.INFO)
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

class Model:
    cell_topic = pd.DataFrame(b, index=[f"Topic{i + 1}" for i in range(K)], columns=cells)
    topic_region = pd.DataFrame(a, index=regions, columns=[f"Topic{i + 1}" for i in range(K)])

class Obj:
    selected_model = Model()
    cell_names, region_names = cells, regions
    cell_data = pd.DataFrame({"celltype": np.array([f"type{g:02d}" for g in range(G)])[labels]}, index=cells)

def timed(label, fn):
    t = time.perf_counter(); out = fn(); dt = time.perf_counter() - t
    print(f"{label}: {dt:.2f} s", flush=True)
    return out

pdf.impute_accessibility(Obj(), selected_regions=regions[:2000], scale_factor=10 ** 6)   # warm-up (library, pool)
imputed = timed(f"impute_accessibility {R} x {C} x K={K} -> int32 host matrix", lambda: pdf.impute_accessibility(Obj(), scale_factor=10 ** 6))
print(f"   {imputed.mtx.nbytes / 1e9:.1f} GB, {len(imputed.feature_names)} regions kept")
normalized = timed("normalize_scores -> float64 host matrix", lambda: pdf.normalize_scores(imputed, 10 ** 4))
print(f"   {normalized.mtx.nbytes / 1e9:.1f} GB")
variable = timed("find_highly_variable_features", lambda: pdf.find_highly_variable_features(normalized, plot=False))
print(f"   {len(variable)} variable regions")
dars = timed(f"find_diff_features, {G} one-vs-rest contrasts, all regions", lambda: pdf.find_diff_features(Obj(), imputed, "celltype"))
print(f"   markers per contrast: {[len(v) for v in dars.values()]}")
block = pdf.CistopicImputedFeatures(np.ascontiguousarray(imputed.mtx[:, :2048]), imputed.feature_names, cells[:2048], "b")
timed("make_rankings, 2 048 cells", lambda: block.make_rankings(seed=123))
