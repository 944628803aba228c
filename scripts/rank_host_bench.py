import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from pycistopic_b200 import diff_features as pdf
rng = np.random.default_rng(0)
x = rng.integers(0, 400, (200000, 2048)).astype(np.int32)
for i in range(4):
    t = time.perf_counter(); r = pdf.rank_columns(x, seed=1); print(f"rank_columns host 200000 x 2048: {time.perf_counter() - t:.3f} s", flush=True)
y = rng.integers(0, 400, (200000, 4096)).astype(np.int32)
for i in range(2):
    t = time.perf_counter(); r = pdf.rank_columns(y[:, :3000], seed=1); print(f"strided 200000 x 3000 of 4096: {time.perf_counter() - t:.3f} s", flush=True)
