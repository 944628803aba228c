#!/usr/bin/env python
"""Two chunks of the C5 leg (bench.py impute_c5) on one GPU, for an ncu capture of the second pass:
normalize_log1p_kernel + row_mean_var_kernel on a 20 000 x 200 000 int32 chunk.  Prints the wall times of the two
ABI calls; run it under `ncu -k regex:normalize_log1p|row_mean_var -c 2 --set full` for the pipe utilisation."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from pycistopic_b200.sharded_consumers import DeviceBackend  # noqa: E402

R, K, C = int(os.environ.get("C5P_REGIONS", 40000)), 100, 200000
rng = np.random.default_rng(4321)
A = rng.random((R, K), dtype=np.float32)
A /= A.sum(0, keepdims=True) * (1_000_000 / R)   # column sums as if R were 10^6: values of the C5 magnitude
B = rng.random((K, C), dtype=np.float32)
B /= B.sum(0, keepdims=True)
backend = DeviceBackend(device=0, chunk_size=20000)
for rep in range(2):
    t0 = time.perf_counter()
    totals, flags = backend.column_totals(A, B, 1e6, True)
    t1 = time.perf_counter()
    mean, var = backend.row_stats(A, B, 1e6, True, 1e4, totals)
    t2 = time.perf_counter()
    print(f"rep {rep}: column_totals {t1 - t0:.3f} s, normalized row stats {t2 - t1:.3f} s "
          f"({R // 20000} chunks of 20000 x {C}); kept {int(flags.sum())}, mean[0] {mean[0]:.6g}", flush=True)
