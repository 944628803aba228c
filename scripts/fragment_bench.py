"""Time the fragments-in-regions count matrix on a PBMC-sized fragments table.
usage: fragment_bench.py [n_fragments n_regions n_cells]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, pandas as pd
from pycistopic_b200.fragment_matrix import count_fragments_in_regions

F = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000
R = int(float(sys.argv[2])) if len(sys.argv) > 2 else 200_000
C = int(float(sys.argv[3])) if len(sys.argv) > 3 else 10_000
rng = np.random.default_rng(0)
n_chrom, chrom_len = 24, 120_000_000
chroms = np.array([f"chr{i + 1}" for i in range(n_chrom)])
rc = rng.integers(0, n_chrom, R); rs = rng.integers(0, chrom_len, R)
regions = pd.DataFrame({"Chromosome": pd.Categorical.from_codes(rc, chroms), "Start": rs, "End": rs + 500})
# half of the fragments fall in peaks (FRiP 0.5): around a random region, the rest anywhere
pick = rng.integers(0, R, F)
in_peak = rng.random(F) < 0.5
fc = np.where(in_peak, rc[pick], rng.integers(0, n_chrom, F)).astype(np.int8)
fs = np.where(in_peak, rs[pick] + rng.integers(-100, 400, F), rng.integers(0, chrom_len, F))
fragments = pd.DataFrame({"Chromosome": pd.Categorical.from_codes(fc, chroms), "Start": fs, "End": fs + rng.integers(30, 600, F),
                          "Name": pd.Categorical.from_codes(rng.integers(0, C, F), [f"BC{j:06d}" for j in range(C)])})
print(f"{F} fragments, {R} regions, {C} barcodes", flush=True)
for rep in range(2):
    os.environ["CGS_B200_TIMING"] = "1"
    tm = {}
    t = time.perf_counter(); m, rn, cn = count_fragments_in_regions(fragments, regions, timings=tm); dt = time.perf_counter() - t
    print({k: round(v, 3) for k, v in tm.items()})
    print(f"count_fragments_in_regions: {dt:.2f} s wall ({F / dt:.3e} fragments/s incl. host factorisation and copies); "
          f"matrix {m.shape}, nnz {m.nnz}, sum {m.sum()}", flush=True)
