"""Oracle log-likelihood traces on a benchmark-SHAPED corpus (C2's 200 000 regions at 3 % density, 4 000 of
its 10 000 cells: 2.6e7 tokens, cells of 6 500 tokens), so that the device sampler can be compared with the
sequential reference at the kernel shapes and concurrency the benchmark uses -- the C1 / 300-cell corpora of the
other parity tests never reach them.  The matrix comes from the torch-CPU path of the generator (seed 1234;
reproducible for a given torch build -- the file stores checksums so that a consumer can tell), the oracle runs
8 seeds per K with refresh = 1.  Takes ~10 min on 8 cores; run in the authoring container:
    python tests/golden/make_fullscale_trace.py
writes tests/golden/trace_c2_4000cells.npz (mean and standard error of the trace per K; a few hundred bytes)."""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import lda_oracle as lo  # noqa: E402
from pycistopic_b200.synth import planted_topic_tokens  # noqa: E402

D, W, DENS, SEED = 4000, 200_000, 0.03, 1234
CASES = {10: 24, 50: 24}  # K -> sweeps
SEEDS = [555, 1, 2, 3, 4, 5, 6, 7]


def tokens():
    """(cell, region) int64 arrays sorted by (cell, region) and their checksums."""
    import torch
    cell, region = planted_topic_tokens(D, W, DENS, seed=SEED, device=torch.device("cpu"))
    cell, region = cell.numpy(), region.numpy()
    return cell, region, np.array([len(cell), int(cell.sum()), int(region.sum())], np.int64)


def matrix():
    cell, region, check = tokens()
    X = sp.csr_matrix((np.ones(len(cell), np.int32), (cell, region)), shape=(D, W), dtype=np.int32)
    return X, check


if __name__ == "__main__":
    t = time.time()
    X, check = matrix()
    print(f"matrix {D} x {W}: {X.nnz} tokens ({time.time() - t:.0f} s)", flush=True)
    out = {"D": D, "W": W, "density": DENS, "data_seed": SEED, "seeds": np.array(SEEDS), "tokens": X.nnz, "checksum": check}
    for K, n_iter in CASES.items():
        def one(s):
            o = lo.OracleLDA(K, n_iter=n_iter, alpha=50.0 / K, eta=0.1, random_state=s, refresh=1).fit(X)
            return np.array(o.loglikelihoods_ + [o.final_loglikelihood_])
        t = time.time()
        with ThreadPoolExecutor(len(SEEDS)) as ex:
            tr = np.array(list(ex.map(one, SEEDS)))
        out[f"mean_K{K}"] = tr.mean(axis=0)
        out[f"se_K{K}"] = tr.std(axis=0, ddof=1) / np.sqrt(len(SEEDS))
        print(f"K={K}: {n_iter} sweeps x {len(SEEDS)} seeds in {time.time() - t:.0f} s; ll0 {tr[0, 0]:.0f} -> {tr[:, -1].mean():.0f}",
              flush=True)
    np.savez(os.path.join(HERE, "trace_c2_4000cells.npz"), **out)
