"""Deterministic mode (cgs_model_set_deterministic): log-likelihood trace against the frozen oracle traces of
tests/golden/trace_c2_4000cells.npz and time per sweep, for several numbers of sub-sweeps.
usage: det_mode_study.py [K ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pycistopic_b200 import Corpus, DeviceModel
from pycistopic_b200.synth import planted_topic_tokens
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
g = np.load(os.path.join(ROOT, "tests", "golden", "trace_c2_4000cells.npz"))
D, W = int(g["D"]), int(g["W"])
cell, region = planted_topic_tokens(D, W, float(g["density"]), seed=int(g["data_seed"]), device=torch.device("cpu"))
assert np.array_equal(np.array([cell.numel(), int(cell.sum()), int(region.sum())]), g["checksum"])
dev = torch.device("cuda:0")
cnt = torch.bincount(cell, minlength=D)
cell_ptr = torch.zeros(D + 1, dtype=torch.int64); cell_ptr[1:] = torch.cumsum(cnt, 0)
cell_ptr, tok = cell_ptr.to(dev), region.to(torch.int32).to(dev)
corpus = Corpus.from_device_tokens(cell_ptr.data_ptr(), tok.data_ptr(), D, W, tok.numel(), keepalive=(cell_ptr, tok))
for K in [int(k) for k in (sys.argv[1:] or ["10", "50"])]:
    ref = g[f"mean_K{K}"]; n_iter = len(ref) - 1
    for parts in (0, 1, 2, 4, 8, 16, 32):
        traces = []
        for seed in (555, 1, 2, 3):
            m = DeviceModel(corpus, K, 50.0 / K, 0.1, seed=seed)
            if parts: m.set_deterministic(parts)
            m.init(); tr = []
            for _ in range(n_iter):
                tr.append(m.loglikelihood()); m.sweep(1)
            tr.append(m.loglikelihood()); traces.append(tr)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(); m.sweep(5); m.sync(); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            m.close()
        rel = (np.mean(traces, axis=0) - ref) / np.abs(ref) * 100
        print(f"K={K} parts={parts:2d}: {ms:.3f} ms/sweep, max |rel| {np.abs(rel).max():.2f} % at sweep {int(np.abs(rel).argmax())}; "
              f"rel% at 4,8,12,16,20,24: {np.round(rel[[4, 8, 12, 16, 20, 24]], 2)}", flush=True)
