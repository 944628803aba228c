"""Device sampler vs the frozen oracle traces of tests/golden/trace_c2_4000cells.npz for several concurrency
settings: usage trace_fullscale.py K [lanes,warps,max_blocks[,nk_div] ...]   (0 = automatic)"""
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
K = int(sys.argv[1])
SEEDS = [555, 1, 2, 3, 4, 5, 6, 7]
ref = g[f"mean_K{K}"]; n_iter = len(ref) - 1
print("oracle se%", np.round(g[f"se_K{K}"] / np.abs(ref) * 100, 2)[[4, 8, 12, 16, 20, 24]])
for cfg in (sys.argv[2:] or ["0,0,0"]):
    parts = [float(x) for x in cfg.split(",")]
    lanes, warps, mb = int(parts[0]), int(parts[1]), int(parts[2])
    nkd = parts[3] if len(parts) > 3 else 0
    traces = []
    for seed in SEEDS:
        m = DeviceModel(corpus, K, 50.0 / K, 0.1, seed=seed)
        if lanes: m.set_sweep_shape(lanes, warps)
        if mb: m.set_max_blocks(mb)
        if nkd: m.set_nk_staleness(nkd)
        m.init(); tr = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(n_iter):
            tr.append(m.loglikelihood()); m.sweep(1)
        tr.append(m.loglikelihood()); traces.append(tr); shape = m.sweep_shape
        s_ = torch.cuda.Stream(); 
        torch.cuda.synchronize(); e0.record(); m.sweep(5); m.sync(); e1.record(); torch.cuda.synchronize(); ms = e0.elapsed_time(e1) / 5
        m.close()
    traces = np.array(traces)
    rel = (traces.mean(axis=0) - ref) / np.abs(ref) * 100
    se = np.sqrt(traces.var(axis=0, ddof=1) / len(traces) + g[f"se_K{K}"] ** 2) / np.abs(ref) * 100
    idx = [4, 8, 12, 14, 16, 18, 20, 24]
    print(f"{cfg:14s} {shape} {ms:.3f} ms/sweep  max|rel| {np.abs(rel).max():.2f}%\n     rel%: {np.round(rel[idx],2)}\n     se% : {np.round(se[idx],2)}", flush=True)
