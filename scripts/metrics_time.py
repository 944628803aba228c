"""Time the device-side metric block (cgs_model_topic_gram / _cell_mixture / _top_regions /
_codocument_frequencies / _loglik_compat) on a named shape after a few sweeps.  usage: metrics_time.py [shape] [K]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pycistopic_b200 import Corpus, DeviceModel
from pycistopic_b200.synth import SHAPES, planted_topic_tokens

shape = sys.argv[1] if len(sys.argv) > 1 else "T"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 50
D, W, dens = SHAPES[shape]
dev = torch.device("cuda:0")
cell, region = planted_topic_tokens(D, W, dens, seed=1234, device=dev)
order = torch.argsort(cell * W + region)
cell, region = cell[order], region[order]
cell_ptr = torch.zeros(D + 1, dtype=torch.int64, device=dev)
cell_ptr[1:] = torch.cumsum(torch.bincount(cell, minlength=D), 0)
tok = region.to(torch.int32).contiguous()
corpus = Corpus.from_device_tokens(cell_ptr.data_ptr(), tok.data_ptr(), D, W, tok.numel(), keepalive=(cell_ptr, tok))
model = DeviceModel(corpus, K, 50.0 / K, 0.1, seed=555)
model.init(); model.sweep(30); model.sync()
def T(name, f, reps=3):
    f(); best = 1e9
    for _ in range(reps):
        t = time.perf_counter(); r = f(); best = min(best, time.perf_counter() - t)
    print(f"{name:32s} {best*1e3:9.3f} ms", flush=True)
    return r
print(f"shape {shape}: {D} cells x {W} regions, {tok.numel()} tokens, K={K}")
T("cell_sizes (uncached)", lambda: (setattr(corpus, "_cell_sizes", None), corpus.cell_sizes())[1])
T("topic_gram", model.topic_gram)
T("cell_mixture", model.cell_mixture)
top = T("top_regions(20)", lambda: model.top_regions(20))
T("codocument_frequencies", lambda: model.codocument_frequencies(top))
T("loglik_compat", lambda: model.loglikelihood_compat(50.0, 0.1))
T("metric_ingredients (all)", model.metric_ingredients)
T("frames export (W x K, K x D f64)", model.frames)
T("counts+distributions export", lambda: (model.counts(), model.distributions()))
model.close(); corpus.close()
