"""Short program for ncu: one rank's share (1/8 of the cells) of the atlas-scale C4 matrix, K = 100, so that
n_wk (1 M regions x 100 topics x 4 B = 400 MB) does not fit in L2 and the sweep really streams rows from HBM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pycistopic_b200 import Corpus, DeviceModel
from pycistopic_b200.synth import SHAPES, planted_topic_tokens
K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
nsweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda:0")
D, W, dens = SHAPES["C4"]
cb, ce = 0, D // 8
cell, region = planted_topic_tokens(D, W, dens, seed=1234, device=dev, cell_range=(cb, ce))
cnt = torch.bincount(cell - cb, minlength=ce - cb)
cp = torch.zeros(ce - cb + 1, dtype=torch.int64, device=dev); cp[1:] = torch.cumsum(cnt, 0)
tok = region.to(torch.int32).contiguous()
del cell, region
N = tok.numel()
corpus = Corpus.from_device_tokens(cp.data_ptr(), tok.data_ptr(), ce - cb, W, N, keepalive=(cp, tok))
m = DeviceModel(corpus, K, 50.0 / K, 0.1, seed=555)
m.init()
m.sweep(nsweeps)
m.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
print("done", K, N, m.loglikelihood())
