"""Short program for ncu: the fused sweep + exchange launch (sweep_kernel<.., XCH = true>) on one GPU -- C2-shaped data,
one K-topic model whose exchange partner is itself (one rank), a few overlapped sweeps."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
K = int(sys.argv[1]) if len(sys.argv) > 1 else 50
nsweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
import bench
from pycistopic_b200 import Corpus, DeviceModel
dev = torch.device("cuda:0")
D, W, cell_ptr, tok = bench.make_workload_on_device(dev)
N = tok.numel()
corpus = Corpus.from_device_tokens(cell_ptr.data_ptr(), tok.data_ptr(), D, W, N, keepalive=(cell_ptr, tok))
per_region = torch.bincount(tok, minlength=W).to(torch.int64)
split = int(torch.searchsorted(torch.cumsum(per_region, 0), torch.tensor([N // 2], device=dev)).item()) + 1
corpus.set_region_split(split)
m = DeviceModel(corpus, K, 50.0 / K, 0.1, seed=555)
m.init()
m.exchange_open([m.exchange_arena()[0]], 0)
m.exchange_sync()
m.sweep_overlapped(nsweeps)
m.exchange_flush()
m.sync()
assert m.verify_counts() == {"n_dk": 0, "n_wk": 0, "n_k": 0} and m.exchange_timeouts() == 0
print("done", K, N, split, m.loglikelihood())
