"""Short program for ncu: C2-shaped data, a few sweeps of one model (K from argv)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.argv = [sys.argv[0]] + sys.argv[1:]
K = int(sys.argv[1]) if len(sys.argv) > 1 else 50
nsweeps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
import bench
from pycistopic_b200 import Corpus, DeviceModel
dev = torch.device("cuda:0")
D, W, cell_ptr, tok = bench.make_workload_on_device(dev)
N = tok.numel()
corpus = Corpus.from_device_tokens(cell_ptr.data_ptr(), tok.data_ptr(), D, W, N, keepalive=(cell_ptr, tok))
m = DeviceModel(corpus, K, 50.0 / K, 0.1, seed=555)
m.init()
m.sweep(nsweeps)
m.sync()
print("done", K, N, m.loglikelihood())
