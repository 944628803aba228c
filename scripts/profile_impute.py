"""Short program for ncu: a few device-resident imputation chunks (R x K x C from argv, default 20000 100 200000)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pycistopic_b200 import _lib
from pycistopic_b200._lib import c_vp, check
R, K, C = [int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (20000, 100, 200000))]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 4
L = _lib.load()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(4321)
A = torch.rand(R, K, device=dev, generator=g); A /= A.sum(0, keepdim=True)
B = torch.rand(K, C, device=dev, generator=g); B /= B.sum(0, keepdim=True)
out = torch.empty(R, C, dtype=torch.int32, device=dev)
fl = torch.empty(R, dtype=torch.uint8, device=dev)
plan = c_vp()
check(L.cgs_impute_plan_create(ctypes.byref(plan), 0, B.data_ptr(), 1, K, C, None))
s = torch.cuda.current_stream().cuda_stream
for _ in range(n):
    check(L.cgs_impute_plan_chunk_device(plan, A.data_ptr(), R, 1e6, 1, out.data_ptr(), fl.data_ptr(), s))
torch.cuda.synchronize()
print("done", R, K, C, int(out[0, :8].sum().item()))
check(L.cgs_impute_plan_destroy(plan))
