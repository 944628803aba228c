"""Where the imputation kernel's output rate saturates: K sweep (K = 8 makes the MMA and the operand traffic
negligible) next to the pure-write rate of a 16 GB fill."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pycistopic_b200 import _lib
from pycistopic_b200._lib import c_vp, check
L = _lib.load()
dev = torch.device("cuda:0")
R, C = 20000, 200000
out = torch.empty(R, C, dtype=torch.int32, device=dev)
fl = torch.empty(R, dtype=torch.uint8, device=dev)
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ms = timeit(lambda: out.fill_(7))
print(f"torch fill_ 16 GB: {ms:.3f} ms -> {R*C*4/ms/1e6:.0f} GB/s", flush=True)
src = torch.empty(R, C // 2, dtype=torch.int32, device=dev)
dst = torch.empty(R, C // 2, dtype=torch.int32, device=dev)
ms = timeit(lambda: dst.copy_(src))
print(f"torch copy 8 GB -> 8 GB: {ms:.3f} ms -> {2*R*(C//2)*4/ms/1e6:.0f} GB/s (read+write)", flush=True)
del src, dst
for K in (8, 24, 50, 100, 200):
    g = torch.Generator(device=dev); g.manual_seed(1)
    A = torch.rand(R, K, device=dev, generator=g); A /= A.sum(0, keepdim=True)
    B = torch.rand(K, C, device=dev, generator=g); B /= B.sum(0, keepdim=True)
    for as_int in (1, 0):
        plan = c_vp()
        check(L.cgs_impute_plan_create(ctypes.byref(plan), 0, B.data_ptr(), 1, K, C, None))
        s = torch.cuda.current_stream().cuda_stream
        ms = timeit(lambda: check(L.cgs_impute_plan_chunk_device(plan, A.data_ptr(), R, 1e6, as_int, out.data_ptr(), fl.data_ptr(), s)))
        print(f"K={K:4d} as_int={as_int}: {ms:.3f} ms -> {R*C*4/ms/1e6:.0f} GB/s written", flush=True)
        check(L.cgs_impute_plan_destroy(plan))
