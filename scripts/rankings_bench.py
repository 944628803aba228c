"""Time the per-cell ranking kernel on a device-resident imputed block.  usage: rankings_bench.py [regions cells]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pycistopic_b200 import _lib
from pycistopic_b200._lib import c_vp, check

R = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
L = _lib.load()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
K = 50
tr = torch.distributions.Dirichlet(torch.full((R,), 0.05, device=dev)).sample((K,)).T.contiguous()   # regions x topics
ct = torch.distributions.Dirichlet(torch.full((K,), 0.3, device=dev)).sample((C,)).T.contiguous()    # topics x cells
x = (tr @ ct * 1e6).to(torch.int32)
out = torch.empty_like(x)
s = torch.cuda.current_stream().cuda_stream
for name, xin, code in (("int32 (imputed: trunc(1e6 p))", x, 0), ("float32 (no ties)", (tr @ ct).contiguous(), 1)):
    def run():
        check(L.cgs_rank_columns_device(0, code, c_vp(xin.data_ptr()), C, R, C, 123, 0, c_vp(out.data_ptr()), C, c_vp(s)))
    for _ in range(2): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(3): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{name}: {R} regions x {C} cells: {ms:.2f} ms per block, {R * C / ms * 1e3:.3e} values/s, "
          f"{C / ms * 1e3:.0f} cells/s, max value {int(xin.max())}", flush=True)
    r = out.cpu().numpy(); xs = xin.cpu().numpy()
    for c in (0, C - 1):
        inv = np.argsort(r[:, c]); assert np.array_equal(np.sort(r[:, c]), np.arange(R)) and (np.diff(xs[inv, c]) <= 0).all()
# the CPU closure of the reference (diff_features.py:294-306) on a few columns of the same block
xs = x[:, :4].cpu().numpy()
rng = np.random.default_rng(123)
t0 = time.perf_counter()
for c in range(4):
    perm = rng.permutation(R); _ = perm[(-xs[:, c])[perm].argsort()].argsort().astype("uint32")
dt = (time.perf_counter() - t0) / 4
print(f"reference closure on one host core: {dt * 1e3:.1f} ms per cell -> {dt * C:.1f} s for {C} cells")
