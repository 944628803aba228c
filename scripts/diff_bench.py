"""Time the rank-sum kernel on a device-resident imputed chunk.  usage: diff_bench.py [rows cells n_fg]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pycistopic_b200 import _lib
from pycistopic_b200._lib import c_vp, check

R = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
n_fg = int(sys.argv[3]) if len(sys.argv) > 3 else C // 10
L = _lib.load()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(1)
for name, dt, code in (("int32", torch.int32, 0), ("float32", torch.float32, 1)):
    x = torch.randint(0, 2000, (R, C), device=dev, generator=g, dtype=torch.int32).to(dt)
    perm = torch.randperm(C, device=dev, generator=g).to(torch.int32)
    fg, bg = perm[:n_fg].sort().values.contiguous(), perm[n_fg:].sort().values.contiguous()  # ascending, as the host mirror passes them
    out = torch.empty(3, R, dtype=torch.float64, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    def run():
        check(L.cgs_ranksum_rows_device(0, code, R, c_vp(x.data_ptr()), C, c_vp(fg.data_ptr()), fg.numel(),
                                        c_vp(x.data_ptr()), C, c_vp(bg.data_ptr()), bg.numel(),
                                        c_vp(out[0].data_ptr()), c_vp(out[1].data_ptr()), c_vp(out[2].data_ptr()), c_vp(s)))
    for _ in range(2): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    gb = R * C * x.element_size() / 1e9
    print(f"{name}: {R} rows x {C} cells, fg {n_fg} / bg {C - n_fg}: {ms:.3f} ms per contrast, "
          f"{gb / ms * 1e3:.0f} GB/s of matrix read, {R / ms * 1e3:.3e} rows/s", flush=True)
    # spot check against scipy on 4 rows
    from scipy.stats import ranksums
    xs = x[:4].cpu().numpy(); f = fg.cpu().numpy(); b = bg.cpu().numpy()
    ref = np.array([ranksums(xs[i, f], xs[i, b]).pvalue for i in range(4)])
    print("   max rel diff vs scipy on 4 rows:", float(np.max(np.abs(out[1, :4].cpu().numpy() - ref) / ref)))
