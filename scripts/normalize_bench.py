"""normalize_scores on the device: time and achieved HBM traffic (read X twice, write the result once)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pycistopic_b200 import _lib
from pycistopic_b200._lib import check
L = _lib.load()
dev = torch.device("cuda:0")
for (R, C) in [(20000, 100000), (20000, 200000)]:
    x = torch.randint(0, 1000, (R, C), dtype=torch.int32, device=dev)
    out = torch.empty(R, C, dtype=torch.float64, device=dev)
    cs = torch.empty(C, dtype=torch.int64, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    run = lambda: check(L.cgs_normalize_scores_device(0, x.data_ptr(), 0, R, C, 1e4, out.data_ptr(), cs.data_ptr(), s))
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    ref = torch.log1p(x[:64].double() / (x.sum(0).double() / 1e4))
    err = float(((out[:64] - ref).abs() / ref.abs().clamp_min(1e-300)).max())
    print(f"normalize_scores int32 {R} x {C}: {ms:.3f} ms, {(R*C*(4+4+8))/ms/1e6:.0f} GB/s of algorithmic traffic "
          f"(2 reads + 1 fp64 write), max rel diff vs torch fp64 on 64 rows {err:.2e}", flush=True)
    del x, out
