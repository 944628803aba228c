"""Tensor-core imputation kernel: parity against an fp64 product on several shapes (ragged edges, K not a
multiple of 32) and a device-resident timing of a C5-shaped chunk (20 000 regions x 100 topics x n_cells)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pycistopic_b200 import _lib
from pycistopic_b200._lib import c_f32p, c_u8p, c_vp, check, ptr

L = _lib.load()
rng = np.random.default_rng(0)
ok_all = True
for (R, K, C) in [(128, 32, 256), (1000, 37, 333), (300, 100, 700), (257, 5, 1030), (4096, 64, 2048)]:
    a = np.ascontiguousarray(rng.dirichlet(np.full(R, 0.1), size=K).T, dtype=np.float32)
    b = np.ascontiguousarray(rng.dirichlet(np.full(K, 0.5), size=C).T, dtype=np.float32)
    ref = a.astype(np.float64) @ b.astype(np.float64)
    for as_int, scale in ((0, 1.0), (1, 1e6)):
        out = np.empty((R, C), np.int32 if as_int else np.float32)
        flags = np.empty(R, np.uint8)
        check(L.cgs_impute_chunk(0, ptr(a, c_f32p), ptr(b, c_f32p), R, K, C, scale, as_int, out.ctypes.data_as(c_vp),
                                 ptr(flags, c_u8p)))
        if as_int:
            err = np.abs(out.astype(np.int64) - np.trunc(ref * 1e6)).max()
            good = err <= 1
            fl = np.array_equal(flags.astype(bool), (out != 0).any(axis=1))
            print(f"R={R} K={K} C={C} int32: max |diff| = {err}  flags_ok={fl}  {'OK' if good and fl else 'FAIL'}", flush=True)
            ok_all &= bool(good and fl)
        else:
            rel = np.abs(out - ref).max() / np.abs(ref).max()
            relw = (np.abs(out - ref) / np.maximum(np.abs(ref), 1e-30)).max()
            good = np.allclose(out, ref, rtol=1e-5, atol=1e-12)
            print(f"R={R} K={K} C={C} fp32: max err / max = {rel:.2e}  worst elementwise rel = {relw:.2e}  {'OK' if good else 'FAIL'}", flush=True)
            ok_all &= bool(good)
print("PARITY", "OK" if ok_all else "FAIL", flush=True)

# ---- timing, device resident ---------------------------------------------------------------------
dev = torch.device("cuda:0")
for (R, K, C) in [(20000, 100, 20000), (20000, 100, 200000), (20000, 50, 100000)]:
    g = torch.Generator(device=dev); g.manual_seed(4321)
    A = torch.rand(R, K, device=dev, generator=g); A /= A.sum(0, keepdim=True)
    B = torch.rand(K, C, device=dev, generator=g); B /= B.sum(0, keepdim=True)
    out = torch.empty(R, C, dtype=torch.int32, device=dev)
    fl = torch.empty(R, dtype=torch.uint8, device=dev)
    plan = c_vp()
    check(L.cgs_impute_plan_create(ctypes.byref(plan), 0, B.data_ptr(), 1, K, C, None))
    s = torch.cuda.current_stream().cuda_stream
    run = lambda: check(L.cgs_impute_plan_chunk_device(plan, A.data_ptr(), R, 1e6, 1, out.data_ptr(), fl.data_ptr(), s))
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    # spot check against fp64 on a few rows
    rows = torch.randint(0, R, (16,), device=dev)
    ref = torch.trunc((A[rows].double() @ B.double()) * 1e6)
    d = (out[rows].double() - ref).abs().max().item()
    print(f"R={R} K={K} C={C}: {ms:.3f} ms per chunk  out {R*C*4/1e9:.2f} GB -> {R*C*4/ms/1e6:.0f} GB/s written, "
          f"{2.0*R*C*K/ms/1e9:.1f} TFLOP/s (2RCK), spot max|diff| = {d}", flush=True)
    check(L.cgs_impute_plan_destroy(plan))
