"""find_highly_variable_features on a C2-shaped model: device-resident reduction rate, the factored path (nothing
materialised) and the reference's arithmetic (np.mean / np.var, dtype=float32) on the host.
usage: variable_features_bench.py [R C K]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pycistopic_b200 import _lib, diff_features as pdf
from pycistopic_b200._lib import c_vp, check

R = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
K = int(sys.argv[3]) if len(sys.argv) > 3 else 50
import logging; logging.disable(logging.INFO)
L = _lib.load()
dev = torch.device("cuda:0")
s = torch.cuda.current_stream().cuda_stream
for name, dt, code, rows in (("int32", torch.int32, 0, 20000), ("float64", torch.float64, 2, 20000)):
    x = torch.randint(0, 500, (rows, C), device=dev, dtype=torch.int32).to(dt)
    m = torch.empty(rows, dtype=torch.float32, device=dev); v = torch.empty_like(m)
    run = lambda: check(L.cgs_row_mean_var_device(0, code, c_vp(x.data_ptr()), C, rows, C, c_vp(m.data_ptr()), c_vp(v.data_ptr()), c_vp(s)))
    for _ in range(2): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"row_mean_var_kernel {name}: {rows} x {C}: {ms:.3f} ms, {x.numel() * x.element_size() / ms / 1e6:.0f} GB/s of matrix read "
          f"(one algorithmic read; the second pass hits L2)", flush=True)
    del x
rng = np.random.default_rng(1)
a = rng.dirichlet(np.full(R, 0.1), size=K).T.astype(np.float32)
b = rng.dirichlet(np.full(K, 0.3), size=C).T.astype(np.float32)
fact = pdf.FactoredImputedFeatures(a, b, [f"r{j}" for j in range(R)], [f"c{j}" for j in range(C)], 10 ** 6, 20000, "bench")
fact.row_mean_var(10 ** 4)  # warm-up
t = time.perf_counter(); sel_f = pdf.find_highly_variable_features(pdf.normalize_scores(fact, 10 ** 4, materialize=False), plot=False); t_f = time.perf_counter() - t
print(f"factored: normalize_scores(materialize=False) -> find_highly_variable_features on {R} x {C} x K={K}: {t_f:.2f} s, {len(sel_f)} features", flush=True)
if R * C <= 2.5e9:
    t = time.perf_counter(); dense = pdf.normalize_scores(fact, 10 ** 4); t_n = time.perf_counter() - t
    t = time.perf_counter(); sel_d = pdf.find_highly_variable_features(dense, plot=False); t_d = time.perf_counter() - t
    print(f"dense:    normalize_scores to host {t_n:.2f} s ({dense.mtx.nbytes / 1e9:.1f} GB) + find_highly_variable_features {t_d:.2f} s; same selection: {sel_d == sel_f}", flush=True)
    n = 2000
    t = time.perf_counter(); np.mean(dense.mtx[:n], axis=1, dtype=np.float32); np.var(dense.mtx[:n], axis=1, dtype=np.float32); t_c = (time.perf_counter() - t) / n
    print(f"reference arithmetic (np.mean + np.var, dtype=float32) on one host core: {t_c * 1e6:.1f} us per region -> {t_c * dense.mtx.shape[0]:.1f} s for the matrix", flush=True)
