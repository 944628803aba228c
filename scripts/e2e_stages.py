"""Where the end-to-end C-ABI call sequence of bench.py spends its wall clock (per stage, host timers
with a device sync after every stage)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pycistopic_b200 import _lib

dev = torch.device("cuda:0")
D, W, cell_ptr, tok = bench.make_workload_on_device(dev)
N = tok.numel()
cell_ptr_h = torch.empty(D + 1, dtype=torch.int64, pin_memory=True); cell_ptr_h.copy_(cell_ptr)
tok_h = torch.empty(N, dtype=torch.int32, pin_memory=True); tok_h.copy_(tok)
torch.cuda.synchronize()
indptr_np, indices_np = cell_ptr_h.numpy(), tok_h.numpy()
L = _lib.load()
topics = [int(k) for k in (sys.argv[1].split(",") if len(sys.argv) > 1 else "50,10".split(","))]
bufs = {K: (torch.empty((K, W), dtype=torch.int32, pin_memory=True).numpy(),
            torch.empty((D, K), dtype=torch.int32, pin_memory=True).numpy(), np.empty(K, np.int32)) for K in topics}


def timed(name, fn, acc):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    acc.setdefault(name, []).append((time.perf_counter() - t) * 1e3); return r


for rep in range(3):
    acc = {}
    h = ctypes.c_void_p()
    timed("corpus_from_cells_by_regions", lambda: _lib.check(L.cgs_corpus_from_cells_by_regions(
        ctypes.byref(h), 0, _lib.ptr(indptr_np, _lib.c_i64p), _lib.ptr(indices_np, _lib.c_i32p), D, W, N, 0, D)), acc)
    for K in topics:
        mh = ctypes.c_void_p()
        timed(f"K{K} create", lambda: _lib.check(L.cgs_model_create(ctypes.byref(mh), h, K, 50.0 / K, 0.1, 555, None)), acc)
        timed(f"K{K} init", lambda: _lib.check(L.cgs_model_init(mh)), acc)
        timed(f"K{K} sweep x10", lambda: _lib.check(L.cgs_model_sweep(mh, 10)), acc)
        a, b, c = bufs[K]
        timed(f"K{K} export_counts", lambda: _lib.check(L.cgs_model_export_counts(
            mh, _lib.ptr(a, _lib.c_i32p), _lib.ptr(b, _lib.c_i32p), _lib.ptr(c, _lib.c_i32p))), acc)
        timed(f"K{K} destroy", lambda: _lib.check(L.cgs_model_destroy(mh)), acc)
    timed("corpus_destroy", lambda: _lib.check(L.cgs_corpus_destroy(h)), acc)
    print(f"rep {rep}: " + "  ".join(f"{k}={v[0]:.1f}ms" for k, v in acc.items()), flush=True)
