"""Tuning probe: ms per sweep of the sweep kernel for several (K, lanes-per-token) shapes on a synthetic
workload, with the count invariants checked after every timing (counts == histogram of z)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pycistopic_b200 import Corpus, DeviceModel
from pycistopic_b200.synth import SHAPES, planted_topic_tokens

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="C2")
ap.add_argument("--cases", default="50:4,50:2,30:2,30:1,20:2,20:1,10:1,100:8,100:4")
ap.add_argument("--sweeps", type=int, default=10)
ap.add_argument("--burn", type=int, default=20)
ap.add_argument("--check", type=int, default=1)
args = ap.parse_args()
D, W, dens = SHAPES[args.shape]
dev = torch.device("cuda:0")
cell, region = planted_topic_tokens(D, W, dens, seed=1234, device=dev)
cnt = torch.bincount(cell, minlength=D)
cell_ptr = torch.zeros(D + 1, dtype=torch.int64, device=dev); cell_ptr[1:] = torch.cumsum(cnt, 0)
tok = region.to(torch.int32).contiguous()
N = tok.numel()
corpus = Corpus.from_device_tokens(cell_ptr.data_ptr(), tok.data_ptr(), D, W, N, keepalive=(cell_ptr, tok))
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
print("lib", os.environ.get("CGS_B200_LIB", "default"), "N", N, flush=True)
for case in args.cases.split(","):
    parts = [int(x) for x in case.split(":")]
    K, lanes = parts[0], parts[1]
    warps = parts[2] if len(parts) > 2 else 8
    m = DeviceModel(corpus, K, 50.0 / K, 0.1, seed=555, stream=ts.cuda_stream)
    try:
        if lanes > 0:
            m.set_sweep_shape(lanes, warps)
    except Exception as e:
        print(json.dumps({"K": K, "lanes": lanes, "error": str(e)[:80]})); m.close(); continue
    m.init()
    ll0 = m.loglikelihood()
    m.sweep(args.burn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.sweep(args.sweeps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.sweeps
    ll1 = m.loglikelihood()
    ok = None
    if args.check:
        z = torch.from_numpy(m.get_z().astype(np.int64)).to(dev)
        nzw, ndz, nz = m.counts()
        cell_ids = torch.repeat_interleave(torch.arange(D, device=dev), cnt)
        ndz_ref = torch.zeros(D * K, dtype=torch.int64, device=dev).scatter_add_(0, cell_ids * K + z, torch.ones_like(z)).view(D, K)
        nzw_ref = torch.zeros(K * W, dtype=torch.int64, device=dev).scatter_add_(0, z * W + tok.long(), torch.ones_like(z)).view(K, W)
        ok = bool((ndz_ref.cpu().numpy() == ndz).all() and (nzw_ref.cpu().numpy() == nzw).all()
                  and (np.bincount(z.cpu().numpy(), minlength=K) == nz).all())
    print(json.dumps({"K": K, "shape": m.sweep_shape, "ms": round(ms, 3), "Gtok_s": round(N / ms / 1e6, 2),
                      "frac": round(N / ms / 1e6 * (4 * K + 28) / 6544.3, 3), "ll0": round(ll0, 1), "ll": round(ll1, 1),
                      "counts_ok": ok}), flush=True)
    m.close()
