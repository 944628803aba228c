"""Scratch throughput probe: tokens/s of the sweep kernel for several K on a synthetic shape."""
import argparse, sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pycistopic_b200 import Corpus, DeviceModel
from pycistopic_b200.synth import SHAPES, planted_topic_tokens

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="C2")
ap.add_argument("--topics", default="2,5,10,20,30,40,50,100")
ap.add_argument("--sweeps", type=int, default=10)
ap.add_argument("--warm", type=int, default=3)
ap.add_argument("--shape_override", default="", help="lanes,warps")
ap.add_argument("--burn", type=int, default=0, help="extra sweeps before timing (mixing state)")
args = ap.parse_args()
D, W, dens = SHAPES[args.shape]
dev = torch.device("cuda:0")
t = time.time()
cell, region = planted_topic_tokens(D, W, dens, seed=1234, device=dev)
cnt = torch.bincount(cell, minlength=D)
cell_ptr = torch.zeros(D + 1, dtype=torch.int64, device=dev); cell_ptr[1:] = torch.cumsum(cnt, 0)
tok = region.to(torch.int32).contiguous()
N = tok.numel()
torch.cuda.synchronize()
print(f"generated {args.shape}: D={D} W={W} N={N} in {time.time()-t:.1f}s", flush=True)
corpus = Corpus.from_device_tokens(cell_ptr.data_ptr(), tok.data_ptr(), D, W, N, keepalive=(cell_ptr, tok))
ts = torch.cuda.Stream(); torch.cuda.set_stream(ts)
stream = ts.cuda_stream
for K in [int(k) for k in args.topics.split(",")]:
    m = DeviceModel(corpus, K, 50.0 / K, 0.1, seed=555, stream=stream)
    if args.shape_override:
        a, b = [int(x) for x in args.shape_override.split(",")]
        m.set_sweep_shape(a, b)
    m.init()
    ll0 = m.loglikelihood()
    m.sweep(args.warm + args.burn)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); m.sweep(args.sweeps); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.sweeps
    ll1 = m.loglikelihood()
    tps = N / (ms * 1e-3)
    bt = 4 * K + 28
    print(json.dumps({"K": K, "ms_per_sweep": round(ms, 3), "Gtok_s": round(tps / 1e9, 3),
                      "alg_GBs": round(tps * bt / 1e9, 1), "frac_6544": round(tps * bt / 1e9 / 6544.3, 3),
                      "ll0": ll0, "ll": ll1, "shape": m.sweep_shape}), flush=True)
    m.close()
