"""Region-blocked sweeps when n_wk does not fit the L2: one rank's share (1/8 of the cells) of the atlas-scale C4 matrix
(K = 100: 1 M regions x 100 topics x 4 B = 400 MB per rank), ms per sweep for a plain sweep and for B launches over
token-balanced region blocks (B x 400 / B MB slices).  Also C3 / K = 60 (128 MB)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pycistopic_b200 import Corpus, DeviceModel
from pycistopic_b200.adlda import balanced_region_blocks
from pycistopic_b200.synth import SHAPES, planted_topic_tokens
dev = torch.device("cuda:0")
for shape, K, blocks in (("C4", 100, (1, 2, 4, 6, 8, 12)), ("C3", 60, (1, 2, 3, 4))):
    D, W, dens = SHAPES[shape]
    cb, ce = 0, D // 8
    cell, region = planted_topic_tokens(D, W, dens, seed=1234, device=dev, cell_range=(cb, ce))
    cnt = torch.bincount(cell - cb, minlength=ce - cb)
    cp = torch.zeros(ce - cb + 1, dtype=torch.int64, device=dev); cp[1:] = torch.cumsum(cnt, 0)
    tok = region.to(torch.int32).contiguous()
    del cell, region
    N = tok.numel()
    per_region = torch.bincount(tok, minlength=W).cpu().numpy()
    corpus = Corpus.from_device_tokens(cp.data_ptr(), tok.data_ptr(), ce - cb, W, N, keepalive=(cp, tok))
    corpus.set_global_size(D, N * 8)
    for nb in blocks:
        m = DeviceModel(corpus, K, 50.0 / K, 0.1, seed=555)
        m.init()
        if nb > 1:
            corpus.set_region_blocks(balanced_region_blocks(per_region, nb))
        run = (lambda n: m.sweep(n)) if nb == 1 else (lambda n: m.sweep_blocked(n))
        run(12)
        m.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); run(8); m.sync(); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 8
        ok = m.verify_counts() == {"n_dk": 0, "n_wk": 0, "n_k": 0}
        print(f"{shape} shard ({ce - cb} cells, {N} tokens, K={K}, table {4 * W * m.padded_topics / 1e6:.0f} MB) blocks={nb:2d}: "
              f"{ms:7.3f} ms/sweep  {N / ms / 1e6:6.2f} Gtok/s  ll/token {m.loglikelihood() / N:.4f}  counts_ok={ok}", flush=True)
        m.close()
    corpus.close()
    del tok, cp
    torch.cuda.empty_cache()
