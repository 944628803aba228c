"""Tiny end-to-end pass over the hot-path kernels for compute-sanitizer (memcheck / racecheck / synccheck):
corpus construction, init, sweeps in five kernel shapes (default and deterministic, fp32 and fp64), the exchange
kernels with two ranks on one device (cooperative launch), log-likelihood, metric block, exports, and one
tensor-core imputation chunk with its consumers.  Sizes are small enough for the sanitizer's ~100x slowdown.
    compute-sanitizer --tool memcheck  python scripts/sanitizer_run.py
    compute-sanitizer --tool racecheck python scripts/sanitizer_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pycistopic_b200 import Corpus, DeviceModel
from pycistopic_b200.adlda import balanced_cell_partition
from pycistopic_b200.diff_features import calculate_imputed_accessibility, calculate_normalized_scores
from pycistopic_b200.synth import make_binary_matrix

m, _, _ = make_binary_matrix(n_cells=120, n_regions=3000, density=0.05, n_true_topics=5, seed=3)
corpus = Corpus.from_regions_by_cells(m)
for K, shape in ((2, None), (10, None), (50, (4, 4)), (100, (8, 4)), (300, (16, 4))):
    mdl = DeviceModel(corpus, K, 50.0 / K, 0.1, seed=1)
    if shape:
        mdl.set_sweep_shape(*shape)
    mdl.init()
    mdl.sweep(3)
    assert mdl.verify_counts() == {"n_dk": 0, "n_wk": 0, "n_k": 0}
    ll = mdl.loglikelihood()
    if K == 10:
        mdl.metric_ingredients(20), mdl.loglikelihood_compat(50.0, 0.1), mdl.frames(), mdl.counts(), mdl.distributions()
        mdl.set_deterministic(2)
        mdl.sweep(2)
        mdl.set_deterministic(0)
        mdl.set_conditional_precision(64)
        mdl.sweep(2)
        assert mdl.verify_counts() == {"n_dk": 0, "n_wk": 0, "n_k": 0}
    mdl.close()
    print("sweep ok", K, ll, flush=True)
# exchange kernels: two ranks on this device
W, D = m.shape
part, cell_ptr = balanced_cell_partition(np.bincount(m.indices, minlength=D), 2)
for variant in ("round", "sync"):
    shards = []
    for r in range(2):
        c = Corpus.from_regions_by_cells(m, cell_range=(int(part[r]), int(part[r + 1])))
        c.set_token_offset(int(cell_ptr[part[r]]))
        c.set_global_size(D, int(cell_ptr[-1]))
        md = DeviceModel(c, 12, 4.0, 0.1, seed=5)
        md.init()
        shards.append((c, md))
    arenas = [md.exchange_arena()[0] for _, md in shards]
    for r, (_, md) in enumerate(shards):
        md.exchange_open(arenas, r)
    models = [md for _, md in shards]
    step = DeviceModel.exchange_round_local if variant == "round" else DeviceModel.exchange_sync_local
    for it in range(3):
        step(models, 4)
        for md in models:
            md.exchange_wait()
            md.sweep(1)
    step(models, 4)
    for md in models:
        md.exchange_wait()
        md.sync()
    hs = [md.state_hash() for md in models]
    assert hs[0] == hs[1] and hs[0]["sum_nwk"] == m.nnz and models[0].exchange_timeouts() == 0
    for c, md in shards:
        md.close(); c.close()
    print("exchange ok", variant, flush=True)
corpus.close()
rng = np.random.default_rng(0)
a = rng.dirichlet(np.full(300, 0.1), size=20).T.astype(np.float32)
b = rng.dirichlet(np.full(20, 0.5), size=260).T.astype(np.float32)
out, kept = calculate_imputed_accessibility(a, b, list(range(300)), 10 ** 6, 128)
calculate_normalized_scores(out, 10 ** 4)
print("impute ok", out.shape, flush=True)
