"""Mean log-likelihood trace over seeds, device sampler (default shapes) vs the sequential oracle.
Separates the bias of the parallel schedule from chain-to-chain noise.  usage: trace_check.py [case ...]
cases: small40 small100 c1k10 c1k5"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp
from concurrent.futures import ThreadPoolExecutor
from oracle import lda_oracle as lo
from pycistopic_b200 import Corpus, DeviceModel
from pycistopic_b200.synth import make_binary_matrix


def oracle_traces(X, K, n_iter, seeds):
    def one(s):
        o = lo.OracleLDA(K, n_iter=n_iter, alpha=50.0 / K, eta=0.1, random_state=s, refresh=1).fit(X)
        return np.array(o.loglikelihoods_ + [o.final_loglikelihood_])
    with ThreadPoolExecutor(8) as ex:
        return np.array(list(ex.map(one, seeds)))


def gpu_traces(corpus, K, n_iter, seeds, override=None):
    out = []
    shape = None
    for s in seeds:
        model = DeviceModel(corpus, K, 50.0 / K, 0.1, seed=s)
        if override:
            if override[0]: model.set_sweep_shape(override[0], override[1])
            if override[2]: model.set_max_blocks(override[2])
        model.init(); tr = []
        for it in range(n_iter):
            tr.append(model.loglikelihood()); model.sweep(1)
        tr.append(model.loglikelihood()); out.append(tr); shape = model.sweep_shape; model.close()
    return np.array(out), shape


CASES = {"c1k50": ("C1", 50, 40), "c1k30": ("C1", 30, 40), "small40": ("small", 40, 60), "small100": ("small", 100, 60), "c1k10": ("C1", 10, 40), "c1k5": ("C1", 5, 40),
         "c1k2": ("C1", 2, 40)}
for case in (sys.argv[1:] or ["small40", "small100", "c1k10"]):
    override = None
    if "@" in case:  # case@lanes,warps,max_blocks
        case, ov = case.split("@")
        override = [int(x) for x in ov.split(",")]
    data, K, n_iter = CASES[case]
    if data == "small":
        m, _, _ = make_binary_matrix(n_cells=300, n_regions=5000, density=0.04, n_true_topics=6, seed=11)
    else:
        m, _, _ = make_binary_matrix(data)
    X = sp.csr_matrix(m.T)
    seeds = list(range(1, 9))
    ref = oracle_traces(X, K, n_iter, seeds)
    corpus = Corpus.from_regions_by_cells(m)
    tr, shape = gpu_traces(corpus, K, n_iter, seeds, override)
    mu = ref.mean(axis=0)
    rel = (tr.mean(axis=0) - mu) / np.abs(mu) * 100
    se = np.sqrt(tr.var(axis=0, ddof=1) / len(tr) + ref.var(axis=0, ddof=1) / len(ref)) / np.abs(mu) * 100
    idx = [i for i in [1, 2, 3, 5, 8, 12, 20, 30, 40, 60] if i < len(mu)]
    print(f"{case:9s} {override} {shape} max|rel|={np.abs(rel).max():.2f}%\n   it  : {idx}\n   rel%: {np.round(rel[idx], 2)}\n   se% : {np.round(se[idx], 2)}", flush=True)
    corpus.close()
