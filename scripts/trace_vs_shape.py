"""Scratch: ll-trace deviation from the oracle as a function of the sweep shape."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp
from oracle import lda_oracle as lo
from pycistopic_b200 import Corpus, DeviceModel
from pycistopic_b200.synth import make_binary_matrix

def run(m, K, n_iter, shapes, seeds=(555,), nkdiv=None):
    X = sp.csr_matrix(m.T)
    alpha, eta = 50.0 / K, 0.1
    o = lo.OracleLDA(K, n_iter=n_iter, alpha=alpha, eta=eta, random_state=555, refresh=1).fit(X)
    ot = np.array(o.loglikelihoods_ + [o.final_loglikelihood_])
    o2 = lo.OracleLDA(K, n_iter=n_iter, alpha=alpha, eta=eta, random_state=7, refresh=1).fit(X)
    ot2 = np.array(o2.loglikelihoods_ + [o2.final_loglikelihood_])
    rel = np.abs(ot2 - ot) / np.abs(ot)
    print(f"K={K} oracle seed7 vs seed555: max {rel.max():.3%} at {rel.argmax()}  first: {np.round(rel[1:9]*100,2)}", flush=True)
    corpus = Corpus.from_regions_by_cells(m)
    for shape in shapes:
        for seed in seeds:
            model = DeviceModel(corpus, K, alpha, eta, seed=seed)
            if shape is not None:
                model.set_sweep_shape(*shape)
            if nkdiv: model.set_nk_staleness(nkdiv)
            model.init()
            tr = []
            for it in range(n_iter):
                tr.append(model.loglikelihood()); model.sweep(1)
            tr.append(model.loglikelihood())
            gt = np.array(tr)
            rel = (gt - ot) / np.abs(ot)
            print(f"K={K} shape={shape} {model.sweep_shape} seed={seed}: max |rel| {np.abs(rel).max():.3%} at {np.abs(rel).argmax()}  "
                  f"first: {np.round(rel[1:9]*100,2)} last: {rel[-1]*100:.3f}", flush=True)
            model.close()
    corpus.close()

m2, _, _ = make_binary_matrix(n_cells=300, n_regions=5000, density=0.04, n_true_topics=6, seed=11)
run(m2, 40, 30, [None, (4, 1), (32, 1)])
m, _, _ = make_binary_matrix("C1")
run(m, 10, 40, [None, (1, 1), (1, 2), (1, 4), (32, 1)])
run(m, 2, 40, [None, (1, 4)])
