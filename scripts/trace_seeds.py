"""Scratch: mean ll-trace over seeds, GPU configurations vs oracle (separates bias from noise)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.sparse as sp
from concurrent.futures import ThreadPoolExecutor
from oracle import lda_oracle as lo
from pycistopic_b200 import Corpus, DeviceModel
from pycistopic_b200.synth import make_binary_matrix

def oracle_traces(X, K, n_iter, seeds):
    def one(s):
        o = lo.OracleLDA(K, n_iter=n_iter, alpha=50.0/K, eta=0.1, random_state=s, refresh=1).fit(X)
        return np.array(o.loglikelihoods_ + [o.final_loglikelihood_])
    with ThreadPoolExecutor(8) as ex:
        return np.array(list(ex.map(one, seeds)))

def gpu_traces(corpus, K, n_iter, seeds, shape=None, max_blocks=0):
    out = []
    for s in seeds:
        model = DeviceModel(corpus, K, 50.0/K, 0.1, seed=s)
        if shape: model.set_sweep_shape(*shape)
        if max_blocks: model.set_max_blocks(max_blocks)
        model.init(); tr = []
        for it in range(n_iter):
            tr.append(model.loglikelihood()); model.sweep(1)
        tr.append(model.loglikelihood()); out.append(tr); model.close()
    return np.array(out)

def report(name, tr, ref):
    mu = ref.mean(axis=0)
    rel = (tr.mean(axis=0) - mu) / np.abs(mu) * 100
    se = np.sqrt(tr.var(axis=0, ddof=1)/len(tr) + ref.var(axis=0, ddof=1)/len(ref)) / np.abs(mu) * 100
    idx = [1, 2, 3, 5, 8, 12, 20, 30, 40]
    idx = [i for i in idx if i < len(mu)]
    print(f"{name:34s} rel%: {np.round(rel[idx], 2)}  se%: {np.round(se[idx], 2)}", flush=True)

if __name__ == "__main__":
    m, _, _ = make_binary_matrix("C1")
    X = sp.csr_matrix(m.T)
    K, n_iter = 10, 40
    seeds = list(range(1, 9))
    ref = oracle_traces(X, K, n_iter, seeds)
    report("oracle seeds 11-18", oracle_traces(X, K, n_iter, [s + 10 for s in seeds]), ref)
    corpus = Corpus.from_regions_by_cells(m)
    for name, shape, mb in [("default", None, 0), ("(32,1) all blocks", (32, 1), 0), ("(32,1) 32 blocks", (32, 1), 32),
                            ("(32,1) 4 blocks", (32, 1), 4), ("(32,1) 1 block", (32, 1), 1),
                            ("(1,1) 16 blocks", (1, 1), 16), ("(1,4) all", (1, 4), 0)]:
        report(name, gpu_traces(corpus, K, n_iter, seeds, shape, mb), ref)
