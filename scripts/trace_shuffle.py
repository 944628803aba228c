"""Scratch: does the region-sorted token order cause the cross-cell lag?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, scipy.sparse as sp, torch
from trace_seeds import gpu_traces, report
from pycistopic_b200 import Corpus
from pycistopic_b200.synth import make_binary_matrix

m, _, _ = make_binary_matrix("C1")
X = sp.csr_matrix(m.T); X.sort_indices()
K, n_iter, seeds = 10, 40, list(range(1, 9))
ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_cache", "c1_K10_oracle8.npy"))
dev = torch.device("cuda:0")
cell_ptr = torch.tensor(X.indptr.astype(np.int64), device=dev)
def corpus_with(indices):
    tok = torch.tensor(indices.astype(np.int32), device=dev)
    return Corpus.from_device_tokens(cell_ptr.data_ptr(), tok.data_ptr(), X.shape[0], X.shape[1], len(indices), keepalive=(cell_ptr, tok))
sorted_c = corpus_with(X.indices)
rng = np.random.default_rng(0)
shuf = X.indices.copy()
for d in range(X.shape[0]):
    rng.shuffle(shuf[X.indptr[d]:X.indptr[d+1]])
shuf_c = corpus_with(shuf)
for name, corpus, shape, mb in [
    ("sorted default all", sorted_c, None, 0), ("shuffled default all", shuf_c, None, 0),
    ("sorted (32,1) all", sorted_c, (32, 1), 0), ("shuffled (32,1) all", shuf_c, (32, 1), 0),
    ("sorted (1,4) all", sorted_c, (1, 4), 0), ("shuffled (1,4) all", shuf_c, (1, 4), 0),
    ("sorted (1,4) 32 blocks", sorted_c, (1, 4), 32), ("sorted (1,4) 64 blocks", sorted_c, (1, 4), 64),
    ("sorted (1,4) 128 blocks", sorted_c, (1, 4), 128), ("sorted (1,2) 128 blocks", sorted_c, (1, 2), 128),
    ("sorted (1,1) 128 blocks", sorted_c, (1, 1), 128), ("sorted (1,1) 256 blocks", sorted_c, (1, 1), 256)]:
    report(name, gpu_traces(corpus, K, n_iter, seeds, shape, mb), ref)
