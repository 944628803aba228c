import sys; sys.path.insert(0,'/root/repo')
import numpy as np, scipy.sparse as sp
from pycistopic_b200 import Corpus, DeviceModel
from pycistopic_b200.synth import make_binary_matrix
for name,m in [("small", make_binary_matrix(n_cells=300, n_regions=5000, density=0.04, n_true_topics=6, seed=11)[0]), ("C1", make_binary_matrix("C1")[0])]:
    corpus = Corpus.from_regions_by_cells(m)
    N = corpus.n_tokens
    for K, shape in [(2,None),(10,None),(50,None),(50,(4,4)),(50,(4,16)),(36,(2,4)),(10,(2,8)),(4,(1,8)),(100,(8,4)),(300,(16,4))]:
        mod = DeviceModel(corpus, K, 50.0/K, 0.1, seed=1)
        if shape: mod.set_sweep_shape(*shape)
        mod.init(); mod.sweep(2)
        v = [mod.debug_sweep_visits() for _ in range(3)]
        print(name, K, shape, mod.sweep_shape, N, v, "OK" if all(x==N for x in v) else "MISMATCH", flush=True)
        mod.close()
    corpus.close()
