import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, scipy.sparse as sp, time
from trace_seeds import gpu_traces, report
from pycistopic_b200 import Corpus, DeviceModel
from pycistopic_b200.synth import make_binary_matrix
m, _, _ = make_binary_matrix("C1")
ref = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_cache", "c1_K10_oracle8.npy"))
corpus = Corpus.from_regions_by_cells(m)
report("C1 K=10 default(auto caps)", gpu_traces(corpus, 10, 40, list(range(1, 9))), ref)
mm = DeviceModel(corpus, 10, 5.0, 0.1, seed=1); mm.init(); mm.sweep(3); mm.sync()
t = time.time(); mm.sweep(50); mm.sync(); print("C1 K=10 ms/sweep", (time.time() - t) / 50 * 1e3, mm.sweep_shape)
