"""world_size-2 `gloo` test of the AD-LDA orchestration (pycistopic_b200/adlda.py) on the CPU.

The device shard is replaced by an oracle-backed shard with the same interface; the distributed run
(two processes, deltas summed with all_reduce) must then reproduce, bit for bit, the single-process
AD-LDA oracle with the same cell partition (oracle_sample_topics_adlda)."""
import os
import pickle
import socket
import sys
import tempfile

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import lda_oracle as lo  # noqa: E402
from pycistopic_b200 import adlda  # noqa: E402
from pycistopic_b200.synth import make_binary_matrix  # noqa: E402


class OracleShard:
    """CPU stand-in for GpuShard: same methods, oracle arithmetic (test infrastructure)."""

    def __init__(self, M, cell_range, token_offset, K, alpha, eta, seed):
        import torch
        self.torch = torch
        X = sp.csr_matrix(M.T)[cell_range[0]:cell_range[1]]
        self.D, self.W = X.shape
        self.K, self.alpha, self.eta, self.offset = K, alpha, eta, token_offset
        self.WS, self.DS = lo.matrix_to_lists(X)
        self.N = len(self.WS)
        self.rng = np.random.RandomState(seed)
        self.rands = np.random.RandomState(seed).rand(1024 ** 2 // 8)
        self.nd = np.bincount(self.DS, minlength=self.D)

    def init(self):
        K, W = self.K, self.W
        self.ZS = ((self.offset + np.arange(self.N)) % K).astype(np.int32)
        self.nzw, self.ndz, self.nz = lo.counts_from_z(self.WS, self.DS, self.ZS, self.D, W, K)
        self.snap = np.zeros_like(self.nzw)
        self.delta = np.zeros_like(self.nzw)

    def sweep_part(self, part, n_parts):
        if part == 0:
            self.rng.shuffle(self.rands)
        sel = np.flatnonzero(self.DS % n_parts == part) if n_parts > 1 else np.arange(self.N)
        WS = np.ascontiguousarray(self.WS[sel]); DS = np.ascontiguousarray(self.DS[sel]); ZS = np.ascontiguousarray(self.ZS[sel])
        # the oracle indexes rands by position: rotate so that token i uses rands[(offset + i) % n]
        n = len(self.rands)
        idx = (self.offset + sel) % n
        r = np.ascontiguousarray(self.rands[idx]) if n_parts > 1 or self.offset % n else None
        alpha = np.full(self.K, self.alpha); eta = np.full(self.W, self.eta)
        L, P = lo.lib(), lo._p
        if r is None:
            rc = L.oracle_sample_topics(P(WS, lo._i32p), P(DS, lo._i32p), P(ZS, lo._i32p), len(WS), self.K, self.W,
                                        P(self.nzw, lo._i32p), P(self.ndz, lo._i32p), P(self.nz, lo._i32p),
                                        P(alpha, lo._f64p), P(eta, lo._f64p), P(self.rands, lo._f64p), n)
        else:
            # per-token random numbers passed explicitly (n_rand == n_tokens => rands[i % n] == r[i])
            rc = L.oracle_sample_topics(P(WS, lo._i32p), P(DS, lo._i32p), P(ZS, lo._i32p), len(WS), self.K, self.W,
                                        P(self.nzw, lo._i32p), P(self.ndz, lo._i32p), P(self.nz, lo._i32p),
                                        P(alpha, lo._f64p), P(eta, lo._f64p), P(r, lo._f64p), max(1, len(r)))
        assert rc == 0
        self.ZS[sel] = ZS

    def delta_compute(self):
        np.subtract(self.nzw, self.snap, out=self.delta)

    def delta_tensor(self):
        return self.torch.from_numpy(self.delta)

    def delta_apply(self):
        self.nzw[:] = self.snap + self.delta
        self.snap[:] = self.nzw
        self.nz[:] = self.nzw.sum(axis=1)

    def loglik_cells(self):
        from scipy.special import gammaln
        a, K = self.alpha, self.K
        m = self.ndz > 0
        return float((gammaln(a * K) - gammaln(a * K + self.nd)).sum() + (gammaln(a + self.ndz[m]) - gammaln(a)).sum())

    def loglik_regions(self):
        from scipy.special import gammaln
        e, W, K = self.eta, self.W, self.K
        m = self.nzw > 0
        return float(K * gammaln(e * W) - gammaln(e * W + self.nz).sum() + (gammaln(e + self.nzw[m]) - gammaln(e)).sum())

    def export(self):
        return self.nzw.copy(), self.ndz.copy(), self.nz.copy()

    def close(self):
        pass


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, sync_per_sweep):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m, _, _ = make_binary_matrix(n_cells=41, n_regions=500, density=0.06, n_true_topics=3, seed=17)
    res = adlda.fit_adlda(m, n_topics=4, n_iter=6, alpha=0.9, eta=0.1, random_state=555, refresh=1,
                          sync_per_sweep=sync_per_sweep, shard_factory=OracleShard)
    with open(os.path.join(out_dir, f"r{rank}.pkl"), "wb") as f:
        pickle.dump({k: getattr(res, k) for k in ("nzw_", "ndz_", "nz_", "loglikelihoods_", "final_loglikelihood_",
                                                  "topic_word_", "doc_topic_", "part_ptr_")}, f)
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_adlda_two_ranks_equals_single_process_oracle(world):
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d, 1), nprocs=world, join=True)
        outs = [pickle.load(open(os.path.join(d, f"r{r}.pkl"), "rb")) for r in range(world)]
    m, _, _ = make_binary_matrix(n_cells=41, n_regions=500, density=0.06, n_true_topics=3, seed=17)
    X = sp.csr_matrix(m.T)
    o = lo.OracleLDA(4, n_iter=6, alpha=0.9, eta=0.1, random_state=555, refresh=1, n_parts=world,
                     part_ptr=outs[0]["part_ptr_"]).fit(X)
    for out in outs:  # every rank holds the full, identical result
        assert np.array_equal(out["nzw_"], o.nzw_)
        assert np.array_equal(out["ndz_"], o.ndz_)
        assert np.array_equal(out["nz_"], o.nz_)
        np.testing.assert_allclose(out["loglikelihoods_"], o.loglikelihoods_, rtol=1e-12)
        assert out["final_loglikelihood_"] == pytest.approx(o.final_loglikelihood_, rel=1e-12)
        np.testing.assert_allclose(out["topic_word_"], o.topic_word_, rtol=1e-12)
        np.testing.assert_allclose(out["doc_topic_"], o.doc_topic_, rtol=1e-12)
    assert outs[0]["nz_"].sum() == X.nnz


def test_adlda_sub_sweep_exchange_keeps_counts_exact():
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, _free_port(), d, 3), nprocs=2, join=True)
        outs = [pickle.load(open(os.path.join(d, f"r{r}.pkl"), "rb")) for r in range(2)]
    m, _, _ = make_binary_matrix(n_cells=41, n_regions=500, density=0.06, n_true_topics=3, seed=17)
    assert np.array_equal(outs[0]["nzw_"], outs[1]["nzw_"]) and np.array_equal(outs[0]["ndz_"], outs[1]["ndz_"])
    assert outs[0]["nz_"].sum() == m.nnz and np.array_equal(outs[0]["nzw_"].sum(axis=1), outs[0]["nz_"])
    assert np.array_equal(outs[0]["ndz_"].sum(axis=1), np.asarray(m.sum(axis=0)).ravel())
    assert outs[0]["loglikelihoods_"][-1] > outs[0]["loglikelihoods_"][0]


def test_balanced_partition_properties():
    rng = np.random.default_rng(0)
    for world in (1, 2, 4, 8):
        cnt = rng.integers(1, 500, size=97)
        ptr, cell_ptr = adlda.balanced_cell_partition(cnt, world)
        assert ptr[0] == 0 and ptr[-1] == 97 and np.all(np.diff(ptr) >= 1)
        loads = np.diff(cell_ptr[ptr])
        assert loads.sum() == cnt.sum() and loads.max() - loads.min() <= 2 * cnt.max()
    ptr, _ = adlda.balanced_cell_partition([1000, 1, 1], 3)
    assert ptr.tolist() == [0, 1, 2, 3]
    with pytest.raises(ValueError):
        adlda.fit_adlda(sp.csr_matrix(np.ones((3, 2), np.int32)), 2, exchange="bogus")
