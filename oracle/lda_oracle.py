"""CPU oracle -- TEST INFRASTRUCTURE ONLY (never imported by ``pycistopic_b200``).

Restates, on the CPU, everything pycisTopic's ``run_cgs_model`` gets from the third-party
packages ``lda`` and ``tmtoolkit`` (call sites /root/reference/src/pycisTopic/lda_models.py:248-305,
:332-357).  Those packages are not vendored in /root/reference and cannot be installed here
(no network), so the algorithms are restated from their published sources as recorded in
SURVEY.md Appendix B.  **PARITY UNPINNED** for these restatements: the reference ships no
tests, golden vectors or fixtures for this path.  The functions that DO
live in the reference tree (``utils.loglikelihood``, ``calculate_imputed_accessibility``,
``calculate_normalized_scores``) are pinned separately: see ``oracle/ref_functions.py`` and
``tests/golden/``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import scipy.sparse as sparse

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcgs_oracle.so")
_lib = None

_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)


def build(force: bool = False) -> str:
    """Compile oracle/cgs_oracle.c with the committed Makefile."""
    src = os.path.join(_HERE, "cgs_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libcgs_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.oracle_matrix_to_lists.restype = ctypes.c_int64
        L.oracle_matrix_to_lists.argtypes = [_i64p, _i32p, _i32p, ctypes.c_int32, _i32p, _i32p]
        L.oracle_initialize.restype = None
        L.oracle_initialize.argtypes = [_i32p, _i32p, _i32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                        _i32p, _i32p, _i32p]
        L.oracle_counts_from_z.restype = None
        L.oracle_counts_from_z.argtypes = L.oracle_initialize.argtypes
        L.oracle_sample_topics.restype = ctypes.c_int
        L.oracle_sample_topics.argtypes = [_i32p, _i32p, _i32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                           _i32p, _i32p, _i32p, _f64p, _f64p, _f64p, ctypes.c_int32]
        L.oracle_sample_topics_adlda.restype = ctypes.c_int
        L.oracle_sample_topics_adlda.argtypes = [_i32p, _i32p, _i32p, ctypes.c_int64, ctypes.c_int32,
                                                 ctypes.c_int32, ctypes.c_int32, _i64p,
                                                 _i32p, _i32p, _i32p, _f64p, _f64p, _f64p, ctypes.c_int32,
                                                 _i32p, ctypes.c_int32]
        L.oracle_loglikelihood.restype = ctypes.c_double
        L.oracle_loglikelihood.argtypes = [_i32p, _i32p, _i32p, _i32p, ctypes.c_int32, ctypes.c_int32,
                                           ctypes.c_int32, ctypes.c_double, ctypes.c_double]
        L.oracle_run_sweeps.restype = ctypes.c_int
        L.oracle_run_sweeps.argtypes = [_i32p, _i32p, _i32p, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32,
                                        _i32p, _i32p, _i32p, ctypes.c_double, ctypes.c_double,
                                        _f64p, ctypes.c_int32, ctypes.c_int32]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


# --------------------------------------------------------------------------------------
# lda/utils.py : matrix_to_lists   
# --------------------------------------------------------------------------------------
def matrix_to_lists(X):
    """(WS, DS) int32 token lists of a D x W integer count matrix, row-major, columns ascending."""
    X = sparse.csr_matrix(X)
    if not np.issubdtype(X.dtype, np.integer):
        raise ValueError("expected sparse matrix with integer values, found float values")
    X = X.copy()
    X.sum_duplicates()
    X.sort_indices()
    X.eliminate_zeros()
    indptr = X.indptr.astype(np.int64)
    indices = X.indices.astype(np.int32)
    data = np.ascontiguousarray(X.data.astype(np.int32))
    N = int(data.sum())
    WS = np.empty(N, np.int32)
    DS = np.empty(N, np.int32)
    n = lib().oracle_matrix_to_lists(_p(indptr, _i64p), _p(indices, _i32p), _p(data, _i32p),
                                     X.shape[0], _p(WS, _i32p), _p(DS, _i32p))
    assert n == N
    return WS, DS


def matrix_to_lists_py(X):
    """Pure-Python/NumPy restatement of the same function (small cases; cross-checks the C one)."""
    X = np.asarray(sparse.csr_matrix(X).todense())
    ii, jj = np.nonzero(X)
    ss = X[ii, jj]
    DS = np.repeat(ii, ss).astype(np.intc)
    WS = np.repeat(jj, ss).astype(np.intc)
    return WS, DS


def counts_from_z(WS, DS, ZS, D, W, K):
    nzw = np.zeros((K, W), np.int32)
    ndz = np.zeros((D, K), np.int32)
    nz = np.zeros(K, np.int32)
    ZS = np.ascontiguousarray(ZS, np.int32)
    lib().oracle_counts_from_z(_p(WS, _i32p), _p(DS, _i32p), _p(ZS, _i32p), len(WS), K, W,
                               _p(nzw, _i32p), _p(ndz, _i32p), _p(nz, _i32p))
    return nzw, ndz, nz


def loglikelihood_true(nzw, ndz, nz, alpha, eta):
    """lda/_lda.pyx _loglikelihood."""
    nzw = np.ascontiguousarray(nzw, np.int32)
    ndz = np.ascontiguousarray(ndz, np.int32)
    nz = np.ascontiguousarray(nz, np.int32)
    nd = np.ascontiguousarray(ndz.sum(axis=1), np.int32)
    K, W = nzw.shape
    return float(lib().oracle_loglikelihood(_p(nzw, _i32p), _p(ndz, _i32p), _p(nz, _i32p), _p(nd, _i32p),
                                            ndz.shape[0], K, W, float(alpha), float(eta)))


def loglikelihood_true_py(nzw, ndz, nz, alpha, eta):
    """Vectorised restatement of the same formula (scipy gammaln); cross-checks the C one."""
    from scipy.special import gammaln
    K, W = nzw.shape
    D = ndz.shape[0]
    nd = ndz.sum(axis=1)
    ll = K * gammaln(eta * W) - gammaln(eta * W + nz).sum()
    m = nzw > 0
    ll += (gammaln(eta + nzw[m]) - gammaln(eta)).sum()
    ll += (gammaln(alpha * K) - gammaln(alpha * K + nd)).sum()
    m = ndz > 0
    ll += (gammaln(alpha + ndz[m]) - gammaln(alpha)).sum()
    _ = D
    return float(ll)


# --------------------------------------------------------------------------------------
# lda/lda.py : LDA      
# --------------------------------------------------------------------------------------
class OracleLDA:
    """Restatement of ``lda.LDA`` with the exact ``RandomState`` protocol of the package.

    ``refresh`` controls how often the true log-likelihood is appended to
    ``loglikelihoods_`` (pycisTopic passes ``refresh=n_iter``, lda_models.py:254; parity
    traces use ``refresh=1``).  ``n_parts > 1`` switches the sweep to the AD-LDA variant
    (cells split into token-balanced contiguous parts).
    """

    def __init__(self, n_topics, n_iter=2000, alpha=0.1, eta=0.01, random_state=None, refresh=10,
                 n_parts=1, keep_z=False, part_ptr=None):
        if alpha <= 0 or eta <= 0:
            raise ValueError("alpha and eta must be greater than zero")
        self.n_topics = n_topics
        self.n_iter = n_iter
        self.alpha = alpha
        self.eta = eta
        self.random_state = random_state
        self.refresh = refresh
        self.n_parts = n_parts
        self.keep_z = keep_z
        self.part_ptr = part_ptr
        rng = np.random.RandomState(random_state)
        self._rands = rng.rand(1024 ** 2 // 8)

    def _initialize(self, X):
        D, W = X.shape
        K = self.n_topics
        self.WS, self.DS = matrix_to_lists(X)
        N = len(self.WS)
        self.ZS = np.empty(N, np.int32)
        self.nzw_ = np.zeros((K, W), np.int32)
        self.ndz_ = np.zeros((D, K), np.int32)
        self.nz_ = np.zeros(K, np.int32)
        lib().oracle_initialize(_p(self.WS, _i32p), _p(self.DS, _i32p), _p(self.ZS, _i32p), N, K, W,
                                _p(self.nzw_, _i32p), _p(self.ndz_, _i32p), _p(self.nz_, _i32p))
        self.loglikelihoods_ = []

    def loglikelihood(self):
        return loglikelihood_true(self.nzw_, self.ndz_, self.nz_, self.alpha, self.eta)

    def _sample_topics(self, rands):
        K, W = self.nzw_.shape
        D = self.ndz_.shape[0]
        alpha = np.full(K, self.alpha, np.float64)
        eta = np.full(W, self.eta, np.float64)
        N = len(self.WS)
        if self.n_parts == 1:
            rc = lib().oracle_sample_topics(_p(self.WS, _i32p), _p(self.DS, _i32p), _p(self.ZS, _i32p), N, K, W,
                                            _p(self.nzw_, _i32p), _p(self.ndz_, _i32p), _p(self.nz_, _i32p),
                                            _p(alpha, _f64p), _p(eta, _f64p), _p(rands, _f64p), len(rands))
        else:
            rc = lib().oracle_sample_topics_adlda(
                _p(self.WS, _i32p), _p(self.DS, _i32p), _p(self.ZS, _i32p), N, K, W, D,
                _p(self._cell_ptr, _i64p), _p(self.nzw_, _i32p), _p(self.ndz_, _i32p), _p(self.nz_, _i32p),
                _p(alpha, _f64p), _p(eta, _f64p), _p(rands, _f64p), len(rands),
                _p(self._part_ptr, _i32p), self.n_parts)
        if rc != 0:
            raise MemoryError("oracle sweep scratch allocation failed")

    def fit(self, X, y=None):
        X = sparse.csr_matrix(X)
        self._initialize(X)
        if self.n_parts > 1:
            cnt = np.bincount(self.DS, minlength=X.shape[0]).astype(np.int64)
            self._cell_ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
            pp = self.part_ptr if self.part_ptr is not None else balanced_cell_partition(self._cell_ptr, self.n_parts)
            self._part_ptr = np.ascontiguousarray(pp, dtype=np.int32)
        random_state = np.random.RandomState(self.random_state)
        rands = self._rands.copy()
        for it in range(self.n_iter):
            random_state.shuffle(rands)
            if it % self.refresh == 0:
                self.loglikelihoods_.append(self.loglikelihood())
            self._sample_topics(rands)
        self.final_loglikelihood_ = self.loglikelihood()
        self.components_ = (self.nzw_ + self.eta).astype(float)
        self.components_ /= np.sum(self.components_, axis=1)[:, np.newaxis]
        self.topic_word_ = self.components_
        self.doc_topic_ = (self.ndz_ + self.alpha).astype(float)
        self.doc_topic_ /= np.sum(self.doc_topic_, axis=1)[:, np.newaxis]
        if not self.keep_z:
            del self.WS, self.DS, self.ZS
        return self


def balanced_cell_partition(cell_ptr, n_parts):
    """Contiguous cell ranges with (nearly) equal token counts: part p = [ptr[p], ptr[p+1])."""
    cell_ptr = np.asarray(cell_ptr, np.int64)
    D = len(cell_ptr) - 1
    N = int(cell_ptr[-1])
    targets = (np.arange(1, n_parts, dtype=np.float64) * N / n_parts)
    cuts = np.searchsorted(cell_ptr, targets, side="left")
    cuts = np.clip(cuts, 0, D)
    return np.concatenate([[0], cuts, [D]]).astype(np.int64)


# --------------------------------------------------------------------------------------
# tmtoolkit.topicmod.evaluate / model_stats   
# --------------------------------------------------------------------------------------
def metric_arun_2010(topic_word_distrib, doc_topic_distrib, doc_lengths):
    cm1 = np.linalg.svd(topic_word_distrib, compute_uv=False)
    doc_lengths = np.asarray(doc_lengths, dtype=float)
    if doc_lengths.ndim == 1 or doc_lengths.shape[0] != 1:
        doc_lengths = doc_lengths.reshape(1, -1) if doc_lengths.ndim == 1 else doc_lengths.T
    cm2 = doc_lengths @ doc_topic_distrib
    cm2 = cm2 / np.linalg.norm(doc_lengths, 2)
    cm2 = cm2[0]
    return float(np.sum(cm1 * np.log(cm1 / cm2)) + np.sum(cm2 * np.log(cm2 / cm1)))


def metric_cao_juan_2009(topic_word_distrib):
    from scipy.spatial.distance import pdist
    cos_sim = 1 - pdist(topic_word_distrib, metric="cosine")
    return float(np.mean(cos_sim))


def metric_coherence_mimno_2011(topic_word_distrib, dtm, top_n=20, eps=1e-12, normalize=True,
                                return_mean=False):
    n_topics, n_vocab = topic_word_distrib.shape
    if top_n > n_vocab:
        raise ValueError("`top_n=%d` is larger than the vocabulary size of %d words" % (top_n, n_vocab))
    top_words = np.argsort(topic_word_distrib, axis=1)[:, :-(top_n + 1):-1]
    dtm = sparse.csc_matrix(dtm)
    coh = []
    for t in range(n_topics):
        c_t = 0.0
        v = top_words[t]
        top_dtm = dtm[:, v]
        bin_dtm = (top_dtm > 0).astype(np.int64)
        df = np.asarray(bin_dtm.sum(axis=0)).ravel()
        codf = (bin_dtm.T @ bin_dtm).toarray()
        for m in range(1, top_n):
            for l in range(m):
                c_t += np.log((codf[m, l] + eps) / df[l])
        coh.append(c_t)
    coh = np.array(coh)
    if normalize:
        coh *= 2 / (top_n * (top_n - 1))
    if return_mean:
        return float(coh.mean())
    return coh


def marginal_topic_distrib(doc_topic_distrib, doc_lengths):
    doc_lengths = np.asarray(doc_lengths, dtype=float).ravel()
    unnorm = (doc_topic_distrib.T * doc_lengths).sum(axis=1)
    return unnorm / unnorm.sum()


# --------------------------------------------------------------------------------------
# Model-level restatement of run_cgs_model's metric block (lda_models.py:291-344)
# --------------------------------------------------------------------------------------
def model_metrics(model, X, alpha_raw, eta_raw, top_topics_coh=5, loglikelihood_fn=None):
    X = sparse.csr_matrix(X)
    doc_len = np.asarray(X.sum(axis=1)).astype(float)
    arun = metric_arun_2010(model.topic_word_, model.doc_topic_, doc_len)
    cao = metric_cao_juan_2009(model.topic_word_)
    mimno = metric_coherence_mimno_2011(model.topic_word_, X, top_n=20, eps=1e-12, normalize=True)
    if len(mimno) <= top_topics_coh:
        coh = float(np.mean(mimno))
    else:
        coh = float(np.mean(mimno[np.argpartition(mimno, -top_topics_coh)[-top_topics_coh:]]))
    out = {"Arun_2010": arun, "Cao_Juan_2009": cao, "Mimno_2011": coh, "mimno_per_topic": mimno,
           "marg_topic": marginal_topic_distrib(model.doc_topic_, doc_len)}
    if loglikelihood_fn is not None:
        out["loglikelihood"] = loglikelihood_fn(model.nzw_, model.ndz_, alpha_raw, eta_raw)
    return out


def normalized_scores(imputed_acc, scale_factor):
    """CPU restatement of ``calculate_normalized_scores`` (diff_features.py:540-548): per-cell totals over
    the regions, ``X / (totals / scale_factor)``, ``log1p``.  Pinned by tests/golden/impute.npz (norm0,
    norm5), which the reference's own nested function produced."""
    x = np.asarray(imputed_acc)
    normalized = x / (np.sum(x, axis=0) / scale_factor)
    return np.log1p(normalized, out=normalized)
