"""``lda.LDA``-protocol front end of the device sampler.

pycisTopic's ``run_cgs_model`` builds ``lda.LDA(n_topics, n_iter, random_state, alpha, eta,
refresh)``, calls ``.fit(X)`` and reads ``topic_word_``, ``doc_topic_``, ``nzw_``, ``ndz_``,
``nz_`` (reference src/pycisTopic/lda_models.py:248-287, :292-305, :345-357).  This module
provides the same object protocol on top of the C ABI (``include/cgs_b200.h``); every count
and every sweep lives on the GPU.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes
import logging

import numpy as np
import scipy.sparse as sparse

from . import _lib
from ._lib import c_f64p, c_i32p, c_i64p, c_vp, check, ptr

logger = logging.getLogger("lda")


def _as_binary_csr(X, what):
    """Canonical int32/int64 CSR arrays of a binary matrix; rejects what `lda` would reject."""
    if not sparse.issparse(X):
        X = sparse.csr_matrix(np.asarray(X))
    X = X.tocsr()
    if not np.issubdtype(X.dtype, np.integer) and not np.issubdtype(X.dtype, np.bool_):
        # lda.utils.matrix_to_lists: "expected sparse matrix with integer values, found float values"
        if X.nnz and not np.all(np.equal(np.mod(X.data, 1), 0)):
            raise ValueError("expected sparse matrix with integer values, found float values")
    if X.nnz and X.data.max() > 1:
        raise ValueError(
            f"{what} must be binary (0/1), as CistopicObject.binary_matrix is; found a value of {X.data.max()}"
        )
    if X.nnz and X.data.min() < 1:
        X = X.copy()
        X.eliminate_zeros()
    indptr = np.ascontiguousarray(X.indptr, dtype=np.int64)
    indices = np.ascontiguousarray(X.indices, dtype=np.int32)
    return X.shape, indptr, indices


class Corpus:
    """Device-resident token list of one binary accessibility matrix (one GPU).

    Built either from the reference's ``binary_matrix`` layout (regions x cells; the transpose
    of lda_models.py:151 happens on the device) or from the cells x regions CSR that
    ``run_cgs_model`` receives.  ``cell_range`` keeps a contiguous shard of cells (AD-LDA).
    """

    def __init__(self, handle, device):
        self._h = handle
        self.device = device
        L = _lib.load()
        d, w, n = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64()
        check(L.cgs_corpus_dims(self._h, ctypes.byref(d), ctypes.byref(w), ctypes.byref(n)))
        self.n_cells, self.n_regions, self.n_tokens = d.value, w.value, n.value

    @staticmethod
    def canonical_csr(binary_matrix, what="binary_matrix"):
        """Validated ``(shape, indptr int64, indices int32)`` of a binary CSR; hand it to
        :meth:`from_regions_by_cells` when several devices build a corpus from the same matrix."""
        return _as_binary_csr(binary_matrix, what)

    @classmethod
    def from_regions_by_cells(cls, binary_matrix, device=0, cell_range=None):
        _lib.require_device()
        if isinstance(binary_matrix, tuple):  # already canonical (canonical_csr)
            (W, D), indptr, indices = binary_matrix
        else:
            (W, D), indptr, indices = _as_binary_csr(binary_matrix, "binary_matrix")
        cb, ce = cell_range if cell_range is not None else (0, D)
        h = c_vp()
        check(_lib.load().cgs_corpus_from_regions_by_cells(ctypes.byref(h), device, ptr(indptr, c_i64p),
                                                          ptr(indices, c_i32p), W, D, len(indices), cb, ce))
        return cls(h, device)

    @classmethod
    def from_cells_by_regions(cls, X, device=0, cell_range=None):
        _lib.require_device()
        (D, W), indptr, indices = _as_binary_csr(X, "X")
        cb, ce = cell_range if cell_range is not None else (0, D)
        h = c_vp()
        check(_lib.load().cgs_corpus_from_cells_by_regions(ctypes.byref(h), device, ptr(indptr, c_i64p),
                                                          ptr(indices, c_i32p), D, W, len(indices), cb, ce))
        return cls(h, device)

    @classmethod
    def from_device_tokens(cls, cell_ptr_addr, tok_region_addr, n_cells, n_regions, n_tokens, device=0,
                           token_offset=0, keepalive=None):
        """Wrap canonical token arrays already resident on ``device`` (raw device addresses)."""
        _lib.require_device()
        h = c_vp()
        check(_lib.load().cgs_corpus_from_device_tokens(ctypes.byref(h), device, c_vp(cell_ptr_addr),
                                                       c_vp(tok_region_addr), n_cells, n_regions, n_tokens,
                                                       token_offset))
        obj = cls(h, device)
        obj._keepalive = keepalive
        return obj

    def set_token_offset(self, offset: int):
        check(_lib.load().cgs_corpus_set_token_offset(self._h, int(offset)))

    def set_global_size(self, n_cells_global: int, n_tokens_global: int):
        """Size of the whole matrix when this corpus is one shard of its cells (AD-LDA): models created afterwards
        size their concurrency caps against it."""
        check(_lib.load().cgs_corpus_set_global_size(self._h, int(n_cells_global), int(n_tokens_global)))

    def set_region_split(self, region_split: int):
        """Region index that cuts a sweep into two half-sweeps (overlapped AD-LDA exchange)."""
        check(_lib.load().cgs_corpus_set_region_split(self._h, int(region_split)))

    def set_region_blocks(self, bounds):
        """Region blocks [bounds[b], bounds[b + 1]) that cut a sweep into one launch per block (bounds[0] = 0,
        bounds[-1] = n_regions)."""
        b = np.ascontiguousarray(bounds, dtype=np.int32)
        check(_lib.load().cgs_corpus_set_region_blocks(self._h, len(b) - 1, ptr(b, c_i32p)))

    def token_lists(self):
        """(WS, DS) exactly as ``lda.utils.matrix_to_lists`` returns them."""
        WS = np.empty(self.n_tokens, np.int32)
        DS = np.empty(self.n_tokens, np.int32)
        check(_lib.load().cgs_corpus_export_tokens(self._h, ptr(WS, c_i32p), ptr(DS, c_i32p)))
        return WS, DS

    def cell_sizes(self):
        """Tokens per cell, int64[n_cells] (= ``X.sum(axis=1)`` of the cells x regions matrix,
        lda_models.py:295); cached."""
        if getattr(self, "_cell_sizes", None) is None:
            s = np.empty(self.n_cells, np.int64)
            check(_lib.load().cgs_corpus_export_cell_sizes(self._h, ptr(s, c_i64p)))
            self._cell_sizes = s
        return self._cell_sizes

    def close(self):
        if self._h is not None:
            check(_lib.load().cgs_corpus_destroy(self._h))
            self._h = None

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and _lib._lib is not None:
                _lib._lib.cgs_corpus_destroy(self._h)
                self._h = None
        except Exception:
            pass


class DeviceModel:
    """One collapsed-Gibbs chain on a :class:`Corpus` (thin wrapper over ``cgs_model_*``)."""

    def __init__(self, corpus: Corpus, n_topics: int, alpha: float, eta: float, seed: int = 0, stream=None):
        if alpha <= 0 or eta <= 0:
            raise ValueError("alpha and eta must be greater than zero")
        self.corpus = corpus
        self.n_topics = int(n_topics)
        self.alpha = float(alpha)
        self.eta = float(eta)
        h = c_vp()
        check(_lib.load().cgs_model_create(ctypes.byref(h), corpus._h, self.n_topics, self.alpha, self.eta,
                                          int(seed) & 0xFFFFFFFFFFFFFFFF, c_vp(stream) if stream else None))
        self._h = h

    # -- state ---------------------------------------------------------------------------
    def init(self):
        check(_lib.load().cgs_model_init(self._h))

    def set_z(self, z):
        z = np.ascontiguousarray(z, dtype=np.int32)
        if z.shape[0] != self.corpus.n_tokens:
            raise ValueError("z has the wrong length")
        check(_lib.load().cgs_model_set_z(self._h, ptr(z, c_i32p)))

    def get_z(self):
        z = np.empty(self.corpus.n_tokens, np.int32)
        check(_lib.load().cgs_model_get_z(self._h, ptr(z, c_i32p)))
        return z

    def set_stream(self, stream):
        check(_lib.load().cgs_model_set_stream(self._h, c_vp(stream) if stream else None))

    def sync(self):
        check(_lib.load().cgs_model_sync(self._h))

    def set_sweep_shape(self, lanes_per_token: int, warps_per_cell: int):
        check(_lib.load().cgs_model_set_sweep_shape(self._h, lanes_per_token, warps_per_cell))

    def debug_sweep_visits(self) -> int:
        """One sweep with a device-side count of the tokens sampled (diagnostics)."""
        n = ctypes.c_int64()
        check(_lib.load().cgs_model_debug_sweep_visits(self._h, ctypes.byref(n)))
        return n.value

    def set_rng_period(self, period: int):
        """Distinct uniform draws per sweep (0 = every token its own; default 131072 as in lda)."""
        check(_lib.load().cgs_model_set_rng_period(self._h, int(period)))

    def set_max_blocks(self, max_blocks: int):
        check(_lib.load().cgs_model_set_max_blocks(self._h, int(max_blocks)))

    def set_nk_staleness(self, divisor: float):
        check(_lib.load().cgs_model_set_nk_staleness(self._h, float(divisor)))

    def set_conditional_precision(self, bits: int):
        """32 (default) or 64: arithmetic of the conditional; 64 is the A/B build in the reference's precision."""
        check(_lib.load().cgs_model_set_conditional_precision(self._h, int(bits)))

    def set_deterministic(self, parts: int):
        """parts > 0: bit-reproducible sweeps (every sweep = ``parts`` sub-sweeps that read copies of the count
        tables, one warp per cell); 0 = the default live-count schedule."""
        check(_lib.load().cgs_model_set_deterministic(self._h, int(parts)))

    @property
    def sweep_shape(self):
        a, b, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        check(_lib.load().cgs_model_get_sweep_shape(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return {"lanes_per_token": a.value, "chunks_per_lane": b.value, "warps_per_cell": c.value}

    # -- hot path ------------------------------------------------------------------------
    def sweep(self, n_sweeps: int = 1):
        check(_lib.load().cgs_model_sweep(self._h, int(n_sweeps)))

    def sweep_part(self, part: int, n_parts: int):
        check(_lib.load().cgs_model_sweep_part(self._h, int(part), int(n_parts)))

    def loglikelihood(self, partial: int = 0) -> float:
        ll = ctypes.c_double()
        check(_lib.load().cgs_model_loglik(self._h, partial, ctypes.byref(ll)))
        return ll.value

    @property
    def launch_count(self) -> int:
        n = ctypes.c_int64()
        check(_lib.load().cgs_model_launch_count(self._h, ctypes.byref(n)))
        return n.value

    @property
    def sweeps_done(self) -> int:
        n = ctypes.c_int64()
        check(_lib.load().cgs_model_sweeps_done(self._h, ctypes.byref(n)))
        return n.value

    # -- exports -------------------------------------------------------------------------
    def counts(self, nzw=True, ndz=True, nz=True):
        K, W, D = self.n_topics, self.corpus.n_regions, self.corpus.n_cells
        a = np.empty((K, W), np.int32) if nzw else None
        b = np.empty((D, K), np.int32) if ndz else None
        c = np.empty(K, np.int32) if nz else None
        check(_lib.load().cgs_model_export_counts(self._h, ptr(a, c_i32p) if nzw else None,
                                                 ptr(b, c_i32p) if ndz else None, ptr(c, c_i32p) if nz else None))
        return a, b, c

    def distributions(self, topic_word=True, doc_topic=True):
        K, W, D = self.n_topics, self.corpus.n_regions, self.corpus.n_cells
        a = np.empty((K, W), np.float64) if topic_word else None
        b = np.empty((D, K), np.float64) if doc_topic else None
        check(_lib.load().cgs_model_export_distributions(self._h, ptr(a, c_f64p) if topic_word else None,
                                                        ptr(b, c_f64p) if doc_topic else None))
        return a, b

    def frames(self, topic_region=True, cell_topic=True):
        """(topic_region W x K, cell_topic K x D) float64: the DataFrame layouts of lda_models.py:348-357."""
        K, W, D = self.n_topics, self.corpus.n_regions, self.corpus.n_cells
        a = np.empty((W, K), np.float64) if topic_region else None
        b = np.empty((K, D), np.float64) if cell_topic else None
        check(_lib.load().cgs_model_export_frames(self._h, ptr(a, c_f64p) if topic_region else None,
                                                 ptr(b, c_f64p) if cell_topic else None))
        return a, b

    # -- model-quality metrics (lda_models.py:291-305, :332-344) -------------------------------
    def topic_gram(self):
        """``topic_word_ @ topic_word_.T`` (K x K float64) from an exact integer Gram of ``n_wk``."""
        g = np.empty((self.n_topics, self.n_topics), np.float64)
        check(_lib.load().cgs_model_topic_gram(self._h, ptr(g, c_f64p)))
        return g

    def cell_mixture(self):
        """(``doc_lengths @ doc_topic_`` as float64[K], ``sum(doc_lengths ** 2)``)."""
        mix = np.empty(self.n_topics, np.float64)
        sq = ctypes.c_double()
        check(_lib.load().cgs_model_cell_mixture(self._h, ptr(mix, c_f64p), ctypes.byref(sq)))
        return mix, sq.value

    def top_regions(self, top_n=20, counts=False):
        """The ``top_n`` most probable regions of every topic, most probable first (K x top_n int32);
        equal counts are ordered by region index descending (a stable argsort read backwards)."""
        if top_n > self.corpus.n_regions:  # tmtoolkit's check and message
            raise ValueError("`top_n=%d` is larger than the vocabulary size of %d words" % (top_n, self.corpus.n_regions))
        top = np.empty((self.n_topics, top_n), np.int32)
        cnt = np.empty((self.n_topics, top_n), np.int32)
        check(_lib.load().cgs_model_top_regions(self._h, int(top_n), ptr(top, c_i32p), ptr(cnt, c_i32p)))
        return (top, cnt) if counts else top

    def codocument_frequencies(self, top_regions):
        """codf[t, a, b] = cells containing both ``top_regions[t, a]`` and ``top_regions[t, b]``
        (diagonal: document frequencies), K x top_n x top_n int32."""
        top = np.ascontiguousarray(top_regions, dtype=np.int32)
        if top.ndim != 2 or top.shape[0] != self.n_topics:
            raise ValueError("top_regions must be n_topics x top_n")
        codf = np.empty((self.n_topics, top.shape[1], top.shape[1]), np.int32)
        check(_lib.load().cgs_model_codocument_frequencies(self._h, top.shape[1], ptr(top, c_i32p),
                                                          ptr(codf, c_i32p)))
        return codf

    def loglikelihood_compat(self, alpha_raw: float, eta_raw: float) -> float:
        """What ``pycisTopic.utils.loglikelihood(nzw_, ndz_, alpha, eta)`` returns (utils.py:129-160)."""
        ll = ctypes.c_double()
        check(_lib.load().cgs_model_loglik_compat(self._h, float(alpha_raw), float(eta_raw), ctypes.byref(ll)))
        return ll.value

    def metric_ingredients(self, top_n=20):
        """Everything ``build_cistopic_model`` needs for the metric block, computed on the device."""
        top = self.top_regions(top_n)
        mix, sq = self.cell_mixture()
        return {"gram": self.topic_gram(), "cell_mixture": mix, "length_sq": sq, "top_regions": top,
                "codf": self.codocument_frequencies(top), "top_n": top_n}

    # -- consistency checks ----------------------------------------------------------------
    def region_histogram(self, device_address: int) -> int:
        """Writes the (region, topic) histogram of z to a device buffer (int32[n_regions * Kp]); returns the number
        of n_dk entries that differ from the per-cell histogram of z."""
        bad = ctypes.c_int64()
        check(_lib.load().cgs_model_region_histogram(self._h, c_vp(device_address), ctypes.byref(bad)))
        return bad.value

    def verify_counts(self, histogram_address: int | None = None):
        """{"n_dk": ..., "n_wk": ..., "n_k": ...} = entries of each table that are NOT the bookkeeping of z (all 0
        for a consistent state).  ``histogram_address``: device histogram summed over all ranks (AD-LDA shards)."""
        out = (ctypes.c_int64 * 3)()
        check(_lib.load().cgs_model_verify_counts(self._h, c_vp(histogram_address) if histogram_address else None, out))
        return {"n_dk": out[0], "n_wk": out[1], "n_k": out[2]}

    def state_hash(self):
        """{"hash_nwk", "sum_nwk", "sum_nk", "hash_nk"} of the device tables."""
        out = (ctypes.c_uint64 * 4)()
        check(_lib.load().cgs_model_state_hash(self._h, out))
        return {"hash_nwk": int(out[0]), "sum_nwk": int(out[1]), "sum_nk": int(out[2]), "hash_nk": int(out[3])}

    # -- AD-LDA exchange -------------------------------------------------------------------
    def allreduce_delta(self, nccl_comm: int):
        """The synchronous exchange step behind the ABI (delta, ncclAllReduce on ``nccl_comm``, apply)."""
        check(_lib.load().cgs_model_allreduce_delta(self._h, c_vp(nccl_comm)))

    def exchange_arena(self):
        """(address, nbytes) of this rank's arena of the asynchronous peer-memory exchange."""
        p, n = c_vp(), ctypes.c_size_t()
        check(_lib.load().cgs_model_exchange_arena(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def exchange_open(self, arena_addresses, rank: int):
        arr = (c_vp * len(arena_addresses))(*[c_vp(a) for a in arena_addresses])
        check(_lib.load().cgs_model_exchange_open(self._h, arr, len(arena_addresses), int(rank)))

    def exchange_round(self, n_ctas: int = 0):
        check(_lib.load().cgs_model_exchange_round(self._h, int(n_ctas)))

    def exchange_wait(self):
        check(_lib.load().cgs_model_exchange_wait(self._h))

    def exchange_sync(self, n_ctas: int = 0):
        """The synchronous fused exchange (one kernel on the model's stream, between two sweeps)."""
        check(_lib.load().cgs_model_exchange_sync(self._h, int(n_ctas)))

    def sweep_overlapped(self, n_sweeps: int = 1, x_ctas: int = 0):
        """Sweeps whose n_wk exchange runs inside the sampling launches (two half-sweeps per sweep)."""
        check(_lib.load().cgs_model_sweep_overlapped(self._h, int(n_sweeps), int(x_ctas)))

    def sweep_blocked(self, n_sweeps: int = 1):
        """Sweeps as one launch per region block, no exchange (keeps the gathered slice of n_wk L2-sized)."""
        check(_lib.load().cgs_model_sweep_blocked(self._h, int(n_sweeps)))

    def exchange_flush(self):
        """Publish the rows of the last half-sweep (before reading counts or log-likelihoods)."""
        check(_lib.load().cgs_model_exchange_flush(self._h))

    def exchange_phase_times(self):
        """Microseconds of the last synchronous exchange per phase: delta, barrier, reduce, barrier, apply."""
        out = (ctypes.c_double * 5)()
        check(_lib.load().cgs_model_exchange_phase_times(self._h, out))
        return dict(zip(("delta", "barrier_a", "reduce", "barrier_b", "apply"), [round(x, 1) for x in out]))

    @staticmethod
    def exchange_sync_local(models, n_ctas: int = 0):
        arr = (c_vp * len(models))(*[m._h for m in models])
        check(_lib.load().cgs_exchange_sync_local(arr, len(models), int(n_ctas)))

    @staticmethod
    def exchange_round_local(models, n_ctas: int = 0):
        """One round of several ranks that share a device, as ONE cooperative launch."""
        arr = (c_vp * len(models))(*[m._h for m in models])
        check(_lib.load().cgs_exchange_round_local(arr, len(models), int(n_ctas)))

    def exchange_timeouts(self) -> int:
        """Barrier waits of the exchange kernels that gave up (0 = all rounds completed)."""
        n = ctypes.c_int32()
        check(_lib.load().cgs_model_exchange_status(self._h, ctypes.byref(n)))
        return n.value

    @property
    def padded_topics(self) -> int:
        kp = ctypes.c_int32()
        check(_lib.load().cgs_model_padded_topics(self._h, ctypes.byref(kp)))
        return kp.value

    _BUF = {"n_wk": (0, "<i4"), "snapshot": (1, "<i4"), "delta": (2, "<i4"), "n_k": (3, "<i4"),
            "n_dk": (4, "<i4"), "z": (5, "<u2"), "tok_region": (6, "<i4"), "cell_ptr": (7, "<i8")}

    def device_buffer(self, name: str):
        """(address, nbytes, zero-copy view with ``__cuda_array_interface__``) of a handle-owned buffer."""
        which, typestr = self._BUF[name]
        p, n = c_vp(), ctypes.c_size_t()
        check(_lib.load().cgs_model_device_buffer(self._h, which, ctypes.byref(p), ctypes.byref(n)))
        itemsize = int(typestr[2:])
        return p.value, n.value, _lib.DeviceArrayView(p.value, (n.value // itemsize,), typestr, owner=self)

    def snapshot(self):
        check(_lib.load().cgs_model_snapshot(self._h))

    def delta_compute(self):
        check(_lib.load().cgs_model_delta_compute(self._h))

    def delta_apply(self):
        check(_lib.load().cgs_model_delta_apply(self._h))

    def delta_apply_peers(self, addresses):
        arr = (c_vp * len(addresses))(*[c_vp(a) for a in addresses])
        check(_lib.load().cgs_model_delta_apply_peers(self._h, arr, len(addresses)))

    def close(self):
        if self._h is not None:
            check(_lib.load().cgs_model_destroy(self._h))
            self._h = None

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and _lib._lib is not None:
                _lib._lib.cgs_model_destroy(self._h)
                self._h = None
        except Exception:
            pass


class LDA:
    """Drop-in for ``lda.LDA`` as pycisTopic uses it (lda_models.py:248-287).

    Parameters follow the ``lda`` package.  ``fit`` accepts the cells x regions binary CSR that
    ``run_cgs_model`` passes; the sweeps run on ``device``.  After ``fit``:
    ``topic_word_``/``components_`` (K x W), ``doc_topic_`` (D x K), ``nzw_``, ``ndz_``, ``nz_``
    and ``loglikelihoods_`` (the true Griffiths-Steyvers log-likelihood every ``refresh``
    sweeps, starting with sweep 0, as the package logs it).
    """

    def __init__(self, n_topics, n_iter=2000, alpha=0.1, eta=0.01, random_state=None, refresh=10,
                 device=0, corpus: Corpus | None = None, exports="all", device_metrics=None, z_init=None,
                 keep_assignments=False, deterministic=0):
        self.n_topics = n_topics
        self.n_iter = n_iter
        self.alpha = alpha
        self.eta = eta
        self.random_state = random_state
        self.refresh = refresh
        self.device = device
        self._corpus = corpus
        # exports: "all" = every attribute of the lda protocol; "frames" = only what CistopicLDAModel keeps
        # (nz_, topic_region_ W x K, cell_topic_ K x D) -- run_cgs_models uses it together with
        # device_metrics = {"alpha": raw alpha, "eta": raw eta, "top_n": 20}, which makes fit() leave the
        # metric ingredients of lda_models.py:291-305 in ``metric_ingredients_`` (computed on the device).
        if exports not in ("all", "frames"):
            raise ValueError("exports must be 'all' or 'frames'")
        self._exports = exports
        self._device_metrics = device_metrics
        # z_init: topic of every token in (cell, region) order instead of the package's ``i mod K`` start -- how a
        # state trained elsewhere (a Mallet --output-state file, mallet_formats.py) is evaluated (n_iter = 0) or
        # continued on the device; keep_assignments leaves the final assignments in ``z_`` (with ``WS_`` / ``DS_``)
        self._z_init = z_init
        self._keep_assignments = bool(keep_assignments)
        # deterministic: 0 = the default live-count schedule (statistically reproducible); n > 0 = bit-reproducible
        # for a fixed random_state, like the `lda` package: every sweep runs as n sub-sweeps on copies of the count
        # tables (True = 16).  Slower and slightly behind in the transient.
        self._deterministic = 16 if deterministic is True else int(deterministic or 0)
        if alpha <= 0 or eta <= 0:
            raise ValueError("alpha and eta must be greater than zero")
        if random_state is None:
            self._seed = int(np.random.SeedSequence().generate_state(1, np.uint64)[0])
        elif isinstance(random_state, (int, np.integer)):
            self._seed = int(random_state)
        else:  # numpy RandomState, as lda.utils.check_random_state accepts
            self._seed = int(random_state.randint(0, 2 ** 31 - 1))

    def fit(self, X, y=None):
        self._fit(X)
        return self

    def fit_transform(self, X, y=None):
        self._fit(X)
        return self.doc_topic_

    def _fit(self, X):
        own_corpus = False
        corpus = self._corpus
        if corpus is None:
            corpus = Corpus.from_cells_by_regions(X, device=self.device)
            own_corpus = True
        model = DeviceModel(corpus, self.n_topics, self.alpha, self.eta, seed=self._seed)
        try:
            if self._deterministic:
                model.set_deterministic(self._deterministic)
            if self._z_init is None:
                model.init()
            else:
                model.set_z(self._z_init)
            self.loglikelihoods_ = []
            done = 0
            refresh = max(1, int(self.refresh))
            import time as _time
            model.sync()
            t_fit = _time.perf_counter()
            while done < self.n_iter:
                if done % refresh == 0:
                    ll = model.loglikelihood()
                    logger.info("<%d> log likelihood: %.0f", done, ll)
                    self.loglikelihoods_.append(ll)
                step = min(refresh - done % refresh, self.n_iter - done)
                model.sweep(step)
                done += step
            ll = model.loglikelihood()
            logger.info("<%d> log likelihood: %.0f", self.n_iter - 1, ll)
            self.final_loglikelihood_ = ll
            # sampling rate of this fit (the log-likelihood evaluations included), for the "lda" logger and bench
            dt = max(_time.perf_counter() - t_fit, 1e-9)
            self.tokens_per_second_ = corpus.n_tokens * self.n_iter / dt
            logger.info("%d sweeps over %d tokens in %.3f s: %.2f Gtok/s per sweep", self.n_iter, corpus.n_tokens, dt,
                        self.tokens_per_second_ / 1e9)
            if self._exports == "all":
                self.nzw_, self.ndz_, self.nz_ = model.counts()
                self.components_, self.doc_topic_ = model.distributions()
                self.topic_word_ = self.components_
            else:
                self.nz_ = model.counts(nzw=False, ndz=False)[2]
            # DataFrame-ready layouts straight from the device (no host transpose)
            self.topic_region_, self.cell_topic_ = model.frames()
            if self._device_metrics is not None:
                dm = self._device_metrics
                ing = model.metric_ingredients(int(dm.get("top_n", 20)))
                ing["loglikelihood_compat"] = model.loglikelihood_compat(dm["alpha"], dm["eta"])
                ing["doc_lengths"] = corpus.cell_sizes().astype(np.float64)
                self.metric_ingredients_ = ing
            if self._keep_assignments:
                self.z_ = model.get_z()
                self.WS_, self.DS_ = corpus.token_lists()
            self.launch_count_ = model.launch_count
        finally:
            model.close()
            if own_corpus:
                corpus.close()
        return self

    def loglikelihood(self):
        """True log-likelihood of the fitted counts (lda.LDA.loglikelihood)."""
        return self.final_loglikelihood_
