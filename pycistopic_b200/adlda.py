"""One large model sharded by cells over several GPUs (AD-LDA).

One process per GPU (``torchrun``); ``torch.distributed`` is the plumbing.  Cells are split into
contiguous, token-balanced ranges; every rank keeps its token/topic shard and its ``n_dk`` rows
private, and a full replica of ``n_wk`` / ``n_k``.  After every sweep (or every 1/``sync_per_sweep``
of a sweep) the ranks exchange their ``n_wk`` deltas:

* ``exchange="nccl"``: ``delta = n_wk - snapshot`` (kernel) -> ``all_reduce(delta)`` (NCCL over
  NVLink) -> ``n_wk = snapshot + delta; snapshot = n_wk; n_k = colsum`` (one kernel);
* ``exchange="auto"`` (default): ``"overlap"`` (``"fused"`` when ``sync_per_sweep > 1``) -- measured on 8 B200s,
  C3 / K = 60 (128 MB table): 0.02 ms of the exchange exposed per sweep against 0.48 (fused), 0.55 (NCCL) and
  1.44 ms (round 1's peer kernel).
* ``exchange="peer"``: every rank maps the other ranks' delta buffers (CUDA IPC) and ONE kernel
  reads them over NVLink, reduces, applies and rebuilds ``n_k``; NCCL only provides the two
  stream-ordered barriers around it.
* ``exchange="nccl_abi"``: the same three steps as ``"nccl"`` but entirely behind the C ABI
  (``cgs_model_allreduce_delta`` with the process group's ``ncclComm_t``), what a non-Python host calls.
* ``exchange="fused"``: the synchronous fused round kernel (``cgs_model_exchange_sync``): ONE kernel between two
  sweeps does delta, reduce-scatter + all-gather through the peers' arenas over NVLink, apply and the ``n_k``
  rebuild -- 5 table passes of local traffic instead of the 8 of the NCCL path, no library call.
* ``exchange="overlap"``: the same fused exchange INSIDE the sampling launches (``cgs_model_sweep_overlapped``):
  a sweep is cut at a region index into two half-sweeps, and while one half is sampled the rows of ``n_wk`` of the
  other half are exchanged by a few blocks of the same grid.  Statistically the synchronous scheme (every row is
  exchanged once per sweep, right after it was sampled), but the exchange hides behind the sampling.
* ``exchange="async"``: the fused peer-memory round kernel (``csrc/exchange_kernels.cuh``) on a
  high-priority stream BESIDE the sweep: the round that publishes sweep s runs while sweep s + 1
  samples, so the exchange costs no time of its own.  With ``sync_per_sweep=1`` a rank sees the other ranks'
  counts one to two sweeps late (the synchronous schemes: exactly one), which lags the transient by a few per
  cent; with ``sync_per_sweep=2`` (a round after each half of the cell queue) half a sweep to one sweep late.
  ``rounds_per_sweep`` rounds can be chained per sweep part.

The reference has no equivalent (its multi-core analogue is Mallet's per-iteration merge of
per-thread count copies, lda_models.py:572-573).  The orchestration below is independent of the
device: tests drive it with a CPU shard backend under ``gloo``.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sparse


def balanced_cell_partition(tokens_per_cell, n_parts):
    """Contiguous cell ranges with nearly equal token counts; returns int64[n_parts + 1]."""
    cell_ptr = np.concatenate([[0], np.cumsum(np.asarray(tokens_per_cell, dtype=np.int64))])
    D = len(cell_ptr) - 1
    N = int(cell_ptr[-1])
    targets = np.arange(1, n_parts, dtype=np.float64) * N / n_parts
    cuts = np.clip(np.searchsorted(cell_ptr, targets, side="left"), 0, D)
    ptr = np.concatenate([[0], cuts, [D]]).astype(np.int64)
    # every part needs at least one cell
    for i in range(1, len(ptr)):
        if ptr[i] <= ptr[i - 1]:
            ptr[i] = min(D, ptr[i - 1] + 1)
    for i in range(len(ptr) - 2, -1, -1):
        if ptr[i] >= ptr[i + 1]:
            ptr[i] = max(0, ptr[i + 1] - 1)
    return ptr, cell_ptr


def balanced_region_blocks(tokens_per_region, n_blocks):
    """Region-block bounds (int32[n_blocks + 1], 0 ... n_regions) with nearly equal token counts per block."""
    cum = np.cumsum(np.asarray(tokens_per_region, dtype=np.int64))
    W = len(cum)
    n_blocks = max(1, min(int(n_blocks), W))
    cuts = np.searchsorted(cum, cum[-1] * np.arange(1, n_blocks) / n_blocks) + 1
    bounds = np.concatenate([[0], np.clip(cuts, 0, W), [W]]).astype(np.int64)
    return np.maximum.accumulate(bounds).astype(np.int32)


def region_blocks_for_table(n_regions, padded_topics):
    """How many region blocks an overlapped sweep uses: two -- what the overlap needs.  More blocks would keep the slice
    of n_wk a launch gathers from inside the L2 (C4, K = 100: 512 MB per rank), but the sweep is bound by issue slots and
    the load/store path, not by DRAM: measured on a C4 shard 18.8 / 18.6 / 18.8 / 19.1 / 19.5 ms per sweep for 1 / 2 / 4 /
    6 / 8 blocks (profiles/r2_region_blocks.txt)."""
    return 2


class GpuShard:
    """Shard backend on one GPU: a :class:`DeviceModel` over this rank's cells."""

    def __init__(self, binary_matrix, cell_range, token_offset, n_topics, alpha, eta, seed, device, exchange,
                 corpus=None):
        import torch

        from .lda import Corpus, DeviceModel
        self.torch = torch
        self.device = device
        torch.cuda.set_device(device)
        self.stream = torch.cuda.Stream(device=device)
        torch.cuda.set_stream(self.stream)
        if corpus is None:
            corpus = Corpus.from_regions_by_cells(binary_matrix, device=device, cell_range=cell_range)
            corpus.set_token_offset(token_offset)
        self.corpus = corpus
        self.model = DeviceModel(self.corpus, n_topics, alpha, eta, seed=seed, stream=self.stream.cuda_stream)
        self.exchange = exchange
        self._seed = seed
        self._delta_t = None
        self._peers = None
        self._token = None

    def set_global_size(self, n_cells, n_tokens):
        """Must precede the model: rebuilds it so that its concurrency caps follow the whole matrix."""
        self.corpus.set_global_size(n_cells, n_tokens)
        m = self.model
        from .lda import DeviceModel
        self.model = DeviceModel(self.corpus, m.n_topics, m.alpha, m.eta, seed=self._seed, stream=self.stream.cuda_stream)
        m.close()

    def init(self):
        self.model.init()

    def sweep_part(self, part, n_parts):
        self.model.sweep_part(part, n_parts)

    def delta_compute(self):
        self.model.delta_compute()

    def delta_tensor(self):
        if self._delta_t is None:
            _, _, view = self.model.device_buffer("delta")
            self._delta_t = self.torch.as_tensor(view, device=self.torch.device("cuda", self.device))
        return self._delta_t

    def delta_apply(self):
        self.model.delta_apply()

    # -- peer-memory exchange -------------------------------------------------------------
    def ipc_handle(self) -> bytes:
        import ctypes

        from . import _lib
        addr, _, _ = self.model.device_buffer("delta")
        buf = ctypes.create_string_buffer(64)
        _lib.check(_lib.load().cgs_ipc_export(ctypes.c_void_p(addr), buf))
        return buf.raw

    def open_peers(self, handles, rank):
        import ctypes

        from . import _lib
        addrs = []
        for r, h in enumerate(handles):
            if r == rank:
                addrs.append(self.model.device_buffer("delta")[0])
            else:
                p = ctypes.c_void_p()
                _lib.check(_lib.load().cgs_ipc_open(self.device, ctypes.create_string_buffer(h, 64), ctypes.byref(p)))
                addrs.append(p.value)
        self._peers = addrs

    def delta_apply_peers(self):
        self.model.delta_apply_peers(self._peers)

    # -- asynchronous fused exchange (csrc/exchange_kernels.cuh) ------------------------------
    def arena_handle(self) -> bytes:
        import ctypes

        from . import _lib
        addr, _ = self.model.exchange_arena()
        buf = ctypes.create_string_buffer(64)
        _lib.check(_lib.load().cgs_ipc_export(ctypes.c_void_p(addr), buf))
        return buf.raw

    def open_arenas(self, handles, rank):
        import ctypes

        from . import _lib
        own = self.model.exchange_arena()[0]
        addrs = []
        for r, h in enumerate(handles):
            if r == rank:
                addrs.append(own)
            else:
                p = ctypes.c_void_p()
                _lib.check(_lib.load().cgs_ipc_open(self.device, ctypes.create_string_buffer(h, 64), ctypes.byref(p)))
                addrs.append(p.value)
        self._arenas = addrs
        self.model.exchange_open(addrs, rank)

    def exchange_round(self, n_ctas=0):
        self.model.exchange_round(n_ctas)

    def exchange_sync(self, n_ctas=0):
        self.model.exchange_sync(n_ctas)

    def set_region_split(self, region_split):
        self.corpus.set_region_split(region_split)

    def set_region_blocks(self, bounds):
        self.corpus.set_region_blocks(bounds)

    def sweep_overlapped(self, x_ctas=0):
        self.model.sweep_overlapped(1, x_ctas)

    def exchange_flush(self):
        self.model.exchange_flush()

    def exchange_wait(self):
        self.model.exchange_wait()

    def allreduce_delta(self, comm_ptr):
        self.model.allreduce_delta(comm_ptr)

    def barrier_token(self):
        if self._token is None:
            self._token = self.torch.zeros(1, dtype=self.torch.int32, device=self.torch.device("cuda", self.device))
        return self._token

    # -- results --------------------------------------------------------------------------
    def loglik_cells(self):
        return self.model.loglikelihood(1)

    def loglik_regions(self):
        return self.model.loglikelihood(2)

    def export(self):
        return self.model.counts()

    def sync(self):
        self.model.sync()

    def close_peers(self):
        """Unmap the other ranks' buffers.  Every rank must have done this before any rank frees its own buffers
        (callers put a barrier between close_peers() and close())."""
        import ctypes

        from . import _lib
        self.model.sync()
        if self._peers is not None:
            own = self.model.device_buffer("delta")[0]
            for a in self._peers:
                if a != own:
                    _lib.load().cgs_ipc_close(self.device, ctypes.c_void_p(a))
            self._peers = None
        if getattr(self, "_arenas", None) is not None:
            own = self.model.exchange_arena()[0]
            for a in self._arenas:
                if a != own:
                    _lib.load().cgs_ipc_close(self.device, ctypes.c_void_p(a))
            self._arenas = None

    def close(self):
        self.close_peers()
        self.model.close()
        self.corpus.close()


class Exchanger:
    """The n_wk exchange of one shard (see the module docstring).  ``fit_adlda`` calls ``before_sweep`` /
    ``after_sweep`` around every sweep part and ``drain`` before it reads the counts."""

    def __init__(self, shard, exchange="auto", rounds_per_sweep=1, n_ctas=0):
        import torch.distributed as dist
        self.dist = dist
        self.shard = shard
        distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size() if distributed else 1
        self.rank = dist.get_rank() if distributed else 0
        if exchange == "auto":  # as fit_adlda resolves it when the sweep is not cut into parts
            exchange = "overlap" if self.world > 1 and hasattr(shard, "arena_handle") else "nccl"
        self.mode = exchange
        self.rounds_per_sweep = max(1, int(rounds_per_sweep))
        self.n_ctas = n_ctas
        self._dirty = True  # sweeps (or the initial counts) not yet published
        self.use_peer = exchange == "peer" and self.world > 1 and hasattr(shard, "ipc_handle")
        self.use_async = exchange == "async" and hasattr(shard, "arena_handle")
        self.use_overlap = exchange == "overlap" and hasattr(shard, "arena_handle")
        self.use_fused = (exchange == "fused" or self.use_overlap) and hasattr(shard, "arena_handle")
        self.comm_ptr = None
        if self.use_peer:
            handles = [None] * self.world
            dist.all_gather_object(handles, shard.ipc_handle())
            shard.open_peers(handles, self.rank)
        elif self.use_async or self.use_fused:
            handles = [shard.arena_handle()]
            if self.world > 1:
                handles = [None] * self.world
                dist.all_gather_object(handles, shard.arena_handle())
            shard.open_arenas(handles, self.rank)
            if self.world > 1:
                dist.barrier()  # every arena is mapped (and zeroed) before the first round touches a peer
        elif exchange == "nccl_abi" and self.world > 1 and hasattr(shard, "allreduce_delta"):
            backend = dist.distributed_c10d._get_default_group()._get_backend(shard.torch.device("cuda", shard.device))
            # the communicator is created lazily by the first collective
            dist.all_reduce(shard.barrier_token())
            self.comm_ptr = backend._comm_ptr()

    def step(self):
        """One synchronous exchange: complete (on the stream) before anything enqueued afterwards."""
        shard, dist = self.shard, self.dist
        if self.use_async:
            shard.exchange_round(self.n_ctas)
            shard.exchange_wait()
        elif self.use_fused:
            shard.exchange_sync(self.n_ctas)
        elif self.comm_ptr is not None:
            shard.allreduce_delta(self.comm_ptr)
        else:
            shard.delta_compute()
            if self.world == 1:
                shard.delta_apply()
            elif self.use_peer:
                dist.all_reduce(shard.barrier_token())   # every rank's delta is complete
                shard.delta_apply_peers()
                dist.all_reduce(shard.barrier_token())   # every rank has read every delta
            else:
                dist.all_reduce(shard.delta_tensor())
                shard.delta_apply()
        self._dirty = False

    def sweep(self, part=0, n_parts=1):
        """One sweep part with whatever exchange belongs to it."""
        if self.use_overlap:
            if n_parts != 1:
                raise ValueError("the overlapped exchange cuts the sweep itself: sync_per_sweep must be 1")
            self.shard.sweep_overlapped(self.n_ctas)
            self._dirty = True
            return
        self.before_sweep()
        self.shard.sweep_part(part, n_parts)
        self.after_sweep()

    def before_sweep(self):
        if self.use_async:
            # publishes everything up to the previous sweep while the next one samples
            for _ in range(self.rounds_per_sweep):
                self.shard.exchange_round(self.n_ctas)
            self._dirty = True

    def after_sweep(self):
        if self.use_async:
            self._dirty = True
        else:
            self.step()

    def drain(self):
        """All replicas identical and complete (needed before log-likelihoods and exports)."""
        if self.use_overlap:
            self.shard.exchange_flush()
            self._dirty = False
        elif self.use_async and self._dirty:
            self.step()
        elif self.use_async:
            self.shard.exchange_wait()


class AdLdaResult:
    """LDA-protocol result (lda_models.py:292-305, :345-357 read these attributes)."""


def fit_adlda(binary_matrix, n_topics, n_iter=150, alpha=0.1, eta=0.01, random_state=555, refresh=None,
              sync_per_sweep=1, exchange="auto", shard_factory=None, device=None, on_sweep=None, rounds_per_sweep=1):
    """Fit ONE model with the cells of ``binary_matrix`` (scipy CSR, regions x cells -- the
    layout of ``CistopicObject.binary_matrix``) sharded over the ranks of the default process
    group.  ``alpha`` / ``eta`` are the values the sampler uses.  Returns, on every rank, an object
    with ``nzw_``, ``ndz_``, ``nz_``, ``topic_word_``, ``doc_topic_``, ``loglikelihoods_``,
    ``final_loglikelihood_`` (full matrices, gathered)."""
    import torch.distributed as dist

    distributed = dist.is_available() and dist.is_initialized()
    world = dist.get_world_size() if distributed else 1
    rank = dist.get_rank() if distributed else 0
    if exchange not in ("auto", "nccl", "peer", "nccl_abi", "async", "fused", "overlap"):
        raise ValueError("exchange must be 'auto', 'nccl', 'nccl_abi', 'peer', 'fused', 'overlap' or 'async'")
    if sync_per_sweep < 1:
        raise ValueError("sync_per_sweep must be >= 1")
    M = sparse.csr_matrix(binary_matrix)
    W, D = M.shape
    if world > D:
        raise ValueError("more ranks than cells")
    tokens_per_cell = np.bincount(M.indices, minlength=D)
    part_ptr, cell_ptr = balanced_cell_partition(tokens_per_cell, world)
    cb, ce = int(part_ptr[rank]), int(part_ptr[rank + 1])
    token_offset = int(cell_ptr[cb])
    if shard_factory is None:
        if device is None:
            import os
            device = int(os.environ.get("LOCAL_RANK", "0"))

        def shard_factory(*a):
            return GpuShard(*a, device=device, exchange=exchange)
    if exchange == "auto" and world > 1:
        # the exchange inside the sampling launches when the sweep is not cut into parts, the fused kernel otherwise
        # (both need the ranks to map each other's memory; a CPU stand-in shard falls through to the plain all-reduce)
        exchange = "overlap" if sync_per_sweep == 1 else "fused"
    shard = shard_factory(M, (cb, ce), token_offset, n_topics, alpha, eta, random_state)
    try:
        if hasattr(shard, "set_global_size"):
            shard.set_global_size(D, int(cell_ptr[-1]))
        shard.init()
        exchanger = Exchanger(shard, exchange, rounds_per_sweep=rounds_per_sweep)
        if exchanger.use_overlap:
            # token-balanced region blocks of the WHOLE matrix (every rank computes the same ones)
            kp = shard.model.padded_topics if hasattr(shard, "model") else n_topics
            shard.set_region_blocks(balanced_region_blocks(np.diff(M.indptr), region_blocks_for_table(W, kp)))

        def loglik():
            exchanger.drain()
            cells = shard.loglik_cells()
            if world > 1:
                parts = [None] * world
                dist.all_gather_object(parts, cells)
                cells = float(np.sum(parts))
            return cells + shard.loglik_regions()

        exchanger.step()  # local init counts -> global n_wk / n_k on every rank
        res = AdLdaResult()
        res.loglikelihoods_ = []
        refresh = n_iter if refresh is None else max(1, int(refresh))
        for it in range(n_iter):
            if it % refresh == 0:
                res.loglikelihoods_.append(loglik())
            for part in range(sync_per_sweep):
                exchanger.sweep(part, sync_per_sweep)
            if on_sweep is not None:
                on_sweep(it, shard)
        res.final_loglikelihood_ = loglik()
        exchanger.drain()
        nzw, ndz_local, nz = shard.export()
        if world > 1:
            shards = [None] * world
            dist.all_gather_object(shards, ndz_local)
            ndz = np.concatenate(shards, axis=0)
        else:
            ndz = ndz_local
        res.nzw_, res.ndz_, res.nz_ = nzw, ndz, nz
        res.components_ = (nzw + eta) / (nz[:, None] + eta * W)
        res.topic_word_ = res.components_
        nd = tokens_per_cell.astype(np.float64)
        res.doc_topic_ = (ndz + alpha) / (nd[:, None] + alpha * n_topics)
        res.part_ptr_ = part_ptr
        if world > 1 and hasattr(shard, "close_peers"):
            shard.close_peers()
            dist.barrier()  # nobody frees a buffer that another rank still has mapped
        return res
    finally:
        shard.close()
