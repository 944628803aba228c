"""The normalised consumers of the imputed matrix with one PROCESS per GPU (SURVEY.md section 8e iii).

Inside one process ``FactoredImputedFeatures(device=[0, 1, ...])`` shards the regions over host threads and adds the
per-cell totals up on the host.  Under ``torch.distributed`` (torchrun: one rank per GPU) the same three steps are

    1. every rank imputes ITS contiguous range of regions and sums each cell's column      (cgs_impute_column_totals)
    2. ``all_reduce(SUM)`` of the n_cells totals -- the only exchange of the path (NCCL or gloo)
    3. every rank normalises and reduces its regions against the global totals              (cgs_impute_row_stats_shard)

followed by an ``all_gather`` of the per-region results (a few floats per region) so that every rank can run the host
selection of ``find_highly_variable_features``.  ``backend`` is the object that talks to the device; tests replace it
with a NumPy stand-in to check the exchange logic with world_size-2 ``gloo`` on the CPU.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import c_f32p, c_u8p, c_vp, check, ptr


def region_range(n_regions: int, rank: int, world: int):
    """Contiguous range of regions of ``rank`` (the same split :func:`diff_features._shards` makes)."""
    bounds = np.linspace(0, n_regions, world + 1).astype(np.int64)
    return int(bounds[rank]), int(bounds[rank + 1])


class DeviceBackend:
    """The two ABI calls of a rank, on CUDA device ``device``."""

    def __init__(self, device=0, chunk_size=20000):
        _lib.require_device()
        self.device, self.chunk_size = int(device), int(chunk_size)

    def _operands(self, topic_region, cell_topic, impute_scale, as_int32):
        R, K = topic_region.shape
        return (self.device, ptr(topic_region, c_f32p), ptr(cell_topic, c_f32p), R, K, cell_topic.shape[1],
                float(impute_scale), int(as_int32))

    def column_totals(self, topic_region, cell_topic, impute_scale, as_int32):
        C = cell_topic.shape[1]
        totals = np.empty(C, np.int64 if as_int32 else np.float64)
        flags = np.empty(topic_region.shape[0], np.uint8)
        check(_lib.load().cgs_impute_column_totals(*self._operands(topic_region, cell_topic, impute_scale, as_int32),
                                                   self.chunk_size, totals.ctypes.data_as(c_vp), ptr(flags, c_u8p)))
        return totals, flags

    def row_stats(self, topic_region, cell_topic, impute_scale, as_int32, norm_scale, totals):
        R = topic_region.shape[0]
        mean, var, flags = np.empty(R, np.float32), np.empty(R, np.float32), np.empty(R, np.uint8)
        check(_lib.load().cgs_impute_row_stats_shard(*self._operands(topic_region, cell_topic, impute_scale, as_int32),
                                                     float(norm_scale), self.chunk_size, totals.ctypes.data_as(c_vp),
                                                     ptr(mean, c_f32p), ptr(var, c_f32p), ptr(flags, c_u8p)))
        return mean, var


def distributed_normalized_row_stats(topic_region, cell_topic, scale_factor=10 ** 6, norm_scale=10 ** 4, backend=None,
                                     group=None):
    """``(mean, var, row_nonzero)`` over ALL regions, identical on every rank: float32 per-region mean and variance of
    ``normalize_scores(impute_accessibility(...), norm_scale)`` (what ``find_highly_variable_features`` reduces,
    diff_features.py:636-639) and the row filter of ``impute_accessibility`` (:468-486).  Every rank passes the same
    full ``topic_region`` (regions x topics) and ``cell_topic`` (topics x cells) and works on its own range of
    regions; call inside an initialised ``torch.distributed`` process group."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    topic_region = np.ascontiguousarray(topic_region, dtype=np.float32)
    cell_topic = np.ascontiguousarray(cell_topic, dtype=np.float32)
    as_int32 = isinstance(scale_factor, int) and scale_factor != 1
    impute_scale = float(np.float32(scale_factor)) if as_int32 else 1.0
    if backend is None:
        backend = DeviceBackend(device=torch.cuda.current_device() if torch.cuda.is_available() else 0)
    R = topic_region.shape[0]
    r0, r1 = region_range(R, rank, world)
    mine = topic_region[r0:r1]
    C = cell_topic.shape[1]
    if r1 > r0:
        totals, flags = backend.column_totals(mine, cell_topic, impute_scale, as_int32)
    else:  # more ranks than regions: this rank contributes nothing, but takes part in the collectives
        totals, flags = np.zeros(C, np.int64 if as_int32 else np.float64), np.zeros(0, np.uint8)
    # the one exchange step of the path: n_cells values
    exchanged = torch.from_numpy(totals)
    on_gpu = dist.get_backend(group) == "nccl"
    if on_gpu:
        exchanged = exchanged.cuda()
    dist.all_reduce(exchanged, op=dist.ReduceOp.SUM, group=group)
    totals = np.ascontiguousarray(exchanged.cpu().numpy())
    if r1 > r0:
        mean, var = backend.row_stats(mine, cell_topic, impute_scale, as_int32, norm_scale, totals)
    else:
        mean = var = np.zeros(0, np.float32)
    # every rank gets every region's result (ranges differ by at most one region: pad to the largest)
    width = max(region_range(R, r, world)[1] - region_range(R, r, world)[0] for r in range(world))
    packed = torch.zeros(3, width, dtype=torch.float32)
    packed[0, : r1 - r0] = torch.from_numpy(np.asarray(mean))
    packed[1, : r1 - r0] = torch.from_numpy(np.asarray(var))
    packed[2, : r1 - r0] = torch.from_numpy(flags.astype(np.float32))
    if on_gpu:
        packed = packed.cuda()
    gathered = [torch.empty_like(packed) for _ in range(world)]
    dist.all_gather(gathered, packed, group=group)
    out = np.empty((3, R), np.float32)
    for r, t in enumerate(gathered):
        a, b = region_range(R, r, world)
        out[:, a:b] = t.cpu().numpy()[:, : b - a]
    return out[0], out[1], out[2] != 0
