"""Drop-in mirror of the imputation entry points of ``pycisTopic.diff_features``:
:class:`CistopicImputedFeatures` (diff_features.py:26-59), :func:`impute_accessibility`
(:340-508) and :func:`normalize_scores` (:511-574).

The chunk product ``topic_region[chunk] @ cell_topic`` (:443-456), the ``* scale_factor`` /
``astype(int32)`` epilogue (:458-464) and the all-zero-row filter (:468-486) run on the GPU
through ``cgs_impute_plan_*`` (tcgen05 tensor-core kernel; the cell operand is prepared on the device
once and reused by every chunk); the return value has the reference's layout (a DENSE ndarray in
``mtx``, the kept region names, the cell names).
"""
from __future__ import annotations

import ctypes
import logging
import sys

import numpy as np
import pandas as pd
import scipy.sparse

from . import _lib
from ._lib import c_f32p, c_u8p, c_vp, check, ptr


def _as_devices(device):
    """``device`` arguments take one CUDA ordinal or a sequence of them (regions are sharded over the sequence)."""
    if isinstance(device, (int, np.integer)):
        return [int(device)]
    devices = [int(d) for d in device]
    if not devices:
        raise ValueError("need at least one device")
    return devices


def _shards(n, devices):
    """``[(device, start, stop), ...]``: contiguous, near-equal, non-empty ranges of ``range(n)``."""
    bounds = np.linspace(0, n, len(devices) + 1).astype(np.int64)
    return [(d, int(a), int(b)) for d, a, b in zip(devices, bounds[:-1], bounds[1:]) if b > a]


def _run_sharded(fn, shards):
    """``fn(*shard)`` for every shard, one host thread per shard (the library calls release the GIL; every shard
    drives its own device on its own stream); results in shard order."""
    if not shards:
        return []
    if len(shards) == 1:
        return [fn(*shards[0])]
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(len(shards)) as pool:
        return list(pool.map(lambda a: fn(*a), shards))


def get_position_index(query_list, target_list):
    """utils.py:105-110."""
    d = {k: idx for idx, k in enumerate(target_list)}
    return [d[k] for k in query_list]


def subset_list(target_list, index_list):
    """utils.py:113-114."""
    return list(map(target_list.__getitem__, index_list))


def non_zero_rows(matrix):
    """utils.py:117-126."""
    if isinstance(matrix, scipy.sparse.csr_matrix):
        matrix.eliminate_zeros()
        return np.nonzero(matrix.getnnz(axis=1))[0]
    matrix = np.asarray(matrix)
    if matrix.ndim == 2 and matrix.shape[1] and matrix.dtype.kind in "iu":
        # a row has a non-zero entry iff the OR of its entries is non-zero (3.5x faster than count_nonzero at 8 GB)
        return np.nonzero(np.bitwise_or.reduce(matrix, axis=1))[0]
    if matrix.ndim == 2 and matrix.shape[1] and matrix.dtype.kind == "f":
        # ... iff its maximum or its minimum is non-zero (NaN counts as non-zero, as for count_nonzero)
        return np.nonzero((matrix.max(axis=1) != 0) | (matrix.min(axis=1) != 0))[0]
    return np.nonzero(np.count_nonzero(matrix, axis=1))[0]


class CistopicImputedFeatures:
    """cisTopic imputation data class (diff_features.py:26-59) with ``subset`` (:61-129)."""

    def __init__(self, imputed_acc, feature_names, cell_names, project):
        self.mtx = imputed_acc
        self.feature_names = feature_names
        self.cell_names = cell_names
        self.project = project

    def __str__(self):
        return (f"CistopicImputedFeatures from project {self.project} with nCells × nFeatures = "
                f"{len(self.cell_names)} × {len(self.feature_names)}")

    def make_rankings(self, seed=123, device=0):
        """Per-cell rankings of the regions by descending score, 0 = best, ties broken at random
        (diff_features.py:274-337) -- on the GPU (``cgs_rank_columns``).

        Returns a :class:`CistopicImputedFeatures` holding the int32 ranking matrix, like the reference.  A column
        without ties gets the reference's ranking bit for bit; tied scores get a pseudo-random order that is a
        pure function of ``(seed, cell index)`` (the reference's comes from NumPy's PCG64 stream and its unstable
        argsort, which nothing outside NumPy reproduces).  Unlike the reference, whose per-column ``.toarray()``
        (:334) only works on a SciPy sparse ``mtx``, dense arrays are accepted as well."""
        if scipy.sparse.issparse(self.mtx):
            mtx = self.mtx.tocsc()
            n_rows, n_cols = mtx.shape
            out = np.empty((n_rows, n_cols), dtype=np.int32)
            step = max(1, min(n_cols, (1 << 28) // max(1, n_rows)))  # dense column blocks of about 1 GB
            for c0 in range(0, n_cols, step):
                block = mtx[:, c0:c0 + step].toarray()
                out[:, c0:c0 + step] = rank_columns(block, seed=seed, first_col=c0, device=device)
        else:
            out = rank_columns(self.mtx, seed=seed, device=device)
        return CistopicImputedFeatures(out, self.feature_names, self.cell_names, self.project)

    def subset(self, cells=None, features=None, copy=False, split_pattern="___"):
        """Keep the given cells and/or features, then drop all-zero features (diff_features.py:61-129)."""
        mtx, cell_names, feature_names = self.mtx, self.cell_names, self.feature_names
        if cells is not None:
            try:
                cells_index = get_position_index(cells, self.cell_names)
            except KeyError:
                tagged = [c.split(split_pattern)[0] if split_pattern in c else c for c in self.cell_names]
                cells_index = get_position_index(cells, tagged)
            mtx = mtx[:, cells_index]
            cell_names = subset_list(cell_names, cells_index)
        if features is not None:
            features_index = get_position_index(features, feature_names)
            mtx = mtx[features_index, :]
            feature_names = subset_list(feature_names, features_index)
        features_index = non_zero_rows(mtx)
        if len(features_index) != mtx.shape[0]:
            mtx = mtx[features_index, :]
            feature_names = subset_list(feature_names, features_index)
        if copy is True:
            return CistopicImputedFeatures(mtx, feature_names, cell_names, self.project)
        self.mtx, self.cell_names, self.feature_names = mtx, cell_names, feature_names


    def merge(self, cistopic_imputed_features_list, project="cisTopic_impute_merge", copy=False):
        """Append the cells of other :class:`CistopicImputedFeatures` (diff_features.py:131-272): features present in
        both keep their values side by side, a feature missing on one side gets zeros there.  Rows come out as
        [common, only in the objects so far, only in the added one]; the reference orders the common block by a Python
        ``set`` (arbitrary), here it keeps the order of the object merged into.  As in the reference a merge that had to
        add zero rows returns float64 (``np.zeros`` joins the stack)."""
        mtx, feature_names, cell_names = self.mtx, list(self.feature_names), list(self.cell_names)
        for other in cistopic_imputed_features_list:
            add_mtx, add_features = other.mtx, list(other.feature_names)
            cell_names = cell_names + list(other.cell_names)
            in_cur, in_add = set(feature_names), set(add_features)
            common = [f for f in feature_names if f in in_add]
            sparse_in = scipy.sparse.issparse(mtx)
            hstack = (lambda blocks: scipy.sparse.hstack(blocks, format="csr")) if sparse_in else np.hstack
            vstack = (lambda blocks: scipy.sparse.vstack(blocks, format="csr")) if sparse_in else np.vstack
            mtx_common = hstack([mtx[get_position_index(common, feature_names),],
                                 add_mtx[get_position_index(common, add_features),]])
            if in_cur != in_add:
                only_cur = sorted(in_cur - in_add)  # np.setdiff1d: sorted
                only_add = sorted(in_add - in_cur)
                diff_1 = hstack([mtx[get_position_index(only_cur, feature_names),],
                                 np.zeros((len(only_cur), add_mtx.shape[1]))])
                diff_2 = hstack([np.zeros((len(only_add), mtx.shape[1])),
                                 add_mtx[get_position_index(only_add, add_features),]])
                mtx = vstack([mtx_common, diff_1, diff_2])
                feature_names = common + only_cur + only_add
            else:
                mtx, feature_names = mtx_common, common
        if copy is True:
            return CistopicImputedFeatures(mtx, feature_names, cell_names, project)
        self.mtx, self.cell_names, self.feature_names, self.project = mtx, cell_names, feature_names, project


class FactoredImputedFeatures:
    """Imputed accessibility kept as its two factors (``impute_accessibility(..., materialize=False)``).

    The reference materialises ``topic_region @ cell_topic`` as one dense regions x cells array
    (diff_features.py:434-441, :503) -- 8 GB for 200 k regions x 10 k cells, 800 GB at atlas scale -- and
    ``normalize_scores`` / ``find_diff_features`` read it back.  This object holds only ``topic_region``
    (regions x topics) and ``cell_topic`` (topics x cells) in fp32; its consumers recompute the product chunk by
    chunk on the GPU and consume it there (``cgs_impute_normalized``, ``cgs_impute_ranksum``).  ``feature_names``
    lists, like the reference, only the regions with a non-zero imputed value (computed on first use);
    ``to_dense()`` gives the reference's :class:`CistopicImputedFeatures`.
    """

    def __init__(self, topic_region, cell_topic, region_names, cell_names, scale_factor, chunk_size, project,
                 device=0):
        self.topic_region = np.ascontiguousarray(topic_region, dtype=np.float32)
        self.cell_topic = np.ascontiguousarray(cell_topic, dtype=np.float32)
        if self.topic_region.shape[1] != self.cell_topic.shape[0]:
            raise ValueError(f"matmul: topic_region has {self.topic_region.shape[1]} topics, "
                             f"cell_topic has {self.cell_topic.shape[0]}")
        self.region_names = list(region_names)
        self.cell_names = list(cell_names)
        self.scale_factor = scale_factor
        self.chunk_size = int(chunk_size)
        self.project = project
        self.devices = _as_devices(device)      # regions (cells for make_rankings) are sharded over them
        self.device = self.devices[0]
        self._row_nonzero = None

    @property
    def as_int32(self):
        return isinstance(self.scale_factor, int) and self.scale_factor != 1

    @property
    def impute_scale(self):
        return float(np.float32(self.scale_factor)) if self.as_int32 else 1.0

    def __str__(self):
        return (f"FactoredImputedFeatures from project {self.project} with nCells × nFeatures = "
                f"{len(self.cell_names)} × {self.topic_region.shape[0]} (not materialised)")

    def _operands(self, device=None, r0=0, r1=None):
        """Leading ABI arguments for the regions [r0, r1) on ``device``."""
        R, K = self.topic_region.shape
        r1 = R if r1 is None else r1
        a = self.topic_region[r0:r1]
        return (self.device if device is None else device, ptr(a, c_f32p), ptr(self.cell_topic, c_f32p), r1 - r0, K,
                self.cell_topic.shape[1], self.impute_scale, int(self.as_int32))

    def _region_shards(self):
        return _shards(self.topic_region.shape[0], self.devices)

    def _column_totals(self, shards):
        """Per-cell totals of the whole imputed matrix + row flags: every shard's partial totals (one pass over its
        recomputed chunks), summed -- the one exchange step of the sharded normalisation (SURVEY.md 8e iii)."""
        R, C = self.topic_region.shape[0], self.cell_topic.shape[1]
        flags = np.empty(R, np.uint8)
        dt = np.int64 if self.as_int32 else np.float64

        def one(dev, r0, r1):
            tot = np.empty(C, dt)
            check(_lib.load().cgs_impute_column_totals(*self._operands(dev, r0, r1), self.chunk_size,
                                                       tot.ctypes.data_as(c_vp), ptr(flags[r0:r1], c_u8p)))
            return tot

        totals = np.sum(_run_sharded(one, shards), axis=0, dtype=dt)
        self._row_nonzero = flags.astype(bool)
        return np.ascontiguousarray(totals)

    def ranksum(self, contrasts_cols):
        """``[(fg column indices, bg column indices), ...]`` -> ``(statistic, pvalue, log2fc, row_nonzero)``,
        the first three ``n_contrasts x n_regions`` float64; one imputation pass serves every contrast."""
        _lib.require_device()
        R, C = self.topic_region.shape[0], self.cell_topic.shape[1]

        def pack(groups):
            offs = np.zeros(len(groups) + 1, np.int64)
            cols = []
            for i, g in enumerate(groups):
                g = np.asarray(g, dtype=np.int64)
                if g.size and (g.max() >= C or g.min() < -C):
                    raise IndexError(f"index {g.max() if g.max() >= C else g.min()} is out of bounds for axis 1 with size {C}")
                cols.append(np.sort(np.where(g < 0, g + C, g)).astype(np.int32))
                offs[i + 1] = offs[i] + g.size
            return np.ascontiguousarray(np.concatenate(cols) if cols else np.zeros(0, np.int32)), offs

        n = len(contrasts_cols)
        stat, pval, l2fc = (np.full((n, R), np.nan) for _ in range(3))
        flags = np.empty(R, np.uint8)
        live = [i for i, (f, b) in enumerate(contrasts_cols) if len(f) and len(b)]
        if live:
            fg, fo = pack([contrasts_cols[i][0] for i in live])
            bg, bo = pack([contrasts_cols[i][1] for i in live])
            def one(dev, r0, r1):  # regions are independent: a shard is a plain call on its rows
                s2, p2, l2 = (np.empty((len(live), r1 - r0), np.float64) for _ in range(3))
                check(_lib.load().cgs_impute_ranksum(*self._operands(dev, r0, r1), self.chunk_size, len(live),
                                                     ptr(fg, _lib.c_i32p), ptr(fo, _lib.c_i64p), ptr(bg, _lib.c_i32p),
                                                     ptr(bo, _lib.c_i64p), ptr(s2, _lib.c_f64p), ptr(p2, _lib.c_f64p),
                                                     ptr(l2, _lib.c_f64p), ptr(flags[r0:r1], c_u8p)))
                return s2, p2, l2

            shards = self._region_shards()
            for (dev, r0, r1), (s2, p2, l2) in zip(shards, _run_sharded(one, shards)):
                for dst, src in ((stat, s2), (pval, p2), (l2fc, l2)):
                    dst[np.ix_(live, range(r0, r1))] = src
            self._row_nonzero = flags.astype(bool)
        return stat, pval, l2fc, self.row_nonzero

    def normalized(self, scale_factor=10 ** 4):
        """``normalize_scores`` of the imputed matrix without materialising it: a dense
        :class:`CistopicImputedFeatures` (float64 for the int32 imputation, float32 otherwise)."""
        _lib.require_device()
        R, C = self.topic_region.shape[0], self.cell_topic.shape[1]
        out = np.empty((R, C), np.float64 if self.as_int32 else np.float32)
        shards = self._region_shards()
        if len(shards) == 1:
            flags = np.empty(R, np.uint8)
            kept = ctypes.c_int64()
            check(_lib.load().cgs_impute_normalized(*self._operands(), float(scale_factor), self.chunk_size,
                                                    out.ctypes.data_as(c_vp), ptr(flags, c_u8p), None, ctypes.byref(kept)))
            self._row_nonzero = flags.astype(bool)
            return CistopicImputedFeatures(out[: kept.value], self.feature_names, self.cell_names, self.project)
        totals = self._column_totals(shards)                 # sets the row flags
        first = np.concatenate([[0], np.cumsum(self._row_nonzero)])   # kept regions before every region

        def one(dev, r0, r1):
            kept = ctypes.c_int64()
            check(_lib.load().cgs_impute_normalized_shard(*self._operands(dev, r0, r1), float(scale_factor), self.chunk_size,
                                                          totals.ctypes.data_as(c_vp), out[first[r0]:].ctypes.data_as(c_vp),
                                                          None, ctypes.byref(kept)))
            if kept.value != first[r1] - first[r0]:
                raise RuntimeError("row filter changed between the two passes")

        _run_sharded(one, shards)
        return CistopicImputedFeatures(out[: first[-1]], self.feature_names, self.cell_names, self.project)

    def row_mean_var(self, norm_scale=0):
        """float32 per-region mean and variance over the cells (NumPy's ``mean`` / ``var`` with ``dtype=float32``) of
        the kept regions: of the imputed values (``norm_scale = 0``) or of ``normalize_scores(self, norm_scale)``;
        computed chunk by chunk on the device (``cgs_impute_row_stats``)."""
        _lib.require_device()
        R = self.topic_region.shape[0]
        mean, var = np.empty(R, np.float32), np.empty(R, np.float32)
        flags = np.empty(R, np.uint8)
        shards = self._region_shards()
        L = _lib.load()
        if len(shards) > 1 and norm_scale:
            totals = self._column_totals(shards)

            def one(dev, r0, r1):
                check(L.cgs_impute_row_stats_shard(*self._operands(dev, r0, r1), float(norm_scale), self.chunk_size,
                                                   totals.ctypes.data_as(c_vp), ptr(mean[r0:r1], c_f32p),
                                                   ptr(var[r0:r1], c_f32p), ptr(flags[r0:r1], c_u8p)))
        else:
            def one(dev, r0, r1):
                check(L.cgs_impute_row_stats(*self._operands(dev, r0, r1), float(norm_scale), self.chunk_size,
                                             ptr(mean[r0:r1], c_f32p), ptr(var[r0:r1], c_f32p), ptr(flags[r0:r1], c_u8p)))
        _run_sharded(one, shards)
        self._row_nonzero = flags.astype(bool)
        return mean[self._row_nonzero], var[self._row_nonzero]

    def make_rankings(self, seed=123, cell_block=0):
        """``make_rankings`` of the imputed matrix without materialising it (``cgs_impute_rankings``): the regions
        ``impute_accessibility`` would keep are ranked per cell, block of cells by block of cells, on the GPU; only
        the int32 ranking matrix comes back.  Same result as ``self.to_dense().make_rankings(seed)``."""
        _lib.require_device()
        keep = self.row_nonzero
        tr = np.ascontiguousarray(self.topic_region[keep])
        R, K = tr.shape
        C = self.cell_topic.shape[1]
        out = np.empty((R, C), np.int32)
        if R and C:
            def one(dev, c0, c1):  # cells are independent; the tie order is keyed by the global cell index
                ct = np.ascontiguousarray(self.cell_topic[:, c0:c1])
                check(_lib.load().cgs_impute_rankings(dev, ptr(tr, c_f32p), ptr(ct, c_f32p), R, K, c1 - c0,
                                                      self.impute_scale, int(self.as_int32), int(seed) & (2 ** 64 - 1),
                                                      int(cell_block), c0, out[:, c0:].ctypes.data_as(c_vp), C))

            _run_sharded(one, _shards(C, self.devices))
        return CistopicImputedFeatures(out, self.feature_names, self.cell_names, self.project)

    @property
    def row_nonzero(self):
        if self._row_nonzero is None:  # one imputation pass with a one-cell contrast is the cheapest way to the flags
            self.ranksum([([0], [self.cell_topic.shape[1] - 1])])
        return self._row_nonzero

    @property
    def feature_names(self):
        return [n for n, k in zip(self.region_names, self.row_nonzero) if k]

    def subset(self, cells=None, features=None, copy=False, split_pattern="___"):
        """Same selection as :meth:`CistopicImputedFeatures.subset`, applied to the factors."""
        tr, ct, regions, cell_names = self.topic_region, self.cell_topic, self.region_names, self.cell_names
        if cells is not None:
            try:
                idx = get_position_index(cells, cell_names)
            except KeyError:
                idx = get_position_index(cells, [c.split(split_pattern)[0] if split_pattern in c else c for c in cell_names])
            ct, cell_names = ct[:, idx], subset_list(cell_names, idx)
        if features is not None:
            idx = get_position_index(features, regions)
            tr, regions = tr[idx], subset_list(regions, idx)
        out = FactoredImputedFeatures(tr, ct, regions, cell_names, self.scale_factor, self.chunk_size, self.project,
                                      self.devices)
        if copy is True:
            return out
        self.__dict__.update(out.__dict__)

    def to_dense(self):
        imputed_acc, keep = calculate_imputed_accessibility(self.topic_region, self.cell_topic, self.region_names,
                                                            self.scale_factor, self.chunk_size, device=self.devices)
        return CistopicImputedFeatures(imputed_acc, keep, self.cell_names, self.project)


def _logger():
    level = logging.INFO
    log_format = "%(asctime)s %(name)-12s %(levelname)-8s %(message)s"
    handlers = [logging.StreamHandler(stream=sys.stdout)]
    logging.basicConfig(level=level, format=log_format, handlers=handlers)
    return logging.getLogger("cisTopic")


def calculate_imputed_accessibility(topic_region, cell_topic, region_names, scale_factor, chunk_size, device=0,
                                    log=None):
    """Chunked imputation on the GPU; same contract as the nested function at
    diff_features.py:398-492: returns ``(imputed_acc, region_names_to_keep)`` with all-zero rows
    dropped, int32 when ``scale_factor`` is an ``int`` other than 1, else float32."""
    _lib.require_device()
    L = _lib.load()
    topic_region = np.ascontiguousarray(topic_region, dtype=np.float32)
    cell_topic = np.ascontiguousarray(cell_topic, dtype=np.float32)
    R, K = topic_region.shape
    K2, C = cell_topic.shape
    if K != K2:
        raise ValueError(f"matmul: topic_region has {K} topics, cell_topic has {K2}")
    as_int = isinstance(scale_factor, int) and scale_factor != 1
    out_dtype = np.int32 if as_int else np.float32
    imputed_acc = np.empty((R, C), dtype=out_dtype)
    scale = float(np.float32(scale_factor)) if as_int else 1.0
    # Two passes over the (cheap) device product instead of the reference's chunk -> filter -> copy on the host
    # (:443-486): the first one yields the row filter, the second one writes every kept row straight to its final
    # place in ``imputed_acc`` (cgs_impute_rows).  Regions are sharded over ``device`` when it is a sequence.
    shards = _shards(R, _as_devices(device))
    flags = np.empty(R, np.uint8)
    totals_dtype = np.int64 if as_int else np.float64

    def first_pass(dev, r0, r1):
        tot = np.empty(C, totals_dtype)
        check(L.cgs_impute_column_totals(dev, ptr(topic_region[r0:r1], c_f32p), ptr(cell_topic, c_f32p), r1 - r0, K, C, scale,
                                         int(as_int), int(chunk_size), tot.ctypes.data_as(c_vp), ptr(flags[r0:r1], c_u8p)))

    _run_sharded(first_pass, shards)
    first = np.concatenate([[0], np.cumsum(flags, dtype=np.int64)])   # kept regions before every region

    def second_pass(dev, r0, r1):
        if log is not None:
            log.info(f"Impute region accessibility for regions {r0}-{r1}")
        written = ctypes.c_int64()
        check(L.cgs_impute_rows(dev, ptr(topic_region[r0:r1], c_f32p), ptr(cell_topic, c_f32p), r1 - r0, K, C, scale,
                                int(as_int), int(chunk_size), ptr(flags[r0:r1], c_u8p),
                                imputed_acc[first[r0]:].ctypes.data_as(c_vp), ctypes.byref(written)))
        if written.value != first[r1] - first[r0]:
            raise RuntimeError("row filter changed between the two passes")

    if first[-1]:
        _run_sharded(second_pass, shards)
    output_chunk_end = int(first[-1])
    region_names_to_keep = [region_names[i] for i in np.flatnonzero(flags)]
    imputed_acc = imputed_acc[0:output_chunk_end]
    return imputed_acc, region_names_to_keep


def impute_accessibility(
    cistopic_obj,
    selected_cells=None,
    selected_regions=None,
    scale_factor=10 ** 6,
    chunk_size=20000,
    project="cisTopic_Impute",
    device=0,
    materialize=True,
):
    """Impute region accessibility (diff_features.py:340-508).  ``materialize=False`` (extension) returns a
    :class:`FactoredImputedFeatures` whose consumers (``normalize_scores``, ``markers``, ``find_diff_features``)
    recompute the product on the GPU instead of reading a dense regions x cells array."""
    log = _logger()
    model = cistopic_obj.selected_model
    cell_names = cistopic_obj.cell_names
    cell_topic = model.cell_topic.loc[:, cell_names]
    region_names = cistopic_obj.region_names
    topic_region = model.topic_region.loc[region_names]
    if selected_cells is not None:
        cell_topic = cell_topic.loc[:, selected_cells]
        cell_names = selected_cells
    if selected_regions is not None:
        topic_region = topic_region.loc[selected_regions]
        region_names = selected_regions
    cell_topic = cell_topic.to_numpy().astype(np.float32)
    topic_region = topic_region.to_numpy().astype(np.float32)

    if not materialize:
        return FactoredImputedFeatures(topic_region, cell_topic, region_names, cell_names, scale_factor, chunk_size,
                                       project, device)
    log.info("Imputing region accessibility")
    imputed_acc, region_names_to_keep = calculate_imputed_accessibility(
        topic_region=topic_region,
        cell_topic=cell_topic,
        region_names=region_names,
        scale_factor=scale_factor,
        chunk_size=chunk_size,
        device=device,
        log=log,
    )
    imputed_acc_obj = CistopicImputedFeatures(imputed_acc, region_names_to_keep, cell_names, project)
    log.info("Done!")
    return imputed_acc_obj


_NORM_DTYPES = {np.dtype(np.int32): (0, np.float64), np.dtype(np.float32): (1, np.float32),
                np.dtype(np.float64): (2, np.float64)}


def calculate_normalized_scores(imputed_acc, scale_factor, device=0):
    """diff_features.py:540-548: ``X / (colsum(X) / scale_factor)``, then ``log1p`` -- on the GPU
    (``cgs_normalize_scores``: one pass for the per-cell totals, one for the quotient and log1p, fp64 in the
    reference's operation order).  int32 and float64 input give float64, float32 gives float32, as NumPy
    does; other dtypes are converted to float64 first, which is what NumPy's true division produces."""
    _lib.require_device()
    x = np.asarray(imputed_acc)
    if x.ndim != 2:
        raise ValueError("imputed_acc must be a 2-D array (regions x cells)")
    if x.dtype not in _NORM_DTYPES:
        x = x.astype(np.float64)
    x = np.ascontiguousarray(x)
    code, out_dtype = _NORM_DTYPES[x.dtype]
    out = np.empty(x.shape, dtype=out_dtype)
    if x.size == 0:
        return out
    check(_lib.load().cgs_normalize_scores(device, x.ctypes.data_as(c_vp), code, x.shape[0], x.shape[1],
                                           float(scale_factor), out.ctypes.data_as(c_vp)))
    return out


def normalize_scores(imputed_acc, scale_factor=10 ** 4, materialize=True):
    """Log-normalise imputation data (diff_features.py:511-574); the arithmetic runs on the GPU
    (SURVEY.md section 8f, row N1).  ``materialize=False`` (factored input only) returns a
    :class:`FactoredNormalizedFeatures` that :func:`find_highly_variable_features` reduces on the device without the
    normalised matrix ever existing."""
    log = _logger()
    log.info("Normalizing imputed data")
    if isinstance(imputed_acc, FactoredImputedFeatures):
        output = imputed_acc.normalized(scale_factor) if materialize else FactoredNormalizedFeatures(imputed_acc, scale_factor)
    elif isinstance(imputed_acc, CistopicImputedFeatures):
        mtx = imputed_acc.mtx.toarray() if scipy.sparse.issparse(imputed_acc.mtx) else imputed_acc.mtx
        output = CistopicImputedFeatures(calculate_normalized_scores(mtx, scale_factor), imputed_acc.feature_names,
                                         imputed_acc.cell_names, imputed_acc.project)
    elif isinstance(imputed_acc, pd.DataFrame):
        output = pd.DataFrame(calculate_normalized_scores(imputed_acc.to_numpy(), scale_factor),
                              index=imputed_acc.index, columns=imputed_acc.columns)
    else:
        raise TypeError("imputed_acc must be a CistopicImputedFeatures or a pandas DataFrame")
    log.info("Done!")
    return output


class FactoredNormalizedFeatures:
    """``normalize_scores(impute_accessibility(..., materialize=False), scale_factor, materialize=False)``: the
    normalised matrix as a recipe (the two factors + the scale factor).  ``to_dense()`` gives what
    ``normalize_scores`` returns; ``row_mean_var()`` is what ``find_highly_variable_features`` needs of it."""

    def __init__(self, factored, scale_factor):
        self.factored = factored
        self.scale_factor = scale_factor
        self.cell_names = factored.cell_names
        self.project = factored.project

    @property
    def feature_names(self):
        return self.factored.feature_names

    def to_dense(self):
        return self.factored.normalized(self.scale_factor)

    def row_mean_var(self):
        return self.factored.row_mean_var(self.scale_factor)


def row_mean_var(matrix, device=0):
    """``(np.mean(matrix, axis=1, dtype=np.float32), np.var(matrix, axis=1, dtype=np.float32))`` on the GPU
    (``cgs_row_mean_var``; diff_features.py:636-639)."""
    _lib.require_device()
    x = np.asarray(matrix)
    if x.ndim != 2:
        raise ValueError("matrix must be a 2-D array (regions x cells)")
    if x.dtype not in _RANK_DTYPES:
        x = x.astype(np.float64)
    if x.strides[1] != x.itemsize or x.strides[0] < x.shape[1] * x.itemsize or x.strides[0] % x.itemsize:
        x = np.ascontiguousarray(x)
    mean, var = np.empty(x.shape[0], np.float32), np.empty(x.shape[0], np.float32)
    if x.shape[0] == 0:
        return mean, var
    if x.shape[1] == 0:
        mean[:], var[:] = np.nan, np.nan
        return mean, var
    check(_lib.load().cgs_row_mean_var(device, _RANK_DTYPES[x.dtype], x.ctypes.data_as(c_vp), x.strides[0] // x.itemsize,
                                       x.shape[0], x.shape[1], ptr(mean, c_f32p), ptr(var, c_f32p)))
    return mean, var


def select_variable_features(mean, var, features, min_disp=0.05, min_mean=0.0125, max_disp=np.inf, max_mean=3,
                             n_bins=20, n_top_features=None, return_frame=False):
    """The Seurat-style selection of diff_features.py:641-699 on per-region means and variances (host work on
    vectors of length n_regions): log dispersion, ``n_bins`` equal-width mean bins, dispersion normalised by its
    bin's mean and standard deviation (a bin with one region: normalised dispersion 1), then either the
    ``n_top_features`` largest or the mean / dispersion window."""
    mean = np.array(mean, dtype=np.float32)
    mean[mean == 0] = 1e-12
    dispersion = np.asarray(var, dtype=np.float32) / mean
    dispersion[dispersion == 0] = np.nan
    with np.errstate(invalid="ignore", divide="ignore"):
        dispersion = np.log(dispersion)
    df = pd.DataFrame({"means": mean, "dispersions": dispersion})
    df["mean_bin"] = pd.cut(df["means"], bins=n_bins)
    grouped = df.groupby("mean_bin", observed=False)["dispersions"]
    bin_mean, bin_std = grouped.mean(), grouped.std(ddof=1)
    lonely = bin_std.isnull()                      # one region in the bin (or none)
    bin_std[lonely.values] = bin_mean[lonely.values].values
    bin_mean[lonely.values] = 0
    codes = df["mean_bin"].values
    df["dispersions_norm"] = (df["dispersions"].values - bin_mean[codes].values) / bin_std[codes].values
    norm = df["dispersions_norm"].values.astype("float32")
    if n_top_features is not None:
        ranked = norm[~np.isnan(norm)]
        ranked[::-1].sort()
        cut = ranked[n_top_features - 1]
        chosen = np.nan_to_num(df["dispersions_norm"].values) >= cut
    else:
        norm[np.isnan(norm)] = 0
        chosen = np.logical_and.reduce((mean > min_mean, mean < max_mean, norm > min_disp, norm < max_disp))
    df["highly_variable"] = chosen
    selected = [features[i] for i in np.flatnonzero(chosen)]
    return (selected, df) if return_frame else selected


def find_highly_variable_features(input_mat, min_disp=0.05, min_mean=0.0125, max_disp=np.inf, max_mean=3, n_bins=20,
                                  n_top_features=None, plot=True, save=None, device=0):
    """Find highly variable features (diff_features.py:577-714).  The pass over the matrix -- per-region mean and
    variance, :636-639 -- runs on the GPU: ``cgs_row_mean_var`` for a dense matrix / DataFrame, and straight from
    the factors (``cgs_impute_row_stats``; the matrix never exists) for a :class:`FactoredNormalizedFeatures` or a
    :class:`FactoredImputedFeatures`.  The selection on the two vectors is host work (:func:`select_variable_features`)."""
    log = _logger()
    if isinstance(input_mat, pd.DataFrame):
        features = input_mat.index.tolist()
        log.info("Calculating mean and variance")
        mean, var = row_mean_var(input_mat.values, device)
    elif isinstance(input_mat, (FactoredNormalizedFeatures, FactoredImputedFeatures)):
        features = input_mat.feature_names
        log.info("Calculating mean and variance")
        mean, var = input_mat.row_mean_var()
    else:
        features = input_mat.feature_names
        mat = input_mat.mtx
        log.info("Calculating mean and variance")
        if scipy.sparse.issparse(mat):   # sklearn.utils.sparsefuncs.mean_variance_axis in the reference (:633)
            mat = scipy.sparse.csr_matrix(mat)
            step = max(1, (1 << 27) // max(1, mat.shape[1]))
            parts = [row_mean_var(mat[r0:r0 + step].toarray(), device) for r0 in range(0, mat.shape[0], step)]
            mean, var = (np.concatenate([p[i] for p in parts]) if parts else np.empty(0, np.float32) for i in (0, 1))
        else:
            mean, var = row_mean_var(mat, device)
    var_features, df = select_variable_features(mean, var, features, min_disp, min_mean, max_disp, max_mean, n_bins,
                                                n_top_features, return_frame=True)
    if plot:
        try:
            import matplotlib
            import matplotlib.pyplot as plt
        except ImportError:
            log.warning("matplotlib is not installed: the mean / dispersion plot is skipped")
        else:
            fig = plt.figure()
            matplotlib.rcParams["agg.path.chunksize"] = 10000
            plt.scatter(df["means"], df["dispersions_norm"], c=df["highly_variable"], s=10, alpha=0.1)
            plt.xlabel("Mean measurement of features")
            plt.ylabel("Normalized dispersion of the features")
            if save is not None:
                fig.savefig(save)
            plt.show()
    log.info("Done!")
    return var_features


def rank_columns(scores, seed=123, first_col=0, device=0):
    """Ranking of every column of ``scores`` (regions x cells) by descending value, 0 = best, ties in the order of a
    keyed pseudo-random bijection (``cgs_rank_columns``; the closure at diff_features.py:294-306).  ``first_col`` is
    the cell index of column 0, so a matrix ranked in column blocks gives the same result as in one call."""
    _lib.require_device()
    x = np.asarray(scores)
    if x.ndim != 2:
        raise ValueError("scores must be a 2-D array (regions x cells)")
    if x.dtype not in _RANK_DTYPES:
        x = x.astype(np.float64)
    if x.strides[1] != x.itemsize or x.strides[0] < x.shape[1] * x.itemsize or x.strides[0] % x.itemsize:
        x = np.ascontiguousarray(x)
    out = np.empty(x.shape, dtype=np.int32)
    if x.size == 0:
        return out
    check(_lib.load().cgs_rank_columns(device, _RANK_DTYPES[x.dtype], x.ctypes.data_as(c_vp), x.strides[0] // x.itemsize,
                                       x.shape[0], x.shape[1], int(seed) & (2 ** 64 - 1), int(first_col),
                                       out.ctypes.data_as(_lib.c_i32p), out.shape[1]))
    return out


# ---- differential features (diff_features.py:716-1120) -----------------------------
_RANK_DTYPES = {np.dtype(np.int32): 0, np.dtype(np.float32): 1, np.dtype(np.float64): 2}


def _rank_matrix(a):
    a = np.asarray(a)
    if a.ndim != 2:
        raise ValueError("expected a 2-D matrix (features x cells)")
    if a.dtype not in _RANK_DTYPES:
        a = a.astype(np.float64)
    return np.ascontiguousarray(a)


def ranksum_rows(fg_mat, bg_mat=None, fg_cols=None, bg_cols=None, device=0):
    """Per row: ``scipy.stats.ranksums`` statistic and p-value and the reference's log2 fold change, on the GPU
    (``cgs_ranksum_rows``).  Either two matrices (features x fg cells, features x bg cells), or ONE matrix with
    the two column-index lists.  Returns ``(statistic, pvalue, log2fc)``, float64 each."""
    _lib.require_device()
    fg = _rank_matrix(fg_mat)
    bg = fg if bg_mat is None else _rank_matrix(bg_mat)
    if fg.shape[0] != bg.shape[0]:
        raise ValueError("Foreground matrix and background matrix have a different first dimension:"
                         f" {fg.shape[0]} vs {bg.shape[0]}")
    if bg is not fg and bg.dtype != fg.dtype:
        common = np.result_type(fg.dtype, bg.dtype)
        common = common if common in _RANK_DTYPES else np.dtype(np.float64)
        fg, bg = np.ascontiguousarray(fg.astype(common)), np.ascontiguousarray(bg.astype(common))

    def cols(c, width):
        if c is None:
            return None, width
        c = np.ascontiguousarray(np.asarray(c, dtype=np.int64))
        if c.size and (c.max() >= width or c.min() < -width):
            bad = c.max() if c.max() >= width else c.min()
            raise IndexError(f"index {bad} is out of bounds for axis 1 with size {width}")
        # the statistic does not depend on the order inside a group: ascending columns make the gathers of a
        # warp touch neighbouring addresses
        c = np.sort(np.where(c < 0, c + width, c)).astype(np.int32)
        return c, c.size

    fc, n_fg = cols(fg_cols, fg.shape[1])
    bc, n_bg = cols(bg_cols, bg.shape[1])
    R = fg.shape[0]
    stat, pval, l2fc = (np.full(R, np.nan) for _ in range(3))
    if R == 0 or n_fg == 0 or n_bg == 0:  # scipy / numpy give NaN for an empty group
        return stat, pval, l2fc
    check(_lib.load().cgs_ranksum_rows(
        device, _RANK_DTYPES[fg.dtype], R,
        fg.ctypes.data_as(c_vp), fg.shape[1], ptr(fc, _lib.c_i32p) if fc is not None else None, n_fg,
        bg.ctypes.data_as(c_vp), bg.shape[1], ptr(bc, _lib.c_i32p) if bc is not None else None, n_bg,
        ptr(stat, _lib.c_f64p), ptr(pval, _lib.c_f64p), ptr(l2fc, _lib.c_f64p)))
    return stat, pval, l2fc


def ranksum_contrasts(matrix, contrasts_cols, device=0):
    """``ranksum_rows(matrix, None, fg, bg)`` for every ``(fg column indices, bg column indices)`` pair, with ONE upload
    of the matrix (``cgs_ranksum_contrasts``): ``(statistic, pvalue, log2fc)``, each ``n_contrasts x n_rows`` float64;
    a contrast with an empty group gives NaN, as scipy / NumPy do."""
    _lib.require_device()
    x = _rank_matrix(matrix)
    R, C = x.shape
    n = len(contrasts_cols)
    stat, pval, l2fc = (np.full((n, R), np.nan) for _ in range(3))

    def pack(groups):
        offs = np.zeros(len(groups) + 1, np.int64)
        cols = []
        for i, g in enumerate(groups):
            g = np.asarray(g, dtype=np.int64)
            if g.size and (g.max() >= C or g.min() < -C):
                raise IndexError(f"index {g.max() if g.max() >= C else g.min()} is out of bounds for axis 1 with size {C}")
            cols.append(np.sort(np.where(g < 0, g + C, g)).astype(np.int32))
            offs[i + 1] = offs[i] + g.size
        return np.ascontiguousarray(np.concatenate(cols) if cols else np.zeros(0, np.int32)), offs

    live = [i for i, (f, b) in enumerate(contrasts_cols) if len(f) and len(b)]
    if live and R:
        fg, fo = pack([contrasts_cols[i][0] for i in live])
        bg, bo = pack([contrasts_cols[i][1] for i in live])
        s2, p2, l2 = (np.empty((len(live), R), np.float64) for _ in range(3))
        check(_lib.load().cgs_ranksum_contrasts(device, _RANK_DTYPES[x.dtype], R, x.ctypes.data_as(c_vp), C, len(live),
                                                ptr(fg, _lib.c_i32p), ptr(fo, _lib.c_i64p), ptr(bg, _lib.c_i32p),
                                                ptr(bo, _lib.c_i64p), ptr(s2, _lib.c_f64p), ptr(p2, _lib.c_f64p),
                                                ptr(l2, _lib.c_f64p)))
        stat[live], pval[live], l2fc[live] = s2, p2, l2
    return stat, pval, l2fc


def get_wilcox_test_pvalues(fg_mat, bg_mat, device=0):
    """Wilcoxon rank-sum p-value of every row of ``fg_mat`` against the same row of ``bg_mat``
    (diff_features.py:968-995: ``scipy.stats.ranksums(fg_mat[i], y=bg_mat[i]).pvalue``), as a list."""
    return ranksum_rows(fg_mat, bg_mat, device=device)[1].tolist()


def get_log2_fc(fg_mat, bg_mat, device=0):
    """``log2((mean(fg, axis=1) + 1e-12) / (mean(bg, axis=1) + 1e-12))`` (diff_features.py:1095-1120)."""
    return ranksum_rows(fg_mat, bg_mat, device=device)[2]


def p_adjust_bh(p):
    """Benjamini-Hochberg correction (diff_features.py:1031-1041)."""
    p = np.asarray(p, dtype=np.float64)
    by_descend = p.argsort()[::-1]
    steps = float(len(p)) / np.arange(len(p), 0, -1)
    q = np.minimum(1, np.minimum.accumulate(steps * p[by_descend]))
    out = np.empty_like(q)
    out[by_descend] = q   # == q[by_descend.argsort()]: the inverse of a permutation, without the second sort
    return out


def _adjust_all(pvalue_vectors):
    """Benjamini-Hochberg for every contrast (10 ms per 200 000 regions: not worth threads)."""
    return [p_adjust_bh(p) for p in pvalue_vectors]


def markers(input_mat, barcode_group, contrast_name, adjpval_thr=0.05, log2fc_thr=1, n_cpu=1, device=0):
    """Differential features of one contrast (diff_features.py:839-965): rank-sum p-value per feature between
    the foreground and the background cells, Benjamini-Hochberg adjustment, log2 fold change, both thresholds,
    sorted by ``Log2FC`` descending then ``Adjusted_pval``.  The matrix is handed to the GPU ONCE with the two
    column lists (the reference copies out ``fg_mat`` and ``bg_mat`` first); ``n_cpu`` is ignored."""
    log = _logger()
    if isinstance(input_mat, FactoredImputedFeatures):
        return _markers_factored(input_mat, [barcode_group], [contrast_name], adjpval_thr, log2fc_thr, log)[0]
    if isinstance(input_mat, pd.DataFrame):
        mat, features, samples = input_mat.values, input_mat.index.tolist(), input_mat.columns
    else:
        mat, features, samples = input_mat.mtx, input_mat.feature_names, input_mat.cell_names
    fg_cells_index = np.array(get_position_index(barcode_group[0], samples), dtype=np.int64)
    bg_cells_index = np.array(get_position_index(barcode_group[1], samples), dtype=np.int64)
    log.info(f"Subsetting data for {contrast_name} ({fg_cells_index.shape[0]} of {mat.shape[1]})")
    if scipy.sparse.issparse(mat):
        mat = mat.toarray()
    log.info(f"Computing p-value for {contrast_name}")
    _, wilcox_test_pvalues, log2_fc = ranksum_rows(mat, None, fg_cells_index, bg_cells_index, device=device)
    log.info(f"Computing log2FC for {contrast_name}")
    adj_pvalues = p_adjust_bh(wilcox_test_pvalues)
    markers_dataframe = _markers_frame(log2_fc, adj_pvalues, contrast_name, features, adjpval_thr, log2fc_thr)
    log.info(f"{contrast_name} done!")
    return markers_dataframe


def _markers_frame(log2_fc, adj_pvalues, contrast_name, features, adjpval_thr, log2fc_thr):
    """The DataFrame, thresholds and ordering of diff_features.py:941-963."""
    # the two .loc filters of the reference, applied before the frame exists (a frame of every tested region,
    # with an object column and a string index, costs more than the test itself on the GPU)
    log2_fc, adj_pvalues = np.asarray(log2_fc), np.asarray(adj_pvalues)
    with np.errstate(invalid="ignore"):
        keep = np.flatnonzero((adj_pvalues <= adjpval_thr) & (log2_fc >= log2fc_thr))
    contrast = pd.Series([contrast_name] * max(1, len(keep)))[: len(keep)]   # keeps the string dtype when empty
    if not isinstance(features, pd.Index):   # callers with several contrasts convert once and pass the Index
        features = pd.Index(features) if len(features) else pd.Index([], dtype=pd.Index([""]).dtype)
    markers_dataframe = pd.DataFrame(
        {"Log2FC": log2_fc[keep], "Adjusted_pval": adj_pvalues[keep], "Contrast": contrast.to_numpy()},
        index=features.take(keep),
    )
    markers_dataframe["Contrast"] = markers_dataframe["Contrast"].astype(contrast.dtype)
    return markers_dataframe.sort_values(["Log2FC", "Adjusted_pval"], ascending=[False, True])


def _markers_dense(imputed, barcode_groups, contrast_names, adjpval_thr, log2fc_thr, log, device=0):
    """markers() of every contrast of find_diff_features on a dense matrix with ONE upload of it (the reference, and
    a loop over :func:`markers`, hand the whole matrix to every contrast)."""
    mat, features, samples = imputed.mtx, imputed.feature_names, imputed.cell_names
    if scipy.sparse.issparse(mat):
        mat = mat.toarray()
    cols = [(get_position_index(g[0], samples), get_position_index(g[1], samples)) for g in barcode_groups]
    for name, (f, _) in zip(contrast_names, cols):
        log.info(f"Computing p-value and log2FC for {name} ({len(f)} of {len(samples)})")
    _, pval, l2fc = ranksum_contrasts(mat, cols, device=device)
    adjusted = _adjust_all([pval[k] for k in range(len(contrast_names))])
    features = pd.Index(features)
    out = []
    for k, name in enumerate(contrast_names):
        out.append(_markers_frame(l2fc[k], adjusted[k], name, features, adjpval_thr, log2fc_thr))
        log.info(f"{name} done!")
    return out


def _markers_factored(factored, barcode_groups, contrast_names, adjpval_thr, log2fc_thr, log):
    """markers() of every contrast from ONE pass over the recomputed imputation chunks."""
    samples = factored.cell_names
    cols = [(get_position_index(g[0], samples), get_position_index(g[1], samples)) for g in barcode_groups]
    for name, (f, _) in zip(contrast_names, cols):
        log.info(f"Computing p-value and log2FC for {name} ({len(f)} of {len(samples)})")
    _, pval, l2fc, keep = factored.ranksum(cols)
    features = factored.feature_names
    adjusted = _adjust_all([pval[k][keep] for k in range(len(contrast_names))])  # all-zero regions are not tested
    features = pd.Index(features)
    out = []
    for k, name in enumerate(contrast_names):
        out.append(_markers_frame(l2fc[k][keep], adjusted[k], name, features, adjpval_thr, log2fc_thr))
        log.info(f"{name} done!")
    return out


def find_diff_features(cistopic_obj, imputed_features_obj, variable, var_features=None, contrasts=None,
                       adjpval_thr=0.05, log2fc_thr=np.log2(1.5), split_pattern="___", n_cpu=1, device=0, **kwargs):
    """Differential imputed features per contrast (diff_features.py:716-836): same arguments and the same
    ``{contrast name: DataFrame}`` result; ``n_cpu`` and Ray keyword arguments are ignored."""
    selected_cells = list(set(cistopic_obj.cell_data.index.tolist()) & set(imputed_features_obj.cell_names))
    group_var = cistopic_obj.cell_data.loc[selected_cells, variable].dropna()
    if contrasts is None:
        levels = sorted(list(set(group_var.tolist())))
        contrasts = [[[x], levels[: levels.index(x)] + levels[levels.index(x) + 1:]] for x in levels]
        contrasts_names = levels
    else:
        contrasts_names = ["_".join(contrasts[i][0]) + "_VS_" + "_".join(contrasts[i][1]) for i in range(len(contrasts))]
    barcode_groups = [
        [group_var[group_var.isin(contrasts[i][0])].index.tolist(),
         group_var[group_var.isin(contrasts[i][1])].index.tolist()]
        for i in range(len(contrasts))
    ]
    subset_imputed_features_obj = imputed_features_obj.subset(cells=None, features=var_features, copy=True,
                                                              split_pattern=split_pattern)
    if isinstance(subset_imputed_features_obj, FactoredImputedFeatures):
        markers_list = _markers_factored(subset_imputed_features_obj, barcode_groups, contrasts_names, adjpval_thr,
                                         log2fc_thr, _logger())
        return {name: marker for name, marker in zip(contrasts_names, markers_list)}
    markers_list = _markers_dense(subset_imputed_features_obj, barcode_groups, contrasts_names, adjpval_thr, log2fc_thr,
                                  _logger(), device=device)
    return {name: marker for name, marker in zip(contrasts_names, markers_list)}
