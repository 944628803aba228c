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


class CistopicImputedFeatures:
    """cisTopic imputation data class (diff_features.py:26-59)."""

    def __init__(self, imputed_acc, feature_names, cell_names, project):
        self.mtx = imputed_acc
        self.feature_names = feature_names
        self.cell_names = cell_names
        self.project = project

    def __str__(self):
        return (f"CistopicImputedFeatures from project {self.project} with nCells × nFeatures = "
                f"{len(self.cell_names)} × {len(self.feature_names)}")


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
    region_names_to_keep = []
    output_chunk_end = 0
    scale = float(np.float32(scale_factor)) if as_int else 1.0
    plan = c_vp()
    check(L.cgs_impute_plan_create(ctypes.byref(plan), device, ptr(cell_topic, c_f32p), 0, K, C, None))
    try:
        for input_chunk_start in range(0, R, chunk_size):
            input_chunk_end = input_chunk_start + chunk_size
            output_chunk_start = output_chunk_end
            if log is not None:
                log.info(f"Impute region accessibility for regions {input_chunk_start}-{input_chunk_end}")
            a = topic_region[input_chunk_start:input_chunk_end]
            n = a.shape[0]
            chunk = np.empty((n, C), dtype=out_dtype)
            flags = np.empty(n, dtype=np.uint8)
            check(L.cgs_impute_plan_chunk(plan, ptr(a, c_f32p), n, scale, int(as_int), chunk.ctypes.data_as(c_vp),
                                          ptr(flags, c_u8p)))
            keep = np.flatnonzero(flags)
            names_chunk = region_names[input_chunk_start:input_chunk_end]
            region_names_to_keep.extend([names_chunk[i] for i in keep])
            output_chunk_end = output_chunk_start + len(keep)
            imputed_acc[output_chunk_start:output_chunk_end, :] = chunk[keep]
    finally:
        check(L.cgs_impute_plan_destroy(plan))
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
):
    """Impute region accessibility (diff_features.py:340-508)."""
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


def normalize_scores(imputed_acc, scale_factor=10 ** 4):
    """Log-normalise imputation data (diff_features.py:511-574); the arithmetic runs on the GPU
    (SURVEY.md section 8f, row N1)."""
    log = _logger()
    log.info("Normalizing imputed data")
    if isinstance(imputed_acc, CistopicImputedFeatures):
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
