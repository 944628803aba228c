"""ctypes binding of ``libcgs_b200.so`` (the C ABI declared in ``include/cgs_b200.h``).

There is no CPU fallback: if the library is missing, cannot be loaded, or no CUDA device is
visible, the calls raise.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

from . import _build

_lib = None

c_i32p = ctypes.POINTER(ctypes.c_int32)
c_i64p = ctypes.POINTER(ctypes.c_int64)
c_f64p = ctypes.POINTER(ctypes.c_double)
c_f32p = ctypes.POINTER(ctypes.c_float)
c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_vp = ctypes.c_void_p

# name -> (restype, argtypes); every symbol include/cgs_b200.h declares
SIGNATURES = {
    "cgs_last_error": (ctypes.c_char_p, []),
    "cgs_version": (ctypes.c_int, []),
    "cgs_device_count": (ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    "cgs_max_topics": (ctypes.c_int, []),
    "cgs_corpus_from_regions_by_cells": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int, c_i64p, c_i32p,
                                                        ctypes.c_int32, ctypes.c_int32, ctypes.c_int64,
                                                        ctypes.c_int32, ctypes.c_int32]),
    "cgs_corpus_from_cells_by_regions": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int, c_i64p, c_i32p,
                                                        ctypes.c_int32, ctypes.c_int32, ctypes.c_int64,
                                                        ctypes.c_int32, ctypes.c_int32]),
    "cgs_corpus_from_device_tokens": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int, c_vp, c_vp,
                                                     ctypes.c_int32, ctypes.c_int32, ctypes.c_int64, ctypes.c_int64]),
    "cgs_corpus_destroy": (ctypes.c_int, [c_vp]),
    "cgs_corpus_dims": (ctypes.c_int, [c_vp, c_i32p, c_i32p, c_i64p]),
    "cgs_corpus_export_tokens": (ctypes.c_int, [c_vp, c_i32p, c_i32p]),
    "cgs_corpus_set_token_offset": (ctypes.c_int, [c_vp, ctypes.c_int64]),
    "cgs_model_create": (ctypes.c_int, [ctypes.POINTER(c_vp), c_vp, ctypes.c_int32, ctypes.c_double,
                                        ctypes.c_double, ctypes.c_uint64, c_vp]),
    "cgs_model_destroy": (ctypes.c_int, [c_vp]),
    "cgs_model_set_stream": (ctypes.c_int, [c_vp, c_vp]),
    "cgs_model_sync": (ctypes.c_int, [c_vp]),
    "cgs_model_init": (ctypes.c_int, [c_vp]),
    "cgs_model_set_z": (ctypes.c_int, [c_vp, c_i32p]),
    "cgs_model_get_z": (ctypes.c_int, [c_vp, c_i32p]),
    "cgs_model_sweep": (ctypes.c_int, [c_vp, ctypes.c_int32]),
    "cgs_model_sweep_part": (ctypes.c_int, [c_vp, ctypes.c_int32, ctypes.c_int32]),
    "cgs_model_debug_sweep_visits": (ctypes.c_int, [c_vp, c_i64p]),
    "cgs_model_set_sweep_shape": (ctypes.c_int, [c_vp, ctypes.c_int32, ctypes.c_int32]),
    "cgs_model_set_rng_period": (ctypes.c_int, [c_vp, ctypes.c_int64]),
    "cgs_model_set_max_blocks": (ctypes.c_int, [c_vp, ctypes.c_int32]),
    "cgs_model_set_nk_staleness": (ctypes.c_int, [c_vp, ctypes.c_double]),
    "cgs_model_get_sweep_shape": (ctypes.c_int, [c_vp, c_i32p, c_i32p, c_i32p]),
    "cgs_model_launch_count": (ctypes.c_int, [c_vp, c_i64p]),
    "cgs_model_sweeps_done": (ctypes.c_int, [c_vp, c_i64p]),
    "cgs_model_loglik": (ctypes.c_int, [c_vp, ctypes.c_int, c_f64p]),
    "cgs_model_export_counts": (ctypes.c_int, [c_vp, c_i32p, c_i32p, c_i32p]),
    "cgs_model_export_distributions": (ctypes.c_int, [c_vp, c_f64p, c_f64p]),
    "cgs_model_export_frames": (ctypes.c_int, [c_vp, c_f64p, c_f64p]),
    "cgs_corpus_export_cell_sizes": (ctypes.c_int, [c_vp, c_i64p]),
    "cgs_model_topic_gram": (ctypes.c_int, [c_vp, c_f64p]),
    "cgs_model_cell_mixture": (ctypes.c_int, [c_vp, c_f64p, c_f64p]),
    "cgs_model_top_regions": (ctypes.c_int, [c_vp, ctypes.c_int32, c_i32p, c_i32p]),
    "cgs_model_codocument_frequencies": (ctypes.c_int, [c_vp, ctypes.c_int32, c_i32p, c_i32p]),
    "cgs_model_loglik_compat": (ctypes.c_int, [c_vp, ctypes.c_double, ctypes.c_double, c_f64p]),
    "cgs_model_device_buffer": (ctypes.c_int, [c_vp, ctypes.c_int, ctypes.POINTER(c_vp),
                                               ctypes.POINTER(ctypes.c_size_t)]),
    "cgs_model_padded_topics": (ctypes.c_int, [c_vp, c_i32p]),
    "cgs_model_delta_compute": (ctypes.c_int, [c_vp]),
    "cgs_model_delta_apply": (ctypes.c_int, [c_vp]),
    "cgs_model_snapshot": (ctypes.c_int, [c_vp]),
    "cgs_model_delta_apply_peers": (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp), ctypes.c_int32]),
    "cgs_corpus_set_global_size": (ctypes.c_int, [c_vp, ctypes.c_int64, ctypes.c_int64]),
    "cgs_model_set_deterministic": (ctypes.c_int, [c_vp, ctypes.c_int32]),
    "cgs_model_set_conditional_precision": (ctypes.c_int, [c_vp, ctypes.c_int32]),
    "cgs_model_region_histogram": (ctypes.c_int, [c_vp, c_vp, c_i64p]),
    "cgs_model_verify_counts": (ctypes.c_int, [c_vp, c_vp, c_i64p]),
    "cgs_model_state_hash": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_uint64)]),
    "cgs_model_allreduce_delta": (ctypes.c_int, [c_vp, c_vp]),
    "cgs_nccl_unique_id": (ctypes.c_int, [c_vp]),
    "cgs_nccl_comm_create": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int, ctypes.c_int32, ctypes.c_int32, c_vp]),
    "cgs_nccl_comm_destroy": (ctypes.c_int, [c_vp]),
    "cgs_model_exchange_arena": (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp), ctypes.POINTER(ctypes.c_size_t)]),
    "cgs_model_exchange_open": (ctypes.c_int, [c_vp, ctypes.POINTER(c_vp), ctypes.c_int32, ctypes.c_int32]),
    "cgs_model_exchange_round": (ctypes.c_int, [c_vp, ctypes.c_int32]),
    "cgs_model_exchange_wait": (ctypes.c_int, [c_vp]),
    "cgs_model_exchange_sync": (ctypes.c_int, [c_vp, ctypes.c_int32]),
    "cgs_corpus_set_region_split": (ctypes.c_int, [c_vp, ctypes.c_int32]),
    "cgs_corpus_set_region_blocks": (ctypes.c_int, [c_vp, ctypes.c_int32, c_i32p]),
    "cgs_model_sweep_blocked": (ctypes.c_int, [c_vp, ctypes.c_int32]),
    "cgs_model_sweep_overlapped": (ctypes.c_int, [c_vp, ctypes.c_int32, ctypes.c_int32]),
    "cgs_model_exchange_flush": (ctypes.c_int, [c_vp]),
    "cgs_model_exchange_phase_times": (ctypes.c_int, [c_vp, c_f64p]),
    "cgs_exchange_sync_local": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int32, ctypes.c_int32]),
    "cgs_exchange_round_local": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int32, ctypes.c_int32]),
    "cgs_model_exchange_status": (ctypes.c_int, [c_vp, c_i32p]),
    "cgs_ipc_export": (ctypes.c_int, [c_vp, c_vp]),
    "cgs_ipc_open": (ctypes.c_int, [ctypes.c_int, c_vp, ctypes.POINTER(c_vp)]),
    "cgs_ipc_close": (ctypes.c_int, [ctypes.c_int, c_vp]),
    "cgs_device_malloc": (ctypes.c_int, [ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(c_vp)]),
    "cgs_device_free": (ctypes.c_int, [ctypes.c_int, c_vp]),
    "cgs_impute_chunk": (ctypes.c_int, [ctypes.c_int, c_f32p, c_f32p, ctypes.c_int32, ctypes.c_int32,
                                        ctypes.c_int32, ctypes.c_float, ctypes.c_int, c_vp, c_u8p]),
    "cgs_impute_chunk_device": (ctypes.c_int, [ctypes.c_int, c_vp, c_vp, ctypes.c_int32, ctypes.c_int32,
                                               ctypes.c_int32, ctypes.c_float, ctypes.c_int, c_vp, c_vp, c_vp]),
    "cgs_impute_plan_create": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int, c_vp, ctypes.c_int,
                                              ctypes.c_int32, ctypes.c_int32, c_vp]),
    "cgs_impute_plan_destroy": (ctypes.c_int, [c_vp]),
    "cgs_impute_plan_chunk": (ctypes.c_int, [c_vp, c_f32p, ctypes.c_int32, ctypes.c_float, ctypes.c_int, c_vp, c_u8p]),
    "cgs_impute_plan_chunk_device": (ctypes.c_int, [c_vp, c_vp, ctypes.c_int32, ctypes.c_float, ctypes.c_int,
                                                    c_vp, c_vp, c_vp]),
    "cgs_impute_plan_launch_count": (ctypes.c_int, [c_vp, c_i64p]),
    "cgs_normalize_scores": (ctypes.c_int, [ctypes.c_int, c_vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                            ctypes.c_double, c_vp]),
    "cgs_ranksum_rows": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int64, c_vp, ctypes.c_int64, c_i32p,
                                        ctypes.c_int32, c_vp, ctypes.c_int64, c_i32p, ctypes.c_int32,
                                        c_f64p, c_f64p, c_f64p]),
    "cgs_ranksum_rows_device": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int64, c_vp, ctypes.c_int64, c_vp,
                                               ctypes.c_int32, c_vp, ctypes.c_int64, c_vp, ctypes.c_int32,
                                               c_vp, c_vp, c_vp, c_vp]),
    "cgs_ranksum_contrasts": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int64, c_vp, ctypes.c_int64, ctypes.c_int32,
                                             c_i32p, c_i64p, c_i32p, c_i64p, c_f64p, c_f64p, c_f64p]),
    "cgs_impute_normalized": (ctypes.c_int, [ctypes.c_int, c_f32p, c_f32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                             ctypes.c_float, ctypes.c_int, ctypes.c_double, ctypes.c_int32, c_vp, c_u8p,
                                             c_vp, c_i64p]),
    "cgs_impute_ranksum": (ctypes.c_int, [ctypes.c_int, c_f32p, c_f32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                          ctypes.c_float, ctypes.c_int, ctypes.c_int32, ctypes.c_int32, c_i32p, c_i64p,
                                          c_i32p, c_i64p, c_f64p, c_f64p, c_f64p, c_u8p]),
    "cgs_fragment_counts_create": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_int, ctypes.c_int32, c_i32p, c_i32p, c_i32p,
                                                  ctypes.c_int64, c_i32p, c_i32p, c_i32p, c_i32p, ctypes.c_int32]),
    "cgs_fragment_counts_destroy": (ctypes.c_int, [c_vp]),
    "cgs_fragment_counts_dims": (ctypes.c_int, [c_vp, c_i32p, c_i32p, c_i64p, c_i64p]),
    "cgs_fragment_counts_export": (ctypes.c_int, [c_vp, c_i64p, c_i32p, c_i32p]),
    "cgs_bed_read": (ctypes.c_int, [ctypes.POINTER(c_vp), ctypes.c_char_p, ctypes.c_int32, ctypes.c_int32]),
    "cgs_bed_destroy": (ctypes.c_int, [c_vp]),
    "cgs_bed_dims": (ctypes.c_int, [c_vp, c_i64p, c_i32p, c_i32p, c_i64p, c_i32p, c_i64p, c_i32p]),
    "cgs_bed_export": (ctypes.c_int, [c_vp, c_i32p, c_i32p, c_i32p, c_i32p, c_i32p, ctypes.c_char_p, c_i64p,
                                      ctypes.c_char_p, c_i64p]),
    "cgs_bed_collapse_duplicates": (ctypes.c_int, [c_vp, ctypes.c_int32, c_i64p]),
    "cgs_row_mean_var": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                        c_f32p, c_f32p]),
    "cgs_row_mean_var_device": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int64, ctypes.c_int64,
                                               ctypes.c_int64, c_vp, c_vp, c_vp]),
    "cgs_impute_row_stats": (ctypes.c_int, [ctypes.c_int, c_f32p, c_f32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                            ctypes.c_float, ctypes.c_int, ctypes.c_double, ctypes.c_int32, c_f32p, c_f32p,
                                            c_u8p]),
    "cgs_impute_rankings": (ctypes.c_int, [ctypes.c_int, c_f32p, c_f32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                           ctypes.c_float, ctypes.c_int, ctypes.c_uint64, ctypes.c_int32, ctypes.c_int64,
                                           c_vp, ctypes.c_int64]),
    "cgs_impute_rows": (ctypes.c_int, [ctypes.c_int, c_f32p, c_f32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                       ctypes.c_float, ctypes.c_int, ctypes.c_int32, c_u8p, c_vp, c_i64p]),
    "cgs_impute_column_totals": (ctypes.c_int, [ctypes.c_int, c_f32p, c_f32p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32,
                                                ctypes.c_float, ctypes.c_int, ctypes.c_int32, c_vp, c_u8p]),
    "cgs_impute_normalized_shard": (ctypes.c_int, [ctypes.c_int, c_f32p, c_f32p, ctypes.c_int32, ctypes.c_int32,
                                                   ctypes.c_int32, ctypes.c_float, ctypes.c_int, ctypes.c_double,
                                                   ctypes.c_int32, c_vp, c_vp, c_u8p, c_i64p]),
    "cgs_impute_row_stats_shard": (ctypes.c_int, [ctypes.c_int, c_f32p, c_f32p, ctypes.c_int32, ctypes.c_int32,
                                                  ctypes.c_int32, ctypes.c_float, ctypes.c_int, ctypes.c_double,
                                                  ctypes.c_int32, c_vp, c_f32p, c_f32p, c_u8p]),
    "cgs_rank_columns": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
                                        ctypes.c_uint64, ctypes.c_int64, c_i32p, ctypes.c_int64]),
    "cgs_rank_columns_device": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int64, ctypes.c_int64,
                                               ctypes.c_int64, ctypes.c_uint64, ctypes.c_int64, c_vp, ctypes.c_int64, c_vp]),
    "cgs_normalize_scores_device": (ctypes.c_int, [ctypes.c_int, c_vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int64,
                                                   ctypes.c_double, c_vp, c_vp, c_vp]),
}


class CgsError(RuntimeError):
    """Raised when a C-ABI call returns a non-zero code; carries cgs_last_error()."""


def _point_at_bundled_nccl():
    """cgs_nccl_* resolve NCCL at first use (csrc/cgs_b200.cu, nccl_api): the copy the process has already loaded, else
    ``$CGS_NCCL_LIBRARY``, else whatever the loader finds as libnccl.so.2.  In a Python process that has not imported torch
    yet the last one would be the SYSTEM library, and a later ``import torch`` would then be bound to it by soname instead
    of the (newer) one its wheel ships -- and fail.  So name the wheel's copy; nothing is loaded here."""
    if os.environ.get("CGS_NCCL_LIBRARY"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia")
        for loc in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(loc, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["CGS_NCCL_LIBRARY"] = cand
                return
    except Exception:  # no such wheel: the loader's default applies
        pass


def load() -> ctypes.CDLL:
    """Load (building first if the sources are newer) the CUDA library.  Fails loudly."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("CGS_B200_LIB") or _build.LIB_PATH  # CGS_B200_LIB: a prebuilt variant (tuning runs)
    if path == _build.LIB_PATH and _build.needs_build():
        if os.environ.get("CGS_B200_NO_BUILD") == "1" and os.path.exists(path):
            pass
        else:
            _build.build()
    _point_at_bundled_nccl()
    try:
        L = ctypes.CDLL(path)
    except OSError as e:  # pragma: no cover
        raise RuntimeError(f"cannot load {path}: {e}. The CUDA extension is required; there is no CPU fallback.")
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        msg = load().cgs_last_error()
        raise CgsError(msg.decode("utf-8", "replace") if msg else f"cgs error {rc}")


def require_device() -> int:
    """Number of visible CUDA devices; raises if there is none (no CPU path exists)."""
    n = ctypes.c_int(0)
    L = load()
    rc = L.cgs_device_count(ctypes.byref(n))
    if rc != 0 or n.value < 1:
        msg = L.cgs_last_error().decode("utf-8", "replace") if rc != 0 else "no CUDA device visible"
        raise RuntimeError("pycistopic_b200 needs an NVIDIA B200 (sm_100a) GPU: " + msg)
    return n.value


def ptr(a: np.ndarray, t):
    return a.ctypes.data_as(t)


class DeviceArrayView:
    """Zero-copy description of a handle-owned device buffer (``__cuda_array_interface__``),
    so a host framework (torch) can run collectives on it without copying."""

    def __init__(self, address: int, shape, typestr: str, owner=None):
        self._owner = owner
        self.__cuda_array_interface__ = {
            "shape": tuple(int(s) for s in shape),
            "typestr": typestr,
            "data": (int(address), False),
            "version": 3,
            "strides": None,
        }
