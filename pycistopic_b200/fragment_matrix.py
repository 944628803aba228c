"""Count matrix upstream of the topic-modelling path.

``create_cistopic_object_from_fragments`` (reference src/pycisTopic/cistopic_class.py:700-935) joins the
fragments with the regions through pyranges (``regions.join(fragments)``, :868), counts the pairs with
``groupby(["Name", "regionID"]).size().unstack()`` (:880-889, a DENSE regions x cells intermediate with a
partitioned fallback on MemoryError, :901-929) and hands the result to ``create_cistopic_object`` (:508-655),
which binarises it (:592).  Here the overlap join and the counting run on the GPU (``cgs_fragment_counts_*``,
``fragment_kernels.cuh``) and come back as a regions x cells CSR; the binarisation, the all-zero filters and
the per-cell / per-region statistics follow the reference line by line on that sparse matrix.

File parsing (pyranges / polars readers of the reference) is not rebuilt: fragments and regions are pandas
DataFrames with the BED column names pyranges uses (``Chromosome``, ``Start``, ``End``, and ``Name`` = barcode).
"""
from __future__ import annotations

import ctypes

import numpy as np
import pandas as pd
import scipy.sparse as sparse

from . import _lib
from ._lib import c_i32p, c_i64p, c_vp, check, ptr


def _codes_in(values: pd.Series, categories: pd.Index) -> np.ndarray:
    """int32 position of every value in ``categories`` (-1 when absent).  A categorical column (how fragment
    files are read) is recoded through its few categories instead of its millions of values."""
    if isinstance(values.dtype, pd.CategoricalDtype):
        lut = np.append(categories.get_indexer(values.cat.categories.astype(categories.dtype)), -1).astype(np.int32)
        return np.ascontiguousarray(lut[values.cat.codes.to_numpy()])  # code -1 (missing) reads the appended -1
    if categories.dtype == object or str(categories.dtype).startswith("str"):
        values = values.astype(str)
    return np.ascontiguousarray(categories.get_indexer(values).astype(np.int32))


def count_fragments_in_regions(fragments: pd.DataFrame, regions: pd.DataFrame, valid_bc=None, device=0, timings=None):
    """Fragments-in-regions counts as a sparse matrix.

    Returns ``(fragment_matrix, region_names, cell_names)``: CSR int32 regions x cells where entry (r, c) is the
    number of fragments of barcode c that overlap region r (half-open intervals, pyranges' join: a fragment
    counts once for every region it overlaps), region names ``chr:start-end`` in the order of ``regions`` and
    barcodes in the order of ``valid_bc`` (or sorted when it is None).  Like the reference's
    ``groupby().size().unstack()`` only regions and barcodes with at least one counted fragment are kept.
    """
    import time
    t0 = time.perf_counter()
    _lib.require_device()
    chrom = pd.Categorical(regions["Chromosome"].astype(str))
    r_chrom = chrom.codes.astype(np.int32)
    r_start = regions["Start"].to_numpy().astype(np.int32)
    r_end = regions["End"].to_numpy().astype(np.int32)
    order = np.lexsort((r_end, r_start, r_chrom))  # by chromosome, then start (ties by end: any order works)
    n_chrom = len(chrom.categories)
    chrom_ptr = np.zeros(n_chrom + 1, np.int32)
    np.cumsum(np.bincount(r_chrom, minlength=n_chrom), out=chrom_ptr[1:])
    s_start, s_end = np.ascontiguousarray(r_start[order]), np.ascontiguousarray(r_end[order])

    f_chrom = _codes_in(fragments["Chromosome"], chrom.categories)  # -1: chromosome without regions
    if valid_bc is None:
        name = fragments["Name"]
        if isinstance(name.dtype, pd.CategoricalDtype):  # used categories, sorted (what Categorical() would give)
            used = np.flatnonzero(np.bincount(name.cat.codes.to_numpy()[name.cat.codes.to_numpy() >= 0],
                                              minlength=len(name.cat.categories)))
            cell_categories = pd.Index(sorted(name.cat.categories[used]))
        else:
            cell_categories = pd.Index(pd.Categorical(name).categories)
    else:
        cell_categories = pd.Index(list(dict.fromkeys(valid_bc)))
    f_cell = _codes_in(fragments["Name"], cell_categories)  # -1: barcode not kept
    f_start = np.ascontiguousarray(fragments["Start"].to_numpy().astype(np.int32))
    f_end = np.ascontiguousarray(fragments["End"].to_numpy().astype(np.int32))
    n_cells = max(1, len(cell_categories))

    L = _lib.load()
    h = c_vp()
    t1 = time.perf_counter()
    check(L.cgs_fragment_counts_create(ctypes.byref(h), device, n_chrom, ptr(chrom_ptr, c_i32p), ptr(s_start, c_i32p),
                                       ptr(s_end, c_i32p), len(f_cell), ptr(f_chrom, c_i32p),
                                       ptr(f_start, c_i32p), ptr(f_end, c_i32p), ptr(f_cell, c_i32p), n_cells))
    try:
        nnz = ctypes.c_int64()
        check(L.cgs_fragment_counts_dims(h, None, None, ctypes.byref(nnz), None))
        R = len(order)
        indptr = np.empty(R + 1, np.int64)
        indices = np.empty(nnz.value, np.int32)
        data = np.empty(nnz.value, np.int32)
        check(L.cgs_fragment_counts_export(h, ptr(indptr, c_i64p), ptr(indices, c_i32p), ptr(data, c_i32p)))
    finally:
        check(L.cgs_fragment_counts_destroy(h))
    t2 = time.perf_counter()
    m = sparse.csr_matrix((data, indices, indptr), shape=(R, n_cells))
    inv = np.empty(R, np.int64)
    inv[order] = np.arange(R)
    m = m[inv]  # back to the caller's region order
    names = (regions["Chromosome"].astype(str) + ":" + regions["Start"].astype(str) + "-" + regions["End"].astype(str)).to_list()
    keep_r = np.flatnonzero(np.diff(m.indptr))
    keep_c = np.flatnonzero(np.bincount(m.indices, minlength=n_cells)[:n_cells])
    m = sparse.csr_matrix(m[keep_r][:, keep_c], dtype=np.int32)
    cats = list(cell_categories)
    if timings is not None:  # seconds: host recoding / the C-ABI calls (copies + kernels) / scipy reordering
        timings.update(host_prepare=t1 - t0, device=t2 - t1, host_finish=time.perf_counter() - t2)
    return m, [names[i] for i in keep_r], [cats[i] for i in keep_c]


class CistopicObject:
    """The attributes of the reference's ``CistopicObject`` (cistopic_class.py:31-51) plus ``add_LDA_model``
    (:491-504): what ``run_cgs_models`` / ``impute_accessibility`` / ``find_diff_features`` read."""

    def __init__(self, fragment_matrix, binary_matrix, cell_names, region_names, cell_data, region_data,
                 path_to_fragments, project="cisTopic"):
        self.fragment_matrix = fragment_matrix
        self.binary_matrix = binary_matrix
        self.cell_names = cell_names
        self.region_names = region_names
        self.cell_data = cell_data
        self.region_data = region_data
        self.project = project
        if isinstance(path_to_fragments, str):
            path_to_fragments = {project: path_to_fragments}
        self.path_to_fragments = path_to_fragments
        self.selected_model = None
        self.projections = {"cell": {}, "region": {}}

    def __str__(self):
        return f"CistopicObject from project {self.project} with n_cells × n_regions = {len(self.cell_names)} × {len(self.region_names)}"

    def add_LDA_model(self, model):
        model.region_topic = model.topic_region  # the reference's spelling at cistopic_class.py:502
        model.cell_topic = model.cell_topic.loc[:, self.cell_names]
        model.topic_region = model.topic_region.loc[self.region_names]
        self.selected_model = model


def region_names_to_coordinates(region_names):
    """utils.py:60-102: ``chrom:start-end`` -> DataFrame(Chromosome, Start, End) indexed by the names."""
    parts = [n.replace(":", "-", 1).rsplit("-", 2) for n in region_names]
    df = pd.DataFrame(parts, columns=["Chromosome", "Start", "End"], index=list(region_names))
    df["Start"] = df["Start"].astype(np.int32)
    df["End"] = df["End"].astype(np.int32)
    return df


def create_cistopic_object(fragment_matrix, cell_names=None, region_names=None, min_frag=1, min_cell=1, is_acc=1,
                           path_to_fragments=None, project="cisTopic", tag_cells=True, split_pattern="___"):
    """``create_cistopic_object`` (cistopic_class.py:508-655) without the blacklist step: binarise at ``is_acc``,
    drop all-zero regions, per-cell and per-region statistics, ``min_frag`` / ``min_cell`` filters."""
    if isinstance(fragment_matrix, pd.DataFrame):
        region_names = list(fragment_matrix.index)
        cell_names = list(fragment_matrix.columns.values)
        fragment_matrix = sparse.csr_matrix(fragment_matrix.to_numpy(), dtype=np.int32)
    fragment_matrix = sparse.csr_matrix(fragment_matrix, dtype=np.int32)
    cell_names, region_names = list(cell_names), list(region_names)
    if tag_cells:
        cell_names = [c + split_pattern + project for c in cell_names]
    # sklearn.preprocessing.binarize(X, threshold=is_acc - 1): 1 where the count exceeds the threshold
    binary_matrix = fragment_matrix.copy()
    binary_matrix.data = (binary_matrix.data > is_acc - 1).astype(np.int32)
    binary_matrix.eliminate_zeros()
    selected_regions = np.nonzero(binary_matrix.getnnz(axis=1))[0]
    fragment_matrix = fragment_matrix[selected_regions,]
    binary_matrix = binary_matrix[selected_regions,]
    region_names = [region_names[i] for i in selected_regions]

    nr_frag = np.array(fragment_matrix.sum(axis=0)).flatten()
    nr_acc = np.array(binary_matrix.sum(axis=0)).flatten()
    with np.errstate(divide="ignore"):
        cell_data = pd.DataFrame(
            [nr_frag, np.log10(nr_frag), nr_acc, np.log10(nr_acc), [project] * len(cell_names)],
            columns=cell_names,
            index=["cisTopic_nr_frag", "cisTopic_log_nr_frag", "cisTopic_nr_acc", "cisTopic_log_nr_acc", "sample_id"],
        ).transpose()
    if min_frag != 1:
        selected_cells = (cell_data.cisTopic_nr_frag >= min_frag).to_numpy()
        fragment_matrix = fragment_matrix[:, selected_cells]
        binary_matrix = binary_matrix[:, selected_cells]
        cell_data = cell_data.loc[selected_cells,]
        cell_names = cell_data.index.to_list()
    region_data = region_names_to_coordinates(region_names)
    region_data["Width"] = abs(region_data.End - region_data.Start).astype(np.int32)
    region_data["cisTopic_nr_frag"] = np.array(fragment_matrix.sum(axis=1)).flatten()
    with np.errstate(divide="ignore"):
        region_data["cisTopic_log_nr_frag"] = np.log10(region_data["cisTopic_nr_frag"])
        region_data["cisTopic_nr_acc"] = np.array(binary_matrix.sum(axis=1)).flatten()
        region_data["cisTopic_log_nr_acc"] = np.log10(region_data["cisTopic_nr_acc"])
    if min_cell != 1:
        selected_regions = (region_data.cisTopic_nr_acc >= min_cell).to_numpy()
        fragment_matrix = fragment_matrix[selected_regions, :]
        binary_matrix = binary_matrix[selected_regions, :]
        region_data = region_data.loc[selected_regions, :]  # (the reference indexes with [sel, :], a pandas error)
        region_names = region_data.index.to_list()
    return CistopicObject(sparse.csr_matrix(fragment_matrix), sparse.csr_matrix(binary_matrix), cell_names, region_names,
                          cell_data, region_data, path_to_fragments or {}, project)


def create_cistopic_object_from_fragments(fragments, regions, valid_bc=None, min_frag=1, min_cell=1, is_acc=1,
                                          project="cisTopic", split_pattern="___", device=0, **_ignored):
    """The counting half of the reference function (cistopic_class.py:700-935) on the GPU.  ``fragments`` and
    ``regions`` are DataFrames (``Chromosome``, ``Start``, ``End``; fragments also ``Name``) instead of file
    paths; ``valid_bc`` selects barcodes (:855-857); ``n_cpu``, ``partition`` and blacklist / metrics arguments
    of the reference are accepted and ignored."""
    m, region_names, cell_names = count_fragments_in_regions(fragments, regions, valid_bc, device)
    return create_cistopic_object(m, cell_names, region_names, min_frag=min_frag, min_cell=min_cell, is_acc=is_acc,
                                  project=project, split_pattern=split_pattern)
