"""Count matrix upstream of the topic-modelling path.

``create_cistopic_object_from_fragments`` (reference src/pycisTopic/cistopic_class.py:700-935) joins the
fragments with the regions through pyranges (``regions.join(fragments)``, :868), counts the pairs with
``groupby(["Name", "regionID"]).size().unstack()`` (:880-889, a DENSE regions x cells intermediate with a
partitioned fallback on MemoryError, :901-929) and hands the result to ``create_cistopic_object`` (:508-655),
which binarises it (:592).  Here the overlap join and the counting run on the GPU (``cgs_fragment_counts_*``,
``fragment_kernels.cuh``) and come back as a regions x cells CSR; the binarisation, the all-zero filters and
the per-cell / per-region statistics follow the reference line by line on that sparse matrix.

File parsing: ``read_fragments_from_file`` / ``read_bed`` stand where the reference calls
``read_fragments_to_pyranges`` (fragments.py:21-165, at cistopic_class.py:814-817) and ``pyranges.read_bed`` (:835, :578):
a native multi-threaded reader (``cgs_bed_*``, csrc/bed_reader.cpp; plain, gzip or bgzip text) that hands back the
columns dictionary-coded, so the strings of a 10^8-line fragments file never become Python objects.  Fragments and
regions may also be given as pandas DataFrames with the BED column names pyranges uses (``Chromosome``, ``Start``,
``End``, and ``Name`` = barcode).
"""
from __future__ import annotations

import ctypes
import os
import re

import numpy as np
import pandas as pd
import scipy.sparse as sparse

from . import _lib
from ._lib import c_i32p, c_i64p, c_vp, check, ptr


def _log():
    import logging
    return logging.getLogger("cisTopic")


def _codes_in(values: pd.Series, categories: pd.Index) -> np.ndarray:
    """int32 position of every value in ``categories`` (-1 when absent).  A categorical column (how fragment
    files are read) is recoded through its few categories instead of its millions of values."""
    if isinstance(values.dtype, pd.CategoricalDtype):
        lut = np.append(categories.get_indexer(values.cat.categories.astype(categories.dtype)), -1).astype(np.int32)
        return np.ascontiguousarray(lut[values.cat.codes.to_numpy()])  # code -1 (missing) reads the appended -1
    if categories.dtype == object or str(categories.dtype).startswith("str"):
        values = values.astype(str)
    return np.ascontiguousarray(categories.get_indexer(values).astype(np.int32))


def _strings(chars: bytes, offsets: np.ndarray) -> list[str]:
    return [chars[offsets[i]:offsets[i + 1]].decode("utf-8", "replace") for i in range(len(offsets) - 1)]


def read_bed_table(path, min_columns=3, n_threads=0, collapse_duplicates=False) -> dict:
    """One BED-like text file (plain / gzip / bgzip) through ``cgs_bed_read``.  Returns a dict with ``n_columns`` (of the
    file), int32 arrays ``chromosome`` (codes), ``start``, ``end``, ``name`` (codes, 4+ columns) and ``score`` (5+ columns
    of integers, or after ``collapse_duplicates``), the SORTED dictionaries ``chromosomes`` / ``names`` and
    ``score_is_integer``.  ``collapse_duplicates``: utils.py:284-293 on a file WITHOUT a Score column (the condition of
    cistopic_class.py:819); ``duplicates_removed`` tells how many records went."""
    path = os.path.expanduser(os.fspath(path))  # utils.normalise_filepath
    L = _lib.load()
    h = c_vp()
    check(L.cgs_bed_read(ctypes.byref(h), path.encode(), int(min_columns), int(n_threads)))
    try:
        n_rows, n_cols, n_chrom, n_names, state = (ctypes.c_int64(), ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32(),
                                                   ctypes.c_int32())
        chrom_chars, name_chars = ctypes.c_int64(), ctypes.c_int64()
        removed = ctypes.c_int64(0)
        check(L.cgs_bed_dims(h, None, ctypes.byref(n_cols), None, None, None, None, ctypes.byref(state)))
        collapsed = bool(collapse_duplicates) and n_cols.value == 4
        if collapsed:
            check(L.cgs_bed_collapse_duplicates(h, int(n_threads), ctypes.byref(removed)))
        check(L.cgs_bed_dims(h, ctypes.byref(n_rows), ctypes.byref(n_cols), ctypes.byref(n_chrom), ctypes.byref(chrom_chars),
                             ctypes.byref(n_names), ctypes.byref(name_chars), ctypes.byref(state)))
        n = n_rows.value
        out = {"n_columns": n_cols.value, "duplicates_removed": removed.value if collapsed else None,
               "chromosome": np.empty(n, np.int32), "start": np.empty(n, np.int32), "end": np.empty(n, np.int32)}
        has_name, has_score = n_cols.value >= 4, state.value != 0
        if has_name:
            out["name"] = np.empty(n, np.int32)
        if has_score:
            out["score"] = np.empty(n, np.int32)
        out["score_is_integer"] = state.value == 1
        cbuf = ctypes.create_string_buffer(max(1, chrom_chars.value))
        coff = np.empty(n_chrom.value + 1, np.int64)
        nbuf = ctypes.create_string_buffer(max(1, name_chars.value))
        noff = np.empty(n_names.value + 1, np.int64)
        check(L.cgs_bed_export(h, ptr(out["chromosome"], c_i32p), ptr(out["start"], c_i32p), ptr(out["end"], c_i32p),
                               ptr(out["name"], c_i32p) if has_name else None,
                               ptr(out["score"], c_i32p) if has_score else None,
                               cbuf, ptr(coff, c_i64p), nbuf if has_name else None, ptr(noff, c_i64p) if has_name else None))
        out["chromosomes"] = _strings(cbuf.raw, coff)
        out["names"] = _strings(nbuf.raw, noff) if has_name else []
        return out
    finally:
        check(L.cgs_bed_destroy(h))


def _table_to_frame(t: dict) -> pd.DataFrame:
    df = pd.DataFrame({
        "Chromosome": pd.Categorical.from_codes(t["chromosome"], categories=pd.Index(t["chromosomes"], dtype=object)),
        "Start": t["start"], "End": t["end"]})
    if "name" in t:
        df["Name"] = pd.Categorical.from_codes(t["name"], categories=pd.Index(t["names"], dtype=object))
    if "score" in t:
        df["Score"] = t["score"]
    return df


def read_fragments_from_file(path_to_fragments, n_threads=0, collapse_duplicates=False) -> pd.DataFrame:
    """The DataFrame behind ``read_fragments_to_pyranges(path).df`` (fragments.py:21-165) for the columns the count
    matrix reads: ``Chromosome`` and ``Name`` categorical, ``Start`` / ``End`` int32, ``Score`` int32 when the file has
    a fifth column.  Fewer than four columns raise the reference's ValueError."""
    try:
        t = read_bed_table(path_to_fragments, 4, n_threads, collapse_duplicates)
    except _lib.CgsError as e:
        if "needs to have at least" in str(e):
            raise ValueError(re.sub(r" \[[^\]]*\]$", "", str(e))) from None
        raise
    return _table_to_frame(t)


def read_bed(path_to_regions, n_threads=0) -> pd.DataFrame:
    """``pyranges.read_bed(path)[["Chromosome", "Start", "End"]]`` as a DataFrame, records in file order
    (cistopic_class.py:835-836, :578)."""
    return _table_to_frame(read_bed_table(path_to_regions, 3, n_threads))[["Chromosome", "Start", "End"]]


def regions_outside_blacklist(regions: pd.DataFrame, blacklist: pd.DataFrame) -> np.ndarray:
    """Boolean mask of ``regions.overlap(blacklist, invert=True)`` (cistopic_class.py:575-579): a region goes when it
    shares at least one base with a blacklisted interval of its chromosome (half-open intervals)."""
    return ~_overlaps_any(regions["Chromosome"].astype(str).to_numpy(), regions["Start"].to_numpy(),
                          regions["End"].to_numpy(), blacklist)


def _overlaps_any(chrom, start, end, regions: pd.DataFrame) -> np.ndarray:
    """For every interval (chrom names, starts, ends): does it share a base with at least one of ``regions``?"""
    hit = np.zeros(len(start), bool)
    chrom = np.asarray(chrom)
    r_chrom = regions["Chromosome"].astype(str).to_numpy()
    for name in pd.unique(r_chrom):
        rows = np.flatnonzero(chrom == name)
        if len(rows) == 0:
            continue
        sel = r_chrom == name
        rs = regions["Start"].to_numpy()[sel].astype(np.int64)
        re_ = regions["End"].to_numpy()[sel].astype(np.int64)
        order = np.argsort(rs, kind="stable")
        rs, far = rs[order], np.maximum.accumulate(re_[order])  # furthest end among the regions starting at or before
        j = np.searchsorted(rs, np.asarray(end)[rows], side="left")  # regions that start before the interval ends
        hit[rows] = (j > 0) & (far[np.maximum(j, 1) - 1] > np.asarray(start)[rows])
    return hit


def get_fragments_in_peaks(fragments: pd.DataFrame, regions: pd.DataFrame) -> pd.DataFrame:
    """``get_fragments_in_peaks`` (fragments.py:945-1021) on pandas frames: per barcode, over the fragments that overlap
    at least one region (each such fragment once, ``how="first"``), ``total_fragments_in_peaks_count`` = sum of
    ``Score`` and ``unique_fragments_in_peaks_count`` = their number; barcodes in the order they first appear among
    those fragments (``group_by(maintain_order=True)``).  Host NumPy (a sorted-interval search per chromosome)."""
    inside = _overlaps_any(fragments["Chromosome"].astype(str).to_numpy(), fragments["Start"].to_numpy(),
                           fragments["End"].to_numpy(), regions)
    sub = fragments.loc[inside, ["Name", "Score"]]
    codes, barcodes = pd.factorize(sub["Name"].astype(str), sort=False)
    return pd.DataFrame({"CB": barcodes,
                         "total_fragments_in_peaks_count": np.bincount(codes, weights=sub["Score"].to_numpy(),
                                                                       minlength=len(barcodes)).astype(np.int64),
                         "unique_fragments_in_peaks_count": np.bincount(codes, minlength=len(barcodes)).astype(np.int64)})


def _device_counts(device, n_chrom, chrom_ptr, s_start, s_end, f_chrom, f_start, f_end, f_cell, n_cells):
    """The device step: CSR (indptr int64, indices int32, counts int32) of the sorted regions x cells counts."""
    _lib.require_device()
    L = _lib.load()
    h = c_vp()
    check(L.cgs_fragment_counts_create(ctypes.byref(h), device, n_chrom, ptr(chrom_ptr, c_i32p), ptr(s_start, c_i32p),
                                       ptr(s_end, c_i32p), len(f_cell), ptr(f_chrom, c_i32p),
                                       ptr(f_start, c_i32p), ptr(f_end, c_i32p), ptr(f_cell, c_i32p), n_cells))
    try:
        nnz = ctypes.c_int64()
        check(L.cgs_fragment_counts_dims(h, None, None, ctypes.byref(nnz), None))
        indptr = np.empty(len(s_start) + 1, np.int64)
        indices = np.empty(nnz.value, np.int32)
        data = np.empty(nnz.value, np.int32)
        check(L.cgs_fragment_counts_export(h, ptr(indptr, c_i64p), ptr(indices, c_i32p), ptr(data, c_i32p)))
    finally:
        check(L.cgs_fragment_counts_destroy(h))
    return indptr, indices, data


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

    t1 = time.perf_counter()
    R = len(order)
    indptr, indices, data = _device_counts(device, n_chrom, chrom_ptr, s_start, s_end, f_chrom, f_start, f_end, f_cell, n_cells)
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

    def add_cell_data(self, cell_data: pd.DataFrame, split_pattern: str = "___"):
        """cistopic_class.py:88-142: join metadata indexed by cell name -- or by the bare barcode, when the object's
        names carry the ``___project`` tag and the metadata's do not -- onto ``cell_data``; columns that exist are
        overwritten, cells absent from the metadata get NaN."""
        obj_cell_data = self.cell_data.copy()
        obj_cell_names = list(self.cell_names)
        by_barcode = False
        present = set(obj_cell_names) & set(cell_data.index)
        if len(present) < len(obj_cell_names):
            bare = prepare_tag_cells(obj_cell_names, split_pattern)
            if len(set(bare) & set(cell_data.index)) < len(present):
                print("Warning: Some cells in this CistopicObject are not present in this cell_data. Values will be "
                      "filled with Nan\n")
            else:
                by_barcode = True
        common = [c for c in obj_cell_data.columns if c in set(cell_data.columns)]
        if common:
            print(f"Columns {common} will be overwritten")
            obj_cell_data = obj_cell_data.loc[:, [c for c in obj_cell_data.columns if c not in set(common)]]
        if by_barcode:
            bare = prepare_tag_cells(obj_cell_names, split_pattern)
            obj_cell_data.index = bare
            cell_data = cell_data.loc[[b for b in dict.fromkeys(bare) if b in set(cell_data.index)],]
            new_cell_data = pd.concat([obj_cell_data, cell_data], axis=1, sort=False).loc[bare]
            new_cell_data.index = obj_cell_names
        else:
            cell_data = cell_data.loc[[c for c in obj_cell_names if c in present],]
            new_cell_data = pd.concat([obj_cell_data, cell_data], axis=1, sort=False)
        self.cell_data = new_cell_data.loc[obj_cell_names]

    def add_LDA_model(self, model):
        model.region_topic = model.topic_region  # the reference's spelling at cistopic_class.py:502
        model.cell_topic = model.cell_topic.loc[:, self.cell_names]
        model.topic_region = model.topic_region.loc[self.region_names]
        self.selected_model = model


def prepare_tag_cells(cell_names, split_pattern="___"):
    """utils.py:205-223: the barcode in front of the sample tag."""
    if split_pattern == "-":
        out = []
        for x in cell_names:
            m = re.findall(r"^[ACGT]*-[0-9]+-", x)
            y = m[0].rstrip("-") if m else x
            m2 = re.findall(r"^\w*-[0-9]*", y)
            out.append(m2[0].rstrip("-") if m2 and y == x else y)
        return out
    return [x.split(split_pattern)[0] for x in cell_names]


def region_names_to_coordinates(region_names):
    """utils.py:60-102: ``chrom:start-end`` -> DataFrame(Chromosome, Start, End) indexed by the names."""
    parts = [n.replace(":", "-", 1).rsplit("-", 2) for n in region_names]
    df = pd.DataFrame(parts, columns=["Chromosome", "Start", "End"], index=list(region_names))
    df["Start"] = df["Start"].astype(np.int32)
    df["End"] = df["End"].astype(np.int32)
    return df


def create_cistopic_object(fragment_matrix, cell_names=None, region_names=None, path_to_blacklist=None, min_frag=1,
                           min_cell=1, is_acc=1, path_to_fragments=None, project="cisTopic", tag_cells=True,
                           split_pattern="___"):
    """``create_cistopic_object`` (cistopic_class.py:508-655): blacklisted regions out (``path_to_blacklist``: a BED
    file, or a DataFrame of intervals), binarise at ``is_acc``, drop all-zero regions, per-cell and per-region statistics,
    ``min_frag`` / ``min_cell`` filters.  The regions keep their input order (the reference's blacklist step returns them
    grouped by chromosome, a pyranges habit)."""
    if isinstance(fragment_matrix, pd.DataFrame):
        region_names = list(fragment_matrix.index)
        cell_names = list(fragment_matrix.columns.values)
        fragment_matrix = sparse.csr_matrix(fragment_matrix.to_numpy(), dtype=np.int32)
    fragment_matrix = sparse.csr_matrix(fragment_matrix, dtype=np.int32)
    cell_names, region_names = list(cell_names), list(region_names)
    if tag_cells:
        cell_names = [c + split_pattern + project for c in cell_names]
    if path_to_blacklist is not None:
        blacklist = path_to_blacklist if isinstance(path_to_blacklist, pd.DataFrame) else read_bed(path_to_blacklist)
        keep = np.flatnonzero(regions_outside_blacklist(region_names_to_coordinates(region_names), blacklist))
        fragment_matrix = fragment_matrix[keep,]
        region_names = [region_names[i] for i in keep]
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


def create_cistopic_object_from_matrix_file(fragment_matrix_file, path_to_blacklist=None, compression=None, min_frag=1,
                                            min_cell=1, is_acc=1, path_to_fragments=None, sample_id=None,
                                            project="cisTopic", split_pattern="___"):
    """``create_cistopic_object_from_matrix_file`` (cistopic_class.py:656-735): a tab-separated count matrix read as the
    reference reads it (``pd.read_csv(sep="\\t", header=0)``): a header of cell names and records that start with the
    region name, one field more than the header, which pandas then takes as the row names."""
    if compression is not None:
        fragment_matrix = pd.read_csv(fragment_matrix_file, sep="\t", header=0, compression=compression)
    else:
        fragment_matrix = pd.read_csv(fragment_matrix_file, sep="\t", header=0)
    obj = create_cistopic_object(fragment_matrix, path_to_blacklist=path_to_blacklist, min_frag=min_frag, min_cell=min_cell,
                                 is_acc=is_acc, path_to_fragments=path_to_fragments, project=project,
                                 split_pattern=split_pattern)
    if sample_id is not None:
        if isinstance(path_to_fragments, dict):
            obj.add_cell_data(sample_id, split_pattern)
        else:
            _log().error("Provide path_to_fragments with keys matching levels in sample_id!")
    return obj


def _collapse_duplicates_frame(fragments: pd.DataFrame) -> pd.DataFrame:
    """utils.py:284-293 on a DataFrame: one record per (Chromosome, Start, End, Name), ``Score`` = how many there were."""
    keys = ["Chromosome", "Start", "End", "Name"]
    out = fragments.groupby(keys, sort=False, observed=True).size().reset_index(name="Score")
    out["Score"] = out["Score"].astype(np.int32)
    return out


def create_cistopic_object_from_fragments(path_to_fragments, path_to_regions, path_to_blacklist=None, metrics=None,
                                          valid_bc=None, n_cpu=1, min_frag=1, min_cell=1, is_acc=1,
                                          check_for_duplicates=True, project="cisTopic", partition=5, fragments_df=None,
                                          split_pattern="___", use_polars=True, device=0):
    """``create_cistopic_object_from_fragments`` (cistopic_class.py:738-935), same arguments in the same order.

    ``path_to_fragments`` / ``path_to_regions`` are BED files (plain, gzip, bgzip) read by the native reader, or -- as
    ``fragments_df`` of the reference -- DataFrames with ``Chromosome``, ``Start``, ``End`` (and ``Name``).  Without a
    ``Score`` column duplicated fragments are collapsed when ``check_for_duplicates`` (:819-831; in the reader for a
    file).  ``metrics`` (CellRanger ``singlecell.csv`` path or DataFrame, :844-853) and ``valid_bc`` (:854-856) select
    barcodes; without ``metrics`` the number of fragments of every barcode becomes ``Unique_nr_frag`` (:857-866, :934).
    The overlap join and the counting (:868-889) run on the GPU; ``n_cpu`` sets the reader's threads; ``partition`` and
    ``use_polars`` (the reference's fallback for its dense intermediate, :896-929, and its reader switch) are accepted
    and not needed."""
    source = path_to_fragments if isinstance(path_to_fragments, (str, os.PathLike)) else None
    if isinstance(path_to_fragments, pd.DataFrame) and fragments_df is None:
        fragments_df = path_to_fragments
    if hasattr(fragments_df, "df") and not isinstance(fragments_df, pd.DataFrame):
        fragments_df = fragments_df.df  # a PyRanges
    if isinstance(fragments_df, pd.DataFrame):
        fragments = fragments_df
        if "Score" not in fragments.columns:
            fragments = _collapse_duplicates_frame(fragments) if check_for_duplicates else fragments.assign(Score=1)
    else:
        if source is None:
            raise ValueError("path_to_fragments must be a file name when fragments_df is not given")
        threads = 0 if n_cpu is None or n_cpu <= 1 else int(n_cpu)  # the reference's default n_cpu = 1 is about pyranges
        fragments = read_fragments_from_file(source, threads, collapse_duplicates=check_for_duplicates)
        if "Score" not in fragments.columns:
            fragments["Score"] = np.int32(1)
    regions = path_to_regions if isinstance(path_to_regions, pd.DataFrame) else read_bed(path_to_regions)
    regions = regions[["Chromosome", "Start", "End"]]

    if metrics is not None:
        if isinstance(metrics, (str, os.PathLike)):
            metrics = pd.read_csv(metrics)
        if "is__cell_barcode" in metrics.columns:
            metrics = metrics[metrics.is__cell_barcode == 1]
            metrics.index = metrics.barcode
            metrics = metrics.iloc[:, 2:]
        fragments = fragments[fragments.Name.isin(set(metrics.index))]
    if isinstance(valid_bc, list):
        fragments = fragments[fragments.Name.isin(set(valid_bc))]
    if metrics is None:
        name = fragments["Name"]
        if not isinstance(name.dtype, pd.CategoricalDtype):
            name = name.astype("category")
        per_barcode = np.bincount(name.cat.codes.to_numpy()[name.cat.codes.to_numpy() >= 0],
                                  minlength=len(name.cat.categories))
        used = np.flatnonzero(per_barcode)
        barcodes = [str(b) for b in name.cat.categories[used]]
        FPB_DF = pd.DataFrame({"Unique_nr_frag": per_barcode[used]}, index=barcodes)

    # the filters above already removed every other barcode: the cells come out sorted, as without valid_bc
    m, region_names, cell_names = count_fragments_in_regions(fragments, regions, None, device)
    obj = create_cistopic_object(m, cell_names, region_names, path_to_blacklist=path_to_blacklist, min_frag=min_frag,
                                 min_cell=min_cell, is_acc=is_acc, path_to_fragments={project: source}, project=project,
                                 split_pattern=split_pattern)
    if metrics is not None:
        metrics = metrics.copy()
        metrics["barcode"] = metrics.index.tolist()
        obj.add_cell_data(metrics, split_pattern)
    else:
        FPB_DF["barcode"] = FPB_DF.index.tolist()
        obj.add_cell_data(FPB_DF, split_pattern)
    return obj
