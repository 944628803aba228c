"""CPU tests of the file front end of the count matrix: the native BED / fragments reader
(``cgs_bed_*``, csrc/bed_reader.cpp -- host code, no device) against the reference's own reader run live
(fragments.py:21-165, its pandas and pyarrow engines) and against pandas directly; ``collapse_duplicates``,
``create_cistopic_object`` and ``add_cell_data`` against the reference's functions run live; the blacklist filter
against a brute-force overlap test; the whole ``create_cistopic_object_from_fragments`` host logic with the device
counting replaced by the CPU oracle."""
import gzip
import os
import zlib

import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

from oracle import fragment_oracle as fo
from oracle import ref_functions as rf
from pycistopic_b200 import fragment_matrix as fm

needs_reference = pytest.mark.skipif(not rf.available(), reason="/root/reference is not here")


def _fragments(seed, n=5000, n_cells=40, score=True, duplicates=False, n_chrom=5):
    rng = np.random.default_rng(seed)
    chroms = [f"chr{c}" for c in list(range(1, n_chrom)) + ["X"]]
    start = rng.integers(0, 200000, n)
    df = pd.DataFrame({"Chromosome": [chroms[i] for i in rng.integers(0, len(chroms), n)], "Start": start,
                       "End": start + rng.integers(1, 800, n),
                       "Name": [f"{''.join(rng.choice(list('ACGT'), 8))}-1" for _ in range(n_cells)] * (n // n_cells)})
    if duplicates:
        df = pd.concat([df, df.iloc[rng.integers(0, n, n // 3)]], ignore_index=True)
        df = df.iloc[rng.permutation(len(df))].reset_index(drop=True)
    if score:
        df["Score"] = rng.integers(1, 5, len(df))
    return df


def _bgzf(raw: bytes, block=65280, eof=True) -> bytes:
    """BGZF (bgzip) as CellRanger / tabix write it: gzip members of at most 64 KiB whose extra field 'BC' names the
    member's size, closed by an empty member."""
    import struct
    out = []
    for i in range(0, len(raw), block):
        part = raw[i:i + block]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = c.compress(part) + c.flush()
        out.append(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, len(body) + 25)
                   + body + struct.pack("<II", zlib.crc32(part) & 0xFFFFFFFF, len(part)))
    if eof:
        out.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    return b"".join(out)


def _write(df, path, header="", newline="\n", final_newline=True, compress=None):
    text = header + newline.join("\t".join(str(v) for v in row) for row in df.itertuples(index=False))
    if final_newline:
        text += newline
    raw = text.encode()
    if compress == "gzip":
        raw = gzip.compress(raw)
    elif compress == "members":  # a series of plain gzip members (cat a.gz b.gz): zlib's sequential reader
        parts = [raw[i:i + 7001] for i in range(0, len(raw), 7001)]
        raw = b"".join(gzip.compress(p) for p in parts)
    elif compress == "bgzip":  # the members inflated side by side
        raw = _bgzf(raw, block=5000 + len(raw) % 997)
    elif compress == "bgzip64k":
        raw = _bgzf(raw, eof=False)
    with open(path, "wb") as f:
        f.write(raw)
    return path


def _same_frame(got: pd.DataFrame, ref: pd.DataFrame):
    assert list(got.columns) == list(ref.columns)[:len(got.columns)] and len(got) == len(ref)
    for col in got.columns:
        a, b = got[col], ref[col]
        if isinstance(a.dtype, pd.CategoricalDtype):
            assert list(a.astype(str)) == list(b.astype(str))
        else:
            np.testing.assert_array_equal(a.to_numpy(), b.to_numpy())


@pytest.mark.parametrize("compress", [None, "gzip", "members", "bgzip", "bgzip64k"])
@pytest.mark.parametrize("score", [True, False])
def test_reader_equals_pandas(tmp_path, compress, score, monkeypatch):
    df = _fragments(1, score=score)
    path = _write(df, str(tmp_path / ("f.tsv" + (".gz" if compress else ""))), compress=compress)
    for block in (None, "64", "1000", "100000"):  # tiny blocks: every line is carried over a block boundary
        if block is None:
            monkeypatch.delenv("CGS_BED_BLOCK_BYTES", raising=False)
        else:
            monkeypatch.setenv("CGS_BED_BLOCK_BYTES", block)
        for threads in (1, 3, 0):
            got = fm.read_fragments_from_file(path, n_threads=threads)
            _same_frame(got, df)
            assert got["Start"].dtype == np.int32 and got["End"].dtype == np.int32
            # sorted dictionaries, as pandas' "category" dtype gives
            assert list(got["Name"].cat.categories) == sorted(set(df["Name"]))
            assert list(got["Chromosome"].cat.categories) == sorted(set(df["Chromosome"]))


@needs_reference
@pytest.mark.parametrize("engine", ["pandas", "pyarrow"])
def test_reader_equals_the_reference_reader_run_live(tmp_path, engine):
    ref_read = rf.reference_read_fragments_to_pyranges()
    df = _fragments(2, n=3000)
    header = "# id=sample\n#\tsecond comment\n\n   \n"
    cases = [(_write(df, str(tmp_path / "a.tsv"), header=header), True),
             (_write(df, str(tmp_path / "b.tsv.gz"), header=header, compress="gzip"), True),
             (_write(df.iloc[:, :4], str(tmp_path / "c.tsv"), final_newline=False), False)]
    for path, _ in cases:
        ref = ref_read(path, engine)
        got = fm.read_fragments_from_file(path)
        _same_frame(got, ref)
        for col in ("Chromosome", "Name"):
            assert isinstance(ref[col].dtype, pd.CategoricalDtype) and isinstance(got[col].dtype, pd.CategoricalDtype)
            assert sorted(got[col].cat.categories) == sorted(ref[col].cat.categories.astype(str))
        assert got["Start"].dtype == ref["Start"].dtype == np.int32
    # the reference's error for a file with too few columns, word for word
    short = _write(df.iloc[:, :3], str(tmp_path / "short.tsv"))
    with pytest.raises(ValueError) as e_ref:
        ref_read(short, engine)
    with pytest.raises(ValueError) as e_got:
        fm.read_fragments_from_file(short)
    assert str(e_got.value) == str(e_ref.value)


def test_reader_corner_cases(tmp_path, monkeypatch):
    df = _fragments(3, n=400)
    # Windows line ends, blank lines between records, no newline at the end
    text = "\r\n".join("\t".join(str(v) for v in row) for row in df.itertuples(index=False))
    text = text.replace("\r\n", "\r\n\r\n", 5)
    p = str(tmp_path / "crlf.tsv")
    open(p, "wb").write(text.encode())
    for block in ("16", "97", "1000000"):
        monkeypatch.setenv("CGS_BED_BLOCK_BYTES", block)
        _same_frame(fm.read_fragments_from_file(p, n_threads=4), df)
    monkeypatch.delenv("CGS_BED_BLOCK_BYTES")
    # a regions file: three columns are enough, later columns are ignored, a non-integer fifth column is tolerated
    bed = pd.DataFrame({"Chromosome": ["chr2", "chr1", "chr10"], "Start": [5, 10, 0], "End": [50, 20, 2147483647],
                        "Name": ["a", "b", "c"], "Score": [".", "0.5", "7"], "Strand": ["+", "-", "."]})
    p = _write(bed, str(tmp_path / "r.bed"), header="# comment\n")
    got = fm.read_bed(p)
    assert list(got.columns) == ["Chromosome", "Start", "End"] and list(got["Chromosome"].astype(str)) == ["chr2", "chr1", "chr10"]
    np.testing.assert_array_equal(got["End"].to_numpy(), [50, 20, 2147483647])
    t = fm.read_bed_table(p, 3)
    assert t["n_columns"] == 6 and t["score_is_integer"] is False and t["chromosomes"] == ["chr1", "chr10", "chr2"]
    # empty and header-only files
    for content in ("", "# only a comment\n\n"):
        p = str(tmp_path / "empty.tsv")
        open(p, "w").write(content)
        with pytest.raises(ValueError, match="contains only 0 columns"):
            fm.read_fragments_from_file(p)
    # errors name the line
    lines = ["chr1\t1\t2\tA", "chr1\t3\t4\tB", "chr1\tx\t4\tB", "chr1\t5\t6\tC"]
    p = str(tmp_path / "bad.tsv")
    open(p, "w").write("#h\n" + "\n".join(lines) + "\n")
    with pytest.raises(RuntimeError, match=r"line 4: Start / End is not a 32-bit integer"):
        fm.read_fragments_from_file(p, n_threads=2)
    open(p, "w").write("\n".join(lines[:2] + ["chr1\t3\t4\tB\textra"]) + "\n")
    with pytest.raises(RuntimeError, match=r"line 3: expected 4 columns, saw 5"):
        fm.read_fragments_from_file(p)
    open(p, "w").write("chr1\t1\t3000000000\tA\n")
    with pytest.raises(RuntimeError, match="32-bit"):
        fm.read_fragments_from_file(p)
    with pytest.raises(RuntimeError, match="cannot open"):
        fm.read_fragments_from_file(str(tmp_path / "absent.tsv"))
    # a truncated gzip stream is an error, not a short table
    raw = gzip.compress(("\n".join(lines[:2]) + "\n").encode() * 2000)
    p = str(tmp_path / "cut.tsv.gz")
    open(p, "wb").write(raw[:len(raw) // 2])
    with pytest.raises(RuntimeError):
        fm.read_fragments_from_file(p)
    # the same for BGZF: cut inside a member, and a member whose bytes were damaged (CRC)
    raw = _bgzf(("\n".join(lines[:2]) + "\n").encode() * 20000)
    open(p, "wb").write(raw[:len(raw) // 2])
    with pytest.raises(RuntimeError, match="ends inside a BGZF member"):
        fm.read_fragments_from_file(p)
    damaged = bytearray(raw)
    damaged[len(raw) // 3] ^= 0x55
    open(p, "wb").write(bytes(damaged))
    with pytest.raises(RuntimeError, match="BGZF"):
        fm.read_fragments_from_file(p)
    open(p, "wb").write(raw)
    assert len(fm.read_fragments_from_file(p)) == 40000


def test_reader_large_file_many_threads(tmp_path):
    """A file of several default-size pieces per thread: row order and codes survive the parallel tokeniser."""
    rng = np.random.default_rng(0)
    n = 400000
    names = np.array([f"BC{j:05d}" for j in range(3000)])
    chrom = np.array(["chr1", "chr2", "chr3", "chrX"])
    c, b = rng.integers(0, 4, n), rng.integers(0, 3000, n)
    start = rng.integers(0, 10**8, n)
    end = start + rng.integers(1, 1000, n)
    p = str(tmp_path / "big.tsv.gz")
    with gzip.open(p, "wt", compresslevel=1) as f:
        f.write("\n".join(f"{chrom[i]}\t{s}\t{e}\t{names[j]}\t1" for i, s, e, j in zip(c, start, end, b)) + "\n")
    os.environ["CGS_BED_BLOCK_BYTES"] = str(1 << 20)
    try:
        t = fm.read_bed_table(p, 4, n_threads=8)
    finally:
        del os.environ["CGS_BED_BLOCK_BYTES"]
    np.testing.assert_array_equal(np.array(t["chromosomes"])[t["chromosome"]], chrom[c])
    np.testing.assert_array_equal(np.array(t["names"])[t["name"]], names[b])
    np.testing.assert_array_equal(t["start"], start)
    np.testing.assert_array_equal(t["end"], end)
    assert t["score"].sum() == n and t["score_is_integer"]


@needs_reference
def test_collapse_duplicates_equals_the_reference_function(tmp_path):
    collapse = rf.reference_collapse_duplicates()
    df = _fragments(4, n=6000, score=False, duplicates=True)
    ref = pd.concat([collapse(df[df.Chromosome == c]) for c in sorted(set(df.Chromosome))])  # cistopic_class.py:822-827
    ref = ref.astype({"Start": np.int64, "End": np.int64, "Score": np.int64})
    key = ["Chromosome", "Start", "End", "Name"]
    ref = ref.sort_values(key).reset_index(drop=True)
    path = _write(df, str(tmp_path / "dup.tsv"))
    for threads in (1, 4):
        t = fm.read_bed_table(path, 4, n_threads=threads, collapse_duplicates=True)
        got = fm._table_to_frame(t)
        assert t["duplicates_removed"] == len(df) - len(ref) > 0
        # the first record of every group keeps its place
        first = df.drop_duplicates(key).reset_index(drop=True)
        _same_frame(got[key], first)
        got = got.astype({"Chromosome": str, "Name": str, "Start": np.int64, "End": np.int64, "Score": np.int64})
        pd.testing.assert_frame_equal(got.sort_values(key).reset_index(drop=True), ref.astype({"Chromosome": str, "Name": str}),
                                      check_dtype=False)
    # the DataFrame route does the same
    got = fm._collapse_duplicates_frame(df).astype({"Start": np.int64, "End": np.int64, "Score": np.int64})
    pd.testing.assert_frame_equal(got.sort_values(key).reset_index(drop=True), ref, check_dtype=False, check_categorical=False)
    # a file WITH a Score column is left alone (cistopic_class.py:819)
    with_score = _write(df.assign(Score=1), str(tmp_path / "dup5.tsv"))
    t = fm.read_bed_table(with_score, 4, collapse_duplicates=True)
    assert len(t["start"]) == len(df) and t["duplicates_removed"] is None


def test_collapse_duplicates_many_rows(tmp_path):
    """Above the size where the de-duplication runs one hash partition per thread."""
    rng = np.random.default_rng(5)
    n = 150000
    base = pd.DataFrame({"Chromosome": rng.choice(["chr1", "chr2"], n), "Start": rng.integers(0, 3000, n),
                         "End": rng.integers(3000, 3050, n), "Name": rng.choice([f"B{j}" for j in range(30)], n)})
    path = _write(base, str(tmp_path / "many.tsv"))
    key = ["Chromosome", "Start", "End", "Name"]
    want = base.groupby(key, sort=True).size().reset_index(name="Score")
    for threads in (1, 5):
        got = fm._table_to_frame(fm.read_bed_table(path, 4, n_threads=threads, collapse_duplicates=True))
        got = got.astype({"Chromosome": str, "Name": str}).sort_values(key).reset_index(drop=True)
        pd.testing.assert_frame_equal(got, want, check_dtype=False)
        assert got["Score"].sum() == n


def test_blacklist_mask_equals_brute_force():
    rng = np.random.default_rng(6)
    n, m = 3000, 200
    regions = pd.DataFrame({"Chromosome": rng.choice(["chr1", "chr2", "chr3"], n), "Start": rng.integers(0, 100000, n)})
    regions["End"] = regions["Start"] + rng.integers(1, 500, n)
    black = pd.DataFrame({"Chromosome": rng.choice(["chr1", "chr2", "chrM"], m), "Start": rng.integers(0, 100000, m)})
    black["End"] = black["Start"] + rng.integers(1, 3000, m)  # long intervals that swallow shorter, later ones
    # touching intervals do not overlap
    regions.loc[0, ["Chromosome", "Start", "End"]] = ["chr1", 1000, 2000]
    black.loc[0, ["Chromosome", "Start", "End"]] = ["chr1", 2000, 2500]
    black.loc[1, ["Chromosome", "Start", "End"]] = ["chr1", 500, 1000]
    keep = fm.regions_outside_blacklist(regions, black)
    want = np.ones(n, bool)
    for i, r in enumerate(regions.itertuples(index=False)):
        b = black[black.Chromosome == r.Chromosome]
        want[i] = not ((b.Start < r.End) & (r.Start < b.End)).any()
    np.testing.assert_array_equal(keep, want)
    assert 0 < keep.sum() < n and keep[regions.Chromosome == "chr3"].all()


def test_fragments_in_peaks_equals_brute_force():
    """fragments.py:945-1021: per barcode, the fragments that overlap at least one region -- how many, and their scores."""
    frags = _fragments(14, n=3000, n_cells=20)
    rng = np.random.default_rng(14)
    regions = pd.DataFrame({"Chromosome": rng.choice(["chr1", "chr2", "chrX", "chr7"], 80), "Start": rng.integers(0, 200000, 80)})
    regions["End"] = regions["Start"] + rng.integers(50, 4000, 80)  # overlapping regions: a fragment still counts once
    got = fm.get_fragments_in_peaks(frags, regions)
    inside = np.zeros(len(frags), bool)
    for i, f in enumerate(frags.itertuples(index=False)):
        r = regions[regions.Chromosome == f.Chromosome]
        inside[i] = ((r.Start < f.End) & (f.Start < r.End)).any()
    sub = frags[inside]
    want = sub.groupby("Name", sort=False).agg(total=("Score", "sum"), unique=("Score", "size")).reset_index()
    assert 0 < inside.sum() < len(frags)
    assert got["CB"].tolist() == want["Name"].tolist()  # order of first appearance
    np.testing.assert_array_equal(got["total_fragments_in_peaks_count"], want["total"])
    np.testing.assert_array_equal(got["unique_fragments_in_peaks_count"], want["unique"])


def _count_matrix(seed=7, n_regions=120, n_cells=30):
    rng = np.random.default_rng(seed)
    m = sp.random(n_regions, n_cells, density=0.2, random_state=seed, format="csr", dtype=np.float64)
    m.data = rng.integers(1, 4, m.nnz).astype(np.int32)
    m = sp.csr_matrix(m, dtype=np.int32)
    regions = [f"chr{1 + i % 3}:{1000 * i}-{1000 * i + 400 + i}" for i in range(n_regions)]
    cells = [f"CELL{j:03d}-1" for j in range(n_cells)]
    return m, cells, regions


@needs_reference
@pytest.mark.parametrize("kw", [dict(), dict(is_acc=2), dict(is_acc=3, tag_cells=False)])
def test_create_cistopic_object_equals_the_reference_function(kw):
    # (min_frag / min_cell != 1 cannot be run live: the reference indexes a scipy matrix with a pandas boolean Series,
    #  which the scipy of this image refuses, and region_data with [mask, :], a pandas error -- see the next test)
    ref_create = rf.reference_create_cistopic_object(fm.region_names_to_coordinates)
    m, cells, regions = _count_matrix()
    ref = ref_create(m.copy(), cell_names=list(cells), region_names=list(regions), project="s1", **kw)
    got = fm.create_cistopic_object(m.copy(), cell_names=list(cells), region_names=list(regions), project="s1", **kw)
    assert got.cell_names == ref.cell_names and got.region_names == ref.region_names
    assert (got.fragment_matrix != ref.fragment_matrix).nnz == 0 and (got.binary_matrix != ref.binary_matrix).nnz == 0
    pd.testing.assert_frame_equal(got.cell_data, ref.cell_data)
    pd.testing.assert_frame_equal(got.region_data, ref.region_data, check_names=False)
    # a DataFrame goes in the same way (cistopic_class.py:564-568)
    frame = pd.DataFrame(m.toarray(), index=regions, columns=cells)
    got2 = fm.create_cistopic_object(frame, project="s1", **kw)
    assert got2.cell_names == ref.cell_names and (got2.binary_matrix != ref.binary_matrix).nnz == 0


@needs_reference
def test_object_from_matrix_file_equals_the_reference_function(tmp_path):
    ref_create = rf.reference_create_cistopic_object_from_matrix_file(fm.region_names_to_coordinates)
    m, cells, regions = _count_matrix(13)
    frame = pd.DataFrame(m.toarray(), index=regions, columns=cells)
    plain, packed = str(tmp_path / "matrix.tsv"), str(tmp_path / "matrix.tsv.gz")
    # the format the reference expects: a header of cell names only, so that pandas takes the first field of every
    # record -- the region -- as the row name
    frame.to_csv(plain, sep="\t", index_label=False)
    frame.to_csv(packed, sep="\t", index_label=False, compression="gzip")
    sample = pd.DataFrame({"sample_id": "run7", "batch": np.arange(len(cells))}, index=[c + "___mx" for c in cells])
    for path, kw in ((plain, {}), (packed, {"compression": "gzip"}),
                     (plain, {"is_acc": 2, "sample_id": sample, "path_to_fragments": {"mx": "f.tsv.gz"}})):
        ref = ref_create(path, project="mx", **{k: (v.copy() if isinstance(v, pd.DataFrame) else v) for k, v in kw.items()})
        got = fm.create_cistopic_object_from_matrix_file(path, project="mx", **kw)
        assert got.cell_names == ref.cell_names and got.region_names == ref.region_names
        assert (got.binary_matrix != ref.binary_matrix).nnz == 0 and (got.fragment_matrix != ref.fragment_matrix).nnz == 0
        pd.testing.assert_frame_equal(got.cell_data.sort_index(axis=1), ref.cell_data.sort_index(axis=1))
        pd.testing.assert_frame_equal(got.region_data, ref.region_data, check_names=False)
        assert got.path_to_fragments == ref.path_to_fragments


def test_create_cistopic_object_min_frag_min_cell():
    """cistopic_class.py:617-638: cells with fewer than min_frag fragments go first; the region statistics, and with
    them the min_cell filter, are taken over the cells that are left."""
    m, cells, regions = _count_matrix(11)
    dense = m.toarray()
    keep_r = (dense > 0).any(axis=1)
    dense, regions_k = dense[keep_r], [r for r, k in zip(regions, keep_r) if k]
    min_frag, min_cell = int(np.median(dense.sum(axis=0))), 5
    obj = fm.create_cistopic_object(m.copy(), list(cells), list(regions), min_frag=min_frag, min_cell=min_cell, project="q")
    keep_c = dense.sum(axis=0) >= min_frag
    dense = dense[:, keep_c]
    keep_r2 = (dense > 0).sum(axis=1) >= min_cell
    assert 0 < keep_c.sum() < len(cells) and 0 < keep_r2.sum() < len(regions_k)
    np.testing.assert_array_equal(obj.fragment_matrix.toarray(), dense[keep_r2])
    np.testing.assert_array_equal(obj.binary_matrix.toarray(), (dense[keep_r2] > 0).astype(np.int32))
    assert obj.cell_names == [c + "___q" for c, k in zip(cells, keep_c) if k]
    assert obj.region_names == [r for r, k in zip(regions_k, keep_r2) if k]
    assert list(obj.cell_data.index) == obj.cell_names and list(obj.region_data.index) == obj.region_names
    np.testing.assert_array_equal(obj.region_data["cisTopic_nr_acc"].to_numpy(), (dense[keep_r2] > 0).sum(axis=1))


@needs_reference
def test_add_cell_data_equals_the_reference_method():
    ref_add = rf.reference_add_cell_data()
    m, cells, regions = _count_matrix(8)
    rng = np.random.default_rng(8)
    bare = pd.DataFrame({"Unique_nr_frag": rng.integers(1, 1000, len(cells)), "barcode": cells}, index=cells)
    tagged = bare.copy()
    tagged.index = [c + "___s1" for c in cells]
    partial = bare.iloc[5:].copy()
    overwrite = pd.DataFrame({"cisTopic_nr_frag": np.arange(len(cells)), "extra": "x"}, index=tagged.index)
    for meta in (bare, tagged, partial, overwrite, bare.iloc[::-1]):
        got = fm.create_cistopic_object(m.copy(), list(cells), list(regions), project="s1")
        ref = fm.create_cistopic_object(m.copy(), list(cells), list(regions), project="s1")
        got.add_cell_data(meta.copy(), "___")
        ref_add(ref, meta.copy(), "___")
        pd.testing.assert_frame_equal(got.cell_data.sort_index(axis=1), ref.cell_data.sort_index(axis=1))
        assert list(got.cell_data.index) == got.cell_names
    tag = rf.reference_prepare_tag_cells()
    names = ["ACGTACGT-1___s", "ACGT-1-sample", "TTTT-12-x-y", "plain", "cell_7-3", "ACGT-1"]
    for pattern in ("___", "-", "_"):
        assert fm.prepare_tag_cells(names, pattern) == tag(names, pattern)


def _host_counts(device, n_chrom, chrom_ptr, s_start, s_end, f_chrom, f_start, f_end, f_cell, n_cells):
    """What cgs_fragment_counts_* computes (include/cgs_b200.h), by brute force: the stand-in for the device step."""
    rows, cols = [], []
    for c in range(n_chrom):
        lo, hi = chrom_ptr[c], chrom_ptr[c + 1]
        sel = np.flatnonzero((f_chrom == c) & (f_cell >= 0) & (f_cell < n_cells))
        if hi == lo or len(sel) == 0:
            continue
        hit = (s_start[lo:hi, None] < f_end[None, sel]) & (f_start[None, sel] < s_end[lo:hi, None])
        r, f = np.nonzero(hit)
        rows.append(r + lo)
        cols.append(f_cell[sel][f])
    rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
    cols = np.concatenate(cols) if cols else np.zeros(0, np.int64)
    m = sp.coo_matrix((np.ones(len(rows), np.int32), (rows, cols)), shape=(len(s_start), n_cells)).tocsr()
    m.sum_duplicates()
    m.sort_indices()
    return m.indptr.astype(np.int64), m.indices.astype(np.int32), m.data.astype(np.int32)


def test_count_fragments_host_side_with_categorical_columns(monkeypatch):
    """count_fragments_in_regions around the device step: categorical columns as the reader returns them (with
    unused categories after a barcode filter), chromosomes without regions, a barcode whitelist."""
    monkeypatch.setattr(fm, "_device_counts", _host_counts)
    frags = _fragments(12, n=3000, n_cells=30)
    rng = np.random.default_rng(12)
    regions = pd.DataFrame({"Chromosome": rng.choice(["chr1", "chr3", "chrX", "chr9"], 150), "Start": rng.integers(0, 200000, 150)})
    regions["End"] = regions["Start"] + rng.integers(100, 2000, 150)
    regions = regions.drop_duplicates().reset_index(drop=True)
    want = fo.fragment_matrix(regions, frags)
    cat = frags.astype({"Chromosome": "category", "Name": "category"})
    for f in (frags, cat):
        m, rn, cn = fm.count_fragments_in_regions(f, regions)
        got = pd.DataFrame(m.toarray(), index=rn, columns=cn)
        assert cn == sorted(want.columns) and sorted(rn) == sorted(want.index)
        np.testing.assert_array_equal(got.loc[want.index, want.columns].to_numpy(), want.to_numpy())
    valid = sorted(set(frags.Name))[5:20][::-1]
    sub = cat[cat.Name.isin(set(valid))]  # the categories of the other barcodes are still there
    want_v = fo.fragment_matrix(regions, frags, valid_bc=valid)
    for f, bc in ((sub, None), (cat, valid)):
        m, rn, cn = fm.count_fragments_in_regions(f, regions, valid_bc=bc)
        got = pd.DataFrame(m.toarray(), index=rn, columns=cn)
        assert cn == (sorted(want_v.columns) if bc is None else [b for b in valid if b in set(want_v.columns)])
        np.testing.assert_array_equal(got.loc[want_v.index, want_v.columns].to_numpy(), want_v.to_numpy())


def test_object_from_fragment_files_host_logic(tmp_path, monkeypatch):
    """Everything of create_cistopic_object_from_fragments but the device counting: files in, duplicates collapsed,
    barcodes selected by ``valid_bc`` / CellRanger metrics, blacklist applied, Unique_nr_frag joined."""
    monkeypatch.setattr(fm, "_device_counts", _host_counts)
    frags = _fragments(9, n=4000, n_cells=25, score=False, duplicates=True)
    rng = np.random.default_rng(9)
    starts = np.sort(rng.choice(200, 90, replace=False)) * 1000
    regions = pd.DataFrame({"Chromosome": rng.choice(["chr1", "chr2", "chr3", "chrX"], 90), "Start": starts,
                            "End": starts + rng.integers(300, 900, 90)})
    black = pd.DataFrame({"Chromosome": ["chr1", "chr2"], "Start": [0, 50000], "End": [40000, 90000]})
    f_path = _write(frags, str(tmp_path / "fragments.tsv.gz"), header="# id=s\n", compress="bgzip")
    r_path = _write(regions, str(tmp_path / "regions.bed"))
    b_path = _write(black, str(tmp_path / "blacklist.bed"))

    unique = frags.drop_duplicates()
    obj = fm.create_cistopic_object_from_fragments(f_path, r_path, project="s")
    want = fo.fragment_matrix(regions, unique)
    got = pd.DataFrame(obj.fragment_matrix.toarray(), index=obj.region_names, columns=[c.split("___")[0] for c in obj.cell_names])
    np.testing.assert_array_equal(got.loc[want.index, want.columns].to_numpy(), want.to_numpy())
    assert obj.path_to_fragments == {"s": f_path} and obj.project == "s"
    per_bc = unique.groupby("Name").size()
    np.testing.assert_array_equal(obj.cell_data["Unique_nr_frag"].to_numpy().astype(np.int64),
                                  per_bc.loc[[c.split("___")[0] for c in obj.cell_names]].to_numpy())
    assert list(obj.cell_data["barcode"]) == [c.split("___")[0] for c in obj.cell_names]
    # duplicates kept when asked (every record counts, cistopic_class.py:830-831)
    obj_dup = fm.create_cistopic_object_from_fragments(f_path, r_path, project="s", check_for_duplicates=False)
    assert obj_dup.fragment_matrix.sum() == fo.fragment_matrix(regions, frags).to_numpy().sum() > obj.fragment_matrix.sum()
    # the same from DataFrames (``fragments_df`` of the reference)
    obj_df = fm.create_cistopic_object_from_fragments(None, regions, project="s", fragments_df=frags)
    assert obj_df.cell_names == obj.cell_names and (obj_df.fragment_matrix != obj.fragment_matrix).nnz == 0
    # blacklist
    obj_b = fm.create_cistopic_object_from_fragments(f_path, r_path, path_to_blacklist=b_path, project="s")
    gone = set(obj.region_names) - set(obj_b.region_names)
    coords = fm.region_names_to_coordinates(sorted(gone))
    assert len(gone) > 0 and all(((c == "chr1") and s < 40000) or ((c == "chr2") and e > 50000 and s < 90000)
                                 for c, s, e in coords.itertuples(index=False))
    # valid_bc, then CellRanger metrics
    valid = sorted(set(frags.Name))[3:15]
    obj_v = fm.create_cistopic_object_from_fragments(f_path, r_path, valid_bc=valid, project="s")
    assert [c.split("___")[0] for c in obj_v.cell_names] == [b for b in valid if b in set(got.columns)]
    metrics = pd.DataFrame({"barcode": ["NO_BARCODE"] + sorted(set(frags.Name)), "total": 0,
                            "is__cell_barcode": [0] + [1, 0] * 12 + [1], "passed_filters": np.arange(26)})
    m_path = str(tmp_path / "singlecell.csv")
    metrics.to_csv(m_path, index=False)
    obj_m = fm.create_cistopic_object_from_fragments(f_path, r_path, metrics=m_path, project="s")
    cells_m = set(metrics.barcode[metrics.is__cell_barcode == 1])
    assert {c.split("___")[0] for c in obj_m.cell_names} == cells_m & set(got.columns)
    assert "passed_filters" in obj_m.cell_data.columns and "Unique_nr_frag" not in obj_m.cell_data.columns
    assert list(obj_m.cell_data["barcode"]) == [c.split("___")[0] for c in obj_m.cell_names]
