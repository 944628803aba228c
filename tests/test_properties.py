"""Property tests (hypothesis) of the size-independent invariants SURVEY.md section 4(iii) lists: count
conservation, z round trips, equivalence of the two matrix layouts.  The CPU half exercises the oracle and
the host logic; the half marked ``gpu`` runs the same properties through the C ABI on the device."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import lda_oracle as lo  # noqa: E402
from pycistopic_b200 import adlda, lda_models  # noqa: E402

COMMON = dict(deadline=None, suppress_health_check=list(HealthCheck))


@st.composite
def binary_matrices(draw, max_cells=40, max_regions=120):
    """regions x cells int32 CSR with at least one token; may hold empty cells and empty regions."""
    D = draw(st.integers(1, max_cells))
    W = draw(st.integers(1, max_regions))
    density = draw(st.floats(0.01, 0.6))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    rng = np.random.default_rng(seed)
    dense = (rng.random((W, D)) < density).astype(np.int32)
    if dense.sum() == 0:
        dense[rng.integers(W), rng.integers(D)] = 1
    return sp.csr_matrix(dense)


# ---- CPU: the oracle and the host logic ---------------------------------------------------------------
@settings(max_examples=40, **COMMON)
@given(m=binary_matrices(), K=st.integers(1, 12), sweeps=st.integers(0, 4), seed=st.integers(0, 10 ** 6))
def test_oracle_conserves_counts(m, K, sweeps, seed):
    X = sp.csr_matrix(m.T)
    o = lo.OracleLDA(K, n_iter=sweeps, alpha=50.0 / K, eta=0.1, random_state=seed, refresh=max(1, sweeps), keep_z=True).fit(X)
    assert o.nz_.sum() == X.nnz and np.array_equal(o.nzw_.sum(axis=1), o.nz_)
    assert np.array_equal(o.ndz_.sum(axis=1), np.diff(X.indptr))
    assert np.array_equal(o.nzw_.sum(axis=0), np.asarray(X.sum(axis=0)).ravel())
    nzw, ndz, nz = lo.counts_from_z(o.WS, o.DS, o.ZS, X.shape[0], X.shape[1], K)
    assert np.array_equal(nzw, o.nzw_) and np.array_equal(ndz, o.ndz_) and np.array_equal(nz, o.nz_)
    assert (o.nzw_ >= 0).all() and (o.ndz_ >= 0).all()


@settings(max_examples=40, **COMMON)
@given(m=binary_matrices())
def test_oracle_token_lists_two_statements_agree(m):
    X = sp.csr_matrix(m.T)
    WS, DS = lo.matrix_to_lists(X)
    WS2, DS2 = lo.matrix_to_lists_py(X)
    assert np.array_equal(WS, WS2) and np.array_equal(DS, DS2)
    assert np.array_equal(WS, X.indices) and np.array_equal(DS, np.repeat(np.arange(X.shape[0]), np.diff(X.indptr)))


@settings(max_examples=60, **COMMON)
@given(counts=st.lists(st.integers(0, 5000), min_size=1, max_size=200), parts=st.integers(1, 16))
def test_balanced_partition_is_a_partition(counts, parts):
    D = len(counts)
    parts = min(parts, D)
    ptr, cell_ptr = adlda.balanced_cell_partition(counts, parts)
    assert ptr[0] == 0 and ptr[-1] == D and np.all(np.diff(ptr) >= 1)
    assert np.diff(cell_ptr[ptr]).sum() == sum(counts)


@settings(max_examples=60, **COMMON)
@given(counts=st.lists(st.integers(0, 3000), min_size=1, max_size=300), blocks=st.integers(1, 12))
def test_balanced_region_blocks_cover_the_regions(counts, blocks):
    """adlda.balanced_region_blocks: bounds start at 0, end at n_regions, never decrease, and no block holds more than
    its share of the tokens plus one region's worth."""
    if sum(counts) == 0:
        counts = counts[:-1] + [1]
    bounds = adlda.balanced_region_blocks(counts, blocks)
    W = len(counts)
    assert bounds[0] == 0 and bounds[-1] == W and np.all(np.diff(bounds) >= 0) and len(bounds) == min(blocks, W) + 1
    cum = np.concatenate([[0], np.cumsum(counts)])
    per_block = np.diff(cum[bounds])
    assert per_block.sum() == sum(counts)
    assert per_block.max() <= sum(counts) / (len(bounds) - 1) + 2 * max(counts) + 1e-9


@settings(max_examples=60, **COMMON)
@given(topics=st.lists(st.integers(1, 500), min_size=1, max_size=24), workers=st.integers(1, 16))
def test_schedule_covers_every_model_once(topics, workers):
    plan = lda_models.schedule_models(topics, workers)
    assert len(plan) == workers and sorted(i for p in plan for i in p) == list(range(len(topics)))
    loads = [sum(lda_models.model_cost(topics[i]) for i in p) for p in plan]
    longest = max(lda_models.model_cost(k) for k in topics)
    assert max(loads) <= sum(lda_models.model_cost(k) for k in topics) / workers + longest + 1e-9


@settings(max_examples=80, **COMMON)
@given(costs=st.lists(st.floats(0.2, 5.0), min_size=1, max_size=16), workers=st.integers(1, 12),
       ov=st.floats(0.0, 0.3))
def test_split_schedule_partitions_every_model(costs, workers, ov):
    """schedule_models_split: every model's parts tile [0, 1) in order on CONSECUTIVE workers, no part is a sliver,
    no worker exceeds the returned level.  (The minimum part size can make this plan worse than whole-model placement
    on few workers; schedule_models_paired is the one that is never worse.)"""
    overhead = [ov] * len(costs)
    plan, level = lda_models.schedule_models_split(costs, workers, overhead, min_fraction=0.2)
    assert len(plan) == workers
    where = {}
    for w, items in enumerate(plan):
        load = 0.0
        for (i, f0, f1) in items:
            assert 0.0 <= f0 < f1 <= 1.0
            whole = f0 == 0.0 and f1 == 1.0
            assert whole or (f1 - f0) >= 0.2 - 1e-9
            load += (f1 - f0) * costs[i] + (0.0 if whole else overhead[i])
            where.setdefault(i, []).append((f0, f1, w))
        assert load <= level * (1 + 1e-6) + 1e-9
    assert sorted(where) == list(range(len(costs)))
    for i, parts in where.items():
        parts.sort()
        assert parts[0][0] == 0.0 and parts[-1][1] == 1.0
        for a, b in zip(parts, parts[1:]):
            assert abs(a[1] - b[0]) < 1e-12 and b[2] == a[2] + 1   # contiguous fractions on neighbouring workers
    assert level <= sum(costs) + len(costs) * ov + 1e-9


@settings(max_examples=60, **COMMON)
@given(costs=st.lists(st.floats(0.2, 5.0), min_size=1, max_size=12), workers=st.integers(1, 10))
def test_paired_schedule_properties(costs, workers):
    """schedule_models_paired: every model is placed exactly once (whole, or as two parts that tile [0, 1) on two
    different workers), a worker holds at most one part and holds it FIRST, no part is below the minimum fraction, and
    the predicted makespan is never worse than plain longest-first placement of whole models."""
    plan, makespan = lda_models.schedule_models_paired(costs, workers)
    assert len(plan) == workers
    seen = {}
    for w, items in enumerate(plan):
        parts_here = [(i, f0, f1) for (i, f0, f1) in items if not (f0 == 0.0 and f1 == 1.0)]
        assert len(parts_here) <= 1
        if parts_here:
            assert items[0] == parts_here[0] and parts_here[0][2] - parts_here[0][1] >= 0.25 - 1e-9
        for (i, f0, f1) in items:
            seen.setdefault(i, []).append((f0, f1, w))
    assert sorted(seen) == list(range(len(costs)))
    for i, parts in seen.items():
        parts.sort()
        if len(parts) == 1:
            assert parts[0][:2] == (0.0, 1.0)
        else:
            assert len(parts) == 2 and parts[0][0] == 0.0 and parts[1][1] == 1.0 and parts[0][1] == parts[1][0]
            assert parts[0][2] != parts[1][2]
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * workers
    for i in order:
        j = min(range(workers), key=lambda j: (loads[j], j))
        loads[j] += costs[i]
    assert makespan <= max(loads) + 1e-9


# ---- GPU: the same invariants through the C ABI --------------------------------------------------------
# ---- CPU: the fragments / BED reader (host code behind the C ABI) ----------------------------------------
_FIELD = st.text(alphabet=st.characters(min_codepoint=33, max_codepoint=126), min_size=1, max_size=12)


@st.composite
def bed_files(draw):
    """Records of a BED-like file with its layout choices: 3 to 6 columns, '\n' or '\r\n', blank lines between
    records, '#' / blank header lines, with or without a last line end."""
    n_cols = draw(st.integers(3, 6))
    n = draw(st.integers(1, 60))
    chroms = draw(st.lists(_FIELD.filter(lambda f: not f.startswith("#")), min_size=1, max_size=4, unique=True))
    names = draw(st.lists(_FIELD, min_size=1, max_size=6, unique=True))
    rows = []
    for _ in range(n):
        row = [draw(st.sampled_from(chroms)), draw(st.integers(-5, 2 ** 31 - 1)), draw(st.integers(0, 2 ** 31 - 1))]
        if n_cols >= 4:
            row.append(draw(st.sampled_from(names)))
        if n_cols >= 5:
            row.append(draw(st.integers(-3, 1000)))
        if n_cols >= 6:
            row.append(draw(st.sampled_from(["+", "-", "."])))
        rows.append(row)
    eol = draw(st.sampled_from(["\n", "\r\n"]))
    header = "".join(draw(st.lists(st.sampled_from(["# comment" + eol, eol, "#" + eol, "  " + eol]), max_size=3)))
    body = []
    for row in rows:
        body.append("\t".join(str(v) for v in row))
        body.extend([""] * draw(st.integers(0, 1) if n < 20 else st.just(0)))  # blank lines between records
    while body and body[-1] == "":
        body.pop()
    text = header + eol.join(body) + (eol if draw(st.booleans()) else "")
    return rows, n_cols, text


@settings(max_examples=60, **COMMON)
@given(case=bed_files(), block=st.sampled_from([16, 23, 64, 257, 1 << 20]), threads=st.integers(1, 5),
       packing=st.sampled_from(["plain", "gzip", "bgzf"]))
def test_bed_reader_returns_the_records_of_the_file(case, block, threads, packing, tmp_path_factory):
    import gzip

    from pycistopic_b200 import fragment_matrix as fm
    from test_cpu_fragment_files import _bgzf
    rows, n_cols, text = case
    raw = text.encode()
    if packing == "gzip":
        raw = gzip.compress(raw)
    elif packing == "bgzf":
        raw = _bgzf(raw, block=37)
    path = str(tmp_path_factory.mktemp("bed") / "records.bed")
    with open(path, "wb") as f:
        f.write(raw)
    os.environ["CGS_BED_BLOCK_BYTES"] = str(block)
    try:
        t = fm.read_bed_table(path, 3, n_threads=threads)
    finally:
        del os.environ["CGS_BED_BLOCK_BYTES"]
    assert t["n_columns"] == n_cols and len(t["start"]) == len(rows)
    assert t["chromosomes"] == sorted({r[0] for r in rows}) and [t["chromosomes"][c] for c in t["chromosome"]] == [r[0] for r in rows]
    assert t["start"].tolist() == [r[1] for r in rows] and t["end"].tolist() == [r[2] for r in rows]
    if n_cols >= 4:
        assert t["names"] == sorted({r[3] for r in rows}) and [t["names"][c] for c in t["name"]] == [r[3] for r in rows]
    if n_cols >= 5:
        assert t["score_is_integer"] and t["score"].tolist() == [r[4] for r in rows]
    if n_cols == 4:  # duplicates collapse to their first record, Score = multiplicity (utils.py:284-293)
        c = fm.read_bed_table(path, 4, n_threads=threads, collapse_duplicates=True)
        seen = {}
        for r in rows:
            seen[tuple(r)] = seen.get(tuple(r), 0) + 1
        got = [(c["chromosomes"][a], s, e, c["names"][b]) for a, s, e, b in zip(c["chromosome"], c["start"].tolist(),
                                                                                c["end"].tolist(), c["name"])]
        assert got == list(seen) and c["score"].tolist() == list(seen.values())
        assert c["duplicates_removed"] == len(rows) - len(seen)


_INTERVAL = st.tuples(st.sampled_from(["chr1", "chr2", "chrX"]), st.integers(0, 60), st.integers(0, 12))


@settings(max_examples=150, **COMMON)
@given(a=st.lists(_INTERVAL, min_size=1, max_size=25), b=st.lists(_INTERVAL, min_size=0, max_size=25))
def test_interval_overlap_search_equals_all_pairs(a, b):
    """The sorted-interval search behind the blacklist filter and get_fragments_in_peaks: half-open intervals, touching
    ends do not overlap, nested and duplicated intervals, chromosomes present on one side only."""
    import pandas as pd

    from pycistopic_b200 import fragment_matrix as fm
    frame = lambda rows: pd.DataFrame({"Chromosome": [r[0] for r in rows], "Start": [r[1] for r in rows],  # noqa: E731
                                       "End": [r[1] + r[2] + 1 for r in rows]}).astype({"Start": np.int64, "End": np.int64})
    A, B = frame(a), frame(b)
    want = np.array([any(c == c2 and s < e2 and s2 < e for c2, s2, e2 in B.itertuples(index=False))
                     for c, s, e in A.itertuples(index=False)], dtype=bool)
    np.testing.assert_array_equal(fm.regions_outside_blacklist(A, B), ~want)


@pytest.mark.gpu
@settings(max_examples=25, **COMMON)
@given(m=binary_matrices(), K=st.integers(1, 70), sweeps=st.integers(0, 3), seed=st.integers(0, 10 ** 6))
def test_device_conserves_counts_and_round_trips_z(m, K, sweeps, seed):
    from pycistopic_b200 import Corpus, DeviceModel
    X = sp.csr_matrix(m.T)
    corpus = Corpus.from_regions_by_cells(m)
    WS, DS = corpus.token_lists()
    oWS, oDS = lo.matrix_to_lists(X)
    assert np.array_equal(WS, oWS) and np.array_equal(DS, oDS)
    mdl = DeviceModel(corpus, K, 50.0 / K, 0.1, seed=seed)
    mdl.init()
    mdl.sweep(sweeps)
    z = mdl.get_z()
    assert z.min() >= 0 and z.max() < K
    nzw, ndz, nz = mdl.counts()
    rnzw, rndz, rnz = lo.counts_from_z(oWS, oDS, z, X.shape[0], X.shape[1], K)
    assert np.array_equal(nzw, rnzw) and np.array_equal(ndz, rndz) and np.array_equal(nz, rnz)
    assert mdl.verify_counts() == {"n_dk": 0, "n_wk": 0, "n_k": 0}
    h = mdl.state_hash()
    assert h["sum_nwk"] == X.nnz == h["sum_nk"]
    # set_z / get_z round trip with an arbitrary assignment
    rng = np.random.default_rng(seed)
    z2 = rng.integers(0, K, size=len(oWS)).astype(np.int32)
    mdl.set_z(z2)
    assert np.array_equal(mdl.get_z(), z2)
    nzw, ndz, nz = mdl.counts()
    rnzw, rndz, rnz = lo.counts_from_z(oWS, oDS, z2, X.shape[0], X.shape[1], K)
    assert np.array_equal(nzw, rnzw) and np.array_equal(ndz, rndz) and np.array_equal(nz, rnz)
    mdl.close()
    corpus.close()


@pytest.mark.gpu
@settings(max_examples=25, **COMMON)
@given(m=binary_matrices(), shuffle_seed=st.integers(0, 10 ** 6))
def test_device_token_list_is_layout_independent(m, shuffle_seed):
    """regions x cells (the reference's binary_matrix, transposed on the device) and cells x regions with
    shuffled, duplicated column indices give the same canonical token list."""
    from pycistopic_b200 import Corpus
    X = sp.csr_matrix(m.T)
    a = Corpus.from_regions_by_cells(m)
    rng = np.random.default_rng(shuffle_seed)
    indptr, indices = X.indptr.astype(np.int64), X.indices.copy()
    for d in range(X.shape[0]):
        seg = indices[indptr[d]:indptr[d + 1]]
        rng.shuffle(seg)
    Y = sp.csr_matrix((np.ones(len(indices), np.int32), indices, indptr), shape=X.shape)
    b = Corpus.from_cells_by_regions(Y)
    wa, da = a.token_lists()
    wb, db = b.token_lists()
    assert np.array_equal(wa, wb) and np.array_equal(da, db) and a.n_tokens == X.nnz
    a.close()
    b.close()
