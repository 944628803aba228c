"""Seeded inputs and (de)serialisation shared by tests/golden/make_host_mirrors.py (runs the REFERENCE's functions, needs
/root/reference) and tests/test_cpu_host_mirrors_golden.py (runs the mirrors against the frozen outputs, anywhere)."""
import numpy as np
import pandas as pd
import scipy.sparse as sp

PACKING_CASES = [(7, 5, True, False), (3, 5, True, True), (6, 2, False, False), (5, 5, False, True)]  # K, top, a_by, e_by
OBJECT_CASES = [dict(), dict(is_acc=2), dict(is_acc=3, tag_cells=False)]
DIFF_CASES = [dict(), dict(var_features=slice(5, 100, 2)),
              dict(contrasts=[[["A"], ["B"]], [["A", "B"], ["C"]]], adjpval_thr=0.2, log2fc_thr=0.5)]
READER_TEXT = ("# id=sample\n#\tsecond comment\n\n   \n" + "".join(
    f"chr{1 + (i * 7) % 4}\t{1000 + 37 * i}\t{1400 + 41 * i}\tBC{(i * 11) % 13:02d}-1\t{1 + i % 3}\n" for i in range(200)))


def packing_matrix():
    from pycistopic_b200.synth import make_binary_matrix
    return make_binary_matrix(n_cells=40, n_regions=200, density=0.1, n_true_topics=3, seed=3)


def count_matrix(seed=7, n_regions=120, n_cells=30):
    rng = np.random.default_rng(seed)
    dense = (rng.random((n_regions, n_cells)) < 0.2) * rng.integers(1, 4, (n_regions, n_cells))
    regions = [f"chr{1 + i % 3}:{1000 * i}-{1000 * i + 400 + i}" for i in range(n_regions)]
    cells = [f"CELL{j:03d}-1" for j in range(n_cells)]
    return sp.csr_matrix(dense.astype(np.int32)), cells, regions


def cell_metadata(cells):
    rng = np.random.default_rng(8)
    bare = pd.DataFrame({"Unique_nr_frag": rng.integers(1, 1000, len(cells)), "barcode": cells}, index=cells)
    tagged = bare.copy()
    tagged.index = [c + "___s1" for c in cells]
    overwrite = pd.DataFrame({"cisTopic_nr_frag": np.arange(len(cells)), "extra": "x"}, index=tagged.index)
    return {"bare": bare, "tagged": tagged, "partial": bare.iloc[5:].copy(), "overwrite": overwrite, "reversed": bare.iloc[::-1]}


def diff_inputs():
    rng = np.random.default_rng(41)
    R, C = 120, 60
    cells, regions = [f"c{j}___s" for j in range(C)], [f"chr2:{10 * i}-{10 * i + 5}" for i in range(R)]
    kind = np.array(["A", "B", "C"])[np.arange(C) % 3]
    mtx = rng.poisson(3.0, (R, C)).astype(np.float64)
    for t, letter in enumerate("ABC"):
        mtx[40 * t:40 * t + 15, kind == letter] *= 4
    cell_data = pd.DataFrame({"celltype": kind, "other": 1}, index=cells)
    cell_data.loc[cells[7], "celltype"] = np.nan
    return mtx, cells, regions, cell_data


def diff_kwargs(case, regions):
    kw = dict(case)
    if isinstance(kw.get("var_features"), slice):
        kw["var_features"] = regions[kw["var_features"]]
    return kw


def imputed_pairs():
    """(feature lists, cell lists, matrices) of the merge cases: overlapping, disjoint, identical sets, three objects."""
    regions = [f"chr1:{100 * i}-{100 * i + 50}" for i in range(90)]
    cells = [f"c{j}___s" for j in range(25)]
    sets = [(regions[:40], regions[20:60]), (regions[:20], regions[20:30]), (regions[:30], regions[:30][::-1]),
            (regions[:40], regions[30:50], regions[45:80])]
    out = []
    for group in sets:
        objs = []
        for i, feats in enumerate(group):
            cl = [f"s{i}_{c}" for c in cells[:4 + i]]
            m = np.random.default_rng(10 + i).integers(0, 3, (len(feats), len(cl))).astype(np.int32)
            objs.append((list(feats), cl, m))
        out.append(objs)
    return out


# ---- DataFrames <-> flat arrays (npz without pickles) ------------------------------------------------------------------
def pack_frame(prefix, frame, out):
    out[prefix + "/index"] = np.array([str(x) for x in frame.index])
    out[prefix + "/columns"] = np.array([str(x) for x in frame.columns])
    out[prefix + "/dtypes"] = np.array([str(t) for t in frame.dtypes])
    for j in range(frame.shape[1]):
        col = frame.iloc[:, j]
        if pd.api.types.is_numeric_dtype(col.dtype) and not pd.api.types.is_bool_dtype(col.dtype):
            out[f"{prefix}/c{j}"] = col.to_numpy()
        else:  # strings, booleans, mixed objects: their printed form
            out[f"{prefix}/c{j}"] = np.array([repr(x) for x in col])


def assert_frame_is(prefix, frame, gold, rtol=1e-12):
    mine = {}
    pack_frame(prefix, frame, mine)
    for key, value in mine.items():
        want = gold[key]
        if value.dtype.kind in "fc":
            np.testing.assert_allclose(value, want, rtol=rtol, atol=0, equal_nan=True, err_msg=key)
        else:
            assert value.dtype.kind == want.dtype.kind and np.array_equal(value, want), key
    assert sum(k.startswith(prefix + "/") for k in gold.files) == len(mine), prefix
