"""GPU parity tests of the differential-feature kernels: ``cgs_ranksum_rows`` behind the mirror
of ``markers`` / ``get_wilcox_test_pvalues`` / ``get_log2_fc`` (reference diff_features.py:839-1120) against the
golden vectors produced by the reference's own functions and against the scipy-based oracle."""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import diff_oracle as do
from oracle import lda_oracle as lo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ranksum_matches_reference_golden():
    from pycistopic_b200 import diff_features as pdf
    g = np.load(os.path.join(ROOT, "tests", "golden", "diff_features.npz"))
    for i in range(int(g["n"])):
        fg, bg = g[f"fg{i}"], g[f"bg{i}"]
        stat, p, l2fc = pdf.ranksum_rows(fg, bg)
        np.testing.assert_allclose(p, g[f"p{i}"], rtol=1e-11, atol=1e-300)
        f32 = fg.dtype == np.float32
        np.testing.assert_allclose(l2fc, g[f"l2fc{i}"], rtol=1e-6 if f32 else 1e-12, atol=2e-6 if f32 else 0,
                                   equal_nan=True)
        assert pdf.get_wilcox_test_pvalues(fg, bg) == p.tolist()
        np.testing.assert_array_equal(pdf.get_log2_fc(fg, bg), l2fc)
        np.testing.assert_allclose(pdf.p_adjust_bh(p), g[f"padj{i}"], rtol=1e-11)


def test_markers_matches_reference_golden():
    """markers() (diff_features.py:839-965) on a DataFrame: same features, same order, same numbers."""
    from pycistopic_b200 import diff_features as pdf
    g = np.load(os.path.join(ROOT, "tests", "golden", "diff_features.npz"))
    mat = g["m_mat"]
    cells = [f"cell{j}" for j in range(mat.shape[1])]
    feats = [f"chr1:{j * 500}-{j * 500 + 500}" for j in range(mat.shape[0])]
    df = pd.DataFrame(mat, index=feats, columns=cells)
    groups = [[cells[j] for j in g["m_fg"]], [cells[j] for j in g["m_bg"]]]
    res = pdf.markers(df, groups, "A_VS_B", adjpval_thr=0.05, log2fc_thr=np.log2(1.5), n_cpu=1)
    assert list(res.columns) == ["Log2FC", "Adjusted_pval", "Contrast"] and set(res["Contrast"]) == {"A_VS_B"}
    assert [feats.index(f) for f in res.index] == g["m_index"].tolist()
    np.testing.assert_allclose(res["Log2FC"].to_numpy(), g["m_log2fc"], rtol=1e-12)
    np.testing.assert_allclose(res["Adjusted_pval"].to_numpy(), g["m_padj"], rtol=1e-10)
    # the imputed-features object gives the same answer as the DataFrame
    obj = pdf.CistopicImputedFeatures(mat, feats, cells, "t")
    res2 = pdf.markers(obj, groups, "A_VS_B", adjpval_thr=0.05, log2fc_thr=np.log2(1.5))
    assert res2.index.tolist() == res.index.tolist()


@pytest.mark.parametrize("dtype", [np.int32, np.float32, np.float64])
@pytest.mark.parametrize("R,C,n_fg,n_bg", [(37, 300, 40, 200), (5, 3000, 2500, 400), (3, 90000, 40000, 45000),
                                           (8, 64, 1, 63), (6, 70000, 33000, 37000)])
def test_ranksum_one_matrix_two_index_lists(dtype, R, C, n_fg, n_bg):
    """One matrix + two (unsorted) column lists, as markers() uses it: ties, either group larger, groups beyond one
    shared-memory piece (32 768 values for 4-byte, 16 384 for 8-byte keys), negative column indices."""
    from pycistopic_b200 import diff_features as pdf
    rng = np.random.default_rng(R * 1000 + C)
    if dtype == np.int32:
        x = rng.integers(-3, 40, (R, C)).astype(dtype)
    else:
        x = np.round(rng.normal(5, 2, (R, C)), 2).astype(dtype)  # rounding produces ties
    x[0, : C // 2] = 0
    perm = rng.permutation(C)
    fg_cols, bg_cols = perm[:n_fg].copy(), perm[n_fg:n_fg + n_bg].copy()
    fg_cols[0] -= C  # a negative index addresses the same column, like NumPy
    stat, p, l2fc = pdf.ranksum_rows(x, None, fg_cols, bg_cols)
    fg, bg = x[:, fg_cols], x[:, bg_cols]
    # scipy (1.18) ranks float32 input IN float32, so its rank sum is already inexact at 10^4 values
    # (6e-5 relative at 85 000); the kernel counts in integers, i.e. it returns what scipy gives for the same
    # values in float64 -- compare there, and loosely with the float32 call
    ostat, op = do.ranksums_rows(fg.astype(np.float64), bg.astype(np.float64))
    if dtype == np.float32:
        np.testing.assert_allclose(stat, do.ranksums_rows(fg, bg)[0], rtol=2e-3, atol=1e-3)
    np.testing.assert_allclose(stat, ostat, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(p, op, rtol=1e-10, atol=1e-300)
    np.testing.assert_allclose(l2fc, do.log2_fc(fg, bg), rtol=1e-9, equal_nan=True)
    # two pre-subset matrices (the reference's get_wilcox_test_pvalues call shape) agree bit for bit
    stat2, p2, l2fc2 = pdf.ranksum_rows(np.ascontiguousarray(fg), np.ascontiguousarray(bg))
    np.testing.assert_array_equal(stat2, stat)
    np.testing.assert_array_equal(p2, p)
    np.testing.assert_allclose(l2fc2, l2fc, rtol=1e-9, atol=1e-12, equal_nan=True)  # other summation order


def test_ranksum_counting_path_and_its_fallback():
    """int32 rows of at least 2048 values take the shared-memory counting path when max - min < 8192 and the sort +
    search path otherwise; the decision is per row, so one launch mixes both.  Rows at the boundary (range 8191 /
    8192), a constant row, and rows spanning INT32_MIN..INT32_MAX must all agree with scipy, and with the launch
    that has the counting path switched off (same integer U, so bit for bit)."""
    import os
    from pycistopic_b200 import diff_features as pdf
    rng = np.random.default_rng(77)
    R, C, n_fg = 9, 6000, 2500
    x = rng.integers(0, 50, (R, C)).astype(np.int32)
    x[1] = rng.integers(-4000, 4192, C); x[1, 0], x[1, 1] = -4000, 4191       # range 8191: counted
    x[2] = rng.integers(-4000, 4193, C); x[2, 0], x[2, 1] = -4000, 4192       # range 8192: searched
    x[3] = 7                                                                  # constant row
    x[4] = rng.integers(-2**31, 2**31 - 1, C); x[4, 0], x[4, 1] = -2**31, 2**31 - 1
    x[5] = rng.integers(2**31 - 9000, 2**31 - 1, C, dtype=np.int64)            # near INT32_MAX, range > 8192
    x[6] = rng.integers(2**31 - 8000, 2**31 - 1, C, dtype=np.int64); x[6, 5] = 2**31 - 1  # near INT32_MAX, counted
    x[7] = (rng.random(C) < 0.01) * 1000000                                    # sparse, wide
    perm = rng.permutation(C)
    fg_cols, bg_cols = perm[:n_fg], perm[n_fg:]
    stat, p, l2fc = pdf.ranksum_rows(x, None, fg_cols, bg_cols)
    ostat, op = do.ranksums_rows(x[:, fg_cols].astype(np.float64), x[:, bg_cols].astype(np.float64))
    np.testing.assert_allclose(stat, ostat, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(p, op, rtol=1e-10, atol=1e-300)
    os.environ["CGS_RANKSUM_NO_HIST"] = "1"
    try:
        stat2, p2, l2fc2 = pdf.ranksum_rows(x, None, fg_cols, bg_cols)
    finally:
        del os.environ["CGS_RANKSUM_NO_HIST"]
    np.testing.assert_array_equal(stat2, stat)
    np.testing.assert_array_equal(p2, p)
    np.testing.assert_allclose(l2fc2, l2fc, rtol=1e-9, atol=1e-12, equal_nan=True)


def test_ranksum_errors_and_empty_groups():
    from pycistopic_b200 import diff_features as pdf
    x = np.arange(12, dtype=np.int32).reshape(3, 4)
    with pytest.raises(ValueError, match="different first dimension"):
        pdf.ranksum_rows(x, x[:2])
    with pytest.raises(IndexError, match="out of bounds"):
        pdf.ranksum_rows(x, None, [0, 4], [1])
    stat, p, l2fc = pdf.ranksum_rows(x, None, [], [1, 2])
    assert np.isnan(p).all() and np.isnan(l2fc).all()


def test_find_diff_features_one_vs_rest():
    """find_diff_features (diff_features.py:716-836) with the default one-vs-rest contrasts on a planted signal."""
    from pycistopic_b200 import diff_features as pdf
    rng = np.random.default_rng(5)
    R, C = 200, 150
    mat = rng.integers(0, 20, (R, C)).astype(np.int32)
    labels = np.array(["A"] * 50 + ["B"] * 60 + ["C"] * 40)
    mat[:10, :50] += 40      # markers of A
    mat[10:25, 50:110] += 40  # markers of B
    mat[199, :] = 0           # an all-zero feature is dropped by subset()
    cells = [f"c{j}" for j in range(C)]
    feats = [f"r{j}" for j in range(R)]

    class Obj:
        cell_data = pd.DataFrame({"celltype": labels}, index=cells)

    imputed = pdf.CistopicImputedFeatures(mat, feats, cells, "t")
    res = pdf.find_diff_features(Obj(), imputed, "celltype", n_cpu=4, _temp_dir="/tmp/x")
    assert list(res) == ["A", "B", "C"]
    assert sorted(res["A"].index) == sorted(feats[:10]) and sorted(res["B"].index) == sorted(feats[10:25])
    assert len(res["C"]) == 0
    # against the oracle, contrast A
    fg, bg = mat[:199, :50], mat[:199, 50:]
    _, op = do.ranksums_rows(fg, bg)
    padj = do.p_adjust_bh(op)
    l2 = do.log2_fc(fg, bg)
    keep = np.flatnonzero((padj <= 0.05) & (l2 >= np.log2(1.5)))
    assert sorted(res["A"].index) == sorted(feats[i] for i in keep)
    np.testing.assert_allclose(res["A"].loc[[feats[i] for i in keep], "Adjusted_pval"].to_numpy(), padj[keep], rtol=1e-10)


# ---- fused consumers of the imputed matrix (cgs_impute_normalized / cgs_impute_ranksum) --------------------------
def _planted_factors(R=3000, K=12, C=700, seed=9):
    rng = np.random.default_rng(seed)
    a = rng.dirichlet(np.full(R, 0.1), size=K).T.astype(np.float32)          # regions x topics
    groups = np.repeat(np.arange(3), [250, 250, 200])
    conc = np.full((3, K), 0.3)
    conc[0, :4] = 6.0
    conc[1, 4:8] = 6.0
    conc[2, 8:] = 6.0
    b = np.stack([rng.dirichlet(conc[g]) for g in groups]).T.astype(np.float32)  # topics x cells
    a[40:55] = 0  # all-zero regions: dropped by impute_accessibility
    cells = [f"c{j}" for j in range(C)]
    regions = [f"chr1:{j * 500}-{j * 500 + 500}" for j in range(R)]
    return a, b, cells, regions, np.array(["A", "B", "C"])[groups]


class _Model:
    def __init__(self, a, b, cells, regions):
        topics = [f"Topic{i + 1}" for i in range(a.shape[1])]
        self.cell_topic = pd.DataFrame(b.astype(np.float64), index=topics, columns=cells)
        self.topic_region = pd.DataFrame(a.astype(np.float64), index=regions, columns=topics)


class _Obj:
    def __init__(self, a, b, cells, regions, labels):
        self.selected_model = _Model(a, b, cells, regions)
        self.cell_names, self.region_names = cells, regions
        self.cell_data = pd.DataFrame({"celltype": labels}, index=cells)


def test_factored_normalize_matches_reference_golden():
    """normalize_scores(impute_accessibility(..., materialize=False)) == the reference's two functions chained
    (golden vectors of calculate_imputed_accessibility + calculate_normalized_scores, diff_features.py:398-548)."""
    from pycistopic_b200 import diff_features as pdf
    g = np.load(os.path.join(ROOT, "tests", "golden", "impute.npz"))
    for i in (0, 5):
        a, b = g[f"a{i}"], g[f"b{i}"]
        scale = int(g[f"scale{i}"]) if bool(g[f"scale_is_int{i}"]) else float(g[f"scale{i}"])
        names = [f"r{j}" for j in range(a.shape[0])]
        f = pdf.FactoredImputedFeatures(a, b, names, [f"c{j}" for j in range(b.shape[1])], scale, int(g[f"chunk{i}"]), "t")
        out = pdf.normalize_scores(f, 10 ** 4)
        assert out.feature_names == [names[j] for j in g[f"kept{i}"]]
        # the truncated imputed value may be off by ONE where the fp32 product (3xTF32 here, sgemm there; 1e-5
        # relative, the north-star tolerance) straddles an integer -- the reference golden then differs in those
        # elements by one count (a handful per 10 000); everything else agrees to fp64 round-off of the totals
        # (with scale 10^4 the counts are small, so more elements sit on an integer boundary and one count is a
        # larger step of log1p)
        close = np.isclose(out.mtx, g[f"norm{i}"], rtol=1e-4, atol=1e-9)
        assert close.mean() > 0.98, close.mean()
        # ... and it is exactly normalize_scores of this package's own (dense) imputation
        dense = f.to_dense()
        np.testing.assert_allclose(out.mtx, lo.normalized_scores(dense.mtx.astype(np.float64), 10 ** 4), rtol=1e-13)


def test_factored_consumers_match_dense_path():
    """Every consumer gives the same answer from the factors (chunks recomputed on the device, nothing
    materialised) as from the dense matrix the reference builds."""
    from pycistopic_b200 import diff_features as pdf
    a, b, cells, regions, labels = _planted_factors()
    obj = _Obj(a, b, cells, regions, labels)
    dense = pdf.impute_accessibility(obj, scale_factor=10 ** 6, chunk_size=1000)
    fact = pdf.impute_accessibility(obj, scale_factor=10 ** 6, chunk_size=1000, materialize=False)
    assert isinstance(fact, pdf.FactoredImputedFeatures) and fact.feature_names == dense.feature_names
    assert len(dense.feature_names) == len(regions) - 15
    np.testing.assert_array_equal(fact.to_dense().mtx, dense.mtx)
    # normalize_scores
    nd, nf = pdf.normalize_scores(dense, 10 ** 4), pdf.normalize_scores(fact, 10 ** 4)
    assert nf.feature_names == nd.feature_names and nf.mtx.dtype == nd.mtx.dtype == np.float64
    np.testing.assert_allclose(nf.mtx, nd.mtx, rtol=1e-13)
    # find_diff_features, default one-vs-rest contrasts: identical tables
    rd = pdf.find_diff_features(obj, dense, "celltype")
    rf = pdf.find_diff_features(obj, fact, "celltype")
    assert list(rd) == list(rf) == ["A", "B", "C"]
    for name in rd:
        assert len(rd[name]) > 20
        assert rf[name].index.tolist() == rd[name].index.tolist()
        np.testing.assert_array_equal(rf[name]["Log2FC"].to_numpy(), rd[name]["Log2FC"].to_numpy())
        np.testing.assert_array_equal(rf[name]["Adjusted_pval"].to_numpy(), rd[name]["Adjusted_pval"].to_numpy())
    # a feature subset and an explicit contrast
    var = dense.feature_names[100:900]
    c = [[["A"], ["C"]]]
    rd = pdf.find_diff_features(obj, dense, "celltype", var_features=var, contrasts=c)
    rf = pdf.find_diff_features(obj, fact, "celltype", var_features=var, contrasts=c)
    assert list(rf) == ["A_VS_C"] and rf["A_VS_C"].index.tolist() == rd["A_VS_C"].index.tolist()
    # markers() on one contrast
    groups = [[x for x, l in zip(cells, labels) if l == "B"], [x for x, l in zip(cells, labels) if l != "B"]]
    md, mf = pdf.markers(dense, groups, "B"), pdf.markers(fact, groups, "B")
    assert mf.index.tolist() == md.index.tolist() and len(md) > 5


def test_ranksum_contrasts_one_upload_equals_per_contrast_calls():
    """cgs_ranksum_contrasts (the dense find_diff_features path: one upload, every contrast on the device) gives what
    cgs_ranksum_rows gives contrast by contrast -- also across upload rounds (the matrix below is 1.4 GB: two rounds
    of the double-buffered upload) and for an empty group (NaN)."""
    from pycistopic_b200 import diff_features as pdf
    rng = np.random.default_rng(12)
    for R, C, dtype in ((57, 300, np.float64), (900, 400_000, np.int32)):
        x = rng.integers(0, 30, (R, C)).astype(dtype)
        perm = rng.permutation(C)
        contrasts = [(perm[: C // 10], perm[C // 10:]), (perm[C // 2:], perm[: C // 3]), ([], perm[:5]), (perm[:1], perm[1:3])]
        stat, pval, l2fc = pdf.ranksum_contrasts(x, contrasts)
        assert stat.shape == (4, R)
        for k, (f, b) in enumerate(contrasts):
            s1, p1, l1 = pdf.ranksum_rows(x, None, f, b)
            np.testing.assert_array_equal(stat[k], s1)
            np.testing.assert_array_equal(pval[k], p1)
            np.testing.assert_array_equal(l2fc[k], l1)
        assert np.isnan(pval[2]).all()
    with pytest.raises(IndexError, match="out of bounds"):
        pdf.ranksum_contrasts(x[:3, :10], [([0, 10], [1])])
