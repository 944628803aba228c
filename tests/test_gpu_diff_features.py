"""GPU parity tests of the differential-feature kernels: ``cgs_ranksum_rows`` behind the mirror
of ``markers`` / ``get_wilcox_test_pvalues`` / ``get_log2_fc`` (reference diff_features.py:839-1120) against the
golden vectors produced by the reference's own functions and against the scipy-based oracle."""
import os

import numpy as np
import pandas as pd
import pytest

from oracle import diff_oracle as do

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
