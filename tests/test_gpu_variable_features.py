"""GPU parity tests of find_highly_variable_features (reference diff_features.py:577-714): the reduction over the
matrix (np.mean / np.var with dtype=float32, :636-639) on the device -- dense (``cgs_row_mean_var``) and straight
from the factors (``cgs_impute_row_stats``) -- against tests/golden/variable_features.npz, produced by the
reference's own function run from the tree."""
import os

import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# float32 results: NumPy sums pairwise in float32, the kernel in fp64 (correctly rounded) -- a few float32 ulps
RTOL = 2e-6


def _golden():
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "variable_features.npz")))
    imputed = g["imputed"]
    norm64 = np.log1p(imputed / (imputed.sum(axis=0) / 10 ** 4))
    return g, {"imputed": imputed, "norm64": norm64, "norm32": norm64.astype(np.float32)}


def _same_selection(got, want, all_features, slack):
    """Identical lists, or -- where a float32 ulp moves a region across a threshold or a bin edge -- at most
    ``slack`` regions of difference."""
    if list(got) == list(want):
        return
    diff = set(got) ^ set(want)
    assert len(diff) <= slack, (len(diff), sorted(diff)[:10])
    order = {f: i for i, f in enumerate(all_features)}
    assert [order[f] for f in got] == sorted(order[f] for f in got)


def test_row_mean_var_and_selection_match_reference_golden():
    from pycistopic_b200 import diff_features as pdf
    g, mats = _golden()
    for name, mat in mats.items():
        mean, var = pdf.row_mean_var(mat)
        assert mean.dtype == np.float32 and var.dtype == np.float32
        np.testing.assert_allclose(mean, g[f"{name}_mean"], rtol=RTOL)
        np.testing.assert_allclose(var, g[f"{name}_var"], rtol=10 * RTOL, atol=1e-12)
        # ... and against the same definition in float64
        m32 = np.mean(mat.astype(np.float32), axis=1, dtype=np.float64).astype(np.float32)
        np.testing.assert_allclose(mean, m32, rtol=2e-7)
        feats = [f"r{i}" for i in range(mat.shape[0])]
        frame = pd.DataFrame(mat, index=feats)
        obj = pdf.CistopicImputedFeatures(mat, feats, [f"c{j}" for j in range(mat.shape[1])], "t")
        for inp in (frame, obj, pdf.CistopicImputedFeatures(sp.csr_matrix(mat), feats, obj.cell_names, "t")):
            _same_selection(pdf.find_highly_variable_features(inp, plot=False), g[f"{name}_default"], feats, 4)
            _same_selection(pdf.find_highly_variable_features(inp, n_top_features=100, plot=False), g[f"{name}_top100"], feats, 2)
            _same_selection(pdf.find_highly_variable_features(inp, min_disp=0.5, min_mean=0.05, max_mean=2.0, max_disp=3.0,
                                                              n_bins=10, plot=False), g[f"{name}_window"], feats, 4)


@pytest.mark.parametrize("dtype", [np.int32, np.float32, np.float64])
@pytest.mark.parametrize("R,C", [(1, 1), (3, 7), (70, 2048), (5, 2049), (4, 100_003)])
def test_row_mean_var_shapes(dtype, R, C):
    """Column counts around the 2 048-value step of a CTA, a strided (sliced) input, constant rows (variance 0)."""
    from pycistopic_b200 import diff_features as pdf
    rng = np.random.default_rng(R + C)
    x = (rng.normal(3, 2, (R, C + 3)) * (1000 if dtype == np.int32 else 1)).astype(dtype)
    x[0] = 5
    view = x[:, 1:C + 1]                                     # row stride > n_cols
    mean, var = pdf.row_mean_var(view)
    np.testing.assert_allclose(mean, np.mean(view, axis=1, dtype=np.float32), rtol=RTOL)
    np.testing.assert_allclose(var, np.var(view, axis=1, dtype=np.float32), rtol=20 * RTOL, atol=1e-10)
    assert var[0] == 0 and mean[0] == 5


def test_factored_variable_features_equal_dense_path():
    """normalize_scores(factored, materialize=False) -> find_highly_variable_features: statistics of the normalised
    values from the two factors (the regions x cells matrix never exists) == the dense path of this package."""
    from pycistopic_b200 import diff_features as pdf
    rng = np.random.default_rng(21)
    R, C, K = 6000, 500, 10
    tr = rng.dirichlet(np.full(R, 0.3), K).T.astype(np.float32)
    tr[rng.random(R) < 0.1] = 0
    ct = rng.dirichlet(np.full(K, 0.4), C).T.astype(np.float32)
    regions, cells = [f"r{i}" for i in range(R)], [f"c{i}" for i in range(C)]
    for scale in (10 ** 6, 1.0):
        f = pdf.FactoredImputedFeatures(tr, ct, regions, cells, scale, 1500, "t")
        lazy = pdf.normalize_scores(f, 10 ** 4, materialize=False)
        assert isinstance(lazy, pdf.FactoredNormalizedFeatures)
        dense = pdf.normalize_scores(f, 10 ** 4)
        assert lazy.feature_names == dense.feature_names and 0 < len(dense.feature_names) < R
        m1, v1 = lazy.row_mean_var()
        m2, v2 = pdf.row_mean_var(dense.mtx)
        np.testing.assert_array_equal(m1, m2)
        np.testing.assert_array_equal(v1, v2)
        assert pdf.find_highly_variable_features(lazy, plot=False) == pdf.find_highly_variable_features(dense, plot=False)
        assert pdf.find_highly_variable_features(lazy, n_top_features=300, plot=False) == \
            pdf.find_highly_variable_features(dense, n_top_features=300, plot=False)
        np.testing.assert_array_equal(lazy.to_dense().mtx, dense.mtx)
        # raw imputed values (no normalisation)
        m3, v3 = f.row_mean_var()
        m4, v4 = pdf.row_mean_var(f.to_dense().mtx)
        np.testing.assert_array_equal(m3, m4)
        np.testing.assert_array_equal(v3, v4)
