"""GPU parity tests of the per-cell rankings: ``cgs_rank_columns`` behind the mirror of
``CistopicImputedFeatures.make_rankings`` (reference diff_features.py:274-337).

Three anchors: (1) tests/golden/rankings.npz -- the reference's own method run from the tree -- bit for bit on every
column without ties, and as the same sorted score sequence on columns with ties (any tie order is a valid draw);
(2) oracle/rankings_oracle.py::rank_columns, the CPU restatement of the device algorithm, bit for bit INCLUDING the
ties; (3) size-independent properties at sizes the oracle does not reach (every column a permutation, scores
non-increasing along the ranking, determinism in (seed, cell index), independence of the column blocking)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import rankings_oracle as ro

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _check_valid(x, rank):
    n = x.shape[0]
    for c in range(x.shape[1]):
        inv = np.empty(n, np.int64)
        inv[rank[:, c]] = np.arange(n)                      # fails loudly if not a permutation
        assert np.array_equal(np.sort(rank[:, c]), np.arange(n))
        col = x[inv, c].astype(np.float64)
        nan = np.isnan(col)
        assert not nan[: n - nan.sum()].any()               # NaN last
        assert (np.diff(col[~nan]) <= 0).all()              # descending scores


def test_make_rankings_matches_reference_golden():
    from pycistopic_b200 import diff_features as pdf
    g = np.load(os.path.join(ROOT, "tests", "golden", "rankings.npz"))
    for xk, rk, seed in (("x_int", "rank_int", 123), ("x_float", "rank_float", 7)):
        x, ref = g[xk], g[rk]
        obj = pdf.CistopicImputedFeatures(x, [f"r{i}" for i in range(x.shape[0])], [f"c{i}" for i in range(x.shape[1])], "t")
        got = obj.make_rankings(seed=seed)
        assert isinstance(got, pdf.CistopicImputedFeatures) and got.mtx.dtype == np.int32
        assert got.feature_names == obj.feature_names and got.cell_names == obj.cell_names
        _check_valid(x, got.mtx)
        for c in range(x.shape[1]):
            if len(np.unique(x[:, c])) == x.shape[0]:        # no ties: the ranking is unique
                np.testing.assert_array_equal(got.mtx[:, c], ref[:, c])
            else:                                           # ties: same scores at every rank
                np.testing.assert_array_equal(x[np.argsort(got.mtx[:, c]), c], x[np.argsort(ref[:, c]), c])
        np.testing.assert_array_equal(got.mtx, ro.rank_columns(x, seed=seed))
        # the reference's input type (SciPy sparse) gives the same matrix
        sparse_obj = pdf.CistopicImputedFeatures(sp.csr_matrix(x), obj.feature_names, obj.cell_names, "t")
        np.testing.assert_array_equal(sparse_obj.make_rankings(seed=seed).mtx, got.mtx)


@pytest.mark.parametrize("dtype", [np.int32, np.float32, np.float64])
@pytest.mark.parametrize("R,C", [(1, 3), (2, 2), (31, 5), (4096, 3), (4097, 4), (70001, 3)])
def test_rank_columns_matches_oracle_bit_for_bit(dtype, R, C):
    """Ties, negative values, constant and two-valued columns, NaN / +-0 / +-inf for floats, column lengths around
    the 4 096-item tile and beyond one 16-bit Feistel domain."""
    from pycistopic_b200 import diff_features as pdf
    rng = np.random.default_rng(R * 31 + C)
    if dtype == np.int32:
        x = rng.integers(-5, 300, (R, C)).astype(np.int32)
        if C > 2:
            x[:, 2] = rng.integers(-2 ** 31, 2 ** 31 - 1, R)     # all four digits vary
    else:
        x = np.round(rng.normal(0, 3, (R, C)), 1).astype(dtype)
        if R > 8:
            x[1, 0], x[2, 0], x[3, 0], x[4, 0], x[5, 0] = np.nan, 0.0, -0.0, np.inf, -np.inf
            x[6, 0] = -np.nan
    x[:, 1] = 7                                                   # constant column
    x[R // 2:, C - 1] = 0                                         # half the column tied
    got = pdf.rank_columns(x, seed=99, first_col=5)
    _check_valid(x, got)
    np.testing.assert_array_equal(got, ro.rank_columns(x, seed=99, first_col=5))


def test_rank_columns_blocking_strides_and_seeds():
    from pycistopic_b200 import diff_features as pdf
    rng = np.random.default_rng(5)
    x = rng.integers(0, 4, (3000, 40)).astype(np.int32)
    whole = pdf.rank_columns(x, seed=1)
    # a column block is keyed by its cell index, not by its position in the call
    np.testing.assert_array_equal(pdf.rank_columns(x[:, 10:25], seed=1, first_col=10), whole[:, 10:25])
    # other seed, other tie order; same seed, same result
    other = pdf.rank_columns(x, seed=2)
    assert (other != whole).mean() > 0.5
    np.testing.assert_array_equal(pdf.rank_columns(x, seed=1), whole)
    # tie order carries no trace of the row order (the reference permutes first for exactly this reason)
    zeros = np.where(x[:, 0] == 0)[0]
    assert abs(np.corrcoef(zeros, whole[zeros, 0])[0, 1]) < 0.1
    # two cells with identical scores get different tie orders
    x[:, 1] = x[:, 0]
    both = pdf.rank_columns(x[:, :2], seed=1)
    assert (both[:, 0] != both[:, 1]).mean() > 0.5
    # Fortran-ordered and other dtypes are converted, not misread
    np.testing.assert_array_equal(pdf.rank_columns(np.asfortranarray(x), seed=1), pdf.rank_columns(x, seed=1))
    np.testing.assert_array_equal(pdf.rank_columns(x.astype(np.int64), seed=1), pdf.rank_columns(x.astype(np.float64), seed=1))
    with pytest.raises(ValueError, match="2-D"):
        pdf.rank_columns(np.arange(5))
    assert pdf.rank_columns(np.empty((0, 3), np.int32)).shape == (0, 3)


def test_rank_columns_imputed_shape_properties():
    """An imputed-accessibility-shaped block (trunc(1e6 p): small integers, many ties), 200 000 regions: every
    column is a permutation with non-increasing scores; agrees with the oracle on the columns it is run on."""
    from pycistopic_b200 import diff_features as pdf
    rng = np.random.default_rng(11)
    R, C, K = 200_000, 64, 20
    tr = rng.dirichlet(np.full(R, 0.05), K).T.astype(np.float32)          # regions x topics
    ct = rng.dirichlet(np.full(K, 0.3), C).T.astype(np.float32)           # topics x cells
    x = (tr @ ct * 1e6).astype(np.int32)
    got = pdf.rank_columns(x, seed=123)
    _check_valid(x, got)
    np.testing.assert_array_equal(got[:, :2], ro.rank_columns(x[:, :2], seed=123))


@pytest.mark.parametrize("scale", [10 ** 6, 1.0])
def test_factored_make_rankings_equals_dense_path(scale):
    """impute_accessibility(..., materialize=False).make_rankings() == impute_accessibility(...).make_rankings():
    the product is recomputed per block of cells on the device (cgs_impute_rankings), all-zero regions dropped first
    as the reference does (diff_features.py:468-486), and the tie order depends on the cell index only -- so any
    cell blocking gives the matrix of the dense path."""
    from pycistopic_b200 import diff_features as pdf
    rng = np.random.default_rng(3)
    R, C, K = 5000, 333, 12
    tr = rng.dirichlet(np.full(R, 0.05), K).T.astype(np.float32)
    tr[rng.random(R) < 0.2] = 0                                       # regions the row filter drops
    ct = rng.dirichlet(np.full(K, 0.3), C).T.astype(np.float32)
    regions, cells = [f"r{i}" for i in range(R)], [f"c{i}" for i in range(C)]
    f = pdf.FactoredImputedFeatures(tr, ct, regions, cells, scale, 1000, "t")
    dense = f.to_dense()
    assert 0 < len(dense.feature_names) < R
    want = dense.make_rankings(seed=5)
    for block in (0, 100, 32):
        got = f.make_rankings(seed=5, cell_block=block)
        assert got.feature_names == dense.feature_names and got.cell_names == cells
        np.testing.assert_array_equal(got.mtx, want.mtx)
    _check_valid(dense.mtx, want.mtx)
