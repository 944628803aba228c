"""GPU parity tests of the device-side metric block: the ingredients the kernels of
``metrics_kernels.cuh`` return through the C ABI against the host restatement (``pycistopic_b200.metrics``,
itself checked against the oracle's tmtoolkit restatement and the reference's own ``utils.loglikelihood``
golden vectors in the CPU suite) and against the oracle directly."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import lda_oracle as lo
from pycistopic_b200.synth import make_binary_matrix

pytestmark = pytest.mark.gpu


def _host_codf(m, top):
    M = sp.csr_matrix(m)
    out = np.empty((top.shape[0], top.shape[1], top.shape[1]), np.int64)
    for t in range(top.shape[0]):
        sub = M[top[t]]
        sub = sp.csr_matrix((np.ones(sub.nnz, np.int64), sub.indices, sub.indptr), shape=sub.shape)
        out[t] = (sub @ sub.T).toarray()
    return out


def _check_state(m, model, K, alpha, eta, alpha_raw, eta_raw, top_n=20):
    from pycistopic_b200 import metrics as pm
    X = sp.csr_matrix(m.T)
    nzw, ndz, nz = model.counts()
    tw, dt = model.distributions()
    L = np.asarray(X.sum(axis=1)).ravel().astype(np.float64)
    np.testing.assert_array_equal(model.corpus.cell_sizes(), L.astype(np.int64))
    # Gram of topic_word: exact integer Gram on the device, fp64 finish
    gram = model.topic_gram()
    np.testing.assert_allclose(gram, tw @ tw.T, rtol=1e-12, atol=0)
    # cell mixture
    mix, sq = model.cell_mixture()
    np.testing.assert_allclose(mix, L @ dt, rtol=1e-12)
    assert sq == pytest.approx(float(L @ L), rel=1e-14)
    # top regions: count descending, ties by region index descending == stable argsort read backwards
    top, cnt = model.top_regions(top_n, counts=True)
    ref_top = pm.top_regions_stable(tw, top_n)
    np.testing.assert_array_equal(top, ref_top)
    np.testing.assert_array_equal(cnt, np.take_along_axis(nzw, ref_top, axis=1))
    # (co-)document frequencies: exact
    codf = model.codocument_frequencies(top)
    np.testing.assert_array_equal(codf, _host_codf(m, top))
    # the bug-compatible log-likelihood metric
    assert model.loglikelihood_compat(alpha_raw, eta_raw) == pytest.approx(
        pm.loglikelihood_compat(nzw, ndz, alpha_raw, eta_raw), rel=1e-12)
    # the finished metrics against the oracle's restatement of tmtoolkit (same top lists)
    doc_len = L.reshape(-1, 1)
    if K > 1:
        assert pm.arun_from_gram(gram, mix, sq, tw) == pytest.approx(lo.metric_arun_2010(tw, dt, doc_len), rel=1e-8)
        assert pm.cao_juan_from_gram(gram) == pytest.approx(lo.metric_cao_juan_2009(tw), rel=1e-10)
    np.testing.assert_allclose(pm.marginal_from_mixture(mix), lo.marginal_topic_distrib(dt, L), rtol=1e-12)
    np.testing.assert_allclose(pm.mimno_from_codf(codf), pm.metric_coherence_mimno_2011(tw, m, top_n=top_n, top=top),
                               rtol=1e-12)
    return tw, top


@pytest.fixture(scope="module")
def c1():
    return make_binary_matrix("C1")


@pytest.mark.parametrize("K", [1, 2, 10, 50, 130, 300, 512])
def test_metric_ingredients_after_sweeps(c1, K):
    from pycistopic_b200 import Corpus, DeviceModel
    m, _, _ = c1
    corpus = Corpus.from_regions_by_cells(m)
    alpha, eta = 50.0 / K, 0.1
    model = DeviceModel(corpus, K, alpha, eta, seed=555)
    model.init()
    model.sweep(15)
    tw, top = _check_state(m, model, K, alpha, eta, 50.0, 0.1)
    # wherever numpy's default argsort agrees with the stable order, the oracle's coherence must be identical
    from pycistopic_b200 import metrics as pm
    ours = pm.mimno_from_codf(model.codocument_frequencies(top))
    default_top = np.argsort(tw, axis=1)[:, :-21:-1]
    same = np.all(default_top == top, axis=1)
    ref = lo.metric_coherence_mimno_2011(tw, sp.csr_matrix(m.T))
    np.testing.assert_allclose(ours[same], ref[same], rtol=1e-10)
    model.close()
    corpus.close()


def test_metric_ingredients_with_ties_and_empty_rows():
    """Initial state z = i mod K on a tiny corpus: many equal counts (ties in the top lists); then a state whose
    LAST topic is empty and a corpus whose last cells are empty -- the cases utils.loglikelihood's overwritten
    loops are sensitive to (utils.py:139-157)."""
    from pycistopic_b200 import Corpus, DeviceModel
    m, _, _ = make_binary_matrix(n_cells=120, n_regions=900, density=0.05, n_true_topics=4, seed=3)
    m = sp.csr_matrix(m)
    # append 3 empty cells (columns) and make region 7 empty
    m = sp.hstack([m, sp.csr_matrix((m.shape[0], 3), dtype=m.dtype)]).tolil()
    m[7, :] = 0
    m = sp.csr_matrix(m)
    m.eliminate_zeros()
    K = 7
    corpus = Corpus.from_regions_by_cells(m)
    model = DeviceModel(corpus, K, 50.0 / K, 0.1, seed=1)
    model.init()
    _check_state(m, model, K, 50.0 / K, 0.1, 50.0, 0.1)
    z = model.get_z()
    z[z >= K - 2] = 0  # the last two topics are empty now
    model.set_z(z)
    _check_state(m, model, K, 50.0 / K, 0.1, 50.0, 0.1)
    _check_state(m, model, K, 50.0 / K, 0.1, 3.0, 0.01, top_n=5)
    tiny = Corpus.from_regions_by_cells(sp.csr_matrix(np.ones((10, 4), np.int32)))
    mm = DeviceModel(tiny, 2, 1.0, 0.1)
    mm.init()
    with pytest.raises(ValueError, match="larger than the vocabulary size"):
        mm.top_regions(20)
    # z = i mod 2 over 4 cells x 10 regions: topic 0 holds the even regions (count 4), the odd ones tie at 0
    np.testing.assert_array_equal(mm.top_regions(10), [[8, 6, 4, 2, 0, 9, 7, 5, 3, 1], [9, 7, 5, 3, 1, 8, 6, 4, 2, 0]])
    mm.close()
    tiny.close()
    model.close()
    corpus.close()


def test_run_cgs_model_device_metrics_match_host_path(c1):
    """The CistopicLDAModel built from the device ingredients equals the one built from the exported host
    arrays with the reference's formulas (lda_models.py:291-344) for the SAME fitted state."""
    from pycistopic_b200 import LDA
    from pycistopic_b200 import metrics as pm
    from pycistopic_b200.lda_models import build_cistopic_model
    m, cells, regions = c1
    X = sp.csr_matrix(m.T)
    K = 10
    fitted = LDA(K, n_iter=40, alpha=50.0 / K, eta=0.1, random_state=555, refresh=40, exports="all",
                 device_metrics={"alpha": 50.0, "eta": 0.1, "top_n": 20}).fit(X)
    args = (m, K, cells, regions, 40, 555, 50.0, True, 0.1, False, 5, 0.0)
    dev = build_cistopic_model(fitted, *args)
    ing = fitted.metric_ingredients_
    del fitted.metric_ingredients_
    host = build_cistopic_model(fitted, *args)
    for col in ("Arun_2010", "Cao_Juan_2009", "loglikelihood"):
        assert dev.metrics.loc["Metric", col] == pytest.approx(host.metrics.loc["Metric", col], rel=1e-9)
    np.testing.assert_allclose(dev.marg_topic["Marg_Topic"].values, host.marg_topic["Marg_Topic"].values, rtol=1e-12)
    # coherence: identical for the same top-region lists; the host path's default argsort may order tied counts
    # differently (NumPy leaves that unspecified), so the reference formula is evaluated on the device's lists ...
    np.testing.assert_allclose(dev.coherence["Mimno_2011"].values,
                               pm.metric_coherence_mimno_2011(fitted.topic_word_, m, top=ing["top_regions"]), rtol=1e-10)
    np.testing.assert_array_equal(ing["top_regions"], pm.top_regions_stable(fitted.topic_word_, 20))
    # ... and wherever the default argsort happens to agree, the two packed models agree as they are
    same = np.all(np.argsort(fitted.topic_word_, axis=1)[:, :-21:-1] == ing["top_regions"], axis=1)
    np.testing.assert_allclose(dev.coherence["Mimno_2011"].values[same], host.coherence["Mimno_2011"].values[same],
                               rtol=1e-10)
    np.testing.assert_array_equal(dev.topic_region.values, host.topic_region.values)
    np.testing.assert_array_equal(dev.cell_topic.values, host.cell_topic.values)
    assert list(dev.topic_ass["Assignments"]) == list(host.topic_ass["Assignments"])
