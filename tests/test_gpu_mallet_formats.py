"""GPU tests of the Mallet-format bridge (reference lda_models.py:472-494, :601-646): a sampler state
written as a Mallet --output-state file and read back onto the device reproduces the counts bit for bit (bookkeeping
is exact for identical z), can be evaluated with the device metric block and sampled further."""
import io

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import lda_oracle as lo

pytestmark = pytest.mark.gpu


def test_state_roundtrip_through_the_device(tmp_path):
    from pycistopic_b200 import lda as plda
    from pycistopic_b200 import lda_models as plm
    from pycistopic_b200 import mallet_formats as mf
    rng = np.random.default_rng(8)
    R, C, K = 3000, 120, 7
    m = sp.csr_matrix((rng.random((R, C)) < 0.04).astype(np.int32))       # regions x cells
    X = sp.csr_matrix(m.T)
    alpha, eta = 50.0 / K, 0.1
    fitted = plda.LDA(n_topics=K, n_iter=30, alpha=alpha, eta=eta, random_state=555, keep_assignments=True).fit(X)
    assert fitted.z_.shape == (m.nnz,) and np.array_equal(np.bincount(fitted.z_, minlength=K), fitted.nz_)
    # (WS_, DS_) is the token order of lda.utils.matrix_to_lists: cell by cell, regions ascending
    Xc = sp.csc_matrix(m); Xc.sort_indices()
    np.testing.assert_array_equal(fitted.WS_, Xc.indices)
    np.testing.assert_array_equal(fitted.DS_, np.repeat(np.arange(C), np.diff(Xc.indptr)))
    # the corpus text lists exactly those tokens
    buf = io.StringIO()
    mf.corpus_to_mallet(m, buf)
    toks = [line.split("\t")[2].split() for line in buf.getvalue().splitlines()]
    np.testing.assert_array_equal(np.concatenate([np.array(t, dtype=np.int64) for t in toks if t]), fitted.WS_)

    path = str(tmp_path / "state.mallet.gz")
    mf.state_from_lda(fitted, path)
    # the reference's reader view of the file: topics x regions assignment counts == nzw
    np.testing.assert_array_equal(mf.load_word_topics(path, K, R), fitted.nzw_.astype(np.float64))
    # ... and back on the device, evaluated as it is
    back = mf.lda_from_state(m, path, K, alpha, eta, n_iter=0, keep_assignments=True)
    np.testing.assert_array_equal(back.z_, fitted.z_)
    np.testing.assert_array_equal(back.nzw_, fitted.nzw_)
    np.testing.assert_array_equal(back.ndz_, fitted.ndz_)
    np.testing.assert_array_equal(back.nz_, fitted.nz_)
    np.testing.assert_array_equal(back.topic_word_, fitted.topic_word_)
    assert back.final_loglikelihood_ == fitted.final_loglikelihood_
    # counts agree with the oracle's bookkeeping for the same z
    nzw, ndz, nz = lo.counts_from_z(fitted.WS_, fitted.DS_, fitted.z_, C, R, K)
    np.testing.assert_array_equal(back.nzw_, nzw)
    np.testing.assert_array_equal(back.ndz_, ndz)
    np.testing.assert_array_equal(back.nz_, nz)
    # packed like any other model, metric block from the device
    cells, regions = [f"c{i}" for i in range(C)], [f"r{i}" for i in range(R)]
    ev = mf.lda_from_state(m, path, K, alpha, eta, n_iter=0, exports="frames",
                           device_metrics={"alpha": 50.0, "eta": eta, "top_n": 20})
    model = plm.build_cistopic_model(ev, m, K, cells, regions, 0, 555, 50.0, True, eta, False, 5, 0.0)
    host = plm.build_cistopic_model(back, m, K, cells, regions, 0, 555, 50.0, True, eta, False, 5, 0.0)
    got, want = model.metrics.to_numpy().astype(float)[0], host.metrics.to_numpy().astype(float)[0]
    np.testing.assert_allclose(got[[0, 1, 3]], want[[0, 1, 3]], rtol=1e-9)     # Arun, Cao-Juan, loglikelihood
    # Mimno takes the top 20 regions per topic: with counts this small most of them tie, and the device and
    # NumPy's argsort pick different members of the tie (tests/test_gpu_metrics.py covers the tie-free case exactly)
    np.testing.assert_allclose(got[2], want[2], rtol=0.05)
    assert model.topic_ass["Assignments"].sum() == m.nnz
    # sampled further from the imported state: the chain moves on and the likelihood does not collapse
    more = mf.lda_from_state(m, path, K, alpha, eta, n_iter=20, random_state=1)
    assert more.nz_.sum() == m.nnz and not np.array_equal(more.nzw_, fitted.nzw_)
    assert more.final_loglikelihood_ > fitted.loglikelihoods_[0]
    # a state of another matrix is refused
    other = m.copy().tolil(); other[0, 0] = 1 - other[0, 0]
    with pytest.raises(ValueError, match="tokens"):
        mf.lda_from_state(sp.csr_matrix(other), path, K, alpha, eta)
    with pytest.raises(ValueError, match="outside"):
        mf.lda_from_state(m, path, K - 3, alpha, eta)
