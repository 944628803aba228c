"""CPU tests (no GPU): the oracle against golden vectors and against independent restatements,
the host-side logic (metrics, DataFrame packing, model selection, scheduling), and that the
C-ABI library loads and exports every symbol include/cgs_b200.h declares."""
import ctypes
import os
import pickle
import re
import sys

import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

from oracle import lda_oracle as lo
from oracle import ref_functions as rf
from pycistopic_b200 import _lib, lda_models
from pycistopic_b200 import metrics as pm
from pycistopic_b200.synth import SyntheticCistopicObject, make_binary_matrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


# ---- C ABI ------------------------------------------------------------------------------------
def test_abi_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "cgs_b200.h")).read()
    declared = set(re.findall(r"\b(cgs_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    L = _lib.load()
    for name in declared:
        assert hasattr(L, name), f"{name} is declared in include/cgs_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert L.cgs_version() >= 100 and L.cgs_max_topics() == 512


def test_abi_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: the header compiles as C99 (no C++ types, no torch types in the signatures)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "use_header.c"
    src.write_text('#include "cgs_b200.h"\nint main(void) { return cgs_version() == 0; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-fsyntax-only",
                    "-I", os.path.join(ROOT, "include"), str(src)], check=True, capture_output=True)


def build_c_host_example(workdir):
    """gcc examples/host_example.c against the in-tree library; returns the binary (None without gcc)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        return None
    _lib.load()  # builds the library first if the sources are newer
    exe = os.path.join(str(workdir), "host_example")
    lib_dir = os.path.join(ROOT, "pycistopic_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "host_example.c"), "-L", lib_dir, "-lcgs_b200",
                    f"-Wl,-rpath,{lib_dir}", "-lm", "-o", exe], check=True, capture_output=True)
    return exe


def test_c_host_example_builds_and_runs_its_host_part(tmp_path):
    """A host without Python links the C ABI (examples/host_example.c): the file reader runs anywhere; without a GPU
    the program says that the sampler has no CPU path and stops."""
    import subprocess
    exe = build_c_host_example(tmp_path)
    if exe is None:
        pytest.skip("no gcc")
    frag = tmp_path / "f.tsv"
    frag.write_text("chr1\t1\t5\tA\nchr1\t1\t5\tA\nchr2\t3\t9\tB\n")
    res = subprocess.run([exe, str(frag)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "3 records, 4 columns, 2 chromosomes, 2 barcodes" in res.stdout and "1 duplicate records collapsed" in res.stdout
    n = ctypes.c_int(0)
    if _lib.load().cgs_device_count(ctypes.byref(n)) != 0 or n.value < 1:
        assert "no CUDA device" in res.stdout
    else:
        assert res.stdout.rstrip().endswith("OK")
    bad = subprocess.run([exe, str(tmp_path / "absent.tsv")], capture_output=True, text=True, timeout=600)
    assert bad.returncode == 1 and "cannot open" in bad.stderr


def test_no_cpu_fallback_without_device():
    n = ctypes.c_int(0)
    rc = _lib.load().cgs_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a GPU is visible")
    m, cells, regions = make_binary_matrix(n_cells=20, n_regions=100, density=0.1, n_true_topics=2, seed=1)
    with pytest.raises(RuntimeError):
        lda_models.run_cgs_models(SyntheticCistopicObject(m, cells, regions), n_topics=[2], n_iter=2)
    from pycistopic_b200 import LDA
    with pytest.raises(RuntimeError):
        LDA(2, n_iter=1).fit(sp.csr_matrix(m.T))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pycistopic_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert "oracle" not in src.replace("# oracle", ""), f"{f} mentions the oracle"


# ---- oracle vs golden / independent restatements -------------------------------------------------
def test_oracle_known_answer_vectors():
    g = np.load(os.path.join(GOLD, "oracle_kat.npz"))
    X = sp.csr_matrix(g["X"])
    K, alpha, eta, seed = int(g["K"]), float(g["alpha"]), float(g["eta"]), int(g["seed"])
    o = lo.OracleLDA(K, n_iter=0, alpha=alpha, eta=eta, random_state=seed, keep_z=True)
    o._initialize(X)
    assert np.array_equal(o.WS, g["WS"]) and np.array_equal(o.DS, g["DS"])
    rs = np.random.RandomState(seed)
    rands = o._rands.copy()
    assert np.array_equal(o.ZS, g["z"][0])
    for it in range(3):
        rs.shuffle(rands)
        o._sample_topics(rands)
        assert np.array_equal(o.ZS, g["z"][it + 1])
        assert o.loglikelihood() == pytest.approx(float(g["ll"][it + 1]), rel=1e-13)
    assert np.array_equal(o.nzw_, g["nzw"]) and np.array_equal(o.ndz_, g["ndz"]) and np.array_equal(o.nz_, g["nz"])
    Xs = sp.csr_matrix((np.ones(len(g["trace_X_indices"]), np.int32), g["trace_X_indices"], g["trace_X_indptr"]),
                       shape=tuple(g["trace_shape"]))
    o2 = lo.OracleLDA(5, n_iter=20, alpha=10.0, eta=0.1, random_state=555, refresh=1).fit(Xs)
    np.testing.assert_allclose(o2.loglikelihoods_ + [o2.final_loglikelihood_], g["trace_ll"], rtol=1e-13)
    assert np.array_equal(o2.nz_, g["trace_nz"])
    np.testing.assert_allclose(lo.metric_coherence_mimno_2011(o2.topic_word_, Xs), g["trace_mimno"], rtol=1e-12)


def _py_sample_topics(WS, DS, ZS, nzw, ndz, nz, alpha, eta, rands):
    """Line-by-line Python restatement of lda._lda._sample_topics."""
    K, W = nzw.shape
    eta_sum = eta * W
    for i in range(len(WS)):
        w, d, z = WS[i], DS[i], ZS[i]
        nzw[z, w] -= 1; ndz[d, z] -= 1; nz[z] -= 1
        dist = np.cumsum((nzw[:, w] + eta) / (nz + eta_sum) * (ndz[d] + alpha))
        r = rands[i % len(rands)] * dist[-1]
        z_new = int(np.searchsorted(dist, r, side="left"))
        ZS[i] = z_new
        nzw[z_new, w] += 1; ndz[d, z_new] += 1; nz[z_new] += 1


def test_oracle_c_matches_python_restatement():
    rng = np.random.default_rng(0)
    X = sp.csr_matrix((rng.random((15, 40)) < 0.2).astype(np.int32))
    K, alpha, eta, seed = 4, 0.8, 0.05, 3
    o = lo.OracleLDA(K, n_iter=0, alpha=alpha, eta=eta, random_state=seed, keep_z=True)
    o._initialize(X)
    WS, DS = lo.matrix_to_lists_py(X)
    assert np.array_equal(WS, o.WS) and np.array_equal(DS, o.DS)
    ZS = np.arange(len(WS)) % K
    nzw = np.zeros((K, X.shape[1]), np.int64); ndz = np.zeros((X.shape[0], K), np.int64); nz = np.zeros(K, np.int64)
    np.add.at(nzw, (ZS, WS), 1); np.add.at(ndz, (DS, ZS), 1); np.add.at(nz, ZS, 1)
    assert np.array_equal(nzw, o.nzw_) and np.array_equal(ndz, o.ndz_)
    rs = np.random.RandomState(seed)
    rands = o._rands.copy()
    for _ in range(4):
        rs.shuffle(rands)
        o._sample_topics(rands)
        _py_sample_topics(WS, DS, ZS, nzw, ndz, nz, alpha, eta, rands)
        assert np.array_equal(ZS, o.ZS)
    assert np.array_equal(nzw, o.nzw_) and np.array_equal(ndz, o.ndz_) and np.array_equal(nz, o.nz_)
    assert lo.loglikelihood_true(nzw, ndz, nz, alpha, eta) == pytest.approx(
        lo.loglikelihood_true_py(nzw, ndz, nz, alpha, eta), rel=1e-12)
    # count matrices (not binary) expand into repeated tokens
    Xc = sp.csr_matrix(np.array([[2, 0, 1], [0, 3, 0]], dtype=np.int32))
    WS2, DS2 = lo.matrix_to_lists(Xc)
    assert WS2.tolist() == [0, 0, 2, 1, 1, 1] and DS2.tolist() == [0, 0, 0, 1, 1, 1]
    with pytest.raises(ValueError):
        lo.matrix_to_lists(sp.csr_matrix(np.array([[0.5, 0.0]])))
    with pytest.raises(ValueError):
        lo.OracleLDA(2, alpha=0.0)


def test_oracle_adlda_single_part_equals_sequential():
    m, _, _ = make_binary_matrix(n_cells=40, n_regions=300, density=0.08, n_true_topics=3, seed=5)
    X = sp.csr_matrix(m.T)
    a = lo.OracleLDA(4, n_iter=5, alpha=1.0, eta=0.1, random_state=9, refresh=1).fit(X)
    b = lo.OracleLDA(4, n_iter=5, alpha=1.0, eta=0.1, random_state=9, refresh=1, n_parts=1).fit(X)
    assert np.array_equal(a.nzw_, b.nzw_)
    c = lo.OracleLDA(4, n_iter=5, alpha=1.0, eta=0.1, random_state=9, refresh=1, n_parts=3).fit(X)
    assert c.nz_.sum() == a.nz_.sum() == X.nnz and np.array_equal(c.nzw_.sum(axis=1), c.nz_)
    ptr = lo.balanced_cell_partition(np.concatenate([[0], np.cumsum(np.diff(X.indptr))]), 3)
    assert ptr[0] == 0 and ptr[-1] == X.shape[0] and np.all(np.diff(ptr) > 0)


# ---- functions that live in the reference tree: golden fixtures ------------------------------------
def test_loglikelihood_compat_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "loglikelihood_compat.npz"))
    for i in range(int(g["n"])):
        alpha, eta = g[f"par{i}"]
        got = pm.loglikelihood_compat(g[f"nzw{i}"], g[f"ndz{i}"], float(alpha), float(eta))
        assert got == pytest.approx(float(g[f"ll{i}"]), rel=1e-12, abs=1e-9)


@pytest.mark.skipif(not rf.available(), reason="/root/reference not present")
def test_loglikelihood_compat_matches_live_reference():
    f = rf.reference_loglikelihood()
    rng = np.random.default_rng(1)
    for _ in range(20):
        K, W, D = rng.integers(1, 8), rng.integers(2, 30), rng.integers(1, 12)
        nzw = rng.integers(0, 4, (K, W)) * (rng.random((K, W)) < 0.4)
        ndz = rng.integers(0, 4, (D, K)) * (rng.random((D, K)) < 0.5)
        alpha, eta = float(rng.uniform(0.05, 60)), float(rng.uniform(0.01, 1))
        assert pm.loglikelihood_compat(nzw, ndz, alpha, eta) == pytest.approx(f(nzw, ndz, alpha, eta), rel=1e-11, abs=1e-8)


def test_normalized_scores_oracle_matches_reference_golden():
    """The oracle restatement of calculate_normalized_scores against the vectors the reference's own function
    produced (the product path runs on the GPU: tests/test_gpu_parity.py::test_normalize_scores_*)."""
    g = np.load(os.path.join(GOLD, "impute.npz"))
    for i in (0, 5):
        out = g[f"out{i}"]
        np.testing.assert_allclose(lo.normalized_scores(out.astype(np.float64), 10 ** 4), g[f"norm{i}"], rtol=1e-13)
        np.testing.assert_allclose(lo.normalized_scores(out, 10 ** 4), g[f"norm{i}"], rtol=1e-6)


# ---- metrics: product implementation vs oracle restatement -----------------------------------------
@pytest.fixture(scope="module")
def fitted():
    m, cells, regions = make_binary_matrix(n_cells=80, n_regions=1200, density=0.05, n_true_topics=4, seed=21)
    X = sp.csr_matrix(m.T)
    o = lo.OracleLDA(7, n_iter=30, alpha=50 / 7, eta=0.1, random_state=555, refresh=30).fit(X)
    return m, cells, regions, X, o


def test_metrics_match_oracle_restatements(fitted):
    m, _, _, X, o = fitted
    doc_len = np.asarray(X.sum(axis=1)).astype(float)
    assert pm.metric_arun_2010(o.topic_word_, o.doc_topic_, doc_len) == pytest.approx(
        lo.metric_arun_2010(o.topic_word_, o.doc_topic_, doc_len), rel=1e-9)
    assert pm.metric_cao_juan_2009(o.topic_word_) == pytest.approx(lo.metric_cao_juan_2009(o.topic_word_), rel=1e-11)
    np.testing.assert_allclose(pm.metric_coherence_mimno_2011(o.topic_word_, m),
                               lo.metric_coherence_mimno_2011(o.topic_word_, X), rtol=1e-11)
    np.testing.assert_allclose(pm.marginal_topic_distrib(o.doc_topic_, doc_len),
                               lo.marginal_topic_distrib(o.doc_topic_, doc_len), rtol=1e-12)
    with pytest.raises(ValueError):
        pm.metric_coherence_mimno_2011(o.topic_word_[:, :10], m[:10], top_n=20)


def test_model_packing_layout(fitted, tmp_path):
    """CistopicLDAModel layout of lda_models.py:307-388."""
    m, cells, regions, X, o = fitted
    K = 7
    model = lda_models.build_cistopic_model(o, m, K, cells, regions, 30, 555, 50, True, 0.1, False, 5, 1.25,
                                            save_path=str(tmp_path / "models"))
    assert list(model.metrics.columns) == ["Arun_2010", "Cao_Juan_2009", "Mimno_2011", "loglikelihood"]
    assert list(model.metrics.index) == ["Metric"] and model.metrics.shape == (1, 4)
    assert list(model.coherence.columns) == ["Topic", "Mimno_2011"] and model.coherence.shape == (K, 2)
    assert model.coherence["Topic"].tolist() == list(range(1, K + 1))
    assert list(model.marg_topic.columns) == ["Topic", "Marg_Topic"]
    assert model.marg_topic["Marg_Topic"].sum() == pytest.approx(1.0)
    assert list(model.topic_ass.columns) == ["Topic", "Assignments"]
    assert model.topic_ass["Assignments"].tolist() == o.nz_.tolist()
    assert model.cell_topic.shape == (K, len(cells)) and list(model.cell_topic.index) == [f"Topic{i}" for i in range(1, K + 1)]
    assert list(model.cell_topic.columns) == cells
    assert model.topic_region.shape == (len(regions), K) and list(model.topic_region.index) == regions
    np.testing.assert_allclose(model.cell_topic.to_numpy(), o.doc_topic_.T)
    np.testing.assert_allclose(model.topic_region.to_numpy(), o.topic_word_.T)
    assert list(model.parameters.index) == ["package", "n_topics", "n_iter", "random_state", "alpha", "alpha_by_topic",
                                            "eta", "eta_by_topic", "top_topics_coh", "time"]
    assert model.parameters.loc["package", "Parameter"] == "lda" and model.parameters.loc["alpha", "Parameter"] == 50
    assert model.cell_topic_harmony == [] and (model.n_cells, model.n_regions, model.n_topic) == (len(cells), len(regions), K)
    assert str(model).startswith("CistopicLDAModel with 7 topics")
    # metrics: the buggy-compatible loglikelihood is called with the RAW alpha/eta (lda_models.py:305)
    assert model.metrics.loc["Metric", "loglikelihood"] == pytest.approx(pm.loglikelihood_compat(o.nzw_, o.ndz_, 50, 0.1))
    mim = lo.metric_coherence_mimno_2011(o.topic_word_, X)
    assert model.metrics.loc["Metric", "Mimno_2011"] == pytest.approx(np.mean(np.sort(mim)[-5:]))
    with open(tmp_path / "models" / "Topic7.pkl", "rb") as f:
        again = pickle.load(f)
    assert again.n_topic == K and again.topic_region.equals(model.topic_region)
    obj = SyntheticCistopicObject(m, cells, regions)
    obj.add_LDA_model(model)
    assert obj.selected_model is model


@pytest.mark.skipif(not rf.available(), reason="/root/reference is not here")
def test_run_cgs_model_equals_the_reference_function_run_live(tmp_path, monkeypatch):
    """A2 / A14: the reference's ``run_cgs_model`` (lda_models.py:178-396) executed from the tree -- ``lda.LDA`` stood
    in for by the oracle, the four tmtoolkit functions by their restatements -- against the mirror on the same fitted
    model: the seven DataFrames (values, dtypes, index and column names) and the model's attributes are identical, and
    alpha / eta reach the sampler scaled the same way in all four ``*_by_topic`` combinations."""
    made = []

    class Recorder(lo.OracleLDA):
        def __init__(self, **kw):
            made.append(dict(kw))
            super().__init__(**kw)
            made[-1]["fitted"] = self

    # tmtoolkit hands back a (K, 1) matrix for the np.matrix document lengths the reference passes (lda_models.py:337-341)
    fns = dict(metric_arun_2010=lo.metric_arun_2010, metric_cao_juan_2009=lo.metric_cao_juan_2009,
               metric_coherence_mimno_2011=lo.metric_coherence_mimno_2011,
               marginal_topic_distrib=lambda dt, dl: np.asmatrix(lo.marginal_topic_distrib(dt, dl)).T)
    ref_run, ref_class = rf.reference_run_cgs_model(Recorder, fns)
    m, cells, regions = make_binary_matrix(n_cells=60, n_regions=400, density=0.08, n_true_topics=3, seed=3)
    X = sp.csr_matrix(m.T)
    frames = ["metrics", "coherence", "marg_topic", "topic_ass", "cell_topic", "topic_region", "parameters"]
    for K, top, a_by, e_by in [(7, 5, True, False), (3, 5, True, True), (6, 2, False, False), (5, 5, False, True)]:
        ref = ref_run(X, K, cells, regions, n_iter=12, random_state=555, alpha=50, alpha_by_topic=a_by, eta=0.1,
                      eta_by_topic=e_by, top_topics_coh=top, save_path=str(tmp_path / "ref"))
        fitted = made[-1].pop("fitted")
        got = lda_models.build_cistopic_model(fitted, m, K, cells, regions, 12, 555, 50, a_by, 0.1, e_by, top,
                                              ref.parameters.loc["time", "Parameter"], save_path=str(tmp_path / "got"))
        for name in frames:
            pd.testing.assert_frame_equal(getattr(got, name), getattr(ref, name), check_exact=False, rtol=1e-12)
        assert (got.n_cells, got.n_regions, got.n_topic) == (ref.n_cells, ref.n_regions, ref.n_topic) == (60, len(regions), K)
        assert str(got) == str(ref) and got.cell_topic_harmony == ref.cell_topic_harmony == []
        assert sorted(os.listdir(tmp_path / "ref")) == sorted(os.listdir(tmp_path / "got"))
        # what the mirror hands to ITS sampler (the device LDA, recorded instead of run)
        seen = {}

        class DeviceLDA:
            def __init__(self, **kw):
                seen.update(kw)

            def fit(self, X):
                return self

        monkeypatch.setattr(lda_models, "LDA", DeviceLDA)
        import types
        lda_models._fit_cgs_model_on_corpus(types.SimpleNamespace(device=0), K, 12, 555, 50, a_by, 0.1, e_by)
        for key in ("n_topics", "n_iter", "random_state", "alpha", "eta", "refresh"):
            assert seen[key] == made[-1][key], key
        assert seen["device_metrics"]["alpha"] == 50 and seen["device_metrics"]["eta"] == 0.1  # RAW values (:305)
    assert [f for f in os.listdir(tmp_path / "ref")] and isinstance(ref, ref_class)


@pytest.mark.skipif(not rf.available(), reason="/root/reference is not here")
def test_run_cgs_models_equals_the_reference_function_run_live(tmp_path, monkeypatch):
    """A1: the reference's ``run_cgs_models`` (lda_models.py:99-175; Ray stood in for by direct calls, ``lda.LDA`` by
    the oracle) against the mirror's orchestration -- the device replaced, for this test only, by the same oracle
    behind the two classes the orchestration talks to -- on two fake GPUs: same models in the order of ``n_topics``,
    the same matrix orientation reaching the sampler, the same files under ``save_path``."""
    fns = dict(metric_arun_2010=lo.metric_arun_2010, metric_cao_juan_2009=lo.metric_cao_juan_2009,
               metric_coherence_mimno_2011=lo.metric_coherence_mimno_2011,
               marginal_topic_distrib=lambda dt, dl: np.asmatrix(lo.marginal_topic_distrib(dt, dl)).T)
    ref_run, ray = rf.reference_run_cgs_models(lambda **kw: lo.OracleLDA(**kw), fns)
    m, cells, regions = make_binary_matrix(n_cells=50, n_regions=300, density=0.1, n_true_topics=3, seed=5)
    obj = SyntheticCistopicObject(m, cells, regions)
    topics = [6, 2, 9, 4]  # not sorted: the result keeps this order
    kw = dict(n_iter=8, random_state=555, alpha=50, alpha_by_topic=True, eta=0.1, eta_by_topic=False, top_topics_coh=3)
    ref = ref_run(obj, topics, n_cpu=3, save_path=str(tmp_path / "ref"), _temp_dir="/tmp/ray", **kw)
    assert ray.init_calls == [{"num_cpus": 3, "_temp_dir": "/tmp/ray"}] and ray.shutdowns == 1

    class HostCorpus:  # what lda_models asks of a Corpus
        def __init__(self, regions_by_cells, device):
            self.X, self.device = sp.csr_matrix(sp.csr_matrix(regions_by_cells).T), device

        @staticmethod
        def canonical_csr(matrix):
            return sp.csr_matrix(matrix)

        @classmethod
        def from_regions_by_cells(cls, canonical, device=0):
            return cls(canonical, device)

        def close(self):
            pass

    class HostLDA:  # what lda_models asks of the device LDA
        def __init__(self, n_topics, n_iter, random_state, alpha, eta, refresh, corpus=None, **_device_only):
            self.inner, self.corpus = lo.OracleLDA(n_topics, n_iter, alpha, eta, random_state, refresh), corpus

        def fit(self, X):
            assert X is None  # the corpus is already on the "device"
            o = self.inner.fit(self.corpus.X)
            for name in ("topic_word_", "doc_topic_", "nzw_", "ndz_", "nz_", "final_loglikelihood_"):
                setattr(self, name, getattr(o, name))
            return self

    monkeypatch.setattr(lda_models, "Corpus", HostCorpus)
    monkeypatch.setattr(lda_models, "LDA", HostLDA)
    monkeypatch.setattr(lda_models._lib, "require_device", lambda: 2)
    got = lda_models.run_cgs_models(obj, topics, n_cpu=3, save_path=str(tmp_path / "got"), _temp_dir="/tmp/ray", **kw)
    assert [g.n_topic for g in got] == [r.n_topic for r in ref] == topics
    for g, r in zip(got, ref):
        for name in ("metrics", "coherence", "marg_topic", "topic_ass", "cell_topic", "topic_region"):
            pd.testing.assert_frame_equal(getattr(g, name), getattr(r, name), check_exact=False, rtol=1e-9)
        pd.testing.assert_frame_equal(g.parameters.drop("time"), r.parameters.drop("time"))
        assert (g.n_cells, g.n_regions) == (r.n_cells, r.n_regions) == (50, len(regions))
    assert sorted(os.listdir(tmp_path / "got")) == sorted(os.listdir(tmp_path / "ref")) == sorted(f"Topic{k}.pkl" for k in topics)


@pytest.mark.skipif(not rf.available(), reason="/root/reference is not here")
def test_imputed_features_class_equals_the_reference_class_run_live(monkeypatch):
    """A16's host side: ``impute_accessibility`` (diff_features.py:340-508: cell / region selection, dtypes, the
    result class) and the ``subset`` / ``merge`` methods of ``CistopicImputedFeatures`` (:61-272), the reference's
    code executed from the tree against the mirror.  The product itself -- the device's part, compared with the
    reference's matrices in the GPU tests -- is the reference's nested function on both sides here."""
    from pycistopic_b200 import diff_features as df
    ref_class, ref_impute = rf.reference_imputed_features_class()
    ref_product = rf.reference_calculate_imputed_accessibility()
    monkeypatch.setattr(df, "calculate_imputed_accessibility",
                        lambda topic_region, cell_topic, region_names, scale_factor, chunk_size, device=0, log=None:
                        ref_product(topic_region, cell_topic, region_names, scale_factor, chunk_size))
    rng = np.random.default_rng(31)
    R, C, K = 90, 25, 4
    regions, cells = [f"chr1:{100 * i}-{100 * i + 50}" for i in range(R)], [f"c{j}___s" for j in range(C)]
    tr = rng.dirichlet(np.full(R, 0.05), K).T  # sparse-ish topics: some regions impute to zero and are dropped
    ct = rng.dirichlet(np.full(K, 0.5), C).T
    model = lda_models.CistopicLDAModel(None, None, None, None, pd.DataFrame(ct, index=[f"Topic{k + 1}" for k in range(K)], columns=cells),
                                        pd.DataFrame(tr, index=regions, columns=[f"Topic{k + 1}" for k in range(K)]), None)
    obj = SyntheticCistopicObject(sp.csr_matrix((R, C), dtype=np.int32), cells, regions)
    obj.selected_model = model
    for kw in (dict(), dict(scale_factor=1), dict(selected_cells=cells[3:15], selected_regions=regions[10:70][::-1], chunk_size=17)):
        a, b = df.impute_accessibility(obj, project="p", **kw), ref_impute(obj, project="p", **kw)
        assert a.mtx.dtype == b.mtx.dtype and np.array_equal(a.mtx, b.mtx)
        if not kw:  # the int32 path truncates: regions that impute to zero everywhere are dropped
            assert a.mtx.dtype == np.int32 and a.mtx.shape[0] < len(regions)
        assert list(a.feature_names) == list(b.feature_names) and list(a.cell_names) == list(b.cell_names)
        assert a.project == b.project and str(a) == str(b)

    def pair(features, cell_list, seed, sparse_mtx=False):
        m = np.random.default_rng(seed).integers(0, 3, (len(features), len(cell_list))).astype(np.int32)
        m = sp.csr_matrix(m) if sparse_mtx else m
        return (df.CistopicImputedFeatures(m.copy(), list(features), list(cell_list), "p"),
                ref_class(m.copy(), list(features), list(cell_list), "p"))

    # subset: by tagged names, by bare barcodes, by features; all-zero features go
    for cells_arg, feats_arg in ((cells[2:9], None), ([c.split("___")[0] for c in cells[5:12]], regions[0:40:3]), (None, regions[50:60])):
        a, b = pair(regions, cells, 1)
        a2, b2 = a.subset(cells=cells_arg, features=feats_arg, copy=True), b.subset(cells=cells_arg, features=feats_arg, copy=True)
        assert np.array_equal(a2.mtx, b2.mtx) and a2.feature_names == b2.feature_names and a2.cell_names == b2.cell_names
        a.subset(cells=cells_arg, features=feats_arg)
        b.subset(cells=cells_arg, features=feats_arg)
        assert np.array_equal(a.mtx, b.mtx) and a.feature_names == b.feature_names and a.cell_names == b.cell_names
    # merge: overlapping, disjoint and identical feature sets; dense and sparse; several objects at once
    def same_up_to_row_order(a, b):
        assert a.cell_names == b.cell_names and sorted(a.feature_names) == sorted(b.feature_names) and a.project == b.project
        am, bm = (x.toarray() if sp.issparse(x) else np.asarray(x) for x in (a.mtx, b.mtx))
        assert a.mtx.dtype == b.mtx.dtype and sp.issparse(a.mtx) == sp.issparse(b.mtx)
        order = [b.feature_names.index(f) for f in a.feature_names]
        assert np.array_equal(am, bm[order])

    for sparse_mtx in (False, True):
        for sets in ((regions[:40], regions[20:60]), (regions[:20], regions[20:30]), (regions[:30], regions[:30][::-1]),
                     (regions[:40], regions[30:50], regions[45:80])):
            mine, theirs = zip(*[pair(f, [f"s{i}_{c}" for c in cells[:4 + i]], 10 + i, sparse_mtx) for i, f in enumerate(sets)])
            merged_a = mine[0].merge(list(mine[1:]), project="m", copy=True)
            merged_b = theirs[0].merge(list(theirs[1:]), project="m", copy=True)
            same_up_to_row_order(merged_a, merged_b)
            mine[0].merge(list(mine[1:]), project="m")
            same_up_to_row_order(mine[0], merged_b)


@pytest.mark.skipif(not rf.available(), reason="/root/reference is not here")
def test_find_diff_features_equals_the_reference_function_run_live(monkeypatch):
    """N2's host side: ``find_diff_features`` (diff_features.py:716-836: contrasts from a cell_data column or given,
    their names, the barcode groups, ``var_features``, BH adjustment, thresholds, ordering of the tables) -- the
    reference's function with its own ``markers`` beneath it against the mirror, whose device step (rank sums and
    log2FC of every contrast, compared with SciPy in the GPU tests) is the reference's arithmetic here."""
    from pycistopic_b200 import diff_features as df
    ref_find, ref_class, ref_ns = rf.reference_find_diff_features()

    def host_contrasts(matrix, contrasts_cols, device=0):
        n, R = len(contrasts_cols), matrix.shape[0]
        stat, pval, l2fc = (np.full((n, R), np.nan) for _ in range(3))
        for i, (f, b) in enumerate(contrasts_cols):
            fg, bg = np.asarray(matrix)[:, f], np.asarray(matrix)[:, b]
            pval[i] = ref_ns["get_wilcox_test_pvalues"](fg, bg)
            l2fc[i] = ref_ns["get_log2_fc"](fg.astype(np.float64), bg.astype(np.float64))
        return stat, pval, l2fc

    monkeypatch.setattr(df, "ranksum_contrasts", host_contrasts)
    rng = np.random.default_rng(41)
    R, C = 120, 60
    cells, regions = [f"c{j}___s" for j in range(C)], [f"chr2:{10 * i}-{10 * i + 5}" for i in range(R)]
    kind = np.array(["A", "B", "C"])[np.arange(C) % 3]
    mtx = rng.poisson(3.0, (R, C)).astype(np.float64)
    for t, letter in enumerate("ABC"):  # planted signal: a block of regions per cell type
        mtx[40 * t:40 * t + 15, kind == letter] *= 4
    cell_data = pd.DataFrame({"celltype": kind, "other": 1}, index=cells)
    cell_data.loc[cells[7], "celltype"] = np.nan  # a cell without a label is left out (dropna, :762)
    obj = SyntheticCistopicObject(sp.csr_matrix((R, C), dtype=np.int32), cells, regions)
    obj.cell_data = cell_data
    cases = [dict(), dict(var_features=regions[5:100:2]), dict(contrasts=[[["A"], ["B"]], [["A", "B"], ["C"]]], adjpval_thr=0.2,
                                                             log2fc_thr=0.5)]
    for kw in cases:
        got = df.find_diff_features(obj, df.CistopicImputedFeatures(mtx.copy(), list(regions), list(cells), "p"), "celltype", **kw)
        want = ref_find(obj, ref_class(mtx.copy(), list(regions), list(cells), "p"), "celltype", **kw)
        assert list(got) == list(want) and len(got) in (2, 3)
        for name in want:
            assert len(want[name]) >= 3
            pd.testing.assert_frame_equal(got[name], want[name], check_exact=False, rtol=1e-12)


def _fake_model(k, arun, cao, mimno, ll):
    metrics = pd.DataFrame([arun, cao, mimno, ll], index=["Arun_2010", "Cao_Juan_2009", "Mimno_2011", "loglikelihood"],
                           columns=["Metric"]).transpose()
    ct = pd.DataFrame(np.zeros((k, 3)))
    tr = pd.DataFrame(np.zeros((4, k)))
    return lda_models.CistopicLDAModel(metrics, None, None, None, ct, tr, None)


def test_evaluate_models_selection():
    """Model selection arithmetic of lda_models.py:1109-1216, checked against a direct computation."""
    ks = [2, 5, 10, 15, 20]
    arun = [9.0, 7.0, 4.0, 5.0, 6.0]
    cao = [0.5, 0.3, 0.2, 0.25, 0.4]
    mim = [-0.2, -0.5, -0.3, -0.35, -0.6]
    ll = [-100.0, -80.0, -60.0, -65.0, -90.0]
    models = [_fake_model(k, a, c, m, l) for k, a, c, m, l in zip(ks, arun, cao, mim, ll)]
    shuffled = [models[i] for i in (3, 0, 4, 1, 2)]
    best = lda_models.evaluate_models(shuffled, plot=False)
    resc = lambda v: (np.array(v) - min(v)) / (max(v) - min(v))
    idx = [i for i, k in enumerate(ks) if k >= 5]
    comb = (resc([-x for x in arun])[idx] + resc([-x for x in cao])[idx] + resc([mim[i] for i in idx]) + resc(ll)[idx])
    assert best.n_topic == ks[idx[int(np.argmax(comb))]] == 10
    assert lda_models.evaluate_models(shuffled, select_model=15, plot=False).n_topic == 15
    assert lda_models.evaluate_models(shuffled, plot=False, return_model=False) is None
    only = lda_models.evaluate_models(shuffled, metrics=["loglikelihood", "Arun_2010"], plot=False)
    assert only.n_topic == 10
    best2 = lda_models.evaluate_models(shuffled, metrics=["Cao_Juan_2009"], plot=False)
    assert best2.n_topic == 10


@pytest.mark.skipif(not rf.available(), reason="needs /root/reference (authoring container)")
def test_evaluate_models_equals_the_reference_function_run_live():
    """The mirror against the reference's OWN evaluate_models (lda_models.py:1048-1270, AST-executed from the tree,
    matplotlib replaced by an inert stand-in): same selected model for every metric subset, min_topics_coh and
    model order, on random metric tables."""
    ref = rf.reference_evaluate_models()
    rng = np.random.default_rng(1048)
    all_metrics = ["Minmo_2011", "loglikelihood", "Cao_Juan_2009", "Arun_2010"]
    for trial in range(60):
        ks = sorted(rng.choice(np.arange(2, 60), size=int(rng.integers(3, 9)), replace=False).tolist())
        models = [_fake_model(k, float(rng.uniform(1, 20)), float(rng.uniform(0.05, 0.9)), float(rng.uniform(-1, -0.1)),
                              float(rng.uniform(-1e6, 1e6))) for k in ks]
        models = [models[i] for i in rng.permutation(len(models))]
        n_sel = int(rng.integers(1, 5))
        metrics = [all_metrics[i] for i in sorted(rng.choice(4, size=n_sel, replace=False))]
        min_coh = int(rng.choice([2, 5, 5, 10]))
        if "Minmo_2011" in metrics and sum(k >= min_coh for k in ks) < 2:
            continue  # rescaling a single value divides by zero in the reference too
        want = ref(list(models), metrics=list(metrics), min_topics_coh=min_coh, plot=False)
        got = lda_models.evaluate_models(list(models), metrics=list(metrics), min_topics_coh=min_coh, plot=False)
        assert got.n_topic == want.n_topic, (ks, metrics, min_coh)
        pick = int(rng.choice(ks))
        assert lda_models.evaluate_models(list(models), select_model=pick, plot=False).n_topic == \
            ref(list(models), select_model=pick, plot=False).n_topic == pick
    assert ref(list(models), return_model=False, plot=False) is None


def test_schedule_models_lpt():
    topics = [2, 5, 10, 15, 20, 25, 30, 35, 40, 45, 50]
    for n in (1, 2, 4, 8):
        plan = lda_models.schedule_models(topics, n)
        assert sorted(i for p in plan for i in p) == list(range(len(topics)))
        loads = [sum(lda_models.model_cost(topics[i]) for i in p) for p in plan]
        total = sum(lda_models.model_cost(k) for k in topics)
        assert max(loads) <= max(total / n * 4 / 3 + 1e-9, lda_models.model_cost(50))
    assert lda_models.schedule_models(topics, 1) == [sorted(range(len(topics)), key=lambda i: -topics[i])]


def test_synthetic_shapes_and_duck_object():
    m, cells, regions = make_binary_matrix(n_cells=50, n_regions=2000, density=0.03, n_true_topics=5, seed=2)
    assert m.dtype == np.int32 and m.max() == 1 and m.shape == (len(regions), len(cells))
    assert m.has_sorted_indices
    dens = m.nnz / (2000 * 50)  # against the nominal shape (empty regions are dropped)
    assert 0.02 < dens < 0.05
    m2, _, _ = make_binary_matrix(n_cells=50, n_regions=2000, density=0.03, n_true_topics=5, seed=2)
    assert (m != m2).nnz == 0


def test_pickle_interop_with_reference_class_path(fitted, tmp_path):
    """after alias_pycistopic() a saved Topic{K}.pkl names the reference's class
    (pycisTopic.lda_models.CistopicLDAModel, lda_models.py:394-395) and loads back; run in a subprocess because
    the alias renames the class for the whole interpreter."""
    import subprocess
    m, cells, regions, X, o = fitted
    model = lda_models.build_cistopic_model(o, m, 7, cells, regions, 30, 555, 50, True, 0.1, False, 5, 1.25)
    plain = tmp_path / "plain.pkl"
    with open(plain, "wb") as f:
        pickle.dump(model, f)
    code = f"""
import pickle, sys
sys.path.insert(0, {ROOT!r})
from pycistopic_b200.compat import alias_pycistopic
assert alias_pycistopic() is True
import pycisTopic.lda_models as ref
from pycisTopic.lda_models import CistopicLDAModel
m = pickle.load(open({str(plain)!r}, "rb"))
blob = pickle.dumps(m)
assert b"pycisTopic.lda_models" in blob and b"pycistopic_b200" not in blob
again = pickle.loads(blob)
assert type(again) is CistopicLDAModel and again.n_topic == 7 and again.topic_region.equals(m.topic_region)
print("ok")
"""
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.strip().endswith("ok"), res.stderr


def test_saved_models_load_in_a_plain_pycistopic_environment(fitted, tmp_path):
    """`run_cgs_models(save_path=...)` / the CLI must write pickles a pycisTopic environment WITHOUT this
    package can load, in both situations: (a) pycisTopic not installed where the file is written (the class is
    pickled under the reference's qualified name), (b) pycisTopic installed there (models are re-wrapped in its own
    class).  "pycisTopic" is a minimal stand-in package holding the reference's class layout (lda_models.py:31-96,
    cistopic_class.py:31-51); the reader process has only that on its path.  Also: a reference cisTopic-object pickle
    (pycisTopic.cistopic_class.CistopicObject) loads in stub mode -- what the CLI's --input needs."""
    import subprocess
    import textwrap
    m, cells, regions, X, o = fitted
    fake = tmp_path / "site"
    (fake / "pycisTopic").mkdir(parents=True)
    (fake / "pycisTopic" / "__init__.py").write_text("")
    (fake / "pycisTopic" / "lda_models.py").write_text(textwrap.dedent("""
        class CistopicLDAModel:
            def __init__(self, metrics, coherence, marg_topic, topic_ass, cell_topic, topic_region, parameters):
                self.metrics, self.coherence, self.marg_topic, self.topic_ass = metrics, coherence, marg_topic, topic_ass
                self.cell_topic, self.cell_topic_harmony, self.topic_region = cell_topic, [], topic_region
                self.parameters = parameters
                self.n_cells, self.n_regions, self.n_topic = cell_topic.shape[1], topic_region.shape[0], cell_topic.shape[0]
        """))
    (fake / "pycisTopic" / "cistopic_class.py").write_text(textwrap.dedent("""
        class CistopicObject:
            def __init__(self, fragment_matrix, binary_matrix, cell_names, region_names, cell_data, region_data,
                         path_to_fragments, project="cisTopic"):
                self.fragment_matrix, self.binary_matrix = fragment_matrix, binary_matrix
                self.cell_names, self.region_names = cell_names, region_names
                self.cell_data, self.region_data = cell_data, region_data
                self.path_to_fragments, self.project = path_to_fragments, project
                self.selected_model, self.projections = None, {"cell": {}, "region": {}}
        """))
    reader = textwrap.dedent(f"""
        import pickle, sys
        sys.path.insert(0, {str(fake)!r})
        assert not any("pycistopic_b200" in p for p in sys.modules)
        m = pickle.load(open(sys.argv[1], "rb"))
        m = m[0] if isinstance(m, list) else m
        import pycisTopic.lda_models as ref
        assert type(m) is ref.CistopicLDAModel and m.n_topic == 7 and m.topic_region.shape[1] == 7
        assert "pycistopic_b200" not in sys.modules
        print("ok")
        """)
    for mode, extra_path in (("stub", None), ("installed", str(fake))):
        save = tmp_path / f"models_{mode}"
        writer = textwrap.dedent(f"""
            import pickle, sys
            sys.path.insert(0, {ROOT!r})
            if {extra_path!r}:
                sys.path.insert(0, {extra_path!r})
            from pycistopic_b200 import compat, lda_models
            o, m, cells, regions = pickle.load(open({str(tmp_path / "fit.pkl")!r}, "rb"))
            model = lda_models.build_cistopic_model(o, m, 7, cells, regions, 30, 555, 50, True, 0.1, False, 5, 1.25,
                                                    save_path={str(save)!r})
            assert compat.alias_pycistopic() is ({mode!r} == "stub")
            with open({str(save / "list.pkl")!r}, "wb") as f:
                pickle.dump(compat.pickle_ready([model]), f)
            print("ok")
            """)
        class _Fit:
            pass
        keep = _Fit()
        for a in ("topic_word_", "doc_topic_", "nzw_", "ndz_", "nz_", "final_loglikelihood_"):
            if hasattr(o, a):
                setattr(keep, a, getattr(o, a))
        with open(tmp_path / "fit.pkl", "wb") as f:
            pickle.dump(({k: v for k, v in vars(keep).items()}, m, cells, regions), f)
        writer = writer.replace("o, m, cells, regions = pickle.load", "d, m, cells, regions = pickle.load")
        writer = writer.replace("model = lda_models", "o = type('Fit', (), d)()\nmodel = lda_models")
        res = subprocess.run([sys.executable, "-c", writer], capture_output=True, text=True)
        assert res.returncode == 0 and res.stdout.strip().endswith("ok"), (mode, res.stderr)
        for name in ("Topic7.pkl", "list.pkl"):
            res = subprocess.run([sys.executable, "-c", reader, str(save / name)], capture_output=True, text=True)
            assert res.returncode == 0 and res.stdout.strip().endswith("ok"), (mode, name, res.stderr)
    # a reference cisTopic object pickle (written by the stand-in package) loads in stub mode
    obj_pkl = tmp_path / "cistopic_obj.pkl"
    res = subprocess.run([sys.executable, "-c", textwrap.dedent(f"""
        import pickle, sys
        sys.path.insert(0, {str(fake)!r})
        from pycisTopic.cistopic_class import CistopicObject
        pickle.dump(CistopicObject(None, None, ["c1"], ["chr1:1-2"], None, None, "f.tsv.gz"), open({str(obj_pkl)!r}, "wb"))
        """)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    res = subprocess.run([sys.executable, "-c", textwrap.dedent(f"""
        import pickle, sys
        sys.path.insert(0, {ROOT!r})
        from pycistopic_b200 import compat, fragment_matrix
        assert compat.alias_pycistopic() is True
        obj = pickle.load(open({str(obj_pkl)!r}, "rb"))
        assert type(obj) is fragment_matrix.CistopicObject and obj.cell_names == ["c1"]
        print("ok")
        """)], capture_output=True, text=True)
    assert res.returncode == 0 and res.stdout.strip().endswith("ok"), res.stderr


def test_diff_feature_oracle_matches_reference_golden():
    """oracle/diff_oracle.py against tests/golden/diff_features.npz (the reference's own get_wilcox_test_pvalues /
    get_log2_fc / p_adjust_bh, diff_features.py:968-1120, executed from the tree) and its NumPy restatement of
    scipy.stats.ranksums against scipy itself."""
    from oracle import diff_oracle as do
    from pycistopic_b200 import diff_features as pdf
    g = np.load(os.path.join(ROOT, "tests", "golden", "diff_features.npz"))
    for i in range(int(g["n"])):
        fg, bg = g[f"fg{i}"], g[f"bg{i}"]
        stat, p = do.ranksums_rows(fg, bg)
        np.testing.assert_allclose(p, g[f"p{i}"], rtol=1e-12)
        # float32 input: the reference's mean accumulates in float32, the oracle (and the GPU) in float64
        f32 = fg.dtype == np.float32
        np.testing.assert_allclose(do.log2_fc(fg, bg), g[f"l2fc{i}"], rtol=1e-6 if f32 else 1e-12,
                                   atol=2e-6 if f32 else 0, equal_nan=True)
        np.testing.assert_allclose(do.p_adjust_bh(p), g[f"padj{i}"], rtol=1e-14)
        np.testing.assert_allclose(pdf.p_adjust_bh(p), g[f"padj{i}"], rtol=1e-14)
        for r in range(0, fg.shape[0], 7):
            z, pr = do.ranksums_restated(fg[r], bg[r])
            assert z == pytest.approx(stat[r], rel=1e-12, abs=1e-15) and pr == pytest.approx(p[r], rel=1e-12)


def test_rankings_oracle_matches_reference_golden():
    """oracle/rankings_oracle.py against the reference's own make_rankings (diff_features.py:274-337, executed
    from the tree into tests/golden/rankings.npz): the NumPy restatement of the reference closure reproduces it
    entirely (same PCG64 stream, same argsort); the restatement of the DEVICE algorithm (keyed Feistel bijection +
    stable sort) reproduces it on every column without ties and yields the same scores at every rank elsewhere."""
    from oracle import rankings_oracle as ro
    g = np.load(os.path.join(ROOT, "tests", "golden", "rankings.npz"))
    for xk, rk, seed in (("x_int", "rank_int", 123), ("x_float", "rank_float", 7)):
        x, ref = g[xk], g[rk]
        np.testing.assert_array_equal(ro.reference_make_rankings(x, seed=seed), ref)
        got = ro.rank_columns(x, seed=seed)
        for c in range(x.shape[1]):
            assert np.array_equal(np.sort(got[:, c]), np.arange(x.shape[0]))
            if len(np.unique(x[:, c])) == x.shape[0]:
                np.testing.assert_array_equal(got[:, c], ref[:, c])
            else:
                np.testing.assert_array_equal(x[np.argsort(got[:, c]), c], x[np.argsort(ref[:, c]), c])
    for n in (1, 2, 3, 17, 256, 257, 65536, 65537, 100_003):
        perm = ro.permutation(n, ro.column_key(123, n))
        assert np.array_equal(np.sort(perm), np.arange(n))


def test_mallet_formats_roundtrip_and_reference_text(tmp_path):
    """mallet_formats: corpus text identical to the reference's corpus_to_mallet
    (lda_models.py:472-494; executed from the tree when it is there, else the documented line format), state files
    round-trip, load_word_topics counts columns 4 / 5 as the reference does (:620-646)."""
    import gzip
    import io
    import scipy.sparse as sp
    from pycistopic_b200 import mallet_formats as mf
    rng = np.random.default_rng(4)
    X = (rng.random((40, 9)) < 0.25).astype(np.int32)          # regions x cells
    X[:, 3] = 0                                                # an empty cell
    buf = io.StringIO()
    mf.corpus_to_mallet(sp.csr_matrix(X), buf)
    lines = buf.getvalue().splitlines()
    assert len(lines) == 9 and lines[3] == "3\t0\t"
    for c, line in enumerate(lines):
        assert line == f"{c}\t0\t" + " ".join(str(r) for r in np.flatnonzero(X[:, c]))
    bow = [[(int(r), 1) for r in np.flatnonzero(X[:, c])] for c in range(9)]   # gensim Sparse2Corpus view (:797-803)
    buf2 = io.StringIO()
    mf.corpus_to_mallet(bow, buf2)
    assert buf2.getvalue() == buf.getvalue()
    if rf.available():
        buf3 = io.StringIO()
        rf.reference_corpus_to_mallet()(bow, buf3)
        assert buf3.getvalue() == buf.getvalue()
    # state file: write -> read -> counts
    Xc = sp.csc_matrix(X)
    doc = np.repeat(np.arange(9), np.diff(Xc.indptr))
    region, K = Xc.indices, 5
    z = rng.integers(0, K, len(region))
    for name in ("state.gz", "state.txt"):
        path = str(tmp_path / name)
        mf.write_state(path, doc, region, z, np.full(K, 2.5), 0.1)
        opener = gzip.open if name.endswith(".gz") else open
        with opener(path, "rt") as f:
            head = [f.readline() for _ in range(4)]
        assert head[0] == "#doc source pos typeindex type topic\n" and head[1].startswith("#alpha : 2.5 2.5")
        assert head[2] == "#beta : 0.1\n" and head[3] == f"{doc[0]} NA 0 0 {region[0]} {z[0]}\n"
        st = mf.read_state(path)
        np.testing.assert_array_equal(st["alpha"], np.full(K, 2.5))
        assert st["beta"] == 0.1
        np.testing.assert_array_equal(st["doc"], doc)
        np.testing.assert_array_equal(st["region"], region)
        np.testing.assert_array_equal(st["topic"], z)
        wt = mf.load_word_topics(path, K, 40)
        want = np.zeros((K, 40))
        np.add.at(want, (z, region), 1)
        np.testing.assert_array_equal(wt, want)
        with pytest.raises(AssertionError, match="Mismatch between MALLET"):
            mf.load_word_topics(path, K + 1, 40)
    bad = tmp_path / "bad.txt"
    bad.write_text("not a state\n")
    with pytest.raises(ValueError, match="not a Mallet state file"):
        mf.read_state(str(bad))


def test_variable_feature_selection_matches_reference_golden():
    """select_variable_features (host half of find_highly_variable_features) against the reference's own function
    (diff_features.py:577-714, run from the tree into tests/golden/variable_features.npz) on NumPy's float32 mean /
    var: identical feature lists for the default window, n_top_features and a custom window, three dtypes."""
    from pycistopic_b200 import diff_features as pdf
    g = np.load(os.path.join(ROOT, "tests", "golden", "variable_features.npz"))
    for name in ("norm64", "imputed", "norm32"):
        mean, var = g[f"{name}_mean"], g[f"{name}_var"]
        feats = [f"r{i}" for i in range(len(mean))]
        assert pdf.select_variable_features(mean, var, feats) == list(g[f"{name}_default"])
        assert pdf.select_variable_features(mean, var, feats, n_top_features=100) == list(g[f"{name}_top100"])
        assert pdf.select_variable_features(mean, var, feats, min_disp=0.5, min_mean=0.05, max_mean=2.0, max_disp=3.0,
                                            n_bins=10) == list(g[f"{name}_window"])
    assert len(g["norm64_default"]) > 100 and len(g["norm64_window"]) > 20


def test_cli_flags_mirror_the_reference():
    """python -m pycistopic_b200 topic_modeling lda: flags, types and defaults of the reference parser
    (cli/subcommand/topic_modeling.py:171-302); parsing only -- the run itself is a GPU test."""
    from pycistopic_b200.__main__ import build_parser
    p = build_parser()
    a = p.parse_args("topic_modeling lda -i in.pkl -o out.pkl -t 2 5 10 -p 4".split())
    assert (a.input, a.output, a.topics, a.parallel) == ("in.pkl", "out.pkl", [2, 5, 10], 4)
    assert (a.iterations, a.alpha, a.alpha_by_topic, a.eta, a.eta_by_topic) == (500, 50, True, 0.1, False)
    assert (a.keep_intermediate_topic_models, a.seed, a.temp_dir) == (False, 555, None)
    a = p.parse_args("topic_modeling lda --input i --output o --topics 7 --parallel 1 --iterations 20 --alpha 10 "
                     "--alpha_by_topic no --eta 0.5 --eta_by_topic TRUE --keep y --seed 3 --temp_dir /tmp/x".split())
    assert (a.iterations, a.alpha, a.alpha_by_topic, a.eta, a.eta_by_topic) == (20, 10, False, 0.5, True)
    assert (a.keep_intermediate_topic_models, a.seed, a.temp_dir) == (True, 3, "/tmp/x")
    for bad in ("topic_modeling lda -i a -o b -t 2", "topic_modeling lda -i a -o b -p 1", "topic_modeling lda -i a -o b -t 2 -p 1 -A maybe"):
        with pytest.raises(SystemExit):
            p.parse_args(bad.split())


def test_region_shards_cover_the_range():
    """Host bookkeeping of the sharded consumers (SURVEY.md 8e iii): contiguous, disjoint, non-empty ranges that
    cover range(n) in order, one per device, fewer when there is less work than devices."""
    from pycistopic_b200 import diff_features as pdf
    assert pdf._as_devices(2) == [2] and pdf._as_devices((0, 1)) == [0, 1] and pdf._as_devices(np.int64(3)) == [3]
    with pytest.raises(ValueError, match="at least one device"):
        pdf._as_devices([])
    for n in (0, 1, 2, 7, 100, 200_001):
        for devices in ([0], [0, 1], [3, 2, 1, 0], list(range(8))):
            shards = pdf._shards(n, devices)
            assert len(shards) <= min(n, len(devices)) and (n < len(devices) or len(shards) == len(devices))
            pos = 0
            for dev, a, b in shards:
                assert a == pos and b > a and dev in devices
                pos = b
            assert pos == n
            sizes = [b - a for _, a, b in shards]
            assert not sizes or max(sizes) - min(sizes) <= 1 or n < len(devices)
    assert pdf._run_sharded(lambda d, a, b: (d, b - a), pdf._shards(10, [5, 6, 7])) == [(5, 3), (6, 3), (7, 4)]
    assert pdf._run_sharded(lambda d, a, b: 1, []) == []


def test_mallet_state_is_validated_before_the_device_is_touched(tmp_path):
    """lda_from_state refuses a state whose tokens are not those of the matrix, or whose topics do not fit, with a
    ValueError raised on the host (no GPU needed to get there); shuffled state lines are put back in corpus order."""
    import scipy.sparse as sp
    from pycistopic_b200 import mallet_formats as mf
    rng = np.random.default_rng(6)
    X = sp.csr_matrix((rng.random((30, 8)) < 0.3).astype(np.int32))      # regions x cells
    Xc = sp.csc_matrix(X); Xc.sort_indices()
    doc = np.repeat(np.arange(8), np.diff(Xc.indptr))
    z = rng.integers(0, 4, Xc.nnz)
    path = str(tmp_path / "s.gz")
    mf.write_state(path, doc, Xc.indices, z, np.full(4, 1.0), 0.1)
    other = X.tolil(copy=True); other[0, 0] = 1 - other[0, 0]
    with pytest.raises(ValueError, match="tokens"):
        mf.lda_from_state(sp.csr_matrix(other), path, 4, 1.0, 0.1)
    with pytest.raises(ValueError, match="outside"):
        mf.lda_from_state(X, path, 3, 1.0, 0.1)
    state = mf.read_state(path)
    state["region"] = state["region"].copy(); state["region"][0] = 29 if state["region"][0] != 29 else 28
    with pytest.raises(ValueError, match="not those of the matrix"):
        mf.lda_from_state(X, state, 4, 1.0, 0.1)


@pytest.mark.skipif(not rf.available(), reason="/root/reference not present")
def test_variable_feature_selection_matches_live_reference():
    """select_variable_features against the reference's function run live from the tree, on inputs that hit its
    corner cases: all-zero regions (mean 0 -> 1e-12), constant regions (dispersion 0 -> NaN), bins holding a single
    region (std NaN -> normalised dispersion 1), few regions, n_top_features ties."""
    import logging
    from pycistopic_b200 import diff_features as pdf
    fn = rf.reference_find_highly_variable_features()
    logging.disable(logging.INFO)
    try:
        rng = np.random.default_rng(31)
        for R, C, n_bins in ((400, 60, 20), (35, 20, 20), (120, 33, 5), (7, 9, 3)):
            mat = np.abs(rng.normal(0.3, 0.4, (R, C))) * (rng.random((R, 1)) < 0.9)
            mat[1] = 0.7                       # constant region
            mat[2, :] = 0; mat[2, 0] = 40.0    # an outlier mean: alone in its bin
            frame = pd.DataFrame(mat, index=[f"r{i}" for i in range(R)])
            mean = np.mean(mat, axis=1, dtype=np.float32)
            var = np.var(mat, axis=1, dtype=np.float32)
            feats = frame.index.tolist()
            for kw in ({}, {"n_top_features": max(1, R // 5)}, {"min_disp": -0.5, "min_mean": 0.0, "max_mean": 50, "max_disp": 2.0}):
                want = fn(frame, n_bins=n_bins, plot=False, **kw)
                got = pdf.select_variable_features(mean, var, feats, n_bins=n_bins, **kw)
                assert got == want, (R, C, n_bins, kw)
    finally:
        logging.disable(logging.NOTSET)


def test_pin_against_pypi_runs_or_reports_absent():
    """oracle/pin_against_pypi.py compares the oracle with the real `lda` / `tmtoolkit` wherever they are
    installed; here (no network) it must report them absent -- and if a later image does ship them, it must pin."""
    from oracle import pin_against_pypi as pin
    code, report = pin.run(verbose=False)
    if pin.packages()[0] is None:
        assert code == pin.EXIT_ABSENT
    else:
        assert code == pin.EXIT_PINNED, report


def test_pin_against_pyranges_runs_or_reports_absent():
    """oracle/pin_against_pyranges.py: the interval join / blacklist / read_bed half of the count-matrix oracle against
    the real pyranges wherever it is installed; here it must report it absent (and must not write a pin file)."""
    from oracle import pin_against_pyranges as pin
    try:
        import pyranges  # noqa: F401
        present = True
    except Exception:
        present = False
    code, report = pin.run(verbose=False)
    assert code == (pin.EXIT_PINNED if present else pin.EXIT_ABSENT), report
    assert present or not os.path.exists(os.path.join(GOLD, "pyranges_pin.json"))


def test_lazy_nccl_resolution_does_not_break_a_later_torch_import():
    """cgs_nccl_* dlopen NCCL at first use.  In a process that has not imported torch yet the loader's default is the
    SYSTEM libnccl.so.2; torch would then be bound to it by soname and fail on the symbols only its own (newer) wheel
    has.  _lib.load() therefore names the wheel's copy in CGS_NCCL_LIBRARY: the id call must load that file and
    `import torch` must still work afterwards (fresh interpreter; ncclGetUniqueId needs no GPU)."""
    import subprocess
    code = (
        "import ctypes, os, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from pycistopic_b200 import _lib\n"
        "L = _lib.load()\n"
        "assert 'torch' not in sys.modules\n"
        "uid = ctypes.create_string_buffer(128)\n"
        "assert L.cgs_nccl_unique_id(uid) == 0, L.cgs_last_error()\n"
        "assert any(uid.raw)\n"
        "loaded = sorted(set(l.split()[-1] for l in open('/proc/self/maps') if 'libnccl' in l))\n"
        "named = os.environ.get('CGS_NCCL_LIBRARY')\n"
        "assert named is None or loaded == [os.path.realpath(named)] or loaded == [named], (named, loaded)\n"
        "import torch\n"
        "print('ok', loaded)\n"
    )
    env = {k: v for k, v in os.environ.items() if k != "CGS_NCCL_LIBRARY"}
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr


def test_bench_extra_leg_reports_failures_where_no_rank_is_left_waiting():
    """bench.py: a failing extra leg is reported under its key in one process or when rank 0 runs it alone; a leg
    with collectives on several ranks still fails the run."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    def boom(what):
        raise ValueError("bad " + what)

    assert bench.run_extra_leg(1, False, lambda x, y=1: {"v": x + y}, 2, y=3) == {"v": 5}
    assert bench.run_extra_leg(1, False, boom, "leg") == {"ok": False, "error": "ValueError: bad leg"}
    assert bench.run_extra_leg(8, True, boom, "solo")["ok"] is False
    with pytest.raises(ValueError):
        bench.run_extra_leg(8, False, boom, "collective")

