"""The host-side mirrors against what the REFERENCE's own functions returned on the same seeded inputs, frozen in
tests/golden/host_mirrors.npz by tests/golden/make_host_mirrors.py (which executes run_cgs_model,
create_cistopic_object, add_cell_data, find_diff_features + markers, CistopicImputedFeatures.subset / merge and
read_fragments_to_pyranges from /root/reference).  Unlike the "run live" tests of test_cpu_oracle_and_host.py and
test_cpu_fragment_files.py these need no reference tree: they travel with the repository."""
import os
import sys
import types

import numpy as np
import pandas as pd
import pytest
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import host_mirror_cases as hc  # noqa: E402
from pycistopic_b200 import diff_features as df  # noqa: E402
from pycistopic_b200 import fragment_matrix as fm  # noqa: E402
from pycistopic_b200 import lda_models  # noqa: E402


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(HERE, "golden", "host_mirrors.npz"))


def test_model_packing_and_hyperparameter_scaling(gold, monkeypatch):
    """A2 / A14 (lda_models.py:247-388)."""
    m, cells, regions = hc.packing_matrix()
    for i, (K, top, a_by, e_by) in enumerate(hc.PACKING_CASES):
        fitted = types.SimpleNamespace(**{name: gold[f"packing{i}/{name}"] for name in
                                          ("nzw_", "ndz_", "nz_", "topic_word_", "doc_topic_")})
        got = lda_models.build_cistopic_model(fitted, m, K, cells, regions, 12, 555, 50, a_by, 0.1, e_by, top, 0.0)
        for frame in ("metrics", "coherence", "marg_topic", "topic_ass", "cell_topic", "topic_region", "parameters"):
            hc.assert_frame_is(f"packing{i}/{frame}", getattr(got, frame), gold, rtol=1e-9)
        seen = {}

        class DeviceLDA:
            def __init__(self, **kw):
                seen.update(kw)

            def fit(self, X):
                return self

        monkeypatch.setattr(lda_models, "LDA", DeviceLDA)
        lda_models._fit_cgs_model_on_corpus(types.SimpleNamespace(device=0), K, 12, 555, 50, a_by, 0.1, e_by)
        assert [seen["alpha"], seen["eta"]] == gold[f"packing{i}/sampler_alpha_eta"].tolist()
        assert [seen["n_topics"], seen["n_iter"], seen["random_state"], seen["refresh"]] == gold[f"packing{i}/sampler_ints"].tolist()


def test_create_cistopic_object_and_add_cell_data(gold):
    """N4 (cistopic_class.py:508-655, :88-142)."""
    cm, cells, regions = hc.count_matrix()
    for i, kw in enumerate(hc.OBJECT_CASES):
        got = fm.create_cistopic_object(cm.copy(), cell_names=list(cells), region_names=list(regions), project="s1", **kw)
        assert got.cell_names == gold[f"object{i}/cell_names"].tolist() and got.region_names == gold[f"object{i}/region_names"].tolist()
        assert np.array_equal(got.binary_matrix.toarray(), gold[f"object{i}/binary"])
        assert np.array_equal(got.fragment_matrix.toarray(), gold[f"object{i}/fragments"])
        hc.assert_frame_is(f"object{i}/cell_data", got.cell_data, gold)
        hc.assert_frame_is(f"object{i}/region_data", got.region_data, gold)
    for name, meta in hc.cell_metadata(cells).items():
        got = fm.create_cistopic_object(cm.copy(), list(cells), list(regions), project="s1")
        got.add_cell_data(meta.copy(), "___")
        hc.assert_frame_is(f"add_cell_data/{name}", got.cell_data.sort_index(axis=1), gold)


def test_find_diff_features_tables(gold, monkeypatch):
    """N2 (diff_features.py:716-1120); the device step is SciPy's ranksums and the log2FC formula here."""
    from scipy.stats import ranksums

    def host_contrasts(matrix, contrasts_cols, device=0):
        n, R = len(contrasts_cols), matrix.shape[0]
        stat, pval, l2fc = (np.full((n, R), np.nan) for _ in range(3))
        for i, (f, b) in enumerate(contrasts_cols):
            fg, bg = np.asarray(matrix)[:, f], np.asarray(matrix)[:, b]
            pval[i] = [ranksums(fg[r], bg[r]).pvalue for r in range(R)]  # get_wilcox_test_pvalues, :968-1028
            l2fc[i] = np.log2((fg.mean(axis=1) + 10 ** -12) / (bg.mean(axis=1) + 10 ** -12))  # get_log2_fc, :1095-1120
        return stat, pval, l2fc

    monkeypatch.setattr(df, "ranksum_contrasts", host_contrasts)
    mtx, cells, regions, cell_data = hc.diff_inputs()
    obj = types.SimpleNamespace(cell_data=cell_data)
    for i, case in enumerate(hc.DIFF_CASES):
        got = df.find_diff_features(obj, df.CistopicImputedFeatures(mtx.copy(), list(regions), list(cells), "p"), "celltype",
                                    **hc.diff_kwargs(case, regions))
        assert list(got) == gold[f"diff{i}/names"].tolist()
        for name, table in got.items():
            hc.assert_frame_is(f"diff{i}/{name}", table, gold, rtol=1e-10)


def test_imputed_features_merge_and_subset(gold):
    """A16's result class (diff_features.py:61-272)."""
    for i, group in enumerate(hc.imputed_pairs()):
        objs = [df.CistopicImputedFeatures(mm.copy(), list(f), list(c), "p") for f, c, mm in group]
        merged = objs[0].merge(objs[1:], project="m", copy=True)
        order = np.argsort(merged.feature_names)
        assert np.array(merged.feature_names)[order].tolist() == gold[f"merge{i}/features"].tolist()
        assert merged.cell_names == gold[f"merge{i}/cells"].tolist() and merged.project == "m"
        want = gold[f"merge{i}/mtx"]
        assert merged.mtx.dtype == want.dtype and np.array_equal(np.asarray(merged.mtx)[order], want)
    feats, cl, mm = hc.imputed_pairs()[0][0]
    sub = df.CistopicImputedFeatures(mm.copy(), list(feats), list(cl), "p").subset(cells=cl[1:3], features=feats[0:30:3], copy=True)
    assert sub.feature_names == gold["subset/features"].tolist() and sub.cell_names == gold["subset/cells"].tolist()
    assert np.array_equal(sub.mtx, gold["subset/mtx"])


def test_fragments_reader(gold, tmp_path):
    """N4's reader (fragments.py:21-165)."""
    path = tmp_path / "fragments.tsv"
    path.write_text(hc.READER_TEXT)
    got = fm.read_fragments_from_file(str(path))
    assert list(got.columns) == gold["reader/columns"].tolist()
    for c, dtype in zip(got.columns, gold["reader/dtypes"].tolist()):
        want = gold[f"reader/{c}"]
        if dtype == "category":
            assert isinstance(got[c].dtype, pd.CategoricalDtype) and got[c].astype(str).tolist() == want.tolist()
        else:
            assert np.array_equal(got[c].to_numpy(), want) and (c == "Score" or str(got[c].dtype) == dtype)
