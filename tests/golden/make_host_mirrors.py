"""Generates tests/golden/host_mirrors.npz: what the REFERENCE's own host functions return on the seeded inputs of
host_mirror_cases.py -- run_cgs_model (lda_models.py:178-396), create_cistopic_object / add_cell_data
(cistopic_class.py:508-655, :88-142), find_diff_features with its markers chain (diff_features.py:716-1120),
CistopicImputedFeatures.subset / merge (:61-272) and read_fragments_to_pyranges (fragments.py:21-165) -- executed from
/root/reference (oracle/ref_functions.py).  Run in the authoring container:  python tests/golden/make_host_mirrors.py"""
import os
import sys
import tempfile

import numpy as np
import pandas as pd
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import host_mirror_cases as hc  # noqa: E402
from oracle import lda_oracle as lo  # noqa: E402
from oracle import ref_functions as rf  # noqa: E402
from pycistopic_b200 import fragment_matrix as fm  # noqa: E402  (region_names_to_coordinates: the reference's is polars)


def main():
    out = {}
    # ---- run_cgs_model: scaling of alpha / eta, metric block, the seven DataFrames --------------------------------
    made = []

    class Recorder(lo.OracleLDA):
        def __init__(self, **kw):
            made.append(dict(kw))
            super().__init__(**kw)
            made[-1]["fitted"] = self

    fns = dict(metric_arun_2010=lo.metric_arun_2010, metric_cao_juan_2009=lo.metric_cao_juan_2009,
               metric_coherence_mimno_2011=lo.metric_coherence_mimno_2011,
               marginal_topic_distrib=lambda dt, dl: np.asmatrix(lo.marginal_topic_distrib(dt, dl)).T)
    ref_run, _ = rf.reference_run_cgs_model(Recorder, fns)
    m, cells, regions = hc.packing_matrix()
    X = sp.csr_matrix(m.T)
    for i, (K, top, a_by, e_by) in enumerate(hc.PACKING_CASES):
        ref = ref_run(X, K, cells, regions, n_iter=12, random_state=555, alpha=50, alpha_by_topic=a_by, eta=0.1,
                      eta_by_topic=e_by, top_topics_coh=top, save_path=None)
        o = made[-1].pop("fitted")
        for name in ("nzw_", "ndz_", "nz_", "topic_word_", "doc_topic_"):
            out[f"packing{i}/{name}"] = getattr(o, name)
        out[f"packing{i}/sampler_alpha_eta"] = np.array([made[-1]["alpha"], made[-1]["eta"]])
        out[f"packing{i}/sampler_ints"] = np.array([made[-1]["n_topics"], made[-1]["n_iter"], made[-1]["random_state"],
                                                    made[-1]["refresh"]])
        ref.parameters.loc["time", "Parameter"] = 0.0
        for frame in ("metrics", "coherence", "marg_topic", "topic_ass", "cell_topic", "topic_region", "parameters"):
            hc.pack_frame(f"packing{i}/{frame}", getattr(ref, frame), out)
    # ---- create_cistopic_object, add_cell_data --------------------------------------------------------------------
    ref_create = rf.reference_create_cistopic_object(fm.region_names_to_coordinates)
    ref_add = rf.reference_add_cell_data()
    cm, ccells, cregions = hc.count_matrix()
    for i, kw in enumerate(hc.OBJECT_CASES):
        ref = ref_create(cm.copy(), cell_names=list(ccells), region_names=list(cregions), project="s1", **kw)
        out[f"object{i}/cell_names"], out[f"object{i}/region_names"] = np.array(ref.cell_names), np.array(ref.region_names)
        out[f"object{i}/binary"], out[f"object{i}/fragments"] = ref.binary_matrix.toarray(), ref.fragment_matrix.toarray()
        hc.pack_frame(f"object{i}/cell_data", ref.cell_data, out)
        hc.pack_frame(f"object{i}/region_data", ref.region_data, out)
    for name, meta in hc.cell_metadata(ccells).items():
        ref = ref_create(cm.copy(), cell_names=list(ccells), region_names=list(cregions), project="s1")
        ref_add(ref, meta.copy(), "___")
        hc.pack_frame(f"add_cell_data/{name}", ref.cell_data.sort_index(axis=1), out)
    # ---- find_diff_features ---------------------------------------------------------------------------------------
    ref_find, ref_class, _ = rf.reference_find_diff_features()
    mtx, dcells, dregions, cell_data = hc.diff_inputs()

    class Obj:
        pass
    obj = Obj()
    obj.cell_data = cell_data
    for i, case in enumerate(hc.DIFF_CASES):
        want = ref_find(obj, ref_class(mtx.copy(), list(dregions), list(dcells), "p"), "celltype", **hc.diff_kwargs(case, dregions))
        out[f"diff{i}/names"] = np.array(list(want))
        for name, table in want.items():
            hc.pack_frame(f"diff{i}/{name}", table, out)
    # ---- CistopicImputedFeatures.merge / subset -------------------------------------------------------------------
    for i, group in enumerate(hc.imputed_pairs()):
        objs = [ref_class(mm.copy(), list(f), list(c), "p") for f, c, mm in group]
        merged = objs[0].merge(objs[1:], project="m", copy=True)
        order = np.argsort(merged.feature_names)
        out[f"merge{i}/features"] = np.array(merged.feature_names)[order]
        out[f"merge{i}/cells"] = np.array(merged.cell_names)
        out[f"merge{i}/mtx"] = np.asarray(merged.mtx)[order]
    feats, cl, mm = hc.imputed_pairs()[0][0]
    sub = ref_class(mm.copy(), list(feats), list(cl), "p").subset(cells=cl[1:3], features=feats[0:30:3], copy=True)
    out["subset/features"], out["subset/cells"], out["subset/mtx"] = np.array(sub.feature_names), np.array(sub.cell_names), sub.mtx
    # ---- read_fragments_to_pyranges -------------------------------------------------------------------------------
    ref_read = rf.reference_read_fragments_to_pyranges()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "fragments.tsv")
        with open(path, "w") as f:
            f.write(hc.READER_TEXT)
        frame = ref_read(path, "pandas")
    out["reader/columns"] = np.array(list(frame.columns))
    out["reader/dtypes"] = np.array([str(t) for t in frame.dtypes])
    for c in frame.columns:
        out[f"reader/{c}"] = frame[c].astype(str).to_numpy().astype(str) if isinstance(frame[c].dtype, pd.CategoricalDtype) else frame[c].to_numpy()
    np.savez_compressed(os.path.join(HERE, "host_mirrors.npz"), **out)
    print("wrote host_mirrors.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
