"""Generates tests/golden/*.npz by executing the REFERENCE's own functions (AST-extracted from
/root/reference, see oracle/ref_functions.py) and the CPU oracle on small seeded inputs.
Run in the authoring container (needs /root/reference):   python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import lda_oracle as lo  # noqa: E402
from oracle import ref_functions as rf  # noqa: E402


def golden_loglikelihood():
    """utils.loglikelihood (utils.py:129-160) on random count tables, incl. empty rows/topics."""
    f = rf.reference_loglikelihood()
    rng = np.random.default_rng(20261019)
    out = {}
    cases = [(3, 7, 5, 50.0, 0.1), (10, 40, 12, 50.0, 0.1), (5, 30, 9, 0.7, 0.02), (2, 11, 4, 25.0, 0.1),
             (6, 25, 8, 50.0, 0.1)]
    for i, (K, W, D, alpha, eta) in enumerate(cases):
        nzw = rng.integers(0, 6, (K, W)) * (rng.random((K, W)) < 0.5)
        ndz = rng.integers(0, 5, (D, K)) * (rng.random((D, K)) < 0.6)
        if i == 2:
            nzw[-1] = 0  # empty last topic
        if i == 3:
            ndz[-1] = 0  # empty last cell
        if i == 4:
            nzw[:] = 0
            nzw[1, 3] = 4  # a single non-empty topic followed by empty ones
        out[f"nzw{i}"], out[f"ndz{i}"] = nzw.astype(np.int32), ndz.astype(np.int32)
        out[f"par{i}"] = np.array([alpha, eta])
        out[f"ll{i}"] = np.array(f(nzw, ndz, alpha, eta))
    out["n"] = np.array(len(cases))
    np.savez(os.path.join(HERE, "loglikelihood_compat.npz"), **out)


def golden_impute():
    """calculate_imputed_accessibility (diff_features.py:398-492) and calculate_normalized_scores (:540-548)."""
    f = rf.reference_calculate_imputed_accessibility()
    g = rf.reference_calculate_normalized_scores()
    rng = np.random.default_rng(4321)
    out = {}
    cases = [(257, 10, 33, 10 ** 6, 64), (1000, 37, 129, 10 ** 6, 300), (500, 100, 64, 10 ** 6, 20000),
             (300, 5, 50, 1, 100), (300, 5, 50, 1e6, 100), (640, 16, 256, 10 ** 4, 128)]
    for i, (R, K, C, scale, chunk) in enumerate(cases):
        a = rng.dirichlet(np.full(R, 0.1), size=K).T.astype(np.float32)  # every topic sums to 1 over regions
        b = rng.dirichlet(np.full(K, 50.0 / K), size=C).T.astype(np.float32)  # every cell sums to 1 over topics
        if i == 1:
            a[10:20] = 0  # all-zero regions must be dropped
        names = [f"chr1:{j * 500}-{j * 500 + 500}" for j in range(R)]
        m, kept = f(a, b, names, scale, chunk)
        out[f"a{i}"], out[f"b{i}"] = a, b
        out[f"scale{i}"] = np.array(scale, dtype=object if False else np.float64)
        out[f"scale_is_int{i}"] = np.array(isinstance(scale, int))
        out[f"chunk{i}"] = np.array(chunk)
        out[f"out{i}"] = m
        out[f"kept{i}"] = np.array([names.index(k) for k in kept], np.int64)
        if i in (0, 5):
            out[f"norm{i}"] = g(m.astype(np.float64) if m.dtype != np.float64 else m, 10 ** 4)
    out["n"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "impute.npz"), **out)


def golden_diff_features():
    """markers / get_wilcox_test_pvalues / get_log2_fc / p_adjust_bh (diff_features.py:839-1120) executed from the
    reference tree on top of the real scipy.stats.ranksums."""
    import pandas as pd
    ns = rf.reference_diff_feature_functions()
    rng = np.random.default_rng(716)
    out = {}
    # (rows, n_fg, n_bg, dtype, kind): ties everywhere (small ints), continuous, negative values and zeros,
    # a foreground larger than the background, a one-cell group
    cases = [(64, 17, 40, np.int32, "ties"), (50, 33, 9, np.float32, "cont"), (40, 12, 31, np.float64, "signed"),
             (30, 1, 25, np.int32, "ties"), (25, 60, 60, np.int32, "wide")]
    for i, (R, n1, n2, dt, kind) in enumerate(cases):
        if kind == "ties":
            fg, bg = rng.integers(0, 4, (R, n1)), rng.integers(0, 5, (R, n2))
        elif kind == "wide":
            fg, bg = rng.integers(0, 10 ** 6, (R, n1)), rng.integers(0, 10 ** 6, (R, n2))
            fg[:5] += 200000  # clearly shifted rows
        elif kind == "cont":
            fg, bg = rng.gamma(2.0, 1.0, (R, n1)), rng.gamma(2.2, 1.0, (R, n2))
        else:
            fg, bg = np.round(rng.normal(0, 1, (R, n1)), 1), np.round(rng.normal(0.3, 1, (R, n2)), 1)
            fg[0, :] = 0.0
            bg[0, :3] = -0.0
        fg, bg = np.ascontiguousarray(fg.astype(dt)), np.ascontiguousarray(bg.astype(dt))
        p = np.array(ns["get_wilcox_test_pvalues"](fg, bg), dtype=np.float64)
        out[f"fg{i}"], out[f"bg{i}"], out[f"p{i}"] = fg, bg, p
        out[f"l2fc{i}"] = np.asarray(ns["get_log2_fc"](fg, bg), dtype=np.float64)
        out[f"padj{i}"] = ns["p_adjust_bh"](p)
    # markers() on a DataFrame: one matrix, two barcode lists in scrambled order
    R, C = 120, 90
    mat = rng.integers(0, 30, (R, C)).astype(np.int32)
    mat[:15, :30] += rng.integers(20, 60, (15, 30)).astype(np.int32)  # 15 true markers of the first 30 cells
    cells = [f"cell{j}" for j in range(C)]
    feats = [f"chr1:{j * 500}-{j * 500 + 500}" for j in range(R)]
    fg_names = [cells[j] for j in rng.permutation(30)]
    bg_names = [cells[j] for j in 30 + rng.permutation(60)[:50]]
    df = pd.DataFrame(mat, index=feats, columns=cells)
    res = ns["markers"](df, [fg_names, bg_names], "A_VS_B", adjpval_thr=0.05, log2fc_thr=np.log2(1.5), n_cpu=1)
    out["m_mat"], out["m_fg"], out["m_bg"] = mat, np.array([cells.index(c) for c in fg_names]), \
        np.array([cells.index(c) for c in bg_names])
    out["m_index"] = np.array([feats.index(f) for f in res.index])
    out["m_log2fc"], out["m_padj"] = res["Log2FC"].to_numpy(), res["Adjusted_pval"].to_numpy()
    out["n"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "diff_features.npz"), **out)


def golden_oracle_kat():
    """Known-answer vectors frozen from the CPU oracle (pins the oracle against regressions;
    parity with the real `lda` package stays UNPINNED, see oracle/lda_oracle.py)."""
    X = sp.csr_matrix(np.array([[1, 1, 0, 1, 0, 1],
                                [0, 1, 1, 0, 1, 0],
                                [1, 0, 1, 1, 1, 1],
                                [0, 1, 0, 1, 0, 1]], dtype=np.int32))
    K, alpha, eta, seed = 2, 0.5, 0.1, 7
    o = lo.OracleLDA(K, n_iter=0, alpha=alpha, eta=eta, random_state=seed, keep_z=True)
    o._initialize(X)
    rs = np.random.RandomState(seed)
    rands = o._rands.copy()
    zs, lls = [o.ZS.copy()], [o.loglikelihood()]
    for _ in range(3):
        rs.shuffle(rands)
        o._sample_topics(rands)
        zs.append(o.ZS.copy())
        lls.append(o.loglikelihood())
    out = {"X": X.toarray(), "K": np.array(K), "alpha": np.array(alpha), "eta": np.array(eta), "seed": np.array(seed),
           "WS": o.WS, "DS": o.DS, "z": np.array(zs), "ll": np.array(lls), "nzw": o.nzw_, "ndz": o.ndz_, "nz": o.nz_}
    # a C1-like small trace
    from pycistopic_b200.synth import make_binary_matrix
    m, _, _ = make_binary_matrix(n_cells=60, n_regions=800, density=0.05, n_true_topics=4, seed=3)
    Xs = sp.csr_matrix(m.T)
    o2 = lo.OracleLDA(5, n_iter=20, alpha=10.0, eta=0.1, random_state=555, refresh=1).fit(Xs)
    out["trace_X_indptr"], out["trace_X_indices"] = Xs.indptr, Xs.indices
    out["trace_shape"] = np.array(Xs.shape)
    out["trace_ll"] = np.array(o2.loglikelihoods_ + [o2.final_loglikelihood_])
    out["trace_nz"] = o2.nz_
    out["trace_mimno"] = lo.metric_coherence_mimno_2011(o2.topic_word_, Xs)
    out["trace_arun"] = np.array(lo.metric_arun_2010(o2.topic_word_, o2.doc_topic_,
                                                      np.asarray(Xs.sum(axis=1)).astype(float)))
    out["trace_cao"] = np.array(lo.metric_cao_juan_2009(o2.topic_word_))
    np.savez_compressed(os.path.join(HERE, "oracle_kat.npz"), **out)


def golden_rankings():
    """make_rankings of the reference (diff_features.py:274-337, executed from the tree) on small matrices: columns
    without ties (unique ranking), columns with ties (any valid draw), float scores."""
    run = rf.reference_make_rankings()
    rng = np.random.default_rng(2024)
    R, C = 700, 6
    x = np.empty((R, C), np.int32)
    x[:, 0] = rng.permutation(R)                       # no ties
    x[:, 1] = rng.permutation(R) * 3 - 500             # no ties, negative values
    x[:, 2] = rng.integers(0, 6, R)                    # heavy ties
    x[:, 3] = 0                                        # constant
    x[:, 4] = (rng.random(R) < 0.1) * rng.integers(1, 1000, R)   # mostly zero
    x[:, 5] = rng.permutation(R) + 70000               # no ties, needs three digit passes
    xf = rng.normal(0, 1, (R, 3)).astype(np.float32)   # no ties with probability 1
    xf[:, 2] = np.round(xf[:, 2], 1)                   # ties
    # the reference method only accepts a SciPy sparse mtx (it calls .toarray() on every column, :334)
    out = {"x_int": x, "rank_int": run(sp.csr_matrix(x), seed=123), "x_float": xf,
           "rank_float": run(sp.csr_matrix(xf), seed=7)}
    np.savez_compressed(os.path.join(HERE, "rankings.npz"), **out)


def golden_variable_features():
    """find_highly_variable_features of the reference (diff_features.py:577-714, executed from the tree) on a
    normalised-imputation-shaped float64 matrix, an int32 imputed one and a float32 one: NumPy's float32 mean / var
    per region and the selected features for the default window and for n_top_features."""
    import pandas as pd
    fn = rf.reference_find_highly_variable_features()
    rng = np.random.default_rng(77)
    R, C, K = 900, 160, 8
    tr = rng.dirichlet(np.full(R, 0.3), K).T
    ct = rng.dirichlet(np.full(K, 0.4), C).T
    imputed = (tr @ ct * 1e6).astype(np.int32)
    imputed = imputed[imputed.any(axis=1)]
    norm64 = np.log1p(imputed / (imputed.sum(axis=0) / 10 ** 4))
    out = {"imputed": imputed}   # norm64 = np.log1p(imputed / (imputed.sum(axis=0) / 10 ** 4)), recomputed by the tests
    for name, mat in (("norm64", norm64), ("imputed", imputed), ("norm32", norm64.astype(np.float32))):
        frame = pd.DataFrame(mat, index=[f"r{i}" for i in range(mat.shape[0])])
        out[f"{name}_mean"] = np.mean(mat, axis=1, dtype=np.float32)
        out[f"{name}_var"] = np.var(mat, axis=1, dtype=np.float32)
        out[f"{name}_default"] = np.array(fn(frame, plot=False), dtype=str)
        out[f"{name}_top100"] = np.array(fn(frame, n_top_features=100, plot=False), dtype=str)
        out[f"{name}_window"] = np.array(fn(frame, min_disp=0.5, min_mean=0.05, max_mean=2.0, max_disp=3.0, n_bins=10,
                                            plot=False), dtype=str)
    np.savez_compressed(os.path.join(HERE, "variable_features.npz"), **out)


if __name__ == "__main__":
    if not rf.available():
        sys.exit("the reference tree is not available: cannot regenerate the golden fixtures")
    golden_loglikelihood()
    golden_impute()
    golden_oracle_kat()
    golden_diff_features()
    golden_rankings()
    golden_variable_features()
    print("golden fixtures written to", HERE)
