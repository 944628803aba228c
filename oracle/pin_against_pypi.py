#!/usr/bin/env python
"""Pin the CPU oracle against the REAL third-party packages pycisTopic's hot path delegates to
(TEST INFRASTRUCTURE; run it on any machine where they can be installed):

    pip install lda tmtoolkit          # unpinned in the reference: pyproject.toml:42,44, requirements.txt:13,48
    python oracle/pin_against_pypi.py  # exit 0 = pinned, 3 = packages absent (parity stays UNPINNED), 1 = mismatch

What is compared, on seeded planted-topic matrices (the C1-shaped toy and two tiny ones), for the call sequence of
``run_cgs_model`` (lda_models.py:248-305: ``lda.LDA(n_topics, n_iter, random_state, alpha, eta, refresh)``, ``fit``,
the three tmtoolkit metrics, ``marginal_topic_distrib``):

* ``lda.utils.matrix_to_lists``                     == ``oracle.matrix_to_lists``           (bit-exact WS, DS)
* ``lda.LDA.fit``: ``nzw_``, ``ndz_``, ``nz_``       == ``OracleLDA`` with the same seed     (bit-exact: the sampler is
  deterministic given ``random_state``, so equal counts after every tested number of sweeps mean an identical z
  trajectory), ``loglikelihoods_`` / ``loglikelihood()`` to 1e-12, ``topic_word_`` / ``doc_topic_`` to 1e-12
* ``tmtoolkit.topicmod.evaluate.metric_arun_2010`` / ``metric_cao_juan_2009`` / ``metric_coherence_mimno_2011`` and
  ``tmtoolkit.topicmod.model_stats.marginal_topic_distrib`` == the oracle's restatements to 1e-10

On success it writes ``tests/golden/pypi_pin.json`` (package versions, what matched) -- commit that file and drop
the "PARITY UNPINNED" notes in oracle/lda_oracle.py, oracle/cgs_oracle.c and DESIGN.md.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

EXIT_PINNED, EXIT_MISMATCH, EXIT_ABSENT = 0, 1, 3


def packages():
    """(lda module, tmtoolkit evaluate module, tmtoolkit model_stats module) or None when one is missing."""
    try:
        import lda  # noqa: F401
        import lda.utils  # noqa: F401
        from tmtoolkit.topicmod import evaluate, model_stats
    except Exception as e:  # ImportError, or a broken optional dependency of tmtoolkit
        return None, repr(e)
    return (lda, evaluate, model_stats), None


def cases():
    from pycistopic_b200.synth import make_binary_matrix
    yield "tiny", make_binary_matrix(n_cells=30, n_regions=400, density=0.08, n_true_topics=3, seed=5)[0], [2, 5], [1, 3, 12]
    yield "small", make_binary_matrix(n_cells=300, n_regions=5000, density=0.04, n_true_topics=6, seed=11)[0], [10, 40], [25]
    yield "C1", make_binary_matrix("C1")[0], [2, 5, 10], [150]


def run(verbose=True):
    pk, why = packages()
    if pk is None:
        if verbose:
            print(f"lda / tmtoolkit not importable ({why}): the sampler oracle stays PARITY UNPINNED")
        return EXIT_ABSENT, {}
    lda, evaluate, model_stats = pk
    import logging

    from oracle import lda_oracle as lo
    logging.getLogger("lda").setLevel(logging.WARNING)
    report = {"lda": getattr(lda, "__version__", "?"),
              "tmtoolkit": getattr(__import__("tmtoolkit"), "__version__", "?"), "cases": []}
    ok = True

    def check(name, cond):
        nonlocal ok
        ok = ok and bool(cond)
        if verbose:
            print(f"  {'ok ' if cond else 'MISMATCH'}  {name}", flush=True)
        return bool(cond)

    for label, m, topics, iters in cases():
        X = sp.csr_matrix(m.T)
        X.sort_indices()
        WS, DS = lda.utils.matrix_to_lists(X)
        oWS, oDS = lo.matrix_to_lists(X)
        check(f"{label}: matrix_to_lists", np.array_equal(WS, oWS) and np.array_equal(DS, oDS))
        for K in topics:
            for n_iter in iters:
                alpha, eta = 50.0 / K, 0.1
                ref = lda.LDA(n_topics=K, n_iter=n_iter, random_state=555, alpha=alpha, eta=eta, refresh=1).fit(X)
                o = lo.OracleLDA(K, n_iter=n_iter, alpha=alpha, eta=eta, random_state=555, refresh=1).fit(X)
                tag = f"{label} K={K} n_iter={n_iter}"
                same = (np.array_equal(ref.nzw_, o.nzw_) and np.array_equal(ref.ndz_, o.ndz_)
                        and np.array_equal(ref.nz_, o.nz_))
                check(f"{tag}: nzw_/ndz_/nz_ bit-identical", same)
                check(f"{tag}: log-likelihood trace", np.allclose(ref.loglikelihoods_, o.loglikelihoods_, rtol=1e-12)
                      and np.isclose(ref.loglikelihood(), o.final_loglikelihood_, rtol=1e-12))
                check(f"{tag}: topic_word_/doc_topic_", np.allclose(ref.topic_word_, o.topic_word_, rtol=1e-12)
                      and np.allclose(ref.doc_topic_, o.doc_topic_, rtol=1e-12))
                doc_len = np.asarray(X.sum(axis=1)).astype(float)
                a = evaluate.metric_arun_2010(ref.topic_word_, ref.doc_topic_, doc_len)
                check(f"{tag}: metric_arun_2010", np.isclose(a, lo.metric_arun_2010(o.topic_word_, o.doc_topic_, doc_len), rtol=1e-10))
                c = evaluate.metric_cao_juan_2009(ref.topic_word_)
                check(f"{tag}: metric_cao_juan_2009", np.isclose(c, lo.metric_cao_juan_2009(o.topic_word_), rtol=1e-10))
                mim = evaluate.metric_coherence_mimno_2011(ref.topic_word_, dtm=X, top_n=20, eps=1e-12, normalize=True,
                                                           return_mean=False)
                check(f"{tag}: metric_coherence_mimno_2011", np.allclose(mim, lo.metric_coherence_mimno_2011(o.topic_word_, X), rtol=1e-10))
                marg = model_stats.marginal_topic_distrib(ref.doc_topic_, doc_len.ravel())
                check(f"{tag}: marginal_topic_distrib", np.allclose(np.asarray(marg).ravel(),
                                                                     lo.marginal_topic_distrib(o.doc_topic_, doc_len), rtol=1e-12))
                report["cases"].append({"case": tag, "identical_counts": bool(same)})
    report["pinned"] = bool(ok)
    if ok:
        with open(os.path.join(ROOT, "tests", "golden", "pypi_pin.json"), "w") as f:
            json.dump(report, f, indent=1)
        if verbose:
            print("PINNED: the oracle reproduces lda", report["lda"], "and tmtoolkit", report["tmtoolkit"],
                  "-> tests/golden/pypi_pin.json")
    return (EXIT_PINNED if ok else EXIT_MISMATCH), report


if __name__ == "__main__":
    sys.exit(run()[0])
