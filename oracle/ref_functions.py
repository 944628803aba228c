"""Load the reference's OWN functions for this path straight from the read-only tree.

TEST INFRASTRUCTURE ONLY.  ``pycisTopic`` cannot be imported here (its first third-party
imports fail), but the functions of interest are plain NumPy/``math`` code, so their
``FunctionDef`` nodes are extracted with ``ast`` and executed unmodified:

* ``utils.loglikelihood``            /root/reference/src/pycisTopic/utils.py:129-160
* ``utils.subset_list`` / ``non_zero_rows``                          utils.py:113-126
* ``calculate_imputed_accessibility`` (nested)   diff_features.py:398-492
* ``calculate_normalized_scores`` (nested)       diff_features.py:540-548
* ``CistopicImputedFeatures.make_rankings`` (method, run on a stand-in object)   diff_features.py:274-337
* ``find_highly_variable_features`` (matplotlib, absent here, replaced by a no-op stand-in; plot=False)
                                                 diff_features.py:577-714
* ``evaluate_models`` (model selection; matplotlib replaced by an inert stand-in)   lda_models.py:1048-1270
* ``LDAMallet.corpus_to_mallet`` (method)                                          lda_models.py:472-494
* ``markers`` / ``get_wilcox_test_pvalues`` / ``p_adjust_bh`` / ``get_log2_fc`` / ``mean_axis1`` /
  ``subset_array_second_axis``                   diff_features.py:839-1120 (decorators dropped: numba.prange
  becomes ``range``; ``np.asfarray``, removed in NumPy 2, is supplied as ``np.asarray(..., float)``)

Used in THIS container to generate ``tests/golden/*.npz`` (``tests/golden/make_golden.py``);
the GPU box has no /root/reference, so tests there read the committed fixtures.
"""
from __future__ import annotations

import ast
import logging
import math
import os
import typing

import numpy as np
import scipy.sparse as sparse

REFERENCE_ROOT = os.environ.get("PYCISTOPIC_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.exists(os.path.join(REFERENCE_ROOT, "src", "pycisTopic", "utils.py"))


def _find(tree, name):
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return node
    raise KeyError(name)


def _load(relpath, names, extra=None):
    path = os.path.join(REFERENCE_ROOT, "src", "pycisTopic", relpath)
    with open(path) as f:
        tree = ast.parse(f.read())
    ns = {"math": math, "np": np, "sparse": sparse, "typing": typing, "Union": typing.Union,
          "Sequence": typing.Sequence, "log": logging.getLogger("reference")}
    if extra:
        ns.update(extra)
    for name in names:
        node = _find(tree, name)
        node.decorator_list = []
        node.returns = None
        for a in node.args.args + node.args.kwonlyargs:
            a.annotation = None
        mod = ast.Module(body=[node], type_ignores=[])
        ast.fix_missing_locations(mod)
        exec(compile(mod, path, "exec"), ns)
    return ns


def reference_loglikelihood():
    return _load("utils.py", ["loglikelihood"])["loglikelihood"]


def reference_calculate_imputed_accessibility():
    ns = _load("utils.py", ["subset_list", "non_zero_rows"])
    ns = _load("diff_features.py", ["calculate_imputed_accessibility"],
               extra={"subset_list": ns["subset_list"], "non_zero_rows": ns["non_zero_rows"]})
    return ns["calculate_imputed_accessibility"]


def reference_calculate_normalized_scores():
    return _load("diff_features.py", ["calculate_normalized_scores"])["calculate_normalized_scores"]


class _NumpyWithAsfarray:
    """NumPy 2 dropped ``asfarray`` (used at diff_features.py:1035); everything else is NumPy's."""

    def __getattr__(self, name):
        if name == "asfarray":
            return lambda a: np.asarray(a, dtype=np.float64)
        return getattr(np, name)


def reference_diff_feature_functions():
    """{'markers', 'get_wilcox_test_pvalues', 'p_adjust_bh', 'get_log2_fc'}: the reference's own code
    (diff_features.py:839-1120) on top of the real ``scipy.stats.ranksums``."""
    import sys
    import types

    import pandas as pd
    from scipy.stats import ranksums

    ns = _load("utils.py", ["get_position_index"])
    extra = {"get_position_index": ns["get_position_index"], "ranksums": ranksums, "pd": pd, "sys": sys,
             "logging": logging, "numba": types.SimpleNamespace(prange=range), "np": _NumpyWithAsfarray(),
             "CistopicImputedFeatures": type("CistopicImputedFeatures", (), {})}
    return _load("diff_features.py", ["subset_array_second_axis", "mean_axis1", "get_log2_fc", "p_adjust_bh",
                                      "get_wilcox_test_pvalues", "markers"], extra=extra)


class _ImputedStandIn:
    """The four attributes ``make_rankings`` touches (diff_features.py:26-59)."""

    def __init__(self, imputed_acc, feature_names, cell_names, project):
        self.mtx, self.feature_names, self.cell_names, self.project = imputed_acc, feature_names, cell_names, project


def reference_make_rankings():
    """``make_rankings(obj, seed)`` -> ranking matrix: the reference's own method (diff_features.py:274-337)
    executed on a stand-in for ``CistopicImputedFeatures``."""
    ns = _load("diff_features.py", ["make_rankings"], extra={"CistopicImputedFeatures": _ImputedStandIn})
    method = ns["make_rankings"]

    def run(mtx, seed=123):
        obj = _ImputedStandIn(mtx, [f"r{i}" for i in range(mtx.shape[0])], [f"c{i}" for i in range(mtx.shape[1])], "ref")
        return method(obj, seed=seed).mtx

    return run


def reference_corpus_to_mallet():
    """``corpus_to_mallet(corpus, file_like)``: the reference's own method (lda_models.py:472-494), self unused."""
    from itertools import chain
    method = _load("lda_models.py", ["corpus_to_mallet"], extra={"chain": chain})["corpus_to_mallet"]
    return lambda corpus, file_like: method(None, corpus, file_like)


def reference_find_highly_variable_features():
    """The reference's own function (diff_features.py:577-714); ``plt`` / ``matplotlib`` are inert stand-ins (the
    function calls ``plt.figure()`` even with ``plot=False``)."""
    import sys
    import types

    import pandas as pd
    import sklearn.utils.sparsefuncs  # noqa: F401  (the sparse branch, :633)
    import sklearn

    inert = types.SimpleNamespace(figure=lambda *a, **k: None, scatter=lambda *a, **k: None,
                                  xlabel=lambda *a, **k: None, ylabel=lambda *a, **k: None, show=lambda *a, **k: None)
    extra = {"pd": pd, "sys": sys, "logging": logging, "sklearn": sklearn, "plt": inert,
             "matplotlib": types.SimpleNamespace(rcParams={}), "CistopicImputedFeatures": _ImputedStandIn}
    return _load("diff_features.py", ["find_highly_variable_features"], extra=extra)["find_highly_variable_features"]


class _Inert:
    """Accepts any attribute access and any call (stand-in for matplotlib.pyplot / matplotlib)."""

    def __getattr__(self, name):
        return _Inert()

    def __call__(self, *a, **k):
        return _Inert()


def reference_evaluate_models():
    """The reference's own ``evaluate_models`` (lda_models.py:1048-1270); plotting goes to an inert stand-in, so call
    it with ``plot=False`` or ``True`` alike."""
    import pandas as pd
    ns = _load("utils.py", ["subset_list"])
    extra = {"pd": pd, "plt": _Inert(), "matplotlib": _Inert(), "subset_list": ns["subset_list"],
             "CistopicLDAModel": object}
    return _load("lda_models.py", ["evaluate_models"], extra=extra)["evaluate_models"]


# ---- the count-matrix front end -------------------------------------------------------------------
class _PyRangesStandIn:
    """pyranges is absent: ``PyRanges(df)`` only has to hand the DataFrame back as ``.df``."""

    def __init__(self, df):
        self.df = df


def reference_read_fragments_to_pyranges():
    """fragments.py:21-165 with its "pandas" and "pyarrow" engines (both installed here; polars is not).  Returns a
    function ``(path, engine) -> DataFrame``."""
    import gzip
    import types

    import pandas as pd
    import pyarrow as pa
    import pyarrow.csv  # noqa: F401  (the reference reaches it as pa.csv)
    ns = _load("utils.py", ["normalise_filepath"], extra={"os": os})
    ns = _load("fragments.py", ["read_fragments_to_pyranges"],
               extra={"normalise_filepath": ns["normalise_filepath"], "gzip": gzip, "pd": pd, "pa": pa, "pl": None,
                      "pr": types.SimpleNamespace(PyRanges=_PyRangesStandIn), "Literal": typing.Literal})
    fn = ns["read_fragments_to_pyranges"]
    return lambda path, engine="pandas": fn(path, engine=engine).df


def reference_collapse_duplicates():
    """utils.py:284-293."""
    import pandas as pd
    return _load("utils.py", ["collapse_duplicates"], extra={"pd": pd})["collapse_duplicates"]


def reference_prepare_tag_cells():
    """utils.py:205-223."""
    import re
    return _load("utils.py", ["prepare_tag_cells"], extra={"re": re})["prepare_tag_cells"]


class _ObjectStandIn:
    """Records what ``create_cistopic_object`` hands to ``CistopicObject`` (cistopic_class.py:643-652)."""

    def __init__(self, fragment_matrix, binary_matrix, cell_names, region_names, cell_data, region_data,
                 path_to_fragments, project="cisTopic"):
        self.fragment_matrix, self.binary_matrix = fragment_matrix, binary_matrix
        self.cell_names, self.region_names = cell_names, region_names
        self.cell_data, self.region_data = cell_data, region_data
        self.path_to_fragments, self.project = path_to_fragments, project


def reference_create_cistopic_object(region_names_to_coordinates):
    """cistopic_class.py:508-655 without a blacklist (that step is pyranges).  ``region_names_to_coordinates`` of the
    reference is polars code (utils.py:60-102): the caller supplies an equivalent."""
    import sys

    import pandas as pd
    import sklearn.preprocessing as sp
    u = _load("utils.py", ["subset_list", "non_zero_rows"])
    ns = _load("cistopic_class.py", ["create_cistopic_object"],
               extra={"pd": pd, "sp": sp, "sys": sys, "logging": logging, "CistopicObject": _ObjectStandIn,
                      "subset_list": u["subset_list"], "non_zero_rows": u["non_zero_rows"],
                      "region_names_to_coordinates": region_names_to_coordinates})
    return ns["create_cistopic_object"]


def reference_create_cistopic_object_from_matrix_file(region_names_to_coordinates):
    """cistopic_class.py:656-735 on top of the reference's own ``create_cistopic_object``; the stand-in object gets the
    reference's ``add_cell_data``."""
    import sys

    import pandas as pd
    _ObjectStandIn.add_cell_data = reference_add_cell_data()
    ns = _load("cistopic_class.py", ["create_cistopic_object_from_matrix_file"],
               extra={"pd": pd, "sys": sys, "logging": logging,
                      "create_cistopic_object": reference_create_cistopic_object(region_names_to_coordinates)})
    return ns["create_cistopic_object_from_matrix_file"]


def reference_add_cell_data():
    """``CistopicObject.add_cell_data`` (cistopic_class.py:88-142) as a function of (object, cell_data, split_pattern)."""
    import pandas as pd
    ns = _load("cistopic_class.py", ["add_cell_data"],
               extra={"pd": pd, "prepare_tag_cells": reference_prepare_tag_cells()})
    return ns["add_cell_data"]


# ---- run_cgs_model itself (A2 / A14): hyper-parameter scaling, metric block, DataFrame packing ----------------------
def reference_run_cgs_model(lda_class, metric_functions):
    """``run_cgs_model`` (lda_models.py:178-396, the Ray decorator dropped) and ``CistopicLDAModel`` (:31-96) executed
    from the tree.  The two third-party packages it calls are absent and are stood in for: ``lda.LDA`` by ``lda_class``
    (the oracle's restatement, or a recorder) and the four ``tmtoolkit`` functions by ``metric_functions`` -- a dict
    ``metric_arun_2010`` / ``metric_cao_juan_2009`` / ``metric_coherence_mimno_2011`` / ``marginal_topic_distrib``.
    Everything else -- the scaling of alpha / eta, which arguments reach the metrics and ``utils.loglikelihood``, the
    seven DataFrames, the pickle -- is the reference's own code.  Returns ``(run_cgs_model, CistopicLDAModel)``."""
    import pickle
    import sys
    import time
    import types
    import warnings
    from itertools import chain

    import pandas as pd
    path = os.path.join(REFERENCE_ROOT, "src", "pycisTopic", "lda_models.py")
    with open(path) as f:
        tree = ast.parse(f.read())
    cls = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == "CistopicLDAModel")
    for node in ast.walk(cls):
        if isinstance(node, ast.FunctionDef):
            node.returns = None
            for a in node.args.args + node.args.kwonlyargs:
                a.annotation = None
    tm = types.SimpleNamespace(
        topicmod=types.SimpleNamespace(
            evaluate=types.SimpleNamespace(**{k: metric_functions[k] for k in (
                "metric_arun_2010", "metric_cao_juan_2009", "metric_coherence_mimno_2011")}),
            model_stats=types.SimpleNamespace(marginal_topic_distrib=metric_functions["marginal_topic_distrib"])))
    extra = {"pd": pd, "sys": sys, "time": time, "os": os, "pickle": pickle, "warnings": warnings, "logging": logging,
             "chain": chain, "lda": types.SimpleNamespace(LDA=lda_class), "tmtoolkit": tm,
             "loglikelihood": reference_loglikelihood()}
    # the class must be importable for the reference's pickle.dump (:394-395): give it a module of its own
    home = types.ModuleType("oracle._reference_lda_models")
    sys.modules[home.__name__] = home
    extra["__name__"] = home.__name__
    mod = ast.Module(body=[cls], type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, path, "exec"), extra)
    home.CistopicLDAModel = extra["CistopicLDAModel"]
    ns = _load("lda_models.py", ["run_cgs_model"], extra=extra)
    return ns["run_cgs_model"], ns["CistopicLDAModel"]


class _RayStandIn:
    """Ray is absent: tasks run where they are submitted (``f.remote(...)`` is the call, ``ray.get`` hands the list
    back); ``init`` records what it was given (``n_cpu`` and the forwarded keyword arguments, lda_models.py:154)."""

    def __init__(self):
        self.init_calls, self.shutdowns = [], 0

    def init(self, **kw):
        self.init_calls.append(kw)

    def get(self, results):
        return results

    def shutdown(self):
        self.shutdowns += 1


def reference_run_cgs_models(lda_class, metric_functions):
    """``run_cgs_models`` (lda_models.py:99-175) on top of the reference's own ``run_cgs_model``; returns
    ``(run_cgs_models, ray_stand_in)``."""
    import types
    run_one, _ = reference_run_cgs_model(lda_class, metric_functions)
    ray = _RayStandIn()
    ns = _load("lda_models.py", ["run_cgs_models"],
               extra={"ray": ray, "run_cgs_model": types.SimpleNamespace(remote=run_one)})
    return ns["run_cgs_models"], ray


def reference_imputed_features_class():
    """``CistopicImputedFeatures`` (diff_features.py:26-337: constructor, ``subset``, ``merge``, ``make_rankings``) and the
    outer ``impute_accessibility`` (:340-508) executed from the tree; returns ``(class, impute_accessibility)``."""
    import sys

    import pandas as pd
    path = os.path.join(REFERENCE_ROOT, "src", "pycisTopic", "diff_features.py")
    with open(path) as f:
        tree = ast.parse(f.read())
    cls = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == "CistopicImputedFeatures")
    for node in ast.walk(cls):
        if isinstance(node, ast.FunctionDef):
            node.returns = None
            for a in node.args.args + node.args.kwonlyargs:
                a.annotation = None
    u = _load("utils.py", ["subset_list", "non_zero_rows", "get_position_index"])
    extra = {"pd": pd, "sys": sys, "logging": logging, "prepare_tag_cells": reference_prepare_tag_cells(),
             "subset_list": u["subset_list"], "non_zero_rows": u["non_zero_rows"],
             "get_position_index": u["get_position_index"], "__name__": "oracle._reference_diff_features"}
    ns = {"math": math, "np": np, "sparse": sparse, "typing": typing}
    ns.update(extra)
    mod = ast.Module(body=[cls], type_ignores=[])
    ast.fix_missing_locations(mod)
    exec(compile(mod, path, "exec"), ns)
    extra["CistopicImputedFeatures"] = ns["CistopicImputedFeatures"]
    fn = _load("diff_features.py", ["impute_accessibility"], extra=extra)["impute_accessibility"]
    return ns["CistopicImputedFeatures"], fn


def reference_find_diff_features():
    """``find_diff_features`` (diff_features.py:716-836) with the reference's own ``markers`` chain beneath it (on the real
    ``scipy.stats.ranksums``) and its own ``CistopicImputedFeatures``; Ray is not reached for ``n_cpu=1`` and is the
    direct-call stand-in otherwise.  Returns ``(find_diff_features, CistopicImputedFeatures, helper functions)``."""
    import sys
    import types

    import pandas as pd
    from scipy.stats import ranksums
    cls, _ = reference_imputed_features_class()
    u = _load("utils.py", ["get_position_index"])
    extra = {"get_position_index": u["get_position_index"], "ranksums": ranksums, "pd": pd, "sys": sys,
             "logging": logging, "numba": types.SimpleNamespace(prange=range), "np": _NumpyWithAsfarray(),
             "CistopicImputedFeatures": cls, "ray": _RayStandIn()}
    ns = _load("diff_features.py", ["subset_array_second_axis", "mean_axis1", "get_log2_fc", "p_adjust_bh",
                                    "get_wilcox_test_pvalues", "markers", "find_diff_features"], extra=extra)
    return ns["find_diff_features"], cls, ns
