"""Pickle interoperability with the reference package.

pycisTopic pickles its models by qualified class name (``pickle.dump(model, ...)`` at
lda_models.py:394-395 and cli/subcommand/topic_modeling.py:69-70), so a ``Topic{K}.pkl`` written here
only loads in a pycisTopic environment -- and one written there only loads here -- if the class is
importable as ``pycisTopic.lda_models.CistopicLDAModel``.  :func:`alias_pycistopic` arranges that.
"""
from __future__ import annotations

import importlib
import sys
import types


def alias_pycistopic(force: bool = False) -> bool:
    """Make ``pycisTopic.lda_models`` / ``pycisTopic.diff_features`` / ``pycisTopic.cistopic_class`` resolve to this
    package's classes.

    * pycisTopic is NOT importable (or ``force``): stub modules are registered in ``sys.modules`` and
      :class:`CistopicLDAModel` / :class:`CistopicImputedFeatures` / :class:`CistopicObject` take the reference's
      module names, so pickles written from now on carry ``pycisTopic.lda_models.CistopicLDAModel`` and reference
      pickles -- the cisTopic object the CLI reads (``pycisTopic.cistopic_class.CistopicObject``) included -- load
      as this package's (attribute-for-attribute identical) classes.  Returns True.
    * pycisTopic IS importable: nothing is touched (its own classes load its own pickles); models are handed to
      it through :func:`to_reference_model` (:func:`pickle_ready` does that before every ``pickle.dump`` of this
      package).  Returns False.
    """
    from . import diff_features, fragment_matrix, lda_models
    if not force:
        try:
            found = importlib.import_module("pycisTopic.lda_models")
            return bool(getattr(found, "__cgs_stub__", False))  # our own stub from an earlier call: still stub mode
        except Exception:
            pass
    root = sys.modules.get("pycisTopic")
    if root is None or force:
        root = types.ModuleType("pycisTopic")
        root.__path__ = []  # a package, so that submodule imports resolve through sys.modules
        sys.modules["pycisTopic"] = root
    for name, mod in (("lda_models", lda_models), ("diff_features", diff_features), ("cistopic_class", fragment_matrix)):
        stub = types.ModuleType("pycisTopic." + name)
        stub.__cgs_stub__ = True
        for attr in getattr(mod, "__all__", None) or [a for a in dir(mod) if not a.startswith("_")]:
            setattr(stub, attr, getattr(mod, attr))
        sys.modules["pycisTopic." + name] = stub
        setattr(root, name, stub)
    lda_models.CistopicLDAModel.__module__ = "pycisTopic.lda_models"
    diff_features.CistopicImputedFeatures.__module__ = "pycisTopic.diff_features"
    fragment_matrix.CistopicObject.__module__ = "pycisTopic.cistopic_class"
    return True


def pickle_ready(obj):
    """What to hand to ``pickle.dump`` so that a plain pycisTopic environment can load the file
    (``Topic{K}.pkl`` of ``run_cgs_models(save_path=...)``, lda_models.py:390-395; the model list of the CLI,
    cli/subcommand/topic_modeling.py:69-70).  pycisTopic installed: models are re-wrapped in ITS class.  Not
    installed: the stub modules are registered, after which this package's class pickles under the reference's
    qualified name.  Lists are converted element-wise; anything else is returned as is."""
    from . import lda_models
    if isinstance(obj, (list, tuple)):
        return type(obj)(pickle_ready(o) for o in obj)
    if not isinstance(obj, lda_models.CistopicLDAModel):
        return obj
    if alias_pycistopic():
        return obj
    return to_reference_model(obj)


def to_reference_model(model):
    """Re-wrap a :class:`pycistopic_b200.CistopicLDAModel` in the installed pycisTopic's own class (same seven
    constructor arguments, lda_models.py:78-96) so that ``cistopic_obj.add_LDA_model`` / pickling use it."""
    ref = importlib.import_module("pycisTopic.lda_models").CistopicLDAModel
    out = ref(model.metrics, model.coherence, model.marg_topic, model.topic_ass, model.cell_topic,
              model.topic_region, model.parameters)
    out.cell_topic_harmony = model.cell_topic_harmony
    for extra in ("loglikelihood_true",):  # attributes this package adds; harmless for the reference
        if hasattr(model, extra):
            setattr(out, extra, getattr(model, extra))
    return out
