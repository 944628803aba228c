"""B200-native drop-in for pycisTopic's collapsed-Gibbs topic-modelling hot path.

Public surface (same names as the reference's ``pycisTopic.lda_models`` /
``pycisTopic.diff_features``): ``run_cgs_models``, ``run_cgs_model``, ``evaluate_models``,
``CistopicLDAModel``, ``load_cisTopic_model``, ``impute_accessibility``, ``normalize_scores``,
``CistopicImputedFeatures``; plus ``LDA`` (the ``lda.LDA`` object protocol on the GPU).
All compute goes through ``libcgs_b200.so`` (``include/cgs_b200.h``); there is no CPU path.
"""
from .diff_features import (CistopicImputedFeatures, FactoredImputedFeatures, FactoredNormalizedFeatures, find_diff_features,
                            find_highly_variable_features, impute_accessibility, markers, normalize_scores)
from .lda import LDA, Corpus, DeviceModel
from .lda_models import (CistopicLDAModel, evaluate_models, load_cisTopic_model, run_cgs_model,
                         run_cgs_models)

__version__ = "0.1.0"
__all__ = [
    "LDA", "Corpus", "DeviceModel", "CistopicLDAModel", "run_cgs_models", "run_cgs_model", "evaluate_models",
    "load_cisTopic_model", "impute_accessibility", "normalize_scores", "CistopicImputedFeatures",
    "FactoredImputedFeatures", "FactoredNormalizedFeatures", "markers", "find_diff_features",
    "find_highly_variable_features",
]
