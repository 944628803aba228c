"""Model-quality metrics stored in ``CistopicLDAModel.metrics`` / ``coherence`` / ``marg_topic``.

The reference takes three of them from ``tmtoolkit`` (lda_models.py:291-304, :332-344) and the
fourth from its own ``utils.loglikelihood`` (utils.py:129-160, called at lda_models.py:305).
They are O(K^2 W) / O(K W) host post-processing on the exported distributions, not the hot
path; the formulas are restated here for speed and to remove the tmtoolkit dependency.
"""
from __future__ import annotations

import math

import numpy as np
import scipy.sparse as sparse


def arun_from_gram(gram, cell_mixture, length_sq, topic_word=None) -> float:
    """Arun et al. 2010 from ``gram = topic_word @ topic_word.T``, ``cell_mixture = doc_lengths @ doc_topic``
    and ``length_sq = sum(doc_lengths ** 2)``.

    cm1 = singular values of topic_word (descending) = sqrt(eig(gram)); cm2 = cell_mixture / ||doc_lengths||_2
    (topic order, NOT sorted); symmetric KL of the two (un-normalised) vectors.  ``topic_word`` (an array or
    a callable returning it) is only touched when the Gram is too ill-conditioned for the eigenvalue route.
    """
    ev = np.linalg.eigvalsh(gram)
    if ev.min() <= 1e-12 * ev.max():  # ill-conditioned: fall back on the direct SVD
        if topic_word is None:
            raise np.linalg.LinAlgError("topic Gram matrix is singular to working precision")
        cm1 = np.linalg.svd(topic_word() if callable(topic_word) else topic_word, compute_uv=False)
    else:
        cm1 = np.sqrt(ev[::-1])
    cm2 = np.asarray(cell_mixture, np.float64) / math.sqrt(length_sq)
    return float(np.sum(cm1 * np.log(cm1 / cm2)) + np.sum(cm2 * np.log(cm2 / cm1)))


def metric_arun_2010(topic_word, doc_topic, doc_lengths) -> float:
    """Arun et al. 2010 (tmtoolkit.topicmod.evaluate.metric_arun_2010)."""
    L = np.asarray(doc_lengths, dtype=np.float64).reshape(-1)
    # singular values of the K x W matrix = sqrt(eig(T T^T)): the Gram is only K x K
    return arun_from_gram(topic_word @ topic_word.T, L @ doc_topic, float(L @ L), topic_word)


def cao_juan_from_gram(gram) -> float:
    """Cao Juan et al. 2009 from ``topic_word @ topic_word.T``: mean pairwise cosine similarity."""
    K = gram.shape[0]
    if K < 2:
        return float("nan")
    norms = np.sqrt(np.diagonal(gram))
    cos = gram / np.outer(norms, norms)
    iu = np.triu_indices(K, 1)
    return float(np.mean(cos[iu]))


def metric_cao_juan_2009(topic_word) -> float:
    """Cao Juan et al. 2009: mean pairwise cosine similarity of the topic-word rows."""
    if topic_word.shape[0] < 2:
        return float("nan")
    return cao_juan_from_gram(topic_word @ topic_word.T)


def top_regions_stable(topic_word, top_n):
    """tmtoolkit's ``argsort(topic_word, axis=1)[:, :-(top_n + 1):-1]`` with a STABLE sort: equal
    probabilities come out by region index descending (numpy's default sort leaves ties unspecified);
    the order the device kernel (cgs_model_top_regions) produces."""
    return np.argsort(topic_word, axis=1, kind="stable")[:, :-(top_n + 1):-1]


def mimno_from_codf(codf, eps=1e-12, normalize=True):
    """Mimno et al. 2011 coherence per topic from the co-document frequencies of every topic's top regions
    (K x top_n x top_n, most probable region first, diagonal = document frequencies):
    ``sum_{m > l} log((codf[m, l] + eps) / df[l])``, times ``2 / (n (n - 1))`` when normalised."""
    codf = np.asarray(codf, dtype=np.float64)
    K, top_n = codf.shape[0], codf.shape[1]
    tri = np.tril_indices(top_n, -1)  # pairs (m, l) with l < m
    coh = np.empty(K, np.float64)
    for t in range(K):
        df = np.diagonal(codf[t])
        coh[t] = np.sum(np.log((codf[t][tri] + eps) / df[tri[1]]))
    if normalize:
        coh *= 2.0 / (top_n * (top_n - 1))
    return coh


def metric_coherence_mimno_2011(topic_word, regions_by_cells, top_n=20, eps=1e-12, normalize=True, top=None):
    """Mimno et al. 2011 coherence per topic (tmtoolkit ... metric_coherence_mimno_2011 with
    ``return_mean=False``).  ``regions_by_cells`` is the binary matrix in the reference's own
    layout (CSR, regions x cells), so the 20 top regions of a topic are 20 CSR rows."""
    K, W = topic_word.shape
    if top_n > W:
        raise ValueError("`top_n=%d` is larger than the vocabulary size of %d words" % (top_n, W))
    M = sparse.csr_matrix(regions_by_cells)
    # indices of the top_n largest probabilities, most probable first (ties: as argsort gives them)
    if top is None:
        top = np.argsort(topic_word, axis=1)[:, :-(top_n + 1):-1]
    codf = np.empty((K, top_n, top_n), np.float64)
    for t in range(K):
        sub = M[top[t]]
        sub = sparse.csr_matrix((np.ones(sub.nnz, np.float64), sub.indices, sub.indptr), shape=sub.shape)
        codf[t] = (sub @ sub.T).toarray()
    return mimno_from_codf(codf, eps, normalize)


def marginal_topic_distrib(doc_topic, doc_lengths):
    """tmtoolkit.topicmod.model_stats.marginal_topic_distrib."""
    L = np.asarray(doc_lengths, dtype=np.float64).reshape(-1)
    return marginal_from_mixture(doc_topic.T @ L)


def marginal_from_mixture(cell_mixture):
    """marginal_topic_distrib from ``doc_topic.T @ doc_lengths``."""
    u = np.asarray(cell_mixture, np.float64)
    return u / u.sum()


def loglikelihood_compat(nzw, ndz, alpha, eta) -> float:
    """The number the reference stores as ``metrics['loglikelihood']``.

    ``utils.loglikelihood`` (utils.py:129-160) overwrites ``topic_ll`` / ``doc_ll`` inside its
    loops (``=`` where ``+=`` is meant, :145 and :155), so only the LAST positive entry of the
    last non-empty row survives, minus the ``lgamma`` of the row totals seen since.  This is the
    O(K + D) closed form of exactly that recurrence; it is called with the RAW alpha and eta
    (lda_models.py:305), not the per-topic scaled ones.
    """
    nzw = np.asarray(nzw)
    ndz = np.asarray(ndz)
    D, K = ndz.shape
    W = nzw.shape[1]
    const_prior = (K * math.lgamma(alpha) - math.lgamma(alpha * K)) * D
    const_ll = (W * math.lgamma(eta) - math.lgamma(eta * W)) * K

    def tail(counts, prior, width):
        """value of the recurrence  t = 0; for row: (if any>0: t = lgamma(last_pos + prior)); t -= lgamma(prior*width + rowsum)"""
        pos = counts > 0
        has = pos.any(axis=1)
        rowsum = counts.sum(axis=1, dtype=np.float64)  # only entries > 0 matter; counts are >= 0
        t = 0.0
        start = 0
        if has.any():
            r = int(np.flatnonzero(has)[-1])
            c = int(np.flatnonzero(pos[r])[-1])
            t = math.lgamma(float(counts[r, c]) + prior)
            start = r
        for i in range(start, counts.shape[0]):
            t -= math.lgamma(prior * width + float(rowsum[i]))
        return t

    topic_ll = tail(nzw, eta, W)
    doc_ll = tail(ndz, alpha, K)
    return doc_ll - const_prior + topic_ll - const_ll
