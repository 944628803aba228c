"""Mallet text formats either side of the path: what ``LDAMallet`` of the reference
writes before and reads after a Mallet run (src/pycisTopic/lda_models.py:472-494 ``corpus_to_mallet``, :601-646
``load_word_topics``), so that corpora and sampler states move between Mallet and the device sampler.

* :func:`corpus_to_mallet` writes the corpus text Mallet imports -- one line per cell, ``"<cell>\\t0\\t<region ids>"``
  (:488-494) -- from a regions x cells binary matrix or a gensim-style bag-of-words corpus.
* :func:`read_state` / :func:`load_word_topics` read a Mallet ``--output-state`` file (gzip text: three ``#`` header
  lines, then ``doc source pos typeindex type topic`` per token); ``load_word_topics`` returns the topics x regions
  count matrix exactly as the reference builds it from columns 4 and 5 (:620-646).
* :func:`lda_from_state` puts such a state on the GPU: the assignments become the ``z`` of a device model
  (``cgs_model_set_z``; counts are rebuilt there, bit-exact bookkeeping), which is evaluated as it is (``n_iter = 0``)
  or sampled further, and comes back as a fitted ``lda.LDA``-protocol object that ``build_cistopic_model`` packs into a
  :class:`CistopicLDAModel` with the device metric block.
* :func:`write_state` writes a fitted device model's assignments in the same format.

No Mallet binary, no Java: only the formats.  Parsing is host text work, and ``load_word_topics`` -- a reader whose
result is a function of the file alone, a polars group-by in the reference -- tallies the parsed columns on the host
like the reference does; everything that involves the corpus (counts per cell and region, metrics, further sweeps)
goes through :func:`lda_from_state`, i.e. the device.
"""
from __future__ import annotations

import gzip

import numpy as np
import scipy.sparse as sparse

from .lda import LDA

STATE_HEADER = "#doc source pos typeindex type topic\n"


def _open_text(path, mode):
    with open(path, "rb") as probe:
        magic = probe.read(2) if "r" in mode else b""
    if "r" in mode and magic != b"\x1f\x8b":
        return open(path, mode + "t")
    return gzip.open(path, mode + "t")


def corpus_to_mallet(corpus, file_like):
    """``LDAMallet.corpus_to_mallet`` (lda_models.py:472-494): ``corpus`` is a regions x cells binary matrix (what
    ``run_cgs_models_mallet`` starts from, :797-803) or an iterable of ``[(region, count), ...]`` per cell."""
    if sparse.issparse(corpus) or isinstance(corpus, np.ndarray):
        X = sparse.csc_matrix(corpus)           # column = cell, row indices = accessible regions
        X.sort_indices()
        indptr, indices = X.indptr, X.indices
        for cell in range(X.shape[1]):
            regions = indices[indptr[cell]:indptr[cell + 1]]
            file_like.write(f'{cell}\t0\t{" ".join(map(str, regions.tolist()))}\n')
        return
    for cell, doc in enumerate(corpus):
        file_like.write(f'{cell}\t0\t{" ".join(str(region) for region, _cnt in doc)}\n')


def read_state(path):
    """A Mallet state file -> ``{"alpha": float64[K], "beta": float, "doc", "region", "topic": int32[N]}``; tokens in
    file order (cell by cell, position by position).  ``region`` is column 4 (the token text, which
    ``corpus_to_mallet`` made the region index -- the column the reference reads, :625-631)."""
    with _open_text(path, "r") as f:
        header = f.readline()
        if not header.startswith("#doc"):
            raise ValueError(f"{path}: not a Mallet state file (first line {header!r})")
        alpha_line = f.readline().split()
        beta_line = f.readline().split()
        if alpha_line[:2] != ["#alpha", ":"] or beta_line[:2] != ["#beta", ":"]:
            raise ValueError(f"{path}: malformed #alpha / #beta header")
        alpha = np.array(alpha_line[2:], dtype=np.float64)
        beta = float(beta_line[2])
        import pandas as pd
        try:
            cols = pd.read_csv(f, sep=" ", header=None, usecols=[0, 4, 5], dtype=np.int64).to_numpy()
        except pd.errors.EmptyDataError:
            cols = np.zeros((0, 3), np.int64)
    return {"alpha": alpha, "beta": beta, "doc": cols[:, 0].astype(np.int32), "region": cols[:, 1].astype(np.int32),
            "topic": cols[:, 2].astype(np.int32)}


def load_word_topics(path, num_topics, num_terms):
    """``LDAMallet.load_word_topics`` (lda_models.py:601-646): topics x regions float64 matrix of assignment counts."""
    state = read_state(path)
    if len(state["alpha"]) != num_topics:
        raise AssertionError("Mismatch between MALLET vs. requested topics")
    word_topics = np.zeros((num_topics, num_terms), dtype=np.float64)
    np.add.at(word_topics, (state["topic"], state["region"]), 1.0)
    return word_topics


def write_state(path, doc, region, topic, alpha, beta):
    """Tokens (in order) -> a Mallet ``--output-state`` file (gzip unless ``path`` does not end in ``.gz``)."""
    doc, region, topic = (np.asarray(a, dtype=np.int64) for a in (doc, region, topic))
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "wt") as f:
        f.write(STATE_HEADER)
        f.write("#alpha : " + " ".join(repr(float(a)) for a in np.atleast_1d(alpha)) + " \n")
        f.write(f"#beta : {float(beta)!r}\n")
        if len(doc):
            start = np.r_[0, np.flatnonzero(np.diff(doc)) + 1]
            pos = np.arange(len(doc)) - np.repeat(start, np.diff(np.r_[start, len(doc)]))
            # typeindex: Mallet numbers the types by first appearance; the reader uses the token text (column 4)
            _, first = np.unique(region, return_index=True)
            order = np.empty(region.max() + 1, np.int64)
            order[region[np.sort(first)]] = np.arange(len(first))
            rows = np.column_stack([doc, np.zeros_like(doc), pos, order[region], region, topic])
            f.write("\n".join(f"{d} NA {p} {ti} {r} {t}" for d, _s, p, ti, r, t in rows.tolist()) + "\n")


def lda_from_state(regions_by_cells, state, n_topics, alpha, eta, n_iter=0, random_state=555, device=0, **lda_kwargs):
    """A Mallet state (path or :func:`read_state` result) -> fitted ``lda.LDA``-protocol object on the GPU.

    The state's tokens must be those of ``regions_by_cells`` (the matrix ``corpus_to_mallet`` exported): cell by
    cell, regions ascending inside a cell -- checked.  ``n_iter = 0`` evaluates the state as it is; ``n_iter > 0``
    continues sampling on the device from it.  ``alpha`` / ``eta`` are per-topic values as ``lda.LDA`` takes them."""
    if not isinstance(state, dict):
        state = read_state(state)
    X = sparse.csc_matrix(regions_by_cells)
    X.sort_indices()
    if len(state["region"]) != X.nnz:
        raise ValueError(f"the state holds {len(state['region'])} tokens, the matrix {X.nnz}")
    doc = np.repeat(np.arange(X.shape[1], dtype=np.int32), np.diff(X.indptr))
    order = np.lexsort((state["region"], state["doc"]))      # Mallet keeps file order; be robust to shuffled lines
    if not (np.array_equal(state["doc"][order], doc) and np.array_equal(state["region"][order], X.indices)):
        raise ValueError("the state's (cell, region) tokens are not those of the matrix")
    z = state["topic"][order]
    if z.size and (z.min() < 0 or z.max() >= n_topics):
        raise ValueError("the state assigns a topic outside [0, n_topics)")
    cells_by_regions = sparse.csr_matrix((X.data, X.indices, X.indptr), shape=(X.shape[1], X.shape[0]))
    model = LDA(n_topics=n_topics, n_iter=n_iter, alpha=alpha, eta=eta, random_state=random_state, device=device,
                z_init=z, **lda_kwargs)
    return model.fit(cells_by_regions)


def state_from_lda(model, path, beta=None):
    """Assignments of a model fitted with ``keep_assignments=True`` -> Mallet state file."""
    if not hasattr(model, "z_"):
        raise ValueError("fit the model with keep_assignments=True")
    write_state(path, model.DS_, model.WS_, model.z_, np.full(model.n_topics, model.alpha), model.eta if beta is None else beta)
