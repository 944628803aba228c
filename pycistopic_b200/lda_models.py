"""Drop-in mirror of ``pycisTopic.lda_models`` for the collapsed-Gibbs path.

Same public names, argument meaning and result layout as the reference
(src/pycisTopic/lda_models.py): :class:`CistopicLDAModel` (:31-96), :func:`run_cgs_models`
(:99-175), :func:`run_cgs_model` (:178-396), :func:`evaluate_models` (:1048-1270),
:func:`load_cisTopic_model` (:1273-1289).  What changed is underneath: the token list is built
on the GPU, the sweeps run in hand-written sm_100a kernels behind the C ABI, and the
``n_topics`` models are spread over the visible GPUs (one model per GPU at a time, longest
first) instead of over Ray worker processes.  Ray-only keyword arguments (``_temp_dir`` ...)
are accepted and ignored; ``n_cpu`` is accepted and ignored.
"""
from __future__ import annotations

import logging
import os
import pickle
import sys
import threading
import time
import warnings

import numpy as np
import pandas as pd
from scipy import sparse

from . import _lib
from . import metrics as _metrics
from .lda import LDA, Corpus

__all__ = ["CistopicLDAModel", "run_cgs_models", "run_cgs_model", "evaluate_models", "load_cisTopic_model"]


class CistopicLDAModel:
    """cisTopic LDA model container -- attribute-for-attribute the reference class
    (lda_models.py:31-96): ``metrics``, ``coherence``, ``marg_topic``, ``topic_ass``,
    ``cell_topic`` (topics x cells), ``cell_topic_harmony``, ``topic_region`` (regions x topics),
    ``parameters``, ``n_cells``, ``n_regions``, ``n_topic``."""

    def __init__(self, metrics, coherence, marg_topic, topic_ass, cell_topic, topic_region, parameters):
        self.metrics = metrics
        self.coherence = coherence
        self.marg_topic = marg_topic
        self.topic_ass = topic_ass
        self.cell_topic = cell_topic
        self.cell_topic_harmony = []
        self.topic_region = topic_region
        self.parameters = parameters
        self.n_cells = cell_topic.shape[1]
        self.n_regions = topic_region.shape[0]
        self.n_topic = cell_topic.shape[0]

    def __str__(self):
        return (f"CistopicLDAModel with {self.n_topic} topics and n_cells × n_regions = "
                f"{self.n_cells} × {self.n_regions}")


def _logger():
    level = logging.INFO
    log_format = "%(asctime)s %(name)-12s %(levelname)-8s %(message)s"
    handlers = [logging.StreamHandler(stream=sys.stdout)]
    logging.basicConfig(level=level, format=log_format, handlers=handlers)
    return logging.getLogger("cisTopic")


def model_cost(n_topics: int) -> float:
    """Relative cost of one sweep: algorithmic bytes per token-update, 4K + 28."""
    return 4.0 * n_topics + 28.0


def schedule_models(n_topics, n_workers):
    """Longest-processing-time assignment of models to workers.  Returns a list (per worker) of
    indices into ``n_topics``; deterministic."""
    order = sorted(range(len(n_topics)), key=lambda i: (-model_cost(n_topics[i]), i))
    loads = [0.0] * n_workers
    plan = [[] for _ in range(n_workers)]
    for i in order:
        w = min(range(n_workers), key=lambda j: (loads[j], j))
        plan[w].append(i)
        loads[w] += model_cost(n_topics[i])
    return plan


def sweep_cost(n_topics: int) -> float:
    """Relative time of one sweep over a fixed corpus as measured on B200 (C2, round 2: 0.59 ms at K = 2, 0.95 at
    10, 1.37 at 25, 2.35 at 50): the small models are latency-bound, so time grows like K + 15 rather than 4K + 28."""
    return float(n_topics) + 15.0


def schedule_models_split(costs, n_workers, overhead=None, min_fraction=0.2):
    """Schedule for more workers than whole models can fill: models are DIVISIBLE (a model's cells can be sharded
    over consecutive workers, AD-LDA: one n_wk exchange per sweep between them), so the workers are filled one
    after the other up to a common level T (McNaughton's wrap-around rule) and a model that does not fit is
    continued on the next worker.  ``costs[i]`` = time of one sweep of model i on one whole worker, ``overhead[i]``
    = what every PART of a split model pays on top (its share of the exchange).  Models are taken longest first,
    so the split ones are the big ones, where the overhead matters least.  A part is never smaller than
    ``min_fraction`` of a model.  Returns ``(plan, T)``: ``plan[w]`` = ordered list of ``(model, f0, f1)`` -- worker w
    samples the cells whose cumulative token fraction lies in [f0, f1) -- with the part shared with worker w - 1
    first and the part shared with worker w + 1 last, so that neighbours meet at their common model once per step.
    Deterministic."""
    n = len(costs)
    overhead = [0.0] * n if overhead is None else list(overhead)
    order = sorted(range(n), key=lambda i: (-costs[i], i))

    def fill(T, whole_first):
        plan = [[] for _ in range(n_workers)]
        todo = list(order)           # models not started yet, longest first
        carry = None                 # (model, fraction already placed): continues on the next worker, first
        for w in range(n_workers):
            load = 0.0
            head, middle, tail = [], [], []
            if carry is not None:
                i, done = carry
                rest = (1.0 - done) * costs[i] + overhead[i]
                if rest <= T + 1e-12:
                    head.append((i, done, 1.0))
                    load += rest
                    carry = None
                else:  # a model longer than two levels: another full worker of it
                    frac = min(1.0 - done - min_fraction, (T - overhead[i]) / costs[i])
                    if frac < min_fraction:
                        return None  # the level is too low for parts of the minimum size
                    head.append((i, done, done + frac))
                    carry = (i, done + frac)
                    plan[w] = head
                    continue
            while todo:
                room = T - load
                # whole_first: the longest model that fits whole goes in before anything is split (few splits);
                # otherwise only the longest remaining model is considered (pure wrap-around: equal levels)
                fits = next((i for i in (todo if whole_first else todo[:1]) if costs[i] <= room + 1e-12), None)
                if fits is not None:
                    middle.append((fits, 0.0, 1.0))
                    load += costs[fits]
                    todo.remove(fits)
                    continue
                # nothing fits whole: start the longest remaining model here and continue it on the next worker
                i = todo[0]
                frac = (room - overhead[i]) / costs[i]
                if frac < min_fraction:
                    break
                if frac > 1.0 - min_fraction:  # never leave a sliver behind: place a little less, keep filling
                    frac = 1.0 - min_fraction
                tail.append((i, 0.0, frac))
                load += frac * costs[i] + overhead[i]
                todo.pop(0)
                carry = (i, frac)
                break
            if carry is not None and tail and todo:
                # room left after a clamped part: whole models that still fit go in the middle
                for i in list(todo):
                    if costs[i] <= T - load + 1e-12:
                        middle.append((i, 0.0, 1.0))
                        load += costs[i]
                        todo.remove(i)
            plan[w] = head + middle + tail
        if todo or carry is not None:
            return None
        return plan

    best = None
    for whole_first in (True, False):
        lo = sum(costs) / n_workers
        hi = lo * 2 + max(costs) + max(overhead)
        plan = fill(hi, whole_first)
        if plan is None:
            continue
        for _ in range(40):
            mid = 0.5 * (lo + hi)
            trial = fill(mid, whole_first)
            if trial is None:
                lo = mid
            else:
                hi, plan = mid, trial
        if best is None or hi < best[1] - 1e-9:
            best = (plan, hi)
    if best is None:  # cannot happen for sane inputs; fall back to whole models
        whole = schedule_models([0] * n, n_workers)
        return [[(i, 0.0, 1.0) for i in p] for p in whole], max(costs)
    return best


def schedule_models_paired(costs, n_workers, penalty=1.07, overhead=0.05, min_fraction=0.25, max_splits=None,
                           equal_parts=True):
    """Whole models, longest first -- except that the j longest models are split in TWO parts that run on two workers
    at the same time (first in both workers' steps, so that their overlapped exchange meets), a worker taking at most
    one such part.  For every j: the parts and the whole models are placed longest-first, a split model is cut in equal
    halves (``equal_parts``: the partners' launches then take the same time and never wait for each other -- measured
    better than fractions that level the two workers, which is what ``equal_parts=False`` does), and a local search (move or swap whole models between workers, re-levelling
    after each try) lowers the makespan further; the j with the smallest predicted makespan wins (j = 0 is plain LPT
    plus the local search).  A part costs ``penalty`` x its share of the model (a shard fills the GPU less evenly) plus
    ``overhead``.  Returns ``(plan, makespan)`` with ``plan[w]`` = ordered ``(model, f0, f1)`` items, split part first.
    Deterministic."""
    n = len(costs)
    order = sorted(range(n), key=lambda i: (-costs[i], i))
    max_splits = min(n, n_workers // 2) if max_splits is None else min(max_splits, n, n_workers // 2)

    def evaluate(whole, pairs):
        """whole[w] = list of whole models on w, pairs = {model: (wa, wb)} -> (loads, fractions)."""
        loads = [sum(costs[i] for i in whole[w]) for w in range(n_workers)]
        fractions = {}
        for i, (wa, wb) in pairs.items():
            c = costs[i] * penalty
            f = min(1.0 - min_fraction, max(min_fraction, (loads[wb] - loads[wa] + c) / (2.0 * c)))
            if equal_parts:
                f = 0.5  # the two parts then take the same time, so the partners' fused launches meet without waiting
            loads[wa] += f * c + overhead
            loads[wb] += (1.0 - f) * c + overhead
            fractions[i] = f
        return loads, fractions

    def score(loads):
        return (round(max(loads), 9), round(sum(x * x for x in loads), 9))

    best = None
    for j in range(0, max_splits + 1):
        split = order[:j]
        items = []
        for i in order:
            if i in split:
                half = 0.5 * costs[i] * penalty + overhead
                items += [(half, i, 0), (half, i, 1)]
            else:
                items.append((costs[i], i, None))
        items.sort(key=lambda t: (-t[0], t[1], t[2] if t[2] is not None else -1))
        loads = [0.0] * n_workers
        part_on = {}
        place = {}
        whole = [[] for _ in range(n_workers)]
        ok = True
        for c, i, part in items:
            cands = [w for w in range(n_workers) if part is None or (w not in part_on and place.get((i, 1 - part)) != w)]
            if not cands:
                ok = False
                break
            w = min(cands, key=lambda w: (loads[w], w))
            loads[w] += c
            if part is not None:
                part_on[w] = i
                place[(i, part)] = w
            else:
                whole[w].append(i)
        if not ok:
            continue
        pairs = {i: (place[(i, 0)], place[(i, 1)]) for i in split}
        loads, fractions = evaluate(whole, pairs)
        def one_better(whole, loads):
            """First move / swap of whole models between two workers that lowers (makespan, sum of squares)."""
            cur = score(loads)
            for a in range(n_workers):
                for b in range(n_workers):
                    if a == b:
                        continue
                    trials = [([x], []) for x in whole[a]] + [([x], [y]) for x in whole[a] for y in whole[b] if a < b]
                    for out_a, out_b in trials:
                        cand = [list(w) for w in whole]
                        for x in out_a:
                            cand[a].remove(x)
                            cand[b].append(x)
                        for y in out_b:
                            cand[b].remove(y)
                            cand[a].append(y)
                        l2, f2 = evaluate(cand, pairs)
                        if score(l2) < cur:
                            return cand, l2, f2
            return None

        for _ in range(200):  # local search
            nxt = one_better(whole, loads)
            if nxt is None:
                break
            whole, loads, fractions = nxt
        plan = [[(i, 0.0, 1.0) for i in sorted(whole[w], key=lambda i: (-costs[i], i))] for w in range(n_workers)]
        for i, (wa, wb) in pairs.items():
            plan[wa].insert(0, (i, 0.0, fractions[i]))
            plan[wb].insert(0, (i, fractions[i], 1.0))
        if best is None or max(loads) < best[1] - 1e-9:
            best = (plan, max(loads))
    return best


def run_cgs_models(
    cistopic_obj,
    n_topics,
    n_cpu=1,
    n_iter=150,
    random_state=555,
    alpha=50,
    alpha_by_topic=True,
    eta=0.1,
    eta_by_topic=False,
    top_topics_coh=5,
    save_path=None,
    **kwargs,
):
    """Run LDA with collapsed Gibbs sampling for every value in ``n_topics``.

    Signature and return value of the reference (lda_models.py:99-175): a list of
    :class:`CistopicLDAModel` in the order of ``n_topics``.  ``n_cpu`` and Ray keyword
    arguments are ignored; ``devices=[...]`` (keyword-only extra) restricts the GPUs used;
    ``deterministic=True`` (or a number of sub-sweeps) makes the result a pure function of ``random_state`` as the
    reference's is (lda_models.py:104) -- the default schedule is only statistically reproducible.
    """
    devices = kwargs.pop("devices", None)
    deterministic = kwargs.pop("deterministic", 0)
    n_gpu = _lib.require_device()
    if devices is None:
        devices = list(range(n_gpu))
    # one Index object shared by the K DataFrames of every model (pandas would rebuild it per model)
    region_names = pd.Index(cistopic_obj.region_names)
    cell_names = pd.Index(cistopic_obj.cell_names)
    binary_matrix = cistopic_obj.binary_matrix  # regions x cells; transposed on the device
    canonical = Corpus.canonical_csr(binary_matrix)  # validated once, shared by every device's corpus build
    devices = devices[: max(1, min(len(devices), len(n_topics)))]
    # whole models, longest first by the measured cost model, improved by moves / swaps between GPUs
    plan = [[i for (i, _, _) in p]
            for p in schedule_models_paired([sweep_cost(k) for k in n_topics], len(devices), max_splits=0)[0]]
    results = [None] * len(n_topics)
    errors = []

    def worker(dev, idxs):
        # The sweeps of model i+1 run on the GPU while the host packs model i (metrics + DataFrames,
        # about a second per model at 300 k regions): one packing thread per device, results in order.
        from concurrent.futures import ThreadPoolExecutor
        try:
            corpus = Corpus.from_regions_by_cells(canonical, device=dev)
            pending = {}
            try:
                with ThreadPoolExecutor(max_workers=1) as packer:
                    for i in idxs:
                        fitted, end_time, log = _fit_cgs_model_on_corpus(
                            corpus, n_topics[i], n_iter, random_state, alpha, alpha_by_topic, eta, eta_by_topic,
                            deterministic)
                        pending[i] = packer.submit(
                            build_cistopic_model, fitted, binary_matrix, n_topics[i], cell_names, region_names,
                            n_iter, random_state, alpha, alpha_by_topic, eta, eta_by_topic, top_topics_coh, end_time,
                            save_path, log)
                    for i, fut in pending.items():
                        results[i] = fut.result()
            finally:
                corpus.close()
        except BaseException as e:  # propagate to the caller like ray.get would
            errors.append(e)

    if len(devices) == 1:
        worker(devices[0], plan[0])
    else:
        threads = [threading.Thread(target=worker, args=(d, p)) for d, p in zip(devices, plan) if p]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
    return results


def run_cgs_model(
    binary_matrix,
    n_topics,
    cell_names,
    region_names,
    n_iter=150,
    random_state=555,
    alpha=50,
    alpha_by_topic=True,
    eta=0.1,
    eta_by_topic=False,
    top_topics_coh=5,
    save_path=None,
    device=0,
):
    """One model (lda_models.py:178-396).  ``binary_matrix`` is cells x regions, as there."""
    X = sparse.csr_matrix(binary_matrix)
    corpus = Corpus.from_cells_by_regions(X, device=device)
    try:
        return _run_cgs_model_on_corpus(corpus, sparse.csr_matrix(X.T), n_topics, cell_names, region_names, n_iter,
                                        random_state, alpha, alpha_by_topic, eta, eta_by_topic, top_topics_coh,
                                        save_path)
    finally:
        corpus.close()


def _fit_cgs_model_on_corpus(corpus, n_topics, n_iter, random_state, alpha, alpha_by_topic, eta, eta_by_topic,
                             deterministic=0):
    """The device part of run_cgs_model (lda_models.py:247-289): returns (fitted LDA, seconds, logger)."""
    # The reference silences warnings and the `lda` logger inside its Ray worker PROCESS (lda_models.py:234-238);
    # here the fit runs in the caller's process, so neither is changed beyond this call.
    log = _logger()
    lda_log = logging.getLogger("lda")
    if not any(isinstance(h, logging.NullHandler) for h in lda_log.handlers):
        lda_log.addHandler(logging.NullHandler())

    # hyper-parameter scaling of lda_models.py:247-282
    model = LDA(
        n_topics=n_topics,
        n_iter=n_iter,
        random_state=random_state,
        alpha=alpha / n_topics if alpha_by_topic else alpha,
        eta=eta / n_topics if eta_by_topic else eta,
        refresh=n_iter,
        device=corpus.device,
        corpus=corpus,
        exports="frames",
        device_metrics={"alpha": alpha, "eta": eta, "top_n": 20},
        deterministic=deterministic,
    )
    log.info(f"Running model with {n_topics} topics")
    start_time = time.time()
    propagate = lda_log.propagate
    lda_log.propagate = False
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            model.fit(None)
    finally:
        lda_log.propagate = propagate
    return model, time.time() - start_time, log


def _run_cgs_model_on_corpus(corpus, regions_by_cells, n_topics, cell_names, region_names, n_iter, random_state,
                             alpha, alpha_by_topic, eta, eta_by_topic, top_topics_coh, save_path):
    model, end_time, log = _fit_cgs_model_on_corpus(corpus, n_topics, n_iter, random_state, alpha, alpha_by_topic,
                                                    eta, eta_by_topic)
    return build_cistopic_model(model, regions_by_cells, n_topics, cell_names, region_names, n_iter, random_state,
                                alpha, alpha_by_topic, eta, eta_by_topic, top_topics_coh, end_time, save_path, log)


def build_cistopic_model(model, regions_by_cells, n_topics, cell_names, region_names, n_iter, random_state, alpha,
                         alpha_by_topic, eta, eta_by_topic, top_topics_coh, end_time, save_path=None, log=None):
    """Metric block and DataFrame packing of lda_models.py:291-388 from a fitted LDA-protocol
    object (needs topic_word_, doc_topic_, nzw_, ndz_, nz_)."""
    ing = getattr(model, "metric_ingredients_", None)
    if ing is not None:
        # Gram matrix, cell mixture, top regions and their (co-)document frequencies came from the device
        # (cgs_model_topic_gram / _cell_mixture / _top_regions / _codocument_frequencies / _loglik_compat)
        arun_2010 = _metrics.arun_from_gram(ing["gram"], ing["cell_mixture"], ing["length_sq"],
                                            lambda: np.ascontiguousarray(model.topic_region_.T))
        cao_juan_2009 = _metrics.cao_juan_from_gram(ing["gram"])
        mimno_2011 = _metrics.mimno_from_codf(ing["codf"], eps=1e-12, normalize=True)
        ll = ing["loglikelihood_compat"]
        marg = _metrics.marginal_from_mixture(ing["cell_mixture"])
    else:  # any other fitted lda-protocol object: the same formulas from its host arrays
        M = sparse.csr_matrix(regions_by_cells)
        doc_lengths = np.asarray(M.sum(axis=0)).ravel().astype(float)  # tokens per cell
        arun_2010 = _metrics.metric_arun_2010(model.topic_word_, model.doc_topic_, doc_lengths)
        cao_juan_2009 = _metrics.metric_cao_juan_2009(model.topic_word_)
        mimno_2011 = _metrics.metric_coherence_mimno_2011(model.topic_word_, M, top_n=20, eps=1e-12, normalize=True)
        ll = _metrics.loglikelihood_compat(model.nzw_, model.ndz_, alpha, eta)
        marg = _metrics.marginal_topic_distrib(model.doc_topic_, doc_lengths)

    if len(mimno_2011) <= top_topics_coh:
        model_coh = np.mean(mimno_2011)
    else:
        model_coh = np.mean(mimno_2011[np.argpartition(mimno_2011, -top_topics_coh)[-top_topics_coh:]])
    metrics = pd.DataFrame(
        [arun_2010, cao_juan_2009, model_coh, ll],
        index=["Arun_2010", "Cao_Juan_2009", "Mimno_2011", "loglikelihood"],
        columns=["Metric"],
    ).transpose()
    topics = np.arange(1, n_topics + 1)
    coherence = pd.DataFrame({"Topic": topics.astype(np.float64), "Mimno_2011": np.asarray(mimno_2011, np.float64)})
    marg_topic = pd.DataFrame({"Topic": topics.astype(np.float64), "Marg_Topic": np.asarray(marg, np.float64)})
    topic_ass = pd.DataFrame({"Topic": topics.astype(np.int64), "Assignments": np.asarray(model.nz_, np.int64)})
    topic_names = ["Topic" + str(i) for i in range(1, n_topics + 1)]
    cell_topic_arr = getattr(model, "cell_topic_", None)
    if cell_topic_arr is None:
        cell_topic_arr = np.ascontiguousarray(model.doc_topic_.T)
    topic_region_arr = getattr(model, "topic_region_", None)
    if topic_region_arr is None:
        topic_region_arr = np.ascontiguousarray(model.topic_word_.T)
    cell_topic = pd.DataFrame(cell_topic_arr, index=topic_names, columns=cell_names)
    topic_region = pd.DataFrame(topic_region_arr, index=region_names, columns=topic_names)
    parameters = pd.DataFrame(
        ["lda", n_topics, n_iter, random_state, alpha, alpha_by_topic, eta, eta_by_topic, top_topics_coh, end_time],
        index=["package", "n_topics", "n_iter", "random_state", "alpha", "alpha_by_topic", "eta", "eta_by_topic",
               "top_topics_coh", "time"],
        columns=["Parameter"],
    )
    out = CistopicLDAModel(metrics, coherence, marg_topic, topic_ass, cell_topic, topic_region, parameters)
    # extras (not in the reference): the true Griffiths-Steyvers log-likelihood of the final state
    out.loglikelihood_true = getattr(model, "final_loglikelihood_", None)
    if log is not None:
        log.info(f"Model with {n_topics} topics done!")
    if isinstance(save_path, str):
        if log is not None:
            log.info(f"Saving model with {n_topics} topics at {save_path}")
        from . import compat
        os.makedirs(save_path, exist_ok=True)  # one packing thread per GPU may get here at the same time
        with open(os.path.join(save_path, f"Topic{n_topics}.pkl"), "wb") as f:
            pickle.dump(compat.pickle_ready(out), f)
    return out


def _rescale(values):
    values = np.asarray(values, dtype=float)
    return (values - values.min()) / (values.max() - values.min())


def evaluate_models(
    models,
    select_model=None,
    return_model=True,
    metrics=["Minmo_2011", "loglikelihood", "Cao_Juan_2009", "Arun_2010"],
    min_topics_coh=5,
    plot=True,
    figsize=(6.4, 4.8),
    plot_metrics=False,
    save=None,
):
    """Model selection on the four quality metrics (lda_models.py:1048-1270).

    Behaviour kept from the reference: models are sorted by ``n_topic``; every metric is
    min-max rescaled (Arun and Cao-Juan negated first); the coherence -- spelled
    ``"Minmo_2011"`` in ``metrics``, as in the reference -- only uses models with at least
    ``min_topics_coh`` topics and, when present, restricts the other metrics to the same
    models; the best model maximises the sum.  Plotting needs matplotlib and is skipped
    (with a log line) when it is not installed.
    """
    models = [models[i] for i in np.argsort([m.n_topic for m in models])]
    all_topics = sorted([models[x].n_topic for x in range(0, len(models))])
    metrics_dict = {}
    curves = []  # (x, y, label) for the combined plot
    raw = {}
    use_coh = "Minmo_2011" in metrics
    if use_coh:
        in_index = [i for i in range(len(all_topics)) if all_topics[i] >= min_topics_coh]

    def pick(values):
        return np.array([values[i] for i in in_index]) if use_coh else values

    if "Arun_2010" in metrics:
        arun = [m.metrics.loc["Metric", "Arun_2010"] for m in models]
        raw["Arun_2010"] = (all_topics, arun, "Arun_2010 - Minimize")
        res = _rescale([-x for x in arun])
        metrics_dict["Arun_2010"] = pick(res)
        curves.append((all_topics, res, "Inv_Arun_2010"))
    if "Cao_Juan_2009" in metrics:
        cao = [m.metrics.loc["Metric", "Cao_Juan_2009"] for m in models]
        raw["Cao_Juan_2009"] = (all_topics, cao, "Cao_Juan_2009 - Minimize")
        res = _rescale([-x for x in cao])
        metrics_dict["Cao_Juan_2009"] = pick(res)
        curves.append((all_topics, res, "Inv_Cao_Juan_2009"))
    if use_coh:
        mimno = [models[i].metrics.loc["Metric", "Mimno_2011"] for i in in_index]
        mimno_topics = [all_topics[i] for i in in_index]
        raw["Minmo_2011"] = (mimno_topics, mimno, "Mimno_2011 - Maximize")
        res = _rescale(mimno)
        metrics_dict["Minmo_2011"] = np.array(res)
        curves.append((mimno_topics, res, "Mimno_2011"))
    if "loglikelihood" in metrics:
        ll = [m.metrics.loc["Metric", "loglikelihood"] for m in models]
        raw["loglikelihood"] = (all_topics, ll, "Loglikelihood - Maximize")
        res = _rescale(ll)
        metrics_dict["loglikelihood"] = pick(res)
        curves.append((all_topics, res, "Loglikelihood"))

    if select_model is None:
        combined_metric = sum(metrics_dict.values())
        pool = mimno_topics if use_coh else all_topics
        best_model = pool[combined_metric.tolist().index(max(combined_metric))]
    else:
        combined_metric = None
        best_model = select_model

    if plot or save is not None or plot_metrics:
        try:
            import matplotlib.backends.backend_pdf
            import matplotlib.pyplot as plt
        except ImportError:
            logging.getLogger("cisTopic").info("matplotlib is not installed: evaluate_models skips plotting")
        else:
            fig = plt.figure(figsize=figsize)
            for x, y, label in curves:
                plt.plot(x, y, linestyle="--", marker="o", label=label)
            plt.axvline(best_model, linestyle="--", color="grey")
            plt.xlabel("Number of topics\nOptimal number of topics: " + str(best_model))
            plt.ylabel("Rescaled metric")
            plt.legend(bbox_to_anchor=(1.04, 1), loc="upper left")
            pdf = None
            if save is not None:
                pdf = matplotlib.backends.backend_pdf.PdfPages(save)
                pdf.savefig(fig, bbox_inches="tight")
            if plot is True:
                plt.show()
            else:
                plt.close(fig)
            if plot_metrics:
                for name in ("Arun_2010", "Cao_Juan_2009", "Minmo_2011", "loglikelihood"):
                    if name in raw:
                        x, y, title = raw[name]
                        fig = plt.figure()
                        plt.plot(x, y, linestyle="--", marker="o")
                        plt.axvline(best_model, linestyle="--", color="grey")
                        plt.title(title)
                        if pdf is not None:
                            pdf.savefig(fig)
                        plt.show()
            if pdf is not None:
                pdf.close()

    if return_model:
        return models[all_topics.index(best_model)]


def load_cisTopic_model(path_to_cisTopic_model_matrices):
    """Feather-matrix loader of lda_models.py:1273-1289."""
    cell_topic = pd.read_feather(path_to_cisTopic_model_matrices + "cell_topic.feather")
    cell_topic.index = ["Topic" + str(x) for x in range(1, cell_topic.shape[0] + 1)]
    topic_region = pd.read_feather(path_to_cisTopic_model_matrices + "topic_region.feather")
    topic_region.index = ["Topic" + str(x) for x in range(1, topic_region.shape[0] + 1)]
    topic_region = topic_region.T
    return CistopicLDAModel(None, None, None, None, cell_topic, topic_region, None)
