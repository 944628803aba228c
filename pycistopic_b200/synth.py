"""Seeded planted-topic binary accessibility matrices.

The reference ships no data; the benchmark shapes of BASELINE.json are filled with a
planted-topic model: ``K*`` ground-truth topics over ``W`` regions, ``phi_k ~ Dirichlet(0.05)``
modulated by a Zipf(1.0) region-popularity prior, ``theta_d ~ Dirichlet(0.3)``, per-cell token
budget ``n_d ~ LogNormal(log(density * W), 0.4)``, tokens drawn from the mixture and
de-duplicated into a binary matrix (a region is accessible in a cell or it is not, as
``sklearn.preprocessing.binarize`` leaves it at cistopic_class.py:592).

The returned matrix is scipy CSR **regions x cells**, int32 -- the layout of
``CistopicObject.binary_matrix`` that ``run_cgs_models`` transposes (lda_models.py:151).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sparse

# (cells, regions, density) of the configurations named in BASELINE.json
SHAPES = {
    "C1": (500, 50_000, 0.03),
    "C2": (10_000, 200_000, 0.03),
    "C3": (35_000, 500_000, 0.02),
    "C4": (200_000, 1_000_000, 0.01),
    "T": (20_000, 300_000, 0.03),
}


def planted_topic_tokens(n_cells, n_regions, density, n_true_topics=20, seed=1234, device=None, max_rounds=12,
                         cell_range=None):
    """Return (cell_ids, region_ids) int64 arrays of UNIQUE (cell, region) pairs sorted by
    (cell, region).  Runs in NumPy, or in torch on ``device`` when given (same algorithm,
    different generator -- the result is only reproducible per backend).  Cells keep drawing
    from their topic mixture until they hold ``n_d`` distinct regions (at most ``max_rounds``
    rounds), so the matrix reaches the requested density despite the heavy-tailed region
    popularity."""
    if device is not None:
        return _planted_torch(n_cells, n_regions, density, n_true_topics, seed, device, max_rounds, cell_range)
    if cell_range is not None:
        raise ValueError("cell_range needs the torch backend (device=...)")
    rng = np.random.default_rng(seed)
    W, D, Kt = n_regions, n_cells, n_true_topics
    pop = 1.0 / np.arange(1, W + 1) ** 1.0
    pop = pop[rng.permutation(W)]
    phi = rng.gamma(0.05, 1.0, size=(Kt, W)) + 1e-12
    phi *= pop
    phi /= phi.sum(axis=1, keepdims=True)
    cdf = np.cumsum(phi, axis=1)
    cdf[:, -1] = 1.0
    theta = rng.dirichlet(np.full(Kt, 0.3), size=D)
    tcdf = np.cumsum(theta, axis=1)
    tcdf[:, -1] = 1.0
    n_d = np.clip(np.rint(rng.lognormal(np.log(density * W), 0.4, size=D)), 1, W // 2).astype(np.int64)
    key = np.empty(0, np.int64)
    have = np.zeros(D, np.int64)
    for _ in range(max_rounds):
        missing = n_d - have
        if not (missing > 0).any():
            break
        n_draw = np.where(missing > 0, missing + missing // 2 + 8, 0)
        cell = np.repeat(np.arange(D, dtype=np.int64), n_draw)
        u = rng.random(cell.shape[0])
        topic = np.empty(cell.shape[0], np.int64)
        step = 2_000_000
        for s0 in range(0, cell.shape[0], step):
            e0 = min(s0 + step, cell.shape[0])
            topic[s0:e0] = (u[s0:e0, None] > tcdf[cell[s0:e0]]).sum(axis=1)
        region = np.empty(cell.shape[0], np.int64)
        u2 = rng.random(cell.shape[0])
        for k in range(Kt):
            msk = topic == k
            region[msk] = np.searchsorted(cdf[k], u2[msk], side="left")
        region = np.minimum(region, W - 1)
        key = np.unique(np.concatenate([key, cell * W + region]))
        have = np.bincount(key // W, minlength=D)
    cell, region = key // W, key % W
    # trim every cell back to its budget n_d (keep a random subset; order stays sorted)
    cnt = np.bincount(cell, minlength=D)
    start = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    rank = rng.random(key.shape[0])
    order = np.lexsort((rank, cell))
    pos_in_cell = np.arange(key.shape[0]) - np.repeat(start, cnt)
    keep_sorted = pos_in_cell < np.repeat(n_d, cnt)
    keep = np.zeros(key.shape[0], bool)
    keep[order] = keep_sorted
    return cell[keep], region[keep]


def _planted_torch(n_cells, n_regions, density, n_true_topics, seed, device, max_rounds, cell_range=None):
    """``cell_range=(c0, c1)`` draws only those cells (same topics and budgets as the full matrix, its own
    random stream): lets every rank of a sharded run generate its own cells of an atlas-scale matrix."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed)
    W, D, Kt = n_regions, n_cells, n_true_topics
    f64 = torch.float64
    pop = 1.0 / torch.arange(1, W + 1, device=device, dtype=f64)
    pop = pop[torch.randperm(W, device=device, generator=g)]
    torch.manual_seed(seed)
    gam = torch.distributions.Gamma(torch.tensor(0.05, device=device, dtype=f64),
                                    torch.tensor(1.0, device=device, dtype=f64))
    phi = gam.sample((Kt, W)) + 1e-12
    phi = phi * pop
    phi = phi / phi.sum(dim=1, keepdim=True)
    cdf = torch.cumsum(phi, dim=1)
    cdf[:, -1] = 1.0
    flat_r = (cdf + torch.arange(Kt, device=device, dtype=f64)[:, None] * 2.0).reshape(-1)
    gam_t = torch.distributions.Gamma(torch.tensor(0.3, device=device, dtype=f64),
                                      torch.tensor(1.0, device=device, dtype=f64))
    theta = gam_t.sample((D, Kt)) + 1e-300
    theta = theta / theta.sum(dim=1, keepdim=True)
    tcdf = torch.cumsum(theta, dim=1)
    tcdf[:, -1] = 1.0
    n_d = torch.clamp(torch.round(torch.exp(torch.randn(D, device=device, generator=g, dtype=f64) * 0.4
                                            + float(np.log(density * W)))), 1, W // 2).to(torch.int64)
    max_tok = 48_000_000  # draws per chunk (bounds the temporaries)
    keys = []
    c0 = 0
    c_end = D
    if cell_range is not None:
        c0, c_end = int(cell_range[0]), int(cell_range[1])
        g.manual_seed(seed + 7919 * (c0 + 1))
    budget_ptr = torch.cumsum(n_d, 0)
    while c0 < c_end:
        base = int(budget_ptr[c0 - 1]) if c0 > 0 else 0
        c1 = int(torch.searchsorted(budget_ptr, torch.tensor(base + max_tok // 2, device=device)).item())
        c1 = max(c0 + 1, min(c1, c_end))
        nc = c1 - c0
        flat_t = (tcdf[c0:c1] + torch.arange(nc, device=device, dtype=f64)[:, None] * 2.0).reshape(-1)
        key = torch.empty(0, dtype=torch.int64, device=device)
        have = torch.zeros(nc, dtype=torch.int64, device=device)
        for _ in range(max_rounds):
            missing = n_d[c0:c1] - have
            if not bool((missing > 0).any()):
                break
            n_draw = torch.where(missing > 0, missing + missing // 2 + 8, torch.zeros_like(missing))
            lc = torch.repeat_interleave(torch.arange(nc, device=device), n_draw)  # local cell id
            u = torch.rand(lc.shape[0], device=device, generator=g, dtype=f64)
            topic = torch.searchsorted(flat_t, u + lc.to(f64) * 2.0) - lc * Kt
            topic = torch.clamp(topic, 0, Kt - 1)
            u2 = torch.rand(lc.shape[0], device=device, generator=g, dtype=f64)
            region = torch.searchsorted(flat_r, u2 + topic.to(f64) * 2.0) - topic * W
            region = torch.clamp(region, 0, W - 1)
            key = torch.unique(torch.cat([key, (lc + c0) * W + region]))
            have = torch.bincount(key // W - c0, minlength=nc)
        kc = key // W - c0
        cnt = torch.bincount(kc, minlength=nc)
        rank = torch.rand(key.shape[0], device=device, generator=g, dtype=f64)
        order = torch.argsort(kc.to(f64) * 2.0 + rank)
        start = torch.cumsum(cnt, 0) - cnt
        pos = torch.arange(key.shape[0], device=device) - torch.repeat_interleave(start, cnt)
        keep_sorted = pos < torch.repeat_interleave(n_d[c0:c1], cnt)
        keep = torch.zeros(key.shape[0], dtype=torch.bool, device=device)
        keep[order] = keep_sorted
        keys.append(key[keep])
        c0 = c1
    key = torch.cat(keys)
    return key // W, key % W


def make_binary_matrix(shape="C1", n_cells=None, n_regions=None, density=None, n_true_topics=20,
                       seed=1234, drop_empty=True):
    """scipy CSR (regions x cells) int32 binary matrix + (cell_names, region_names)."""
    if n_cells is None:
        n_cells, n_regions, density = SHAPES[shape]
    cell, region = planted_topic_tokens(n_cells, n_regions, density, n_true_topics, seed)
    data = np.ones(cell.shape[0], np.int32)
    m = sparse.csr_matrix((data, (region, cell)), shape=(n_regions, n_cells), dtype=np.int32)
    if drop_empty:
        keep_r = np.flatnonzero(np.diff(m.indptr) > 0)
        m = m[keep_r]
        keep_c = np.flatnonzero(np.asarray(m.sum(axis=0)).ravel() > 0)
        m = sparse.csr_matrix(m[:, keep_c])
    else:
        keep_r = np.arange(n_regions)
        keep_c = np.arange(n_cells)
    m.sort_indices()
    cell_names = ["cell_%d" % i for i in keep_c]
    region_names = ["chr1:%d-%d" % (int(i) * 500, int(i) * 500 + 500) for i in keep_r]
    return m, cell_names, region_names


class SyntheticCistopicObject:
    """Duck-typed stand-in for ``CistopicObject``: only the attributes the hot path reads
    (lda_models.py:151-153, diff_features.py:380-384)."""

    def __init__(self, binary_matrix, cell_names, region_names, project="synthetic"):
        self.binary_matrix = binary_matrix
        self.fragment_matrix = binary_matrix
        self.cell_names = list(cell_names)
        self.region_names = list(region_names)
        self.selected_model = None
        self.project = project

    def add_LDA_model(self, model):
        # cistopic_class.py:491-504 (reorders to the object's names)
        model.cell_topic = model.cell_topic.loc[:, self.cell_names]
        model.topic_region = model.topic_region.loc[self.region_names]
        self.selected_model = model
