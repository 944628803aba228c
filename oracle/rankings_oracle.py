"""CPU oracle of the per-cell rankings.  TEST INFRASTRUCTURE ONLY.

``CistopicImputedFeatures.make_rankings`` (src/pycisTopic/diff_features.py:274-337) ranks every column of the
regions x cells matrix by descending score, 0 = best, breaking ties with a random permutation:

    perm = rng.permutation(R);  ranking = perm[(-scores)[perm].argsort()].argsort()        (:294-306)

``reference_make_rankings`` is that closure verbatim in behaviour (NumPy's default_rng + default argsort), used on
columns WITHOUT ties -- where the ranking is unique -- to pin the device path bit for bit, and on columns with ties
for the properties every valid draw has.  NumPy's own tie order (PCG64 stream + unstable introsort) is not
reproducible outside NumPy, so the device path uses its own keyed bijection instead of ``rng.permutation``;
``rank_columns`` restates THAT algorithm (csrc/rankings_kernels.cuh: 4-round Feistel network with cycle walking,
then a stable sort by descending score) so the GPU result can be checked bit for bit, ties included.
"""
from __future__ import annotations

import numpy as np

_M32 = np.uint64(0xFFFFFFFF)


def _mix32(x):
    x = np.asarray(x, np.uint64) & _M32
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7FEB352D)) & _M32
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846CA68B)) & _M32
    x ^= x >> np.uint64(16)
    return x


def column_key(seed: int, col: int) -> int:
    seed, col = int(seed) & (2 ** 64 - 1), int(col) & (2 ** 64 - 1)
    k = int(_mix32((seed & 0xFFFFFFFF) ^ 0x9E3779B9))
    k = int(_mix32(k ^ (seed >> 32)))
    k = int(_mix32(k ^ (col & 0xFFFFFFFF)))
    return int(_mix32((k + (col >> 32)) & 0xFFFFFFFF))


def half_bits(n: int) -> int:
    h = 1
    while h < 16 and (1 << (2 * h)) < n:
        h += 1
    return h


def permutation(n: int, key: int) -> np.ndarray:
    """pi(0 .. n-1): 4 Feistel rounds on 2h bits, walked until the image is < n (rk_permute)."""
    h = half_bits(n)
    mask = np.uint64((1 << h) - 1)
    hh = np.uint64(h)
    x = np.arange(n, dtype=np.uint64)
    out = np.empty(n, dtype=np.uint64)
    todo = np.arange(n)
    first = True
    while todo.size:
        l, r = x >> hh, x & mask
        for i in range(1, 5):
            rk = np.uint64((key + 0x9E3779B9 * i) & 0xFFFFFFFF)
            f = _mix32(r ^ rk) & mask
            l, r = r, l ^ f
        x = (l << hh) | r
        done = x < np.uint64(n)
        out[todo[done]] = x[done]
        todo, x = todo[~done], x[~done]
    return out.astype(np.int64)


def rank_columns(x: np.ndarray, seed: int = 123, first_col: int = 0) -> np.ndarray:
    """The device algorithm on the CPU: read column c through pi_(seed, first_col + c), sort stably by descending
    score (NaN last), ranking[row] = final position."""
    x = np.asarray(x)
    n_rows, n_cols = x.shape
    out = np.empty((n_rows, n_cols), dtype=np.int32)
    for c in range(n_cols):
        perm = permutation(n_rows, column_key(seed, first_col + c))
        col = x[perm, c]
        if np.issubdtype(col.dtype, np.floating):
            order = np.argsort(-col, kind="stable")          # NaN (also -NaN) sorts to the end
        else:
            order = np.argsort(-col.astype(np.int64), kind="stable")
        out[perm[order], c] = np.arange(n_rows, dtype=np.int32)
    return out


def reference_make_rankings(x: np.ndarray, seed: int = 123) -> np.ndarray:
    """diff_features.py:289-335 restated: one rng for all columns, permutation + double argsort per column."""
    rng = np.random.default_rng(seed=seed)
    x = np.asarray(x)
    out = np.zeros(x.shape, dtype=np.int32)
    for c in range(x.shape[1]):
        scores = x[:, c]
        perm = rng.permutation(scores.shape[0])
        out[:, c] = perm[(-scores)[perm].argsort()].argsort().astype("uint32")
    return out
