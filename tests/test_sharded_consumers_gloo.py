"""world_size-2 / 3 `gloo` test on the CPU of the process-per-GPU consumers (pycistopic_b200/sharded_consumers.py,
SURVEY.md 8e iii): the device backend is replaced by a NumPy stand-in with the same two methods, and the distributed
run -- per-rank column totals, ONE all_reduce of n_cells values, per-rank statistics, all_gather -- must reproduce the
single-process NumPy computation of normalize_scores + mean / var on the whole matrix, on every rank."""
import os
import pickle
import socket
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pycistopic_b200 import sharded_consumers as sc  # noqa: E402


class NumpyBackend:
    """CPU stand-in for DeviceBackend (test infrastructure): the reference's arithmetic in NumPy
    (diff_features.py:443-464 imputation, :540-548 normalisation, :636-639 mean / var)."""

    @staticmethod
    def _impute(topic_region, cell_topic, impute_scale, as_int32):
        x = topic_region @ cell_topic
        return (x * np.float32(impute_scale)).astype(np.int32) if as_int32 else x

    def column_totals(self, topic_region, cell_topic, impute_scale, as_int32):
        x = self._impute(topic_region, cell_topic, impute_scale, as_int32)
        totals = x.sum(axis=0, dtype=np.int64 if as_int32 else np.float64)
        return totals, (np.count_nonzero(x, axis=1) != 0).astype(np.uint8)

    def row_stats(self, topic_region, cell_topic, impute_scale, as_int32, norm_scale, totals):
        x = self._impute(topic_region, cell_topic, impute_scale, as_int32)
        norm = np.log1p(x / (totals / norm_scale))
        if not as_int32:
            norm = norm.astype(np.float32)
        return np.mean(norm, axis=1, dtype=np.float32), np.var(norm, axis=1, dtype=np.float32)


def _factors(R, C=23, K=4, seed=3):
    rng = np.random.default_rng(seed)
    tr = rng.dirichlet(np.full(R, 0.4), K).T.astype(np.float32)
    tr[rng.random(R) < 0.2] = 0
    ct = rng.dirichlet(np.full(K, 0.5), C).T.astype(np.float32)
    return tr, ct


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, R, scale):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tr, ct = _factors(R)
    res = sc.distributed_normalized_row_stats(tr, ct, scale_factor=scale, norm_scale=10 ** 4, backend=NumpyBackend())
    with open(os.path.join(out_dir, f"r{rank}.pkl"), "wb") as f:
        pickle.dump(res, f)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,R,scale", [(2, 301, 10 ** 6), (3, 301, 10 ** 6), (2, 300, 1.0), (3, 2, 10 ** 6)])
def test_distributed_row_stats_equal_single_process(world, R, scale):
    import torch.multiprocessing as mp
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d, R, scale), nprocs=world, join=True)
        outs = [pickle.load(open(os.path.join(d, f"r{r}.pkl"), "rb")) for r in range(world)]
    tr, ct = _factors(R)
    as_int = isinstance(scale, int)
    b = NumpyBackend()
    totals, flags = b.column_totals(tr, ct, float(np.float32(scale)) if as_int else 1.0, as_int)
    mean, var = b.row_stats(tr, ct, float(np.float32(scale)) if as_int else 1.0, as_int, 10 ** 4, totals)
    for m, v, f in outs:  # every rank holds the full, identical result
        assert m.shape == (R,) and f.dtype == bool
        np.testing.assert_array_equal(f, flags != 0)
        if as_int:   # exact int64 totals: the sharded normalisation IS the global one
            np.testing.assert_array_equal(m, mean)
            np.testing.assert_array_equal(v, var)
        else:        # float totals are partial sums in another order
            np.testing.assert_allclose(m, mean, rtol=1e-5)
            np.testing.assert_allclose(v, var, rtol=1e-4, atol=1e-12)
    assert all(sc.region_range(R, r, world)[1] == sc.region_range(R, r + 1, world)[0] for r in range(world - 1))
    assert sc.region_range(R, 0, world)[0] == 0 and sc.region_range(R, world - 1, world)[1] == R
