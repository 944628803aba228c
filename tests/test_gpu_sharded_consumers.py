"""The consumers of the imputed matrix sharded over devices (SURVEY.md 8e iii): regions are split into contiguous
ranges, one host thread + one device per range; normalize_scores needs the per-cell totals of the whole matrix, i.e.
one exchange of n_cells values between the shards; make_rankings shards by cells.  ``device=[0, 0, 0]`` runs three
shards on one GPU -- the sharding logic is the same -- and ``[0, 1]`` is added when the box has two GPUs.  Every
result must equal the single-device one (bit for bit for the int32 imputation, whose totals are exact integers)."""
import ctypes

import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


def _device_sets():
    from pycistopic_b200 import _lib
    n = ctypes.c_int(0)
    _lib.load().cgs_device_count(ctypes.byref(n))
    sets = [[0, 0, 0]]
    if n.value >= 2:
        sets.append([0, 1])
    return sets


def _factors(R=7000, C=300, K=9, seed=2):
    rng = np.random.default_rng(seed)
    tr = rng.dirichlet(np.full(R, 0.3), K).T.astype(np.float32)
    tr[rng.random(R) < 0.15] = 0
    ct = rng.dirichlet(np.full(K, 0.4), C).T.astype(np.float32)
    return tr, ct, [f"r{i}" for i in range(R)], [f"c{i}" for i in range(C)]


@pytest.mark.parametrize("scale", [10 ** 6, 1.0])
def test_sharded_consumers_equal_single_device(scale):
    from pycistopic_b200 import diff_features as pdf
    tr, ct, regions, cells = _factors()
    one = pdf.FactoredImputedFeatures(tr, ct, regions, cells, scale, 1000, "t", device=0)
    dense1 = one.to_dense()
    norm1 = one.normalized(10 ** 4)
    rank1 = one.make_rankings(seed=9)
    stats1, raw1 = one.row_mean_var(10 ** 4), one.row_mean_var()
    C = len(cells)
    contrasts = [(list(range(0, 100)), list(range(100, C))), (list(range(50, 250)), list(range(0, 50)))]
    rs1 = one.ranksum(contrasts)
    exact = isinstance(scale, int)
    for devices in _device_sets():
        many = pdf.FactoredImputedFeatures(tr, ct, regions, cells, scale, 1000, "t", device=devices)
        assert many.devices == devices
        dense = many.to_dense()
        assert dense.feature_names == dense1.feature_names
        np.testing.assert_array_equal(dense.mtx, dense1.mtx)
        norm = many.normalized(10 ** 4)
        assert norm.feature_names == norm1.feature_names and norm.mtx.dtype == norm1.mtx.dtype
        if exact:   # int64 totals: the shards' partial sums add up to exactly the single-device totals
            np.testing.assert_array_equal(norm.mtx, norm1.mtx)
        else:       # float totals are sums of doubles in another order
            np.testing.assert_allclose(norm.mtx, norm1.mtx, rtol=1e-5)
        for (m, v), (m1, v1) in ((many.row_mean_var(10 ** 4), stats1), (many.row_mean_var(), raw1)):
            if exact:
                np.testing.assert_array_equal(m, m1)
                np.testing.assert_array_equal(v, v1)
            else:
                np.testing.assert_allclose(m, m1, rtol=1e-5)
                np.testing.assert_allclose(v, v1, rtol=1e-4)
        rank = many.make_rankings(seed=9)
        np.testing.assert_array_equal(rank.mtx, rank1.mtx)
        rs = many.ranksum(contrasts)
        for a, b in zip(rs, rs1):
            np.testing.assert_array_equal(a, b)
        lazy = pdf.normalize_scores(many, 10 ** 4, materialize=False)
        assert pdf.find_highly_variable_features(lazy, plot=False) == \
            pdf.find_highly_variable_features(pdf.normalize_scores(one, 10 ** 4, materialize=False), plot=False)


def test_sharded_impute_accessibility_public_api():
    """impute_accessibility(device=[...]) / find_diff_features on the sharded factored object: same objects as on one
    device; more devices than chunks or regions is fine."""
    from pycistopic_b200 import diff_features as pdf
    tr, ct, regions, cells = _factors(R=2500, C=120, K=5, seed=4)
    labels = np.array(["A", "B", "C"])[np.arange(len(cells)) % 3]

    class Model:
        cell_topic = pd.DataFrame(ct, index=[f"Topic{i + 1}" for i in range(ct.shape[0])], columns=cells)
        topic_region = pd.DataFrame(tr, index=regions, columns=[f"Topic{i + 1}" for i in range(tr.shape[1])])

    class Obj:
        selected_model = Model()
        cell_names, region_names = cells, regions
        cell_data = pd.DataFrame({"celltype": labels}, index=cells)

    want = pdf.impute_accessibility(Obj(), scale_factor=10 ** 6, chunk_size=700)
    for devices in _device_sets() + [[0] * 7]:
        got = pdf.impute_accessibility(Obj(), scale_factor=10 ** 6, chunk_size=700, device=devices)
        assert got.feature_names == want.feature_names
        np.testing.assert_array_equal(got.mtx, want.mtx)
        fact = pdf.impute_accessibility(Obj(), scale_factor=10 ** 6, chunk_size=700, device=devices, materialize=False)
        assert fact.devices == devices
        a = pdf.find_diff_features(Obj(), fact, "celltype")
        b = pdf.find_diff_features(Obj(), want, "celltype")
        assert list(a) == list(b)
        for k in a:
            assert a[k].index.tolist() == b[k].index.tolist()
    tiny = pdf.FactoredImputedFeatures(tr[:2], ct, regions[:2], cells, 10 ** 6, 700, "t", device=[0, 0, 0, 0])
    assert tiny.to_dense().mtx.shape[0] <= 2 and len(tiny.row_mean_var(10 ** 4)[0]) == len(tiny.feature_names)


def _nccl_worker(rank, world, port, out_dir):
    import os
    import pickle
    import torch
    import torch.distributed as dist
    from pycistopic_b200 import sharded_consumers as sc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    tr, ct, _, _ = _factors()
    res = sc.distributed_normalized_row_stats(tr, ct, scale_factor=10 ** 6, norm_scale=10 ** 4,
                                              backend=sc.DeviceBackend(device=rank, chunk_size=1000))
    with open(os.path.join(out_dir, f"r{rank}.pkl"), "wb") as f:
        pickle.dump(res, f)
    dist.destroy_process_group()


def test_process_per_gpu_row_stats_over_nccl(tmp_path):
    """One process per GPU (pycistopic_b200/sharded_consumers.py): per-rank column totals, one NCCL all_reduce of
    n_cells values, per-rank statistics, all_gather -- equal to the single-device result on every rank."""
    import pickle
    import socket
    import torch
    import torch.multiprocessing as mp
    from pycistopic_b200 import diff_features as pdf
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    tr, ct, regions, cells = _factors()
    one = pdf.FactoredImputedFeatures(tr, ct, regions, cells, 10 ** 6, 1000, "t", device=0)
    mean1, var1 = one.row_mean_var(10 ** 4)
    for r in range(2):
        with open(tmp_path / f"r{r}.pkl", "rb") as f:
            mean, var, flags = pickle.load(f)
        np.testing.assert_array_equal(flags, one.row_nonzero)
        np.testing.assert_array_equal(mean[flags], mean1)
        np.testing.assert_array_equal(var[flags], var1)
