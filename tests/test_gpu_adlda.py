"""Multi-GPU AD-LDA on real GPUs (needs >= 2 devices; skipped on a 1-GPU box): NCCL all-reduce (through
torch.distributed and behind the C ABI), the peer-memory apply kernel and the asynchronous fused round kernel
must keep the count tables exact and identical on all ranks, and the log-likelihood trace must follow the
AD-LDA oracle with the same partition within 1 %.  (tests/test_gpu_exchange.py runs the same round kernel with
several ranks on ONE device, for boxes with a single GPU.)"""
import os
import pickle
import socket
import sys
import tempfile

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import lda_oracle as lo  # noqa: E402
from pycistopic_b200.synth import make_binary_matrix  # noqa: E402

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir, exchange, sync_per_sweep, K, n_iter):
    import torch
    import torch.distributed as dist

    from pycistopic_b200 import adlda
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["LOCAL_RANK"] = str(rank)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    m, _, _ = make_binary_matrix("C1")
    res = adlda.fit_adlda(m, n_topics=K, n_iter=n_iter, alpha=50.0 / K, eta=0.1, random_state=555, refresh=1,
                          sync_per_sweep=sync_per_sweep, exchange=exchange, device=rank)
    with open(os.path.join(out_dir, f"r{rank}.pkl"), "wb") as f:
        pickle.dump({k: getattr(res, k) for k in ("nzw_", "ndz_", "nz_", "loglikelihoods_", "final_loglikelihood_",
                                                  "part_ptr_")}, f)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("exchange,sync_per_sweep", [("nccl", 1), ("peer", 1), ("peer", 4), ("nccl_abi", 1), ("fused", 1), ("overlap", 1),
                                                     ("async", 2), ("async", 4)])
def test_adlda_two_gpus(exchange, sync_per_sweep):
    import torch.multiprocessing as mp
    K, n_iter, world = 10, 30, 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), d, exchange, sync_per_sweep, K, n_iter), nprocs=world, join=True)
        outs = [pickle.load(open(os.path.join(d, f"r{r}.pkl"), "rb")) for r in range(world)]
    m, _, _ = make_binary_matrix("C1")
    X = sp.csr_matrix(m.T)
    a, b = outs
    assert np.array_equal(a["nzw_"], b["nzw_"]) and np.array_equal(a["ndz_"], b["ndz_"]) and np.array_equal(a["nz_"], b["nz_"])
    assert a["nz_"].sum() == X.nnz and np.array_equal(a["nzw_"].sum(axis=1), a["nz_"])
    assert np.array_equal(a["ndz_"].sum(axis=1), np.diff(X.indptr))
    assert np.array_equal(a["nzw_"].sum(axis=0), np.asarray(X.sum(axis=0)).ravel())
    o = lo.OracleLDA(K, n_iter=n_iter, alpha=50.0 / K, eta=0.1, random_state=555, refresh=1, n_parts=world,
                     part_ptr=a["part_ptr_"]).fit(X)
    ot = np.array(o.loglikelihoods_ + [o.final_loglikelihood_])
    gt = np.array(a["loglikelihoods_"] + [a["final_loglikelihood_"]])
    assert gt[0] == pytest.approx(ot[0], rel=1e-10)
    rel = (gt - ot) / np.abs(ot)
    if sync_per_sweep == 1:
        assert np.abs(rel).max() < 0.01, f"trace deviates {np.abs(rel).max():.3%} at {np.abs(rel).argmax()}"
    else:
        # more exchanges per sweep can only bring the chain closer to the sequential one (never worse than -1 %)
        assert rel.min() > -0.01
