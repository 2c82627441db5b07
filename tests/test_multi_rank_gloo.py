"""N>1 host logic on the CPU: world_size-2 gloo process group, static contiguous partition of the
index list, partial (nllh, gradient) per rank, one all-reduce -- the plumbing bench.py uses with
NCCL.  The per-rank partial sums come from the oracle here (no GPU in this test); the GPU test
`test_partition_sums_to_whole` checks the same property on the CUDA path."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle.binding import OracleProblem
    from parpe_b200 import capi, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tab = synth.make_problem("steadystate_scaled", 7, seed=9, condition_specific=1)
    theta = synth.random_theta(tab, seed=3, width=0.2)
    idx = np.array([6, 2, 5, 0, 1, 4, 3], dtype=np.int32)
    lo, hi = capi.partition_bounds(len(idx), rank, world)
    o = OracleProblem(tab)
    r = o.evaluate(theta, idx=idx[lo:hi], nthreads=1)
    buf = torch.zeros(2 + tab.n_opt, dtype=torch.float64)
    buf[0] = r["fval"]
    buf[1] = float(np.sum(r["cond_status"] != 0))
    buf[2:] = torch.from_numpy(r["grad"])
    dist.all_reduce(buf)
    if rank == 0:
        full = o.evaluate(theta, idx=idx, nthreads=1)
        np.save(out, np.concatenate([buf.numpy(), [full["fval"]], full["grad"]]))
    dist.destroy_process_group()


def test_two_ranks_sum_to_whole(oracle_available, tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "res.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    v = np.load(out)
    n = (len(v) - 3) // 2
    red, full = v[:2 + n], v[2 + n:]
    assert red[1] == 0.0
    assert red[0] == pytest.approx(full[0], rel=1e-12)
    np.testing.assert_allclose(red[2:], full[1:], rtol=1e-10, atol=1e-12)
