"""Hierarchical evaluation of the Fujita-shape problem (BASELINE config 2: 6 doses tiled to C conditions, three
scalings and three sigmas analytic), one GPU or -- under torchrun -- one rank per GPU with the conditions
partitioned and the collective hook (pb200_set_allreduce) served by torch.distributed / NCCL.

    python tools/benchmarks/hier_bench.py [C]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/benchmarks/hier_bench.py [C]
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from parpe_b200 import capi, synth  # noqa: E402
from parpe_b200.problem import B200SummedGradientFunction  # noqa: E402

C = int(sys.argv[1]) if len(sys.argv) > 1 else 10002
world = int(os.environ.get("WORLD_SIZE", 1))
rank = int(os.environ.get("RANK", 0))
local_rank = int(os.environ.get("LOCAL_RANK", 0))
tab = synth.fujita_hierarchical(C, seed=0)
gpu = B200SummedGradientFunction(tab, device=local_rank)
if world > 1:
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    gpu.set_partition(rank, world)

    def allreduce(arr, op):
        t = torch.from_numpy(arr).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM if op == capi.REDUCE_SUM else dist.ReduceOp.MAX)
        arr[:] = t.cpu().numpy()
    gpu.set_allreduce(allreduce)
n_outer = gpu.hier_num_outer()
full = synth.random_theta(tab, seed=1, width=0.3)
outer = np.delete(full, tab.hier.inner_idx)
assert len(outer) == n_outer
for rep in range(4):
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    r = gpu.hier_evaluate(outer, want_gradient=True)
    wall = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps(dict(n_gpus=world, conditions=C, status=r["status"], fval=r["fval"], wall_s=wall,
                              solves_per_s_wall=2 * C / wall)))
t0 = time.perf_counter()
r0 = gpu.hier_evaluate(outer, want_gradient=False)
wall0 = time.perf_counter() - t0
if rank == 0:
    print(json.dumps(dict(cost_only=True, fval=r0["fval"], wall_s=wall0, solves_per_s_wall=C / wall0)))
if world == 1:
    from oracle import binding as ob
    if ob.available():
        o = ob.OracleProblem(tab)
        t0 = time.perf_counter()
        ro = o.hier_evaluate(outer, want_grad=True)
        wo = time.perf_counter() - t0
        g, go = r["grad"], ro["grad"]
        gs = np.max(np.abs(go))
        print(json.dumps(dict(oracle_wall_s=wo, oracle_solves_per_s=2 * C / wo,
                              rel_fval=abs(r["fval"] - ro["fval"]) / abs(ro["fval"]),
                              rel_grad=float(np.max(np.abs(g - go) / np.maximum(np.abs(go), 1e-4 * gs))))))
else:
    dist.destroy_process_group()
