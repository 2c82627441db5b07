#!/usr/bin/env python
"""bench.py -- forward-sensitivity ODE solves/sec per objective+gradient evaluation.

A "step" is one pass of the hot path -- AmiciSummedGradientFunction::evaluate(parameters,
dataIndices, fval, gradient) -- over one batch of synthetic conditions.

  python bench.py --gpus N --steps K --warmup W            (our arm, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  (reference CPU path, rank 0)

Workload (BASELINE.json configs[1]): Boehm_JProteomeRes2014-shape model (8 states, 9
parameters, 3 observables, 16 output times), forward sensitivities w.r.t. all 9
parameters, rtol 1e-8 / atol 1e-12, random parameters in bounds, CONDITIONS_PER_GPU (18 944)
conditions per GPU (weak scaling: the global problem has N x that many conditions, each
rank simulates its contiguous share and one NCCL all-reduce combines nllh and gradient).
The real PEtab problem is not vendored in the reference tree: model structure is a
stand-in of the published shape and the measurement table is synthetic (say so in `data`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "forward-sensitivity ODE solves/sec per obj+grad eval"
UNIT = "solves/s"
MODEL = "boehm"   # --workload: boehm (the metric's configuration, default) | chen | froehlich | net32 | fujita ...
DEFAULT_CONDITIONS = {"chen": 296, "froehlich": 148, "net32": 4736}
# conditions per GPU: four resident waves of the small-model kernel on a B200 (148 SMs x 32 conditions per SM);
# any number works -- the kernel pulls conditions from a queue -- but a partial last wave costs a whole one
# (16 384 conditions = 3.46 waves run at 145 k solves/s, whole waves at 163 k)
CONDITIONS_PER_GPU = int(os.environ.get("PB200_BENCH_CONDITIONS", 4 * 148 * 32))
LARGE_MODELS = ("chen", "froehlich", "net32", "net31")
CPU_SAMPLE_SECONDS = float(os.environ.get("PB200_BENCH_CPU_SECONDS", 15.0))


def workload_name(n_gpus):
    if MODEL != "boehm":
        from parpe_b200.models import get_spec
        sp = get_spec(MODEL)
        shape = {"chen": "Chen_MSB2009-shape", "froehlich": "Froehlich_CellSystems2018-shape"}.get(MODEL, MODEL)
        return ("%s generated mass-action network (nx=%d, np=%d, ny=%d, nt=8), forward sensitivities, obj+grad, "
                "%d conditions/GPU x %d GPU" % (shape, sp.nx, sp.np, sp.ny, CONDITIONS_PER_GPU, n_gpus)
                if MODEL in LARGE_MODELS else "%s, %d conditions/GPU x %d GPU" % (MODEL, CONDITIONS_PER_GPU, n_gpus))
    return ("Boehm_JProteomeRes2014-shape (nx=8, np=9, ny=3, nt=16), forward sensitivities, obj+grad, "
            "%d conditions/GPU x %d GPU" % (CONDITIONS_PER_GPU, n_gpus))


def make_tables(n_conditions):
    from parpe_b200 import synth
    tab = synth.make_problem(MODEL, n_conditions, scale="log10", seed=0)
    if os.environ.get("PB200_BENCH_STLDET"):  # experiments only; AMICI's default (on) is the benchmark setting
        tab.stldet = int(os.environ["PB200_BENCH_STLDET"])
    theta = synth.random_theta(tab, seed=1, width=0.5)
    return tab, theta


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def algorithmic_flops(stats, ops_info, nx, nplist):
    """Flops of one launch from the integrator statistics -- the counting model of SURVEY.md
    section 8(d): F = nfe*F_f + nje*F_J + nsetups*(2/3 nx^3 + nx^2) + nni*2nx^2 +
    nniS*nplist*2nx^2 + nfSe*(2 nnzJ + F_dfdp/nplist) + V, with the vector work
    V = nst*(q+1)(q+2)*nx*(1+nplist) + nst*2(q+1)*nx*(1+nplist) + (nni + nniS*nplist)*6nx."""
    from parpe_b200.tables import STAT_NAMES
    s = {n: stats[:, i].astype(np.float64) for i, n in enumerate(STAT_NAMES)}
    q = np.where(s["nst"] > 0, s["qsum"] / np.maximum(s["nst"], 1), 1.0)
    nc = 1 + nplist
    F = (s["nfe"] * ops_info["xdot"] + s["nje"] * ops_info["J"] + s["nsetups"] * (2.0 / 3.0 * nx ** 3 + nx ** 2)
         + s["nni"] * 2 * nx ** 2 + s["nniS"] * nplist * 2 * nx ** 2
         + s["nfSe"] * (2 * ops_info["nnzJ"] + ops_info["dxdotdp"] / max(nplist, 1)))
    V = (s["nst"] * (q + 1) * (q + 2) * nx * nc + s["nst"] * 2 * (q + 1) * nx * nc
         + (s["nni"] + s["nniS"] * nplist) * 6 * nx)
    return float(np.sum(F + V)), {k: float(np.mean(v)) for k, v in s.items()}


def algorithmic_bytes(tab, n, want_grad=True):
    """Small-nx regime: inputs + outputs only:
    8*(np + nk + nt + 2*nt*ny + 1 + nplist) per solve."""
    nt = tab.nt(0)
    return float(n) * 8.0 * (tab.np_ + tab.nk + nt + 2 * nt * tab.ny + 1 + (tab.np_ if want_grad else 0))


def model_op_counts():
    from parpe_b200.codegen.modelgen import count_flops, derive
    from parpe_b200.models import get_spec
    if MODEL in LARGE_MODELS:
        return get_spec(MODEL).flops()
    return count_flops(derive(get_spec(MODEL)))


def algorithmic_bytes_large(stats, nx, nplist, nnz, tab, n):
    """Large-nx regime (state not cache-resident), SURVEY.md 8(d):
    B = 8*[nst*(1+nplist)*nx*(2(q+1) + 2(q+1) + 6) + (nni + nniS)*nx^2 + nsetups*2*nx^2 + nje*nnzJ]
        + inputs + outputs."""
    from parpe_b200.tables import STAT_NAMES
    s = {nm: stats[:, i].astype(np.float64) for i, nm in enumerate(STAT_NAMES)}
    q = np.where(s["nst"] > 0, s["qsum"] / np.maximum(s["nst"], 1), 1.0)
    B = 8.0 * (s["nst"] * (1 + nplist) * nx * (4 * (q + 1) + 6) + (s["nni"] + s["nniS"]) * nx ** 2
               + s["nsetups"] * 2 * nx ** 2 + s["nje"] * nnz)
    return float(np.sum(B)) + algorithmic_bytes(tab, n)


def cpu_baseline(tab, theta, seconds=CPU_SAMPLE_SECONDS, nthreads=0):
    """Reference AMICI/CVODES path (oracle/_ref, KLU, staggered) + restated parPE glue, OpenMP over
    conditions on this box's host cores, on a bounded sample of the same workload."""
    from oracle import binding as ob
    if not ob.available():
        return None
    ncores = os.cpu_count() or 1
    o = ob.OracleProblem(tab, ob.ISM_STAGGERED, ob.LINSOL_KLU)
    n0 = min(tab.n_conditions, (1 if MODEL in LARGE_MODELS else 8) * ncores)
    t0 = time.perf_counter()
    o.evaluate(theta, idx=np.arange(n0, dtype=np.int32), nthreads=nthreads)
    dt = time.perf_counter() - t0
    rate = n0 / dt
    n = int(min(tab.n_conditions, max(n0, rate * seconds)))
    t0 = time.perf_counter()
    r = o.evaluate(theta, idx=np.arange(n, dtype=np.int32), nthreads=nthreads)
    dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": ncores, "kind": "reference",
            "sample": "%d of %d conditions, one obj+grad evaluation, %.1f s; AMICI 0.31.0/CVODES 7.1.1 (KLU, "
                      "staggered) built from the reference sources, OpenMP dynamic over conditions" % (
                          n, tab.n_conditions, dt),
            "status": int(r["status"])}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    n_gpus = args.gpus
    tab, theta = make_tables(CONDITIONS_PER_GPU * n_gpus)
    from oracle import binding as ob
    ncores = os.cpu_count() or 1
    o = ob.OracleProblem(tab, ob.ISM_STAGGERED, ob.LINSOL_KLU)
    # bounded sample per step: sized from a probe so that the whole run ends within minutes
    n0 = min(tab.n_conditions, (1 if MODEL in LARGE_MODELS else 8) * ncores)
    t0 = time.perf_counter()
    o.evaluate(theta, idx=np.arange(n0, dtype=np.int32))
    rate = n0 / (time.perf_counter() - t0)
    budget = 120.0 / max(args.steps + args.warmup, 1)
    n = int(min(tab.n_conditions, max(n0, rate * min(budget, 20.0))))
    idx = np.arange(n, dtype=np.int32)
    for _ in range(args.warmup):
        o.evaluate(theta, idx=idx)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = o.evaluate(theta, idx=idx)
    dt = time.perf_counter() - t0
    value = args.steps * n / dt
    sample = ("%d of %d conditions per step (bounded sample of the same workload); AMICI 0.31.0 / CVODES 7.1.1 "
              "(KLU, staggered) compiled from the reference sources + restated parPE glue, OpenMP over "
              "conditions" % (n, tab.n_conditions))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": workload_name(n_gpus), "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": ncores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "status": int(r["status"])}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lanes", type=int, default=0, help="lanes per condition (0 = model default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="boehm",
                    help="boehm = the configuration the metric is quoted on (default); chen / froehlich = the "
                         "large-model path (one CTA per condition), reported for DESIGN.md, not the driver's line")
    args = ap.parse_args()
    global MODEL, CONDITIONS_PER_GPU
    MODEL = args.workload
    if "PB200_BENCH_CONDITIONS" not in os.environ:
        CONDITIONS_PER_GPU = DEFAULT_CONDITIONS.get(MODEL, CONDITIONS_PER_GPU)
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from parpe_b200 import capi
    from parpe_b200.problem import B200SummedGradientFunction

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE %d != --gpus %d" % (world, args.gpus))
    if capi.device_count() <= 0:
        raise SystemExit("no CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n_total = CONDITIONS_PER_GPU * world
    tab, theta = make_tables(n_total)
    tab.lanes_per_condition = args.lanes
    fun = B200SummedGradientFunction(tab, device=local_rank)
    fun.set_partition(rank, world)
    idx = np.arange(n_total, dtype=np.int32)
    n_opt = tab.n_opt
    d_out = torch.zeros(2 + n_opt, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step_device():
        fun.evaluate_device(theta, idx, True, d_out.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_reduce(d_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput (`value`) --------------------------------------
    # one evaluation through the host API first: it leaves the cost-aware processing order (previous
    # evaluation's step counts) in the handle, which the device-resident path then uses as the
    # host path does from its second evaluation on
    fun.evaluate(theta, idx)
    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_s = []
    wall0 = time.perf_counter()
    for k in range(args.steps):
        flush.fill_(k & 0xFF)  # L2 flush between timed iterations (untimed)
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
        if world == 1:
            torch.cuda.synchronize()
            kernel_s.append(fun.last_launch()["seconds"])
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_s = torch.tensor([sum(step_ms) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_s, op=dist.ReduceOp.MAX)
    total_s = float(total_s.item())
    out_host = d_out.cpu().numpy()
    value = args.steps * n_total / total_s

    # ---- objective only (gradient.empty() in the reference): same workload, no sensitivities ----
    def step_cost_only():
        fun.evaluate_device(theta, idx, False, d_out.data_ptr(), stream.cuda_stream)
        if world > 1:
            dist.all_reduce(d_out)
    for _ in range(2):
        step_cost_only()
    barrier()
    ev0 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        flush.fill_(k & 0xFF)
        ev0[k][0].record(stream)
        step_cost_only()
        ev0[k][1].record(stream)
    barrier()
    cost_s = torch.tensor([sum(a.elapsed_time(b) for a, b in ev0) * 1e-3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(cost_s, op=dist.ReduceOp.MAX)
    cost_only_value = args.steps * n_total / float(cost_s.item())
    # leave the gradient evaluation's per-condition results in the handle (read below)
    step_device()
    barrier()

    # ---- end to end through the public API with host buffers (`e2e`) ----------------
    h_out = torch.empty(2 + n_opt, dtype=torch.float64).pin_memory()
    e2e_t = []
    for k in range(args.warmup + args.steps):
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            st, fval, grad, _ = fun.evaluate(theta, idx)  # pb200_evaluate: H2D theta/idx, kernels, D2H
        else:
            step_device()
            h_out.copy_(d_out, non_blocking=False)
            torch.cuda.synchronize()
            fval = float(h_out[0])
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if k >= args.warmup:
            e2e_t.append(float(dt.item()))
    e2e_value = args.steps * n_total / sum(e2e_t)

    if rank == 0:
        res = fun.condition_results()
        lo, hi = capi.partition_bounds(n_total, rank, world)
        ops_info = model_op_counts()
        from parpe_b200.models import get_spec
        nx_model = get_spec(MODEL).nx
        flops, mean_stats = algorithmic_flops(res["stats"], ops_info, nx_model, tab.np_)
        kern_avg = float(np.mean(kernel_s)) if kernel_s else total_s / args.steps
        peak_tf = capi.measure_fp64_peak(local_rank)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        if MODEL in LARGE_MODELS:
            abytes = algorithmic_bytes_large(res["stats"], nx_model, tab.np_, ops_info["nnzJ"], tab, hi - lo)
        else:
            abytes = algorithmic_bytes(tab, hi - lo)
        roofline = {
            "bound": "fp64", "achieved": flops / kern_avg / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": flops / kern_avg / 1e12 / peak_tf,
            # DRAM bytes per launch: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full`
            # capture of this kernel (profiles/r01_g_summary.txt: 5 435 648 B for 4736 conditions), scaled
            # to the conditions of this launch; the large-model workloads have no full-length capture
            "traffic": (5435648.0 / 4736.0) * (hi - lo) if MODEL == "boehm" else None,
            "traffic_source": "profiles/r01_g_summary.txt (ncu, 4736 conditions), scaled per condition" if MODEL == "boehm" else None,
            "peak_source": "DFMA microbenchmark measured live on this device (MEASURED_PEAKS.json has no FP64 figure)",
            "kernel": ("simulate_team_kernel<%d> (%s)" % (fun.last_launch()["block"], MODEL)) if MODEL in LARGE_MODELS
            else "simulate_kernel<Model_%s,G=%d>" % (MODEL, fun.last_launch()["lanes"]),
            "kernel_ms": 1e3 * kern_avg, "algorithmic_flops_per_launch": flops,
            "hbm": {"achieved": abytes / kern_avg / 1e9, "peak": hbm_peak, "unit": "GB/s",
                    "frac": abytes / kern_avg / 1e9 / hbm_peak,
                    "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
            "mean_stats_per_solve": mean_stats,
        }
        nerr = float(out_host[1])
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(world), "conditions_total": n_total, "rtol": tab.rtol,
                       "atol": tab.atol, "l2": "flushed between timed steps (256 MiB fill, untimed)",
                       "timing": "CUDA events on the launching stream per step, max over ranks",
                       "order": "positions handed out by the previous evaluation's step counts, costliest first "
                                "(PB200_REORDER=0: index order); results do not depend on it",
                       "lanes_per_condition": fun.last_launch()["lanes"], "grid": fun.last_launch()["grid"],
                       "block": fun.last_launch()["block"]},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 48 * n_opt,
                    "d2h_bytes_per_step": 8 * (2 + n_opt),
                    "note": "pb200_evaluate with host theta/idx/fval/gradient; the index list is re-uploaded "
                            "only when it changes (4 B/condition)"},
            "cost_only": {"value": cost_only_value, "unit": UNIT,
                          "note": "objective without gradient (no sensitivity columns), device-resident, same timing rules"},
            "gpu_launches": 2 * args.steps, "clocks": clocks, "roofline": roofline,
            "nllh": float(out_host[0]), "failed_conditions": nerr, "wall_s": wall,
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(tab, theta)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
