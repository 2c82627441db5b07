#!/usr/bin/env python
"""bench.py -- forward-sensitivity ODE solves/sec per objective+gradient evaluation.

A "step" is one pass of the hot path -- AmiciSummedGradientFunction::evaluate(parameters,
dataIndices, fval, gradient) -- over one batch of synthetic conditions.

  python bench.py --gpus N --steps K --warmup W            (our arm, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  (reference CPU path, rank 0)

Default workload (the driver's line): the largest single-GPU configuration that runs in minutes,
BASELINE.json configs[3] -- Chen_MSB2009-shape generated mass-action network (nx = 500, nplist = 50,
nt = 8 output times to t = 300, stiff), forward sensitivities w.r.t. all 50 parameters, rtol 1e-8 /
atol 1e-12, random parameters in bounds, 148 conditions per GPU (weak scaling: the global problem
has N x that many conditions, each rank simulates its contiguous share and ONE NCCL all-reduce --
issued by the library itself, pb200_nccl_init -- combines nllh and gradient).  `--workload
froehlich` is BASELINE configs[4] (nx = 1200, nplist = 100), `--workload boehm` configs[1] (nx = 8,
the small-model kernel; its numbers also ride in the default line under "boehm").  `--scaling
strong` fixes the TOTAL number of conditions (default 10^4 for boehm, 1184 for the large models)
and divides it over the ranks.  The real PEtab problems are not vendored in the reference tree:
model structures are stand-ins of the published shapes and the measurement tables are synthetic
(say so in `data`).
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
DEFAULT_WORKLOAD = "chen"
# conditions per GPU (weak scaling).  Large models: one condition per SM (one resident wave of the
# team kernel on a B200).  boehm: four resident waves of the small-model kernel (148 SMs x 32 conditions);
# any number works -- the kernels pull conditions from a queue -- but a partial last wave costs a whole one.
DEFAULT_CONDITIONS = {"chen": 148, "froehlich": 148, "net32": 4736, "boehm": 4 * 148 * 32}
STRONG_TOTAL = {"boehm": 10000, "chen": 1184, "froehlich": 1184}
LARGE_MODELS = ("chen", "froehlich", "net32", "net31")
CPU_SAMPLE_SECONDS = 15.0
THETA_WIDTH = {"boehm": 0.5}  # (round-1 continuity; the heterogeneous Boehm set uses SURVEY 8(d)'s +-1)


def host_threads():
    """Threads the reference arm uses: all host cores this process may run on (torchrun exports
    OMP_NUM_THREADS=1, so the count is passed explicitly, never taken from the OpenMP default)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def workload_name(model, per_gpu, n_gpus, total, hetero=False):
    name = _workload_name(model, per_gpu, n_gpus, total, hetero)
    if STEADY:
        name += "; steady-state variant: 3 of 4 conditions pre-equilibrated, last two output times t = inf"
    return name


def _workload_name(model, per_gpu, n_gpus, total, hetero=False):
    if model in LARGE_MODELS:
        from parpe_b200.models import get_spec
        sp = get_spec(model)
        shape = {"chen": "Chen_MSB2009-shape", "froehlich": "Froehlich_CellSystems2018-shape"}.get(model, model)
        return ("%s generated mass-action network (nx=%d, np=%d, ny=%d, nt=8), forward sensitivities, obj+grad, "
                "%d conditions total on %d GPU" % (shape, sp.nx, sp.np, sp.ny, total, n_gpus))
    if model == "boehm":
        return ("Boehm_JProteomeRes2014-shape (nx=8, np=9, ny=3, nt=16)%s, forward sensitivities, obj+grad, "
                "%d conditions total on %d GPU" % (", heterogeneous conditions (constant spans 300x, theta +-1)" if hetero else "",
                                                   total, n_gpus))
    return "%s, %d conditions total on %d GPU" % (model, total, n_gpus)


STEADY = False  # --steady: three of four conditions pre-equilibrated, the last two output times t = inf


def make_tables(model, n_conditions, hetero=False):
    from parpe_b200 import synth
    kw = dict(preeq=True, posteq=2) if STEADY else {}
    tab = synth.make_problem(model, n_conditions, scale="log10", seed=0, heterogeneous=hetero, **kw)
    theta = synth.random_theta(tab, seed=1, width=1.0 if hetero else THETA_WIDTH.get(model, 0.5))
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


def _stat_dict(stats):
    from parpe_b200.tables import STAT_NAMES
    return {n: stats[:, i].astype(np.float64) for i, n in enumerate(STAT_NAMES)}


def algorithmic_flops(stats, ops_info, nx, nplist, linsol=None):
    """Flops of one launch from the integrator statistics -- the counting model of SURVEY.md
    section 8(d): F = nfe*F_f + nje*F_J + nsetups*(2/3 nx^3 + nx^2) + nni*2nx^2 +
    nniS*nplist*2nx^2 + nfSe*(2 nnzJ + F_dfdp/nplist) + V, with the vector work
    V = nst*(q+1)(q+2)*nx*(1+nplist) + nst*2(q+1)*nx*(1+nplist) + (nni + nniS*nplist)*6nx.
    Sparse-LU variant (8(d): "replace 2/3 nx^3, 2nx^2 by the factor/solve flop counts of the fixed
    symbolic pattern"): linsol = dict(factor_flops, solve_flops, m_dense) of the symbolic phase --
    factor = sparse part + LU of the dense m x m Schur block, solve = sparse triangular solves + 2m^2."""
    s = _stat_dict(stats)
    q = np.where(s["nst"] > 0, s["qsum"] / np.maximum(s["nst"], 1), 1.0)
    nc = 1 + nplist
    if linsol:
        m = float(linsol["m_dense"])
        f_factor = linsol["factor_flops"] + 2.0 / 3.0 * m ** 3 + m ** 2
        f_solve = linsol["solve_flops"] + 2.0 * m ** 2
    else:
        f_factor = 2.0 / 3.0 * nx ** 3 + nx ** 2
        f_solve = 2.0 * nx ** 2
    F = (s["nfe"] * ops_info["xdot"] + s["nje"] * ops_info["J"] + s["nsetups"] * f_factor
         + s["nni"] * f_solve + s["nniS"] * nplist * f_solve
         + s["nfSe"] * (2 * ops_info["nnzJ"] + ops_info["dxdotdp"] / max(nplist, 1)))
    V = (s["nst"] * (q + 1) * (q + 2) * nx * nc + s["nst"] * 2 * (q + 1) * nx * nc
         + (s["nni"] + s["nniS"] * nplist) * 6 * nx)
    return float(np.sum(F + V)), {k: float(np.mean(v)) for k, v in s.items()}


def algorithmic_bytes(tab, n, want_grad=True):
    """Small-nx regime: inputs + outputs only:
    8*(np + nk + nt + 2*nt*ny + 1 + nplist) per solve."""
    nt = tab.nt(0)
    return float(n) * 8.0 * (tab.np_ + tab.nk + nt + 2 * nt * tab.ny + 1 + (tab.np_ if want_grad else 0))


def model_op_counts(model):
    from parpe_b200.codegen.modelgen import count_flops, derive
    from parpe_b200.models import get_spec
    if model in LARGE_MODELS:
        return get_spec(model).flops()
    return count_flops(derive(get_spec(model)))


def algorithmic_bytes_large(stats, nx, nplist, nnz, tab, n, linsol=None):
    """Large-nx regime (state not cache-resident), SURVEY.md 8(d):
    B = 8*[nst*(1+nplist)*nx*(2(q+1) + 2(q+1) + 6) + (nni + nniS)*(nx^2 or nnzLU) + nsetups*2*(nx^2 or nnzLU)
        + nje*nnzJ] + inputs + outputs;  nnzLU of the sparse variant = sparse factor values + m^2."""
    s = _stat_dict(stats)
    q = np.where(s["nst"] > 0, s["qsum"] / np.maximum(s["nst"], 1), 1.0)
    lu = (linsol["solve_bytes"] / 8.0) if linsol else float(nx) ** 2
    B = 8.0 * (s["nst"] * (1 + nplist) * nx * (4 * (q + 1) + 6) + (s["nni"] + s["nniS"]) * lu
               + s["nsetups"] * 2 * lu + s["nje"] * nnz)
    return float(np.sum(B)) + algorithmic_bytes(tab, n)


def cpu_baseline(model, tab, theta, seconds=CPU_SAMPLE_SECONDS):
    """Reference AMICI/CVODES path (oracle/_ref, KLU, staggered) + restated parPE glue, OpenMP over
    conditions on this box's host cores, on a bounded sample of the same workload."""
    from oracle import binding as ob
    if not ob.available():
        return None
    nthreads = host_threads()
    o = ob.OracleProblem(tab, ob.ISM_STAGGERED, ob.LINSOL_KLU)
    n0 = min(tab.n_conditions, (1 if model in LARGE_MODELS else 8) * nthreads)
    t0 = time.perf_counter()
    r = o.evaluate(theta, idx=np.arange(n0, dtype=np.int32), nthreads=nthreads)
    dt = time.perf_counter() - t0
    n = n0
    if dt < 0.5 * seconds:  # the probe was short: a second, bounded sample of about `seconds`
        n = int(min(tab.n_conditions, max(n0, (n0 / dt) * seconds)))
        t0 = time.perf_counter()
        r = o.evaluate(theta, idx=np.arange(n, dtype=np.int32), nthreads=nthreads)
        dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": nthreads, "kind": "reference",
            "sample": "%d of %d conditions, one obj+grad evaluation, %.1f s, %d OpenMP threads; AMICI 0.31.0/CVODES 7.1.1 "
                      "(KLU, staggered) built from the reference sources, dynamic schedule over conditions" % (
                          n, tab.n_conditions, dt, nthreads),
            "status": int(r["status"])}


def run_reference(args, model, per_gpu):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    n_gpus = args.gpus
    total = per_gpu * n_gpus if args.scaling == "weak" else args.conditions
    tab, theta = make_tables(model, total)
    from oracle import binding as ob
    nthreads = host_threads()
    o = ob.OracleProblem(tab, ob.ISM_STAGGERED, ob.LINSOL_KLU)
    # bounded sample per step: sized from a probe so that the whole run ends within minutes; never
    # fewer conditions than threads (every core busy)
    n0 = min(tab.n_conditions, (1 if model in LARGE_MODELS else 8) * nthreads)
    t0 = time.perf_counter()
    o.evaluate(theta, idx=np.arange(n0, dtype=np.int32), nthreads=nthreads)
    probe_s = time.perf_counter() - t0
    rate = n0 / probe_s
    budget = 150.0 / max(args.steps + min(args.warmup, 1), 1)
    n = int(min(tab.n_conditions, max(n0, rate * min(budget, 20.0))))
    idx = np.arange(n, dtype=np.int32)
    step_est = probe_s * n / n0
    warm = args.warmup if step_est * args.warmup < 30.0 else min(args.warmup, 1)  # (nothing to warm on a CPU)
    # One step cannot be shorter than one solve on one core (Chen shape: ~17 s), so for slow workloads K steps of
    # even the smallest full-machine sample (one condition per thread) would run for many minutes: the rate is then
    # measured over fewer timed steps (at least 2), and the line says so.
    k_timed = args.steps
    if step_est * args.steps > 200.0:
        k_timed = max(2, min(args.steps, int(200.0 / max(step_est, 1e-9))))
    for _ in range(warm):
        o.evaluate(theta, idx=idx, nthreads=nthreads)
    t0 = time.perf_counter()
    for _ in range(k_timed):
        r = o.evaluate(theta, idx=idx, nthreads=nthreads)
    dt = time.perf_counter() - t0
    value = k_timed * n / dt
    sample = ("%d of %d conditions per step (bounded sample of the same workload), %d OpenMP threads (passed "
              "explicitly); AMICI 0.31.0 / CVODES 7.1.1 (KLU, staggered) compiled from the reference sources + "
              "restated parPE glue, dynamic schedule over conditions" % (n, tab.n_conditions, nthreads))
    if k_timed != args.steps:
        sample += "; %d of the %d steps timed (a step is >= one solve on one core: %.0f s)" % (k_timed, args.steps, step_est)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / k_timed,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": workload_name(model, per_gpu, n_gpus, total), "sample": sample,
                                            "warmup_steps_run": warm, "steps_timed": k_timed},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthreads, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "status": int(r["status"])}
    print(json.dumps(line))
    return 0


class Rig:
    """One rank's view of the job: device, NCCL plumbing, timing helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from parpe_b200 import capi
        self.torch, self.dist, self.capi = torch, dist, capi
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.rank = int(os.environ.get("RANK", 0))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        if self.world != args.gpus and self.world > 1:
            raise SystemExit("WORLD_SIZE %d != --gpus %d" % (self.world, args.gpus))
        if capi.device_count() <= 0:
            raise SystemExit("no CUDA device: the hot path has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self.stream = torch.cuda.current_stream()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def make_function(self, tab, lanes=0, linsol=0):
        """Handle on this rank's GPU; with several ranks the library's own NCCL communicator (one
        all-reduce of 2 + n_opt doubles inside every evaluation)."""
        from parpe_b200.problem import B200SummedGradientFunction
        tab.lanes_per_condition = lanes
        fun = B200SummedGradientFunction(tab, device=self.local_rank)
        if linsol:
            fun.set_option(self.capi.OPT_LINSOL, linsol)
        if self.world > 1:
            uid = self.torch.zeros(128, dtype=self.torch.uint8, device=self.dev)
            if self.rank == 0:
                uid.copy_(self.torch.frombuffer(bytearray(B200SummedGradientFunction.nccl_unique_id()),
                                                dtype=self.torch.uint8))
            self.dist.broadcast(uid, 0)
            fun.nccl_init(bytes(uid.cpu().numpy().tobytes()), self.rank, self.world)
        return fun


def measure(rig, model, n_total, steps, warmup, lanes=0, hetero=False, linsol=0, want_cost_only=True, sample_clocks=True):
    """Device-resident throughput, objective-only rate and end-to-end rate of one workload."""
    torch = rig.torch
    tab, theta = make_tables(model, n_total, hetero)
    fun = rig.make_function(tab, lanes, linsol)
    idx = np.arange(n_total, dtype=np.int32)
    n_opt = tab.n_opt
    d_out = torch.zeros(2 + n_opt, dtype=torch.float64, device=rig.dev)
    stream = rig.stream

    def step_device(want_grad=True):
        # simulate + reduce (+ the library's NCCL all-reduce when partitioned), all on `stream`
        fun.evaluate_device(theta, idx, want_grad, d_out.data_ptr(), stream.cuda_stream)

    # one evaluation through the host API first: it leaves the cost-aware processing order (previous
    # evaluation's step counts) in the handle, which the device-resident path then uses as the
    # host path does from its second evaluation on
    fun.evaluate(theta, idx)
    for _ in range(warmup):
        step_device()
    rig.barrier()
    sampler = ClockSampler(rig.local_rank)
    if rig.rank == 0 and sample_clocks:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    kernel_s = []
    wall0 = time.perf_counter()
    for k in range(steps):
        rig.flush.fill_(k & 0xFF)  # L2 flush between timed iterations (untimed)
        ev[k][0].record(stream)
        step_device()
        ev[k][1].record(stream)
        if rig.world == 1:
            torch.cuda.synchronize()
            kernel_s.append(fun.last_launch()["seconds"])
    rig.barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if (rig.rank == 0 and sample_clocks) else None
    total_s = rig.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev) * 1e-3)
    out_host = d_out.cpu().numpy().copy()
    value = steps * n_total / total_s

    cost_only_value = None
    if want_cost_only:
        # objective only (gradient.empty() in the reference): same workload, no sensitivities
        for _ in range(2):
            step_device(False)
        rig.barrier()
        ev0 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for k in range(steps):
            rig.flush.fill_(k & 0xFF)
            ev0[k][0].record(stream)
            step_device(False)
            ev0[k][1].record(stream)
        rig.barrier()
        cost_s = rig.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev0) * 1e-3)
        cost_only_value = steps * n_total / cost_s
        # leave the gradient evaluation's per-condition results in the handle (read below)
        step_device()
        rig.barrier()

    # ---- end to end through the public API with host buffers (`e2e`): pb200_evaluate on every rank --
    # H2D of the parameter table (+ index list when it changed), kernels, the library's all-reduce, D2H
    e2e_t, h2d, d2h = [], [], []
    for k in range(warmup + steps):
        rig.barrier()
        t0 = time.perf_counter()
        st, fval, grad, _ = fun.evaluate(theta, idx)
        dt = rig.max_over_ranks(time.perf_counter() - t0)
        if k >= warmup:
            e2e_t.append(dt)
            a, b = fun.last_transfer_bytes()
            h2d.append(a)
            d2h.append(b)
    e2e_value = steps * n_total / sum(e2e_t)
    res = fun.condition_results() if rig.rank == 0 else None
    return dict(tab=tab, theta=theta, fun=fun, value=value, total_s=total_s, wall=wall, clocks=clocks,
                kernel_s=kernel_s, out_host=out_host, cost_only=cost_only_value, e2e=e2e_value,
                e2e_status=int(st), h2d=float(np.mean(h2d)), d2h=float(np.mean(d2h)), res=res, n_total=n_total,
                steps=steps)


def roofline_of(rig, model, m, peaks, peak_tf):
    """Achieved fraction of the FP64 and HBM rooflines for the dominant kernel of measurement `m`
    (rank 0's launch): the larger fraction is the binding one."""
    from parpe_b200.models import get_spec
    tab, fun, res = m["tab"], m["fun"], m["res"]
    lo, hi = rig.capi.partition_bounds(m["n_total"], rig.rank, rig.world)
    ops_info = model_op_counts(model)
    nx_model = get_spec(model).nx
    linsol = None
    if model in LARGE_MODELS:
        info = fun.linsol_info()
        if info.get("kind") == 2 and m.get("linsol", 0) != 1:
            linsol = info
    flops, mean_stats = algorithmic_flops(res["stats"], ops_info, nx_model, tab.np_, linsol)
    kern_avg = float(np.mean(m["kernel_s"])) if m["kernel_s"] else m["total_s"] / m["steps"]
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    if model in LARGE_MODELS:
        abytes = algorithmic_bytes_large(res["stats"], nx_model, tab.np_, ops_info["nnzJ"], tab, hi - lo, linsol)
    else:
        abytes = algorithmic_bytes(tab, hi - lo)
    f_frac = flops / kern_avg / 1e12 / peak_tf
    b_frac = abytes / kern_avg / 1e9 / hbm_peak
    ll = fun.last_launch()
    fp64 = {"achieved": flops / kern_avg / 1e12, "peak": peak_tf, "unit": "TFLOP/s", "frac": f_frac,
            "peak_source": "DFMA microbenchmark (pb200_measure_fp64_peak) measured live on this device; "
                           "MEASURED_PEAKS.json has no FP64 figure"}
    hbm = {"achieved": abytes / kern_avg / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": b_frac,
           "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"}
    bound = "hbm" if b_frac >= f_frac else "fp64"
    top = dict(hbm if bound == "hbm" else fp64)
    traffic, traffic_source = None, None
    tr_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_file):
        try:
            tr = json.load(open(tr_file)).get(model)
            if tr:
                traffic = tr["dram_bytes_per_condition"] * (hi - lo)
                traffic_source = tr["source"]
        except Exception:
            pass
    rf = {"bound": bound, "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"], "frac": top["frac"],
          "traffic": traffic, "traffic_source": traffic_source, "peak_source": top["peak_source"],
          "kernel": ("simulate_team_kernel<%d> (%s)" % (ll["block"], model)) if model in LARGE_MODELS
          else "simulate_kernel<Model_%s,G=%d>" % (model, ll["lanes"]),
          "kernel_ms": 1e3 * kern_avg, "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": abytes,
          "fp64": fp64, "hbm": hbm, "mean_stats_per_solve": mean_stats,
          "counting": "SURVEY.md 8(d)" + (": sparse-LU variant (factor/solve counts of the fixed symbolic pattern + "
                                          "LU of the dense Schur block)" if linsol else ": dense")}
    if linsol:
        rf["linear_solver"] = {k: linsol[k] for k in ("nx", "n_sparse", "m_dense", "nnz_L", "nnz_U", "updates",
                                                     "levels_fwd", "levels_bwd")}
        st = res["stats"]
        rf["sparse_fallbacks_per_solve"] = float(np.mean(st[:, 16]))
    if model in LARGE_MODELS:
        st = res["stats"]
        cyc = {"dense_panel": (st[:, 14] >> 32), "dense_update": (st[:, 14] & 0xffffffff), "solve": (st[:, 15] >> 32),
               "rhs": (st[:, 15] & 0xffffffff), "sparse_factor": (st[:, 17] >> 32), "sparse_tri": (st[:, 17] & 0xffffffff)}
        rf["phase_kcycles_per_solve"] = {k: float(np.mean(v)) for k, v in cyc.items()}
    return rf


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lanes", type=int, default=0, help="small-model kernel: lanes per condition (0 = model default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra small-model (boehm) measurements")
    ap.add_argument("--no-cost-only", action="store_true", help="skip the objective-only (no gradient) measurement")
    ap.add_argument("--quick", action="store_true",
                    help="scaling records of workloads whose single evaluation takes tens of seconds (NOT the driver's "
                         "line): one host evaluation as warm-up, ONE timed step, one end-to-end evaluation; says so in config")
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD,
                    help="chen (default: BASELINE configs[3], nx = 500) | froehlich (configs[4], nx = 1200) | boehm "
                         "(configs[1], the small-model kernel) | net32 | fujita ...")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--conditions", type=int, default=0,
                    help="weak: conditions per GPU; strong: total conditions (0 = default of the workload)")
    ap.add_argument("--linsol", type=int, default=0, help="large models: 0 automatic, 1 dense inverse, 2 sparse LU")
    ap.add_argument("--heterogeneous", action="store_true", help="boehm: conditions that differ by up to 300x, theta +-1")
    ap.add_argument("--steady", action="store_true",
                    help="steady-state variant of the workload (DESIGN 3.5): three of four conditions pre-equilibrated, "
                         "the last two output times t = inf; the integrator statistics (and so the roofline figures) "
                         "count the main simulation only")
    args = ap.parse_args()
    global STEADY
    STEADY = bool(args.steady)
    model = args.workload
    if args.scaling == "weak":
        per_gpu = args.conditions or DEFAULT_CONDITIONS.get(model, 4 * 148 * 32)
    else:
        args.conditions = args.conditions or STRONG_TOTAL.get(model, 10000)
        per_gpu = args.conditions // max(args.gpus, 1)
    if args.impl == "reference":
        return run_reference(args, model, per_gpu)
    args.warmup = max(args.warmup, 3)
    if args.quick:
        args.warmup, args.steps, args.no_cost_only, args.no_extra, args.no_cpu_baseline = 0, 1, True, True, True

    rig = Rig(args)
    world, rank = rig.world, rig.rank
    n_total = per_gpu * world if args.scaling == "weak" else args.conditions
    m = measure(rig, model, n_total, args.steps, args.warmup, lanes=args.lanes, hetero=args.heterogeneous,
                linsol=args.linsol, want_cost_only=not args.no_cost_only)
    m["linsol"] = args.linsol

    extra = {}
    if not args.no_extra and model != "boehm":
        # the small-model kernel beside the default line (BASELINE configs[1], round 1's driver line), in its
        # near-identical and its heterogeneous form; short runs (5 steps)
        nb = DEFAULT_CONDITIONS["boehm"] * world
        for key, het in (("boehm", False), ("boehm_heterogeneous", True)):
            mb = measure(rig, "boehm", nb, 5, 3, hetero=het, want_cost_only=False, sample_clocks=False)
            if rank == 0:
                extra[key] = mb
        # ... and its steady-state variant (three of four conditions pre-equilibrated, the last two output
        # times t = inf -- the steady-state instantiation of the kernel; `--workload boehm --steady` as a line of its own)
        was = STEADY
        STEADY = True
        try:
            mb = measure(rig, "boehm", nb, 5, 3, hetero=False, want_cost_only=False, sample_clocks=False)
            if rank == 0:
                extra["boehm_steady"] = mb
        except Exception as e:  # (an extra key must not cost the driver its line)
            if world > 1:
                raise
            extra["boehm_steady_error"] = repr(e)
        finally:
            STEADY = was
    if rank == 0:
        peak_tf = rig.capi.measure_fp64_peak(rig.local_rank)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        roofline = roofline_of(rig, model, m, peaks, peak_tf)
        tab, fun = m["tab"], m["fun"]
        ll = fun.last_launch()
        line = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * m["total_s"] / args.steps, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(model, per_gpu, world, n_total, args.heterogeneous),
                       "conditions_total": n_total, "conditions_per_gpu": per_gpu, "rtol": tab.rtol,
                       "atol": tab.atol, "l2": "flushed between timed steps (256 MiB fill, untimed)",
                       "timing": "CUDA events on the launching stream per step, max over ranks",
                       "order": "positions handed out by an earlier evaluation's step counts, costliest first "
                                "(PB200_OPT_REORDER); results do not depend on it",
                       "collective": ("one ncclAllReduce of %d doubles per evaluation, issued by libparpe_b200 "
                                      "(pb200_nccl_init) on the evaluation's stream" % (2 + tab.n_opt)) if world > 1 else "none (1 GPU)",
                       "lanes_per_condition": ll["lanes"], "grid": ll["grid"], "block": ll["block"],
                       **({"quick": "one host evaluation as the only warm-up, one timed step, one e2e evaluation "
                                    "(scaling record of a long-running workload, not a driver line)"} if args.quick else {})},
            "e2e": {"value": m["e2e"], "unit": UNIT, "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                    "note": "pb200_evaluate with host theta/idx/fval/gradient on every rank; bytes as counted by the "
                            "library (pb200_last_transfer_bytes): parameter table 48 B/parameter, index list 4 B/condition "
                            "when it changed, result 8(2+n_opt) B, and every 16th evaluation the processing-order refresh "
                            "(8 B/condition down, 4 B/condition up)"},
            "cost_only": {"value": m["cost_only"], "unit": UNIT,
                          "note": "objective without gradient (no sensitivity columns), device-resident, same timing rules"},
            "gpu_launches": 2 * args.steps, "clocks": m["clocks"], "roofline": roofline,
            "nllh": float(m["out_host"][0]), "failed_conditions": float(m["out_host"][1]), "wall_s": m["wall"],
        }
        for key, mb in extra.items():
            if key.endswith("_error"):
                line[key] = mb
                continue
            rb = roofline_of(rig, "boehm", mb, peaks, peak_tf)
            line[key] = {"value": mb["value"], "unit": UNIT, "e2e": mb["e2e"], "steps": 5,
                         "workload": workload_name("boehm", DEFAULT_CONDITIONS["boehm"], world, mb["n_total"],
                                                   key.endswith("heterogeneous")) +
                                     ("; steady-state variant: 3 of 4 conditions pre-equilibrated, last two output times "
                                      "t = inf (the integrator statistics count the main simulation only)"
                                      if key.endswith("steady") else ""),
                         "roofline_frac_fp64": rb["fp64"]["frac"], "kernel_ms": rb["kernel_ms"],
                         "mean_nst": rb["mean_stats_per_solve"]["nst"], "failed_conditions": float(mb["out_host"][1])}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(model, tab, m["theta"])
        print(json.dumps(line))
    if world > 1:
        rig.dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
