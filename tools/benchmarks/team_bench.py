"""Phase split of the large-model kernel (cycle counters of the statistics record).
usage: python tools/benchmarks/team_bench.py MODEL CONDITIONS N_OUTPUT_TIMES [LINSOL 0|1|2] [REPS] [FUSE_OFF bits]"""
import sys, time, numpy as np
sys.path.insert(0, '.')
from parpe_b200 import capi, synth
from parpe_b200.problem import B200SummedGradientFunction
model, C, ntp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
linsol = int(sys.argv[4]) if len(sys.argv) > 4 else 0
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
fuse_off = int(sys.argv[6]) if len(sys.argv) > 6 else 0
tab = synth.make_problem(model, C, seed=0, timepoints=synth.DEFAULT_TIMEPOINTS[model][:ntp])
theta = synth.random_theta(tab, seed=1, width=0.5)
gpu = B200SummedGradientFunction(tab)
gpu.set_option(capi.OPT_LINSOL, linsol)
gpu.set_option(capi.OPT_FUSE_OFF, fuse_off)
print("linsol", gpu.linsol_info())
for rep in range(reps):
    t0 = time.time(); st, fval, grad, sec = gpu.evaluate(theta); wall = time.time() - t0
    res = gpu.condition_results()
    S = res["stats"]
    ms = S.mean(axis=0)
    k = lambda v: float(np.mean(v)) / 1e3
    print(model, "C", C, "linsol", linsol, "fuse_off", fuse_off, "status", st, "fval %.10g" % fval, "device_s %.3f wall %.3f" % (sec, wall),
          "solves/s %.2f" % (C / sec),
          "Mcyc/solve: dense panel %.0f update %.0f | solve(all) %.0f of which sparse tri %.0f | rhs %.0f | sparse factor %.0f | total %.0f" % (
              k(S[:, 14] >> 32), k(S[:, 14] & 0xffffffff), k(S[:, 15] >> 32), k(S[:, 17] & 0xffffffff), k(S[:, 15] & 0xffffffff),
              k(S[:, 17] >> 32), sec * 1.965e3 * min(1.0, 148.0 / C)),
          "| history %.0f state-phase %.0f sens-phase %.0f whole %.0f |" % (k(S[:, 18] >> 32), k(S[:, 18] & 0xffffffff), k(S[:, 19] >> 32), k(S[:, 19] & 0xffffffff)),
          "nst %.0f nsetups %.0f nni %.0f nniS %.0f fallbacks %.2f" % (ms[0], ms[2], ms[4], ms[5], ms[16]), gpu.last_launch(), flush=True)
