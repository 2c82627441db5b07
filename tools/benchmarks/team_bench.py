import sys, time, numpy as np
sys.path.insert(0, '.')
from parpe_b200 import synth
from parpe_b200.problem import B200SummedGradientFunction
model, C, ntp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
tab = synth.make_problem(model, C, seed=0, timepoints=synth.DEFAULT_TIMEPOINTS[model][:ntp])
theta = synth.random_theta(tab, seed=1, width=0.5)
gpu = B200SummedGradientFunction(tab)
for rep in range(2):
    t0 = time.time(); st, fval, grad, sec = gpu.evaluate(theta); wall = time.time() - t0
    res = gpu.condition_results()
    ms = res["stats"].mean(axis=0)
    print(model, "C", C, "status", st, "fval %.10g" % fval, "device_s %.3f wall %.3f" % (sec, wall), "solves/s %.1f" % (C / sec),
          "Mcyc panel %.0f update %.0f solve %.0f rhs %.0f total %.0f" % (
              (res["stats"][:,14] >> 32).mean()/1e3, (res["stats"][:,14] & 0xffffffff).mean()/1e3,
              (res["stats"][:,15] >> 32).mean()/1e3, (res["stats"][:,15] & 0xffffffff).mean()/1e3, sec*1.965e3),
          "nst %.0f nsetups %.0f nni %.0f nniS %.0f" % (ms[0], ms[2], ms[4], ms[5]), gpu.last_launch())
