import sys, time, json, numpy as np
sys.path.insert(0, '.')
from parpe_b200 import synth
from parpe_b200.problem import B200SummedGradientFunction
C = int(sys.argv[1]) if len(sys.argv) > 1 else 10002
tab = synth.fujita_hierarchical(C, seed=0)
gpu = B200SummedGradientFunction(tab)
n_outer = gpu.hier_num_outer()
full = synth.random_theta(tab, seed=1, width=0.3)
outer = np.delete(full, tab.hier.inner_idx)
assert len(outer) == n_outer
out = {}
for rep in range(4):
    t0 = time.perf_counter(); r = gpu.hier_evaluate(outer, want_gradient=True); wall = time.perf_counter() - t0
    out = dict(conditions=C, status=r["status"], fval=r["fval"], device_s=r["seconds"], wall_s=wall,
               solves_per_s_wall=2 * C / wall, evaluations_per_s=1.0 / wall)
    print(json.dumps(out))
t0 = time.perf_counter(); r0 = gpu.hier_evaluate(outer, want_gradient=False); wall0 = time.perf_counter() - t0
print(json.dumps(dict(cost_only=True, fval=r0["fval"], wall_s=wall0, solves_per_s_wall=C / wall0)))
from oracle import binding as ob
if ob.available():
    o = ob.OracleProblem(tab)
    t0 = time.perf_counter(); ro = o.hier_evaluate(outer, want_grad=True); wo = time.perf_counter() - t0
    g = r["grad"]; go = ro["grad"]; gs = np.max(np.abs(go))
    print(json.dumps(dict(oracle_wall_s=wo, oracle_solves_per_s=2 * C / wo, rel_fval=abs(r["fval"] - ro["fval"]) / abs(ro["fval"]),
                          rel_grad=float(np.max(np.abs(g - go) / np.maximum(np.abs(go), 1e-4 * gs))))))
