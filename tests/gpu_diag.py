"""Diagnostic (not a test): print where GPU and oracle differ for a config."""
import sys
import numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from oracle_binding import Oracle
from oscar_mpc_planner_mr_modification_b200 import engine, synthetic

cfg, planners, num_iter, n_sets = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
seed = int(sys.argv[5]) if len(sys.argv) > 5 else 1234
eng = engine.Engine(cfg, 0, max(64, n_sets * planners)); orc = Oracle(cfg)
b = synthetic.make_batch(eng.parameter_map, eng.dims, n_sets, planners, seed=seed)
out = eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=num_iter)
print("kernel ms", eng.last_kernel_ms(), "n", b["n"])
ref = orc.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=num_iter)
scale = np.maximum(1.0, np.abs(ref["xtraj"]).max(axis=1))
ex = np.abs(out["xtraj"] - ref["xtraj"]).max(axis=1) / scale
eu = np.abs(out["utraj"] - ref["utraj"]).max(axis=1)
for i in range(b["n"]):
    flag = ""
    if out["exit_code"][i] != ref["exit_code"][i] or out["qp_status"][i] != ref["qp_status"][i] or out["ipm_iters"][i] != ref["ipm_iters"][i] or (ref["exit_code"][i] == 1 and ex[i] > 1e-6):
        flag = "  <<<<"
        print(i, "exit", out["exit_code"][i], ref["exit_code"][i], "qps", out["qp_status"][i], ref["qp_status"][i], "ipm", out["ipm_iters"][i], ref["ipm_iters"][i],
              "ex %.2e eu %.2e" % (ex[i], eu[i]), "pobj %.9g %.9g" % (out["pobj"][i], ref["pobj"][i]), "req %.3g %.3g" % (out["res_eq"][i], ref["res_eq"][i]), flag)
ok = ref["exit_code"] == 1
print("ok frac", ok.mean(), "max ex over ok", ex[ok].max() if ok.any() else None, "mean ipm", ref["ipm_iters"].mean())
