#!/usr/bin/env python3
"""Parity campaign on a GPU box: every compiled configuration x 4 seeds x {1, 10} SQP-RTI iterations x both solve kernels
against the CPU oracle (exit codes bit-exact, worst relative trajectory error).  ~22 000 problem-solves, about 10 s."""
import os
import sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))      # (test infrastructure: the only place besides smoke / bench that may use oracle/)
from oscar_mpc_planner_mr_modification_b200 import engine, synthetic
from oracle_binding import Oracle
tot = bad = 0
worst = 0.0
for cfg, pl, sets in (("c2_tmpc12", 9, 64), ("tmpc_shipped", 5, 96), ("c1_basic", 1, 400), ("c5_ccmpc", 1, 200), ("c6_goal_unicycle", 1, 400), ("c7_linearized", 1, 400)):
    eng = engine.Engine(cfg, 0, 4096); orc = Oracle(cfg)
    has_split = eng.set_kernel_mode(0)
    for seed in (101, 202, 303, 404):
        b = synthetic.make_batch(eng.parameter_map, eng.dims, sets, pl, seed=seed, gaussian=(cfg == "c5_ccmpc")) if cfg == "c5_ccmpc" else synthetic.make_batch(eng.parameter_map, eng.dims, sets, pl, seed=seed)
        for nit in (1, 10):
            ref = orc.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=nit)
            for mode in ((1, 2) if has_split else (1,)):
                eng.set_kernel_mode(mode)
                out = eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=nit)
                ok = ref["exit_code"] == 1
                mism = int((out["exit_code"] != ref["exit_code"]).sum())
                err = float(np.abs(out["xtraj"][ok] - ref["xtraj"][ok]).max() / max(1.0, np.abs(ref["xtraj"][ok]).max())) if ok.any() else 0.0
                tot += b["n"]; bad += mism; worst = max(worst, err)
                if mism or err > 1e-7:
                    print(cfg, "seed", seed, "iters", nit, "mode", mode, "exit mismatches", mism, "max rel err %.2e" % err, flush=True)
    eng.close()
print("problems x kernels checked:", tot, "exit-code mismatches:", bad, "worst relative trajectory error: %.2e" % worst)
