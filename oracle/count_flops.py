#!/usr/bin/env python3
"""Canonical algorithmic FLOPs per solve: run the flop-counting oracle build on a
seeded sample of the bench workload and record the mean per config in oracle/flops.json.
TEST/MEASUREMENT INFRASTRUCTURE."""
import ctypes
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
sys.path.insert(0, os.path.join(HERE, "..", "tests"))
from oracle_binding import Oracle  # noqa: E402
from oscar_mpc_planner_mr_modification_b200 import synthetic  # noqa: E402

WORKLOADS = {"c1_basic": 1, "tmpc_shipped": 5, "c2_tmpc12": 9, "c5_ccmpc": 1, "c6_goal_unicycle": 1, "c7_linearized": 1}


def count(cfg, planners, num_iter, n_sets=16, seed=1234):
    orc = Oracle(cfg)
    lib = ctypes.CDLL(os.path.join(HERE, "_build", "libflops_%s.so" % cfg))
    b = synthetic.make_batch(orc.parameter_map, orc.dims, n_sets, planners, seed=seed)
    n = b["n"]
    fl = np.zeros(n); tr = np.zeros(n); ipm = np.zeros(n, np.int32); ec = np.zeros(n, np.int32)
    ni = np.full(n, num_iter, np.int32)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    lib.flopcount_solve(n, P(b["xinit"]), P(b["x0"]), P(b["params"]), P(ni), P(fl), P(tr), P(ipm), P(ec))
    # split into a per-IPM-iteration and a per-solve part by least squares (flops = a + b * ipm_iters)
    A = np.stack([np.ones(n), ipm.astype(float)], axis=1)
    coef, *_ = np.linalg.lstsq(A, fl, rcond=None)
    return dict(config=cfg, planners_per_set=planners, num_iter=num_iter, sample=n, seed=seed,
                flops_per_solve_mean=float(fl.mean()), transcendentals_per_solve_mean=float(tr.mean()),
                ipm_iters_mean=float(ipm.mean()), flops_fixed_part=float(coef[0]), flops_per_ipm_iter=float(coef[1]),
                convention="add/sub/mul/div/sqrt = 1 each (FMA = 2); transcendental calls counted separately; "
                           "dense counting by the plain CPU restatement")


if __name__ == "__main__":
    out = {}
    for cfg, pl in WORKLOADS.items():
        for nit in (1, 10):
            r = count(cfg, pl, nit)
            out["%s/iter%d" % (cfg, nit)] = r
            print(cfg, nit, "MFLOP/solve %.3f  transc %.0f  ipm %.1f  per-ipm-iter %.1f kFLOP" %
                  (r["flops_per_solve_mean"] / 1e6, r["transcendentals_per_solve_mean"], r["ipm_iters_mean"], r["flops_per_ipm_iter"] / 1e3))
    with open(os.path.join(HERE, "flops.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
