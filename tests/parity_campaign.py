#!/usr/bin/env python3
"""Parity campaign on a GPU box: every compiled configuration x seeds x {1, 10} SQP-RTI iterations x both solve kernels
against the CPU oracle (exit codes bit-exact, worst relative trajectory error).  Run as a script for the full campaign
(4 seeds, ~30 000 problem-solves); tests/test_parity_campaign.py runs a 2-seed cut as a `-m gpu` test so that the driver
sees the verdict."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))      # (test infrastructure: the only place besides smoke / bench that may use oracle/)
from oscar_mpc_planner_mr_modification_b200 import engine, synthetic  # noqa: E402
from oracle_binding import Oracle  # noqa: E402

CASES = (("c2_tmpc12", 9, 64), ("tmpc_shipped", 5, 96), ("c1_basic", 1, 400), ("c5_ccmpc", 1, 200), ("c6_goal_unicycle", 1, 400),
         ("c7_linearized", 1, 400))


def campaign(seeds=(101, 202, 303, 404), scale=1.0, verbose=True):
    """returns dict(problems, exit_mismatches, worst_rel_err, per_config)"""
    tot = bad = 0
    worst = 0.0
    per = {}
    for cfg, pl, sets in CASES:
        sets = max(1, int(sets * scale))
        eng = engine.Engine(cfg, 0, 4096); orc = Oracle(cfg)
        has_split = eng.set_kernel_mode(0)
        c_tot = c_bad = 0
        c_worst = 0.0
        for seed in seeds:
            b = synthetic.make_batch(eng.parameter_map, eng.dims, sets, pl, seed=seed)
            for nit in (1, 10):
                ref = orc.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=nit)
                for mode in ((1, 2) if has_split else (1,)):
                    eng.set_kernel_mode(mode)
                    out = eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=nit)
                    ok = ref["exit_code"] == 1
                    mism = int((out["exit_code"] != ref["exit_code"]).sum())
                    err = float(np.abs(out["xtraj"][ok] - ref["xtraj"][ok]).max() / max(1.0, np.abs(ref["xtraj"][ok]).max())) if ok.any() else 0.0
                    c_tot += b["n"]; c_bad += mism; c_worst = max(c_worst, err)
                    if verbose and (mism or err > 1e-7):
                        print(cfg, "seed", seed, "iters", nit, "mode", mode, "exit mismatches", mism, "max rel err %.2e" % err, flush=True)
        eng.close()
        per[cfg] = dict(problems=c_tot, exit_mismatches=c_bad, worst_rel_err=c_worst, kernels=2 if has_split else 1)
        tot += c_tot; bad += c_bad; worst = max(worst, c_worst)
    return dict(problems=tot, exit_mismatches=bad, worst_rel_err=worst, per_config=per)


if __name__ == "__main__":
    r = campaign()
    for cfg, v in r["per_config"].items():
        print("%-18s problems x kernels %6d  exit-code mismatches %d  worst relative trajectory error %.2e" % (cfg, v["problems"], v["exit_mismatches"], v["worst_rel_err"]))
    print("problems x kernels checked:", r["problems"], "exit-code mismatches:", r["exit_mismatches"], "worst relative trajectory error: %.2e" % r["worst_rel_err"])
