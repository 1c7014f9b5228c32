#!/usr/bin/env python3
"""Generate the golden fixtures (container only: imports the reference's Python problem definition).

  model_<cfg>.npz   values of the REFERENCE's own symbolic expressions (dynamics f, stage cost l,
                    constraints h, and their first derivatives) at seeded points, evaluated with
                    sympy.lambdify straight from the reference scripts -- independent of this repo's
                    code generators.  Pins oracle/generated/model_<cfg>.h and generated/<cfg>/model.cuh.
  solve_<cfg>.npz   frozen oracle outputs for seeded synthetic problems (regression pin of the oracle and
                    size-independent reference for the GPU tests).
"""
import os
import sys

import numpy as np
import sympy as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import reference_problem as rp  # noqa: E402
from oracle_binding import Oracle  # noqa: E402
from oscar_mpc_planner_mr_modification_b200 import synthetic  # noqa: E402

PLANNERS = {"c1_basic": 1, "tmpc_shipped": 5, "c2_tmpc12": 9, "c5_ccmpc": 1, "c6_goal_unicycle": 1, "c7_linearized": 1}


def model_golden(cfg, npts=12, seed=11):
    pb = rp.build(cfg)
    z, p = pb["z"], pb["p"]
    args = list(z) + list(p)
    mods = ["math", {"erf": __import__("math").erf, "_fmod": __import__("math").fmod}]
    f_fn = sp.lambdify(args, pb["f"], modules=mods)
    c_fn = sp.lambdify(args, pb["cost"], modules=mods)
    g_fn = sp.lambdify(args, [sp.diff(pb["cost"], v) for v in z], modules=mods)
    h_fn = sp.lambdify(args, pb["h"], modules=mods)
    jf_fn = sp.lambdify(args, [[sp.diff(e, v) for v in z] for e in pb["f"]], modules=mods)
    jh_fn = sp.lambdify(args, [[sp.diff(e, v) for v in z] for e in pb["h"]], modules=mods)
    orc = Oracle(cfg)
    b = synthetic.make_batch(orc.parameter_map, orc.dims, 2, PLANNERS[cfg], seed=seed)
    rng = np.random.default_rng(seed)
    N, nz, npar = orc.N, orc.nz, orc.npar
    Z, P, F, C, G, H, JF, JH = [], [], [], [], [], [], [], []
    for i in range(npts):
        prob = int(rng.integers(0, b["n"]))
        k = int(rng.integers(1, N))
        zz = b["x0"][prob].reshape(N + 1, nz)[k].copy()
        zz[:2] = rng.uniform(-0.5, 0.5, 2)
        zz[2:] += rng.normal(0, 0.1, nz - 2)
        pp = b["params"][prob].reshape(N, npar)[k].copy()
        if "ego_disc_0_offset" in orc.parameter_map and i % 2 == 1:
            pp[orc.parameter_map["ego_disc_0_offset"]] = 0.25      # exercise the psi-dependence of the disc position
        a = list(zz) + list(pp)
        Z.append(zz); P.append(pp); F.append(f_fn(*a)); C.append(c_fn(*a)); G.append(g_fn(*a)); H.append(h_fn(*a))
        JF.append(jf_fn(*a)); JH.append(jh_fn(*a))
    np.savez_compressed(os.path.join(HERE, "model_%s.npz" % cfg), z=np.array(Z), p=np.array(P), f=np.array(F, float),
                        cost=np.array(C, float), grad=np.array(G, float), h=np.array(H, float), jf=np.array(JF, float),
                        jh=np.array(JH, float), lh=np.array(pb["lh"]), uh=np.array(pb["uh"]), lb=np.array(pb["lb"]),
                        ub=np.array(pb["ub"]), param_names=np.array(pb["param_names"]))


def solve_golden(cfg, n_sets, seed=2024):
    orc = Oracle(cfg)
    b = synthetic.make_batch(orc.parameter_map, orc.dims, n_sets, PLANNERS[cfg], seed=seed)
    out = {}
    for nit in (1, 10):
        r = orc.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=nit)
        best = orc.select_best(b["set_offsets"], r["pobj"], r["exit_code"])
        for k, v in r.items():
            out["%s_it%d" % (k, nit)] = v
        out["best_it%d" % nit] = best
    np.savez_compressed(os.path.join(HERE, "solve_%s.npz" % cfg), seed=seed, n_sets=n_sets, planners=PLANNERS[cfg], **out)


if __name__ == "__main__":
    for cfg in PLANNERS:
        if "--solve-only" not in sys.argv:      # the model fixtures need /root/reference; the solve fixtures only the oracle
            model_golden(cfg)
        solve_golden(cfg, 8 if PLANNERS[cfg] > 1 else 32)
        print("golden", cfg)
