"""The oracle's QP and SQP-RTI levels: independent KKT verification of the interior-point solution
(numpy), frozen solve goldens, NLP-level sanity of converged problems, persistent-memory and
selection semantics."""
import ctypes
import os

import numpy as np
import pytest

from oracle_binding import Oracle
from oscar_mpc_planner_mr_modification_b200 import synthetic

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
PLANNERS = {"c1_basic": 1, "tmpc_shipped": 5, "c2_tmpc12": 9, "c5_ccmpc": 1, "c6_goal_unicycle": 1, "c7_linearized": 1}


def qp_debug(orc, xinit, x0, params):
    N, nx, nz, nh, nc = orc.N, orc.nx, orc.nz, orc.nh, orc.nc
    H = np.zeros((N + 1, nz, nz)); g = np.zeros((N + 1, nz)); W = np.zeros((N, nx, nz)); b = np.zeros((N, nx))
    C = np.zeros((N, max(nh, 1), nz)); d = np.zeros((N, nc)); v = np.zeros((N + 1, nz)); pi = np.zeros((N + 1, nx))
    lam = np.zeros((N, nc)); t = np.zeros((N, nc)); it = ctypes.c_int()
    st = orc.lib.oracle_qp_debug(P(xinit), P(x0), P(params), P(H), P(g), P(W), P(b), P(C), P(d), P(v), P(pi), P(lam), P(t), ctypes.byref(it))
    return st, it.value, dict(H=H, g=g, W=W, b=b, C=C, d=d, v=v, pi=pi, lam=lam, t=t)


@pytest.mark.parametrize("cfg", ["c1_basic", "c2_tmpc12"])
def test_qp_solution_satisfies_kkt(cfg):
    """Stationarity, primal/dual feasibility and complementarity of the first QP, recomputed in numpy
    from the exported QP data: the IPM's answer is checked without trusting the IPM."""
    orc = Oracle(cfg)
    b = synthetic.make_batch(orc.parameter_map, orc.dims, 3, PLANNERS[cfg], seed=77)
    lh, uh = orc.bounds(2), orc.bounds(3)
    hrow, hs = [], []
    for r in range(orc.nh):
        if lh[r] > -1e10: hrow.append(r); hs.append(1.0)
        if uh[r] < 1e10: hrow.append(r); hs.append(-1.0)
    N, nu, nz, nc = orc.N, orc.nu, orc.nz, orc.nc
    checked = 0
    for i in range(b["n"]):
        st, iters, q = qp_debug(orc, b["xinit"][i], b["x0"][i], b["params"][i])
        if st != 0:
            continue       # infeasible QP (reported through the status): nothing to verify
        checked += 1
        assert iters <= 50
        for k in range(N + 1):
            if k == N:
                r = q["H"][k][nu:, nu:] @ q["v"][k][nu:] - q["pi"][k]
                assert np.abs(r).max() < 1e-5
                continue
            Chat = np.vstack([np.eye(nz), -np.eye(nz)] + [hs[j] * q["C"][k, hrow[j]][None] for j in range(len(hrow))])
            act = np.ones(nc, bool)
            if k == 0:
                act[nu:nz] = False; act[nz + nu:2 * nz] = False
            r = q["H"][k] @ q["v"][k] + q["g"][k] + q["W"][k].T @ q["pi"][k + 1] - Chat[act].T @ q["lam"][k][act]
            if k > 0:
                r[nu:] -= q["pi"][k]
            else:
                r[nu:] = 0
            assert np.abs(r).max() < 1e-5                                                        # stationarity
            assert np.abs(q["W"][k] @ q["v"][k] + q["b"][k] - q["v"][k + 1][nu:]).max() < 1e-5      # dynamics
            slack = Chat[act] @ q["v"][k] - q["d"][k][act]
            assert slack.min() > -1e-5                                                           # primal feasibility
            assert (q["lam"][k][act] >= 0).all()                                                 # dual feasibility
            assert np.abs(q["lam"][k][act] * slack).max() < 2e-5                                 # complementarity
            assert np.linalg.eigvalsh(q["H"][k]).min() > 0.99e-4                                 # MIRROR floor
    assert checked >= 2


@pytest.mark.parametrize("cfg", sorted(PLANNERS))
def test_frozen_solve_golden(cfg):
    gd = np.load(os.path.join(GOLD, "solve_%s.npz" % cfg))
    orc = Oracle(cfg)
    b = synthetic.make_batch(orc.parameter_map, orc.dims, int(gd["n_sets"]), int(gd["planners"]), seed=int(gd["seed"]))
    for nit in (1, 10):
        r = orc.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=nit)
        np.testing.assert_array_equal(r["exit_code"], gd["exit_code_it%d" % nit])
        ok = r["exit_code"] == 1
        np.testing.assert_array_equal(r["qp_status"][ok], gd["qp_status_it%d" % nit][ok])
        np.testing.assert_allclose(r["xtraj"][ok], gd["xtraj_it%d" % nit][ok], rtol=0, atol=1e-8)
        np.testing.assert_allclose(r["utraj"][ok], gd["utraj_it%d" % nit][ok], rtol=0, atol=1e-8)
        np.testing.assert_allclose(r["pobj"][ok], gd["pobj_it%d" % nit][ok], rtol=1e-9)
        best = orc.select_best(b["set_offsets"], r["pobj"], r["exit_code"])
        np.testing.assert_array_equal(best, gd["best_it%d" % nit])


@pytest.mark.parametrize("cfg", ["c1_basic", "c2_tmpc12"])
def test_converged_solutions_are_feasible(cfg):
    """After 10 SQP-RTI iterations the successful problems satisfy dynamics, input/state bounds and the
    path constraints (stages 1..N-1) -- NLP-level sanity that is independent of the solver internals."""
    from test_oracle_model import model_eval
    orc = Oracle(cfg)
    b = synthetic.make_batch(orc.parameter_map, orc.dims, 4, PLANNERS[cfg], seed=5)
    r = orc.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=10)
    ok = np.nonzero(r["exit_code"] == 1)[0]
    assert len(ok) >= b["n"] // 2
    lb, ub, lh, uh = orc.bounds(0), orc.bounds(1), orc.bounds(2), orc.bounds(3)
    N, nx, nu, nz = orc.N, orc.nx, orc.nu, orc.nz
    for i in ok:
        x = r["xtraj"][i].reshape(N + 1, nx); u = r["utraj"][i].reshape(N, nu)
        assert r["res_eq"][i] < 1e-2
        np.testing.assert_allclose(x[0], b["xinit"][i], atol=1e-9)
        assert (u >= lb[:nu] - 1e-6).all() and (u <= ub[:nu] + 1e-6).all()
        assert (x[1:N] >= lb[nu:] - 1e-6).all() and (x[1:N] <= ub[nu:] + 1e-6).all()
        pp = b["params"][i].reshape(N, orc.npar)
        for k in range(1, N):
            h = model_eval(orc, np.concatenate([u[k], x[k]]), pp[k].copy(), np.zeros(nx), np.zeros(max(orc.nh, 1)))["h"]
            assert (h >= lh[:orc.nh] - 1e-3).all() and (h <= uh[:orc.nh] + 1e-3).all()


def test_persistent_memory_semantics():
    """mem blob: flag 2 + multipliers after success, zeroed after failure (Solver_acados_reset,
    acados_solver_interface.cpp:187-191); a second call that continues from the blob (exact Hessian with
    the stored multipliers, HPIPM warm start 2) is a different (multiplier-aware) step than a fresh capsule takes."""
    orc = Oracle("c1_basic")
    b = synthetic.make_batch(orc.parameter_map, orc.dims, 8, 1, seed=9)
    mem = np.zeros((b["n"], orc.mem_doubles))
    r1 = orc.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=3, mem=mem)
    ok = r1["exit_code"] == 1
    assert ok.any()
    assert (mem[ok, 0] == 2.0).all() and (np.abs(mem[ok, 1:]).sum(axis=1) > 0).all()
    assert (mem[~ok] == 0.0).all()
    x0b = np.zeros_like(b["x0"]).reshape(b["n"], orc.N + 1, orc.nz)
    x0b[:, :, orc.nu:] = r1["xtraj"].reshape(b["n"], orc.N + 1, orc.nx)
    x0b[:, :orc.N, :orc.nu] = r1["utraj"].reshape(b["n"], orc.N, orc.nu)
    x0b = x0b.reshape(b["n"], -1)
    r2 = orc.solve_batch(b["xinit"], x0b, b["params"], num_iter=1, mem=mem.copy())
    r2c = orc.solve_batch(b["xinit"], x0b, b["params"], num_iter=1, mem=None)
    assert (r2["exit_code"][ok] == 1).all()
    assert not np.array_equal(r2["xtraj"][ok], r2c["xtraj"][ok])     # the blob is really used


def test_select_best_semantics():
    """FindBestPlanner (guidance_constraints.cpp:572-590): strict '<' from 1e10 in ascending index order,
    success = exit_code == 1, disabled planners skipped, all-fail -> -1; objective post-processing :373-420."""
    orc = Oracle("c1_basic")
    off = np.array([0, 3, 6, 9, 12], np.int32)
    pobj = np.array([3.0, 2.0, 2.0, 1.0, 5.0, 0.5, 1.0, 1.0, 1.0, 4.0, 8.0, 2e10])
    ec = np.array([1, 1, 1, 4, 1, 0, 2, 3, 4, 1, 1, 1], np.int32)
    assert orc.select_best(off, pobj, ec).tolist() == [1, 1, -1, 0]       # tie -> first; failures skipped; all fail -> -1
    dis = np.zeros(12, np.uint8); dis[1] = 1; dis[9] = 1
    assert orc.select_best(off, pobj, ec, disabled=dis).tolist() == [2, 1, -1, 1]
    scale = np.ones(12); scale[0] = 0.5          # previously selected guidance x selection weight (:418-419)
    assert orc.select_best(off, pobj, ec, obj_scale=scale).tolist() == [0, 1, -1, 0]
    sub = np.zeros(12); sub[2] = 1.5             # consistency cost subtracted before the weight (:384-388)
    assert orc.select_best(off, pobj, ec, obj_scale=scale, obj_sub=sub).tolist() == [2, 1, -1, 0]
    assert orc.select_best(np.array([0, 1], np.int32), np.array([1e10]), np.array([1], np.int32)).tolist() == [-1]   # not < 1e10
