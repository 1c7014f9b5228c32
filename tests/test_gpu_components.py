"""GPU component / edge-case / full-size tests, all through the C ABI."""
import os

import numpy as np
import pytest

from oracle_binding import Oracle
from oscar_mpc_planner_mr_modification_b200 import engine, synthetic

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PLANNERS = {"c1_basic": 1, "tmpc_shipped": 5, "c2_tmpc12": 9, "c5_ccmpc": 1, "c6_goal_unicycle": 1, "c7_linearized": 1}
REL_TOL = 1e-6


def unpack(pk, n=7):
    H = np.zeros((n, n))
    for i in range(n):
        for j in range(i + 1):
            H[i, j] = H[j, i] = pk[i * (i + 1) // 2 + j]
    return H


@pytest.mark.parametrize("cfg", sorted(PLANNERS))
def test_emitted_device_functions_match_reference_expressions(cfg):
    """generated/<cfg>/model.cuh evaluated ON THE GPU against the golden values of the reference's own
    symbolic expressions, and against the oracle for the second-order terms."""
    from test_oracle_model import model_eval
    gd = np.load(os.path.join(GOLD, "model_%s.npz" % cfg))
    eng = engine.Engine(cfg, 0, 64)
    orc = Oracle(cfg)
    n = gd["z"].shape[0]
    rng = np.random.default_rng(1)
    pi, mh = rng.normal(size=(n, eng.nx)), rng.normal(size=(n, eng.nh))
    r = eng.model_eval(gd["z"], gd["p"], pi, mh)
    nhs, dt = r["nhs"], 0.2
    sup = [i for i in range(eng.nz) if np.abs(gd["jh"][:, :, i]).max() > 0]
    assert len(sup) <= nhs
    np.testing.assert_allclose(r["cost"][:, 0], gd["cost"], rtol=1e-12)
    np.testing.assert_allclose(r["g"], dt * gd["grad"], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(r["h"], gd["h"], rtol=1e-12, atol=1e-12)
    C = r["C"].reshape(n, eng.nh, nhs)
    hsup = [2, 3, 4][:nhs]
    np.testing.assert_allclose(C, gd["jh"][:, :, hsup], rtol=1e-10, atol=1e-11)
    for i in range(n):
        z, p = gd["z"][i].copy(), gd["p"][i].copy()
        xn = np.zeros(eng.nx); W = np.zeros((eng.nx, eng.nz)); Hd = np.zeros((eng.nz, eng.nz))
        P = lambda a: a.ctypes.data_as(__import__("ctypes").c_void_p)
        orc.lib.oracle_integrate(P(z[eng.nu:].copy()), P(z[:eng.nu].copy()), P(p), P(pi[i].copy()), P(xn), P(W), P(Hd))
        np.testing.assert_allclose(r["xn"][i], xn, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(r["W"][i].reshape(eng.nx, eng.nz), W, rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(unpack(r["Hdyn"][i], eng.nz), Hd, rtol=1e-9, atol=1e-11)
        o = model_eval(orc, z, p, np.zeros(eng.nx), mh[i].copy())
        np.testing.assert_allclose(unpack(r["Hcost"][i], eng.nz), dt * o["Hc"], rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(unpack(r["Hcon"][i], eng.nz), o["Hh"], rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("cfg", sorted(PLANNERS))
def test_frozen_golden_solves(cfg):
    gd = np.load(os.path.join(GOLD, "solve_%s.npz" % cfg))
    eng = engine.Engine(cfg, 0, 512)
    b = synthetic.make_batch(eng.parameter_map, eng.dims, int(gd["n_sets"]), int(gd["planners"]), seed=int(gd["seed"]))
    for nit in (1, 10):
        r = eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=nit)
        np.testing.assert_array_equal(r["exit_code"], gd["exit_code_it%d" % nit])
        ok = r["exit_code"] == 1
        np.testing.assert_array_equal(r["qp_status"][ok], gd["qp_status_it%d" % nit][ok])
        if ok.any():       # (N = 50 with a single iteration: every problem is still above the res_eq threshold)
            scale = np.maximum(1.0, np.abs(gd["xtraj_it%d" % nit][ok]).max(axis=1, keepdims=True))
            assert (np.abs(r["xtraj"][ok] - gd["xtraj_it%d" % nit][ok]) / scale).max() < REL_TOL
            assert np.abs(r["utraj"][ok] - gd["utraj_it%d" % nit][ok]).max() < REL_TOL
        best = eng.select_best(b["set_offsets"], r["pobj"], r["exit_code"])
        np.testing.assert_array_equal(best, gd["best_it%d" % nit])


def test_select_best_bit_exact():
    eng = engine.Engine("c1_basic", 0, 64)
    orc = Oracle("c1_basic")
    off = np.array([0, 3, 6, 9, 12], np.int32)
    pobj = np.array([3.0, 2.0, 2.0, 1.0, 5.0, 0.5, 1.0, 1.0, 1.0, 4.0, 8.0, 2e10])
    ec = np.array([1, 1, 1, 4, 1, 0, 2, 3, 4, 1, 1, 1], np.int32)
    dis = np.zeros(12, np.uint8); dis[1] = 1; dis[9] = 1
    scale = np.ones(12); scale[0] = 0.5
    sub = np.zeros(12); sub[2] = 1.5
    for kw in ({}, dict(disabled=dis), dict(obj_scale=scale), dict(obj_scale=scale, obj_sub=sub), dict(obj_sub=sub, disabled=dis)):
        np.testing.assert_array_equal(eng.select_best(off, pobj, ec, **kw), orc.select_best(off, pobj, ec, **kw))
    assert eng.select_best(off, pobj, ec).tolist() == [1, 1, -1, 0]
    rng = np.random.default_rng(0)
    n_sets = 500
    sizes = rng.integers(0, 10, n_sets)                       # ragged sets, including empty ones
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    eng2 = engine.Engine("c1_basic", 0, int(off[-1]) + 8)
    pobj = np.round(rng.uniform(0, 3, off[-1]), 1)           # many exact ties
    ec = rng.choice([0, 1, 1, 1, 2, 4], off[-1]).astype(np.int32)
    dis = (rng.uniform(size=off[-1]) < 0.2).astype(np.uint8)
    np.testing.assert_array_equal(eng2.select_best(off, pobj, ec, disabled=dis), orc.select_best(off, pobj, ec, disabled=dis))


@pytest.mark.parametrize("n", [1, 3, 5, 33])
def test_ragged_batch_sizes_and_per_problem_iterations(n):
    cfg = "tmpc_shipped"
    eng = engine.Engine(cfg, 0, 64)
    orc = Oracle(cfg)
    b = synthetic.make_batch(eng.parameter_map, eng.dims, 7, 5, seed=3)
    sl = slice(0, n)
    ni = (np.arange(n) % 5).astype(np.int32) * 2            # 0, 2, 4, 6, 8 iterations: includes "no iteration"
    out = eng.solve_batch(b["xinit"][sl], b["x0"][sl], b["params"][sl], num_iter=ni)
    ref = orc.solve_batch(b["xinit"][sl], b["x0"][sl], b["params"][sl], num_iter=ni)
    np.testing.assert_array_equal(out["exit_code"], ref["exit_code"])
    ok = ref["exit_code"] == 1
    np.testing.assert_array_equal(out["qp_status"][ok], ref["qp_status"][ok])
    if ok.any():
        assert np.abs(out["xtraj"][ok] - ref["xtraj"][ok]).max() < 1e-6 * max(1.0, np.abs(ref["xtraj"][ok]).max())
    z = ni == 0                                             # zero iterations: the warm start comes back untouched
    x0 = b["x0"][sl].reshape(n, eng.N + 1, eng.nz)
    np.testing.assert_array_equal(out["xtraj"][z].reshape(-1, eng.N + 1, eng.nx), x0[z][:, :, eng.nu:])


def test_empty_batch_and_oversized_batch():
    eng = engine.Engine("c1_basic", 0, 8)
    assert eng.lib.mpcgpu_solve_batch(eng.handle, 0, *([None] * 4), 10, *([None] * 8)) in (0, -1)
    b = synthetic.make_batch(eng.parameter_map, eng.dims, 9, 1, seed=3)
    with pytest.raises(engine.MpcGpuError):
        eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=1)      # 9 > max_batch: rejected, not truncated


def test_persistent_memory_matches_oracle():
    cfg = "c1_basic"
    eng = engine.Engine(cfg, 0, 64)
    orc = Oracle(cfg)
    assert eng.mem_doubles == orc.mem_doubles
    b = synthetic.make_batch(eng.parameter_map, eng.dims, 24, 1, seed=9)
    mg, mo = np.zeros((b["n"], eng.mem_doubles)), np.zeros((b["n"], orc.mem_doubles))
    og = eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=3, mem=mg)
    oo = orc.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=3, mem=mo)
    np.testing.assert_array_equal(og["exit_code"], oo["exit_code"])
    np.testing.assert_array_equal(mg[:, 0], mo[:, 0])                        # flags: 2 after success, 0 after failure
    ok = oo["exit_code"] == 1
    assert (mg[~ok] == 0).all()
    assert np.abs(mg[ok] - mo[ok]).max() < 1e-6 * max(1.0, np.abs(mo[ok]).max())
    x0b = np.zeros((b["n"], eng.N + 1, eng.nz))
    x0b[:, :, eng.nu:] = oo["xtraj"].reshape(b["n"], eng.N + 1, eng.nx)
    x0b[:, :eng.N, :eng.nu] = oo["utraj"].reshape(b["n"], eng.N, eng.nu)
    x0b = x0b.reshape(b["n"], -1)
    mo2 = mo.copy()
    og2 = eng.solve_batch(b["xinit"], x0b, b["params"], num_iter=2, mem=mo.copy())   # both continue from the SAME blob
    oo2 = orc.solve_batch(b["xinit"], x0b, b["params"], num_iter=2, mem=mo2)
    np.testing.assert_array_equal(og2["exit_code"], oo2["exit_code"])
    ok2 = oo2["exit_code"] == 1
    assert np.abs(og2["xtraj"][ok2] - oo2["xtraj"][ok2]).max() < 1e-6 * max(1.0, np.abs(oo2["xtraj"][ok2]).max())


def test_full_size_batch_properties():
    """BASELINE-size batch (4096 sets x 9 planners = 36 864 problems): properties that need no oracle run
    at that size -- x_0 = xinit, dynamics residual and bounds of the successes, determinism under
    duplication, argmin recomputed in numpy -- plus a random subsample against the oracle."""
    cfg, planners, n_sets = "c2_tmpc12", 9, 4096
    eng = engine.Engine(cfg, 0, n_sets * planners)
    orc = Oracle(cfg)
    half = synthetic.make_batch(eng.parameter_map, eng.dims, n_sets // 2, planners, seed=123)
    xinit = np.concatenate([half["xinit"], half["xinit"]]); x0 = np.concatenate([half["x0"], half["x0"]])
    params = np.concatenate([half["params"], half["params"]])
    off = np.arange(0, n_sets * planners + 1, planners, dtype=np.int32)
    out = eng.solve_batch(xinit, x0, params, num_iter=10)
    n = n_sets * planners
    h = n // 2
    for k in ("xtraj", "utraj", "pobj", "exit_code", "qp_status", "res_eq", "ipm_iters"):
        np.testing.assert_array_equal(out[k][:h], out[k][h:])                # same problem, different warp: same bits
    ok = out["exit_code"] == 1
    assert 0.7 < ok.mean() < 1.0
    X = out["xtraj"].reshape(n, eng.N + 1, eng.nx); U = out["utraj"].reshape(n, eng.N, eng.nu)
    np.testing.assert_allclose(X[ok][:, 0], xinit[ok], atol=1e-9)
    assert (out["res_eq"][ok] <= 1e-2).all()
    lb, ub = orc.bounds(0), orc.bounds(1)
    assert (U[ok] >= lb[:2] - 1e-6).all() and (U[ok] <= ub[:2] + 1e-6).all()
    assert (X[ok][:, 1:eng.N] >= lb[2:] - 1e-6).all() and (X[ok][:, 1:eng.N] <= ub[2:] + 1e-6).all()
    best = eng.select_best(off, out["pobj"], out["exit_code"])
    obj = np.where(ok, out["pobj"], np.inf).reshape(n_sets, planners)
    expect = np.where(np.isfinite(obj.min(axis=1)), obj.argmin(axis=1), -1)    # numpy argmin = first minimum = strict '<'
    np.testing.assert_array_equal(best, expect)
    idx = np.random.default_rng(0).choice(h, 96, replace=False)
    ref = orc.solve_batch(xinit[idx], x0[idx], params[idx], num_iter=10)
    np.testing.assert_array_equal(out["exit_code"][idx], ref["exit_code"])
    okr = ref["exit_code"] == 1
    scale = np.maximum(1.0, np.abs(ref["xtraj"][okr]).max(axis=1))
    assert (np.abs(out["xtraj"][idx][okr] - ref["xtraj"][okr]).max(axis=1) / scale).max() < REL_TOL


def compact_sets(batch, eng, planners):
    """Split a flat synthetic batch into (shared block per set, indices that differ between the planners
    of a set, their per-planner values) -- the inputs of mpcgpu_solve_sets."""
    n = batch["n"]
    n_sets = n // planners
    P = batch["params"].reshape(n_sets, planners, eng.N, eng.npar)
    differs = np.nonzero((P != P[:, :1]).any(axis=(0, 1, 2)))[0].astype(np.int32)
    shared = np.ascontiguousarray(P[:, 0])
    vals = np.ascontiguousarray(P[..., differs])
    xs = np.ascontiguousarray(batch["xinit"].reshape(n_sets, planners, eng.nx)[:, 0])
    return n_sets, xs, shared, differs, vals


@pytest.mark.parametrize("cfg,planners", [("tmpc_shipped", 5), ("c2_tmpc12", 9)])
def test_compact_set_entry_equals_flat_entry(cfg, planners):
    """mpcgpu_solve_sets (shared parameter block per set + per-planner overrides + fused selection) gives
    bit-identical results to mpcgpu_solve_batch + mpcgpu_select_best on the expanded inputs."""
    eng = engine.Engine(cfg, 0, 512)
    b = synthetic.make_batch(eng.parameter_map, eng.dims, 12, planners, seed=17)
    n_sets, xs, shared, idx, vals = compact_sets(b, eng, planners)
    assert 0 < idx.size < eng.npar // 2                       # only the guidance halfspaces (+ consistency reference) differ
    flat = eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=4)
    best = eng.select_best(b["set_offsets"], flat["pobj"], flat["exit_code"])
    out = eng.solve_sets(n_sets, planners, xs, shared, b["x0"], idx, vals, num_iter=4)
    for k in ("xtraj", "utraj", "pobj", "exit_code", "qp_status", "res_eq"):
        np.testing.assert_array_equal(out[k], flat[k])
    np.testing.assert_array_equal(out["best"], best)


def test_gpu_converged_iterate_is_a_kkt_point_of_the_reference_nlp():
    """Same algorithm-independent NLP-level check as tests/test_nlp_optimality.py, on the GPU result."""
    from test_nlp_optimality import nlp_kkt
    cfg, planners = "c2_tmpc12", 9
    eng = engine.Engine(cfg, 0, 64)
    orc = Oracle(cfg)               # model functions only (pinned to the reference by the golden tests)
    b = synthetic.make_batch(eng.parameter_map, eng.dims, 1, planners, seed=41)
    mem = np.zeros((b["n"], eng.mem_doubles))
    r = eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=40, mem=mem)
    checked = 0
    for i in np.nonzero(r["exit_code"] == 1)[0]:
        x = r["xtraj"][i].reshape(eng.N + 1, eng.nx); u = r["utraj"][i].reshape(eng.N, eng.nu)
        stat, dyn, feas, comp, x0err = nlp_kkt(orc, b["xinit"][i], b["params"][i], x, u, mem[i])
        if dyn > 1e-6:
            continue
        checked += 1
        assert stat < 5e-4 and feas < 1e-5 and comp < 1e-3 and x0err < 1e-9, (i, stat, feas, comp)
    assert checked >= 5
