"""Guidance halfspace construction: LinearizedConstraints::update / projectToSafety / setParameters for
the topology constraints (mpc_planner_modules/src/linearized_constraints.cpp:43-189), restated in the oracle and built on
the device by mpcgpu_guidance_halfspaces_device / mpcgpu_solve_sets_guided."""
import numpy as np
import pytest

from oracle_binding import Oracle
from oscar_mpc_planner_mr_modification_b200 import engine, synthetic

CFG, PLANNERS = "c2_tmpc12", 9


def lin_block(pm):
    cnt = sum(1 for k in pm if k.startswith("lin_constraint_") and k.endswith("_a1"))
    return pm["lin_constraint_0_a1"], cnt


def numpy_halfspaces(b, dims, pm, n_sets, planners, project=True):
    """Independent numpy restatement of linearized_constraints.cpp:49-189 (loops written after the C++ source)."""
    N, nx, nu, npar = dims["N"], dims["nx"], dims["nu"], dims["npar"]
    nz = nx + nu
    base, cnt = lin_block(pm)
    P = b["params"].reshape(-1, N, npar).copy()
    x0 = b["x0"].reshape(-1, N + 1, nz)
    r = 1e-3 + b["robot_radius"]
    ob_all = b["obst_pred"]
    M = ob_all.shape[2]
    for q in range(n_sets * planners):
        s = q // planners
        dummy_b = b["xinit"][q, 0] + 100.0
        for k in range(N):
            P[q, k, base:base + 3 * cnt] = np.tile([1.0, 0.0, dummy_b], cnt)
            if k == 0 or not b["guided"][q]:
                continue
            ob = ob_all[s, k - 1]
            pos = x0[q, k, nu:nu + 2].copy()
            if project:
                for _ in range(3):
                    for j in range(M):
                        if np.sqrt(((pos - ob[j]) ** 2).sum()) < r:
                            def proj(p, c, toward):
                                d = p - c
                                if np.sqrt((d ** 2).sum()) < r:
                                    t = toward - c
                                    return c + t / np.sqrt((t ** 2).sum()) * r
                                return p
                            ra = 2.0 * proj(pos, ob[0], pos) - pos
                            pos = 0.5 * (pos + 2.0 * proj(ra, ob[j], pos) - ra)
            for j in range(min(M, cnt)):
                d = ob[j] - pos
                dist = np.sqrt(d[0] * d[0] + d[1] * d[1])
                a1, a2 = d[0] / dist, d[1] / dist
                P[q, k, base + 3 * j:base + 3 * j + 3] = [a1, a2, a1 * ob[j, 0] + a2 * ob[j, 1] - r]
    return P.reshape(-1, N * npar)


def make(n_sets, seed, inside=False):
    orc = Oracle(CFG)
    b = synthetic.make_batch(orc.parameter_map, orc.dims, n_sets, PLANNERS, seed=seed)
    if inside:      # move some warm-start positions INTO an obstacle's circle so that projectToSafety has work to do
        N, nz, nu = orc.N, orc.nz, orc.nu
        x0 = b["x0"].reshape(-1, N + 1, nz)
        rng = np.random.default_rng(seed)
        for q in range(0, b["n"], 2):
            s = q // PLANNERS
            for k in rng.choice(np.arange(1, N), 4, replace=False):
                j = rng.integers(0, b["obst_pred"].shape[2])
                x0[q, k, nu:nu + 2] = b["obst_pred"][s, k - 1, j] + rng.uniform(-0.1, 0.1, 2)
    return orc, b


@pytest.mark.parametrize("inside", [False, True])
def test_oracle_halfspaces_match_numpy_restatement(inside):
    orc, b = make(6, 11, inside)
    base, cnt = lin_block(orc.parameter_map)
    assert cnt == 12 and b["obst_pred"].shape == (6, orc.N, 12, 2)
    want = numpy_halfspaces(b, orc.dims, orc.parameter_map, 6, PLANNERS)
    got = b["params"].copy()
    got.reshape(-1, orc.N, orc.npar)[:, :, base:base + 3 * cnt] = np.nan      # every slot must be (re)written
    xs = b["xinit"].reshape(6, PLANNERS, -1)[:, 0]
    orc.guidance_halfspaces(6, PLANNERS, xs, b["x0"], b["obst_pred"], b["guided"], b["robot_radius"], base, cnt, got)
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-13)
    if not inside:
        # the synthetic generator builds the same halfspaces on the host (warm starts are collision free: projection is a no-op)
        np.testing.assert_allclose(got, b["params"], rtol=0, atol=1e-12)
    else:
        # after projectToSafety the linearisation point is outside every obstacle circle it was pushed out of: b - a.pos <= 0
        # cannot be asserted for all obstacles (3 sweeps only); what must hold: unit normals, finite offsets
        G = got.reshape(-1, orc.N, orc.npar)[:, 1:, base:base + 3 * cnt].reshape(-1, cnt, 3)
        np.testing.assert_allclose(np.hypot(G[..., 0], G[..., 1]), 1.0, atol=1e-12)


def test_dummy_rows_for_stage0_and_nonguided_planner():
    orc, b = make(3, 5)
    base, cnt = lin_block(orc.parameter_map)
    P = b["params"].copy()
    xs = b["xinit"].reshape(3, PLANNERS, -1)[:, 0]
    orc.guidance_halfspaces(3, PLANNERS, xs, b["x0"], b["obst_pred"][:, :, :5], b["guided"], b["robot_radius"], base, cnt, P)
    P = P.reshape(3, PLANNERS, orc.N, orc.npar)[..., base:base + 3 * cnt].reshape(3, PLANNERS, orc.N, cnt, 3)
    dummy = np.stack([np.ones(3), np.zeros(3), xs[:, 0] + 100.0], axis=1)
    assert (P[:, :, 0] == dummy[:, None, None, :]).all()                 # k = 0 (linearized_constraints.cpp:155-166)
    assert (P[:, PLANNERS - 1] == dummy[:, None, None, :]).all()         # non-guided planner: update(state, empty_data_)
    assert (P[:, :PLANNERS - 1, 1:, 5:] == dummy[:, None, None, None, :]).all()   # only 5 obstacles: slots 5.. are dummies
    assert not (P[:, 0, 1:, :5, 0] == 1.0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("inside", [False, True])
def test_device_halfspaces_bit_identical_to_oracle(inside):
    import torch
    orc, b = make(8, 21, inside)
    eng = engine.Engine(CFG, 0, 128)
    base, cnt = eng.lin_constraint_block()
    want = b["params"].copy()
    xs = np.ascontiguousarray(b["xinit"].reshape(8, PLANNERS, -1)[:, 0])
    orc.guidance_halfspaces(8, PLANNERS, xs, b["x0"], b["obst_pred"], b["guided"], b["robot_radius"], base, cnt, want)
    dev = torch.device("cuda", 0)
    d_params = torch.from_numpy(b["params"].copy()).to(dev)
    d_params.view(-1, orc.N, orc.npar)[:, :, base:base + 3 * cnt] = float("nan")
    d_xs, d_x0 = torch.from_numpy(xs).to(dev), torch.from_numpy(b["x0"]).to(dev)
    d_ob, d_g = torch.from_numpy(b["obst_pred"]).to(dev), torch.from_numpy(b["guided"]).to(dev)
    torch.cuda.synchronize()
    eng.guidance_halfspaces_device(8, PLANNERS, d_xs.data_ptr(), d_x0.data_ptr(), d_ob.data_ptr(), b["obst_pred"].shape[2],
                                   d_g.data_ptr(), b["robot_radius"], d_params.data_ptr())
    eng.sync()
    got = d_params.cpu().numpy()
    np.testing.assert_array_equal(got, want)          # products and sums are left unfused on the device: bit-exact


@pytest.mark.gpu
def test_solve_sets_guided_matches_host_built_parameters():
    eng = engine.Engine(CFG, 0, 512)
    n_sets = 24
    b = synthetic.make_batch(eng.parameter_map, eng.dims, n_sets, PLANNERS, seed=31)
    N, npar, nx = eng.N, eng.npar, eng.nx
    ref = eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=5)
    best_ref = eng.select_best(b["set_offsets"], ref["pobj"], ref["exit_code"])
    Pv = b["params"].reshape(n_sets, PLANNERS, N, npar)
    shared = np.ascontiguousarray(Pv[:, 0]).copy()
    base, cnt = eng.lin_constraint_block()
    shared[:, :, base:base + 3 * cnt] = -7.0          # garbage: must be overwritten on the device
    xs = np.ascontiguousarray(b["xinit"].reshape(n_sets, PLANNERS, nx)[:, 0])
    out = eng.solve_sets_guided(n_sets, PLANNERS, xs, shared, b["x0"], b["obst_pred"], b["guided"], b["robot_radius"], num_iter=5)
    np.testing.assert_array_equal(out["exit_code"], ref["exit_code"])
    np.testing.assert_array_equal(out["best"], best_ref)
    ok = ref["exit_code"] == 1
    assert ok.sum() > 100
    assert np.abs(out["xtraj"][ok] - ref["xtraj"][ok]).max() < 1e-6


def static_rows(n_sets, N, n_static, seed):
    rng = np.random.default_rng(seed)
    ang = rng.uniform(-np.pi, np.pi, (n_sets, N, n_static))
    return np.ascontiguousarray(np.stack([np.cos(ang), np.sin(ang), rng.uniform(5.0, 9.0, ang.shape)], axis=-1))


def test_oracle_static_halfspaces_follow_the_obstacle_rows():
    """module_data.static_obstacles (linearized_constraints.cpp:107-127): behind the obstacle rows of a guided planner, from
    slot 0 for the non-guided one (empty obstacle list), never at stage 0."""
    orc, b = make(3, 7)
    base, cnt = lin_block(orc.parameter_map)
    b["obst_pred"] = np.ascontiguousarray(b["obst_pred"][:, :, :5])
    st = static_rows(3, orc.N, 2, 1)
    xs = np.ascontiguousarray(b["xinit"].reshape(3, PLANNERS, -1)[:, 0])
    plain = orc.guidance_halfspaces(3, PLANNERS, xs, b["x0"], b["obst_pred"], b["guided"], b["robot_radius"], base, cnt, b["params"].copy())
    P = orc.guidance_halfspaces(3, PLANNERS, xs, b["x0"], b["obst_pred"], b["guided"], b["robot_radius"], base, cnt, b["params"].copy(), st)
    blk = lambda a: a.reshape(3, PLANNERS, orc.N, orc.npar)[..., base:base + 3 * cnt].reshape(3, PLANNERS, orc.N, cnt, 3)
    P, plain = blk(P), blk(plain)
    assert (P[:, :, 0] == plain[:, :, 0]).all()
    assert (P[:, :PLANNERS - 1, 1:, 5:7] == st[:, None, 1:]).all() and (P[:, PLANNERS - 1, 1:, 0:2] == st[:, 1:]).all()
    keep = np.ones(cnt, bool); keep[5:7] = False
    assert (P[:, :PLANNERS - 1][:, :, :, keep] == plain[:, :PLANNERS - 1][:, :, :, keep]).all()
    assert (P[:, PLANNERS - 1, :, 2:] == plain[:, PLANNERS - 1, :, 2:]).all()


@pytest.mark.gpu
def test_solve_sets_guided_with_static_halfspaces():
    """static halfspaces through struct mpcgpu_set_options: the device-built parameter block equals the oracle's"""
    eng = engine.Engine(CFG, 0, 256)
    orc = Oracle(CFG)
    n_sets = 6
    b = synthetic.make_batch(eng.parameter_map, eng.dims, n_sets, PLANNERS, seed=33)
    b["obst_pred"] = np.ascontiguousarray(b["obst_pred"][:, :, :8])      # 8 obstacles + 3 static rows + 1 dummy slot
    st = static_rows(n_sets, eng.N, 3, 2)
    base, cnt = eng.lin_constraint_block()
    xs = np.ascontiguousarray(b["xinit"].reshape(n_sets, PLANNERS, eng.nx)[:, 0])
    want = orc.guidance_halfspaces(n_sets, PLANNERS, xs, b["x0"], b["obst_pred"], b["guided"], b["robot_radius"], base, cnt, b["params"].copy(), st)
    ref = eng.solve_batch(b["xinit"], b["x0"], want, num_iter=3)
    shared = np.ascontiguousarray(b["params"].reshape(n_sets, PLANNERS, eng.N, eng.npar)[:, 0]).copy()
    out = eng.solve_sets_guided(n_sets, PLANNERS, xs, shared, b["x0"], b["obst_pred"], b["guided"], b["robot_radius"], num_iter=3,
                                static_halfspaces=st)
    for k in ("xtraj", "utraj", "pobj", "exit_code"):
        np.testing.assert_array_equal(out[k], ref[k])
