"""Options of the homotopy-set entries (struct mpcgpu_set_options): the consistency-cost post-processing of
GuidanceConstraints::optimize (mpc_planner_modules/src/guidance_constraints.cpp:384-388,405-408,418-419,1025-1050), the
planners' persistent capsules across control cycles (:17-25,323; acados_solver_interface.cpp:67-77) and static halfspaces
(linearized_constraints.cpp:107-127)."""
import numpy as np
import pytest

from oracle_binding import Oracle
from oscar_mpc_planner_mr_modification_b200 import engine, synthetic

CFG, PLANNERS = "tmpc_shipped", 5          # the repository default: consistency module on


def consistency_case(pmap, dims, n_sets, seed, weight, shift):
    """A control cycle with a previous selection: the previous trajectory of every set is the warm start of planner
    s % 4 displaced by `shift` metres; only that planner's topology matches (has_consistency_enabled)."""
    b = synthetic.make_batch(pmap, dims, n_sets, PLANNERS, seed=seed)
    N, nz, nu = dims["N"], dims["nx"] + dims["nu"], dims["nu"]
    x0 = b["x0"].reshape(n_sets, PLANNERS, N + 1, nz)
    sel = np.arange(n_sets) % 4
    prev = x0[np.arange(n_sets), sel, :N, nu:nu + 2] + np.array([0.0, shift])
    enabled = np.zeros((n_sets, PLANNERS), np.uint8)
    enabled[np.arange(n_sets), sel] = 1
    en = synthetic.apply_consistency(b, pmap, dims, PLANNERS, prev, enabled, weight)
    scale = np.where(enabled.reshape(-1) == 1, 0.75, 1.0)      # previously_selected_: selection_weight_consistency_ (:418-419)
    return b, np.ascontiguousarray(prev), en, scale


def numpy_consistency(xtraj, prev, en, weight, N, nx, planners):
    """calculateConsistencyCostForSolver (:1025-1050), loops as in the C++ source"""
    out = np.zeros(xtraj.shape[0])
    for i in range(xtraj.shape[0]):
        if not en[i]:
            continue
        x = xtraj[i].reshape(N + 1, nx)
        s = 0.0
        for k in range(1, N - 1):
            dx = x[k, 0] - prev[i // planners, k, 0]
            dy = x[k, 1] - prev[i // planners, k, 1]
            s += dx * dx + dy * dy
        out[i] = weight * s
    return out


def test_oracle_consistency_cost_is_the_reference_formula():
    orc = Oracle(CFG)
    b, prev, en, scale = consistency_case(orc.parameter_map, orc.dims, 6, 5, 0.05, 0.4)
    r = orc.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=4)
    best, obj, cons = orc.select_best_cons(b["set_offsets"], r["pobj"], r["exit_code"], r["xtraj"], prev, 0.05, en, obj_scale=scale)
    want = numpy_consistency(r["xtraj"], prev, en, 0.05, orc.N, orc.nx, PLANNERS)
    np.testing.assert_array_equal(cons, want)
    np.testing.assert_array_equal(obj, (r["pobj"] - want) * scale)
    assert (cons[en == 1] > 0).all() and (cons[en == 0] == 0).all()
    # FindBestPlanner on the post-processed objective (:572-590)
    for s in range(6):
        o = np.where(r["exit_code"][s * PLANNERS:(s + 1) * PLANNERS] == 1, obj[s * PLANNERS:(s + 1) * PLANNERS], np.inf)
        assert best[s] == (int(np.argmin(o)) if np.isfinite(o).any() else -1)


@pytest.mark.gpu
@pytest.mark.parametrize("weight,shift", [(0.05, 0.4), (5.0, 1.0)])
def test_set_entry_consistency_on_device(weight, shift):
    """The fused set entry computes the consistency cost ON THE DEVICE from the solved trajectory and subtracts it before
    the selection weight -- same argmin as the oracle, including sets where the subtraction flips it."""
    eng = engine.Engine(CFG, 0, 512)
    orc = Oracle(CFG)
    n_sets = 32
    b, prev, en, scale = consistency_case(eng.parameter_map, eng.dims, n_sets, 9, weight, shift)
    ref = orc.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=10)
    best_ref, obj_ref, cons_ref = orc.select_best_cons(b["set_offsets"], ref["pobj"], ref["exit_code"], ref["xtraj"], prev, weight, en,
                                                        obj_scale=scale)
    best_raw = orc.select_best(b["set_offsets"], ref["pobj"], ref["exit_code"], obj_scale=scale)
    from test_gpu_components import compact_sets
    ns, xs, shared, idx, vals = compact_sets(b, eng, PLANNERS)
    out = eng.solve_sets(ns, PLANNERS, xs, shared, b["x0"], idx, vals, num_iter=10, obj_scale=scale, prev_traj=prev, cons_weight=weight,
                         cons_enabled=en)
    np.testing.assert_array_equal(out["exit_code"], ref["exit_code"])
    np.testing.assert_array_equal(out["best"], best_ref)
    # the device value IS the reference formula evaluated on the device's own trajectory (unfused, stage order): bit-exact
    mine = numpy_consistency(out["xtraj"], prev, en, weight, eng.N, eng.nx, PLANNERS)
    np.testing.assert_array_equal(out["consistency_cost"], mine)
    np.testing.assert_array_equal(out["objective"], (out["pobj"] - mine) * scale)
    ok = ref["exit_code"] == 1
    assert np.abs(out["objective"][ok] - obj_ref[ok]).max() <= 1e-6 * np.maximum(1.0, np.abs(obj_ref[ok])).max()
    if weight > 1.0:
        assert (best_ref != best_raw).sum() >= 1      # the subtraction decides the argmin in at least one set


@pytest.mark.gpu
def test_set_entry_carries_the_planner_capsules_across_cycles():
    """Two control cycles through mpcgpu_solve_sets with the persistent capsule memory: identical to the flat entry with the
    `*solver = *_solver` flag downgrade (QP memory reset, multipliers kept: acados_solver_interface.cpp:67-77) applied by hand."""
    eng = engine.Engine(CFG, 0, 256)
    n_sets = 8
    b = synthetic.make_batch(eng.parameter_map, eng.dims, n_sets, PLANNERS, seed=3)
    from test_gpu_components import compact_sets
    ns, xs, shared, idx, vals = compact_sets(b, eng, PLANNERS)
    mem_a = np.zeros((b["n"], eng.mem_doubles)); mem_b = np.zeros((b["n"], eng.mem_doubles))
    for cycle in range(2):
        flat = eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=2, mem=mem_a)
        out = eng.solve_sets(ns, PLANNERS, xs, shared, b["x0"], idx, vals, num_iter=2, mem=mem_b)
        for k in ("xtraj", "utraj", "pobj", "exit_code", "qp_status", "res_eq"):
            np.testing.assert_array_equal(out[k], flat[k], err_msg="cycle %d %s" % (cycle, k))
        np.testing.assert_array_equal(mem_a, mem_b)
        assert (mem_a[flat["exit_code"] == 1, 0] == 2.0).all() and (mem_a[flat["exit_code"] != 1] == 0.0).all()
        mem_a[:, 0] = np.minimum(mem_a[:, 0], 1.0)      # what the set entry does itself at the start of the next cycle
    fresh = eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=2)
    okk = (flat["exit_code"] == 1) & (fresh["exit_code"] == 1)
    assert np.abs(flat["xtraj"][okk] - fresh["xtraj"][okk]).max() > 1e-9      # the surviving multipliers do change the second cycle


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,planners", [("c2_tmpc12", 9), ("tmpc_shipped", 5)])
def test_struct_of_tables_entry_is_bit_identical_and_small(cfg, planners):
    """stage-invariant parameters once per set + obstacle table + warm starts, expanded on the device --
    bit-identical to the flat entry on the expanded inputs, <= 3 KB host->device per solve for the benchmark configuration."""
    eng = engine.Engine(cfg, 0, 512)
    n_sets = 16
    b = synthetic.make_batch(eng.parameter_map, eng.dims, n_sets, planners, seed=23)
    N, npar, nx = eng.N, eng.npar, eng.nx
    P = b["params"].reshape(n_sets, planners, N, npar)
    lay = eng.table_layout()
    inv_idx = lay["invariant_idx"]
    assert (P[..., inv_idx] == P[:, :1, :1][..., inv_idx]).all()            # what the generator calls stage-invariant IS
    invariant = np.ascontiguousarray(P[:, 0, 0][:, inv_idx])
    radius = np.full((n_sets, b["obst_pred"].shape[2]), synthetic.OBSTACLE_RADIUS)
    stage_idx = stage = None
    if "prev_traj_x" in eng.parameter_map:                                   # the consistency reference: per stage, shared? no: per planner
        pidx = np.array([eng.parameter_map[k] for k in ("consistency_weight", "prev_traj_x", "prev_traj_y")], np.int32)
        pvals = np.ascontiguousarray(P[..., pidx])
    else:
        pidx = pvals = None
    xs = np.ascontiguousarray(b["xinit"].reshape(n_sets, planners, nx)[:, 0])
    flat = eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=5)
    best = eng.select_best(b["set_offsets"], flat["pobj"], flat["exit_code"])
    out = eng.solve_sets_tables(n_sets, planners, xs, invariant, b["obst_pred"], b["x0"], guided=b["guided"], robot_radius=b["robot_radius"],
                                obstacle_radius=radius, param_idx=pidx, planner_params=pvals, num_iter=5)
    np.testing.assert_array_equal(out["exit_code"], flat["exit_code"])
    np.testing.assert_array_equal(out["best"], best)
    ok = flat["exit_code"] == 1
    # the halfspaces are rebuilt on the device (bit-identical to the oracle restatement, not to numpy's): trajectories agree to rounding
    assert np.abs(out["xtraj"][ok] - flat["xtraj"][ok]).max() < 1e-6
    per_solve = out["h2d_bytes"] / b["n"]
    if cfg == "c2_tmpc12":
        assert per_solve <= 3072, per_solve
    # and with the obstacle table in the (x, y, psi, r) form of the wire path: same result bit for bit
    tab4 = np.concatenate([b["obst_pred"], np.zeros_like(b["obst_pred"][..., :1]), np.full_like(b["obst_pred"][..., :1], synthetic.OBSTACLE_RADIUS)], axis=-1)
    out4 = eng.solve_sets_tables(n_sets, planners, xs, invariant, tab4, b["x0"], guided=b["guided"], robot_radius=b["robot_radius"],
                                 param_idx=pidx, planner_params=pvals, num_iter=5)
    for k in ("xtraj", "utraj", "pobj", "exit_code", "qp_status", "res_eq", "best"):
        np.testing.assert_array_equal(out4[k], out[k])
    # against the guided entry fed with the host-built shared block: the device-built block is the same block
    shared = np.ascontiguousarray(P[:, 0])
    g = eng.solve_sets_guided(n_sets, planners, xs, shared, b["x0"], b["obst_pred"], b["guided"], b["robot_radius"], num_iter=5,
                              param_idx=pidx, planner_params=pvals)
    for k in ("xtraj", "utraj", "pobj", "exit_code", "qp_status", "res_eq", "best"):
        np.testing.assert_array_equal(g[k], out[k])
