"""Several GPUs behind one handle (mpcgpu_multi_*): by-set partition into contiguous ranges, one host thread and
stream per device, no collective; the gathered per-set table must equal the single-GPU result bit for bit (north_star: "only
the per-problem cost and feasibility flags are gathered to pick the best trajectory")."""
import ctypes

import numpy as np
import pytest

from oscar_mpc_planner_mr_modification_b200 import engine, sharding, synthetic

CFG, PLANNERS = "c2_tmpc12", 9


def test_shard_range_matches_the_python_layout():
    lib = engine.load_library()
    for n in (0, 1, 7, 16, 4097):
        for w in (1, 2, 3, 8):
            for i in range(w):
                b, e = ctypes.c_int(), ctypes.c_int()
                assert lib.mpcgpu_multi_shard_range(n, w, i, ctypes.byref(b), ctypes.byref(e)) == 0
                assert (b.value, e.value) == sharding.shard_range(n, w, i)
    assert lib.mpcgpu_multi_shard_range(4, 2, 2, ctypes.byref(b), ctypes.byref(e)) == -1


def _devices():
    import torch
    n = torch.cuda.device_count()
    lists = [[0, 0, 0]]                      # three engines on one GPU: exercises the uneven by-set partition on any box
    if n >= 2:
        lists.append(list(range(n)))         # every GPU of the box
    return lists


@pytest.mark.gpu
def test_multi_device_sets_equal_single_device_bit_for_bit():
    from test_gpu_components import compact_sets
    single = engine.Engine(CFG, 0, 512)
    single.set_kernel_mode(engine.KERNEL_STAGE)      # bit-for-bit needs the same kernel on both sides (AUTO picks by batch size)
    n_sets = 20                              # 20 sets over 3 engines: 6 + 6 + 8
    b = synthetic.make_batch(single.parameter_map, single.dims, n_sets, PLANNERS, seed=77)
    ns, xs, shared, idx, vals = compact_sets(b, single, PLANNERS)
    scale = np.where(np.arange(b["n"]) % PLANNERS == 2, 0.75, 1.0)
    ref = single.solve_sets(ns, PLANNERS, xs, shared, b["x0"], idx, vals, num_iter=4, obj_scale=scale)
    for devs in _devices():
        multi = engine.MultiEngine(CFG, devs, 256)
        multi.set_kernel_mode(engine.KERNEL_STAGE)
        out = multi.solve_sets(ns, PLANNERS, xs, shared, b["x0"], idx, vals, num_iter=4, obj_scale=scale)
        for k in ("xtraj", "utraj", "pobj", "exit_code", "qp_status", "res_eq", "best"):
            np.testing.assert_array_equal(out[k], ref[k], err_msg="%s devices %s" % (k, devs))
        # decision record + ONE trajectory per set instead of every planner's: the north_star gather
        slim = multi.solve_sets(ns, PLANNERS, xs, shared, b["x0"], idx, vals, num_iter=4, obj_scale=scale, best_only=True)
        np.testing.assert_array_equal(slim["best"], ref["best"])
        np.testing.assert_array_equal(slim["pobj"], ref["pobj"])
        np.testing.assert_array_equal(slim["exit_code"], ref["exit_code"])
        chosen = np.arange(ns) * PLANNERS + np.maximum(ref["best"], 0)
        np.testing.assert_array_equal(slim["best_xtraj"], ref["xtraj"][chosen])
        np.testing.assert_array_equal(slim["best_utraj"], ref["utraj"][chosen])
        assert multi.last_kernel_ms() > 0
        multi.close()


@pytest.mark.gpu
def test_multi_device_flat_batch_and_guided_sets():
    single = engine.Engine(CFG, 0, 512)
    single.set_kernel_mode(engine.KERNEL_STAGE)
    n_sets = 11
    b = synthetic.make_batch(single.parameter_map, single.dims, n_sets, PLANNERS, seed=78)
    ref = single.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=3)
    xs = np.ascontiguousarray(b["xinit"].reshape(n_sets, PLANNERS, single.nx)[:, 0])
    shared = np.ascontiguousarray(b["params"].reshape(n_sets, PLANNERS, single.N, single.npar)[:, 0])
    gref = single.solve_sets_guided(n_sets, PLANNERS, xs, shared, b["x0"], b["obst_pred"], b["guided"], b["robot_radius"], num_iter=3)
    lin_base, lin_count = single.lin_constraint_block()
    for devs in _devices():
        multi = engine.MultiEngine(CFG, devs, 256)
        multi.set_kernel_mode(engine.KERNEL_STAGE)
        out = multi.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=3)
        for k in ("xtraj", "utraj", "pobj", "exit_code", "qp_status", "res_eq", "ipm_iters"):
            np.testing.assert_array_equal(out[k], ref[k])
        g = multi.solve_sets(n_sets, PLANNERS, xs, shared, b["x0"], None, None, num_iter=3,
                             guided_args=(b["obst_pred"], b["guided"], b["robot_radius"], lin_base, lin_count))
        for k in ("xtraj", "pobj", "exit_code", "best"):
            np.testing.assert_array_equal(g[k], gref[k])
        # struct-of-tables entry over the devices: the tables are sliced by set (mpcgpu_multi_solve_sets_tables)
        lay = single.table_layout()
        P = b["params"].reshape(n_sets, PLANNERS, single.N, single.npar)
        invariant = np.ascontiguousarray(P[:, 0, 0][:, lay["invariant_idx"]])
        radius = np.full((n_sets, b["obst_pred"].shape[2]), synthetic.OBSTACLE_RADIUS)
        targs = dict(guided=b["guided"], robot_radius=b["robot_radius"], obstacle_radius=radius, num_iter=3)
        tref = single.solve_sets_tables(n_sets, PLANNERS, xs, invariant, b["obst_pred"], b["x0"], **targs)
        t = multi.solve_sets_tables(n_sets, PLANNERS, xs, invariant, b["obst_pred"], b["x0"], **targs)
        for k in ("xtraj", "utraj", "pobj", "exit_code", "qp_status", "best"):
            np.testing.assert_array_equal(t[k], tref[k])
        np.testing.assert_array_equal(t["best"], gref["best"])
        multi.close()
