"""Wire formats: serialized mpc_planner_msgs/ObstacleGMM | ObstacleArray -> obstacle tables -> ellipsoid
parameter slots, and the engine's records -> serialized MPCMetrics.  The messages are built here by an independent
struct-based ROS 1 serializer written after the .msg definitions (mpc_planner_msgs/msg/*.msg)."""
import ctypes
import struct

import numpy as np
import pytest

from oscar_mpc_planner_mr_modification_b200 import engine, synthetic

ERR_ARG = -1


class Track(ctypes.Structure):
    _fields_ = [("id", ctypes.c_int), ("x", ctypes.c_double), ("y", ctypes.c_double), ("angle", ctypes.c_double),
                ("radius", ctypes.c_double), ("n_steps", ctypes.c_int)]


# ---- ROS 1 serialization of the reference's messages, written from the .msg files --------------------------------
def s_string(s):
    b = s.encode()
    return struct.pack("<I", len(b)) + b


def s_header(seq=0, sec=0, nsec=0, frame=""):
    return struct.pack("<III", seq, sec, nsec) + s_string(frame)


def s_pose(x, y, yaw, z=0.0):
    return struct.pack("<7d", x, y, z, 0.0, 0.0, np.sin(yaw / 2.0), np.cos(yaw / 2.0))


def s_f64s(a):
    a = np.asarray(a, np.float64)
    return struct.pack("<I", a.size) + a.tobytes()


def s_gaussian(path_xyyaw, frame="map"):
    out = s_header(1, 2, 3, frame) + struct.pack("<I", len(path_xyyaw))
    for k, (x, y, yaw) in enumerate(path_xyyaw):
        out += s_header(k, 10, 20, frame) + s_pose(x, y, yaw)
    return out + s_f64s([-1.0] * len(path_xyyaw)) + s_f64s([-1.0] * len(path_xyyaw))


def s_obstacle_gmm(oid, pose, paths, probs):
    out = struct.pack("<i", oid) + s_pose(*pose) + struct.pack("<I", len(paths))
    for p in paths:
        out += s_gaussian(p)
    return out + s_f64s(probs)


def s_obstacle_array(obstacles):
    out = s_header(7, 8, 9, "world") + struct.pack("<I", len(obstacles))
    for o in obstacles:
        out += s_obstacle_gmm(*o)
    return out


def lib():
    L = engine.load_library()
    L.mpcgpu_wire_parse_obstacle_gmm.restype = ctypes.c_long
    L.mpcgpu_wire_parse_obstacle_gmm.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.POINTER(Track), ctypes.c_int, ctypes.c_void_p]
    L.mpcgpu_wire_parse_obstacle_array.restype = ctypes.c_long
    L.mpcgpu_wire_parse_obstacle_array.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.POINTER(Track),
                                                   ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
    L.mpcgpu_obstacle_table.argtypes = [ctypes.POINTER(Track), ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.c_void_p]
    L.mpcgpu_wire_serialize_metrics.restype = ctypes.c_long
    L.mpcgpu_wire_serialize_metrics.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]
    L.mpcgpu_metrics_from_set.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                          ctypes.c_void_p]
    return L


def rand_paths(rng, n, steps):
    return [[(float(rng.uniform(-5, 25)), float(rng.uniform(-6, 6)), float(rng.uniform(-3, 3))) for _ in range(steps)] for _ in range(n)]


def test_parse_obstacle_gmm_roundtrip_and_truncation():
    L = lib()
    rng = np.random.default_rng(0)
    paths = rand_paths(rng, 2, 30)          # two Gaussians: only the first is used (jules_ros1_jackalplanner.cpp:593-601)
    msg = s_obstacle_gmm(42, (1.5, -2.5, 0.7), paths, [0.6, 0.4])
    t, steps = Track(), np.full((40, 3), np.nan)
    used = L.mpcgpu_wire_parse_obstacle_gmm(msg, len(msg), ctypes.byref(t), 40, steps.ctypes.data)
    assert used == len(msg)
    assert (t.id, t.n_steps, t.radius) == (42, 30, 0.0)
    np.testing.assert_allclose([t.x, t.y, t.angle], [1.5, -2.5, 0.7], atol=1e-15)
    np.testing.assert_allclose(steps[:30], np.array(paths[0]), atol=1e-14)         # quaternion -> yaw
    assert np.isnan(steps[30:]).all()
    # clamped to max_steps, still consumes the whole message
    steps2 = np.zeros((10, 3))
    assert L.mpcgpu_wire_parse_obstacle_gmm(msg, len(msg), ctypes.byref(t), 10, steps2.ctypes.data) == len(msg) and t.n_steps == 10
    # message without a trajectory (the callback ignores it): n_steps = 0
    empty = s_obstacle_gmm(3, (0, 0, 0), [], [])
    assert L.mpcgpu_wire_parse_obstacle_gmm(empty, len(empty), ctypes.byref(t), 10, steps2.ctypes.data) == len(empty) and t.n_steps == 0
    # every truncation is rejected, never read past the end
    for cut in (0, 3, 4 + 55, len(msg) // 2, len(msg) - 1):
        assert L.mpcgpu_wire_parse_obstacle_gmm(msg[:cut], cut, ctypes.byref(t), 40, steps.ctypes.data) == ERR_ARG
    bad = bytearray(msg)
    bad[4 + 56:4 + 60] = struct.pack("<I", 0x7fffffff)      # absurd Gaussian count
    assert L.mpcgpu_wire_parse_obstacle_gmm(bytes(bad), len(bad), ctypes.byref(t), 40, steps.ctypes.data) == ERR_ARG


def numpy_table(tracks, steps, N, M, state):
    """ensureObstacleSize + table, written after mpc_planner/src/data_preparation.cpp:97-170"""
    n = len(tracks)
    idx = list(range(n))
    if n > M:
        d = np.array([np.cos(state[2]), np.sin(state[2])])
        dist = []
        for i in range(n):
            md = 1e5
            for k in range(N):
                ego = np.array(state[:2]) + state[3] * k * d
                dd = (k + 1) * 0.6 * np.sqrt(((steps[i][k, :2] - ego) ** 2).sum())
                md = min(md, dd)
            dist.append(md)
        idx = sorted(idx, key=lambda i: dist[i])[:M]
    T = np.zeros((N, M, 4))
    for j in range(M):
        if j < len(idx):
            T[:, j, :3] = steps[idx[j]][:N]
            T[:, j, 3] = tracks[idx[j]][1]
        else:
            T[:, j] = [state[0] + 100.0, state[1] + 100.0, 0.0, 0.0]
    return T, min(n, M)


@pytest.mark.parametrize("n_obs", [0, 5, 12, 17])
def test_obstacle_array_to_table(n_obs):
    L = lib()
    N, M, S = 30, 12, 32
    rng = np.random.default_rng(n_obs)
    paths = rand_paths(rng, n_obs, N)
    obstacles = [(100 + i, (paths[i][0][0], paths[i][0][1], 0.1 * i), [paths[i]], [1.0]) for i in range(n_obs)]
    msg = s_obstacle_array(obstacles)
    tracks = (Track * 20)()
    steps = np.zeros((20, S, 3))
    nt = ctypes.c_int(-1)
    assert L.mpcgpu_wire_parse_obstacle_array(msg, len(msg), 20, S, tracks, steps.ctypes.data, ctypes.byref(nt)) == len(msg)
    assert nt.value == n_obs
    for i in range(n_obs):
        assert tracks[i].id == 100 + i and tracks[i].n_steps == N
        tracks[i].radius = 0.3 + 0.01 * i                       # CONFIG["obstacle_radius"]: not on the wire
    state = np.array([0.4, -0.2, 0.15, 1.7])
    table = np.full((N, M, 4), np.nan)
    kept = L.mpcgpu_obstacle_table(tracks, steps.ctypes.data, n_obs, S, N, M, state.ctypes.data, table.ctypes.data)
    want, wkept = numpy_table([(100 + i, 0.3 + 0.01 * i) for i in range(n_obs)], [np.array(p) for p in paths], N, M, state)
    assert kept == wkept
    np.testing.assert_allclose(table, want, atol=1e-13)
    # a track shorter than the horizon is an error (the reference would read past the prediction)
    if n_obs:
        tracks[0].n_steps = N - 1
        assert L.mpcgpu_obstacle_table(tracks, steps.ctypes.data, n_obs, S, N, M, state.ctypes.data, table.ctypes.data) == ERR_ARG


class Metrics(ctypes.Structure):
    _fields_ = [("seq", ctypes.c_uint), ("stamp_sec", ctypes.c_uint), ("stamp_nsec", ctypes.c_uint), ("frame_id", ctypes.c_char_p),
                ("robot_name", ctypes.c_char_p), ("solve_time_ms", ctypes.c_double), ("success_rate", ctypes.c_double),
                ("iterations", ctypes.c_int), ("exit_code", ctypes.c_int), ("objective_value", ctypes.c_double),
                ("objective_values_all_planners", ctypes.POINTER(ctypes.c_double)), ("n_planners", ctypes.c_int),
                ("selected_planner_index", ctypes.c_int), ("num_of_guidance_found", ctypes.c_int), ("used_guidance", ctypes.c_ubyte)]


def test_metrics_message_bytes():
    L = lib()
    pobj = np.array([3.5, 1.25, 9.0, 2.0, 7.0])
    exit_code = np.array([1, 1, 4, 0, 1], np.int32)
    guided = np.array([1, 1, 1, 1, 0], np.uint8)
    vals = np.zeros(5)
    m = Metrics(seq=11, stamp_sec=100, stamp_nsec=200, frame_id=b"map", robot_name=b"jackal1", solve_time_ms=3.25, iterations=10)
    assert L.mpcgpu_metrics_from_set(ctypes.byref(m), 5, pobj.ctypes.data, exit_code.ctypes.data, 1, guided.ctypes.data, vals.ctypes.data) == 0
    assert (m.selected_planner_index, m.exit_code, m.objective_value, m.used_guidance, m.num_of_guidance_found) == (1, 1, 1.25, 1, 4)
    np.testing.assert_array_equal(vals, [3.5, 1.25, -1.0, -1.0, 7.0])
    buf = ctypes.create_string_buffer(512)
    n = L.mpcgpu_wire_serialize_metrics(ctypes.byref(m), buf, 512)
    want = (s_header(11, 100, 200, "map") + s_string("jackal1") + struct.pack("<ddiid", 3.25, 0.6, 10, 1, 1.25) + s_f64s(vals) +
            struct.pack("<iiBBii", 0, 0, 0, 1, 1, 4) + s_string("") + s_string("") + s_f64s([]) + struct.pack("<dd", 0.0, 0.0) +
            s_string("") + struct.pack("<iid", 0, 0, 0.0) + struct.pack("<IIII", 0, 0, 0, 0))
    assert n == len(want) and buf.raw[:n] == want
    assert L.mpcgpu_wire_serialize_metrics(ctypes.byref(m), buf, 40) == ERR_ARG       # buffer too small: nothing half-written is reported
    # all planners failed: planner 0's exit code is reported (guidance_constraints.cpp:441)
    ec = np.array([4, 0, 4, 4, 4], np.int32)
    L.mpcgpu_metrics_from_set(ctypes.byref(m), 5, pobj.ctypes.data, ec.ctypes.data, -1, guided.ctypes.data, vals.ctypes.data)
    assert (m.selected_planner_index, m.exit_code, m.used_guidance) == (-1, 4, 0)


def ellipsoid_layout(pm):
    base = pm["ellipsoid_obst_0_x"]
    stride = pm["ellipsoid_obst_1_x"] - base
    off = [pm["ellipsoid_obst_0_" + k] - base for k in ("x", "y", "psi", "major", "minor", "chi", "r")]
    M = sum(1 for k in pm if k.startswith("ellipsoid_obst_") and k.endswith("_x"))
    return base, stride, np.array(off, np.int32), M


@pytest.mark.gpu
def test_solve_sets_tracks_matches_host_built_parameters():
    """Wire bytes -> tracks -> tables -> (device) ellipsoid slots + guidance halfspaces -> solve == the same solve from
    host-built parameter blocks (synthetic.make_batch follows ellipsoid_constraints.cpp:34-90 on the host)."""
    L = lib()
    cfg, planners, n_sets = "c2_tmpc12", 9, 12
    eng = engine.Engine(cfg, 0, 256)
    b = synthetic.make_batch(eng.parameter_map, eng.dims, n_sets, planners, seed=77)
    N, npar, nx = eng.N, eng.npar, eng.nx
    base, stride, off, M = ellipsoid_layout(eng.parameter_map)
    assert (stride, M) == (7, 12) and sorted(off.tolist()) == list(range(7))
    # the obstacles of every set travel as one serialized ObstacleArray, like on the ROS topic
    tables = np.zeros((n_sets, N, M, 4))
    xs = np.ascontiguousarray(b["xinit"].reshape(n_sets, planners, nx)[:, 0])
    for s_ in range(n_sets):
        ob = b["obst_pred"][s_]                                              # (N, M, 2)
        msg = s_obstacle_array([(j, (ob[0, j, 0], ob[0, j, 1], 0.0), [[(ob[k, j, 0], ob[k, j, 1], 0.0) for k in range(N)]], [1.0])
                                for j in range(M)])
        tracks, steps, nt = (Track * M)(), np.zeros((M, N, 3)), ctypes.c_int()
        assert L.mpcgpu_wire_parse_obstacle_array(msg, len(msg), M, N, tracks, steps.ctypes.data, ctypes.byref(nt)) == len(msg)
        for j in range(M):
            tracks[j].radius = synthetic.OBSTACLE_RADIUS
        st = np.array([xs[s_, 0], xs[s_, 1], xs[s_, 2], xs[s_, 3]])
        assert L.mpcgpu_obstacle_table(tracks, steps.ctypes.data, nt.value, N, N, M, st.ctypes.data, tables[s_].ctypes.data) == M
    ref = eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=5)
    best_ref = eng.select_best(b["set_offsets"], ref["pobj"], ref["exit_code"])
    shared = np.ascontiguousarray(b["params"].reshape(n_sets, planners, N, npar)[:, 0]).copy()
    lin_base, lin_count = eng.lin_constraint_block()
    shared[:, :, lin_base:lin_base + 3 * lin_count] = np.nan             # both blocks must be rewritten on the device
    shared[:, :, base:base + stride * M] = np.nan
    out = eng.alloc_outputs(b["n"])
    best = np.zeros(n_sets, np.int32)
    vp = ctypes.c_void_p
    L.mpcgpu_solve_sets_tracks.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, ctypes.c_int, vp, vp, ctypes.c_int, ctypes.c_int,
                                           ctypes.c_double, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int] + [vp] * 11
    P = lambda a: a.ctypes.data
    rc = L.mpcgpu_solve_sets_tracks(eng.handle, n_sets, planners, P(xs), P(shared), P(b["x0"]), M, P(tables), P(b["guided"]), lin_base, lin_count,
                                    b["robot_radius"], base, stride, P(off), None, 5, P(out["xtraj"]), P(out["utraj"]), P(out["pobj"]),
                                    P(out["exit_code"]), P(out["qp_status"]), P(out["res_eq"]), None, None, None, P(best), None)
    assert rc == 0, eng.last_error()
    np.testing.assert_array_equal(out["exit_code"], ref["exit_code"])
    np.testing.assert_array_equal(best, best_ref)
    ok = ref["exit_code"] == 1
    assert ok.sum() > 50 and np.abs(out["xtraj"][ok] - ref["xtraj"][ok]).max() < 1e-9
