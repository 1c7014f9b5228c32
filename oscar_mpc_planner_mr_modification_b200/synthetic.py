"""Synthetic obstacle / path / warm-start data for the batched MPC solve path.

Every array is laid out exactly as the reference's `AcadosParameters` block
(mpc_planner_solver/include/mpc_planner_solver/acados_solver_interface.h:51-91):
  xinit[nx], x0[(nu+nx)(N+1)] stage-major [u_k, x_k], all_parameters[N*npar] (index k*npar+idx).
Parameter VALUES follow the conventions of the reference's C++ modules:
  * weights                    mpc_planner_jackalsimulator/config/settings.yaml:78-92
  * spline block per stage     mpc_planner_modules/src/contouring.cpp:52-126 (same for every k)
  * ellipsoid obstacles        ellipsoid_constraints.cpp:34-90   (stage k <- prediction k-1, dummies at k=0)
  * guidance halfspaces        linearized_constraints.cpp:49-189 (from the warm-start positions, k>=1;
                               dummy (1,0,x+100) at k=0 and for the non-guided planner)
  * constant-velocity predictions  mpc_planner/src/data_preparation.cpp:64-81
  * braking roll-out           mpc_planner_solver/src/acados_solver_interface.cpp:303-342
The guidance planner itself (external `guidance_planner` package) is replaced by a synthetic
stand-in: planner h follows a laterally shifted corridor chosen by h, pushed out of obstacles.
"""
import numpy as np

WEIGHTS = dict(acceleration=0.34, angular_velocity=0.85, velocity=0.55, reference_velocity=2.0,
               contour=0.05, lag=0.75, terminal_angle=100.0, terminal_contouring=10.0,
               consistency_weight=0.05)
ROBOT_RADIUS = 0.325
OBSTACLE_RADIUS = 0.325
DECELERATION = 3.0
NUM_SEGMENTS = 5
_LATERAL = np.array([0.0, 1.0, -1.0, 2.0, -2.0, 3.0, -3.0, 0.5])


def _natural_cubic(t, y):
    """Batched natural cubic spline. t: (B,n) strictly increasing knots, y: (B,n).
    Returns per-segment coefficients a,b,c,d (B,n-1) of a*s^3+b*s^2+c*s+d, s = t - t_i."""
    B, n = t.shape
    h = np.diff(t, axis=1)
    A = np.zeros((B, n, n))
    rhs = np.zeros((B, n))
    A[:, 0, 0] = 1.0
    A[:, -1, -1] = 1.0
    idx = np.arange(1, n - 1)
    A[:, idx, idx - 1] = h[:, :-1]
    A[:, idx, idx] = 2.0 * (h[:, :-1] + h[:, 1:])
    A[:, idx, idx + 1] = h[:, 1:]
    rhs[:, 1:-1] = 3.0 * ((y[:, 2:] - y[:, 1:-1]) / h[:, 1:] - (y[:, 1:-1] - y[:, :-2]) / h[:, :-1])
    c = np.linalg.solve(A, rhs[..., None])[..., 0]
    b = (y[:, 1:] - y[:, :-1]) / h - h * (2.0 * c[:, :-1] + c[:, 1:]) / 3.0
    a = (c[:, 1:] - c[:, :-1]) / (3.0 * h)
    return a, c[:, :-1], b, y[:, :-1]


def make_batch(pmap, dims, n_sets, planners_per_set=1, seed=1234, guided=None, gaussian=False):
    """Build `n_sets * planners_per_set` problems.

    pmap: parameter name -> index (parameter_map.yaml); dims: dict(N, nx, nu, npar, dt).
    guided: if None, inferred from the presence of `lin_constraint_0_a1` in pmap.
    Problems of a set share state, path and obstacles and differ in warm start + halfspaces
    (planner index h < planners_per_set-1: guided; the last one is the non-guided planner when
    planners_per_set > 1), matching GuidanceConstraints (guidance_constraints.cpp:304-370).
    Returns dict of float64 arrays + `set_offsets`.
    """
    N, nx, nu, npar, dt = dims["N"], dims["nx"], dims["nu"], dims["npar"], dims["dt"]
    nz = nx + nu
    rng = np.random.default_rng(seed)
    S, Pn = n_sets, planners_per_set
    B = S * Pn
    has_lin = "lin_constraint_0_a1" in pmap
    if guided is None:
        guided = has_lin
    M = sum(1 for k in pmap if k.startswith("ellipsoid_obst_") and k.endswith("_x"))
    Mg = sum(1 for k in pmap if k.startswith("gaussian_obst_") and k.endswith("_x"))
    Md = sum(1 for k in pmap if k.startswith("disc_0_decomp_") and k.endswith("_a1"))
    Ml = sum(1 for k in pmap if k.startswith("disc_0_lin_constraint_") and k.endswith("_a1"))      # LinearizedConstraintModule rows
    Ml_dyn = max(Ml - 2, 0)                      # the last two rows play static halfspaces (road boundaries)
    Mall = max(M, Mg, Ml_dyn)
    has_spline = nx > 4                          # ContouringSecondOrderUnicycleModel; the plain unicycle has no path state

    # ---- per-set scenario ---------------------------------------------------------------------
    st = np.stack([rng.uniform(-0.5, 0.5, S), rng.uniform(-0.5, 0.5, S), rng.uniform(-0.3, 0.3, S),
                   rng.uniform(0.0, 2.5, S), np.zeros(S)], axis=1)
    nw = NUM_SEGMENTS + 2
    wx = np.concatenate([np.zeros((S, 1)), np.cumsum(rng.uniform(3.0, 6.0, (S, nw - 1)), axis=1)], axis=1)
    wy = np.concatenate([np.zeros((S, 1)), rng.uniform(-2.0, 2.0, (S, nw - 1))], axis=1)
    ts = np.concatenate([np.zeros((S, 1)), np.cumsum(np.hypot(np.diff(wx, axis=1), np.diff(wy, axis=1)), axis=1)], axis=1)
    ax, bx, cx, dx = _natural_cubic(ts, wx)
    ay, by, cy, dy = _natural_cubic(ts, wy)
    op0 = np.stack([rng.uniform(2.0, 20.0, (S, Mall)), rng.uniform(-4.0, 4.0, (S, Mall))], axis=2)
    ospeed = rng.uniform(0.5, 1.5, (S, Mall))
    ohead = rng.uniform(-np.pi, np.pi, (S, Mall))
    ovel = np.stack([ospeed * np.cos(ohead), ospeed * np.sin(ohead)], axis=2)
    steps = np.arange(N)[None, :, None, None]
    opred = op0[:, None] + ovel[:, None] * dt * steps          # (S, N, M, 2): prediction index i

    # ---- warm starts --------------------------------------------------------------------------
    x0 = np.zeros((S, Pn, N + 1, nz))
    kk = np.arange(N + 1)[None, :]
    for h in range(Pn):
        nonguided = (Pn > 1 and h == Pn - 1) and guided
        X = np.zeros((S, N + 1, nz))
        if guided and not nonguided:
            vg = np.clip(st[:, 3:4] * 0 + WEIGHTS["reference_velocity"], 0.5, 2.5)
            lat = _LATERAL[h % len(_LATERAL)]
            px = st[:, 0:1] + vg * dt * kk
            py = st[:, 1:2] + lat * np.sin(np.pi * kk / N) ** 2
            pos = np.stack([px, py], axis=2)                    # (S, N+1, 2)
            # push out of obstacles (stand-in for projectToSafety, linearized_constraints.cpp:130-148)
            for _ in range(3):
                for j in range(Mall):
                    o = np.concatenate([opred[:, :1, j], opred[:, :, j]], axis=1)  # stage k uses k-1
                    dvec = pos - o
                    dist = np.linalg.norm(dvec, axis=2, keepdims=True)
                    need = ROBOT_RADIUS + OBSTACLE_RADIUS + 0.1
                    push = np.where(dist < need, (need - dist) / np.maximum(dist, 1e-9), 0.0)
                    pos = pos + dvec * push
            pos[:, 0] = st[:, 0:2]
            vel = np.gradient(pos, dt, axis=1)
            X[:, :, nu + 0] = pos[:, :, 0]
            X[:, :, nu + 1] = pos[:, :, 1]
            X[:, :, nu + 2] = np.arctan2(vel[:, :, 1], vel[:, :, 0])
            X[:, :, nu + 3] = np.linalg.norm(vel, axis=2)
            if has_spline:
                X[:, :, nu + 4] = np.concatenate([np.zeros((S, 1)), np.cumsum(X[:, :-1, nu + 3] * dt, axis=1)], axis=1)
            X[:, 0, nu:] = st[:, :nx]
        else:
            # braking roll-out for the non-guided planner; constant-velocity cruise otherwise
            a = -DECELERATION if nonguided else 0.0
            x, y, psi, v, s = (st[:, i].copy() for i in range(5))
            for k in range(N + 1):
                X[:, k, 0] = a
                X[:, k, nu:] = np.stack([x, y, psi, v, s], axis=1)[:, :nx]
                x = x + v * dt * np.cos(psi)
                y = y + v * dt * np.sin(psi)
                s = s + v * dt
                v = np.maximum(v + a * dt, 0.0)
        x0[:, h] = X

    # ---- parameters ---------------------------------------------------------------------------
    P = np.zeros((S, Pn, N, npar))

    def setp(name, val):
        if name in pmap:
            P[..., pmap[name]] = val

    for name, val in WEIGHTS.items():
        setp(name, val)
    for i in range(NUM_SEGMENTS):
        for nm, arr in (("spline_x%d_a", ax), ("spline_x%d_b", bx), ("spline_x%d_c", cx), ("spline_x%d_d", dx),
                        ("spline_y%d_a", ay), ("spline_y%d_b", by), ("spline_y%d_c", cy), ("spline_y%d_d", dy)):
            setp(nm % i, arr[:, i][:, None, None])
        setp("spline%d_start" % i, ts[:, i][:, None, None])
    setp("ego_disc_radius", ROBOT_RADIUS)
    setp("ego_disc_0_offset", 0.0)
    if "goal_x" in pmap:        # GoalModule (goal_module.py:22-36): the third waypoint of the synthetic path; weights.goal of settings.yaml:79
        setp("goal_weight", 1.0)
        setp("goal_x", wx[:, 3][:, None, None])
        setp("goal_y", wy[:, 3][:, None, None])
    if "prev_traj_x" in pmap:  # consistency reference: the planner's own warm start positions
        P[..., pmap["prev_traj_x"]] = x0[:, :, :N, nu + 0]
        P[..., pmap["prev_traj_y"]] = x0[:, :, :N, nu + 1]

    dummy_xy = st[:, None, 0:2] + 50.0
    for j in range(M):
        pre = "ellipsoid_obst_%d_" % j
        P[:, :, 1:, pmap[pre + "x"]] = opred[:, None, :N - 1, j, 0]
        P[:, :, 1:, pmap[pre + "y"]] = opred[:, None, :N - 1, j, 1]
        P[:, :, 1:, pmap[pre + "r"]] = OBSTACLE_RADIUS
        P[:, :, 0, pmap[pre + "x"]] = dummy_xy[:, :, 0]
        P[:, :, 0, pmap[pre + "y"]] = dummy_xy[:, :, 1]
        P[:, :, 0, pmap[pre + "r"]] = 0.1
        P[..., pmap[pre + "chi"]] = 1.0
    for j in range(Mg):
        pre = "gaussian_obst_%d_" % j
        # propagatePredictionUncertainty (data_preparation.cpp:175-191): sigma grows with sqrt of the step
        sig = np.sqrt(np.cumsum(np.full(N, (0.3 * dt) ** 2)))
        P[:, :, 1:, pmap[pre + "x"]] = opred[:, None, :N - 1, j, 0]
        P[:, :, 1:, pmap[pre + "y"]] = opred[:, None, :N - 1, j, 1]
        P[:, :, 1:, pmap[pre + "major"]] = sig[None, None, :N - 1]
        P[:, :, 1:, pmap[pre + "minor"]] = sig[None, None, :N - 1]
        P[:, :, 1:, pmap[pre + "r"]] = OBSTACLE_RADIUS
        P[:, :, 0, pmap[pre + "x"]] = dummy_xy[:, :, 0]
        P[:, :, 0, pmap[pre + "y"]] = dummy_xy[:, :, 1]
        P[:, :, 0, pmap[pre + "major"]] = 0.1
        P[:, :, 0, pmap[pre + "minor"]] = 0.1
        P[:, :, 0, pmap[pre + "r"]] = 0.1
        P[..., pmap[pre + "risk"]] = 0.05
    if Md:
        # supporting halfspaces of a regular polygon of half-width 6 m around the reference position
        ang = 2.0 * np.pi * np.arange(Md) / Md
        cxk = st[:, None, 0:1] + WEIGHTS["reference_velocity"] * dt * np.arange(N)[None, :, None]   # (S,N,1)
        cyk = st[:, None, 1:2] + 0.0 * cxk
        for j in range(Md):
            a1, a2 = np.cos(ang[j]), np.sin(ang[j])
            P[..., pmap["disc_0_decomp_%d_a1" % j]] = a1
            P[..., pmap["disc_0_decomp_%d_a2" % j]] = a2
            P[..., pmap["disc_0_decomp_%d_b" % j]] = (a1 * cxk + a2 * cyk + 6.0)[:, None, :, 0]
    if has_lin:
        nlin = sum(1 for k in pmap if k.startswith("lin_constraint_") and k.endswith("_a1"))
        dummy_b = st[:, 0] + 100.0
        for j in range(nlin):
            pre = "lin_constraint_%d_" % j
            P[..., pmap[pre + "a1"]] = 1.0
            P[..., pmap[pre + "a2"]] = 0.0
            P[..., pmap[pre + "b"]] = dummy_b[:, None, None]
        for h in range(Pn):
            nonguided = (Pn > 1 and h == Pn - 1)
            if nonguided or not guided:
                continue
            pos = x0[:, h, 1:N, nu:nu + 2]                       # stages 1..N-1
            for j in range(min(nlin, M)):
                o = opred[:, :N - 1, j]                          # stage k <- prediction k-1
                dvec = o - pos
                dist = np.maximum(np.linalg.norm(dvec, axis=2), 1e-9)
                a1 = dvec[..., 0] / dist
                a2 = dvec[..., 1] / dist
                b = a1 * o[..., 0] + a2 * o[..., 1] - (1e-3 + ROBOT_RADIUS)
                pre = "lin_constraint_%d_" % j
                P[:, h, 1:, pmap[pre + "a1"]] = a1
                P[:, h, 1:, pmap[pre + "a2"]] = a2
                P[:, h, 1:, pmap[pre + "b"]] = b

    if Ml:
        # LinearizedConstraints::update / setParameters with _use_guidance = false (linearized_constraints.cpp:49-189): halfspaces
        # from the warm-start DISC position towards obstacle j's prediction k-1, radius = obstacle + robot radius; stage 0 and
        # unused slots: dummies (1, 0, x + 100).  The last two rows stand in for module_data.static_obstacles (:107-127):
        # road boundaries |y - y0| <= 3.5.
        dummy_b = st[:, 0] + 100.0
        for j in range(Ml):
            pre = "disc_0_lin_constraint_%d_" % j
            P[..., pmap[pre + "a1"]] = 1.0
            P[..., pmap[pre + "a2"]] = 0.0
            P[..., pmap[pre + "b"]] = dummy_b[:, None, None]
        for h in range(Pn):
            pos = x0[:, h, 1:N, nu:nu + 2]
            for j in range(Ml_dyn):
                o = opred[:, :N - 1, j]
                dvec = o - pos
                dist = np.maximum(np.linalg.norm(dvec, axis=2), 1e-9)
                a1, a2 = dvec[..., 0] / dist, dvec[..., 1] / dist
                pre = "disc_0_lin_constraint_%d_" % j
                P[:, h, 1:, pmap[pre + "a1"]] = a1
                P[:, h, 1:, pmap[pre + "a2"]] = a2
                P[:, h, 1:, pmap[pre + "b"]] = a1 * o[..., 0] + a2 * o[..., 1] - (OBSTACLE_RADIUS + ROBOT_RADIUS)
            for j, sgn in zip(range(Ml_dyn, Ml), (1.0, -1.0)):
                pre = "disc_0_lin_constraint_%d_" % j
                P[:, h, 1:, pmap[pre + "a1"]] = 0.0
                P[:, h, 1:, pmap[pre + "a2"]] = sgn
                P[:, h, 1:, pmap[pre + "b"]] = (sgn * st[:, 1] + 3.5)[:, None]

    xinit = np.repeat(st[:, None, :nx], Pn, axis=1)
    return dict(
        xinit=np.ascontiguousarray(xinit.reshape(B, nx)),
        x0=np.ascontiguousarray(x0.reshape(B, (N + 1) * nz)),
        params=np.ascontiguousarray(P.reshape(B, N * npar)),
        set_offsets=np.arange(0, B + 1, Pn, dtype=np.int32),
        n=B,
        # inputs of the device-side halfspace construction (mpcgpu_solve_sets_guided): per-set obstacle predictions
        # (prediction index i = stage k-1) and which planners follow a guidance trajectory
        obst_pred=np.ascontiguousarray(opred[:, :, :M] if has_lin else opred[:, :, :0]),
        guided=np.ascontiguousarray(np.tile(np.array([1 if (guided and not (Pn > 1 and h == Pn - 1)) else 0 for h in range(Pn)],
                                                      np.uint8), S)),
        robot_radius=ROBOT_RADIUS,
    )


def make_multi_robot_batch(pmap, dims, n_scenarios, robots=3, planners_per_set=9, seed=1234, peer_spacing=1.7):
    """BASELINE.json configs[3]: R robots x (8 guided + 1 non-guided) planners per scenario, one batch.

    In the reference "inter-robot constraints" are the ordinary ellipsoid constraints with a peer's communicated
    trajectory as the obstacle prediction (trajectoryCallback -> prepareObstacleData,
    mpc_planner_jackalsimulator/src/jules_ros1_jackalplanner.cpp:521-678,800-834; data_preparation.cpp:202-237:
    type ROBOT, radius = robot radius).  Here every robot of a scenario is one homotopy set of `make_batch`; the first
    R-1 obstacle slots of robot r are overwritten with the peers' trajectories (their planner-0 warm starts, running
    side by side `peer_spacing` apart), and the guidance halfspaces are rebuilt from the new predictions exactly as
    linearized_constraints.cpp:49-189 does (numpy, no projection: see tests for the projected variant).
    Problem index = (scenario * robots + robot) * planners + planner; `set_offsets` delimits one set per robot."""
    N, nx, nu, npar, dt = dims["N"], dims["nx"], dims["nu"], dims["npar"], dims["dt"]
    nz = nx + nu
    S, Rn, Pn = n_scenarios, robots, planners_per_set
    b = make_batch(pmap, dims, S * Rn, Pn, seed=seed)
    M = b["obst_pred"].shape[2]
    assert M >= Rn - 1, "configuration has too few obstacle slots for the peers"
    x0 = b["x0"].reshape(S, Rn, Pn, N + 1, nz)
    P = b["params"].reshape(S, Rn, Pn, N, npar)
    ob = b["obst_pred"].reshape(S, Rn, N, M, 2)
    for r in range(Rn):
        slot = 0
        for q in range(Rn):
            if q == r:
                continue
            # peer q's communicated plan in robot r's frame: its planner-0 warm start, shifted sideways
            traj = x0[:, q, 0, 1:, nu:nu + 2].copy()                       # (S, N, 2): positions of stages 1..N
            traj = traj - x0[:, q, 0, :1, nu:nu + 2] + x0[:, r, 0, :1, nu:nu + 2]
            traj[..., 1] += peer_spacing * (q - r)
            ob[:, r, :, slot] = traj                                       # prediction index i = stage i+1
            pre = "ellipsoid_obst_%d_" % slot
            P[:, r, :, 1:, pmap[pre + "x"]] = traj[:, None, :N - 1, 0]
            P[:, r, :, 1:, pmap[pre + "y"]] = traj[:, None, :N - 1, 1]
            P[:, r, :, 1:, pmap[pre + "r"]] = ROBOT_RADIUS
            slot += 1
    # guidance halfspaces from the new predictions (linearized_constraints.cpp:84-105)
    if "lin_constraint_0_a1" in pmap:
        nlin = sum(1 for k in pmap if k.startswith("lin_constraint_") and k.endswith("_a1"))
        for h in range(Pn - 1 if Pn > 1 else Pn):
            pos = x0[:, :, h, 1:N, nu:nu + 2]                              # (S, R, N-1, 2)
            for j in range(min(nlin, M)):
                o = ob[:, :, :N - 1, j]
                dvec = o - pos
                dist = np.sqrt(dvec[..., 0] * dvec[..., 0] + dvec[..., 1] * dvec[..., 1])
                a1, a2 = dvec[..., 0] / dist, dvec[..., 1] / dist
                pre = "lin_constraint_%d_" % j
                P[:, :, h, 1:, pmap[pre + "a1"]] = a1
                P[:, :, h, 1:, pmap[pre + "a2"]] = a2
                P[:, :, h, 1:, pmap[pre + "b"]] = a1 * o[..., 0] + a2 * o[..., 1] - (1e-3 + ROBOT_RADIUS)
    b["params"] = np.ascontiguousarray(P.reshape(-1, N * npar))
    b["obst_pred"] = np.ascontiguousarray(ob.reshape(S * Rn, N, M, 2))
    b["robots"] = Rn
    return b


def apply_consistency(batch, pmap, dims, planners, prev_traj, enabled, weight):
    """Load the consistency parameters the way GuidanceConstraints::setConsistencyParametersForPlanner does
    (mpc_planner_modules/src/guidance_constraints.cpp:985-1023): ONE interpolated previous trajectory per homotopy set
    (prev_traj [n_sets, N, 2]); for a planner with has_consistency_enabled the stages 1..N-2 get (weight, X_k, Y_k), every
    other stage and every other planner gets (0, 0, 0).  Modifies batch["params"] in place and returns the per-planner
    flags as uint8 [n]."""
    N, npar = dims["N"], dims["npar"]
    n = batch["n"]
    n_sets = n // planners
    P = batch["params"].reshape(n_sets, planners, N, npar)
    en = np.ascontiguousarray(enabled, np.uint8).reshape(n_sets, planners)
    valid = np.zeros(N, bool)
    valid[1:N - 1] = True
    act = (en[:, :, None] != 0) & valid[None, None, :]
    P[..., pmap["consistency_weight"]] = np.where(act, weight, 0.0)
    P[..., pmap["prev_traj_x"]] = np.where(act, prev_traj[:, None, :, 0], 0.0)
    P[..., pmap["prev_traj_y"]] = np.where(act, prev_traj[:, None, :, 1], 0.0)
    return en.reshape(-1)


# ---- counter-based generator: host mirror of csrc/mpcgpu_synth.cu --------------------------------------------
# A set's data depend on (seed, global set index) only: Philox4x32-10, counter = (set index lo, hi, draw block, 0),
# key = (seed lo, hi).  The arithmetic below is, operation for operation, the one of the device kernel (which is compiled
# without FMA contraction); sin / cos / arctan2 are the only functions whose last bit may differ between the two.
_M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al. 2011) on uint32 arrays (held in uint64 for the 32x32 -> 64 products)."""
    c0, c1, c2, c3 = (np.asarray(c, np.uint64) & _M32 for c in (c0, c1, c2, c3))
    k0 = np.uint64(k0) & _M32
    k1 = np.uint64(k1) & _M32
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ k0
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ k1
        c0, c1, c2, c3 = n0, p1 & _M32, n2, p0 & _M32
        k0 = (k0 + np.uint64(0x9E3779B9)) & _M32
        k1 = (k1 + np.uint64(0xBB67AE85)) & _M32
    return c0, c1, c2, c3


def philox_uniform(seed, gs, j):
    """j-th uniform double in [0, 1) of every global set index in `gs` (53 bits, as the device's draw())."""
    gs = np.asarray(gs, np.uint64)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    r = philox4x32_10(gs & _M32, gs >> np.uint64(32), np.full(gs.shape, j >> 1, np.uint64), np.zeros(gs.shape, np.uint64),
                      seed & 0xFFFFFFFF, seed >> 32)
    a, b = r[2 * (j & 1)], r[2 * (j & 1) + 1]
    return ((a >> np.uint64(5)).astype(np.float64) * 67108864.0 + (b >> np.uint64(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)


def synth_layout(pmap, dims, guided=None):
    """Parameter indices + constants of the counter-based generator for one configuration (the fields of
    mpcgpu_synth_layout, include/mpcgpu.h).  Supported: base weights, contouring spline, goal, consistency reference, ellipsoid
    obstacles and guidance halfspaces -- i.e. c1_basic, tmpc_shipped, c2_tmpc12, c6_goal_unicycle; configurations with
    Gaussian / decomp / per-disc linearised constraints raise ValueError (use make_batch)."""
    for pre in ("gaussian_obst_", "disc_0_decomp_", "disc_0_lin_constraint_"):
        if any(k.startswith(pre) for k in pmap):
            raise ValueError("counter-based generator: constraint family %s* is not supported" % pre)
    N = dims["N"]
    g = lambda name: int(pmap.get(name, -1))
    M = sum(1 for k in pmap if k.startswith("ellipsoid_obst_") and k.endswith("_x"))
    nlin = sum(1 for k in pmap if k.startswith("lin_constraint_") and k.endswith("_a1"))
    if M > 16 or nlin > 16 or N + 1 > 64:
        raise ValueError("counter-based generator: at most 16 obstacles / halfspaces and N <= 63")
    kk = np.arange(64, dtype=np.float64)
    return dict(
        N=N, nx=dims["nx"], nu=dims["nu"], npar=dims["npar"], guided=int((nlin > 0) if guided is None else bool(guided)),
        weights=[g(n) for n in WEIGHTS],
        spline=[[g(t % i) for t in ("spline_x%d_a", "spline_x%d_b", "spline_x%d_c", "spline_x%d_d", "spline_y%d_a", "spline_y%d_b",
                                    "spline_y%d_c", "spline_y%d_d", "spline%d_start")] for i in range(NUM_SEGMENTS)],
        ego_disc_radius=g("ego_disc_radius"), ego_disc_0_offset=g("ego_disc_0_offset"),
        goal=[g("goal_weight"), g("goal_x"), g("goal_y")], prev_traj_x=g("prev_traj_x"), prev_traj_y=g("prev_traj_y"),
        n_obst=M, obst=[[g("ellipsoid_obst_%d_%s" % (j, f)) for f in ("x", "y", "psi", "r", "major", "minor", "chi")] for j in range(M)],
        n_lin=nlin, lin=[[g("lin_constraint_%d_%s" % (j, f)) for f in ("a1", "a2", "b")] for j in range(nlin)],
        dt=float(dims["dt"]), pi=float(np.pi), vg=float(np.clip(WEIGHTS["reference_velocity"], 0.5, 2.5)),
        need=ROBOT_RADIUS + OBSTACLE_RADIUS + 0.1, deceleration=DECELERATION, robot_radius=ROBOT_RADIUS,
        obstacle_radius=OBSTACLE_RADIUS, lin_margin=1e-3 + ROBOT_RADIUS,
        weight_values=[float(v) for v in WEIGHTS.values()], lateral=[float(v) for v in _LATERAL],
        lat_profile=[float(v) for v in np.where(kk <= N, np.sin(np.pi * kk / N) ** 2, 0.0)],
    )


def _natural_cubic_thomas(t, y):
    """Natural cubic spline, Thomas algorithm in the device kernel's operation order.  t, y: (S, n)."""
    n = t.shape[1]
    h = t[:, 1:] - t[:, :-1]
    cp = np.zeros_like(t)
    dp = np.zeros_like(t)
    for i in range(1, n - 1):
        r = 3.0 * ((y[:, i + 1] - y[:, i]) / h[:, i] - (y[:, i] - y[:, i - 1]) / h[:, i - 1])
        den = 2.0 * (h[:, i - 1] + h[:, i]) - h[:, i - 1] * cp[:, i - 1]
        cp[:, i] = h[:, i] / den
        dp[:, i] = (r - h[:, i - 1] * dp[:, i - 1]) / den
    c = np.zeros_like(t)
    for i in range(n - 2, 0, -1):
        c[:, i] = dp[:, i] - cp[:, i] * c[:, i + 1]
    b = (y[:, 1:] - y[:, :-1]) / h - h * (2.0 * c[:, :-1] + c[:, 1:]) / 3.0
    a = (c[:, 1:] - c[:, :-1]) / (3.0 * h)
    return a, c[:, :-1], b, y[:, :-1]


def make_batch_philox(pmap, dims, n_sets, planners_per_set=1, seed=1234, first_set=0, guided=None):
    """The scenario definition of make_batch on the counter-based generator: homotopy sets first_set .. first_set + n_sets - 1
    of the global batch of `seed`.  Same return value as make_batch."""
    L = synth_layout(pmap, dims, guided)
    N, nx, nu, npar, dt = L["N"], L["nx"], L["nu"], L["npar"], L["dt"]
    nz, S, Pn, M = nx + nu, n_sets, planners_per_set, L["n_obst"]
    gs = np.arange(first_set, first_set + S, dtype=np.uint64)
    U = lambda lo, hi, j: lo + (hi - lo) * philox_uniform(seed, gs, j)
    st = np.stack([U(-0.5, 0.5, 0), U(-0.5, 0.5, 1), U(-0.3, 0.3, 2), U(0.0, 2.5, 3), np.zeros(S)], axis=1)
    wx, wy, ts = np.zeros((S, 7)), np.zeros((S, 7)), np.zeros((S, 7))
    for i in range(1, 7):
        wx[:, i] = wx[:, i - 1] + U(3.0, 6.0, 3 + i)
        wy[:, i] = U(-2.0, 2.0, 9 + i)
    for i in range(1, 7):
        ddx, ddy = wx[:, i] - wx[:, i - 1], wy[:, i] - wy[:, i - 1]
        ts[:, i] = ts[:, i - 1] + np.sqrt(ddx * ddx + ddy * ddy)
    spx = _natural_cubic_thomas(ts, wx)
    spy = _natural_cubic_thomas(ts, wy)
    op0 = np.zeros((S, M, 2))
    ovel = np.zeros((S, M, 2))
    for j in range(M):
        op0[:, j, 0], op0[:, j, 1] = U(2.0, 20.0, 16 + 4 * j), U(-4.0, 4.0, 17 + 4 * j)
        sp, hd = U(0.5, 1.5, 18 + 4 * j), U(-L["pi"], L["pi"], 19 + 4 * j)
        ovel[:, j, 0], ovel[:, j, 1] = sp * np.cos(hd), sp * np.sin(hd)
    steps = np.arange(N, dtype=np.float64)[None, :, None, None]
    opred = op0[:, None] + ovel[:, None] * dt * steps                      # (S, N, M, 2)
    ostage = np.concatenate([opred[:, :1], opred], axis=1)                 # stage k <- prediction max(k - 1, 0)

    x0 = np.zeros((S, Pn, N + 1, nz))
    kk = np.arange(N + 1, dtype=np.float64)[None, :]
    follow = [bool(L["guided"]) and not (Pn > 1 and h == Pn - 1) for h in range(Pn)]
    for h in range(Pn):
        X = np.zeros((S, N + 1, nz))
        if follow[h]:
            px = st[:, 0:1] + L["vg"] * dt * kk
            py = st[:, 1:2] + L["lateral"][h % 8] * np.asarray(L["lat_profile"][:N + 1])[None, :]
            for _ in range(3):
                for j in range(M):
                    ddx, ddy = px - ostage[:, :, j, 0], py - ostage[:, :, j, 1]
                    dist = np.sqrt(ddx * ddx + ddy * ddy)
                    push = np.where(dist < L["need"], (L["need"] - dist) / np.maximum(dist, 1e-9), 0.0)
                    px, py = px + ddx * push, py + ddy * push
            px[:, 0], py[:, 0] = st[:, 0], st[:, 1]
            vel = np.zeros((S, N + 1, 2))
            for c, q in enumerate((px, py)):
                vel[:, 1:N, c] = (q[:, 2:] - q[:, :-2]) / (2.0 * dt)
                vel[:, 0, c] = (q[:, 1] - q[:, 0]) / dt
                vel[:, N, c] = (q[:, N] - q[:, N - 1]) / dt
            X[:, :, nu + 0], X[:, :, nu + 1] = px, py
            X[:, :, nu + 2] = np.arctan2(vel[:, :, 1], vel[:, :, 0])
            X[:, :, nu + 3] = np.sqrt(vel[:, :, 0] * vel[:, :, 0] + vel[:, :, 1] * vel[:, :, 1])
            if nx > 4:
                sacc = np.zeros(S)
                for q in range(1, N + 1):
                    sacc = sacc + X[:, q - 1, nu + 3] * dt
                    X[:, q, nu + 4] = sacc
            X[:, 0, nu:] = st[:, :nx]
        else:
            a = -L["deceleration"] if (L["guided"] and Pn > 1 and h == Pn - 1) else 0.0
            x, y, psi, v, s = (st[:, i].copy() for i in range(5))
            for q in range(N + 1):
                X[:, q, 0] = a
                X[:, q, nu:] = np.stack([x, y, psi, v, s], axis=1)[:, :nx]
                x = x + v * dt * np.cos(psi)
                y = y + v * dt * np.sin(psi)
                s = s + v * dt
                v = np.maximum(v + a * dt, 0.0)
        x0[:, h] = X

    P = np.zeros((S, Pn, N, npar))

    def setp(idx, val):
        if idx >= 0:
            P[..., idx] = val

    for idx, val in zip(L["weights"], L["weight_values"]):
        setp(idx, val)
    for i in range(NUM_SEGMENTS):
        for c in range(4):
            setp(L["spline"][i][c], spx[c][:, i][:, None, None])
            setp(L["spline"][i][4 + c], spy[c][:, i][:, None, None])
        setp(L["spline"][i][8], ts[:, i][:, None, None])
    setp(L["ego_disc_radius"], L["robot_radius"])
    setp(L["ego_disc_0_offset"], 0.0)
    setp(L["goal"][0], 1.0)
    setp(L["goal"][1], wx[:, 3][:, None, None])
    setp(L["goal"][2], wy[:, 3][:, None, None])
    setp(L["prev_traj_x"], x0[:, :, :N, nu + 0])
    setp(L["prev_traj_y"], x0[:, :, :N, nu + 1])
    for j in range(M):
        ix, iy, _, ir, _, _, ichi = L["obst"][j]
        P[:, :, 1:, ix] = opred[:, None, :N - 1, j, 0]
        P[:, :, 1:, iy] = opred[:, None, :N - 1, j, 1]
        P[:, :, 1:, ir] = L["obstacle_radius"]
        P[:, :, 0, ix] = (st[:, 0] + 50.0)[:, None]
        P[:, :, 0, iy] = (st[:, 1] + 50.0)[:, None]
        P[:, :, 0, ir] = 0.1
        P[..., ichi] = 1.0
    for j in range(L["n_lin"]):
        ia1, ia2, ib = L["lin"][j]
        P[..., ia1], P[..., ia2] = 1.0, 0.0
        P[..., ib] = (st[:, 0] + 100.0)[:, None, None]
        if j >= M:
            continue
        for h in range(Pn):
            if not follow[h]:
                continue
            ox, oy = opred[:, :N - 1, j, 0], opred[:, :N - 1, j, 1]
            ddx, ddy = ox - x0[:, h, 1:N, nu + 0], oy - x0[:, h, 1:N, nu + 1]
            dist = np.maximum(np.sqrt(ddx * ddx + ddy * ddy), 1e-9)
            a1, a2 = ddx / dist, ddy / dist
            P[:, h, 1:, ia1], P[:, h, 1:, ia2] = a1, a2
            P[:, h, 1:, ib] = a1 * ox + a2 * oy - L["lin_margin"]
    B = S * Pn
    return dict(
        xinit=np.ascontiguousarray(np.repeat(st[:, None, :nx], Pn, axis=1).reshape(B, nx)),
        x0=np.ascontiguousarray(x0.reshape(B, (N + 1) * nz)),
        params=np.ascontiguousarray(P.reshape(B, N * npar)),
        set_offsets=np.arange(0, B + 1, Pn, dtype=np.int32), n=B,
        obst_pred=np.ascontiguousarray(opred if L["n_lin"] else opred[:, :, :0]),
        guided=np.ascontiguousarray(np.tile(np.array([1 if f else 0 for f in follow], np.uint8), S)),
        robot_radius=ROBOT_RADIUS,
    )
