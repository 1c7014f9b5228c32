#!/usr/bin/env python3
"""Optional cross-check against the GENUINE acados-generated solver (VERDICT r01 #8a).

The reference's arithmetic (acados + HPIPM + BLASFEO + CasADi) is not vendored and not installable in the build
container, which is why the oracle's solver level is "parity unpinned" (DESIGN.md section 5).  On a machine that HAS
casadi and acados_template (ACADOS_SOURCE_DIR set, t_renderer built: README.md:226-245 of the reference) this script

  1. copies the reference's solver_generator/ and mpc_planner_modules/scripts/ to a scratch directory (the generator
     writes its C code next to itself; nothing is copied into this repository),
  2. builds (modules, model, settings) exactly as tools/reference_problem.py does -- but with the REAL casadi -- and calls
     the reference's own generate_acados_solver(modules, settings, model, False) (generate_acados_solver.py:68-200),
  3. replays Solver::solve() (mpc_planner_solver/src/acados_solver_interface.cpp:86-204) on the seeded inputs of
     tests/golden/solve_<cfg>.npz through the AcadosOcpSolver handle: x0 bounds = xinit (:124-125), parameters per stage with
     stage N reusing N-1 (:128-134), warm start (:274-284), rti_phase 0, `num_iter` calls of solve() with the wrapper's
     early exit on qp_status != 0 (:99-106), cost (:167), res_eq rule (:176-181), exit-code map (:197-203),
  4. writes tests/golden/acados_<cfg>.npz.  tests/test_acados_crosscheck.py then compares the oracle with it (exit flags
     bit-exact, x / u to the tolerance north_star states) and is skipped while the file is absent.

Without casadi / acados_template the script prints why and exits 0: it is never required.
Usage: ACADOS_SOURCE_DIR=... python tools/acados_crosscheck.py [config ...]
UNTESTED in the build container (acados absent there): the replay follows the reference's C++ wrapper line by line."""
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")


def available():
    try:
        import casadi  # noqa: F401
        if "casadi_compat" in os.path.abspath(getattr(casadi, "__file__", "")):
            return "only the sympy-backed casadi stand-in is importable"
        import acados_template  # noqa: F401
    except Exception as ex:      # noqa: BLE001
        return "%s: %s" % (type(ex).__name__, ex)
    if not os.environ.get("ACADOS_SOURCE_DIR"):
        return "ACADOS_SOURCE_DIR is not set"
    return None


def build_reference_solver(cfg, scratch):
    """the reference's generator on the reference's modules, real casadi + acados"""
    ref = os.environ.get("MPC_REFERENCE_ROOT", "/root/reference")
    gen = os.path.join(scratch, "solver_generator")
    mods = os.path.join(scratch, "mpc_planner_modules", "scripts")
    shutil.copytree(os.path.join(ref, "solver_generator"), gen)
    shutil.copytree(os.path.join(ref, "mpc_planner_modules", "scripts"), mods)
    os.makedirs(os.path.join(scratch, "mpc_planner_solver"), exist_ok=True)
    sys.path[:0] = [mods, gen]
    if not hasattr(np, "Inf"):
        np.Inf = np.inf
    sys.path.insert(0, HERE)
    import reference_problem as rp
    rp._prepare_imports = lambda: None          # keep the REAL casadi: do not put the stand-in on sys.path
    modules, model, settings = rp.build_modules(cfg)
    from generate_acados_solver import generate_acados_solver
    cwd = os.getcwd()
    os.chdir(scratch)
    try:
        solver, _sim = generate_acados_solver(modules, settings, model, False)
    finally:
        os.chdir(cwd)
    return solver, model, settings


def replay_solve(solver, N, nx, nu, npar, xinit, x0, params, num_iter):
    """Solver::solve() through the Python handle of the generated solver"""
    nz = nx + nu
    solver.reset()                                                   # a fresh capsule per problem, like the golden inputs assume
    solver.set(0, "lbx", xinit); solver.set(0, "ubx", xinit)          # :124-125
    P = params.reshape(N, npar)
    for k in range(N + 1):
        solver.set(k, "p", P[min(k, N - 1)])                         # :128-134
    X = x0.reshape(N + 1, nz)
    for k in range(N + 1):                                           # loadWarmstart :274-284
        solver.set(k, "x", X[k, nu:])
        if k < N:
            solver.set(k, "u", X[k, :nu])
    solver.options_set("rti_phase", 0)                               # :139-141
    status, qp_status = 0, 0
    for _ in range(num_iter):                                        # :99
        status = solver.solve()                                      # :149
        qp_status = int(np.atleast_1d(solver.get_stats("qp_stat"))[-1])     # :155
        if qp_status != 0:                                           # :105-106
            break
    pobj = float(solver.get_cost())                                  # :167-168
    xt = np.concatenate([solver.get(k, "x") for k in range(N + 1)])  # :171-174
    ut = np.concatenate([solver.get(k, "u") for k in range(N)])
    res = solver.get_residuals()                                     # [res_stat, res_eq, res_ineq, res_comp]
    res_eq = float(res[1])
    if res_eq > 1e-2 and status == 0:                                # :176-181
        status = 4
    exit_code = 1 if status == 0 else (0 if status == 1 else status)  # :197-203
    return xt, ut, pobj, exit_code, qp_status, res_eq


def main():
    why = available()
    if why:
        print("acados cross-check unavailable: %s -- nothing written (the check is optional)" % why)
        return 0
    sys.path.insert(0, ROOT)
    from oscar_mpc_planner_mr_modification_b200 import engine, synthetic
    for cfg in (sys.argv[1:] or ["c1_basic", "tmpc_shipped", "c2_tmpc12"]):
        gd = np.load(os.path.join(GOLD, "solve_%s.npz" % cfg))
        pmap, _mmap, st = engine.load_maps(cfg)
        dims = dict(N=st["N"], nx=st["nx"], nu=st["nu"], npar=st["npar"], dt=0.2)
        b = synthetic.make_batch(pmap, dims, int(gd["n_sets"]), int(gd["planners"]), seed=int(gd["seed"]))
        scratch = tempfile.mkdtemp(prefix="acados_crosscheck_")
        solver, _model, _settings = build_reference_solver(cfg, scratch)
        out = {}
        for nit in (1, 10):
            rows = [replay_solve(solver, dims["N"], dims["nx"], dims["nu"], dims["npar"], b["xinit"][i], b["x0"][i], b["params"][i], nit)
                    for i in range(b["n"])]
            for j, key in enumerate(("xtraj", "utraj", "pobj", "exit_code", "qp_status", "res_eq")):
                out["%s_it%d" % (key, nit)] = np.array([r[j] for r in rows])
        np.savez_compressed(os.path.join(GOLD, "acados_%s.npz" % cfg), seed=int(gd["seed"]), n_sets=int(gd["n_sets"]), planners=int(gd["planners"]), **out)
        print("wrote tests/golden/acados_%s.npz (%d problems)" % (cfg, b["n"]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
