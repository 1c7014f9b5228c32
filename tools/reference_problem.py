"""Load the REFERENCE's own symbolic problem definition (generation-time tooling, container only).

Imports the unmodified reference scripts from /root/reference through the sympy-backed `casadi`
stand-in (oscar_mpc_planner_mr_modification_b200/solver_generator/casadi_compat) and returns plain sympy expressions for the dynamics, the stage cost and
the constraints of a named configuration.  /root/reference does not exist on the GPU box, so nothing
here is used at run time: the oracle model code (oracle/generated/*.c) and the golden vectors
(tests/golden/*.npz) are produced from it ONCE and committed.

Configurations restate `mpc_planner_jackalsimulator/scripts/generate_jackalsimulator_solver.py:37-116`
(that script runs the acados generator and calls exit() at import, so it cannot be imported).
"""
import os
import sys

import numpy as np
import sympy as sp
import yaml

REFERENCE_ROOT = os.environ.get("MPC_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "solver_generator"))


def _prepare_imports():
    shim = os.path.join(_HERE, "..", "oscar_mpc_planner_mr_modification_b200", "solver_generator", "casadi_compat")
    gen = os.path.join(REFERENCE_ROOT, "solver_generator")
    mods = os.path.join(REFERENCE_ROOT, "mpc_planner_modules", "scripts")
    for p in (mods, gen, shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not hasattr(np, "Inf"):  # gaussian_constraints.py:65 uses np.Inf (removed in numpy 2)
        np.Inf = np.inf


CONFIGS = {
    # name: (configuration, max_obstacles, N, consistency)
    "c1_basic": dict(kind="basic", max_obstacles=4, N=30),
    "tmpc_shipped": dict(kind="tmpc_consistency", max_obstacles=4, N=30),
    "c2_tmpc12": dict(kind="tmpc", max_obstacles=12, N=30),
    "c5_ccmpc": dict(kind="ccmpc_decomp", max_obstacles=4, N=50),
    # north_star's second dynamics model and the goal module: SecondOrderUnicycleModel (nx = 4, solver_model.py:170-190) with
    # base weights + GoalModule (goal_module.py:22-36) + ellipsoid obstacles -- configuration_lmpcc
    # (generate_jackalsimulator_solver.py:118-134) without the path-velocity module, which needs the spline state
    "c6_goal_unicycle": dict(kind="goal_unicycle", max_obstacles=4, N=30),
    # LinearizedConstraintModule (linearized_constraints.py:17-95; an option of generate_rosnavigation_solver.py:57):
    # MPCC + 6 halfspaces a1 x_disc + a2 y_disc <= b per stage on the disc position
    "c7_linearized": dict(kind="linearized", max_obstacles=6, N=30),
}


def load_settings(max_obstacles, N):
    path = os.path.join(REFERENCE_ROOT, "mpc_planner_jackalsimulator", "config", "settings.yaml")
    with open(path) as f:
        settings = yaml.safe_load(f)
    settings["max_obstacles"] = max_obstacles
    settings["N"] = N
    return settings


def build_modules(config_name):
    """(modules, model, settings) built from the reference's own module classes -- the three objects
    the reference passes to generate_acados_solver()."""
    _prepare_imports()
    import contextlib
    import io

    cfg = CONFIGS[config_name]
    settings = load_settings(cfg["max_obstacles"], cfg["N"])

    with contextlib.redirect_stdout(io.StringIO()):
        from control_modules import ModuleManager
        from mpc_base import MPCBaseModule
        from contouring import ContouringModule
        from consistency_module import ConsistencyModule
        from ellipsoid_constraints import EllipsoidConstraintModule
        from gaussian_constraints import GaussianConstraintModule
        from guidance_constraints import GuidanceConstraintModule
        from decomp_constraints import DecompConstraintModule
        from goal_module import GoalModule
        from linearized_constraints import LinearizedConstraintModule
        from solver_model import ContouringSecondOrderUnicycleModel, SecondOrderUnicycleModel

    kind = cfg["kind"]
    modules = ModuleManager()
    if kind == "goal_unicycle":  # generate_jackalsimulator_solver.py:118-134 on the plain unicycle (solver_model.py:170-190)
        model = SecondOrderUnicycleModel()
        base = modules.add_module(MPCBaseModule(settings))
        base.weigh_variable(var_name="a", weight_names="acceleration")
        base.weigh_variable(var_name="w", weight_names="angular_velocity")
        modules.add_module(GoalModule(settings))
        modules.add_module(EllipsoidConstraintModule(settings))
        return modules, model, settings

    # generate_jackalsimulator_solver.py:37-58 (configuration_no_obstacles)
    model = ContouringSecondOrderUnicycleModel()
    base = modules.add_module(MPCBaseModule(settings))
    base.weigh_variable(var_name="a", weight_names="acceleration")
    base.weigh_variable(var_name="w", weight_names="angular_velocity")
    base.weigh_variable(var_name="v", weight_names=["velocity", "reference_velocity"],
                        cost_function=lambda x, w: w[0] * (x - w[1]) ** 2)
    modules.add_module(ContouringModule(settings))

    if kind == "linearized":  # generate_rosnavigation_solver.py:57 (commented alternative): LinearizedConstraintModule
        modules.add_module(LinearizedConstraintModule(settings))
    elif kind == "basic":  # :61-67
        modules.add_module(EllipsoidConstraintModule(settings))
    elif kind == "tmpc":  # :95-105
        modules.add_module(GuidanceConstraintModule(settings, constraint_submodule=EllipsoidConstraintModule))
    elif kind == "tmpc_consistency":  # :107-116
        modules.add_module(ConsistencyModule(settings))
        modules.add_module(GuidanceConstraintModule(settings, constraint_submodule=EllipsoidConstraintModule))
    elif kind == "ccmpc_decomp":  # BASELINE config 5: CC-MPC + decomp polytopes (synthetic combination)
        modules.add_module(GaussianConstraintModule(settings))
        modules.add_module(DecompConstraintModule(settings))
    else:
        raise KeyError(kind)
    return modules, model, settings


def build(config_name):
    """Returns dict with sympy symbols/expressions taken from the reference scripts."""
    import contextlib
    import io

    cfg = CONFIGS[config_name]
    modules, model, settings = build_modules(config_name)
    with contextlib.redirect_stdout(io.StringIO()):
        from solver_definition import (define_parameters, objective, constraints,
                                       constraint_lower_bounds, constraint_upper_bounds)
        from util.parameters import AcadosParameters
        import casadi as cd

    # generate_acados_solver.py:68-75, 27-65
    with contextlib.redirect_stdout(io.StringIO()):
        params = AcadosParameters()
        define_parameters(modules, params, settings)
        params.load_acados_parameters()
        settings["params"] = params
        z = model.acados_symbolics()
        f_expl, _ = model.get_acados_dynamics()
        p = params.get_acados_p()
        h = constraints(modules, z, p, model, settings, 1)
        cost = objective(modules, z, p, model, settings, 1)
        lh = list(constraint_lower_bounds(modules))
        uh = list(constraint_upper_bounds(modules))

    names = list(params._params.keys())
    z_syms = cd.to_sympy(z)
    p_syms = [cd.to_sympy(q)[0] for q in p]
    h_exprs = []
    for c in h:
        h_exprs.extend(cd.to_sympy(c))
    big = 1e15  # generate_acados_solver.py:17-24
    lh = [(-big if v == -np.inf else (big if v == np.inf else float(v))) for v in lh]
    uh = [(-big if v == -np.inf else (big if v == np.inf else float(v))) for v in uh]
    return dict(
        name=config_name, N=cfg["N"], dt=float(settings["integrator_step"]),
        nx=model.nx, nu=model.nu, states=list(model.states), inputs=list(model.inputs),
        lb=[float(v) for v in model.lower_bound], ub=[float(v) for v in model.upper_bound],
        z=z_syms, p=p_syms, param_names=names, bundles=dict(params.parameter_bundles),
        f=cd.to_sympy(f_expl), cost=cd.to_sympy(cost)[0], h=h_exprs, lh=lh, uh=uh,
        settings=settings,
    )


if __name__ == "__main__":
    for name in CONFIGS:
        pb = build(name)
        print(name, "npar", len(pb["p"]), "nh", len(pb["h"]), "lh", pb["lh"][:2], "uh", pb["uh"][-2:])
