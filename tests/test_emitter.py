"""The CUDA emitter (solver_generator/generate_cuda_solver.py).  With the reference present (build
container) the committed generated files must be exactly what the emitter produces from the reference's
own module objects, and the reference's own generator tests must pass on top of the casadi stand-in.
Without it (GPU box) only the self-consistency checks run."""
import filecmp
import os
import subprocess
import sys

import pytest

from oracle_binding import Oracle
from oscar_mpc_planner_mr_modification_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import reference_problem as rp  # noqa: E402

needs_ref = pytest.mark.skipif(not rp.reference_available(), reason="/root/reference not present")
COMPAT = os.path.join(ROOT, "oscar_mpc_planner_mr_modification_b200", "solver_generator", "casadi_compat")


@pytest.mark.parametrize("cfg", ["c1_basic", "tmpc_shipped", "c2_tmpc12", "c5_ccmpc", "c6_goal_unicycle", "c7_linearized"])
def test_maps_agree_with_oracle_model(cfg):
    """Two independent extraction paths (reference solver_definition.py for the oracle, the emitter's own
    loops for the product) give the same parameter order and dimensions."""
    pmap, mmap, st = engine.load_maps(cfg)
    orc = Oracle(cfg)
    assert pmap == orc.parameter_map
    assert (st["N"], st["nx"], st["nu"], st["npar"]) == (orc.N, orc.nx, orc.nu, orc.npar)
    assert [n for n, _ in sorted(((k, v[1]) for k, v in mmap.items()), key=lambda t: t[1])] == orc.var_names
    hdr = open(os.path.join(engine.config_dir(cfg), "model.cuh")).read()
    assert "NH = %d" % orc.nh in hdr and "NCG = %d" % (orc.nc - 2 * orc.nz) in hdr
    par = open(os.path.join(engine.config_dir(cfg), "mpc_planner_parameters.h")).read()
    fns = ["setSolverParameterAcceleration"]
    fns += ["setSolverParameterSplineXA"] if "spline_x0_a" in pmap else ["setSolverParameterGoalX", "setSolverParameterGoalWeight"]
    fns += ["setSolverParameterEgoDiscRadius"] if "ego_disc_radius" in pmap else ["setSolverParameterEgoDiscOffset", "setSolverParameterLinConstraintA1"]
    for fn in fns:
        assert fn in par        # generate_cpp_files.py:235-254 naming rule


@needs_ref
@pytest.mark.parametrize("cfg", ["c1_basic", "c2_tmpc12"])
def test_committed_files_are_current(cfg, tmp_path):
    from oscar_mpc_planner_mr_modification_b200.solver_generator.generate_cuda_solver import generate_cuda_solver
    modules, model, settings = rp.build_modules(cfg)
    generate_cuda_solver(modules, settings, model, cfg, str(tmp_path))
    for f in ("model.cuh", "parameter_map.yaml", "model_map.yaml", "solver_settings.yaml", "mpc_planner_parameters.h"):
        assert filecmp.cmp(os.path.join(str(tmp_path), f), os.path.join(engine.config_dir(cfg), f), shallow=False), f


@needs_ref
def test_reference_generator_tests_pass_on_the_casadi_stand_in():
    """solver_generator/test/test_base_classes.py and test_control_modules.py of the reference, unmodified,
    with `casadi` resolved to the sympy-backed stand-in the emitter uses."""
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([COMPAT, os.path.join(ROOT, "tests", "ref_shims")])
    tests = [os.path.join(rp.REFERENCE_ROOT, "solver_generator", "test", t) for t in ("test_base_classes.py", "test_control_modules.py")]
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider", "--rootdir", "/tmp"] + tests, env=env,
                       cwd="/tmp", stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]
