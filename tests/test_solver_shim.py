"""The C++ host shim (MPCPlanner::Solver on the engine): compiled test binary mirroring the
reference's mpc_planner_solver/test/test_solver.cpp."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "oscar_mpc_planner_mr_modification_b200", "host")
SETTINGS = os.path.join(HOST, "test", "settings.yaml")


def _build():
    subprocess.run(["make", "-C", HOST], check=True, stdout=subprocess.DEVNULL)


@pytest.mark.parametrize("cfg", ["tmpc_shipped", "c2_tmpc12"])
def test_shim_host_side(cfg):
    _build()
    r = subprocess.run([os.path.join(HOST, "build", "test_solver_shim_" + cfg), SETTINGS, "--no-gpu"], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["tmpc_shipped", "c2_tmpc12"])
def test_shim_solve_and_batch(cfg):
    _build()
    r = subprocess.run([os.path.join(HOST, "build", "test_solver_shim_" + cfg), SETTINGS], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout
