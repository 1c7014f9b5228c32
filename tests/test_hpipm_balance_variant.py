"""The -DMPC_HPIPM_BALANCE=1 build (HPIPM's conditional Mehrotra predictor-corrector, which DESIGN.md section 4 leaves out of
the default algorithm contract): kernel variant (lib/libmpcgpu_balance.so, benchmark configuration) against the oracle built
with the same switch.  Exists so that the closer variant can be chosen the day tools/acados_crosscheck.py pins the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle_binding import Oracle
from oscar_mpc_planner_mr_modification_b200 import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oscar_mpc_planner_mr_modification_b200", "lib", "libmpcgpu_balance.so")


def test_balance_oracle_differs_from_the_default_contract_in_a_few_percent_of_the_iterations():
    cfg = "c2_tmpc12"
    plain, bal = Oracle(cfg), Oracle(cfg, "balance_")
    b = synthetic.make_batch(plain.parameter_map, plain.dims, 6, 9, seed=1234)
    r0 = plain.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=10)
    r1 = bal.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=10)
    changed = (r0["ipm_iters"] != r1["ipm_iters"]).mean()
    assert 0.0 < changed < 1.0                       # the fallback triggers (about 5 % of the interior-point iterations) ...
    both = (r0["exit_code"] == 1) & (r1["exit_code"] == 1)
    assert both.sum() >= 30                          # ... and most problems still converge, to (nearly) the same point after 10 iterations
    assert np.median(np.abs(r0["xtraj"][both] - r1["xtraj"][both]).max(axis=1)) < 1e-3


_CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from oscar_mpc_planner_mr_modification_b200 import engine, synthetic
from oracle_binding import Oracle
eng = engine.Engine("c2_tmpc12", 0, 1024); orc = Oracle("c2_tmpc12", "balance_")
assert eng.set_kernel_mode(0) is False               # no role-split kernel in this variant
b = synthetic.make_batch(eng.parameter_map, eng.dims, 40, 9, seed=5)
for nit in (1, 10):
    ref = orc.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=nit)
    out = eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=nit)
    ok = ref["exit_code"] == 1
    assert (out["exit_code"] == ref["exit_code"]).all()
    assert np.abs(out["xtraj"][ok] - ref["xtraj"][ok]).max() < 1e-6 * max(1.0, np.abs(ref["xtraj"][ok]).max())
print("BALANCE-OK")
"""


@pytest.mark.gpu
def test_balance_kernel_variant_matches_balance_oracle():
    assert os.path.exists(LIB), "lib/libmpcgpu_balance.so is not built (python __graft_entry__.py)"
    env = dict(os.environ, MPCGPU_LIB=LIB)
    r = subprocess.run([sys.executable, "-c", _CHILD % (ROOT, os.path.join(ROOT, "tests"))], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=600)
    assert r.returncode == 0 and "BALANCE-OK" in r.stdout, r.stdout[-2000:]
