"""The -DMPC_CHECK=1 diagnostic build (lib/libmpcgpu_check.so, benchmark configuration): bounds-checked shared-memory accessors,
canary words between the shared-memory regions of both solve kernels, labelled rendezvous in the role-split kernel.  It stands
in for compute-sanitizer (memcheck / racecheck), which this pool's GPUs refuse: the parity
batches run under it with every counter at zero, and the detector's self-test shows that it does detect."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oscar_mpc_planner_mr_modification_b200", "lib", "libmpcgpu_check.so")

_CHILD = r"""
import ctypes, sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
from oscar_mpc_planner_mr_modification_b200 import engine, synthetic
from oracle_binding import Oracle
lib = engine.load_library()
lib.mpcgpu_check_report.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
lib.mpcgpu_check_selftest.argtypes = [ctypes.c_void_p]
eng = engine.Engine("c2_tmpc12", 0, 2048); orc = Oracle("c2_tmpc12")
cnt = np.zeros(8, np.uint64)
def report(reset=1):
    assert lib.mpcgpu_check_report(eng.handle, cnt.ctypes.data, reset) == 0
    return cnt.copy()
report()
# the detector detects: one bad index, one overwritten canary
assert lib.mpcgpu_check_selftest(eng.handle) == 0
c = report()
assert c[0] == 1 and c[1] == 1, c
# parity batches under the checker: both kernels, small and larger batches, with capsule memory and the set entry
b = synthetic.make_batch(eng.parameter_map, eng.dims, 40, 9, seed=11)
ref = orc.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=10)
total = 0
for mode in (engine.KERNEL_STAGE, engine.KERNEL_SPLIT):
    eng.set_kernel_mode(mode)
    for sl in (slice(0, 9), slice(0, 360)):
        mem = np.zeros((sl.stop - sl.start, eng.mem_doubles))
        for rep in range(2):
            out = eng.solve_batch(b["xinit"][sl], b["x0"][sl], b["params"][sl], num_iter=10, mem=mem if rep else None)
            total += sl.stop - sl.start
        first = eng.solve_batch(b["xinit"][sl], b["x0"][sl], b["params"][sl], num_iter=10)
        total += sl.stop - sl.start
        assert (first["exit_code"] == ref["exit_code"][sl]).all()
        ok = ref["exit_code"][sl] == 1
        assert np.abs(first["xtraj"][ok] - ref["xtraj"][sl][ok]).max() < 1e-6 * max(1.0, np.abs(ref["xtraj"][sl][ok]).max())
eng.set_kernel_mode(engine.KERNEL_AUTO)
xs = np.ascontiguousarray(b["xinit"].reshape(40, 9, -1)[:, 0])
shared = np.ascontiguousarray(b["params"].reshape(40, 9, eng.N, eng.npar)[:, 0])
eng.solve_sets_guided(40, 9, xs, shared, b["x0"], b["obst_pred"], b["guided"], b["robot_radius"], num_iter=10)
total += 360
c = report()
assert c[0] == 0 and c[1] == 0 and c[2] == 0 and c[3] == 0, c
assert c[4] == total, (c, total)          # every problem went through the checks
print("CHECK-OK", c.tolist())
"""


def test_check_library_is_built_and_normal_library_refuses_the_report():
    import ctypes
    from oscar_mpc_planner_mr_modification_b200 import engine
    assert os.path.exists(LIB), "lib/libmpcgpu_check.so is not built (python __graft_entry__.py)"
    lib = engine.load_library()
    assert lib.mpcgpu_check_report(None, None, 0) == -1


@pytest.mark.gpu
def test_parity_batches_run_clean_under_the_check_build():
    env = dict(os.environ, MPCGPU_LIB=LIB)
    r = subprocess.run([sys.executable, "-c", _CHILD % (ROOT, os.path.join(ROOT, "tests"))], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                       text=True, timeout=900)
    assert r.returncode == 0 and "CHECK-OK" in r.stdout, r.stdout[-3000:]
