"""The C-ABI shared library: loads, exports every symbol include/mpcgpu.h declares, registers the
compiled configurations, and FAILS LOUDLY without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

from oscar_mpc_planner_mr_modification_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    syms = set()
    for h in sorted(os.listdir(os.path.join(ROOT, "include"))):      # every header of the C ABI (mpcgpu.h, mpcgpu_wire.h)
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        syms |= set(re.findall(r"\b(mpcgpu_[a-z0-9_]+)\s*\(", src))
    return sorted(syms)


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), "libmpcgpu.so does not export %s" % s
    assert set(engine.SYMBOLS) <= set(syms)


def test_registered_configurations_and_maps():
    lib = engine.load_library()
    names = [lib.mpcgpu_config_name(i).decode() for i in range(lib.mpcgpu_num_configs())]
    assert set(names) == {"c1_basic", "tmpc_shipped", "c2_tmpc12", "c5_ccmpc", "c6_goal_unicycle", "c7_linearized"}
    expect = {"c1_basic": (83, 4, 30), "tmpc_shipped": (98, 8, 30), "c2_tmpc12": (175, 24, 30), "c5_ccmpc": (115, 16, 50),
              "c6_goal_unicycle": (35, 4, 30),      # base 2 + goal 3 + ellipsoid 2 + 4 x 7 (goal_module.py:22-26, ellipsoid_constraints.py:41-49)
              "c7_linearized": (72, 6, 30)}         # 53 + disc offset 1 + 6 x 3 (linearized_constraints.py:41-49)
    for n in names:
        pmap, mmap, st = engine.load_maps(n)
        nx = 4 if n == "c6_goal_unicycle" else 5
        assert st == dict(N=expect[n][2], nx=nx, nu=2, nvar=nx + 2, npar=expect[n][0])
        assert len(pmap) == expect[n][0] and sorted(pmap.values()) == list(range(expect[n][0]))
        if nx == 5:      # ContouringSecondOrderUnicycleModel (solver_model.py:193-205)
            assert mmap["a"][:2] == ["u", 0] and mmap["spline"][:2] == ["x", 6] and mmap["v"][2:] == [-0.01, 3.0]
        else:            # SecondOrderUnicycleModel (solver_model.py:170-181): no spline state, its own bounds
            assert "spline" not in mmap and mmap["w"][2:] == [-2.0, 2.0] and mmap["v"] == ["x", 5, -2.0, 3.0] and mmap["x"][2:] == [-200.0, 200.0]
    pm = engine.load_maps("c2_tmpc12")[0]
    assert pm["acceleration"] == 0 and pm["spline_x0_a"] == 8 and pm["lin_constraint_0_a1"] == 53
    assert pm["ego_disc_radius"] == 89 and pm["ellipsoid_obst_0_x"] == 91 and pm["ellipsoid_obst_11_r"] == 174
    pm = engine.load_maps("tmpc_shipped")[0]
    assert pm["consistency_weight"] == 53 and pm["lin_constraint_0_a1"] == 56 and pm["ellipsoid_obst_0_x"] == 70


def test_bad_arguments_are_rejected():
    lib = engine.load_library()
    h = ctypes.c_void_p()
    assert lib.mpcgpu_engine_create(b"no_such_config", 0, 16, ctypes.byref(h)) == -1
    assert lib.mpcgpu_engine_create(b"c1_basic", 0, 0, ctypes.byref(h)) == -1
    assert lib.mpcgpu_engine_destroy(None) == -1


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(engine.MpcGpuError) as ei:
        engine.Engine("c1_basic", device=0, max_batch=8)
    assert "status -3" in str(ei.value)       # MPCGPU_ERR_NO_DEVICE: no silent CPU path
