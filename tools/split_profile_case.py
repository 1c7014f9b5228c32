#!/usr/bin/env python3
"""One launch of the role-split kernel on 16 homotopy sets (144 problems, one CTA each) -- the ncu target for the
latency path (the thread-per-stage kernel is profiled through bench.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oscar_mpc_planner_mr_modification_b200 import engine, synthetic  # noqa: E402

eng = engine.Engine("c2_tmpc12", 0, 256)
b = synthetic.make_batch(eng.parameter_map, eng.dims, 16, 9, seed=1234)
eng.set_kernel_mode(engine.KERNEL_SPLIT)
for _ in range(3):
    out = eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=10)
print("split kernel: n", b["n"], "kernel ms %.3f" % eng.last_kernel_ms(), "ok", float((out["exit_code"] == 1).mean()),
      "ipm mean", float(out["ipm_iters"].mean()))
