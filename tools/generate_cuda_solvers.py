#!/usr/bin/env python3
"""Run the CUDA emitter on the reference's own module definitions for every benchmark
configuration (container only: needs /root/reference).  The output directory
oscar_mpc_planner_mr_modification_b200/generated/<config>/ is committed."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, ".."))
import reference_problem as rp  # noqa: E402
from oscar_mpc_planner_mr_modification_b200.solver_generator.generate_cuda_solver import generate_cuda_solver  # noqa: E402

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(rp.CONFIGS)):
        modules, model, settings = rp.build_modules(name)
        out = os.path.join(HERE, "..", "oscar_mpc_planner_mr_modification_b200", "generated", name)
        pb = generate_cuda_solver(modules, settings, model, name, out)
        print("generated", name, "npar", len(pb["p"]), "nh", len(pb["h"]), "->", os.path.relpath(out))
