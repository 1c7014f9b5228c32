#!/usr/bin/env python3
"""FP64 operations the solve kernel actually EXECUTES, from an ncu --set full capture ("no executed-FLOP figure is
recorded beside the algorithmic one").  Reads smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on (thread-level
instruction counts per elapsed cycle, summed over the GPU) x elapsed cycles; FMA = 2 flops.
Usage: ncu_executed_flops.py report.ncu-rep <key e.g. c2_tmpc12/iter10> <solves in the profiled launch> <mean IPM iterations of that batch>
Appends to profiles/executed_flops.json (read by bench.py: roofline.executed)."""
import csv
import io
import json
import os
import subprocess
import sys

rep, key, n_solves, ipm = sys.argv[1], sys.argv[2], int(sys.argv[3]), float(sys.argv[4])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                                  text=True).stdout)))
m = dict(zip(raw[0], raw[2]))
cyc = float(m["smsp__cycles_elapsed.avg"])
ops = {k: float(m["smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % k]) * cyc for k in ("dfma", "dadd", "dmul")}
flops = 2.0 * ops["dfma"] + ops["dadd"] + ops["dmul"]
out_path = os.path.join(ROOT, "profiles", "executed_flops.json")
try:
    out = json.load(open(out_path))
except Exception:
    out = {}
out[key] = {"source": "ncu thread-instruction counts of %s" % os.path.basename(rep), "solves_in_launch": n_solves, "ipm_iters_mean": ipm,
            "thread_instructions": ops, "executed_flops_per_solve": flops / n_solves, "executed_flops_per_ipm_iter": flops / n_solves / ipm,
            "kernel_ms": float(m["gpu__time_duration.sum"]) if "gpu__time_duration.sum" in m else None,
            "fp64_pipe_busy_pct": float(m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]),
            "dram_bytes_per_solve": (float(m["dram__bytes_read.sum"]) + float(m["dram__bytes_write.sum"])) * (1e9 if "Gbyte" in raw[1][raw[0].index("dram__bytes_read.sum")] else 1.0) / n_solves}
json.dump(out, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(out[key], indent=1))
