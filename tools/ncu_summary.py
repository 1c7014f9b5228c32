#!/usr/bin/env python3
"""Summarise an .ncu-rep of mpc_solve_kernel: headline metrics, stall mix, executed-instruction share
per kernel region / model function, serial vs lane-parallel split.  Usage: ncu_summary.py rep [model.cuh]"""
import csv
import io
import os
import subprocess
import sys

rep = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
model = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "oscar_mpc_planner_mr_modification_b200/generated/c2_tmpc12/model.cuh")
kern = os.path.join(ROOT, "oscar_mpc_planner_mr_modification_b200/csrc/mpc_solve_kernel.cuh")


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep] + list(args), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
hdr, units, vals = raw[0], raw[1], raw[2]
m = dict(zip(hdr, zip(units, vals)))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
print("== headline")
for k in keys:
    if k in m:
        print("%-70s %s %s" % (k, m[k][1], m[k][0]))
print("== stall mix (samples)")
st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v[1]) for k, v in m.items()
      if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")}
tot = sum(st.values())
for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]:
    print("  %-22s %5.1f%%" % (k, 100 * v / tot))

rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--print-source", "cuda,sass", "--csv"))))
cur, hdr, agg = None, None, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) > 3 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 9 or r[2] != "-":
        continue
    try:
        agg[(cur, int(r[0]))] = (int(r[7]), int(r[8]), int(r[6]), r[1].strip()[:100])
    except ValueError:
        pass
tot = sum(v[0] for v in agg.values())
tots = sum(v[2] for v in agg.values())
ser = sum(v[0] for v in agg.values() if v[0] and v[1] / v[0] <= 2)
print("== executed warp-instructions: total %.3g, single-lane share %.1f%%" % (tot, 100 * ser / tot))
src = open(kern).read().splitlines()
marks = [(i + 1, l.strip()) for i, l in enumerate(src) if l.strip().startswith("// ----") or l.strip().startswith("// ====")]
marks.append((len(src) + 1, "end"))
print("== kernel regions")
for (a, name), (b, _) in zip(marks, marks[1:]):
    sel = [v for (f, ln), v in agg.items() if f == "mpc_solve_kernel.cuh" and a <= ln < b]
    e = sum(v[0] for v in sel)
    if e > 0.002 * tot:
        print("  %5.1f%% exec %5.1f%% samples  avgthr %4.1f | %s" % (100 * e / tot, 100 * sum(v[2] for v in sel) / tots, sum(v[1] for v in sel) / e, name[:80]))
msrc = open(model).read().splitlines()
fm = [(i + 1, l.split("(")[0].split()[-1]) for i, l in enumerate(msrc) if l.startswith("__device__ __forceinline__")]
fm.append((len(msrc) + 1, "end"))
print("== model.cuh functions")
for (a, name), (b, _) in zip(fm, fm[1:]):
    sel = [v for (f, ln), v in agg.items() if f == "model.cuh" and a <= ln < b]
    e = sum(v[0] for v in sel)
    if e > 0.002 * tot:
        print("  %5.1f%% exec %5.1f%% samples  avgthr %4.1f | %s" % (100 * e / tot, 100 * sum(v[2] for v in sel) / tots, sum(v[1] for v in sel) / e, name))
other = {}
for (f, ln), v in agg.items():
    if f not in ("mpc_solve_kernel.cuh", "model.cuh"):
        o = other.setdefault(f, [0, 0]); o[0] += v[0]; o[1] += v[2]
for f, (e, s_) in other.items():
    if e > 0.002 * tot:
        print("  %5.1f%% exec %5.1f%% samples | %s" % (100 * e / tot, 100 * s_ / tots, f))
print("== top source lines by stall samples")
for (f, ln), v in sorted(agg.items(), key=lambda x: -x[1][2])[:14]:
    print("  %-20s %4d exec %5.2f%% samp %5.2f%% | %s" % (f[:20], ln, 100 * v[0] / tot, 100 * v[2] / tots, v[3][:90]))
