#!/usr/bin/env python3
"""Attribute the SASS-level counters of an .ncu-rep (thread-per-stage solve kernel) to the CALL SITE inside solve_problem():
inlined helpers (ineq_final, step_limit, the emitted model functions ...) are charged to the line of solve_problem that called
them, which the plain source page cannot do.  Joins `ncu --page source --csv` (per SASS instruction, in address order) with
`nvdisasm --print-line-info-inline` of the same cubin (same instruction order).
Usage: ncu_callsite.py report.ncu-rep cfg_object.o [kernel-symbol-substring]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj = sys.argv[1], sys.argv[2]
want = sys.argv[3] if len(sys.argv) > 3 else "mpc_solve_kernelILi8"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
kern = os.path.join(ROOT, "oscar_mpc_planner_mr_modification_b200/csrc/mpc_solve_kernel.cuh")
src = open(kern).read().splitlines()
lo = next(i + 1 for i, l in enumerate(src) if l.startswith("__device__ void solve_problem("))
hi = next(i + 1 for i, l in enumerate(src) if i + 1 > lo and l.startswith("}"))
marks = [(i + 1, l.strip()) for i, l in enumerate(src) if lo <= i + 1 <= hi and (l.strip().startswith("// ----") or l.strip().startswith("// ===="))]
marks = [(lo, "// prologue: load iterate / capsule memory")] + marks + [(hi + 1, "end")]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info-inline", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                     text=True).stdout.splitlines()
insts, chain, active = [], [], False
for l in dis:
    if l.startswith(".text."):
        active = want in l
        chain = []
        continue
    if not active:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        if "inlined at" in m.group(3) or not chain or chain_closed:
            if not ("inlined at" in m.group(3)) and chain and not chain_closed:
                pass
        # a block of consecutive //## lines forms one chain (innermost first)
        if not chain or chain_closed:
            chain, chain_closed = [], False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        chain_closed = True
        insts.append((int(m.group(1), 16), m.group(2).strip(), list(chain)))

rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                                   text=True).stdout)))
data = [r for r in rows[2:] if len(r) >= 64]
assert len(data) == len(insts), (len(data), len(insts))
f = lambda x: float(x) if x not in ("", "-") else 0.0
agg = collections.defaultdict(lambda: [0.0] * 6)      # samples, exec, thread-exec, long_sb, wait, short_sb
tot_s = sum(f(r[4]) for r in data); tot_e = sum(f(r[5]) for r in data)
for r, (off, txt, ch) in zip(data, insts):
    site = None
    for fn, ln in ch:                                   # innermost first: the first frame inside solve_problem is the call site
        if fn == "mpc_solve_kernel.cuh" and lo <= ln <= hi:
            site = ln
            break
    a = agg[site]
    a[0] += f(r[4]); a[1] += f(r[5]); a[2] += f(r[6]); a[3] += f(r[35]); a[4] += f(r[46]); a[5] += f(r[43])
print("total samples %.0f, executed warp-instructions %.3g" % (tot_s, tot_e))
print("== regions of solve_problem (inlined helpers charged to their call site)")
for (a0, name), (a1, _) in zip(marks[:-1], marks[1:]):
    v = [sum(agg[k][i] for k in agg if k is not None and a0 <= k < a1) for i in range(6)]
    if v[1]:
        print("  %5.1f%% samp %5.1f%% exec  avgthr %4.1f  long_sb %4.1f%% wait %4.1f%% | %s" % (100 * v[0] / tot_s, 100 * v[1] / tot_e, v[2] / v[1],
                                                                                    100 * v[3] / tot_s, 100 * v[4] / tot_s, name[:90]))
v = agg[None]
print("  %5.1f%% samp %5.1f%% exec  (outside solve_problem: work loop, non-inlined functions)" % (100 * v[0] / tot_s, 100 * v[1] / tot_e))
print("== top call-site lines")
for k, v in sorted(((k, v) for k, v in agg.items() if k is not None), key=lambda kv: -kv[1][0])[:int(os.environ.get("NCU_TOP", "25"))]:
    print("  line %4d  %5.2f%% samp %5.2f%% exec avgthr %4.1f long_sb %4.1f%% | %s" % (k, 100 * v[0] / tot_s, 100 * v[1] / tot_e, v[2] / max(v[1], 1), 100 * v[3] / tot_s,
                                                                          src[k - 1].strip()[:100]))
