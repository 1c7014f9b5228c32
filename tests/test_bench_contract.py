"""bench.py contract: the CPU reference arm prints exactly ONE JSON line on stdout with the keys the driver reads; the
GPU arm does the same on a GPU box (marked gpu) and carries roofline / cpu_baseline / e2e / clocks / gpu_launches."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config", "e2e", "cpu_baseline"}


def run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, "stdout must carry exactly one line, got %d: %r" % (len(lines), r.stdout[:300])
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line():
    d = run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-sets", "4"], 300)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "MPC solves/sec (N=30 SQP-RTI batch)" and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["dtype"] == "f64" and d["vs_baseline"] is None


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_line():
    d = run(["--steps", "2", "--warmup", "3", "--sets", "512", "--latency-reps", "10", "--cpu-sets", "8"], 600)
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["scaling"] == "weak" and d["value"] > 1e4
    assert d["gpu_launches"] >= 2 * d["steps"]
    rf = d["roofline"]
    assert rf["bound"] == "fp64" and 0.0 < rf["frac"] < 1.0 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and rf["traffic"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and 0 < d["e2e"]["value"] <= 1.15 * d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["latency"]["p50"] > 0 and d["e2e_sets"]["guided"]["value"] > 0
    # SURVEY 8(d) beside the headline: one iteration of the same workload, the other configurations, executed FLOPs, one CPU policy
    ex = d["extra"]
    assert ex["c2_tmpc12/iter1"]["solves_per_s_per_gpu"] > d["value"] and 0 < ex["c2_tmpc12/iter1"]["fp64_frac"] < 1
    assert ex["c2_tmpc12/iter1_warm"]["success_frac"] > 0.5 > ex["c2_tmpc12/iter1"]["success_frac"]
    for k in ("c1_basic/iter10", "tmpc_shipped/iter10", "c5_ccmpc/iter10", "c6_goal_unicycle/iter10", "c7_linearized/iter1"):
        assert ex[k]["solves_per_s_per_gpu"] > 0 and 0 < ex[k]["fp64_frac"] < 1, k
    assert 0 < rf["executed"]["frac"] < rf["frac"]
    assert "all host cores" in d["cpu_baseline"]["cores_policy"] and d["config"]["resident_batches"] == 2
    assert d["config"]["generator"].startswith("device") and d["config"]["generation_s"] < 5.0
