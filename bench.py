#!/usr/bin/env python3
"""bench.py -- throughput of the batched MPC solve path (BASELINE.json metric) on N B200s.

A step = one pass of the hot path over one batch of synthetic homotopy sets: every problem runs
`num_iter` SQP-RTI iterations (one `Solver::solve()` of the reference,
mpc_planner_solver/src/acados_solver_interface.cpp:86-204) and the best planner of every set is
picked (guidance_constraints.cpp:572-590).  Workload = BASELINE.json configs[1]: T-MPC++ with
8 guided + 1 non-guided planner, 12 dynamic obstacles, N=30 (generated config `c2_tmpc12`).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                     CPU arm: the oracle port on the host cores

`value`   device-resident inputs, CUDA-event timed, max over ranks.
`e2e`     same metric through the C ABI with pinned HOST buffers (H2D + kernel + D2H inside).
`roofline`  FP64 pipe: algorithmic FLOPs (oracle/flops.json, counted by the oracle built
          with a counting scalar type) / measured solve-kernel time, against the DFMA peak measured
          here (MEASURED_PEAKS.json has no FP64 entry); HBM streaming reported beside it.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "MPC solves/sec (N=30 SQP-RTI batch)"
PLANNERS = {"c1_basic": 1, "tmpc_shipped": 5, "c2_tmpc12": 9, "c5_ccmpc": 1, "c6_goal_unicycle": 1, "c7_linearized": 1}


def load_json(path, default=None):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "50"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def mark(self):
        """number of samples written so far (call at the start of the timed region)"""
        try:
            with open(self.f.name) as fh:
                return sum(1 for _ in fh)
        except Exception:
            return 0

    def stop(self, first=0):
        """statistics of the samples taken after `first` (the timed region); a region shorter than the sampling period
        falls back to the samples just before it (same load: the warm-up steps)"""
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                rows.append((float(c[0]), float(c[1]), [nm for nm, v in zip(names, c[3:7]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        os.unlink(self.f.name)
        sel = rows[first:] if len(rows) > first else rows[-3:]
        if not sel:
            return None
        reasons = set()
        for r in sel:
            reasons.update(r[2])
        return {"sm_mhz": float(np.median([r[0] for r in sel])), "sm_max_mhz": float(max(r[1] for r in sel)), "reasons": sorted(reasons),
                "samples": len(sel), "in_timed_region": len(rows) > first}


def bytes_per_solve(d):
    """in 8(nx + nvar(N+1) + N npar) + out 8(nx(N+1) + nu N + 3) + 8"""
    nz = d["nx"] + d["nu"]
    return 8 * (d["nx"] + nz * (d["N"] + 1) + d["N"] * d["npar"]) + 8 * (d["nx"] * (d["N"] + 1) + d["nu"] * d["N"] + 3) + 8


def cpu_oracle_rate(cfg, planners, num_iter, n_sets, threads, seed=4321):
    from oracle_binding import Oracle
    from oscar_mpc_planner_mr_modification_b200 import synthetic
    orc = Oracle(cfg)
    b = synthetic.make_batch(orc.parameter_map, orc.dims, n_sets, planners, seed=seed)
    t0 = time.perf_counter()
    out = orc.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=num_iter, threads=threads)
    orc.select_best(b["set_offsets"], out["pobj"], out["exit_code"])
    dt = time.perf_counter() - t0
    return b["n"] / dt, dt, b["n"]


def run_reference(args):
    """CPU arm.  The reference's own implementation of the path (acados/HPIPM generated solver) cannot
    be built here (not vendored, no network): this times the oracle port -- the only other place
    bench.py executes oracle/ -- OpenMP over problems like guidance_constraints.cpp:304."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, planners = args.config, PLANNERS[args.config]
    threads = os.cpu_count() or 1
    n_sets = args.ref_sets
    for _ in range(min(args.warmup, 1)):
        cpu_oracle_rate(cfg, planners, args.num_iter, max(1, n_sets // 8), threads)
    rates, times = [], []
    for s in range(args.steps):
        r, dt, n = cpu_oracle_rate(cfg, planners, args.num_iter, n_sets, threads, seed=4321 + s)
        rates.append(r); times.append(dt)
    value = float(n * len(times) / sum(times))
    sample = "%d homotopy sets x %d planners per step (bounded sample of the %s workload), %d SQP-RTI iterations" % (
        n_sets, planners, cfg, args.num_iter)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(cfg, planners, args.num_iter), "sample": sample},
            "cpu_baseline": {"value": value, "unit": "solves/s", "cores": threads, "kind": "port", "sample": sample,
                             "cores_policy": "all host cores (os.cpu_count()), OpenMP over problems -- the same policy as cpu_baseline of the GPU arm"},
            "e2e": {"value": value, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit_json(line)


def workload_name(cfg, planners, num_iter):
    return "%s: T-MPC++ homotopy sets, %d planners/set, N=%d, dt=0.2, %d SQP-RTI iterations/solve" % (cfg, planners, 50 if cfg == "c5_ccmpc" else 30, num_iter)


_REAL_STDOUT = None


def emit_json(line):
    """the ONE JSON line of the contract, on the process's original stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # Libraries chat on stdout (NCCL prints its version banner there at NCCL_DEBUG=VERSION/WARN): everything written to fd 1
    # during the run goes to stderr, the JSON line goes to the saved descriptor.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2_tmpc12", choices=sorted(PLANNERS))
    ap.add_argument("--sets", type=int, default=4096, help="homotopy sets per GPU per step")
    ap.add_argument("--num-iter", type=int, default=10, help="SQP-RTI iterations per solve (settings.yaml:18)")
    ap.add_argument("--ref-sets", type=int, default=128, help="homotopy sets per step of the CPU arm")
    ap.add_argument("--cpu-sets", type=int, default=0, help="homotopy sets of the cpu_baseline sample (0: ~20 s of CPU work on all cores)")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra measurements (other iteration counts / configurations)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--latency-reps", type=int, default=100, help="single-set latency repetitions (0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)

    import torch
    from oscar_mpc_planner_mr_modification_b200 import engine, synthetic

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the solve path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    cfg, planners = args.config, PLANNERS[args.config]
    n_sets = args.sets
    n = n_sets * planners
    eng = engine.Engine(cfg, device=local_rank, max_batch=n)
    d = eng.dims
    N, nx, nu, npar = d["N"], d["nx"], d["nu"], d["npar"]
    nz = nx + nu

    # ---- synthetic batches of this rank's shard (independent homotopy sets; no data-path collective).  TWO resident batches are
    #      rotated between steps so that no step re-solves what the previous one left in L2 (each is larger than L2 anyway).
    #      The data come from the counter-based generator ON THE DEVICE (Philox4x32-10 keyed by (seed, global set index):
    #      rank r holds sets [r n_sets, (r+1) n_sets) of the global batch) and are copied back once for the host-array entries;
    #      configurations the generator does not cover fall back to the numpy generator.
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.float64 if a.dtype == np.float64 else torch.int32, pin_memory=True)
        t.numpy()[...] = a
        return t

    t_gen = time.perf_counter()
    try:
        M_syn = synthetic.synth_layout(eng.parameter_map, d)["n_obst"]
        generator = "device: Philox4x32-10 keyed by (seed, set index), mpcgpu_generate_synthetic_device"
    except ValueError:
        M_syn, generator = None, "host: numpy (synthetic.make_batch)"
    if M_syn is not None:
        d_in, d_obst = [], []
        for i in range(2):
            bufs = (torch.empty((n, nx), dtype=torch.float64, device=dev), torch.empty((n, (N + 1) * nz), dtype=torch.float64, device=dev),
                    torch.empty((n, N * npar), dtype=torch.float64, device=dev))
            ob = torch.empty((n_sets, N, M_syn, 2), dtype=torch.float64, device=dev)
            engine.generate_synthetic(eng.parameter_map, d, n_sets, planners, seed=1234 + 104729 * i, first_set=rank * n_sets,
                                      device=local_rank, device_buffers=[b.data_ptr() for b in bufs] + [ob.data_ptr()])
            d_in.append(bufs)
            d_obst.append(ob)
        torch.cuda.synchronize()
        t_gen = (time.perf_counter() - t_gen) / 2
        h_in = []
        for bufs in d_in:
            hb = tuple(torch.empty(b.shape, dtype=torch.float64, pin_memory=True) for b in bufs)
            for h_, b in zip(hb, bufs):
                h_.copy_(b)
            h_in.append(hb)
        has_lin = "lin_constraint_0_a1" in eng.parameter_map
        follow = [1 if (has_lin and not (planners > 1 and h == planners - 1)) else 0 for h in range(planners)]
        batches = [dict(xinit=h[0].numpy(), x0=h[1].numpy(), params=h[2].numpy(), n=n, set_offsets=np.arange(0, n + 1, planners, dtype=np.int32),
                        obst_pred=(ob.cpu().numpy() if has_lin else np.zeros((n_sets, N, 0, 2))), guided=np.tile(np.array(follow, np.uint8), n_sets),
                        robot_radius=synthetic.ROBOT_RADIUS) for h, ob in zip(h_in, d_obst)]
        del d_obst
    else:
        batches = [synthetic.make_batch(eng.parameter_map, d, n_sets, planners, seed=1234 + 7919 * rank + 104729 * i) for i in range(2)]
        t_gen = (time.perf_counter() - t_gen) / len(batches)
        h_in = [(pinned(b["xinit"]), pinned(b["x0"]), pinned(b["params"])) for b in batches]
        d_in = [tuple(t.to(dev) for t in h) for h in h_in]
    batch = batches[0]
    h_xinit, h_x0, h_params = h_in[0]
    h_offsets = pinned(batch["set_offsets"])
    d_offsets = h_offsets.to(dev)
    d_xtraj = torch.empty((n, (N + 1) * nx), dtype=torch.float64, device=dev)
    d_utraj = torch.empty((n, N * nu), dtype=torch.float64, device=dev)
    d_pobj = torch.empty(n, dtype=torch.float64, device=dev)
    d_res = torch.empty(n, dtype=torch.float64, device=dev)
    d_exit = torch.empty(n, dtype=torch.int32, device=dev)
    d_qps = torch.empty(n, dtype=torch.int32, device=dev)
    d_ipm = torch.empty(n, dtype=torch.int32, device=dev)
    d_best = torch.empty(n_sets, dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)      # a real (non-default) stream: the engine launches on it, events record on it
    torch.cuda.synchronize()
    step_no = [0]

    def step_device(which=None, num_iter=None):
        i = step_no[0] % len(d_in) if which is None else which
        step_no[0] += 1
        xi_, x0_, pr_ = d_in[i]
        eng.solve_batch_device(n, xi_.data_ptr(), x0_.data_ptr(), pr_.data_ptr(), args.num_iter if num_iter is None else num_iter,
                               d_xtraj.data_ptr(), d_utraj.data_ptr(), d_pobj.data_ptr(), d_exit.data_ptr(), d_qps.data_ptr(), d_res.data_ptr(),
                               ipm_iters=d_ipm.data_ptr(), stream=stream.cuda_stream)
        eng.select_best_device(n_sets, d_offsets.data_ptr(), d_pobj.data_ptr(), d_exit.data_ptr(), d_best.data_ptr(),
                               stream=stream.cuda_stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    fp64_peak = engine.measure_fp64_peak(local_rank)

    # ---- value: device-resident, CUDA events on the launching stream, max over ranks
    sampler = ClockSampler(local_rank)          # started before the warm-up so that a short timed region still has samples nearby
    for _ in range(args.warmup):
        step_device()
    barrier()
    clock_mark = sampler.mark()
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    assert stream.cuda_stream != 0
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    torch.cuda.synchronize()
    total_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    # solve-kernel duration (events recorded by the engine around the kernel, same stream): one more pass
    ipm_means = []
    for i in (1, 0, 1, 0):
        step_device(which=i)
        torch.cuda.synchronize()
        kernel_ms.append(eng.last_kernel_ms())
        ipm_means.append(float(d_ipm.float().mean().item()))
    barrier()
    clocks = sampler.stop(clock_mark)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * n * args.steps / (total_ms_max * 1e-3)

    exit_codes = d_exit.cpu().numpy()           # (batch 0: the last pass above)
    ipm_mean = float(np.mean(ipm_means))        # over both resident batches, like the kernel time
    best = d_best.cpu().numpy()

    # ---- e2e: the C-ABI call a user makes, pinned host buffers, H2D + kernel + D2H + select inside
    out = eng.alloc_outputs(n)
    h_out = {k: pinned(v) for k, v in out.items()}
    np_out = {k: v.numpy() for k, v in h_out.items()}
    xi_np, x0_np, p_np = h_xinit.numpy(), h_x0.numpy(), h_params.numpy()
    off_np = h_offsets.numpy()

    h_np = [tuple(t.numpy() for t in h) for h in h_in]

    def step_e2e(which):
        a, b_, c = h_np[which]
        eng.solve_batch(a, b_, c, num_iter=args.num_iter, out=np_out)
        return eng.select_best(off_np, np_out["pobj"], np_out["exit_code"])

    step_e2e(1)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e((i + 1) % 2)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    best_e2e = step_e2e(0)
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * n * args.steps / float(t.item())
    h2d = (h_xinit.numel() + h_x0.numel() + h_params.numel()) * 8 + (n_sets + 1) * 4 + n * 12
    d2h = (n * ((N + 1) * nx + N * nu + 2)) * 8 + n * 12 + n_sets * 4
    assert (best_e2e == best).all(), "device-resident and host-API paths disagree on the selected planners"

    # ---- e2e through the compact homotopy-set entry (mpcgpu_solve_sets): shared parameter block per set
    e2e_sets = None
    if planners > 1:
        Pv = p_np.reshape(n_sets, planners, N, npar)
        differs = np.nonzero((Pv[:64] != Pv[:64, :1]).any(axis=(0, 1, 2)))[0].astype(np.int32)
        h_shared = pinned(np.ascontiguousarray(Pv[:, 0]))
        h_vals = pinned(np.ascontiguousarray(Pv[..., differs]))
        h_xs = pinned(np.ascontiguousarray(xi_np.reshape(n_sets, planners, nx)[:, 0]))
        so = dict(np_out)
        so["best"] = pinned(np.zeros(n_sets, np.int32)).numpy()
        eng.solve_sets(n_sets, planners, h_xs.numpy(), h_shared.numpy(), x0_np, differs, h_vals.numpy(), num_iter=args.num_iter, out=so)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            eng.solve_sets(n_sets, planners, h_xs.numpy(), h_shared.numpy(), x0_np, differs, h_vals.numpy(), num_iter=args.num_iter, out=so)
        torch.cuda.synchronize()
        ts_ = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(ts_, op=dist.ReduceOp.MAX)
        assert (so["best"] == best).all()
        e2e_sets = {"value": world * n * args.steps / float(ts_.item()), "unit": "solves/s",
                    "h2d_bytes_per_step": int((h_shared.numel() + h_vals.numel() + h_xs.numel() + h_x0.numel()) * 8 + differs.size * 4),
                    "what": "mpcgpu_solve_sets: shared parameter block per set + %d per-planner parameters, selection fused" % differs.size}

        # ---- the same with the guidance halfspaces built on the device (mpcgpu_solve_sets_guided)
        lin_base, lin_count = eng.lin_constraint_block()
        if lin_count > 0 and set(differs.tolist()) <= set(range(lin_base, lin_base + 3 * lin_count)):
            h_ob, h_g = pinned(batch["obst_pred"]), torch.from_numpy(batch["guided"]).pin_memory()
            go = dict(np_out)
            go["best"] = pinned(np.zeros(n_sets, np.int32)).numpy()
            run = lambda: eng.solve_sets_guided(n_sets, planners, h_xs.numpy(), h_shared.numpy(), x0_np, h_ob.numpy(), h_g.numpy(),
                                                batch["robot_radius"], num_iter=args.num_iter, out=go)
            run()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                run()
            torch.cuda.synchronize()
            tg_ = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(tg_, op=dist.ReduceOp.MAX)
            assert (go["best"] == best).all()
            # ---- struct-of-tables (mpcgpu_solve_sets_tables): stage-invariant parameters once per set, obstacle table,
            #      warm starts; the parameter block is built on the device
            lay = eng.table_layout()
            inv_idx = lay["invariant_idx"]
            if lay.get("ellipsoid") and set(range(npar)) - set(inv_idx.tolist()) <= (set(range(lin_base, lin_base + 3 * lin_count)) |
                                                                                    set(range(lay["ellipsoid"]["base"], lay["ellipsoid"]["base"] + lay["ellipsoid"]["count"] * lay["ellipsoid"]["stride"]))):
                h_inv = pinned(np.ascontiguousarray(Pv[:, 0, 0][:, inv_idx]))
                h_rad = pinned(np.full((n_sets, batch["obst_pred"].shape[2]), synthetic.OBSTACLE_RADIUS))
                to = dict(np_out)
                to["best"] = pinned(np.zeros(n_sets, np.int32)).numpy()
                run_t = lambda: eng.solve_sets_tables(n_sets, planners, h_xs.numpy(), h_inv.numpy(), h_ob.numpy(), x0_np, guided=h_g.numpy(),
                                                      robot_radius=batch["robot_radius"], obstacle_radius=h_rad.numpy(), num_iter=args.num_iter, out=to)
                ot = run_t()
                barrier()
                t0 = time.perf_counter()
                for _ in range(args.steps):
                    run_t()
                torch.cuda.synchronize()
                tt_ = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
                if dist is not None:
                    dist.all_reduce(tt_, op=dist.ReduceOp.MAX)
                assert (to["best"] == best).all()
                e2e_sets["tables"] = {"value": world * n * args.steps / float(tt_.item()), "unit": "solves/s", "h2d_bytes_per_step": int(ot["h2d_bytes"]),
                                      "h2d_bytes_per_solve": ot["h2d_bytes"] / n,
                                      "what": "mpcgpu_solve_sets_tables: %d stage-invariant parameters once per set + obstacle table + warm starts; the [N][npar] "
                                              "block is built on the device (ellipsoid slots, guidance halfspaces)" % inv_idx.size}
            e2e_sets["guided"] = {"value": world * n * args.steps / float(tg_.item()), "unit": "solves/s",
                                  "h2d_bytes_per_step": int((h_shared.numel() + h_xs.numel() + h_x0.numel() + h_ob.numel()) * 8 + h_g.numel()),
                                  "what": "mpcgpu_solve_sets_guided: halfspaces built on the device from %d obstacle predictions per set "
                                          "and the warm starts (linearized_constraints.cpp:49-189), nothing per planner uploaded but x0" % batch["obst_pred"].shape[2]}

    # ---- latency of ONE homotopy set end to end (H2D -> solve -> select -> D2H), BASELINE.json's second metric
    latency = None
    if rank == 0 and args.latency_reps > 0:
        eng1 = engine.Engine(cfg, device=local_rank, max_batch=planners)
        lat = []
        o1 = {k: pinned(v).numpy() for k, v in eng1.alloc_outputs(planners).items()}      # pinned like the inputs
        for rep in range(args.latency_reps + 5):
            s0 = (rep * 37) % n_sets
            sl = slice(s0 * planners, (s0 + 1) * planners)
            t0 = time.perf_counter()
            o1 = eng1.solve_batch(xi_np[sl], x0_np[sl], p_np[sl], num_iter=args.num_iter, out=o1)
            eng1.select_best(np.array([0, planners], np.int32), o1["pobj"], o1["exit_code"])
            lat.append((time.perf_counter() - t0) * 1e3)
        lat = np.array(lat[5:])
        latency = {"unit": "ms", "what": "one homotopy set (%d planners) through the C ABI, host buffers" % planners,
                   "p50": float(np.percentile(lat, 50)), "p95": float(np.percentile(lat, 95)), "reps": int(args.latency_reps)}
        eng1.close()

    # ---- SURVEY 8(d) beside the headline: the same workload at ONE iteration (the fork's effective behaviour) and the
    #      other configurations (throughput + FP64 fraction each), device-resident, kernel timed by the engine's CUDA events
    flops = load_json(os.path.join(ROOT, "oracle", "flops.json"), {})
    extra = None
    if not args.no_extra:
        extra = {}

        def frac_of(cfgname, nit, rate, ipm):
            fl = flops.get("%s/iter%d" % (cfgname, nit))
            if not fl:
                return None, None
            fps = fl["flops_fixed_part"] + fl["flops_per_ipm_iter"] * ipm
            return fps, fps * rate / 1e12 / fp64_peak

        kms1 = []
        for i in (0, 1, 0, 1):
            step_device(which=i, num_iter=1)
            torch.cuda.synchronize()
            kms1.append(eng.last_kernel_ms())
        ipm1 = float(d_ipm.float().mean().item())
        r1 = n / (float(np.mean(kms1[1:])) * 1e-3)
        fps1, fr1 = frac_of(cfg, 1, r1, ipm1)
        extra["%s/iter1" % cfg] = {"solves_per_s_per_gpu": r1, "kernel_ms": float(np.mean(kms1[1:])), "ipm_iters_mean": ipm1,
                                   "success_frac": float((d_exit == 1).float().mean().item()), "flops_per_solve": fps1, "fp64_frac": fr1, "n": n}
        # the fork's steady state: ONE RTI iteration per control cycle, started from the previous cycle's solution and
        # the capsule's multipliers / QP warm start (acados_solver_interface.cpp:344-376 keeps the unshifted previous output)
        xi_, x0_, pr_ = d_in[0]
        d_mem = torch.zeros((n, eng.mem_doubles), dtype=torch.float64, device=dev)
        eng.solve_batch_device(n, xi_.data_ptr(), x0_.data_ptr(), pr_.data_ptr(), args.num_iter, d_xtraj.data_ptr(), d_utraj.data_ptr(),
                               d_pobj.data_ptr(), d_exit.data_ptr(), d_qps.data_ptr(), d_res.data_ptr(), ipm_iters=d_ipm.data_ptr(),
                               mem=d_mem.data_ptr(), stream=stream.cuda_stream)
        torch.cuda.synchronize()
        ok_prev = (d_exit == 1)
        prev = torch.cat([torch.cat([d_utraj.view(n, N, nu), torch.zeros((n, 1, nu), dtype=torch.float64, device=dev)], dim=1),
                          d_xtraj.view(n, N + 1, nx)], dim=2)
        x0w = torch.where(ok_prev[:, None, None], prev, x0_.view(n, N + 1, nz)).contiguous()
        d_mem0 = d_mem.clone()
        kmsw = []
        for _ in range(4):
            d_mem.copy_(d_mem0)
            torch.cuda.synchronize()
            eng.solve_batch_device(n, xi_.data_ptr(), x0w.data_ptr(), pr_.data_ptr(), 1, d_xtraj.data_ptr(), d_utraj.data_ptr(),
                                   d_pobj.data_ptr(), d_exit.data_ptr(), d_qps.data_ptr(), d_res.data_ptr(), ipm_iters=d_ipm.data_ptr(),
                                   mem=d_mem.data_ptr(), stream=stream.cuda_stream)
            torch.cuda.synchronize()
            kmsw.append(eng.last_kernel_ms())
        ipmw = float(d_ipm.float().mean().item())
        rw = n / (float(np.mean(kmsw[1:])) * 1e-3)
        fpsw, frw = frac_of(cfg, 1, rw, ipmw)
        extra["%s/iter1_warm" % cfg] = {"solves_per_s_per_gpu": rw, "kernel_ms": float(np.mean(kmsw[1:])), "ipm_iters_mean": ipmw,
                                        "success_frac": float((d_exit == 1).float().mean().item()),
                                        "success_frac_previous_cycle": float(ok_prev.float().mean().item()), "flops_per_solve": fpsw,
                                        "fp64_frac": frw, "n": n,
                                        "what": "one RTI iteration from the previous cycle's solution + capsule memory (multipliers, QP warm start)"}
        del d_mem, d_mem0, x0w, prev
        step_device(which=0)                      # leave the buffers as the e2e comparison expects them
        torch.cuda.synchronize()
        for ocfg in [c for c in sorted(PLANNERS) if c != cfg]:
            opl = PLANNERS[ocfg]
            osets = max(148 * 8 * 4 // opl, 1)     # >= 4 waves of the persistent grid
            oeng = engine.Engine(ocfg, device=local_rank, max_batch=osets * opl)
            ob = synthetic.make_batch(oeng.parameter_map, oeng.dims, osets, opl, seed=99 + rank)
            o_d = [torch.from_numpy(ob[k]).to(dev) for k in ("xinit", "x0", "params")]
            on = ob["n"]
            o_out = [torch.empty((on, (oeng.N + 1) * oeng.nx), dtype=torch.float64, device=dev), torch.empty((on, oeng.N * oeng.nu), dtype=torch.float64, device=dev),
                     torch.empty(on, dtype=torch.float64, device=dev), torch.empty(on, dtype=torch.int32, device=dev),
                     torch.empty(on, dtype=torch.int32, device=dev), torch.empty(on, dtype=torch.float64, device=dev), torch.empty(on, dtype=torch.int32, device=dev)]
            torch.cuda.synchronize()
            for nit in (10, 1):
                ms_ = []
                for _ in range(3):
                    oeng.solve_batch_device(on, o_d[0].data_ptr(), o_d[1].data_ptr(), o_d[2].data_ptr(), nit, o_out[0].data_ptr(), o_out[1].data_ptr(),
                                            o_out[2].data_ptr(), o_out[3].data_ptr(), o_out[4].data_ptr(), o_out[5].data_ptr(), ipm_iters=o_out[6].data_ptr(),
                                            stream=stream.cuda_stream)
                    torch.cuda.synchronize()
                    ms_.append(oeng.last_kernel_ms())
                rate_ = on / (float(np.mean(ms_[1:])) * 1e-3)
                ipm_ = float(o_out[6].float().mean().item())
                fps_, fr_ = frac_of(ocfg, nit, rate_, ipm_)
                extra["%s/iter%d" % (ocfg, nit)] = {"solves_per_s_per_gpu": rate_, "kernel_ms": float(np.mean(ms_[1:])), "ipm_iters_mean": ipm_,
                                                    "success_frac": float((o_out[3] == 1).float().mean().item()), "flops_per_solve": fps_, "fp64_frac": fr_,
                                                    "n": on, "N": oeng.N, "nx": oeng.nx, "npar": oeng.npar, "nh": oeng.nh}
            del o_d, o_out
            oeng.close()
        torch.cuda.empty_cache()

    # ---- several GPUs from ONE process (mpcgpu_multi_*: one host thread + streams per device, by-set partition, only the decision
    #      records and the selected trajectory of every set come back): rank 0 drives all GPUs of the job while the other ranks wait
    e2e_multi = None
    if world > 1:
        cpu_group = dist.new_group(backend="gloo")      # the waiting ranks must wait on the HOST: an NCCL barrier would spin on their GPUs,
        barrier()                                       # which rank 0 is about to drive from its own process
        if rank == 0:
            multi = engine.MultiEngine(cfg, list(range(world)), n)
            Pv = h_np[0][2].reshape(n_sets, planners, N, npar)
            differs_m = np.nonzero((Pv[:64] != Pv[:64, :1]).any(axis=(0, 1, 2)))[0].astype(np.int32)
            tot_sets = n_sets * world
            m_xs = pinned(np.tile(np.ascontiguousarray(h_np[0][0].reshape(n_sets, planners, nx)[:, 0]), (world, 1))).numpy()
            m_sh = pinned(np.tile(np.ascontiguousarray(Pv[:, 0]), (world, 1, 1))).numpy()
            m_x0 = pinned(np.tile(h_np[0][1], (world, 1))).numpy()
            m_pv = pinned(np.tile(np.ascontiguousarray(Pv[..., differs_m]), (world, 1, 1, 1))).numpy()
            lb_, lc_ = eng.lin_constraint_block()
            what_m = "mpcgpu_multi_solve_sets"
            if lc_ > 0 and set(differs_m.tolist()) <= set(range(lb_, lb_ + 3 * lc_)) and batches[0]["obst_pred"].size:
                # guidance halfspaces built on each device (7 KB per solve from the host: every device copies and expands its range
                # first and solves it in one launch)
                m_ob = pinned(np.tile(batches[0]["obst_pred"], (world, 1, 1, 1))).numpy()
                m_g = np.tile(batches[0]["guided"], world)
                run_m = lambda: multi.solve_sets(tot_sets, planners, m_xs, m_sh, m_x0, None, None, num_iter=args.num_iter, best_only=True,
                                                 guided_args=(m_ob, m_g, batches[0]["robot_radius"], lb_, lc_))
                what_m = "mpcgpu_multi_solve_sets_guided"
                m_pv = m_ob
                lay_m = eng.table_layout()
                if lay_m.get("ellipsoid") and set(range(npar)) - set(lay_m["invariant_idx"].tolist()) <= (
                        set(range(lb_, lb_ + 3 * lc_)) | set(range(lay_m["ellipsoid"]["base"], lay_m["ellipsoid"]["base"] + lay_m["ellipsoid"]["count"] * lay_m["ellipsoid"]["stride"]))):
                    # struct-of-tables entry: 2.4 KB per solve from the host, the parameter blocks are built on every device
                    m_inv = pinned(np.tile(np.ascontiguousarray(Pv[:, 0, 0][:, lay_m["invariant_idx"]]), (world, 1))).numpy()
                    m_rad = pinned(np.full((tot_sets, m_ob.shape[2]), synthetic.OBSTACLE_RADIUS)).numpy()
                    run_m = lambda: multi.solve_sets_tables(tot_sets, planners, m_xs, m_inv, m_ob, m_x0, guided=m_g, robot_radius=batches[0]["robot_radius"],
                                                            obstacle_radius=m_rad, num_iter=args.num_iter, best_only=True)
                    what_m = "mpcgpu_multi_solve_sets_tables"
                    m_sh = m_inv
            else:
                run_m = lambda: multi.solve_sets(tot_sets, planners, m_xs, m_sh, m_x0, differs_m, m_pv, num_iter=args.num_iter, best_only=True)
            om = run_m()
            t0 = time.perf_counter()
            for _ in range(max(2, args.steps // 2)):
                om = run_m()
            dtm = (time.perf_counter() - t0) / max(2, args.steps // 2)
            assert (om["best"][:n_sets] == best).all() and (om["best"].reshape(world, n_sets) == best[None]).all()
            e2e_multi = {"value": tot_sets * planners / dtm, "unit": "solves/s", "devices": world, "process": "one (rank 0), a host thread per GPU",
                         "h2d_bytes_per_step": int((m_xs.size + m_sh.size + m_x0.size + m_pv.size) * 8),
                         "d2h_bytes_per_step": int(tot_sets * planners * 28 + tot_sets * (4 + ((N + 1) * nx + N * nu) * 8)),
                         "kernel_ms_max_over_devices": multi.last_kernel_ms(),
                         "what": what_m + ": contiguous ranges of whole sets per GPU, device-side selection, decision records + the selected trajectory gathered"}
            multi.close()
        dist.barrier(group=cpu_group)
        barrier()

    # ---- roofline of the dominant kernel (mpc_solve_kernel)
    fkey = "%s/iter%d" % (cfg, args.num_iter)
    kms = float(np.mean(kernel_ms))
    roofline = None
    if fkey in flops:
        fl = flops[fkey]
        # scale the canonical sample figure to this batch's measured interior-point iteration count
        fps = fl["flops_fixed_part"] + fl["flops_per_ipm_iter"] * ipm_mean
        achieved = fps * n / (kms * 1e-3) / 1e12
        peaks = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"), {})
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        hbm_ach = bytes_per_solve(d) * n / (kms * 1e-3) / 1e9
        tr = load_json(os.path.join(ROOT, "profiles", "traffic.json"), {}).get(fkey)
        traffic = tr["dram_bytes_per_solve"] * n if tr else None     # per launch, from the committed ncu --set full capture
        ex = load_json(os.path.join(ROOT, "profiles", "executed_flops.json"), {}).get(fkey)
        executed = None
        if ex:      # FP64 operations the kernel actually EXECUTES (ncu thread-instruction counts; the sparsity of W is exploited,
                    # so it is about half the dense algorithmic count the fraction above is quoted on)
            exf = ex["executed_flops_per_ipm_iter"] * ipm_mean
            executed = {"flops_per_solve": exf, "achieved": exf * n / (kms * 1e-3) / 1e12, "unit": "TFLOP/s",
                        "frac": exf * n / (kms * 1e-3) / 1e12 / fp64_peak, "source": ex["source"]}
        roofline = {"bound": "fp64", "kernel": "mpc_solve_kernel", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp64_peak, "traffic": traffic,
                    "traffic_note": "DRAM bytes per launch (ncu) vs %d algorithmic: parameters re-read in most SQP iterations + thread-local lines written back "
                                    "(L2 holds the stack frames and parameter blocks of 1 184 problems in flight); latency bound, DRAM bus < 2 %% busy" % (bytes_per_solve(d) * n),
                    "peak_source": "DFMA micro-kernel measured in this run (MEASURED_PEAKS.json has no FP64 entry)",
                    "flops_per_solve": fps, "kernel_ms": kms, "executed": executed,
                    "hbm": {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak,
                            "bytes_per_solve": bytes_per_solve(d),
                            "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}}

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on the box's host cores
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # ONE core-count policy for every CPU throughput figure of this file (cpu_baseline here and the --impl reference arm):
        # all host cores, OpenMP over problems.  The single-set LATENCY uses the reference's own team size instead
        # (`omp parallel for num_threads(8)` over the planners of one set, guidance_constraints.cpp:304).
        threads = os.cpu_count() or 1
        cpu_sets = args.cpu_sets if args.cpu_sets > 0 else max(32, int(20.0 * 140.0 * threads / planners))      # ~20 s of CPU work
        rate, dt, cnt = cpu_oracle_rate(cfg, planners, args.num_iter, cpu_sets, threads)
        if latency is not None:      # the same single-set latency on the host cores (OpenMP over the planners)
            from oracle_binding import Oracle
            orc = Oracle(cfg)
            cl = []
            lat_threads = min(8, threads)
            for rep in range(12):
                sl = slice(rep * planners, (rep + 1) * planners)
                t0 = time.perf_counter()
                r = orc.solve_batch(xi_np[sl], x0_np[sl], p_np[sl], num_iter=args.num_iter, threads=lat_threads)
                orc.select_best(np.array([0, planners], np.int32), r["pobj"], r["exit_code"])
                cl.append((time.perf_counter() - t0) * 1e3)
            latency["cpu_port_p50"] = float(np.percentile(cl[2:], 50))
            latency["cpu_cores"] = lat_threads
            latency["reference_measured_ms"] = "35.3 mean / 34.6 p50 ('Optimization' scope, 5 planners, reference traces, BASELINE.md)"
        cpu_baseline = {"value": rate, "unit": "solves/s", "cores": threads, "kind": "port",
                        "cores_policy": "all host cores (os.cpu_count()), OpenMP over problems -- the same policy as the --impl reference arm",
                        "sample": "%d homotopy sets x %d planners (%d solves, %.1f s wall) of the same workload" % (
                            cpu_sets, planners, cnt, dt),
                        "reference_measured": "20-25 ms per planner solve on the reference's own traces (BASELINE.md), i.e. 40-50 solves/s/core"}

    if dist is not None:
        lt = torch.tensor([launches], dtype=torch.float64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(cfg, planners, args.num_iter), "homotopy_sets_per_gpu": n_sets,
                           "solves_per_gpu_per_step": n, "l2": "inputs (%.2f GB/GPU) larger than L2" % (h2d / 1e9),
                           "parallelism": "sets sharded over %d GPU(s), no collective on the solve path" % world,
                           "success_frac": float((exit_codes == 1).mean()), "ipm_iters_mean": ipm_mean,
                           "generator": generator, "generation_s": t_gen, "resident_batches": len(batches),
                           "inputs": "two resident batches rotated between steps: no step re-solves the inputs of the previous one"},
                "e2e": {"value": e2e_value, "unit": "solves/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
                "latency": latency, "e2e_sets": e2e_sets, "e2e_multi": e2e_multi, "extra": extra}
        emit_json(line)


if __name__ == "__main__":
    main()
