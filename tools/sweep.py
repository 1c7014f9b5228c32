#!/usr/bin/env python3
"""Batch-size / iteration-count sweep of the solve kernel on one GPU (BASELINE configs[2]: 4096 .. 1M
independent N=30 problems).  Inputs: every problem distinct, generated ON THE DEVICE by the counter-based generator
(mpcgpu_generate_synthetic_device) where it covers the configuration; otherwise generated once on the host for `--base-sets`
homotopy sets and tiled on the device.  Device-resident timing with CUDA events.  Writes one JSON line per point."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from oscar_mpc_planner_mr_modification_b200 import engine, synthetic
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2_tmpc12")
    ap.add_argument("--planners", type=int, default=9)
    ap.add_argument("--base-sets", type=int, default=512)
    ap.add_argument("--sizes", default="4096,16384,65536,262144")
    ap.add_argument("--iters", default="1,10")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    sizes = [int(v) for v in args.sizes.split(",")]
    eng = engine.Engine(args.config, 0, max(sizes))
    try:
        synthetic.synth_layout(eng.parameter_map, eng.dims)
        on_device = True
    except ValueError:
        on_device = False
        base = synthetic.make_batch(eng.parameter_map, eng.dims, args.base_sets, args.planners, seed=1234)
        nb = base["n"]
        bx, b0, bp = (torch.from_numpy(base[k]).to(dev) for k in ("xinit", "x0", "params"))
    stream = torch.cuda.Stream(device=dev)
    for n in sizes:
        if on_device:
            ns = (n + args.planners - 1) // args.planners
            n = min(ns * args.planners, max(sizes)) // args.planners * args.planners
            ns = n // args.planners
            xi = torch.empty((n, eng.nx), dtype=torch.float64, device=dev)
            x0 = torch.empty((n, (eng.N + 1) * eng.nz), dtype=torch.float64, device=dev)
            pr = torch.empty((n, eng.N * eng.npar), dtype=torch.float64, device=dev)
            engine.generate_synthetic(eng.parameter_map, eng.dims, ns, args.planners, seed=1234, device_buffers=[xi.data_ptr(), x0.data_ptr(), pr.data_ptr(), None])
        else:
            reps = (n + nb - 1) // nb
            xi = bx.repeat(reps, 1)[:n].contiguous(); x0 = b0.repeat(reps, 1)[:n].contiguous(); pr = bp.repeat(reps, 1)[:n].contiguous()
        xt = torch.empty((n, (eng.N + 1) * eng.nx), dtype=torch.float64, device=dev); ut = torch.empty((n, eng.N * eng.nu), dtype=torch.float64, device=dev)
        po = torch.empty(n, dtype=torch.float64, device=dev); rq = torch.empty(n, dtype=torch.float64, device=dev)
        ec = torch.empty(n, dtype=torch.int32, device=dev); qs = torch.empty(n, dtype=torch.int32, device=dev); ip = torch.empty(n, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        for nit in [int(v) for v in args.iters.split(",")]:
            ms = []
            for rep in range(4):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                eng.solve_batch_device(n, xi.data_ptr(), x0.data_ptr(), pr.data_ptr(), nit, xt.data_ptr(), ut.data_ptr(), po.data_ptr(),
                                       ec.data_ptr(), qs.data_ptr(), rq.data_ptr(), ipm_iters=ip.data_ptr(), stream=stream.cuda_stream)
                e1.record(stream)
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            t = min(ms[1:])
            print(json.dumps({"config": args.config, "n": n, "num_iter": nit, "kernel_ms": t, "solves_per_s": n / t * 1e3,
                              "success_frac": float((ec == 1).float().mean().item()), "ipm_iters_mean": float(ip.float().mean().item()),
                              "input_gb": (xi.numel() + x0.numel() + pr.numel()) * 8 / 1e9,
                              "inputs": "distinct, generated on the device" if on_device else "tiled from %d host-generated sets" % args.base_sets}), flush=True)
        del xi, x0, pr, xt, ut
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
