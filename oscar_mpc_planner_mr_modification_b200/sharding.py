"""Multi-GPU layout of the solve path: homotopy sets are independent, so the batch is
partitioned BY SET into contiguous ranges, one range per rank / GPU, and solved with no collective on
the data path.  Only the small per-set decision records are gathered afterwards (off the timed path).

The reference's counterpart is the OpenMP team over the planners of one set
(mpc_planner_modules/src/guidance_constraints.cpp:304) and, across robots, one ROS node per robot.
"""
import numpy as np


def shard_range(n_sets, world, rank):
    """Contiguous [begin, end) range of sets of `rank`; the remainder goes to the last rank."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    per = n_sets // world
    begin = rank * per
    end = n_sets if rank == world - 1 else begin + per
    return begin, end


def shard_batch(batch, world, rank):
    """Slice a synthetic.make_batch() dict (problems of a set are contiguous) to this rank's sets."""
    off = np.asarray(batch["set_offsets"])
    b, e = shard_range(off.size - 1, world, rank)
    lo, hi = int(off[b]), int(off[e])
    return dict(xinit=batch["xinit"][lo:hi], x0=batch["x0"][lo:hi], params=batch["params"][lo:hi],
                set_offsets=(off[b:e + 1] - off[b]).astype(np.int32), n=hi - lo, set_begin=b, set_end=e)


def solve_sharded(batch, solve_fn, select_fn, num_iter, dist=None):
    """Solve this rank's shard with `solve_fn(xinit, x0, params, num_iter) -> dict` and
    `select_fn(set_offsets, pobj, exit_code) -> best`, then assemble the global per-set table
    {best index, objective and exit code of the chosen planner} on every rank with ONE all_gather of a
    few bytes per set.  `dist` is torch.distributed (any backend) or None for a single process."""
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    sh = shard_batch(batch, world, rank)
    out = solve_fn(sh["xinit"], sh["x0"], sh["params"], num_iter)
    best = np.asarray(select_fn(sh["set_offsets"], out["pobj"], out["exit_code"]), np.int64)
    chosen = sh["set_offsets"][:-1] + np.maximum(best, 0)
    rec = np.stack([best.astype(np.float64), np.where(best >= 0, out["pobj"][chosen], np.nan),
                    np.where(best >= 0, out["exit_code"][chosen], out["exit_code"][sh["set_offsets"][:-1]]).astype(np.float64)],
                   axis=1)
    if dist is None or world == 1:
        return rec, out, sh
    import torch
    n_sets = np.asarray(batch["set_offsets"]).size - 1
    sizes = [shard_range(n_sets, world, r)[1] - shard_range(n_sets, world, r)[0] for r in range(world)]
    pad = max(sizes)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = torch.zeros((pad, 3), dtype=torch.float64, device=dev)
    mine[:rec.shape[0]] = torch.from_numpy(rec).to(dev)
    parts = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    table = np.concatenate([p.cpu().numpy()[:sizes[r]] for r, p in enumerate(parts)], axis=0)
    return table, out, sh
