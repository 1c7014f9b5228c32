"""N>1 path on CPU: world_size 2 over gloo.  The engine call is replaced by the oracle (test-only) so
the host-side sharding / gathering logic is exercised exactly as bench.py uses it on GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oscar_mpc_planner_mr_modification_b200 import sharding, synthetic


def test_shard_ranges_partition_everything():
    for n in (0, 1, 7, 16, 4097):
        for w in (1, 2, 3, 8):
            r = [sharding.shard_range(n, w, k) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
    with pytest.raises(ValueError):
        sharding.shard_range(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_binding import Oracle
    orc = Oracle("tmpc_shipped")
    batch = synthetic.make_batch(orc.parameter_map, orc.dims, 5, 5, seed=31)       # 5 sets: uneven split 2 + 3
    table, out, sh = sharding.solve_sharded(batch, lambda a, b, c, n: orc.solve_batch(a, b, c, num_iter=n, threads=2),
                                            orc.select_best, 3, dist)
    q.put((rank, table, sh["set_begin"], sh["set_end"]))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_matches_single_process():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert (res[0][2], res[0][3], res[1][2], res[1][3]) == (0, 2, 2, 5)
    from oracle_binding import Oracle
    orc = Oracle("tmpc_shipped")
    batch = synthetic.make_batch(orc.parameter_map, orc.dims, 5, 5, seed=31)
    ref, _, _ = sharding.solve_sharded(batch, lambda a, b, c, n: orc.solve_batch(a, b, c, num_iter=n, threads=2),
                                       orc.select_best, 3, None)
    for rank, table, _, _ in res:
        np.testing.assert_array_equal(np.nan_to_num(table, nan=-7.0), np.nan_to_num(ref, nan=-7.0))   # same table on every rank
