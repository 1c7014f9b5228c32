"""mpcgpu_solve_batch from pinned host memory: ONE gated launch whose inputs arrive chunk by chunk on the copy stream
(csrc/mpcgpu_capi.cu, solve_batch_impl) must give exactly what the chunked launches from pageable memory give."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oscar_mpc_planner_mr_modification_b200 import engine, synthetic  # noqa: E402

pytestmark = pytest.mark.gpu
KEYS = ("xtraj", "utraj", "pobj", "exit_code", "qp_status", "res_eq", "ipm_iters")


def _pinned_outputs(eng, n):
    hold, out = [], {}
    for k, v in eng.alloc_outputs(n).items():
        h = engine.PinnedArray(v.shape, v.dtype)
        hold.append(h)
        out[k] = h.array
    return hold, out


@pytest.mark.parametrize("cfg,planners,n_sets", [("c2_tmpc12", 9, 1100), ("c1_basic", 1, 9000)])
def test_gated_launch_equals_chunked_launches(cfg, planners, n_sets):
    eng = engine.Engine(cfg, 0, n_sets * planners)
    eng.set_kernel_mode(engine.KERNEL_STAGE)
    b = synthetic.make_batch_philox(eng.parameter_map, eng.dims, n_sets, planners, seed=11)
    n = b["n"]
    assert n >= 8192
    rng = np.random.default_rng(0)
    nit = rng.integers(1, 4, n).astype(np.int32)           # per-problem iteration counts travel with their chunk too
    ref = eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=nit)      # pageable memory: chunked launches
    ref = {k: np.array(ref[k]) for k in KEYS}
    holders = [engine.pinned_copy(b[k]) for k in ("xinit", "x0", "params")] + [engine.pinned_copy(nit)]
    xi, x0, pr, nit_p = (h[1] for h in holders)
    # pinned inputs, pageable outputs: gated launch, results staged and copied at the end
    o1 = eng.solve_batch(xi, x0, pr, num_iter=nit_p)
    for k in KEYS:
        assert np.array_equal(o1[k], ref[k]), k
    # pinned inputs and outputs: the kernel writes the results into the caller's arrays
    hold, out = _pinned_outputs(eng, n)
    for v in out.values():
        v[...] = 0
    o2 = eng.solve_batch(xi, x0, pr, num_iter=nit_p, out=out)
    for k in KEYS:
        assert np.array_equal(o2[k], ref[k]), k
    assert eng.last_kernel_ms() > 0


def test_gated_launch_with_capsule_memory():
    eng = engine.Engine("tmpc_shipped", 0, 8200)
    eng.set_kernel_mode(engine.KERNEL_STAGE)
    b = synthetic.make_batch_philox(eng.parameter_map, eng.dims, 1640, 5, seed=3)
    n = b["n"]
    mem_a = np.zeros((n, eng.mem_doubles))
    ref = eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=2, mem=mem_a)
    ref = {k: np.array(ref[k]) for k in KEYS}
    hx, hx0, hp = (engine.pinned_copy(b[k]) for k in ("xinit", "x0", "params"))
    hm = engine.PinnedArray((n, eng.mem_doubles))
    hm.array[...] = 0
    o = eng.solve_batch(hx[1], hx0[1], hp[1], num_iter=2, mem=hm.array)
    for k in KEYS:
        assert np.array_equal(o[k], ref[k]), k
    assert np.array_equal(hm.array, mem_a)
    # second cycle from the stored capsules, both ways
    ref2 = eng.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=1, mem=mem_a)
    o2 = eng.solve_batch(hx[1], hx0[1], hp[1], num_iter=1, mem=hm.array)
    for k in KEYS:
        assert np.array_equal(o2[k], ref2[k]), k
    assert np.array_equal(hm.array, mem_a)
