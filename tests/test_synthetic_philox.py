"""Counter-based synthetic generator: Philox4x32-10 known answers, shard independence, and equivalence of the
scenario definition with synthetic.make_batch (the numpy mirror of the device kernel csrc/mpcgpu_synth.cu)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oscar_mpc_planner_mr_modification_b200 import synthetic  # noqa: E402
from oracle_binding import Oracle  # noqa: E402

# Random123 known-answer vectors for philox4x32-10: counter, key -> output
KAT = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
       ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
       ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


def test_philox_known_answers():
    for ctr, key, out in KAT:
        r = synthetic.philox4x32_10(*[np.array([c]) for c in ctr], *key)
        assert tuple(int(x[0]) for x in r) == out


def test_uniforms_are_in_range_and_keyed_by_set_index():
    gs = np.arange(1000, 1100)
    u = np.stack([synthetic.philox_uniform(1234, gs, j) for j in range(64)])
    assert (u >= 0).all() and (u < 1).all() and 0.45 < u.mean() < 0.55
    one = np.stack([synthetic.philox_uniform(1234, np.array([1042]), j) for j in range(64)])[:, 0]
    assert np.array_equal(one, u[:, 42])
    assert not np.array_equal(synthetic.philox_uniform(1235, gs, 0), u[0])


class _PhiloxAsRng:
    """Feeds make_batch the uniforms of the counter-based generator in make_batch's own call order."""

    def __init__(self, seed, n_sets, M):
        self.seed, self.gs, self.M, self.calls = seed, np.arange(n_sets, dtype=np.uint64), M, 0

    def uniform(self, lo, hi, shape):
        c = self.calls
        self.calls += 1
        if c < 4:
            js = [c]
        elif c < 6:
            js = list(range(4 + 6 * (c - 4), 10 + 6 * (c - 4)))
        else:
            js = [16 + 4 * m + (c - 6) for m in range(self.M)]
        u = np.stack([synthetic.philox_uniform(self.seed, self.gs, j) for j in js], axis=1)
        return lo + (hi - lo) * (u[:, 0] if c < 4 else u)


@pytest.mark.parametrize("cfg,planners", [("c1_basic", 1), ("tmpc_shipped", 5), ("c2_tmpc12", 9), ("c6_goal_unicycle", 1)])
def test_same_scenario_definition_as_make_batch(cfg, planners, monkeypatch):
    orc = Oracle(cfg)
    M = sum(1 for k in orc.parameter_map if k.startswith("ellipsoid_obst_") and k.endswith("_x"))
    monkeypatch.setattr(synthetic.np.random, "default_rng", lambda seed: _PhiloxAsRng(seed, 6, M))
    a = synthetic.make_batch(orc.parameter_map, orc.dims, 6, planners, seed=77)
    b = synthetic.make_batch_philox(orc.parameter_map, orc.dims, 6, planners, seed=77)
    for k in ("xinit", "x0", "params", "obst_pred"):
        assert a[k].shape == b[k].shape, k
        # LAPACK solve vs Thomas, hypot / norm vs sqrt of sums: rounding-level differences only
        assert np.allclose(a[k], b[k], rtol=1e-11, atol=1e-11), (k, np.abs(a[k] - b[k]).max())
    assert np.array_equal(a["guided"], b["guided"]) and np.array_equal(a["set_offsets"], b["set_offsets"])


def test_a_shard_is_a_slice_of_the_global_batch():
    orc = Oracle("c2_tmpc12")
    full = synthetic.make_batch_philox(orc.parameter_map, orc.dims, 10, 9, seed=5)
    part = synthetic.make_batch_philox(orc.parameter_map, orc.dims, 4, 9, seed=5, first_set=3)
    for k in ("xinit", "x0", "params"):
        assert np.array_equal(part[k], full[k][27:63]), k
    assert np.array_equal(part["obst_pred"], full["obst_pred"][3:7])


def test_unsupported_constraint_families_are_refused():
    orc = Oracle("c5_ccmpc")
    with pytest.raises(ValueError):
        synthetic.synth_layout(orc.parameter_map, orc.dims)
