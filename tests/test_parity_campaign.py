"""The parity campaign's verdict as a `-m gpu` test: all six configurations x 2 seeds x {1, 10} iterations x both
solve kernels against the oracle -- exit codes bit-exact, trajectories within north_star's 1e-6 relative."""
import pytest

pytestmark = pytest.mark.gpu


def test_parity_campaign_two_seeds():
    from parity_campaign import campaign
    r = campaign(seeds=(101, 202), scale=0.5, verbose=True)
    print(r)
    assert r["problems"] >= 6000
    assert r["exit_mismatches"] == 0, r["per_config"]
    assert r["worst_rel_err"] < 1e-6, r["per_config"]
    assert set(r["per_config"]) == {"c1_basic", "tmpc_shipped", "c2_tmpc12", "c5_ccmpc", "c6_goal_unicycle", "c7_linearized"}
