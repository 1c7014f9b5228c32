"""Device-side synthetic generator (mpcgpu_generate_synthetic*, csrc/mpcgpu_synth.cu) against its numpy mirror
(synthetic.make_batch_philox): SURVEY 8d "host and device generate identical data"."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oscar_mpc_planner_mr_modification_b200 import engine, synthetic  # noqa: E402

pytestmark = pytest.mark.gpu
DIMS = {"c1_basic": (30, 5, 2, 83), "tmpc_shipped": (30, 5, 2, 98), "c2_tmpc12": (30, 5, 2, 175), "c6_goal_unicycle": (30, 4, 2, 35)}


def _maps(cfg):
    pm, _, _ = engine.load_maps(cfg)
    N, nx, nu, npar = DIMS[cfg]
    return pm, dict(N=N, nx=nx, nu=nu, npar=npar, dt=0.2)


@pytest.mark.parametrize("cfg,planners", [("c1_basic", 1), ("tmpc_shipped", 5), ("c2_tmpc12", 9), ("c6_goal_unicycle", 1)])
def test_device_equals_host_mirror(cfg, planners):
    pm, dims = _maps(cfg)
    dev = engine.generate_synthetic(pm, dims, 64, planners, seed=1234, first_set=1000)
    host = synthetic.make_batch_philox(pm, dims, 64, planners, seed=1234, first_set=1000)
    # the state, the path and everything else made of + - * / sqrt only: bit for bit
    assert np.array_equal(dev["xinit"], host["xinit"])
    sp = [pm[k] for k in pm if k.startswith("spline")]
    P_d, P_h = dev["params"].reshape(-1, dims["N"], dims["npar"]), host["params"].reshape(-1, dims["N"], dims["npar"])
    assert np.array_equal(P_d[:, :, sp], P_h[:, :, sp])
    # obstacle headings (sin / cos), polyline headings (atan2), braking roll-out (sin / cos): last-bit differences of the two
    # math libraries, amplified at most by the push-out
    for k in ("x0", "params", "obst_pred"):
        assert dev[k].shape == host[k].shape
        assert np.allclose(dev[k], host[k], rtol=1e-12, atol=1e-12), (k, np.abs(dev[k] - host[k]).max())
    assert np.array_equal(dev["guided"], host["guided"])
    frac_exact = float((dev["params"] == host["params"]).mean())
    assert frac_exact > 0.9, frac_exact


def test_shards_reproduce_their_slice_on_the_device():
    pm, dims = _maps("c2_tmpc12")
    full = engine.generate_synthetic(pm, dims, 40, 9, seed=7)
    part = engine.generate_synthetic(pm, dims, 11, 9, seed=7, first_set=20)
    for k in ("xinit", "x0", "params"):
        assert np.array_equal(part[k], full[k][180:279]), k
    other = engine.generate_synthetic(pm, dims, 11, 9, seed=8, first_set=20)
    assert not np.array_equal(other["xinit"], part["xinit"])


def test_generated_in_place_and_solved_like_the_host_mirror():
    torch = pytest.importorskip("torch")
    pm, dims = _maps("c2_tmpc12")
    S, Pn, N, nx, nz, npar = 32, 9, 30, 5, 7, 175
    B = S * Pn
    d = [torch.empty(B * nx, dtype=torch.float64, device="cuda"), torch.empty(B * (N + 1) * nz, dtype=torch.float64, device="cuda"),
         torch.empty(B * N * npar, dtype=torch.float64, device="cuda"), torch.empty(S * N * 12 * 2, dtype=torch.float64, device="cuda")]
    engine.generate_synthetic(pm, dims, S, Pn, seed=99, device_buffers=[t.data_ptr() for t in d], stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ref = engine.generate_synthetic(pm, dims, S, Pn, seed=99)
    assert np.array_equal(d[0].cpu().numpy().reshape(B, nx), ref["xinit"])
    assert np.array_equal(d[1].cpu().numpy().reshape(B, -1), ref["x0"])
    assert np.array_equal(d[2].cpu().numpy().reshape(B, -1), ref["params"])
    assert np.array_equal(d[3].cpu().numpy().reshape(ref["obst_pred"].shape), ref["obst_pred"])
    eng = engine.Engine("c2_tmpc12", 0, 512)
    host = synthetic.make_batch_philox(pm, dims, S, Pn, seed=99)
    a = eng.solve_batch(ref["xinit"], ref["x0"], ref["params"], num_iter=10)
    a = {k: np.array(v) for k, v in a.items()}
    b = eng.solve_batch(host["xinit"], host["x0"], host["params"], num_iter=10)
    assert (a["exit_code"] == b["exit_code"]).mean() > 0.97
    ok = (a["exit_code"] == 1) & (b["exit_code"] == 1)
    assert ok.mean() > 0.7
    assert np.allclose(a["xtraj"][ok], b["xtraj"][ok], atol=1e-6)
