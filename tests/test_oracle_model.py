"""The oracle's model code against golden vectors taken from the REFERENCE's own symbolic expressions
(tests/golden/model_<cfg>.npz, produced by tests/golden/make_golden.py with sympy.lambdify), plus
finite-difference checks of every second-order quantity and of the ERK4 sensitivities."""
import ctypes
import os

import numpy as np
import pytest

from oracle_binding import Oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CONFIGS = ["c1_basic", "tmpc_shipped", "c2_tmpc12", "c5_ccmpc", "c6_goal_unicycle", "c7_linearized"]
P = lambda a: a.ctypes.data_as(ctypes.c_void_p)


def model_eval(orc, z, p, mu, mh):
    nx, nz, nh = orc.nx, orc.nz, orc.nh
    f = np.zeros(nx); Jf = np.zeros((nx, nz)); Hf = np.zeros((nz, nz)); c = ctypes.c_double(); g = np.zeros(nz)
    Hc = np.zeros((nz, nz)); h = np.zeros(max(nh, 1)); Jh = np.zeros((max(nh, 1), nz)); Hh = np.zeros((nz, nz))
    orc.lib.oracle_model_eval(P(z), P(p), P(mu), P(mh), P(f), P(Jf), P(Hf), ctypes.byref(c), P(g), P(Hc), P(h), P(Jh), P(Hh))
    return dict(f=f, Jf=Jf, Hf=Hf, cost=c.value, g=g, Hc=Hc, h=h[:nh], Jh=Jh[:nh], Hh=Hh)


@pytest.mark.parametrize("cfg", CONFIGS)
def test_model_matches_reference_expressions(cfg):
    gd = np.load(os.path.join(GOLD, "model_%s.npz" % cfg))
    orc = Oracle(cfg)
    assert [n for n in gd["param_names"]] == [orc.lib.oracle_param_name(i).decode() for i in range(orc.npar)]
    np.testing.assert_array_equal(orc.bounds(0), gd["lb"])
    np.testing.assert_array_equal(orc.bounds(1), gd["ub"])
    np.testing.assert_array_equal(orc.bounds(2)[:orc.nh], gd["lh"])
    np.testing.assert_array_equal(orc.bounds(3)[:orc.nh], gd["uh"])
    for i in range(gd["z"].shape[0]):
        r = model_eval(orc, gd["z"][i].copy(), gd["p"][i].copy(), np.zeros(orc.nx), np.zeros(max(orc.nh, 1)))
        np.testing.assert_allclose(r["f"], gd["f"][i], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(r["Jf"], gd["jf"][i], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(r["cost"], gd["cost"][i], rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(r["g"], gd["grad"][i], rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(r["h"], gd["h"][i], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(r["Jh"], gd["jh"][i], rtol=1e-10, atol=1e-11)


def test_reference_asserted_facts():
    """The two numeric facts the reference's own tests assert (solver_generator/test/test_control_modules.py):
    stage objective > 0 (:55-66) and an ellipsoid constraint strictly inside its bounds (:83-103)."""
    orc = Oracle("c1_basic")
    z = np.zeros(orc.nz); z[3] = 5.0
    p = np.ones(orc.npar)
    r = model_eval(orc, z, p, np.zeros(orc.nx), np.zeros(orc.nh))
    assert r["cost"] > 0
    pm = orc.parameter_map
    p = np.zeros(orc.npar)
    for j in range(4):       # obstacle at (5, 10), radius 1, robot at the origin (p[2]=5, p[3]=10, p[-1]=1 in the 1-obstacle test)
        p[pm["ellipsoid_obst_%d_x" % j]] = 5.0; p[pm["ellipsoid_obst_%d_y" % j]] = 10.0; p[pm["ellipsoid_obst_%d_r" % j]] = 1.0
    r = model_eval(orc, np.zeros(orc.nz), p, np.zeros(orc.nx), np.zeros(orc.nh))
    lh, uh = orc.bounds(2), orc.bounds(3)
    assert (r["h"] > lh).all() and (r["h"] < uh).all()
    np.testing.assert_allclose(r["h"], 125.0)       # (5^2 + 10^2) / 1^2
    assert len(lh) == 4 and orc.nx == 5 and orc.nu == 2      # test_acados.py:68-71 shape facts (M obstacles -> M rows)


@pytest.mark.parametrize("cfg", CONFIGS)
def test_second_derivatives_finite_difference(cfg):
    gd = np.load(os.path.join(GOLD, "model_%s.npz" % cfg))
    orc = Oracle(cfg)
    rng = np.random.default_rng(5)
    eps = 1e-6
    for i in range(0, gd["z"].shape[0], 4):
        z, p = gd["z"][i].copy(), gd["p"][i].copy()
        mu, mh = rng.normal(size=orc.nx), rng.normal(size=max(orc.nh, 1))
        r = model_eval(orc, z, p, mu, mh)
        Hc, Hh, Hf = np.zeros((orc.nz,) * 2), np.zeros((orc.nz,) * 2), np.zeros((orc.nz,) * 2)
        for j in range(orc.nz):
            zp, zm = z.copy(), z.copy()
            zp[j] += eps; zm[j] -= eps
            a, b = model_eval(orc, zp, p, mu, mh), model_eval(orc, zm, p, mu, mh)
            Hc[:, j] = (a["g"] - b["g"]) / (2 * eps)
            Hh[:, j] = ((a["Jh"] - b["Jh"]) / (2 * eps)).T @ mh[:orc.nh]
            Hf[:, j] = ((a["Jf"] - b["Jf"]) / (2 * eps)).T @ mu
        for name, fd in (("Hc", Hc), ("Hh", Hh), ("Hf", Hf)):
            scale = max(1.0, np.abs(fd).max())
            assert np.abs(r[name] - fd).max() / scale < 2e-7, (cfg, name)
            assert np.abs(r[name] - r[name].T).max() == 0.0


@pytest.mark.parametrize("cfg", ["c1_basic"])
def test_erk4_sensitivities_and_adjoint_hessian(cfg):
    orc = Oracle(cfg)
    rng = np.random.default_rng(0)
    nx, nu, nz = orc.nx, orc.nu, orc.nz

    def integ(z, pi):
        xn = np.zeros(nx); W = np.zeros((nx, nz)); Hc = np.zeros((nz, nz)); p = np.zeros(orc.npar)
        orc.lib.oracle_integrate(P(z[nu:].copy()), P(z[:nu].copy()), P(p), P(pi), P(xn), P(W), P(Hc))
        return xn, W, Hc

    for _ in range(3):
        z = rng.normal(size=nz); z[nu + 3] = rng.uniform(0, 3)
        pi = rng.normal(size=nx)
        xn, W, Hc = integ(z, pi)
        # independent plain RK4 (3 steps of dt/3) on the unicycle written out by hand (solver_model.py:207-214)
        f = lambda x, u: np.array([x[3] * np.cos(x[2]), x[3] * np.sin(x[2]), u[1], u[0], x[3]])
        x, u, h = z[nu:].copy(), z[:nu], 0.2 / 3
        for _s in range(3):
            k1 = f(x, u); k2 = f(x + h / 2 * k1, u); k3 = f(x + h / 2 * k2, u); k4 = f(x + h * k3, u)
            x = x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        np.testing.assert_allclose(xn, x, rtol=1e-13, atol=1e-14)
        eps = 1e-6
        Wfd, Hfd = np.zeros((nx, nz)), np.zeros((nz, nz))
        for j in range(nz):
            zp, zm = z.copy(), z.copy()
            zp[j] += eps; zm[j] -= eps
            xp, Wp, _ = integ(zp, pi); xm, Wm, _ = integ(zm, pi)
            Wfd[:, j] = (xp - xm) / (2 * eps)
            Hfd[:, j] = ((Wp - Wm) / (2 * eps)).T @ pi
        assert np.abs(W - Wfd).max() < 1e-8
        assert np.abs(Hc - Hfd).max() < 1e-8


def test_mirror_matches_eigendecomposition():
    orc = Oracle("c1_basic")
    rng = np.random.default_rng(3)
    cases = [np.zeros((7, 7)), np.diag([1e-6, -1e-6, 2.0, -3.0, 0.0, 5e-5, -5e-5])]
    for _ in range(20):
        A = rng.normal(size=(7, 7)); cases.append(A + A.T)
    B = np.zeros((7, 7)); B[:4, :4] = rng.normal(size=(4, 4)); B[4:, 4:] = rng.normal(size=(3, 3)); cases.append(B + B.T)
    for A in cases:
        w, V = np.linalg.eigh(A)
        w2 = np.where(np.abs(w) <= 1e-4, 1e-4, np.abs(w))
        ref = (V * w2) @ V.T
        M = np.ascontiguousarray(A.copy())
        orc.lib.oracle_mirror(P(M), 7)
        assert np.abs(M - ref).max() < 1e-10 * max(1.0, np.abs(A).max())
        assert np.linalg.eigvalsh(M).min() > 1e-4 * (1 - 1e-6)
