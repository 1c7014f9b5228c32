"""Algorithm-independent check at the NLP level: after enough SQP iterations the oracle's iterate must be
a KKT point of the REFERENCE's nonlinear program -- stationarity of the Lagrangian,
dynamics, bounds, path constraints and complementarity are re-assembled here in numpy from the model
functions (which the golden tests pin to the reference's own expressions) and the multipliers the solver
returns in its capsule memory.  Nothing of the Riccati / interior-point code is trusted."""
import ctypes

import numpy as np
import pytest

from oracle_binding import Oracle
from oscar_mpc_planner_mr_modification_b200 import synthetic
from test_oracle_model import model_eval

P = lambda a: a.ctypes.data_as(ctypes.c_void_p)


def nlp_kkt(orc, xinit, params, x, u, mem):
    N, nx, nu, nz, nc, nh = orc.N, orc.nx, orc.nu, orc.nz, orc.nc, orc.nh
    m = mem[1:]
    pi = m[:(N + 1) * nx].reshape(N + 1, nx); m = m[(N + 1) * nx:]
    lam = m[:N * nc].reshape(N, nc)
    lb, ub, lh, uh = orc.bounds(0), orc.bounds(1), orc.bounds(2), orc.bounds(3)
    hrow, hs, hb = [], [], []
    for r in range(nh):
        if lh[r] > -1e10: hrow.append(r); hs.append(1.0); hb.append(lh[r])
        if uh[r] < 1e10: hrow.append(r); hs.append(-1.0); hb.append(uh[r])
    pp = params.reshape(N, orc.npar)
    stat = dyn = feas = comp = 0.0
    for k in range(N):
        z = np.concatenate([u[k], x[k]])
        xn = np.zeros(nx); W = np.zeros((nx, nz)); Hd = np.zeros((nz, nz))
        orc.lib.oracle_integrate(P(x[k].copy()), P(u[k].copy()), P(pp[k].copy()), P(np.zeros(nx)), P(xn), P(W), P(Hd))
        dyn = max(dyn, np.abs(xn - x[k + 1]).max())
        ev = model_eval(orc, z, pp[k].copy(), np.zeros(nx), np.zeros(max(nh, 1)))
        chat = np.vstack([np.eye(nz), -np.eye(nz)] + [hs[j] * ev["Jh"][hrow[j]][None] for j in range(len(hrow))])
        slack = np.concatenate([z - lb, ub - z] + [[hs[j] * (ev["h"][hrow[j]] - hb[j])] for j in range(len(hrow))])
        act = np.ones(nc, bool)
        if k == 0:
            act[nu:nz] = False; act[nz + nu:2 * nz] = False
        r = 0.2 * ev["g"] + W.T @ pi[k + 1] - chat[act].T @ lam[k][act]
        if k > 0:
            r[nu:] -= pi[k]
            stat = max(stat, np.abs(r).max())
        else:
            stat = max(stat, np.abs(r[:nu]).max())
        feas = max(feas, max(0.0, (-slack[act]).max()))
        comp = max(comp, np.abs(lam[k][act] * slack[act]).max())
        assert (lam[k][act] >= 0).all()
    return stat, dyn, feas, comp, np.abs(x[0] - xinit).max()


@pytest.mark.parametrize("cfg,planners", [("c1_basic", 1), ("tmpc_shipped", 5)])
def test_converged_iterate_is_a_kkt_point_of_the_reference_nlp(cfg, planners):
    orc = Oracle(cfg)
    b = synthetic.make_batch(orc.parameter_map, orc.dims, 6 if planners == 1 else 2, planners, seed=41)
    mem = np.zeros((b["n"], orc.mem_doubles))
    r = orc.solve_batch(b["xinit"], b["x0"], b["params"], num_iter=40, mem=mem)
    checked = 0
    for i in np.nonzero(r["exit_code"] == 1)[0]:
        x = r["xtraj"][i].reshape(orc.N + 1, orc.nx); u = r["utraj"][i].reshape(orc.N, orc.nu)
        stat, dyn, feas, comp, x0err = nlp_kkt(orc, b["xinit"][i], b["params"][i], x, u, mem[i])
        if dyn > 1e-6:
            continue            # SQP (fixed full steps, no globalisation) has not settled on this instance
        checked += 1
        # the multipliers come from the LAST QP, linearised one step earlier: first-order accurate
        assert stat < 5e-4, (i, stat)
        assert feas < 1e-5 and comp < 1e-3 and x0err < 1e-9, (i, feas, comp, x0err)
    assert checked >= 2
