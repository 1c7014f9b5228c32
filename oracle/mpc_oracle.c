/* mpc_oracle.c -- CPU ORACLE for the batched MPC solve path.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this.  The product (libmpcgpu.so) never links or calls it.
 *
 * PARITY UNPINNED: the reference's arithmetic for this path lives in acados + HPIPM + BLASFEO +
 * CasADi-generated C, none of which is vendored in /root/reference (acados is not even version
 * pinned: pyproject.toml:18, README.md:226-234) and none of which is installable here.  The
 * reference's tests hold no golden solve results (solver_generator/test/test_acados.py:48-77 never
 * solves).  This file therefore RESTATES the published algorithms with every upstream-only choice
 * written down explicitly (see DESIGN.md "Algorithm contract"); the problem definition (dynamics,
 * cost, constraints, bounds, parameter order) IS pinned: oracle/generated/model_*.h is derived
 * mechanically from the reference's own Python scripts (oracle/gen_model.py).
 *
 * What is restated, with the reference call sites:
 *   Solver::solve() loop, early exits, res_eq rule, exit-code map
 *                         mpc_planner_solver/src/acados_solver_interface.cpp:86-204
 *   loadWarmstart         acados_solver_interface.cpp:274-284
 *   stage-N parameter reuse                              acados_solver_interface.cpp:128-134
 *   OCP formulation/options (ERK 4 stages x 3 steps, EXACT Hessian, MIRROR, FIXED_STEP, qp_tol 1e-5,
 *   HPIPM iter_max 50, warm start 2)  solver_generator/generate_acados_solver.py:84-177
 *   FindBestPlanner / objective post-processing
 *                         mpc_planner_modules/src/guidance_constraints.cpp:373-420,572-590,1025-1050
 *
 * Written in the common subset of C and C++: compile as C for the oracle proper; compile as C++
 * with -DREAL=<counting type> to obtain the canonical algorithmic FLOP count (oracle/flopcount.cpp).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef ORACLE_TRACE
#include <stdio.h>
#endif

#ifndef REAL
#define REAL double
#endif

#ifndef MODEL_HEADER
#error "compile with -DMODEL_HEADER=\"generated/model_<config>.h\""
#endif
#include MODEL_HEADER

#define NX MODEL_NX
#define NU MODEL_NU
#define NZ MODEL_NZ
#define NP MODEL_NP
#define NH MODEL_NH
#define NN MODEL_N

/* ---- explicit upstream-only choices (DESIGN.md "Algorithm contract") ------------------------- */
#define SIM_STEPS 3          /* generate_acados_solver.py:150 */
#define REG_EPS 1e-4         /* acados reg_epsilon default for MIRROR */
#define BOUND_INF 1e10       /* |bound| >= this => bound absent */
#define IPM_ITER_MAX 50      /* generate_acados_solver.py:172 */
#define IPM_TOL 1e-5         /* generate_acados_solver.py:162 (all four residual tolerances) */
#define IPM_MU0 10.0         /* HPIPM BALANCE mode */
#define IPM_THR0 0.1         /* HPIPM init threshold for slacks */
#define IPM_ALPHA_MIN 1e-12
#define IPM_LAM_MIN 1e-16
#define IPM_T_MIN 1e-16
#define IPM_STEP_SCALE 0.995
#define RES_EQ_MAX 1e-2      /* acados_solver_interface.cpp:177 */
#define JACOBI_MAX_SWEEPS 30
#define JACOBI_TOL 1e-24     /* stop when sum offdiag^2 <= tol * sum all^2 */

/* compact constraint list of one path stage: [u lower NU][u upper NU][x lower NX][x upper NX][h rows] */
#define NCB (2 * NZ)
static int g_nc = -1;                 /* entries per stage */
static int g_hrow[2 * NH + 1];        /* h index of general entry e-NCB */
static REAL g_hsgn[2 * NH + 1];       /* +1 lower, -1 upper */
static REAL g_hbnd[2 * NH + 1];       /* bound value */
#define NCMAX (NCB + 2 * NH)

static void setup_constraints(void)
{
    if (g_nc >= 0) return;
    int e = 0;
    for (int i = 0; i < NH; i++) {
        if (model_lh[i] > -BOUND_INF) { g_hrow[e] = i; g_hsgn[e] = 1.0; g_hbnd[e] = model_lh[i]; e++; }
        if (model_uh[i] < BOUND_INF) { g_hrow[e] = i; g_hsgn[e] = -1.0; g_hbnd[e] = model_uh[i]; e++; }
    }
    g_nc = NCB + e;
}

#ifdef __cplusplus
extern "C" {
#endif
int oracle_nc(void) { setup_constraints(); return g_nc; }
int oracle_dims(int *N, int *nx, int *nu, int *np, int *nh)
{
    *N = NN; *nx = NX; *nu = NU; *np = NP; *nh = NH;
    return 0;
}
/* doubles per problem in the persistent memory blob */
int oracle_mem_doubles(void) { setup_constraints(); return 1 + (NN + 1) * NX + 2 * NN * g_nc + (NN + 1) * NZ; }
#ifdef __cplusplus
}
#endif

typedef struct {
    /* NLP iterate and multipliers */
    REAL x[NN + 1][NX], u[NN][NU];
    REAL pi[NN + 1][NX];             /* pi[k]: multiplier of x_k = Phi(x_{k-1},u_{k-1}), k >= 1 */
    REAL lam[NN][NCMAX], t[NN][NCMAX];
    /* QP data */
    REAL W[NN][NX * NZ], b[NN][NX];
    REAL H[NN + 1][NZ * NZ], g[NN + 1][NZ];
    REAL C[NN][(NH > 0 ? NH : 1) * NZ], hval[NN][NH > 0 ? NH : 1];
    REAL d[NN][NCMAX];               /* signed bound: chat'v >= d */
    /* QP iterate */
    REAL v[NN + 1][NZ], qpi[NN + 1][NX];
    /* IPM work */
    REAL rg[NN + 1][NZ], rg0[NN + 1][NZ], rb[NN][NX], rd[NN][NCMAX], rm[NN][NCMAX];
    REAL Ht[NN + 1][NZ * NZ], gt[NN + 1][NZ];
    REAL P[NN + 1][NX * NX], pv[NN + 1][NX], Lx0[NN][NX], Lx1[NN][NX], L10[NN], iL[NN][NU], lv[NN][NU];
    REAL dv[NN + 1][NZ], dpi[NN + 1][NX], dlam[NN][NCMAX], dt[NN][NCMAX];
    REAL dva[NN + 1][NZ];
#ifdef MPC_HPIPM_BALANCE
    REAL dt_aff[NN][NCMAX], dlam_aff[NN][NCMAX];
#endif
    int qp_warm;                     /* previous QP solution available (HPIPM warm start 2) */
    int ipm_iters_total, qp_iters_last;
} work_t;

/* ------------------------------------------------------------------------------------------------
 * K1: ERK4 with forward sensitivities and adjoint-weighted second-order term  [upstream: acados ERK]
 * ---------------------------------------------------------------------------------------------- */
/* one RK4 step h from (x,u): xn, S = d xn / d[u;x] (NX x NZ); if lam != NULL also
 * Hc (NZ x NZ) = sum_j lam_j d2 xn_j / d[u;x]^2                                               */
static void rk4_step(const REAL *x, const REAL *u, const REAL *p, REAL h, REAL *xn, REAL *S, const REAL *lam, REAL *Hc)
{
    static const double ac[4] = {0.0, 0.5, 0.5, 1.0};
    static const double bc[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
    REAL X[4][NX], Kf[4][NX], Jf[4][NX * NZ], Z[4][NZ * NZ], dK[NX * NZ];
    for (int i = 0; i < NX; i++) xn[i] = x[i];
    for (int i = 0; i < NX * NZ; i++) S[i] = 0.0;
    for (int i = 0; i < NX; i++) S[i * NZ + NU + i] = 1.0;
    for (int s = 0; s < 4; s++) {
        /* stage state and its tangent Z = d[u;X_s]/dz */
        for (int i = 0; i < NZ * NZ; i++) Z[s][i] = 0.0;
        for (int i = 0; i < NU; i++) Z[s][i * NZ + i] = 1.0;
        for (int i = 0; i < NX; i++) {
            X[s][i] = x[i];
            Z[s][(NU + i) * NZ + NU + i] = 1.0;
        }
        if (s > 0) {
            for (int i = 0; i < NX; i++) {
                X[s][i] += ac[s] * h * Kf[s - 1][i];
                for (int j = 0; j < NZ; j++) Z[s][(NU + i) * NZ + j] += ac[s] * h * dK[i * NZ + j];
            }
        }
        model_f(X[s], u, p, Kf[s]);
        model_f_jac(X[s], u, p, Jf[s]);
        for (int i = 0; i < NX; i++)
            for (int j = 0; j < NZ; j++) {
                REAL a = 0.0;
                for (int k = 0; k < NZ; k++) a += Jf[s][i * NZ + k] * Z[s][k * NZ + j];
                dK[i * NZ + j] = a;
                S[i * NZ + j] += bc[s] * h * a;
            }
        for (int i = 0; i < NX; i++) xn[i] += bc[s] * h * Kf[s][i];
    }
    if (lam) {
        REAL nu_[4][NX], Hf[NZ * NZ], T[NZ * NZ];
        for (int i = 0; i < NZ * NZ; i++) Hc[i] = 0.0;
        for (int s = 3; s >= 0; s--) {
            for (int i = 0; i < NX; i++) {
                nu_[s][i] = bc[s] * h * lam[i];
                if (s < 3) {
                    REAL a = 0.0; /* a_{s+1} h Jfx_{s+1}' nu_{s+1} */
                    for (int k = 0; k < NX; k++) a += Jf[s + 1][k * NZ + NU + i] * nu_[s + 1][k];
                    nu_[s][i] += ac[s + 1] * h * a;
                }
            }
            model_f_hess(X[s], u, p, nu_[s], Hf);
            for (int i = 0; i < NZ; i++)
                for (int j = 0; j < NZ; j++) {
                    REAL a = 0.0;
                    for (int k = 0; k < NZ; k++) a += Hf[i * NZ + k] * Z[s][k * NZ + j];
                    T[i * NZ + j] = a;
                }
            for (int i = 0; i < NZ; i++)
                for (int j = 0; j < NZ; j++) {
                    REAL a = 0.0;
                    for (int k = 0; k < NZ; k++) a += Z[s][k * NZ + i] * T[k * NZ + j];
                    Hc[i * NZ + j] += a;
                }
        }
    }
}

/* SIM_STEPS RK4 steps over one shooting interval: xn = Phi(x,u), W = dPhi/d[u;x];
 * if pi != NULL, Hc = sum_j pi_j d2 Phi_j                                                        */
static void integrate(const REAL *x, const REAL *u, const REAL *p, REAL *xn, REAL *W, const REAL *pi, REAL *Hc)
{
    const REAL h = MODEL_DT / SIM_STEPS;
    REAL y[SIM_STEPS + 1][NX], S[SIM_STEPS][NX * NZ], T[SIM_STEPS + 1][NZ * NZ], dummy[NZ * NZ];
    for (int i = 0; i < NX; i++) y[0][i] = x[i];
    for (int i = 0; i < NZ * NZ; i++) T[0][i] = 0.0;
    for (int i = 0; i < NZ; i++) T[0][i * NZ + i] = 1.0;
    for (int s = 0; s < SIM_STEPS; s++) {
        rk4_step(y[s], u, p, h, y[s + 1], S[s], NULL, dummy);
        for (int i = 0; i < NZ * NZ; i++) T[s + 1][i] = 0.0;
        for (int i = 0; i < NU; i++) T[s + 1][i * NZ + i] = 1.0;
        for (int i = 0; i < NX; i++)
            for (int j = 0; j < NZ; j++) {
                REAL a = 0.0;
                for (int k = 0; k < NZ; k++) a += S[s][i * NZ + k] * T[s][k * NZ + j];
                T[s + 1][(NU + i) * NZ + j] = a;
            }
    }
    for (int i = 0; i < NX; i++) xn[i] = y[SIM_STEPS][i];
    if (W)
        for (int i = 0; i < NX; i++)
            for (int j = 0; j < NZ; j++) W[i * NZ + j] = T[SIM_STEPS][(NU + i) * NZ + j];
    if (pi) {
        REAL lam[NX], lamp[NX], Hs[NZ * NZ], tmp[NX], St[NX * NZ], M[NZ * NZ];
        for (int i = 0; i < NX; i++) lam[i] = pi[i];
        for (int i = 0; i < NZ * NZ; i++) Hc[i] = 0.0;
        for (int s = SIM_STEPS - 1; s >= 0; s--) {
            rk4_step(y[s], u, p, h, tmp, St, lam, Hs);
            for (int i = 0; i < NZ; i++)
                for (int j = 0; j < NZ; j++) {
                    REAL a = 0.0;
                    for (int k = 0; k < NZ; k++) a += Hs[i * NZ + k] * T[s][k * NZ + j];
                    M[i * NZ + j] = a;
                }
            for (int i = 0; i < NZ; i++)
                for (int j = 0; j < NZ; j++) {
                    REAL a = 0.0;
                    for (int k = 0; k < NZ; k++) a += T[s][k * NZ + i] * M[k * NZ + j];
                    Hc[i * NZ + j] += a;
                }
            for (int i = 0; i < NX; i++) {
                REAL a = 0.0;
                for (int k = 0; k < NX; k++) a += St[k * NZ + NU + i] * lam[k];
                lamp[i] = a;
            }
            for (int i = 0; i < NX; i++) lam[i] = lamp[i];
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * K4: MIRROR regularisation  [upstream: acados regularize_mirror; eigen-decomposition restated as
 * cyclic Jacobi]  A (n x n, row-major, leading dimension ld) <- V max(|lambda|, eps) V'
 * ---------------------------------------------------------------------------------------------- */
static void mirror(REAL *A, int n, int ld)
{
    /* Pair order: round-robin tournament on np = n + (n odd) positions (an odd n gets a decoupled dummy
     * position): every round rotates the position pairs (0,np-1) (1,np-2) ... and then shifts positions
     * 1..np-1 cyclically; after np-1 rounds (one sweep) the arrangement is back to the identity.
     * Rotation: ir = 1/sqrt(tau^2 + 4 a_pq^2), cos^2 = (1 + |tau| ir)/2, c = sqrt(cos^2),
     * s = sign(tau) a_pq ir / c, t = s / c  (tau = a_qq - a_pp).                                    */
    enum { MAXP = NZ + 1 };
    const int np = n + (n & 1);
    REAL a[MAXP][MAXP], V[NZ][MAXP], ev[NZ];
    int pos[MAXP];                       /* pos[j] = variable sitting at position j (np-1 may be the dummy) */
    for (int i = 0; i < np; i++) {
        pos[i] = i;
        for (int j = 0; j < np; j++) a[i][j] = (i < n && j < n) ? 0.5 * (A[i * ld + j] + A[j * ld + i]) : 0.0;
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < np; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < JACOBI_MAX_SWEEPS; sweep++) {
        REAL off = 0.0, dia = 0.0;
        for (int i = 0; i < n; i++)
            for (int j = 0; j <= i; j++) {
                if (i == j) dia += a[i][j] * a[i][j];
                else off += a[i][j] * a[i][j];
            }
        off *= 2.0;
        if (!(off > JACOBI_TOL * (off + dia))) break;
        for (int round = 0; round < np - 1; round++) {
            for (int pr = 0; pr < np / 2; pr++) {
                const int p = pos[pr], q = pos[np - 1 - pr];   /* variables (matrix kept in variable order) */
                REAL apq = a[q][p], q2 = apq * apq;
                if (!(q2 > 0.0)) continue;
                /* orientation as in the kernel: "p" is the lower position, "q" the higher one */
                REAL tau = a[q][q] - a[p][p];
                REAL ir = 1.0 / sqrt(tau * tau + 4.0 * q2);
                REAL c2 = 0.5 + 0.5 * fabs(tau) * ir;
                REAL ic = 1.0 / sqrt(c2), c = c2 * ic;
                REAL sn = (tau >= 0.0 ? apq : -apq) * ir * ic, tt = sn * ic;
                for (int k = 0; k < np; k++) {
                    if (k == p || k == q) continue;
                    REAL akp = a[k][p], akq = a[k][q];
                    a[k][p] = a[p][k] = c * akp - sn * akq;
                    a[k][q] = a[q][k] = sn * akp + c * akq;
                }
                a[p][p] -= tt * apq;
                a[q][q] += tt * apq;
                a[p][q] = a[q][p] = 0.0;
                for (int k = 0; k < n; k++) {
                    REAL vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - sn * vkq;
                    V[k][q] = sn * vkp + c * vkq;
                }
            }
            int last = pos[np - 1];
            for (int j = np - 1; j >= 2; j--) pos[j] = pos[j - 1];
            pos[1] = last;
        }
    }
    for (int i = 0; i < n; i++) {
        REAL e = a[i][i];
        if (e >= -REG_EPS && e <= REG_EPS) e = REG_EPS;
        else if (e < 0.0) e = -e;
        ev[i] = e;
    }
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            REAL s = 0.0;
            for (int k = 0; k < n; k++) s += V[i][k] * ev[k] * V[j][k];
            A[i * ld + j] = s;
        }
}

/* ------------------------------------------------------------------------------------------------
 * K1-K4 for all stages: build the QP at the current iterate
 * ---------------------------------------------------------------------------------------------- */
static const REAL *stage_params(const REAL *params, int k) { return params + (size_t)(k < NN ? k : NN - 1) * NP; }

static void linearize(work_t *w, const REAL *xinit, const REAL *params)
{
    for (int k = 0; k < NN; k++) {
        const REAL *p = stage_params(params, k);
        REAL z[NZ], xn[NX], Hd[NZ * NZ], gc[NZ], Hl[NZ * NZ], Hh[NZ * NZ], m[NH > 0 ? NH : 1];
        for (int i = 0; i < NU; i++) z[i] = w->u[k][i];
        for (int i = 0; i < NX; i++) z[NU + i] = w->x[k][i];
        integrate(w->x[k], w->u[k], p, xn, w->W[k], w->pi[k + 1], Hd);
        for (int i = 0; i < NX; i++) w->b[k][i] = xn[i] - w->x[k + 1][i];
        model_cost_grad_hess(z, p, gc, Hl);
        for (int e = 0; e < NH; e++) m[e] = 0.0;
        for (int e = NCB; e < g_nc; e++) m[g_hrow[e - NCB]] -= g_hsgn[e - NCB] * w->lam[k][e]; /* lam_u - lam_l */
        model_h_hess(z, p, m, Hh);
        for (int i = 0; i < NZ; i++) w->g[k][i] = MODEL_DT * gc[i];
        for (int i = 0; i < NZ * NZ; i++) w->H[k][i] = MODEL_DT * Hl[i] + Hd[i] + Hh[i];
        mirror(w->H[k], NZ, NZ);
        model_h(z, p, w->hval[k]);
        model_h_jac(z, p, w->C[k]);
        /* signed bounds: box on u (all k), on x (k >= 1), general rows */
        for (int i = 0; i < NZ; i++) {
            w->d[k][i] = model_lbz[i] - z[i];
            w->d[k][NZ + i] = -(model_ubz[i] - z[i]);
        }
        for (int e = NCB; e < g_nc; e++) {
            int r = g_hrow[e - NCB];
            w->d[k][e] = g_hsgn[e - NCB] * (g_hbnd[e - NCB] - w->hval[k][r]);
        }
    }
    /* terminal stage: no cost, no constraints => H_N = mirror(0) = eps I, g_N = 0 */
    for (int i = 0; i < NZ * NZ; i++) w->H[NN][i] = 0.0;
    for (int i = 0; i < NZ; i++) w->g[NN][i] = 0.0;
    mirror(&w->H[NN][NU * NZ + NU], NX, NZ);
    (void)xinit;
}

#ifdef MPC_HPIPM_BALANCE
long g_cond_pc_triggers, g_cond_pc_checks;      /* diagnostics of the conditional predictor-corrector (not thread safe: run with 1 thread) */
#endif
/* entry e of stage k is active? (x box rows are absent at k = 0: x_0 is fixed) */
static int active(int k, int e)
{
    if (k == 0 && ((e >= NU && e < NZ) || (e >= NZ + NU && e < NCB))) return 0;
    return 1;
}
/* chat_e' y for a 7-vector y */
static REAL crow_dot(const work_t *w, int k, int e, const REAL *y)
{
    if (e < NZ) return y[e];
    if (e < NCB) return -y[e - NZ];
    int r = g_hrow[e - NCB];
    REAL s = 0.0;
    for (int j = 0; j < NZ; j++) s += w->C[k][r * NZ + j] * y[j];
    return g_hsgn[e - NCB] * s;
}
/* y += a * chat_e */
static void crow_axpy(const work_t *w, int k, int e, REAL a, REAL *y)
{
    if (e < NZ) { y[e] += a; return; }
    if (e < NCB) { y[e - NZ] -= a; return; }
    int r = g_hrow[e - NCB];
    for (int j = 0; j < NZ; j++) y[j] += a * g_hsgn[e - NCB] * w->C[k][r * NZ + j];
}

/* ------------------------------------------------------------------------------------------------
 * K5: primal-dual interior point QP with Riccati  [upstream: HPIPM ocp_qp_ipm, restated]
 * ---------------------------------------------------------------------------------------------- */
static void qp_init(work_t *w, const REAL *dx0)
{
    if (!w->qp_warm) {
        memset(w->v, 0, sizeof(w->v));
        memset(w->qpi, 0, sizeof(w->qpi));
    }
    for (int i = 0; i < NX; i++) w->v[0][NU + i] = dx0[i];
    if (w->qp_warm) {
        for (int k = 0; k < NN; k++)
            for (int e = 0; e < g_nc; e++) {
                if (!active(k, e)) continue;
                if (w->lam[k][e] < IPM_THR0) w->lam[k][e] = IPM_THR0;
                if (w->t[k][e] < IPM_THR0) w->t[k][e] = IPM_THR0;
            }
        return;
    }
    for (int k = 0; k < NN; k++) {
        for (int i = 0; i < NZ; i++) {
            if (!active(k, i)) continue;
            REAL dl = w->d[k][i], du = -w->d[k][NZ + i];
            REAL tl = w->v[k][i] - dl, tu = du - w->v[k][i];
            if (tl < IPM_THR0) {
                if (tu < IPM_THR0) { w->v[k][i] = 0.5 * (dl + du); tl = IPM_THR0; tu = IPM_THR0; }
                else { tl = IPM_THR0; w->v[k][i] = dl + IPM_THR0; }
            } else if (tu < IPM_THR0) { tu = IPM_THR0; w->v[k][i] = du - IPM_THR0; }
            w->t[k][i] = tl; w->t[k][NZ + i] = tu;
            w->lam[k][i] = IPM_MU0 / tl; w->lam[k][NZ + i] = IPM_MU0 / tu;
        }
        for (int e = NCB; e < g_nc; e++) {
            REAL tt = crow_dot(w, k, e, w->v[k]) - w->d[k][e];
            if (tt < IPM_THR0) tt = IPM_THR0;
            w->t[k][e] = tt;
            w->lam[k][e] = IPM_MU0 / tt;
        }
    }
}

/* max that propagates NaN: a NaN anywhere must surface as QP status 3 */
static REAL nanmax(REAL a, REAL b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }

static void qp_residuals(work_t *w, REAL nrm[4], REAL *mu)
{
    REAL ng = 0.0, nb = 0.0, nd = 0.0, nm = 0.0, sm = 0.0;
    int cnt = 0;
    for (int k = 0; k <= NN; k++) {
        REAL *r = w->rg[k];
        int j0 = (k == NN) ? NU : 0;
        for (int i = 0; i < NZ; i++) r[i] = 0.0;
        for (int i = j0; i < NZ; i++) {
            REAL s = w->g[k][i];
            for (int j = j0; j < NZ; j++) s += w->H[k][i * NZ + j] * w->v[k][j];
            r[i] = s;
        }
        if (k < NN)
            for (int j = 0; j < NZ; j++) {
                REAL s = 0.0;
                for (int i = 0; i < NX; i++) s += w->W[k][i * NZ + j] * w->qpi[k + 1][i];
                r[j] += s;
            }
        if (k > 0)
            for (int i = 0; i < NX; i++) r[NU + i] -= w->qpi[k][i];
        for (int i = 0; i < NZ; i++) w->rg0[k][i] = r[i];
        if (k < NN)
            for (int e = 0; e < g_nc; e++)
                if (active(k, e)) crow_axpy(w, k, e, -w->lam[k][e], r);
        if (k == 0)
            for (int i = NU; i < NZ; i++) r[i] = 0.0; /* x_0 is not a variable */
        for (int i = 0; i < NZ; i++) ng = nanmax(ng, fabs(r[i]));
    }
    for (int k = 0; k < NN; k++) {
        for (int i = 0; i < NX; i++) {
            REAL s = w->b[k][i] - w->v[k + 1][NU + i];
            for (int j = 0; j < NZ; j++) s += w->W[k][i * NZ + j] * w->v[k][j];
            w->rb[k][i] = s;
            nb = nanmax(nb, fabs(s));
        }
        for (int e = 0; e < g_nc; e++) {
            if (!active(k, e)) { w->rd[k][e] = 0.0; w->rm[k][e] = 0.0; continue; }
            REAL s = crow_dot(w, k, e, w->v[k]) - w->d[k][e] - w->t[k][e];
            REAL m = w->lam[k][e] * w->t[k][e];
            w->rd[k][e] = s; w->rm[k][e] = m;
            nd = nanmax(nd, fabs(s));
            nm = nanmax(nm, fabs(m));
            sm += m; cnt++;
        }
    }
    nrm[0] = ng; nrm[1] = nb; nrm[2] = nd; nrm[3] = nm;
    *mu = sm / (REAL)cnt;
}

/* Htilde = H + sum Gamma chat chat'  (only depends on lam, t) */
static void kkt_hessian(work_t *w)
{
    for (int k = 0; k <= NN; k++) {
        memcpy(w->Ht[k], w->H[k], sizeof(w->H[k]));
        if (k == NN) break;
        for (int e = 0; e < g_nc; e++) {
            if (!active(k, e)) continue;
            REAL G = w->lam[k][e] * (1.0 / w->t[k][e]);
            if (e < NCB) { int i = (e < NZ) ? e : e - NZ; w->Ht[k][i * NZ + i] += G; }
            else {
                int r = g_hrow[e - NCB];
                for (int i = 0; i < NZ; i++)
                    for (int j = 0; j < NZ; j++) w->Ht[k][i * NZ + j] += G * w->C[k][r * NZ + i] * w->C[k][r * NZ + j];
            }
        }
    }
}
/* gtilde = rg + sum chat gamma,  gamma = (rm + lam rd)/t.  With rm = lam t (+ corrector term) the lam
 * part cancels against the -chat lam inside rg, so (rg0 = residual without the inequality multipliers)
 *   gtilde = rg0 + sum chat (Gamma rd + corr),   Gamma = lam/t,
 *   corr = 0 (predictor)  or  (dt_aff dlam_aff - sigma mu)/t (corrector).                          */
static void kkt_gradient(work_t *w, int corrector, REAL sigmu)
{
    for (int k = 0; k <= NN; k++) {
        for (int i = 0; i < NZ; i++) w->gt[k][i] = w->rg0[k][i];
        if (k == NN) break;
        for (int e = 0; e < g_nc; e++) {
            if (!active(k, e)) continue;
            REAL invt = 1.0 / w->t[k][e];
            REAL gam = w->lam[k][e] * invt * w->rd[k][e];
            if (corrector == 1) gam += w->dt[k][e] * w->dlam[k][e] * invt - sigmu * invt;
            else if (corrector == 2) gam += -sigmu * invt;      /* centering direction only (conditional predictor-corrector) */
            crow_axpy(w, k, e, gam, w->gt[k]);
        }
    }
}

/* Backward Riccati factorisation, square-root form on the input block: G_k = Ht_k + W_k' P_{k+1} W_k,
 * two Cholesky pivots give [Luu 0; Lxu I], and P_k = Gxx - Lxu Lxu' is the trailing update of the
 * Cholesky factorisation (backward stable; no explicit inverse of Guu).                          */
static void riccati_factor(work_t *w)
{
#if NU != 2
#error "riccati_factor assumes NU == 2"
#endif
    for (int i = 0; i < NX; i++)
        for (int j = 0; j < NX; j++) w->P[NN][i * NX + j] = w->Ht[NN][(NU + i) * NZ + NU + j];
    for (int k = NN - 1; k >= 0; k--) {
        REAL PW[NX * NZ], G[NZ * NZ];
        for (int i = 0; i < NX; i++)
            for (int j = 0; j < NZ; j++) {
                REAL s = 0.0;
                for (int l = 0; l < NX; l++) s += w->P[k + 1][i * NX + l] * w->W[k][l * NZ + j];
                PW[i * NZ + j] = s;
            }
        for (int i = 0; i < NZ; i++)
            for (int j = 0; j <= i; j++) {
                REAL s = w->Ht[k][i * NZ + j];
                for (int l = 0; l < NX; l++) s += w->W[k][l * NZ + i] * PW[l * NZ + j];
                G[i * NZ + j] = s;
            }
        REAL i0 = 1.0 / sqrt(G[0]);
        REAL l10 = G[NZ] * i0;
        REAL i1 = 1.0 / sqrt(G[NZ + 1] - l10 * l10);
        w->iL[k][0] = i0; w->iL[k][1] = i1; w->L10[k] = l10;
        for (int i = 0; i < NX; i++) {
            w->Lx0[k][i] = G[(NU + i) * NZ] * i0;
            w->Lx1[k][i] = (G[(NU + i) * NZ + 1] - w->Lx0[k][i] * l10) * i1;
        }
        for (int i = 0; i < NX; i++)
            for (int j = 0; j <= i; j++) {
                REAL s = G[(NU + i) * NZ + NU + j] - w->Lx0[k][i] * w->Lx0[k][j] - w->Lx1[k][i] * w->Lx1[k][j];
                w->P[k][i * NX + j] = s; w->P[k][j * NX + i] = s;
            }
    }
}

/* backward vector sweep + forward sweep with rhs (gt, rb): Newton step dv, dpi.
 *   l = Luu^-1 q_u ; p = q_x - Lxu l ; du = -Luu^-T (Lxu' dx + l)                                  */
static void riccati_solve(work_t *w, REAL (*dv)[NZ])
{
#if NU != 2
#error "riccati_solve assumes NU == 2"
#endif
    for (int i = 0; i < NX; i++) w->pv[NN][i] = w->gt[NN][NU + i];
    for (int k = NN - 1; k >= 0; k--) {
        REAL y[NX], q[NZ];
        for (int i = 0; i < NX; i++) {
            REAL s = 0.0;
            for (int l = 0; l < NX; l++) s += w->P[k + 1][i * NX + l] * w->rb[k][l];
            y[i] = w->pv[k + 1][i] + s;
        }
        for (int j = 0; j < NZ; j++) {
            REAL s = w->gt[k][j];
            for (int i = 0; i < NX; i++) s += w->W[k][i * NZ + j] * y[i];
            q[j] = s;
        }
        w->lv[k][0] = q[0] * w->iL[k][0];
        w->lv[k][1] = (q[1] - w->L10[k] * w->lv[k][0]) * w->iL[k][1];
        for (int i = 0; i < NX; i++)
            w->pv[k][i] = q[NU + i] - w->Lx0[k][i] * w->lv[k][0] - w->Lx1[k][i] * w->lv[k][1];
    }
    for (int i = 0; i < NX; i++) dv[0][NU + i] = 0.0; /* dx_0 = 0: x_0 is fixed */
    for (int k = 0; k < NN; k++) {
        REAL r0 = w->lv[k][0], r1 = w->lv[k][1];
        for (int j = 0; j < NX; j++) { r0 += w->Lx0[k][j] * dv[k][NU + j]; r1 += w->Lx1[k][j] * dv[k][NU + j]; }
        dv[k][1] = -r1 * w->iL[k][1];
        dv[k][0] = -(r0 + w->L10[k] * dv[k][1]) * w->iL[k][0];
        for (int i = 0; i < NX; i++) {
            REAL s = w->rb[k][i];
            for (int j = 0; j < NZ; j++) s += w->W[k][i * NZ + j] * dv[k][j];
            dv[k + 1][NU + i] = s;
        }
        for (int i = 0; i < NX; i++) {
            REAL s = w->pv[k + 1][i];
            for (int j = 0; j < NX; j++) s += w->P[k + 1][i * NX + j] * dv[k + 1][NU + j];
            w->dpi[k + 1][i] = s;
        }
    }
    for (int i = 0; i < NU; i++) dv[NN][i] = 0.0;
}

/* dt, dlam from the Newton direction; returns the max step alpha in (0,1] keeping lam, t >= 0.
 *   dt = chat'dv + rd ;  dlam = -(rm_c + lam dt)/t  with rm_c = lam t (+ dt_aff dlam_aff - sigma mu)
 *      = -(lam + Gamma dt [+ (dt_aff dlam_aff - sigma mu)/t])
 * corrector != 0: w->dt/w->dlam hold the affine step on entry and are overwritten.                 */
/* ratio test without a division per entry: the running minimum of val/(-dval) over entries with
 * dval < 0 is kept as a fraction bn/bd (bd > 0) and compared by cross-multiplication */
static void step_limit(REAL val, REAL dval, REAL *bn, REAL *bd)
{
    if (dval < 0.0 && val * *bd < *bn * (-dval)) { *bn = val; *bd = -dval; }
}
static REAL ipm_step_ineq(work_t *w, REAL (*dv)[NZ], int corrector, REAL sigmu)
{
    REAL alpha = 1.0;
    for (int k = 0; k < NN; k++) {       /* per-stage fraction, then min over stages (as the warp does) */
        REAL bn = 1.0, bd = 1.0;
        for (int e = 0; e < g_nc; e++) {
            if (!active(k, e)) { w->dt[k][e] = 0.0; w->dlam[k][e] = 0.0; continue; }
            REAL lam = w->lam[k][e], t = w->t[k][e], invt = 1.0 / t;
            REAL dt = crow_dot(w, k, e, dv[k]) + w->rd[k][e];
            REAL dl;
            if (corrector == 1) {
                REAL corr = w->dt[k][e] * w->dlam[k][e] * invt;
                dl = -(lam + lam * invt * dt + (corr - sigmu * invt));
            } else if (corrector == 2) {
                dl = -(lam + lam * invt * dt - sigmu * invt);
            } else {
                dl = -(lam + lam * invt * dt);
            }
            w->dt[k][e] = dt; w->dlam[k][e] = dl;
            step_limit(lam, dl, &bn, &bd);
            step_limit(t, dt, &bn, &bd);
        }
        REAL a = bn / bd;
        if (a < alpha) alpha = a;
    }
    return alpha;
}

/* returns HPIPM-style status: 0 ok, 1 max iter, 2 min step, 3 NaN */
static int qp_solve(work_t *w, const REAL *dx0)
{
    REAL nrm[4], mu, alpha = 1.0;
    int kk;
    qp_init(w, dx0);
    qp_residuals(w, nrm, &mu);
#ifdef ORACLE_TRACE
    printf("cpu ipm 0: rg %.3e rb %.3e rd %.3e rm %.3e mu %.3e alpha %.3e\n", (double)nrm[0], (double)nrm[1], (double)nrm[2], (double)nrm[3], (double)mu, (double)alpha);
#endif
#define QP_ISNAN() ((mu != mu) || (nrm[0] != nrm[0]) || (nrm[1] != nrm[1]) || (nrm[2] != nrm[2]) || (nrm[3] != nrm[3]))
    for (kk = 0; kk < IPM_ITER_MAX && alpha > IPM_ALPHA_MIN && !QP_ISNAN() &&
                 (nrm[0] > IPM_TOL || nrm[1] > IPM_TOL || nrm[2] > IPM_TOL || nrm[3] > IPM_TOL); kk++) {
        /* predictor (affine scaling) */
        kkt_hessian(w);
        kkt_gradient(w, 0, 0.0);
        riccati_factor(w);
        riccati_solve(w, w->dva);
        REAL alpha_aff = ipm_step_ineq(w, w->dva, 0, 0.0);
        /* mu_aff = sum (lam + a dlam)(t + a dt) / count, expanded in powers of a */
        REAL S1 = 0.0, S2 = 0.0;
        for (int k = 0; k < NN; k++)
            for (int e = 0; e < g_nc; e++)
                if (active(k, e)) {
                    S1 += w->lam[k][e] * w->dt[k][e] + w->t[k][e] * w->dlam[k][e];
                    S2 += w->dt[k][e] * w->dlam[k][e];
                }
        REAL cnt = (REAL)(NN * (2 * NU + (g_nc - NCB)) + (NN - 1) * 2 * NX);
        REAL mu_aff = (mu * cnt + alpha_aff * S1 + alpha_aff * alpha_aff * S2) / cnt;
        REAL rat = mu_aff / mu, sigmu = rat * rat * rat * mu;
        /* corrector: rm += dt_aff dlam_aff - sigma mu (folded into gtilde and dlam) */
        kkt_gradient(w, 1, sigmu);
        riccati_solve(w, w->dv);
#ifdef MPC_HPIPM_BALANCE
        /* [upstream, from memory of hpipm ocp_qp_ipm.c: mode BALANCE has cond_pred_corr = 1] conditional Mehrotra predictor-
         * corrector: if the corrected step would leave the duality measure above twice the predictor's, fall back to the
         * pure centering direction (complementarity right-hand side lam t - sigma mu, no second-order term).  The affine
         * dt / dlam are needed again by the fallback, so they are saved first.  BALANCE's two iterative-refinement steps
         * (itref_corr_max = 2) only run while the residual of the Newton system exceeds the exit tolerances (1e-5 here): the
         * Riccati solve leaves ~1e-12 (tests/test_oracle_solve.py checks the QP KKT conditions to 1e-10), so they never start. */
        memcpy(w->dt_aff, w->dt, sizeof(w->dt)); memcpy(w->dlam_aff, w->dlam, sizeof(w->dlam));
#endif
        alpha = ipm_step_ineq(w, w->dv, 1, sigmu);
#ifdef MPC_HPIPM_BALANCE
        {
            REAL T1 = 0.0, T2 = 0.0;
            for (int k = 0; k < NN; k++)
                for (int e = 0; e < g_nc; e++)
                    if (active(k, e)) {
                        T1 += w->lam[k][e] * w->dt[k][e] + w->t[k][e] * w->dlam[k][e];
                        T2 += w->dt[k][e] * w->dlam[k][e];
                    }
            REAL mu_pc = (mu * cnt + alpha * T1 + alpha * alpha * T2) / cnt;
            if (mu_pc > 2.0 * mu_aff) {
                g_cond_pc_triggers++;
                memcpy(w->dt, w->dt_aff, sizeof(w->dt)); memcpy(w->dlam, w->dlam_aff, sizeof(w->dlam));
                kkt_gradient(w, 2, sigmu);
                riccati_solve(w, w->dv);
                alpha = ipm_step_ineq(w, w->dv, 2, sigmu);
            }
            g_cond_pc_checks++;
        }
#endif
        REAL a = alpha < 1.0 ? alpha * IPM_STEP_SCALE : alpha;
        for (int k = 0; k <= NN; k++) {
            for (int i = 0; i < NZ; i++) w->v[k][i] += a * w->dv[k][i];
            if (k > 0) for (int i = 0; i < NX; i++) w->qpi[k][i] += a * w->dpi[k][i];
            if (k < NN)
                for (int e = 0; e < g_nc; e++) {
                    if (!active(k, e)) continue;
                    w->lam[k][e] += a * w->dlam[k][e];
                    w->t[k][e] += a * w->dt[k][e];
                    if (w->lam[k][e] < IPM_LAM_MIN) w->lam[k][e] = IPM_LAM_MIN;
                    if (w->t[k][e] < IPM_T_MIN) w->t[k][e] = IPM_T_MIN;
                }
        }
        qp_residuals(w, nrm, &mu);
#ifdef ORACLE_TRACE
        printf("cpu ipm %d: rg %.3e rb %.3e rd %.3e rm %.3e mu %.3e alpha %.3e\n", kk + 1, (double)nrm[0], (double)nrm[1], (double)nrm[2], (double)nrm[3], (double)mu, (double)alpha);
#endif
    }
    w->qp_iters_last = kk;
    w->ipm_iters_total += kk;
    if (QP_ISNAN()) return 3;
    if (kk == IPM_ITER_MAX) return 1;
    if (alpha <= IPM_ALPHA_MIN) return 2;
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * One Solver::solve() call   (acados_solver_interface.cpp:86-204)
 * ---------------------------------------------------------------------------------------------- */
static void solve_one(work_t *w, const REAL *xinit, const REAL *x0, const REAL *params, int num_iter, REAL *mem,
                      REAL *xtraj, REAL *utraj, REAL *pobj, int *exit_code, int *qp_status, REAL *res_eq, int *ipm_iters)
{
    /* num_iter < 0: |num_iter| iterations, completion deferred (stepwise interface, acados_solver_interface.cpp:145-160:
     * no res_eq demotion, no reset on failure -- both belong to completeOneIteration, :176-191) */
    const int defer = num_iter < 0;
    if (defer) num_iter = -num_iter;
    /* persistent solver memory (multipliers survive between solve() calls on one capsule) */
    memset(w->pi, 0, sizeof(w->pi));
    memset(w->lam, 0, sizeof(w->lam));
    memset(w->t, 0, sizeof(w->t));
    memset(w->v, 0, sizeof(w->v));
    w->qp_warm = 0;
    w->ipm_iters_total = 0;
    if (mem && mem[0] != 0.0) {
        const REAL *m = mem + 1;
        memcpy(w->pi, m, sizeof(REAL) * (NN + 1) * NX); m += (NN + 1) * NX;
        for (int k = 0; k < NN; k++) { memcpy(w->lam[k], m, sizeof(REAL) * g_nc); m += g_nc; }
        for (int k = 0; k < NN; k++) { memcpy(w->t[k], m, sizeof(REAL) * g_nc); m += g_nc; }
        memcpy(w->v, m, sizeof(REAL) * (NN + 1) * NZ);
        w->qp_warm = (mem[0] >= 2.0);
        memcpy(w->qpi, w->pi, sizeof(w->pi));
    }
    /* loadWarmstart: x0 = [u_k, x_k] per stage  (acados_solver_interface.cpp:274-284) */
    for (int k = 0; k <= NN; k++) {
        if (k < NN) for (int i = 0; i < NU; i++) w->u[k][i] = x0[k * NZ + i];
        for (int i = 0; i < NX; i++) w->x[k][i] = x0[k * NZ + NU + i];
    }
    int status = 0, qps = 0;
    for (int it = 0; it < num_iter; it++) {
        REAL dx0[NX];
        linearize(w, xinit, params);
        for (int i = 0; i < NX; i++) dx0[i] = xinit[i] - w->x[0][i];
        qps = qp_solve(w, dx0);
        if (qps != 0 && qps != 1) { status = 4; break; }       /* ACADOS_QP_FAILURE; iterate unchanged */
        for (int k = 0; k <= NN; k++) {                        /* full step (FIXED_STEP) */
            if (k < NN) for (int i = 0; i < NU; i++) w->u[k][i] += w->v[k][i];
            for (int i = 0; i < NX; i++) w->x[k][i] += w->v[k][NU + i];
        }
        memcpy(w->pi, w->qpi, sizeof(w->pi));
        w->qp_warm = 1;
        status = 0;
        if (qps != 0) break;                                   /* wrapper breaks on qp_status != 0 (:105-106) */
    }
    /* completeOneIteration (:162-204) */
    REAL cost = 0.0, req = 0.0;
    for (int k = 0; k < NN; k++) {
        const REAL *p = stage_params(params, k);
        REAL z[NZ], xn[NX];
        for (int i = 0; i < NU; i++) z[i] = w->u[k][i];
        for (int i = 0; i < NX; i++) z[NU + i] = w->x[k][i];
        cost += MODEL_DT * model_cost(z, p);
        integrate(w->x[k], w->u[k], p, xn, NULL, NULL, NULL);
        for (int i = 0; i < NX; i++) req = nanmax(req, fabs(xn[i] - w->x[k + 1][i]));
    }
    for (int k = 0; k <= NN; k++) {
        for (int i = 0; i < NX; i++) xtraj[k * NX + i] = w->x[k][i];
        if (k < NN) for (int i = 0; i < NU; i++) utraj[k * NU + i] = w->u[k][i];
    }
    if (!(req <= RES_EQ_MAX) && status == 0 && !defer) status = 4;
    *pobj = cost; *res_eq = req; *qp_status = (qps == 0) ? 0 : qps + 1;   /* acados numbering decoded at acados_solver_interface.cpp:409-420: 2 max-iter, 3 min-step, 4 NaN */
    *exit_code = (status == 0) ? 1 : (status == 1 ? 0 : status);
    if (ipm_iters) *ipm_iters = w->ipm_iters_total;
    if (mem) {
        if (status != 0) {
            if (!defer) memset(mem, 0, sizeof(REAL) * oracle_mem_doubles()); /* Solver_acados_reset + reset_qp_memory (:187-191) */
        } else {
            REAL *m = mem + 1;
            mem[0] = 2.0;
            memcpy(m, w->pi, sizeof(REAL) * (NN + 1) * NX); m += (NN + 1) * NX;
            for (int k = 0; k < NN; k++) { memcpy(m, w->lam[k], sizeof(REAL) * g_nc); m += g_nc; }
            for (int k = 0; k < NN; k++) { memcpy(m, w->t[k], sizeof(REAL) * g_nc); m += g_nc; }
            memcpy(m, w->v, sizeof(REAL) * (NN + 1) * NZ);
        }
    }
}

#ifndef ORACLE_NO_API
#ifdef __cplusplus
extern "C" {
#endif

/* Batched solve; OpenMP over problems mirrors guidance_constraints.cpp:304 (one capsule per thread). */
int oracle_solve_batch(int n, const double *xinit, const double *x0, const double *params, const int *num_iter,
                       double *mem_inout, double *xtraj, double *utraj, double *pobj, int *exit_code,
                       int *qp_status, double *res_eq, int *ipm_iters, int num_threads)
{
    setup_constraints();
    int md = oracle_mem_doubles();
#pragma omp parallel num_threads(num_threads > 0 ? num_threads : 1)
    {
        work_t *w = (work_t *)malloc(sizeof(work_t));
#pragma omp for schedule(dynamic, 1)
        for (int i = 0; i < n; i++)
            solve_one(w, xinit + (size_t)i * NX, x0 + (size_t)i * NZ * (NN + 1), params + (size_t)i * NN * NP, num_iter[i],
                      mem_inout ? mem_inout + (size_t)i * md : NULL, xtraj + (size_t)i * NX * (NN + 1),
                      utraj + (size_t)i * NU * NN, pobj + i, exit_code + i, qp_status + i, res_eq + i,
                      ipm_iters ? ipm_iters + i : NULL);
        free(w);
    }
    return 0;
}

/* K7: FindBestPlanner (guidance_constraints.cpp:572-590) with the objective post-processing of
 * :373-420: obj = (pobj - obj_sub) * obj_scale; success = exit_code == 1; strict <, ascending.   */
/* ---- guidance halfspaces -------------------------------------------------------------------------
 * Restates what GuidanceConstraints::optimize does per planner right before solve() for its topology constraints:
 * LinearizedConstraints::update + projectToSafety + setParameters with _use_guidance = true, one disc, radius 1e-3
 * (mpc_planner_modules/src/linearized_constraints.cpp:43-189; guidance_constraints.cpp:321-352).
 *   stage 0:           every slot is the dummy (1, 0, state.x + 100)                              (:155-166, :54)
 *   guided planner:    position = the warm start getEgoPrediction(k, x|y) (:63), pushed out of the obstacles by
 *                      projectToSafety (:130-148), a = (o - pos)/|o - pos|, b = a.o - (1e-3 + robot_radius) (:84-105)
 *                      with o = prediction.modes[0][k-1] of obstacle j; remaining slots dummies (:181-187)
 *   non-guided planner (update(state, empty_data_): no obstacles): dummies in every slot            (:326-329, :181-187)
 * projectToSafety calls RosTools::DouglasRachford::douglasRachfordProjection, which lives in the un-vendored, unpinned
 * `ros_tools` package (not in the repository): restated here as the published Douglas-Rachford step
 * p <- (p + R_obstacle(R_anchor(p))) / 2 with R = 2 Proj - I and Proj = radial projection onto the circle of radius r
 * around the centre (identity outside it), applied only when p is inside the obstacle's circle.  PARITY UNPINNED for
 * this sub-step; the halfspace formulas are the reference's own lines.  xinit_sets / obst_pred are per homotopy set. */
static void dr_proj(const double *p, const double *c, double r, const double *toward, double *out)
{
    const double dx = p[0] - c[0], dy = p[1] - c[1];
    if (sqrt(dx * dx + dy * dy) < r) {
        const double sx = toward[0] - c[0], sy = toward[1] - c[1], n = sqrt(sx * sx + sy * sy);
        out[0] = c[0] + sx / n * r; out[1] = c[1] + sy / n * r;
    } else { out[0] = p[0]; out[1] = p[1]; }
}
void oracle_guidance_halfspaces_static(int n_sets, int planners, int N, int nx, int nu, int npar, int lin_base, int lin_count,
                                       int n_obs, const double *xinit_sets, const double *x0, const double *obst_pred,
                                       const unsigned char *guided, double robot_radius, const double *stat, int n_static,
                                       double *params);
void oracle_guidance_halfspaces(int n_sets, int planners, int N, int nx, int nu, int npar, int lin_base, int lin_count,
                                int n_obs, const double *xinit_sets, const double *x0, const double *obst_pred,
                                const unsigned char *guided, double robot_radius, double *params)
{
    oracle_guidance_halfspaces_static(n_sets, planners, N, nx, nu, npar, lin_base, lin_count, n_obs, xinit_sets, x0, obst_pred, guided,
                                      robot_radius, NULL, 0, params);
}
/* + module_data.static_obstacles (linearized_constraints.cpp:107-127): n_static rows (a1, a2, b) per set and stage k >= 1,
 * behind the obstacle rows (guided) or from slot 0 (non-guided: empty obstacle list, guidance_constraints.cpp:326-329) */
void oracle_guidance_halfspaces_static(int n_sets, int planners, int N, int nx, int nu, int npar, int lin_base, int lin_count,
                                       int n_obs, const double *xinit_sets, const double *x0, const double *obst_pred,
                                       const unsigned char *guided, double robot_radius, const double *stat, int n_static,
                                       double *params)
{
    const int nz = nx + nu;
    const double r = 1e-3 + robot_radius;
    for (int q = 0; q < n_sets * planners; q++) {
        const int s = q / planners;
        const double dummy_b = xinit_sets[(size_t)s * nx] + 100.0;
        for (int k = 0; k < N; k++) {
            double *P = params + ((size_t)q * N + k) * npar + lin_base;
            for (int j = 0; j < lin_count; j++) { P[3 * j] = 1.0; P[3 * j + 1] = 0.0; P[3 * j + 2] = dummy_b; }
            if (stat && k > 0) {
                const int first = guided[q] ? n_obs : 0;
                const double *sh = stat + ((size_t)s * N + k) * n_static * 3;
                for (int h = 0; h < n_static && first + h < lin_count; h++) {
                    P[3 * (first + h)] = sh[3 * h]; P[3 * (first + h) + 1] = sh[3 * h + 1]; P[3 * (first + h) + 2] = sh[3 * h + 2];
                }
            }
            if (k == 0 || !guided[q]) continue;
            const double *ob = obst_pred + ((size_t)s * N + (k - 1)) * n_obs * 2;
            double pos[2] = {x0[((size_t)q * (N + 1) + k) * nz + nu], x0[((size_t)q * (N + 1) + k) * nz + nu + 1]};
            if (n_obs > 0)
                for (int it = 0; it < 3; it++)
                    for (int j = 0; j < n_obs; j++) {
                        const double dx = pos[0] - ob[2 * j], dy = pos[1] - ob[2 * j + 1];
                        if (sqrt(dx * dx + dy * dy) < r) {
                            double pa[2], ra[2], pb[2];
                            dr_proj(pos, ob, r, pos, pa);                       /* anchor = obstacle 0 (:143) */
                            ra[0] = 2.0 * pa[0] - pos[0]; ra[1] = 2.0 * pa[1] - pos[1];
                            dr_proj(ra, ob + 2 * j, r, pos, pb);
                            pos[0] = 0.5 * (pos[0] + 2.0 * pb[0] - ra[0]); pos[1] = 0.5 * (pos[1] + 2.0 * pb[1] - ra[1]);
                        }
                    }
            for (int j = 0; j < n_obs && j < lin_count; j++) {
                const double ox = ob[2 * j], oy = ob[2 * j + 1];
                const double dx = ox - pos[0], dy = oy - pos[1], dist = sqrt(dx * dx + dy * dy);
                const double a1 = dx / dist, a2 = dy / dist;
                P[3 * j] = a1; P[3 * j + 1] = a2; P[3 * j + 2] = a1 * ox + a2 * oy - r;
            }
        }
    }
}

/* calculateConsistencyCostForSolver (guidance_constraints.cpp:1025-1050): weight * sum_{k=1}^{N-2} (dx^2 + dy^2) between the
 * SOLVED trajectory (getOutput(k, "x"/"y")) and _interpolated_prev_trajectory[k]; not scaled by dt.  Plain C++ semantics
 * of the reference: no FMA contraction (oracle/Makefile compiles with -ffp-contract=off). */
double oracle_consistency_cost(const double *xtraj, const double *prev, int N, int nx, int ix, int iy, double weight)
{
    double sum = 0.0;
    for (int k = 1; k <= N - 2; k++) {
        const double dx = xtraj[k * nx + ix] - prev[2 * k], dy = xtraj[k * nx + iy] - prev[2 * k + 1];
        sum += dx * dx + dy * dy;
    }
    return weight * sum;
}
/* Objective post-processing + FindBestPlanner (guidance_constraints.cpp:373-420,572-590):
 * objective = pobj [- consistency cost if has_consistency_enabled (:384-388,405-408)] [- obj_sub] [* obj_scale (:418-419)] */
int oracle_select_best_cons(int n_sets, const int *set_offsets, const double *pobj, const int *exit_code,
                            const double *obj_scale, const double *obj_sub, const unsigned char *disabled, int *best_idx,
                            const double *xtraj, const double *prev_traj, const unsigned char *cons_enabled, double cons_weight,
                            int N, int nx, int ix, int iy, double *objective_out, double *cons_out)
{
    for (int s = 0; s < n_sets; s++) {
        double best = 1e10;
        int bi = -1;
        for (int i = set_offsets[s]; i < set_offsets[s + 1]; i++) {
            double obj = pobj[i], cons = 0.0;
            if (prev_traj && (!cons_enabled || cons_enabled[i])) {
                cons = oracle_consistency_cost(xtraj + (size_t)i * (N + 1) * nx, prev_traj + (size_t)s * N * 2, N, nx, ix, iy, cons_weight);
                obj -= cons;
            }
            if (obj_sub) obj -= obj_sub[i];
            if (obj_scale) obj *= obj_scale[i];
            if (objective_out) objective_out[i] = obj;
            if (cons_out) cons_out[i] = cons;
            if (disabled && disabled[i]) continue;
            if (exit_code[i] == 1 && obj < best) { best = obj; bi = i - set_offsets[s]; }
        }
        best_idx[s] = bi;
    }
    return 0;
}
int oracle_select_best(int n_sets, const int *set_offsets, const double *pobj, const int *exit_code,
                       const double *obj_scale, const double *obj_sub, const unsigned char *disabled, int *best_idx)
{
    return oracle_select_best_cons(n_sets, set_offsets, pobj, exit_code, obj_scale, obj_sub, disabled, best_idx, NULL, NULL, NULL, 0.0, 0, 0,
                                   0, 1, NULL, NULL);
}

/* ---- component entry points used by the unit tests ------------------------------------------ */
void oracle_model_eval(const double *z, const double *p, const double *mu, const double *mh, double *f, double *Jf,
                       double *Hf, double *cost, double *gc, double *Hc, double *h, double *Jh, double *Hh)
{
    model_f(z + NU, z, p, f); model_f_jac(z + NU, z, p, Jf); model_f_hess(z + NU, z, p, mu, Hf);
    *cost = model_cost(z, p); model_cost_grad_hess(z, p, gc, Hc);
    model_h(z, p, h); model_h_jac(z, p, Jh); model_h_hess(z, p, mh, Hh);
}
void oracle_integrate(const double *x, const double *u, const double *p, const double *pi, double *xn, double *W, double *Hc)
{
    integrate(x, u, p, xn, W, pi, Hc);
}
void oracle_mirror(double *A, int n) { mirror(A, n, n); }
const double *oracle_bounds(int which)
{
    return which == 0 ? model_lbz : which == 1 ? model_ubz : which == 2 ? model_lh : model_uh;
}
const char *oracle_param_name(int i) { return model_param_names[i]; }
const char *oracle_var_name(int i) { return model_var_names[i]; }

/* Linearise at (x0 warm start, multipliers zero) and solve ONE QP; export QP data and solution so
 * that tests can verify the KKT conditions independently (numpy).                               */
int oracle_qp_debug(const double *xinit, const double *x0, const double *params, double *Hout, double *gout,
                    double *Wout, double *bout, double *Cout, double *dout, double *vout, double *piout,
                    double *lamout, double *tout, int *iters)
{
    setup_constraints();
    work_t *w = (work_t *)calloc(1, sizeof(work_t));
    for (int k = 0; k <= NN; k++) {
        if (k < NN) for (int i = 0; i < NU; i++) w->u[k][i] = x0[k * NZ + i];
        for (int i = 0; i < NX; i++) w->x[k][i] = x0[k * NZ + NU + i];
    }
    linearize(w, xinit, params);
    double dx0[NX];
    for (int i = 0; i < NX; i++) dx0[i] = xinit[i] - w->x[0][i];
    int st = qp_solve(w, dx0);
    memcpy(Hout, w->H, sizeof(w->H)); memcpy(gout, w->g, sizeof(w->g));
    memcpy(Wout, w->W, sizeof(w->W)); memcpy(bout, w->b, sizeof(w->b));
    memcpy(Cout, w->C, sizeof(w->C));
    for (int k = 0; k < NN; k++)
        for (int e = 0; e < g_nc; e++) {
            dout[k * g_nc + e] = w->d[k][e]; lamout[k * g_nc + e] = w->lam[k][e]; tout[k * g_nc + e] = w->t[k][e];
        }
    memcpy(vout, w->v, sizeof(w->v)); memcpy(piout, w->qpi, sizeof(w->qpi));
    *iters = w->qp_iters_last;
    free(w);
    return st;
}
#ifdef __cplusplus
}
#endif
#endif /* ORACLE_NO_API */
