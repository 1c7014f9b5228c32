// mpc_solve_split.cuh -- role-split solve kernel: one CTA per problem, the work of a stage shared out over warps.
//
// Included inside namespace MPC_NS by mpc_solve_kernel.cuh (it reuses the MIRROR, interior-point entry algebra and
// cooperative Riccati routines defined there).  Same algorithm contract, same reference
// boundary (Solver::solve(), mpc_planner_solver/src/acados_solver_interface.cpp:86-204).
//
// Why: the thread-per-stage kernel keeps the WHOLE stage (iterate, Hessian, 14 box and NCG general inequality entries)
// in one thread -- 255 registers plus ~2.7 KB of thread-local memory, 8 problems per SM, each advancing slowly on one
// warp.  Here a problem owns 2 + NBR warps:
//   role A (warp 0)      lane k = stage k: iterate, cost / dynamics linearisation, MIRROR, residuals, the scalar
//                        decisions of the interior-point loop, and the cooperative Riccati recursion
//   role X (warp 1)      lane k = stage k: the 2 NZ box entries of that stage (multipliers, slacks in registers)
//   roles B (warps 2..)  lane k = stage k: a contiguous slice of the general inequality entries of that stage
//                        (constraint evaluation, multiplier-weighted constraint Hessian, the per-entry Newton algebra)
// Roles meet at named-barrier points; B hands A its per-stage contributions (Hessian / gradient / residual terms,
// step-length ratios, complementarity sums) through shared-memory slots that A adds in a fixed role order, so results
// do not depend on timing.  A and B run mirrored loop structures with the same barrier sequence; every loop decision is
// taken by A and published in shared memory.
#pragma once

#ifndef MPC_SPLIT_XROLES         // roles X: the box entries of a stage shared out over 1 or 2 warps.  One warp with all 2 NZ entries is the
#define MPC_SPLIT_XROLES 2       // straggler of every pass (28 entry updates against 4-5 per B role); two warps halve what role A waits for
#endif
#ifndef MPC_SPLIT_ROLES          // roles B: slices of 4-5 general entries; with role A and the X roles the CTA stays at 8 warps (255 registers each)
#define MPC_SPLIT_ROLES (NCG >= 24 ? 7 - MPC_SPLIT_XROLES : (NCG >= 12 ? 3 : 2))
#endif
#ifndef MPC_SPLIT_MIN_CTAS       // 1: all 255 registers for role A (the kernel serves small batches: one CTA per SM)
#define MPC_SPLIT_MIN_CTAS 1
#endif
#ifndef MPC_SPLIT_B_UNROLL       // entries of a B role's slice in flight together (2 / 3 / 5: kernel 2.19 / 2.20 / 2.23 ms for the 9-planner set)
#define MPC_SPLIT_B_UNROLL 2
#endif
constexpr int SPLIT_B_UNROLL = MPC_SPLIT_B_UNROLL;
constexpr int NBR = MPC_SPLIT_ROLES;
constexpr int NXR = MPC_SPLIT_XROLES;            // warps sharing the box entries
constexpr int XSPLIT = NXR == 2 ? (NZ + 1) / 2 : NZ;   // variables 0..XSPLIT-1: X role 0, the rest: X role 1
constexpr int SPLIT_WARPS = 1 + NXR + NBR;       // role A, NXR roles X (box entries), NBR roles B
constexpr int SPLIT_THREADS = SPLIT_WARPS * 32;
constexpr int RPB = (NCG + NBR - 1) / NBR > 0 ? (NCG + NBR - 1) / NBR : 1;     // general entries per B role
constexpr int NHP = NHS * (NHS + 1) / 2;                                       // packed block over the support of h
constexpr int XS = NHP + 2 * NHS + 3;            // exchange slots per B role and stage: Hs | g | rg | nd nm sm
constexpr int XSX = 3 * NZ + 3;                  // exchange slots of role X per stage: diag Ht | g | rg | nd nm sm
constexpr bool SPLIT_OK = (NSTAGE + 1 <= 32) && NPAD <= 32 && NCG >= 6 && NCG >= 2 * NBR && !BALANCE;      // (the BALANCE variant exists in the thread-per-stage kernel only)
// shared memory of one problem (doubles)
constexpr int SP_GAP = MPC_CHECK ? 2 : 0;        // MPC_CHECK: two canary doubles behind every region
constexpr int SP_RS = 0;
constexpr int SP_XCH = ((RS_DOUBLES + 1) & ~1) + SP_GAP;    // [NBR][XS][32]
constexpr int SP_XCX = SP_XCH + NBR * XS * 32 + SP_GAP;   // [NXR][XSX][32] slots of the X roles
constexpr int SP_PUB = SP_XCX + NXR * XSX * 32 + SP_GAP;        // [2 NZ][32]: z and v of every stage, published by role A (v: amended by role X)
constexpr int SP_DEC = SP_PUB + 2 * NZ * 32 + SP_GAP;     // decisions published by role A
constexpr int SP_SW = SP_DEC + 8 + SP_GAP;                // blocked-sweep workspace (SW_DOUBLES)
constexpr int SP_DOUBLES = SP_SW + SW_DOUBLES + SP_GAP;
#if MPC_CHECK
__device__ constexpr int SP_CANARY_AT[6] = {SP_XCH - SP_GAP, SP_XCX - SP_GAP, SP_PUB - SP_GAP, SP_DEC - SP_GAP, SP_SW - SP_GAP, SP_DOUBLES - SP_GAP};
#endif
enum { DEC_CONT = 0, DEC_SIGMU = 1, DEC_STEP = 2, DEC_SQP = 3, DEC_STATUS = 4 };

#ifdef MPC_PROF          // cycle accounting of role A (lane 0 of problem 0 prints it): a diagnostic build, never shipped
#define PROF_DECL long long pt_[24] = {0}; long long pc_ = clock64();
#define PROF(i) { const long long c_ = clock64(); pt_[i] += c_ - pc_; pc_ = c_; }
#else
#define PROF_DECL
#define PROF(i)
#endif
__host__ __device__ constexpr int hidx(int a, int b) { return a * (a + 1) / 2 + b; }      // a >= b
__device__ __forceinline__ void split_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(SPLIT_THREADS) : "memory"); }
// Rendezvous with a label: the roles run mirrored loop structures and must meet at the SAME point of the protocol.  The MPC_CHECK
// build verifies it (every role posts the label it arrived with; role A compares after the barrier) -- the race evidence the
// closed compute-sanitizer cannot give.
#if MPC_CHECK
__device__ __forceinline__ void split_barrier_l(int label)
{
    __shared__ int s_label[SPLIT_WARPS];
    const int role = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) s_label[role] = label;
    split_barrier();
    if (threadIdx.x == 0) {
        for (int r = 0; r < SPLIT_WARPS; r++)
            if (s_label[r] != label) MPC_CHECK_FAIL(2);
    }
    split_barrier();
}
#else
__device__ __forceinline__ void split_barrier_l(int) { split_barrier(); }
#endif

// ------------------------------------------------------------------------------------------------------------------
// role B: a slice [e_lo, e_hi) of the general inequality entries of every path stage
// ------------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void split_role_b(const int prob, const double* __restrict__ params_g, const int num_iter, double* mem_g,
                                          const int mem_doubles, double* sm, const int rb_)
{
    const int k = threadIdx.x & 31;
    const bool path = k < NSTAGE;
    const double* __restrict__ p = params_g + ((size_t)prob * NSTAGE + (path ? k : NSTAGE - 1)) * NP;
    double* const rs = sm + SP_RS;
    double* const xs = sm + SP_XCH + (size_t)rb_ * XS * 32 + k;        // slot s of this role and stage: xs[s * 32]
    double* const pub = sm + SP_PUB + k;
    const double* const dec = sm + SP_DEC;
    const double* const blk = rs + (path ? k : 0) * RSTRIDE;
    const int e_lo = rb_ * RPB < NCG ? rb_ * RPB : NCG, e_hi = (rb_ + 1) * RPB < NCG ? (rb_ + 1) * RPB : NCG;
    const int ne = path ? e_hi - e_lo : 0;
    const int r_lo = e_hi > e_lo ? HROW[e_lo] : 0, r_hi = e_hi > e_lo ? HROW[e_hi - 1] + 1 : 0;

    double C[RPB * NHS], Ce[RPB * NHS], sge[RPB], dg[RPB], lamg[RPB], tg[RPB], itg[RPB];
    double v3[NHS], vo3[NHS], dva3[NHS], dv3[NHS];
#pragma unroll
    for (int e = 0; e < RPB; e++) { lamg[e] = 0.0; tg[e] = 0.0; itg[e] = 0.0; dg[e] = 0.0; }
    for (int i = 0; i < RPB * NHS; i++) { C[i] = 0.0; Ce[i] = 0.0; }
#pragma unroll
    for (int e = 0; e < RPB; e++) sge[e] = 1.0;
#pragma unroll
    for (int a = 0; a < NHS; a++) { v3[a] = 0.0; vo3[a] = 0.0; dva3[a] = 0.0; dv3[a] = 0.0; }
    int qp_warm = 0;
    double* mem = mem_g ? mem_g + (size_t)prob * mem_doubles : nullptr;
    if (mem && mem[0] != 0.0) {
        const double* m = mem + 1 + (NSTAGE + 1) * NX;
#pragma unroll
        for (int e = 0; e < RPB; e++) if (e < ne) { lamg[e] = m[k * NC + NCB + e_lo + e]; tg[e] = m[NSTAGE * NC + k * NC + NCB + e_lo + e]; }
        qp_warm = (mem[0] >= 2.0);
    }

    for (int it = 0; it < num_iter; it++) {
        split_barrier_l(101);                                             // L1: z on the support of h is published
        {
            double zz[NZ], Hh[NPK];
#pragma unroll
            for (int i = 0; i < NZ; i++) zz[i] = 0.0;
#pragma unroll
            for (int a = 0; a < NHS; a++) zz[HSUP[a]] = pub[HSUP[a] * 32];
#pragma unroll
            for (int i = 0; i < NPK; i++) Hh[i] = 0.0;
            if (ne > 0) {
                double mh[RPB], hv[RPB];
                for (int j = 0; j < RPB; j++) mh[j] = 0.0;
#pragma unroll
                for (int e = 0; e < RPB; e++) if (e < ne) mh[HROW[e_lo + e] - r_lo] -= HSGN[e_lo + e] * lamg[e];      // lam_u - lam_l
                con_lin_rows(zz, p, r_lo, r_hi, mh, Hh, hv, C);
#pragma unroll
                for (int e = 0; e < RPB; e++) if (e < ne) dg[e] = HSGN[e_lo + e] * (HBND[e_lo + e] - hv[HROW[e_lo + e] - r_lo]);
                // per ENTRY copies (row of the entry, its sign): the passes below index them with compile-time constants, so they
                // live in registers -- indexed through HROW[e_lo + e] they sat in thread-local memory and every dot product
                // of a pass waited for local loads
#pragma unroll
                for (int e = 0; e < RPB; e++) {
                    const int r = e < ne ? HROW[e_lo + e] - r_lo : 0;
                    sge[e] = e < ne ? HSGN[e_lo + e] : 1.0;
#pragma unroll
                    for (int a = 0; a < NHS; a++) Ce[e * NHS + a] = C[r * NHS + a];
                }
            }
#pragma unroll
            for (int a = 0; a < NHS; a++)
#pragma unroll
                for (int b = 0; b <= a; b++) xs[hidx(a, b) * 32] = Hh[pk(HSUP[a], HSUP[b])];
        }
        split_barrier_l(102);                                             // L2: constraint Hessian terms handed to A
        split_barrier_l(103);                                             // L3: A has initialised v
        split_barrier_l(104);                                             // L4: role X has centred v inside the boxes
#pragma unroll
        for (int a = 0; a < NHS; a++) v3[a] = pub[(NZ + HSUP[a]) * 32];
        if (qp_warm) {
#pragma unroll
            for (int e = 0; e < RPB; e++) if (e < ne) { lamg[e] = clamp_lo(lamg[e], IPM_THR0); tg[e] = clamp_lo(tg[e], IPM_THR0); }
        } else {
#pragma unroll
            for (int e = 0; e < RPB; e++) {
                if (e < ne) {
                    double s = 0.0;
#pragma unroll
                    for (int a = 0; a < NHS; a++) s += Ce[e * NHS + a] * v3[a];
                    double tt = sge[e] * s - dg[e];
                    if (tt < IPM_THR0) tt = IPM_THR0;
                    tg[e] = tt; lamg[e] = IPM_MU0 * rcp_nb(tt);
                }
            }
        }

        double a_ = 0.0, sigmu = 0.0;
        for (int kk = 0;; kk++) {
            // ---- pass DA: apply the previous step to this slice, then its terms of Htilde, gtilde, the stationarity
            //      residual, and the slice's residual / complementarity measures
            const bool upd = kk > 0;
#pragma unroll
            for (int a = 0; a < NHS; a++) { vo3[a] = v3[a]; if (upd) v3[a] += a_ * dv3[a]; }
            double Hs[NHP], gs[NHS], rgs[NHS], nd = 0.0, nm = 0.0, sm_ = 0.0;
#pragma unroll
            for (int i = 0; i < NHP; i++) Hs[i] = 0.0;
#pragma unroll
            for (int a = 0; a < NHS; a++) { gs[a] = 0.0; rgs[a] = 0.0; }
#pragma unroll
            for (int e = 0; e < RPB; e++) if (e < ne) {
                const double sg = sge[e];
                double lam = lamg[e], t = tg[e];
                double cv = 0.0;
                if (upd) {
                    double cvo = 0.0, cda = 0.0, cd = 0.0;
#pragma unroll
                    for (int a = 0; a < NHS; a++) {
                        const double ca = Ce[e * NHS + a];
                        cvo += ca * vo3[a]; cda += ca * dva3[a]; cd += ca * dv3[a];
                    }
                    const IneqStep st = ineq_final(lam, itg[e], sg * cvo - dg[e] - t, sg * cda, sg * cd, sigmu);
                    lam = clamp_lo(lam + a_ * st.dlam, IPM_LAM_MIN); t = clamp_lo(t + a_ * st.dt, IPM_T_MIN);
                    lamg[e] = lam; tg[e] = t;
                }
#pragma unroll
                for (int a = 0; a < NHS; a++) cv += Ce[e * NHS + a] * v3[a];
                const double it_ = rcp_nb(t);
                const double rd = sg * cv - dg[e] - t, G = lam * it_, m = lam * t;
                itg[e] = it_;
#pragma unroll
                for (int a = 0; a < NHS; a++) {
                    const double ca = Ce[e * NHS + a];
#pragma unroll
                    for (int bb = 0; bb <= a; bb++) Hs[hidx(a, bb)] += G * ca * Ce[e * NHS + bb];
                    gs[a] += sg * ca * (G * rd);
                    rgs[a] -= sg * ca * lam;
                }
                nd = nanmax(nd, fabs(rd)); nm = nanmax(nm, fabs(m)); sm_ += m;
            }
#pragma unroll
            for (int i = 0; i < NHP; i++) xs[i * 32] = Hs[i];
#pragma unroll
            for (int a = 0; a < NHS; a++) { xs[(NHP + a) * 32] = gs[a]; xs[(NHP + NHS + a) * 32] = rgs[a]; }
            xs[(NHP + 2 * NHS) * 32] = nd; xs[(NHP + 2 * NHS + 1) * 32] = nm; xs[(NHP + 2 * NHS + 2) * 32] = sm_;
            split_barrier_l(1);                                         // 1: DA terms handed to A
            split_barrier_l(2);                                         // 2: A has decided
            if (dec[DEC_CONT] == 0.0) break;
            split_barrier_l(3);                                         // 3: factorisation + predictor sweep done
#pragma unroll
            for (int a = 0; a < NHS; a++) dva3[a] = path ? blk[RO_DZ + HSUP[a]] : 0.0;

            // ---- pass B: affine step length, mu_aff sums, corrector vectors of this slice
            StepFrac sfa;
            double S1 = 0.0, S2 = 0.0, V1[NHS], V2[NHS];
#pragma unroll
            for (int a = 0; a < NHS; a++) { V1[a] = 0.0; V2[a] = 0.0; }
#pragma unroll
            for (int e = 0; e < RPB; e++) if (e < ne) {
                const double sg = sge[e], lam = lamg[e], t = tg[e];
                double cv = 0.0, cd = 0.0;
#pragma unroll
                for (int a = 0; a < NHS; a++) { cv += Ce[e * NHS + a] * v3[a]; cd += Ce[e * NHS + a] * dva3[a]; }
                const double it_ = itg[e];
                const IneqStep st = ineq_affine(lam, it_, sg * cv - dg[e] - t, sg * cd);
                sfa.add(lam, st.dlam, t, st.dt);
                S1 += lam * st.dt + t * st.dlam; S2 += st.dt * st.dlam;
#pragma unroll
                for (int a = 0; a < NHS; a++) {
                    V1[a] += sg * Ce[e * NHS + a] * st.corr;
                    V2[a] += sg * Ce[e * NHS + a] * it_;
                }
            }
#pragma unroll
            for (int a = 0; a < NHS; a++) { xs[a * 32] = V1[a]; xs[(NHS + a) * 32] = V2[a]; }
            xs[(2 * NHS) * 32] = sfa.ratio(); xs[(2 * NHS + 1) * 32] = S1; xs[(2 * NHS + 2) * 32] = S2;
            split_barrier_l(4);                                         // 4: pass-B terms handed to A
            split_barrier_l(5);                                         // 5: sigma mu published
            sigmu = dec[DEC_SIGMU];
            split_barrier_l(6);                                         // 6: corrector sweeps done
#pragma unroll
            for (int a = 0; a < NHS; a++) dv3[a] = path ? blk[RO_DZ + HSUP[a]] : 0.0;

            // ---- pass C: step length of the corrected direction
            StepFrac sfc;
#pragma unroll
            for (int e = 0; e < RPB; e++) if (e < ne) {
                const double sg = sge[e], lam = lamg[e], t = tg[e];
                double cv = 0.0, cda = 0.0, cd = 0.0;
#pragma unroll
                for (int a = 0; a < NHS; a++) {
                    const double ca = Ce[e * NHS + a];
                    cv += ca * v3[a]; cda += ca * dva3[a]; cd += ca * dv3[a];
                }
                const IneqStep st = ineq_final(lam, itg[e], sg * cv - dg[e] - t, sg * cda, sg * cd, sigmu);
                sfc.add(lam, st.dlam, t, st.dt);
            }
            xs[0] = sfc.ratio();
            split_barrier_l(7);                                         // 7: ratios handed to A
            split_barrier_l(8);                                         // 8: step length published
            a_ = dec[DEC_STEP];
        }
        split_barrier_l(9);                                             // 9: outcome of this SQP iteration published
        qp_warm = 1;
        if (dec[DEC_SQP] == 0.0) break;
    }
    split_barrier_l(99);                                                 // F: final status published
    if (mem && dec[DEC_STATUS] == 0.0) {
        double* m = mem + 1 + (NSTAGE + 1) * NX;
#pragma unroll
        for (int e = 0; e < RPB; e++) if (e < ne) { m[k * NC + NCB + e_lo + e] = lamg[e]; m[NSTAGE * NC + k * NC + NCB + e_lo + e] = tg[e]; }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// role X: the 2 NZ box entries of every stage (input box on path stages, state box on stages 1..N-1)
// ------------------------------------------------------------------------------------------------------------------
template <int XR>
__device__ __noinline__ void split_role_x(const int prob, const int num_iter, double* mem_g, const int mem_doubles, double* sm)
{
    constexpr int I_LO = XR == 0 ? 0 : XSPLIT, I_HI = XR == 0 ? XSPLIT : NZ;      // this role's variables
    const int k = threadIdx.x & 31;
    const bool path = k < NSTAGE, xbox = path && k >= 1;
    double* const rs = sm + SP_RS;
    double* const xs = sm + SP_XCX + XR * XSX * 32 + k;                 // slot s of this role and stage: xs[s * 32]
    double* const pub = sm + SP_PUB + k;
    const double* const dec = sm + SP_DEC;
    const double* const blk = rs + (path ? k : 0) * RSTRIDE;

    double lamb[NCB], tb[NCB], itb[NCB], zl[NZ], zu[NZ], v[NZ], vo[NZ], dva[NZ], dv[NZ];
#pragma unroll
    for (int e = 0; e < NCB; e++) { lamb[e] = 0.0; tb[e] = 0.0; itb[e] = 0.0; }
#pragma unroll
    for (int i = 0; i < NZ; i++) { v[i] = 0.0; vo[i] = 0.0; dva[i] = 0.0; dv[i] = 0.0; zl[i] = 0.0; zu[i] = 0.0; }
    int qp_warm = 0;
    double* mem = mem_g ? mem_g + (size_t)prob * mem_doubles : nullptr;
    if (mem && mem[0] != 0.0) {
        const double* m = mem + 1 + (NSTAGE + 1) * NX;
        if (path) for (int e = 0; e < NCB; e++) {
            const int i = e < NZ ? e : e - NZ;
            if (i >= I_LO && i < I_HI) { lamb[e] = m[k * NC + e]; tb[e] = m[NSTAGE * NC + k * NC + e]; }
        }
        qp_warm = (mem[0] >= 2.0);
    }

    for (int it = 0; it < num_iter; it++) {
        split_barrier_l(101);                                             // L1: z is published
#pragma unroll
        for (int i = 0; i < NZ; i++) { const double zi = pub[i * 32]; zl[i] = LBZ[i] - zi; zu[i] = UBZ[i] - zi; }
        split_barrier_l(102);                                             // L2
        split_barrier_l(103);                                             // L3: A has initialised v
#pragma unroll
        for (int i = 0; i < NZ; i++) v[i] = pub[(NZ + i) * 32];
        if (qp_warm) {
#pragma unroll
            for (int i = I_LO; i < I_HI; i++) {
                const bool act = (i < NU) ? path : xbox;
                if (act) {
                    lamb[i] = clamp_lo(lamb[i], IPM_THR0); tb[i] = clamp_lo(tb[i], IPM_THR0);
                    lamb[NZ + i] = clamp_lo(lamb[NZ + i], IPM_THR0); tb[NZ + i] = clamp_lo(tb[NZ + i], IPM_THR0);
                }
            }
        } else {
#pragma unroll
            for (int i = I_LO; i < I_HI; i++) {
                const bool act = (i < NU) ? path : xbox;
                if (act) {
                    const double dl = zl[i], du = zu[i];
                    double tl = v[i] - dl, tu = du - v[i];
                    if (tl < IPM_THR0) {
                        if (tu < IPM_THR0) { v[i] = 0.5 * (dl + du); tl = IPM_THR0; tu = IPM_THR0; }
                        else { tl = IPM_THR0; v[i] = dl + IPM_THR0; }
                    } else if (tu < IPM_THR0) { tu = IPM_THR0; v[i] = du - IPM_THR0; }
                    tb[i] = tl; tb[NZ + i] = tu;
                    lamb[i] = IPM_MU0 * rcp_nb(tl); lamb[NZ + i] = IPM_MU0 * rcp_nb(tu);
                }
            }
#pragma unroll
            for (int i = I_LO; i < I_HI; i++) pub[(NZ + i) * 32] = v[i];
        }
        split_barrier_l(104);                                             // L4: v is final for every role

        double a_ = 0.0, sigmu = 0.0;
        for (int kk = 0;; kk++) {
            const bool upd = kk > 0;
#pragma unroll
            for (int i = 0; i < NZ; i++) { vo[i] = v[i]; if (upd) v[i] += a_ * dv[i]; }
            double nd = 0.0, nm = 0.0, sm_ = 0.0;
#pragma unroll
            for (int i = I_LO; i < I_HI; i++) {
                const bool act = (i < NU) ? path : xbox;
                double hd = 0.0, gg = 0.0, rr = 0.0;
                if (act) {
                    const double dl = zl[i], du = zu[i];
                    {   // lower: chat = +e_i, d = dl
                        double lam = lamb[i], t = tb[i];
                        if (upd) {
                            const IneqStep st = ineq_final(lam, itb[i], vo[i] - dl - t, dva[i], dv[i], sigmu);
                            lam = clamp_lo(lam + a_ * st.dlam, IPM_LAM_MIN); t = clamp_lo(t + a_ * st.dt, IPM_T_MIN);
                            lamb[i] = lam; tb[i] = t;
                        }
                        const double it_ = rcp_nb(t);
                        const double rd = v[i] - dl - t, G = lam * it_, m = lam * t;
                        itb[i] = it_;
                        hd += G; gg += G * rd; rr -= lam;
                        nd = nanmax(nd, fabs(rd)); nm = nanmax(nm, fabs(m)); sm_ += m;
                    }
                    {   // upper: chat = -e_i, d = -du
                        double lam = lamb[NZ + i], t = tb[NZ + i];
                        if (upd) {
                            const IneqStep st = ineq_final(lam, itb[NZ + i], du - vo[i] - t, -dva[i], -dv[i], sigmu);
                            lam = clamp_lo(lam + a_ * st.dlam, IPM_LAM_MIN); t = clamp_lo(t + a_ * st.dt, IPM_T_MIN);
                            lamb[NZ + i] = lam; tb[NZ + i] = t;
                        }
                        const double it_ = rcp_nb(t);
                        const double rd = du - v[i] - t, G = lam * it_, m = lam * t;
                        itb[NZ + i] = it_;
                        hd += G; gg -= G * rd; rr += lam;
                        nd = nanmax(nd, fabs(rd)); nm = nanmax(nm, fabs(m)); sm_ += m;
                    }
                }
                xs[i * 32] = hd; xs[(NZ + i) * 32] = gg; xs[(2 * NZ + i) * 32] = rr;
            }
            xs[(3 * NZ) * 32] = nd; xs[(3 * NZ + 1) * 32] = nm; xs[(3 * NZ + 2) * 32] = sm_;
            split_barrier_l(1);                                         // 1
            split_barrier_l(2);                                         // 2
            if (dec[DEC_CONT] == 0.0) break;
            split_barrier_l(3);                                         // 3
#pragma unroll
            for (int i = 0; i < NZ; i++) dva[i] = path ? blk[RO_DZ + i] : 0.0;

            StepFrac sfa;
            double S1 = 0.0, S2 = 0.0;
#pragma unroll
            for (int i = I_LO; i < I_HI; i++) {
                const bool act = (i < NU) ? path : xbox;
                double V1 = 0.0, V2 = 0.0;
                if (act) {
                    const double dl = zl[i], du = zu[i];
                    {
                        const double lam = lamb[i], t = tb[i];
                        const double it_ = itb[i];
                        const IneqStep st = ineq_affine(lam, it_, v[i] - dl - t, dva[i]);
                        sfa.add(lam, st.dlam, t, st.dt);
                        S1 += lam * st.dt + t * st.dlam; S2 += st.dt * st.dlam;
                        V1 += st.corr; V2 += it_;
                    }
                    {
                        const double lam = lamb[NZ + i], t = tb[NZ + i];
                        const double it_ = itb[NZ + i];
                        const IneqStep st = ineq_affine(lam, it_, du - v[i] - t, -dva[i]);
                        sfa.add(lam, st.dlam, t, st.dt);
                        S1 += lam * st.dt + t * st.dlam; S2 += st.dt * st.dlam;
                        V1 -= st.corr; V2 -= it_;
                    }
                }
                xs[i * 32] = V1; xs[(NZ + i) * 32] = V2;
            }
            xs[(2 * NZ) * 32] = sfa.ratio(); xs[(2 * NZ + 1) * 32] = S1; xs[(2 * NZ + 2) * 32] = S2;
            split_barrier_l(4);                                         // 4
            split_barrier_l(5);                                         // 5
            sigmu = dec[DEC_SIGMU];
            split_barrier_l(6);                                         // 6
#pragma unroll
            for (int i = 0; i < NZ; i++) dv[i] = path ? blk[RO_DZ + i] : 0.0;
            StepFrac sfc;
#pragma unroll
            for (int i = I_LO; i < I_HI; i++) {
                const bool act = (i < NU) ? path : xbox;
                if (act) {
                    const double dl = zl[i], du = zu[i];
                    {
                        const double lam = lamb[i], t = tb[i];
                        const IneqStep st = ineq_final(lam, itb[i], v[i] - dl - t, dva[i], dv[i], sigmu);
                        sfc.add(lam, st.dlam, t, st.dt);
                    }
                    {
                        const double lam = lamb[NZ + i], t = tb[NZ + i];
                        const IneqStep st = ineq_final(lam, itb[NZ + i], du - v[i] - t, -dva[i], -dv[i], sigmu);
                        sfc.add(lam, st.dlam, t, st.dt);
                    }
                }
            }
            xs[0] = sfc.ratio();
            split_barrier_l(7);                                         // 7
            split_barrier_l(8);                                         // 8
            a_ = dec[DEC_STEP];
        }
        split_barrier_l(9);                                             // 9
        qp_warm = 1;
        if (dec[DEC_SQP] == 0.0) break;
    }
    split_barrier_l(99);                                                 // F
    if (mem && dec[DEC_STATUS] == 0.0 && path) {
        double* m = mem + 1 + (NSTAGE + 1) * NX;
        for (int e = 0; e < NCB; e++) {
            const int i = e < NZ ? e : e - NZ;
            if (i >= I_LO && i < I_HI) { m[k * NC + e] = lamb[e]; m[NSTAGE * NC + k * NC + e] = tb[e]; }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------------
// role A: everything else of Solver::solve()
// ------------------------------------------------------------------------------------------------------------------
__device__ __noinline__ void split_role_a(const int prob, const double* __restrict__ xinit_g, const double* __restrict__ x0_g,
                                          const double* __restrict__ params_g, const int num_iter, double* mem_g, const int mem_doubles,
                                          double* xtraj_g, double* utraj_g, double* pobj_g, int* exit_g, int* qps_g, double* reseq_g,
                                          int* ipm_g, double* sm, const bool defer)
{
    const int k = threadIdx.x & 31;
    const bool path = k < NSTAGE, term = k == NSTAGE, live = k <= NSTAGE, xbox = path && k >= 1;
    const double* __restrict__ p = params_g + ((size_t)prob * NSTAGE + (path ? k : NSTAGE - 1)) * NP;
    double* const rs = sm + SP_RS;
    const double* const xch = sm + SP_XCH + k;                          // slot s of role r: xch[(r * XS + s) * 32]
    const double* const xcx = sm + SP_XCX + k;                          // slot s of role X: xcx[s * 32]
    double* const pub = sm + SP_PUB + k;
    double* const dec = sm + SP_DEC;
    double* const blk = rs + (live ? k : 0) * RSTRIDE;
    double* const sw = sm + SP_SW;                                      // blocked-sweep workspace
    double* const swk = sw + k * SWS;                                   // (k < 32 <= NPAD + 1 blocks)
    Grp grp;                                                            // warp-level helpers (a role is one warp here)
    grp.xch = nullptr; grp.gid = 0; grp.wig = 0;
    static_assert(GW == 1 || !SPLIT_OK, "role-split kernel: one warp per role");

    double z[NZ], pi[NX], v[NZ], qpi[NX], xi[NX];
#pragma unroll
    for (int i = 0; i < NZ; i++) {
        z[i] = live ? x0_g[(size_t)prob * NZ * (NSTAGE + 1) + k * NZ + i] : 0.0;      // loadWarmstart (:274-284)
        v[i] = 0.0;
    }
    if (term) { z[0] = 0.0; z[1] = 0.0; }
#pragma unroll
    for (int i = 0; i < NX; i++) { xi[i] = xinit_g[(size_t)prob * NX + i]; pi[i] = 0.0; qpi[i] = 0.0; }
    int qp_warm = 0;
    double* mem = mem_g ? mem_g + (size_t)prob * mem_doubles : nullptr;
    if (mem && mem[0] != 0.0) {                                         // persistent capsule memory: [flag][pi][lam][t][v]
        const double* m = mem + 1;
        if (live) for (int i = 0; i < NX; i++) pi[i] = m[k * NX + i];
        m += (NSTAGE + 1) * NX;
        m += 2 * NSTAGE * NC;                                            // (box multipliers: role X, general: roles B)
        if (live) for (int i = 0; i < NZ; i++) v[i] = m[k * NZ + i];
        qp_warm = (mem[0] >= 2.0);
#pragma unroll
        for (int i = 0; i < NX; i++) qpi[i] = pi[i];
    }

    int status = 0, qps = 0, ipm_total = 0;
    PROF_DECL
    for (int it = 0; it < num_iter; it++) {
        // ======================= linearise at the current iterate ====================================
        double H[NPK], g[NZ], Wv[NWV], b[NX];
        {
            double pin[NX], xnx[NX], zx_[NX];
#pragma unroll
            for (int i = 0; i < NX; i++) zx_[i] = z[NU + i];
            grp.shift_down(pi, pin);
            grp.shift_down(zx_, xnx);
#pragma unroll
            for (int i = 0; i < NZ; i++) pub[i * 32] = z[i];
            split_barrier_l(101);                                         // L1
#pragma unroll
            for (int i = 0; i < NPK; i++) H[i] = 0.0;
#pragma unroll
            for (int i = 0; i < NZ; i++) g[i] = 0.0;
#pragma unroll
            for (int i = 0; i < NX; i++) b[i] = 0.0;
#pragma unroll
            for (int i = 0; i < NWV; i++) Wv[i] = 0.0;
            if (path) {
                double xn[NX];
                cost_lin(z, p, g, H);
                dyn_lin(z, pin, xn, Wv, H);
#pragma unroll
                for (int i = 0; i < NX; i++) b[i] = xn[i] - xnx[i];
            }
            PROF(0)
            split_barrier_l(102);                                         // L2: constraint Hessian terms are in the slots
            PROF(1)
            if (path) {
#pragma unroll 1
                for (int r = 0; r < NBR; r++)
#pragma unroll
                    for (int a = 0; a < NHS; a++)
#pragma unroll
                        for (int c = 0; c <= a; c++) H[pk(HSUP[a], HSUP[c])] += xch[(r * XS + hidx(a, c)) * 32];
                mirror_packed(H);
                double Wd[NX * NZ];
                w_to_dense(Wv, Wd);
#pragma unroll
                for (int l = 0; l < NX; l++)
#pragma unroll
                    for (int j = 0; j < NZ; j++) blk[RO_B + l * NB + j] = Wd[l * NZ + j];
            } else if (term) {
#pragma unroll
                for (int i = NU; i < NZ; i++) H[pk(i, i)] = REG_EPS;   // mirror(0) = eps I; no terminal cost
            }
        }

        // ======================= interior-point QP: initialisation ===================================
        if (!qp_warm) {
#pragma unroll
            for (int i = 0; i < NZ; i++) v[i] = 0.0;
#pragma unroll
            for (int i = 0; i < NX; i++) qpi[i] = 0.0;
        }
        if (k == 0) {
#pragma unroll
            for (int i = 0; i < NX; i++) v[NU + i] = xi[i] - z[NU + i];
        }
#pragma unroll
        for (int i = 0; i < NZ; i++) pub[(NZ + i) * 32] = v[i];
        PROF(2)
        split_barrier_l(103);                                             // L3
        split_barrier_l(104);                                             // L4: role X has centred v inside the boxes (cold start)
#pragma unroll
        for (int i = 0; i < NZ; i++) v[i] = pub[(NZ + i) * 32];
        PROF(3)

        double alpha = 1.0, mu = 0.0;
        int kk = 0;
        bool isnan_ = false;
        double dva[NZ], dv[NZ], dpi[NX], sigmu = 0.0, a_ = 0.0;
#pragma unroll
        for (int i = 0; i < NZ; i++) { dva[i] = 0.0; dv[i] = 0.0; }
#pragma unroll
        for (int i = 0; i < NX; i++) dpi[i] = 0.0;
        for (;; kk++) {
            // ---- pass DA (role A part): apply the previous step, residuals, box entries
            double Ht[NPK], gt[NZ], rb[NX], rg[NZ];
            double ng = 0.0, nb = 0.0, nd = 0.0, nm = 0.0, sm_ = 0.0;
            const bool upd = kk > 0;
            double vo[NZ];
#pragma unroll
            for (int i = 0; i < NZ; i++) { vo[i] = v[i]; if (upd) v[i] += a_ * dv[i]; }
            if (upd) {
#pragma unroll
                for (int i = 0; i < NX; i++) qpi[i] += a_ * dpi[i];
            }
            {
                double qpn[NX], vxn[NX], vx_[NX];
#pragma unroll
                for (int i = 0; i < NX; i++) vx_[i] = v[NU + i];
                grp.shift_down(qpi, qpn);
                grp.shift_down(vx_, vxn);
#pragma unroll
                for (int i = 0; i < NPK; i++) Ht[i] = H[i];
#pragma unroll
                for (int i = 0; i < NZ; i++) {
                    double s = g[i];
#pragma unroll
                    for (int j = 0; j < NZ; j++) s += H[pk(i, j)] * v[j];
                    rg[i] = s;
                }
#pragma unroll
                for (int i = 0; i < NX; i++) rb[i] = 0.0;
                if (path) {
                    wt_mul_add(Wv, qpn, rg);
#pragma unroll
                    for (int i = 0; i < NX; i++) rb[i] = b[i] - vxn[i];
                    w_mul_add(Wv, v, rb);
                }
                if (k >= 1) {
#pragma unroll
                    for (int i = 0; i < NX; i++) rg[NU + i] -= qpi[i];
                }
#pragma unroll
                for (int i = 0; i < NZ; i++) gt[i] = rg[i];
            }
            PROF(4)
            split_barrier_l(1);                                         // 1: the slices' DA terms are in the slots
            PROF(5)
#pragma unroll
            for (int i = 0; i < NZ; i++) {                           // box entries (roles X: variable i lives in role i >= XSPLIT)
                const double* xr = xcx + (i >= XSPLIT ? XSX * 32 : 0);
                Ht[pk(i, i)] += xr[i * 32]; gt[i] += xr[(NZ + i) * 32]; rg[i] += xr[(2 * NZ + i) * 32];
            }
#pragma unroll
            for (int r = 0; r < NXR; r++) {
                const double* xr = xcx + r * XSX * 32;
                nd = nanmax(nd, xr[(3 * NZ) * 32]); nm = nanmax(nm, xr[(3 * NZ + 1) * 32]); sm_ += xr[(3 * NZ + 2) * 32];
            }
            if (path) {
#pragma unroll 1
                for (int r = 0; r < NBR; r++) {
                    const double* x = xch + (size_t)r * XS * 32;
#pragma unroll
                    for (int a = 0; a < NHS; a++) {
#pragma unroll
                        for (int c = 0; c <= a; c++) Ht[pk(HSUP[a], HSUP[c])] += x[hidx(a, c) * 32];
                        gt[HSUP[a]] += x[(NHP + a) * 32];
                        rg[HSUP[a]] += x[(NHP + NHS + a) * 32];
                    }
                    nd = nanmax(nd, x[(NHP + 2 * NHS) * 32]); nm = nanmax(nm, x[(NHP + 2 * NHS + 1) * 32]);
                    sm_ += x[(NHP + 2 * NHS + 2) * 32];
                }
            }
            if (k == 0) {
#pragma unroll
                for (int i = NU; i < NZ; i++) rg[i] = 0.0;          // x_0 is not a variable
            }
            if (live) {
#pragma unroll
                for (int i = 0; i < NZ; i++) ng = nanmax(ng, fabs(rg[i]));
#pragma unroll
                for (int i = 0; i < NX; i++) nb = nanmax(nb, fabs(rb[i]));
            }
            const bool lane_nan = (ng != ng) || (nb != nb) || (nd != nd) || (nm != nm);
            const bool lane_ok = (ng <= IPM_TOL) && (nb <= IPM_TOL) && (nd <= IPM_TOL) && (nm <= IPM_TOL);
            const bool any_nan = __any_sync(FULL, lane_nan);
            const bool all_ok = __all_sync(FULL, lane_ok);
            mu = warp_sum(sm_) / (double)IPM_COUNT;
            isnan_ = (mu != mu) || any_nan;
            const bool cont = (kk < IPM_ITER_MAX && alpha > IPM_ALPHA_MIN && !isnan_ && !all_ok);
            if (k == 0) dec[DEC_CONT] = cont ? 1.0 : 0.0;
            if (cont) {                                              // park Ht, gt, rb for the cooperative recursion
                if (live) {
#pragma unroll
                    for (int i = 0; i < NPK; i++) blk[RO_G + i] = Ht[i];
#pragma unroll
                    for (int i = 0; i < NZ; i++) blk[RO_Q + i] = gt[i];
                }
                if (path) {
#pragma unroll
                    for (int i = 0; i < NX; i++) blk[RO_B + i * NB + NZ] = rb[i];
                }
            }
            PROF(6)
            split_barrier_l(2);                                         // 2
            PROF(7)
            if (!cont) break;
            __syncwarp();
            riccati_factor_coop(rs);
            PROF(8)
            // closed-loop matrices of every stage (lane-parallel), then the predictor sweep
            double Lx0[NX], Lx1[NX], Prb[NX], lv[NU], Bd[NX * NU], L10 = 0.0, iL0 = 0.0, iL1 = 0.0;
#pragma unroll
            for (int i = 0; i < NX; i++) { Lx0[i] = 0.0; Lx1[i] = 0.0; Prb[i] = 0.0; }
#pragma unroll
            for (int i = 0; i < NX * NU; i++) Bd[i] = 0.0;
            lv[0] = lv[1] = 0.0;
            if (path) {
                iL0 = blk[RO_G + pk(0, 0)]; L10 = blk[RO_G + pk(1, 0)]; iL1 = blk[RO_G + pk(1, 1)];
                lv[0] = blk[RO_Q]; lv[1] = blk[RO_Q + 1];
#pragma unroll
                for (int i = 0; i < NX; i++) {
                    Lx0[i] = blk[RO_G + pk(NU + i, 0)]; Lx1[i] = blk[RO_G + pk(NU + i, 1)]; Prb[i] = blk[RO_PRB + i];
                }
                double Acl[NX * NX];
                closed_loop(Wv, Lx0, Lx1, L10, iL0, iL1, Acl, Bd);
                const double k1 = -lv[1] * iL1, k0 = -(lv[0] + L10 * k1) * iL0;      // kff = -Luu^-T l
#pragma unroll
                for (int i = 0; i < NX * NX; i++) swk[SW_A + i] = Acl[i];
#pragma unroll
                for (int i = 0; i < NX; i++) swk[SW_B + i] = rb[i] + Bd[i * NU] * k0 + Bd[i * NU + 1] * k1;
            } else if (k < NPAD) {                                   // padding stages: identity
#pragma unroll
                for (int i = 0; i < NX * NX; i++) swk[SW_A + i] = (i / NX == i % NX) ? 1.0 : 0.0;
#pragma unroll
                for (int i = 0; i < NX; i++) swk[SW_B + i] = 0.0;
            }
            __syncwarp();
            PROF(22)
            sweep_products(sw);
            sweep_forward_blocked(sw);
            PROF(23)
            if (live) {
#pragma unroll
                for (int i = 0; i < NX; i++) dva[NU + i] = swk[SW_X + i];
            }
            {
                double r0 = lv[0], r1 = lv[1];           // du = -Luu^-T (Lxu' dx + l)
#pragma unroll
                for (int j = 0; j < NX; j++) { r0 += Lx0[j] * dva[NU + j]; r1 += Lx1[j] * dva[NU + j]; }
                dva[1] = path ? -r1 * iL1 : 0.0;
                dva[0] = path ? -(r0 + L10 * dva[1]) * iL0 : 0.0;
            }
            if (path) {                                              // roles X and B read the step from the blocks
#pragma unroll
                for (int i = 0; i < NZ; i++) blk[RO_DZ + i] = dva[i];
            }
            PROF(9)
            split_barrier_l(3);                                         // 3
            PROF(10)

            // ---- pass B (box entries): affine step length, mu_aff sums, corrector vectors
            double S1 = 0.0, S2 = 0.0, V1[NZ], V2[NZ];
            double ratio = 1.0;
            PROF(11)
            split_barrier_l(4);                                         // 4: the slices' pass-B terms are in the slots
            PROF(12)
#pragma unroll
            for (int i = 0; i < NZ; i++) {                           // box entries (roles X)
                const double* xr = xcx + (i >= XSPLIT ? XSX * 32 : 0);
                V1[i] = xr[i * 32]; V2[i] = xr[(NZ + i) * 32];
            }
#pragma unroll
            for (int r = 0; r < NXR; r++) {
                const double* xr = xcx + r * XSX * 32;
                ratio = fmin(ratio, xr[(2 * NZ) * 32]); S1 += xr[(2 * NZ + 1) * 32]; S2 += xr[(2 * NZ + 2) * 32];
            }
            if (path) {
#pragma unroll 1
                for (int r = 0; r < NBR; r++) {
                    const double* x = xch + (size_t)r * XS * 32;
#pragma unroll
                    for (int a = 0; a < NHS; a++) { V1[HSUP[a]] += x[a * 32]; V2[HSUP[a]] += x[(NHS + a) * 32]; }
                    ratio = fmin(ratio, x[(2 * NHS) * 32]);
                    S1 += x[(2 * NHS + 1) * 32]; S2 += x[(2 * NHS + 2) * 32];
                }
            }
            const double alpha_aff = warp_min(ratio);
            S1 = warp_sum(S1); S2 = warp_sum(S2);
            const double mu_aff = (mu * (double)IPM_COUNT + alpha_aff * S1 + alpha_aff * alpha_aff * S2) / (double)IPM_COUNT;
            const double rat = mu_aff / mu;
            sigmu = rat * rat * rat * mu;
            if (k == 0) dec[DEC_SIGMU] = sigmu;
#pragma unroll
            for (int i = 0; i < NZ; i++) gt[i] += V1[i] - sigmu * V2[i];
            PROF(13)
            split_barrier_l(5);                                         // 5
            PROF(14)
            double pv[NX];
            {   // backward vector sweep: p_k = Acl_k' (p_{k+1} + P_{k+1} rb_k) + (gt_x - Lxu Luu^-1 gt_u)
                if (path) {
                    const double lg0 = gt[0] * iL0, lg1 = (gt[1] - L10 * lg0) * iL1;
#pragma unroll
                    for (int i = 0; i < NX; i++) {
                        double c_ = gt[NU + i] - Lx0[i] * lg0 - Lx1[i] * lg1;
#pragma unroll
                        for (int j = 0; j < NX; j++) c_ += swk[SW_A + j * NX + i] * Prb[j];
                        swk[SW_B + i] = c_;
                    }
                } else if (k < NPAD) {
#pragma unroll
                    for (int i = 0; i < NX; i++) swk[SW_B + i] = 0.0;
                }
                if (term) {
#pragma unroll
                    for (int i = 0; i < NX; i++) sw[NPAD * SWS + SW_X + i] = gt[NU + i];      // p_N (= p_NPAD: identity padding)
                }
                __syncwarp();
                sweep_backward_blocked(sw);
                double pn[NX];
#pragma unroll
                for (int i = 0; i < NX; i++) { pv[i] = live ? swk[SW_X + i] : 0.0; pn[i] = path ? swk[SWS + SW_X + i] : 0.0; }
                if (path) {
                    double q0 = gt[0], q1 = gt[1];
#pragma unroll
                    for (int i = 0; i < NX; i++) {
                        const double y = pn[i] + Prb[i];
                        q0 += Bd[i * NU] * y; q1 += Bd[i * NU + 1] * y;
                    }
                    lv[0] = q0 * iL0;
                    lv[1] = (q1 - L10 * lv[0]) * iL1;
                    const double k1 = -lv[1] * iL1, k0 = -(lv[0] + L10 * k1) * iL0;
#pragma unroll
                    for (int i = 0; i < NX; i++) swk[SW_B + i] = rb[i] + Bd[i * NU] * k0 + Bd[i * NU + 1] * k1;
                }
                __syncwarp();
            }
            PROF(15)
            sweep_forward_blocked(sw);
            if (live) {
#pragma unroll
                for (int i = 0; i < NX; i++) dv[NU + i] = swk[SW_X + i];
            }
            {
                double r0 = lv[0], r1 = lv[1];
#pragma unroll
                for (int j = 0; j < NX; j++) { r0 += Lx0[j] * dv[NU + j]; r1 += Lx1[j] * dv[NU + j]; }
                dv[1] = path ? -r1 * iL1 : 0.0;
                dv[0] = path ? -(r0 + L10 * dv[1]) * iL0 : 0.0;
            }
            if (path) {
#pragma unroll
                for (int i = 0; i < NZ; i++) blk[RO_DZ + i] = dv[i];
            }
            PROF(16)
            split_barrier_l(6);                                         // 6
            PROF(17)
            if (k >= 1 && live) {                                    // dpi_k = P_k dx_k + p_k
#pragma unroll
                for (int i = 0; i < NX; i++) {
                    double a = pv[i];
#pragma unroll
                    for (int j = 0; j < NX; j++) a += blk[RO_G + pk(NU + i, NU + j)] * dv[NU + j];
                    dpi[i] = a;
                }
            }

            // ---- pass C (box entries): step length of the corrected direction
            double ratc = 1.0;
            PROF(18)
            split_barrier_l(7);                                         // 7: the slices' ratios are in slot 0
            PROF(19)
#pragma unroll
            for (int r = 0; r < NXR; r++) ratc = fmin(ratc, xcx[r * XSX * 32]);      // box entries (roles X): slot 0 of each role's region
            if (path) {
#pragma unroll 1
                for (int r = 0; r < NBR; r++) ratc = fmin(ratc, xch[(size_t)r * XS * 32]);
            }
            alpha = warp_min(ratc);
            a_ = alpha < 1.0 ? alpha * IPM_STEP_SCALE : alpha;      // applied by the next pass DA
            if (k == 0) dec[DEC_STEP] = a_;
            PROF(20)
            split_barrier_l(8);                                         // 8
            PROF(21)
        }
        ipm_total += kk;
        qps = isnan_ ? 3 : ((kk == IPM_ITER_MAX) ? 1 : (alpha <= IPM_ALPHA_MIN ? 2 : 0));

        // ======================= SQP-RTI full step ===================================================
        const bool qp_failed = (qps != 0 && qps != 1);
        if (k == 0) dec[DEC_SQP] = (qps == 0) ? 1.0 : 0.0;
        split_barrier_l(9);                                             // 9
        if (qp_failed) { status = 4; break; }                        // ACADOS_QP_FAILURE: iterate unchanged
#pragma unroll
        for (int i = 0; i < NZ; i++) z[i] += v[i];
        if (term) { z[0] = 0.0; z[1] = 0.0; }
#pragma unroll
        for (int i = 0; i < NX; i++) pi[i] = qpi[i];
        qp_warm = 1;
        status = 0;
        if (qps != 0) break;                                         // wrapper breaks on qp_status != 0 (:105-106)
    }

    // ======================= completeOneIteration (:162-204) =========================================
    double cst = 0.0, req = 0.0;
    {
        double xnx[NX], zx_[NX];
#pragma unroll
        for (int i = 0; i < NX; i++) zx_[i] = z[NU + i];
        grp.shift_down(zx_, xnx);
        if (path) {
            double xn[NX];
            cst = DT * cost_val(z, p);
            dyn_phi(z, xn);
#pragma unroll
            for (int i = 0; i < NX; i++) req = nanmax(req, fabs(xn[i] - xnx[i]));
        }
    }
    double cost = 0.0;                                               // stage-order sum (matches the oracle's accumulation)
#pragma unroll 1
    for (int l = 0; l < 32; l++) cost += __shfl_sync(FULL, cst, l);
    req = warp_max(req);
    if (!(req <= RES_EQ_MAX) && status == 0 && !defer) status = 4;      // defer: completion belongs to the caller (stepwise interface)
    if (k == 0) dec[DEC_STATUS] = (double)status;
    split_barrier_l(99);                                                 // F
    const int exit_code = (status == 0) ? 1 : (status == 1 ? 0 : status);
    if (live) {
#pragma unroll
        for (int i = 0; i < NX; i++) xtraj_g[(size_t)prob * NX * (NSTAGE + 1) + k * NX + i] = z[NU + i];
    }
    if (path) {
#pragma unroll
        for (int i = 0; i < NU; i++) utraj_g[(size_t)prob * NU * NSTAGE + k * NU + i] = z[i];
    }
#ifdef MPC_PROF
    if (k == 0 && prob == 0) {
        printf("PROF ipm %d:", ipm_total);
        for (int i = 0; i < 24; i++) printf(" %d:%lld", i, pt_[i]);
        printf("\n");
    }
#endif
    if (k == 0) {
        pobj_g[prob] = cost; exit_g[prob] = exit_code; qps_g[prob] = qp_status_acados(qps); reseq_g[prob] = req;
        if (ipm_g) ipm_g[prob] = ipm_total;
    }
    if (mem) {
        if (status != 0) {                                           // Solver_acados_reset + reset_qp_memory (:187-191)
            if (!defer)
                for (int i = k; i < mem_doubles; i += 32) mem[i] = 0.0;
        } else {
            double* m = mem + 1;
            if (k == 0) mem[0] = 2.0;
            if (live) for (int i = 0; i < NX; i++) m[k * NX + i] = pi[i];
            m += (NSTAGE + 1) * NX;
            m += 2 * NSTAGE * NC;
            if (live) for (int i = 0; i < NZ; i++) m[k * NZ + i] = v[i];
        }
    }
}

// Persistent grid of one-problem CTAs; CTAs pull problem indices from a global counter.
__global__ void __launch_bounds__(SPLIT_THREADS, MPC_SPLIT_MIN_CTAS)
mpc_solve_split_kernel(int n, const double* __restrict__ xinit, const double* __restrict__ x0, const double* __restrict__ params,
                       const int* __restrict__ num_iter, int num_iter_all, double* mem, int mem_doubles, double* xtraj,
                       double* utraj, double* pobj, int* exit_code, int* qp_status, double* res_eq, int* ipm_iters,
                       int* work_counter)
{
    extern __shared__ double s_split[];          // SP_DOUBLES
    __shared__ int s_prob;
    const int role = threadIdx.x >> 5;
    for (;;) {
        __syncthreads();                         // the previous problem is finished in every role
        if (threadIdx.x == 0) s_prob = atomicAdd(work_counter, 1);
        __syncthreads();
        const int prob = s_prob;
        if (prob >= n) return;
#if MPC_CHECK
        if (threadIdx.x < 12) s_split[SP_CANARY_AT[threadIdx.x >> 1] + (threadIdx.x & 1)] = CANARY;
        __syncthreads();
#endif
        const int nit_raw = num_iter ? num_iter[prob] : num_iter_all;
        const int nit = nit_raw < 0 ? -nit_raw : nit_raw;      // < 0: completion deferred (see solve_problem)
        if (role == 0)
            split_role_a(prob, xinit, x0, params, nit, mem, mem_doubles, xtraj, utraj, pobj, exit_code, qp_status, res_eq, ipm_iters,
                         s_split, nit_raw < 0);
        else if (role == 1)
            split_role_x<0>(prob, nit, mem, mem_doubles, s_split);
        else if (NXR == 2 && role == 2)
            split_role_x<NXR - 1>(prob, nit, mem, mem_doubles, s_split);
        else
            split_role_b(prob, params, nit, mem, mem_doubles, s_split, role - 1 - NXR);
#if MPC_CHECK
        __syncthreads();
        if (threadIdx.x < 12 && s_split[SP_CANARY_AT[threadIdx.x >> 1] + (threadIdx.x & 1)] != CANARY) MPC_CHECK_FAIL(1);
        if (threadIdx.x == 0) MPC_CHECK_FAIL(4);
#endif
    }
}
