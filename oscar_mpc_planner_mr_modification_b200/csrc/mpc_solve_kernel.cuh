// mpc_solve_kernel.cuh -- one warp solves one N-stage MPC problem (SQP-RTI + interior-point QP), FP64.
//
// Replaces what `Solver::solve()` runs through acados for one GuidanceConstraints homotopy
// (reference: mpc_planner_solver/src/acados_solver_interface.cpp:86-204).  The algorithm contract is
// DESIGN.md "Algorithm contract"; the model-specific device functions come from the emitter
// (solver_generator/generate_cuda_solver.py -> generated/<config>/model.cuh).
//
// Mapping: lane k of the warp owns stage k (k = 0..N, N <= 31).  Everything that is independent per
// stage (RK4 roll-out with sensitivities, cost/constraint linearisation, MIRROR, the per-constraint
// interior-point algebra, residuals) runs lane-parallel with the stage data in registers / local
// memory; stage-to-stage coupling (x_{k+1}, pi_{k+1}, the Riccati recursion) goes through warp
// shuffles; scalar decisions (step length, mu, residual norms, exit) are warp reductions.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#if defined(MPC_TRACE) || defined(MPC_PROF)
#include <cstdio>
#endif

#ifndef MPC_GEN_UNROLL
#define MPC_GEN_UNROLL 0   // general-constraint loops: entries in flight together.  0 = automatic: 2 where the entries' rows come from
                           // thread-local or shared memory (hides their latency), 1 where they come from tensor memory (12-cycle loads:
                           // 261.1 k solves/s against 257.5 / 256.2 / 247.7 k for 2 / 3 / 4 on the 18 432-problem batch)
#endif
#ifndef MPC_SCAN_ALWAYS
#define MPC_SCAN_ALWAYS 0   // 1: substitution sweeps as prefix scans in the throughput instantiation too
#endif
#ifndef MPC_WARPS_PER_CTA
#define MPC_WARPS_PER_CTA 8
#endif
#ifndef MPC_MIN_CTAS
#define MPC_MIN_CTAS 1
#endif
#ifndef MPC_CFG_TAG
#error "compile with -DMPC_CFG_TAG=<config> -DMPC_MODEL_HEADER='\"model.cuh\"'"
#endif
#define MPC_CAT2(a, b) a##b
#define MPC_CAT(a, b) MPC_CAT2(a, b)
#define MPC_NS MPC_CAT(mpck_, MPC_CFG_TAG)   // one namespace per compiled configuration

namespace MPC_NS {
// ---- MPC_CHECK build (diagnostic variant library, never shipped): what compute-sanitizer would look for, since the tool is closed
//      on this pool.  Counters (read through mpcgpu_check_report): 0 index outside an array of the shared-memory placement,
//      1 canary word between shared-memory regions overwritten, 2 roles of the role-split kernel met at different barrier labels,
//      3 exchange slot read with a stale generation, 4 problems checked.
#ifndef MPC_CHECK
#define MPC_CHECK 0
#endif
#if MPC_CHECK
__device__ unsigned long long g_check[8];
#define MPC_CHECK_FAIL(i) atomicAdd(&g_check[i], 1ull)
#define MPCK(n) , (n)
constexpr double CANARY = -7.0e300;
#else
#define MPCK(n)
#endif
// Branch-free reciprocal square root: hardware approximation (rsqrt.approx.f64, ~2^-23) + one cubically convergent
// correction y (1 + e/2 + 3 e^2/8), e = 1 - a y^2 (error below the rounding of a double).  Half the instructions of
// rsqrt() and no slow-path branch, so two of them interleave in one in-order instruction stream (the two Cholesky
// pivots of a Riccati stage); NaN for a < 0 or NaN input, which is what the failure detection relies on.
#ifndef MPC_RSQRT_NB
#define MPC_RSQRT_NB 1
#endif
__device__ __forceinline__ double rsqrt_nb(double a)
{
#if MPC_RSQRT_NB
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double m = a * y;
    const double e = fma(-m, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    return fma(y * e, p, y);
#else
    return rsqrt(a);
#endif
}
// Branch-free reciprocal for the 1/t of every inequality entry (38 per stage and interior-point iteration): hardware
// approximation (rcp.approx.f64, ~2^-23) + one cubically convergent correction y + y e (1 + e), e = 1 - t y.  The library
// division is correctly rounded but costs ~3x the instructions plus a slow-path branch per call; this one is within one
// ulp, which the parity tolerance (1e-6) does not see.  t > 0 here (slacks are clamped at 1e-16); NaN propagates.
__device__ __forceinline__ double rcp_nb(double t)
{
#if MPC_RSQRT_NB
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(t));
    const double e = fma(-t, y, 1.0);
    return fma(y * e, 1.0 + e, y);
#else
    return 1.0 / t;
#endif
}
// the emitted model code takes its reciprocals through this macro
#define MPCGEN_RCP(x) rcp_nb(x)
#include MPC_MODEL_HEADER
using namespace mpcgen;

// One thread per stage: a problem is solved by a GROUP of GW consecutive warps of one CTA (GW = 1 for
// N <= 31, GW = 2 for N <= 63, e.g. the N = 50 CC-MPC configuration).  Stage k = warp-in-group * 32 + lane.
constexpr int GW = (NSTAGE + 1 + 31) / 32;
// Columns of the per-stage shared-memory arrays ([entry][stage slot]): one slot per LIVE lane (stages 0..N) instead of one per lane
// of the group.  Inequality entries exist on path stages only, so the terminal stage's slot is a dummy and the dead lanes of the
// group (lane 31 for N = 30) share it.  For N = 30 this is 31/32 of the footprint -- exactly what lets the right-hand sides d of
// the 24 general entries of c2 sit beside the multipliers and slacks at 8 warps per CTA (they have since moved to tensor memory).
#ifndef MPC_COL_COMPACT
#define MPC_COL_COMPACT 1
#endif
constexpr int LCOL = MPC_COL_COMPACT ? (NSTAGE + 1) : GW * 32;
static_assert(GW == 1 || GW == 2, "thread-per-stage kernel needs N <= 63");
static_assert(MPC_WARPS_PER_CTA % ((NSTAGE + 1 + 31) / 32) == 0, "warps per CTA must be a multiple of the group size");
static_assert(NU == 2, "Riccati input block elimination is written for nu == 2");


constexpr int NPX = NX * (NX + 1) / 2;      // packed P
// Shared-memory placement of per-stage state (one column [entry][thread of the group] per value: conflict free).  What pays is
// what takes REGISTER pressure out of the interior-point loop (255 registers per thread, ~270 doubles of live state: the rest
// is spilled to thread-local memory, which 8 warps x ~60 KB cannot keep in the L1).  Measured on c2_tmpc12 (18 432 problems):
//   general multipliers + slacks + d (147 KB)                                   221 k solves/s   (round-1 layout)
//   + box multipliers and slacks (205 KB)                                       234 k
//   general multipliers + slacks, box multipliers + slacks + 1/t (184 KB)       245 k            <- default
//   (+ g, b: 244 k; box without 1/t but H: 232 k; Jacobian columns x, y instead of d: 217 k; 7 warps with everything: 227 k)
//   ... + Jacobian rows, d and the box multipliers + slacks in TENSOR memory (MPC_TMEM below; 123 KB of shared memory left)  263 k
// The default takes the arrays in that priority order while they fit beside each other for all warps of the CTA:
// MPC_LT_MASK bits: 0 general multipliers, 1 general slacks, 2 right-hand sides d, 3 Jacobian rows C; MPC_BOX_SMEM: 1 = 1/t of the
// box entries, 2 = their multipliers and slacks, 3 = all three.  Either macro can be pinned on the command line.
// Budget per stage slot.  If EVERYTHING (incl. the Jacobian rows) fits into the 227 KB maximum, take it: nothing thread-local is
// left for the L1 to serve (c5: +4 % over the smaller configuration).  Otherwise stay within the 196 KB shared-memory
// configuration: the step to 228 KB costs 32 KB of L1 (60 -> 28 KB) that the thread-local rest (Jacobian rows, spills) needs more
// than d needs shared memory (c2 with d inside at 228 KB 251 k solves/s, without at 196 KB 253.6 k; 250.8 k before compact columns).
#ifndef MPC_SMEM_KB
#define MPC_SMEM_KB 196
#endif
constexpr int SMEM_BUDGET_DOUBLES = (MPC_SMEM_KB * 1024 - 4096) / 8 / ((MPC_WARPS_PER_CTA / GW) * LCOL);
constexpr int SMEM_MAX_DOUBLES = (227 * 1024 - 4096) / 8 / ((MPC_WARPS_PER_CTA / GW) * LCOL);
constexpr int auto_lt_mask()
{
    int used = 2 * NCG + 6 * (NX + NU), mask = 3;
    if (used + NCG + NH * NHS <= SMEM_MAX_DOUBLES) return 15;
    if (used + NCG <= SMEM_BUDGET_DOUBLES) { mask |= 4; used += NCG; }
    if (used + NH * NHS <= SMEM_BUDGET_DOUBLES) mask |= 8;
    return mask;
}
#ifdef MPC_LT_SMEM
#ifndef MPC_LT_MASK          // (older knob: 0 none, 1 multipliers + slacks, 2 + d, 3 + C)
#define MPC_LT_MASK (MPC_LT_SMEM == 0 ? 0 : (MPC_LT_SMEM == 1 ? 3 : (MPC_LT_SMEM == 2 ? 7 : 15)))
#endif
#endif
#ifndef MPC_LT_MASK
#define MPC_LT_MASK auto_lt_mask()
#endif
constexpr bool LT_SMEM = MPC_LT_MASK != 0;
constexpr bool LT_LAM = (MPC_LT_MASK & 1) != 0, LT_T = (MPC_LT_MASK & 2) != 0, LT_D = (MPC_LT_MASK & 4) != 0, LT_C = (MPC_LT_MASK & 8) != 0;
constexpr int LT_OFF_LAM = 0, LT_OFF_T = LT_OFF_LAM + (LT_LAM ? NCG : 0), LT_OFF_D = LT_OFF_T + (LT_T ? NCG : 0),
              LT_OFF_C = LT_OFF_D + (LT_D ? NCG : 0);
#ifndef MPC_BOX_SMEM
#define MPC_BOX_SMEM 3
#endif
// MPC_C_SPLIT: the first MPC_C_SPLIT support columns of every Jacobian row in shared memory, the remaining ones thread-local and
// SKIPPED when they are zero in every lane of the warp (e.g. the psi column with a zero disc offset) -- a third of the loads of
// the inequality passes then never leave the SM although the full rows do not fit beside the multipliers (needs !LT_C)
#ifndef MPC_C_SPLIT
#define MPC_C_SPLIT 0
#endif
constexpr int CSPL = (MPC_C_SPLIT > NHS ? NHS : MPC_C_SPLIT);
// MPC_C_ZSKIP (default off: measured 240.5 k vs 246.0 k solves/s, the predication costs more than the loads it saves): the LAST support column of the Jacobian rows (psi for the shipped modules: it only enters through the
// disc offset) is not loaded in the inequality passes while it is zero in every lane of the warp -- decided once per linearisation
#ifndef MPC_C_ZSKIP
#define MPC_C_ZSKIP 0
#endif
constexpr int CTAIL0 = CSPL > 0 ? CSPL : ((MPC_C_ZSKIP != 0 && NHS > 1) ? NHS - 1 : NHS);      // first column of the skippable tail
static_assert(CSPL == 0 || !LT_C, "MPC_C_SPLIT needs the Jacobian rows outside shared memory (MPC_LT_MASK bit 3 clear)");
constexpr int LT_OFF_CS = LT_OFF_C + (LT_C ? NH * NHS : 0);
constexpr int LT_ENTRY_DOUBLES = LT_OFF_CS + NH * CSPL;
#ifndef MPC_TMEM
#define MPC_TMEM 2
#endif
constexpr int TM_ECOLS = 8;                      // tensor memory: 32-bit columns per general entry = (NHS + 1) doubles, padded
constexpr bool TMEM_CD = (MPC_TMEM != 0) && !LT_C && !LT_D && NCG > 0 && NCG * TM_ECOLS <= 256 && NHS + 1 <= 4 && GW == 1 && CSPL == 0 && MPC_C_ZSKIP == 0 &&
                         MPC_WARPS_PER_CTA == 8;
// MPC_TMEM >= 2: the multipliers and slacks of the box entries go to tensor memory as well (variable i = 8 columns {lam_l, t_l, lam_u,
// t_u} behind the general entries); their 4 (NX + NU) shared-memory columns are not allocated, so the SM runs the next smaller
// shared-memory configuration and the L1 -- which serves the register spills -- doubles.
constexpr bool TMEM_BOX = TMEM_CD && (MPC_TMEM >= 2) && MPC_BOX_SMEM == 3 && (NCG * TM_ECOLS + (NX + NU) * 8 <= 256);
constexpr int TM_BOX_COL0 = NCG * TM_ECOLS;
constexpr int BOX_SM_DOUBLES = TMEM_BOX ? 2 * (NX + NU) : (MPC_BOX_SMEM == 1 ? 2 * (NX + NU) : (MPC_BOX_SMEM == 2 ? 4 * (NX + NU) : (MPC_BOX_SMEM == 3 ? 6 * (NX + NU) : 0)));      // 1: 1/t; 2: multipliers + slacks; 3: all three
// MPC_QP_SMEM: per-stage QP data that the interior-point loop only READS, once per iteration in pass DA, parked in shared memory
// after the linearisation instead of staying in registers (where they are spilled): bit 0 g and b (12 doubles), bit 1 H (28)
// Default (-1): g and b where the box multipliers have left shared memory for tensor memory -- there the shared-memory configuration
// has room and the L1 is not shared with thread-local arrays any more: 270.2 k solves/s against 263.9 k (H as well: 261.8 k, no gain);
// before tensor memory the same move was neutral (244 against 245 k).
#ifndef MPC_QP_SMEM
#define MPC_QP_SMEM -1
#endif
constexpr int QP_SMEM_EFF = (MPC_QP_SMEM) >= 0 ? (MPC_QP_SMEM) : (TMEM_BOX ? 1 : 0);
constexpr bool QPS_GB = (QP_SMEM_EFF & 1) != 0, QPS_H = (QP_SMEM_EFF & 2) != 0;
constexpr int QP_OFF_G = LT_ENTRY_DOUBLES + BOX_SM_DOUBLES, QP_OFF_B = QP_OFF_G + (QPS_GB ? NZ : 0), QP_OFF_H = QP_OFF_B + (QPS_GB ? NX : 0);
// MPC_RIC_SMEM: outputs of the Riccati factorisation that are produced by ONE serial step and read much later: bit 0 the
// cost-to-go matrix P (15 doubles; used by the multiplier step dpi = P dx + p at the very end of the iteration), bit 1 P+ rb (5)
#ifndef MPC_RIC_SMEM
#define MPC_RIC_SMEM 0
#endif
constexpr bool RIC_P = (MPC_RIC_SMEM & 1) != 0, RIC_PRB = (MPC_RIC_SMEM & 2) != 0;
constexpr int RIC_OFF_P = QP_OFF_H + (QPS_H ? NPK : 0), RIC_OFF_PRB = RIC_OFF_P + (RIC_P ? NPX : 0);
constexpr int LT_DOUBLES = (RIC_OFF_PRB + (RIC_PRB ? NX : 0)) * LCOL;
static_assert(!MPC_COL_COMPACT || MPC_RIC_SMEM == 0, "compact columns: the terminal stage's slot is shared with the dead lanes");
constexpr int GEN_UNROLL = MPC_GEN_UNROLL > 0 ? MPC_GEN_UNROLL : (TMEM_CD ? 1 : 2);
constexpr int LT_STRIDE = LT_DOUBLES + (MPC_CHECK ? GW * 32 : 0);      // MPC_CHECK: a row of canaries behind every group's region      // per group: lam, t (d, C) of the general entries (+ 1/t of the boxes)
// column accessor: stride GW*32 doubles in shared memory ([entry][thread of the group]), stride 1 for a thread-local array
template <bool SM>
struct LtCol {
    double* p;
#if MPC_CHECK
    int n;
#endif
    __device__ __forceinline__ double& operator[](int e) const
    {
#if MPC_CHECK
        if (e < 0 || e >= n) { MPC_CHECK_FAIL(0); e = 0; }
#endif
        return p[SM ? e * LCOL : e];
    }
};
// Jacobian rows of the general constraints: at(r, a) = d h_r / d z_HSUP[a].  Plain (all in one place, shared or thread-local)
// or split by column (MPC_C_SPLIT).  operator[] (flat index r * NHS + a) is what the emitted con_lin writes through.
template <bool SM>
struct CRows {
    double* p;          // flat rows: shared-memory column accessor stride or thread-local
    double* sm;         // split: [NH * CSPL][threads of the group] in shared memory
    bool tail_zero;     // split: the thread-local tail columns are zero in every lane (set after the linearisation)
    __device__ __forceinline__ double& operator[](int i) const
    {
#if MPC_CHECK
        if (i < 0 || i >= NH * NHS) { MPC_CHECK_FAIL(0); i = 0; }
#endif
        if constexpr (CSPL > 0) {
            const int r = i / NHS, a = i - r * NHS;
            if (a < CSPL) return sm[(r * CSPL + a) * LCOL];
            return p[i];
        } else {
            return p[SM ? i * LCOL : i];
        }
    }
    __device__ __forceinline__ double at(int r, int a) const
    {
#if MPC_CHECK
        if (r < 0 || r >= NH || a < 0 || a >= NHS) { MPC_CHECK_FAIL(0); r = 0; a = 0; }
#endif
        if constexpr (CSPL > 0) {
            if (a < CSPL) return sm[(r * CSPL + a) * LCOL];
            return tail_zero ? 0.0 : p[r * NHS + a];
        } else {
            if (a >= CTAIL0 && tail_zero) return 0.0;
            return p[SM ? (r * NHS + a) * LCOL : r * NHS + a];
        }
    }
};
struct SmemCol {
    double* p;
#if MPC_CHECK
    int n;
#endif
    __device__ __forceinline__ double& operator[](int e) const
    {
#if MPC_CHECK
        if (e < 0 || e >= n) { MPC_CHECK_FAIL(0); e = 0; }
#endif
        return p[e * LCOL];
    }
};
constexpr int NCB = 2 * NZ;                 // box entries: lower(z_i) i<NZ, then upper(z_i)
constexpr int NC = NCB + NCG;               // inequality entries per path stage
constexpr unsigned FULL = 0xffffffffu;

// ---- algorithm constants: identical to oracle/mpc_oracle.c -------------------------------------
constexpr double REG_EPS = 1e-4;
constexpr int IPM_ITER_MAX = 50;
constexpr double IPM_TOL = 1e-5;
constexpr double IPM_MU0 = 10.0;
constexpr double IPM_THR0 = 0.1;
constexpr double IPM_ALPHA_MIN = 1e-12;
constexpr double IPM_LAM_MIN = 1e-16;
constexpr double IPM_T_MIN = 1e-16;
constexpr double IPM_STEP_SCALE = 0.995;
constexpr double RES_EQ_MAX = 1e-2;
constexpr int JACOBI_MAX_SWEEPS = 30;
constexpr double JACOBI_TOL = 1e-24;   // stop when sum offdiag^2 <= tol * sum all^2
// inequality entries over the whole horizon: u box + general on N stages, x box on stages 1..N-1
constexpr int IPM_COUNT = NSTAGE * (2 * NU + NCG) + (NSTAGE - 1) * 2 * NX;

// AcadosInfo::qp_status is read with ocp_nlp_get("qp_status") and decoded by Solver::explainExitFlag
// (acados_solver_interface.cpp:409-420) in acados' return-value numbering: 2 max iterations, 3 minimal step, 4 NaN.
// The interior-point loop keeps the HPIPM-raw code (0 ok, 1 max-iter, 2 min-step, 3 NaN); outputs carry the acados one.
__host__ __device__ constexpr int qp_status_acados(int raw) { return raw == 0 ? 0 : raw + 1; }

__host__ __device__ constexpr int pk(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

__device__ __forceinline__ double shfl(double v, int src) { return __shfl_sync(FULL, v, src); }
__device__ __forceinline__ double shfl_down1(double v) { return __shfl_down_sync(FULL, v, 1); }
// max that PROPAGATES NaN (fmax would drop it): a NaN anywhere must surface as QP status 3
// for NON-NEGATIVE arguments (|x|): IEEE-754 doubles order like unsigned integers and every NaN pattern is
// larger than +inf, so an integer max is a NaN-propagating max -- on the ALU pipe, not the FP64 pipe.
__device__ __forceinline__ double nanmax(double a, double b)
{
    const unsigned long long ua = (unsigned long long)__double_as_longlong(a) & 0x7fffffffffffffffull;
    const unsigned long long ub = (unsigned long long)__double_as_longlong(b) & 0x7fffffffffffffffull;
    return __longlong_as_double((long long)(ua > ub ? ua : ub));
}
__device__ __forceinline__ double clamp_lo(double x, double lo) { return (x < lo) ? lo : x; }   // keeps NaN
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_min(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// ---- group primitives: warp shuffles inside a warp, shared-memory exchange + named barrier across the
//      GW warps of a problem.  For GW == 1 every function reduces to the plain warp intrinsic.
constexpr int XCH = 32;                          // exchange doubles per warp (an affine map: 30)
// ---- Tensor memory as a third on-chip store (MPC_TMEM, thread-per-stage throughput instantiation).  The engine has no MMA, so
//      the SM's 256 KB of tensor memory are free: 128 lanes x 512 columns x 32 bit; warp w of the CTA reaches lanes
//      32 (w % 4) .. +31, and with 8 warps every warp gets 256 columns = 128 doubles per thread, 12-cycle loads.  The Jacobian rows
//      and right-hand sides of the general entries (written once per linearisation, read in every pass) live there instead of in
//      thread-local memory: entry e = 8 columns {c0, c1, c2, d}.  tcgen05.ld / .st are warp-collective (.sync.aligned): the entry
//      loops then run for all 32 lanes and only the commits are predicated.
__device__ __forceinline__ void tmem_ld8(unsigned taddr, double& a, double& b, double& c, double& d)
{
    unsigned r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    a = __hiloint2double((int)r1, (int)r0); b = __hiloint2double((int)r3, (int)r2);
    c = __hiloint2double((int)r5, (int)r4); d = __hiloint2double((int)r7, (int)r6);
}
__device__ __forceinline__ void tmem_st8(unsigned taddr, double a, double b, double c, double d)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(__double2loint(a)), "r"(__double2hiint(a)), "r"(__double2loint(b)), "r"(__double2hiint(b)),
                    "r"(__double2loint(c)), "r"(__double2hiint(c)), "r"(__double2loint(d)), "r"(__double2hiint(d)) : "memory");
}
struct Grp {
    unsigned tmem = 0;                           // MPC_TMEM: tensor-memory address of this warp's slot (lane quarter, column base)
    double* xch;                                 // [GW][XCH] shared-memory exchange area of this group
    int gid, wig;                                // group index inside the CTA, warp index inside the group
    __device__ __forceinline__ void sync() const
    {
        if constexpr (GW == 1) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "r"(GW * 32) : "memory");
    }
    // out[i] <- in[i] of the thread owning stage k+1 (undefined for the last stage)
    template <int n> __device__ __forceinline__ void shift_down(const double (&in)[n], double (&out)[n]) const
    {
#pragma unroll
        for (int i = 0; i < n; i++) out[i] = __shfl_down_sync(FULL, in[i], 1);
        if constexpr (GW > 1) {
            const int lane = threadIdx.x & 31;
            if (wig == 1 && lane == 0) {
#pragma unroll
                for (int i = 0; i < n; i++) xch[i] = in[i];
            }
            sync();
            if (wig == 0 && lane == 31) {
#pragma unroll
                for (int i = 0; i < n; i++) out[i] = xch[i];
            }
            sync();
        }
    }
    // out[i] <- in[i] of the thread owning stage k-1 (zero for stage 0)
    template <int n> __device__ __forceinline__ void shift_up(const double (&in)[n], double (&out)[n]) const
    {
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int i = 0; i < n; i++) {
            const double t = __shfl_up_sync(FULL, in[i], 1);
            out[i] = (lane == 0) ? 0.0 : t;
        }
        if constexpr (GW > 1) {
            if (wig == 0 && lane == 31) {
#pragma unroll
                for (int i = 0; i < n; i++) xch[i] = in[i];
            }
            sync();
            if (wig == 1 && lane == 0) {
#pragma unroll
                for (int i = 0; i < n; i++) out[i] = xch[i];
            }
            sync();
        }
    }
    // cross-warp stage of a reduction: v is already warp-reduced (identical in all lanes of a warp)
    template <int OP> __device__ __forceinline__ double across(double v) const      // OP 0 sum, 1 min, 2 nan-max
    {
        if constexpr (GW > 1) {
            if ((threadIdx.x & 31) == 0) xch[wig * XCH] = v;
            sync();
            const double a = xch[0], b = xch[XCH];
            sync();
            v = (OP == 0) ? a + b : ((OP == 1) ? fmin(a, b) : nanmax(a, b));
        }
        return v;
    }
    __device__ __forceinline__ double sum(double v) const { return across<0>(warp_sum(v)); }
    __device__ __forceinline__ double min(double v) const { return across<1>(warp_min(v)); }
    __device__ __forceinline__ double max(double v) const { return across<2>(warp_max(v)); }
    __device__ __forceinline__ bool any(bool p) const
    {
        bool r = __any_sync(FULL, p);
        if constexpr (GW > 1) r = across<2>(r ? 1.0 : 0.0) > 0.5;
        return r;
    }
    __device__ __forceinline__ bool all(bool p) const
    {
        bool r = __all_sync(FULL, p);
        if constexpr (GW > 1) r = across<1>(r ? 1.0 : 0.0) > 0.5;
        return r;
    }
    // sum of one value per stage in stage order (warp-local stage order, then warp 0 + warp 1)
    __device__ __forceinline__ double ordered_sum(double v) const
    {
        double acc = 0.0;
#pragma unroll 1
        for (int l = 0; l < 32; l++) acc += __shfl_sync(FULL, v, l);
        return across<0>(acc);
    }
};

// ---- Parallel-in-stage substitution sweeps.  With the factorisation in hand both the forward sweep
//      dx_{k+1} = Acl_k dx_k + bcl_k  and the backward vector sweep  p_k = Acl_k' p_{k+1} + c_k  are affine
//      recurrences with the closed-loop matrices Acl_k = A_k + B_k K_k, K_k = -Luu^-T Lxu'.  Affine maps compose
//      associatively, so the recurrences are evaluated as a Hillis-Steele scan over the lanes (log2(32) levels of
//      5x5 products with all lanes busy) instead of 30 dependent single-lane steps.
struct Aff {
    double M[NX * NX], c[NX];                    // x -> M x + c
};
// a <- a o r  (apply r first):  M = a.M r.M,  c = a.M r.c + a.c ; row by row, in place
__device__ __forceinline__ void aff_compose(Aff& a, const Aff& r, bool with_matrix)
{
#pragma unroll
    for (int i = 0; i < NX; i++) {
        double row[NX], ci = a.c[i];
#pragma unroll
        for (int l = 0; l < NX; l++) ci += a.M[i * NX + l] * r.c[l];
        a.c[i] = ci;
        if (with_matrix) {
#pragma unroll
            for (int j = 0; j < NX; j++) {
                double m = 0.0;
#pragma unroll
                for (int l = 0; l < NX; l++) m += a.M[i * NX + l] * r.M[l * NX + j];
                row[j] = m;
            }
#pragma unroll
            for (int j = 0; j < NX; j++) a.M[i * NX + j] = row[j];
        }
    }
}
// Inclusive scan over the stages.  UP: lane k ends with a_k o a_{k-1} o ... o a_0 (forward sweep);
// !UP: lane k ends with a_k o a_{k+1} o ... (backward sweep).  Only the vector part is needed afterwards.
template <bool UP>
__device__ __noinline__ void aff_scan(Aff* ap, const Grp grp)      // ONE copy of the scan per direction: code size matters
{
    Aff a = *ap;
    const int lane = threadIdx.x & 31;
#pragma unroll 1
    for (int d = 1; d < 32; d <<= 1) {
        Aff r;
#pragma unroll
        for (int i = 0; i < NX * NX; i++) r.M[i] = UP ? __shfl_up_sync(FULL, a.M[i], d) : __shfl_down_sync(FULL, a.M[i], d);
#pragma unroll
        for (int i = 0; i < NX; i++) r.c[i] = UP ? __shfl_up_sync(FULL, a.c[i], d) : __shfl_down_sync(FULL, a.c[i], d);
        const bool valid = UP ? (lane >= d) : (lane + d < 32);
        if (valid) aff_compose(a, r, GW > 1 || d < 16);
    }
    if constexpr (GW > 1) {
        // second warp (UP) / first warp (!UP) continues from the other warp's total
        const bool src = UP ? (grp.wig == 0 && lane == 31) : (grp.wig == 1 && lane == 0);
        double* x = grp.xch;                     // needs NX*NX + NX <= GW * XCH doubles
        static_assert(NX * NX + NX <= GW * XCH, "exchange area too small for an affine map");
        if (src) {
#pragma unroll
            for (int i = 0; i < NX * NX; i++) x[i] = a.M[i];
#pragma unroll
            for (int i = 0; i < NX; i++) x[NX * NX + i] = a.c[i];
        }
        grp.sync();
        if (UP ? (grp.wig == 1) : (grp.wig == 0)) {
            Aff r;
#pragma unroll
            for (int i = 0; i < NX * NX; i++) r.M[i] = x[i];
#pragma unroll
            for (int i = 0; i < NX; i++) r.c[i] = x[NX * NX + i];
            aff_compose(a, r, false);
        }
        grp.sync();
    }
#pragma unroll
    for (int i = 0; i < NX; i++) ap->c[i] = a.c[i];      // only the vector part is used by the callers
}
// closed-loop matrix of one stage from the factorisation: Acl = A + B K, K = -Luu^-T Lxu'
__device__ __forceinline__ void closed_loop(const double* Wv, const double* Lx0, const double* Lx1, double L10, double iL0, double iL1,
                                            double* Acl /*NX*NX*/, double* Bd /*NX*NU*/)
{
    double Wd[NX * NZ], K0[NX], K1[NX];
    w_to_dense(Wv, Wd);
#pragma unroll
    for (int j = 0; j < NX; j++) {
        K1[j] = -Lx1[j] * iL1;
        K0[j] = -(Lx0[j] + L10 * K1[j]) * iL0;
    }
#pragma unroll
    for (int i = 0; i < NX; i++) {
        Bd[i * NU] = Wd[i * NZ]; Bd[i * NU + 1] = Wd[i * NZ + 1];
#pragma unroll
        for (int j = 0; j < NX; j++) Acl[i * NX + j] = Wd[i * NZ + NU + j] + Wd[i * NZ] * K0[j] + Wd[i * NZ + 1] * K1[j];
    }
}
// forward sweep by scan: returns du (2) and dx (NX) of this lane's stage in dvec = [du; dx]
__device__ __noinline__ void forward_scan(const double* Wv, const double* Lx0, const double* Lx1, double L10, double iL0, double iL1,
                                             const double* lv, const double* rb, bool path, double* dvec, const Grp grp)
{
    Aff a;
    double Bd[NX * NU];
    if (path) {
        closed_loop(Wv, Lx0, Lx1, L10, iL0, iL1, a.M, Bd);
        const double k1 = -lv[1] * iL1, k0 = -(lv[0] + L10 * k1) * iL0;      // kff = -Luu^-T l
#pragma unroll
        for (int i = 0; i < NX; i++) a.c[i] = rb[i] + Bd[i * NU] * k0 + Bd[i * NU + 1] * k1;
    } else {                                     // terminal / idle lanes: identity
#pragma unroll
        for (int i = 0; i < NX * NX; i++) a.M[i] = (i / NX == i % NX) ? 1.0 : 0.0;
#pragma unroll
        for (int i = 0; i < NX; i++) a.c[i] = 0.0;
    }
    aff_scan<true>(&a, grp);                     // lane k: a.c = dx_{k+1}   (dx_0 = 0)
    double dxn[NX], dx[NX];
#pragma unroll
    for (int i = 0; i < NX; i++) dxn[i] = a.c[i];
    grp.shift_up(dxn, dx);                       // dx_k from lane k-1 (stage 0: zero)
#pragma unroll
    for (int i = 0; i < NX; i++) dvec[NU + i] = dx[i];
    double r0 = lv[0], r1 = lv[1];               // du = -Luu^-T (Lxu' dx + l)
#pragma unroll
    for (int j = 0; j < NX; j++) { r0 += Lx0[j] * dx[j]; r1 += Lx1[j] * dx[j]; }
    dvec[1] = path ? -r1 * iL1 : 0.0;
    dvec[0] = path ? -(r0 + L10 * dvec[1]) * iL0 : 0.0;
}


// ---- Cooperative Riccati recursion.  The recursion is sequential in the stage index, so with one thread per stage
//      only one lane of the warp would work at a time.  Instead the per-stage QP data are parked in shared memory
//      (one block of RSTRIDE doubles per stage) and ALL 32 lanes execute stage s together: every lane owns one entry
//      of the small matrix products (5-term dot products, uniform code, operands read from shared memory), with a
//      __syncwarp between the dependent steps.  Block layout (offsets in doubles):
//        RO_G   packed 8x8 lower triangle of the augmented matrix [[Ht, gt], [gt', -]] (35 entries: Ht 28, gt 7).
//               After the factorisation of the stage, in place: iL0 @(0,0), L10 @(1,0), iL1 @(1,1), Lxu[i][0..1]
//               @(2+i,0..1), P @(2+i,2+j), l = Luu^-1 q_u @(7,0..1), p @(7,2+i)
//        RO_B   [W | rb], NX x (NZ+1) row-major      RO_PRB  P+ rb      RO_DZ  step of the stage [du; dx]
// HPIPM "BALANCE" mode behaviour that DESIGN.md section 4 leaves out by default: the conditional Mehrotra predictor-corrector
// (fall back to the pure centering direction when the corrected step would leave the duality measure above twice the
// predictor's).  1 = on, in this kernel and in the oracle (-DMPC_HPIPM_BALANCE); the role-split kernel is not built then.
#ifndef MPC_HPIPM_BALANCE
#define MPC_HPIPM_BALANCE 0
#endif
constexpr bool BALANCE = MPC_HPIPM_BALANCE != 0;
#ifndef MPC_COOP
#define MPC_COOP 0   // thread-per-stage kernel: 1 = cooperative Riccati recursion (all lanes on one stage, shared-memory workspace)
#endif
constexpr bool COOP = (MPC_COOP != 0) && (((NSTAGE + 1 + 31) / 32) == 1);
constexpr int NB = NZ + 1;
constexpr int RO_G = 0;
constexpr int RO_Q = RO_G + NPK;                 // row 7 of the packed augmented matrix: pk(NZ, j) == NPK + j
constexpr int RO_B = RO_Q + NZ;
constexpr int RO_PRB = RO_B + NX * NB;
constexpr int RO_DZ = RO_PRB + NX;
constexpr int RO_END = RO_DZ + NZ;
constexpr int RSTRIDE = RO_END | 1;              // odd stride: lane-parallel accesses (lane = stage) are conflict-free
constexpr int RS_DOUBLES = (NSTAGE + 1) * RSTRIDE + NX * NB;   // per problem, incl. the T = P+ [W | rb] scratch
static_assert(pk(NZ, 0) == NPK, "augmented row must follow the packed Hessian");
// lane roles of the cooperative factor step (lanes 0..NPK-1 own G(i,j), the next ceil(NZ/2) lanes two entries of q each;
// step C: NPX lanes own P(i,j), NX lanes a row of Lxu and p, one lane the pivots) for any nx, nu = 2 that fits a warp
static_assert(NPK + (NZ + 1) / 2 <= 32 && NPX + NX < 32, "cooperative Riccati step: one warp must cover the lane roles");

__device__ __forceinline__ void tri_unpack(int e, int& i, int& j)
{
    i = 0;
    while ((i + 1) * (i + 2) / 2 <= e) i++;
    j = e - i * (i + 1) / 2;
}

// Same recursion, shorter dependency chain per stage (the recursion is latency bound: 30 dependent stages):
//  * one step for G = M + [W|rb]' P+ [W|rb]: every lane forms t = P+ c for its own column c of [W|rb] redundantly
//    (25 independent-enough FMAs) and then its entry; no round trip through shared memory for T
//  * the two Cholesky pivots run side by side: iL1 = rsqrt(g00 g11 - g10^2) * g00 * iL0
// Lanes 0..27 own G(i,j); lanes 28..31 own q_j, two each (y = p+ + P+ rb); lane 28 also stores P+ rb.
__device__ __noinline__ void riccati_factor_coop(double* __restrict__ rs)
{
    const int lane = threadIdx.x & 31;
    const bool ql = lane >= NPK;                        // q lanes
    int gi, gj;
    tri_unpack(lane < NPK ? lane : 0, gi, gj);
    const int cc = ql ? NZ : gi;                        // column of [W|rb] that t is formed with
    const int jq = 2 * (lane - NPK);
    const bool qact = ql && jq < NZ;                    // (q lanes beyond ceil(NZ/2) idle: they recompute q_0 and store nothing)
    const int j0 = ql ? (qact ? jq : 0) : gj;           // output columns
    const int j1 = ql ? ((qact && jq + 1 < NZ) ? jq + 1 : j0) : gj;
    const int o0 = ql ? RO_Q + j0 : RO_G + lane;        // in/out offsets of the two entries
    const int o1 = ql ? RO_Q + j1 : o0;
    const bool st0 = !ql || qact, st1 = qact && (jq + 1 < NZ);
    int ci = 0, cj = 0;                                 // step C: lanes 0..14 own P(ci,cj); lanes 15..19 own row ci of Lxu and p
    if (lane < NPX) tri_unpack(lane, ci, cj);
    else if (lane < NPX + NX) ci = lane - NPX;
#pragma unroll 1
    for (int s = NSTAGE - 1; s >= 0; s--) {
        double* __restrict__ blk = rs + s * RSTRIDE;
        const double* __restrict__ nxt = blk + RSTRIDE;
        {
            double op[NPX + NX + 3 * NX + 2];           // P+ | p+ | c | b0 | b1 | M0 M1
#pragma unroll
            for (int l = 0; l < NX; l++)
#pragma unroll
                for (int m = 0; m <= l; m++) op[pk(l, m)] = nxt[RO_G + pk(NU + l, NU + m)];
#pragma unroll
            for (int l = 0; l < NX; l++) {
                op[NPX + l] = nxt[RO_Q + NU + l];
                op[NPX + NX + l] = blk[RO_B + l * NB + cc];
                op[NPX + 2 * NX + l] = blk[RO_B + l * NB + j0];
                op[NPX + 3 * NX + l] = blk[RO_B + l * NB + j1];
            }
            op[NPX + 4 * NX] = blk[o0];
            op[NPX + 4 * NX + 1] = blk[o1];
            double t[NX];                               // (loads and FMAs are left to the compiler's interleaving: the 25 independent
#pragma unroll                                           //  FMAs overlap the load latency; forcing the loads first cost 120 cycles per stage)
            for (int l = 0; l < NX; l++) t[l] = 0.0;
#pragma unroll
            for (int m = 0; m < NX; m++)
#pragma unroll
                for (int l = 0; l < NX; l++) t[l] += op[pk(l, m)] * op[NPX + NX + m];
            if (lane == NPK) {
#pragma unroll
                for (int l = 0; l < NX; l++) blk[RO_PRB + l] = t[l];
            }
            double a0 = op[NPX + 4 * NX], a1 = op[NPX + 4 * NX + 1];
#pragma unroll
            for (int l = 0; l < NX; l++) {
                const double y = ql ? t[l] + op[NPX + l] : t[l];
                a0 += op[NPX + 2 * NX + l] * y;
                a1 += op[NPX + 3 * NX + l] * y;
            }
            if (st0) blk[o0] = a0;
            if (st1) blk[o1] = a1;
        }
        __syncwarp();
        {
            double op[11];
            op[0] = blk[RO_G + pk(0, 0)]; op[1] = blk[RO_G + pk(1, 0)]; op[2] = blk[RO_G + pk(1, 1)];
            op[3] = blk[RO_G + pk(NU + ci, 0)]; op[4] = blk[RO_G + pk(NU + ci, 1)];
            op[5] = blk[RO_G + pk(NU + cj, 0)]; op[6] = blk[RO_G + pk(NU + cj, 1)];
            op[7] = blk[RO_G + pk(NU + ci, NU + cj)];
            op[8] = blk[RO_Q]; op[9] = blk[RO_Q + 1]; op[10] = blk[RO_Q + NU + ci];
            const double g00 = op[0], g10 = op[1], g11 = op[2];
            const double gi0 = op[3], gi1 = op[4], gj0 = op[5], gj1 = op[6], gxx = op[7], q0 = op[8], q1 = op[9], qi = op[10];
            const double iL0 = rsqrt_nb(g00), dd = rsqrt_nb(g00 * g11 - g10 * g10);
            const double L10 = g10 * iL0, iL1 = dd * (g00 * iL0);
            const double li0 = gi0 * iL0, li1 = (gi1 - li0 * L10) * iL1;
            const double lj0 = gj0 * iL0, lj1 = (gj1 - lj0 * L10) * iL1;
            const double pij = gxx - li0 * lj0 - li1 * lj1;
            const double l0 = q0 * iL0, l1 = (q1 - L10 * l0) * iL1;
            const double pvi = qi - li0 * l0 - li1 * l1;
            __syncwarp();
            if (lane < NPX) blk[RO_G + pk(NU + ci, NU + cj)] = pij;
            else if (lane < NPX + NX) {
                blk[RO_G + pk(NU + ci, 0)] = li0; blk[RO_G + pk(NU + ci, 1)] = li1; blk[RO_Q + NU + ci] = pvi;
            } else if (lane == NPX + NX) {
                blk[RO_G + pk(0, 0)] = iL0; blk[RO_G + pk(1, 0)] = L10; blk[RO_G + pk(1, 1)] = iL1;
                blk[RO_Q] = l0; blk[RO_Q + 1] = l1;
            }
        }
        __syncwarp();
    }
}

// Cooperative substitution sweeps of the thread-per-stage kernel (MPC_COOP): all lanes on one stage, operands in the
// shared-memory blocks, one __syncwarp per stage.  ~45 / ~35 warp-instructions per stage instead of ~60 single-lane ones
// AND no register-resident factor (P, Lxu, p, P+ rb: 40 doubles per lane less).
// forward: du_s = -Luu^-T (Lxu' dx_s + l_s), dx_{s+1} = rb_s + W_s [du_s; dx_s], dx_0 = 0  ->  RO_DZ of every block
__device__ __noinline__ void riccati_forward_coop(double* __restrict__ rs)
{
    const int lane = threadIdx.x & 31;
    const int i = lane < NX ? lane : 0;
    if (lane < NX) rs[RO_DZ + NU + lane] = 0.0;
    __syncwarp();
#pragma unroll 1
    for (int s = 0; s < NSTAGE; s++) {
        double* __restrict__ blk = rs + s * RSTRIDE;
        double dx[NX], l0c[NX], l1c[NX], wx[NX];
#pragma unroll
        for (int j = 0; j < NX; j++) {
            dx[j] = blk[RO_DZ + NU + j];
            l0c[j] = blk[RO_G + pk(NU + j, 0)];
            l1c[j] = blk[RO_G + pk(NU + j, 1)];
            wx[j] = blk[RO_B + i * NB + NU + j];
        }
        const double l0 = blk[RO_Q], l1 = blk[RO_Q + 1];
        const double iL0 = blk[RO_G + pk(0, 0)], L10 = blk[RO_G + pk(1, 0)], iL1 = blk[RO_G + pk(1, 1)];
        const double rbi = blk[RO_B + i * NB + NZ], wu0 = blk[RO_B + i * NB], wu1 = blk[RO_B + i * NB + 1];
        double r0 = l0, r1 = l1, acc = rbi;
#pragma unroll
        for (int j = 0; j < NX; j++) {
            r0 += l0c[j] * dx[j];
            r1 += l1c[j] * dx[j];
            acc += wx[j] * dx[j];
        }
        const double du1 = -r1 * iL1;
        const double du0 = -(r0 + L10 * du1) * iL0;
        acc += wu0 * du0 + wu1 * du1;
        if (lane < NX) blk[RSTRIDE + RO_DZ + NU + i] = acc;
        else if (lane == NX) { blk[RO_DZ] = du0; blk[RO_DZ + 1] = du1; }
        __syncwarp();
    }
}
// backward vector sweep of a new right-hand side (factorisation reused).  In: gt in row 7 of every block (block N:
// p_N = gt_x).  q = gt + W' (p+ + P+ rb), l = Luu^-1 q_u, p = q_x - Lxu l  ->  row 7, in place
__device__ __noinline__ void riccati_backvec_coop(double* __restrict__ rs)
{
    const int lane = threadIdx.x & 31;
    const int j = lane < NZ ? lane : 0;
#pragma unroll 1
    for (int s = NSTAGE - 1; s >= 0; s--) {
        double* __restrict__ blk = rs + s * RSTRIDE;
        const double* __restrict__ nxt = blk + RSTRIDE;
        double wc[NX], y[NX];
#pragma unroll
        for (int l = 0; l < NX; l++) {
            wc[l] = blk[RO_B + l * NB + j];
            y[l] = nxt[RO_Q + NU + l] + blk[RO_PRB + l];
        }
        const double iL0 = blk[RO_G + pk(0, 0)], L10 = blk[RO_G + pk(1, 0)], iL1 = blk[RO_G + pk(1, 1)];
        const double lj0 = blk[RO_G + pk(j, 0)], lj1 = blk[RO_G + pk(j, 1)];      // lanes 2..6: row j - NU of Lxu
        double acc = blk[RO_Q + j];
#pragma unroll
        for (int l = 0; l < NX; l++) acc += wc[l] * y[l];
        const double q0 = __shfl_sync(FULL, acc, 0), q1 = __shfl_sync(FULL, acc, 1);
        const double l0 = q0 * iL0;
        const double l1 = (q1 - L10 * l0) * iL1;
        const double pvi = acc - lj0 * l0 - lj1 * l1;
        if (lane < NU) blk[RO_Q + lane] = lane == 0 ? l0 : l1;
        else if (lane < NZ) blk[RO_Q + lane] = pvi;
        __syncwarp();
    }
}

// ---- Substitution sweeps in closed-loop form, blocked (role-split kernel).  With Acl_k = A_k + B_k K_k both sweeps are affine
//      recurrences, x_{k+1} = Acl_k x_k + b_k (forward) and p_k = Acl_k' p_{k+1} + c_k (backward).  A dependent step costs
//      ~170 cycles (exchange of the five components + a 5-term FMA tree), so the chain is shortened 4x: lane-parallel over
//      the stages, M1_k = Acl_{k+1} Acl_k and M2_k = M1_{k+2} M1_k are formed once per factorisation (they serve all three
//      sweeps; the backward sweep uses their transposes), every sweep condenses its vectors the same way (two lane-parallel
//      levels), runs the serial recursion over every 4th stage only, and fills the stages in between lane-parallel again.
//      Horizon padded to a multiple of 4 with identity stages.  Workspace: one block of SWS doubles per stage.
constexpr int NPAD = ((NSTAGE + 3) / 4) * 4;
constexpr int SW_A = 0, SW_M1 = NX * NX, SW_M2 = 2 * NX * NX, SW_B = 3 * NX * NX, SW_B1 = SW_B + NX, SW_B2 = SW_B1 + NX, SW_X = SW_B2 + NX;
constexpr int SWS = (SW_X + NX) | 1;
constexpr int SW_DOUBLES = (NPAD + 1) * SWS;

// out = w + M v (TR: M' v); M row-major in shared memory
template <bool TR>
__device__ __forceinline__ void affine5(const double* __restrict__ M, const double* v, const double* w, double* out)
{
    double m[NX * NX];
#pragma unroll
    for (int i = 0; i < NX * NX; i++) m[i] = M[i];
#pragma unroll
    for (int i = 0; i < NX; i++) {
        double acc = w[i];
#pragma unroll
        for (int j = 0; j < NX; j++) acc += (TR ? m[j * NX + i] : m[i * NX + j]) * v[j];
        out[i] = acc;
    }
}
// C = A B (row-major, registers in, shared memory out)
__device__ __forceinline__ void matmul5(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C)
{
    double a[NX * NX], b[NX * NX];
#pragma unroll
    for (int i = 0; i < NX * NX; i++) { a[i] = A[i]; b[i] = B[i]; }
#pragma unroll
    for (int i = 0; i < NX; i++)
#pragma unroll
        for (int j = 0; j < NX; j++) {
            double acc = 0.0;
#pragma unroll
            for (int l = 0; l < NX; l++) acc += a[i * NX + l] * b[l * NX + j];
            C[i * NX + j] = acc;
        }
}
// once per factorisation: SW_A of every stage (identity on the padding) -> SW_M1, SW_M2.  One full warp, lane = stage.
__device__ __noinline__ void sweep_products(double* __restrict__ sw)
{
    // (one warp covers the padded horizon: callers require NPAD <= 32, SPLIT_OK)
    const int k = threadIdx.x & 31;
    const int k1 = k + 1 < NPAD ? k + 1 : NPAD - 1, k2 = k + 2 < NPAD ? k + 2 : NPAD - 1;      // (clamped lanes produce unused values)
    double* __restrict__ me = sw + k * SWS;
    if (k < NPAD) matmul5(sw + k1 * SWS + SW_A, me + SW_A, me + SW_M1);
    __syncwarp();
    if (k < NPAD) matmul5(sw + k2 * SWS + SW_M1, me + SW_M1, me + SW_M2);
    __syncwarp();
}
// serial part over every 4th stage: lanes 0..4 own one component, exchange by shuffles
template <bool FWD>
__device__ __forceinline__ void sweep_serial4(double* __restrict__ sw)
{
    const int lane = threadIdx.x & 31;
    const int i = lane < NX ? lane : NX - 1;
    double x = FWD ? 0.0 : sw[NPAD * SWS + SW_X + i];
    if (FWD && lane < NX) sw[SW_X + i] = 0.0;
#pragma unroll 1
    for (int q = 0; q < NPAD / 4; q++) {
        const int k = FWD ? 4 * q : NPAD - 4 - 4 * q;
        const double* __restrict__ blk = sw + k * SWS;
        double a[NX], xs[NX];
#pragma unroll
        for (int j = 0; j < NX; j++) a[j] = FWD ? blk[SW_M2 + i * NX + j] : blk[SW_M2 + j * NX + i];
        const double c = blk[SW_B2 + i];
#pragma unroll
        for (int j = 0; j < NX; j++) xs[j] = __shfl_sync(FULL, x, j);
        if constexpr (NX == 5) {                    // balanced FMA tree: the serial step is latency bound
            const double t0 = a[0] * xs[0] + c, t1 = a[1] * xs[1], t2 = a[4] * xs[4];
            x = ((a[2] * xs[2] + t0) + (a[3] * xs[3] + t1)) + t2;
        } else {
            double t0 = c, t1 = 0.0;
#pragma unroll
            for (int j = 0; j < NX; j += 2) t0 += a[j] * xs[j];
#pragma unroll
            for (int j = 1; j < NX; j += 2) t1 += a[j] * xs[j];
            x = t0 + t1;
        }
        if (lane < NX) sw[(FWD ? k + 4 : k) * SWS + SW_X + i] = x;
    }
    __syncwarp();
}
// forward: in SW_B (b_k; 0 on the padding), out SW_X of stage k = x_k (x_0 = 0)
__device__ __noinline__ void sweep_forward_blocked(double* __restrict__ sw)
{
    const int k = threadIdx.x & 31;
    const int k1 = k + 1 < NPAD ? k + 1 : NPAD - 1, k2 = k + 2 < NPAD ? k + 2 : NPAD - 1;
    double* __restrict__ me = sw + k * SWS;
    double v[NX], w[NX], o[NX];
    if (k < NPAD) {                                  // b1_k = A_{k+1} b_k + b_{k+1}
#pragma unroll
        for (int i = 0; i < NX; i++) { v[i] = me[SW_B + i]; w[i] = sw[k1 * SWS + SW_B + i]; }
        affine5<false>(sw + k1 * SWS + SW_A, v, w, o);
#pragma unroll
        for (int i = 0; i < NX; i++) me[SW_B1 + i] = o[i];
    }
    __syncwarp();
    if (k < NPAD) {                                  // b2_k = M1_{k+2} b1_k + b1_{k+2}
#pragma unroll
        for (int i = 0; i < NX; i++) { v[i] = o[i]; w[i] = sw[k2 * SWS + SW_B1 + i]; }
        affine5<false>(sw + k2 * SWS + SW_M1, v, w, o);
#pragma unroll
        for (int i = 0; i < NX; i++) me[SW_B2 + i] = o[i];
    }
    __syncwarp();
    sweep_serial4<true>(sw);                         // x_0, x_4, x_8, ...
    if (k < NPAD && (k & 3) == 0) {                  // x_{k+2} = M1_k x_k + b1_k
#pragma unroll
        for (int i = 0; i < NX; i++) { v[i] = me[SW_X + i]; w[i] = me[SW_B1 + i]; }
        affine5<false>(me + SW_M1, v, w, o);
#pragma unroll
        for (int i = 0; i < NX; i++) sw[(k + 2) * SWS + SW_X + i] = o[i];
    }
    __syncwarp();
    if (k < NPAD && (k & 1) == 0) {                  // x_{k+1} = A_k x_k + b_k
#pragma unroll
        for (int i = 0; i < NX; i++) { v[i] = me[SW_X + i]; w[i] = me[SW_B + i]; }
        affine5<false>(me + SW_A, v, w, o);
#pragma unroll
        for (int i = 0; i < NX; i++) sw[(k + 1) * SWS + SW_X + i] = o[i];
    }
    __syncwarp();
}
// backward: in SW_B (c_k; 0 on the padding) and p_NPAD in SW_X of block NPAD; out SW_X of stage k = p_k
__device__ __noinline__ void sweep_backward_blocked(double* __restrict__ sw)
{
    const int k = threadIdx.x & 31;
    const int k1 = k + 1 < NPAD ? k + 1 : NPAD - 1, k2 = k + 2 < NPAD ? k + 2 : NPAD - 1;
    double* __restrict__ me = sw + k * SWS;
    double v[NX], w[NX], o[NX];
    if (k < NPAD) {                                  // c1_k = A_k' c_{k+1} + c_k
#pragma unroll
        for (int i = 0; i < NX; i++) { v[i] = sw[k1 * SWS + SW_B + i]; w[i] = me[SW_B + i]; }
        if (k + 1 >= NPAD) {
#pragma unroll
            for (int i = 0; i < NX; i++) v[i] = 0.0;
        }
        affine5<true>(me + SW_A, v, w, o);
#pragma unroll
        for (int i = 0; i < NX; i++) me[SW_B1 + i] = o[i];
    }
    __syncwarp();
    if (k < NPAD) {                                  // c2_k = M1_k' c1_{k+2} + c1_k
#pragma unroll
        for (int i = 0; i < NX; i++) { w[i] = o[i]; v[i] = (k + 2 < NPAD) ? sw[k2 * SWS + SW_B1 + i] : 0.0; }
        affine5<true>(me + SW_M1, v, w, o);
#pragma unroll
        for (int i = 0; i < NX; i++) me[SW_B2 + i] = o[i];
    }
    __syncwarp();
    sweep_serial4<false>(sw);                        // p_{NPAD-4}, ..., p_4, p_0
    if (k < NPAD && (k & 3) == 2) {                  // p_k = M1_k' p_{k+2} + c1_k
#pragma unroll
        for (int i = 0; i < NX; i++) { v[i] = sw[(k + 2) * SWS + SW_X + i]; w[i] = me[SW_B1 + i]; }
        affine5<true>(me + SW_M1, v, w, o);
#pragma unroll
        for (int i = 0; i < NX; i++) me[SW_X + i] = o[i];
    }
    __syncwarp();
    if (k < NPAD && (k & 1) == 1) {                  // p_k = A_k' p_{k+1} + c_k
#pragma unroll
        for (int i = 0; i < NX; i++) { v[i] = sw[(k + 1) * SWS + SW_X + i]; w[i] = me[SW_B + i]; }
        affine5<true>(me + SW_A, v, w, o);
#pragma unroll
        for (int i = 0; i < NX; i++) me[SW_X + i] = o[i];
    }
    __syncwarp();
}

// ---- K4: MIRROR regularisation of one packed symmetric NZ x NZ block (cyclic Jacobi) -----------
// Register-resident AND compact: the pair order is the round-robin tournament on JP = NZ + (NZ odd) positions
// (for 7 variables: 8 positions, position 7 is a decoupled dummy), i.e. every round rotates the FIXED position pairs
// (0,JP-1)(1,JP-2)... and then shifts positions 1..JP-1 cyclically, so that one rolled loop body (JP/2 rotations + a register
// permutation) serves all JP-1 rounds of a sweep -- ~10 KB of code instead of >100 KB fully unrolled.
// Rotation without a division: ir = rsqrt(tau^2 + 4 a_pq^2), cos^2 = (1 + |tau| ir)/2, ic = rsqrt(cos^2),
// c = cos^2 ic, s = sign(tau) a_pq ir ic, t = s ic   (tau = a_qq - a_pp).
constexpr int JP = NZ + (NZ & 1);                // positions: an odd variable count gets a decoupled dummy position
constexpr int JPK = JP * (JP + 1) / 2;
__host__ __device__ constexpr int jsigma(int j) { return j == 0 ? 0 : (j == 1 ? JP - 1 : j - 1); }   // new position j <- old position
__device__ __noinline__ void mirror_generic(double* Hp)
{
    double a[JPK], V[NZ][JP];
#pragma unroll
    for (int i = 0; i < JP; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) a[pk(i, j)] = (i < NZ) ? Hp[pk(i, j)] : 0.0;
#pragma unroll
    for (int i = 0; i < NZ; i++)
#pragma unroll
        for (int j = 0; j < JP; j++) V[i][j] = (i == j) ? 1.0 : 0.0;
#pragma unroll 1
    for (int sweep = 0; sweep < JACOBI_MAX_SWEEPS; sweep++) {
        double off = 0.0, dia = 0.0;
#pragma unroll
        for (int i = 0; i < NZ; i++)
#pragma unroll
            for (int j = 0; j <= i; j++) {
                if (i == j) dia += a[pk(i, j)] * a[pk(i, j)];
                else off += a[pk(i, j)] * a[pk(i, j)];
            }
        off *= 2.0;
        if (!(off > JACOBI_TOL * (off + dia))) break;
#pragma unroll 1
        for (int round = 0; round < JP - 1; round++) {
#pragma unroll
            for (int pr = 0; pr < JP / 2; pr++) {
                const int p = pr, q = JP - 1 - pr;
                const double apq = a[pk(q, p)], q2 = apq * apq;
                if (q2 > 0.0) {
                    const double tau = a[pk(q, q)] - a[pk(p, p)];
                    const double ir = rsqrt_nb(tau * tau + 4.0 * q2);
                    const double c2 = 0.5 + 0.5 * fabs(tau) * ir;
                    const double ic = rsqrt_nb(c2), c = c2 * ic;
                    const double sn = (tau >= 0.0 ? apq : -apq) * ir * ic, tt = sn * ic;
#pragma unroll
                    for (int k = 0; k < JP; k++) {
                        if (k != p && k != q) {
                            const double akp = a[pk(k, p)], akq = a[pk(k, q)];
                            a[pk(k, p)] = c * akp - sn * akq;
                            a[pk(k, q)] = sn * akp + c * akq;
                        }
                    }
                    a[pk(p, p)] -= tt * apq;
                    a[pk(q, q)] += tt * apq;
                    a[pk(q, p)] = 0.0;
#pragma unroll
                    for (int k = 0; k < NZ; k++) {
                        const double vkp = V[k][p], vkq = V[k][q];
                        V[k][p] = c * vkp - sn * vkq;
                        V[k][q] = sn * vkp + c * vkq;
                    }
                }
            }
            // shift positions 1..7 by one (position j now holds what position jsigma(j) held)
            double an[JPK], Vn[NZ][JP];
#pragma unroll
            for (int i = 0; i < JP; i++)
#pragma unroll
                for (int j = 0; j <= i; j++) an[pk(i, j)] = a[pk(jsigma(i), jsigma(j))];
#pragma unroll
            for (int i = 0; i < NZ; i++)
#pragma unroll
                for (int j = 0; j < JP; j++) Vn[i][j] = V[i][jsigma(j)];
#pragma unroll
            for (int i = 0; i < JPK; i++) a[i] = an[i];
#pragma unroll
            for (int i = 0; i < NZ; i++)
#pragma unroll
                for (int j = 0; j < JP; j++) V[i][j] = Vn[i][j];
        }
    }
    double ev[NZ];
#pragma unroll
    for (int i = 0; i < NZ; i++) {
        double e = a[pk(i, i)];
        if (e >= -REG_EPS && e <= REG_EPS) e = REG_EPS;
        else if (e < 0.0) e = -e;
        ev[i] = e;
    }
#pragma unroll
    for (int i = 0; i < NZ; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < NZ; k++) s += V[i][k] * ev[k] * V[j][k];
            Hp[pk(i, j)] = s;
        }
}

// Block path: when the entries that couple the emitter's Hessian blocks (HBLK_*) are exactly zero in every lane
// of the warp (e.g. zero disc offset: the constraints do not depend on psi), MIRROR acts on each diagonal block
// separately -- a 4x4 and a 3x3 Jacobi instead of a 7x7 one.  Row-cyclic order, same rotation and stopping rule.
template <int B>
__device__ __forceinline__ void mirror_block(double* Hp)
{
    constexpr int n = HBLK_SIZE[B];
    double a[n][n], V[n][n];
#pragma unroll
    for (int i = 0; i < n; i++)
#pragma unroll
        for (int j = 0; j < n; j++) {
            a[i][j] = Hp[pk(HBLK_IDX[B][i], HBLK_IDX[B][j])];
            V[i][j] = (i == j) ? 1.0 : 0.0;
        }
#pragma unroll 1
    for (int sweep = 0; sweep < JACOBI_MAX_SWEEPS; sweep++) {
        double off = 0.0, dia = 0.0;
#pragma unroll
        for (int i = 0; i < n; i++)
#pragma unroll
            for (int j = 0; j <= i; j++) {
                if (i == j) dia += a[i][j] * a[i][j];
                else off += a[i][j] * a[i][j];
            }
        off *= 2.0;
        if (!(off > JACOBI_TOL * (off + dia))) break;
#pragma unroll
        for (int p = 0; p < n - 1; p++)
#pragma unroll
            for (int q = p + 1; q < n; q++) {
                const double apq = a[q][p], q2 = apq * apq;
                if (q2 > 0.0) {
                    const double tau = a[q][q] - a[p][p];
                    const double ir = rsqrt_nb(tau * tau + 4.0 * q2);
                    const double c2 = 0.5 + 0.5 * fabs(tau) * ir;
                    const double ic = rsqrt_nb(c2), c = c2 * ic;
                    const double sn = (tau >= 0.0 ? apq : -apq) * ir * ic, tt = sn * ic;
#pragma unroll
                    for (int k = 0; k < n; k++) {
                        if (k != p && k != q) {
                            const double akp = a[k][p], akq = a[k][q];
                            const double np_ = c * akp - sn * akq, nq_ = sn * akp + c * akq;
                            a[k][p] = np_; a[p][k] = np_;
                            a[k][q] = nq_; a[q][k] = nq_;
                        }
                    }
                    a[p][p] -= tt * apq;
                    a[q][q] += tt * apq;
                    a[q][p] = 0.0; a[p][q] = 0.0;
#pragma unroll
                    for (int k = 0; k < n; k++) {
                        const double vkp = V[k][p], vkq = V[k][q];
                        V[k][p] = c * vkp - sn * vkq;
                        V[k][q] = sn * vkp + c * vkq;
                    }
                }
            }
    }
#pragma unroll
    for (int i = 0; i < n; i++) {
        double e = a[i][i];
        if (e >= -REG_EPS && e <= REG_EPS) e = REG_EPS;
        else if (e < 0.0) e = -e;
        a[i][i] = e;
    }
#pragma unroll
    for (int i = 0; i < n; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < n; k++) s += V[i][k] * a[k][k] * V[j][k];
            Hp[pk(HBLK_IDX[B][i], HBLK_IDX[B][j])] = s;
        }
}
template <int B>
__device__ __forceinline__ void mirror_blocks_from(double* Hp)
{
    if constexpr (B < HBLK_N) {
        mirror_block<B>(Hp);
        mirror_blocks_from<B + 1>(Hp);
    }
}
// Joint block path for the 4 + 3 structure every shipped configuration has: the two blocks are rotated in the SAME sweep
// loop, in round-robin order -- per round two disjoint rotations of the 4x4 block and one of the 3x3 block, whose angles are
// mutually independent (a rotation (p,q) leaves the (r,s) sub-block of a disjoint pair untouched).  A Jacobi rotation is a
// dependent chain of two reciprocal square roots (~150 cycles); three of them now overlap in one instruction stream, which
// cuts the chain of a sweep from 9 to 3 rotations.  Rotations are branch-free (identity when the pivot is exactly zero).
#ifndef MPC_MIRROR_JOINT
#define MPC_MIRROR_JOINT 1
#endif
struct JRot {
    double c, s, t;
    bool act;
};
__device__ __forceinline__ JRot jacobi_rot(double app, double aqq, double apq)
{
    const double q2 = apq * apq;
    JRot r;
    r.act = q2 > 0.0;
    const double tau = aqq - app;
    const double ir = rsqrt_nb(r.act ? tau * tau + 4.0 * q2 : 1.0);
    const double c2 = 0.5 + 0.5 * fabs(tau) * ir;
    const double ic = rsqrt_nb(c2), c = c2 * ic;
    const double sn = (tau >= 0.0 ? apq : -apq) * ir * ic;
    r.c = r.act ? c : 1.0; r.s = r.act ? sn : 0.0; r.t = r.act ? sn * ic : 0.0;
    return r;
}
template <int n>
__device__ __forceinline__ void jacobi_apply(double (&a)[n][n], double (&V)[n][n], int p, int q, const JRot& r)
{
    const double apq = a[q][p];
#pragma unroll
    for (int k = 0; k < n; k++) {
        if (k != p && k != q) {
            const double akp = a[k][p], akq = a[k][q];
            const double np_ = r.c * akp - r.s * akq, nq_ = r.s * akp + r.c * akq;
            a[k][p] = np_; a[p][k] = np_;
            a[k][q] = nq_; a[q][k] = nq_;
        }
    }
    if (r.act) {                                   // (predicated moves; a NaN pivot stays in place and surfaces later)
        a[p][p] -= r.t * apq;
        a[q][q] += r.t * apq;
        a[q][p] = 0.0; a[p][q] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < n; k++) {
        const double vkp = V[k][p], vkq = V[k][q];
        V[k][p] = r.c * vkp - r.s * vkq;
        V[k][q] = r.s * vkp + r.c * vkq;
    }
}
template <int n>
__device__ __forceinline__ bool jacobi_unconverged(const double (&a)[n][n])
{
    double off = 0.0, dia = 0.0;
#pragma unroll
    for (int i = 0; i < n; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            if (i == j) dia += a[i][j] * a[i][j];
            else off += a[i][j] * a[i][j];
        }
    off *= 2.0;
    return off > JACOBI_TOL * (off + dia);
}
template <int n>
__device__ __forceinline__ void mirror_writeback(double (&a)[n][n], const double (&V)[n][n], const int* idx, double* Hp)
{
#pragma unroll
    for (int i = 0; i < n; i++) {
        double e = a[i][i];
        if (e >= -REG_EPS && e <= REG_EPS) e = REG_EPS;
        else if (e < 0.0) e = -e;
        a[i][i] = e;
    }
#pragma unroll
    for (int i = 0; i < n; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < n; k++) s += V[i][k] * a[k][k] * V[j][k];
            Hp[pk(idx[i], idx[j])] = s;
        }
}
__device__ __forceinline__ void mirror_blocks_43(double* Hp)
{
    double a[4][4], V[4][4], b[3][3], W[3][3];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { a[i][j] = Hp[pk(HBLK_IDX[0][i], HBLK_IDX[0][j])]; V[i][j] = (i == j) ? 1.0 : 0.0; }
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) { b[i][j] = Hp[pk(HBLK_IDX[1][i], HBLK_IDX[1][j])]; W[i][j] = (i == j) ? 1.0 : 0.0; }
#pragma unroll 1
    for (int sweep = 0; sweep < JACOBI_MAX_SWEEPS; sweep++) {
        if (!(jacobi_unconverged(a) || jacobi_unconverged(b))) break;
        {   // round 0: (0,1) (2,3) | (0,1)
            const JRot r0 = jacobi_rot(a[0][0], a[1][1], a[1][0]), r1 = jacobi_rot(a[2][2], a[3][3], a[3][2]), r2 = jacobi_rot(b[0][0], b[1][1], b[1][0]);
            jacobi_apply(a, V, 0, 1, r0); jacobi_apply(a, V, 2, 3, r1); jacobi_apply(b, W, 0, 1, r2);
        }
        {   // round 1: (0,2) (1,3) | (0,2)
            const JRot r0 = jacobi_rot(a[0][0], a[2][2], a[2][0]), r1 = jacobi_rot(a[1][1], a[3][3], a[3][1]), r2 = jacobi_rot(b[0][0], b[2][2], b[2][0]);
            jacobi_apply(a, V, 0, 2, r0); jacobi_apply(a, V, 1, 3, r1); jacobi_apply(b, W, 0, 2, r2);
        }
        {   // round 2: (0,3) (1,2) | (1,2)
            const JRot r0 = jacobi_rot(a[0][0], a[3][3], a[3][0]), r1 = jacobi_rot(a[1][1], a[2][2], a[2][1]), r2 = jacobi_rot(b[1][1], b[2][2], b[2][1]);
            jacobi_apply(a, V, 0, 3, r0); jacobi_apply(a, V, 1, 2, r1); jacobi_apply(b, W, 1, 2, r2);
        }
    }
    mirror_writeback(a, V, HBLK_IDX[0], Hp);
    mirror_writeback(b, W, HBLK_IDX[1], Hp);
}
constexpr bool HBLK_IS_43 = (HBLK_N == 2) && (HBLK_SIZE[0] == 4) && (HBLK_SIZE[HBLK_N - 1] == 3) && (MPC_MIRROR_JOINT != 0);

__host__ __device__ constexpr int hblk_of(int v)
{
    for (int b = 0; b < HBLK_N; b++)
        for (int i = 0; i < HBLK_SIZE[b]; i++)
            if (HBLK_IDX[b][i] == v) return b;
    return -1;
}
__device__ __noinline__ void mirror_packed(double* Hp)
{
    if constexpr (HBLK_N > 1) {
        bool coupled = false;
#pragma unroll
        for (int i = 0; i < NZ; i++)
#pragma unroll
            for (int j = 0; j < i; j++)
                if (hblk_of(i) != hblk_of(j)) coupled = coupled || (Hp[pk(i, j)] != 0.0);
        if (!__any_sync(__activemask(), coupled)) {
            if constexpr (HBLK_IS_43) mirror_blocks_43(Hp);
            else mirror_blocks_from<0>(Hp);
            return;
        }
    }
    mirror_generic(Hp);
}

// Software prefetch of the thread-local Jacobian rows (MPC_C_PREFETCH = entries ahead, 0 = off): the rows of C are read once
// per pass in entry order and do not survive in the L1 between passes (8 warps x 18 KB); a row is 3 doubles per lane,
// i.e. 6 lines of the lane-interleaved local window, fetched PF entries ahead so that the loads of the entry loop hit.
#ifndef MPC_C_PREFETCH
#define MPC_C_PREFETCH 0
#endif
#ifndef MPC_PARAM_PREFETCH
#define MPC_PARAM_PREFETCH 1
#endif
template <class CM>
__device__ __forceinline__ void c_prefetch(const CM& C, int e)
{
#if MPC_C_PREFETCH > 0
    if constexpr (!LT_C) {
        if (e + MPC_C_PREFETCH < NCG) {
            const double* q = &C.p[HROW[e + MPC_C_PREFETCH] * NHS];
            const size_t a = __cvta_generic_to_local(q);
#pragma unroll
            for (int w = 0; w < 2 * NHS; w++) asm volatile("prefetch.local.L1 [%0];" ::"l"(a + 4 * w));
        }
    }
#endif
}

// chat_e' y for general entry e (y indexed by z component)
template <class CM>
__device__ __forceinline__ double gen_dot(const CM& C, int e, const double* y)
{
    const int r = HROW[e];
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < NHS; a++) s += C.at(r, a) * y[HSUP[a]];
    return HSGN[e] * s;
}

// One inequality entry of the Newton step.  lam, t, invt = 1/t, rd = chat'v - d - t, chat'dva / chat'dv =
// entry row times the affine / final direction, sigmu = sigma*mu.
//   dt = chat'dv + rd ;  dlam = -(lam + Gamma dt [+ (dt_aff dlam_aff - sigma mu)/t]),  Gamma = lam/t
struct IneqStep {
    double dt, dlam, corr;
};
__device__ __forceinline__ IneqStep ineq_affine(double lam, double invt, double rd, double cdva)
{
    IneqStep s;
    s.dt = cdva + rd;
    s.dlam = -(lam + lam * invt * s.dt);
    s.corr = s.dt * s.dlam * invt;
    return s;
}
// cen: pure centering direction (conditional predictor-corrector fallback, MPC_HPIPM_BALANCE): no second-order term
__device__ __forceinline__ IneqStep ineq_final(double lam, double invt, double rd, double cdva, double cdv, double sigmu, bool cen = false)
{
    const IneqStep a = ineq_affine(lam, invt, rd, cdva);
    IneqStep s;
    s.dt = cdv + rd;
    s.dlam = -(lam + lam * invt * s.dt + (((BALANCE && cen) ? 0.0 : a.corr) - sigmu * invt));
    s.corr = 0.0;
    return s;
}
// Step-length ratio test without a division per entry: the running minimum of val/(-dval) over the
// entries with dval < 0 is kept as a fraction bn/bd (bd > 0) and compared by cross-multiplication.
__device__ __forceinline__ void step_limit(double val, double dval, double& bn, double& bd)
{
    if (dval < 0.0 && val * bd < bn * (-dval)) { bn = val; bd = -dval; }
}
// The running minimum is a serial dependency across the entries (two multiplies + compare + selects per call): callers keep
// TWO fractions -- one fed by the multipliers, one by the slacks -- and merge them at the end, which halves the chain.
struct StepFrac {
    double n0 = 1.0, d0 = 1.0, n1 = 1.0, d1 = 1.0;
    __device__ __forceinline__ void add(double lam, double dlam, double t, double dt)
    {
        step_limit(lam, dlam, n0, d0);
        step_limit(t, dt, n1, d1);
    }
    __device__ __forceinline__ double ratio() const
    {
        const bool second = n1 * d0 < n0 * d1;
        return (second ? n1 : n0) / (second ? d1 : d0);
    }
};

// ------------------------------------------------------------------------------------------------
// One Solver::solve() on one warp
// ------------------------------------------------------------------------------------------------
// SCAN: substitution sweeps as a parallel prefix scan over the stages (lowest latency: the latency-mode kernel)
// or lane-serial (fewest instructions: the throughput kernel).
template <bool SCAN>
__device__ void solve_problem(int prob, const double* __restrict__ xinit_g, const double* __restrict__ x0_g,
                              const double* __restrict__ params_g, int num_iter, double* mem_g, int mem_doubles,
                              double* xtraj_g, double* utraj_g, double* pobj_g, int* exit_g, int* qps_g,
                              double* reseq_g, int* ipm_g, double* hb, double* lt_sm, double* rs, const Grp grp)
{
    // num_iter < 0: |num_iter| iterations with the completion step DEFERRED -- the stepwise interface (solveOneIteration,
    // acados_solver_interface.cpp:145-160) keeps multipliers and QP memory between iterations; the res_eq demotion and the
    // reset on failure belong to completeOneIteration (:176-191) and are applied by the caller
    constexpr bool TM = TMEM_CD && !SCAN;      // Jacobian rows + right-hand sides of the general entries in tensor memory
    const bool defer = num_iter < 0;
    if (defer) num_iter = -num_iter;
    const int k = grp.wig * 32 + (threadIdx.x & 31);   // stage owned by this thread
    const bool path = k < NSTAGE;             // has inputs, cost, constraints, dynamics
    const bool term = k == NSTAGE;
    const bool live = k <= NSTAGE;
    const bool xbox = path && k >= 1;         // x_0 is fixed: no state bounds at stage 0
    const double* __restrict__ p = params_g + ((size_t)prob * NSTAGE + (path ? k : NSTAGE - 1)) * NP;

    const int slot = MPC_COL_COMPACT ? (k < NSTAGE ? k : NSTAGE) : k;      // stage slot of the shared-memory columns (dead lanes: the terminal stage's dummy)
    double z[NZ], pi[NX], v[NZ], qpi[NX];
#if MPC_BOX_SMEM >= 2
    // multipliers and slacks of the box entries in shared memory too: 56 registers less in the interior-point loop
    // (TMEM_BOX: the columns do not exist -- tensor memory in the throughput instantiation, thread-local arrays in the other)
    const SmemCol lamb{lt_sm + LT_ENTRY_DOUBLES * LCOL + slot MPCK(NCB)};
    const SmemCol tb{lt_sm + (LT_ENTRY_DOUBLES + NCB) * LCOL + slot MPCK(NCB)};
#else
    double lamb[NCB], tb[NCB];
#endif
    constexpr bool TMB = TM && TMEM_BOX;
    double lamb_l[TMEM_BOX ? NCB : 1], tb_l[TMEM_BOX ? NCB : 1];
    // the four values of variable i's box entries {lam_l, t_l, lam_u, t_u}: a register window over tensor memory (loaded and
    // stored by ALL lanes, warp-collective), or the entries where they live otherwise
    auto box_ld = [&](int i, double (&bx)[4]) {
        if constexpr (TMB) tmem_ld8(grp.tmem + TM_BOX_COL0 + 8 * i, bx[0], bx[1], bx[2], bx[3]);
    };
    auto box_st = [&](int i, const double (&bx)[4]) {
        if constexpr (TMB) tmem_st8(grp.tmem + TM_BOX_COL0 + 8 * i, bx[0], bx[1], bx[2], bx[3]);
    };
    auto box_st_done = [&]() { if constexpr (TMB) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); };
#define BXLL (*(TMB ? &bx[0] : (TMEM_BOX ? &lamb_l[i] : &lamb[i])))
#define BXTL (*(TMB ? &bx[1] : (TMEM_BOX ? &tb_l[i] : &tb[i])))
#define BXLU (*(TMB ? &bx[2] : (TMEM_BOX ? &lamb_l[NZ + i] : &lamb[NZ + i])))
#define BXTU (*(TMB ? &bx[3] : (TMEM_BOX ? &tb_l[NZ + i] : &tb[NZ + i])))
    // general-entry state: shared-memory columns or thread-local arrays (MPC_LT_MASK); the unused alternative is optimised away
    double* const lt_me = lt_sm + slot;
    double lamg_loc[(!LT_LAM && NCG > 0) ? NCG : 1], tg_loc[(!LT_T && NCG > 0) ? NCG : 1];
    const LtCol<LT_LAM> lamg{LT_LAM ? lt_me + LT_OFF_LAM * LCOL : lamg_loc MPCK(NCG)};
    const LtCol<LT_T> tg{LT_T ? lt_me + LT_OFF_T * LCOL : tg_loc MPCK(NCG)};
#pragma unroll
    for (int i = 0; i < NZ; i++) {
        z[i] = live ? x0_g[(size_t)prob * NZ * (NSTAGE + 1) + k * NZ + i] : 0.0;   // loadWarmstart (:274-284)
        v[i] = 0.0;
    }
    if (term) { z[0] = 0.0; z[1] = 0.0; }
    double xi[NX];
#pragma unroll
    for (int i = 0; i < NX; i++) { xi[i] = xinit_g[(size_t)prob * NX + i]; pi[i] = 0.0; qpi[i] = 0.0; }
#pragma unroll
    for (int i = 0; i < NZ; i++) { double bx[4] = {0.0, 0.0, 0.0, 0.0}; BXLL = 0.0; BXTL = 0.0; BXLU = 0.0; BXTU = 0.0; box_st(i, bx); }
    for (int e = 0; e < NCG; e++) { lamg[e] = 0.0; tg[e] = 0.0; }
    int qp_warm = 0;
    double* mem = mem_g ? mem_g + (size_t)prob * mem_doubles : nullptr;
    if (mem && mem[0] != 0.0) {               // persistent capsule memory: [flag][pi][lam][t][v]
        const double* m = mem + 1;
        if (live) for (int i = 0; i < NX; i++) pi[i] = m[k * NX + i];
        m += (NSTAGE + 1) * NX;
#pragma unroll
        for (int i = 0; i < NZ; i++) {
            double bx[4] = {0.0, 0.0, 0.0, 0.0};
            if (path) {
                BXLL = m[k * NC + i]; BXTL = m[NSTAGE * NC + k * NC + i];
                BXLU = m[k * NC + NZ + i]; BXTU = m[NSTAGE * NC + k * NC + NZ + i];
            }
            box_st(i, bx);
        }
        if (path) {
            for (int e = 0; e < NCG; e++) { lamg[e] = m[k * NC + NCB + e]; tg[e] = m[NSTAGE * NC + k * NC + NCB + e]; }
        }
        m += 2 * NSTAGE * NC;
        if (live) for (int i = 0; i < NZ; i++) v[i] = m[k * NZ + i];
        qp_warm = (mem[0] >= 2.0);
#pragma unroll
        for (int i = 0; i < NX; i++) qpi[i] = pi[i];
    }

    int status = 0, qps = 0, ipm_total = 0;
    for (int it = 0; it < num_iter; it++) {
        // ======================= K1-K4: linearise at the current iterate ============================
        double H[NPK], g[NZ], Wv[NWV], b[NX];
        double C_loc[(!LT_C && NH > 0) ? NH * NHS : 1], dg_loc[(!LT_D && NCG > 0) ? NCG : 1];
        CRows<LT_C> C{LT_C ? lt_me + LT_OFF_C * LCOL : C_loc, lt_me + LT_OFF_CS * LCOL, false};
        const LtCol<LT_D> dg{LT_D ? lt_me + LT_OFF_D * LCOL : dg_loc MPCK(NCG)};
        // row and right-hand side of general entry e: from tensor memory (all lanes, warp-collective) or from wherever C / d live
        auto entry_cd = [&](int e, double (&cr)[NHS > 0 ? NHS : 1], double& de) {
            if constexpr (TM) {
                double c0, c1, c2;
                tmem_ld8(grp.tmem + e * TM_ECOLS, c0, c1, c2, de);
                cr[0] = c0;
                if constexpr (NHS > 1) cr[1] = c1;
                if constexpr (NHS > 2) cr[2] = c2;
            } else {
                const int r = HROW[e];
#pragma unroll
                for (int a = 0; a < NHS; a++) cr[a] = C.at(r, a);
                de = dg[e];
            }
        };
        {
            double pin[NX], xnx[NX], zx_[NX];
#pragma unroll
            for (int i = 0; i < NX; i++) zx_[i] = z[NU + i];
            grp.shift_down(pi, pin);
            grp.shift_down(zx_, xnx);
#pragma unroll
            for (int i = 0; i < NPK; i++) H[i] = 0.0;
#pragma unroll
            for (int i = 0; i < NZ; i++) g[i] = 0.0;
#pragma unroll
            for (int i = 0; i < NX; i++) b[i] = 0.0;
            if (path) {
                double xn[NX];
#if MPC_PARAM_PREFETCH
                // the stage's parameter block (NP doubles, stride NP between lanes) is read piecemeal by the emitted model code:
                // ask for all of its lines now so that the constraint rows find them in the L1
#pragma unroll 1
                for (int o = 0; o < NP; o += 16) asm volatile("prefetch.global.L1 [%0];" ::"l"(p + o));
#endif
                cost_lin(z, p, g, H);
                dyn_lin(z, pin, xn, Wv, H);
#pragma unroll
                for (int i = 0; i < NX; i++) b[i] = xn[i] - xnx[i];
                if (NH > 0) {
                    double mh[NH > 0 ? NH : 1], hv[NH > 0 ? NH : 1];
                    for (int r = 0; r < NH; r++) mh[r] = 0.0;
                    for (int e = 0; e < NCG; e++) mh[HROW[e]] -= HSGN[e] * lamg[e];   // lam_u - lam_l
                    con_lin(z, p, mh, H, hv, C);      // Jacobian + multiplier-weighted Hessian in one traversal (shared subexpressions)
                    for (int e = 0; e < NCG; e++) dg[e] = HSGN[e] * (HBND[e] - hv[HROW[e]]);
                }
                mirror_packed(H);
            } else if (term) {
#pragma unroll
                for (int i = NU; i < NZ; i++) H[pk(i, i)] = REG_EPS;   // mirror(0) = eps I; no terminal cost
            }
        }
        if constexpr (TM) {      // every lane of the warp (the stores are warp-collective); lanes without a path stage park don't-cares
#pragma unroll 2
            for (int e = 0; e < NCG; e++) {
                const int r = HROW[e];
                tmem_st8(grp.tmem + e * TM_ECOLS, C_loc[r * NHS], NHS > 1 ? C_loc[r * NHS + (NHS > 1 ? 1 : 0)] : 0.0,
                         NHS > 2 ? C_loc[r * NHS + (NHS > 2 ? 2 : 0)] : 0.0, dg_loc[e]);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        const SmemCol gs{lt_me + QP_OFF_G * LCOL MPCK(NZ)}, bs{lt_me + QP_OFF_B * LCOL MPCK(NX)}, Hs{lt_me + QP_OFF_H * LCOL MPCK(NPK)};
        if constexpr (QPS_GB) {      // (live lanes only: with compact columns the dead lanes share the terminal stage's slot)
            if (live) {
#pragma unroll
                for (int i = 0; i < NZ; i++) gs[i] = g[i];
#pragma unroll
                for (int i = 0; i < NX; i++) bs[i] = b[i];
            }
        }
        if constexpr (QPS_H) {
            if (live) {
#pragma unroll
                for (int i = 0; i < NPK; i++) Hs[i] = H[i];
            }
        }
        if constexpr (CTAIL0 < NHS && NH > 0) {      // are the tail columns of the Jacobian zero in every lane?
            bool nzt = false;
            if (path)
                for (int r = 0; r < NH; r++)
#pragma unroll
                    for (int a = CTAIL0; a < NHS; a++) nzt = nzt || (C.p[LT_C ? (r * NHS + a) * LCOL : r * NHS + a] != 0.0);
            C.tail_zero = !grp.any(nzt);
        }

        double* const blk = rs + (live ? k : 0) * RSTRIDE;      // this stage's block of the cooperative Riccati workspace (MPC_COOP)
        if constexpr (COOP) {
            if (path) {
                double Wd[NX * NZ];
                w_to_dense(Wv, Wd);
#pragma unroll
                for (int l = 0; l < NX; l++)
#pragma unroll
                    for (int j = 0; j < NZ; j++) blk[RO_B + l * NB + j] = Wd[l * NZ + j];
            }
        }

        // ======================= K5: interior-point QP ============================================
        // ---- initialisation (HPIPM-style; warm start 2 keeps v, pi, lam, t and thresholds lam, t)
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
        if (qp_warm) {
#pragma unroll
            for (int i = 0; i < NZ; i++) {
                const bool act = (i < NU) ? path : xbox;
                double bx[4];
                box_ld(i, bx);
                if (act) {
                    BXLL = clamp_lo(BXLL, IPM_THR0); BXTL = clamp_lo(BXTL, IPM_THR0);
                    BXLU = clamp_lo(BXLU, IPM_THR0); BXTU = clamp_lo(BXTU, IPM_THR0);
                }
                box_st(i, bx);
            }
            if (path) for (int e = 0; e < NCG; e++) { lamg[e] = clamp_lo(lamg[e], IPM_THR0); tg[e] = clamp_lo(tg[e], IPM_THR0); }
        } else {
#pragma unroll
            for (int i = 0; i < NZ; i++) {
                const bool act = (i < NU) ? path : xbox;
                double bx[4] = {0.0, 0.0, 0.0, 0.0};
                if (act) {
                    const double dl = LBZ[i] - z[i], du = UBZ[i] - z[i];
                    double tl = v[i] - dl, tu = du - v[i];
                    if (tl < IPM_THR0) {
                        if (tu < IPM_THR0) { v[i] = 0.5 * (dl + du); tl = IPM_THR0; tu = IPM_THR0; }
                        else { tl = IPM_THR0; v[i] = dl + IPM_THR0; }
                    } else if (tu < IPM_THR0) { tu = IPM_THR0; v[i] = du - IPM_THR0; }
                    BXTL = tl; BXTU = tu;
                    BXLL = IPM_MU0 * rcp_nb(tl); BXLU = IPM_MU0 * rcp_nb(tu);
                }
                box_st(i, bx);
            }
            if (TM || path) for (int e = 0; e < NCG; e++) {
                double cr[NHS > 0 ? NHS : 1], de;
                entry_cd(e, cr, de);
                if (!TM || path) {
                    double s_ = 0.0;
#pragma unroll
                    for (int a = 0; a < NHS; a++) s_ += cr[a] * v[HSUP[a]];
                    double tt = HSGN[e] * s_ - de;
                    if (tt < IPM_THR0) tt = IPM_THR0;
                    tg[e] = tt; lamg[e] = IPM_MU0 * rcp_nb(tt);
                }
            }
        }

        box_st_done();
        double alpha = 1.0, mu = 0.0;
#ifdef MPC_TRACE
        double nrm_g, nrm_b, nrm_d, nrm_m;
#endif
        int kk = 0;
        bool isnan_ = false;
        // state of the previous iteration's step, applied at the top of the next pass (fused update):
#if MPC_BOX_SMEM == 1 || MPC_BOX_SMEM == 3
        // 1/t of the box entries (reused by passes B, C and the update) parked in shared memory: 28 registers less in the loop
        const SmemCol itb{lt_sm + (LT_ENTRY_DOUBLES + ((MPC_BOX_SMEM == 3 && !TMEM_BOX) ? 2 * NCB : 0)) * LCOL + slot MPCK(NCB)};
#else
        double itb[NCB];                             // 1/t of the box entries, reused by passes B, C and the update (general entries: recomputed)
#endif
        double dva[NZ], dv[NZ], dpi[NX], sigmu = 0.0, a_ = 0.0;
        bool cen = false;                            // the step being applied is the centering fallback (MPC_HPIPM_BALANCE)
#pragma unroll
        for (int e = 0; e < NCB; e++) itb[e] = 0.0;
#pragma unroll
        for (int i = 0; i < NZ; i++) { dva[i] = 0.0; dv[i] = 0.0; }
#pragma unroll
        for (int i = 0; i < NX; i++) dpi[i] = 0.0;
        for (;; kk++) {
            // ---- pass DA (one traversal of the inequality entries): apply the step of the previous iteration
            //      (v, pi, lam, t), then residuals, norms, mu, Htilde = H + sum Gamma chat chat', gtilde (affine)
            double Ht[NPK], gt[NZ], rb[NX];
            double ng = 0.0, nb = 0.0, nd = 0.0, nm = 0.0, sm = 0.0;
            const bool upd = kk > 0;
            double vo[NZ];                           // v before the update (the step's residuals refer to it)
#pragma unroll
            for (int i = 0; i < NZ; i++) { vo[i] = v[i]; if (upd) v[i] += a_ * dv[i]; }
            if (upd) {
#pragma unroll
                for (int i = 0; i < NX; i++) qpi[i] += a_ * dpi[i];
            }
            {
                double qpn[NX], vxn[NX], rg[NZ], vx_[NX];
#pragma unroll
                for (int i = 0; i < NX; i++) vx_[i] = v[NU + i];
                grp.shift_down(qpi, qpn);
                grp.shift_down(vx_, vxn);
#pragma unroll
                for (int i = 0; i < NPK; i++) Ht[i] = QPS_H ? Hs[i] : H[i];
#pragma unroll
                for (int i = 0; i < NZ; i++) {
                    double s = QPS_GB ? gs[i] : g[i];
#pragma unroll
                    for (int j = 0; j < NZ; j++) s += Ht[pk(i, j)] * v[j];
                    rg[i] = s;
                }
#pragma unroll
                for (int i = 0; i < NX; i++) rb[i] = 0.0;
                if (path) {
                    wt_mul_add(Wv, qpn, rg);
#pragma unroll
                    for (int i = 0; i < NX; i++) rb[i] = (QPS_GB ? bs[i] : b[i]) - vxn[i];
                    w_mul_add(Wv, v, rb);
                }
                if (k >= 1) {
#pragma unroll
                    for (int i = 0; i < NX; i++) rg[NU + i] -= qpi[i];
                }
#pragma unroll
                for (int i = 0; i < NZ; i++) gt[i] = rg[i];
#pragma unroll
                for (int i = 0; i < NZ; i++) {
                    const bool act = (i < NU) ? path : xbox;
                    double bx[4];
                    box_ld(i, bx);
                    if (act) {
                        const double dl = LBZ[i] - z[i], du = UBZ[i] - z[i];
                        {   // lower: chat = +e_i, d = dl
                            double lam = BXLL, t = BXTL;
                            if (upd) {
                                const IneqStep st = ineq_final(lam, itb[i], vo[i] - dl - t, dva[i], dv[i], sigmu, cen);
                                lam = clamp_lo(lam + a_ * st.dlam, IPM_LAM_MIN); t = clamp_lo(t + a_ * st.dt, IPM_T_MIN);
                                BXLL = lam; BXTL = t;
                            }
                            const double it_ = rcp_nb(t);
                            const double rd = v[i] - dl - t, G = lam * it_, m = lam * t;
                            itb[i] = it_;
                            Ht[pk(i, i)] += G; gt[i] += G * rd; rg[i] -= lam;
                            nd = nanmax(nd, fabs(rd)); nm = nanmax(nm, fabs(m)); sm += m;
                        }
                        {   // upper: chat = -e_i, d = -du
                            double lam = BXLU, t = BXTU;
                            if (upd) {
                                const IneqStep st = ineq_final(lam, itb[NZ + i], du - vo[i] - t, -dva[i], -dv[i], sigmu, cen);
                                lam = clamp_lo(lam + a_ * st.dlam, IPM_LAM_MIN); t = clamp_lo(t + a_ * st.dt, IPM_T_MIN);
                                BXLU = lam; BXTU = t;
                            }
                            const double it_ = rcp_nb(t);
                            const double rd = du - v[i] - t, G = lam * it_, m = lam * t;
                            itb[NZ + i] = it_;
                            Ht[pk(i, i)] += G; gt[i] -= G * rd; rg[i] += lam;
                            nd = nanmax(nd, fabs(rd)); nm = nanmax(nm, fabs(m)); sm += m;
                        }
                    }
                    if (upd) box_st(i, bx);
                }
                if (upd) box_st_done();
                if (TM || path) {      // TM: every lane runs the loop (tensor-memory loads are warp-collective), commits predicated
#pragma unroll GEN_UNROLL
                    for (int e = 0; e < NCG; e++) {
                        c_prefetch(C, e);
                        double cr[NHS > 0 ? NHS : 1], de;
                        entry_cd(e, cr, de);
                        if (!TM || path) {
                            const double sg = HSGN[e];
                            double lam = lamg[e], t = tg[e];
                            double cv = 0.0;
                            if (upd) {
                                double cvo = 0.0, cda = 0.0, cd = 0.0;
#pragma unroll
                                for (int a = 0; a < NHS; a++) {
                                    const double ca = cr[a];
                                    cvo += ca * vo[HSUP[a]]; cda += ca * dva[HSUP[a]]; cd += ca * dv[HSUP[a]];
                                }
                                const IneqStep st = ineq_final(lam, rcp_nb(t), sg * cvo - de - t, sg * cda, sg * cd, sigmu, cen);      // 1/t recomputed: cheaper than a thread-local array
                                lam = clamp_lo(lam + a_ * st.dlam, IPM_LAM_MIN); t = clamp_lo(t + a_ * st.dt, IPM_T_MIN);
                                lamg[e] = lam; tg[e] = t;
                            }
#pragma unroll
                            for (int a = 0; a < NHS; a++) cv += cr[a] * v[HSUP[a]];
                            const double it_ = rcp_nb(t);
                            const double rd = sg * cv - de - t, G = lam * it_, m = lam * t;
#pragma unroll
                            for (int a = 0; a < NHS; a++) {
                                const double ca = cr[a];
#pragma unroll
                                for (int bb = 0; bb <= a; bb++) Ht[pk(HSUP[a], HSUP[bb])] += G * ca * cr[bb];
                                gt[HSUP[a]] += sg * ca * (G * rd);
                                rg[HSUP[a]] -= sg * ca * lam;
                            }
                            nd = nanmax(nd, fabs(rd)); nm = nanmax(nm, fabs(m)); sm += m;
                        }
                    }
                }
                if (k == 0) {
#pragma unroll
                    for (int i = NU; i < NZ; i++) rg[i] = 0.0;      // x_0 is not a variable
                }
                if (live) {
#pragma unroll
                    for (int i = 0; i < NZ; i++) ng = nanmax(ng, fabs(rg[i]));
#pragma unroll
                    for (int i = 0; i < NX; i++) nb = nanmax(nb, fabs(rb[i]));
                }
            }
            // convergence / NaN decisions by warp votes (max-norm <= tol  <=>  every lane's maximum <= tol)
            const bool lane_nan = (ng != ng) || (nb != nb) || (nd != nd) || (nm != nm);
            const bool lane_ok = (ng <= IPM_TOL) && (nb <= IPM_TOL) && (nd <= IPM_TOL) && (nm <= IPM_TOL);
            const bool any_nan = grp.any(lane_nan);
            const bool all_ok = grp.all(lane_ok);
            mu = grp.sum(sm) / (double)IPM_COUNT;
#ifdef MPC_TRACE
            nrm_g = grp.max(ng); nrm_b = grp.max(nb); nrm_d = grp.max(nd); nrm_m = grp.max(nm);
#endif
#ifdef MPC_TRACE
            if (k == 0 && prob == MPC_TRACE)
                printf("gpu sqp %d ipm %d: rg %.3e rb %.3e rd %.3e rm %.3e mu %.3e alpha %.3e\n", it, kk, nrm_g, nrm_b, nrm_d, nrm_m, mu, alpha);
#endif
            isnan_ = (mu != mu) || any_nan;
            if (!(kk < IPM_ITER_MAX && alpha > IPM_ALPHA_MIN && !isnan_ && !all_ok))
                break;

            // ---- Riccati factorisation + predictor solve.  Stage k lives on lane k; the recursion
            //      walks down the lanes, (P, p) of stage k+1 arrive by shuffle.  Square-root form on the
            //      input block: G = Ht + W'P+W, [Luu 0; Lxu I] from two Cholesky pivots, P = Gxx - Lxu Lxu'
            //      (the trailing update of the Cholesky factorisation: backward stable, no explicit
            //      inverse), l = Luu^-1 q_u, p = q_x - Lxu l.
            if constexpr (COOP) {
                // park Ht, gt, rb of every stage in shared memory; all lanes then run the recursion stage by stage
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
                __syncwarp();
                riccati_factor_coop(rs);
                riccati_forward_coop(rs);
#pragma unroll
                for (int i = 0; i < NZ; i++) dva[i] = (live && (path || i >= NU)) ? blk[RO_DZ + i] : 0.0;
            }
            double P[RIC_P ? 1 : NPX], pv[NX], Lx0[NX], Lx1[NX], Prb[RIC_PRB ? 1 : NX], lv[NU], L10 = 0.0, iL0 = 0.0, iL1 = 0.0;
            const SmemCol Ps{lt_me + RIC_OFF_P * LCOL MPCK(NPX)}, Prbs{lt_me + RIC_OFF_PRB * LCOL MPCK(NX)};      // (MPC_RIC_SMEM)
            auto prb = [&](int i) -> double { if constexpr (RIC_PRB) return Prbs[i]; else return Prb[i]; };
            auto pmat = [&](int i) -> double { if constexpr (RIC_P) return Ps[i]; else return P[i]; };
            if constexpr (!COOP) {
            if constexpr (!RIC_P) {
#pragma unroll
                for (int i = 0; i < NPX; i++) P[i] = 0.0;
            }
#pragma unroll
            for (int i = 0; i < NX; i++) { pv[i] = 0.0; Lx0[i] = 0.0; Lx1[i] = 0.0; }
            if constexpr (!RIC_PRB) {
#pragma unroll
                for (int i = 0; i < NX; i++) Prb[i] = 0.0;
            }
            lv[0] = lv[1] = 0.0;
            if (term) {
#pragma unroll
                for (int i = 0; i < NX; i++) {
                    pv[i] = gt[NU + i];
#pragma unroll
                    for (int j = 0; j <= i; j++) {
                        const double pij = Ht[pk(NU + i, NU + j)];
                        hb[pk(i, j)] = pij;
                        if constexpr (RIC_P) Ps[pk(i, j)] = pij; else P[pk(i, j)] = pij;
                    }
                }
#pragma unroll
                for (int i = 0; i < NX; i++) hb[NPX + i] = pv[i];
            }
#pragma unroll 1
            for (int s = NSTAGE - 1; s >= 0; s--) {
                grp.sync();
                if (k == s) {
                    double Pn[NPX], pn[NX];
#pragma unroll
                    for (int i = 0; i < NPX; i++) Pn[i] = hb[i];
#pragma unroll
                    for (int i = 0; i < NX; i++) pn[i] = hb[NPX + i];
                    double G[NPK], q[NZ], y[NX];
#pragma unroll
                    for (int i = 0; i < NPK; i++) G[i] = Ht[i];
                    wtpw_add(Wv, Pn, G);
#pragma unroll
                    for (int i = 0; i < NX; i++) {
                        double a = 0.0;
#pragma unroll
                        for (int j = 0; j < NX; j++) a += Pn[pk(i, j)] * rb[j];
                        if constexpr (RIC_PRB) Prbs[i] = a; else Prb[i] = a;
                        y[i] = pn[i] + a;
                    }
#pragma unroll
                    for (int i = 0; i < NZ; i++) q[i] = gt[i];
                    wt_mul_add(Wv, y, q);
                    iL0 = rsqrt_nb(G[pk(0, 0)]);
                    L10 = G[pk(1, 0)] * iL0;
                    iL1 = rsqrt_nb(G[pk(1, 1)] - L10 * L10);
#pragma unroll
                    for (int i = 0; i < NX; i++) {
                        Lx0[i] = G[pk(NU + i, 0)] * iL0;
                        Lx1[i] = (G[pk(NU + i, 1)] - Lx0[i] * L10) * iL1;
                    }
#pragma unroll
                    for (int i = 0; i < NX; i++)
#pragma unroll
                        for (int j = 0; j <= i; j++) {
                            const double pij = G[pk(NU + i, NU + j)] - Lx0[i] * Lx0[j] - Lx1[i] * Lx1[j];
                            hb[pk(i, j)] = pij;
                            if constexpr (RIC_P) Ps[pk(i, j)] = pij; else P[pk(i, j)] = pij;
                        }
                    lv[0] = q[0] * iL0;
                    lv[1] = (q[1] - L10 * lv[0]) * iL1;
#pragma unroll
                    for (int i = 0; i < NX; i++) pv[i] = q[NU + i] - Lx0[i] * lv[0] - Lx1[i] * lv[1];
#pragma unroll
                    for (int i = 0; i < NX; i++) hb[NPX + i] = pv[i];
                }
            }
            // forward sweep: dva
#pragma unroll
            for (int i = 0; i < NX; i++) dpi[i] = 0.0;
            if constexpr (SCAN) {
            forward_scan(Wv, Lx0, Lx1, L10, iL0, iL1, lv, rb, path, dva, grp);
            } else {
#pragma unroll
            for (int i = 0; i < NZ; i++) dva[i] = 0.0;
#pragma unroll 1
            for (int s = 0; s < NSTAGE; s++) {
                grp.sync();
                if (k == s) {
                    if (s > 0) {
#pragma unroll
                        for (int i = 0; i < NX; i++) dva[NU + i] = hb[i];
                    }
                    double dxn[NX];
                    double r0 = lv[0], r1 = lv[1];      // du = -Luu^-T (Lxu' dx + l)
#pragma unroll
                    for (int j = 0; j < NX; j++) { r0 += Lx0[j] * dva[NU + j]; r1 += Lx1[j] * dva[NU + j]; }
                    dva[1] = -r1 * iL1;
                    dva[0] = -(r0 + L10 * dva[1]) * iL0;
#pragma unroll
                    for (int i = 0; i < NX; i++) dxn[i] = rb[i];
                    w_mul_add(Wv, dva, dxn);
#pragma unroll
                    for (int i = 0; i < NX; i++) hb[i] = dxn[i];
                }
            }
            grp.sync();
            if (term) {
#pragma unroll
                for (int i = 0; i < NX; i++) dva[NU + i] = hb[i];
            }
            }

            }   // !COOP

            // ---- pass B: affine step length, mu_aff sums, corrector vectors
            StepFrac sfa;
            double S1 = 0.0, S2 = 0.0, V1[NZ], V2[NZ];   // alpha_aff = sfa.ratio()
#pragma unroll
            for (int i = 0; i < NZ; i++) { V1[i] = 0.0; V2[i] = 0.0; }
#pragma unroll
            for (int i = 0; i < NZ; i++) {
                const bool act = (i < NU) ? path : xbox;
                double bx[4];
                box_ld(i, bx);
                if (act) {
                    const double dl = LBZ[i] - z[i], du = UBZ[i] - z[i];
                    {
                        const double lam = BXLL, t = BXTL;
                        const double it_ = itb[i];
                        const IneqStep st = ineq_affine(lam, it_, v[i] - dl - t, dva[i]);
                        sfa.add(lam, st.dlam, t, st.dt);
                        S1 += lam * st.dt + t * st.dlam; S2 += st.dt * st.dlam;
                        V1[i] += st.corr; V2[i] += it_;
                    }
                    {
                        const double lam = BXLU, t = BXTU;
                        const double it_ = itb[NZ + i];
                        const IneqStep st = ineq_affine(lam, it_, du - v[i] - t, -dva[i]);
                        sfa.add(lam, st.dlam, t, st.dt);
                        S1 += lam * st.dt + t * st.dlam; S2 += st.dt * st.dlam;
                        V1[i] -= st.corr; V2[i] -= it_;
                    }
                }
            }
            if (TM || path) {
#pragma unroll GEN_UNROLL
                for (int e = 0; e < NCG; e++) {
                    c_prefetch(C, e);
                    double cr[NHS > 0 ? NHS : 1], de;
                    entry_cd(e, cr, de);
                    if (!TM || path) {
                        const double sg = HSGN[e], lam = lamg[e], t = tg[e];
                        double cv = 0.0, cd = 0.0;
#pragma unroll
                        for (int a = 0; a < NHS; a++) { const double ca = cr[a]; cv += ca * v[HSUP[a]]; cd += ca * dva[HSUP[a]]; }
                        const double it_ = rcp_nb(t);
                        const IneqStep st = ineq_affine(lam, it_, sg * cv - de - t, sg * cd);
                        sfa.add(lam, st.dlam, t, st.dt);
                        S1 += lam * st.dt + t * st.dlam; S2 += st.dt * st.dlam;
#pragma unroll
                        for (int a = 0; a < NHS; a++) {
                            V1[HSUP[a]] += sg * cr[a] * st.corr;
                            V2[HSUP[a]] += sg * cr[a] * it_;
                        }
                    }
                }
            }
            const double alpha_aff = grp.min(sfa.ratio());
            S1 = grp.sum(S1); S2 = grp.sum(S2);
            const double mu_aff = (mu * (double)IPM_COUNT + alpha_aff * S1 + alpha_aff * alpha_aff * S2) / (double)IPM_COUNT;
            const double rat = mu_aff / mu;
            sigmu = rat * rat * rat * mu;
#pragma unroll
            for (int i = 0; i < NZ; i++) gt[i] += V1[i] - sigmu * V2[i];

            cen = false;
#pragma unroll 1
            for (int attempt = 0; attempt < (BALANCE ? 2 : 1); attempt++) {
            // ---- corrector solve (factorisation reused): backward vector sweep + forward sweep
            if constexpr (COOP) {
                if (live) {
#pragma unroll
                    for (int i = 0; i < NZ; i++) blk[RO_Q + i] = gt[i];
                }
                __syncwarp();
                riccati_backvec_coop(rs);
                riccati_forward_coop(rs);
#pragma unroll
                for (int i = 0; i < NZ; i++) dv[i] = (live && (path || i >= NU)) ? blk[RO_DZ + i] : 0.0;
                if (k >= 1 && live) {                   // dpi_k = P_k dx_k + p_k (lane-parallel)
#pragma unroll
                    for (int i = 0; i < NX; i++) {
                        double a = blk[RO_Q + NU + i];
#pragma unroll
                        for (int j = 0; j < NX; j++) a += blk[RO_G + pk(NU + i, NU + j)] * dv[NU + j];
                        dpi[i] = a;
                    }
                }
            } else {
            if constexpr (SCAN) {
            {   // backward vector sweep as a scan: p_k = Acl_k' (p_{k+1} + P_{k+1} rb_k) + (gt_x - Lxu Luu^-1 gt_u)
                Aff am;
                double Bd[NX * NU];
                if (path) {
                    double Acl[NX * NX];
                    closed_loop(Wv, Lx0, Lx1, L10, iL0, iL1, Acl, Bd);
                    const double lg0 = gt[0] * iL0, lg1 = (gt[1] - L10 * lg0) * iL1;
#pragma unroll
                    for (int i = 0; i < NX; i++) {
                        double c_ = gt[NU + i] - Lx0[i] * lg0 - Lx1[i] * lg1;
#pragma unroll
                        for (int j = 0; j < NX; j++) {
                            am.M[i * NX + j] = Acl[j * NX + i];
                            c_ += Acl[j * NX + i] * prb(j);
                        }
                        am.c[i] = c_;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < NX * NX; i++) am.M[i] = (!term && i / NX == i % NX) ? 1.0 : 0.0;   // terminal: absorbing
#pragma unroll
                    for (int i = 0; i < NX; i++) am.c[i] = term ? gt[NU + i] : 0.0;
#pragma unroll
                    for (int i = 0; i < NX * NU; i++) Bd[i] = 0.0;
                }
                aff_scan<false>(&am, grp);
                double pn[NX];
#pragma unroll
                for (int i = 0; i < NX; i++) pv[i] = am.c[i];
                grp.shift_down(pv, pn);
                if (path) {
                    double q0 = gt[0], q1 = gt[1];
#pragma unroll
                    for (int i = 0; i < NX; i++) {
                        const double y = pn[i] + prb(i);
                        q0 += Bd[i * NU] * y; q1 += Bd[i * NU + 1] * y;
                    }
                    lv[0] = q0 * iL0;
                    lv[1] = (q1 - L10 * lv[0]) * iL1;
                }
            }
            forward_scan(Wv, Lx0, Lx1, L10, iL0, iL1, lv, rb, path, dv, grp);
            } else {
            grp.sync();
            if (term) {
#pragma unroll
                for (int i = 0; i < NX; i++) { pv[i] = gt[NU + i]; hb[NPX + i] = pv[i]; }
            }
#pragma unroll 1
            for (int s = NSTAGE - 1; s >= 0; s--) {
                grp.sync();
                if (k == s) {
                    double pn[NX];
#pragma unroll
                    for (int i = 0; i < NX; i++) pn[i] = hb[NPX + i];
                    double q[NZ], y[NX];
#pragma unroll
                    for (int i = 0; i < NX; i++) y[i] = pn[i] + prb(i);
#pragma unroll
                    for (int i = 0; i < NZ; i++) q[i] = gt[i];
                    wt_mul_add(Wv, y, q);
                    lv[0] = q[0] * iL0;
                    lv[1] = (q[1] - L10 * lv[0]) * iL1;
#pragma unroll
                    for (int i = 0; i < NX; i++) { pv[i] = q[NU + i] - Lx0[i] * lv[0] - Lx1[i] * lv[1]; hb[NPX + i] = pv[i]; }
                }
            }
#pragma unroll
            for (int i = 0; i < NZ; i++) dv[i] = 0.0;
#pragma unroll 1
            for (int s = 0; s < NSTAGE; s++) {
                grp.sync();
                if (k == s) {
                    if (s > 0) {
#pragma unroll
                        for (int i = 0; i < NX; i++) dv[NU + i] = hb[i];
                    }
                    double dxn[NX];
                    double r0 = lv[0], r1 = lv[1];      // du = -Luu^-T (Lxu' dx + l)
#pragma unroll
                    for (int j = 0; j < NX; j++) { r0 += Lx0[j] * dv[NU + j]; r1 += Lx1[j] * dv[NU + j]; }
                    dv[1] = -r1 * iL1;
                    dv[0] = -(r0 + L10 * dv[1]) * iL0;
#pragma unroll
                    for (int i = 0; i < NX; i++) dxn[i] = rb[i];
                    w_mul_add(Wv, dv, dxn);
#pragma unroll
                    for (int i = 0; i < NX; i++) hb[i] = dxn[i];
                }
            }
            grp.sync();
            if (term) {
#pragma unroll
                for (int i = 0; i < NX; i++) dv[NU + i] = hb[i];
            }
            }
            if (k >= 1 && live) {                       // dpi_k = P_k dx_k + p_k (lane-parallel)
#pragma unroll
                for (int i = 0; i < NX; i++) {
                    double a = pv[i];
#pragma unroll
                    for (int j = 0; j < NX; j++) a += pmat(pk(i, j)) * dv[NU + j];
                    dpi[i] = a;
                }
            }
            }   // !COOP

            // ---- pass C: step length of the corrected direction
            StepFrac sfc;                  // alpha = sfc.ratio()
            double T1 = 0.0, T2 = 0.0;     // (MPC_HPIPM_BALANCE) duality measure of the corrected step, expanded in powers of alpha
#pragma unroll
            for (int i = 0; i < NZ; i++) {
                const bool act = (i < NU) ? path : xbox;
                double bx[4];
                box_ld(i, bx);
                if (act) {
                    const double dl = LBZ[i] - z[i], du = UBZ[i] - z[i];
                    {
                        const double lam = BXLL, t = BXTL;
                        const IneqStep st = ineq_final(lam, itb[i], v[i] - dl - t, dva[i], dv[i], sigmu, cen);
                        sfc.add(lam, st.dlam, t, st.dt);
                        if constexpr (BALANCE) { T1 += lam * st.dt + t * st.dlam; T2 += st.dt * st.dlam; }
                    }
                    {
                        const double lam = BXLU, t = BXTU;
                        const IneqStep st = ineq_final(lam, itb[NZ + i], du - v[i] - t, -dva[i], -dv[i], sigmu, cen);
                        sfc.add(lam, st.dlam, t, st.dt);
                        if constexpr (BALANCE) { T1 += lam * st.dt + t * st.dlam; T2 += st.dt * st.dlam; }
                    }
                }
            }
            if (TM || path) {
#pragma unroll GEN_UNROLL
                for (int e = 0; e < NCG; e++) {
                    c_prefetch(C, e);
                    double cr[NHS > 0 ? NHS : 1], de;
                    entry_cd(e, cr, de);
                    if (!TM || path) {
                        const double sg = HSGN[e], lam = lamg[e], t = tg[e];
                        double cv = 0.0, cda = 0.0, cd = 0.0;
#pragma unroll
                        for (int a = 0; a < NHS; a++) {
                            const double ca = cr[a];
                            cv += ca * v[HSUP[a]]; cda += ca * dva[HSUP[a]]; cd += ca * dv[HSUP[a]];
                        }
                        const IneqStep st = ineq_final(lam, rcp_nb(t), sg * cv - de - t, sg * cda, sg * cd, sigmu, cen);
                        sfc.add(lam, st.dlam, t, st.dt);
                        if constexpr (BALANCE) { T1 += lam * st.dt + t * st.dlam; T2 += st.dt * st.dlam; }
                    }
                }
            }
            alpha = grp.min(sfc.ratio());
            if constexpr (BALANCE) {
                // conditional predictor-corrector [upstream: hpipm ocp_qp_ipm.c, cond_pred_corr = 1 in mode BALANCE]: the duality
                // measure the corrected step would leave, against twice the predictor's
                if (attempt == 0) {
                    T1 = grp.sum(T1); T2 = grp.sum(T2);
                    const double mu_pc = (mu * (double)IPM_COUNT + alpha * T1 + alpha * alpha * T2) / (double)IPM_COUNT;
                    if (mu_pc > 2.0 * mu_aff) {
                        cen = true;                                  // redo the solve with the centering right-hand side only
#pragma unroll
                        for (int i = 0; i < NZ; i++) gt[i] -= V1[i];
                        continue;
                    }
                }
                break;
            }
            }   // attempt
            a_ = alpha < 1.0 ? alpha * IPM_STEP_SCALE : alpha;      // applied by the next pass DA
        }
        ipm_total += kk;
        qps = isnan_ ? 3 : ((kk == IPM_ITER_MAX) ? 1 : (alpha <= IPM_ALPHA_MIN ? 2 : 0));

        // ======================= K6: SQP-RTI full step ============================================
        if (qps != 0 && qps != 1) { status = 4; break; }   // ACADOS_QP_FAILURE: iterate unchanged
#pragma unroll
        for (int i = 0; i < NZ; i++) z[i] += v[i];
        if (term) { z[0] = 0.0; z[1] = 0.0; }
#pragma unroll
        for (int i = 0; i < NX; i++) pi[i] = qpi[i];
        qp_warm = 1;
        status = 0;
        if (qps != 0) break;                               // wrapper breaks on qp_status != 0 (:105-106)
    }

    // ======================= completeOneIteration (:162-204) ======================================
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
    // deterministic stage-order sum (matches the oracle's sequential accumulation)
    const double cost = grp.ordered_sum(cst);
    req = grp.max(req);
    if (!(req <= RES_EQ_MAX) && status == 0 && !defer) status = 4;
    const int exit_code = (status == 0) ? 1 : (status == 1 ? 0 : status);
    if (live) {
#pragma unroll
        for (int i = 0; i < NX; i++) xtraj_g[(size_t)prob * NX * (NSTAGE + 1) + k * NX + i] = z[NU + i];
    }
    if (path) {
#pragma unroll
        for (int i = 0; i < NU; i++) utraj_g[(size_t)prob * NU * NSTAGE + k * NU + i] = z[i];
    }
    if (k == 0) {
        pobj_g[prob] = cost; exit_g[prob] = exit_code; qps_g[prob] = qp_status_acados(qps); reseq_g[prob] = req;
        if (ipm_g) ipm_g[prob] = ipm_total;
    }
    if (mem) {
        if (status != 0) {                                 // Solver_acados_reset + reset_qp_memory (:187-191)
            if (!defer)
                for (int i = k; i < mem_doubles; i += 32 * GW) mem[i] = 0.0;
        } else {
            double* m = mem + 1;
            if (k == 0) mem[0] = 2.0;
            if (live) for (int i = 0; i < NX; i++) m[k * NX + i] = pi[i];
            m += (NSTAGE + 1) * NX;
#pragma unroll
            for (int i = 0; i < NZ; i++) {
                double bx[4];
                box_ld(i, bx);
                if (path) {
                    m[k * NC + i] = BXLL; m[NSTAGE * NC + k * NC + i] = BXTL;
                    m[k * NC + NZ + i] = BXLU; m[NSTAGE * NC + k * NC + NZ + i] = BXTU;
                }
            }
            if (path) {
                for (int e = 0; e < NCG; e++) { m[k * NC + NCB + e] = lamg[e]; m[NSTAGE * NC + k * NC + NCB + e] = tg[e]; }
            }
            m += 2 * NSTAGE * NC;
            if (live) for (int i = 0; i < NZ; i++) m[k * NZ + i] = v[i];
        }
    }
}
#undef BXLL
#undef BXTL
#undef BXLU
#undef BXTU

constexpr int WARPS_PER_CTA = MPC_WARPS_PER_CTA;

// Persistent grid: warp groups pull problem indices from a global counter (work per problem is data
// dependent: 50-100 interior-point iterations), so late finishers do not idle a whole CTA.
// Two instantiations: WPC = WARPS_PER_CTA (throughput: 8 warps share an SM) and WPC = GW (latency: one
// problem per CTA, so a small batch -- one homotopy set -- spreads over as many SMs as it has problems).
template <int WPC>
__global__ void __launch_bounds__(WPC * 32, (WPC == WARPS_PER_CTA) ? MPC_MIN_CTAS : 1)
mpc_solve_kernel(int n, const double* __restrict__ xinit, const double* __restrict__ x0, const double* __restrict__ params,
                 const int* __restrict__ num_iter, int num_iter_all, double* mem, int mem_doubles, double* xtraj,
                 double* utraj, double* pobj, int* exit_code, int* qp_status, double* res_eq, int* ipm_iters,
                 int* work_counter)
{
    // per-group shared memory: hand-off buffer of the Riccati sweeps ((P, p) of stage k+1 -> stage k,
    // dx_k+1 -> stage k+1), the cross-warp exchange area, and the fetched problem index
    constexpr int GROUPS = WPC / GW;
    __shared__ double s_hand[GROUPS][NPX + NX + 4];
    __shared__ double s_xch[GROUPS][GW * XCH];
    __shared__ int s_prob[GROUPS];
    extern __shared__ double s_lt[];            // [GROUPS][LT_DOUBLES] per-entry state of the general inequality entries
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Grp grp;
    grp.gid = warp / GW;
    grp.wig = warp % GW;
    grp.xch = s_xch[grp.gid];
#if MPC_CHECK
    s_lt[(size_t)grp.gid * LT_STRIDE + LT_DOUBLES + grp.wig * 32 + lane] = CANARY;
#endif
    constexpr bool TM = TMEM_CD && (WPC == WARPS_PER_CTA) && (MPC_SCAN_ALWAYS == 0);
    __shared__ unsigned s_tmem;
    if constexpr (TM) {      // the CTA owns its SM: all 512 columns; warp w: lanes 32 (w % 4) .., columns 256 (w / 4) ..
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(&s_tmem)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        grp.tmem = s_tmem + ((32u * (warp & 3)) << 16) + 256u * (warp >> 2);
    }
    for (;;) {
        int prob = 0;
        if (grp.wig == 0 && lane == 0) prob = atomicAdd(work_counter, 1);
        if constexpr (GW == 1) {
            prob = __shfl_sync(FULL, prob, 0);
        } else {
            if (grp.wig == 0 && lane == 0) s_prob[grp.gid] = prob;
            grp.sync();
            prob = s_prob[grp.gid];
            grp.sync();
        }
        if (prob >= n) break;
        // gated launch (host pipeline of mpcgpu_solve_batch): the word behind the work counter is 0 (inputs resident) or -(m + 1)
        // where the inputs of problems 0 .. m-1 have arrived; the copy stream raises it behind every chunk it has delivered
        {
            const volatile int* gate = work_counter + 1;
            for (int g = *gate; g < 0 && -(g + 1) <= prob; g = *gate) __nanosleep(500);
            __threadfence();
        }
        const int nit = num_iter ? num_iter[prob] : num_iter_all;
        solve_problem<(WPC != WARPS_PER_CTA) || (MPC_SCAN_ALWAYS != 0)>(prob, xinit, x0, params, nit, mem, mem_doubles, xtraj, utraj, pobj, exit_code, qp_status, res_eq,
                      ipm_iters, s_hand[grp.gid], s_lt + (size_t)grp.gid * LT_STRIDE,
                      s_lt + (size_t)GROUPS * LT_STRIDE + (size_t)grp.gid * (COOP ? RS_DOUBLES : 0), grp);
#if MPC_CHECK
        grp.sync();
        if (s_lt[(size_t)grp.gid * LT_STRIDE + LT_DOUBLES + grp.wig * 32 + lane] != CANARY) MPC_CHECK_FAIL(1);
        if (grp.wig == 0 && lane == 0) MPC_CHECK_FAIL(4);
#endif
    }
    if constexpr (TM) {      // every warp has left the work loop before the columns go back
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512u) : "memory");
    }
}

// Diagnostic kernel (tests): evaluates the emitted model functions at n points, one thread per point.
// out layout per point: xn[NX] | W[NX*NZ] | Hdyn[NPK] | cost | g[NZ] | Hcost[NPK] | h[NH] | C[NH*NHS] | Hcon[NPK]
constexpr int MODEL_EVAL_DOUBLES = NX + NX * NZ + NPK + 1 + NZ + NPK + NH + NH * NHS + NPK;
__global__ void model_eval_kernel(int n, const double* __restrict__ z_g, const double* __restrict__ p_g,
                                  const double* __restrict__ pi_g, const double* __restrict__ mh_g, double* __restrict__ out_g)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double z[NZ], pi[NX], mh[NH > 0 ? NH : 1], xn[NX], Wv[NWV], Wd[NX * NZ], H[NPK], g[NZ], hv[NH > 0 ? NH : 1], C[NH > 0 ? NH * NHS : 1];
    const double* p = p_g + (size_t)i * NP;
    for (int j = 0; j < NZ; j++) z[j] = z_g[(size_t)i * NZ + j];
    for (int j = 0; j < NX; j++) pi[j] = pi_g[(size_t)i * NX + j];
    for (int j = 0; j < NH; j++) mh[j] = mh_g[(size_t)i * NH + j];
    double* o = out_g + (size_t)i * MODEL_EVAL_DOUBLES;
    for (int j = 0; j < NPK; j++) H[j] = 0.0;
    dyn_lin(z, pi, xn, Wv, H);
    w_to_dense(Wv, Wd);
    for (int j = 0; j < NX; j++) *o++ = xn[j];
    for (int j = 0; j < NX * NZ; j++) *o++ = Wd[j];
    for (int j = 0; j < NPK; j++) *o++ = H[j];
    *o++ = cost_val(z, p);
    cost_lin(z, p, g, H);
    for (int j = 0; j < NZ; j++) *o++ = g[j];
    for (int j = 0; j < NPK; j++) *o++ = H[j];
    for (int j = 0; j < NPK; j++) H[j] = 0.0;
    con_lin(z, p, mh, H, hv, C);                 // the function the solve kernels use
    for (int j = 0; j < NH; j++) *o++ = hv[j];
    for (int j = 0; j < NH * NHS; j++) *o++ = C[j];
    for (int j = 0; j < NPK; j++) *o++ = H[j];
}

#include "mpc_solve_split.cuh"

}  // namespace MPC_NS
