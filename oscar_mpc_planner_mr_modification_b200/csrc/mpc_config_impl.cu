// Compiled once per problem configuration:  nvcc -DMPC_MODEL_HEADER='"<cfg>/model.cuh"' -DMPC_CFG_TAG=<cfg>
#include <cstdlib>
#include "mpc_solve_kernel.cuh"
#include "mpc_registry.h"

namespace MPC_NS {

constexpr size_t SMEM_THR = (size_t)(WARPS_PER_CTA / GW) * (LT_STRIDE + (COOP ? RS_DOUBLES : 0)) * sizeof(double);   // per-entry state of the general inequality entries
constexpr size_t SMEM_LAT = (size_t)(LT_STRIDE + (COOP ? RS_DOUBLES : 0)) * sizeof(double);
#ifndef MPC_SPLIT
#define MPC_SPLIT 1
#endif
// Role-split kernel (one CTA of several warps per problem, mpc_solve_split.cuh) for SMALL batches -- one or a few homotopy
// sets, where latency is what counts -- when the configuration has enough general constraints to share out.  Large
// batches use the thread-per-stage kernel (8 problems per SM; higher throughput).  MPC_SPLIT=2 forces it for every batch.
constexpr bool USE_SPLIT = (MPC_SPLIT != 0) && SPLIT_OK;
constexpr bool SPLIT_ALWAYS = (MPC_SPLIT == 2) && SPLIT_OK;
constexpr size_t SMEM_SPLIT = (size_t)SP_DOUBLES * sizeof(double);
constexpr int MEM_DOUBLES = 1 + (NSTAGE + 1) * NX + 2 * NSTAGE * NC + (NSTAGE + 1) * NZ;

// dynamic shared memory beyond 48 KB is an opt-in per kernel (and per device context: set on every call, it is cheap)
static cudaError_t set_smem_attributes()
{
    cudaError_t err = cudaSuccess;
    if (USE_SPLIT) {
        err = cudaFuncSetAttribute(mpc_solve_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_SPLIT);
        if (err != cudaSuccess || SPLIT_ALWAYS) return err;
    }
    err = cudaFuncSetAttribute(mpc_solve_kernel<GW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_LAT);
    if (err == cudaSuccess && GW != WARPS_PER_CTA)
        err = cudaFuncSetAttribute(mpc_solve_kernel<WARPS_PER_CTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_THR);
    if (const char* co = getenv("MPCGPU_CARVEOUT"))      // experiment: preferred shared-memory carve-out in percent
        cudaFuncSetAttribute(mpc_solve_kernel<WARPS_PER_CTA>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(co));
    return err;
}

static cudaError_t launch_solve(int grid, int mode, cudaStream_t stream, int n, const double* xinit, const double* x0, const double* params,
                         const int* num_iter, int num_iter_all, double* mem, double* xtraj, double* utraj, double* pobj,
                         int* exit_code, int* qp_status, double* res_eq, int* ipm_iters, int* work_counter)
{
    // work_counter[0]: next problem index; [1]: input gate of the thread-per-stage kernel (0 = inputs resident).  A gated launch
    // (mode bit 8, host pipeline) finds both words prepared by the caller on its copy stream.
    const bool gated = (mode & 0x100) != 0;
    mode &= 0xff;
    cudaError_t err = gated ? cudaSuccess : cudaMemsetAsync(work_counter, 0, 2 * sizeof(int), stream);
    if (err != cudaSuccess) return err;
    err = set_smem_attributes();
    if (err != cudaSuccess) return err;
    // small batches (grid < 0, at most one problem per SM) -> role-split kernel unless the caller pins a kernel
    int sms = 148, dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const bool split = USE_SPLIT && !gated && (mode == 2 || SPLIT_ALWAYS || (mode == 0 && grid < 0 && -grid <= sms));
    if (split)
        mpc_solve_split_kernel<<<grid < 0 ? -grid : (n < sms * 4 ? n : sms * 4), SPLIT_THREADS, SMEM_SPLIT, stream>>>(n, xinit, x0, params, num_iter, num_iter_all, mem,
                                                                                          MEM_DOUBLES, xtraj, utraj, pobj, exit_code,
                                                                                          qp_status, res_eq, ipm_iters, work_counter);
    else if (grid < 0)      // latency mode: one problem per CTA, -grid CTAs
        mpc_solve_kernel<GW><<<-grid, GW * 32, SMEM_LAT, stream>>>(n, xinit, x0, params, num_iter, num_iter_all, mem, MEM_DOUBLES, xtraj, utraj,
                                                             pobj, exit_code, qp_status, res_eq, ipm_iters, work_counter);
    else
        mpc_solve_kernel<WARPS_PER_CTA><<<grid, WARPS_PER_CTA * 32, SMEM_THR, stream>>>(n, xinit, x0, params, num_iter, num_iter_all, mem,
                                                                                  MEM_DOUBLES, xtraj, utraj, pobj, exit_code, qp_status,
                                                                                  res_eq, ipm_iters, work_counter);
    return cudaGetLastError();
}

static cudaError_t occupancy(int* ctas_per_sm, int* threads_per_cta)
{
    cudaError_t err = set_smem_attributes();
    if (err != cudaSuccess) return err;
    if (SPLIT_ALWAYS) {
        *threads_per_cta = SPLIT_THREADS;
        return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, mpc_solve_split_kernel, SPLIT_THREADS, SMEM_SPLIT);
    }
    *threads_per_cta = WARPS_PER_CTA * 32;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, mpc_solve_kernel<WARPS_PER_CTA>, WARPS_PER_CTA * 32, SMEM_THR);
}

static cudaError_t launch_model_eval(cudaStream_t stream, int n, const double* z, const double* p, const double* pi,
                                     const double* mh, double* out)
{
    model_eval_kernel<<<(n + 63) / 64, 64, 0, stream>>>(n, z, p, pi, mh, out);
    return cudaGetLastError();
}

#if MPC_CHECK
static cudaError_t check_report(unsigned long long* out8, int reset)
{
    cudaError_t err = cudaDeviceSynchronize();
    if (err == cudaSuccess) err = cudaMemcpyFromSymbol(out8, g_check, sizeof(g_check));
    if (err == cudaSuccess && reset) {
        const unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        err = cudaMemcpyToSymbol(g_check, z, sizeof(z));
    }
    return err;
}
// the detector detecting: one out-of-range index through a shared-memory column and one overwritten canary word
__global__ void check_selftest_kernel()
{
    extern __shared__ double sm[];
    sm[threadIdx.x] = 0.0;
    sm[64 + threadIdx.x] = CANARY;
    __syncthreads();
    const SmemCol col{sm + threadIdx.x, 1};
    if (threadIdx.x == 3) col[1] = 5.0;                    // index 1 of a 1-entry column: counted, redirected to entry 0
    if (threadIdx.x == 0) sm[64 + 7] = 1.0;                // a stray store into the canary row
    __syncthreads();
    if (sm[64 + threadIdx.x] != CANARY) MPC_CHECK_FAIL(1);
}
static cudaError_t check_selftest(cudaStream_t stream)
{
    check_selftest_kernel<<<1, 32, 2 * 64 * sizeof(double), stream>>>();
    return cudaGetLastError();
}
#endif

static const MpcConfigOps ops = {MPCGEN_CONFIG_NAME, NSTAGE, NX, NU, NP, NH, NC, MEM_DOUBLES, launch_solve, occupancy,
                                 NHS, MODEL_EVAL_DOUBLES, launch_model_eval, SPLIT_ALWAYS ? SPLIT_WARPS : GW, USE_SPLIT ? 1 : 0,
#if MPC_CHECK
                                 check_report, check_selftest};
#else
                                 nullptr, nullptr};
#endif

static struct Registrar {
    Registrar() { mpc_register_config(&ops); }
} registrar;
}  // namespace MPC_NS
