// Compiled once per problem configuration:  nvcc -DMPC_MODEL_HEADER='"<cfg>/model.cuh"' -DMPC_CFG_TAG=<cfg>
#include "mpc_solve_kernel.cuh"
#include "mpc_registry.h"

namespace MPC_NS {

constexpr int MEM_DOUBLES = 1 + (NSTAGE + 1) * NX + 2 * NSTAGE * NC + (NSTAGE + 1) * NZ;

static cudaError_t launch_solve(int grid, cudaStream_t stream, int n, const double* xinit, const double* x0, const double* params,
                         const int* num_iter, int num_iter_all, double* mem, double* xtraj, double* utraj, double* pobj,
                         int* exit_code, int* qp_status, double* res_eq, int* ipm_iters, int* work_counter)
{
    cudaError_t err = cudaMemsetAsync(work_counter, 0, sizeof(int), stream);
    if (err != cudaSuccess) return err;
    if (grid < 0)      // latency mode: one problem per CTA, -grid CTAs
        mpc_solve_kernel<GW><<<-grid, GW * 32, 0, stream>>>(n, xinit, x0, params, num_iter, num_iter_all, mem, MEM_DOUBLES, xtraj, utraj,
                                                             pobj, exit_code, qp_status, res_eq, ipm_iters, work_counter);
    else
        mpc_solve_kernel<WARPS_PER_CTA><<<grid, WARPS_PER_CTA * 32, 0, stream>>>(n, xinit, x0, params, num_iter, num_iter_all, mem,
                                                                                  MEM_DOUBLES, xtraj, utraj, pobj, exit_code, qp_status,
                                                                                  res_eq, ipm_iters, work_counter);
    return cudaGetLastError();
}

static cudaError_t occupancy(int* ctas_per_sm, int* threads_per_cta)
{
    *threads_per_cta = WARPS_PER_CTA * 32;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, mpc_solve_kernel<WARPS_PER_CTA>, WARPS_PER_CTA * 32, 0);
}

static cudaError_t launch_model_eval(cudaStream_t stream, int n, const double* z, const double* p, const double* pi,
                                     const double* mh, double* out)
{
    model_eval_kernel<<<(n + 63) / 64, 64, 0, stream>>>(n, z, p, pi, mh, out);
    return cudaGetLastError();
}

static const MpcConfigOps ops = {MPCGEN_CONFIG_NAME, NSTAGE, NX, NU, NP, NH, NC, MEM_DOUBLES, launch_solve, occupancy,
                                 NHS, MODEL_EVAL_DOUBLES, launch_model_eval, GW};

static struct Registrar {
    Registrar() { mpc_register_config(&ops); }
} registrar;
}  // namespace MPC_NS
