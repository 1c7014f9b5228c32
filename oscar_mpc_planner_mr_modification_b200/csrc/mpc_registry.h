// Internal registry of compiled problem configurations (one mpc_config_impl.cu object per config).
#pragma once
#include <cuda_runtime.h>

struct MpcConfigOps {
    const char* name;
    int N, nx, nu, np, nh, nc, mem_doubles;
    // grid > 0: throughput kernel with `grid` CTAs; grid < 0: latency kernel, one problem per CTA, -grid CTAs
    // mode: MPCGPU_KERNEL_AUTO / _STAGE / _SPLIT (include/mpcgpu.h); a configuration without a role-split kernel ignores _SPLIT
    cudaError_t (*launch_solve)(int grid, int mode, cudaStream_t stream, int n, const double* xinit, const double* x0,
                                const double* params, const int* num_iter, int num_iter_all, double* mem,
                                double* xtraj, double* utraj, double* pobj, int* exit_code, int* qp_status,
                                double* res_eq, int* ipm_iters, int* work_counter);
    cudaError_t (*occupancy)(int* ctas_per_sm, int* threads_per_cta);
    int nhs, model_eval_doubles;
    cudaError_t (*launch_model_eval)(cudaStream_t stream, int n, const double* z, const double* p, const double* pi,
                                     const double* mh, double* out);
    int group_warps;   // warps cooperating on one problem (1 for N <= 31, 2 for N <= 63)
    int has_split;     // 1 when the role-split kernel exists for this configuration
    // MPC_CHECK build only (nullptr otherwise): read (and optionally clear) the diagnostic counters; run the detector's self-test
    cudaError_t (*check_report)(unsigned long long* out8, int reset);
    cudaError_t (*check_selftest)(cudaStream_t stream);
};

void mpc_register_config(const MpcConfigOps* ops);
