// C ABI of the batched MPC solve engine (include/mpcgpu.h).  Host-side plumbing only: device buffers,
// H2D/D2H staging, stream + events, the persistent-grid launch, and the best-planner selection kernel.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mpcgpu.h"
#include "mpc_registry.h"

namespace {
std::vector<const MpcConfigOps*>& registry()
{
    static std::vector<const MpcConfigOps*> r;
    return r;
}

// K7: one thread per homotopy set -- FindBestPlanner (guidance_constraints.cpp:572-590) with the
// objective post-processing of :373-420.  Sequential, ascending, strict '<': bit-exact by construction.
// Consistency term (calculateConsistencyCostForSolver, :1025-1050) from the SOLVED trajectory: squares and sums are kept
// unfused (__dmul_rn / __dadd_rn) and accumulated in stage order so that the value is the one plain C++ computes.
struct ConsArgs {
    const double* xtraj = nullptr;         // [n][N+1][nx] solver output
    const double* prev = nullptr;          // [n_sets][N][2], nullptr: no consistency term
    const unsigned char* enabled = nullptr;
    double weight = 0.0;
    int N = 0, nx = 0, ix = 0, iy = 1;
    double* obj_out = nullptr;
    double* cons_out = nullptr;
};
__global__ void select_best_kernel(int n_sets, const int* __restrict__ set_offsets, const double* __restrict__ pobj,
                                   const int* __restrict__ exit_code, const double* __restrict__ obj_scale,
                                   const double* __restrict__ obj_sub, const unsigned char* __restrict__ disabled,
                                   int* __restrict__ best_idx, const ConsArgs ca)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sets) return;
    double best = 1e10;
    int bi = -1;
    for (int i = set_offsets[s]; i < set_offsets[s + 1]; i++) {
        double obj = pobj[i], cons = 0.0;
        if (ca.prev && (!ca.enabled || ca.enabled[i])) {
            const double* x = ca.xtraj + (size_t)i * (ca.N + 1) * ca.nx;
            const double* pr = ca.prev + (size_t)s * ca.N * 2;
            double sum = 0.0;
            for (int k = 1; k <= ca.N - 2; k++) {
                const double dx = x[k * ca.nx + ca.ix] - pr[2 * k], dy = x[k * ca.nx + ca.iy] - pr[2 * k + 1];
                sum = __dadd_rn(sum, __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            }
            cons = __dmul_rn(ca.weight, sum);
            obj = obj - cons;
        }
        if (obj_sub) obj = obj - obj_sub[i];
        if (obj_scale) obj = __dmul_rn(obj, obj_scale[i]);
        if (ca.obj_out) ca.obj_out[i] = obj;
        if (ca.cons_out) ca.cons_out[i] = cons;
        if (disabled && disabled[i]) continue;
        if (exit_code[i] == 1 && obj < best) { best = obj; bi = i - set_offsets[s]; }
    }
    best_idx[s] = bi;
}
// `*solver = *_solver` resets the QP memory of the planner's capsule and keeps the NLP multipliers
// (acados_solver_interface.cpp:67-77): flag 2 -> 1 in the persistent memory blob.
__global__ void mem_flag_downgrade_kernel(int n, int mem_doubles, double* __restrict__ mem)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && mem[(size_t)i * mem_doubles] > 1.0) mem[(size_t)i * mem_doubles] = 1.0;
}

// The selected planner's trajectory of every set (guidance_constraints.cpp:520-522 copies best_solver->_output into the main
// solver); planner 0 of the set when none succeeded.  One CTA per set, coalesced.
__global__ void gather_best_kernel(int n_sets, const int* __restrict__ set_offsets, const int* __restrict__ best_idx, int sx, int su,
                                   const double* __restrict__ xtraj, const double* __restrict__ utraj, double* __restrict__ bx,
                                   double* __restrict__ bu)
{
    const int s = blockIdx.x;
    if (s >= n_sets) return;
    const int b = best_idx[s];
    const size_t src = (size_t)set_offsets[s] + (b >= 0 ? b : 0);
    for (int i = threadIdx.x; i < sx; i += blockDim.x) bx[(size_t)s * sx + i] = xtraj[src * sx + i];
    for (int i = threadIdx.x; i < su; i += blockDim.x) bu[(size_t)s * su + i] = utraj[src * su + i];
}

// Compact parameter path: params[prob][k][:] <- shared[set][k][:], then the per-planner
// parameters (guidance halfspaces, consistency reference, ...) are scattered over it.  One thread per
// (problem, stage, parameter); coalesced in the parameter index.
__global__ void expand_params_kernel(int n, int planners, int N, int npar, int nidx, const double* __restrict__ shared,
                                     const int* __restrict__ idx, const double* __restrict__ vals, double* __restrict__ params)
{
    const size_t total = (size_t)n * N * npar;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(t % npar);
        const size_t pk_ = t / npar;               // prob * N + k
        const size_t prob = pk_ / N;
        const int k = (int)(pk_ % N);
        params[t] = shared[((prob / planners) * N + k) * npar + j];
    }
}
__global__ void scatter_params_kernel(int n, int N, int npar, int nidx, const int* __restrict__ idx, const double* __restrict__ vals,
                                      double* __restrict__ params)
{
    const size_t total = (size_t)n * N * nidx;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(t % nidx);
        const size_t pk_ = t / nidx;
        params[pk_ * npar + idx[j]] = vals[t];
    }
}
// struct-of-tables path: the per-set block [N][npar] from the stage-invariant values (broadcast over the stages)
// and the per-stage shared values; every other slot has been zeroed
__global__ void scatter_invariant_kernel(int n_sets, int N, int npar, int n_inv, const int* __restrict__ idx, const double* __restrict__ vals,
                                         double* __restrict__ shared)
{
    const size_t total = (size_t)n_sets * N * n_inv;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t % n_inv);
        const size_t sk = t / n_inv;               // set * N + k
        shared[sk * npar + idx[i]] = vals[(sk / N) * n_inv + i];
    }
}
__global__ void scatter_stage_kernel(int n_sets, int N, int npar, int n_stg, const int* __restrict__ idx, const double* __restrict__ vals,
                                     double* __restrict__ shared)
{
    const size_t total = (size_t)n_sets * N * n_stg;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x)
        shared[(t / n_stg) * npar + idx[t % n_stg]] = vals[t];
}
__global__ void repeat_xinit_kernel(int n, int planners, int nx, const double* __restrict__ xs, double* __restrict__ xinit)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n * nx) xinit[t] = xs[(size_t)(t / nx / planners) * nx + t % nx];
}

// FP64 roofline denominator: MEASURED_PEAKS.json has no FP64 entry, so bench.py measures the DFMA peak
// with this kernel: 8 independent FMA chains per thread, 2 flops per FMA.
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 0.999999, c = 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}
}  // namespace

void mpc_register_config(const MpcConfigOps* ops) { registry().push_back(ops); }

struct mpcgpu_engine {
    const MpcConfigOps* ops = nullptr;
    int device = 0, max_batch = 0, grid = 0, sms = 0, threads_per_cta = 0, kernel_mode = 0;
    void* d_obst = nullptr; size_t cap_obst = 0;   // obstacle predictions + guided flags of mpcgpu_solve_sets_guided
    cudaStream_t stream = nullptr, stream2 = nullptr;   // stream2: second lane of the chunked host pipeline
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    static constexpr int MAX_CHUNKS = 12;      // gated path: up to 12 input chunks; legacy path: up to 6 launches
    cudaEvent_t cev0[MAX_CHUNKS] = {}, cev1[MAX_CHUNKS] = {};   // per chunk (host path)
    int chunks_timed = 0;      // > 0: last_kernel_ms sums the chunk kernels of the last host call
    // device staging for the host-pointer entry points
    double *d_xinit = nullptr, *d_x0 = nullptr, *d_params = nullptr, *d_mem = nullptr, *d_xtraj = nullptr, *d_utraj = nullptr,
           *d_pobj = nullptr, *d_res_eq = nullptr, *d_scale = nullptr, *d_sub = nullptr;
    int *d_num_iter = nullptr, *d_exit = nullptr, *d_qps = nullptr, *d_ipm = nullptr, *d_counters = nullptr, *d_offsets = nullptr,
        *d_best = nullptr;
    unsigned long long launch_seq = 0;   // every solve launch takes its own work counter from the ring d_counters[MPCGPU_MAX_INFLIGHT]
    int* next_counter() { return d_counters + 2 * (launch_seq++ % MPCGPU_MAX_INFLIGHT); }      // {work counter, input gate}
    int* h_gate = nullptr;               // pinned: gate values of the host pipeline
    cudaEvent_t ev_gate = nullptr;
    double *d_prev = nullptr, *d_objout = nullptr, *d_consout = nullptr, *d_static = nullptr;   // set options, allocated on first use
    double* d_tab = nullptr; size_t cap_tab = 0;      // struct-of-tables inputs: [invariant | stage | obstacle radius] doubles, then the index arrays
    int* d_tabidx = nullptr; size_t cap_tabidx = 0;
    unsigned char* d_consen = nullptr;
    size_t cap_prev = 0, cap_static = 0;
    unsigned char* d_disabled = nullptr;
    double *d_shared = nullptr, *d_pvals = nullptr, *d_xs = nullptr;   // compact (per-set) inputs, allocated on first use
    int* d_pidx = nullptr;
    size_t cap_shared = 0, cap_pvals = 0;
    long long launches = 0;
    std::string err;
};

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t _e = (call);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            e->err = std::string(#call) + ": " + cudaGetErrorString(_e);                              \
            return MPCGPU_ERR_CUDA;                                                                   \
        }                                                                                             \
    } while (0)

extern "C" {

int mpcgpu_num_configs(void) { return (int)registry().size(); }
const char* mpcgpu_config_name(int i) { return (i >= 0 && i < (int)registry().size()) ? registry()[i]->name : nullptr; }

int mpcgpu_engine_create(const char* config_name, int device, int max_batch, mpcgpu_engine** out)
{
    if (!config_name || !out || max_batch <= 0) return MPCGPU_ERR_ARG;
    *out = nullptr;
    const MpcConfigOps* ops = nullptr;
    for (auto* o : registry())
        if (std::strcmp(o->name, config_name) == 0) ops = o;
    if (!ops) return MPCGPU_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
        std::fprintf(stderr, "mpcgpu: no usable CUDA device (requested %d of %d); there is no CPU fallback\n", device, ndev);
        return MPCGPU_ERR_NO_DEVICE;
    }
    mpcgpu_engine* e = new mpcgpu_engine();
    e->ops = ops; e->device = device; e->max_batch = max_batch;
    auto fail = [&](int code) { *out = e; return code; };   // caller can read last_error, then destroy
    if (cudaSetDevice(device) != cudaSuccess) { e->err = "cudaSetDevice failed"; return fail(MPCGPU_ERR_CUDA); }
    const size_t B = (size_t)max_batch;
    const int N = ops->N, nx = ops->nx, nu = ops->nu, nz = nx + nu;
#define AL(ptr, count) do { cudaError_t _e = cudaMalloc((void**)&(ptr), (count)); if (_e != cudaSuccess) { e->err = std::string("cudaMalloc ") + #ptr + ": " + cudaGetErrorString(_e); return fail(MPCGPU_ERR_CUDA); } } while (0)
    AL(e->d_xinit, B * nx * 8); AL(e->d_x0, B * nz * (N + 1) * 8); AL(e->d_params, B * N * ops->np * 8);
    AL(e->d_mem, B * ops->mem_doubles * 8); AL(e->d_xtraj, B * nx * (N + 1) * 8); AL(e->d_utraj, B * nu * N * 8);
    AL(e->d_pobj, B * 8); AL(e->d_res_eq, B * 8); AL(e->d_scale, B * 8); AL(e->d_sub, B * 8);
    AL(e->d_num_iter, B * 4); AL(e->d_exit, B * 4); AL(e->d_qps, B * 4); AL(e->d_ipm, B * 4); AL(e->d_counters, 8 * MPCGPU_MAX_INFLIGHT);
    AL(e->d_offsets, (B + 1) * 4); AL(e->d_best, B * 4); AL(e->d_disabled, B);
#undef AL
    if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&e->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&e->ev0) != cudaSuccess || cudaEventCreate(&e->ev1) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_gate, cudaEventDisableTiming) != cudaSuccess ||
        cudaHostAlloc((void**)&e->h_gate, (mpcgpu_engine::MAX_CHUNKS + 2) * sizeof(int), cudaHostAllocDefault) != cudaSuccess ||
        cudaMemset(e->d_counters, 0, 8 * MPCGPU_MAX_INFLIGHT) != cudaSuccess ||
        [&] { for (int i = 0; i < mpcgpu_engine::MAX_CHUNKS; i++) if (cudaEventCreate(&e->cev0[i]) != cudaSuccess || cudaEventCreate(&e->cev1[i]) != cudaSuccess) return true; return false; }()) {
        e->err = "stream/event creation failed";
        return fail(MPCGPU_ERR_CUDA);
    }
    int ctas = 0, threads = 0, sms = 0;
    if (ops->occupancy(&ctas, &threads) != cudaSuccess || ctas <= 0 ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) {
        e->err = "occupancy query failed";
        return fail(MPCGPU_ERR_CUDA);
    }
    if (const char* ov = std::getenv("MPCGPU_CTAS_PER_SM")) {      // tuning knob (tools/sweep experiments): cap the resident CTAs per SM
        const int c = std::atoi(ov);
        if (c > 0 && c < ctas) ctas = c;
    }
    e->grid = sms * ctas;   // persistent grid: every SM holds its full complement of CTAs
    e->sms = sms;
    e->threads_per_cta = threads;
    *out = e;
    return MPCGPU_OK;
}

int mpcgpu_engine_destroy(mpcgpu_engine* e)
{
    if (!e) return MPCGPU_ERR_ARG;
    cudaSetDevice(e->device);
    void* ptrs[] = {e->d_obst, e->d_shared, e->d_pvals, e->d_xs, e->d_pidx, e->d_xinit, e->d_x0, e->d_params, e->d_mem, e->d_xtraj, e->d_utraj, e->d_pobj, e->d_res_eq, e->d_scale,
                    e->d_sub, e->d_num_iter, e->d_exit, e->d_qps, e->d_ipm, e->d_counters, e->d_offsets, e->d_best, e->d_disabled,
                    e->d_prev, e->d_objout, e->d_consout, e->d_static, e->d_consen, e->d_tab, e->d_tabidx};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->ev_gate) cudaEventDestroy(e->ev_gate);
    if (e->h_gate) cudaFreeHost(e->h_gate);
    for (int i = 0; i < mpcgpu_engine::MAX_CHUNKS; i++) { if (e->cev0[i]) cudaEventDestroy(e->cev0[i]); if (e->cev1[i]) cudaEventDestroy(e->cev1[i]); }
    if (e->stream) cudaStreamDestroy(e->stream);
    if (e->stream2) cudaStreamDestroy(e->stream2);
    delete e;
    return MPCGPU_OK;
}

int mpcgpu_desc_query(const mpcgpu_engine* e, int* N, int* nx, int* nu, int* npar, int* nh)
{
    if (!e) return MPCGPU_ERR_ARG;
    if (N) *N = e->ops->N;
    if (nx) *nx = e->ops->nx;
    if (nu) *nu = e->ops->nu;
    if (npar) *npar = e->ops->np;
    if (nh) *nh = e->ops->nh;
    return MPCGPU_OK;
}
int mpcgpu_mem_doubles(const mpcgpu_engine* e) { return e ? e->ops->mem_doubles : MPCGPU_ERR_ARG; }

static int launch_solve_on(mpcgpu_engine* e, cudaStream_t st, int* counter, cudaEvent_t t0, cudaEvent_t t1, int n, const double* xinit,
                           const double* x0, const double* params, const int* num_iter, int num_iter_all, double* mem_inout,
                           double* xtraj, double* utraj, double* pobj, int* exit_code, int* qp_status, double* res_eq, int* ipm_iters,
                           bool gated = false);

int mpcgpu_solve_batch_device(mpcgpu_engine* e, int n, const double* xinit, const double* x0, const double* params,
                              const int* num_iter, int num_iter_all, double* mem_inout, double* xtraj, double* utraj,
                              double* pobj, int* exit_code, int* qp_status, double* res_eq, int* ipm_iters, void* stream)
{
    if (!e || n < 0 || !xinit || !x0 || !params || !xtraj || !utraj || !pobj || !exit_code || !qp_status || !res_eq)
        return MPCGPU_ERR_ARG;
    if (n == 0) return MPCGPU_OK;
    CK(cudaSetDevice(e->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    e->chunks_timed = 0;
    return launch_solve_on(e, st, e->next_counter(), e->ev0, e->ev1, n, xinit, x0, params, num_iter, num_iter_all, mem_inout, xtraj, utraj,
                           pobj, exit_code, qp_status, res_eq, ipm_iters);
}

static int launch_solve_on(mpcgpu_engine* e, cudaStream_t st, int* counter, cudaEvent_t t0, cudaEvent_t t1, int n, const double* xinit,
                           const double* x0, const double* params, const int* num_iter, int num_iter_all, double* mem_inout,
                           double* xtraj, double* utraj, double* pobj, int* exit_code, int* qp_status, double* res_eq, int* ipm_iters,
                           bool gated)
{
    // small batches (one or a few homotopy sets): one problem per CTA so every problem gets its own SM;
    // otherwise the persistent throughput grid
    int grid;
    if (n <= 2 * e->sms) {
        grid = -n;
    } else {
        const int groups_per_cta = e->threads_per_cta / 32 / e->ops->group_warps;
        grid = (n + groups_per_cta - 1) / groups_per_cta;
        if (grid > e->grid) grid = e->grid;
    }
    CK(cudaEventRecord(t0, st));
    CK(e->ops->launch_solve(grid, e->kernel_mode | (gated ? 0x100 : 0), st, n, xinit, x0, params, num_iter, num_iter_all, mem_inout, xtraj, utraj, pobj, exit_code,
                            qp_status, res_eq, ipm_iters, counter));
    CK(cudaEventRecord(t1, st));
    e->launches += 1;
    return MPCGPU_OK;
}

int mpcgpu_sync(mpcgpu_engine* e)
{
    if (!e) return MPCGPU_ERR_ARG;
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaStreamSynchronize(e->stream2));
    CK(cudaEventSynchronize(e->ev1));      // recorded behind the most recent device call, whatever stream it went to
    return MPCGPU_OK;
}

// drain both engine streams before an error return: earlier asynchronous copies may still touch the caller's host buffers
static int drain(mpcgpu_engine* e, int rc)
{
    if (rc != MPCGPU_OK) { cudaStreamSynchronize(e->stream); cudaStreamSynchronize(e->stream2); }
    return rc;
}
static int solve_batch_impl(mpcgpu_engine* e, int n, const double* xinit, const double* x0, const double* params,
                            const int* num_iter, int num_iter_all, double* mem_inout, double* xtraj, double* utraj, double* pobj,
                            int* exit_code, int* qp_status, double* res_eq, int* ipm_iters);

int mpcgpu_solve_batch(mpcgpu_engine* e, int n, const double* xinit, const double* x0, const double* params,
                       const int* num_iter, int num_iter_all, double* mem_inout, double* xtraj, double* utraj, double* pobj,
                       int* exit_code, int* qp_status, double* res_eq, int* ipm_iters)
{
    if (!e || n < 0 || n > e->max_batch || !xinit || !x0 || !params || !xtraj || !utraj || !pobj || !exit_code ||
        !qp_status || !res_eq)
        return MPCGPU_ERR_ARG;
    if (n == 0) return MPCGPU_OK;
    return drain(e, solve_batch_impl(e, n, xinit, x0, params, num_iter, num_iter_all, mem_inout, xtraj, utraj, pobj, exit_code, qp_status,
                                     res_eq, ipm_iters));
}

static int solve_batch_impl(mpcgpu_engine* e, int n, const double* xinit, const double* x0, const double* params,
                       const int* num_iter, int num_iter_all, double* mem_inout, double* xtraj, double* utraj, double* pobj,
                       int* exit_code, int* qp_status, double* res_eq, int* ipm_iters)
{
    if (!e || n < 0 || n > e->max_batch || !xinit || !x0 || !params || !xtraj || !utraj || !pobj || !exit_code ||
        !qp_status || !res_eq)
        return MPCGPU_ERR_ARG;
    if (n == 0) return MPCGPU_OK;
    CK(cudaSetDevice(e->device));
    const MpcConfigOps* o = e->ops;
    const size_t B = (size_t)n;
    const int N = o->N, nx = o->nx, nu = o->nu, nz = nx + nu;
    const size_t sx0 = (size_t)nz * (N + 1), spar = (size_t)N * o->np, sxt = (size_t)nx * (N + 1), sut = (size_t)nu * N,
                 smem_ = (size_t)o->mem_doubles;
    // ---- Gated single launch (large batches from PINNED host memory).  Separate launches per chunk leave every SM partly idle
    // while the last warps of a chunk finish (a CTA of the next launch needs the whole SM): ~6 % of a 36 864-problem step.  Instead
    // ONE persistent launch takes the whole batch; the copy stream delivers the inputs chunk by chunk and raises a gate word
    // behind every chunk, and a warp that draws a problem beyond the gate waits for it (copies run ~4x ahead of the solves, so
    // only the first chunk is ever waited for).  Chunk boundaries are multiples of 16 problems: a boundary then falls on a 128-byte
    // line of every input array, so no SM can have cached a line that a later copy completes.  Results: written by the kernel
    // straight into the caller's arrays when those are pinned (mapped) host memory, else staged and copied once at the end.
    auto pinned_dev = [](const void* h, bool need) -> void* {
        if (!h) return need ? nullptr : (void*)1;
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, h) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        return (a.type == cudaMemoryTypeHost) ? a.devicePointer : nullptr;
    };
    if (B >= 8192 && e->kernel_mode != MPCGPU_KERNEL_SPLIT && pinned_dev(xinit, true) && pinned_dev(x0, true) && pinned_dev(params, true) && pinned_dev(num_iter, false) &&
        pinned_dev(mem_inout, false)) {
        size_t gb[mpcgpu_engine::MAX_CHUNKS + 1];
        int ng = 0;
        gb[0] = 0;
        gb[++ng] = 2048;                                    // enough to start every warp of the grid; 1.6 ms of copy for c2
        size_t per_ = (B - 2048 + (mpcgpu_engine::MAX_CHUNKS - 2)) / (mpcgpu_engine::MAX_CHUNKS - 1);
        if (per_ < 4096) per_ = 4096;
        per_ = (per_ + 15) & ~(size_t)15;
        while (gb[ng] < B) { const size_t hi = gb[ng] + per_; gb[ng + 1] = hi < B ? hi : B; ng++; }
        int* slot = e->next_counter();
        e->h_gate[0] = -1;
        for (int c = 0; c < ng; c++) e->h_gate[c + 1] = (c + 1 == ng) ? 0 : -(int)gb[c + 1] - 1;      // the last chunk opens the gate for good
        cudaStream_t cs = e->stream2;
        CK(cudaMemsetAsync(slot, 0, sizeof(int), cs));
        CK(cudaMemcpyAsync(slot + 1, &e->h_gate[0], sizeof(int), cudaMemcpyHostToDevice, cs));
        CK(cudaEventRecord(e->ev_gate, cs));
        CK(cudaStreamWaitEvent(e->stream, e->ev_gate, 0));
        for (int c = 0; c < ng; c++) {
            const size_t b0 = gb[c], m = gb[c + 1] - b0;
            CK(cudaMemcpyAsync(e->d_xinit + b0 * nx, xinit + b0 * nx, m * nx * 8, cudaMemcpyHostToDevice, cs));
            CK(cudaMemcpyAsync(e->d_x0 + b0 * sx0, x0 + b0 * sx0, m * sx0 * 8, cudaMemcpyHostToDevice, cs));
            CK(cudaMemcpyAsync(e->d_params + b0 * spar, params + b0 * spar, m * spar * 8, cudaMemcpyHostToDevice, cs));
            if (num_iter) CK(cudaMemcpyAsync(e->d_num_iter + b0, num_iter + b0, m * 4, cudaMemcpyHostToDevice, cs));
            if (mem_inout) CK(cudaMemcpyAsync(e->d_mem + b0 * smem_, mem_inout + b0 * smem_, m * smem_ * 8, cudaMemcpyHostToDevice, cs));
            CK(cudaMemcpyAsync(slot + 1, &e->h_gate[c + 1], sizeof(int), cudaMemcpyHostToDevice, cs));
        }
        // every copy and every gate update is enqueued: the launch cannot be left waiting for a copy that was never issued
        void* zx = pinned_dev(xtraj, true); void* zu = pinned_dev(utraj, true); void* zp = pinned_dev(pobj, true);
        void* ze = pinned_dev(exit_code, true); void* zq = pinned_dev(qp_status, true); void* zr = pinned_dev(res_eq, true);
        void* zi = ipm_iters ? pinned_dev(ipm_iters, true) : nullptr;
        const bool direct = zx && zu && zp && ze && zq && zr && (!ipm_iters || zi);
        int rc = launch_solve_on(e, e->stream, slot, e->ev0, e->ev1, n, e->d_xinit, e->d_x0, e->d_params, num_iter ? e->d_num_iter : nullptr, num_iter_all,
                                 mem_inout ? e->d_mem : nullptr, direct ? (double*)zx : e->d_xtraj, direct ? (double*)zu : e->d_utraj,
                                 direct ? (double*)zp : e->d_pobj, direct ? (int*)ze : e->d_exit, direct ? (int*)zq : e->d_qps,
                                 direct ? (double*)zr : e->d_res_eq, direct ? (int*)zi : e->d_ipm, true);
        if (rc != MPCGPU_OK) return rc;
        if (!direct) {
            CK(cudaMemcpyAsync(xtraj, e->d_xtraj, B * sxt * 8, cudaMemcpyDeviceToHost, e->stream));
            CK(cudaMemcpyAsync(utraj, e->d_utraj, B * sut * 8, cudaMemcpyDeviceToHost, e->stream));
            CK(cudaMemcpyAsync(pobj, e->d_pobj, B * 8, cudaMemcpyDeviceToHost, e->stream));
            CK(cudaMemcpyAsync(exit_code, e->d_exit, B * 4, cudaMemcpyDeviceToHost, e->stream));
            CK(cudaMemcpyAsync(qp_status, e->d_qps, B * 4, cudaMemcpyDeviceToHost, e->stream));
            CK(cudaMemcpyAsync(res_eq, e->d_res_eq, B * 8, cudaMemcpyDeviceToHost, e->stream));
            if (ipm_iters) CK(cudaMemcpyAsync(ipm_iters, e->d_ipm, B * 4, cudaMemcpyDeviceToHost, e->stream));
        }
        if (mem_inout) CK(cudaMemcpyAsync(mem_inout, e->d_mem, B * smem_ * 8, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        CK(cudaStreamSynchronize(e->stream2));
        e->chunks_timed = 0;
        return MPCGPU_OK;
    }
    // Chunked two-stream pipeline: the H2D copy of chunk c+1 and the D2H copy of chunk c-1 overlap the solve
    // kernel of chunk c (pinned host memory makes the copies truly asynchronous).  Chunks stay >= 8192
    // problems so that the persistent grid keeps several waves per launch.
    // A small first chunk (1/16 of the batch, at least 2048 problems) shortens the one copy nothing can overlap.
    size_t bounds[mpcgpu_engine::MAX_CHUNKS + 1];
    int nchunk = 0;
    bounds[0] = 0;
    if (B >= 4 * 8192) {
        size_t first = B / 16;
        if (first < 2048) first = 2048;
        bounds[++nchunk] = first;
    }
    {
        const size_t rest = B - bounds[nchunk];
        int k = (int)(rest / 8192);
        if (k < 1) k = 1;
        if (k > 4) k = 4;
        const size_t per_ = (rest + k - 1) / k;
        for (int i = 0; i < k; i++) {
            const size_t hi = bounds[nchunk] + per_;
            bounds[nchunk + 1] = hi < B ? hi : B;
            nchunk++;
        }
    }
    const bool out_pinned = nchunk == 1 || (pinned_dev(xtraj, true) && pinned_dev(utraj, true) && pinned_dev(pobj, true) && pinned_dev(exit_code, true) &&
                                            pinned_dev(qp_status, true) && pinned_dev(res_eq, true) && pinned_dev(ipm_iters, false) &&
                                            pinned_dev(mem_inout, false));      // (one chunk: nothing to overlap, copy in stream order)
    for (int c = 0; c < nchunk; c++) {
        const size_t b0 = bounds[c];
        if (b0 >= B || bounds[c + 1] <= b0) { nchunk = c; break; }
        const size_t m = bounds[c + 1] - b0;
        cudaStream_t st = (c & 1) ? e->stream2 : e->stream;
        int* counter = e->next_counter();
        CK(cudaMemcpyAsync(e->d_xinit + b0 * nx, xinit + b0 * nx, m * nx * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(e->d_x0 + b0 * sx0, x0 + b0 * sx0, m * sx0 * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(e->d_params + b0 * spar, params + b0 * spar, m * spar * 8, cudaMemcpyHostToDevice, st));
        if (num_iter) CK(cudaMemcpyAsync(e->d_num_iter + b0, num_iter + b0, m * 4, cudaMemcpyHostToDevice, st));
        if (mem_inout) CK(cudaMemcpyAsync(e->d_mem + b0 * smem_, mem_inout + b0 * smem_, m * smem_ * 8, cudaMemcpyHostToDevice, st));
        int rc = launch_solve_on(e, st, counter, e->cev0[c], e->cev1[c], (int)m, e->d_xinit + b0 * nx, e->d_x0 + b0 * sx0, e->d_params + b0 * spar,
                                 num_iter ? e->d_num_iter + b0 : nullptr, num_iter_all, mem_inout ? e->d_mem + b0 * smem_ : nullptr,
                                 e->d_xtraj + b0 * sxt, e->d_utraj + b0 * sut, e->d_pobj + b0, e->d_exit + b0, e->d_qps + b0,
                                 e->d_res_eq + b0, e->d_ipm + b0);
        if (rc != MPCGPU_OK) return rc;
        if (!out_pinned) continue;      // a copy into pageable memory blocks the host until the chunk is solved: it would serialise the pipeline
        CK(cudaMemcpyAsync(xtraj + b0 * sxt, e->d_xtraj + b0 * sxt, m * sxt * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(utraj + b0 * sut, e->d_utraj + b0 * sut, m * sut * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(pobj + b0, e->d_pobj + b0, m * 8, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(exit_code + b0, e->d_exit + b0, m * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(qp_status + b0, e->d_qps + b0, m * 4, cudaMemcpyDeviceToHost, st));
        CK(cudaMemcpyAsync(res_eq + b0, e->d_res_eq + b0, m * 8, cudaMemcpyDeviceToHost, st));
        if (ipm_iters) CK(cudaMemcpyAsync(ipm_iters + b0, e->d_ipm + b0, m * 4, cudaMemcpyDeviceToHost, st));
        if (mem_inout) CK(cudaMemcpyAsync(mem_inout + b0 * smem_, e->d_mem + b0 * smem_, m * smem_ * 8, cudaMemcpyDeviceToHost, st));
    }
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaStreamSynchronize(e->stream2));
    if (!out_pinned) {                  // pageable outputs: everything at the end, one copy per array
        CK(cudaMemcpy(xtraj, e->d_xtraj, B * sxt * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(utraj, e->d_utraj, B * sut * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(pobj, e->d_pobj, B * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(exit_code, e->d_exit, B * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(qp_status, e->d_qps, B * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(res_eq, e->d_res_eq, B * 8, cudaMemcpyDeviceToHost));
        if (ipm_iters) CK(cudaMemcpy(ipm_iters, e->d_ipm, B * 4, cudaMemcpyDeviceToHost));
        if (mem_inout) CK(cudaMemcpy(mem_inout, e->d_mem, B * smem_ * 8, cudaMemcpyDeviceToHost));
    }
    e->chunks_timed = nchunk;
    return MPCGPU_OK;
}

int mpcgpu_select_best_device(mpcgpu_engine* e, int n_sets, const int* set_offsets, const double* pobj, const int* exit_code,
                              const double* obj_scale, const double* obj_sub, const unsigned char* disabled, int* best_idx,
                              void* stream)
{
    if (!e || n_sets < 0 || !set_offsets || !pobj || !exit_code || !best_idx) return MPCGPU_ERR_ARG;
    if (n_sets == 0) return MPCGPU_OK;
    CK(cudaSetDevice(e->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    select_best_kernel<<<(n_sets + 127) / 128, 128, 0, st>>>(n_sets, set_offsets, pobj, exit_code, obj_scale, obj_sub, disabled,
                                                              best_idx, ConsArgs());
    CK(cudaGetLastError());
    e->launches += 1;
    return MPCGPU_OK;
}

// ---- guidance halfspaces on the device (contract: include/mpcgpu.h, restated in oracle/mpc_oracle.c) ----
// One thread per (problem, stage): the work GuidanceConstraints::optimize does per planner before solve()
// (linearized_constraints.cpp:49-189) -- instead of uploading 3 * max_obstacles doubles per planner and stage.
// products and sums are kept unfused (__dmul_rn / __dadd_rn) so that the result is bit-identical to the host restatement
__device__ __forceinline__ double norm2_dev(double dx, double dy) { return sqrt(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))); }
__device__ __forceinline__ void dr_proj_dev(const double* p, const double* c, double r, const double* toward, double* out)
{
    const double dx = p[0] - c[0], dy = p[1] - c[1];
    if (norm2_dev(dx, dy) < r) {
        const double sx = toward[0] - c[0], sy = toward[1] - c[1], n = norm2_dev(sx, sy);
        out[0] = __dadd_rn(c[0], __dmul_rn(sx / n, r)); out[1] = __dadd_rn(c[1], __dmul_rn(sy / n, r));
    } else { out[0] = p[0]; out[1] = p[1]; }
}
__global__ void guidance_halfspaces_kernel(int n, int planners, int N, int nx, int nu, int npar, int lin_base, int lin_count, int n_obs, int obs,
                                           const double* __restrict__ xinit_sets, const double* __restrict__ x0,
                                           const double* __restrict__ obst_pred, const unsigned char* __restrict__ guided,
                                           double robot_radius, const double* __restrict__ stat, int n_static, double* __restrict__ params)
{
    const long long total = (long long)n * N;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(t / N), k = (int)(t % N), s = q / planners, nz = nx + nu;
        const double r = 1e-3 + robot_radius;
        const double dummy_b = xinit_sets[(size_t)s * nx] + 100.0;
        double* P = params + ((size_t)q * N + k) * npar + lin_base;
        const bool act = k > 0 && guided[q];
        double pos[2] = {0.0, 0.0};
        const double* ob = obst_pred + ((size_t)s * N + (k > 0 ? k - 1 : 0)) * n_obs * obs;      // obs doubles per obstacle, (x, y) first
        if (act) {
            pos[0] = x0[((size_t)q * (N + 1) + k) * nz + nu]; pos[1] = x0[((size_t)q * (N + 1) + k) * nz + nu + 1];
            for (int it = 0; it < 3; it++)
                for (int j = 0; j < n_obs; j++) {
                    const double dx = pos[0] - ob[obs * j], dy = pos[1] - ob[obs * j + 1];
                    if (norm2_dev(dx, dy) < r) {
                        double pa[2], ra[2], pb[2];
                        dr_proj_dev(pos, ob, r, pos, pa);
                        ra[0] = 2.0 * pa[0] - pos[0]; ra[1] = 2.0 * pa[1] - pos[1];
                        dr_proj_dev(ra, ob + obs * j, r, pos, pb);
                        pos[0] = 0.5 * (pos[0] + 2.0 * pb[0] - ra[0]); pos[1] = 0.5 * (pos[1] + 2.0 * pb[1] - ra[1]);
                    }
                }
        }
        for (int j = 0; j < lin_count; j++) {
            double a1 = 1.0, a2 = 0.0, b = dummy_b;
            if (act && j < n_obs) {
                const double ox = ob[obs * j], oy = ob[obs * j + 1];
                const double dx = ox - pos[0], dy = oy - pos[1], dist = norm2_dev(dx, dy);
                a1 = dx / dist; a2 = dy / dist;
                b = __dsub_rn(__dadd_rn(__dmul_rn(a1, ox), __dmul_rn(a2, oy)), r);      // no FMA contraction: bit-identical to the host restatement
            }
            P[3 * j] = a1; P[3 * j + 1] = a2; P[3 * j + 2] = b;
        }
        // module_data.static_obstacles (linearized_constraints.cpp:107-127): behind the obstacle rows; a non-guided planner's
        // obstacle list is empty (update(state, empty_data_, ...), guidance_constraints.cpp:326-329), so its rows start at 0
        if (stat && k > 0) {
            const int first = guided[q] ? n_obs : 0;
            const double* sh = stat + ((size_t)s * N + k) * n_static * 3;
            for (int h = 0; h < n_static && first + h < lin_count; h++) {
                P[3 * (first + h)] = sh[3 * h]; P[3 * (first + h) + 1] = sh[3 * h + 1]; P[3 * (first + h) + 2] = sh[3 * h + 2];
            }
        }
    }
}

static int guidance_halfspaces_launch(mpcgpu_engine* e, int n_sets, int planners, const double* xinit_sets, const double* x0,
                                      const double* obst_pred, int n_obs, int ob_stride, const unsigned char* guided, int lin_base,
                                      int lin_count, double robot_radius, const double* stat, int n_static, double* params, void* stream);

int mpcgpu_guidance_halfspaces_device(mpcgpu_engine* e, int n_sets, int planners, const double* xinit_sets, const double* x0,
                                      const double* obst_pred, int n_obs, const unsigned char* guided, int lin_base, int lin_count,
                                      double robot_radius, double* params, void* stream)
{
    return guidance_halfspaces_launch(e, n_sets, planners, xinit_sets, x0, obst_pred, n_obs, 2, guided, lin_base, lin_count, robot_radius, nullptr, 0,
                                      params, stream);
}

static int guidance_halfspaces_launch(mpcgpu_engine* e, int n_sets, int planners, const double* xinit_sets, const double* x0,
                                      const double* obst_pred, int n_obs, int ob_stride, const unsigned char* guided, int lin_base,
                                      int lin_count, double robot_radius, const double* stat, int n_static, double* params, void* stream)
{
    if (!e || n_sets < 0 || planners <= 0 || !xinit_sets || !x0 || (n_obs > 0 && !obst_pred) || n_obs < 0 || !guided || !params || lin_count < 0 ||
        lin_base < 0 || lin_base + 3 * lin_count > e->ops->np)
        return MPCGPU_ERR_ARG;
    const long long n = (long long)n_sets * planners;
    if (n == 0 || lin_count == 0) return MPCGPU_OK;
    CK(cudaSetDevice(e->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
    const MpcConfigOps* o = e->ops;
    const long long total = n * o->N;
    const int blocks = (int)((total + 127) / 128 < 148 * 16 ? (total + 127) / 128 : 148 * 16);
    guidance_halfspaces_kernel<<<blocks, 128, 0, st>>>((int)n, planners, o->N, o->nx, o->nu, o->np, lin_base, lin_count, n_obs, ob_stride, xinit_sets, x0,
                                                        obst_pred, guided, robot_radius, stat, n_static, params);
    CK(cudaGetLastError());
    e->launches += 1;
    return MPCGPU_OK;
}

struct GuidedArgs {
    int n_obs, lin_base, lin_count;
    const double* obst_pred;             // [n_sets][N][n_obs][ob_stride], (x, y) first
    const unsigned char* guided;
    double robot_radius;
    int ob_stride = 2;                   // 4: obstacle tables (x, y, psi, r) of include/mpcgpu_wire.h
    int ell_base = -1, ell_stride = 0, ell_off[7] = {0, 0, 0, 0, 0, 0, 0};      // >= 0: ellipsoid slots written from the table
    const mpcgpu_param_tables* tab = nullptr;      // struct-of-tables path: the shared block is BUILT on the device (no shared_params)
};
static int pack_obstacles_launch(mpcgpu_engine* e, int n_sets, const double* xinit_sets, const double* table, int obs, const double* radius, int M,
                                 int ell_base, int ell_stride, const int* off, double* params, void* stream);
static int solve_sets_impl(mpcgpu_engine* e, int n_sets, int planners, const double* xinit_sets, const double* shared_params,
                           const double* x0, int nidx, const int* param_idx, const double* planner_params, const GuidedArgs* ga,
                           const int* num_iter, int num_iter_all, double* xtraj, double* utraj, double* pobj, int* exit_code,
                           int* qp_status, double* res_eq, const double* obj_scale, const double* obj_sub, const unsigned char* disabled,
                           int* best_idx, const mpcgpu_set_options* opt);
extern "C" int mpcgpu_pack_obstacles_device(mpcgpu_engine* e, int n_sets, const double* xinit_sets, const double* table, int M, int ell_base,
                                            int ell_stride, const int* ell_offsets, double* params, void* stream);

int mpcgpu_solve_sets(mpcgpu_engine* e, int n_sets, int planners, const double* xinit_sets, const double* shared_params,
                      const double* x0, int nidx, const int* param_idx, const double* planner_params, const int* num_iter,
                      int num_iter_all, double* xtraj, double* utraj, double* pobj, int* exit_code, int* qp_status, double* res_eq,
                      const double* obj_scale, const double* obj_sub, const unsigned char* disabled, int* best_idx,
                      const mpcgpu_set_options* opt)
{
    if (!e) return MPCGPU_ERR_ARG;
    return drain(e, solve_sets_impl(e, n_sets, planners, xinit_sets, shared_params, x0, nidx, param_idx, planner_params, nullptr, num_iter,
                                    num_iter_all, xtraj, utraj, pobj, exit_code, qp_status, res_eq, obj_scale, obj_sub, disabled, best_idx, opt));
}

int mpcgpu_solve_sets_guided(mpcgpu_engine* e, int n_sets, int planners, const double* xinit_sets, const double* shared_params,
                             const double* x0, int n_obs, const double* obst_pred, const unsigned char* guided, int lin_base,
                             int lin_count, double robot_radius, int nidx, const int* param_idx, const double* planner_params,
                             const int* num_iter, int num_iter_all, double* xtraj, double* utraj, double* pobj, int* exit_code,
                             int* qp_status, double* res_eq, const double* obj_scale, const double* obj_sub,
                             const unsigned char* disabled, int* best_idx, const mpcgpu_set_options* opt)
{
    if (!e || n_obs < 0 || (n_obs > 0 && !obst_pred) || !guided || lin_base < 0 || lin_count < 0 || lin_base + 3 * lin_count > e->ops->np)
        return MPCGPU_ERR_ARG;
    const GuidedArgs ga = {n_obs, lin_base, lin_count, obst_pred, guided, robot_radius};
    return drain(e, solve_sets_impl(e, n_sets, planners, xinit_sets, shared_params, x0, nidx, param_idx, planner_params, &ga, num_iter, num_iter_all,
                                    xtraj, utraj, pobj, exit_code, qp_status, res_eq, obj_scale, obj_sub, disabled, best_idx, opt));
}

static int solve_sets_impl(mpcgpu_engine* e, int n_sets, int planners, const double* xinit_sets, const double* shared_params,
                           const double* x0, int nidx, const int* param_idx, const double* planner_params, const GuidedArgs* ga,
                           const int* num_iter, int num_iter_all, double* xtraj, double* utraj, double* pobj, int* exit_code,
                           int* qp_status, double* res_eq, const double* obj_scale, const double* obj_sub, const unsigned char* disabled,
                           int* best_idx, const mpcgpu_set_options* opt)
{
    const mpcgpu_param_tables* tab = ga ? ga->tab : nullptr;
    if (!e || n_sets < 0 || planners <= 0 || !xinit_sets || (!shared_params && !tab) || !x0 || nidx < 0 || (nidx > 0 && (!param_idx || !planner_params)) ||
        !pobj || !exit_code || !qp_status || !res_eq || !best_idx)
        return MPCGPU_ERR_ARG;
    // per-planner trajectories are optional when the selected trajectory of every set is asked for instead
    if ((!xtraj || !utraj) && !(opt && opt->best_xtraj && opt->best_utraj)) return MPCGPU_ERR_ARG;
    const long long nll = (long long)n_sets * planners;
    if (nll > e->max_batch) return MPCGPU_ERR_ARG;
    const int n = (int)nll;
    if (n == 0) return MPCGPU_OK;
    CK(cudaSetDevice(e->device));
    const MpcConfigOps* o = e->ops;
    const int N = o->N, nx = o->nx, nu = o->nu, nz = nx + nu, np = o->np;
    for (int j = 0; j < nidx; j++)
        if (param_idx[j] < 0 || param_idx[j] >= np) return MPCGPU_ERR_ARG;
    cudaStream_t st = e->stream;
    const size_t need_sh = (size_t)n_sets * N * np, need_pv = (size_t)n * N * (nidx > 0 ? nidx : 1);
    if (need_sh > e->cap_shared) {
        if (e->d_shared) cudaFree(e->d_shared);
        if (e->d_xs) cudaFree(e->d_xs);
        e->d_shared = nullptr; e->d_xs = nullptr; e->cap_shared = 0;
        CK(cudaMalloc((void**)&e->d_shared, need_sh * 8));
        CK(cudaMalloc((void**)&e->d_xs, (size_t)n_sets * nx * 8));
        e->cap_shared = need_sh;
    }
    if (need_pv > e->cap_pvals) {
        if (e->d_pvals) cudaFree(e->d_pvals);
        e->d_pvals = nullptr; e->cap_pvals = 0;
        CK(cudaMalloc((void**)&e->d_pvals, need_pv * 8));
        e->cap_pvals = need_pv;
    }
    if (!e->d_pidx) CK(cudaMalloc((void**)&e->d_pidx, (size_t)np * 4));
    // optional arguments (include/mpcgpu.h: mpcgpu_set_options)
    const bool cons = opt && opt->prev_traj;
    const bool want_obj = opt && (opt->objective_out || opt->consistency_cost_out);
    double* mem_host = opt ? opt->mem_inout : nullptr;
    const int n_static = (opt && opt->static_halfspaces && ga) ? opt->n_static : 0;
    if (opt && (opt->n_static < 0 || (cons && (opt->ix < 0 || opt->ix >= nx || opt->iy < 0 || opt->iy >= nx)))) return MPCGPU_ERR_ARG;
    if (cons || want_obj) {
        if ((size_t)n_sets * N * 2 > e->cap_prev || !e->d_objout) {
            for (void* q : {(void*)e->d_prev, (void*)e->d_objout, (void*)e->d_consout, (void*)e->d_consen})
                if (q) cudaFree(q);
            e->d_prev = e->d_objout = e->d_consout = nullptr; e->d_consen = nullptr; e->cap_prev = 0;
            const size_t sets_cap = (size_t)(e->max_batch / planners + 1);
            CK(cudaMalloc((void**)&e->d_prev, (sets_cap > (size_t)n_sets ? sets_cap : (size_t)n_sets) * N * 2 * 8));
            CK(cudaMalloc((void**)&e->d_objout, (size_t)e->max_batch * 8));
            CK(cudaMalloc((void**)&e->d_consout, (size_t)e->max_batch * 8));
            CK(cudaMalloc((void**)&e->d_consen, (size_t)e->max_batch));
            e->cap_prev = (sets_cap > (size_t)n_sets ? sets_cap : (size_t)n_sets) * N * 2;
        }
    }
    const size_t st_per_set = (size_t)N * n_static * 3;
    if (n_static > 0 && (size_t)n_sets * st_per_set > e->cap_static) {
        if (e->d_static) cudaFree(e->d_static);
        e->d_static = nullptr; e->cap_static = 0;
        CK(cudaMalloc((void**)&e->d_static, (size_t)n_sets * st_per_set * 8));
        e->cap_static = (size_t)n_sets * st_per_set;
    }
    const size_t smem_ = (size_t)o->mem_doubles;
    const size_t ob_per_set = ga ? (size_t)N * ga->n_obs * ga->ob_stride : 0;       // doubles
    unsigned char* d_guided = nullptr;
    size_t tab_inv = 0, tab_stg = 0, tab_rad = 0;      // offsets (doubles) inside d_tab
    if (tab) {
        if (tab->n_invariant < 0 || tab->n_stage < 0 || (tab->n_invariant > 0 && (!tab->invariant_idx || !tab->invariant)) ||
            (tab->n_stage > 0 && (!tab->stage_idx || !tab->stage)) || tab->n_invariant > np || tab->n_stage > np)
            return MPCGPU_ERR_ARG;
        for (int i = 0; i < tab->n_invariant; i++) if (tab->invariant_idx[i] < 0 || tab->invariant_idx[i] >= np) return MPCGPU_ERR_ARG;
        for (int i = 0; i < tab->n_stage; i++) if (tab->stage_idx[i] < 0 || tab->stage_idx[i] >= np) return MPCGPU_ERR_ARG;
        tab_stg = (size_t)n_sets * tab->n_invariant;
        tab_rad = tab_stg + (size_t)n_sets * N * tab->n_stage;
        const size_t need = tab_rad + (size_t)n_sets * (tab->obstacle_radius ? tab->M : 0) + 1;
        if (need > e->cap_tab) {
            if (e->d_tab) cudaFree(e->d_tab);
            e->d_tab = nullptr; e->cap_tab = 0;
            CK(cudaMalloc((void**)&e->d_tab, need * 8));
            e->cap_tab = need;
        }
        if ((size_t)(tab->n_invariant + tab->n_stage + 1) > e->cap_tabidx) {
            if (e->d_tabidx) cudaFree(e->d_tabidx);
            e->d_tabidx = nullptr; e->cap_tabidx = 0;
            CK(cudaMalloc((void**)&e->d_tabidx, (size_t)(2 * np + 1) * 4));
            e->cap_tabidx = (size_t)(2 * np + 1);
        }
        if (tab->n_invariant) CK(cudaMemcpyAsync(e->d_tabidx, tab->invariant_idx, (size_t)tab->n_invariant * 4, cudaMemcpyHostToDevice, st));
        if (tab->n_stage) CK(cudaMemcpyAsync(e->d_tabidx + tab->n_invariant, tab->stage_idx, (size_t)tab->n_stage * 4, cudaMemcpyHostToDevice, st));
    }
    if (ga && (ga->guided || ob_per_set)) {      // obstacle predictions / tables + guided flags of the device-side constraint construction
        const size_t ob_bytes = (size_t)n_sets * ob_per_set * 8;
        if (ob_bytes + (size_t)n > e->cap_obst) {
            if (e->d_obst) cudaFree(e->d_obst);
            e->d_obst = nullptr; e->cap_obst = 0;
            CK(cudaMalloc((void**)&e->d_obst, ob_bytes + (size_t)n + 16));
            e->cap_obst = ob_bytes + (size_t)n;
        }
        d_guided = (unsigned char*)e->d_obst + ob_bytes;
    }
    // small data shared by every chunk
    if (nidx > 0) {
        if (nidx > np) return MPCGPU_ERR_ARG;
        CK(cudaMemcpyAsync(e->d_pidx, param_idx, (size_t)nidx * 4, cudaMemcpyHostToDevice, st));
    }
    std::vector<int> off((size_t)n_sets + 1);
    for (int s_ = 0; s_ <= n_sets; s_++) off[s_] = s_ * planners;
    CK(cudaMemcpyAsync(e->d_offsets, off.data(), (size_t)(n_sets + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));      // d_pidx / d_offsets are read by both streams; `off` may go out of scope

    // Chunks of whole homotopy sets on two streams: the copies of chunk c+1 overlap the kernels of chunk c (same scheme as
    // mpcgpu_solve_batch: a small first chunk, then up to four more of >= ~8192 problems).
    int bounds[mpcgpu_engine::MAX_CHUNKS + 1];
    int nchunk = 0;
    bounds[0] = 0;
    const int sets8k = (8192 + planners - 1) / planners;
    // Separate solve launches per chunk cost ~6 % (SMs idle while the last warps of a chunk finish; a gate as in
    // mpcgpu_solve_batch is no option here: the expansion kernels of a later chunk could not run beside the persistent grid).
    // When the compact inputs are small -- under ~8 KB per problem, i.e. their copy takes less than that -- everything is
    // copied and expanded first and ONE launch solves the batch.
    size_t in_bytes = (size_t)nz * (N + 1) * 8 + (size_t)nidx * N * 8 + (mem_host ? smem_ * 8 : 0) +
                      ((size_t)nx * 8 + (size_t)ob_per_set * 8 + (size_t)st_per_set * 8 + (cons ? (size_t)N * 16 : 0) +
                       (tab ? ((size_t)tab->n_invariant + (size_t)N * tab->n_stage + (size_t)tab->M) * 8 : (size_t)N * np * 8)) / planners;
    if (in_bytes <= 8192) {
        bounds[++nchunk] = n_sets;
    } else if (n_sets >= 4 * sets8k) {
        int first = n_sets / 16;
        if (first * planners < 2048) first = (2048 + planners - 1) / planners;
        bounds[++nchunk] = first;
    }
    if (bounds[nchunk] < n_sets) {
        const int rest = n_sets - bounds[nchunk];
        int k = rest / sets8k;
        if (k < 1) k = 1;
        if (k > 4) k = 4;
        const int per_ = (rest + k - 1) / k;
        for (int i = 0; i < k; i++) {
            const int hi = bounds[nchunk] + per_;
            bounds[nchunk + 1] = hi < n_sets ? hi : n_sets;
            nchunk++;
        }
    }
    // results: copied behind the chunk's kernels when the destination is page-locked; a copy into PAGEABLE memory would block
    // the host until the chunk is solved and so serialise the chunks -- those are made at the end
    struct Pending { void* dst; const void* src; size_t bytes; };
    std::vector<Pending> pending;
    auto out_copy = [&](void* dst, const void* src, size_t bytes, cudaStream_t cs) -> cudaError_t {
        cudaPointerAttributes a;
        const bool pinned = nchunk == 1 || (cudaPointerGetAttributes(&a, dst) == cudaSuccess && a.type == cudaMemoryTypeHost);
        if (pinned) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, cs);
        cudaGetLastError();
        pending.push_back({dst, src, bytes});
        return cudaSuccess;
    };
    for (int c = 0; c < nchunk; c++) {
        const int s0 = bounds[c], ns = bounds[c + 1] - s0;
        if (ns <= 0) { nchunk = c; break; }
        const size_t p0 = (size_t)s0 * planners, m = (size_t)ns * planners;          // first problem, problems
        cudaStream_t cs = (c & 1) ? e->stream2 : e->stream;
        int* counter = e->next_counter();
        double* d_xs = e->d_xs + (size_t)s0 * nx;
        double* d_sh = e->d_shared + (size_t)s0 * N * np;
        CK(cudaMemcpyAsync(d_xs, xinit_sets + (size_t)s0 * nx, (size_t)ns * nx * 8, cudaMemcpyHostToDevice, cs));
        if (tab) {      // build the per-set block on the device: zero, invariant values broadcast over the stages, per-stage values
            CK(cudaMemsetAsync(d_sh, 0, (size_t)ns * N * np * 8, cs));
            if (tab->n_invariant) {
                double* dv = e->d_tab + tab_inv + (size_t)s0 * tab->n_invariant;
                CK(cudaMemcpyAsync(dv, tab->invariant + (size_t)s0 * tab->n_invariant, (size_t)ns * tab->n_invariant * 8, cudaMemcpyHostToDevice, cs));
                scatter_invariant_kernel<<<296, 256, 0, cs>>>(ns, N, np, tab->n_invariant, e->d_tabidx, dv, d_sh);
                e->launches += 1;
            }
            if (tab->n_stage) {
                double* dv = e->d_tab + tab_stg + (size_t)s0 * N * tab->n_stage;
                CK(cudaMemcpyAsync(dv, tab->stage + (size_t)s0 * N * tab->n_stage, (size_t)ns * N * tab->n_stage * 8, cudaMemcpyHostToDevice, cs));
                scatter_stage_kernel<<<296, 256, 0, cs>>>(ns, N, np, tab->n_stage, e->d_tabidx + tab->n_invariant, dv, d_sh);
                e->launches += 1;
            }
            if (tab->obstacle_radius)
                CK(cudaMemcpyAsync(e->d_tab + tab_rad + (size_t)s0 * tab->M, tab->obstacle_radius + (size_t)s0 * tab->M, (size_t)ns * tab->M * 8,
                                   cudaMemcpyHostToDevice, cs));
            CK(cudaGetLastError());
        } else
            CK(cudaMemcpyAsync(d_sh, shared_params + (size_t)s0 * N * np, (size_t)ns * N * np * 8, cudaMemcpyHostToDevice, cs));
        CK(cudaMemcpyAsync(e->d_x0 + p0 * nz * (N + 1), x0 + p0 * nz * (N + 1), m * nz * (N + 1) * 8, cudaMemcpyHostToDevice, cs));
        if (nidx > 0)
            CK(cudaMemcpyAsync(e->d_pvals + p0 * N * nidx, planner_params + p0 * N * nidx, m * N * nidx * 8, cudaMemcpyHostToDevice, cs));
        if (num_iter) CK(cudaMemcpyAsync(e->d_num_iter + p0, num_iter + p0, m * 4, cudaMemcpyHostToDevice, cs));
        if (mem_host) {      // the planners' persistent capsules; `*solver = *_solver` keeps the multipliers, resets the QP memory
            CK(cudaMemcpyAsync(e->d_mem + p0 * smem_, mem_host + p0 * smem_, m * smem_ * 8, cudaMemcpyHostToDevice, cs));
            mem_flag_downgrade_kernel<<<(int)((m + 127) / 128), 128, 0, cs>>>((int)m, (int)smem_, e->d_mem + p0 * smem_);
            e->launches += 1;
        }
        if (cons) {
            CK(cudaMemcpyAsync(e->d_prev + (size_t)s0 * N * 2, opt->prev_traj + (size_t)s0 * N * 2, (size_t)ns * N * 2 * 8, cudaMemcpyHostToDevice, cs));
            if (opt->consistency_enabled) CK(cudaMemcpyAsync(e->d_consen + p0, opt->consistency_enabled + p0, m, cudaMemcpyHostToDevice, cs));
        }
        if (n_static > 0)
            CK(cudaMemcpyAsync(e->d_static + (size_t)s0 * st_per_set, opt->static_halfspaces + (size_t)s0 * st_per_set, (size_t)ns * st_per_set * 8,
                               cudaMemcpyHostToDevice, cs));
        const double* d_ob = nullptr;
        if (ga) {
            d_ob = (const double*)e->d_obst + (size_t)s0 * ob_per_set;
            if (ob_per_set) CK(cudaMemcpyAsync((void*)d_ob, ga->obst_pred + (size_t)s0 * ob_per_set, (size_t)ns * ob_per_set * 8, cudaMemcpyHostToDevice, cs));
            if (ga->guided) CK(cudaMemcpyAsync(d_guided + p0, ga->guided + p0, m, cudaMemcpyHostToDevice, cs));
            if (ga->ell_base >= 0) {      // ellipsoid slots of the shared block from the tables, before it is expanded per planner
                int rc_ = pack_obstacles_launch(e, ns, d_xs, d_ob, ga->ob_stride, (tab && tab->obstacle_radius) ? e->d_tab + tab_rad + (size_t)s0 * tab->M : nullptr,
                                                ga->n_obs, ga->ell_base, ga->ell_stride, ga->ell_off, d_sh, cs);
                if (rc_ != MPCGPU_OK) return rc_;
            }
        }
        double* d_par = e->d_params + p0 * N * np;
        repeat_xinit_kernel<<<(int)((m * nx + 255) / 256), 256, 0, cs>>>((int)m, planners, nx, d_xs, e->d_xinit + p0 * nx);
        expand_params_kernel<<<1184, 256, 0, cs>>>((int)m, planners, N, np, nidx, d_sh, e->d_pidx, e->d_pvals + p0 * N * nidx, d_par);
        if (nidx > 0) scatter_params_kernel<<<592, 256, 0, cs>>>((int)m, N, np, nidx, e->d_pidx, e->d_pvals + p0 * N * nidx, d_par);
        CK(cudaGetLastError());
        e->launches += (nidx > 0) ? 3 : 2;
        if (ga && ga->guided) {      // guidance halfspaces built on the device from the obstacle predictions and the warm starts
            int rc_ = guidance_halfspaces_launch(e, ns, planners, d_xs, e->d_x0 + p0 * nz * (N + 1), d_ob, ga->n_obs, ga->ob_stride, d_guided + p0,
                                                 ga->lin_base, ga->lin_count, ga->robot_radius,
                                                 n_static > 0 ? e->d_static + (size_t)s0 * st_per_set : nullptr, n_static, d_par, cs);
            if (rc_ != MPCGPU_OK) return rc_;
        }
        int rc = launch_solve_on(e, cs, counter, e->cev0[c], e->cev1[c], (int)m, e->d_xinit + p0 * nx, e->d_x0 + p0 * nz * (N + 1), d_par,
                                 num_iter ? e->d_num_iter + p0 : nullptr, num_iter_all, mem_host ? e->d_mem + p0 * smem_ : nullptr,
                                 e->d_xtraj + p0 * nx * (N + 1),
                                 e->d_utraj + p0 * nu * N, e->d_pobj + p0, e->d_exit + p0, e->d_qps + p0, e->d_res_eq + p0, e->d_ipm + p0);
        if (rc != MPCGPU_OK) return rc;
        // selection on the device, then everything back
        if (obj_scale) CK(cudaMemcpyAsync(e->d_scale + p0, obj_scale + p0, m * 8, cudaMemcpyHostToDevice, cs));
        if (obj_sub) CK(cudaMemcpyAsync(e->d_sub + p0, obj_sub + p0, m * 8, cudaMemcpyHostToDevice, cs));
        if (disabled) CK(cudaMemcpyAsync(e->d_disabled + p0, disabled + p0, m, cudaMemcpyHostToDevice, cs));
        // set offsets are absolute problem indices: the chunk passes the offset table from s0 on with the full arrays
        {
            ConsArgs ca;
            ca.xtraj = e->d_xtraj; ca.N = N; ca.nx = nx;
            if (cons) {
                ca.prev = e->d_prev + (size_t)s0 * N * 2; ca.enabled = opt->consistency_enabled ? e->d_consen : nullptr;
                ca.weight = opt->consistency_weight; ca.ix = opt->ix; ca.iy = opt->iy;
            }
            if (want_obj) { ca.obj_out = e->d_objout; ca.cons_out = e->d_consout; }
            select_best_kernel<<<(ns + 127) / 128, 128, 0, cs>>>(ns, e->d_offsets + s0, e->d_pobj, e->d_exit, obj_scale ? e->d_scale : nullptr,
                                                                  obj_sub ? e->d_sub : nullptr, disabled ? e->d_disabled : nullptr, e->d_best + s0, ca);
            CK(cudaGetLastError());
            e->launches += 1;
        }
        if (opt && opt->objective_out) CK(out_copy(opt->objective_out + p0, e->d_objout + p0, m * 8, cs));
        if (opt && opt->consistency_cost_out) CK(out_copy(opt->consistency_cost_out + p0, e->d_consout + p0, m * 8, cs));
        if (mem_host) CK(out_copy(mem_host + p0 * smem_, e->d_mem + p0 * smem_, m * smem_ * 8, cs));
        if (xtraj) CK(out_copy(xtraj + p0 * nx * (N + 1), e->d_xtraj + p0 * nx * (N + 1), m * nx * (N + 1) * 8, cs));
        if (utraj) CK(out_copy(utraj + p0 * nu * N, e->d_utraj + p0 * nu * N, m * nu * N * 8, cs));
        if (opt && opt->best_xtraj && opt->best_utraj) {
            // the input staging of this chunk is free once its solve has run: the per-set trajectories are gathered into it
            double* bx = e->d_x0 + p0 * nz * (N + 1);
            double* bu = bx + (size_t)ns * nx * (N + 1);
            gather_best_kernel<<<ns, 128, 0, cs>>>(ns, e->d_offsets + s0, e->d_best + s0, nx * (N + 1), nu * N, e->d_xtraj, e->d_utraj, bx, bu);
            CK(cudaGetLastError());
            e->launches += 1;
            CK(out_copy(opt->best_xtraj + (size_t)s0 * nx * (N + 1), bx, (size_t)ns * nx * (N + 1) * 8, cs));
            CK(out_copy(opt->best_utraj + (size_t)s0 * nu * N, bu, (size_t)ns * nu * N * 8, cs));
        }
        CK(out_copy(pobj + p0, e->d_pobj + p0, m * 8, cs));
        CK(out_copy(exit_code + p0, e->d_exit + p0, m * 4, cs));
        CK(out_copy(qp_status + p0, e->d_qps + p0, m * 4, cs));
        CK(out_copy(res_eq + p0, e->d_res_eq + p0, m * 8, cs));
        CK(out_copy(best_idx + s0, e->d_best + s0, (size_t)ns * 4, cs));
    }
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaStreamSynchronize(e->stream2));
    for (const Pending& pc : pending) CK(cudaMemcpy(pc.dst, pc.src, pc.bytes, cudaMemcpyDeviceToHost));
    e->chunks_timed = nchunk;
    return MPCGPU_OK;
}

int mpcgpu_solve_sets_tables(mpcgpu_engine* e, int n_sets, int planners, const double* xinit_sets, const mpcgpu_param_tables* tab,
                             const double* x0, int nidx, const int* param_idx, const double* planner_params, const int* num_iter,
                             int num_iter_all, double* xtraj, double* utraj, double* pobj, int* exit_code, int* qp_status, double* res_eq,
                             const double* obj_scale, const double* obj_sub, const unsigned char* disabled, int* best_idx,
                             const mpcgpu_set_options* opt)
{
    if (!e || !tab || tab->M < 0 || (tab->M > 0 && (!tab->obstacles || (tab->ob_stride != 2 && tab->ob_stride != 4))) ||
        (tab->M > 0 && tab->ob_stride == 2 && tab->ell_base >= 0 && !tab->obstacle_radius))
        return MPCGPU_ERR_ARG;
    const int np = e->ops->np;
    if (tab->ell_base >= 0) {
        if (!tab->ell_offsets || tab->ell_stride <= 0 || tab->ell_base + tab->M * tab->ell_stride > np) return MPCGPU_ERR_ARG;
        for (int i = 0; i < 7; i++)
            if (tab->ell_offsets[i] < 0 || tab->ell_offsets[i] >= tab->ell_stride) return MPCGPU_ERR_ARG;
    }
    if (tab->guided && (tab->lin_base < 0 || tab->lin_count < 0 || tab->lin_base + 3 * tab->lin_count > np)) return MPCGPU_ERR_ARG;
    GuidedArgs ga = {tab->M, tab->lin_base, tab->lin_count, tab->obstacles, tab->guided, tab->robot_radius};
    ga.ob_stride = tab->M > 0 ? tab->ob_stride : 2;
    ga.ell_base = tab->M > 0 ? tab->ell_base : -1; ga.ell_stride = tab->ell_stride;
    if (ga.ell_base >= 0)
        for (int i = 0; i < 7; i++) ga.ell_off[i] = tab->ell_offsets[i];
    ga.tab = tab;
    return drain(e, solve_sets_impl(e, n_sets, planners, xinit_sets, nullptr, x0, nidx, param_idx, planner_params, &ga, num_iter, num_iter_all, xtraj,
                                    utraj, pobj, exit_code, qp_status, res_eq, obj_scale, obj_sub, disabled, best_idx, opt));
}

int mpcgpu_select_best(mpcgpu_engine* e, int n_sets, const int* set_offsets, const double* pobj, const int* exit_code,
                       const double* obj_scale, const double* obj_sub, const unsigned char* disabled, int* best_idx)
{
    if (!e || n_sets < 0 || !set_offsets || !pobj || !exit_code || !best_idx) return MPCGPU_ERR_ARG;
    if (n_sets == 0) return MPCGPU_OK;
    const int n = set_offsets[n_sets];
    if (n > e->max_batch || n_sets > e->max_batch) return MPCGPU_ERR_ARG;
    CK(cudaSetDevice(e->device));
    cudaStream_t st = e->stream;
    CK(cudaMemcpyAsync(e->d_offsets, set_offsets, (size_t)(n_sets + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->d_pobj, pobj, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(e->d_exit, exit_code, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    if (obj_scale) CK(cudaMemcpyAsync(e->d_scale, obj_scale, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    if (obj_sub) CK(cudaMemcpyAsync(e->d_sub, obj_sub, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    if (disabled) CK(cudaMemcpyAsync(e->d_disabled, disabled, (size_t)n, cudaMemcpyHostToDevice, st));
    int rc = mpcgpu_select_best_device(e, n_sets, e->d_offsets, e->d_pobj, e->d_exit, obj_scale ? e->d_scale : nullptr,
                                       obj_sub ? e->d_sub : nullptr, disabled ? e->d_disabled : nullptr, e->d_best, st);
    if (rc != MPCGPU_OK) return rc;
    CK(cudaMemcpyAsync(best_idx, e->d_best, (size_t)n_sets * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return MPCGPU_OK;
}

int mpcgpu_model_eval_doubles(const mpcgpu_engine* e, int* nhs) { if (!e) return MPCGPU_ERR_ARG; if (nhs) *nhs = e->ops->nhs; return e->ops->model_eval_doubles; }

int mpcgpu_model_eval(mpcgpu_engine* e, int n, const double* z, const double* p, const double* pi, const double* mh, double* out)
{
    if (!e || n <= 0 || !z || !p || !pi || !mh || !out) return MPCGPU_ERR_ARG;
    CK(cudaSetDevice(e->device));
    const MpcConfigOps* o = e->ops;
    const size_t sz[5] = {(size_t)n * (o->nx + o->nu) * 8, (size_t)n * o->np * 8, (size_t)n * o->nx * 8,
                          (size_t)n * (o->nh > 0 ? o->nh : 1) * 8, (size_t)n * o->model_eval_doubles * 8};
    const void* src[4] = {z, p, pi, mh};
    void* d[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    int rc = MPCGPU_OK;
    for (int i = 0; i < 5 && rc == MPCGPU_OK; i++)
        if (cudaMalloc(&d[i], sz[i]) != cudaSuccess) rc = MPCGPU_ERR_CUDA;
    for (int i = 0; i < 4 && rc == MPCGPU_OK; i++)
        if (cudaMemcpy(d[i], src[i], i == 3 ? (size_t)n * o->nh * 8 : sz[i], cudaMemcpyHostToDevice) != cudaSuccess) rc = MPCGPU_ERR_CUDA;
    if (rc == MPCGPU_OK) {
        if (o->launch_model_eval(e->stream, n, (double*)d[0], (double*)d[1], (double*)d[2], (double*)d[3], (double*)d[4]) != cudaSuccess ||
            cudaStreamSynchronize(e->stream) != cudaSuccess || cudaMemcpy(out, d[4], sz[4], cudaMemcpyDeviceToHost) != cudaSuccess)
            rc = MPCGPU_ERR_CUDA;
    }
    if (rc != MPCGPU_OK) e->err = std::string("mpcgpu_model_eval: ") + cudaGetErrorString(cudaGetLastError());
    for (void* q : d)
        if (q) cudaFree(q);
    return rc;
}

int mpcgpu_measure_fp64_peak(int device, double* tflops)
{
    if (!tflops) return MPCGPU_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return MPCGPU_ERR_NO_DEVICE;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    double* out = nullptr;
    if (cudaMalloc((void**)&out, (size_t)blocks * threads * 8) != cudaSuccess) return MPCGPU_ERR_CUDA;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0);
        dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0 + rep);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return MPCGPU_ERR_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * (double)iters * (double)blocks * (double)threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return MPCGPU_OK;
}

int mpcgpu_check_report(mpcgpu_engine* e, unsigned long long* counters8, int reset)
{
    if (!e || !counters8) return MPCGPU_ERR_ARG;
    if (!e->ops->check_report) return MPCGPU_ERR_ARG;      // not a diagnostic (-DMPC_CHECK=1) build
    CK(cudaSetDevice(e->device));
    CK(e->ops->check_report(counters8, reset));
    return MPCGPU_OK;
}
int mpcgpu_check_selftest(mpcgpu_engine* e)
{
    if (!e || !e->ops->check_selftest) return MPCGPU_ERR_ARG;
    CK(cudaSetDevice(e->device));
    CK(e->ops->check_selftest(e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return MPCGPU_OK;
}

int mpcgpu_alloc_pinned(size_t bytes, void** out)
{
    if (!out || bytes == 0) return MPCGPU_ERR_ARG;
    *out = nullptr;
    return cudaHostAlloc(out, bytes, cudaHostAllocPortable) == cudaSuccess ? MPCGPU_OK : MPCGPU_ERR_CUDA;
}
int mpcgpu_free_pinned(void* p) { return (!p || cudaFreeHost(p) == cudaSuccess) ? MPCGPU_OK : MPCGPU_ERR_CUDA; }

int mpcgpu_set_kernel_mode(mpcgpu_engine* e, int mode)
{
    if (!e || mode < MPCGPU_KERNEL_AUTO || mode > MPCGPU_KERNEL_SPLIT) return MPCGPU_ERR_ARG;
    e->kernel_mode = mode;
    return e->ops->has_split ? 1 : 0;
}

long long mpcgpu_launch_count(const mpcgpu_engine* e) { return e ? e->launches : 0; }

float mpcgpu_last_kernel_ms(mpcgpu_engine* e)
{
    if (!e) return -1.0f;
    float ms = -1.0f;
    if (e->chunks_timed > 0) {      // host call: sum of the chunk kernels
        float tot = 0.0f;
        for (int c = 0; c < e->chunks_timed; c++) {
            if (cudaEventSynchronize(e->cev1[c]) != cudaSuccess || cudaEventElapsedTime(&ms, e->cev0[c], e->cev1[c]) != cudaSuccess) return -1.0f;
            tot += ms;
        }
        return tot;
    }
    if (cudaEventSynchronize(e->ev1) != cudaSuccess) return -1.0f;
    if (cudaEventElapsedTime(&ms, e->ev0, e->ev1) != cudaSuccess) return -1.0f;
    return ms;
}

const char* mpcgpu_last_error(const mpcgpu_engine* e) { return e ? e->err.c_str() : "null engine"; }

}  // extern "C"

#include "mpcgpu_wire.inl"
