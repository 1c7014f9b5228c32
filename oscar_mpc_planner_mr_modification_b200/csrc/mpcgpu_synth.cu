// Device-side generator of the synthetic workload ("counter-based Philox keyed by problem index so host and device
// generate identical data").  Measurement infrastructure, not part of the reference's interface: it fills the reference's own
// AcadosParameters layouts (xinit, x0, all_parameters) for whole homotopy sets, with the value conventions of the reference's
// C++ modules -- the same scenario definition as synthetic.make_batch, see there for the file:line of every convention.
// The host mirror is synthetic.make_batch_philox (numpy).  This file is compiled with -fmad=false: every operation the mirror
// performs is one IEEE operation here too, so the two agree bit for bit except where sin / cos / atan2 of the two math
// libraries round differently (obstacle headings, the heading of the guidance polyline, the braking roll-out).
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/mpcgpu.h"

namespace {

// Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0, k1)
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// j-th uniform double in [0, 1) of homotopy set `gs` (53 bits: 27 from one word, 26 from the next)
__device__ double draw(unsigned long long seed, unsigned long long gs, int j)
{
    uint32_t r[4];
    philox4x32_10((uint32_t)gs, (uint32_t)(gs >> 32), (uint32_t)(j >> 1), 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const uint32_t a = r[2 * (j & 1)], b = r[2 * (j & 1) + 1];
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double uni(double lo, double hi, double u) { return lo + (hi - lo) * u; }

constexpr int NWP = 7;          // waypoints: num_segments + 2
constexpr int MAXM = MPCGPU_SYNTH_MAX_OBST;
constexpr int MAXK = 64;        // N + 1 <= 64

// natural cubic spline through (t_i, y_i), i < NWP: Thomas algorithm; coefficients of segment i: a s^3 + b s^2 + c s + d
__device__ void natural_cubic(const double* t, const double* y, double* ca, double* cb, double* cc, double* cd)
{
    double h[NWP - 1], c[NWP], cp[NWP], dp[NWP];
    for (int i = 0; i < NWP - 1; i++) h[i] = t[i + 1] - t[i];
    // rows 1..NWP-2: h[i-1] c[i-1] + 2 (h[i-1] + h[i]) c[i] + h[i] c[i+1] = r_i;  c[0] = c[NWP-1] = 0
    cp[0] = 0.0; dp[0] = 0.0;
    for (int i = 1; i < NWP - 1; i++) {
        const double r = 3.0 * ((y[i + 1] - y[i]) / h[i] - (y[i] - y[i - 1]) / h[i - 1]);
        const double den = 2.0 * (h[i - 1] + h[i]) - h[i - 1] * cp[i - 1];
        cp[i] = h[i] / den;
        dp[i] = (r - h[i - 1] * dp[i - 1]) / den;
    }
    c[NWP - 1] = 0.0;
    for (int i = NWP - 2; i >= 1; i--) c[i] = dp[i] - cp[i] * c[i + 1];
    c[0] = 0.0;
    for (int i = 0; i < NWP - 1; i++) {
        cb[i] = c[i];
        cc[i] = (y[i + 1] - y[i]) / h[i] - h[i] * (2.0 * c[i] + c[i + 1]) / 3.0;
        ca[i] = (c[i + 1] - c[i]) / (3.0 * h[i]);
        cd[i] = y[i];
    }
}

// One CTA of 64 threads per homotopy set; thread k owns stage k where the work is stage-parallel.
__global__ void __launch_bounds__(64) synth_kernel(const mpcgpu_synth_layout L, unsigned long long seed, long long first_set, int n_sets,
                                                   int planners, double* __restrict__ xinit, double* __restrict__ x0,
                                                   double* __restrict__ params, double* __restrict__ obst_pred)
{
    __shared__ double st[5], sx[5][NWP - 1], sy[5][NWP - 1], ts[NWP], op0[MAXM][2], ovel[MAXM][2];
    __shared__ double pos[MAXK][2], vel[MAXK][2], X[MAXK][8];
    const int N = L.N, nx = L.nx, nu = L.nu, nz = nx + nu, npar = L.npar, M = L.n_obst, k = threadIdx.x;
    const double dt = L.dt;
    for (int s = blockIdx.x; s < n_sets; s += gridDim.x) {
        const unsigned long long gs = (unsigned long long)(first_set + s);
        __syncthreads();
        // ---- per-set scenario: state, path, obstacles
        if (k == 0) {
            st[0] = uni(-0.5, 0.5, draw(seed, gs, 0)); st[1] = uni(-0.5, 0.5, draw(seed, gs, 1));
            st[2] = uni(-0.3, 0.3, draw(seed, gs, 2)); st[3] = uni(0.0, 2.5, draw(seed, gs, 3)); st[4] = 0.0;
            double wx[NWP], wy[NWP];
            wx[0] = 0.0; wy[0] = 0.0; ts[0] = 0.0;
            for (int i = 1; i < NWP; i++) {
                wx[i] = wx[i - 1] + uni(3.0, 6.0, draw(seed, gs, 3 + i));
                wy[i] = uni(-2.0, 2.0, draw(seed, gs, 9 + i));
            }
            for (int i = 1; i < NWP; i++) {
                const double dx = wx[i] - wx[i - 1], dy = wy[i] - wy[i - 1];
                ts[i] = ts[i - 1] + sqrt(dx * dx + dy * dy);
            }
            natural_cubic(ts, wx, sx[0], sx[1], sx[2], sx[3]);
            natural_cubic(ts, wy, sy[0], sy[1], sy[2], sy[3]);
            sx[4][0] = wx[3]; sy[4][0] = wy[3];          // third waypoint (GoalModule configurations)
        }
        if (k < M) {
            op0[k][0] = uni(2.0, 20.0, draw(seed, gs, 16 + 4 * k)); op0[k][1] = uni(-4.0, 4.0, draw(seed, gs, 17 + 4 * k));
            const double sp = uni(0.5, 1.5, draw(seed, gs, 18 + 4 * k)), hd = uni(-L.pi, L.pi, draw(seed, gs, 19 + 4 * k));
            ovel[k][0] = sp * cos(hd); ovel[k][1] = sp * sin(hd);
        }
        __syncthreads();
        // obstacle prediction i of obstacle j: p0 + (v dt) i  (data_preparation.cpp:74-75)
        auto opred = [&](int i, int j, int c) { return op0[j][c] + ovel[j][c] * dt * (double)i; };
        if (obst_pred)
            for (int t = k; t < N * M * 2; t += 64) obst_pred[(size_t)s * N * M * 2 + t] = opred(t / (2 * M), (t / 2) % M, t & 1);
        if (k < nx) for (int h = 0; h < planners; h++) xinit[((size_t)s * planners + h) * nx + k] = st[k];

        for (int h = 0; h < planners; h++) {
            const size_t prob = (size_t)s * planners + h;
            const bool nonguided = L.guided && planners > 1 && h == planners - 1;
            const bool follow = L.guided && !nonguided;
            __syncthreads();
            if (k <= N) for (int i = 0; i < 8; i++) X[k][i] = 0.0;
            __syncthreads();
            // ---- warm start
            if (follow) {
                if (k <= N) {
                    double px = st[0] + L.vg * dt * (double)k, py = st[1] + L.lateral[h % 8] * L.lat_profile[k];
                    for (int it = 0; it < 3; it++)
                        for (int j = 0; j < M; j++) {                 // stand-in for projectToSafety: push out of the discs
                            const int i = k >= 1 ? k - 1 : 0;
                            const double dx = px - opred(i, j, 0), dy = py - opred(i, j, 1);
                            const double dist = sqrt(dx * dx + dy * dy);
                            const double push = dist < L.need ? (L.need - dist) / fmax(dist, 1e-9) : 0.0;
                            px = px + dx * push; py = py + dy * push;
                        }
                    if (k == 0) { px = st[0]; py = st[1]; }
                    pos[k][0] = px; pos[k][1] = py;
                }
                __syncthreads();
                if (k <= N) {                                          // np.gradient: central inside, one-sided at the ends
                    for (int c = 0; c < 2; c++)
                        vel[k][c] = k == 0 ? (pos[1][c] - pos[0][c]) / dt
                                           : (k == N ? (pos[N][c] - pos[N - 1][c]) / dt : (pos[k + 1][c] - pos[k - 1][c]) / (2.0 * dt));
                    X[k][nu + 0] = pos[k][0]; X[k][nu + 1] = pos[k][1];
                    X[k][nu + 2] = atan2(vel[k][1], vel[k][0]);
                    X[k][nu + 3] = sqrt(vel[k][0] * vel[k][0] + vel[k][1] * vel[k][1]);
                }
                __syncthreads();
                if (k == 0) {
                    if (nx > 4) {
                        double sacc = 0.0;
                        for (int q = 1; q <= N; q++) { sacc = sacc + X[q - 1][nu + 3] * dt; X[q][nu + 4] = sacc; }
                    }
                    for (int i = 0; i < nx; i++) X[0][nu + i] = st[i];
                }
            } else if (k == 0) {                                       // braking roll-out (non-guided planner) / cruise
                const double a = nonguided ? -L.deceleration : 0.0;
                double x = st[0], y = st[1], psi = st[2], v = st[3], sp = st[4];
                for (int q = 0; q <= N; q++) {
                    X[q][0] = a;
                    X[q][nu + 0] = x; X[q][nu + 1] = y; X[q][nu + 2] = psi; X[q][nu + 3] = v;
                    if (nx > 4) X[q][nu + 4] = sp;
                    x = x + v * dt * cos(psi);
                    y = y + v * dt * sin(psi);
                    sp = sp + v * dt;
                    v = fmax(v + a * dt, 0.0);
                }
            }
            __syncthreads();
            for (int t = k; t < (N + 1) * nz; t += 64) x0[prob * (size_t)(N + 1) * nz + t] = X[t / nz][t % nz];

            // ---- parameters of the problem: zero, then the values of every stage (thread k: stage k)
            double* P = params + prob * (size_t)N * npar;
            for (int t = k; t < N * npar; t += 64) P[t] = 0.0;
            __syncthreads();
            if (k < N) {
                double* p = P + (size_t)k * npar;
                auto set = [&](int idx, double v) { if (idx >= 0) p[idx] = v; };
                for (int i = 0; i < 9; i++) set(L.weights[i], L.weight_values[i]);
                for (int i = 0; i < 5; i++) {
                    for (int c = 0; c < 4; c++) { set(L.spline[i][c], sx[c][i]); set(L.spline[i][4 + c], sy[c][i]); }
                    set(L.spline[i][8], ts[i]);
                }
                set(L.ego_disc_radius, L.robot_radius);
                set(L.ego_disc_0_offset, 0.0);
                set(L.goal[0], 1.0); set(L.goal[1], sx[4][0]); set(L.goal[2], sy[4][0]);
                set(L.prev_traj_x, X[k][nu + 0]);
                set(L.prev_traj_y, X[k][nu + 1]);
                for (int j = 0; j < M; j++) {                          // ellipsoid_constraints.cpp:34-90: stage k <- prediction k-1
                    set(L.obst[j][0], k >= 1 ? opred(k - 1, j, 0) : st[0] + 50.0);
                    set(L.obst[j][1], k >= 1 ? opred(k - 1, j, 1) : st[1] + 50.0);
                    set(L.obst[j][3], k >= 1 ? L.obstacle_radius : 0.1);
                    set(L.obst[j][6], 1.0);
                }
                for (int j = 0; j < L.n_lin; j++) {                    // linearized_constraints.cpp:49-189 (guidance halfspaces)
                    double a1 = 1.0, a2 = 0.0, b = st[0] + 100.0;
                    if (follow && k >= 1 && j < M) {
                        const double ox = opred(k - 1, j, 0), oy = opred(k - 1, j, 1);
                        const double dx = ox - X[k][nu + 0], dy = oy - X[k][nu + 1];
                        const double dist = fmax(sqrt(dx * dx + dy * dy), 1e-9);
                        a1 = dx / dist; a2 = dy / dist;
                        b = a1 * ox + a2 * oy - L.lin_margin;
                    }
                    set(L.lin[j][0], a1); set(L.lin[j][1], a2); set(L.lin[j][2], b);
                }
            }
        }
    }
}

}   // namespace

extern "C" int mpcgpu_generate_synthetic_device(int device, const mpcgpu_synth_layout* layout, unsigned long long seed, long long first_set,
                                                int n_sets, int planners, double* d_xinit, double* d_x0, double* d_params,
                                                double* d_obst_pred, void* stream)
{
    if (!layout || n_sets < 0 || planners < 1 || first_set < 0 || !d_xinit || !d_x0 || !d_params) return MPCGPU_ERR_ARG;
    if (layout->N < 2 || layout->N + 1 > MAXK || layout->nx < 4 || layout->nx > 5 || layout->nu != 2 || layout->n_obst < 0 ||
        layout->n_obst > MAXM || layout->n_lin < 0 || layout->n_lin > MAXM || layout->npar < 1)
        return MPCGPU_ERR_ARG;
    if (n_sets == 0) return MPCGPU_OK;
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) return MPCGPU_ERR_NO_DEVICE;
    if (cur != device && cudaSetDevice(device) != cudaSuccess) return MPCGPU_ERR_CUDA;
    const int grid = n_sets < 148 * 16 ? n_sets : 148 * 16;
    synth_kernel<<<grid, 64, 0, (cudaStream_t)stream>>>(*layout, seed, first_set, n_sets, planners, d_xinit, d_x0, d_params, d_obst_pred);
    const cudaError_t err = cudaGetLastError();
    if (cur != device) cudaSetDevice(cur);
    return err == cudaSuccess ? MPCGPU_OK : MPCGPU_ERR_CUDA;
}

// Same data into HOST arrays (temporary device buffers; for tests and for the host-side copy of a benchmark batch).
extern "C" int mpcgpu_generate_synthetic(int device, const mpcgpu_synth_layout* layout, unsigned long long seed, long long first_set,
                                         int n_sets, int planners, double* xinit, double* x0, double* params, double* obst_pred)
{
    if (!layout || n_sets < 0 || planners < 1 || !xinit || !x0 || !params) return MPCGPU_ERR_ARG;
    if (n_sets == 0) return MPCGPU_OK;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return MPCGPU_ERR_NO_DEVICE;
    if (cudaSetDevice(device) != cudaSuccess) return MPCGPU_ERR_CUDA;
    const size_t B = (size_t)n_sets * planners, nz = layout->nx + layout->nu;
    const size_t bx = B * layout->nx * 8, b0 = B * (layout->N + 1) * nz * 8, bp = B * layout->N * (size_t)layout->npar * 8,
                 bo = obst_pred ? (size_t)n_sets * layout->N * layout->n_obst * 2 * 8 : 0;
    double *dx = nullptr, *d0 = nullptr, *dp = nullptr, *dob = nullptr;
    int rc = MPCGPU_ERR_CUDA;
    if (cudaMalloc((void**)&dx, bx) == cudaSuccess && cudaMalloc((void**)&d0, b0) == cudaSuccess && cudaMalloc((void**)&dp, bp) == cudaSuccess &&
        (bo == 0 || cudaMalloc((void**)&dob, bo) == cudaSuccess)) {
        rc = mpcgpu_generate_synthetic_device(device, layout, seed, first_set, n_sets, planners, dx, d0, dp, dob, nullptr);
        if (rc == MPCGPU_OK &&
            (cudaMemcpy(xinit, dx, bx, cudaMemcpyDeviceToHost) != cudaSuccess || cudaMemcpy(x0, d0, b0, cudaMemcpyDeviceToHost) != cudaSuccess ||
             cudaMemcpy(params, dp, bp, cudaMemcpyDeviceToHost) != cudaSuccess ||
             (bo && cudaMemcpy(obst_pred, dob, bo, cudaMemcpyDeviceToHost) != cudaSuccess)))
            rc = MPCGPU_ERR_CUDA;
    }
    cudaFree(dx); cudaFree(d0); cudaFree(dp); cudaFree(dob);
    return rc;
}
