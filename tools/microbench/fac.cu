// micro-benchmark of the cooperative factorisation variants on one warp (real kernel header, synthetic SPD data)
#include <cstdio>
#include <vector>
#include <cmath>
#include "mpc_solve_kernel.cuh"
using namespace MPC_NS;

#include "fac_variants.cuh"

__global__ void kfac(const double* init, double* out, long long* cyc, int variant, int oz)
{
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < RS_DOUBLES; i += 32) sm[i] = init[i];
    __syncwarp();
    const long long t0 = clock64();
    if (variant == 0) riccati_factor_coop(sm);
    else if (variant == 1) factor_v1(sm);
    else if (variant == 2) factor_v2(sm);
    const long long t1 = clock64();
    __syncwarp();
    if (threadIdx.x == 0) cyc[variant] = t1 - t0;
    for (int i = threadIdx.x; i < RS_DOUBLES; i += 32) out[i] = sm[i];
}

int main()
{
    std::vector<double> h(RS_DOUBLES, 0.0);
    srand(1);
    auto rnd = []() { return (double)rand() / RAND_MAX - 0.5; };
    for (int s = 0; s <= NSTAGE; s++) {
        double* b = h.data() + s * RSTRIDE;
        // SPD augmented matrix: M = R'R + 2I (8x8), packed lower
        double Rm[8][8];
        for (auto& r : Rm) for (auto& v : r) v = rnd();
        for (int i = 0; i < 8; i++) for (int j = 0; j <= i; j++) {
            double a = (i == j) ? 2.0 : 0.0;
            for (int l = 0; l < 8; l++) a += Rm[l][i] * Rm[l][j];
            if (i * (i + 1) / 2 + j < NPK + NZ) b[RO_G + i * (i + 1) / 2 + j] = a;
        }
        for (int l = 0; l < NX; l++) for (int j = 0; j < NB; j++) b[RO_B + l * NB + j] = (l + NU == j ? 1.0 : 0.0) + 0.2 * rnd();
    }
    double *d_init, *d_out; long long* cyc;
    const size_t bytes = RS_DOUBLES * sizeof(double);
    cudaMalloc(&d_init, bytes); cudaMalloc(&d_out, bytes); cudaMallocManaged(&cyc, 64);
    cudaMemcpy(d_init, h.data(), bytes, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(kfac, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    std::vector<double> ref(RS_DOUBLES), got(RS_DOUBLES);
    for (int v = 0; v < 3; v++) {
        for (int rep = 0; rep < 2; rep++) { kfac<<<1, 32, bytes>>>(d_init, d_out, cyc, v, 0); cudaDeviceSynchronize(); }
        cudaMemcpy(got.data(), d_out, bytes, cudaMemcpyDeviceToHost);
        if (v == 0) ref = got;
        double err = 0.0;
        for (int s = 0; s < NSTAGE; s++) for (int i = 0; i < NPK + NZ; i++) err = fmax(err, fabs(got[s * RSTRIDE + i] - ref[s * RSTRIDE + i]) / (1.0 + fabs(ref[s * RSTRIDE + i])));
        printf("variant %d: %lld cycles/stage  (max rel diff vs v0 %.2e)  %s\n", v, cyc[v] / NSTAGE, err, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
