#include <cstdio>
#include <cuda_runtime.h>
// single-warp shared-memory load throughput: n independent LDS.64 / LDS.128 per lane, lane stride as in the sweep workspace
template <int STRIDE, int NLD, bool V2>
__global__ void k(double* out, long long* cyc, int reps)
{
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 32 * STRIDE + 64; i += 32) sm[i] = 1.0 + i * 1e-3;
    __syncwarp();
    const double* me = sm + threadIdx.x * STRIDE;
    double acc = 0.0;
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) {
        double v[NLD];
        if constexpr (V2) {
#pragma unroll
            for (int i = 0; i < NLD; i += 2) { const double2 t = *reinterpret_cast<const double2*>(me + i); v[i] = t.x; v[i + 1] = t.y; }
        } else {
#pragma unroll
            for (int i = 0; i < NLD; i++) v[i] = me[i];
        }
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
        for (int i = 0; i < NLD; i += 4) { s0 += v[i]; s1 += v[i + 1]; s2 += v[i + 2]; s3 += v[i + 3]; }
        acc += (s0 + s1) + (s2 + s3);
        sm[threadIdx.x * STRIDE + (r & 3)] = acc;      // keep the loads inside the loop
        __syncwarp();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    out[threadIdx.x] = acc;
}
template <int STRIDE, int NLD, bool V2>
void run(const char* name, double* out, long long* cyc)
{
    const int reps = 1000;
    cudaFuncSetAttribute(k<STRIDE, NLD, V2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (32 * STRIDE + 64) * 8);
    for (int i = 0; i < 2; i++) { k<STRIDE, NLD, V2><<<1, 32, (32 * STRIDE + 64) * 8>>>(out, cyc, reps); cudaDeviceSynchronize(); }
    printf("%-40s %.1f cycles per iteration (%d loads of %d doubles, %d DADD)\n", name, (double)cyc[0] / reps, V2 ? NLD / 2 : NLD, V2 ? 2 : 1, NLD + 4);
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 8);
    run<95, 32, false>("LDS.64  x32, stride 95", out, cyc);
    run<95, 16, false>("LDS.64  x16, stride 95", out, cyc);
    run<95, 8, false>("LDS.64  x8, stride 95", out, cyc);
    run<98, 32, true>("LDS.128 x16, stride 98", out, cyc);
    run<98, 16, true>("LDS.128 x8, stride 98", out, cyc);
    run<96, 32, false>("LDS.64  x32, stride 96 (conflicts)", out, cyc);
    run<1, 32, false>("LDS.64  x32, stride 1 (overlapping)", out, cyc);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
