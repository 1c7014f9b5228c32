#include <cstdio>
#include <cuda_runtime.h>
constexpr int NS = 30, ST = 31;   // stages, doubles per stage (25 A + 5 b + pad)
__device__ __forceinline__ double gate5(const double* a, int oz)
{
    int t = (__double2hiint(a[0]) | __double2hiint(a[1]) | __double2hiint(a[2])) + (__double2hiint(a[3]) | __double2hiint(a[4]));
    return __hiloint2double(t & oz, 0);
}
__global__ void k(double* out, long long* cyc, int oz, double seed)
{
    __shared__ double sm[(NS + 1) * ST + 64];
    const int lane = threadIdx.x;
    for (int i = lane; i < (NS + 1) * ST; i += 32) sm[i] = 0.01 * ((i * 7) % 13) * seed;
    __syncwarp();
    long long t0, t1;
    // (0) SHFL dependent chain
    double x = seed + lane;
    t0 = clock64();
    for (int i = 0; i < 256; i++) x = __shfl_sync(0xffffffffu, x, (lane + 1) & 31) + 1.0;
    t1 = clock64(); cyc[0] = (t1 - t0) / 256;    // shfl64 + dadd
    // (1) sweep v1: 5 lanes, shuffle exchange, interleaved by ptxas
    const int i = lane < 5 ? lane : 4;
    x = seed * lane;
    t0 = clock64();
    for (int s = 0; s < NS; s++) {
        const double* b = sm + s * ST;
        double xs[5];
        for (int j = 0; j < 5; j++) xs[j] = __shfl_sync(0xffffffffu, x, j);
        x = b[25 + i];
        for (int j = 0; j < 5; j++) x += b[i * 5 + j] * xs[j];
        if (lane < 5) sm[(s + 1) * ST + 25 + 0] = x;
    }
    t1 = clock64(); cyc[1] = (t1 - t0) / NS;
    // (2) sweep v2: gate + tree
    x = seed * lane;
    t0 = clock64();
    for (int s = 0; s < NS; s++) {
        const double* b = sm + s * ST;
        double a[5], c = b[25 + i];
        for (int j = 0; j < 5; j++) a[j] = b[i * 5 + j];
        double xs[5];
        for (int j = 0; j < 5; j++) xs[j] = __shfl_sync(0xffffffffu, x, j);
        const double z0 = gate5(xs, oz);
        const double u0 = a[0] * xs[0] + (c + z0), u1 = a[1] * xs[1] + z0, u2 = a[4] * xs[4] + z0;
        x = ((a[2] * xs[2] + u0) + (a[3] * xs[3] + u1)) + u2;
        if (lane < 5) sm[(s + 1) * ST + 26] = x;
    }
    t1 = clock64(); cyc[2] = (t1 - t0) / NS;
    // (3) redundant: every lane all 5 components, no exchange
    double y[5] = {seed, seed * 2, seed * 3, seed * 4, seed * 5};
    t0 = clock64();
    for (int s = 0; s < NS; s++) {
        const double* b = sm + s * ST;
        double yn[5];
        for (int r = 0; r < 5; r++) {
            double acc = b[25 + r];
            for (int j = 0; j < 5; j++) acc += b[r * 5 + j] * y[j];
            yn[r] = acc;
        }
        for (int r = 0; r < 5; r++) y[r] = yn[r];
        if (lane == 0) sm[(s + 1) * ST + 27] = y[0];
    }
    t1 = clock64(); cyc[3] = (t1 - t0) / NS;
    // (4) smem exchange: 5 lanes, STS -> syncwarp -> LDS
    x = seed * lane;
    t0 = clock64();
    for (int s = 0; s < NS; s++) {
        const double* b = sm + s * ST;
        double* xch = sm + (NS + 1) * ST;
        if (lane < 5) xch[lane + (s & 1) * 8] = x;
        __syncwarp();
        double acc = b[25 + i];
        for (int j = 0; j < 5; j++) acc += b[i * 5 + j] * xch[j + (s & 1) * 8];
        x = acc;
    }
    t1 = clock64(); cyc[4] = (t1 - t0) / NS;
    // (5) empty-ish loop with only the LDS prefetch and STS
    t0 = clock64();
    for (int s = 0; s < NS; s++) {
        const double* b = sm + s * ST;
        x += b[25 + i];
        if (lane < 5) sm[(s + 1) * ST + 28] = x;
    }
    t1 = clock64(); cyc[5] = (t1 - t0) / NS;
    out[lane] = x + y[0] + y[1] + y[2] + y[3] + y[4];
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 16 * 8);
    for (int rep = 0; rep < 2; rep++) { k<<<1, 32>>>(out, cyc, 0, 1.0000001); cudaDeviceSynchronize(); }
    const char* nm[] = {"SHFL64+DADD chain", "sweep v1 shuffle", "sweep v2 gate+tree", "sweep v3 redundant (no exchange)", "sweep v4 smem exchange", "LDS+DADD+STS loop"};
    for (int i = 0; i < 6; i++) printf("%-34s %lld cycles/step\n", nm[i], cyc[i]);
    return 0;
}
