#include <cstdio>
#include <cuda_runtime.h>
// single-warp dependent-chain latencies (cycles per op)
__global__ void k(double* out, long long* cyc, double a, double b, int n)
{
    __shared__ double sm[64];
    sm[threadIdx.x] = a; sm[threadIdx.x + 32] = b;
    __syncwarp();
    double x = a + threadIdx.x;
    long long t0, t1;
    // DFMA chain
    t0 = clock64();
    for (int i = 0; i < n; i++) { x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); x = fma(x, b, a); }
    t1 = clock64(); cyc[0] = t1 - t0;
    // DADD chain
    t0 = clock64();
    for (int i = 0; i < n; i++) { x += a; x += b; x += a; x += b; }
    t1 = clock64(); cyc[1] = t1 - t0;
    // DMUL chain
    t0 = clock64();
    for (int i = 0; i < n; i++) { x *= b; x *= b; x *= b; x *= b; }
    t1 = clock64(); cyc[2] = t1 - t0;
    // rsqrt chain
    double y = fabs(x) + 2.0;
    t0 = clock64();
    for (int i = 0; i < n; i++) { y = rsqrt(y) + 1.5; y = rsqrt(y) + 1.5; y = rsqrt(y) + 1.5; y = rsqrt(y) + 1.5; }
    t1 = clock64(); cyc[3] = t1 - t0;
    // 4 independent DFMA chains (ILP)
    double x0 = x, x1 = x + 1, x2 = x + 2, x3 = x + 3;
    t0 = clock64();
    for (int i = 0; i < n; i++) { x0 = fma(x0, b, a); x1 = fma(x1, b, a); x2 = fma(x2, b, a); x3 = fma(x3, b, a); }
    t1 = clock64(); cyc[4] = t1 - t0;
    // LDS dependent chain (pointer chase through values)
    int idx = threadIdx.x & 31;
    t0 = clock64();
    for (int i = 0; i < n; i++) { idx = (int)sm[idx] & 31; idx = (int)sm[idx + 32] & 31; idx = (int)sm[idx] & 31; idx = (int)sm[idx+32] & 31; }
    t1 = clock64(); cyc[5] = t1 - t0;
    // LOP3 chain
    int q = idx + (int)y;
    t0 = clock64();
    for (int i = 0; i < n; i++) { q = (q | i) ^ 0x55; q = (q & ~i) ^ 0x33; q = (q | i) ^ 0x0f; q = (q & ~i) ^ 0x71; }
    t1 = clock64(); cyc[6] = t1 - t0;
    // 1/x division chain
    double w = y + 3.0;
    t0 = clock64();
    for (int i = 0; i < n; i++) { w = 1.0 / w + 2.0; w = 1.0 / w + 2.0; w = 1.0 / w + 2.0; w = 1.0 / w + 2.0; }
    t1 = clock64(); cyc[7] = t1 - t0;
    // shfl chain (64-bit)
    double s = w;
    t0 = clock64();
    for (int i = 0; i < n; i++) { s = __shfl_sync(0xffffffffu, s, 1) ; s = __shfl_sync(0xffffffffu, s, 2); s = __shfl_sync(0xffffffffu, s, 3); s = __shfl_sync(0xffffffffu, s, 4); }
    t1 = clock64(); cyc[8] = t1 - t0;
    // sts -> syncwarp -> lds round trip
    double r = s;
    t0 = clock64();
    for (int i = 0; i < n; i++) {
        sm[threadIdx.x] = r; __syncwarp(); r = sm[(threadIdx.x + 1) & 31] + 1.0; __syncwarp();
        sm[threadIdx.x] = r; __syncwarp(); r = sm[(threadIdx.x + 1) & 31] + 1.0; __syncwarp();
        sm[threadIdx.x] = r; __syncwarp(); r = sm[(threadIdx.x + 1) & 31] + 1.0; __syncwarp();
        sm[threadIdx.x] = r; __syncwarp(); r = sm[(threadIdx.x + 1) & 31] + 1.0; __syncwarp();
    }
    t1 = clock64(); cyc[9] = t1 - t0;
    out[threadIdx.x] = x + y + x0 + x1 + x2 + x3 + idx + q + w + s + r;
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 16 * 8);
    const int n = 2000;
    for (int rep = 0; rep < 2; rep++) { k<<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, n); cudaDeviceSynchronize(); }
    const char* nm[] = {"DFMA", "DADD", "DMUL", "rsqrt+add", "DFMA x4 ILP (per 4)", "LDS+cvt chase", "LOP3 x2", "1/x + add", "SHFL64", "STS-sync-LDS-add"};
    for (int i = 0; i < 10; i++) printf("%-24s %.1f cycles/op\n", nm[i], (double)cyc[i] / (4.0 * n));
    return 0;
}
