// Latency microbenchmark (one warp): dependent chains of the operations the solve kernel is
// made of. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat tools/lat_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void k(double* out, long long* cyc, double seed)
{
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = seed + i * 1e-9;
    __syncwarp();
    double a = seed, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    // DFMA dependent chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = fma(a, b, c);
    t1 = clock64(); cyc[0] = t1 - t0;
    // DADD chain
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) a = a + c;
    t1 = clock64(); cyc[1] = t1 - t0;
    // 4 independent DFMA chains
    double a1 = a + 1, a2 = a + 2, a3 = a + 3;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) { a = fma(a, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c); }
    t1 = clock64(); cyc[2] = t1 - t0; a += a1 + a2 + a3;
    // LDS.64 dependent (pointer chase through values)
    int idx = threadIdx.x;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) { double v = sm[idx]; idx = (idx + (int)(v * 0.0) + 33) & 1023; }
    t1 = clock64(); cyc[3] = t1 - t0; a += idx;
    // SHFL double dependent
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) a = __shfl_xor_sync(0xffffffffu, a, 1);
    t1 = clock64(); cyc[4] = t1 - t0;
    // division dependent
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) a = b / (a + 2.0);
    t1 = clock64(); cyc[5] = t1 - t0;
    // rsqrt dependent
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) a = rsqrt(a + 2.0);
    t1 = clock64(); cyc[6] = t1 - t0;
    // __drcp_rn dependent
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) a = __drcp_rn(a + 2.0);
    t1 = clock64(); cyc[7] = t1 - t0;
    // syncwarp
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) __syncwarp();
    t1 = clock64(); cyc[8] = t1 - t0;
    // FFMA dependent chain (fp32)
    float f = (float)seed, g = 1.0001f, h = 1e-6f;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) f = fmaf(f, g, h);
    t1 = clock64(); cyc[9] = t1 - t0; a += f;
    // redux
    unsigned u = threadIdx.x;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; ++i) u = __reduce_max_sync(0xffffffffu, u + 1);
    t1 = clock64(); cyc[10] = t1 - t0; a += u;
    // LDS.128 + 2 DFMA streaming (independent loads, 2 chains)
    const double2* s2 = (const double2*)sm;
    double z0 = 0, z1 = 0;
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; i += 2) { double2 x = s2[(i + threadIdx.x) & 511], y = s2[(i + 1 + threadIdx.x) & 511];
        z0 = fma(x.x, y.x, z0); z1 = fma(x.y, y.y, z1); }
    t1 = clock64(); cyc[11] = t1 - t0; a += z0 + z1;
    // sqrt
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) a = sqrt(a + 2.0);
    t1 = clock64(); cyc[12] = t1 - t0;
    out[threadIdx.x] = a;
}
int main()
{
    double* out; long long* cyc;
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 16 * 8);
    k<<<1, 32>>>(out, cyc, 1.5); k<<<1, 32>>>(out, cyc, 1.5);
    long long h[16];
    cudaMemcpy(h, cyc, 16 * 8, cudaMemcpyDeviceToHost);
    const char* names[] = {"DFMA dep", "DADD dep", "DFMA x4 indep (per group)", "LDS.64 dep", "SHFL f64 dep", "DDIV dep",
                           "rsqrt f64 dep", "drcp dep", "syncwarp", "FFMA dep", "redux.max dep", "LDS.128x2+2DFMA iter(2 elem)", "sqrt f64 dep"};
    for (int i = 0; i < 13; ++i) printf("%-32s %.1f cycles\n", names[i], (double)h[i] / (i == 11 ? 256 : 512));
    return 0;
}
