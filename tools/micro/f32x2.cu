// Microbenchmark: issue throughput of scalar FMUL/FADD vs packed FMUL2/FFMA2 on sm_100a, and a check that the
// FMUL2 + FFMA2(x, ONE, y) form rounds like separate mul/add (ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 upk(u64 v) { float2 o; asm("mov.b64 {%0,%1}, %2;" : "=f"(o.x), "=f"(o.y) : "l"(v)); return o; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void bench(float *out, float w, float one, int iters) {
    float a[8];
    u64 p[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { a[q] = threadIdx.x * 0.001f + q; p[q] = pk(a[q], a[q] + 0.5f); }
    const u64 W = pk(w, w), ONE = pk(one, one);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (MODE == 0) a[q] = __fmul_rn(a[q], w);                       // 8 scalar FMUL
            if (MODE == 1) p[q] = mul2(p[q], W);                            // 8 FMUL2 (16 mults)
            if (MODE == 2) a[q] = __fadd_rn(__fmul_rn(a[q], w), a[(q + 1) & 7]);   // FMUL + FADD
            if (MODE == 3) p[q] = fma2(mul2(p[q], W), ONE, p[(q + 1) & 7]);        // FMUL2 + FFMA2
        }
    }
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) { float2 u = upk(p[q]); s += a[q] + u.x + u.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void exact(const float *x, const float *y, const float *z, float *o_sep, float *o_pk, float *o_fused, float one, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    o_sep[i] = __fadd_rn(__fmul_rn(x[i], y[i]), z[i]);
    float2 u = upk(fma2(mul2(pk(x[i], x[i]), pk(y[i], y[i])), pk(one, one), pk(z[i], z[i])));
    o_pk[i] = u.x;
    o_fused[i] = __fmaf_rn(x[i], y[i], z[i]);
}

int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 256 * 4 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 4096, blocks = 148 * 8, threads = 256;
    const char *names[4] = {"FMUL x8", "FMUL2 x8", "FMUL+FADD x8", "FMUL2+FFMA2 x8"};
    for (int m = 0; m < 4; ++m) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (m == 0) bench<0><<<blocks, threads>>>(out, 1.0000001f, 1.f, iters);
            if (m == 1) bench<1><<<blocks, threads>>>(out, 1.0000001f, 1.f, iters);
            if (m == 2) bench<2><<<blocks, threads>>>(out, 0.5f, 1.f, iters);
            if (m == 3) bench<3><<<blocks, threads>>>(out, 0.5f, 1.f, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double inst = (double)blocks * threads / 32 * iters * 8 * (m >= 2 ? 2 : 1);
        printf("%-16s %.3f ms  %.1f G warp-inst/s  (%.2f inst/clk/SM at 1.965 GHz)\n", names[m], ms, inst / ms / 1e6,
               inst / (ms * 1e-3) / 148 / 1.965e9);
    }
    const int n = 1 << 20;
    float *h = (float *)malloc(3 * n * sizeof(float)), *d, *o;
    srand(1);
    for (int i = 0; i < 3 * n; ++i) h[i] = (float)rand() / RAND_MAX * 4.f - 2.f;
    cudaMalloc(&d, 3 * n * sizeof(float)); cudaMalloc(&o, 3 * n * sizeof(float));
    cudaMemcpy(d, h, 3 * n * sizeof(float), cudaMemcpyHostToDevice);
    exact<<<n / 256, 256>>>(d, d + n, d + 2 * n, o, o + n, o + 2 * n, 1.f, n);
    float *r = (float *)malloc(3 * n * sizeof(float));
    cudaMemcpy(r, o, 3 * n * sizeof(float), cudaMemcpyDeviceToHost);
    int diff_pk = 0, diff_fused = 0;
    for (int i = 0; i < n; ++i) { diff_pk += r[i] != r[n + i]; diff_fused += r[i] != r[2 * n + i]; }
    printf("separate vs packed(FMUL2+FFMA2*1): %d mismatches;  separate vs fused fma: %d mismatches (of %d)\n", diff_pk, diff_fused, n);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
