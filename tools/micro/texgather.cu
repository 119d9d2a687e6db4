// Micro-benchmark: the 8-tap trilinear gather of k_gen_warp_pk through (A) 64-bit global loads, (B) tld4 texture
// gathers (4 per voxel: one per x plane and component, each returning the 2x2 (y, z) footprint exactly), (C) point-
// sampled 3-D texture fetches, (D) a hybrid.  Same coordinates (rotation + scale + smooth displacement), same lerps;
// variants are compared bit for bit.    nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o texgather texgather.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int N = 160, B = 8;

struct Aff { float A[9], c[3]; };
struct Params { Aff a[B]; };

__device__ __forceinline__ void coords(const Aff &a, int i, int j, int k, float &x, float &y, float &z) {
    const float xc = i - 79.5f, yc = j - 79.5f, zc = k - 79.5f;
    // smooth displacement, +-3 voxels
    const float dx = 3.f * __sinf(0.021f * j + 0.017f * k), dy = 3.f * __sinf(0.019f * i + 0.023f * k),
                dz = 3.f * __sinf(0.018f * i + 0.02f * j);
    const float X = xc + dx, Y = yc + dy, Z = zc + dz;
    x = a.A[0] * X + a.A[1] * Y + a.A[2] * Z + a.c[0];
    y = a.A[3] * X + a.A[4] * Y + a.A[5] * Z + a.c[1];
    z = a.A[6] * X + a.A[7] * Y + a.A[8] * Z + a.c[2];
    x = fminf(fmaxf(x, 0.f), N - 2.f); y = fminf(fmaxf(y, 0.f), N - 2.f); z = fminf(fmaxf(z, 0.f), N - 2.f);
}

__device__ __forceinline__ float2 tri(const float2 t[8], float ax, float ay, float az) {
    // t: 000 001 100 101 010 011 110 111 (x y z)
    float2 r;
    {
        const float c00 = fmaf(ax, t[2].x - t[0].x, t[0].x), c01 = fmaf(ax, t[3].x - t[1].x, t[1].x);
        const float c10 = fmaf(ax, t[6].x - t[4].x, t[4].x), c11 = fmaf(ax, t[7].x - t[5].x, t[5].x);
        const float c0 = fmaf(ay, c10 - c00, c00), c1 = fmaf(ay, c11 - c01, c01);
        r.x = fmaf(az, c1 - c0, c0);
    }
    {
        const float c00 = fmaf(ax, t[2].y - t[0].y, t[0].y), c01 = fmaf(ax, t[3].y - t[1].y, t[1].y);
        const float c10 = fmaf(ax, t[6].y - t[4].y, t[4].y), c11 = fmaf(ax, t[7].y - t[5].y, t[5].y);
        const float c0 = fmaf(ay, c10 - c00, c00), c1 = fmaf(ay, c11 - c01, c01);
        r.y = fmaf(az, c1 - c0, c0);
    }
    return r;
}

struct Tex { cudaTextureObject_t pair2d[B], syn2d[B], t12d[B], pair3d[B]; };

// MODE 0: LDG.64 pairs; 1: tld4 on float2 array (4 gathers); 2: tld4 on two float arrays (4 gathers);
// 3: tex3D point float2 (8 fetches); 4: hybrid: x0 plane by tld4 (2 gathers), x0+1 plane by 4 LDG.64
// 5: syn via tld4 (2 gathers, float array), T1 via 8 LDG.32 from a separate linear volume
template <int MODE>
__global__ void __launch_bounds__(160) k_warp(const float2 *__restrict__ vol, const float *__restrict__ t1lin,
                                               const __grid_constant__ Params P, const __grid_constant__ Tex T,
                                               float *__restrict__ o0, float *__restrict__ o1, int rpb) {
    const int b = blockIdx.z, i = blockIdx.y, k = threadIdx.x;
    const Aff &a = P.a[b];
    const float2 *v = vol + (size_t)b * N * N * N;
    const float *t1 = t1lin + (size_t)b * N * N * N;
    const int j0 = blockIdx.x * rpb;
#pragma unroll 2
    for (int j = j0; j < j0 + rpb; ++j) {
        float x, y, z;
        coords(a, i, j, k, x, y, z);
        const int ix = (int)x, iy = (int)y, iz = (int)z;
        const float ax = x - ix, ay = y - iy, az = z - iz;
        float2 t[8];
        if (MODE == 0) {
            const float2 *b00 = v + (ix * N + iy) * N + iz, *b10 = b00 + N * N, *b01 = b00 + N, *b11 = b10 + N;
            t[0] = __ldg(b00); t[1] = __ldg(b00 + 1); t[2] = __ldg(b10); t[3] = __ldg(b10 + 1);
            t[4] = __ldg(b01); t[5] = __ldg(b01 + 1); t[6] = __ldg(b11); t[7] = __ldg(b11 + 1);
        } else if (MODE == 1 || MODE == 2) {
            // 2-D array: width = z, height = x * N + y.  gather at (u, v) = (iz + 1, row + 1) returns texels
            // (iz, row+1) (iz+1, row+1) (iz+1, row) (iz, row) as .x .y .z .w
            const float u = iz + 1.0f, r0 = (float)(ix * N + iy) + 1.0f, r1 = r0 + (float)N;
            float4 s0, s1, q0, q1;
            if (MODE == 1) {
                s0 = tex2Dgather<float4>(T.pair2d[b], u, r0, 0); q0 = tex2Dgather<float4>(T.pair2d[b], u, r0, 1);
                s1 = tex2Dgather<float4>(T.pair2d[b], u, r1, 0); q1 = tex2Dgather<float4>(T.pair2d[b], u, r1, 1);
            } else {
                s0 = tex2Dgather<float4>(T.syn2d[b], u, r0, 0); q0 = tex2Dgather<float4>(T.t12d[b], u, r0, 0);
                s1 = tex2Dgather<float4>(T.syn2d[b], u, r1, 0); q1 = tex2Dgather<float4>(T.t12d[b], u, r1, 0);
            }
            t[0] = make_float2(s0.w, q0.w); t[1] = make_float2(s0.z, q0.z); t[4] = make_float2(s0.x, q0.x); t[5] = make_float2(s0.y, q0.y);
            t[2] = make_float2(s1.w, q1.w); t[3] = make_float2(s1.z, q1.z); t[6] = make_float2(s1.x, q1.x); t[7] = make_float2(s1.y, q1.y);
        } else if (MODE == 3) {
            // 3-D array: width = z, height = y, depth = x; point sampling at texel centres
            const float fz = iz + 0.5f, fy = iy + 0.5f, fx = ix + 0.5f;
            t[0] = tex3D<float2>(T.pair3d[b], fz, fy, fx);       t[1] = tex3D<float2>(T.pair3d[b], fz + 1, fy, fx);
            t[2] = tex3D<float2>(T.pair3d[b], fz, fy, fx + 1);   t[3] = tex3D<float2>(T.pair3d[b], fz + 1, fy, fx + 1);
            t[4] = tex3D<float2>(T.pair3d[b], fz, fy + 1, fx);   t[5] = tex3D<float2>(T.pair3d[b], fz + 1, fy + 1, fx);
            t[6] = tex3D<float2>(T.pair3d[b], fz, fy + 1, fx + 1); t[7] = tex3D<float2>(T.pair3d[b], fz + 1, fy + 1, fx + 1);
        } else if (MODE == 4) {
            const float u = iz + 1.0f, r0 = (float)(ix * N + iy) + 1.0f;
            const float4 s0 = tex2Dgather<float4>(T.pair2d[b], u, r0, 0), q0 = tex2Dgather<float4>(T.pair2d[b], u, r0, 1);
            t[0] = make_float2(s0.w, q0.w); t[1] = make_float2(s0.z, q0.z); t[4] = make_float2(s0.x, q0.x); t[5] = make_float2(s0.y, q0.y);
            const float2 *b10 = v + ((ix + 1) * N + iy) * N + iz, *b11 = b10 + N;
            t[2] = __ldg(b10); t[3] = __ldg(b10 + 1); t[6] = __ldg(b11); t[7] = __ldg(b11 + 1);
        } else {
            const float u = iz + 1.0f, r0 = (float)(ix * N + iy) + 1.0f, r1 = r0 + (float)N;
            const float4 s0 = tex2Dgather<float4>(T.syn2d[b], u, r0, 0), s1 = tex2Dgather<float4>(T.syn2d[b], u, r1, 0);
            const float *b00 = t1 + (ix * N + iy) * N + iz, *b10 = b00 + N * N, *b01 = b00 + N, *b11 = b10 + N;
            t[0] = make_float2(s0.w, __ldg(b00)); t[1] = make_float2(s0.z, __ldg(b00 + 1));
            t[4] = make_float2(s0.x, __ldg(b01)); t[5] = make_float2(s0.y, __ldg(b01 + 1));
            t[2] = make_float2(s1.w, __ldg(b10)); t[3] = make_float2(s1.z, __ldg(b10 + 1));
            t[6] = make_float2(s1.x, __ldg(b11)); t[7] = make_float2(s1.y, __ldg(b11 + 1));
        }
        const float2 r = tri(t, ax, ay, az);
        const size_t o = (((size_t)b * N + i) * N + j) * N + k;
        o0[o] = r.x; o1[o] = r.y;
    }
}

// write cost of the synthetic image: linear float2 stores vs surface stores into the 2-D array
__global__ void k_fill_lin(float2 *v) {
    const size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < (size_t)B * N * N * N) v[p] = make_float2((float)(p % 977), (float)(p % 331));
}
__global__ void k_fill_surf(cudaSurfaceObject_t s, int b) {
    const int row = blockIdx.x, z = threadIdx.x;
    const size_t p = ((size_t)b * N * N + row) * N + z;
    surf2Dwrite(make_float2((float)(p % 977), (float)(p % 331)), s, z * 8, row);
}
__global__ void k_fill_surf1(cudaSurfaceObject_t s, int b) {     // float array, 4 z per thread
    const int row = blockIdx.x, z = threadIdx.x * 4;
    const size_t p = ((size_t)b * N * N + row) * N + z;
    surf2Dwrite(make_float4((float)(p % 977), (float)((p + 1) % 977), (float)((p + 2) % 977), (float)((p + 3) % 977)), s, z * 4, row);
}

static cudaTextureObject_t mk_tex(cudaArray_t arr) {
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    cudaTextureObject_t t; CK(cudaCreateTextureObject(&t, &rd, &td, nullptr));
    return t;
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    printf("device %s, maxTexture2DGather %d x %d, maxTexture2D %d x %d, maxTexture2DLayered %d x %d x %d, SMs %d\n", pr.name,
           pr.maxTexture2DGather[0], pr.maxTexture2DGather[1], pr.maxTexture2D[0], pr.maxTexture2D[1],
           pr.maxTexture2DLayered[0], pr.maxTexture2DLayered[1], pr.maxTexture2DLayered[2], pr.multiProcessorCount);
    const size_t V = (size_t)N * N * N;
    std::vector<float2> h(V * B);
    std::vector<float> h1(V * B), hs(V * B);
    srand(1);
    for (size_t p = 0; p < V * B; ++p) { h[p] = make_float2((float)(rand() % 1000) * 0.37f, (float)(rand() % 1000) * 0.11f); hs[p] = h[p].x; h1[p] = h[p].y; }
    float2 *vol; float *t1, *o0, *o1, *r0, *r1;
    CK(cudaMalloc(&vol, V * B * 8)); CK(cudaMalloc(&t1, V * B * 4));
    CK(cudaMalloc(&o0, V * B * 4)); CK(cudaMalloc(&o1, V * B * 4)); CK(cudaMalloc(&r0, V * B * 4)); CK(cudaMalloc(&r1, V * B * 4));
    CK(cudaMemcpy(vol, h.data(), V * B * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(t1, h1.data(), V * B * 4, cudaMemcpyHostToDevice));
    Tex T = {};
    cudaArray_t a_pair[B], a_syn[B], a_t1[B], a_3d[B];
    cudaChannelFormatDesc c2 = cudaCreateChannelDesc<float2>(), c1 = cudaCreateChannelDesc<float>();
    for (int b = 0; b < B; ++b) {
        CK(cudaMallocArray(&a_pair[b], &c2, N, N * N, cudaArrayTextureGather | cudaArraySurfaceLoadStore));
        CK(cudaMallocArray(&a_syn[b], &c1, N, N * N, cudaArrayTextureGather | cudaArraySurfaceLoadStore));
        CK(cudaMallocArray(&a_t1[b], &c1, N, N * N, cudaArrayTextureGather));
        CK(cudaMemcpy2DToArray(a_pair[b], 0, 0, h.data() + V * b, N * 8, N * 8, N * N, cudaMemcpyHostToDevice));
        CK(cudaMemcpy2DToArray(a_syn[b], 0, 0, hs.data() + V * b, N * 4, N * 4, N * N, cudaMemcpyHostToDevice));
        CK(cudaMemcpy2DToArray(a_t1[b], 0, 0, h1.data() + V * b, N * 4, N * 4, N * N, cudaMemcpyHostToDevice));
        CK(cudaMalloc3DArray(&a_3d[b], &c2, make_cudaExtent(N, N, N), 0));
        cudaMemcpy3DParms cp = {};
        cp.srcPtr = make_cudaPitchedPtr(h.data() + V * b, N * 8, N, N); cp.dstArray = a_3d[b];
        cp.extent = make_cudaExtent(N, N, N); cp.kind = cudaMemcpyHostToDevice;
        CK(cudaMemcpy3D(&cp));
        T.pair2d[b] = mk_tex(a_pair[b]); T.syn2d[b] = mk_tex(a_syn[b]); T.t12d[b] = mk_tex(a_t1[b]); T.pair3d[b] = mk_tex(a_3d[b]);
    }
    Params P;
    for (int b = 0; b < B; ++b) {
        // rotations about the three axes (5..15 degrees), scale 0.9..1.1, shear 0.1
        const double rx = (5 + 10.0 * b / 7) * M_PI / 180 * ((b & 1) ? -1 : 1), ry = (15 - 10.0 * b / 7) * M_PI / 180, rz = 8 * M_PI / 180 * ((b & 2) ? -1 : 1);
        const double sc = 0.9 + 0.2 * b / 7;
        double Rx[9] = {1, 0, 0, 0, cos(rx), -sin(rx), 0, sin(rx), cos(rx)}, Ry[9] = {cos(ry), 0, sin(ry), 0, 1, 0, -sin(ry), 0, cos(ry)},
               Rz[9] = {cos(rz), -sin(rz), 0, sin(rz), cos(rz), 0, 0, 0, 1}, Sh[9] = {sc, 0.1, 0.05, 0.02, sc, 0.1, 0.05, 0.03, sc}, M1[9], M2[9], M3[9];
        auto mm = [](const double *a, const double *bb, double *c) { for (int r = 0; r < 3; ++r) for (int q = 0; q < 3; ++q) { c[r * 3 + q] = 0; for (int t = 0; t < 3; ++t) c[r * 3 + q] += a[r * 3 + t] * bb[t * 3 + q]; } };
        mm(Rx, Ry, M1); mm(M1, Rz, M2); mm(M2, Sh, M3);
        for (int q = 0; q < 9; ++q) P.a[b].A[q] = (float)M3[q];
        for (int q = 0; q < 3; ++q) P.a[b].c[q] = 79.5f + (q - 1) * 2.f;
    }
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int rpb = 32;
    dim3 grid(N / rpb, N, B), blk(N);
    std::vector<float> ref0(V * B), ref1(V * B), g0(V * B), g1(V * B);
    auto run = [&](int mode, const char *name) {
        float *a0 = mode == 0 ? r0 : o0, *a1 = mode == 0 ? r1 : o1;
        auto launch = [&]() {
            switch (mode) {
                case 0: k_warp<0><<<grid, blk>>>(vol, t1, P, T, a0, a1, rpb); break;
                case 1: k_warp<1><<<grid, blk>>>(vol, t1, P, T, a0, a1, rpb); break;
                case 2: k_warp<2><<<grid, blk>>>(vol, t1, P, T, a0, a1, rpb); break;
                case 3: k_warp<3><<<grid, blk>>>(vol, t1, P, T, a0, a1, rpb); break;
                case 4: k_warp<4><<<grid, blk>>>(vol, t1, P, T, a0, a1, rpb); break;
                case 5: k_warp<5><<<grid, blk>>>(vol, t1, P, T, a0, a1, rpb); break;
            }
        };
        for (int w = 0; w < 3; ++w) launch();
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        const int reps = 20;
        for (int w = 0; w < reps; ++w) launch();
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaGetLastError());
        double maxd = -1;
        if (mode == 0) {
            CK(cudaMemcpy(ref0.data(), r0, V * B * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(ref1.data(), r1, V * B * 4, cudaMemcpyDeviceToHost));
        } else {
            CK(cudaMemcpy(g0.data(), o0, V * B * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(g1.data(), o1, V * B * 4, cudaMemcpyDeviceToHost));
            maxd = 0;
            for (size_t p = 0; p < V * B; ++p) { maxd = fmax(maxd, fabs((double)g0[p] - ref0[p])); maxd = fmax(maxd, fabs((double)g1[p] - ref1[p])); }
        }
        printf("%-44s %8.1f us per batch of %d  (%.2f us/sample)  max|diff vs LDG| %g\n", name, ms * 1000 / reps, B, ms * 1000 / reps / B, maxd);
    };
    run(0, "A  8 x LDG.64 {syn,T1}");
    run(1, "B1 4 x tld4 on one float2 array");
    run(2, "B2 4 x tld4 on two float arrays");
    run(3, "C  8 x tex3D point float2");
    run(4, "D  2 x tld4 (plane x0) + 4 x LDG.64 (x0+1)");
    run(5, "E  2 x tld4 syn + 8 x LDG.32 T1");
    // ---- fill cost
    {
        cudaSurfaceObject_t s2[B], s1[B];
        for (int b = 0; b < B; ++b) {
            cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray;
            rd.res.array.array = a_pair[b]; CK(cudaCreateSurfaceObject(&s2[b], &rd));
            rd.res.array.array = a_syn[b]; CK(cudaCreateSurfaceObject(&s1[b], &rd));
        }
        for (int rep = 0; rep < 2; ++rep) {
            float ms;
            CK(cudaEventRecord(e0));
            k_fill_lin<<<(unsigned)((V * B + 255) / 256), 256>>>(vol);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("fill linear float2 (8 vols): %.1f us\n", ms * 1000);
            CK(cudaEventRecord(e0));
            for (int b = 0; b < B; ++b) k_fill_surf<<<N * N, N>>>(s2[b], b);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("fill surface float2 (8 arrays): %.1f us\n", ms * 1000);
            CK(cudaEventRecord(e0));
            for (int b = 0; b < B; ++b) k_fill_surf1<<<N * N, N / 4>>>(s1[b], b);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("fill surface float (8 arrays, 16 B per thread): %.1f us\n", ms * 1000);
        }
        CK(cudaGetLastError());
    }
    return 0;
}
