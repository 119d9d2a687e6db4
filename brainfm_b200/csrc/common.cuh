// Shared device helpers for libbfm (sm_100a).  No torch types anywhere in csrc/.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/bfm.h"

namespace bfm {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char *fmt, const char *a = "") {
    snprintf(g_err, sizeof(g_err), fmt, a);
    return code;
}

inline int check_launch(const char *what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return BFM_E_CUDA;
    }
    return BFM_OK;
}

#define BFM_REQUIRE(cond, msg) \
    do {                       \
        if (!(cond)) return bfm::fail(BFM_E_INVALID, "%s", msg); \
    } while (0)

// ---- exactly-rounded building blocks --------------------------------------------------------
// The reference evaluates `w0*a + w1*b` as three separate ATen kernels (mul, mul, add), i.e. three
// separately rounded fp32 operations.  __fmul_rn/__fadd_rn are never contracted into an FMA.
__device__ __forceinline__ float lerp_rn(float w0, float a, float w1, float b) {
    return __fadd_rn(__fmul_rn(w0, a), __fmul_rn(w1, b));
}

__device__ __forceinline__ float nan_to_num(float v) {
    if (isnan(v)) return 0.f;
    if (isinf(v)) return v > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
    return v;
}

// order-preserving float <-> int encoding for atomic min/max on arbitrary floats
__device__ __forceinline__ int f2ord(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) {
    return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff);
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- Philox4x32-10 counter-based generator + Box-Muller --------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

// four independent N(0,1) variates for (seed, stream, 64-bit group index)
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint32_t stream, uint64_t group) {
    uint4 r = philox4x32_10(make_uint4((uint32_t)group, (uint32_t)(group >> 32), stream, 0u),
                            make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float S = 2.3283064365386963e-10f;  // 2^-32
    float u0 = ((float)r.x + 0.5f) * S, u1 = (float)r.y * S;
    float u2 = ((float)r.z + 0.5f) * S, u3 = (float)r.w * S;
    u0 = fminf(u0, 0.99999994f);
    u2 = fminf(u2, 0.99999994f);
    const float ta = -1.3862943611198906f * __log2f(u0), tb = -1.3862943611198906f * __log2f(u2);   // -2 ln u
    float ra = ta * rsqrtf(ta), rb = tb * rsqrtf(tb);
    float sa, ca, sb, cb;
    __sincosf(6.283185307179586f * u1 - 3.14159265358979f, &sa, &ca);
    __sincosf(6.283185307179586f * u3 - 3.14159265358979f, &sb, &cb);
    return make_float4(ra * ca, ra * sa, rb * cb, rb * sb);
}

// ---- deformation of one output row -------------------------------------------------------------
// A warp owns the output row (i, j, 0..sz-1).  The first two passes of the separable zoom
// (myzoom_torch axis 0 then axis 1, Generator/utils.py:239-244) only depend on (i, j, kz), so the
// warp evaluates them once per small-grid z node into shared memory; each voxel then needs only
// the third pass.  Identical operations, identical rounding as the reference.
constexpr int kMaxSmallZ = 96;   // >= fs[2], bs[2]

__device__ __forceinline__ void row_zoom_setup(const float *__restrict__ small, int n1, int n2, int C,
                                               const bfm_zoom_tab &t, int i, int j, float *sm, int lane) {
    const int lo0 = t.lo[0][i], hi0 = t.hi[0][i], lo1 = t.lo[1][j], hi1 = t.hi[1][j];
    const float wl0 = t.wl[0][i], wh0 = t.wh[0][i], wl1 = t.wl[1][j], wh1 = t.wh[1][j];
    const int rowlen = n2 * C;
    const float *p00 = small + ((int64_t)lo0 * n1 + lo1) * rowlen;
    const float *p10 = small + ((int64_t)hi0 * n1 + lo1) * rowlen;
    const float *p01 = small + ((int64_t)lo0 * n1 + hi1) * rowlen;
    const float *p11 = small + ((int64_t)hi0 * n1 + hi1) * rowlen;
    for (int q = lane; q < rowlen; q += 32) {
        float a = lerp_rn(wl0, __ldg(p00 + q), wh0, __ldg(p10 + q));
        float b = lerp_rn(wl0, __ldg(p01 + q), wh0, __ldg(p11 + q));
        sm[q] = lerp_rn(wl1, a, wh1, b);
    }
}

// Loop-invariant deformation state, loaded once into registers.
struct DefRegs {
    float A[9], c2[3], mx, my, mz, ctr0, ctr1, ctr2;
    int s0, s1, s2, photo;
    const float *F_full, *fsmall;
};

__device__ __forceinline__ DefRegs load_def(const bfm_deform &d) {
    DefRegs g;
#pragma unroll
    for (int q = 0; q < 9; ++q) g.A[q] = d.A[q];
#pragma unroll
    for (int q = 0; q < 3; ++q) g.c2[q] = d.c2[q];
    g.mx = (float)(d.src[0] - 1); g.my = (float)(d.src[1] - 1); g.mz = (float)(d.src[2] - 1);
    g.ctr0 = d.ctr[0]; g.ctr1 = d.ctr[1]; g.ctr2 = d.ctr[2];
    g.s0 = d.size[0]; g.s1 = d.size[1]; g.s2 = d.size[2];
    g.photo = d.photo;
    g.F_full = d.F_full;
    g.fsmall = d.F_full ? nullptr : d.fsmall;
    return g;
}

// ((A0*x + A1*y) + A2*z) + c, evaluated left to right with separately rounded ops (datasets.py:276-278),
// then the clamp to the source volume (datasets.py:279-284).
__device__ __forceinline__ void affine_clamp(const DefRegs &g, float x1, float y1, float z1, float &px, float &py,
                                             float &pz) {
    px = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(g.A[0], x1), __fmul_rn(g.A[1], y1)), __fmul_rn(g.A[2], z1)), g.c2[0]);
    py = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(g.A[3], x1), __fmul_rn(g.A[4], y1)), __fmul_rn(g.A[5], z1)), g.c2[1]);
    pz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(g.A[6], x1), __fmul_rn(g.A[7], y1)), __fmul_rn(g.A[8], z1)), g.c2[2]);
    px = px < 0.f ? 0.f : px; py = py < 0.f ? 0.f : py; pz = pz < 0.f ? 0.f : pz;
    px = px > g.mx ? g.mx : px; py = py > g.my ? g.my : py; pz = pz > g.mz ? g.mz : pz;
}

// Row-blocked traversal of the output grid.  A warp owns R consecutive rows (i, j, 0..s2-1) starting at row0.
// For every k-chunk the per-k table entries of the third zoom pass are loaded ONCE into registers
// (`kfn(k)` lets the caller load its own k-dependent state) and reused for the R rows;
// `fn(r, row, i, j, k, px, py, pz)` receives the clamped source coordinates.
// smF: this warp's shared scratch, R * 3*fs2 floats (rows of the first two zoom passes).
template <int R, typename KFn, typename VFn>
__device__ __forceinline__ void deform_rows(const bfm_deform &d, const DefRegs &g, float *smF, int row0,
                                            int n_rows, int lane, KFn kfn, VFn fn) {
    const int fs2x3 = d.fs[2] * 3;
    int ri[R], rj[R];
    float xc[R], yc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int row = min(row0 + r, n_rows - 1);
        ri[r] = row / g.s1;
        rj[r] = row - ri[r] * g.s1;
        xc[r] = __fsub_rn((float)ri[r], g.ctr0);
        yc[r] = __fsub_rn((float)rj[r], g.ctr1);
        if (g.fsmall) row_zoom_setup(g.fsmall, d.fs[1], d.fs[2], 3, d.ftab, ri[r], rj[r], smF + r * fs2x3, lane);
    }
    __syncwarp();
    const int nr = min(R, n_rows - row0);
    for (int k = lane; k < g.s2; k += 32) {
        int lo = 0, hi = 0;
        float wl = 0.f, wh = 0.f;
        if (g.fsmall) {
            lo = __ldg(d.ftab.lo[2] + k) * 3; hi = __ldg(d.ftab.hi[2] + k) * 3;
            wl = __ldg(d.ftab.wl[2] + k); wh = __ldg(d.ftab.wh[2] + k);
        }
        const float zc = __fsub_rn((float)k, g.ctr2);
        kfn(k);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r >= nr) break;
            float x1 = xc[r], y1 = yc[r], z1 = zc;
            if (g.F_full) {
                const float *f = g.F_full + ((int64_t)(row0 + r) * g.s2 + k) * 3;
                x1 = __fadd_rn(x1, f[0]); y1 = __fadd_rn(y1, f[1]); z1 = __fadd_rn(z1, f[2]);
            } else if (g.fsmall) {
                const float *sm = smF + r * fs2x3;
                const float f0 = lerp_rn(wl, sm[lo], wh, sm[hi]);
                const float f1 = g.photo ? 0.f : lerp_rn(wl, sm[lo + 1], wh, sm[hi + 1]);
                const float f2 = lerp_rn(wl, sm[lo + 2], wh, sm[hi + 2]);
                x1 = __fadd_rn(x1, f0); y1 = __fadd_rn(y1, f1); z1 = __fadd_rn(z1, f2);
            }
            float px, py, pz;
            affine_clamp(g, x1, y1, z1, px, py, pz);
            fn(r, row0 + r, ri[r], rj[r], k, px, py, pz);
        }
    }
}

// Bounding-box-relative trilinear taps with 32-bit element indices (volumes < 2^31 elements).
struct BoxRegs {
    float l0, l1, l2, h0, h1, h2;   // bbox origin and (crop extent - 1) as floats
    int b0, b1, b2, c0, c1, c2;     // bbox origin, crop extent - 1
    int n1n2, n2;                   // source strides
};
__device__ __forceinline__ BoxRegs load_box(const int *bb, int n1, int n2) {
    BoxRegs q;
    q.b0 = bb[0]; q.b1 = bb[1]; q.b2 = bb[2];
    q.c0 = bb[3] - bb[0] - 1; q.c1 = bb[4] - bb[1] - 1; q.c2 = bb[5] - bb[2] - 1;
    q.l0 = (float)q.b0; q.l1 = (float)q.b1; q.l2 = (float)q.b2;
    q.h0 = (float)q.c0; q.h1 = (float)q.c1; q.h2 = (float)q.c2;
    q.n1n2 = n1 * n2; q.n2 = n2;
    return q;
}
struct Taps32 {
    int base, dx, dy, dz;           // element index of the (lo,lo,lo) tap and the offsets to the hi taps
    float ax0, ax1, ay0, ay1, az0, az1;
    bool ok;
};
__device__ __forceinline__ Taps32 make_taps32(float px, float py, float pz, const BoxRegs &q) {
    Taps32 t;
    const float rx = __fsub_rn(px, q.l0), ry = __fsub_rn(py, q.l1), rz = __fsub_rn(pz, q.l2);
    t.ok = (rx > 0.f) & (ry > 0.f) & (rz > 0.f) & (rx <= q.h0) & (ry <= q.h1) & (rz <= q.h2);
    const float fx = floorf(rx), fy = floorf(ry), fz = floorf(rz);
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    t.ax1 = __fsub_rn(rx, fx); t.ax0 = __fsub_rn(1.f, t.ax1);
    t.ay1 = __fsub_rn(ry, fy); t.ay0 = __fsub_rn(1.f, t.ay1);
    t.az1 = __fsub_rn(rz, fz); t.az0 = __fsub_rn(1.f, t.az1);
    // invalid points gather from the crop origin (always in range) and are masked by the caller: the loads
    // stay unconditional, so the rows a warp owns can be in flight together
    t.base = t.ok ? (q.b0 + ix) * q.n1n2 + (q.b1 + iy) * q.n2 + (q.b2 + iz) : q.b0 * q.n1n2 + q.b1 * q.n2 + q.b2;
    t.dx = (t.ok && ix < q.c0) ? q.n1n2 : 0;       // hi = min(lo+1, n-1)  (utils.py:148-149)
    t.dy = (t.ok && iy < q.c1) ? q.n2 : 0;
    t.dz = (t.ok && iz < q.c2) ? 1 : 0;
    return t;
}
template <typename Fetch>
__device__ __forceinline__ float trilerp32(const Taps32 &t, Fetch at) {
    const int b = t.base;
    float e00 = lerp_rn(at(b), t.ax0, at(b + t.dx), t.ax1);
    float e01 = lerp_rn(at(b + t.dz), t.ax0, at(b + t.dx + t.dz), t.ax1);
    float e10 = lerp_rn(at(b + t.dy), t.ax0, at(b + t.dx + t.dy), t.ax1);
    float e11 = lerp_rn(at(b + t.dy + t.dz), t.ax0, at(b + t.dx + t.dy + t.dz), t.ax1);
    float f0 = lerp_rn(e00, t.ay0, e10, t.ay1);
    float f1 = lerp_rn(e01, t.ay0, e11, t.ay1);
    return lerp_rn(f0, t.az0, f1, t.az1);
}

// Trilinear taps of a bbox-relative coordinate triple (fast_3D_interp_torch, utils.py:141-166).
struct Taps {
    int x0, x1, y0, y1, z0, z1;   // ABSOLUTE source indices
    float ax0, ax1, ay0, ay1, az0, az1;
    bool ok;
};

__device__ __forceinline__ Taps make_taps(float px, float py, float pz, const int *bb) {
    Taps t;
    const float rx = __fsub_rn(px, (float)bb[0]), ry = __fsub_rn(py, (float)bb[1]), rz = __fsub_rn(pz, (float)bb[2]);
    const int nx = bb[3] - bb[0], ny = bb[4] - bb[1], nz = bb[5] - bb[2];
    t.ok = (rx > 0.f) && (ry > 0.f) && (rz > 0.f) && (rx <= (float)(nx - 1)) && (ry <= (float)(ny - 1)) &&
           (rz <= (float)(nz - 1));
    const float fx = floorf(rx), fy = floorf(ry), fz = floorf(rz);
    const int ix = (int)fx, iy = (int)fy, iz = (int)fz;
    t.ax1 = __fsub_rn(rx, fx); t.ax0 = __fsub_rn(1.f, t.ax1);
    t.ay1 = __fsub_rn(ry, fy); t.ay0 = __fsub_rn(1.f, t.ay1);
    t.az1 = __fsub_rn(rz, fz); t.az0 = __fsub_rn(1.f, t.az1);
    t.x0 = bb[0] + ix; t.x1 = bb[0] + min(ix + 1, nx - 1);
    t.y0 = bb[1] + iy; t.y1 = bb[1] + min(iy + 1, ny - 1);
    t.z0 = bb[2] + iz; t.z1 = bb[2] + min(iz + 1, nz - 1);
    return t;
}

// x, then y, then z; every product and sum separately rounded (utils.py:177-185)
template <typename Fetch>
__device__ __forceinline__ float trilerp(const Taps &t, Fetch at) {
    float e00 = lerp_rn(at(t.x0, t.y0, t.z0), t.ax0, at(t.x1, t.y0, t.z0), t.ax1);
    float e01 = lerp_rn(at(t.x0, t.y0, t.z1), t.ax0, at(t.x1, t.y0, t.z1), t.ax1);
    float e10 = lerp_rn(at(t.x0, t.y1, t.z0), t.ax0, at(t.x1, t.y1, t.z0), t.ax1);
    float e11 = lerp_rn(at(t.x0, t.y1, t.z1), t.ax0, at(t.x1, t.y1, t.z1), t.ax1);
    float f0 = lerp_rn(e00, t.ay0, e10, t.ay1);
    float f1 = lerp_rn(e01, t.ay0, e11, t.ay1);
    return lerp_rn(f0, t.az0, f1, t.az1);
}

constexpr int kRowWarps = 8;   // warps per block in the row-wise kernels
#ifndef KROWS
#define KROWS 4
#endif
constexpr int kRowsPerWarp = KROWS;   // output rows owned by one warp

}  // namespace bfm
