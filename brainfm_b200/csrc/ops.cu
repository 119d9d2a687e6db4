// Op-level entry points: one reference function each (see include/bfm.h for the citations).
#include "common.cuh"

namespace bfm {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

// ------------------------------------------------------------------------------------------------
// fast_3D_interp_torch 'linear' on explicit coordinate arrays
// ------------------------------------------------------------------------------------------------
__global__ void k_trilerp_pull(const float *__restrict__ X, int nx, int ny, int nz, int C,
                               const float *__restrict__ I, const float *__restrict__ J,
                               const float *__restrict__ K, int64_t n, float dflt,
                               const float *__restrict__ dflt_dev, float *__restrict__ out) {
    const int bb[6] = {0, 0, 0, nx, ny, nz};
    const float dv = dflt_dev ? *dflt_dev : dflt;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        Taps t = make_taps(I[p], J[p], K[p], bb);
        if (!t.ok) {
            for (int c = 0; c < C; ++c) out[p * C + c] = dv;
            continue;
        }
        for (int c = 0; c < C; ++c) {
            out[p * C + c] = trilerp(t, [&](int x, int y, int z) {
                return __ldg(X + (((int64_t)x * ny + y) * nz + z) * C + c);
            });
        }
    }
}

// Small host -> device copy done by the SMs: reads mapped pinned host memory over PCIe.  Used for the plan arena
// (~16 KB per sample) so that it never queues behind bulk uploads in the copy engine's FIFO.
__global__ void k_upload(uint4 *__restrict__ dst, const uint4 *__restrict__ src, int64_t n) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
        dst[p] = src[p];
}

// torch.nan_to_num in place (Generator/utils.py:305)
__global__ void k_sanitize(float *__restrict__ x, int64_t n) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const float v = x[p];
        if (!(fabsf(v) <= 3.4028234663852886e38f)) x[p] = nan_to_num(v);
    }
}

template <typename T>
__global__ void k_nearest_pull(const T *__restrict__ X, int nx, int ny, int nz, int C,
                               const float *__restrict__ I, const float *__restrict__ J,
                               const float *__restrict__ K, int64_t n, T *__restrict__ out) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        // torch.round == round-half-even == rintf; conversion saturates, then clamp (utils.py:125-133)
        int x = min(max(__float2int_rn(I[p]), 0), nx - 1);
        int y = min(max(__float2int_rn(J[p]), 0), ny - 1);
        int z = min(max(__float2int_rn(K[p]), 0), nz - 1);
        const T *src = X + (((int64_t)x * ny + y) * nz + z) * C;
        for (int c = 0; c < C; ++c) out[p * C + c] = src[c];
    }
}

// ------------------------------------------------------------------------------------------------
// myzoom_torch: warp per output row (i,j); first two passes once per source z node
// ------------------------------------------------------------------------------------------------
__global__ void k_zoom_linear(const float *__restrict__ X, int b, int c, int C, bfm_zoom_tab t, int A, int B,
                              int Cc, float *__restrict__ out) {
    extern __shared__ float smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rowlen = c * C;
    float *sm = smem + warp * rowlen;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
    if (row >= (int64_t)A * B) return;
    const int i = (int)(row / B), j = (int)(row % B);
    row_zoom_setup(X, b, c, C, t, i, j, sm, lane);
    __syncwarp();
    float *o = out + row * (int64_t)Cc * C;
    for (int q = lane; q < Cc * C; q += 32) {
        const int k = q / C, ch = q - k * C;
        o[q] = lerp_rn(t.wl[2][k], sm[t.lo[2][k] * C + ch], t.wh[2][k], sm[t.hi[2][k] * C + ch]);
    }
}

// ------------------------------------------------------------------------------------------------
// banded linear map along one axis (+ optional noise epilogue)
// ------------------------------------------------------------------------------------------------
__global__ void k_band_axis(const float *__restrict__ in, float *__restrict__ out, int n0, int n1, int n2,
                            int axis, int n_out, const int *__restrict__ start, const float *__restrict__ w,
                            int T, float noise_std, const float *__restrict__ eps, uint64_t seed) {
    int o0 = n0, o1 = n1, o2 = n2;
    (axis == 0 ? o0 : axis == 1 ? o1 : o2) = n_out;
    const int64_t total = (int64_t)o0 * o1 * o2;
    const int64_t stride = axis == 0 ? (int64_t)n1 * n2 : axis == 1 ? n2 : 1;
    const int n_in = axis == 0 ? n0 : axis == 1 ? n1 : n2;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(p % o2);
        const int j = (int)((p / o2) % o1);
        const int i = (int)(p / ((int64_t)o1 * o2));
        const int q = axis == 0 ? i : axis == 1 ? j : k;
        const int s = start[q];
        int64_t base;
        if (axis == 0) base = ((int64_t)s * n1 + j) * n2 + k;
        else if (axis == 1) base = ((int64_t)i * n1 + s) * n2 + k;
        else base = ((int64_t)i * n1 + j) * n2 + s;
        const float *wr = w + (int64_t)q * T;
        float acc = 0.f;
        for (int t = 0; t < T; ++t) {
            const int src = s + t;
            if (src >= 0 && src < n_in) acc = fmaf(__ldg(wr + t), __ldg(in + base + t * stride), acc);
        }
        if (noise_std >= 0.f) {
            float e;
            if (eps) e = eps[p];
            else {
                float4 g = philox_normal4(seed, 1u, (uint64_t)p >> 2);
                const int r = (int)(p & 3);
                e = r == 0 ? g.x : r == 1 ? g.y : r == 2 ? g.z : g.w;
            }
            acc = __fadd_rn(acc, __fmul_rn(noise_std, e));
            if (acc < 0.f) acc = 0.f;
        }
        out[p] = acc;
    }
}

// Axis 0 / 1 with n2 % 4 == 0: one thread = 4 consecutive z outputs (128-bit loads and stores; the tap weight is
// uniform across the four).  Same tap order and the same noise mapping as k_band_axis: bit-identical results.
__global__ void __launch_bounds__(256) k_band_axis4(const float *__restrict__ in, float *__restrict__ out, int n0,
                                                    int n1, int n2, int axis, int n_out,
                                                    const int *__restrict__ start, const float *__restrict__ w, int T,
                                                    float noise_std, const float *__restrict__ eps, uint64_t seed) {
    const int o0 = axis == 0 ? n_out : n0, o1 = axis == 1 ? n_out : n1, o2v = n2 >> 2;
    const int64_t total = (int64_t)o0 * o1 * o2v;
    const int64_t stride = axis == 0 ? (int64_t)n1 * n2 : n2;
    const int n_in = axis == 0 ? n0 : n1;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int kv = (int)(p % o2v);
        const int64_t r = p / o2v;
        const int j = (int)(r % o1), i = (int)(r / o1);
        const int q = axis == 0 ? i : j;
        const int s = __ldg(start + q);
        const int64_t base = axis == 0 ? ((int64_t)s * n1 + j) * n2 + 4 * kv : ((int64_t)i * n1 + s) * n2 + 4 * kv;
        const float *wr = w + (int64_t)q * T;
        const int t0 = max(0, -s), t1 = min(T, n_in - s);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 4
        for (int t = t0; t < t1; ++t) {
            const float wt = __ldg(wr + t);
            const float4 x = __ldg((const float4 *)(in + base + t * stride));
            a0 = fmaf(wt, x.x, a0); a1 = fmaf(wt, x.y, a1); a2 = fmaf(wt, x.z, a2); a3 = fmaf(wt, x.w, a3);
        }
        const int64_t op = ((int64_t)i * o1 + j) * n2 + 4 * kv;
        if (noise_std >= 0.f) {
            float4 e;
            if (eps) e = __ldg((const float4 *)(eps + op));
            else e = philox_normal4(seed, 1u, (uint64_t)op >> 2);
            a0 = __fadd_rn(a0, __fmul_rn(noise_std, e.x)); a1 = __fadd_rn(a1, __fmul_rn(noise_std, e.y));
            a2 = __fadd_rn(a2, __fmul_rn(noise_std, e.z)); a3 = __fadd_rn(a3, __fmul_rn(noise_std, e.w));
            a0 = a0 < 0.f ? 0.f : a0; a1 = a1 < 0.f ? 0.f : a1; a2 = a2 < 0.f ? 0.f : a2; a3 = a3 < 0.f ? 0.f : a3;
        }
        *(float4 *)(out + op) = make_float4(a0, a1, a2, a3);
    }
}

__global__ void k_blur_axis(const float *__restrict__ in, float *__restrict__ out, int n0, int n1, int n2,
                            int axis, const float *__restrict__ taps, int half) {
    const int64_t total = (int64_t)n0 * n1 * n2;
    const int64_t stride = axis == 0 ? (int64_t)n1 * n2 : axis == 1 ? n2 : 1;
    const int n_in = axis == 0 ? n0 : axis == 1 ? n1 : n2;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(p % n2);
        const int j = (int)((p / n2) % n1);
        const int i = (int)(p / ((int64_t)n1 * n2));
        const int q = axis == 0 ? i : axis == 1 ? j : k;
        float acc = 0.f;
        for (int t = -half; t <= half; ++t) {
            const int src = q + t;
            if (src >= 0 && src < n_in) acc = fmaf(__ldg(taps + t + half), __ldg(in + p + t * stride), acc);
        }
        out[p] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// reductions / elementwise
// ------------------------------------------------------------------------------------------------
__global__ void k_minmax_init(int *mm) {
    mm[0] = f2ord(INFINITY);
    mm[1] = f2ord(-INFINITY);
}
__global__ void k_minmax(const float *__restrict__ x, int64_t n, int *mm) {
    __shared__ float red[8][2];
    float lo = INFINITY, hi = -INFINITY;
    const int64_t n4 = ((uintptr_t)x & 15) == 0 ? n / 4 : 0;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n4; p += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg((const float4 *)x + p);
        lo = fminf(fminf(lo, v.x), fminf(fminf(v.y, v.z), v.w));
        hi = fmaxf(fmaxf(hi, v.x), fmaxf(fmaxf(v.y, v.z), v.w));
    }
    for (int64_t p = n4 * 4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const float v = x[p];
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
    lo = warp_min(lo);
    hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = lo; red[threadIdx.x >> 5][1] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = fminf(lo, red[w][0]); hi = fmaxf(hi, red[w][1]); }
        atomicMin(mm, f2ord(lo));
        atomicMax(mm + 1, f2ord(hi));
    }
}
__global__ void k_minmax_decode(int *mm) {
    float lo = ord2f(mm[0]), hi = ord2f(mm[1]);
    ((float *)mm)[0] = lo;
    ((float *)mm)[1] = hi;
}

__global__ void k_shift_scale_flip(const float *__restrict__ x, float *__restrict__ out, int nx, int64_t plane,
                                   const float *sub_dev, const float *div_dev, float post, int flip) {
    const float sub = sub_dev ? *sub_dev : 0.f;
    float div = div_dev ? *div_dev : 1.f;
    // `Idef -= min; Idef /= max(Idef)`: the divisor is the maximum AFTER the subtraction
    if (sub_dev && div_dev) div = __fsub_rn(div, sub);
    const int64_t total = (int64_t)nx * plane;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int i = (int)(p / plane);
        const int64_t r = p - (int64_t)i * plane;
        float v = x[p];
        if (sub_dev) v = __fsub_rn(v, sub);
        if (div_dev) v = __fdiv_rn(v, div);
        if (post != 1.f) v = __fmul_rn(v, post);
        const int oi = flip ? nx - 1 - i : i;
        out[(int64_t)oi * plane + r] = v;
    }
}

// 128-bit form: blockIdx.y = plane, no 64-bit divisions (plane % 4 == 0, 16-byte aligned pointers)
__global__ void __launch_bounds__(256) k_shift_scale_flip4(const float4 *__restrict__ x, float4 *__restrict__ out,
                                                           int nx, int plane4, const float *sub_dev,
                                                           const float *div_dev, float post, int flip) {
    const float sub = sub_dev ? *sub_dev : 0.f;
    float div = div_dev ? *div_dev : 1.f;
    if (sub_dev && div_dev) div = __fsub_rn(div, sub);
    for (int i = blockIdx.y; i < nx; i += gridDim.y) {
        const float4 *src = x + (int64_t)i * plane4;
        float4 *dst = out + (int64_t)(flip ? nx - 1 - i : i) * plane4;
        for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < plane4; r += gridDim.x * blockDim.x) {
            float4 v = __ldg(src + r);
            float a[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (sub_dev) a[c] = __fsub_rn(a[c], sub);
                if (div_dev) a[c] = __fdiv_rn(a[c], div);
                if (post != 1.f) a[c] = __fmul_rn(a[c], post);
            }
            dst[r] = make_float4(a[0], a[1], a[2], a[3]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// deformation-driven kernels (warp per output row)
// ------------------------------------------------------------------------------------------------
__global__ void k_bbox_init(int *bb) {
    if (threadIdx.x < 3) bb[threadIdx.x] = 0x7f7fffff;     // +FLT_MAX bits (coords are >= 0)
    else if (threadIdx.x < 6) bb[threadIdx.x] = 0;
}

__global__ void __launch_bounds__(kRowWarps * 32)
k_deform_bbox(const __grid_constant__ bfm_deform d, int *bb_bits, int fstride) {
    extern __shared__ float smem[];
    __shared__ float red[kRowWarps][6];
    const DefRegs g = load_def(d);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *smF = smem + warp * kRowsPerWarp * fstride;
    const int n_rows = g.s0 * g.s1;
    const int row0 = (blockIdx.x * kRowWarps + warp) * kRowsPerWarp;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {0.f, 0.f, 0.f};
    if (row0 < n_rows) {
        deform_rows<kRowsPerWarp>(d, g, smF, row0, n_rows, lane, [](int) {},
                                  [&](int, int, int, int, int, float px, float py, float pz) {
                                      lo[0] = fminf(lo[0], px); hi[0] = fmaxf(hi[0], px);
                                      lo[1] = fminf(lo[1], py); hi[1] = fmaxf(hi[1], py);
                                      lo[2] = fminf(lo[2], pz); hi[2] = fmaxf(hi[2], pz);
                                  });
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        lo[a] = warp_min(lo[a]);
        hi[a] = warp_max(hi[a]);
    }
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            red[warp][a] = lo[a];
            red[warp][3 + a] = hi[a];
        }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        float v = red[0][threadIdx.x];
        for (int w = 1; w < kRowWarps; ++w)
            v = threadIdx.x < 3 ? fminf(v, red[w][threadIdx.x]) : fmaxf(v, red[w][threadIdx.x]);
        // all coordinates are >= 0 after the clamp, so the raw bit pattern is order preserving
        if (threadIdx.x < 3) atomicMin(bb_bits + threadIdx.x, __float_as_int(v));
        else atomicMax(bb_bits + threadIdx.x, __float_as_int(v));
    }
}

__global__ void k_bbox_finish(int *bb) {
    // lo = floor(min), hi = 1 + ceil(max)   (datasets.py:288-293)
    if (threadIdx.x < 3) bb[threadIdx.x] = (int)floorf(__int_as_float(bb[threadIdx.x]));
    else if (threadIdx.x < 6) bb[threadIdx.x] = 1 + (int)ceilf(__int_as_float(bb[threadIdx.x]));
}

__global__ void __launch_bounds__(kRowWarps * 32)
k_deform_coords(const __grid_constant__ bfm_deform d, const int *__restrict__ bb, float *__restrict__ out, int fstride) {
    extern __shared__ float smem[];
    const DefRegs g = load_def(d);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *smF = smem + warp * kRowsPerWarp * fstride;
    const int n_rows = g.s0 * g.s1;
    const int row0 = (blockIdx.x * kRowWarps + warp) * kRowsPerWarp;
    if (row0 >= n_rows) return;
    const int64_t N = (int64_t)n_rows * g.s2;
    const float l0 = (float)bb[0], l1 = (float)bb[1], l2 = (float)bb[2];
    deform_rows<kRowsPerWarp>(d, g, smF, row0, n_rows, lane, [](int) {},
                              [&](int, int row, int, int, int k, float px, float py, float pz) {
                                  const int64_t p = (int64_t)row * g.s2 + k;
                                  out[p] = __fsub_rn(px, l0);
                                  out[N + p] = __fsub_rn(py, l1);
                                  out[2 * N + p] = __fsub_rn(pz, l2);
                              });
}

// crop maximum of (nan_to_num(src) - mean) / scale  (default_value_linear_mode == 'max')
__global__ void k_crop_max_init(float *m) { *(int *)m = f2ord(-INFINITY); }
__global__ void k_crop_max(const float *__restrict__ src, int n1, int n2, const int *__restrict__ bb, float mean,
                           float scale, float *m) {
    __shared__ float red[8];
    const int c0 = bb[3] - bb[0], c1 = bb[4] - bb[1], c2 = bb[5] - bb[2];
    const int64_t total = (int64_t)c0 * c1 * c2;
    float hi = -INFINITY;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int z = (int)(p % c2), y = (int)((p / c2) % c1), x = (int)(p / ((int64_t)c1 * c2));
        float v = nan_to_num(src[((int64_t)(x + bb[0]) * n1 + (y + bb[1])) * n2 + z + bb[2]]);
        v = __fdiv_rn(__fsub_rn(v, mean), scale);
        hi = fmaxf(hi, v);
    }
    hi = warp_max(hi);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = hi;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) hi = fmaxf(hi, red[w]);
        atomicMax((int *)m, f2ord(hi));
    }
}
__global__ void k_crop_max_decode(float *m) { *m = ord2f(*(int *)m); }

// read_and_deform: warp of a full source volume; optional fused min/max of the result (for the
// `Idef -= min; Idef /= max` normalisation of read_and_deform_image)
__global__ void __launch_bounds__(kRowWarps * 32)
k_warp_volume(const __grid_constant__ bfm_deform d, const int *__restrict__ bb, const float *__restrict__ src,
              float mean, float scale, const float *__restrict__ dflt_dev, float *__restrict__ out, int *mm_ord,
              int fstride) {
    extern __shared__ float smem[];
    __shared__ float red[kRowWarps][2];
    const DefRegs g = load_def(d);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *smF = smem + warp * kRowsPerWarp * fstride;
    const int n_rows = g.s0 * g.s1;
    const int row0 = (blockIdx.x * kRowWarps + warp) * kRowsPerWarp;
    float vlo = INFINITY, vhi = -INFINITY;
    if (row0 < n_rows) {
        const float dv = dflt_dev ? *dflt_dev : 0.f;
        const BoxRegs box = load_box(bb, d.src[1], d.src[2]);
        const bool plain = (mean == 0.f && scale == 1.f);
        deform_rows<kRowsPerWarp>(d, g, smF, row0, n_rows, lane, [](int) {},
                                  [&](int, int row, int, int, int k, float px, float py, float pz) {
                                      const Taps32 t = make_taps32(px, py, pz, box);
                                      float v = trilerp32(t, [&](int e) {
                                          const float s = nan_to_num(__ldg(src + e));
                                          return plain ? s : __fdiv_rn(__fsub_rn(s, mean), scale);
                                      });
                                      v = t.ok ? v : dv;
                                      out[row * g.s2 + k] = v;
                                      vlo = fminf(vlo, v);
                                      vhi = fmaxf(vhi, v);
                                  });
    }
    if (mm_ord) {
        vlo = warp_min(vlo);
        vhi = warp_max(vhi);
        if (lane == 0) { red[warp][0] = vlo; red[warp][1] = vhi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < kRowWarps; ++w) { vlo = fminf(vlo, red[w][0]); vhi = fmaxf(vhi, red[w][1]); }
            atomicMin(mm_ord, f2ord(vlo));
            atomicMax(mm_ord + 1, f2ord(vhi));
        }
    }
}

__global__ void __launch_bounds__(kRowWarps * 32)
k_label_warp(const __grid_constant__ bfm_deform d, const int *__restrict__ bb, const int32_t *__restrict__ labels,
             const int32_t *__restrict__ lut, int lut_n, int n_classes, const int32_t *__restrict__ vflip, int flip,
             float *__restrict__ onehot, int32_t *__restrict__ label_out, int fstride) {
    extern __shared__ float smem[];
    const DefRegs g = load_def(d);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *smF = smem + warp * kRowsPerWarp * fstride;
    const int n_rows = g.s0 * g.s1;
    const int row0 = (blockIdx.x * kRowWarps + warp) * kRowsPerWarp;
    if (row0 >= n_rows) return;
    const int n1 = d.src[1], n2 = d.src[2];
    const int b0 = bb[0], b1 = bb[1], b2 = bb[2];
    const int cx = bb[3] - b0, cy = bb[4] - b1, cz = bb[5] - b2;
    const float l0 = (float)b0, l1 = (float)b1, l2 = (float)b2;
    const int64_t N = (int64_t)n_rows * g.s2;
    deform_rows<kRowsPerWarp>(
        d, g, smF, row0, n_rows, lane, [](int) {},
        [&](int, int row, int i, int j, int k, float px, float py, float pz) {
            // nearest: round-half-even of the bbox-relative coordinate, clamped to the crop (utils.py:124-138)
            const int x = min(max(__float2int_rn(__fsub_rn(px, l0)), 0), cx - 1) + b0;
            const int y = min(max(__float2int_rn(__fsub_rn(py, l1)), 0), cy - 1) + b1;
            const int z = min(max(__float2int_rn(__fsub_rn(pz, l2)), 0), cz - 1) + b2;
            const int32_t lab = __ldg(labels + (x * n1 + y) * n2 + z);
            const int32_t cls = (lab >= 0 && lab < lut_n) ? __ldg(lut + lab) : 0;
            if (label_out) label_out[row * g.s2 + k] = cls;
            if (onehot) {
                const int64_t p = (int64_t)((flip ? g.s0 - 1 - i : i) * g.s1 + j) * g.s2 + k;
                // flipped output channel c holds input channel vflip[c]
                for (int c = 0; c < n_classes; ++c) {
                    const int srcc = flip ? __ldg(vflip + c) : c;
                    onehot[(int64_t)c * N + p] = (srcc == cls) ? 1.f : 0.f;
                }
            }
        });
}

__global__ void k_svf_step(const float *__restrict__ Fin, float *__restrict__ Fout, int sx, int sy, int sz) {
    const int bb[6] = {0, 0, 0, sx, sy, sz};
    const int64_t total = (int64_t)sx * sy * sz;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(p % sz), j = (int)((p / sz) % sy), i = (int)(p / ((int64_t)sy * sz));
        const float f0 = Fin[p * 3], f1 = Fin[p * 3 + 1], f2 = Fin[p * 3 + 2];
        Taps t = make_taps(__fadd_rn((float)i, f0), __fadd_rn((float)j, f1), __fadd_rn((float)k, f2), bb);
        float g[3] = {0.f, 0.f, 0.f};
        if (t.ok) {
#pragma unroll
            for (int c = 0; c < 3; ++c)
                g[c] = trilerp(t, [&](int x, int y, int z) { return __ldg(Fin + (((int64_t)x * sy + y) * sz + z) * 3 + c); });
        }
        Fout[p * 3] = __fadd_rn(f0, g[0]);
        Fout[p * 3 + 1] = __fadd_rn(f1, g[1]);
        Fout[p * 3 + 2] = __fadd_rn(f2, g[2]);
    }
}

// The same step on 16-byte {f0, f1, f2, 0} records: one 128-bit load per trilinear tap instead of three 32-bit loads
// 12 bytes apart (bfm_svf_integrate keeps the field in this layout between its steps; identical arithmetic).
__global__ void __launch_bounds__(256) k_svf_step4(const float4 *__restrict__ Fin, float4 *__restrict__ Fout, int sx,
                                                   int sy, int sz) {
    const int bb[6] = {0, 0, 0, sx, sy, sz};
    const int64_t total = (int64_t)sx * sy * sz;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(p % sz), j = (int)((p / sz) % sy), i = (int)(p / ((int64_t)sy * sz));
        const float4 f = __ldg(Fin + p);
        Taps t = make_taps(__fadd_rn((float)i, f.x), __fadd_rn((float)j, f.y), __fadd_rn((float)k, f.z), bb);
        float g[3] = {0.f, 0.f, 0.f};
        if (t.ok) {
            g[0] = trilerp(t, [&](int x, int y, int z) { return __ldg(Fin + ((int64_t)x * sy + y) * sz + z).x; });
            g[1] = trilerp(t, [&](int x, int y, int z) { return __ldg(Fin + ((int64_t)x * sy + y) * sz + z).y; });
            g[2] = trilerp(t, [&](int x, int y, int z) { return __ldg(Fin + ((int64_t)x * sy + y) * sz + z).z; });
        }
        Fout[p] = make_float4(__fadd_rn(f.x, g[0]), __fadd_rn(f.y, g[1]), __fadd_rn(f.z, g[2]), 0.f);
    }
}
__global__ void __launch_bounds__(256) k_svf_pack(const float *__restrict__ src, float4 *__restrict__ dst, int64_t n,
                                                  float scale) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
        dst[q] = make_float4(__fmul_rn(src[q * 3], scale), __fmul_rn(src[q * 3 + 1], scale), __fmul_rn(src[q * 3 + 2], scale), 0.f);
}
__global__ void __launch_bounds__(256) k_svf_unpack(const float4 *__restrict__ src, float *__restrict__ dst, int64_t n) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(src + q);
        dst[q * 3] = v.x; dst[q * 3 + 1] = v.y; dst[q * 3 + 2] = v.z;
    }
}

// Cached 'f32' volume <- a volume in its stored dtype, converted on the device: dst = nan_to_num(src * slope + inter)
// (nib get_fdata scaling + torch.nan_to_num, Generator/utils.py:304-305).  16 source elements per thread for the
// narrow dtypes so that loads and stores are both 128-bit.
template <typename T>
__global__ void __launch_bounds__(256) k_ingest(float *__restrict__ dst, const T *__restrict__ src, int64_t n,
                                                float slope, float inter) {
    constexpr int V = 16 / sizeof(T);                       // elements per 128-bit load
    const int64_t groups = n / V;
    const bool aligned = (((uintptr_t)src & 15) == 0) && (((uintptr_t)dst & 15) == 0);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if (aligned) {
        for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += stride) {
            const uint4 raw = __ldg((const uint4 *)src + g);
            const T *e = (const T *)&raw;
            float v[V];
#pragma unroll
            for (int q = 0; q < V; ++q) v[q] = nan_to_num(__fadd_rn(__fmul_rn((float)e[q], slope), inter));
            float4 *o = (float4 *)(dst + g * V);
#pragma unroll
            for (int q = 0; q < V / 4; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
    }
    const int64_t done = aligned ? groups * V : 0;
    for (int64_t p = done + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride)
        dst[p] = nan_to_num(__fadd_rn(__fmul_rn((float)src[p], slope), inter));
}

// out[4g .. 4g+3] = philox_normal4(seed, stream, first_group + g): the N(0,1) source of the fused chain, exposed
// so that its distribution can be tested directly (tests/test_noise_gpu.py).
__global__ void __launch_bounds__(256) k_philox_normal(float *__restrict__ out, int64_t n, uint64_t seed,
                                                       uint32_t stream, uint64_t first_group) {
    const int64_t groups = (n + 3) / 4;
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (int64_t)gridDim.x * blockDim.x) {
        const float4 e = philox_normal4(seed, stream, first_group + (uint64_t)g);
        const float v[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (4 * g + q < n) out[4 * g + q] = v[q];
    }
}

// x[i] = max(0, x[i] + noise_std * eps[first + i]), eps = the chain's stream-`stream` normal sequence of `seed`
// (element e = component e & 3 of Philox group e >> 2): add_noise (Generator/utils.py:633-638) for a PART of a volume
// whose first element has the absolute index `first` -- a slab of a volume cut across GPUs draws exactly the numbers
// the whole volume would.
__global__ void __launch_bounds__(256) k_add_noise_at(float *__restrict__ x, int64_t n, float noise_std, uint64_t seed,
                                                      uint32_t stream, int64_t first) {
    const int64_t g0 = first >> 2, g1 = (first + n - 1) >> 2;
    for (int64_t g = g0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g <= g1; g += (int64_t)gridDim.x * blockDim.x) {
        const float4 e = philox_normal4(seed, stream, (uint64_t)g);
        const float ev[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int64_t i = 4 * g + c - first;
            if (i >= 0 && i < n) {
                const float v = __fadd_rn(x[i], __fmul_rn(noise_std, ev[c]));
                x[i] = v < 0.f ? 0.f : v;
            }
        }
    }
}

static inline int grid_for(int64_t n, int block = 256) {
    int64_t g = (n + block - 1) / block;
    const int64_t cap = 148LL * 32;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}
static inline int check_deform(const bfm_deform *d) {
    if (!d) return fail(BFM_E_INVALID, "%s", "null deformation");
    for (int a = 0; a < 3; ++a)
        if (d->size[a] <= 0 || d->src[a] <= 0) return fail(BFM_E_INVALID, "%s", "non-positive size");
    if (d->fsmall && !d->F_full && (d->fs[2] > kMaxSmallZ || d->fs[2] <= 0))
        return fail(BFM_E_UNSUPPORTED, "%s", "small-grid z extent exceeds kMaxSmallZ");
    if ((int64_t)d->src[0] * d->src[1] * d->src[2] >= (1LL << 31) ||
        (int64_t)d->size[0] * d->size[1] * d->size[2] >= (1LL << 31))
        return fail(BFM_E_UNSUPPORTED, "%s", "volumes of 2^31 voxels or more are not supported");
    return BFM_OK;
}
}  // namespace bfm

using namespace bfm;

extern "C" {

int bfm_abi_version(void) { return BFM_ABI_VERSION; }

int bfm_ingest_volume(float *dst, const void *src, int src_dtype, int64_t n, float slope, float inter, void *stream) {
    BFM_REQUIRE(dst && src && n >= 0, "bfm_ingest_volume: null pointer");
    if (n == 0) return BFM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (src_dtype) {
        case 0: k_ingest<uint8_t><<<grid_for((n + 15) / 16), 256, 0, st>>>(dst, (const uint8_t *)src, n, slope, inter); break;
        case 1: k_ingest<int16_t><<<grid_for((n + 7) / 8), 256, 0, st>>>(dst, (const int16_t *)src, n, slope, inter); break;
        case 2: k_ingest<int32_t><<<grid_for((n + 3) / 4), 256, 0, st>>>(dst, (const int32_t *)src, n, slope, inter); break;
        case 3: k_ingest<float><<<grid_for((n + 3) / 4), 256, 0, st>>>(dst, (const float *)src, n, slope, inter); break;
        case 4: k_ingest<int8_t><<<grid_for((n + 15) / 16), 256, 0, st>>>(dst, (const int8_t *)src, n, slope, inter); break;
        default: return fail(BFM_E_INVALID, "%s", "bfm_ingest_volume: src_dtype must be 0 (u8), 1 (i16), 2 (i32), 3 (f32) or 4 (i8)");
    }
    return check_launch("bfm_ingest_volume");
}

int bfm_add_noise_at(float *x, int64_t n, float noise_std, uint64_t seed, uint32_t stream_id, int64_t first_element,
                     void *stream) {
    BFM_REQUIRE(x && n >= 0 && first_element >= 0, "bfm_add_noise_at: bad argument");
    if (n == 0) return BFM_OK;
    k_add_noise_at<<<grid_for((n + 3) / 4 + 1), 256, 0, (cudaStream_t)stream>>>(x, n, noise_std, seed, stream_id,
                                                                                  first_element);
    return check_launch("bfm_add_noise_at");
}

int bfm_philox_normal(float *out, int64_t n, uint64_t seed, uint32_t stream_id, uint64_t first_group, void *stream) {
    BFM_REQUIRE(out && n >= 0, "bfm_philox_normal: null output");
    if (n == 0) return BFM_OK;
    k_philox_normal<<<grid_for((n + 3) / 4), 256, 0, (cudaStream_t)stream>>>(out, n, seed, stream_id, first_group);
    return check_launch("bfm_philox_normal");
}

int bfm_upload_pinned(void *dst, const void *src_pinned, int64_t nbytes, void *stream) {
    BFM_REQUIRE(dst && src_pinned && nbytes >= 0, "bfm_upload_pinned: null pointer");
    BFM_REQUIRE(((uintptr_t)dst & 15) == 0 && ((uintptr_t)src_pinned & 15) == 0 && (nbytes & 15) == 0,
                "bfm_upload_pinned: addresses and size must be multiples of 16 bytes");
    if (nbytes == 0) return BFM_OK;
    const int64_t n = nbytes / 16;
    const int64_t blocks = (n + 255) / 256;
    k_upload<<<(unsigned)(blocks < 148 * 4 ? blocks : 148 * 4), 256, 0, (cudaStream_t)stream>>>(
        (uint4 *)dst, (const uint4 *)src_pinned, n);
    return check_launch("bfm_upload_pinned");
}

int bfm_sanitize_f32(float *x, int64_t n, void *stream) {
    BFM_REQUIRE(x && n >= 0, "bfm_sanitize_f32: null pointer");
    if (n == 0) return BFM_OK;
    const int64_t blocks = (n + 4 * 256 - 1) / (4 * 256);
    k_sanitize<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, (cudaStream_t)stream>>>(x, n);
    return check_launch("bfm_sanitize_f32");
}
const char *bfm_last_error(void) { return g_err; }
uint64_t bfm_launch_count(void) { return g_launches.load(); }

int bfm_trilerp_pull(const float *X, int nx, int ny, int nz, int C, const float *I, const float *J, const float *K,
                     int64_t n, float default_value, const float *default_dev, float *out, void *stream) {
    BFM_REQUIRE(X && I && J && K && out, "bfm_trilerp_pull: null pointer");
    BFM_REQUIRE(nx > 0 && ny > 0 && nz > 0 && C > 0 && n >= 0, "bfm_trilerp_pull: bad shape");
    if (n == 0) return BFM_OK;
    k_trilerp_pull<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(X, nx, ny, nz, C, I, J, K, n, default_value,
                                                                  default_dev, out);
    return check_launch("bfm_trilerp_pull");
}

int bfm_nearest_pull(const void *X, int elem_size, int nx, int ny, int nz, int C, const float *I, const float *J,
                     const float *K, int64_t n, void *out, void *stream) {
    BFM_REQUIRE(X && I && J && K && out, "bfm_nearest_pull: null pointer");
    BFM_REQUIRE(nx > 0 && ny > 0 && nz > 0 && C > 0 && n >= 0, "bfm_nearest_pull: bad shape");
    if (n == 0) return BFM_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (elem_size == 1)
        k_nearest_pull<uint8_t><<<grid_for(n), 256, 0, s>>>((const uint8_t *)X, nx, ny, nz, C, I, J, K, n, (uint8_t *)out);
    else if (elem_size == 4)
        k_nearest_pull<uint32_t><<<grid_for(n), 256, 0, s>>>((const uint32_t *)X, nx, ny, nz, C, I, J, K, n, (uint32_t *)out);
    else if (elem_size == 8)
        k_nearest_pull<uint64_t><<<grid_for(n), 256, 0, s>>>((const uint64_t *)X, nx, ny, nz, C, I, J, K, n, (uint64_t *)out);
    else
        return fail(BFM_E_UNSUPPORTED, "%s", "bfm_nearest_pull: elem_size must be 1, 4 or 8");
    return check_launch("bfm_nearest_pull");
}

int bfm_zoom_linear(const float *X, int a, int b, int c, int C, const int *lo0, const int *hi0, const float *wl0,
                    const float *wh0, int A, const int *lo1, const int *hi1, const float *wl1, const float *wh1, int B,
                    const int *lo2, const int *hi2, const float *wl2, const float *wh2, int Cc, float *out,
                    void *stream) {
    BFM_REQUIRE(X && out && lo0 && lo1 && lo2, "bfm_zoom_linear: null pointer");
    BFM_REQUIRE(a > 0 && b > 0 && c > 0 && C > 0 && A > 0 && B > 0 && Cc > 0, "bfm_zoom_linear: bad shape");
    bfm_zoom_tab t;
    t.lo[0] = lo0; t.hi[0] = hi0; t.wl[0] = wl0; t.wh[0] = wh0;
    t.lo[1] = lo1; t.hi[1] = hi1; t.wl[1] = wl1; t.wh[1] = wh1;
    t.lo[2] = lo2; t.hi[2] = hi2; t.wl[2] = wl2; t.wh[2] = wh2;
    const int warps = 8;
    const size_t smem = (size_t)warps * c * C * sizeof(float);
    if (smem > 200 * 1024) return fail(BFM_E_UNSUPPORTED, "%s", "bfm_zoom_linear: source row too long for shared memory");
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(k_zoom_linear, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t rows = (int64_t)A * B;
    k_zoom_linear<<<(unsigned)((rows + warps - 1) / warps), warps * 32, smem, (cudaStream_t)stream>>>(X, b, c, C, t, A,
                                                                                                  B, Cc, out);
    return check_launch("bfm_zoom_linear");
}

int bfm_blur_axis(const float *in, float *out, int nx, int ny, int nz, int axis, const float *taps, int half,
                  void *stream) {
    BFM_REQUIRE(in && out && taps, "bfm_blur_axis: null pointer");
    BFM_REQUIRE(nx > 0 && ny > 0 && nz > 0 && axis >= 0 && axis < 3 && half >= 0, "bfm_blur_axis: bad argument");
    const int64_t n = (int64_t)nx * ny * nz;
    k_blur_axis<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(in, out, nx, ny, nz, axis, taps, half);
    return check_launch("bfm_blur_axis");
}

int bfm_band_axis(const float *in, float *out, const int *in_shape, int axis, int n_out, const int *start,
                  const float *w, int T, float noise_std, const float *eps, uint64_t seed, void *stream) {
    BFM_REQUIRE(in && out && in_shape && start && w, "bfm_band_axis: null pointer");
    BFM_REQUIRE(axis >= 0 && axis < 3 && n_out > 0 && T > 0, "bfm_band_axis: bad argument");
    int o[3] = {in_shape[0], in_shape[1], in_shape[2]};
    o[axis] = n_out;
    const int64_t n = (int64_t)o[0] * o[1] * o[2];
    if (axis != 2 && (in_shape[2] & 3) == 0 && ((uintptr_t)in % 16) == 0 && ((uintptr_t)out % 16) == 0 &&
        (!eps || ((uintptr_t)eps % 16) == 0)) {
        k_band_axis4<<<grid_for(n / 4), 256, 0, (cudaStream_t)stream>>>(in, out, in_shape[0], in_shape[1], in_shape[2],
                                                                        axis, n_out, start, w, T, noise_std, eps, seed);
        return check_launch("bfm_band_axis");
    }
    k_band_axis<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(in, out, in_shape[0], in_shape[1], in_shape[2], axis,
                                                               n_out, start, w, T, noise_std, eps, seed);
    return check_launch("bfm_band_axis");
}

int bfm_minmax(const float *x, int64_t n, float *minmax_dev, void *stream) {
    BFM_REQUIRE(x && minmax_dev && n > 0, "bfm_minmax: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    k_minmax_init<<<1, 1, 0, s>>>((int *)minmax_dev);
    k_minmax<<<148 * 4, 256, 0, s>>>(x, n, (int *)minmax_dev);
    k_minmax_decode<<<1, 1, 0, s>>>((int *)minmax_dev);
    g_launches.fetch_add(2);
    return check_launch("bfm_minmax");
}

int bfm_shift_scale_flip(const float *x, float *out, int nx, int64_t plane, const float *sub_dev,
                         const float *div_dev, float post_scale, int flip, void *stream) {
    BFM_REQUIRE(x && out && nx > 0 && plane > 0, "bfm_shift_scale_flip: bad argument");
    BFM_REQUIRE(!(flip && x == out), "bfm_shift_scale_flip: in-place flip is not supported");
    if ((plane & 3) == 0 && plane / 4 < (1LL << 30) && ((uintptr_t)x % 16) == 0 && ((uintptr_t)out % 16) == 0) {
        const int plane4 = (int)(plane / 4);
        const unsigned gx = (unsigned)min((int64_t)64, ((int64_t)plane4 + 255) / 256);
        const unsigned gy = (unsigned)min(nx, 65535);
        k_shift_scale_flip4<<<dim3(gx, gy), 256, 0, (cudaStream_t)stream>>>((const float4 *)x, (float4 *)out, nx, plane4,
                                                                            sub_dev, div_dev, post_scale, flip);
        return check_launch("bfm_shift_scale_flip");
    }
    k_shift_scale_flip<<<grid_for((int64_t)nx * plane), 256, 0, (cudaStream_t)stream>>>(x, out, nx, plane, sub_dev,
                                                                                        div_dev, post_scale, flip);
    return check_launch("bfm_shift_scale_flip");
}

static inline unsigned row_blocks(const bfm_deform *d) {
    const int64_t rows = (int64_t)d->size[0] * d->size[1];
    const int per = kRowWarps * kRowsPerWarp;
    return (unsigned)((rows + per - 1) / per);
}
static inline int row_fstride(const bfm_deform *d) { return (d->fsmall && !d->F_full) ? d->fs[2] * 3 : 0; }
static inline size_t row_smem(const bfm_deform *d) {
    return (size_t)kRowWarps * kRowsPerWarp * row_fstride(d) * sizeof(float);
}

int bfm_deform_grid(const bfm_deform *d, int *bbox_dev, float *coords_out, void *stream) {
    int rc = check_deform(d);
    if (rc) return rc;
    BFM_REQUIRE(bbox_dev, "bfm_deform_grid: null bbox");
    cudaStream_t s = (cudaStream_t)stream;
    if (!coords_out) {
        k_bbox_init<<<1, 32, 0, s>>>(bbox_dev);
        k_deform_bbox<<<row_blocks(d), kRowWarps * 32, row_smem(d), s>>>(*d, bbox_dev, row_fstride(d));
        k_bbox_finish<<<1, 32, 0, s>>>(bbox_dev);
        g_launches.fetch_add(2);
        return check_launch("bfm_deform_grid(bbox)");
    }
    k_deform_coords<<<row_blocks(d), kRowWarps * 32, row_smem(d), s>>>(*d, bbox_dev, coords_out, row_fstride(d));
    return check_launch("bfm_deform_grid(coords)");
}

int bfm_warp_volume(const bfm_deform *d, const int *bbox_dev, const float *src, float mean, float scale,
                    int default_max, float *scratch_max_dev, float *out, float *minmax_out_dev, void *stream) {
    int rc = check_deform(d);
    if (rc) return rc;
    BFM_REQUIRE(bbox_dev && src && out, "bfm_warp_volume: null pointer");
    BFM_REQUIRE(!default_max || scratch_max_dev, "bfm_warp_volume: default_max needs a scratch scalar");
    cudaStream_t s = (cudaStream_t)stream;
    if (default_max) {
        k_crop_max_init<<<1, 1, 0, s>>>(scratch_max_dev);
        k_crop_max<<<148 * 8, 256, 0, s>>>(src, d->src[1], d->src[2], bbox_dev, mean, scale, scratch_max_dev);
        k_crop_max_decode<<<1, 1, 0, s>>>(scratch_max_dev);
        g_launches.fetch_add(3);
    }
    if (minmax_out_dev) {
        k_minmax_init<<<1, 1, 0, s>>>((int *)minmax_out_dev);
        g_launches.fetch_add(1);
    }
    k_warp_volume<<<row_blocks(d), kRowWarps * 32, row_smem(d), s>>>(
        *d, bbox_dev, src, mean, scale, default_max ? scratch_max_dev : nullptr, out, (int *)minmax_out_dev,
        row_fstride(d));
    if (minmax_out_dev) {
        k_minmax_decode<<<1, 1, 0, s>>>((int *)minmax_out_dev);
        g_launches.fetch_add(1);
    }
    return check_launch("bfm_warp_volume");
}

int bfm_label_warp_onehot(const bfm_deform *d, const int *bbox_dev, const int32_t *labels, const int32_t *lut,
                          int lut_n, int n_classes, const int32_t *vflip, int flip, float *onehot_out,
                          int32_t *label_out, void *stream) {
    int rc = check_deform(d);
    if (rc) return rc;
    BFM_REQUIRE(bbox_dev && labels && lut && (onehot_out || label_out), "bfm_label_warp_onehot: null pointer");
    BFM_REQUIRE(!flip || vflip, "bfm_label_warp_onehot: flip needs vflip");
    k_label_warp<<<row_blocks(d), kRowWarps * 32, row_smem(d), (cudaStream_t)stream>>>(
        *d, bbox_dev, labels, lut, lut_n, n_classes, vflip, flip, onehot_out, label_out, row_fstride(d));
    return check_launch("bfm_label_warp_onehot");
}

int bfm_svf_step(const float *Fin, float *Fout, int sx, int sy, int sz, void *stream) {
    BFM_REQUIRE(Fin && Fout && Fin != Fout, "bfm_svf_step: bad pointers");
    BFM_REQUIRE(sx > 0 && sy > 0 && sz > 0, "bfm_svf_step: bad shape");
    k_svf_step<<<grid_for((int64_t)sx * sy * sz), 256, 0, (cudaStream_t)stream>>>(Fin, Fout, sx, sy, sz);
    return check_launch("bfm_svf_step");
}

int bfm_svf_integrate(const float *F, float *out, int sx, int sy, int sz, int n_steps, float scale, float *scratch,
                      void *stream) {
    BFM_REQUIRE(F && out && scratch && sx > 0 && sy > 0 && sz > 0 && n_steps >= 0, "bfm_svf_integrate: bad argument");
    BFM_REQUIRE(((uintptr_t)scratch % 16) == 0, "bfm_svf_integrate: scratch must be 16-byte aligned");
    const int64_t n = (int64_t)sx * sy * sz;
    cudaStream_t st = (cudaStream_t)stream;
    float4 *cur = (float4 *)scratch, *nxt = cur + n;
    k_svf_pack<<<grid_for(n), 256, 0, st>>>(F, cur, n, scale);
    for (int q = 0; q < n_steps; ++q) {
        k_svf_step4<<<grid_for(n), 256, 0, st>>>(cur, nxt, sx, sy, sz);
        float4 *t = cur; cur = nxt; nxt = t;
    }
    k_svf_unpack<<<grid_for(n), 256, 0, st>>>(cur, out, n);
    g_launches.fetch_add(n_steps + 1);
    return check_launch("bfm_svf_integrate");
}

}  // extern "C"
