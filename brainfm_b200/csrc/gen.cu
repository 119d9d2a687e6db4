// Fused, batched synthesis chain: BaseGen.generate_sample + augment_sample with the stock steps
// (Generator/datasets.py:306-428).  One launch per stage covers the whole batch (blockIdx.y = sample).
//
//   bbox     deform_grid (datasets.py:264-303): bounding box of the deformed grid.  The coordinate field is
//            multilinear between the nodes of the small random grid, so its extrema sit on the voxels next to
//            node boundaries: only those candidates are evaluated (~10^4 instead of 4*10^6 voxels).  The
//            result is accepted only when floor/ceil cannot be changed by the fp32 evaluation error; otherwise
//            (and for SVF-integrated fields) the full scan runs -- the integers are always exact.
//   gmm      mus[Gr] + sigmas[Gr]*eps, clamp (datasets.py:364-372) over the bbox crop -> syn
//   warp     coordinates -> trilinear gather of syn -> mix -> clamp -> gamma -> bias field
//            (utils.py:140-192, datasets.py:379-411, utils.py:568-589) -> i_bf, bflog_out
//   resample blur o downsample as banded per-axis maps + noise (utils.py:83-94, 591-609, 633-638)
//   finish   myzoom_torch back to the grid, global max, normalise, flip (datasets.py:337-352)
//
// Every kernel is instruction-issue bound on B200 (see profiles/), so the code is organised to minimise
// instructions per voxel: the sample descriptor is staged in shared memory, loop invariants live in registers,
// a warp owns several rows so per-k table entries are loaded once, element indices are 32-bit.
#include "common.cuh"

namespace bfm {

constexpr float kBoxDelta = 1e-3f;   // bound on |computed - exact| source coordinate

__device__ __forceinline__ void stage_desc(bfm_gen_sample *dst, const bfm_gen_sample *src) {
    const int n = (int)(sizeof(bfm_gen_sample) / 4);
    const uint32_t *s = (const uint32_t *)src;
    uint32_t *d = (uint32_t *)dst;
    for (int q = threadIdx.x; q < n; q += blockDim.x) d[q] = __ldg(s + q);
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------- bbox
__global__ void k_gen_bbox_init(const bfm_gen_sample *__restrict__ S) {
    int *bb = S[blockIdx.x].bbox;
    if (threadIdx.x < 3) bb[threadIdx.x] = 0x7f7fffff;
    else if (threadIdx.x < 6) bb[threadIdx.x] = 0;
    else if (threadIdx.x == 6) bb[6] = 1;                   // 1 = full scan still required
    if (threadIdx.x == 7) *S[blockIdx.x].maxval = 0.f;      // chain values are >= 0 after the noise clamp
}

// all three zoom passes for one voxel -- same operations as the row-wise evaluation
__device__ __forceinline__ void field_direct(const bfm_deform &d, int i, int j, int k, float &f0, float &f1, float &f2) {
    const bfm_zoom_tab &t = d.ftab;
    const int n1 = d.fs[1], n2 = d.fs[2];
    const int x0 = t.lo[0][i], x1 = t.hi[0][i], y0 = t.lo[1][j], y1 = t.hi[1][j], z0 = t.lo[2][k], z1 = t.hi[2][k];
    const float wx0 = t.wl[0][i], wx1 = t.wh[0][i], wy0 = t.wl[1][j], wy1 = t.wh[1][j], wz0 = t.wl[2][k], wz1 = t.wh[2][k];
    float out[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        auto at = [&](int x, int y, int z) { return __ldg(d.fsmall + ((x * n1 + y) * n2 + z) * 3 + c); };
        const float a0 = lerp_rn(wx0, at(x0, y0, z0), wx1, at(x1, y0, z0));
        const float b0 = lerp_rn(wx0, at(x0, y1, z0), wx1, at(x1, y1, z0));
        const float a1 = lerp_rn(wx0, at(x0, y0, z1), wx1, at(x1, y0, z1));
        const float b1 = lerp_rn(wx0, at(x0, y1, z1), wx1, at(x1, y1, z1));
        const float c0 = lerp_rn(wy0, a0, wy1, b0);
        const float c1 = lerp_rn(wy0, a1, wy1, b1);
        out[c] = lerp_rn(wz0, c0, wz1, c1);
    }
    f0 = out[0]; f1 = d.photo ? 0.f : out[1]; f2 = out[2];
}

__device__ __forceinline__ void block_minmax_atomic(float lo[3], float hi[3], int *bb_bits) {
    __shared__ float red[32][6];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
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
        for (int w = 1; w < nw; ++w) v = threadIdx.x < 3 ? fminf(v, red[w][threadIdx.x]) : fmaxf(v, red[w][threadIdx.x]);
        // coordinates are >= 0 after the clamp: the raw bit pattern is order preserving
        if (threadIdx.x < 3) atomicMin(bb_bits + threadIdx.x, __float_as_int(v));
        else atomicMax(bb_bits + threadIdx.x, __float_as_int(v));
    }
}

__global__ void __launch_bounds__(256) k_gen_bbox_cand(const bfm_gen_sample *__restrict__ S) {
    __shared__ bfm_gen_sample sd;
    stage_desc(&sd, S + blockIdx.y);
    const bfm_deform &d = sd.d;
    if (d.F_full || d.ncand[0] <= 0) return;                // no structure to exploit: full scan
    const int n0 = d.ncand[0], n1 = d.ncand[1], n2 = d.ncand[2];
    const int total = n0 * n1 * n2;
    if ((int)(blockIdx.x * blockDim.x) >= total) return;
    const DefRegs g = load_def(d);
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {0.f, 0.f, 0.f};
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < total) {
        const int c = p % n2, b = (p / n2) % n1, a = p / (n1 * n2);
        const int i = __ldg(d.cand[0] + a), j = __ldg(d.cand[1] + b), k = __ldg(d.cand[2] + c);
        float x1 = __fsub_rn((float)i, g.ctr0), y1 = __fsub_rn((float)j, g.ctr1), z1 = __fsub_rn((float)k, g.ctr2);
        if (d.fsmall) {
            float f0, f1, f2;
            field_direct(d, i, j, k, f0, f1, f2);
            x1 = __fadd_rn(x1, f0); y1 = __fadd_rn(y1, f1); z1 = __fadd_rn(z1, f2);
        }
        float px, py, pz;
        affine_clamp(g, x1, y1, z1, px, py, pz);
        lo[0] = hi[0] = px; lo[1] = hi[1] = py; lo[2] = hi[2] = pz;
    }
    block_minmax_atomic(lo, hi, sd.bbox);
}

__global__ void k_gen_bbox_decide(const bfm_gen_sample *__restrict__ S) {
    const bfm_gen_sample &s = S[blockIdx.x];
    int *bb = s.bbox;
    __shared__ int ambiguous;
    if (threadIdx.x == 0) ambiguous = (s.d.F_full || s.d.ncand[0] <= 0) ? 1 : 0;
    __syncthreads();
    int result = 0;
    if (threadIdx.x < 6) {
        const int a = threadIdx.x % 3;
        const float v = __int_as_float(bb[threadIdx.x]);
        bool exact;
        if (threadIdx.x < 3) {
            // true minimum over all voxels lies in [v - 2*delta, v] and is >= 0
            exact = (v == 0.f) || (floorf(v - 2.f * kBoxDelta) == floorf(v));
            result = (int)floorf(v);
        } else {
            exact = (v == (float)(s.d.src[a] - 1)) || (ceilf(v + 2.f * kBoxDelta) == ceilf(v));
            result = 1 + (int)ceilf(v);
        }
        if (!exact) atomicOr(&ambiguous, 1);
    }
    __syncthreads();
    if (threadIdx.x < 6) bb[threadIdx.x] = ambiguous ? (threadIdx.x < 3 ? 0x7f7fffff : 0) : result;
    if (threadIdx.x == 6) bb[6] = ambiguous;
}

__global__ void __launch_bounds__(kRowWarps * 32) k_gen_bbox_full(const bfm_gen_sample *__restrict__ S, int fstride) {
    extern __shared__ float smem[];
    __shared__ bfm_gen_sample sd;
    if (S[blockIdx.y].bbox[6] == 0) return;                  // candidate result was provably exact
    stage_desc(&sd, S + blockIdx.y);
    const bfm_deform &d = sd.d;
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
    block_minmax_atomic(lo, hi, sd.bbox);
}

__global__ void k_gen_bbox_finish(const bfm_gen_sample *__restrict__ S) {
    int *bb = S[blockIdx.x].bbox;
    if (bb[6] == 0) return;
    __syncthreads();
    // lo = floor(min), hi = 1 + ceil(max)   (datasets.py:288-293)
    if (threadIdx.x < 3) bb[threadIdx.x] = (int)floorf(__int_as_float(bb[threadIdx.x]));
    else if (threadIdx.x < 6) bb[threadIdx.x] = 1 + (int)ceilf(__int_as_float(bb[threadIdx.x]));
}

// ---------------------------------------------------------------------------------------------- gmm
// One thread = 4 consecutive source voxels along z (one 32-bit label load, one 128-bit store).
__device__ __forceinline__ int label_index(float g) {
    if (g == 77.f) g = 2.f;                       // datasets.py:366
    int r = __float2int_rn(g);                    // torch.round: half to even
    return min(max(r, 0), 255);
}

// Vector path (n2 % 4 == 0): grid = (groups per x-plane / 256, n0, B); no integer division by runtime values
// except one per thread.
__global__ void __launch_bounds__(256) k_gen_gmm_planes(const bfm_gen_sample *__restrict__ S) {
    __shared__ bfm_gen_sample sd;
    __shared__ float lut[512];
    const bfm_gen_sample *sp = S + blockIdx.z;
    const int x = blockIdx.y;
    {
        const int *bb = sp->bbox;
        if (x >= sp->d.src[0] || (sp->d.src[2] & 3) || x < bb[0] || x >= bb[3]) return;
    }
    stage_desc(&sd, sp);
    const bfm_gen_sample &s = sd;
    const int n1 = s.d.src[1], n2 = s.d.src[2];
    const int n2v = n2 >> 2;
    const int b1 = s.bbox[1], b2 = s.bbox[2], e1 = s.bbox[4], e2 = s.bbox[5], b0 = s.bbox[0];
    // block covers 256 consecutive groups of this plane: skip if all of its rows are outside [b1, e1)
    const int g0 = blockIdx.x * blockDim.x;
    if (g0 >= n1 * n2v) return;
    {
        const int ya = g0 / n2v, yb = min(g0 + (int)blockDim.x - 1, n1 * n2v - 1) / n2v;
        if (yb < b1 || ya >= e1) return;
    }
    for (int q = threadIdx.x; q < 512; q += blockDim.x) lut[q] = q < 256 ? __ldg(s.mu + q) : __ldg(s.sigma + q - 256);
    __syncthreads();
    const int gq = g0 + threadIdx.x;
    if (gq >= n1 * n2v) return;
    const int y = gq / n2v, z = (gq - y * n2v) << 2;
    if (y < b1 || y >= e1 || z + 3 < b2 || z >= e2) return;
    const int p0 = (x * n1 + y) * n2 + z;
    float ev[4];
    if (!s.eps_gmm) {
        const float4 e = philox_normal4(s.seed, 0u, (uint64_t)(p0 >> 2));
        ev[0] = e.x; ev[1] = e.y; ev[2] = e.z; ev[3] = e.w;
    } else {
        const int c1 = e1 - b1, c2 = e2 - b2;
        const float *er = s.eps_gmm + ((x - b0) * c1 + (y - b1)) * c2 - b2;
#pragma unroll
        for (int q = 0; q < 4; ++q) ev[q] = (z + q >= b2 && z + q < e2) ? __ldg(er + z + q) : 0.f;
    }
    int lab[4];
    if (s.label_is_u8) {
        const uint32_t w = __ldg((const uint32_t *)((const uint8_t *)s.labels + p0));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int l = (w >> (8 * q)) & 0xff;
            lab[q] = (l == 77) ? 2 : l;
        }
    } else {
        const float4 f = __ldg((const float4 *)((const float *)s.labels + p0));
        lab[0] = label_index(f.x); lab[1] = label_index(f.y); lab[2] = label_index(f.z); lab[3] = label_index(f.w);
    }
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float v = __fadd_rn(lut[lab[q]], __fmul_rn(lut[256 + lab[q]], ev[q]));
        o[q] = v < 0.f ? 0.f : v;
    }
    // positions just outside the crop are never gathered; writing them keeps the store 128-bit
    *(float4 *)(s.syn + p0) = make_float4(o[0], o[1], o[2], o[3]);
}

__global__ void __launch_bounds__(256) k_gen_gmm(const bfm_gen_sample *__restrict__ S) {
    __shared__ bfm_gen_sample sd;
    __shared__ float lut[512];
    stage_desc(&sd, S + blockIdx.y);
    const bfm_gen_sample &s = sd;
    const int n0 = s.d.src[0], n1 = s.d.src[1], n2 = s.d.src[2];
    if ((n2 & 3) == 0) return;                       // handled by k_gen_gmm_planes
    const int total = n0 * n1 * n2;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int blk0 = blockIdx.x * blockDim.x * 4;
    if (blk0 >= total) return;
    {   // whole block outside the crop slab along x?  (blocks cover contiguous flat ranges)
        const int xa = blk0 / (n1 * n2), xb = min(blk0 + (int)blockDim.x * 4 - 1, total - 1) / (n1 * n2);
        if (xb < s.bbox[0] || xa >= s.bbox[3]) return;
    }
    for (int q = threadIdx.x; q < 512; q += blockDim.x) lut[q] = q < 256 ? __ldg(s.mu + q) : __ldg(s.sigma + q - 256);
    __syncthreads();
    const int p0 = g * 4;
    if (p0 >= total) return;
    const int b0 = s.bbox[0], b1 = s.bbox[1], b2 = s.bbox[2], e0 = s.bbox[3], e1 = s.bbox[4], e2 = s.bbox[5];
    const int z = p0 % n2, y = (p0 / n2) % n1, x = p0 / (n1 * n2);
    const bool vec = ((n2 & 3) == 0);
    if (vec && (x < b0 || x >= e0 || y < b1 || y >= e1 || z + 3 < b2 || z >= e2)) return;
    const int c1 = e1 - b1, c2 = e2 - b2;
    float ev[4] = {0.f, 0.f, 0.f, 0.f};
    if (!s.eps_gmm) {
        const float4 e = philox_normal4(s.seed, 0u, (uint64_t)g);
        ev[0] = e.x; ev[1] = e.y; ev[2] = e.z; ev[3] = e.w;
    }
    int lab[4];
    if (vec && s.label_is_u8) {
        const uint32_t w = __ldg((const uint32_t *)((const uint8_t *)s.labels + p0));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int l = (w >> (8 * q)) & 0xff;
            lab[q] = (l == 77) ? 2 : l;
        }
    } else if (vec) {
        const float4 f = __ldg((const float4 *)((const float *)s.labels + p0));
        lab[0] = label_index(f.x); lab[1] = label_index(f.y); lab[2] = label_index(f.z); lab[3] = label_index(f.w);
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int p = p0 + q;
            if (p >= total) { lab[q] = 0; continue; }
            if (s.label_is_u8) {
                const int l = ((const uint8_t *)s.labels)[p];
                lab[q] = (l == 77) ? 2 : l;
            } else {
                lab[q] = label_index(((const float *)s.labels)[p]);
            }
        }
    }
    float outv[4];
    bool inside[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int p = p0 + q;
        int zz = z + q, yy = y, xx = x;
        if (!vec && zz >= n2) {               // group straddles a row end (only when n2 % 4 != 0)
            zz = p % n2; yy = (p / n2) % n1; xx = p / (n1 * n2);
        }
        inside[q] = p < total && xx >= b0 && xx < e0 && yy >= b1 && yy < e1 && zz >= b2 && zz < e2;
        float ee = ev[q];
        if (s.eps_gmm && inside[q]) ee = __ldg(s.eps_gmm + ((xx - b0) * c1 + (yy - b1)) * c2 + (zz - b2));
        const float v = __fadd_rn(lut[lab[q]], __fmul_rn(lut[256 + lab[q]], ee));
        outv[q] = v < 0.f ? 0.f : v;
    }
    if (vec) {
        // positions just outside the crop are never gathered; writing them keeps the store 128-bit
        *(float4 *)(s.syn + p0) = make_float4(outv[0], outv[1], outv[2], outv[3]);
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (inside[q]) s.syn[p0 + q] = outv[q];
    }
}

// ---------------------------------------------------------------------------------------------- warp
// pow / exp forms: ex2.approx(gamma * lg2.approx(x)) and ex2.approx(x*log2e); relative error ~1e-6, inside the
// 1e-5 parity tolerance (the reference's own CPU and CUDA pow differ at that level).
__device__ __forceinline__ float fast_pow(float x, float g) { return exp2f(g * __log2f(x)); }

#ifndef WARP_MINB
#define WARP_MINB 3
#endif
template <bool MIX, bool BFL>
__global__ void __launch_bounds__(kRowWarps * 32, WARP_MINB)
k_gen_warp(const bfm_gen_sample *__restrict__ S, int fstride, int bstride) {
    extern __shared__ float smem[];
    __shared__ bfm_gen_sample sd;
    {   // each instantiation handles the samples of its own kind
        const bfm_gen_sample *sp = S + blockIdx.y;
        if ((sp->mix[0] != nullptr) != MIX || (sp->bflog_out != nullptr) != BFL) return;
    }
    stage_desc(&sd, S + blockIdx.y);
    const bfm_gen_sample &s = sd;
    const bfm_deform &d = s.d;
    const DefRegs g = load_def(d);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *smF = smem + warp * kRowsPerWarp * (fstride + bstride);
    float *smB = smF + kRowsPerWarp * fstride;
    const int n_rows = g.s0 * g.s1;
    const int row0 = (blockIdx.x * kRowWarps + warp) * kRowsPerWarp;
    if (row0 >= n_rows) return;
    const float *__restrict__ bfsmall = s.bfsmall;
    const int bs2 = s.bs[2];
    if (bfsmall) {
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {
            const int row = min(row0 + r, n_rows - 1);
            row_zoom_setup(bfsmall, s.bs[1], bs2, 1, s.btab, row / g.s1, row % g.s1, smB + r * bs2, lane);
        }
    }
    const BoxRegs box = load_box(s.bbox, d.src[1], d.src[2]);
    const float *__restrict__ syn = s.syn;
    const float *__restrict__ mix0 = s.mix[0], *__restrict__ mix1 = s.mix[1], *__restrict__ mix2 = s.mix[2];
    const float mw0 = s.mixw[0], mw1 = s.mixw[1], mw2 = s.mixw[2], mw3 = s.mixw[3];
    const float gamma = s.gamma;
    float *__restrict__ i_bf = s.i_bf;
    float *__restrict__ bfl = s.bflog_out;
    const int flip = s.flip;
    const int *__restrict__ blo = s.btab.lo[2], *__restrict__ bhi = s.btab.hi[2];
    const float *__restrict__ bwl = s.btab.wl[2], *__restrict__ bwh = s.btab.wh[2];
    int klo = 0, khi = 0;
    float kwl = 0.f, kwh = 0.f;
    deform_rows<kRowsPerWarp>(
        d, g, smF, row0, n_rows, lane,
        [&](int k) {
            if (bfsmall) { klo = __ldg(blo + k); khi = __ldg(bhi + k); kwl = __ldg(bwl + k); kwh = __ldg(bwh + k); }
        },
        [&](int r, int row, int i, int j, int k, float px, float py, float pz) {
            const Taps32 t = make_taps32(px, py, pz, box);
            float v = trilerp32(t, [&](int e) { return __ldg(syn + e); });
            v = t.ok ? v : 0.f;
            const int p = row * g.s2 + k;
            if (MIX) {                                        // datasets.py:379-388
                v = __fadd_rn(__fmul_rn(mw0, v), __fmul_rn(mw1, mix0[p]));
                if (mix1) v = __fadd_rn(v, __fmul_rn(mw2, mix1[p]));
                if (mix2) v = __fadd_rn(v, __fmul_rn(mw3, mix2[p]));
            }
            v = v < 0.f ? 0.f : v;                            // datasets.py:411
            // gamma: 300 * (I/300) ** gamma                  utils.py:568-572
            v = 300.f * fast_pow(v * (1.f / 300.f), gamma);
            // bias field: I * exp(zoom(BFsmall))             utils.py:574-589
            if (bfsmall) {
                const float *sb = smB + r * bs2;
                const float bl = lerp_rn(kwl, sb[klo], kwh, sb[khi]);
                v *= exp2f(bl * 1.4426950408889634f);
                if (BFL) bfl[((flip ? g.s0 - 1 - i : i) * g.s1 + j) * g.s2 + k] = bl;
            }
            i_bf[p] = v;
        });
}

// ---------------------------------------------------------------------------------------------- resample
// One banded pass.  Axis 0/1: a thread owns VEC consecutive z outputs (128-bit loads when VEC == 4); the tap
// weight is uniform across the warp.  Axis 2: a thread owns one output, taps are contiguous.
template <int VEC>
__global__ void __launch_bounds__(256) k_gen_band(const bfm_gen_sample *__restrict__ S, int pass) {
    __shared__ bfm_gen_sample sd;
    if (pass >= S[blockIdx.y].n_band) return;
    stage_desc(&sd, S + blockIdx.y);
    const bfm_gen_sample &s = sd;
    int sh0 = s.d.size[0], sh1 = s.d.size[1], sh2 = s.d.size[2];
    for (int q = 0; q < pass; ++q) {
        const int ax = s.band[q].axis, no = s.band[q].n_out;
        if (ax == 0) sh0 = no; else if (ax == 1) sh1 = no; else sh2 = no;
    }
    const bfm_band &b = s.band[pass];
    const int axis = b.axis, T = b.T;
    const int o0 = axis == 0 ? b.n_out : sh0, o1 = axis == 1 ? b.n_out : sh1, o2 = axis == 2 ? b.n_out : sh2;
    if (VEC == 4 && (axis == 2 || (sh2 & 3))) return;        // handled by the scalar instantiation
    if (VEC == 1 && !(axis == 2 || (sh2 & 3))) return;
    const bool last = (pass == s.n_band - 1);
    const float *__restrict__ in = pass == 0 ? s.i_bf : s.tmp[(pass - 1) & 1];
    float *__restrict__ out = last ? s.lowres : s.tmp[pass & 1];
    const int stride = axis == 0 ? sh1 * sh2 : axis == 1 ? sh2 : 1;
    const int n_in = axis == 0 ? sh0 : axis == 1 ? sh1 : sh2;
    const int o2v = o2 / VEC;
    const int total = o0 * o1 * o2v;
    const int zf0 = s.zero_first[0], zf1 = s.zero_first[1], zf2 = s.zero_first[2];
    const float nstd = s.noise_std;
    const float *__restrict__ eps = s.eps_noise;
    const int *__restrict__ bstart = b.start;
    const float *__restrict__ bw = b.w;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
        const int kv = p % o2v, j = (p / o2v) % o1, i = p / (o1 * o2v);
        const int k = kv * VEC;
        const int q = axis == 0 ? i : axis == 1 ? j : k;
        const int st = __ldg(bstart + q);
        const int base = axis == 0 ? (st * sh1 + j) * sh2 + k : axis == 1 ? (i * sh1 + st) * sh2 + k : (i * sh1 + j) * sh2 + st;
        const float *__restrict__ wr = bw + q * T;
        const int t0 = max(0, -st), t1 = min(T, n_in - st);
        float acc[VEC];
#pragma unroll
        for (int c = 0; c < VEC; ++c) acc[c] = 0.f;
        const float *__restrict__ src = in + base;
#pragma unroll 4
        for (int t = t0; t < t1; ++t) {
            const float w = __ldg(wr + t);
            if (VEC == 4) {
                const float4 x = __ldg((const float4 *)(src + t * stride));
                acc[0] = fmaf(w, x.x, acc[0]); acc[1] = fmaf(w, x.y, acc[1]);
                acc[2] = fmaf(w, x.z, acc[2]); acc[3] = fmaf(w, x.w, acc[3]);
            } else {
                acc[0] = fmaf(w, __ldg(src + t * stride), acc[0]);
            }
        }
        const int op = (i * o1 + j) * o2 + k;
        if (last) {
            float4 e4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!eps && VEC == 4) e4 = philox_normal4(s.seed, 1u, (uint64_t)(op >> 2));   // o2 % 4 == 0 here
            const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
            for (int c = 0; c < VEC; ++c) {
                if ((zf0 && i == 0) || (zf1 && j == 0) || (zf2 && k + c == 0)) acc[c] = 0.f;
                float e;
                if (eps) e = __ldg(eps + op + c);
                else if (VEC == 4) e = ev[c];
                else {
                    const float4 g4 = philox_normal4(s.seed, 1u, (uint64_t)(op >> 2));
                    const int r = op & 3;
                    e = r == 0 ? g4.x : r == 1 ? g4.y : r == 2 ? g4.z : g4.w;
                }
                acc[c] = __fadd_rn(acc[c], __fmul_rn(nstd, e));      // utils.py:635-636
                acc[c] = acc[c] < 0.f ? 0.f : acc[c];
            }
        }
        if (VEC == 4) *(float4 *)(out + op) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        else out[op] = acc[0];
    }
}

// ---------------------------------------------------------------------------------------------- finish
// A warp owns kRowsPerWarp output rows; the first two zoom passes are evaluated once per low-res z node.
template <bool WRITE>
__global__ void __launch_bounds__(kRowWarps * 32) k_gen_upsample(const bfm_gen_sample *__restrict__ S, int max_lz) {
    extern __shared__ float smem[];
    __shared__ bfm_gen_sample sd;
    __shared__ float red[kRowWarps];
    stage_desc(&sd, S + blockIdx.y);
    const bfm_gen_sample &s = sd;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *sm = smem + warp * kRowsPerWarp * max_lz;
    const int s0 = s.d.size[0], s1 = s.d.size[1], s2 = s.d.size[2];
    const int n_rows = s0 * s1;
    const int row0 = (blockIdx.x * kRowWarps + warp) * kRowsPerWarp;
    float hi = 0.f;
    if (row0 < n_rows) {
        const int lz = s.new_size[2];
        const int nr = min(kRowsPerWarp, n_rows - row0);
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {
            const int row = min(row0 + r, n_rows - 1);
            row_zoom_setup(s.lowres, s.new_size[1], lz, 1, s.utab, row / s1, row % s1, sm + r * max_lz, lane);
        }
        __syncwarp();
        const int *__restrict__ lo = s.utab.lo[2], *__restrict__ hi2 = s.utab.hi[2];
        const float *__restrict__ wl = s.utab.wl[2], *__restrict__ wh = s.utab.wh[2];
        // I / max(I) (datasets.py:342-343) as a multiplication by the reciprocal: <= 1 ulp from the division
        const float rmx = WRITE ? __frcp_rn(*s.maxval) : 1.f;
        float *__restrict__ outp = s.out;
        float *__restrict__ resid = s.residual;
        const float *__restrict__ hr = s.i_bf;
        int obase[kRowsPerWarp];
#pragma unroll
        for (int r = 0; r < kRowsPerWarp; ++r) {
            const int row = min(row0 + r, n_rows - 1);
            const int i = row / s1, j = row - i * s1;
            obase[r] = ((s.flip ? s0 - 1 - i : i) * s1 + j) * s2;
        }
        for (int k = lane; k < s2; k += 32) {
            const int a = __ldg(lo + k), b = __ldg(hi2 + k);
            const float wa = __ldg(wl + k), wb = __ldg(wh + k);
#pragma unroll
            for (int r = 0; r < kRowsPerWarp; ++r) {
                if (r >= nr) break;
                const float *row_sm = sm + r * max_lz;
                const float v = lerp_rn(wa, row_sm[a], wb, row_sm[b]);
                if (WRITE) {
                    const float y = v * rmx;
                    outp[obase[r] + k] = y;
                    if (resid) resid[obase[r] + k] = __fsub_rn(hr[(row0 + r) * s2 + k] * rmx, y);   // datasets.py:345-347
                } else {
                    hi = fmaxf(hi, v);
                }
            }
        }
    }
    if (!WRITE) {
        hi = warp_max(hi);
        if (lane == 0) red[warp] = hi;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < kRowWarps; ++w) hi = fmaxf(hi, red[w]);
            atomicMax((int *)s.maxval, __float_as_int(hi));           // values >= 0: bit order == float order
        }
    }
}

static int check_batch(const bfm_gen_sample *h, const bfm_gen_sample *d, int B) {
    if (!h || !d || B <= 0) return fail(BFM_E_INVALID, "%s", "bfm_gen: null descriptors or empty batch");
    for (int b = 0; b < B; ++b) {
        const bfm_gen_sample &s = h[b];
        for (int a = 0; a < 3; ++a)
            if (s.d.size[a] <= 0 || s.d.src[a] <= 0 || s.new_size[a] <= 0)
                return fail(BFM_E_INVALID, "%s", "bfm_gen: non-positive size");
        if ((int64_t)s.d.src[0] * s.d.src[1] * s.d.src[2] >= (1LL << 31) ||
            (int64_t)s.d.size[0] * s.d.size[1] * s.d.size[2] >= (1LL << 31))
            return fail(BFM_E_UNSUPPORTED, "%s", "bfm_gen: volumes of 2^31 voxels or more are not supported");
        if (s.d.fsmall && (s.d.fs[2] > kMaxSmallZ || s.d.fs[2] <= 0))
            return fail(BFM_E_UNSUPPORTED, "%s", "bfm_gen: deformation small grid too deep");
        if (s.bfsmall && (s.bs[2] > kMaxSmallZ || s.bs[2] <= 0))
            return fail(BFM_E_UNSUPPORTED, "%s", "bfm_gen: bias small grid too deep");
        if (!s.labels || !s.mu || !s.sigma || !s.syn || !s.bbox || !s.i_bf || !s.lowres || !s.maxval || !s.out)
            return fail(BFM_E_INVALID, "%s", "bfm_gen: null buffer");
        if (s.n_band < 1 || s.n_band > 3) return fail(BFM_E_INVALID, "%s", "bfm_gen: n_band must be 1..3");
        if (s.d.size[0] != h[0].d.size[0] || s.d.size[1] != h[0].d.size[1] || s.d.size[2] != h[0].d.size[2])
            return fail(BFM_E_UNSUPPORTED, "%s", "bfm_gen: all samples of a batch share the output size");
    }
    return BFM_OK;
}

static inline unsigned rows_grid(const bfm_gen_sample *h) {
    const int64_t rows = (int64_t)h[0].d.size[0] * h[0].d.size[1];
    const int per_block = kRowWarps * kRowsPerWarp;
    return (unsigned)((rows + per_block - 1) / per_block);
}
static inline int max_fstride(const bfm_gen_sample *h, int B) {
    int m = 0;
    for (int b = 0; b < B; ++b)
        if (h[b].d.fsmall && !h[b].d.F_full) m = h[b].d.fs[2] * 3 > m ? h[b].d.fs[2] * 3 : m;
    return m;
}
static inline int max_bstride(const bfm_gen_sample *h, int B) {
    int m = 0;
    for (int b = 0; b < B; ++b)
        if (h[b].bfsmall) m = h[b].bs[2] > m ? h[b].bs[2] : m;
    return m;
}
}  // namespace bfm

using namespace bfm;

extern "C" {

int bfm_gen_bbox(const bfm_gen_sample *h, const bfm_gen_sample *d, int B, void *stream) {
    int rc = check_batch(h, d, B);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    int64_t most = 0;
    for (int b = 0; b < B; ++b) {
        const int64_t n = (int64_t)h[b].d.ncand[0] * h[b].d.ncand[1] * h[b].d.ncand[2];
        if (n >= (1LL << 30)) return fail(BFM_E_INVALID, "%s", "bfm_gen_bbox: too many candidate voxels");
        most = n > most ? n : most;
    }
    k_gen_bbox_init<<<B, 32, 0, s>>>(d);
    if (most > 0) k_gen_bbox_cand<<<dim3((unsigned)((most + 255) / 256), B), 256, 0, s>>>(d);
    k_gen_bbox_decide<<<B, 32, 0, s>>>(d);
    const int fstride = max_fstride(h, B);
    const size_t smem = (size_t)kRowWarps * kRowsPerWarp * fstride * sizeof(float);
    k_gen_bbox_full<<<dim3(rows_grid(h), B), kRowWarps * 32, smem, s>>>(d, fstride);
    k_gen_bbox_finish<<<B, 32, 0, s>>>(d);
    g_launches.fetch_add(most > 0 ? 4 : 3);
    return check_launch("bfm_gen_bbox");
}

int bfm_gen_gmm(const bfm_gen_sample *h, const bfm_gen_sample *d, int B, void *stream) {
    int rc = check_batch(h, d, B);
    if (rc) return rc;
    int64_t flat_groups = 0, plane_groups = 0;
    int max_n0 = 0;
    for (int b = 0; b < B; ++b) {
        const int *n = h[b].d.src;
        if (n[2] & 3) {
            const int64_t g = ((int64_t)n[0] * n[1] * n[2] + 3) / 4;
            flat_groups = g > flat_groups ? g : flat_groups;
        } else {
            const int64_t g = (int64_t)n[1] * (n[2] / 4);
            plane_groups = g > plane_groups ? g : plane_groups;
            max_n0 = n[0] > max_n0 ? n[0] : max_n0;
        }
    }
    if (plane_groups > 0) {
        if (max_n0 > 65535 || B > 65535) return fail(BFM_E_UNSUPPORTED, "%s", "bfm_gen_gmm: grid too large");
        k_gen_gmm_planes<<<dim3((unsigned)((plane_groups + 255) / 256), max_n0, B), 256, 0, (cudaStream_t)stream>>>(d);
        rc = check_launch("bfm_gen_gmm");
        if (rc) return rc;
    }
    if (flat_groups > 0) {
        k_gen_gmm<<<dim3((unsigned)((flat_groups + 255) / 256), B), 256, 0, (cudaStream_t)stream>>>(d);
        rc = check_launch("bfm_gen_gmm");
    }
    return rc;
}

int bfm_gen_warp(const bfm_gen_sample *h, const bfm_gen_sample *d, int B, void *stream) {
    int rc = check_batch(h, d, B);
    if (rc) return rc;
    const int fstride = max_fstride(h, B), bstride = max_bstride(h, B);
    const size_t smem = (size_t)kRowWarps * kRowsPerWarp * (fstride + bstride) * sizeof(float);
    if (smem > 200 * 1024) return fail(BFM_E_UNSUPPORTED, "%s", "bfm_gen_warp: small grids too deep for shared memory");
    bool kinds[2][2] = {{false, false}, {false, false}};
    for (int b = 0; b < B; ++b) kinds[h[b].mix[0] != nullptr][h[b].bflog_out != nullptr] = true;
    const dim3 grid(rows_grid(h), B);
    cudaStream_t st = (cudaStream_t)stream;
#define BFM_LAUNCH_WARP(M, L)                                                                                  \
    if (kinds[M][L]) {                                                                                         \
        if (smem > 40 * 1024)                                                                                  \
            cudaFuncSetAttribute(k_gen_warp<M, L>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
        k_gen_warp<M, L><<<grid, kRowWarps * 32, smem, st>>>(d, fstride, bstride);                             \
        rc = check_launch("bfm_gen_warp");                                                                     \
        if (rc) return rc;                                                                                     \
    }
    BFM_LAUNCH_WARP(false, false)
    BFM_LAUNCH_WARP(false, true)
    BFM_LAUNCH_WARP(true, false)
    BFM_LAUNCH_WARP(true, true)
#undef BFM_LAUNCH_WARP
    return BFM_OK;
}

int bfm_gen_resample(const bfm_gen_sample *h, const bfm_gen_sample *d, int B, void *stream) {
    int rc = check_batch(h, d, B);
    if (rc) return rc;
    int maxp = 0;
    for (int b = 0; b < B; ++b) maxp = h[b].n_band > maxp ? h[b].n_band : maxp;
    for (int pass = 0; pass < maxp; ++pass) {
        int64_t most4 = 0, most1 = 0;
        for (int b = 0; b < B; ++b) {
            if (pass >= h[b].n_band) continue;
            int sh[3] = {h[b].d.size[0], h[b].d.size[1], h[b].d.size[2]};
            for (int q = 0; q < pass; ++q) sh[h[b].band[q].axis] = h[b].band[q].n_out;
            const int axis = h[b].band[pass].axis;
            const bool scalar = axis == 2 || (sh[2] & 3);
            sh[axis] = h[b].band[pass].n_out;
            const int64_t n = (int64_t)sh[0] * sh[1] * sh[2];
            if (scalar) most1 = n > most1 ? n : most1;
            else most4 = n / 4 > most4 ? n / 4 : most4;
        }
        if (most4 > 0) {
            k_gen_band<4><<<dim3((unsigned)((most4 + 255) / 256), B), 256, 0, (cudaStream_t)stream>>>(d, pass);
            int rc2 = check_launch("bfm_gen_resample");
            if (rc2) return rc2;
        }
        if (most1 > 0) {
            k_gen_band<1><<<dim3((unsigned)((most1 + 255) / 256), B), 256, 0, (cudaStream_t)stream>>>(d, pass);
            int rc2 = check_launch("bfm_gen_resample");
            if (rc2) return rc2;
        }
    }
    return BFM_OK;
}

int bfm_gen_finish(const bfm_gen_sample *h, const bfm_gen_sample *d, int B, void *stream) {
    int rc = check_batch(h, d, B);
    if (rc) return rc;
    int max_lz = 1;
    for (int b = 0; b < B; ++b) max_lz = h[b].new_size[2] > max_lz ? h[b].new_size[2] : max_lz;
    const size_t smem = (size_t)kRowWarps * kRowsPerWarp * max_lz * sizeof(float);
    if (smem > 200 * 1024) return fail(BFM_E_UNSUPPORTED, "%s", "bfm_gen_finish: low-res row too long");
    if (smem > 40 * 1024) {
        cudaFuncSetAttribute(k_gen_upsample<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(k_gen_upsample<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    cudaStream_t s = (cudaStream_t)stream;
    k_gen_upsample<false><<<dim3(rows_grid(h), B), kRowWarps * 32, smem, s>>>(d, max_lz);
    g_launches.fetch_add(1);
    k_gen_upsample<true><<<dim3(rows_grid(h), B), kRowWarps * 32, smem, s>>>(d, max_lz);
    return check_launch("bfm_gen_finish");
}

int bfm_gen_run(const bfm_gen_sample *h, const bfm_gen_sample *d, int B, void *stream) {
    int rc;
    if ((rc = bfm_gen_bbox(h, d, B, stream))) return rc;
    if ((rc = bfm_gen_gmm(h, d, B, stream))) return rc;
    if ((rc = bfm_gen_warp(h, d, B, stream))) return rc;
    if ((rc = bfm_gen_resample(h, d, B, stream))) return rc;
    return bfm_gen_finish(h, d, B, stream);
}

}  // extern "C"
