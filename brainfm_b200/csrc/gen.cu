// Fused, batched synthesis chain: BaseGen.generate_sample + augment_sample with the stock steps
// (Generator/datasets.py:306-428).  One launch per stage covers the whole batch (blockIdx.y = sample).
//
//   bbox     deform_grid (datasets.py:264-303): coordinates -> bounding box, nothing written to HBM
//   gmm      mus[Gr] + sigmas[Gr]*eps, clamp (datasets.py:364-372) over the bbox crop -> syn
//   warp     coordinates again -> trilinear gather of syn -> mix -> clamp -> gamma -> bias field
//            (utils.py:140-192, datasets.py:379-411, utils.py:568-589) -> i_bf, bflog_out
//   resample blur o downsample as banded per-axis maps + noise (utils.py:83-94, 591-609, 633-638)
//   finish   myzoom_torch back to the grid, global max, normalise, flip (datasets.py:337-352)
#include "common.cuh"

namespace bfm {

// ---------------------------------------------------------------------------------------------- bbox
__global__ void k_gen_bbox_init(const bfm_gen_sample *__restrict__ S) {
    int *bb = S[blockIdx.x].bbox;
    if (threadIdx.x < 3) bb[threadIdx.x] = 0x7f7fffff;
    else if (threadIdx.x < 6) bb[threadIdx.x] = 0;
    if (threadIdx.x == 6) *S[blockIdx.x].maxval = 0.f;   // chain values are >= 0 after the noise clamp
}

__global__ void __launch_bounds__(kRowWarps * 32) k_gen_bbox(const bfm_gen_sample *__restrict__ S) {
    __shared__ float smF[kRowWarps * kMaxSmallZ * 3];
    __shared__ float red[kRowWarps][6];
    const bfm_deform &d = S[blockIdx.y].d;
    int *bb_bits = S[blockIdx.y].bbox;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *sm = smF + warp * (kMaxSmallZ * 3);
    const int64_t row = (int64_t)blockIdx.x * kRowWarps + warp;
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {0.f, 0.f, 0.f};
    if (row < (int64_t)d.size[0] * d.size[1]) {
        const int i = (int)(row / d.size[1]), j = (int)(row % d.size[1]);
        if (d.fsmall && !d.F_full) {
            row_zoom_setup(d.fsmall, d.fs[1], d.fs[2], 3, d.ftab, i, j, sm, lane);
            __syncwarp();
        }
        for (int k = lane; k < d.size[2]; k += 32) {
            float px, py, pz;
            voxel_coords(d, sm, i, j, k, px, py, pz);
            lo[0] = fminf(lo[0], px); hi[0] = fmaxf(hi[0], px);
            lo[1] = fminf(lo[1], py); hi[1] = fmaxf(hi[1], py);
            lo[2] = fminf(lo[2], pz); hi[2] = fmaxf(hi[2], pz);
        }
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
        if (threadIdx.x < 3) atomicMin(bb_bits + threadIdx.x, __float_as_int(v));
        else atomicMax(bb_bits + threadIdx.x, __float_as_int(v));
    }
}

__global__ void k_gen_bbox_finish(const bfm_gen_sample *__restrict__ S) {
    int *bb = S[blockIdx.x].bbox;
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

__global__ void __launch_bounds__(256) k_gen_gmm(const bfm_gen_sample *__restrict__ S) {
    const bfm_gen_sample &s = S[blockIdx.y];
    const int n0 = s.d.src[0], n1 = s.d.src[1], n2 = s.d.src[2];
    const int64_t total = (int64_t)n0 * n1 * n2;
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t p0 = g * 4;
    if (p0 >= total) return;
    const int *bb = s.bbox;
    int z = (int)(p0 % n2), y = (int)((p0 / n2) % n1), x = (int)(p0 / ((int64_t)n1 * n2));
    const bool row_whole = (z + 3 < n2);
    if (row_whole && (x < bb[0] || x >= bb[3] || y < bb[1] || y >= bb[4] || z + 3 < bb[2] || z >= bb[5])) return;
    const int c1 = bb[4] - bb[1], c2 = bb[5] - bb[2];
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!s.eps_gmm) e = philox_normal4(s.seed, 0u, (uint64_t)g);
    float ev[4] = {e.x, e.y, e.z, e.w};
    float outv[4];
    bool inside[4];
    int lab[4];
    const bool vec = row_whole && ((n2 & 3) == 0);
    if (vec && s.label_is_u8) {
        const uint32_t w = __ldg((const uint32_t *)((const uint8_t *)s.labels + p0));
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            int l = (w >> (8 * q)) & 0xff;
            lab[q] = (l == 77) ? 2 : l;
        }
    } else if (vec) {
        const float4 f = __ldg((const float4 *)((const float *)s.labels + p0));
        lab[0] = label_index(f.x); lab[1] = label_index(f.y); lab[2] = label_index(f.z); lab[3] = label_index(f.w);
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int64_t p = p0 + q;
            if (p >= total) { lab[q] = 0; continue; }
            if (s.label_is_u8) {
                int l = ((const uint8_t *)s.labels)[p];
                lab[q] = (l == 77) ? 2 : l;
            } else {
                lab[q] = label_index(((const float *)s.labels)[p]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int64_t p = p0 + q;
        int zz = z + q, yy = y, xx = x;
        if (zz >= n2) {               // group straddles a row end (only when n2 % 4 != 0)
            zz = (int)(p % n2); yy = (int)((p / n2) % n1); xx = (int)(p / ((int64_t)n1 * n2));
        }
        inside[q] = p < total && xx >= bb[0] && xx < bb[3] && yy >= bb[1] && yy < bb[4] && zz >= bb[2] && zz < bb[5];
        float ee = ev[q];
        if (s.eps_gmm && inside[q])
            ee = __ldg(s.eps_gmm + ((int64_t)(xx - bb[0]) * c1 + (yy - bb[1])) * c2 + (zz - bb[2]));
        float v = __fadd_rn(__ldg(s.mu + lab[q]), __fmul_rn(__ldg(s.sigma + lab[q]), ee));
        outv[q] = v < 0.f ? 0.f : v;
    }
    if (vec) {
        // positions outside the crop are never gathered; writing them is harmless and keeps the store 128-bit
        *(float4 *)(s.syn + p0) = make_float4(outv[0], outv[1], outv[2], outv[3]);
    } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (inside[q]) s.syn[p0 + q] = outv[q];
    }
}

// ---------------------------------------------------------------------------------------------- warp
__global__ void __launch_bounds__(kRowWarps * 32) k_gen_warp(const bfm_gen_sample *__restrict__ S) {
    __shared__ float smF[kRowWarps * kMaxSmallZ * 3];
    __shared__ float smB[kRowWarps * kMaxSmallZ];
    const bfm_gen_sample &s = S[blockIdx.y];
    const bfm_deform &d = s.d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *sm = smF + warp * (kMaxSmallZ * 3);
    float *sb = smB + warp * kMaxSmallZ;
    const int64_t row = (int64_t)blockIdx.x * kRowWarps + warp;
    if (row >= (int64_t)d.size[0] * d.size[1]) return;
    const int i = (int)(row / d.size[1]), j = (int)(row % d.size[1]);
    if (d.fsmall && !d.F_full) row_zoom_setup(d.fsmall, d.fs[1], d.fs[2], 3, d.ftab, i, j, sm, lane);
    if (s.bfsmall) row_zoom_setup(s.bfsmall, s.bs[1], s.bs[2], 1, s.btab, i, j, sb, lane);
    __syncwarp();
    const int *bb = s.bbox;
    const int n1 = d.src[1], n2 = d.src[2];
    const float *__restrict__ syn = s.syn;
    const int oi = s.flip ? d.size[0] - 1 - i : i;
    const int64_t orow = ((int64_t)oi * d.size[1] + j) * d.size[2];
    for (int k = lane; k < d.size[2]; k += 32) {
        float px, py, pz;
        voxel_coords(d, sm, i, j, k, px, py, pz);
        Taps t = make_taps(px, py, pz, bb);
        float v = 0.f;
        if (t.ok) v = trilerp(t, [&](int x, int y, int z) { return __ldg(syn + ((int64_t)x * n1 + y) * n2 + z); });
        const int64_t p = row * d.size[2] + k;
        if (s.mix[0]) {                                   // datasets.py:379-388
            v = __fadd_rn(__fmul_rn(s.mixw[0], v), __fmul_rn(s.mixw[1], s.mix[0][p]));
            if (s.mix[1]) v = __fadd_rn(v, __fmul_rn(s.mixw[2], s.mix[1][p]));
            if (s.mix[2]) v = __fadd_rn(v, __fmul_rn(s.mixw[3], s.mix[2][p]));
        }
        if (v < 0.f) v = 0.f;                             // datasets.py:411
        // gamma: 300 * (I/300) ** gamma                  utils.py:568-572
        v = __fmul_rn(300.f, powf(__fdiv_rn(v, 300.f), s.gamma));
        // bias field: I * exp(zoom(BFsmall))             utils.py:574-589
        if (s.bfsmall) {
            const float bl = lerp_rn(s.btab.wl[2][k], sb[s.btab.lo[2][k]], s.btab.wh[2][k], sb[s.btab.hi[2][k]]);
            v = __fmul_rn(v, expf(bl));
            if (s.bflog_out) s.bflog_out[orow + k] = bl;
        }
        s.i_bf[p] = v;
    }
}

// ---------------------------------------------------------------------------------------------- resample
__global__ void __launch_bounds__(256) k_gen_band(const bfm_gen_sample *__restrict__ S, int pass) {
    const bfm_gen_sample &s = S[blockIdx.y];
    if (pass >= s.n_band) return;
    // shape before this pass
    int sh[3] = {s.d.size[0], s.d.size[1], s.d.size[2]};
    for (int q = 0; q < pass; ++q) sh[s.band[q].axis] = s.band[q].n_out;
    const bfm_band &b = s.band[pass];
    const int axis = b.axis;
    int o[3] = {sh[0], sh[1], sh[2]};
    o[axis] = b.n_out;
    const int64_t total = (int64_t)o[0] * o[1] * o[2];
    const bool last = (pass == s.n_band - 1);
    const float *__restrict__ in = pass == 0 ? s.i_bf : s.tmp[(pass - 1) & 1];
    float *__restrict__ out = last ? s.lowres : s.tmp[pass & 1];
    const int64_t stride = axis == 0 ? (int64_t)sh[1] * sh[2] : axis == 1 ? sh[2] : 1;
    const int n_in = sh[axis];
    const int T = b.T;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(p % o[2]);
        const int j = (int)((p / o[2]) % o[1]);
        const int i = (int)(p / ((int64_t)o[1] * o[2]));
        const int q = axis == 0 ? i : axis == 1 ? j : k;
        const int st = __ldg(b.start + q);
        int64_t base;
        if (axis == 0) base = ((int64_t)st * sh[1] + j) * sh[2] + k;
        else if (axis == 1) base = ((int64_t)i * sh[1] + st) * sh[2] + k;
        else base = ((int64_t)i * sh[1] + j) * sh[2] + st;
        const float *wr = b.w + (int64_t)q * T;
        float acc = 0.f;
        for (int t = 0; t < T; ++t) {
            const int src = st + t;
            if (src >= 0 && src < n_in) acc = fmaf(__ldg(wr + t), __ldg(in + base + t * stride), acc);
        }
        if (last) {
            if ((s.zero_first[0] && i == 0) || (s.zero_first[1] && j == 0) || (s.zero_first[2] && k == 0)) acc = 0.f;
            float e;
            if (s.eps_noise) e = __ldg(s.eps_noise + p);
            else {
                float4 g = philox_normal4(s.seed, 1u, (uint64_t)p >> 2);
                const int r = (int)(p & 3);
                e = r == 0 ? g.x : r == 1 ? g.y : r == 2 ? g.z : g.w;
            }
            acc = __fadd_rn(acc, __fmul_rn(s.noise_std, e));       // utils.py:635-636
            if (acc < 0.f) acc = 0.f;
        }
        out[p] = acc;
    }
}

// ---------------------------------------------------------------------------------------------- finish
// warp per output row; the first two zoom passes are evaluated once per low-res z node.
template <bool WRITE>
__global__ void __launch_bounds__(kRowWarps * 32) k_gen_upsample(const bfm_gen_sample *__restrict__ S, int max_lz) {
    extern __shared__ float smem[];
    const bfm_gen_sample &s = S[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float *sm = smem + warp * max_lz;
    const int s0 = s.d.size[0], s1 = s.d.size[1], s2 = s.d.size[2];
    const int64_t row = (int64_t)blockIdx.x * kRowWarps + warp;
    float hi = 0.f;
    if (row < (int64_t)s0 * s1) {
        const int i = (int)(row / s1), j = (int)(row % s1);
        row_zoom_setup(s.lowres, s.new_size[1], s.new_size[2], 1, s.utab, i, j, sm, lane);
        __syncwarp();
        const int *__restrict__ lo = s.utab.lo[2], *__restrict__ hi2 = s.utab.hi[2];
        const float *__restrict__ wl = s.utab.wl[2], *__restrict__ wh = s.utab.wh[2];
        if (WRITE) {
            const float mx = *s.maxval;
            const int oi = s.flip ? s0 - 1 - i : i;
            float *__restrict__ o = s.out + ((int64_t)oi * s1 + j) * s2;
            float *__restrict__ r = s.residual ? s.residual + ((int64_t)oi * s1 + j) * s2 : nullptr;
            const float *__restrict__ hr = s.i_bf + row * s2;
            for (int k = lane; k < s2; k += 32) {
                const float v = lerp_rn(wl[k], sm[lo[k]], wh[k], sm[hi2[k]]);
                const float y = __fdiv_rn(v, mx);                      // datasets.py:342-343
                o[k] = y;
                if (r) r[k] = __fsub_rn(__fdiv_rn(hr[k], mx), y);       // datasets.py:345-347
            }
        } else {
            for (int k = lane; k < s2; k += 32) hi = fmaxf(hi, lerp_rn(wl[k], sm[lo[k]], wh[k], sm[hi2[k]]));
        }
    }
    if (!WRITE) {
        __shared__ float red[kRowWarps];
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
    return (unsigned)((rows + kRowWarps - 1) / kRowWarps);
}
}  // namespace bfm

using namespace bfm;

extern "C" {

int bfm_gen_bbox(const bfm_gen_sample *h, const bfm_gen_sample *d, int B, void *stream) {
    int rc = check_batch(h, d, B);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    k_gen_bbox_init<<<B, 32, 0, s>>>(d);
    k_gen_bbox<<<dim3(rows_grid(h), B), kRowWarps * 32, 0, s>>>(d);
    k_gen_bbox_finish<<<B, 32, 0, s>>>(d);
    g_launches.fetch_add(2);
    return check_launch("bfm_gen_bbox");
}

int bfm_gen_gmm(const bfm_gen_sample *h, const bfm_gen_sample *d, int B, void *stream) {
    int rc = check_batch(h, d, B);
    if (rc) return rc;
    int64_t groups = 0;
    for (int b = 0; b < B; ++b) {
        int64_t n = ((int64_t)h[b].d.src[0] * h[b].d.src[1] * h[b].d.src[2] + 3) / 4;
        groups = n > groups ? n : groups;
    }
    k_gen_gmm<<<dim3((unsigned)((groups + 255) / 256), B), 256, 0, (cudaStream_t)stream>>>(d);
    return check_launch("bfm_gen_gmm");
}

int bfm_gen_warp(const bfm_gen_sample *h, const bfm_gen_sample *d, int B, void *stream) {
    int rc = check_batch(h, d, B);
    if (rc) return rc;
    k_gen_warp<<<dim3(rows_grid(h), B), kRowWarps * 32, 0, (cudaStream_t)stream>>>(d);
    return check_launch("bfm_gen_warp");
}

int bfm_gen_resample(const bfm_gen_sample *h, const bfm_gen_sample *d, int B, void *stream) {
    int rc = check_batch(h, d, B);
    if (rc) return rc;
    int maxp = 0;
    for (int b = 0; b < B; ++b) maxp = h[b].n_band > maxp ? h[b].n_band : maxp;
    for (int pass = 0; pass < maxp; ++pass) {
        int64_t most = 0;
        for (int b = 0; b < B; ++b) {
            if (pass >= h[b].n_band) continue;
            int sh[3] = {h[b].d.size[0], h[b].d.size[1], h[b].d.size[2]};
            for (int q = 0; q <= pass; ++q) sh[h[b].band[q].axis] = h[b].band[q].n_out;
            int64_t n = (int64_t)sh[0] * sh[1] * sh[2];
            most = n > most ? n : most;
        }
        unsigned gx = (unsigned)((most + 255) / 256);
        if (gx < 1) gx = 1;
        k_gen_band<<<dim3(gx, B), 256, 0, (cudaStream_t)stream>>>(d, pass);
        int rc2 = check_launch("bfm_gen_resample");
        if (rc2) return rc2;
    }
    return BFM_OK;
}

int bfm_gen_finish(const bfm_gen_sample *h, const bfm_gen_sample *d, int B, void *stream) {
    int rc = check_batch(h, d, B);
    if (rc) return rc;
    int max_lz = 1;
    for (int b = 0; b < B; ++b) max_lz = h[b].new_size[2] > max_lz ? h[b].new_size[2] : max_lz;
    const size_t smem = (size_t)kRowWarps * max_lz * sizeof(float);
    if (smem > 200 * 1024) return fail(BFM_E_UNSUPPORTED, "%s", "bfm_gen_finish: low-res row too long");
    if (smem > 48 * 1024) {
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
