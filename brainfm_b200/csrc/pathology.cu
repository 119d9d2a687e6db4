// Op-level kernels of the pathology branch of generate_sample / augment_sample
// (reference Generator/datasets.py:357-411, 496-518).  The fused chain (gen.cu) does not paint lesions; samples with a
// pathology map run op by op, and these kernels replace the eager tensor expressions of that path:
//
//   bfm_gmm_crop          SYN = clamp(mus[Gr] + sigmas[Gr] * eps, 0) over the bbox crop        datasets.py:364-372
//   bfm_pathol_cerebral   SYN_cerebral = SYN * (Gr != 0); white / grey matter sums             datasets.py:391-398
//   bfm_zero_where_zero   P[C == 0] = 0                                                        datasets.py:399-400
//   bfm_masked_mean       sum(I * P), sum(P) in float64                                        datasets.py:500
//   bfm_encode_pathology  I += Pprob * (mu[round(P)] + sigma[round(P)] * eps); I[I < 0] = 0    datasets.py:509-513
//
// dtype rules follow the reference's tensor expressions: P / Pprob are float64 for random Perlin shapes (the product
// and the in-place add are then evaluated in float64 and rounded to float32 once), float32 for maps read from disk.
#include "common.cuh"

namespace bfm {

static inline unsigned blocks_for(int64_t n, int per_thread = 1) {
    int64_t g = (n + 256LL * per_thread - 1) / (256LL * per_thread);
    const int64_t cap = 148LL * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

__device__ __forceinline__ int label_of(const void *labels, int is_u8, int64_t p) {
    if (is_u8) {
        const int l = ((const uint8_t *)labels)[p];
        return l == 77 ? 2 : l;                            // datasets.py:366
    }
    float g = ((const float *)labels)[p];
    if (g == 77.f) g = 2.f;
    const int r = __float2int_rn(g);                       // torch.round: half to even
    return min(max(r, 0), 255);
}

struct CropBox { int n1, n2, b0, b1, b2, c0, c1, c2; };

__device__ __forceinline__ int64_t crop_to_src(const CropBox &bx, int64_t q) {
    const int z = (int)(q % bx.c2);
    const int64_t r = q / bx.c2;
    const int y = (int)(r % bx.c1), x = (int)(r / bx.c1);
    return ((int64_t)(bx.b0 + x) * bx.n1 + (bx.b1 + y)) * bx.n2 + (bx.b2 + z);
}

__global__ void __launch_bounds__(256) k_gmm_crop(const void *__restrict__ labels, int is_u8, CropBox bx,
                                                  const float *__restrict__ mu, const float *__restrict__ sigma,
                                                  const float *__restrict__ eps, uint64_t seed, float *__restrict__ out,
                                                  int64_t n) {
    __shared__ float lut[512];
    for (int q = threadIdx.x; q < 512; q += blockDim.x) lut[q] = q < 256 ? __ldg(mu + q) : __ldg(sigma + q - 256);
    __syncthreads();
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = crop_to_src(bx, q);
        const int l = label_of(labels, is_u8, p);
        float e;
        if (eps) {
            e = __ldg(eps + q);
        } else {                                            // the chain's counter-based field: stream 0, absolute voxel
            const float4 t = philox_normal4(seed, 0u, (uint64_t)(p >> 2));
            const int c = (int)(p & 3);
            e = c == 0 ? t.x : c == 1 ? t.y : c == 2 ? t.z : t.w;
        }
        const float v = __fadd_rn(lut[l], __fmul_rn(lut[256 + l], e));
        out[q] = v < 0.f ? 0.f : v;
    }
}

// sums[0..3] = sum(SYN * wm), count(wm), sum(SYN * gm), count(gm); wm = label in {2, 41}, gm = other non-zero labels
__global__ void __launch_bounds__(256) k_pathol_cerebral(const float *__restrict__ syn, const void *__restrict__ labels,
                                                         int is_u8, CropBox bx, float *__restrict__ cerebral,
                                                         double *__restrict__ sums, int64_t n) {
    double a[4] = {0, 0, 0, 0};
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const int l = label_of(labels, is_u8, crop_to_src(bx, q));
        const float v = syn[q];
        cerebral[q] = l == 0 ? 0.f : v;
        if (l == 2 || l == 41) { a[0] += v; a[1] += 1.0; }
        else if (l != 0) { a[2] += v; a[3] += 1.0; }
    }
    __shared__ double red[8][4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a[c] += __shfl_xor_sync(0xffffffffu, a[c], o);
    }
    if ((threadIdx.x & 31) == 0)
        for (int c = 0; c < 4; ++c) red[threadIdx.x >> 5][c] = a[c];
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w][threadIdx.x];
        atomicAdd(sums + threadIdx.x, t);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_zero_where_zero(T *__restrict__ p, const float *__restrict__ c, int64_t n) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
        if (c[q] == 0.f) p[q] = (T)0;
}

template <typename T>
__global__ void __launch_bounds__(256) k_masked_mean(const float *__restrict__ I, const T *__restrict__ P,
                                                     double *__restrict__ sums, int64_t n) {
    double a = 0, b = 0;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const T pv = P[q];
        if (sizeof(T) == 8) a += (double)I[q] * (double)pv;     // float32 * float64 -> float64
        else a += (double)__fmul_rn(I[q], (float)pv);           // float32 * float32 -> float32, summed
        b += (double)pv;
    }
    __shared__ double red[8][2];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = a; red[threadIdx.x >> 5][1] = b; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double t = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w][threadIdx.x];
        atomicAdd(sums + threadIdx.x, t);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) k_encode_pathology(float *__restrict__ I, const T *__restrict__ P,
                                                          const T *__restrict__ Pprob, const float *__restrict__ mus,
                                                          const float *__restrict__ sigmas, int n_tab,
                                                          const float *__restrict__ eps, uint64_t seed, int64_t n) {
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const T pv = P[q];
        // torch.round(P).long(): half to even
        const long long m = sizeof(T) == 8 ? __double2ll_rn((double)pv) : (long long)__float2int_rn((float)pv);
        const int idx = (int)(m < 0 ? m + n_tab : m);            // python-style negative index
        float e;
        if (eps) {
            e = __ldg(eps + q);
        } else {
            const float4 t = philox_normal4(seed, 2u, (uint64_t)(q >> 2));
            const int c = (int)(q & 3);
            e = c == 0 ? t.x : c == 1 ? t.y : c == 2 ? t.z : t.w;
        }
        const float draw = __fadd_rn(__ldg(mus + idx), __fmul_rn(__ldg(sigmas + idx), e));
        float v;
        if (sizeof(T) == 8) v = (float)__dadd_rn((double)I[q], __dmul_rn((double)Pprob[q], (double)draw));
        else v = __fadd_rn(I[q], __fmul_rn((float)Pprob[q], draw));
        I[q] = v < 0.f ? 0.f : v;
    }
}

}  // namespace bfm

using namespace bfm;

extern "C" {

int bfm_gmm_crop(const void *labels, int label_is_u8, const int *src, const int *bbox, const float *mu,
                 const float *sigma, const float *eps, uint64_t seed, float *out, void *stream) {
    BFM_REQUIRE(labels && src && bbox && mu && sigma && out, "bfm_gmm_crop: null pointer");
    CropBox bx = {src[1], src[2], bbox[0], bbox[1], bbox[2], bbox[3] - bbox[0], bbox[4] - bbox[1], bbox[5] - bbox[2]};
    BFM_REQUIRE(bx.c0 > 0 && bx.c1 > 0 && bx.c2 > 0 && bbox[0] >= 0 && bbox[1] >= 0 && bbox[2] >= 0 &&
                bbox[3] <= src[0] && bbox[4] <= src[1] && bbox[5] <= src[2], "bfm_gmm_crop: box outside the volume");
    const int64_t n = (int64_t)bx.c0 * bx.c1 * bx.c2;
    k_gmm_crop<<<blocks_for(n, 2), 256, 0, (cudaStream_t)stream>>>(labels, label_is_u8, bx, mu, sigma, eps, seed, out, n);
    return check_launch("bfm_gmm_crop");
}

int bfm_pathol_cerebral(const float *syn, const void *labels, int label_is_u8, const int *src, const int *bbox,
                        float *cerebral, double *sums_dev, void *stream) {
    BFM_REQUIRE(syn && labels && src && bbox && cerebral && sums_dev, "bfm_pathol_cerebral: null pointer");
    CropBox bx = {src[1], src[2], bbox[0], bbox[1], bbox[2], bbox[3] - bbox[0], bbox[4] - bbox[1], bbox[5] - bbox[2]};
    BFM_REQUIRE(bx.c0 > 0 && bx.c1 > 0 && bx.c2 > 0, "bfm_pathol_cerebral: empty box");
    const int64_t n = (int64_t)bx.c0 * bx.c1 * bx.c2;
    cudaMemsetAsync(sums_dev, 0, 4 * sizeof(double), (cudaStream_t)stream);
    k_pathol_cerebral<<<blocks_for(n, 4), 256, 0, (cudaStream_t)stream>>>(syn, labels, label_is_u8, bx, cerebral, sums_dev, n);
    return check_launch("bfm_pathol_cerebral");
}

int bfm_zero_where_zero(void *p, int p_is_double, const float *c, int64_t n, void *stream) {
    BFM_REQUIRE(p && c && n >= 0, "bfm_zero_where_zero: bad argument");
    if (n == 0) return BFM_OK;
    if (p_is_double) k_zero_where_zero<double><<<blocks_for(n, 4), 256, 0, (cudaStream_t)stream>>>((double *)p, c, n);
    else k_zero_where_zero<float><<<blocks_for(n, 4), 256, 0, (cudaStream_t)stream>>>((float *)p, c, n);
    return check_launch("bfm_zero_where_zero");
}

int bfm_masked_mean(const float *I, const void *P, int p_is_double, int64_t n, double *sums_dev, void *stream) {
    BFM_REQUIRE(I && P && sums_dev && n > 0, "bfm_masked_mean: bad argument");
    cudaMemsetAsync(sums_dev, 0, 2 * sizeof(double), (cudaStream_t)stream);
    if (p_is_double) k_masked_mean<double><<<blocks_for(n, 4), 256, 0, (cudaStream_t)stream>>>(I, (const double *)P, sums_dev, n);
    else k_masked_mean<float><<<blocks_for(n, 4), 256, 0, (cudaStream_t)stream>>>(I, (const float *)P, sums_dev, n);
    return check_launch("bfm_masked_mean");
}

int bfm_encode_pathology(float *I, const void *P, const void *Pprob, int p_is_double, const float *mus,
                         const float *sigmas, int n_tab, const float *eps, uint64_t seed, int64_t n, void *stream) {
    BFM_REQUIRE(I && P && Pprob && mus && sigmas && n_tab > 0 && n >= 0, "bfm_encode_pathology: bad argument");
    if (n == 0) return BFM_OK;
    if (p_is_double)
        k_encode_pathology<double><<<blocks_for(n, 2), 256, 0, (cudaStream_t)stream>>>(
            I, (const double *)P, (const double *)Pprob, mus, sigmas, n_tab, eps, seed, n);
    else
        k_encode_pathology<float><<<blocks_for(n, 2), 256, 0, (cudaStream_t)stream>>>(
            I, (const float *)P, (const float *)Pprob, mus, sigmas, n_tab, eps, seed, n);
    return check_launch("bfm_encode_pathology");
}

}  // extern "C"
