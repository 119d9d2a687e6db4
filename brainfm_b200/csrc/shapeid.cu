// ShapeID kernels: Perlin noise, curl velocity, upwind advection RHS with Neumann boundary, Runge-Kutta stage
// combinations and the dopri5 error ratio.
//   generate_perlin_noise_3d     ShapeID/perlin3d.py:15-90   (numpy float64; reproduced bit for bit)
//   stream_3D / gradient_c       ShapeID/misc.py:66-80, 198-259
//   gradient_f / gradient_b      ShapeID/DiffEqs/pde.py:13-183
//   AdvDiffPDE.forward (adv)     ShapeID/DiffEqs/pde.py:588-640, 301-328, 499-509
//   _runge_kutta_step combos     ShapeID/DiffEqs/rk_common.py:22-61, misc.py:22-25
//   _compute_error_ratio         ShapeID/DiffEqs/misc.py:146-157
// numpy / ATen evaluate every `*` and `+` as a separately rounded operation, so the kernels use the
// non-contracting intrinsics (__dmul_rn, __dadd_rn, __fmul_rn, __fadd_rn).
#include "common.cuh"

namespace bfm {

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dadd_rn(a, -b); }

// t*t*t*(t*(t*6 - 15) + 10)   (perlin3d.py:11-12), numpy evaluation order
__device__ __forceinline__ double fade(double t) {
    return dmul(dmul(dmul(t, t), t), dadd(dmul(t, dsub(dmul(t, 6.0), 15.0)), 10.0));
}

// grad: (r0+1, r1+1, r2+1, 3) unit vectors (tileable copies already applied by the caller)
__global__ void k_perlin3d(const double *__restrict__ grad, int s0, int s1, int s2, int r0, int r1, int r2,
                           double d0, double d1, double d2, double *__restrict__ out) {
    const int64_t total = (int64_t)s0 * s1 * s2;
    const int q0 = s0 / r0, q1 = s1 / r1, q2 = s2 / r2;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(p % s2), j = (int)((p / s2) % s1), i = (int)(p / ((int64_t)s1 * s2));
        // np.mgrid[0:res:delta] % 1  ->  fmod(i*delta, 1) (values are >= 0)
        const double gx = fmod(dmul((double)i, d0), 1.0), gy = fmod(dmul((double)j, d1), 1.0),
                     gz = fmod(dmul((double)k, d2), 1.0);
        const int a = i / q0, b = j / q1, c = k / q2;
        auto ramp = [&](int da, int db, int dc) {
            const double *g = grad + ((((int64_t)(a + da) * (r1 + 1)) + (b + db)) * (r2 + 1) + (c + dc)) * 3;
            // np.sum over the stacked last axis: (x + y) + z
            return dadd(dadd(dmul(dsub(gx, (double)da), g[0]), dmul(dsub(gy, (double)db), g[1])),
                        dmul(dsub(gz, (double)dc), g[2]));
        };
        const double n000 = ramp(0, 0, 0), n100 = ramp(1, 0, 0), n010 = ramp(0, 1, 0), n110 = ramp(1, 1, 0);
        const double n001 = ramp(0, 0, 1), n101 = ramp(1, 0, 1), n011 = ramp(0, 1, 1), n111 = ramp(1, 1, 1);
        const double t0 = fade(gx), t1 = fade(gy), t2 = fade(gz);
        const double u0 = dsub(1.0, t0), u1 = dsub(1.0, t1), u2 = dsub(1.0, t2);
        const double n00 = dadd(dmul(n000, u0), dmul(t0, n100));
        const double n10 = dadd(dmul(n010, u0), dmul(t0, n110));
        const double n01 = dadd(dmul(n001, u0), dmul(t0, n101));
        const double n11 = dadd(dmul(n011, u0), dmul(t0, n111));
        const double n0 = dadd(dmul(u1, n00), dmul(t1, n10));
        const double n1 = dadd(dmul(u1, n01), dmul(t1, n11));
        out[p] = dadd(dmul(u2, n0), dmul(t2, n1));
    }
}

// noise *= (noise >= thr); mask = (noise >= thr)     (perlin3d.py:86-90)
__global__ void k_threshold_mask(double *__restrict__ noise, double *__restrict__ mask, int64_t n, double thr) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        const double m = noise[p] >= thr ? 1.0 : 0.0;
        mask[p] = m;
        noise[p] = dmul(noise[p], m);
    }
}

// one-sided / central differences of gradient_c / gradient_f / gradient_b for a 3-D volume (spacing 1):
// the difference is formed in the input precision and stored as float32, like `dX[...] = ...` in the reference.
template <typename T>
__device__ __forceinline__ float diff_axis(const T *__restrict__ X, int64_t p, int q, int n, int64_t st, int mode) {
    T v;
    if (mode == 0) {         // central (misc.py:243-245)
        if (q == 0) v = X[p + st] - X[p];
        else if (q == n - 1) v = X[p] - X[p - st];
        else v = (X[p + st] - X[p - st]) / T(2);
    } else if (mode == 1) {  // forward (pde.py:48-49)
        v = q == n - 1 ? X[p] - X[p - st] : X[p + st] - X[p];
    } else {                 // backward (pde.py:105-106)
        v = q == 0 ? X[p + st] - X[p] : X[p] - X[p - st];
    }
    return (float)v;
}

template <typename T>
__global__ void k_gradient3d(const T *__restrict__ X, int n0, int n1, int n2, int mode, float i0, float i1, float i2,
                             float *__restrict__ out) {
    const int64_t total = (int64_t)n0 * n1 * n2;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(p % n2), j = (int)((p / n2) % n1), i = (int)(p / ((int64_t)n1 * n2));
        out[p * 3] = __fdiv_rn(diff_axis(X, p, i, n0, (int64_t)n1 * n2, mode), i0);
        out[p * 3 + 1] = __fdiv_rn(diff_axis(X, p, j, n1, n2, mode), i1);
        out[p * 3 + 2] = __fdiv_rn(diff_axis(X, p, k, n2, 1, mode), i2);
    }
}

// stream_3D: V = curl(Phi_a, Phi_b, Phi_c) * multiplier   (misc.py:66-80, perlin3d.py:149-156)
template <typename T>
__global__ void k_curl3d(const T *__restrict__ A, const T *__restrict__ B, const T *__restrict__ Cc, int n0, int n1,
                         int n2, float mult, float *__restrict__ Vx, float *__restrict__ Vy, float *__restrict__ Vz) {
    const int64_t total = (int64_t)n0 * n1 * n2;
    const int64_t s0 = (int64_t)n1 * n2, s1 = n2;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(p % n2), j = (int)((p / n2) % n1), i = (int)(p / s0);
        const float a_y = diff_axis(A, p, j, n1, s1, 0), a_z = diff_axis(A, p, k, n2, 1, 0);
        const float b_x = diff_axis(B, p, i, n0, s0, 0), b_z = diff_axis(B, p, k, n2, 1, 0);
        const float c_x = diff_axis(Cc, p, i, n0, s0, 0), c_y = diff_axis(Cc, p, j, n1, s1, 0);
        Vx[p] = __fmul_rn(__fsub_rn(c_y, b_z), mult);
        Vy[p] = __fmul_rn(__fsub_rn(a_z, c_x), mult);
        Vz[p] = __fmul_rn(__fsub_rn(b_x, a_y), mult);
    }
}

// AdvDiffPDE.forward, perf_pattern 'adv', V_type 'vector_div_free':
//   C <- ReplicationPad3d(1)(C[1:-1,1:-1,1:-1])  (neumann != 0), then
//   out = -(Vx*Cx + Vy*Cy + Vz*Cz) with per-component upwinding (V > 0 -> backward difference)
template <typename T>
__global__ void k_advect_rhs(const T *__restrict__ C, const float *__restrict__ Vx, const float *__restrict__ Vy,
                             const float *__restrict__ Vz, int n0, int n1, int n2, int neumann, float sp0, float sp1,
                             float sp2, float *__restrict__ out) {
    const int64_t total = (int64_t)n0 * n1 * n2;
    const int64_t s0 = (int64_t)n1 * n2, s1 = n2;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < total; p += (int64_t)gridDim.x * blockDim.x) {
        const int k = (int)(p % n2), j = (int)((p / n2) % n1), i = (int)(p / s0);
        auto at = [&](int a, int b, int c) -> T {
            if (neumann) {
                a = min(max(a, 1), n0 - 2); b = min(max(b, 1), n1 - 2); c = min(max(c, 1), n2 - 2);
            }
            return C[(int64_t)a * s0 + (int64_t)b * s1 + c];
        };
        const T c0 = at(i, j, k);
        auto updiff = [&](float v, int q, int n, int da, int db, int dc) -> float {
            // gradient_f / gradient_b of the padded volume, stored as float32
            float df, db_;
            if (q == n - 1) df = (float)(c0 - at(i - da, j - db, k - dc));
            else df = (float)(at(i + da, j + db, k + dc) - c0);
            if (q == 0) db_ = (float)(at(i + da, j + db, k + dc) - c0);
            else db_ = (float)(c0 - at(i - da, j - db, k - dc));
            return v > 0.f ? db_ : df;     // dXf*(1-flag) + dXb*flag with flag in {0,1}
        };
        const float vx = Vx[p], vy = Vy[p], vz = Vz[p];
        const float cx = __fdiv_rn(updiff(vx, i, n0, 1, 0, 0), sp0), cy = __fdiv_rn(updiff(vy, j, n1, 0, 1, 0), sp1),
                    cz = __fdiv_rn(updiff(vz, k, n2, 0, 0, 1), sp2);
        out[p] = -__fadd_rn(__fadd_rn(__fmul_rn(vx, cx), __fmul_rn(vy, cy)), __fmul_rn(vz, cz));
    }
}

// out = y0 + sum_j coef[j] * k_j : the float32 partial sums of _scaled_dot_product, added to the state in the
// state's precision (rk_common.py:47-48).  y0 == NULL: out = the float32 sum itself (error estimate).
struct RkArgs {
    const float *k[8];
    float coef[8];
    int n_terms;
};
template <typename T, typename TO>
__global__ void k_rk_combine(const T *__restrict__ y0, RkArgs a, int64_t n, TO *__restrict__ out) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        float acc = __fmul_rn(a.coef[0], a.k[0][p]);
#pragma unroll 1
        for (int j = 1; j < a.n_terms; ++j) acc = __fadd_rn(acc, __fmul_rn(a.coef[j], a.k[j][p]));
        out[p] = y0 ? (TO)(y0[p] + (T)acc) : (TO)acc;
    }
}

// sum over voxels of (err / (atol + rtol*max(|y0|,|y1|)))^2 in float64 (misc.py:146-157); result[0] += sum
template <typename T>
__global__ void k_rk_error_sum(const float *__restrict__ err, const T *__restrict__ y0, const T *__restrict__ y1,
                               int64_t n, double rtol, double atol, double *__restrict__ result) {
    __shared__ double red[8];
    double s = 0.0;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
        // tolerance in the state's precision, ratio promoted like torch (float32 state stays float32)
        const T tol = (T)atol + (T)rtol * max(abs(y0[p]), abs(y1[p]));
        const T r = (T)err[p] / tol;
        s += (double)(r * r);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        atomicAdd(result, s);
    }
}


// ---- vectorised variants (4 consecutive elements per thread, every stage load issued before the arithmetic) ----
template <int NT>
__device__ __forceinline__ void rk_sum4(const RkArgs &a, int64_t p, float acc[4]) {
    float4 kv[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) kv[j] = __ldg((const float4 *)(a.k[j] + p));
    acc[0] = __fmul_rn(a.coef[0], kv[0].x); acc[1] = __fmul_rn(a.coef[0], kv[0].y);
    acc[2] = __fmul_rn(a.coef[0], kv[0].z); acc[3] = __fmul_rn(a.coef[0], kv[0].w);
#pragma unroll
    for (int j = 1; j < NT; ++j) {
        acc[0] = __fadd_rn(acc[0], __fmul_rn(a.coef[j], kv[j].x)); acc[1] = __fadd_rn(acc[1], __fmul_rn(a.coef[j], kv[j].y));
        acc[2] = __fadd_rn(acc[2], __fmul_rn(a.coef[j], kv[j].z)); acc[3] = __fadd_rn(acc[3], __fmul_rn(a.coef[j], kv[j].w));
    }
}
template <typename T>
__device__ __forceinline__ void load4(const T *p, T v[4]);
template <>
__device__ __forceinline__ void load4<float>(const float *p, float v[4]) {
    const float4 t = __ldg((const float4 *)p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <>
__device__ __forceinline__ void load4<double>(const double *p, double v[4]) {
    const double2 a = __ldg((const double2 *)p), b = __ldg((const double2 *)(p + 2));
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4(float *p, const float v[4]) { *(float4 *)p = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void store4(double *p, const double v[4]) {
    *(double2 *)p = make_double2(v[0], v[1]);
    *(double2 *)(p + 2) = make_double2(v[2], v[3]);
}

template <typename T, int NT, bool HAS_Y0>
__global__ void __launch_bounds__(256) k_rk_combine4(const T *__restrict__ y0, const RkArgs a, int64_t n4,
                                                     T *__restrict__ out) {
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q >= n4) return;
    const int64_t p = 4 * q;
    float acc[4];
    rk_sum4<NT>(a, p, acc);
    T r[4];
    if (HAS_Y0) {
        T y[4];
        load4<T>(y0 + p, y);
#pragma unroll
        for (int e = 0; e < 4; ++e) r[e] = y[e] + (T)acc[e];
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) r[e] = (T)acc[e];
    }
    store4(out + p, r);
}

// error estimate and its scaled square sum in one pass: err = sum_j coef[j]*k_j is never written
// (_runge_kutta_step rk_common.py:50-52 + _compute_error_ratio misc.py:146-157)
template <typename T, int NT>
__global__ void __launch_bounds__(256) k_rk_error4(const RkArgs a, const T *__restrict__ y0, const T *__restrict__ y1,
                                                   int64_t n4, double rtol, double atol, double *__restrict__ result) {
    __shared__ double red[8];
    double s = 0.0;
    const int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (q < n4) {
        const int64_t p = 4 * q;
        float acc[4];
        rk_sum4<NT>(a, p, acc);
        T u[4], v[4];
        load4<T>(y0 + p, u);
        load4<T>(y1 + p, v);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const T tol = (T)atol + (T)rtol * max(abs(u[e]), abs(v[e]));
            const T r = (T)acc[e] / tol;
            s += (double)(r * r);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) s += red[w];
        atomicAdd(result, s);
    }
}

// advection right-hand side.  Block = 64 (k) x 4 (j) voxels of one x plane: no index division, one 32-bit centre
// offset per voxel and six neighbour offsets that are +-1 / +-n2 / +-plane except where the Neumann clamp or the
// volume's edge folds them (the first version spent 347 warp instructions per 32 voxels, mostly 64-bit address
// arithmetic, and ran at 85 % instruction issue: profiles/r2_ncu_advect_summary.csv).  All ten loads are issued
// before anything depends on them; the six neighbours are other threads' centres (L1 / L2 hits).
template <typename T>
__global__ void __launch_bounds__(256) k_advect_rhs_p(const T *__restrict__ C, const float *__restrict__ Vx,
                                                      const float *__restrict__ Vy, const float *__restrict__ Vz, int n0,
                                                      int n1, int n2, int neumann, float sp0, float sp1, float sp2,
                                                      float *__restrict__ out) {
    const int i = blockIdx.z;
    const int j = blockIdx.y * 4 + threadIdx.y, k = blockIdx.x * 64 + threadIdx.x;
    if (j >= n1 || k >= n2) return;
    const int plane = n1 * n2;
    const int lo = neumann ? 1 : 0;
    const int hi0 = neumann ? n0 - 2 : n0 - 1, hi1 = neumann ? n1 - 2 : n1 - 1, hi2 = neumann ? n2 - 2 : n2 - 1;
    // clamped coordinates of the centre and of the six neighbours (set_BC: replicate-padded interior); a neighbour
    // index outside the volume is only ever formed for the side that is not selected below
    const int ic = min(max(i, lo), hi0), jc = min(max(j, lo), hi1), kc = min(max(k, lo), hi2);
    const int im = min(max(max(i - 1, 0), lo), hi0), ip = min(max(min(i + 1, n0 - 1), lo), hi0);
    const int jm = min(max(max(j - 1, 0), lo), hi1), jp = min(max(min(j + 1, n1 - 1), lo), hi1);
    const int km = min(max(max(k - 1, 0), lo), hi2), kp = min(max(min(k + 1, n2 - 1), lo), hi2);
    const int ctr = (ic * n1 + jc) * n2 + kc;                     // < 2^31 (checked on the host)
    const T *__restrict__ Cc = C + ctr;
    const T c0 = __ldg(Cc);
    const T xm = __ldg(Cc + (im - ic) * plane), xp = __ldg(Cc + (ip - ic) * plane);
    const T ym = __ldg(Cc + (jm - jc) * n2), yp = __ldg(Cc + (jp - jc) * n2);
    const T zm = __ldg(Cc + (km - kc)), zp = __ldg(Cc + (kp - kc));
    const int p = (i * n1 + j) * n2 + k;
    const float vx = __ldg(Vx + p), vy = __ldg(Vy + p), vz = __ldg(Vz + p);
    // forward difference at the last index falls back to the backward one and vice versa (gradient_f / gradient_b)
    const bool fx = (vx > 0.f) ? (i == 0) : (i != n0 - 1);
    const bool fy = (vy > 0.f) ? (j == 0) : (j != n1 - 1);
    const bool fz = (vz > 0.f) ? (k == 0) : (k != n2 - 1);
    const float dx = fx ? (float)(xp - c0) : (float)(c0 - xm);
    const float dy = fy ? (float)(yp - c0) : (float)(c0 - ym);
    const float dz = fz ? (float)(zp - c0) : (float)(c0 - zm);
    // x / 1 == x exactly: skip the IEEE division for unit spacing (uniform branch)
    const float cx = sp0 == 1.f ? dx : __fdiv_rn(dx, sp0), cy = sp1 == 1.f ? dy : __fdiv_rn(dy, sp1),
                cz = sp2 == 1.f ? dz : __fdiv_rn(dz, sp2);
    out[p] = -__fadd_rn(__fadd_rn(__fmul_rn(vx, cx), __fmul_rn(vy, cy)), __fmul_rn(vz, cz));
}

// Diffusion right-hand side (AdvDiffPartial.Grad_constantD / Grad_scalarD, ShapeID/DiffEqs/pde.py:331-353) in the
// reference's own composition of one-sided differences:
//   second difference along d  = gradient_b(gradient_f(C)[d])[d]                       (pde.py:551-559)
//   constant D:  D * (ddX + ddY + ddZ)
//   scalar D:    sum_d gradient_c(D)[d] * gradient_c(C)[d]  +  sum_d D * ddC_d
// Every gradient_* result is a float32 buffer divided by the spacing; C is read through the replicate-padded
// interior when `neumann` (set_BC).  accumulate != 0: out += rhs (the advection part is already there).
template <typename T>
__global__ void __launch_bounds__(256) k_diffuse_rhs(const T *__restrict__ C, const float *__restrict__ D, float Dconst,
                                                     int n0, int n1, int n2, int neumann, float sp0, float sp1, float sp2,
                                                     int accumulate, float *__restrict__ out) {
    const int i = blockIdx.y;
    const int plane = n1 * n2;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= plane) return;
    const int j = q / n2, k = q - j * n2;
    const int lo = neumann ? 1 : 0;
    const int hi[3] = {neumann ? n0 - 2 : n0 - 1, neumann ? n1 - 2 : n1 - 1, neumann ? n2 - 2 : n2 - 1};
    const int n[3] = {n0, n1, n2};
    const float sp[3] = {sp0, sp1, sp2};
    const int pos[3] = {i, j, k};
    auto cl = [&](int v, int d) { return neumann ? min(max(v, lo), hi[d]) : v; };
    // C at the point whose coordinate along axis d is m (the other two are this voxel's), through the BC clamp
    auto atC = [&](int d, int m) -> T {
        int a[3] = {cl(pos[0], 0), cl(pos[1], 1), cl(pos[2], 2)};
        a[d] = cl(m, d);
        return __ldg(C + ((int64_t)a[0] * n1 + a[1]) * n2 + a[2]);
    };
    auto atD = [&](int d, int m) -> float {
        int a[3] = {pos[0], pos[1], pos[2]};
        a[d] = m;
        return __ldg(D + ((int64_t)a[0] * n1 + a[1]) * n2 + a[2]);
    };
    // gradient_f along d at index m: forward difference, backward at the last index; float32 result / spacing
    auto gfC = [&](int d, int m) -> float {
        const float v = m != n[d] - 1 ? (float)(atC(d, m + 1) - atC(d, m)) : (float)(atC(d, m) - atC(d, m - 1));
        return __fdiv_rn(v, sp[d]);
    };
    auto gcC = [&](int d, int m) -> float {
        float v;
        if (m == 0) v = (float)(atC(d, 1) - atC(d, 0));
        else if (m == n[d] - 1) v = (float)(atC(d, m) - atC(d, m - 1));
        else v = (float)((atC(d, m + 1) - atC(d, m - 1)) / T(2));
        return __fdiv_rn(v, sp[d]);
    };
    auto gcD = [&](int d, int m) -> float {
        float v;
        if (m == 0) v = __fsub_rn(atD(d, 1), atD(d, 0));
        else if (m == n[d] - 1) v = __fsub_rn(atD(d, m), atD(d, m - 1));
        else v = __fdiv_rn(__fsub_rn(atD(d, m + 1), atD(d, m - 1)), 2.f);
        return __fdiv_rn(v, sp[d]);
    };
    float dd[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int m = pos[d];
        // gradient_b of the float32 forward-difference array: backward difference, forward at index 0
        const float v = m != 0 ? __fsub_rn(gfC(d, m), gfC(d, m - 1)) : __fsub_rn(gfC(d, 1), gfC(d, 0));
        dd[d] = __fdiv_rn(v, sp[d]);
    }
    const int64_t p = (int64_t)i * plane + q;
    float r;
    if (D) {
        const float dv = __ldg(D + p);
        r = __fadd_rn(__fadd_rn(__fmul_rn(gcD(0, i), gcC(0, i)), __fmul_rn(gcD(1, j), gcC(1, j))),
                      __fmul_rn(gcD(2, k), gcC(2, k)));
        r = __fadd_rn(r, __fmul_rn(dv, dd[0]));
        r = __fadd_rn(r, __fmul_rn(dv, dd[1]));
        r = __fadd_rn(r, __fmul_rn(dv, dd[2]));
    } else {
        r = __fmul_rn(Dconst, __fadd_rn(__fadd_rn(dd[0], dd[1]), dd[2]));
    }
    out[p] = accumulate ? __fadd_rn(out[p], r) : r;
}

static inline unsigned g1d(int64_t n) {
    int64_t g = (n + 255) / 256;
    const int64_t cap = 148LL * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}
}  // namespace bfm

using namespace bfm;

extern "C" {

int bfm_perlin3d(const double *grad, const int *shape, const int *res, double *out, void *stream) {
    BFM_REQUIRE(grad && shape && res && out, "bfm_perlin3d: null pointer");
    for (int d = 0; d < 3; ++d) {
        BFM_REQUIRE(shape[d] > 0 && res[d] > 0, "bfm_perlin3d: non-positive shape / res");
        BFM_REQUIRE(shape[d] % res[d] == 0, "bfm_perlin3d: shape must be a multiple of res");
    }
    const int64_t n = (int64_t)shape[0] * shape[1] * shape[2];
    k_perlin3d<<<g1d(n), 256, 0, (cudaStream_t)stream>>>(grad, shape[0], shape[1], shape[2], res[0], res[1], res[2],
                                                         (double)res[0] / shape[0], (double)res[1] / shape[1],
                                                         (double)res[2] / shape[2], out);
    return check_launch("bfm_perlin3d");
}

int bfm_threshold_mask(double *noise, double *mask, int64_t n, double thr, void *stream) {
    BFM_REQUIRE(noise && mask && n > 0, "bfm_threshold_mask: bad argument");
    k_threshold_mask<<<g1d(n), 256, 0, (cudaStream_t)stream>>>(noise, mask, n, thr);
    return check_launch("bfm_threshold_mask");
}

int bfm_gradient3d(const void *X, int is_double, const int *shape, int mode, const float *spacing, float *out,
                   void *stream) {
    BFM_REQUIRE(X && shape && out && spacing, "bfm_gradient3d: null pointer");
    BFM_REQUIRE(mode >= 0 && mode <= 2, "bfm_gradient3d: mode 0 central, 1 forward, 2 backward");
    BFM_REQUIRE(shape[0] > 1 && shape[1] > 1 && shape[2] > 1, "bfm_gradient3d: every axis needs at least 2 samples");
    const int64_t n = (int64_t)shape[0] * shape[1] * shape[2];
    cudaStream_t s = (cudaStream_t)stream;
    if (is_double) k_gradient3d<double><<<g1d(n), 256, 0, s>>>((const double *)X, shape[0], shape[1], shape[2], mode, spacing[0], spacing[1], spacing[2], out);
    else k_gradient3d<float><<<g1d(n), 256, 0, s>>>((const float *)X, shape[0], shape[1], shape[2], mode, spacing[0], spacing[1], spacing[2], out);
    return check_launch("bfm_gradient3d");
}

int bfm_curl3d(const void *A, const void *B, const void *Cc, int is_double, const int *shape, float multiplier,
               float *Vx, float *Vy, float *Vz, void *stream) {
    BFM_REQUIRE(A && B && Cc && shape && Vx && Vy && Vz, "bfm_curl3d: null pointer");
    BFM_REQUIRE(shape[0] > 1 && shape[1] > 1 && shape[2] > 1, "bfm_curl3d: every axis needs at least 2 samples");
    const int64_t n = (int64_t)shape[0] * shape[1] * shape[2];
    cudaStream_t s = (cudaStream_t)stream;
    if (is_double) k_curl3d<double><<<g1d(n), 256, 0, s>>>((const double *)A, (const double *)B, (const double *)Cc, shape[0], shape[1], shape[2], multiplier, Vx, Vy, Vz);
    else k_curl3d<float><<<g1d(n), 256, 0, s>>>((const float *)A, (const float *)B, (const float *)Cc, shape[0], shape[1], shape[2], multiplier, Vx, Vy, Vz);
    return check_launch("bfm_curl3d");
}

int bfm_advect_rhs(const void *C, int is_double, const float *Vx, const float *Vy, const float *Vz, const int *shape,
                   int neumann, const float *spacing, float *out, void *stream) {
    BFM_REQUIRE(C && Vx && Vy && Vz && shape && out && spacing, "bfm_advect_rhs: null pointer");
    BFM_REQUIRE(shape[0] > 2 && shape[1] > 2 && shape[2] > 2, "bfm_advect_rhs: every axis needs at least 3 samples");
    const int64_t n = (int64_t)shape[0] * shape[1] * shape[2];
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t plane = (int64_t)shape[1] * shape[2];
    if (n < (1LL << 31) && shape[0] <= 65535 && (shape[1] + 3) / 4 <= 65535) {
        const dim3 grid((unsigned)((shape[2] + 63) / 64), (unsigned)((shape[1] + 3) / 4), (unsigned)shape[0]);
        const dim3 block(64, 4);
        if (is_double) k_advect_rhs_p<double><<<grid, block, 0, s>>>((const double *)C, Vx, Vy, Vz, shape[0], shape[1], shape[2], neumann, spacing[0], spacing[1], spacing[2], out);
        else k_advect_rhs_p<float><<<grid, block, 0, s>>>((const float *)C, Vx, Vy, Vz, shape[0], shape[1], shape[2], neumann, spacing[0], spacing[1], spacing[2], out);
        return check_launch("bfm_advect_rhs");
    }
    if (is_double) k_advect_rhs<double><<<g1d(n), 256, 0, s>>>((const double *)C, Vx, Vy, Vz, shape[0], shape[1], shape[2], neumann, spacing[0], spacing[1], spacing[2], out);
    else k_advect_rhs<float><<<g1d(n), 256, 0, s>>>((const float *)C, Vx, Vy, Vz, shape[0], shape[1], shape[2], neumann, spacing[0], spacing[1], spacing[2], out);
    return check_launch("bfm_advect_rhs");
}

int bfm_diffuse_rhs(const void *C, int is_double, const float *D, float D_const, const int *shape, int neumann,
                    const float *spacing, int accumulate, float *out, void *stream) {
    BFM_REQUIRE(C && shape && spacing && out, "bfm_diffuse_rhs: null pointer");
    BFM_REQUIRE(shape[0] >= 2 + 2 * (neumann != 0) && shape[1] >= 2 + 2 * (neumann != 0) && shape[2] >= 2 + 2 * (neumann != 0),
                "bfm_diffuse_rhs: every axis needs at least 2 samples (4 with the Neumann boundary)");
    const int64_t plane = (int64_t)shape[1] * shape[2];
    if (plane >= (1LL << 30) || shape[0] > 65535) return fail(BFM_E_UNSUPPORTED, "%s", "bfm_diffuse_rhs: volume too large");
    const dim3 grid((unsigned)((plane + 255) / 256), (unsigned)shape[0]);
    cudaStream_t s = (cudaStream_t)stream;
    if (is_double) k_diffuse_rhs<double><<<grid, 256, 0, s>>>((const double *)C, D, D_const, shape[0], shape[1], shape[2], neumann, spacing[0], spacing[1], spacing[2], accumulate, out);
    else k_diffuse_rhs<float><<<grid, 256, 0, s>>>((const float *)C, D, D_const, shape[0], shape[1], shape[2], neumann, spacing[0], spacing[1], spacing[2], accumulate, out);
    return check_launch("bfm_diffuse_rhs");
}

}  // extern "C" (reopened below)

namespace bfm {
// Dense output of dopri5 (interp.py:5-65): the quartic through y0, y1, y_mid, f0, f1 evaluated at x in [0, 1], with
// the reference's tensor expressions and dtype promotions for a float64 state and float32 stages:
//   a = (-2dt)*f0 + (2dt)*f1 + -8*y0 + -8*y1 + 16*y_mid      ((f32 + f32) promoted to f64 by the first f64 term)
//   b = (5dt)*f0 + (-3dt)*f1 + 18*y0 + 14*y1 + -32*y_mid
//   c = (-4dt)*f0 + dt*f1 + -11*y0 + -5*y1 + 16*y_mid
//   d = dt*f0 (f32),  e = y0
//   out = a*x^4 + b*x^3 + c*x^2 + d*x + e*1                   (d*x in f32: a 0-dim f64 factor does not promote)
// fifteen element-wise passes over 7-10 tensors in the reference's formulation, one here.
__global__ void __launch_bounds__(256) k_dopri5_interp(const double *__restrict__ y0, const double *__restrict__ y1,
                                                       const double *__restrict__ ym, const float *__restrict__ f0,
                                                       const float *__restrict__ f1, double dt, double x, int64_t n,
                                                       double *__restrict__ out) {
    const float m2 = (float)(-2 * dt), p2 = (float)(2 * dt), p5 = (float)(5 * dt), m3 = (float)(-3 * dt),
                m4 = (float)(-4 * dt), p1 = (float)dt, xf = (float)x;
    const double x2 = x * x, x3 = x2 * x, x4 = x3 * x;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
        const double a0 = y0[q], a1 = y1[q], am = ym[q];
        const float g0 = f0[q], g1 = f1[q];
        const double a = __dadd_rn(__dadd_rn(__dadd_rn((double)__fadd_rn(__fmul_rn(m2, g0), __fmul_rn(p2, g1)),
                                                       __dmul_rn(-8.0, a0)), __dmul_rn(-8.0, a1)), __dmul_rn(16.0, am));
        const double b = __dadd_rn(__dadd_rn(__dadd_rn((double)__fadd_rn(__fmul_rn(p5, g0), __fmul_rn(m3, g1)),
                                                       __dmul_rn(18.0, a0)), __dmul_rn(14.0, a1)), __dmul_rn(-32.0, am));
        const double c = __dadd_rn(__dadd_rn(__dadd_rn((double)__fadd_rn(__fmul_rn(m4, g0), __fmul_rn(p1, g1)),
                                                       __dmul_rn(-11.0, a0)), __dmul_rn(-5.0, a1)), __dmul_rn(16.0, am));
        const float d = __fmul_rn(p1, g0);
        double r = __dadd_rn(__dmul_rn(a, x4), __dmul_rn(b, x3));
        r = __dadd_rn(r, __dmul_rn(c, x2));
        r = __dadd_rn(r, (double)__fmul_rn(d, xf));
        out[q] = __dadd_rn(r, __dmul_rn(a0, 1.0));
    }
}
}  // namespace bfm

extern "C" {

int bfm_dopri5_interp(const double *y0, const double *y1, const double *y_mid, const float *f0, const float *f1,
                      double dt, double x, int64_t n, double *out, void *stream) {
    BFM_REQUIRE(y0 && y1 && y_mid && f0 && f1 && out && n > 0, "bfm_dopri5_interp: bad argument");
    const int64_t g = (n + 255) / 256;
    k_dopri5_interp<<<(unsigned)(g < 148 * 32 ? g : 148 * 32), 256, 0, (cudaStream_t)stream>>>(y0, y1, y_mid, f0, f1, dt, x,
                                                                                              n, out);
    return check_launch("bfm_dopri5_interp");
}

int bfm_rk_combine(const void *y0, int is_double, const float *const *k_host, const float *coef_host, int n_terms,
                   int64_t n, void *out, int out_is_double, void *stream) {
    BFM_REQUIRE(k_host && coef_host && out && n > 0, "bfm_rk_combine: bad argument");
    BFM_REQUIRE(n_terms >= 1 && n_terms <= 8, "bfm_rk_combine: 1..8 terms");
    RkArgs a;
    a.n_terms = n_terms;
    for (int j = 0; j < 8; ++j) {
        a.k[j] = j < n_terms ? k_host[j] : nullptr;
        a.coef[j] = j < n_terms ? coef_host[j] : 0.f;
        if (j < n_terms && !k_host[j]) return fail(BFM_E_INVALID, "%s", "bfm_rk_combine: null stage");
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (y0) {
        BFM_REQUIRE((is_double != 0) == (out_is_double != 0), "bfm_rk_combine: state and output precision must agree");
    }
    bool vec = (n % 4) == 0 && ((uintptr_t)out % 16) == 0 && (!y0 || ((uintptr_t)y0 % 16) == 0);
    for (int j = 0; j < n_terms; ++j) vec = vec && ((uintptr_t)k_host[j] % 16) == 0;
    if (vec && !(y0 == nullptr && out_is_double)) {
        const int64_t n4 = n / 4;
        const unsigned g = (unsigned)((n4 + 255) / 256);
#define BFM_RK(NT)                                                                                                   \
    case NT:                                                                                                         \
        if (!y0) k_rk_combine4<float, NT, false><<<g, 256, 0, s>>>(nullptr, a, n4, (float *)out);                    \
        else if (is_double) k_rk_combine4<double, NT, true><<<g, 256, 0, s>>>((const double *)y0, a, n4, (double *)out); \
        else k_rk_combine4<float, NT, true><<<g, 256, 0, s>>>((const float *)y0, a, n4, (float *)out);               \
        break;
        switch (n_terms) { BFM_RK(1) BFM_RK(2) BFM_RK(3) BFM_RK(4) BFM_RK(5) BFM_RK(6) BFM_RK(7) BFM_RK(8) }
#undef BFM_RK
        return check_launch("bfm_rk_combine");
    }
    if (!y0) {
        BFM_REQUIRE(!out_is_double, "bfm_rk_combine: the bare float32 sum is written as float32");
        k_rk_combine<float, float><<<g1d(n), 256, 0, s>>>(nullptr, a, n, (float *)out);
    } else if (is_double) {
        BFM_REQUIRE(out_is_double, "bfm_rk_combine: float64 state needs float64 output");
        k_rk_combine<double, double><<<g1d(n), 256, 0, s>>>((const double *)y0, a, n, (double *)out);
    } else {
        BFM_REQUIRE(!out_is_double, "bfm_rk_combine: float32 state needs float32 output");
        k_rk_combine<float, float><<<g1d(n), 256, 0, s>>>((const float *)y0, a, n, (float *)out);
    }
    return check_launch("bfm_rk_combine");
}

int bfm_rk_error_fused(const float *const *k_host, const float *coef_host, int n_terms, const void *y0, const void *y1,
                       int is_double, int64_t n, double rtol, double atol, float *err_scratch, double *result_dev,
                       void *stream) {
    BFM_REQUIRE(k_host && coef_host && y0 && y1 && result_dev && n > 0, "bfm_rk_error_fused: bad argument");
    BFM_REQUIRE(n_terms >= 1 && n_terms <= 8, "bfm_rk_error_fused: 1..8 terms");
    RkArgs a;
    a.n_terms = n_terms;
    bool vec = (n % 4) == 0 && ((uintptr_t)y0 % 16) == 0 && ((uintptr_t)y1 % 16) == 0;
    for (int j = 0; j < 8; ++j) {
        a.k[j] = j < n_terms ? k_host[j] : nullptr;
        a.coef[j] = j < n_terms ? coef_host[j] : 0.f;
        if (j < n_terms && !k_host[j]) return fail(BFM_E_INVALID, "%s", "bfm_rk_error_fused: null stage");
        if (j < n_terms) vec = vec && ((uintptr_t)k_host[j] % 16) == 0;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (!vec) {                                   // unaligned / ragged: the two-kernel form through err_scratch
        BFM_REQUIRE(err_scratch, "bfm_rk_error_fused: err_scratch needed for unaligned or ragged inputs");
        int rc = bfm_rk_combine(nullptr, 0, k_host, coef_host, n_terms, n, err_scratch, 0, stream);
        if (rc) return rc;
        return bfm_rk_error_sum(err_scratch, y0, y1, is_double, n, rtol, atol, result_dev, stream);
    }
    cudaMemsetAsync(result_dev, 0, sizeof(double), s);
    const int64_t n4 = n / 4;
    const unsigned g = (unsigned)((n4 + 255) / 256);
#define BFM_RE(NT)                                                                                                  \
    case NT:                                                                                                        \
        if (is_double) k_rk_error4<double, NT><<<g, 256, 0, s>>>(a, (const double *)y0, (const double *)y1, n4, rtol, atol, result_dev); \
        else k_rk_error4<float, NT><<<g, 256, 0, s>>>(a, (const float *)y0, (const float *)y1, n4, rtol, atol, result_dev); \
        break;
    switch (n_terms) { BFM_RE(1) BFM_RE(2) BFM_RE(3) BFM_RE(4) BFM_RE(5) BFM_RE(6) BFM_RE(7) BFM_RE(8) }
#undef BFM_RE
    return check_launch("bfm_rk_error_fused");
}

int bfm_rk_error_sum(const float *err, const void *y0, const void *y1, int is_double, int64_t n, double rtol,
                     double atol, double *result_dev, void *stream) {
    BFM_REQUIRE(err && y0 && y1 && result_dev && n > 0, "bfm_rk_error_sum: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    cudaMemsetAsync(result_dev, 0, sizeof(double), s);
    if (is_double) k_rk_error_sum<double><<<g1d(n), 256, 0, s>>>(err, (const double *)y0, (const double *)y1, n, rtol, atol, result_dev);
    else k_rk_error_sum<float><<<g1d(n), 256, 0, s>>>(err, (const float *)y0, (const float *)y1, n, rtol, atol, result_dev);
    return check_launch("bfm_rk_error_sum");
}

}  // extern "C"
