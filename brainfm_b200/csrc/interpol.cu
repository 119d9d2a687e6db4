// Spline resampling kernels behind brainfm_b200.interpol (utils/interpol of the reference = torch-interpol 0.2.3).
//   pull / push / count / grad : utils/interpol/nd.py:81-288, iso0.py:24-100, iso1.py:29-387 (forward semantics)
//   bounds                     : utils/interpol/bounds.py:30-89
//   spline weights             : utils/interpol/splines.py:30-160 (orders 0..7)
//   in-bounds mask             : utils/interpol/jit_utils.py:242-255
//   prefilter                  : utils/interpol/coeff.py:35-281
// One thread per sampled point; every channel of the point is processed by that thread so the (order+1)^3 node
// indices / weights are computed once.  Always 3-D: the Python layer pads lower-dimensional problems with
// singleton axes (order 0, coordinate 0).
#include "common.cuh"

namespace bfm {

// ---- Python-style integer helpers ----------------------------------------------------------------
__device__ __forceinline__ int64_t pymod(int64_t a, int64_t n) {
    int64_t r = a % n;
    return r < 0 ? r + n : r;
}
__device__ __forceinline__ int64_t floordiv(int64_t a, int64_t n) {
    int64_t q = a / n;
    return (a % n != 0 && ((a < 0) != (n < 0))) ? q - 1 : q;
}

// Bound.index (bounds.py:30-60)
__device__ __forceinline__ int bound_index(int type, int64_t i, int n) {
    switch (type) {
        case 0: case 1: return (int)(i < 0 ? 0 : (i > n - 1 ? n - 1 : i));
        case 3: case 5: {
            const int64_t n2 = 2 * (int64_t)n;
            i = i < 0 ? n2 - 1 - pymod(-i - 1, n2) : pymod(i, n2);
            return (int)(i >= n ? n2 - 1 - i : i);
        }
        case 2: {
            if (n == 1) return 0;
            const int64_t n2 = 2 * ((int64_t)n - 1);
            i = pymod(i < 0 ? -i : i, n2);
            return (int)(i >= n ? n2 - i : i);
        }
        case 4: {
            const int64_t n2 = 2 * ((int64_t)n + 1);
            i = i < 0 ? -i - 2 : i;
            i = pymod(i, n2);
            i = i > n ? n2 - 2 - i : i;
            if (i == -1) i = 0;
            if (i == n) i = n - 1;
            return (int)i;
        }
        case 6: return (int)pymod(i, n);
        default: return (int)i;
    }
}
// Bound.transform (bounds.py:62-89): sign applied to the gathered value (1 when the reference returns None)
__device__ __forceinline__ int bound_sign(int type, int64_t i, int n) {
    switch (type) {
        case 4: {
            if (n == 1) return 1;
            const int64_t n2 = 2 * ((int64_t)n + 1);
            i = i < 0 ? -i + (n - 1) : i;
            i = pymod(i, n2);
            int x = (i == 0) ? 0 : 1;
            if (pymod(i, n + 1) == n) x = 0;
            i = floordiv(i, n + 1);
            return pymod(i, 2) > 0 ? -x : x;
        }
        case 5: {
            i = i < 0 ? n - 1 - i : i;
            i = floordiv(i, n);
            return pymod(i, 2) > 0 ? -1 : 1;
        }
        case 0: return (i < 0 || i >= n) ? 0 : 1;
        default: return 1;
    }
}

// Spline.fastweight (splines.py:30-79)
template <typename T>
__device__ __forceinline__ T spline_weight(int order, T x) {
    x = x < 0 ? -x : x;
    switch (order) {
        case 0: return T(1);
        case 1: return T(1) - x;
        case 2: return x < T(0.5) ? T(0.75) - x * x : T(0.5) * (T(1.5) - x) * (T(1.5) - x);
        case 3: return x < T(1) ? (x * x * (x - T(2)) * T(3) + T(4)) / T(6) : (T(2) - x) * (T(2) - x) * (T(2) - x) / T(6);
        case 4: {
            if (x < T(0.5)) { T a = x * x; return a * (a * T(0.25) - T(0.625)) + T(115.) / T(192.); }
            if (x < T(1.5)) return x * (x * (x * (T(5) - x) / T(6) - T(1.25)) + T(5.) / T(24.)) + T(55.) / T(96.);
            T a = x - T(2.5); a = a * a; return a * a / T(24);
        }
        case 5: {
            if (x < T(1)) { T a = x * x; return a * (a * (T(0.25) - x / T(12)) - T(0.5)) + T(0.55); }
            if (x < T(2)) return x * (x * (x * (x * (x / T(24) - T(0.375)) + T(1.25)) - T(1.75)) + T(0.625)) + T(0.425);
            T a = T(3) - x; T a2 = a * a; return a2 * a2 * a / T(120);
        }
        case 6: {
            if (x < T(0.5)) { T a = x * x; return a * (a * (T(7.) / T(48.) - a / T(36)) - T(77.) / T(192.)) + T(5887.) / T(11520.); }
            if (x < T(1.5)) return x * (x * (x * (x * (x * (x / T(48) - T(7.) / T(48.)) + T(0.328125)) - T(35.) / T(288.)) - T(91.) / T(256.)) - T(7.) / T(768.)) + T(7861.) / T(15360.);
            if (x < T(2.5)) return x * (x * (x * (x * (x * (T(7.) / T(60.) - x / T(120)) - T(0.65625)) + T(133.) / T(72.)) - T(2.5703125)) + T(1267.) / T(960.)) + T(1379.) / T(7680.);
            T a = x - T(3.5); T a2 = a * a; return a2 * a2 * a2 / T(720);
        }
        default: {  // 7
            if (x < T(1)) { T a = x * x; return a * (a * (a * (x / T(144) - T(1.) / T(36.)) + T(1.) / T(9.)) - T(1.) / T(3.)) + T(151.) / T(315.); }
            if (x < T(2)) return x * (x * (x * (x * (x * (x * (T(0.05) - x / T(240)) - T(7.) / T(30.)) + T(0.5)) - T(7.) / T(18.)) - T(0.1)) - T(7.) / T(90.)) + T(103.) / T(210.);
            if (x < T(3)) return x * (x * (x * (x * (x * (x * (x / T(720) - T(1.) / T(36.)) + T(7.) / T(30.)) - T(19.) / T(18.)) + T(49.) / T(18.)) - T(23.) / T(6.)) + T(217.) / T(90.)) - T(139.) / T(630.);
            T a = T(4) - x; T a2 = a * a; T a4 = a2 * a2; return a4 * a2 * a / T(5040);
        }
    }
}
// Spline.fastgrad (splines.py:90-147): derivative of the weight wrt the distance
template <typename T>
__device__ __forceinline__ T spline_grad(int order, T xs) {
    if (order == 0) return T(0);
    const T sgn = xs > 0 ? T(1) : (xs < 0 ? T(-1) : T(0));
    const T x = xs < 0 ? -xs : xs;
    T g;
    switch (order) {
        case 1: g = T(1); break;
        case 2: g = x < T(0.5) ? T(-2) * x : x - T(1.5); break;
        case 3: g = x < T(1) ? x * (x * T(1.5) - T(2)) : T(-0.5) * (T(2) - x) * (T(2) - x); break;
        case 4:
            if (x < T(0.5)) g = x * (x * x - T(1.25));
            else if (x < T(1.5)) g = x * (x * (x * (T(-2.) / T(3.)) + T(2.5)) - T(2.5)) + T(5.) / T(24.);
            else { T a = T(2) * x - T(5); g = a * a * a / T(48); }
            break;
        case 5:
            if (x < T(1)) g = x * (x * (x * (x * (T(-5.) / T(12.)) + T(1))) - T(1));
            else if (x < T(2)) g = x * (x * (x * (x * (T(5.) / T(24.)) - T(1.5)) + T(3.75)) - T(3.5)) + T(0.625);
            else { T a = x - T(3); a = a * a; g = a * a / T(-24); }
            break;
        case 6:
            if (x < T(0.5)) { T a = x * x; g = x * (a * (T(7.) / T(12.)) - a * a / T(6) - T(77.) / T(96.)); }
            else if (x < T(1.5)) g = x * (x * (x * (x * (x * T(0.125) - T(35.) / T(48.)) + T(1.3125)) - T(35.) / T(96.)) - T(0.7109375)) - T(7.) / T(768.);
            else if (x < T(2.5)) g = x * (x * (x * (x * (x / T(-20) + T(7.) / T(12.)) - T(2.625)) + T(133.) / T(24.)) - T(5.140625)) + T(1267.) / T(960.);
            else { T a = T(2) * x - T(7); T a2 = a * a; g = a2 * a2 * a / T(3840); }
            break;
        default:
            if (x < T(1)) { T a = x * x; g = x * (a * (a * (x * (T(7.) / T(144.)) - T(1.) / T(6.)) + T(4.) / T(9.)) - T(2.) / T(3.)); }
            else if (x < T(2)) g = x * (x * (x * (x * (x * (x * (T(-7.) / T(240.)) + T(3.) / T(10.)) - T(7.) / T(6.)) + T(2)) - T(7.) / T(6.)) - T(1.) / T(5.)) - T(7.) / T(90.);
            else if (x < T(3)) g = x * (x * (x * (x * (x * (x * (T(7.) / T(720.)) - T(1.) / T(6.)) + T(7.) / T(6.)) - T(38.) / T(9.)) + T(49.) / T(6.)) - T(23.) / T(3.)) + T(217.) / T(90.);
            else { T a = x - T(4); T a2 = a * a; g = a2 * a2 * a2 / T(-720); }
            break;
    }
    return g * sgn;
}

struct InterpolArgs {
    int ishape[3];
    int order[3];
    int bound[3];
    int extrapolate;
    int iso;          // 1: every order is 0 => torch.round (half to even) instead of floor(g+0.5) (iso0.py:10-15)
                      // 2: every order is 1 => iso1 gradients (+v1 - v0, iso1.py:269-387); the generic nd path
                      //    uses sign(dist) for order-1 axes (splines.py:90-97), kept as in the reference
    int B, C, Bi, Bg; // output batch, channels, input batch (1 or B), grid batch (1 or B)
    int64_t P;        // points per batch element
};

enum { MODE_PULL = 0, MODE_PUSH = 1, MODE_GRAD = 2 };

template <typename T, int MAXN, int MODE>
__global__ void __launch_bounds__(128) k_interpol(const T *__restrict__ inp, const T *__restrict__ grid,
                                                   T *__restrict__ out, const InterpolArgs a) {
    const int64_t total = (int64_t)a.B * a.P;
    const int64_t vol = (int64_t)a.ishape[0] * a.ishape[1] * a.ishape[2];
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < total; q += (int64_t)gridDim.x * blockDim.x) {
        const int b = (int)(q / a.P);
        const int64_t p = q - (int64_t)b * a.P;
        const T *g = grid + ((int64_t)(a.Bg == 1 ? 0 : b) * a.P + p) * 3;
        int idx[3][MAXN];
        T w[3][MAXN], gw[3][MAXN];
        bool inb = true;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const T x = g[d];
            const int n = a.ishape[d], o = a.order[d];
            if (a.extrapolate != 1) {
                const T thr = a.extrapolate == 2 ? T(0.55) : T(0.05);
                inb = inb && (x > -thr) && (x < T(n - 1) + thr);
            }
            T f0;
            if (o == 0 && a.iso == 1) f0 = rint(x);
            else f0 = floor(x - T(o - 1) / T(2));
            const T dist0 = x - f0;
            const int64_t i0 = (int64_t)f0;
#pragma unroll
            for (int k = 0; k < MAXN; ++k) {
                if (k <= o) {
                    const int sg = bound_sign(a.bound[d], i0 + k, n);
                    idx[d][k] = bound_index(a.bound[d], i0 + k, n);
                    const T dist = dist0 - T(k);
                    w[d][k] = (o == 0 ? T(1) : spline_weight<T>(o, dist)) * T(sg);
                    if (MODE == MODE_GRAD)
                        gw[d][k] = (a.iso == 2 ? (k == 0 ? T(-1) : T(1)) : spline_grad<T>(o, dist)) * T(sg);
                }
            }
        }
        const T m = inb ? T(1) : T(0);
        if (MODE == MODE_PULL || MODE == MODE_GRAD) {
            for (int c = 0; c < a.C; ++c) {
                const T *src = inp + ((int64_t)(a.Bi == 1 ? 0 : b) * a.C + c) * vol;
                T acc = 0, a0 = 0, a1 = 0, a2 = 0;
                for (int kx = 0; kx <= a.order[0]; ++kx)
                    for (int ky = 0; ky <= a.order[1]; ++ky)
                        for (int kz = 0; kz <= a.order[2]; ++kz) {
                            const T v = src[((int64_t)idx[0][kx] * a.ishape[1] + idx[1][ky]) * a.ishape[2] + idx[2][kz]];
                            if (MODE == MODE_PULL) acc += v * w[0][kx] * w[1][ky] * w[2][kz];
                            else {
                                a0 += v * gw[0][kx] * w[1][ky] * w[2][kz];
                                a1 += v * w[0][kx] * gw[1][ky] * w[2][kz];
                                a2 += v * w[0][kx] * w[1][ky] * gw[2][kz];
                            }
                        }
                if (MODE == MODE_PULL) out[((int64_t)b * a.C + c) * a.P + p] = acc * m;
                else {
                    T *o3 = out + (((int64_t)b * a.C + c) * a.P + p) * 3;
                    o3[0] = a0 * m; o3[1] = a1 * m; o3[2] = a2 * m;
                }
            }
        } else {  // push / count (inp == nullptr => ones)
            for (int c = 0; c < a.C; ++c) {
                const T v0 = (inp ? inp[((int64_t)(a.Bi == 1 ? 0 : b) * a.C + c) * a.P + p] : T(1)) * m;
                T *dst = out + ((int64_t)b * a.C + c) * vol;
                for (int kx = 0; kx <= a.order[0]; ++kx)
                    for (int ky = 0; ky <= a.order[1]; ++ky)
                        for (int kz = 0; kz <= a.order[2]; ++kz)
                            atomicAdd(dst + ((int64_t)idx[0][kx] * a.ishape[1] + idx[1][ky]) * a.ishape[2] + idx[2][kz],
                                      v0 * w[0][kx] * w[1][ky] * w[2][kz]);
            }
        }
    }
}

// ---- prefilter: one thread per line ------------------------------------------------------------------
struct FilterArgs {
    int64_t outer, inner;   // tensor viewed as (outer, n, inner); line stride = inner
    int n;
    int bound;              // 0 zero,1 replicate,2 dct1,3 dct2,6 dft
    int npoles;
    double poles[3];
};

template <typename T>
__global__ void k_spline_filter(T *__restrict__ data, const FilterArgs a) {
    const int64_t lines = a.outer * a.inner;
    const int n = a.n;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < lines; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o = q / a.inner, in = q - o * a.inner;
        T *c = data + o * (int64_t)n * a.inner + in;
        const int64_t st = a.inner;
        double gain = 1.0;
        for (int k = 0; k < a.npoles; ++k) gain *= (1.0 - a.poles[k]) * (1.0 - 1.0 / a.poles[k]);
        const T tg = (T)gain;
        for (int i = 0; i < n; ++i) c[i * st] *= tg;                         // coeff.py:265-266
        for (int k = 0; k < a.npoles; ++k) {
            const double pole = a.poles[k];
            const T tp = (T)pole;
            const int max_iter0 = (int)ceil(-30.0 / log(fabs(pole)));
            // ---- initial value (coeff.py:69-177)
            T init;
            if (a.bound == 0 || a.bound == 2) {                               // dct1
                if (max_iter0 < n) {
                    T s = 0, pw = tp;
                    for (int i = 1; i < max_iter0; ++i) { s += c[i * st] * pw; pw *= tp; }
                    init = s + c[0];
                } else {
                    const double polen = pow(pole, (double)(n - 1));
                    T s = 0, pw = tp;
                    for (int i = 1; i < n - 1; ++i) {
                        s += c[i * st] * (pw + (T)(polen * polen) / pw);
                        pw *= tp;
                    }
                    init = s + (c[0] + (T)polen * c[(int64_t)(n - 1) * st]);
                    const double pl = pow(pole, (double)(n - 1));
                    init = init / (T)(1.0 - pl * pl);
                }
            } else if (a.bound == 1 || a.bound == 3) {                        // dct2
                const double polen = pow(pole, (double)n);
                const double pole_last = polen * (1.0 + 1.0 / (pole + polen * polen));
                T s = 0;
                for (int i = 1; i < n - 1; ++i)
                    s += c[i * st] * ((T)pow(pole, (double)i) + (T)pow(pole, (double)(2 * n - 1 - i)));
                T v = s + (c[0] + (T)pole_last * c[(int64_t)(n - 1) * st]);
                v = v * (T)(pole / (1.0 - polen * polen));
                init = v + c[0];
            } else {                                                          // dft
                const int mi = max_iter0 < n ? max_iter0 : n;
                T s = 0;
                for (int i = 1; i < mi; ++i) s += c[(int64_t)(n - i) * st] * (T)pow(pole, (double)i);
                init = (s + c[0]) / (T)(1.0 - pow(pole, (double)mi));
            }
            c[0] = init;
            for (int i = 1; i < n; ++i) c[i * st] = fma(tp, c[(i - 1) * st], c[i * st]);     // coeff.py:272-273
            // ---- final value (coeff.py:181-224)
            T fin;
            const int64_t l = (int64_t)(n - 1) * st;
            if (a.bound == 0 || a.bound == 2) fin = (tp * c[l - st] + c[l]) * (T)(pole / (pole * pole - 1.0));
            else if (a.bound == 1 || a.bound == 3) fin = c[l] * (T)(pole / (pole - 1.0));
            else {
                const int mi = max_iter0 < n ? max_iter0 : n;
                T s = 0;
                for (int i = 0; i < mi - 1; ++i) s += c[i * st] * (T)pow(pole, (double)(i + 2));
                fin = (s + tp * c[l]) / (T)(pow(pole, (double)mi) - 1.0);
            }
            c[l] = fin;
            for (int i = n - 2; i >= 0; --i) c[i * st] = (c[(i + 1) * st] - c[i * st]) * tp;   // coeff.py:277-278
        }
    }
}

template <typename T>
static int launch_interpol(int mode, const void *inp, const void *grid, void *out, const InterpolArgs &a, cudaStream_t s) {
    const int maxo = a.order[0] > a.order[1] ? (a.order[0] > a.order[2] ? a.order[0] : a.order[2])
                                             : (a.order[1] > a.order[2] ? a.order[1] : a.order[2]);
    const int64_t total = (int64_t)a.B * a.P;
    int64_t gsz = (total + 127) / 128;
    if (gsz > 148 * 64) gsz = 148 * 64;
    if (gsz < 1) gsz = 1;
    const T *ip = (const T *)inp, *gp = (const T *)grid;
    T *op = (T *)out;
#define BFM_GO(MAXN)                                                                                     \
    do {                                                                                                 \
        if (mode == MODE_PULL) k_interpol<T, MAXN, MODE_PULL><<<(unsigned)gsz, 128, 0, s>>>(ip, gp, op, a);    \
        else if (mode == MODE_PUSH) k_interpol<T, MAXN, MODE_PUSH><<<(unsigned)gsz, 128, 0, s>>>(ip, gp, op, a); \
        else k_interpol<T, MAXN, MODE_GRAD><<<(unsigned)gsz, 128, 0, s>>>(ip, gp, op, a);                      \
    } while (0)
    if (maxo <= 1) BFM_GO(2);
    else if (maxo <= 3) BFM_GO(4);
    else BFM_GO(8);
#undef BFM_GO
    return check_launch("bfm_interpol");
}
}  // namespace bfm

using namespace bfm;

extern "C" {

int bfm_interpol(int mode, int is_double, const void *inp, const void *grid, void *out, const int *ishape,
                 const int *order, const int *bound, int extrapolate, int iso, int B, int C, int Bi, int Bg,
                 int64_t P, void *stream) {
    BFM_REQUIRE(grid && out && ishape && order && bound, "bfm_interpol: null pointer");
    BFM_REQUIRE(mode >= 0 && mode <= 2, "bfm_interpol: mode must be 0 (pull), 1 (push/count), 2 (grad)");
    BFM_REQUIRE(inp || mode == MODE_PUSH, "bfm_interpol: input required");
    BFM_REQUIRE(B > 0 && C > 0 && P >= 0 && (Bi == 1 || Bi == B) && (Bg == 1 || Bg == B), "bfm_interpol: bad batch");
    InterpolArgs a;
    BFM_REQUIRE(iso >= 0 && iso <= 2, "bfm_interpol: iso must be 0, 1 or 2");
    for (int d = 0; d < 3; ++d) {
        if (ishape[d] <= 0) return fail(BFM_E_INVALID, "%s", "bfm_interpol: non-positive shape");
        if (order[d] < 0 || order[d] > 7) return fail(BFM_E_INVALID, "%s", "bfm_interpol: order must be 0..7");
        if (bound[d] < 0 || bound[d] > 6) return fail(BFM_E_INVALID, "%s", "bfm_interpol: bound must be 0..6");
        a.ishape[d] = ishape[d]; a.order[d] = order[d]; a.bound[d] = bound[d];
    }
    BFM_REQUIRE(extrapolate >= 0 && extrapolate <= 2, "bfm_interpol: extrapolate must be 0, 1 or 2");
    a.extrapolate = extrapolate; a.iso = iso;
    a.B = B; a.C = C; a.Bi = Bi; a.Bg = Bg; a.P = P;
    if (P == 0) return BFM_OK;
    return is_double ? launch_interpol<double>(mode, inp, grid, out, a, (cudaStream_t)stream)
                     : launch_interpol<float>(mode, inp, grid, out, a, (cudaStream_t)stream);
}

int bfm_spline_filter(void *data, int is_double, int64_t outer, int n, int64_t inner, int bound, const double *poles_host,
                      int npoles, void *stream) {
    BFM_REQUIRE(data && outer > 0 && n > 0 && inner > 0, "bfm_spline_filter: bad argument");
    BFM_REQUIRE(npoles >= 0 && npoles <= 3, "bfm_spline_filter: at most 3 poles (order <= 7)");
    if (!(bound == 0 || bound == 1 || bound == 2 || bound == 3 || bound == 6))
        return fail(BFM_E_UNSUPPORTED, "%s", "bfm_spline_filter: only zero/replicate/dct1/dct2/dft are implemented");
    if (n == 1 || npoles == 0) return BFM_OK;
    FilterArgs a;
    a.outer = outer; a.inner = inner; a.n = n; a.bound = bound; a.npoles = npoles;
    for (int k = 0; k < 3; ++k) a.poles[k] = k < npoles ? poles_host[k] : 0.0;
    const int64_t lines = outer * inner;
    int64_t gsz = (lines + 127) / 128;
    if (gsz > 148 * 32) gsz = 148 * 32;
    if (is_double) k_spline_filter<double><<<(unsigned)gsz, 128, 0, (cudaStream_t)stream>>>((double *)data, a);
    else k_spline_filter<float><<<(unsigned)gsz, 128, 0, (cudaStream_t)stream>>>((float *)data, a);
    return check_launch("bfm_spline_filter");
}

}  // extern "C"
