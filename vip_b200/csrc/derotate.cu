// FFT three-shear frame rotation ("vip-fft" semantics) for a whole cube.
//
// Replaces the per-frame Python loop  cube_derotate -> frame_rotate -> rotate_fft -> _fft_shear
// (reference: src/vip_hci/preproc/derotation.py:331-399, 51-328, 542-622, 625-640).
//
// The reference embeds each S x S frame in a 4x zero-padded plane (n4 x n4), optionally
// rot90's it, and applies three shears  x(a) -> y(b) -> x(a); each shear is, per row (or
// column),  IFFT( FFT(line) * exp(-2 pi i * c * u * f) )  with u the centred index of the line
// and f = fftfreq(N) (Nyquist bin = -1/2).  Only a band of the plane is ever non-zero /
// needed, so we run the *pruned* but exact formulation:
//   pass 1: shear-x on rows  [y0, y0+S]      (S+1 rows, real input gathered through the
//           rot90 index map; all N output columns kept, complex)          -> T1[(S+1) x N]
//   pass 2: shear-y on all N columns, input rows [y0, y0+S], output rows [y0, y0+S)
//                                                                           -> T2[S x N]
//   pass 3: shear-x on rows [y0, y0+S), real part of columns [y0, y0+S)     -> out[S x S]
// followed by restoring mask_val at the originally masked pixels.
//
// Two implementations of a 1-D shear:
//   * N = 512..4096 (power of two): register-resident radix-16 FFT, 16 points per thread,
//     DIF forward -> phase multiply in digit-reversed order -> DIT inverse, two swizzled
//     (bank-conflict-free) shared-memory exchanges each way.  fp32 complex.
//   * any other even N (e.g. 402 for 101 x 101 frames): the shear is a circular convolution
//     with the closed-form Dirichlet kernel  D(t) = e^{-i pi t/N} sin(pi t) / (N sin(pi t/N)),
//     evaluated directly on the non-zero support.
#include "common.cuh"
#include <cstdlib>
#include <vector>

namespace vb {

struct RotParams {
    int S;        // frame size (square frames)
    int N;        // working plane size (even)
    int y0;       // offset of the frame block inside the plane (rows and columns)
    int zero_masked;   // 1: pixels equal to mask_val enter the rotation as 0 (interp_zeros with a numeric mask)
    int mask_is_nan;   // 1: masked pixels are the NaNs of the input; 0: pixels == mask_val
    float mask_val;
};

// Value at (i, j) of the zero-padded plane after np.rot90(plane, k) and dropping the last
// row/column (derotation.py:577-599).  P is the (N+1)^2 plane holding the frame at [y0, y0+S)^2.
__device__ __forceinline__ float plane_sample(const float* __restrict__ frame, const RotParams& g,
                                              int k, int i, int j) {
    int r, c;
    if (k == 0)      { r = i;       c = j; }
    else if (k == 1) { r = j;       c = g.N - i; }
    else if (k == 2) { r = g.N - i; c = g.N - j; }
    else             { r = g.N - j; c = i; }
    const int fy = r - g.y0, fx = c - g.y0;
    if (fy < 0 || fy >= g.S || fx < 0 || fx >= g.S) return 0.f;
    const float v = __ldg(frame + (size_t)fy * g.S + fx);
    if (isnan(v)) return 0.f;
    if (g.zero_masked && v == g.mask_val) return 0.f;
    return v;
}

__device__ __forceinline__ bool is_masked(float v, const RotParams& g) {
    return g.mask_is_nan ? isnan(v) : (v == g.mask_val);
}

// split a shift (in pixels, fp64) into nearest integer + fp32 fraction in [-0.5, 0.5]
__device__ __forceinline__ void split_shift(double s, int& s_int, float& s_frac) {
    const double r = rint(s);
    s_int = (int)r;
    s_frac = (float)(s - r);
}

// =====================================================================================
// Power-of-two path: register radix-16 FFT shear
// =====================================================================================

__host__ __device__ constexpr int ilog2c(int x) { return x <= 1 ? 0 : 1 + ilog2c(x >> 1); }
__host__ __device__ constexpr int brev(int x, int bits) {
    int r = 0;
    for (int b = 0; b < bits; ++b) r |= ((x >> b) & 1) << (bits - 1 - b);
    return r;
}

// (xr + i xi) * w16^(SIGN * m),  w16 = exp(+2 pi i / 16), m in [0, 8)
template <int SIGN>
__device__ __forceinline__ void mul_w16(int m, float xr, float xi, float& yr, float& yi) {
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, RH = 0.70710678118654752f;
    constexpr float sg = (float)SIGN;
    switch (m) {
        case 0: yr = xr; yi = xi; break;
        case 4: yr = -sg * xi; yi = sg * xr; break;
        case 2: yr = RH * (xr - sg * xi); yi = RH * (sg * xr + xi); break;
        case 6: yr = RH * (-xr - sg * xi); yi = RH * (sg * xr - xi); break;
        default: {
            float c, s;
            if (m == 1)      { c = C1;  s = S1; }
            else if (m == 3) { c = S1;  s = C1; }
            else if (m == 5) { c = -S1; s = C1; }
            else             { c = -C1; s = S1; }
            s *= sg;
            yr = xr * c - xi * s;
            yi = xr * s + xi * c;
        }
    }
}

// In-register radix-2 decimation-in-frequency DFT of R points at re/im[OFF .. OFF+R):
//   y[k] = sum_j x[j] exp(SIGN * 2 pi i j k / R);  natural order in, bit-reversed order out.
// (template recursion over the stage length keeps every register index a compile-time constant)
template <int R, int SIGN, int OFF, int LEN>
struct DifStage {
    __device__ __forceinline__ static void run(float (&re)[16], float (&im)[16]) {
        constexpr int half = LEN / 2;
#pragma unroll
        for (int s = 0; s < R; s += LEN) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const int i0 = OFF + s + j, i1 = i0 + half;
                const float ar = re[i0], ai = im[i0], br = re[i1], bi = im[i1];
                re[i0] = ar + br; im[i0] = ai + bi;
                mul_w16<SIGN>(j * (16 / LEN), ar - br, ai - bi, re[i1], im[i1]);
            }
        }
        DifStage<R, SIGN, OFF, LEN / 2>::run(re, im);
    }
};
template <int R, int SIGN, int OFF>
struct DifStage<R, SIGN, OFF, 1> {
    __device__ __forceinline__ static void run(float (&)[16], float (&)[16]) {}
};
template <int R, int SIGN, int OFF>
__device__ __forceinline__ void dif(float (&re)[16], float (&im)[16]) {
    DifStage<R, SIGN, OFF, R>::run(re, im);
}

// Decimation-in-time counterpart: bit-reversed order in, natural order out.
template <int R, int SIGN, int OFF, int LEN>
struct DitStage {
    __device__ __forceinline__ static void run(float (&re)[16], float (&im)[16]) {
        constexpr int half = LEN / 2;
#pragma unroll
        for (int s = 0; s < R; s += LEN) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const int i0 = OFF + s + j, i1 = i0 + half;
                float br, bi;
                mul_w16<SIGN>(j * (16 / LEN), re[i1], im[i1], br, bi);
                const float ar = re[i0], ai = im[i0];
                re[i0] = ar + br; im[i0] = ai + bi;
                re[i1] = ar - br; im[i1] = ai - bi;
            }
        }
        DitStage<R, SIGN, OFF, LEN * 2>::run(re, im);
    }
};
template <int R, int SIGN, int OFF>
struct DitStage<R, SIGN, OFF, 2 * R> {
    __device__ __forceinline__ static void run(float (&)[16], float (&)[16]) {}
};
template <int R, int SIGN, int OFF>
__device__ __forceinline__ void dit(float (&re)[16], float (&im)[16]) {
    DitStage<R, SIGN, OFF, 2>::run(re, im);
}

template <int R, int SIGN, int G>
struct GroupFft {
    __device__ __forceinline__ static void fwd(float (&re)[16], float (&im)[16]) {
        dif<R, SIGN, (G - 1) * R>(re, im);
        GroupFft<R, SIGN, G - 1>::fwd(re, im);
    }
    __device__ __forceinline__ static void inv(float (&re)[16], float (&im)[16]) {
        dit<R, SIGN, (G - 1) * R>(re, im);
        GroupFft<R, SIGN, G - 1>::inv(re, im);
    }
};
template <int R, int SIGN>
struct GroupFft<R, SIGN, 0> {
    __device__ __forceinline__ static void fwd(float (&)[16], float (&)[16]) {}
    __device__ __forceinline__ static void inv(float (&)[16], float (&)[16]) {}
};

__device__ __forceinline__ void cmul(float ar, float ai, float br, float bi, float& cr, float& ci) {
    cr = ar * br - ai * bi;
    ci = ar * bi + ai * br;
}

// w[k] = base^k for k = 0..15 where base^{1,2,4,8} are exact table values tw[(m * 2^b) % N]
// (tw[j] = exp(-2 pi i j / N)); CONJ selects exp(+...).
template <int N, bool CONJ>
__device__ __forceinline__ void twiddle_powers(const float2* __restrict__ tw, int m, float (&wr)[16],
                                               float (&wi)[16]) {
    const float2 w1 = __ldg(tw + (m & (N - 1)));
    const float2 w2 = __ldg(tw + ((2 * m) & (N - 1)));
    const float2 w4 = __ldg(tw + ((4 * m) & (N - 1)));
    const float2 w8 = __ldg(tw + ((8 * m) & (N - 1)));
    const float sg = CONJ ? -1.f : 1.f;
    wr[0] = 1.f; wi[0] = 0.f;
    wr[1] = w1.x; wi[1] = sg * w1.y;
    wr[2] = w2.x; wi[2] = sg * w2.y;
    wr[4] = w4.x; wi[4] = sg * w4.y;
    wr[8] = w8.x; wi[8] = sg * w8.y;
    cmul(wr[1], wi[1], wr[2], wi[2], wr[3], wi[3]);
    cmul(wr[1], wi[1], wr[4], wi[4], wr[5], wi[5]);
    cmul(wr[2], wi[2], wr[4], wi[4], wr[6], wi[6]);
    cmul(wr[3], wi[3], wr[4], wi[4], wr[7], wi[7]);
#pragma unroll
    for (int k = 1; k < 8; ++k) cmul(wr[k], wi[k], wr[8], wi[8], wr[8 + k], wi[8 + k]);
}

// exp(-2 pi i * s * k / N) / scale_den with s = s_int + s_frac, exact integer range reduction
template <int N>
__device__ __forceinline__ void phase_of(int s_int, float s_frac, int k, float& pr, float& pi) {
    // turns = s*k/N ;  (s_int*k mod N)/N is exact, the fractional part is < 1/2 turn
    const int ik = (int)(((long long)s_int * k) & (N - 1));
    const float turns = (float)ik * (1.0f / N) + s_frac * ((float)k * (1.0f / N));
    sincospif(-2.0f * turns, &pi, &pr);
}

// barrier among the T threads of one transform only (transforms of a CTA do not wait for each other)
template <int T>
__device__ __forceinline__ void transform_sync(int tr) {
    if (T <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(tr + 1), "r"(T) : "memory");
}

template <int N>
struct ShearFft {
    static constexpr int T = N / 16;      // threads per transform
    static constexpr int R3 = N / 256;    // last radix (2..16)
    static constexpr int L1 = N / 16;
    static constexpr int L2 = R3;
    static constexpr int LOG_R3 = ilog2c(R3);
    static constexpr int LOG_L1 = ilog2c(L1);
    static constexpr int G3 = 16 / R3;    // radix-R3 butterflies per thread in stage 3
    static constexpr int BUF = 2 * N + 8; // floats per transform buffer (re[N], im[N], stagger pad)
    static_assert(N >= 512 && N <= 4096 && (N & (N - 1)) == 0, "FFT path handles N = 512..4096");

    // bank swizzles (validated conflict-free for both sides of each exchange)
    __device__ __forceinline__ static int sw1(int P) {
        return P ^ (((P >> LOG_L1) & ((1 << (5 - LOG_R3)) - 1)) << LOG_R3);
    }
    __device__ __forceinline__ static int sw2(int P) { return P ^ ((P >> 4) & 31); }

    // In:  re/im[j] = x[t + j*T]  (j = 0..15).  Out: re/im[j] = y[t + j*T] where
    //   y = IFFT( FFT(x) * exp(-2 pi i s f) ),  s = s_int + s_frac pixels.
    // sre/sim: this transform's N-float exchange arrays; ph3: R3 complex per-transform phase
    // constants in shared memory.  Must be called by all T threads of transform `tr` (named barrier tr+1).
    __device__ __forceinline__ static void run(float (&re)[16], float (&im)[16], float* sre, float* sim,
                                               float2* ph3, const float2* __restrict__ tw, int t,
                                               int tr, int s_int, float s_frac) {
        float wr[16], wi[16];
        const int npp = t & (L2 - 1), k1p = t >> LOG_R3;

        // per-transform constants: exp(-2 pi i s f(256*k3)) / N, f wraps to negative for k3 >= R3/2
        if (t < R3) {
            const int k3s = (t >= R3 / 2) ? t - R3 : t;          // signed multiple of N/R3
            const int ik = (int)(((long long)s_int * k3s) % R3); // may be negative: fine for sincospi
            const float turns = (float)ik * (1.0f / R3) + s_frac * ((float)k3s * (1.0f / R3));
            float pr, pi;
            sincospif(-2.0f * turns, &pi, &pr);
            ph3[t] = make_float2(pr * (1.0f / N), pi * (1.0f / N));
        }

        // ---- forward stage 1: radix-16 over x[t + j*T]
        dif<16, -1, 0>(re, im);
        twiddle_powers<N, false>(tw, t, wr, wi);
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
            const int r = brev(k1, 4);
            float yr, yi;
            cmul(re[r], im[r], wr[k1], wi[k1], yr, yi);
            const int a = sw1(k1 * L1 + t);
            sre[a] = yr; sim[a] = yi;
        }
        transform_sync<T>(tr);
        // ---- forward stage 2: radix-16 inside each length-L1 block
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int a = sw1(k1p * L1 + j * L2 + npp);
            re[j] = sre[a]; im[j] = sim[a];
        }
        dif<16, -1, 0>(re, im);
        twiddle_powers<N, false>(tw, npp * 16, wr, wi);
        transform_sync<T>(tr);
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
            const int r = brev(k2, 4);
            float yr, yi;
            cmul(re[r], im[r], wr[k2], wi[k2], yr, yi);
            const int a = sw2(k1p * L1 + k2 * L2 + npp);
            sre[a] = yr; sim[a] = yi;
        }
        transform_sync<T>(tr);
        // ---- forward stage 3: radix-R3 on 16 contiguous points, phase, inverse stage 3
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int a = sw2(16 * t + e);
            re[e] = sre[a]; im[e] = sim[a];
        }
        GroupFft<R3, -1, G3>::fwd(re, im);
#pragma unroll
        for (int g = 0; g < G3; ++g) {
            const int q = t * G3 + g;               // = k1*16 + k2
            const int kb = (q >> 4) + 16 * (q & 15);  // k1 + 16*k2  (< 256)
            float er, ei;
            phase_of<N>(s_int, s_frac, kb, er, ei);
#pragma unroll
            for (int k3 = 0; k3 < R3; ++k3) {
                const int r = g * R3 + brev(k3, LOG_R3);
                const float2 p3 = ph3[k3];
                float pr, pi;
                cmul(er, ei, p3.x, p3.y, pr, pi);
                const float xr = re[r], xi = im[r];
                cmul(xr, xi, pr, pi, re[r], im[r]);
            }
        }
        GroupFft<R3, +1, G3>::inv(re, im);
        transform_sync<T>(tr);
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const int a = sw2(16 * t + e);
            sre[a] = re[e]; sim[a] = im[e];
        }
        transform_sync<T>(tr);
        // ---- inverse stage 2
        twiddle_powers<N, true>(tw, npp * 16, wr, wi);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int a = sw2(k1p * L1 + j * L2 + npp);
            const int r = brev(j, 4);
            cmul(sre[a], sim[a], wr[j], wi[j], re[r], im[r]);
        }
        dit<16, +1, 0>(re, im);
        transform_sync<T>(tr);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int a = sw1(k1p * L1 + j * L2 + npp);
            sre[a] = re[j]; sim[a] = im[j];
        }
        transform_sync<T>(tr);
        // ---- inverse stage 1
        twiddle_powers<N, true>(tw, t, wr, wi);
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
            const int a = sw1(k1 * L1 + t);
            const int r = brev(k1, 4);
            cmul(sre[a], sim[a], wr[k1], wi[k1], re[r], im[r]);
        }
        dit<16, +1, 0>(re, im);
        // caller must __syncthreads() before reusing sre/sim
    }
};

// ---- pass 1: rows [y0, y0+S], real gathered input -> T1[(S+1) x N] complex
template <int N, int NT>
__global__ void __launch_bounds__(NT * N / 16)
shear_rows_first_fft(const float* __restrict__ in, float2* __restrict__ T1, RotParams g,
                     const int* __restrict__ krot, const double* __restrict__ a_coef,
                     const float2* __restrict__ tw, int frame0) {
    using F = ShearFft<N>;
    extern __shared__ float smem[];
    __shared__ float2 ph3s[NT][16];
    const int tr = threadIdx.x / F::T, t = threadIdx.x % F::T;
    const int fl = blockIdx.y, f = frame0 + fl;
    const int row = blockIdx.x * NT + tr;
    const bool valid = row <= g.S;
    const int i = g.y0 + row;
    const int k = krot[f];
    const float* frame = in + (size_t)f * g.S * g.S;
    float re[16], im[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int n = t + j * F::T;
        re[j] = (valid && n >= g.y0 && n <= g.y0 + g.S) ? plane_sample(frame, g, k, i, n) : 0.f;
        im[j] = 0.f;
    }
    int s_int; float s_frac;
    split_shift(a_coef[f] * (double)(i - N / 2), s_int, s_frac);
    float* sre = smem + tr * F::BUF;
    F::run(re, im, sre, sre + N, ph3s[tr], tw, t, tr, s_int, s_frac);
    if (valid) {
        float2* dst = T1 + ((size_t)fl * (g.S + 1) + row) * N;
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[t + j * F::T] = make_float2(re[j], im[j]);
    }
}

// ---- pass 2: all N columns; input rows [y0, y0+S] of T1, output rows [y0, y0+S) -> T2[S x N]
template <int N, int NT>
__global__ void __launch_bounds__(NT * N / 16)
shear_cols_fft(const float2* __restrict__ T1, float2* __restrict__ T2, RotParams g,
               const double* __restrict__ b_coef, const float2* __restrict__ tw, int frame0) {
    using F = ShearFft<N>;
    extern __shared__ float smem[];
    __shared__ float2 ph3s[NT][16];
    const int tr = threadIdx.x / F::T, t = threadIdx.x % F::T;
    const int fl = blockIdx.y, f = frame0 + fl;
    const int c0 = blockIdx.x * NT;
    // stage the (S+1) x NT slab with columns fastest (32-byte global segments)
    const float2* src = T1 + (size_t)fl * (g.S + 1) * N + c0;
    for (int idx = threadIdx.x; idx < (g.S + 1) * NT; idx += blockDim.x) {
        const int row = idx / NT, col = idx % NT;
        const float2 v = src[(size_t)row * N + col];
        float* b = smem + col * F::BUF;
        const int a = F::sw1(g.y0 + row);
        b[a] = v.x; b[N + a] = v.y;
    }
    __syncthreads();
    float* sre = smem + tr * F::BUF;
    float re[16], im[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int n = t + j * F::T;
        if (n >= g.y0 && n <= g.y0 + g.S) {
            const int a = F::sw1(n);
            re[j] = sre[a]; im[j] = sre[N + a];
        } else {
            re[j] = 0.f; im[j] = 0.f;
        }
    }
    __syncthreads();
    int s_int; float s_frac;
    split_shift(b_coef[f] * (double)(c0 + tr - N / 2), s_int, s_frac);
    F::run(re, im, sre, sre + N, ph3s[tr], tw, t, tr, s_int, s_frac);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const int n = t + j * F::T;
        if (n >= g.y0 && n < g.y0 + g.S) {
            const int a = F::sw1(n);
            sre[a] = re[j]; sre[N + a] = im[j];
        }
    }
    __syncthreads();
    float2* dst = T2 + (size_t)fl * g.S * N + c0;
    for (int idx = threadIdx.x; idx < g.S * NT; idx += blockDim.x) {
        const int row = idx / NT, col = idx % NT;
        const float* b = smem + col * F::BUF;
        const int a = F::sw1(g.y0 + row);
        dst[(size_t)row * N + col] = make_float2(b[a], b[N + a]);
    }
}

// ---- pass 3: rows [y0, y0+S); real part of columns [y0, y0+S) -> out, mask restored
template <int N, int NT>
__global__ void __launch_bounds__(NT * N / 16)
shear_rows_last_fft(const float2* __restrict__ T2, const float* __restrict__ in, float* __restrict__ out,
                    RotParams g, const double* __restrict__ a_coef, const float2* __restrict__ tw,
                    int frame0) {
    using F = ShearFft<N>;
    extern __shared__ float smem[];
    __shared__ float2 ph3s[NT][16];
    const int tr = threadIdx.x / F::T, t = threadIdx.x % F::T;
    const int fl = blockIdx.y, f = frame0 + fl;
    const int row = blockIdx.x * NT + tr;
    const bool valid = row < g.S;
    const int i = g.y0 + row;
    float re[16], im[16];
    if (valid) {
        const float2* src = T2 + ((size_t)fl * g.S + row) * N;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float2 v = src[t + j * F::T];
            re[j] = v.x; im[j] = v.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) { re[j] = 0.f; im[j] = 0.f; }
    }
    int s_int; float s_frac;
    split_shift(a_coef[f] * (double)(i - N / 2), s_int, s_frac);
    float* sre = smem + tr * F::BUF;
    F::run(re, im, sre, sre + N, ph3s[tr], tw, t, tr, s_int, s_frac);
    if (valid) {
        const float* src = in + ((size_t)f * g.S + row) * g.S;
        float* dst = out + ((size_t)f * g.S + row) * g.S;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int x = t + j * F::T - g.y0;
            if (x >= 0 && x < g.S) dst[x] = is_masked(__ldg(src + x), g) ? g.mask_val : re[j];
        }
    }
}

// =====================================================================================
// Generic path (any even N): direct circular convolution with the Dirichlet kernel
// =====================================================================================

// K[d] = D(d - s), d in [0, N):  D(t) = (1/N) e^{-i pi t / N} sin(pi t) / sin(pi t / N)
__device__ __forceinline__ void dirichlet_table(float* kr, float* ki, int N, int s_int, float s_frac) {
    const float invN = 1.0f / (float)N;
    const float num0 = -sinpif(s_frac);  // sin(pi (e - s_frac)) = -(-1)^e sin(pi s_frac)
    for (int d = threadIdx.x; d < N; d += blockDim.x) {
        int e = (d - s_int) % N;
        if (e < 0) e += N;
        if (e > N / 2) e -= N;              // D is N-periodic (N even): keep |x/N| <= 1/2 for accuracy
        const float x = (float)e - s_frac;  // in [-N/2-0.5, N/2+0.5]
        float ratio;
        if (fabsf(x) < 1e-4f) {
            const float px = 3.14159265358979f * x;
            ratio = (float)N * (1.0f - px * px * (1.0f / 6.0f));
        } else {
            const float num = (e & 1) ? -num0 : num0;
            ratio = num / sinpif(x * invN);
        }
        float sn, cs;
        sincospif(x * invN, &sn, &cs);
        kr[d] = ratio * cs * invN;
        ki[d] = -ratio * sn * invN;
    }
}

// MODE 0: rows first, MODE 1: columns, MODE 2: rows last.  One CTA per 1-D line.
template <int MODE>
__global__ void __launch_bounds__(128)
shear_direct(const float* __restrict__ in, float* __restrict__ out, float2* __restrict__ T1,
             float2* __restrict__ T2, RotParams g, const int* __restrict__ krot,
             const double* __restrict__ a_coef, const double* __restrict__ b_coef, int frame0) {
    extern __shared__ float smem[];
    const int N = g.N, S = g.S, y0 = g.y0;
    float* kr = smem;
    float* ki = kr + N;
    float* xr = ki + N;
    float* xi = xr + N;
    const int fl = blockIdx.y, f = frame0 + fl;
    const int line = blockIdx.x;
    double coef; int u;
    if (MODE == 1) { coef = b_coef[f]; u = line - N / 2; }
    else           { coef = a_coef[f]; u = y0 + line - N / 2; }
    int s_int; float s_frac;
    split_shift(coef * (double)u, s_int, s_frac);
    dirichlet_table(kr, ki, N, s_int, s_frac);

    int n0, L, m0, M;  // input support [n0, n0+L), outputs [m0, m0+M)
    if (MODE == 0) {
        n0 = y0; L = S + 1; m0 = 0; M = N;
        const float* frame = in + (size_t)f * S * S;
        const int k = krot[f];
        for (int n = threadIdx.x; n < L; n += blockDim.x) {
            xr[n] = plane_sample(frame, g, k, y0 + line, n0 + n);
            xi[n] = 0.f;
        }
    } else if (MODE == 1) {
        n0 = y0; L = S + 1; m0 = y0; M = S;
        const float2* src = T1 + (size_t)fl * (S + 1) * N + line;
        for (int n = threadIdx.x; n < L; n += blockDim.x) {
            const float2 v = src[(size_t)n * N];
            xr[n] = v.x; xi[n] = v.y;
        }
    } else {
        n0 = 0; L = N; m0 = y0; M = S;
        const float2* src = T2 + ((size_t)fl * S + line) * N;
        for (int n = threadIdx.x; n < L; n += blockDim.x) {
            const float2 v = src[n];
            xr[n] = v.x; xi[n] = v.y;
        }
    }
    __syncthreads();
    for (int mm = threadIdx.x; mm < M; mm += blockDim.x) {
        const int m = m0 + mm;
        int d = (m - n0) % N;
        if (d < 0) d += N;
        float accr = 0.f, acci = 0.f;
        for (int n = 0; n < L; ++n) {
            const float a = xr[n], b = xi[n], c = kr[d], e = ki[d];
            accr = fmaf(a, c, accr);
            acci = fmaf(a, e, acci);
            if (MODE != 0) {
                accr = fmaf(-b, e, accr);
                acci = fmaf(b, c, acci);
            }
            d = (d == 0) ? N - 1 : d - 1;
        }
        if (MODE == 0) {
            T1[((size_t)fl * (S + 1) + line) * N + m] = make_float2(accr, acci);
        } else if (MODE == 1) {
            T2[((size_t)fl * S + mm) * N + line] = make_float2(accr, acci);
        } else {
            const size_t o = ((size_t)f * S + line) * S + mm;
            out[o] = is_masked(__ldg(in + o), g) ? g.mask_val : accr;
        }
    }
}

// =====================================================================================
// host side
// =====================================================================================

// Optional per-kernel timing (bench.py): when enabled, CUDA events bracket each of the three
// shear kernels on the launching stream; vb_profile_read() sums the elapsed times.
struct PassTimer {
    bool on = false;
    std::vector<cudaEvent_t> ev;   // 4 events per chunk: before p1, after p1, after p2, after p3
    void mark(cudaStream_t st) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        ev.push_back(e);
    }
};
static PassTimer g_timer;

void profile_enable(int on) {
    for (auto e : g_timer.ev) cudaEventDestroy(e);
    g_timer.ev.clear();
    g_timer.on = on != 0;
}

// out[0..2] = total ms of pass 1/2/3, out[3] = number of chunk launches timed
int profile_read(float* out) {
    out[0] = out[1] = out[2] = out[3] = 0.f;
    for (size_t i = 0; i + 3 < g_timer.ev.size(); i += 4) {
        if (cudaEventSynchronize(g_timer.ev[i + 3]) != cudaSuccess) return -1;
        for (int k = 0; k < 3; ++k) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, g_timer.ev[i + k], g_timer.ev[i + k + 1]);
            out[k] += ms;
        }
        out[3] += 1.f;
    }
    return 0;
}

template <int N, int NT>
static int launch_fft_chunk(const float* in, float* out, float2* T1, float2* T2, const RotParams& g,
                            const int* krot, const double* a, const double* b, const float2* tw,
                            int frame0, int nf, cudaStream_t st) {
    using F = ShearFft<N>;
    const size_t smem = (size_t)NT * F::BUF * sizeof(float);
    static bool configured = false;
    if (!configured) {
        VB_CHECK_CUDA(cudaFuncSetAttribute(shear_rows_first_fft<N, NT>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VB_CHECK_CUDA(cudaFuncSetAttribute(shear_cols_fft<N, NT>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VB_CHECK_CUDA(cudaFuncSetAttribute(shear_rows_last_fft<N, NT>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int threads = NT * F::T;
    g_timer.mark(st);
    shear_rows_first_fft<N, NT><<<dim3(ceil_div(g.S + 1, NT), nf), threads, smem, st>>>(
        in, T1, g, krot, a, tw, frame0);
    VB_CHECK_LAUNCH();
    g_timer.mark(st);
    shear_cols_fft<N, NT><<<dim3(N / NT, nf), threads, smem, st>>>(T1, T2, g, b, tw, frame0);
    VB_CHECK_LAUNCH();
    g_timer.mark(st);
    shear_rows_last_fft<N, NT><<<dim3(ceil_div(g.S, NT), nf), threads, smem, st>>>(
        T2, in, out, g, a, tw, frame0);
    VB_CHECK_LAUNCH();
    g_timer.mark(st);
    return 0;
}

static int launch_direct_chunk(const float* in, float* out, float2* T1, float2* T2, const RotParams& g,
                               const int* krot, const double* a, const double* b, int frame0, int nf,
                               cudaStream_t st) {
    const size_t smem = (size_t)4 * g.N * sizeof(float);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        VB_REQUIRE(smem <= 200 * 1024, "derotate: plane size %d too large for the direct path", g.N);
        VB_CHECK_CUDA(cudaFuncSetAttribute(shear_direct<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VB_CHECK_CUDA(cudaFuncSetAttribute(shear_direct<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VB_CHECK_CUDA(cudaFuncSetAttribute(shear_direct<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    g_timer.mark(st);
    shear_direct<0><<<dim3(g.S + 1, nf), 128, smem, st>>>(in, out, T1, T2, g, krot, a, b, frame0);
    VB_CHECK_LAUNCH();
    g_timer.mark(st);
    shear_direct<1><<<dim3(g.N, nf), 128, smem, st>>>(in, out, T1, T2, g, krot, a, b, frame0);
    VB_CHECK_LAUNCH();
    g_timer.mark(st);
    shear_direct<2><<<dim3(g.S, nf), 128, smem, st>>>(in, out, T1, T2, g, krot, a, b, frame0);
    VB_CHECK_LAUNCH();
    g_timer.mark(st);
    return 0;
}

// transforms per CTA for the 2048-point kernels: 2 (two 256-thread CTAs per SM, phases overlap) or 4
static int fft_nt() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VIP_B200_FFT_NT");
        v = (e && atoi(e) == 4) ? 4 : 2;
    }
    return v;
}

size_t derotate_scratch_bytes_per_frame(int S, int N) {
    return ((size_t)(S + 1) * N + (size_t)S * N) * sizeof(float2);
}

int derotate_run(const float* in, float* out, int nframes, const RotParams& g, const int* krot,
                 const double* a, const double* b, const float2* tw, void* scratch,
                 size_t scratch_bytes, int force_direct, int* launches, cudaStream_t st) {
    const size_t per_frame = derotate_scratch_bytes_per_frame(g.S, g.N);
    VB_REQUIRE(scratch_bytes >= per_frame, "derotate: scratch too small (%zu < %zu)", scratch_bytes,
               per_frame);
    int chunk = (int)(scratch_bytes / per_frame);
    if (chunk > nframes) chunk = nframes;
    if (chunk > 65535) chunk = 65535;
    float2* T1 = reinterpret_cast<float2*>(scratch);
    float2* T2 = T1 + (size_t)chunk * (g.S + 1) * g.N;
    const bool pow2 = (g.N & (g.N - 1)) == 0 && g.N >= 512 && g.N <= 4096;
    int nl = 0;
    for (int f0 = 0; f0 < nframes; f0 += chunk) {
        const int nf = (nframes - f0 < chunk) ? nframes - f0 : chunk;
        int rc;
        if (pow2 && !force_direct) {
            switch (g.N) {
                case 512:  rc = launch_fft_chunk<512, 8>(in, out, T1, T2, g, krot, a, b, tw, f0, nf, st); break;
                case 1024: rc = launch_fft_chunk<1024, 4>(in, out, T1, T2, g, krot, a, b, tw, f0, nf, st); break;
                case 2048:
                    if (fft_nt() == 2) rc = launch_fft_chunk<2048, 2>(in, out, T1, T2, g, krot, a, b, tw, f0, nf, st);
                    else rc = launch_fft_chunk<2048, 4>(in, out, T1, T2, g, krot, a, b, tw, f0, nf, st);
                    break;
                default:   rc = launch_fft_chunk<4096, 2>(in, out, T1, T2, g, krot, a, b, tw, f0, nf, st); break;
            }
        } else {
            rc = launch_direct_chunk(in, out, T1, T2, g, krot, a, b, f0, nf, st);
        }
        if (rc) return rc;
        nl += 3;
    }
    if (launches) *launches = nl;
    return 0;
}

}  // namespace vb
