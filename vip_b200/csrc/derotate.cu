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
#include "fft_packed.cuh"
#include <cstdlib>
#include <type_traits>
#include <vector>

namespace vb {

// Scattered output of the last pass (sharded path): row r of local frame f is written to
//   base[r / rows_per] + (fofs + f) * fstride + (r % rows_per) * S
// i.e. straight into the pixel-shard slab (n x p_g) of the rank that owns those rows -- local or PEER memory over
// NVLink -- so the frame->pixel exchange before the temporal collapse rides on the derotation.  nshards = 0: plain
// (n, S, S) output.
struct OutMap {
    float* base[8];
    int nshards;
    int rows_per;
    long long fstride;
    int fofs;
};

struct RotParams {
    int S;        // frame size (square frames)
    int N;        // working plane size (even)
    int y0;       // offset of the frame block inside the plane (rows and columns)
    int zero_masked;   // 1: pixels equal to mask_val enter the rotation as 0 (interp_zeros with a numeric mask)
    int mask_is_nan;   // 1: masked pixels are the NaNs of the input; 0: pixels == mask_val
    float mask_val;
    OutMap om;
};

// address of output row `row` of local frame f (see OutMap)
__device__ __forceinline__ float* out_row_ptr(float* out, const RotParams& g, int f, int row) {
    if (g.om.nshards == 0) return out + ((size_t)f * g.S + row) * g.S;
    const int h = row / g.om.rows_per;
    return g.om.base[h] + (size_t)(g.om.fofs + f) * (size_t)g.om.fstride + (size_t)(row - h * g.om.rows_per) * g.S;
}

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

// Offset of plane pixel (i, j) inside the frame (same mapping as plane_sample), or -1 outside the frame.
__device__ __forceinline__ int plane_offset(const RotParams& g, int k, int i, int j) {
    int r, c;
    if (k == 0)      { r = i;       c = j; }
    else if (k == 1) { r = j;       c = g.N - i; }
    else if (k == 2) { r = g.N - i; c = g.N - j; }
    else             { r = g.N - j; c = i; }
    const int fy = r - g.y0, fx = c - g.y0;
    if (fy < 0 || fy >= g.S || fx < 0 || fx >= g.S) return -1;
    return fy * g.S + fx;
}

// what enters the rotation for a raw frame sample (NaNs and, optionally, masked pixels count as 0)
__device__ __forceinline__ float clean_sample(float v, const RotParams& g) {
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

// exp(-2 pi i * s * k / N) with s = s_int + s_frac, exact integer range reduction
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

// Inter-stage twiddles w^k, k = 0..15, for a thread-dependent base w = exp(-/+ 2 pi i m / N), applied
// as w^(4a+b) = (w^4)^a (w)^b from six table/derived values (few live registers).
struct Twiddle6 {
    float b1r, b1i, b2r, b2i, b3r, b3i;   // w^1, w^2, w^3
    float a1r, a1i, a2r, a2i, a3r, a3i;   // w^4, w^8, w^12
    template <int N, bool CONJ>
    __device__ __forceinline__ void load(const float2* __restrict__ tw, int m) {
        const float2 w1 = __ldg(tw + (m & (N - 1)));
        const float2 w2 = __ldg(tw + ((2 * m) & (N - 1)));
        const float2 w4 = __ldg(tw + ((4 * m) & (N - 1)));
        const float2 w8 = __ldg(tw + ((8 * m) & (N - 1)));
        const float sg = CONJ ? -1.f : 1.f;
        b1r = w1.x; b1i = sg * w1.y;
        b2r = w2.x; b2i = sg * w2.y;
        a1r = w4.x; a1i = sg * w4.y;
        a2r = w8.x; a2i = sg * w8.y;
        cmul(b1r, b1i, b2r, b2i, b3r, b3i);
        cmul(a1r, a1i, a2r, a2i, a3r, a3i);
    }
    // (xr, xi) *= w^k, k compile-time after unrolling
    __device__ __forceinline__ void apply(int k, float& xr, float& xi) const {
        const int a = k >> 2, b = k & 3;
        float tr_, ti_;
        if (b == 1)      { cmul(xr, xi, b1r, b1i, tr_, ti_); xr = tr_; xi = ti_; }
        else if (b == 2) { cmul(xr, xi, b2r, b2i, tr_, ti_); xr = tr_; xi = ti_; }
        else if (b == 3) { cmul(xr, xi, b3r, b3i, tr_, ti_); xr = tr_; xi = ti_; }
        if (a == 1)      { cmul(xr, xi, a1r, a1i, tr_, ti_); xr = tr_; xi = ti_; }
        else if (a == 2) { cmul(xr, xi, a2r, a2i, tr_, ti_); xr = tr_; xi = ti_; }
        else if (a == 3) { cmul(xr, xi, a3r, a3i, tr_, ti_); xr = tr_; xi = ti_; }
    }
};

// radix-4 DFT of (x0..x3): y[a] = sum_j x_j exp(SIGN * 2 pi i j a / 4), natural order in and out
template <int SIGN>
__device__ __forceinline__ void dft4(float& x0r, float& x0i, float& x1r, float& x1i, float& x2r, float& x2i,
                                     float& x3r, float& x3i) {
    const float s0r = x0r + x2r, s0i = x0i + x2i, d0r = x0r - x2r, d0i = x0i - x2i;
    const float s1r = x1r + x3r, s1i = x1i + x3i, d1r = x1r - x3r, d1i = x1i - x3i;
    // SIGN * i * d1
    const float er = -(float)SIGN * d1i, ei = (float)SIGN * d1r;
    x0r = s0r + s1r; x0i = s0i + s1i;
    x2r = s0r - s1r; x2i = s0i - s1i;
    x1r = d0r + er;  x1i = d0i + ei;
    x3r = d0r - er;  x3i = d0i - ei;
}

// Forward radix-16 DFT when only inputs 0..3 are non-zero:  y[4a+b] = sum_j (x_j w16^(-jb)) (-i)^(ja).
// In: re/im[0..3].  Out: y[k] in register k (natural order).
__device__ __forceinline__ void dft16_in4(float (&re)[16], float (&im)[16]) {
    const float x0r = re[0], x0i = im[0], x1r = re[1], x1i = im[1], x2r = re[2], x2i = im[2], x3r = re[3],
                x3i = im[3];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        float z0r = x0r, z0i = x0i, z1r, z1i, z2r, z2i, z3r, z3i;
        // w16^(-j*b): exponents j*b in {0..9}; mul_w16<-1>(m, ...) handles m in [0,8); m = 9 -> -(m=1)
        mul_w16<-1>((1 * b) & 7, x1r, x1i, z1r, z1i);
        mul_w16<-1>((2 * b) & 7, x2r, x2i, z2r, z2i);
        if (3 * b < 8) mul_w16<-1>(3 * b, x3r, x3i, z3r, z3i);
        else { mul_w16<-1>(3 * b - 8, x3r, x3i, z3r, z3i); z3r = -z3r; z3i = -z3i; }
        dft4<-1>(z0r, z0i, z1r, z1i, z2r, z2i, z3r, z3i);
        re[b] = z0r; im[b] = z0i;
        re[4 + b] = z1r; im[4 + b] = z1i;
        re[8 + b] = z2r; im[8 + b] = z2i;
        re[12 + b] = z3r; im[12 + b] = z3i;
    }
}

// Inverse radix-16 DFT when only outputs 0..3 are needed:  x_j = sum_b w16^(+jb) sum_a v[4a+b] (+i)^(ja).
// In: v[k] in register k (natural order).  Out: re/im[0..3].
__device__ __forceinline__ void dft16_out4(float (&re)[16], float (&im)[16]) {
    float o0r = 0.f, o0i = 0.f, o1r = 0.f, o1i = 0.f, o2r = 0.f, o2i = 0.f, o3r = 0.f, o3i = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        float u0r = re[b], u0i = im[b], u1r = re[4 + b], u1i = im[4 + b], u2r = re[8 + b], u2i = im[8 + b],
              u3r = re[12 + b], u3i = im[12 + b];
        dft4<+1>(u0r, u0i, u1r, u1i, u2r, u2i, u3r, u3i);      // u_j = sum_a v[4a+b] (+i)^(ja)
        float t1r, t1i, t2r, t2i, t3r, t3i;
        mul_w16<+1>((1 * b) & 7, u1r, u1i, t1r, t1i);
        mul_w16<+1>((2 * b) & 7, u2r, u2i, t2r, t2i);
        if (3 * b < 8) mul_w16<+1>(3 * b, u3r, u3i, t3r, t3i);
        else { mul_w16<+1>(3 * b - 8, u3r, u3i, t3r, t3i); t3r = -t3r; t3i = -t3i; }
        o0r += u0r; o0i += u0i;
        o1r += t1r; o1i += t1i;
        o2r += t2r; o2i += t2i;
        o3r += t3r; o3i += t3i;
    }
    re[0] = o0r; im[0] = o0i; re[1] = o1r; im[1] = o1i; re[2] = o2r; im[2] = o2i; re[3] = o3r; im[3] = o3i;
}

// One 1-D shear  y = IFFT( FFT(x) * exp(-2 pi i s f) )  of length N by T = N/16 cooperating threads.
//
// Index convention ("re-indexed" lines): every line is stored starting at the first possibly
// non-zero sample, i.e. position n' = n - y0 (mod N).  A circular shift commutes with the
// (circulant) shear, so the result is the same line re-indexed the same way; with N = 4S the
// non-zero input of passes 1/2 then lives at n' in [0, S] (thread t holds n' = t + j*T: j < 4,
// plus the single sample n' = S owned by thread 0) and passes 2/3 only need n' in [0, S), j < 4.
template <int N>
struct ShearFft {
    static constexpr int T = N / 16;      // threads per transform
    static constexpr int R3 = N / 256;    // last radix (2..16)
    static constexpr int L1 = N / 16;
    static constexpr int L2 = R3;
    static constexpr int LOG_R3 = ilog2c(R3);
    static constexpr int LOG_L1 = ilog2c(L1);
    static constexpr int G3 = 16 / R3;    // radix-R3 butterflies per thread in stage 3
    static constexpr int BUF = N + 4;     // float2 per transform buffer (+ stagger between transforms)
    static_assert(N >= 512 && N <= 4096 && (N & (N - 1)) == 0, "FFT path handles N = 512..4096");

    // 64-bit shared-memory accesses are served per half-warp: 16 lanes must hit 16 distinct
    // 8-byte banks.  XOR swizzles, validated for both sides of both exchanges (tools/fft_model.py).
    __device__ __forceinline__ static int sw1(int P) {
        if (LOG_R3 < 4) return P ^ (((P >> LOG_L1) & ((1 << (4 - LOG_R3)) - 1)) << LOG_R3);
        return P;
    }
    __device__ __forceinline__ static int sw2(int P) { return P ^ ((P >> 4) & 15); }

    // Stages 2 and 3 (forward and inverse) only exchange data inside one length-L1 block of the buffer,
    // i.e. among the N/256 consecutive threads (<= 16, same warp) that own it: a warp-level barrier
    // orders those exchanges, and only the two stage-1 transposes need all T threads.
    __device__ __forceinline__ static void group_sync() { __syncwarp(); }

    // IN4 : only re/im[0..3] (n' = t + j*T, j < 4) and the sample n' = 4T (x4, thread 0) are non-zero.
    // OUT4: only outputs j < 4 are produced (re/im[0..3]).
    // Otherwise re/im[j] <-> n' = t + j*T for j = 0..15 on input and output.
    //
    // forward(): stages 1-3; on return register g*R3 + brev(k3) of thread t holds the spectrum at
    //   k = kb + 256*k3,  kb = (q >> 4) + 16*(q & 15),  q = t*G3 + g.
    template <bool IN4>
    __device__ __forceinline__ static void forward(float (&re)[16], float (&im)[16], float2* buf,
                                                   const float2* __restrict__ tw, int t, int tr, float x4r,
                                                   float x4i) {
        const int npp = t & (L2 - 1), k1p = t >> LOG_R3;
        Twiddle6 w;
        // ---- forward stage 1: radix-16 over n' = t + j*T; result y[k1] -> position k1*L1 + t
        if (IN4) {
            dft16_in4(re, im);
            if (t == 0) {   // the lone sample at n' = 4T adds x4 * (-i)^k1 to every output of thread 0
#pragma unroll
                for (int k1 = 0; k1 < 16; ++k1) {
                    const int q = k1 & 3;
                    re[k1] += (q == 0) ? x4r : (q == 1) ? x4i : (q == 2) ? -x4r : -x4i;
                    im[k1] += (q == 0) ? x4i : (q == 1) ? -x4r : (q == 2) ? -x4i : x4r;
                }
            }
        } else {
            dif<16, -1, 0>(re, im);
        }
        w.load<N, false>(tw, t);
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
            const int r = IN4 ? k1 : brev(k1, 4);
            w.apply(k1, re[r], im[r]);
            buf[sw1(k1 * L1 + t)] = make_float2(re[r], im[r]);
        }
        transform_sync<T>(tr);
        // ---- forward stage 2: radix-16 inside each length-L1 block
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float2 v = buf[sw1(k1p * L1 + j * L2 + npp)];
            re[j] = v.x; im[j] = v.y;
        }
        dif<16, -1, 0>(re, im);
        w.load<N, false>(tw, npp * 16);
        group_sync();
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
            const int r = brev(k2, 4);
            w.apply(k2, re[r], im[r]);
            buf[sw2(k1p * L1 + k2 * L2 + npp)] = make_float2(re[r], im[r]);
        }
        group_sync();
        // ---- forward stage 3: radix-R3 on 16 contiguous points
#pragma unroll
        for (int e = 0; e < 16; ++e) {
            const float2 v = buf[sw2(16 * t + e)];
            re[e] = v.x; im[e] = v.y;
        }
        GroupFft<R3, -1, G3>::fwd(re, im);
    }

    // inverse(): undoes forward() (unnormalised); callers that reuse `buf` must synchronise first
    template <bool OUT4>
    __device__ __forceinline__ static void inverse(float (&re)[16], float (&im)[16], float2* buf,
                                                   const float2* __restrict__ tw, int t, int tr) {
        const int npp = t & (L2 - 1), k1p = t >> LOG_R3;
        Twiddle6 w;
        GroupFft<R3, +1, G3>::inv(re, im);
        group_sync();
#pragma unroll
        for (int e = 0; e < 16; ++e) buf[sw2(16 * t + e)] = make_float2(re[e], im[e]);
        // twiddles of the next stage are fetched BEFORE the barrier so that the load latency hides behind it
        w.load<N, true>(tw, npp * 16);
        group_sync();
        // ---- inverse stage 2
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float2 v = buf[sw2(k1p * L1 + j * L2 + npp)];
            const int r = brev(j, 4);
            re[r] = v.x; im[r] = v.y;
            w.apply(j, re[r], im[r]);
        }
        dit<16, +1, 0>(re, im);
        group_sync();
#pragma unroll
        for (int j = 0; j < 16; ++j) buf[sw1(k1p * L1 + j * L2 + npp)] = make_float2(re[j], im[j]);
        w.load<N, true>(tw, t);
        transform_sync<T>(tr);
        // ---- inverse stage 1
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
            const float2 v = buf[sw1(k1 * L1 + t)];
            const int r = OUT4 ? k1 : brev(k1, 4);
            re[r] = v.x; im[r] = v.y;
            w.apply(k1, re[r], im[r]);
        }
        if (OUT4) dft16_out4(re, im);
        else dit<16, +1, 0>(re, im);
    }

    // exp(-2 pi i s f(256*k3)) * scale, f wraps to negative for k3 >= R3/2 (Nyquist bin = -1/2)
    __device__ __forceinline__ static float2 phase3(int k3, int s_int, float s_frac, float scale) {
        const int k3s = (k3 >= R3 / 2) ? k3 - R3 : k3;          // signed multiple of N/R3
        const int ik = (int)(((long long)s_int * k3s) % R3);     // may be negative: fine for sincospi
        const float turns = (float)ik * (1.0f / R3) + s_frac * ((float)k3s * (1.0f / R3));
        float pr, pi;
        sincospif(-2.0f * turns, &pi, &pr);
        return make_float2(pr * scale, pi * scale);
    }

    // One complex line:  y = IFFT( FFT(x) * exp(-2 pi i s f) ).
    template <bool IN4, bool OUT4>
    __device__ __forceinline__ static void run(float (&re)[16], float (&im)[16], float2* buf, float2* ph3,
                                               const float2* __restrict__ tw, int t, int tr, int s_int,
                                               float s_frac, float x4r, float x4i) {
        if (t < R3) ph3[t] = phase3(t, s_int, s_frac, 1.0f / N);   // visible after the stage-1 barrier
        forward<IN4>(re, im, buf, tw, t, tr, x4r, x4i);
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
        inverse<OUT4>(re, im, buf, tw, t, tr);
    }

    // Two REAL lines a (in re[]) and b (in im[]) with their own shifts through ONE complex transform:
    //   Z = FFT(a + i b),  A[k] = (Z[k] + conj Z[N-k]) / 2,  B[k] = (Z[k] - conj Z[N-k]) / (2i),
    //   W[k] = A[k] Pa[k] + i B[k] Pb[k] = Z[k] (Pa+Pb)/2 + conj(Z[N-k]) (Pa-Pb)/2,   w = IFFT(W),
    // so that re[] = Re shear_a(a), im[] = Re shear_b(b) on return.  A shear of a real line is real except
    // for the Nyquist bin (f = -1/2, multiplier e^{i pi s}): its imaginary part is nyq * (-1)^n' with
    // nyq = X[N/2] sin(pi s) / N, returned for both lines (thread 0 only); Pa/Pb carry cos(pi s) at k = N/2.
    // The mirrored bins live in other threads: one extra exchange through `zbuf` (N+4 float2).
    // tools/shear_real_model.py proves the bookkeeping against the oracle in fp64.
    template <bool IN4, bool OUT4>
    __device__ __forceinline__ static void run_pair(float (&re)[16], float (&im)[16], float2* buf, float2* zbuf,
                                                    float2* ph3, const float2* __restrict__ tw, int t, int tr,
                                                    int sa_int, float sa_frac, int sb_int, float sb_frac,
                                                    float x4r, float x4i, float& nyq_a, float& nyq_b) {
        if (t < 2 * R3) {
            const bool second = t >= R3;
            ph3[t] = phase3(second ? t - R3 : t, second ? sb_int : sa_int, second ? sb_frac : sa_frac,
                            0.5f / N);
        }
        forward<IN4>(re, im, buf, tw, t, tr, x4r, x4i);
#pragma unroll
        for (int e = 0; e < 16; ++e) zbuf[sw2(16 * t + e)] = make_float2(re[e], im[e]);
        transform_sync<T>(tr);
        nyq_a = 0.f; nyq_b = 0.f;
#pragma unroll
        for (int g = 0; g < G3; ++g) {
            const int q = t * G3 + g;
            const int kb = (q >> 4) + 16 * (q & 15);
            float ear, eai, ebr, ebi;
            phase_of<N>(sa_int, sa_frac, kb, ear, eai);
            phase_of<N>(sb_int, sb_frac, kb, ebr, ebi);
            // bin N-k: kb' = (256 - kb) mod 256 and k3' = R3-1-k3 (kb != 0) or (R3 - k3) mod R3 (kb == 0)
            const int kbp = (256 - kb) & 255;
            const int basep = R3 * (((kbp & 15) << 4) | (kbp >> 4));
#pragma unroll
            for (int k3 = 0; k3 < R3; ++k3) {
                const int r = g * R3 + brev(k3, LOG_R3);
                const int i1 = brev(R3 - 1 - k3, LOG_R3), i0 = brev((R3 - k3) & (R3 - 1), LOG_R3);
                const float2 zp = zbuf[sw2(basep + (kb != 0 ? i1 : i0))];
                const float2 p3a = ph3[k3], p3b = ph3[R3 + k3];
                float par, pai, pbr, pbi;
                cmul(ear, eai, p3a.x, p3a.y, par, pai);
                cmul(ebr, ebi, p3b.x, p3b.y, pbr, pbi);
                if (k3 == R3 / 2 && kb == 0) {      // Nyquist bin (thread 0 only)
                    nyq_a = 2.f * re[r] * pai;
                    nyq_b = 2.f * im[r] * pbi;
                    pai = 0.f; pbi = 0.f;
                }
                const float hsr = par + pbr, hsi = pai + pbi, hdr = par - pbr, hdi = pai - pbi;
                const float zr = re[r], zi = im[r];
                re[r] = zr * hsr - zi * hsi + zp.x * hdr + zp.y * hdi;
                im[r] = zr * hsi + zi * hsr + zp.x * hdi - zp.y * hdr;
            }
        }
        inverse<OUT4>(re, im, buf, tw, t, tr);
    }

    // One complex line with the multiplier  M[k] = sum_{u=-N/2}^{N/2-1} exp(-2 pi i c u f_k)
    //   = e^{i pi c f} sin(pi c f N) / sin(pi c f)   (the sum of the N column shears of pass 2), fp64 evaluation.
    template <bool IN4, bool OUT4>
    __device__ __forceinline__ static void run_mult(float (&re)[16], float (&im)[16], float2* buf,
                                                    const float2* __restrict__ tw, int t, int tr, double c,
                                                    float x4r, float x4i) {
        forward<IN4>(re, im, buf, tw, t, tr, x4r, x4i);
#pragma unroll
        for (int g = 0; g < G3; ++g) {
            const int q = t * G3 + g;
            const int kb = (q >> 4) + 16 * (q & 15);
#pragma unroll
            for (int k3 = 0; k3 < R3; ++k3) {
                const int r = g * R3 + brev(k3, LOG_R3);
                const int ks = kb + 256 * ((k3 >= R3 / 2) ? k3 - R3 : k3);
                const double cf = c * (double)ks * (1.0 / N);
                double mr = 1.0, mi = 0.0;             // M / N
                double sden, cden;
                sincospi(cf, &sden, &cden);
                if (fabs(sden) > 1e-14) {
                    const double ratio = sinpi(c * (double)ks) / (sden * (double)N);
                    mr = cden * ratio; mi = sden * ratio;
                }
                const float xr = re[r], xi = im[r];
                cmul(xr, xi, (float)mr, (float)mi, re[r], im[r]);
            }
        }
        inverse<OUT4>(re, im, buf, tw, t, tr);
    }
};

// =====================================================================================
// The same transforms on PACKED fp32 pairs (FFMA2 / FADD2 / FMUL2, csrc/fft_packed.cuh): one complex number per
// aligned register pair.  Identical data flow, shared-memory layout, twiddles and phase arithmetic as ShearFft
// above -- only the instruction selection of the complex arithmetic changes (a complex add is one instruction, a
// complex multiply two, +-i a free operand modifier), which halves the FP32 instruction count of kernels that are
// instruction-issue bound.  Selected at run time (VIP_B200_FFT_F32X2, default 1) so both can be timed and tested.
// =====================================================================================
struct Twiddle6P {
    float2 b1, b2, b3, a1, a2, a3;        // w^1, w^2, w^3, w^4, w^8, w^12
    template <int N, bool CONJ>
    __device__ __forceinline__ void load(const float2* __restrict__ tw, int m) {
        const float2 w1 = __ldg(tw + (m & (N - 1)));
        const float2 w2 = __ldg(tw + ((2 * m) & (N - 1)));
        const float2 w4 = __ldg(tw + ((4 * m) & (N - 1)));
        const float2 w8 = __ldg(tw + ((8 * m) & (N - 1)));
        const float sg = CONJ ? -1.f : 1.f;
        b1 = make_float2(w1.x, sg * w1.y);
        b2 = make_float2(w2.x, sg * w2.y);
        a1 = make_float2(w4.x, sg * w4.y);
        a2 = make_float2(w8.x, sg * w8.y);
        b3 = pk::cmul2(b1, b2);
        a3 = pk::cmul2(a1, a2);
    }
    __device__ __forceinline__ void apply(int k, float2& x) const {
        const int a = k >> 2, b = k & 3;
        if (b == 1)      x = pk::cmul2(x, b1);
        else if (b == 2) x = pk::cmul2(x, b2);
        else if (b == 3) x = pk::cmul2(x, b3);
        if (a == 1)      x = pk::cmul2(x, a1);
        else if (a == 2) x = pk::cmul2(x, a2);
        else if (a == 3) x = pk::cmul2(x, a3);
    }
};

// x * w16^(-e) for e in [0, 16): exponents above 7 are the negated lower half
__device__ __forceinline__ float2 mul_w16n(int e, float2 x) {
    if (e < 8) return pk::mul_w16<-1>(e, x);
    const float2 y = pk::mul_w16<-1>(e - 8, x);
    return make_float2(-y.x, -y.y);
}
__device__ __forceinline__ float2 mul_w16p(int e, float2 x) {
    if (e < 8) return pk::mul_w16<+1>(e, x);
    const float2 y = pk::mul_w16<+1>(e - 8, x);
    return make_float2(-y.x, -y.y);
}

// forward radix-16 DFT with inputs 0..3 only (see dft16_in4)
__device__ __forceinline__ void dft16_in4_p(float2 (&z)[16]) {
    const float2 x0 = z[0], x1 = z[1], x2 = z[2], x3 = z[3];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        float2 y0 = x0, y1 = mul_w16n(b, x1), y2 = mul_w16n(2 * b, x2), y3 = mul_w16n(3 * b, x3);
        pk::dft4<-1>(y0, y1, y2, y3);
        z[b] = y0; z[4 + b] = y1; z[8 + b] = y2; z[12 + b] = y3;
    }
}
// inverse radix-16 DFT with outputs 0..3 only (see dft16_out4)
__device__ __forceinline__ void dft16_out4_p(float2 (&z)[16]) {
    float2 o0 = make_float2(0.f, 0.f), o1 = o0, o2 = o0, o3 = o0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        float2 u0 = z[b], u1 = z[4 + b], u2 = z[8 + b], u3 = z[12 + b];
        pk::dft4<+1>(u0, u1, u2, u3);
        o0 = pk::add2(o0, u0);
        o1 = pk::add2(o1, mul_w16p(b, u1));
        o2 = pk::add2(o2, mul_w16p(2 * b, u2));
        o3 = pk::add2(o3, mul_w16p(3 * b, u3));
    }
    z[0] = o0; z[1] = o1; z[2] = o2; z[3] = o3;
}

// PAD = false: the XOR-swizzled buffer layout of ShearFft (bank-conflict free, but every access pays a LOP3 + an
// address add).  PAD = true: a PADDED layout with the same conflict-freedom in which every shared-memory access of a
// thread is `base register + compile-time offset`:
//   exchange A (stage 1 -> 2): element (row k1, column x)      at  k1 * PITCH + x
//   exchange B (stage 2 -> 3): element (block k1p, q in [0,L1)) at  k1p * PITCH + q + (q >> 4)
// with PITCH = 17 * R3 = L1 + L1/16: the 16 contiguous points a thread owns in stage 3 sit at pitch 17 (bank =
// (tt + e) mod 16 over the lanes), and PITCH = R3 mod 16 puts the R3-thread groups of adjacent blocks on disjoint
// banks for the stride-R3 accesses of stage 2 (checked for R3 = 2, 4, 8, 16 in tools/fft_pad_model.py).
template <int N, bool PAD>
struct ShearFftP {
    using B = ShearFft<N>;
    static constexpr int T = B::T, R3 = B::R3, L1 = B::L1, L2 = B::L2, LOG_R3 = B::LOG_R3, G3 = B::G3;
    static constexpr int LOG_L1 = B::LOG_L1;
    static constexpr int PITCH = 17 * R3;
    static constexpr int BUF = PAD ? 16 * PITCH + 4 : B::BUF;

    // exchange A: row k1 (compile-time at the call sites), column x
    __device__ __forceinline__ static int iA(int k1, int x) { return PAD ? k1 * PITCH + x : B::sw1(k1 * L1 + x); }
    // exchange B, stage-2 side: block k1p, element k2 * L2 + npp (k2 compile-time, npp < L2 <= 16, L2 | 16)
    __device__ __forceinline__ static int iB2(int k1p, int k2, int npp) {
        return PAD ? k1p * PITCH + npp + k2 * L2 + ((k2 * L2) >> 4) : B::sw2(k1p * L1 + k2 * L2 + npp);
    }
    // exchange B, stage-3 side: flat element 16 t + e
    __device__ __forceinline__ static int iB3(int t, int e) {
        return PAD ? (t >> LOG_R3) * PITCH + 17 * (t & (R3 - 1)) + e : B::sw2(16 * t + e);
    }
    // exchange-B position of an arbitrary flat element P (base a multiple of R3, i < R3 added by the caller)
    __device__ __forceinline__ static int iBflat(int P) {
        if (!PAD) return B::sw2(P);
        const int q = P & (L1 - 1);
        return (P >> LOG_L1) * PITCH + q + (q >> 4);
    }

    template <bool IN4>
    __device__ __forceinline__ static void forward(float2 (&z)[16], float2* buf, const float2* __restrict__ tw,
                                                   int t, int tr, float2 x4) {
        const int npp = t & (L2 - 1), k1p = t >> LOG_R3;
        Twiddle6P w;
        if (IN4) {
            dft16_in4_p(z);
            if (t == 0) {   // the lone sample at n' = 4T adds x4 * (-i)^k1 to every output of thread 0
                const float2 c1 = make_float2(x4.y, -x4.x), c2 = make_float2(-x4.x, -x4.y), c3 = make_float2(-x4.y, x4.x);
#pragma unroll
                for (int k1 = 0; k1 < 16; ++k1) {
                    const int q = k1 & 3;
                    z[k1] = pk::add2(z[k1], (q == 0) ? x4 : (q == 1) ? c1 : (q == 2) ? c2 : c3);
                }
            }
        } else {
            pk::dif<16, -1, 0>(z);
        }
        w.template load<N, false>(tw, t);
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
            const int r = IN4 ? k1 : brev(k1, 4);
            w.apply(k1, z[r]);
            buf[iA(k1, t)] = z[r];
        }
        transform_sync<T>(tr);
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = buf[iA(k1p, j * L2 + npp)];
        pk::dif<16, -1, 0>(z);
        w.template load<N, false>(tw, npp * 16);
        B::group_sync();
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
            const int r = brev(k2, 4);
            w.apply(k2, z[r]);
            buf[iB2(k1p, k2, npp)] = z[r];
        }
        B::group_sync();
#pragma unroll
        for (int e = 0; e < 16; ++e) z[e] = buf[iB3(t, e)];
        pk::GroupFft<R3, -1, G3>::fwd(z);
    }

    template <bool OUT4>
    __device__ __forceinline__ static void inverse(float2 (&z)[16], float2* buf, const float2* __restrict__ tw,
                                                   int t, int tr) {
        const int npp = t & (L2 - 1), k1p = t >> LOG_R3;
        Twiddle6P w;
        pk::GroupFft<R3, +1, G3>::inv(z);
        B::group_sync();
#pragma unroll
        for (int e = 0; e < 16; ++e) buf[iB3(t, e)] = z[e];
        w.template load<N, true>(tw, npp * 16);
        B::group_sync();
        {
            float2 y[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                float2 v = buf[iB2(k1p, j, npp)];
                w.apply(j, v);
                y[brev(j, 4)] = v;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) z[j] = y[j];
        }
        pk::dit<16, +1, 0>(z);
        B::group_sync();
#pragma unroll
        for (int j = 0; j < 16; ++j) buf[iA(k1p, j * L2 + npp)] = z[j];
        w.template load<N, true>(tw, t);
        transform_sync<T>(tr);
        {
            float2 y[16];
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
                float2 v = buf[iA(k1, t)];
                w.apply(k1, v);
                y[OUT4 ? k1 : brev(k1, 4)] = v;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) z[j] = y[j];
        }
        if (OUT4) dft16_out4_p(z);
        else pk::dit<16, +1, 0>(z);
    }

    // two real lines per complex transform: see ShearFft::run_pair (same bookkeeping, packed arithmetic)
    template <bool IN4, bool OUT4>
    __device__ __forceinline__ static void run_pair(float (&re)[16], float (&im)[16], float2* buf, float2* zbuf,
                                                    float2* ph3, const float2* __restrict__ tw, int t, int tr,
                                                    int sa_int, float sa_frac, int sb_int, float sb_frac,
                                                    float x4r, float x4i, float& nyq_a, float& nyq_b) {
        float2 z[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = make_float2(re[j], im[j]);
        if (t < 2 * R3) {
            const bool second = t >= R3;
            ph3[t] = B::phase3(second ? t - R3 : t, second ? sb_int : sa_int, second ? sb_frac : sa_frac,
                               0.5f / N);
        }
        forward<IN4>(z, buf, tw, t, tr, make_float2(x4r, x4i));
#pragma unroll
        for (int e = 0; e < 16; ++e) zbuf[iB3(t, e)] = z[e];
        transform_sync<T>(tr);
        nyq_a = 0.f; nyq_b = 0.f;
#pragma unroll
        for (int g = 0; g < G3; ++g) {
            const int q = t * G3 + g;
            const int kb = (q >> 4) + 16 * (q & 15);
            float2 ea, eb;
            phase_of<N>(sa_int, sa_frac, kb, ea.x, ea.y);
            phase_of<N>(sb_int, sb_frac, kb, eb.x, eb.y);
            const int kbp = (256 - kb) & 255;
            const int basep = R3 * (((kbp & 15) << 4) | (kbp >> 4));
            const float2* zmir = zbuf + (PAD ? iBflat(basep) : 0);     // PAD: + i below; swizzled: full index below
#pragma unroll
            for (int k3 = 0; k3 < R3; ++k3) {
                const int r = g * R3 + brev(k3, LOG_R3);
                const int i1 = brev(R3 - 1 - k3, LOG_R3), i0 = brev((R3 - k3) & (R3 - 1), LOG_R3);
                const int isel = (kb != 0 ? i1 : i0);
                const float2 zp = PAD ? zmir[isel] : zbuf[B::sw2(basep + isel)];
                float2 pa = pk::cmul2(ea, ph3[k3]);
                float2 pb = pk::cmul2(eb, ph3[R3 + k3]);
                if (k3 == R3 / 2 && kb == 0) {      // Nyquist bin (thread 0 only)
                    nyq_a = 2.f * z[r].x * pa.y;
                    nyq_b = 2.f * z[r].y * pb.y;
                    pa.y = 0.f; pb.y = 0.f;
                }
                const float2 hs = pk::add2(pa, pb), hd = pk::sub2(pa, pb);
                // z * hs + conj(zp) * hd
                float2 acc = pk::cmul2(z[r], hs);
                acc = pk::fma2(pk::bc(zp.x), hd, acc);
                acc = pk::fma2(pk::bc(zp.y), make_float2(hd.y, -hd.x), acc);
                z[r] = acc;
            }
        }
        inverse<OUT4>(z, buf, tw, t, tr);
#pragma unroll
        for (int j = 0; j < 16; ++j) { re[j] = z[j].x; im[j] = z[j].y; }
    }
};

// transform class of a kernel: MODE 0 = scalar arithmetic (round 1), 1 = packed fp32x2, 2 = packed + padded layout
template <int N, int MODE> struct FftSel { using type = ShearFft<N>; };
template <int N> struct FftSel<N, 1> { using type = ShearFftP<N, false>; };
template <int N> struct FftSel<N, 2> { using type = ShearFftP<N, true>; };

// Column label n' of the re-indexed planes T1/T2 <-> physical plane column (n' + y0) mod N.

// ---- pass 1: rows [y0, y0+S], real gathered input -> T1[(S+1) x N] complex (re-indexed columns)
template <int N, int NT, int MINB>
__global__ void __launch_bounds__(NT * N / 16, MINB)
shear_rows_first_fft(const float* __restrict__ in, float2* __restrict__ T1, RotParams g,
                     const int* __restrict__ krot, const double* __restrict__ a_coef,
                     const float2* __restrict__ tw, int frame0) {
    using F = ShearFft<N>;
    extern __shared__ float2 smem2[];
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
    for (int j = 0; j < 4; ++j) {
        re[j] = valid ? plane_sample(frame, g, k, i, g.y0 + t + j * F::T) : 0.f;
        im[j] = 0.f;
    }
    const float x4 = (valid && t == 0) ? plane_sample(frame, g, k, i, g.y0 + g.S) : 0.f;
    int s_int; float s_frac;
    split_shift(a_coef[f] * (double)(i - N / 2), s_int, s_frac);
    F::template run<true, false>(re, im, smem2 + tr * F::BUF, ph3s[tr], tw, t, tr, s_int, s_frac, x4, 0.f);
    if (valid) {
        float2* dst = T1 + ((size_t)fl * (g.S + 1) + row) * N;
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[t + j * F::T] = make_float2(re[j], im[j]);
    }
}

// ---- pass 2: all N columns; input rows [0, S] of T1, output rows [0, S) -> T2[S x N]
template <int N, int NT, int MINB>
__global__ void __launch_bounds__(NT * N / 16, MINB)
shear_cols_fft(const float2* __restrict__ T1, float2* __restrict__ T2, RotParams g,
               const double* __restrict__ b_coef, const float2* __restrict__ tw, int frame0) {
    using F = ShearFft<N>;
    extern __shared__ float2 smem2[];
    __shared__ float2 ph3s[NT][16];
    const int tr = threadIdx.x / F::T, t = threadIdx.x % F::T;
    const int fl = blockIdx.y, f = frame0 + fl;
    const int c0 = blockIdx.x * NT;
    if (NT == 1) {
        // one column per CTA: every thread gathers its own four samples (and thread 0 the extra row S)
        // straight into registers and scatters its four results -- no staging pass, no CTA-wide barriers
        const float2* src1 = T1 + (size_t)fl * (g.S + 1) * N + c0;
        float re[16], im[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 v = __ldg(src1 + (size_t)(t + j * F::T) * N);
            re[j] = v.x; im[j] = v.y;
        }
        float2 v4 = make_float2(0.f, 0.f);
        if (t == 0) v4 = __ldg(src1 + (size_t)(4 * F::T) * N);
        int s_int; float s_frac;
        const int col_phys = (c0 + g.y0) & (N - 1);
        split_shift(b_coef[f] * (double)(col_phys - N / 2), s_int, s_frac);
        F::template run<true, true>(re, im, smem2, ph3s[0], tw, t, 0, s_int, s_frac, v4.x, v4.y);
        float2* dst1 = T2 + (size_t)fl * g.S * N + c0;
#pragma unroll
        for (int j = 0; j < 4; ++j) dst1[(size_t)(t + j * F::T) * N] = make_float2(re[j], im[j]);
        return;
    }
    // stage the (S+1) x NT slab with columns fastest (NT*8-byte global segments)
    const float2* src = T1 + (size_t)fl * (g.S + 1) * N + c0;
    for (int idx = threadIdx.x; idx < (g.S + 1) * NT; idx += blockDim.x) {
        const int row = idx / NT, col = idx % NT;
        smem2[col * F::BUF + F::sw1(row)] = src[(size_t)row * N + col];
    }
    __syncthreads();
    float2* buf = smem2 + tr * F::BUF;
    float re[16], im[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 v = buf[F::sw1(t + j * F::T)];
        re[j] = v.x; im[j] = v.y;
    }
    const float2 v4 = buf[F::sw1(4 * F::T)];   // row S (only thread 0 uses it)
    __syncthreads();
    int s_int; float s_frac;
    const int col_phys = (c0 + tr + g.y0) & (N - 1);
    split_shift(b_coef[f] * (double)(col_phys - N / 2), s_int, s_frac);
    F::template run<true, true>(re, im, buf, ph3s[tr], tw, t, tr, s_int, s_frac, v4.x, v4.y);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) buf[F::sw1(t + j * F::T)] = make_float2(re[j], im[j]);
    __syncthreads();
    float2* dst = T2 + (size_t)fl * g.S * N + c0;
    for (int idx = threadIdx.x; idx < g.S * NT; idx += blockDim.x) {
        const int row = idx / NT, col = idx % NT;
        dst[(size_t)row * N + col] = smem2[col * F::BUF + F::sw1(row)];
    }
}

// ---- pass 2, persistent variant (one transform per CTA at a time, CTAs loop over the columns) -------
// ncu on the one-shot kernel above shows the long scoreboard (global loads: the strided column gather)
// as the top stall.  Here each CTA prefetches the NEXT column with cp.async into a staging line while
// the current transform runs, and writes its S outputs straight from registers.
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int N, int MINB>
__global__ void __launch_bounds__(N / 16, MINB)
shear_cols_fft_persist(const float2* __restrict__ T1, float2* __restrict__ T2, RotParams g,
                       const double* __restrict__ b_coef, const float2* __restrict__ tw, int frame0, int nf) {
    using F = ShearFft<N>;
    constexpr int T = F::T;
    extern __shared__ float2 smem2[];
    __shared__ float2 ph3s[16];
    float2* buf = smem2;                 // F::BUF exchange buffer
    float2* stage = smem2 + F::BUF;      // S+1 input samples of the prefetched column
    const int t = threadIdx.x;
    const int S = g.S;
    const long long total = (long long)nf * N;
    long long item = blockIdx.x;
    auto prefetch = [&](long long it) {
        const int fl = (int)(it / N), c = (int)(it % N);
        const float2* src = T1 + (size_t)fl * (S + 1) * N + c;
        for (int row = t; row <= S; row += T) cp_async8(stage + row, src + (size_t)row * N);
        cp_async_commit();
    };
    if (item < total) prefetch(item);
    for (; item < total; item += gridDim.x) {
        cp_async_wait_all();
        __syncthreads();                 // staged column visible; previous transform done with buf
        float re[16], im[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 v = stage[t + j * T];
            re[j] = v.x; im[j] = v.y;
        }
        const float2 v4 = stage[4 * T];  // row S (only thread 0 uses it)
        __syncthreads();                 // everyone has read the staging line
        if (item + gridDim.x < total) prefetch(item + gridDim.x);
        const int fl = (int)(item / N), c = (int)(item % N);
        int s_int; float s_frac;
        const int col_phys = (c + g.y0) & (N - 1);
        split_shift(b_coef[frame0 + fl] * (double)(col_phys - N / 2), s_int, s_frac);
        F::template run<true, true>(re, im, buf, ph3s, tw, t, 0, s_int, s_frac, v4.x, v4.y);
        float2* dst = T2 + (size_t)fl * S * N + c;
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[(size_t)(t + j * T) * N] = make_float2(re[j], im[j]);
    }
}

// ---- pass 2, slab variant: one CTA owns NC adjacent columns and transforms them one after the other.
// A single-column gather touches one 32-byte DRAM/L2 sector per 8-byte sample; measured (config 2) that
// costs 3.5 ms of the 11 ms of this pass.  Here the (S+1) x NC input slab is loaded with NC*8-byte
// segments (one full sector per row for NC = 4), kept column-major in shared memory (pitch = 4 mod 16
// complex words: conflict-free both for the segment-wise fill and for the per-thread column reads),
// every result overwrites its own input slot, and the S x NC output slab leaves with the same segments.
template <int N, int NC, int MINB>
__global__ void __launch_bounds__(N / 16, MINB)
shear_cols_fft_slab(const float2* __restrict__ T1, float2* __restrict__ T2, RotParams g,
                    const double* __restrict__ b_coef, const float2* __restrict__ tw, int frame0, int pitch) {
    using F = ShearFft<N>;
    constexpr int T = F::T;
    extern __shared__ float2 smem2[];
    __shared__ float2 ph3s[16];
    float2* buf = smem2;
    float2* slab = smem2 + F::BUF;
    const int t = threadIdx.x;
    const int fl = blockIdx.y, f = frame0 + fl;
    const int c0 = blockIdx.x * NC;
    const int S = g.S;
    const float2* src = T1 + (size_t)fl * (S + 1) * N + c0;
    for (int idx = t; idx < (S + 1) * NC; idx += T) {
        const int row = idx / NC, c = idx % NC;
        slab[c * pitch + row] = __ldg(src + (size_t)row * N + c);
    }
    __syncthreads();
    const double bc = b_coef[f];
#pragma unroll 1
    for (int c = 0; c < NC; ++c) {
        float2* col = slab + c * pitch;
        float re[16], im[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 v = col[t + j * T];
            re[j] = v.x; im[j] = v.y;
        }
        const float2 v4 = col[4 * T];      // row S (only thread 0 uses it)
        int s_int; float s_frac;
        const int col_phys = (c0 + c + g.y0) & (N - 1);
        split_shift(bc * (double)(col_phys - N / 2), s_int, s_frac);
        F::template run<true, true>(re, im, buf, ph3s, tw, t, 0, s_int, s_frac, v4.x, v4.y);
#pragma unroll
        for (int j = 0; j < 4; ++j) col[t + j * T] = make_float2(re[j], im[j]);
        __syncthreads();                   // buf (and ph3s) are reused by the next column
    }
    float2* dst = T2 + (size_t)fl * S * N + c0;
    for (int idx = t; idx < S * NC; idx += T) {
        const int row = idx / NC, c = idx % NC;
        dst[(size_t)row * N + c] = slab[c * pitch + row];
    }
}

// ---- pass 3: rows [0, S); real part of columns n' in [0, S) -> out, mask restored
template <int N, int NT, int MINB>
__global__ void __launch_bounds__(NT * N / 16, MINB)
shear_rows_last_fft(const float2* __restrict__ T2, const float* __restrict__ in, float* __restrict__ out,
                    RotParams g, const double* __restrict__ a_coef, const float2* __restrict__ tw,
                    int frame0) {
    using F = ShearFft<N>;
    extern __shared__ float2 smem2[];
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
    F::template run<false, true>(re, im, smem2 + tr * F::BUF, ph3s[tr], tw, t, tr, s_int, s_frac, 0.f, 0.f);
    if (valid) {
        const float* src = in + ((size_t)f * g.S + row) * g.S;
        float* dst = out + ((size_t)f * g.S + row) * g.S;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = t + j * F::T;     // < S = 4T
            dst[x] = is_masked(__ldg(src + x), g) ? g.mask_val : re[j];
        }
    }
}

// =====================================================================================
// Packed REAL-plane kernels (default for the power-of-two path)
// =====================================================================================
// A shear of a real line is real up to its Nyquist bin, so the planes between the passes are stored as
// REAL planes  T1r[(S+1) x N], T2r[S x N]  plus a few per-frame scalars, and every complex transform carries
// two adjacent real lines (ShearFft::run_pair): half the transforms and half the plane traffic of the
// complex formulation above, same result.  With primed (re-indexed) coordinates r', n' throughout:
//   pass 1  T1 = A + i beta_r' (-1)^n'                       -> A (real plane), beta[r'] (S+1 scalars)
//   aux 1   sigma = sum_r' (-1)^r' beta_r';  Cs[r'] = Re sum_n' shear_{s_n'}(beta)[r']  (one transform, run_mult)
//   pass 2  column n': Re T2 = B_n'[r'] - (-1)^(n'+r') sigma sin(pi s_n')/N -> P (real plane);  gamma[n'] = Nyquist
//           scalar of column n'.  Im T2 = gamma_n' (-1)^r' + (-1)^n' C_n'[r'] is never formed: pass 3 only
//           needs its alternating row sums  q_r' = (-1)^r' Gamma + Cs[r'],  Gamma = sum_n' (-1)^n' gamma_n'.
//   aux 2   Gamma, corr[r'] = q_r' sin(pi s_r') / N
//   pass 3  out[r'][n'] = Re shear(P_r')[n'] - (-1)^n' corr[r']
// Per-frame scalars live in `aux` (2N floats per frame): beta [0,S], sigma [2S-1], Cs [2S,3S), corr [3S,4S),
// gamma [N,2N).
struct AuxLayout {
    __host__ __device__ static size_t stride(int N) { return (size_t)2 * N; }
    __host__ __device__ static int beta(int) { return 0; }
    __host__ __device__ static int sigma(int S) { return 2 * S - 1; }
    __host__ __device__ static int cs(int S) { return 2 * S; }
    __host__ __device__ static int corr(int S) { return 3 * S; }
    __host__ __device__ static int gamma(int S) { return 4 * S; }
};

// sin(pi s) for s = s_int + s_frac
__device__ __forceinline__ float sinpi_split(int s_int, float s_frac) {
    const float v = sinpif(s_frac);
    return (s_int & 1) ? -v : v;
}

// ---- pass 1 (packed): rows 2m, 2m+1 of [0, S] per transform
template <int N, int NT, int MINB>
__global__ void __launch_bounds__(NT * N / 16, MINB)
shear_rows_first_pk(const float* __restrict__ in, float* __restrict__ T1, float* __restrict__ aux, RotParams g,
                    const int* __restrict__ krot, const double* __restrict__ a_coef,
                    const float2* __restrict__ tw, int frame0) {
    using F = ShearFft<N>;
    extern __shared__ float2 smem2[];
    __shared__ float2 ph3s[NT][32];
    const int tr = threadIdx.x / F::T, t = threadIdx.x % F::T;
    const int fl = blockIdx.y, f = frame0 + fl;
    const int row = 2 * (blockIdx.x * NT + tr);
    const bool va = row <= g.S, vb_ = row + 1 <= g.S;
    const int i = g.y0 + row;
    const int k = krot[f];
    const float* frame = in + (size_t)f * g.S * g.S;
    float re[16], im[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        re[j] = va ? plane_sample(frame, g, k, i, g.y0 + t + j * F::T) : 0.f;
        im[j] = vb_ ? plane_sample(frame, g, k, i + 1, g.y0 + t + j * F::T) : 0.f;
    }
    const float x4r = (va && t == 0) ? plane_sample(frame, g, k, i, g.y0 + g.S) : 0.f;
    const float x4i = (vb_ && t == 0) ? plane_sample(frame, g, k, i + 1, g.y0 + g.S) : 0.f;
    int sa_int, sb_int; float sa_frac, sb_frac;
    const double ac = a_coef[f];
    split_shift(ac * (double)(i - N / 2), sa_int, sa_frac);
    split_shift(ac * (double)(i + 1 - N / 2), sb_int, sb_frac);
    float nya, nyb;
    float2* buf = smem2 + (size_t)tr * 2 * F::BUF;
    F::template run_pair<true, false>(re, im, buf, buf + F::BUF, ph3s[tr], tw, t, tr, sa_int, sa_frac, sb_int,
                                      sb_frac, x4r, x4i, nya, nyb);
    float* dst = T1 + ((size_t)fl * (g.S + 1) + row) * N;
    if (va) {
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[t + j * F::T] = re[j];
    }
    if (vb_) {
#pragma unroll
        for (int j = 0; j < 16; ++j) dst[N + t + j * F::T] = im[j];
    }
    if (t == 0) {
        float* beta = aux + (size_t)fl * AuxLayout::stride(N) + AuxLayout::beta(g.S);
        if (va) beta[row] = nya;
        if (vb_) beta[row + 1] = nyb;
    }
}

// ---- pass 1 (packed), looped: every CTA walks PP consecutive row pairs and fetches the samples of the NEXT
// pair with 4-byte cp.async (zero-fill outside the frame) into per-thread staging slots while the current
// transform runs.  The one-shot kernel above starts every CTA with two dependent global round trips
// (krot[f] -> gather; strided by S floats for odd rot90 counts): ncu r01k shows long_scoreboard as its top
// stall (2.5 warps per issue) and 56 % issue-active against 78 % for pass 3.  Staging slots are private to
// the thread that filled them, so no barrier is needed for them.
__device__ __forceinline__ void cp_async4_zfill(void* smem_dst, const void* gsrc, bool pred) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    const int sz = pred ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

template <int N, int NT, int MINB, int PP, int P2 = 0>
__global__ void __launch_bounds__(NT * N / 16, MINB)
shear_rows_first_pk_loop(const float* __restrict__ in, float* __restrict__ T1, float* __restrict__ aux,
                         RotParams g, const int* __restrict__ krot, const double* __restrict__ a_coef,
                         const float2* __restrict__ tw, int frame0) {
    using F = typename FftSel<N, P2>::type;
    constexpr int T = F::T;
    extern __shared__ float2 smem2[];
    __shared__ float2 ph3s[NT][32];
    __shared__ float stage[NT][2][5][T];          // [row a/b][j = 0..3 and the lone sample n' = S][t]
    const int tr = threadIdx.x / T, t = threadIdx.x % T;
    const int fl = blockIdx.y, f = frame0 + fl;
    const int k = krot[f];
    const double ac = a_coef[f];
    const float* frame = in + (size_t)f * g.S * g.S;
    float2* buf = smem2 + (size_t)tr * 2 * F::BUF;
    float* beta = aux + (size_t)fl * AuxLayout::stride(N) + AuxLayout::beta(g.S);

    auto prefetch = [&](int it) {
        const int row = 2 * ((blockIdx.x * PP + it) * NT + tr);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const bool v = row + h <= g.S;
            const int i = g.y0 + row + h;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int off = v ? plane_offset(g, k, i, g.y0 + t + j * T) : -1;
                cp_async4_zfill(&stage[tr][h][j][t], frame + (off < 0 ? 0 : off), off >= 0);
            }
            if (t == 0) {
                const int off = v ? plane_offset(g, k, i, g.y0 + g.S) : -1;
                cp_async4_zfill(&stage[tr][h][4][0], frame + (off < 0 ? 0 : off), off >= 0);
            }
        }
        cp_async_commit();
    };

    prefetch(0);
#pragma unroll 1
    for (int it = 0; it < PP; ++it) {
        const int row0 = 2 * (blockIdx.x * PP + it) * NT;          // first row of this iteration (CTA-uniform)
        if (row0 > g.S) break;
        const int row = row0 + 2 * tr;
        const bool va = row <= g.S, vb_ = row + 1 <= g.S;
        const int i = g.y0 + row;
        cp_async_wait_all();
        float re[16], im[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            re[j] = clean_sample(stage[tr][0][j][t], g);
            im[j] = clean_sample(stage[tr][1][j][t], g);
        }
        float x4r = 0.f, x4i = 0.f;
        if (t == 0) {
            x4r = clean_sample(stage[tr][0][4][0], g);
            x4i = clean_sample(stage[tr][1][4][0], g);
        }
        if (it + 1 < PP) prefetch(it + 1);       // own slots only: already consumed above
        int sa_int, sb_int; float sa_frac, sb_frac;
        split_shift(ac * (double)(i - N / 2), sa_int, sa_frac);
        split_shift(ac * (double)(i + 1 - N / 2), sb_int, sb_frac);
        float nya, nyb;
        F::template run_pair<true, false>(re, im, buf, buf + F::BUF, ph3s[tr], tw, t, tr, sa_int, sa_frac,
                                          sb_int, sb_frac, x4r, x4i, nya, nyb);
        float* dst = T1 + ((size_t)fl * (g.S + 1) + row) * N;
        if (va) {
#pragma unroll
            for (int j = 0; j < 16; ++j) dst[t + j * T] = re[j];
        }
        if (vb_) {
#pragma unroll
            for (int j = 0; j < 16; ++j) dst[N + t + j * T] = im[j];
        }
        if (t == 0) {
            if (va) beta[row] = nya;
            if (vb_) beta[row + 1] = nyb;
        }
        __syncthreads();                         // buf / zbuf / ph3s are reused by the next pair
    }
    cp_async_wait_all();
}

// ---- aux 1: sigma and Cs of one frame (one CTA of N/16 threads per frame)
template <int N>
__global__ void __launch_bounds__(N / 16)
shear_aux_beta(float* __restrict__ aux, RotParams g, const double* __restrict__ b_coef,
               const float2* __restrict__ tw, int frame0) {
    using F = ShearFft<N>;
    constexpr int T = F::T;
    extern __shared__ float2 smem2[];
    __shared__ float red[T / 32];
    const int t = threadIdx.x;
    const int fl = blockIdx.x, f = frame0 + fl;
    const int S = g.S;
    float* ax = aux + (size_t)fl * AuxLayout::stride(N);
    const float* beta = ax + AuxLayout::beta(S);
    float re[16], im[16];
    float part = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        re[j] = beta[t + j * T];
        im[j] = 0.f;
        part += re[j];
    }
    if (t & 1) part = -part;               // (-1)^r', r' = t + j*T with T even
    const float x4 = (t == 0) ? beta[S] : 0.f;
    part += x4;                            // r' = S is even
    part = warp_sum(part);
    if ((t & 31) == 0) red[t >> 5] = part;
    __syncthreads();
    if (t == 0) {
        float sg = 0.f;
        for (int w = 0; w < T / 32; ++w) sg += red[w];
        ax[AuxLayout::sigma(S)] = sg;
    }
    F::template run_mult<true, true>(re, im, smem2, tw, t, 0, b_coef[f], x4, 0.f);
    float* cs = ax + AuxLayout::cs(S);
#pragma unroll
    for (int j = 0; j < 4; ++j) cs[t + j * T] = re[j];
}

// ---- pass 2 (packed): one CTA owns NC adjacent columns = NC/2 transforms, run one after the other.
// The (S+1) x NC real slab is loaded with 16-byte vectors (NC*4-byte segments per row), kept column-major
// in shared memory (pitch = 4 mod 32 words: conflict-free fill and column reads); results overwrite their
// inputs and the S x NC output slab leaves the same way.
template <int N, int NC, int MINB, int P2 = 0>
__global__ void __launch_bounds__(N / 16, MINB)
shear_cols_pk(const float* __restrict__ T1, float* __restrict__ T2, float* __restrict__ aux, RotParams g,
              const double* __restrict__ b_coef, const float2* __restrict__ tw, int frame0, int pitch) {
    using F = typename FftSel<N, P2>::type;
    constexpr int T = F::T;
    constexpr int V = NC / 4;               // float4 vectors per row
    extern __shared__ float2 smem2[];
    __shared__ float2 ph3s[32];
    float2* buf = smem2;
    float2* zbuf = smem2 + F::BUF;
    float* slab = reinterpret_cast<float*>(smem2 + 2 * F::BUF);
    const int t = threadIdx.x;
    const int fl = blockIdx.y, f = frame0 + fl;
    const int c0 = blockIdx.x * NC;
    const int S = g.S;
    const float* src = T1 + (size_t)fl * (S + 1) * N + c0;
    for (int idx = t; idx < (S + 1) * V; idx += T) {
        const int row = idx / V, h = idx % V;
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + (size_t)row * N) + h);
        float* d = slab + (4 * h) * pitch + row;
        d[0] = v.x; d[pitch] = v.y; d[2 * pitch] = v.z; d[3 * pitch] = v.w;
    }
    __syncthreads();
    float* ax = aux + (size_t)fl * AuxLayout::stride(N);
    const float sigma = ax[AuxLayout::sigma(S)] * (1.0f / N);
    float* gamma = ax + AuxLayout::gamma(S);
    const double bc = b_coef[f];
    const float sgn_t = (t & 1) ? -1.f : 1.f;   // (-1)^r' for r' = t + j*T
#pragma unroll 1
    for (int c = 0; c < NC; c += 2) {
        float* ca = slab + c * pitch;
        float* cb = ca + pitch;
        float re[16], im[16];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            re[j] = ca[t + j * T];
            im[j] = cb[t + j * T];
        }
        const float x4r = ca[4 * T], x4i = cb[4 * T];      // row S (only thread 0 uses them)
        int sa_int, sb_int; float sa_frac, sb_frac;
        const int pa = (c0 + c + g.y0) & (N - 1), pb = (c0 + c + 1 + g.y0) & (N - 1);
        split_shift(bc * (double)(pa - N / 2), sa_int, sa_frac);
        split_shift(bc * (double)(pb - N / 2), sb_int, sb_frac);
        float nya, nyb;
        F::template run_pair<true, true>(re, im, buf, zbuf, ph3s, tw, t, 0, sa_int, sa_frac, sb_int, sb_frac,
                                         x4r, x4i, nya, nyb);
        // Re T2 = B - (-1)^(n'+r') sigma sin(pi s)/N ; n' = c0 + c is even for line a, odd for line b
        const float da = sgn_t * sigma * sinpi_split(sa_int, sa_frac);
        const float db = sgn_t * sigma * sinpi_split(sb_int, sb_frac);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            ca[t + j * T] = re[j] - da;
            cb[t + j * T] = im[j] + db;
        }
        if (t == 0) { gamma[c0 + c] = nya; gamma[c0 + c + 1] = nyb; }
        __syncthreads();                   // buf / zbuf / ph3s are reused by the next pair
    }
    float* dst = T2 + (size_t)fl * S * N + c0;
    for (int idx = t; idx < S * V; idx += T) {
        const int row = idx / V, h = idx % V;
        const float* d = slab + (4 * h) * pitch + row;
        const float4 v = make_float4(d[0], d[pitch], d[2 * pitch], d[3 * pitch]);
        *(reinterpret_cast<float4*>(dst + (size_t)row * N) + h) = v;
    }
}

// ---- aux 2: Gamma and the per-row Nyquist correction of pass 3 (one CTA per frame)
__global__ void __launch_bounds__(256)
shear_aux_gamma(float* __restrict__ aux, RotParams g, const double* __restrict__ a_coef, int frame0) {
    __shared__ float red[8];
    __shared__ float Gs;
    const int t = threadIdx.x;
    const int fl = blockIdx.x, f = frame0 + fl;
    const int S = g.S, N = g.N;
    float* ax = aux + (size_t)fl * AuxLayout::stride(N);
    const float* gamma = ax + AuxLayout::gamma(S);
    float part = 0.f;
    for (int n = t; n < N; n += 256) part += gamma[n];      // n = t mod 2 for every term
    if (t & 1) part = -part;
    part = warp_sum(part);
    if ((t & 31) == 0) red[t >> 5] = part;
    __syncthreads();
    if (t == 0) {
        float G = 0.f;
        for (int w = 0; w < 8; ++w) G += red[w];
        Gs = G;
    }
    __syncthreads();
    const float G = Gs;
    const float* cs = ax + AuxLayout::cs(S);
    float* corr = ax + AuxLayout::corr(S);
    const double ac = a_coef[f];
    for (int r = t; r < S; r += 256) {
        int s_int; float s_frac;
        split_shift(ac * (double)(g.y0 + r - N / 2), s_int, s_frac);
        const float q = ((r & 1) ? -G : G) + cs[r];
        corr[r] = q * sinpi_split(s_int, s_frac) * (1.0f / (float)N);
    }
}

// ---- pass 3 (packed): rows 2m, 2m+1 of [0, S); columns n' in [0, S) -> out, mask restored
template <int N, int NT, int MINB, int P2 = 0>
__global__ void __launch_bounds__(NT * N / 16, MINB)
shear_rows_last_pk(const float* __restrict__ T2, const float* __restrict__ aux, const float* __restrict__ in,
                   float* __restrict__ out, RotParams g, const double* __restrict__ a_coef,
                   const float2* __restrict__ tw, int frame0) {
    using F = typename FftSel<N, P2>::type;
    extern __shared__ float2 smem2[];
    __shared__ float2 ph3s[NT][32];
    const int tr = threadIdx.x / F::T, t = threadIdx.x % F::T;
    const int fl = blockIdx.y, f = frame0 + fl;
    const int row = 2 * (blockIdx.x * NT + tr);
    const bool valid = row < g.S;           // S is even on this path: both rows or none
    const int i = g.y0 + row;
    float re[16], im[16];
    if (valid) {
        const float* src = T2 + ((size_t)fl * g.S + row) * N;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            re[j] = __ldg(src + t + j * F::T);
            im[j] = __ldg(src + N + t + j * F::T);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) { re[j] = 0.f; im[j] = 0.f; }
    }
    int sa_int, sb_int; float sa_frac, sb_frac;
    const double ac = a_coef[f];
    split_shift(ac * (double)(i - N / 2), sa_int, sa_frac);
    split_shift(ac * (double)(i + 1 - N / 2), sb_int, sb_frac);
    float nya, nyb;
    float2* buf = smem2 + (size_t)tr * 2 * F::BUF;
    F::template run_pair<false, true>(re, im, buf, buf + F::BUF, ph3s[tr], tw, t, tr, sa_int, sa_frac, sb_int,
                                      sb_frac, 0.f, 0.f, nya, nyb);
    if (valid) {
        const float* corr = aux + (size_t)fl * AuxLayout::stride(N) + AuxLayout::corr(g.S);
        const float sgn_t = (t & 1) ? -1.f : 1.f;      // (-1)^n' for n' = t + j*T
        const float ca = sgn_t * corr[row], cb = sgn_t * corr[row + 1];
        const float* src = in + ((size_t)f * g.S + row) * g.S;
        float* dst = out_row_ptr(out, g, f, row);      // rows row, row + 1 share a shard (rows_per is even)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = t + j * F::T;     // < S = 4T
            dst[x] = is_masked(__ldg(src + x), g) ? g.mask_val : re[j] - ca;
            dst[g.S + x] = is_masked(__ldg(src + g.S + x), g) ? g.mask_val : im[j] - cb;
        }
    }
}

// ---- pass 3 (packed), looped: PP consecutive row pairs per CTA; the 32 samples per thread of the NEXT pair are
// fetched with 4-byte cp.async into thread-private staging slots while the current transform runs (same idea
// as shear_rows_first_pk_loop: the one-shot kernel opens every CTA with a DRAM round trip).
template <int N, int NT, int MINB, int PP>
__global__ void __launch_bounds__(NT * N / 16, MINB)
shear_rows_last_pk_loop(const float* __restrict__ T2, const float* __restrict__ aux, const float* __restrict__ in,
                        float* __restrict__ out, RotParams g, const double* __restrict__ a_coef,
                        const float2* __restrict__ tw, int frame0) {
    using F = ShearFft<N>;
    constexpr int T = F::T;
    extern __shared__ float2 smem2[];
    __shared__ float2 ph3s[NT][32];
    const int tr = threadIdx.x / T, t = threadIdx.x % T;
    const int fl = blockIdx.y, f = frame0 + fl;
    float2* buf = smem2 + (size_t)tr * 2 * F::BUF;
    float* stage = reinterpret_cast<float*>(smem2 + (size_t)NT * 2 * F::BUF) + (size_t)tr * 32 * T;   // [h][j][t]
    const double ac = a_coef[f];
    const float* corr = aux + (size_t)fl * AuxLayout::stride(N) + AuxLayout::corr(g.S);
    const float sgn_t = (t & 1) ? -1.f : 1.f;      // (-1)^n' for n' = t + j*T

    auto prefetch = [&](int it) {
        const int row = 2 * ((blockIdx.x * PP + it) * NT + tr);
        const bool v = row < g.S;                   // S is even on this path: both rows or none
        const float* src = T2 + ((size_t)fl * g.S + (v ? row : 0)) * N;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 16; ++j)
                cp_async4_zfill(stage + (h * 16 + j) * T + t, src + (size_t)h * N + t + j * T, v);
        cp_async_commit();
    };

    prefetch(0);
#pragma unroll 1
    for (int it = 0; it < PP; ++it) {
        const int row0 = 2 * (blockIdx.x * PP + it) * NT;
        if (row0 >= g.S) break;                     // CTA-uniform
        const int row = row0 + 2 * tr;
        const bool valid = row < g.S;
        const int i = g.y0 + row;
        cp_async_wait_all();
        float re[16], im[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            re[j] = stage[j * T + t];
            im[j] = stage[(16 + j) * T + t];
        }
        if (it + 1 < PP) prefetch(it + 1);          // own slots only: already consumed above
        int sa_int, sb_int; float sa_frac, sb_frac;
        split_shift(ac * (double)(i - N / 2), sa_int, sa_frac);
        split_shift(ac * (double)(i + 1 - N / 2), sb_int, sb_frac);
        float nya, nyb;
        F::template run_pair<false, true>(re, im, buf, buf + F::BUF, ph3s[tr], tw, t, tr, sa_int, sa_frac, sb_int,
                                          sb_frac, 0.f, 0.f, nya, nyb);
        if (valid) {
            const float ca = sgn_t * corr[row], cb = sgn_t * corr[row + 1];
            const float* src = in + ((size_t)f * g.S + row) * g.S;
            float* dst = out + ((size_t)f * g.S + row) * g.S;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int x = t + j * T;            // < S = 4T
                dst[x] = is_masked(__ldg(src + x), g) ? g.mask_val : re[j] - ca;
                dst[g.S + x] = is_masked(__ldg(src + g.S + x), g) ? g.mask_val : im[j] - cb;
            }
        }
        __syncthreads();                            // buf / zbuf / ph3s are reused by the next pair
    }
    cp_async_wait_all();
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

static int fft_slab() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VIP_B200_FFT_SLAB");
        v = e ? atoi(e) : 1;
    }
    return v;
}

static int fft_persist() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VIP_B200_FFT_PERSIST");
        v = e ? atoi(e) : 0;       // measured: 12.3 ms vs 11.9 ms for the one-shot kernel at config 2
    }
    return v;
}

template <int N, int NT, int MINB>
static int launch_fft_chunk(const float* in, float* out, float2* T1, float2* T2, const RotParams& g,
                            const int* krot, const double* a, const double* b, const float2* tw,
                            int frame0, int nf, cudaStream_t st) {
    using F = ShearFft<N>;
    const size_t smem = (size_t)NT * F::BUF * sizeof(float2);
    static bool configured = false;
    if (!configured) {
        VB_CHECK_CUDA(cudaFuncSetAttribute(shear_rows_first_fft<N, NT, MINB>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VB_CHECK_CUDA(cudaFuncSetAttribute(shear_cols_fft<N, NT, MINB>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VB_CHECK_CUDA(cudaFuncSetAttribute(shear_rows_last_fft<N, NT, MINB>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    const int threads = NT * F::T;
    g_timer.mark(st);
    shear_rows_first_fft<N, NT, MINB><<<dim3(ceil_div(g.S + 1, NT), nf), threads, smem, st>>>(
        in, T1, g, krot, a, tw, frame0);
    VB_CHECK_LAUNCH();
    g_timer.mark(st);
    if (NT == 1 && N >= 2048 && fft_slab() == 8) {
        constexpr int NC = 8;
        static bool cfg4 = false;
        const int pitch = ((g.S + 1 + 15) / 16) * 16 + 2;      // 8 columns x 2 rows per half-warp: pitch = 2 mod 16
        const size_t smem3 = (size_t)(F::BUF + NC * pitch) * sizeof(float2);
        if (!cfg4) {
            VB_CHECK_CUDA(cudaFuncSetAttribute(shear_cols_fft_slab<N, NC, MINB>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            cfg4 = true;
        }
        shear_cols_fft_slab<N, NC, MINB><<<dim3(N / NC, nf), F::T, smem3, st>>>(T1, T2, g, b, tw, frame0, pitch);
    } else if (NT == 1 && N >= 2048 && fft_slab() > 0) {
        constexpr int NC = 4;
        static bool cfg3 = false;
        const int pitch = ((g.S + 1 + 15) / 16) * 16 + 4;
        const size_t smem3 = (size_t)(F::BUF + NC * pitch) * sizeof(float2);
        if (!cfg3) {
            VB_CHECK_CUDA(cudaFuncSetAttribute(shear_cols_fft_slab<N, NC, MINB>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
            cfg3 = true;
        }
        shear_cols_fft_slab<N, NC, MINB><<<dim3(N / NC, nf), F::T, smem3, st>>>(T1, T2, g, b, tw, frame0, pitch);
    } else if (NT == 1 && N >= 2048 && fft_persist()) {
        static bool cfg2 = false;
        const size_t smem2 = (size_t)(F::BUF + g.S + 8) * sizeof(float2);
        if (!cfg2) {
            VB_CHECK_CUDA(cudaFuncSetAttribute(shear_cols_fft_persist<N, MINB>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            cfg2 = true;
        }
        const long long total = (long long)nf * N;
        const int grid = (int)((total < (long long)kNumSMs * MINB) ? total : (long long)kNumSMs * MINB);
        shear_cols_fft_persist<N, MINB><<<grid, F::T, smem2, st>>>(T1, T2, g, b, tw, frame0, nf);
    } else {
        shear_cols_fft<N, NT, MINB><<<dim3(N / NT, nf), threads, smem, st>>>(T1, T2, g, b, tw, frame0);
    }
    VB_CHECK_LAUNCH();
    g_timer.mark(st);
    shear_rows_last_fft<N, NT, MINB><<<dim3(ceil_div(g.S, NT), nf), threads, smem, st>>>(
        T2, in, out, g, a, tw, frame0);
    VB_CHECK_LAUNCH();
    g_timer.mark(st);
    return 0;
}

static int fft_packed() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VIP_B200_FFT_PACKED");
        v = e ? atoi(e) : 1;       // 0: complex-plane kernels (one line per transform), kept for A/B runs
    }
    return v;
}

// passes 1 / 3: 0 = one row pair per CTA; 1 = looped pass-1 kernel with cp.async prefetch of the next row pair
// (default); 2 = looped pass 1 and pass 3
static int fft_rows_loop() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VIP_B200_FFT_ROWS_LOOP");
        v = e ? atoi(e) : 1;
    }
    return v;
}

static size_t packed_bytes_per_frame(int S, int N) {
    return ((size_t)(S + 1) * N + (size_t)S * N + AuxLayout::stride(N)) * sizeof(float);
}

// pass-2 variant: 0 = 8 columns per CTA at the row kernels' occupancy, 1 = 8 columns, one CTA per SM less
// (no register spills), 2 = 16 columns (64-byte segments), one CTA per SM less
static int fft_cols_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VIP_B200_FFT_COLS");
        v = e ? atoi(e) : 2;       // measured at C2 (pass 2): 5.84 / 5.46 / 5.21 ms for 0 / 1 / 2
    }
    return v;
}

// transforms of the packed-plane kernels: 0 = scalar arithmetic (round 1), 1 = packed fp32x2 (FFMA2 / FADD2 /
// FMUL2) on the swizzled buffers, 2 = packed fp32x2 on the padded buffers (VIP_B200_FFT_F32X2; A/B runs)
static int fft_f32x2() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VIP_B200_FFT_F32X2");
        v = e ? atoi(e) : 2;       // measured at C2 (r02e): derotation 8.47 / 8.41 / 7.72 ms for 0 / 1 / 2
        if (v < 0 || v > 2) v = 2;
    }
    return v;
}

template <int N, int NC, int MINB, int MODE = 0>
static int launch_cols_pk(const float* T1, float* T2, float* aux, const RotParams& g, const double* b,
                          const float2* tw, int frame0, int nf, cudaStream_t st) {
    using F = typename FftSel<N, MODE>::type;
    // conflict-free vector fill: a warp covers 32/(NC/4) rows x NC/4 vectors -> pitch = 4 (NC=8) or 2 (NC=16) mod 32
    const int pitch = ((g.S + 1 + 31) / 32) * 32 + (NC == 16 ? 2 : 4);
    const size_t smem_cols = (size_t)2 * F::BUF * sizeof(float2) + (size_t)NC * pitch * sizeof(float);
    static bool configured = false;
    if (!configured) {
        VB_CHECK_CUDA(cudaFuncSetAttribute(shear_cols_pk<N, NC, MINB, MODE>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = true;
    }
    shear_cols_pk<N, NC, MINB, MODE><<<dim3(N / NC, nf), F::T, smem_cols, st>>>(T1, T2, aux, g, b, tw, frame0, pitch);
    VB_CHECK_LAUNCH();
    return 0;
}

// MODE 0: scalar transforms with every round-1 variant switch; MODE 1 / 2: packed fp32x2 transforms (swizzled /
// padded buffers) on the default variants (looped pass 1, 16-column pass 2, one-shot pass 3)
template <int N, int NT, int MINB, int MODE>
static int launch_fft_chunk_pk_mode(const float* in, float* out, float* T1, float* T2, float* aux, const RotParams& g,
                                    const int* krot, const double* a, const double* b, const float2* tw,
                                    int frame0, int nf, cudaStream_t st) {
    using F = typename FftSel<N, MODE>::type;
    using F0 = ShearFft<N>;
    const size_t smem_rows = (size_t)NT * 2 * F::BUF * sizeof(float2);
    const size_t smem_aux = (size_t)F0::BUF * sizeof(float2);
    constexpr int PP = 4;
    static bool configured = false;
    if (!configured) {
        if (MODE == 0) {
            VB_CHECK_CUDA(cudaFuncSetAttribute(shear_rows_first_pk<N, NT, MINB>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
        }
        VB_CHECK_CUDA(cudaFuncSetAttribute(shear_rows_first_pk_loop<N, NT, MINB, PP, MODE>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
        VB_CHECK_CUDA(cudaFuncSetAttribute(shear_rows_last_pk<N, NT, MINB, MODE>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_rows));
        VB_CHECK_CUDA(cudaFuncSetAttribute(shear_aux_beta<N>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem_aux));
        configured = true;
    }
    const int threads = NT * F::T;
    const int pairs1 = (g.S + 2) / 2, pairs3 = g.S / 2;
    g_timer.mark(st);
    if (MODE != 0 || fft_rows_loop()) {
        shear_rows_first_pk_loop<N, NT, MINB, PP, MODE>
            <<<dim3(ceil_div(pairs1, NT * PP), nf), threads, smem_rows, st>>>(in, T1, aux, g, krot, a, tw, frame0);
    } else if (MODE == 0) {
        shear_rows_first_pk<N, NT, MINB><<<dim3(ceil_div(pairs1, NT), nf), threads, smem_rows, st>>>(
            in, T1, aux, g, krot, a, tw, frame0);
    }
    VB_CHECK_LAUNCH();
    g_timer.mark(st);
    shear_aux_beta<N><<<nf, F0::T, smem_aux, st>>>(aux, g, b, tw, frame0);
    VB_CHECK_LAUNCH();
    {
        constexpr int MC = MINB > 1 ? MINB - 1 : 1;
        int rc;
        if (MODE != 0) {
            rc = launch_cols_pk<N, 16, MC, MODE>(T1, T2, aux, g, b, tw, frame0, nf, st);
        } else {
            const int cv = fft_cols_variant();
            rc = cv == 2 ? launch_cols_pk<N, 16, MC>(T1, T2, aux, g, b, tw, frame0, nf, st)
               : cv == 1 ? launch_cols_pk<N, 8, MC>(T1, T2, aux, g, b, tw, frame0, nf, st)
                         : launch_cols_pk<N, 8, MINB>(T1, T2, aux, g, b, tw, frame0, nf, st);
        }
        if (rc) return rc;
    }
    g_timer.mark(st);
    shear_aux_gamma<<<nf, 256, 0, st>>>(aux, g, a, frame0);
    VB_CHECK_LAUNCH();
    if (MODE == 0 && fft_rows_loop() >= 2) {
        const size_t smem_last = smem_rows + (size_t)NT * 32 * F::T * sizeof(float);
        static bool cfg_last = false;
        if (!cfg_last) {
            VB_CHECK_CUDA(cudaFuncSetAttribute(shear_rows_last_pk_loop<N, NT, MINB, PP>,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_last));
            cfg_last = true;
        }
        shear_rows_last_pk_loop<N, NT, MINB, PP><<<dim3(ceil_div(pairs3, NT * PP), nf), threads, smem_last, st>>>(
            T2, aux, in, out, g, a, tw, frame0);
    } else {
        shear_rows_last_pk<N, NT, MINB, MODE><<<dim3(ceil_div(pairs3, NT), nf), threads, smem_rows, st>>>(
            T2, aux, in, out, g, a, tw, frame0);
    }
    VB_CHECK_LAUNCH();
    g_timer.mark(st);
    return 0;
}

template <int N, int NT, int MINB>
static int launch_fft_chunk_pk(const float* in, float* out, float* T1, float* T2, float* aux, const RotParams& g,
                               const int* krot, const double* a, const double* b, const float2* tw,
                               int frame0, int nf, cudaStream_t st) {
    switch (fft_f32x2()) {
        case 0: return launch_fft_chunk_pk_mode<N, NT, MINB, 0>(in, out, T1, T2, aux, g, krot, a, b, tw, frame0, nf, st);
        case 1: return launch_fft_chunk_pk_mode<N, NT, MINB, 1>(in, out, T1, T2, aux, g, krot, a, b, tw, frame0, nf, st);
        default: return launch_fft_chunk_pk_mode<N, NT, MINB, 2>(in, out, T1, T2, aux, g, krot, a, b, tw, frame0, nf, st);
    }
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
        // 1: one transform per 128-thread CTA, 4 CTAs/SM (default: measured fastest at C2); 2: NT=2 x 2 CTAs/SM;
        // 8 (packed only): NT=1 with 3 CTAs/SM (168 registers).  Measured and dropped on the complex kernels:
        // NT=4 (21.9 vs 18.2 ms), 5 or 6 CTAs/SM (spills).
        v = e ? atoi(e) : 1;
    }
    return v;
}

static bool fft_path(int S, int N) {
    // the FFT kernels assume the exact 4x plane of power-of-two frames (N = 4S)
    return (N & (N - 1)) == 0 && N >= 512 && N <= 4096 && N == 4 * S;
}

// complex planes T1[(S+1) x N], T2[S x N] (direct path, unpacked FFT path)
static size_t complex_bytes_per_frame(int S, int N) {
    return ((size_t)(S + 1) * N + (size_t)S * N) * sizeof(float2);
}

size_t derotate_scratch_bytes_per_frame(int S, int N) {
    if (fft_path(S, N) && fft_packed()) return packed_bytes_per_frame(S, N);
    return complex_bytes_per_frame(S, N);
}

// smallest scratch any path accepts for this geometry (force_direct on a power-of-two plane included)
size_t derotate_scratch_bytes_min(int S, int N) { return complex_bytes_per_frame(S, N); }

int derotate_run(const float* in, float* out, int nframes, const RotParams& g, const int* krot,
                 const double* a, const double* b, const float2* tw, void* scratch,
                 size_t scratch_bytes, int force_direct, int* launches, cudaStream_t st) {
    const bool use_fft = fft_path(g.S, g.N) && !force_direct;
    const bool packed = use_fft && fft_packed();
    VB_REQUIRE(g.om.nshards == 0 || (packed && fft_rows_loop() < 2),
               "derotate: scattered output needs the packed FFT path (power-of-two frames, N = 4S)");
    const size_t per_frame = packed ? packed_bytes_per_frame(g.S, g.N) : complex_bytes_per_frame(g.S, g.N);
    VB_REQUIRE(scratch_bytes >= per_frame, "derotate: scratch too small (%zu < %zu)", scratch_bytes,
               per_frame);
    int chunk = (int)(scratch_bytes / per_frame);
    if (chunk > nframes) chunk = nframes;
    if (chunk > 65535) chunk = 65535;
    float2* T1 = reinterpret_cast<float2*>(scratch);
    float2* T2 = T1 + (size_t)chunk * (g.S + 1) * g.N;
    float* T1r = reinterpret_cast<float*>(scratch);
    float* T2r = T1r + (size_t)chunk * (g.S + 1) * g.N;
    float* aux = T2r + (size_t)chunk * g.S * g.N;
    int nl = 0;
    for (int f0 = 0; f0 < nframes; f0 += chunk) {
        const int nf = (nframes - f0 < chunk) ? nframes - f0 : chunk;
        int rc;
        if (packed) {
            switch (g.N) {
                case 512:  rc = launch_fft_chunk_pk<512, 8, 1>(in, out, T1r, T2r, aux, g, krot, a, b, tw, f0, nf, st); break;
                case 1024: rc = launch_fft_chunk_pk<1024, 4, 1>(in, out, T1r, T2r, aux, g, krot, a, b, tw, f0, nf, st); break;
                case 2048:
                    if (fft_nt() == 2) rc = launch_fft_chunk_pk<2048, 2, 2>(in, out, T1r, T2r, aux, g, krot, a, b, tw, f0, nf, st);
                    else rc = launch_fft_chunk_pk<2048, 1, 4>(in, out, T1r, T2r, aux, g, krot, a, b, tw, f0, nf, st);
                    break;
                default:   rc = launch_fft_chunk_pk<4096, 1, 2>(in, out, T1r, T2r, aux, g, krot, a, b, tw, f0, nf, st); break;
            }
            nl += 5;
        } else if (use_fft) {
            switch (g.N) {
                case 512:  rc = launch_fft_chunk<512, 8, 1>(in, out, T1, T2, g, krot, a, b, tw, f0, nf, st); break;
                case 1024: rc = launch_fft_chunk<1024, 4, 1>(in, out, T1, T2, g, krot, a, b, tw, f0, nf, st); break;
                case 2048:
                    if (fft_nt() == 2) rc = launch_fft_chunk<2048, 2, 2>(in, out, T1, T2, g, krot, a, b, tw, f0, nf, st);
                    else rc = launch_fft_chunk<2048, 1, 4>(in, out, T1, T2, g, krot, a, b, tw, f0, nf, st);
                    break;
                default:   rc = launch_fft_chunk<4096, 1, 2>(in, out, T1, T2, g, krot, a, b, tw, f0, nf, st); break;
            }
            nl += 3;
        } else {
            rc = launch_direct_chunk(in, out, T1, T2, g, krot, a, b, f0, nf, st);
            nl += 3;
        }
        if (rc) return rc;
    }
    if (launches) *launches = nl;
    return 0;
}

}  // namespace vb
