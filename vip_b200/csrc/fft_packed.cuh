// Packed-fp32 (f32x2) complex arithmetic for the register FFTs of csrc/derotate.cu.
//
// sm_100a has FFMA2 / FADD2 / FMUL2: one instruction works on an aligned register pair, i.e. on one complex
// number (lo = re, hi = im), with free operand modifiers for "swap halves" (.LO_HI), per-half negation (.NP / .PN),
// scalar broadcast (.F32) and immediates.  A complex add is ONE instruction instead of two, a complex multiply TWO
// instead of four, a multiplication by +-i disappears into the operand modifiers of the add that consumes it.
// The shear kernels are instruction-issue bound (ncu r01l: issue-active 64-73 %, FMA pipe 41-56 %), so halving the
// FP32 instruction count is the lever (profiles/r02_fp32x2.md has the microbenchmark that motivated this).
//
// Everything is written on float2 values; ptxas keeps them in aligned pairs and picks the modifiers.
#pragma once
#include <cuda_runtime.h>

namespace vb {
namespace pk {

typedef unsigned long long u64;

__device__ __forceinline__ u64 as_u64(float2 a) { return *reinterpret_cast<u64*>(&a); }
__device__ __forceinline__ float2 as_f2(u64 a) { return *reinterpret_cast<float2*>(&a); }

__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(as_u64(a)), "l"(as_u64(b)));
    return as_f2(d);
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(as_u64(a)), "l"(as_u64(b)));
    return as_f2(d);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    u64 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(as_u64(a)), "l"(as_u64(b)));
    return as_f2(d);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(as_u64(a)), "l"(as_u64(b)), "l"(as_u64(c)));
    return as_f2(d);
}

__device__ __forceinline__ float2 swp(float2 a) { return make_float2(a.y, a.x); }
__device__ __forceinline__ float2 bc(float s) { return make_float2(s, s); }

// a + SIGN * i * b  =  (a.x - SIGN b.y, a.y + SIGN b.x)
template <int SIGN>
__device__ __forceinline__ float2 add_i(float2 a, float2 b) {
    return fma2(swp(b), make_float2(-(float)SIGN, (float)SIGN), a);
}
// a - SIGN * i * b
template <int SIGN>
__device__ __forceinline__ float2 sub_i(float2 a, float2 b) {
    return fma2(swp(b), make_float2((float)SIGN, -(float)SIGN), a);
}
// SIGN * i * a
template <int SIGN>
__device__ __forceinline__ float2 mul_i(float2 a) { return make_float2(-(float)SIGN * a.y, (float)SIGN * a.x); }

// x * w for w = (wr, wi):  x * wr + swap(x) * (-wi, wi)
__device__ __forceinline__ float2 cmul2(float2 x, float2 w) {
    return fma2(x, bc(w.x), mul2(swp(x), make_float2(-w.y, w.y)));
}
// x * conj(w)
__device__ __forceinline__ float2 cmulc2(float2 x, float2 w) {
    return fma2(x, bc(w.x), mul2(swp(x), make_float2(w.y, -w.y)));
}

// x * w16^(SIGN * m),  w16 = exp(+2 pi i / 16), m in [0, 8)
template <int SIGN>
__device__ __forceinline__ float2 mul_w16(int m, float2 x) {
    constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, RH = 0.70710678118654752f;
    constexpr float sg = (float)SIGN;
    switch (m) {
        case 0: return x;
        case 4: return mul_i<SIGN>(x);
        case 2: return fma2(swp(x), make_float2(-sg * RH, sg * RH), mul2(x, bc(RH)));
        case 6: return fma2(swp(x), make_float2(-sg * RH, sg * RH), mul2(x, bc(-RH)));
        case 1: return fma2(swp(x), make_float2(-sg * S1, sg * S1), mul2(x, bc(C1)));
        case 3: return fma2(swp(x), make_float2(-sg * C1, sg * C1), mul2(x, bc(S1)));
        case 5: return fma2(swp(x), make_float2(-sg * C1, sg * C1), mul2(x, bc(-S1)));
        default: return fma2(swp(x), make_float2(-sg * S1, sg * S1), mul2(x, bc(-C1)));
    }
}

// (a, b) -> (a + b, (a - b) * w16^(SIGN m)) with the +-i / sqrt(1/2) cases folded into the fewest packed ops
template <int SIGN>
__device__ __forceinline__ void dif_bfly(int m, float2& a, float2& b) {
    const float2 s = add2(a, b);
    const float2 d = sub2(a, b);
    a = s;
    b = mul_w16<SIGN>(m, d);
}
// (a, b) -> (a + b w, a - b w),  w = w16^(SIGN m)
template <int SIGN>
__device__ __forceinline__ void dit_bfly(int m, float2& a, float2& b) {
    if (m == 0) {
        const float2 s = add2(a, b), d = sub2(a, b);
        a = s; b = d;
    } else if (m == 4) {
        const float2 s = add_i<SIGN>(a, b), d = sub_i<SIGN>(a, b);
        a = s; b = d;
    } else {
        const float2 t = mul_w16<SIGN>(m, b);
        const float2 s = add2(a, t), d = sub2(a, t);
        a = s; b = d;
    }
}

// In-register radix-2 DIF DFT of R points at z[OFF .. OFF+R): natural order in, bit-reversed order out.
template <int R, int SIGN, int OFF, int LEN>
struct DifStage {
    __device__ __forceinline__ static void run(float2 (&z)[16]) {
        constexpr int half = LEN / 2;
#pragma unroll
        for (int s = 0; s < R; s += LEN) {
#pragma unroll
            for (int j = 0; j < half; ++j) dif_bfly<SIGN>(j * (16 / LEN), z[OFF + s + j], z[OFF + s + j + half]);
        }
        DifStage<R, SIGN, OFF, LEN / 2>::run(z);
    }
};
template <int R, int SIGN, int OFF>
struct DifStage<R, SIGN, OFF, 1> {
    __device__ __forceinline__ static void run(float2 (&)[16]) {}
};
template <int R, int SIGN, int OFF>
__device__ __forceinline__ void dif(float2 (&z)[16]) { DifStage<R, SIGN, OFF, R>::run(z); }

// DIT counterpart: bit-reversed order in, natural order out.
template <int R, int SIGN, int OFF, int LEN>
struct DitStage {
    __device__ __forceinline__ static void run(float2 (&z)[16]) {
        constexpr int half = LEN / 2;
#pragma unroll
        for (int s = 0; s < R; s += LEN) {
#pragma unroll
            for (int j = 0; j < half; ++j) dit_bfly<SIGN>(j * (16 / LEN), z[OFF + s + j], z[OFF + s + j + half]);
        }
        DitStage<R, SIGN, OFF, LEN * 2>::run(z);
    }
};
template <int R, int SIGN, int OFF>
struct DitStage<R, SIGN, OFF, 2 * R> {
    __device__ __forceinline__ static void run(float2 (&)[16]) {}
};
template <int R, int SIGN, int OFF>
__device__ __forceinline__ void dit(float2 (&z)[16]) { DitStage<R, SIGN, OFF, 2>::run(z); }

template <int R, int SIGN, int G>
struct GroupFft {
    __device__ __forceinline__ static void fwd(float2 (&z)[16]) {
        dif<R, SIGN, (G - 1) * R>(z);
        GroupFft<R, SIGN, G - 1>::fwd(z);
    }
    __device__ __forceinline__ static void inv(float2 (&z)[16]) {
        dit<R, SIGN, (G - 1) * R>(z);
        GroupFft<R, SIGN, G - 1>::inv(z);
    }
};
template <int R, int SIGN>
struct GroupFft<R, SIGN, 0> {
    __device__ __forceinline__ static void fwd(float2 (&)[16]) {}
    __device__ __forceinline__ static void inv(float2 (&)[16]) {}
};

// radix-4 DFT of (x0..x3): y[a] = sum_j x_j exp(SIGN * 2 pi i j a / 4), natural order in and out
template <int SIGN>
__device__ __forceinline__ void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
    const float2 s0 = add2(x0, x2), d0 = sub2(x0, x2), s1 = add2(x1, x3), d1 = sub2(x1, x3);
    x0 = add2(s0, s1);
    x2 = sub2(s0, s1);
    x1 = add_i<SIGN>(d0, d1);
    x3 = sub_i<SIGN>(d0, d1);
}

}  // namespace pk
}  // namespace vb
