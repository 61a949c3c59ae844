// Microbenchmark (run on the GPU box): packed fp32 (FFMA2 / FADD2 / FMUL2) against scalar FFMA / FADD on sm_100a,
//   (1) raw issue rates with register operands,
//   (2) the register part of the shear FFTs: radix-16 DIF + 16 inter-stage twiddle multiplies per thread,
//       scalar (re[16], im[16]; the formulation of csrc/derotate.cu in round 1) against packed (float2 z[16];
//       csrc/fft_packed.cuh), same arithmetic.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I vip_b200/csrc -o fp32x2_rate tools/microbench/fp32x2_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "fft_packed.cuh"

using namespace vb::pk;

template <int OP>
__global__ void __launch_bounds__(256) rate_kernel(float2* out, float2 b, float2 c, int iters) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }      // 2 FFMA
                else if (OP == 1) a[i] = fma2(a[i], b, c);                                               // 1 FFMA2
                else if (OP == 2) { a[i].x += b.x; a[i].y += b.y; }                                       // 2 FADD
                else if (OP == 3) a[i] = add2(a[i], b);                                                   // 1 FADD2
                else if (OP == 4) a[i] = mul2(a[i], b);                                                   // 1 FMUL2
                else if (OP == 5) a[i] = fma2(swp(a[i]), make_float2(-b.x, b.x), a[(i + 1) & 7]);         // swap + NP
                else a[i] = fma2(a[i], a[(i + 1) & 7], a[(i + 2) & 7]);                                   // 3 varying pairs
            }
        }
    }
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- scalar radix-16 DIF (round-1 formulation)
template <int SIGN>
__device__ __forceinline__ void s_mul_w16(int m, float xr, float xi, float& yr, float& yi) {
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
template <int LEN>
__device__ __forceinline__ void s_dif(float (&re)[16], float (&im)[16]) {
    constexpr int half = LEN / 2;
#pragma unroll
    for (int s = 0; s < 16; s += LEN) {
#pragma unroll
        for (int j = 0; j < half; ++j) {
            const int i0 = s + j, i1 = i0 + half;
            const float ar = re[i0], ai = im[i0], br = re[i1], bi = im[i1];
            re[i0] = ar + br; im[i0] = ai + bi;
            s_mul_w16<-1>(j * (16 / LEN), ar - br, ai - bi, re[i1], im[i1]);
        }
    }
    if constexpr (LEN > 2) s_dif<LEN / 2>(re, im);
}

template <int PACKED>
__global__ void __launch_bounds__(128, 3) fft16_kernel(float2* out, const float2* __restrict__ tw, int iters) {
    const int t = threadIdx.x;
    float2 w[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) w[i] = tw[(t * (i + 1)) & 2047];
    if (PACKED) {
        float2 z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = make_float2(t * 0.001f + i, t * 0.002f - i);
        for (int it = 0; it < iters; ++it) {
            dif<16, -1, 0>(z);
#pragma unroll
            for (int k = 1; k < 16; ++k) {
                const int a = k >> 2, b = k & 3;
                if (b) z[k] = cmul2(z[k], w[b - 1]);
                if (a) z[k] = cmul2(z[k], w[2 + a]);
            }
        }
        float2 s = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 16; ++i) s = add2(s, z[i]);
        out[blockIdx.x * blockDim.x + t] = s;
    } else {
        float re[16], im[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { re[i] = t * 0.001f + i; im[i] = t * 0.002f - i; }
        for (int it = 0; it < iters; ++it) {
            s_dif<16>(re, im);
#pragma unroll
            for (int k = 1; k < 16; ++k) {
                const int a = k >> 2, b = k & 3;
                if (b) { const float xr = re[k], xi = im[k]; re[k] = xr * w[b - 1].x - xi * w[b - 1].y; im[k] = xr * w[b - 1].y + xi * w[b - 1].x; }
                if (a) { const float xr = re[k], xi = im[k]; re[k] = xr * w[2 + a].x - xi * w[2 + a].y; im[k] = xr * w[2 + a].y + xi * w[2 + a].x; }
            }
        }
        float sr = 0.f, si = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) { sr += re[i]; si += im[i]; }
        out[blockIdx.x * blockDim.x + t] = make_float2(sr, si);
    }
}

template <typename F>
static float timed(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(16);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch(4096);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

template <int OP>
static void run_rate(const char* name, float2* out, int flop_per_instr_lane, int instr_per_elem) {
    const int grid = 148 * 8, block = 256;
    const float ms = timed([&](int iters) {
        rate_kernel<OP><<<grid, block>>>(out, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f), iters);
    });
    const double elems = (double)grid * block * 4096.0 * 32.0;          // complex elements updated
    const double winstr = elems * instr_per_elem / 32.0;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-34s %8.3f ms  %6.3f warp-instr/clk/SMSP  %7.2f TFLOP/s\n", name, ms,
           winstr / (ms * 1e-3) / (148.0 * 4.0) / (clk * 1e3),
           elems * 2.0 * flop_per_instr_lane / (ms * 1e-3) / 1e12);
}

int main() {
    float2* out;
    cudaMalloc(&out, 148 * 8 * 256 * sizeof(float2));
    float2* tw;
    cudaMalloc(&tw, 2048 * sizeof(float2));
    cudaMemset(tw, 0, 2048 * sizeof(float2));
    run_rate<0>("2 x FFMA (scalar pair)", out, 2, 2);
    run_rate<1>("FFMA2", out, 2, 1);
    run_rate<6>("FFMA2, 3 varying pairs", out, 2, 1);
    run_rate<5>("FFMA2 swap + per-half negate", out, 2, 1);
    run_rate<2>("2 x FADD (scalar pair)", out, 1, 2);
    run_rate<3>("FADD2", out, 1, 1);
    run_rate<4>("FMUL2", out, 1, 1);
    const int grid = 148 * 3, block = 128;
    const float ms_s = timed([&](int iters) { fft16_kernel<0><<<grid, block>>>(out, tw, iters); });
    const float ms_p = timed([&](int iters) { fft16_kernel<1><<<grid, block>>>(out, tw, iters); });
    printf("radix-16 DIF + 15 twiddle multiplies per thread, 4096 iterations: scalar %.3f ms, packed %.3f ms (x%.2f)\n",
           ms_s, ms_p, ms_s / ms_p);
    return 0;
}
