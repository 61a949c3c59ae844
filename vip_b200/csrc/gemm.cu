// Batched fp32 GEMM (CUDA-core FFMA):  C[b] = alpha * A[b'] * op(B[b'']) + beta * C[b].
//
// Role in the reference: the two-sided separable resampling  Re(L X L^T)  that replaces the 2-D FFT
// zoom of scale_fft (src/vip_hci/preproc/rescaling.py:1114-1217) for every (ADI frame, channel) of
// an IFS cube; operators depend on the channel only (index b % a_mod / b % b_mod), frames on b.
// 128x64 tiles, 16-deep k slabs, 8x4 register micro-tiles, double-buffered shared memory.
#include "common.cuh"

namespace vb {

constexpr int MB = 128, NB = 64, KB = 16;

struct GemmArgs {
    const float* A; long long lda, strideA; int a_mod;
    const float* B; long long ldb, strideB; int b_mod;
    float* C; long long ldc, strideC;
    int M, N, K;
    float alpha, beta;
};

template <bool BT>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(GemmArgs g) {
    __shared__ __align__(16) float As[2][KB][MB + 4];
    __shared__ __align__(16) float Bs[2][KB][NB + 4];
    const int b = blockIdx.z;
    const float* A = g.A + (size_t)(g.a_mod ? b % g.a_mod : b) * g.strideA;
    const float* B = g.B + (size_t)(g.b_mod ? b % g.b_mod : b) * g.strideB;
    float* C = g.C + (size_t)b * g.strideC;
    const int m0 = blockIdx.y * MB, n0 = blockIdx.x * NB;
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;       // micro-tile rows ty*8.., cols tx*4..

    // fp32 FMA chains of one k-slab (16 terms), then fp64 accumulation across slabs: the operands
    // are O(1e4) pixel values times O(1) operator entries, and the reference's FFT zoom is only
    // fp32-accurate itself, so this keeps us at that floor for 6 % extra instructions
    float acc[8][4];
    double acc64[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j] = 0.f; acc64[i][j] = 0.0; }

    // loaders: A tile 128x16 (k fastest in memory): thread -> k = tid&15, rows (tid>>4) + 16*i
    // B tile: not transposed (K x N, n fastest): thread -> n = tid&63, k = (tid>>6) + 4*i
    //         transposed     (N x K, k fastest): thread -> k = tid&15, n = (tid>>4) + 16*i
    float ra[8], rb[4];
    auto gload = [&](int k0) {
        {
            const int k = k0 + (tid & 15);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = m0 + (tid >> 4) + 16 * i;
                ra[i] = (k < g.K && m < g.M) ? __ldg(A + (size_t)m * g.lda + k) : 0.f;
            }
        }
        if (BT) {
            const int k = k0 + (tid & 15);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int n = n0 + (tid >> 4) + 16 * i;
                rb[i] = (k < g.K && n < g.N) ? __ldg(B + (size_t)n * g.ldb + k) : 0.f;
            }
        } else {
            const int n = n0 + (tid & 63);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = k0 + (tid >> 6) + 4 * i;
                rb[i] = (k < g.K && n < g.N) ? __ldg(B + (size_t)k * g.ldb + n) : 0.f;
            }
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[buf][tid & 15][(tid >> 4) + 16 * i] = ra[i];
        if (BT) {
#pragma unroll
            for (int i = 0; i < 4; ++i) Bs[buf][tid & 15][(tid >> 4) + 16 * i] = rb[i];
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) Bs[buf][(tid >> 6) + 4 * i][tid & 63] = rb[i];
        }
    };

    gload(0);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < g.K; k0 += KB) {
        const bool more = k0 + KB < g.K;
        if (more) gload(k0 + KB);
#pragma unroll
        for (int k = 0; k < KB; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[4] = {b0.x, b0.y, b0.z, b0.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc64[i][j] += (double)acc[i][j]; acc[i][j] = 0.f; }
        if (more) {
            sstore(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < g.N) {
                float* c = C + (size_t)m * g.ldc + n;
                const double v = (double)g.alpha * acc64[i][j];
                *c = (g.beta == 0.f) ? (float)v : (float)(v + (double)g.beta * (double)(*c));
            }
        }
    }
}

int gemm_f32(const GemmArgs& g, int trans_b, int batch, cudaStream_t st) {
    VB_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0 && batch > 0, "gemm: empty problem");
    VB_REQUIRE(batch <= 65535, "gemm: batch too large");
    dim3 grid(ceil_div(g.N, NB), ceil_div(g.M, MB), batch);
    if (trans_b) gemm_f32_kernel<true><<<grid, 256, 0, st>>>(g);
    else gemm_f32_kernel<false><<<grid, 256, 0, st>>>(g);
    VB_CHECK_LAUNCH();
    return 0;
}

}  // namespace vb
