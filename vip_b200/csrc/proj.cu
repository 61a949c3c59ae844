// Principal components and projection/subtraction (CUDA-core fp32 path).
//
// Role in the reference (src/vip_hci/psfsub/pca_fullfr.py:1727-1732):
//     V = svd_wrapper(ref_lib, ...)               # (k, p)
//     transformed   = np.dot(V, matrix_emp.T)     # (k, n)
//     reconstructed = np.dot(transformed.T, V)    # (n, p)
//     residuals     = matrix - reconstructed
// Here  V = Wt . M  (Wt = diag(1/sigma) E_k^T from the Gramian eigenpairs, svd.py:451-459) and
//       R = M - C . V  with C = transformed^T (n x k).  Both are skinny GEMMs streamed over the
// pixel axis: one thread owns one pixel column and keeps KC accumulators in registers, the
// small coefficient matrix sits in shared memory and is read by warp-wide broadcasts.
#include "common.cuh"

namespace vb {

constexpr int KCMAX = 32;     // components per pass (kernels are instantiated for 8/16/20/24/32)
constexpr int ROWS = 256;     // rows of the small matrix staged in shared memory per step
constexpr int PT = 256;       // threads (= pixels) per CTA

// V[kk][j] = sum_i Wt[kk][i] * M[i][j]   for kk in [k0, k0+kc).
// Coefficients and accumulation are fp64: Wt = diag(1/sigma) E_k^T has rows that nearly annihilate the
// (huge) stellar halo, and the reference's V is an fp64 LAPACK result rounded once to fp32.  Rounding Wt
// to fp32 first was measured to cost 3e-4 of final-frame parity in RDI/ARDI (tests).
template <int KC>
__global__ void __launch_bounds__(PT)
pcs_kernel(const double* __restrict__ Wt, const float* __restrict__ M, int n, size_t p, int k0, int kc,
           float* __restrict__ V) {
    constexpr int WROWS = 128;
    __shared__ __align__(16) double Ws[WROWS][KC];
    const size_t j = (size_t)blockIdx.x * PT + threadIdx.x;
    const bool jin = j < p;
    double acc[KC];
#pragma unroll
    for (int q = 0; q < KC; ++q) acc[q] = 0.0;
    for (int i0 = 0; i0 < n; i0 += WROWS) {
        const int ni = (n - i0 < WROWS) ? n - i0 : WROWS;
        __syncthreads();
        for (int idx = threadIdx.x; idx < ni * KC; idx += PT) {
            const int i = idx / KC, q = idx % KC;
            Ws[i][q] = (q < kc) ? Wt[(size_t)(k0 + q) * n + i0 + i] : 0.0;
        }
        __syncthreads();
        if (jin) {
            // 8 independent loads in flight per thread (the kernel is latency/MLP-bound otherwise)
            const float* src = M + (size_t)i0 * p + j;
            for (int ib = 0; ib < ni; ib += 8) {
                float mv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) mv[u] = (ib + u < ni) ? __ldg(src + (size_t)(ib + u) * p) : 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (ib + u < ni) {
                        const double m = (double)mv[u];
#pragma unroll
                        for (int q2 = 0; q2 < KC; q2 += 2) {
                            const double2 w = *reinterpret_cast<const double2*>(&Ws[ib + u][q2]);
                            acc[q2 + 0] = fma(w.x, m, acc[q2 + 0]);
                            acc[q2 + 1] = fma(w.y, m, acc[q2 + 1]);
                        }
                    }
                }
            }
        }
    }
    if (jin) {
#pragma unroll
        for (int q = 0; q < KC; ++q)
            if (q < kc) V[(size_t)(k0 + q) * p + j] = (float)acc[q];
    }
}

// R[i][j] = Src[i][j] - sum_{kk in [k0,k0+kc)} C[i][kk] * V[kk][j]
template <int KC>
__global__ void __launch_bounds__(PT)
subtract_kernel(const float* Src, const float* __restrict__ C, int ldc, const float* __restrict__ V,
                int n, size_t p, int k0, int kc, float* R) {  // Src may alias R
    __shared__ __align__(16) float Cs[ROWS][KC];
    const size_t j = (size_t)blockIdx.x * PT + threadIdx.x;
    const bool jin = j < p;
    float v[KC];
#pragma unroll
    for (int q = 0; q < KC; ++q) v[q] = (jin && q < kc) ? V[(size_t)(k0 + q) * p + j] : 0.f;
    for (int i0 = 0; i0 < n; i0 += ROWS) {
        const int ni = (n - i0 < ROWS) ? n - i0 : ROWS;
        __syncthreads();
        for (int idx = threadIdx.x; idx < ni * KC; idx += PT) {
            const int i = idx / KC, q = idx % KC;
            Cs[i][q] = (q < kc) ? C[(size_t)(i0 + i) * ldc + k0 + q] : 0.f;
        }
        __syncthreads();
        if (jin) {
            // 8 independent loads in flight per thread; the dot products run while they land
            const float* src = Src + (size_t)i0 * p + j;
            float* dst = R + (size_t)i0 * p + j;
            for (int ib = 0; ib < ni; ib += 8) {
                float mv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) mv[u] = (ib + u < ni) ? src[(size_t)(ib + u) * p] : 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (ib + u < ni) {
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                        for (int q4 = 0; q4 < KC; q4 += 4) {
                            const float4 c = *reinterpret_cast<const float4*>(&Cs[ib + u][q4]);
                            s0 = fmaf(c.x, v[q4 + 0], s0);
                            s1 = fmaf(c.y, v[q4 + 1], s1);
                            s2 = fmaf(c.z, v[q4 + 2], s2);
                            s3 = fmaf(c.w, v[q4 + 3], s3);
                        }
                        dst[(size_t)(ib + u) * p] = mv[u] - ((s0 + s1) + (s2 + s3));
                    }
                }
            }
        }
    }
}

// out = a - b (elementwise), used for `reconstructed = matrix - residuals` when full_output asks for it
__global__ void sub_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                           size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = a[i] - b[i];
}

// V (k x p, fp32) = Wt (k x n, row-major, fp64) . M (n x p, fp32), fp64 accumulation
int pcs_f32(const double* Wt, const float* M, int k, int n, size_t p, float* V, int* launches, cudaStream_t st) {
    VB_REQUIRE(k > 0 && n > 0 && p > 0, "pcs: empty problem");
    const unsigned grid = (unsigned)ceil_div(p, (size_t)PT);
    int nl = 0;
    for (int k0 = 0; k0 < k; k0 += KCMAX) {
        const int kc = (k - k0 < KCMAX) ? k - k0 : KCMAX;
        if (kc <= 8)       pcs_kernel<8><<<grid, PT, 0, st>>>(Wt, M, n, p, k0, kc, V);
        else if (kc <= 16) pcs_kernel<16><<<grid, PT, 0, st>>>(Wt, M, n, p, k0, kc, V);
        else if (kc <= 20) pcs_kernel<20><<<grid, PT, 0, st>>>(Wt, M, n, p, k0, kc, V);
        else if (kc <= 24) pcs_kernel<24><<<grid, PT, 0, st>>>(Wt, M, n, p, k0, kc, V);
        else               pcs_kernel<32><<<grid, PT, 0, st>>>(Wt, M, n, p, k0, kc, V);
        VB_CHECK_LAUNCH();
        ++nl;
    }
    if (launches) *launches = nl;
    return 0;
}

// R (n x p) = M - C (n x k, row-major, leading dimension ldc) . V (k x p).   R may alias M.
int project_subtract_f32(const float* M, const float* C, int ldc, const float* V, int k, int n, size_t p,
                         float* R, int* launches, cudaStream_t st) {
    VB_REQUIRE(k > 0 && n > 0 && p > 0, "project_subtract: empty problem");
    const unsigned grid = (unsigned)ceil_div(p, (size_t)PT);
    int nl = 0;
    for (int k0 = 0; k0 < k; k0 += KCMAX) {
        const int kc = (k - k0 < KCMAX) ? k - k0 : KCMAX;
        const float* src = (k0 == 0) ? M : R;
        if (kc <= 8)       subtract_kernel<8><<<grid, PT, 0, st>>>(src, C, ldc, V, n, p, k0, kc, R);
        else if (kc <= 16) subtract_kernel<16><<<grid, PT, 0, st>>>(src, C, ldc, V, n, p, k0, kc, R);
        else if (kc <= 20) subtract_kernel<20><<<grid, PT, 0, st>>>(src, C, ldc, V, n, p, k0, kc, R);
        else if (kc <= 24) subtract_kernel<24><<<grid, PT, 0, st>>>(src, C, ldc, V, n, p, k0, kc, R);
        else               subtract_kernel<32><<<grid, PT, 0, st>>>(src, C, ldc, V, n, p, k0, kc, R);
        VB_CHECK_LAUNCH();
        ++nl;
    }
    if (launches) *launches = nl;
    return 0;
}

int sub_f32(const float* a, const float* b, float* out, size_t count, cudaStream_t st) {
    sub_kernel<<<(unsigned)ceil_div(count, (size_t)256), 256, 0, st>>>(a, b, out, count);
    VB_CHECK_LAUNCH();
    return 0;
}

}  // namespace vb
