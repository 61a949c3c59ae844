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
#include <cstdlib>

namespace vb {

constexpr int KCMAX = 32;     // components per pass (kernels are instantiated for 8/16/20/24/32)
constexpr int ROWS = 256;     // rows of the small matrix staged in shared memory per step
constexpr int PT = 256;       // threads (= pixels) per CTA

// V[kk][j] = sum_i Wt[kk][i] * M[i][j]   for kk in [k0, k0+kc).
// Coefficients and accumulation are fp64: Wt = diag(1/sigma) E_k^T has rows that nearly annihilate the
// (huge) stellar halo, and the reference's V is an fp64 LAPACK result rounded once to fp32.  Rounding Wt
// to fp32 first was measured to cost 3e-4 of final-frame parity in RDI/ARDI (tests).
// PX = pixels per thread.  With one pixel per thread every DFMA pair needs its own 16-byte broadcast read of the
// coefficients (16 LDS.128 per 32 DFMA at KC = 32): the shared-memory pipe is as loaded as the fp64 pipe and the
// kernel sat at 34 % of the DFMA rate (ncu launch list of the config-5 slice).  Two adjacent pixels per thread
// (one float2 load per frame) share each coefficient read; the per-pixel summation order is unchanged, so the
// result is bit-identical to PX = 1.
template <int KC, int PX>
__global__ void __launch_bounds__(PT / PX)
pcs_kernel(const double* __restrict__ Wt, const float* __restrict__ M, int n, size_t p, int k0, int kc,
           float* __restrict__ V, float* __restrict__ Vlo) {
    constexpr int WROWS = 128;
    constexpr int NT = PT / PX;
    __shared__ __align__(16) double Ws[WROWS][KC];
    const size_t j = ((size_t)blockIdx.x * NT + threadIdx.x) * PX;
    const bool jin = j < p;                      // PX = 2 is only launched for even p: j + 1 < p as well
    double acc[PX][KC];
#pragma unroll
    for (int x = 0; x < PX; ++x)
#pragma unroll
        for (int q = 0; q < KC; ++q) acc[x][q] = 0.0;
    for (int i0 = 0; i0 < n; i0 += WROWS) {
        const int ni = (n - i0 < WROWS) ? n - i0 : WROWS;
        __syncthreads();
        for (int idx = threadIdx.x; idx < ni * KC; idx += NT) {
            const int i = idx / KC, q = idx % KC;
            Ws[i][q] = (q < kc) ? Wt[(size_t)(k0 + q) * n + i0 + i] : 0.0;
        }
        __syncthreads();
        if (jin) {
            // 8 independent loads in flight per thread (the kernel is latency/MLP-bound otherwise)
            const float* src = M + (size_t)i0 * p + j;
            for (int ib = 0; ib < ni; ib += 8) {
                float mv[8][PX];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (PX == 2) {
                        const float2 t = (ib + u < ni)
                            ? __ldg(reinterpret_cast<const float2*>(src + (size_t)(ib + u) * p)) : make_float2(0.f, 0.f);
                        mv[u][0] = t.x;
                        mv[u][PX - 1] = t.y;
                    } else {
                        mv[u][0] = (ib + u < ni) ? __ldg(src + (size_t)(ib + u) * p) : 0.f;
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (ib + u < ni) {
                        double m[PX];
#pragma unroll
                        for (int x = 0; x < PX; ++x) m[x] = (double)mv[u][x];
#pragma unroll
                        for (int q2 = 0; q2 < KC; q2 += 2) {
                            const double2 w = *reinterpret_cast<const double2*>(&Ws[ib + u][q2]);
#pragma unroll
                            for (int x = 0; x < PX; ++x) {
                                acc[x][q2 + 0] = fma(w.x, m[x], acc[x][q2 + 0]);
                                acc[x][q2 + 1] = fma(w.y, m[x], acc[x][q2 + 1]);
                            }
                        }
                    }
                }
            }
        }
    }
    if (jin) {
#pragma unroll
        for (int q = 0; q < KC; ++q) {
            if (q < kc) {
                if (PX == 2)
                    *reinterpret_cast<float2*>(V + (size_t)(k0 + q) * p + j) =
                        make_float2((float)acc[0][q], (float)acc[PX - 1][q]);
                else
                    V[(size_t)(k0 + q) * p + j] = (float)acc[0][q];
                // optional low part: V + Vlo carries the fp64 sum to ~48 bits (error-free split), for sketches whose
                // rows span a dynamic range that fp32 cannot hold (randomized SVD before orthonormalisation)
                if (Vlo != nullptr) {
                    if (PX == 2)
                        *reinterpret_cast<float2*>(Vlo + (size_t)(k0 + q) * p + j) =
                            make_float2((float)(acc[0][q] - (double)(float)acc[0][q]),
                                        (float)(acc[PX - 1][q] - (double)(float)acc[PX - 1][q]));
                    else
                        Vlo[(size_t)(k0 + q) * p + j] = (float)(acc[0][q] - (double)(float)acc[0][q]);
                }
            }
        }
    }
}

// R[i][j] = Src[i][j] - sum_{kk in [k0,k0+kc)} C[i][kk] * V[kk][j]
// PX = 2: two adjacent pixels per thread (float2 loads and stores): twice the bytes in flight per thread for this
// HBM-bound pass (2.8 TB/s with one pixel per thread), same arithmetic per pixel (bit-identical).
template <int KC, int PX>
__global__ void __launch_bounds__(PT / PX)
subtract_kernel(const float* Src, const float* __restrict__ C, int ldc, const float* __restrict__ V,
                int n, size_t p, int k0, int kc, float* R) {  // Src may alias R
    constexpr int NT = PT / PX;
    __shared__ __align__(16) float Cs[ROWS][KC];
    const size_t j = ((size_t)blockIdx.x * NT + threadIdx.x) * PX;
    const bool jin = j < p;
    float v[PX][KC];
#pragma unroll
    for (int q = 0; q < KC; ++q) {
        if (PX == 2) {
            const float2 t = (jin && q < kc) ? *reinterpret_cast<const float2*>(V + (size_t)(k0 + q) * p + j)
                                             : make_float2(0.f, 0.f);
            v[0][q] = t.x;
            v[PX - 1][q] = t.y;
        } else {
            v[0][q] = (jin && q < kc) ? V[(size_t)(k0 + q) * p + j] : 0.f;
        }
    }
    for (int i0 = 0; i0 < n; i0 += ROWS) {
        const int ni = (n - i0 < ROWS) ? n - i0 : ROWS;
        __syncthreads();
        for (int idx = threadIdx.x; idx < ni * KC; idx += NT) {
            const int i = idx / KC, q = idx % KC;
            Cs[i][q] = (q < kc) ? C[(size_t)(i0 + i) * ldc + k0 + q] : 0.f;
        }
        __syncthreads();
        if (jin) {
            // 8 independent loads in flight per thread; the dot products run while they land
            const float* src = Src + (size_t)i0 * p + j;
            float* dst = R + (size_t)i0 * p + j;
            for (int ib = 0; ib < ni; ib += 8) {
                float mv[8][PX];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (PX == 2) {
                        const float2 t = (ib + u < ni) ? *reinterpret_cast<const float2*>(src + (size_t)(ib + u) * p)
                                                       : make_float2(0.f, 0.f);
                        mv[u][0] = t.x;
                        mv[u][PX - 1] = t.y;
                    } else {
                        mv[u][0] = (ib + u < ni) ? src[(size_t)(ib + u) * p] : 0.f;
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (ib + u < ni) {
                        float s[PX][4];
#pragma unroll
                        for (int x = 0; x < PX; ++x) s[x][0] = s[x][1] = s[x][2] = s[x][3] = 0.f;
#pragma unroll
                        for (int q4 = 0; q4 < KC; q4 += 4) {
                            const float4 c = *reinterpret_cast<const float4*>(&Cs[ib + u][q4]);
#pragma unroll
                            for (int x = 0; x < PX; ++x) {
                                s[x][0] = fmaf(c.x, v[x][q4 + 0], s[x][0]);
                                s[x][1] = fmaf(c.y, v[x][q4 + 1], s[x][1]);
                                s[x][2] = fmaf(c.z, v[x][q4 + 2], s[x][2]);
                                s[x][3] = fmaf(c.w, v[x][q4 + 3], s[x][3]);
                            }
                        }
                        const float r0 = mv[u][0] - ((s[0][0] + s[0][1]) + (s[0][2] + s[0][3]));
                        if (PX == 2) {
                            const float r1 = mv[u][PX - 1] -
                                             ((s[PX - 1][0] + s[PX - 1][1]) + (s[PX - 1][2] + s[PX - 1][3]));
                            *reinterpret_cast<float2*>(dst + (size_t)(ib + u) * p) = make_float2(r0, r1);
                        } else {
                            dst[(size_t)(ib + u) * p] = r0;
                        }
                    }
                }
            }
        }
    }
}

// High-precision variant:  R[i][j] = fl32( Src[i][j] - sum_k C[i][k] * (Vhi[k][j] + Vlo[k][j]) )  with fp64 coefficients,
// the principal components as an error-free fp32 pair and fp64 accumulation.  Why (measured at BASELINE config 2,
// tests/test_gpu_configs.py): the cube holds the stellar halo (1e4) while the residuals are O(10); rounding V and C
// to fp32 leaves an error of ~6e-8 * 1e4 in every residual pixel that is THE SAME in every frame (V does not depend
// on the frame, C[:,0] hardly does), so it survives the temporal median: 1.0e-3 of the final frame's peak against the
// reference run in float64, where the fp32 kernel above is 8.7e-5 on a single residual frame.  The pass stays
// HBM-bound: KC DFMA per pixel and frame (2.6 GFLOP at config 2) hide under the 1 GB stream.
template <int KC, int PX>
__global__ void __launch_bounds__(PT / PX)
subtract_hp_kernel(const float* Src, const double* __restrict__ C, int ldc, const float* __restrict__ Vhi,
                   const float* __restrict__ Vlo, int n, size_t p, int k0, int kc, float* R,
                   float* const* __restrict__ Rrows) {  // Src may alias R; Rrows: see project_subtract_hp_f32
    constexpr int NT = PT / PX;
    constexpr int HROWS = 128;
    constexpr int U = 8;                       // rows per batch; the NEXT batch is loaded while this one is computed
    __shared__ __align__(16) double Cs[HROWS][KC];
    const size_t j = ((size_t)blockIdx.x * NT + threadIdx.x) * PX;
    const bool jin = j < p;
    double v[PX][KC];
#pragma unroll
    for (int q = 0; q < KC; ++q) {
#pragma unroll
        for (int x = 0; x < PX; ++x) {
            double t = 0.0;
            if (jin && q < kc) {
                t = (double)Vhi[(size_t)(k0 + q) * p + j + x];
                if (Vlo != nullptr) t += (double)Vlo[(size_t)(k0 + q) * p + j + x];
            }
            v[x][q] = t;
        }
    }
    auto load_batch = [&](const float* src, int ib, int ni, float (&mv)[U][PX]) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (PX == 2) {
                const float2 t = (ib + u < ni) ? *reinterpret_cast<const float2*>(src + (size_t)(ib + u) * p)
                                               : make_float2(0.f, 0.f);
                mv[u][0] = t.x;
                mv[u][PX - 1] = t.y;
            } else {
                mv[u][0] = (ib + u < ni) ? src[(size_t)(ib + u) * p] : 0.f;
            }
        }
    };
    for (int i0 = 0; i0 < n; i0 += HROWS) {
        const int ni = (n - i0 < HROWS) ? n - i0 : HROWS;
        __syncthreads();
        for (int idx = threadIdx.x; idx < ni * KC; idx += NT) {
            const int i = idx / KC, q = idx % KC;
            Cs[i][q] = (q < kc) ? C[(size_t)(i0 + i) * ldc + k0 + q] : 0.0;
        }
        __syncthreads();
        if (jin) {
            const float* src = Src + (size_t)i0 * p + j;
            float* dst = R + (size_t)i0 * p + j;
            float mv[U][PX], nx[U][PX];
            load_batch(src, 0, ni, mv);
            for (int ib = 0; ib < ni; ib += U) {
                if (ib + U < ni) load_batch(src, ib + U, ni, nx);       // in flight during the DFMA chains below
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (ib + u < ni) {
                        // row i of the result goes to R + i p, or -- scattered output -- to the row pointer the caller
                        // supplied (possibly PEER memory over NVLink: the exchange to frame shards rides on this pass)
                        float* drow = (Rrows != nullptr) ? Rrows[i0 + ib + u] + j : dst + (size_t)(ib + u) * p;
                        double s[PX][2];
#pragma unroll
                        for (int x = 0; x < PX; ++x) s[x][0] = s[x][1] = 0.0;
#pragma unroll
                        for (int q2 = 0; q2 < KC; q2 += 2) {
                            const double2 c = *reinterpret_cast<const double2*>(&Cs[ib + u][q2]);
#pragma unroll
                            for (int x = 0; x < PX; ++x) {
                                s[x][0] = fma(c.x, v[x][q2 + 0], s[x][0]);
                                s[x][1] = fma(c.y, v[x][q2 + 1], s[x][1]);
                            }
                        }
                        const float r0 = (float)((double)mv[u][0] - (s[0][0] + s[0][1]));
                        if (PX == 2) {
                            const float r1 = (float)((double)mv[u][PX - 1] - (s[PX - 1][0] + s[PX - 1][1]));
                            *reinterpret_cast<float2*>(drow) = make_float2(r0, r1);
                        } else {
                            *drow = r0;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int x = 0; x < PX; ++x) mv[u][x] = nx[u][x];
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
int pcs_f32(const double* Wt, const float* M, int k, int n, size_t p, float* V, float* Vlo, int* launches,
            cudaStream_t st) {
    VB_REQUIRE(k > 0 && n > 0 && p > 0, "pcs: empty problem");
    static const int px_env = [] { const char* e = getenv("VIP_B200_PCS_PX"); return e ? atoi(e) : 2; }();
    const bool two = px_env == 2 && (p % 2 == 0) && (reinterpret_cast<uintptr_t>(M) % 8 == 0) &&
                     (reinterpret_cast<uintptr_t>(V) % 8 == 0) && (reinterpret_cast<uintptr_t>(Vlo) % 8 == 0);
    const unsigned grid = (unsigned)ceil_div(p, (size_t)PT);      // PT pixels per CTA for both variants
    int nl = 0;
    for (int k0 = 0; k0 < k; k0 += KCMAX) {
        const int kc = (k - k0 < KCMAX) ? k - k0 : KCMAX;
#define VB_PCS_LAUNCH(KCV)                                                                           \
        do {                                                                                         \
            if (two) pcs_kernel<KCV, 2><<<grid, PT / 2, 0, st>>>(Wt, M, n, p, k0, kc, V, Vlo);           \
            else     pcs_kernel<KCV, 1><<<grid, PT, 0, st>>>(Wt, M, n, p, k0, kc, V, Vlo);           \
        } while (0)
        if (kc <= 8)       VB_PCS_LAUNCH(8);
        else if (kc <= 16) VB_PCS_LAUNCH(16);
        else if (kc <= 20) VB_PCS_LAUNCH(20);
        else if (kc <= 24) VB_PCS_LAUNCH(24);
        else               VB_PCS_LAUNCH(32);
#undef VB_PCS_LAUNCH
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
    static const int px_env = [] { const char* e = getenv("VIP_B200_SUB_PX"); return e ? atoi(e) : 2; }();
    const bool two = px_env == 2 && (p % 2 == 0) && (reinterpret_cast<uintptr_t>(M) % 8 == 0) &&
                     (reinterpret_cast<uintptr_t>(V) % 8 == 0) && (reinterpret_cast<uintptr_t>(R) % 8 == 0);
    const unsigned grid = (unsigned)ceil_div(p, (size_t)PT);      // PT pixels per CTA for both variants
    int nl = 0;
    for (int k0 = 0; k0 < k; k0 += KCMAX) {
        const int kc = (k - k0 < KCMAX) ? k - k0 : KCMAX;
        const float* src = (k0 == 0) ? M : R;
#define VB_SUB_LAUNCH(KCV)                                                                                   \
        do {                                                                                                 \
            if (two) subtract_kernel<KCV, 2><<<grid, PT / 2, 0, st>>>(src, C, ldc, V, n, p, k0, kc, R);      \
            else     subtract_kernel<KCV, 1><<<grid, PT, 0, st>>>(src, C, ldc, V, n, p, k0, kc, R);          \
        } while (0)
        if (kc <= 8)       VB_SUB_LAUNCH(8);
        else if (kc <= 16) VB_SUB_LAUNCH(16);
        else if (kc <= 20) VB_SUB_LAUNCH(20);
        else if (kc <= 24) VB_SUB_LAUNCH(24);
        else               VB_SUB_LAUNCH(32);
#undef VB_SUB_LAUNCH
        VB_CHECK_LAUNCH();
        ++nl;
    }
    if (launches) *launches = nl;
    return 0;
}

// R (n x p) = M - C (n x k fp64, leading dimension ldc) . (Vhi + Vlo) (k x p), fp64 accumulation; Vlo may be null.
// Rrows (optional): device array of n row pointers; when given, row i of the FINAL result is written to Rrows[i][0..p)
// instead of R + i p (rows may live in peer memory; each must be 8-byte aligned).  R is then only the scratch of the
// intermediate passes (k > 32) and may be null for k <= 32.
int project_subtract_hp_f32(const float* M, const double* C, int ldc, const float* Vhi, const float* Vlo, int k, int n,
                            size_t p, float* R, float* const* Rrows, int* launches, cudaStream_t st) {
    VB_REQUIRE(k > 0 && n > 0 && p > 0, "project_subtract_hp: empty problem");
    VB_REQUIRE(R != nullptr || (Rrows != nullptr && k <= KCMAX), "project_subtract_hp: R is required");
    const bool al = (p % 2 == 0) && (reinterpret_cast<uintptr_t>(M) % 8 == 0) && (reinterpret_cast<uintptr_t>(R) % 8 == 0);
    const unsigned grid = (unsigned)ceil_div(p, (size_t)PT);
    int nl = 0;
    for (int k0 = 0; k0 < k; k0 += KCMAX) {
        const int kc = (k - k0 < KCMAX) ? k - k0 : KCMAX;
        const float* src = (k0 == 0) ? M : R;
        float* const* rows = (k0 + KCMAX >= k) ? Rrows : nullptr;       // only the last pass scatters
        // two pixels per thread while 2 * KC fp64 components fit the register file comfortably
#define VB_SUBHP_LAUNCH(KCV, TWO)                                                                               \
        do {                                                                                                    \
            if (TWO && al) subtract_hp_kernel<KCV, 2><<<grid, PT / 2, 0, st>>>(src, C, ldc, Vhi, Vlo, n, p, k0, kc, R, rows); \
            else           subtract_hp_kernel<KCV, 1><<<grid, PT, 0, st>>>(src, C, ldc, Vhi, Vlo, n, p, k0, kc, R, rows);     \
        } while (0)
        if (kc <= 8)       VB_SUBHP_LAUNCH(8, true);
        else if (kc <= 16) VB_SUBHP_LAUNCH(16, true);
        else if (kc <= 20) VB_SUBHP_LAUNCH(20, true);
        else if (kc <= 24) VB_SUBHP_LAUNCH(24, true);
        else               VB_SUBHP_LAUNCH(32, false);
#undef VB_SUBHP_LAUNCH
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
