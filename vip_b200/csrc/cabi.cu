// C ABI of libvipb200.so (declared in include/vip_b200.h).
#include "common.cuh"
#include "../../include/vip_b200.h"
#include <cstdarg>
#include <cstdlib>
#include <atomic>
#include <map>
#include <mutex>
#include <vector>

namespace vb {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---- implemented in the other translation units
size_t gram_workspace_bytes(int n, size_t p);
int gram_f32(const float*, int, size_t, int, double*, void*, size_t, int, int*, cudaStream_t);
int cross_gram_f32(const float*, int, const float*, int, size_t, double*, void*, size_t, int, cudaStream_t);
int upload_gram_f32(const float*, int, size_t, float*, double*, void*, size_t, int, int*, cudaStream_t);
size_t eigh_workspace_bytes(int n);
int eigh_f64(const double*, int, double*, double*, int, double, void*, size_t, int*, int*, cudaStream_t);
size_t eigh_topk_workspace_bytes(int n, int B);
int topk_block_width(int k);
int chol_whiten_f64(const double*, int, double*, cudaStream_t);
int eigh_topk_f64(const double*, int, int, double, int, double*, double*, void*, size_t, int*, int*, cudaStream_t,
                  int* async_info = nullptr);
int pcs_f32(const double*, const float*, int, int, size_t, float*, float*, int*, cudaStream_t);
int project_subtract_f32(const float*, const float*, int, const float*, int, int, size_t, float*, int*,
                         cudaStream_t);
int project_subtract_hp_f32(const float*, const double*, int, const float*, const float*, int, int, size_t, float*,
                            float* const*, int*, cudaStream_t);
int sub_f32(const float*, const float*, float*, size_t, cudaStream_t);
struct OutMap { float* base[8]; int nshards; int rows_per; long long fstride; int fofs; };
struct RotParams { int S; int N; int y0; int zero_masked; int mask_is_nan; float mask_val; OutMap om; };
size_t derotate_scratch_bytes_per_frame(int S, int N);
size_t derotate_scratch_bytes_min(int S, int N);
int shift_operators(const double*, const int*, int, int, float*, cudaStream_t);
int checker_correct(const float*, float*, int, int, int, const double*, double*, cudaStream_t);
int derotate_run(const float*, float*, int, const RotParams&, const int*, const double*, const double*,
                 const float2*, void*, size_t, int, int*, cudaStream_t);
int collapse_f32(const float*, int, size_t, int, const double*, int, int, void*, cudaStream_t);
void profile_enable(int on);
struct GemmArgs {
    const float* A; long long lda, strideA; int a_mod;
    const float* B; long long ldb, strideB; int b_mod;
    float* C; long long ldc, strideC;
    int M, N, K;
    float alpha, beta;
};
int gemm_f32(const GemmArgs&, int, int, cudaStream_t);
int split3_bf16(const float*, long long, int, long long, void*, long long, long long, cudaStream_t);
int gemm_bf16x3_tc(const void*, long long, long long, int, const void*, long long, long long, int, int, int, int,
                   float*, long long, long long, void*, long long, long long, int, cudaStream_t);
int annular_weights(const double*, const double*, int, const int*, const int*, const int*, int, int, int, double,
                    int, float*, int*, cudaStream_t);
int annular_auto_weights(const double*, const double*, int, const int*, const int*, const int*, int, int, int,
                         const double*, double, double, const int*, int, float*, int*, int*, double*, cudaStream_t);
int annular_direct_weights(const double*, const double*, int, const int*, const int*, const int*, int, int, int,
                           const int*, int, float*, int*, double*, cudaStream_t);
int gather_columns(const float*, int, size_t, const int*, int, float*, cudaStream_t);
int scatter_columns(const float*, int, int, const int*, size_t, float*, cudaStream_t);
int profile_read(float* out);
int fp32_probe(float*, int, int, cudaStream_t);
int memcpy_h2d_staged(void*, const void*, size_t, cudaStream_t);
int memcpy2d_h2d_staged(void*, const void*, size_t, size_t, size_t, cudaStream_t);
int local_max_mask(const float*, int, int, int, float, unsigned char*, cudaStream_t);
int fits_decode(const void*, int, size_t, double, double, float*, cudaStream_t);
int aperture_sums(const float*, int, int, const double*, const double*, int, double, double*, cudaStream_t);
int snr_points(const float*, const float*, int, int, const int*, const int*, int, double, double, double, int, int,
               double*, double*, cudaStream_t);

// exp(-2 pi i j / N) tables for the FFT path, one per (device, N), built in fp64 on the host
static const float2* twiddle_table(int N) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, float2*> cache;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find({dev, N});
    if (it != cache.end()) return it->second;
    std::vector<float2> h(N);
    for (int j = 0; j < N; ++j) {
        const double ang = -2.0 * 3.14159265358979323846 * (double)j / (double)N;
        h[j] = make_float2((float)cos(ang), (float)sin(ang));
    }
    float2* d = nullptr;
    if (cudaMalloc(&d, N * sizeof(float2)) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h.data(), N * sizeof(float2), cudaMemcpyHostToDevice) != cudaSuccess) return nullptr;
    cache[{dev, N}] = d;
    return d;
}

}  // namespace vb

using namespace vb;

extern "C" {

int vb_version(void) { return 1000; }
const char* vb_last_error(void) { return g_err; }
long long vb_launch_count(void) { return g_launches.load(); }

size_t vb_gram_workspace_bytes(int n, size_t p) { return gram_workspace_bytes(n, p); }

int vb_gram_f32(const float* A, int n, size_t p, int deflate, double* G, void* ws, size_t ws_bytes,
                void* stream) {
    int nl = 0;
    const int rc = gram_f32(A, n, p, deflate, G, ws, ws_bytes, 0, &nl, (cudaStream_t)stream);
    g_launches += nl;
    return rc;
}

int vb_upload_gram_f32(const float* host, int n, size_t p, float* M, double* G, void* ws, size_t ws_bytes,
                       int nslabs, void* stream) {
    int nl = 0;
    const int rc = upload_gram_f32(host, n, p, M, G, ws, ws_bytes, nslabs, &nl, (cudaStream_t)stream);
    g_launches += nl;
    return rc;
}

size_t vb_cross_gram_workspace_bytes(int na, int nb) {
    return (size_t)ceil_div(na, 64) * ceil_div(nb, 64) * sizeof(int2) + 256;   // upper bound: 64- or 128-row tiles
}

int vb_cross_gram_f32(const float* A, int na, const float* B, int nb, size_t p, double* C, void* ws,
                      size_t ws_bytes, void* stream) {
    const int rc = cross_gram_f32(A, na, B, nb, p, C, ws, ws_bytes, 0, (cudaStream_t)stream);
    g_launches += 1;
    return rc;
}

size_t vb_eigh_workspace_bytes(int n) { return eigh_workspace_bytes(n); }

int vb_eigh_f64(const double* G, int n, double* evals, double* evecs, int max_sweeps, double tol, void* ws,
                size_t ws_bytes, int* info_host, void* stream) {
    int nl = 0;
    const int rc = eigh_f64(G, n, evals, evecs, max_sweeps, tol, ws, ws_bytes, info_host, &nl,
                            (cudaStream_t)stream);
    g_launches += nl;
    return rc;
}

int vb_chol_whiten_f64(const double* G, int n, double* Wt, void* stream) {
    g_launches += 1;
    return chol_whiten_f64(G, n, Wt, (cudaStream_t)stream);
}

size_t vb_eigh_topk_workspace_bytes(int n, int k) {
    return eigh_topk_workspace_bytes(n, topk_block_width(k));
}

int vb_eigh_topk_f64(const double* G, int n, int k, double tol, int max_iter, double* evals, double* evecs,
                     void* ws, size_t ws_bytes, int* info_host, void* stream) {
    int nl = 0;
    const int rc = eigh_topk_f64(G, n, k, tol, max_iter, evals, evecs, ws, ws_bytes, info_host, &nl,
                                 (cudaStream_t)stream);
    g_launches += nl;
    return rc;
}

int vb_eigh_topk_async_f64(const double* G, int n, int k, double tol, int max_iter, double* evals, double* evecs,
                           void* ws, size_t ws_bytes, int* info_pinned_host, void* stream) {
    VB_REQUIRE(info_pinned_host != nullptr, "eigh_topk_async: info_pinned_host is required");
    int nl = 0;
    const int rc = eigh_topk_f64(G, n, k, tol, max_iter, evals, evecs, ws, ws_bytes, nullptr, &nl,
                                 (cudaStream_t)stream, info_pinned_host);
    g_launches += nl;
    return rc;
}

int vb_pcs_f32(const double* Wt, const float* M, int k, int n, size_t p, float* V, void* stream) {
    int nl = 0;
    const int rc = pcs_f32(Wt, M, k, n, p, V, nullptr, &nl, (cudaStream_t)stream);
    g_launches += nl;
    return rc;
}

int vb_pcs_hilo_f32(const double* Wt, const float* M, int k, int n, size_t p, float* Vhi, float* Vlo,
                    void* stream) {
    VB_REQUIRE(Vlo != nullptr, "pcs_hilo: Vlo is required");
    int nl = 0;
    const int rc = pcs_f32(Wt, M, k, n, p, Vhi, Vlo, &nl, (cudaStream_t)stream);
    g_launches += nl;
    return rc;
}

int vb_project_subtract_f32(const float* M, const float* C, int ldc, const float* V, int k, int n, size_t p,
                            float* R, void* stream) {
    int nl = 0;
    const int rc = project_subtract_f32(M, C, ldc, V, k, n, p, R, &nl, (cudaStream_t)stream);
    g_launches += nl;
    return rc;
}

int vb_project_subtract_hp_f32(const float* M, const double* C, int ldc, const float* Vhi, const float* Vlo, int k,
                               int n, size_t p, float* R, void* stream) {
    int nl = 0;
    const int rc = project_subtract_hp_f32(M, C, ldc, Vhi, Vlo, k, n, p, R, nullptr, &nl, (cudaStream_t)stream);
    g_launches += nl;
    return rc;
}

int vb_project_subtract_hp_rows_f32(const float* M, const double* C, int ldc, const float* Vhi, const float* Vlo,
                                    int k, int n, size_t p, float* Rscratch, float* const* Rrows, void* stream) {
    VB_REQUIRE(Rrows != nullptr, "project_subtract_hp_rows: Rrows is required");
    int nl = 0;
    const int rc = project_subtract_hp_f32(M, C, ldc, Vhi, Vlo, k, n, p, Rscratch, Rrows, &nl, (cudaStream_t)stream);
    g_launches += nl;
    return rc;
}

int vb_sub_f32(const float* a, const float* b, float* out, size_t count, void* stream) {
    g_launches += 1;
    return sub_f32(a, b, out, count, (cudaStream_t)stream);
}

size_t vb_derotate_scratch_bytes(int nframes, int S, int N, size_t max_bytes) {
    const size_t per = derotate_scratch_bytes_per_frame(S, N);
    size_t want = per * (size_t)nframes;
    if (max_bytes && want > max_bytes) {
        size_t frames = max_bytes / per;
        if (frames < 1) frames = 1;
        want = frames * per;
    }
    const size_t floor_bytes = derotate_scratch_bytes_min(S, N);   // one frame on any path (force_direct)
    return want < floor_bytes ? floor_bytes : want;
}

int vb_derotate_f32(const float* in, float* out, int nframes, int S, int N, int y0, const int* krot,
                    const double* a, const double* b, float mask_val, int mask_is_nan, int zero_masked,
                    void* scratch, size_t scratch_bytes, int force_direct, void* stream) {
    VB_REQUIRE(nframes > 0 && S > 0, "derotate: empty cube");
    VB_REQUIRE(N % 2 == 0 && N > S && y0 >= 0 && y0 + S + 1 <= N, "derotate: bad geometry S=%d N=%d y0=%d", S,
               N, y0);
    RotParams g{S, N, y0, zero_masked, mask_is_nan, mask_val, OutMap{}};
    const float2* tw = nullptr;
    const bool pow2 = (N & (N - 1)) == 0 && N >= 512 && N <= 4096 && N == 4 * S;
    if (pow2 && !force_direct) {
        tw = twiddle_table(N);
        VB_REQUIRE(tw != nullptr, "derotate: could not build the twiddle table");
    }
    int nl = 0;
    const int rc = derotate_run(in, out, nframes, g, krot, a, b, tw, scratch, scratch_bytes, force_direct,
                                &nl, (cudaStream_t)stream);
    g_launches += nl;
    return rc;
}

int vb_derotate_scatter_f32(const float* in, int nframes, int S, int N, int y0, const int* krot, const double* a,
                            const double* b, float mask_val, int mask_is_nan, int zero_masked, void* scratch,
                            size_t scratch_bytes, void* const* out_bases_host, int nshards, int rows_per_shard,
                            long long frame_stride, int frame_offset, void* stream) {
    VB_REQUIRE(nframes > 0 && S > 0, "derotate_scatter: empty cube");
    VB_REQUIRE(N % 2 == 0 && N > S && y0 >= 0 && y0 + S + 1 <= N, "derotate_scatter: bad geometry S=%d N=%d y0=%d",
               S, N, y0);
    VB_REQUIRE(out_bases_host != nullptr && nshards >= 1 && nshards <= 8, "derotate_scatter: 1..8 output shards");
    VB_REQUIRE(rows_per_shard > 0 && rows_per_shard % 2 == 0 && rows_per_shard * nshards == S,
               "derotate_scatter: rows_per_shard must be even and rows_per_shard * nshards == S");
    const bool pow2 = (N & (N - 1)) == 0 && N >= 512 && N <= 4096 && N == 4 * S;
    VB_REQUIRE(pow2, "derotate_scatter: needs power-of-two frames (N = 4S, 512 <= N <= 4096)");
    RotParams g{S, N, y0, zero_masked, mask_is_nan, mask_val, OutMap{}};
    for (int h = 0; h < nshards; ++h) g.om.base[h] = reinterpret_cast<float*>(out_bases_host[h]);
    g.om.nshards = nshards;
    g.om.rows_per = rows_per_shard;
    g.om.fstride = frame_stride;
    g.om.fofs = frame_offset;
    const float2* tw = twiddle_table(N);
    VB_REQUIRE(tw != nullptr, "derotate_scatter: could not build the twiddle table");
    int nl = 0;
    const int rc = derotate_run(in, nullptr, nframes, g, krot, a, b, tw, scratch, scratch_bytes, 0, &nl,
                                (cudaStream_t)stream);
    g_launches += nl;
    return rc;
}

int vb_shift_operators_f32(const double* shift, const int* nplane, int nframes, int L, float* T, void* stream) {
    g_launches += 1;
    return shift_operators(shift, nplane, nframes, L, T, (cudaStream_t)stream);
}

int vb_checker_correct_f32(const float* in, float* out, int nframes, int ny, int nx, const double* coef,
                           double* kappa_ws, void* stream) {
    g_launches += 2;
    return checker_correct(in, out, nframes, ny, nx, coef, kappa_ws, (cudaStream_t)stream);
}

int vb_collapse_f32(const float* cube, int n, size_t p, int mode, const double* w, int trim_k, int trim_n,
                    void* out, void* stream) {
    g_launches += 1;
    return collapse_f32(cube, n, p, mode, w, trim_k, trim_n, out, (cudaStream_t)stream);
}

int vb_annular_weights_f64(const double* G, const double* Gt, int n, const int* idx, const int* len,
                           const int* frame, int nprob, int Lmax, int ncomp, double tol, int max_iter, float* W,
                           int* iters, void* stream) {
    g_launches += 1;
    return annular_weights(G, Gt, n, idx, len, frame, nprob, Lmax, ncomp, tol, max_iter, W, iters,
                           (cudaStream_t)stream);
}

int vb_annular_direct_f64(const double* G, const double* Gt, int n, const int* idx, const int* len,
                          const int* frame, int nprob, int Lmax, int ncomp, const int* plist, int nlist, float* W,
                          int* iters, double* ws, void* stream) {
    g_launches += 1;
    return annular_direct_weights(G, Gt, n, idx, len, frame, nprob, Lmax, ncomp, plist, nlist, W, iters, ws,
                                  (cudaStream_t)stream);
}

int vb_annular_auto_f64(const double* G, const double* Gt, int n, const int* idx, const int* len,
                        const int* frame, int nprob, int Lmax, int kmax, const double* rowsum, double npx,
                        double noise_tol, const int* plist, int nlist, float* W, int* iters, int* ncomp_out,
                        double* ws, void* stream) {
    g_launches += 1;
    return annular_auto_weights(G, Gt, n, idx, len, frame, nprob, Lmax, kmax, rowsum, npx, noise_tol, plist, nlist, W,
                                iters, ncomp_out, ws, (cudaStream_t)stream);
}

int vb_gather_columns_f32(const float* src, int n, size_t p, const int* cols, int npx, float* dst, void* stream) {
    g_launches += 1;
    return gather_columns(src, n, p, cols, npx, dst, (cudaStream_t)stream);
}

int vb_scatter_columns_f32(const float* src, int n, int npx, const int* cols, size_t p, float* dst, void* stream) {
    g_launches += 1;
    return scatter_columns(src, n, npx, cols, p, dst, (cudaStream_t)stream);
}

int vb_gemm_f32(const float* A, long long lda, long long strideA, int a_mod, const float* B, long long ldb,
                long long strideB, int b_mod, int trans_b, float* C, long long ldc, long long strideC, int M,
                int N, int K, float alpha, float beta, int batch, void* stream) {
    GemmArgs g{A, lda, strideA, a_mod, B, ldb, strideB, b_mod, C, ldc, strideC, M, N, K, alpha, beta};
    g_launches += 1;
    return gemm_f32(g, trans_b, batch, (cudaStream_t)stream);
}

int vb_split3_bf16(const float* X, long long rows, int K, long long ldx, void* planes, long long ldp,
                   long long plane_stride, void* stream) {
    g_launches += 1;
    return split3_bf16(X, rows, K, ldx, planes, ldp, plane_stride, (cudaStream_t)stream);
}

int vb_gemm_bf16x3_tc(const void* planesA, long long ldpA, long long strideA, int a_mod, const void* planesB,
                      long long ldpB, long long strideB, int M, int N, int K, int batch, float* C, long long ldc,
                      long long strideC, void* planesC, long long ldpC, long long strideCp, int msplit, void* stream) {
    g_launches += 1;
    return gemm_bf16x3_tc(planesA, ldpA, strideA, a_mod, planesB, ldpB, strideB, M, N, K, batch, C, ldc, strideC,
                          planesC, ldpC, strideCp, msplit, (cudaStream_t)stream);
}

int vb_memcpy2d_h2d(void* dst, size_t dpitch, const void* src_host, size_t spitch, size_t width_bytes,
                    size_t height, void* stream) {
    VB_CHECK_CUDA(cudaMemcpy2DAsync(dst, dpitch, src_host, spitch, width_bytes, height, cudaMemcpyHostToDevice,
                                    (cudaStream_t)stream));
    return 0;
}

int vb_local_max_mask_f32(const float* img, int H, int W, int min_distance, float threshold, unsigned char* mask,
                          void* stream) {
    g_launches += 1;
    return local_max_mask(img, H, W, min_distance, threshold, mask, (cudaStream_t)stream);
}

int vb_fits_decode_f32(const void* raw, int bitpix, size_t count, double bscale, double bzero, float* out,
                       void* stream) {
    g_launches += 1;
    return fits_decode(raw, bitpix, count, bscale, bzero, out, (cudaStream_t)stream);
}

int vb_aperture_sums_f64(const float* img, int H, int W, const double* xs, const double* ys, int nap, double r,
                         double* out, void* stream) {
    g_launches += 1;
    return aperture_sums(img, H, W, xs, ys, nap, r, out, (cudaStream_t)stream);
}

int vb_snr_points_f64(const float* img, const float* img2, int H, int W, const int* px, const int* py, int npts,
                      double fwhm, double cy, double cx, int exclude_negative_lobes, int use2alone, double* snr_out,
                      double* flux_out, void* stream) {
    g_launches += 1;
    return snr_points(img, img2, H, W, px, py, npts, fwhm, cy, cx, exclude_negative_lobes, use2alone, snr_out,
                      flux_out, (cudaStream_t)stream);
}

int vb_fp32_probe(float* out, int blocks, int iters, void* stream) {
    g_launches += 1;
    return fp32_probe(out, blocks, iters, (cudaStream_t)stream);
}

int vb_memcpy_h2d_staged(void* dst, const void* src_host, size_t nbytes, void* stream) {
    return memcpy_h2d_staged(dst, src_host, nbytes, (cudaStream_t)stream);
}

int vb_memcpy2d_h2d_staged(void* dst, const void* src_host, size_t spitch, size_t width_bytes, size_t height,
                           void* stream) {
    return memcpy2d_h2d_staged(dst, src_host, spitch, width_bytes, height, (cudaStream_t)stream);
}

void vb_profile_enable(int on) { profile_enable(on); }
int vb_profile_read(float* out4_host) { return profile_read(out4_host); }

}  // extern "C"
