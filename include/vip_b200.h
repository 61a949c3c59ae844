/*
 * vip_b200 C ABI  --  B200 (sm_100a) kernels behind vip_hci.psfsub.pca / pca_annular and
 * vip_hci.preproc.cube_derotate.
 *
 * The reference (vortex-exoplanet/VIP, vip_hci 2.0.1) is pure Python and has no FFI of its
 * own; each entry point below replaces the numpy/scipy call sequence cited next to it
 * (paths relative to the reference checkout).  INTEGRATION.md shows the ctypes binding a VIP
 * maintainer would add.
 *
 * Conventions
 *   - all array arguments are DEVICE pointers unless the name ends in `_host`;
 *   - matrices are dense row-major, fp32 data / fp64 small matrices as stated;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value: 0 = ok, <0 = error (message from vb_last_error(), thread-local);
 *   - no ownership transfer: the caller allocates outputs and workspaces
 *     (sizes from the *_workspace_bytes queries);
 *   - calls are asynchronous on `stream` unless noted ("synchronises").
 */
#ifndef VIP_B200_H
#define VIP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ABI version (major*1000 + minor). */
int vb_version(void);
/* Message of the last failed call on this thread ("" if none). */
const char* vb_last_error(void);
/* Number of kernel launches issued through this library by this process (all entry points). */
long long vb_launch_count(void);

/* ---- Gramian (PCA decomposition, O(n^2 p) part) ------------------------------------------
 * G[n x n] (fp64) = A A^T for A[n x p] fp32.  deflate=1 routes through the mean-deflated
 * formulation (temporal mean removed, rank-2 terms re-added in fp64).
 * Replaces: np.linalg.svd(matrix.T) / np.dot(matrix, matrix.T)   psfsub/svd.py:466-475, 447-450 */
size_t vb_gram_workspace_bytes(int n, size_t p);
int vb_gram_f32(const float* A, int n, size_t p, int deflate, double* G, void* ws, size_t ws_bytes,
                void* stream);
/* Host cube (n x p fp32, pinned recommended) -> device matrix M AND G = M M^T in one pipelined call:
 * `nslabs` pixel slabs are uploaded with strided 2-D DMAs on a private copy stream and the Gramian of
 * each slab is accumulated on `stream` as soon as it has landed (SYRK hidden behind PCIe).
 * ws: vb_gram_workspace_bytes(n, p). */
int vb_upload_gram_f32(const float* host, int n, size_t p, float* M, double* G, void* ws, size_t ws_bytes,
                       int nslabs, void* stream);
/* C[na x nb] (fp64) = A B^T   (RDI / cube_sig projections: np.dot(V, matrix_emp.T), pca_fullfr.py:1728) */
size_t vb_cross_gram_workspace_bytes(int na, int nb);
int vb_cross_gram_f32(const float* A, int na, const float* B, int nb, size_t p, double* C, void* ws,
                      size_t ws_bytes, void* stream);

/* ---- symmetric eigensolver (fp64 block one-sided Jacobi) ----------------------------------
 * evals[n] descending, evecs[n x n] with ROW j = eigenvector j.  G is not modified.
 * info_host (optional, 2 ints): sweeps executed, converged flag.  Synchronises `stream`.
 * Replaces: LAPACK dgesdd/dsyevd inside numpy                      psfsub/svd.py:450, 470 */
size_t vb_eigh_workspace_bytes(int n);
int vb_eigh_f64(const double* G, int n, double* evals, double* evecs, int max_sweeps, double tol,
                void* ws, size_t ws_bytes, int* info_host, void* stream);

/* Cholesky whitening of a small SPD Gramian (n <= 128): Wt (n x n fp64, lower triangular) = R^-T with G = R^T R,
 * so that Wt G Wt^T = I (CholeskyQR).  Rows whose pivot is numerically zero are dropped (zero row/column of Wt).
 * Asynchronous.  Used by the orthonormalisation passes of the randomized SVD (scikit-learn's QR in
 * randomized_range_finder; call site psfsub/svd.py:487-491). */
int vb_chol_whiten_f64(const double* G, int n, double* Wt, void* stream);

/* Leading k (<= 24, block width <= n) eigenpairs only, by fp64 block subspace iteration with Rayleigh-Ritz
 * (O(n^2 B) per step instead of a full decomposition).  evals[k] descending, evecs[k x n] row j =
 * eigenvector j.  Iterates until max_j ||G x_j - theta_j x_j|| <= tol * theta_k.  info_host
 * (optional, 2 ints): iterations, converged flag.  Synchronises `stream`.
 * Replaces: the truncated use of the SVD, V[:ncomp]               psfsub/svd.py:471-473 */
size_t vb_eigh_topk_workspace_bytes(int n, int k);
int vb_eigh_topk_f64(const double* G, int n, int k, double tol, int max_iter, double* evals, double* evecs,
                     void* ws, size_t ws_bytes, int* info_host, void* stream);
/* Same solver without the host synchronisation: {iterations, converged} are copied to `info_pinned_host`
 * (two ints in PINNED host memory, valid until the stream has passed this call) by an async copy on `stream`;
 * the caller keeps enqueueing dependent work and checks the flag after its own synchronisation. */
int vb_eigh_topk_async_f64(const double* G, int n, int k, double tol, int max_iter, double* evals, double* evecs,
                           void* ws, size_t ws_bytes, int* info_pinned_host, void* stream);

/* ---- principal components and projection/subtraction --------------------------------------
 * V[k x p] (fp32) = Wt[k x n] (fp64) . M[n x p] (fp32), accumulated in fp64   psfsub/svd.py:451-459
 * R[n x p] = M - C[n x k] . V[k x p]   (R may alias M)            psfsub/pca_fullfr.py:1728-1731 */
int vb_pcs_f32(const double* Wt, const float* M, int k, int n, size_t p, float* V, void* stream);
/* Same product with the fp64 sum returned as an error-free pair of fp32 matrices, V = Vhi + Vlo (~48 bits):
 * the raw sketches of the randomized SVD span (sigma_0/sigma_k)^(2q+1) in magnitude before they are
 * orthonormalised, which fp32 storage cannot hold (sklearn's randomized_range_finder, extmath.py; call site
 * psfsub/svd.py:487-491). */
int vb_pcs_hilo_f32(const double* Wt, const float* M, int k, int n, size_t p, float* Vhi, float* Vlo,
                    void* stream);
int vb_project_subtract_f32(const float* M, const float* C, int ldc, const float* V, int k, int n,
                            size_t p, float* R, void* stream);
/* Same projection with fp64 coefficients C (n x k), the components as the error-free pair Vhi + Vlo of
 * vb_pcs_hilo_f32 (Vlo may be NULL) and fp64 accumulation; R is rounded to fp32 once.  The default of the PCA
 * path: with a 1e4-bright halo in the cube, fp32 copies of C and V leave a frame-independent error in every
 * residual pixel that survives the temporal median (measured 1e-3 of the final frame at config 2). */
int vb_project_subtract_hp_f32(const float* M, const double* C, int ldc, const float* Vhi, const float* Vlo, int k,
                               int n, size_t p, float* R, void* stream);
/* Scattered output: row i of the result is written to Rrows[i][0..p) (device array of n row pointers, each 8-byte
 * aligned) instead of R + i p.  The rows may live in PEER memory mapped into this process (NVLink): in the sharded
 * path every rank subtracts on its pixel shard and writes each residual row straight into the frame shard of the
 * rank that will derotate it -- the pixel->frame exchange rides on the projection pass, no separate all-to-all.
 * Rscratch (n x p) is only used by the intermediate passes when k > 32 and may be NULL otherwise. */
int vb_project_subtract_hp_rows_f32(const float* M, const double* C, int ldc, const float* Vhi, const float* Vlo,
                                    int k, int n, size_t p, float* Rscratch, float* const* Rrows, void* stream);
/* out = a - b elementwise (reconstructed = matrix - residuals for full_output) */
int vb_sub_f32(const float* a, const float* b, float* out, size_t count, void* stream);

/* ---- FFT three-shear derotation of a cube --------------------------------------------------
 * out[f] = frame f of `in` rotated exactly as frame_rotate(imlib='vip-fft', edge_blend=None)
 * does for angle -angle_list[f]; the per-frame scalars are computed on the host with the
 * reference's own numpy expressions: krot[f] = rint(angle/90) mod 4 (0 if angle <= 45),
 * a[f] = tan(dangle/2), b[f] = -sin(dangle).  S = frame size (square), N = working plane size,
 * y0 = offset of the frame in the plane (geometry of frame_pad / frame_center).
 * mask: pixels that are NaN (mask_is_nan=1) or == mask_val are restored to mask_val in `out`;
 * zero_masked=1 additionally feeds pixels == mask_val as 0 into the rotation.
 * Replaces: cube_derotate -> frame_rotate -> rotate_fft -> _fft_shear
 *           preproc/derotation.py:331-399, 51-328, 542-622, 625-640 */
size_t vb_derotate_scratch_bytes(int nframes, int S, int N, size_t max_bytes);
int vb_derotate_f32(const float* in, float* out, int nframes, int S, int N, int y0, const int* krot,
                    const double* a, const double* b, float mask_val, int mask_is_nan, int zero_masked,
                    void* scratch, size_t scratch_bytes, int force_direct, void* stream);
/* The same rotation with a SCATTERED output for the sharded path: output row r of local frame f goes to
 *   out_bases[r / rows_per_shard] + (frame_offset + f) * frame_stride + (r % rows_per_shard) * S      (floats)
 * i.e. straight into the (n_frames_total x pixels_of_shard) slab of the rank that owns those image rows; the
 * bases may be PEER allocations mapped into this process (NVLink), so the frame->pixel exchange that precedes the
 * temporal median needs no separate all-to-all.  out_bases_host: HOST array of `nshards` (<= 8) device pointers;
 * rows_per_shard even, rows_per_shard * nshards == S; power-of-two frames only (N = 4S). */
int vb_derotate_scatter_f32(const float* in, int nframes, int S, int N, int y0, const int* krot, const double* a,
                            const double* b, float mask_val, int mask_is_nan, int zero_masked, void* scratch,
                            size_t scratch_bytes, void* const* out_bases_host, int nshards, int rows_per_shard,
                            long long frame_stride, int frame_offset, void* stream);

/* ---- temporal collapse ---------------------------------------------------------------------
 * cube[n x p] -> out[p].  mode: 0 median, 1 mean, 2 sum, 3 max, 4 absmean, 5 wmean (w: n
 * doubles, out: double[p]), 6 trimmean (mean of sorted[trim_k : trim_k+trim_n]).  NaN-aware.
 * Replaces: cube_collapse                                          preproc/subsampling.py:30-116 */
int vb_collapse_f32(const float* cube, int n, size_t p, int mode, const double* w, int trim_k, int trim_n,
                    void* out, void* stream);

/* ---- annular PCA: batched per-frame library eigenproblems -----------------------------------
 * For each problem q (target frame[q], library rows idx[q][0..len[q]) of the segment matrix):
 * top-`ncomp` eigenpairs (theta_j, x_j) of G[idx,idx] by fp64 block subspace iteration, then the
 * projection weights  w = sum_j x_j (x_j . Gt[frame, idx]) / theta_j  written to W[q][idx]
 * (W: nprob x n fp32, zero-initialised by the caller; Gt = NULL means G).
 * The residuals follow as  R = A - W . A_lib  (vb_pcs_f32 + vb_sub_f32).  iters[q] < 0 flags a
 * problem that hit max_iter.
 * Replaces: do_pca_patch -> get_eigenvectors -> svd_wrapper       psfsub/pca_local.py:830-909 */
int vb_annular_weights_f64(const double* G, const double* Gt, int n, const int* idx, const int* len,
                           const int* frame, int nprob, int Lmax, int ncomp, double tol, int max_iter, float* W,
                           int* iters, void* stream);
/* Direct solver for the problems listed in plist[0..nlist) (those the subspace iteration left
 * unconverged: flat, noise-dominated spectra): Householder tridiagonalisation of G[idx,idx] in the
 * workspace ws (nlist * Lmax^2 doubles), Sturm multisection for the ncomp largest eigenvalues,
 * simultaneous inverse iteration, back-transformation; same W/iters outputs (iters = 100000). */
int vb_annular_direct_f64(const double* G, const double* Gt, int n, const int* idx, const int* len,
                          const int* frame, int nprob, int Lmax, int ncomp, const int* plist, int nlist, float* W,
                          int* iters, double* ws, void* stream);
/* ncomp='auto' of pca_annular: the direct solver with `kmax` (<= 24) eigenpairs per problem; the number of components
 * used for problem q is chosen by the reference's noise-decay rule -- smallest m >= 2 whose step
 * std(res_{m-1}) - std(res_m) of the LIBRARY residuals falls below `noise_tol` -- evaluated from the eigenpairs and the
 * row sums `rowsum[n]` of the library matrix (npx pixels per row), and returned in ncomp_out[q] (negative: the rule
 * wanted more than kmax components; |value| were used).
 * Replaces: get_eigenvectors(ncomp='auto', mode='noise')          psfsub/svd.py:622-672 */
int vb_annular_auto_f64(const double* G, const double* Gt, int n, const int* idx, const int* len,
                        const int* frame, int nprob, int Lmax, int kmax, const double* rowsum, double npx,
                        double noise_tol, const int* plist, int nlist, float* W, int* iters, int* ncomp_out,
                        double* ws, void* stream);
/* dst[n x npx] = src[n x p][:, cols]  and the inverse scatter (matrix_segm = array[:, yy, xx],
 * cube_out[fr][yy, xx] = residuals[fr];  psfsub/pca_local.py:713, 786-787) */
int vb_gather_columns_f32(const float* src, int n, size_t p, const int* cols, int npx, float* dst, void* stream);
int vb_scatter_columns_f32(const float* src, int n, int npx, const int* cols, size_t p, float* dst, void* stream);

/* ---- batched fp32 GEMM ---------------------------------------------------------------------------
 * C[b] (M x N, ldc) = alpha * A[ia] (M x K, lda) * op(B[ib]) + beta * C[b]   for b < batch, with
 * ia = a_mod ? b % a_mod : b  (same for ib): operators that depend on the spectral channel only are
 * shared across ADI frames.  trans_b = 0: B is K x N row-major; 1: B is N x K row-major (C = A B^T).
 * Replaces: the FFT zoom of scale_fft, as  Re(L X L^T)            preproc/rescaling.py:1114-1217 */
int vb_gemm_f32(const float* A, long long lda, long long strideA, int a_mod, const float* B, long long ldb,
                long long strideB, int b_mod, int trans_b, float* C, long long ldc, long long strideC, int M,
                int N, int K, float alpha, float beta, int batch, void* stream);

/* ---- batched GEMM on the tensor cores (tcgen05, fp32-grade bf16x3 products) --------------------------
 * vb_split3_bf16: fp32 rows (rows x K, row stride ldx) -> three bf16 planes (error-free split x = x1 + x2 + x3),
 * plane p at planes + p * plane_stride, row stride ldp (elements; multiple of 8, >= K).
 * vb_gemm_bf16x3_tc: C[b] (M x N) = A[b % a_mod] (M x K) . B[b] (N x K)^T, b < batch, from such planes (A planes hold
 * a_mod * M rows, B planes batch * N rows; six cross products per K-block, fp32 accumulation in tensor memory).
 * Output: fp32 C (ldc, strideC; planesC = NULL), or the bf16x3 planes of C (planesC != NULL, row stride ldpC) with
 * the rows r >= msplit stored on row r - msplit at column offset N -- the layout in which  [Lr; Li] X^T  becomes the
 * K-major operand of the second product of  Re(L X L^T) = [Lr | -Li] . [Lr X^T | Li X^T]^T.
 * Replaces: the FFT zoom of scale_fft for IFS cubes              preproc/rescaling.py:1114-1217, 324-475 */
int vb_split3_bf16(const float* X, long long rows, int K, long long ldx, void* planes, long long ldp,
                   long long plane_stride, void* stream);
int vb_gemm_bf16x3_tc(const void* planesA, long long ldpA, long long strideA, int a_mod, const void* planesB,
                      long long ldpB, long long strideB, int M, int N, int K, int batch, float* C, long long ldc,
                      long long strideC, void* planesC, long long ldpC, long long strideCp, int msplit, void* stream);

/* ---- Fourier sub-pixel shift (frame recentring, fake-companion injection) -------------------------
 * The vip-fft shift of a frame is  out = Ty X Tx^T - checkerboard term  with real Toeplitz operators
 * T[m][n] = Re D_N(m - n - s) of the zero-padded even plane of N pixels (csrc/shift.cu).
 * vb_shift_operators_f32: T[f] (L x L, fp32) for shift[f] (pixels, fp64) and plane size nplane[f];
 * the two products are vb_gemm_f32 calls; vb_checker_correct_f32 then applies the Nyquist term
 * out[f][r][c] -= (-1)^(r+c) coef[f] sum_{r',c'} (-1)^(r'+c') in[f][r'][c'],  coef = sin(pi sx) sin(pi sy) / N^2
 * (kappa_ws: nframes doubles).
 * Replaces: cube_shift -> frame_shift(imlib='vip-fft')             preproc/recentering.py:257-305, 122-189 */
int vb_shift_operators_f32(const double* shift, const int* nplane, int nframes, int L, float* T, void* stream);
int vb_checker_correct_f32(const float* in, float* out, int nframes, int ny, int nx, const double* coef,
                           double* kappa_ws, void* stream);

/* ---- host -> device upload of a pixel shard (strided rows of the host cube), one DMA ---------------
 * dst[r][0..width) = src_host[r][0..width) for r < height, pitches in bytes (cudaMemcpy2DAsync). */
int vb_memcpy2d_h2d(void* dst, size_t dpitch, const void* src_host, size_t spitch, size_t width_bytes,
                    size_t height, void* stream);
/* Contiguous host -> device copy.  A PAGEABLE source (a plain numpy array) is staged through a pinned double buffer by
 * several host threads while the previous chunk is on the DMA engine (the driver's own staging is single-threaded,
 * ~10 GB/s); a pinned source is one cudaMemcpyAsync.  The source may be reused when the call returns. */
int vb_memcpy_h2d_staged(void* dst, const void* src_host, size_t nbytes, void* stream);
/* The same for `height` strided rows of `width_bytes` (source pitch spitch) into a CONTIGUOUS device buffer: the
 * frames [f0, f1) of every channel of a (z, n, H, W) IFS cube, uploaded chunk by chunk on a copy stream while the
 * previous chunk is processed.  Returns after the last DMA has completed. */
int vb_memcpy2d_h2d_staged(void* dst, const void* src_host, size_t spitch, size_t width_bytes, size_t height,
                           void* stream);

/* ---- blob detection and FITS decode ------------------------------------------------------
 * mask[y][x] = 1 where img[y][x] is the maximum of its (2 d + 1)^2 neighbourhood (edge-replicated), exceeds
 * `threshold` and lies more than d pixels from the border; NaNs never are peaks and never hide one.
 * Replaces: skimage.feature.peak_local_max as called by detection      metrics/detection.py:277-279 */
int vb_local_max_mask_f32(const float* img, int H, int W, int min_distance, float threshold, unsigned char* mask,
                          void* stream);
/* Data unit of a FITS image HDU on the device: `count` big-endian samples of type BITPIX (8, 16, 32, 64, -32, -64)
 * at `raw` (device) -> out[i] = fp32(BSCALE * sample + BZERO).
 * Replaces: the host-side conversion of open_fits (astropy + np.array(data, dtype))      fits/fits.py:119-146 */
int vb_fits_decode_f32(const void* raw, int bitpix, size_t count, double bscale, double bzero, float* out,
                       void* stream);

/* ---- S/N of test resolution elements, S/N map ------------------------------------------
 * Exact circular-aperture sums: out[a] = sum over pixels of area(circle(xs[a], ys[a], r) ∩ unit pixel) * img.
 * Replaces: photutils CircularAperture + aperture_photometry(method='exact')   metrics/snr_source.py:393-397 */
int vb_aperture_sums_f64(const float* img, int H, int W, const double* xs, const double* ys, int nap, double r,
                         double* out, void* stream);
/* S/N (Mawet et al. 2014 small-sample penalty) at npts integer pixel positions (px, py): one warp per position sums
 * the non-overlapping apertures of its ring.  img2 (optional second frame, e.g. opposite derotation) adds its
 * apertures to the noise sample, use2alone keeps only those; flux_out (optional) receives the source fluxes.
 * (cy, cx) = frame centre.  Replaces: snr() per pixel forked over processes by snrmap()
 * metrics/snr_source.py:32-204, 229-318, 321-455 */
int vb_snr_points_f64(const float* img, const float* img2, int H, int W, const int* px, const int* py, int npts,
                      double fwhm, double cy, double cx, int exclude_negative_lobes, int use2alone, double* snr_out,
                      double* flux_out, void* stream);

/* ---- measurement aid (bench.py) -------------------------------------------------------------
 * Register-only FFMA stream: blocks x 256 threads x iters x 32 fused multiply-adds; `out` receives
 * blocks*256 floats.  Timed by the caller with CUDA events to report the FP32 FMA peak of this GPU at its
 * current clocks (the yardstick of the FFT-bound derotation; no reference counterpart). */
int vb_fp32_probe(float* out, int blocks, int iters, void* stream);

/* ---- measurement hook (bench.py) ----------------------------------------------------------
 * vb_profile_enable(1): CUDA events are recorded around each of the three shear kernels of
 * vb_derotate_f32 on its stream; vb_profile_read(out4_host) synchronises on them and returns the
 * summed elapsed ms of pass 1 / 2 / 3 and the number of chunk launches timed. */
void vb_profile_enable(int on);
int vb_profile_read(float* out4_host);

#ifdef __cplusplus
}
#endif
#endif /* VIP_B200_H */
