// Annular (per-frame library) PCA: batched top-k eigenproblems on sub-Gramians, plus the
// column gather/scatter between the cube and a segment matrix.
//
// Role in the reference: do_pca_patch (src/vip_hci/psfsub/pca_local.py:830-909) runs, for every
// frame f of every annular segment, an SVD of the library  A[I_f]  (I_f from _find_indices_adi,
// at most max_frames_lib rows) and subtracts the projection of the frame on its top-k right
// singular vectors.  With G = A A^T computed once per segment:
//     (theta_j, x_j) = top-k eigenpairs of G[I_f, I_f],   g = G[I_f, f],
//     w = sum_j x_j (x_j . g) / theta_j,                  R_f = A_f - w^T A[I_f].
// One CTA per (segment, frame) problem: block subspace iteration with Rayleigh-Ritz in fp64
// (block width B > k), the B x B Ritz problem solved by a one-sided Jacobi in shared memory.
// Subspace iteration needs O(L^2 B) per step instead of O(L^3) for a full decomposition and
// converges in 4-20 steps on ADI sub-Gramians (gap-dependent; iterated to a residual bound).
#include "common.cuh"

namespace vb {

constexpr int AT = 256;  // threads per problem

// deterministic start vectors: splitmix-style hash -> uniform(-1, 1)
__device__ __forceinline__ double hash_unit(unsigned int a, unsigned int b) {
    unsigned long long z = ((unsigned long long)a << 32 | b) + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

struct AnnularArgs {
    const double* G;        // (n, n) Gramian of the library matrix (A - A_sig, or A)
    const double* Gt;       // (n, n) cross Gramian: Gt[f][i] = target_emp[f] . library[i]  (== G without cube_sig)
    int n;
    const int* idx;         // (nprob, Lmax) library indices, valid prefix of length len[q]
    const int* len;         // (nprob)
    const int* frame;       // (nprob) target frame of each problem
    int Lmax;
    int ncomp;
    double tol;
    int max_iter;
    float* W;               // (nprob, n) output weights, row q = problem q; must be zero-initialised
    int* iters;             // (nprob) iterations used (negative: not converged)
    // ncomp='auto' (direct solver only; get_eigenvectors, psfsub/svd.py:640-672): when rowsum != nullptr the number
    // of components of problem q is chosen by the reference's noise-decay rule among the ncomp computed eigenpairs
    const double* rowsum = nullptr;   // (n) sum over the pixels of every library row
    double npx = 0.0;                 // pixels per row
    double noise_tol = 0.0;           // `tol` of pca_annular
    int* ncomp_out = nullptr;         // (nprob) chosen number of components; negative: the rule wanted more than ncomp
};

// symmetric B x B eigenproblem in shared memory: one-sided Jacobi on the columns of Tm (PSD).
// On exit Tm's columns are mutually orthogonal: column j = theta_j * q_j.
template <int B>
__device__ void small_jacobi(double (*Tm)[B + 1], int nb, int tid) {
    // Round 2 (measured on the B200, tools/microbench/fp64_lat.cu: DFMA 9 cycles, but sqrt 100, divide 126, rsqrt 77
    // cycles of dependent latency): the first version of this routine took three warp reductions (both squared norms
    // and the inner product) and sqrt -> divide -> sqrt -> divide per rotation, ~900 cycles per round and ~80 % of an
    // iteration of the kernel below.  Now the squared column norms are kept up to date analytically, a rotation needs
    // ONE reduction and TWO rsqrt (1/h = rsqrt(d^2 + g^2), cos 2t = |d| / h, c^2 = (1 + cos 2t) / 2, s = sign(d) g /
    // (2 h c)), and a sweep whose rotations all started from couplings below 1e-8 is known to be the last one
    // (quadratic convergence) without a verification sweep.
    const int warp = tid >> 5, lane = tid & 31, nwarps = AT / 32;
    const int m = (nb + 1) & ~1;   // even number of players
    __shared__ int rotated;
    __shared__ double jn[B];
    const double thr = (double)B * 1.1e-16;
    for (int sweep = 0; sweep < 30; ++sweep) {
        if (tid == 0) rotated = 0;
        if (tid < nb) {
            double sq = 0.0;
            for (int i = 0; i < nb; ++i) sq = fma(Tm[i][tid], Tm[i][tid], sq);
            jn[tid] = sq;
        }
        __syncthreads();
        for (int r = 0; r < m - 1; ++r) {
            for (int pr = warp; pr < m / 2; pr += nwarps) {
                int a, b;
                const int mm = m - 1;
                if (pr == 0) { a = mm; b = r % mm; } else { a = (r + pr) % mm; b = (r - pr + mm) % mm; }
                if (a < nb && b < nb) {
                    double ga = 0;
                    for (int i = lane; i < nb; i += 32) ga = fma(Tm[i][a], Tm[i][b], ga);
                    ga = warp_sum(ga);
                    const double al = jn[a], be = jn[b];
                    if (al > 0 && be > 0 && ga * ga > thr * thr * al * be) {
                        const double d = be - al, g2 = 2.0 * ga;
                        const double rh = rsqrt(fma(d, d, g2 * g2));
                        const double c2 = fma(0.5 * fabs(d), rh, 0.5);
                        const double q = rsqrt(c2);
                        const double c = c2 * q;
                        const double s = copysign(0.5, d) * g2 * rh * q;
                        for (int i = lane; i < nb; i += 32) {
                            const double x = Tm[i][a], y = Tm[i][b];
                            Tm[i][a] = c * x - s * y;
                            Tm[i][b] = s * x + c * y;
                        }
                        __syncwarp();
                        if (lane == 0) {
                            const double cs2 = 2.0 * c * s * ga;
                            jn[a] = fma(c * c, al, fma(s * s, be, -cs2));
                            jn[b] = fma(s * s, al, fma(c * c, be, cs2));
                            if (ga * ga > 1e-16 * al * be) rotated = 1;
                        }
                    }
                }
            }
            __syncthreads();
        }
        const int any = rotated;
        __syncthreads();
        if (!any) break;
    }
}

template <int B>
__global__ void __launch_bounds__(AT)
annular_weights_kernel(AnnularArgs p) {
    extern __shared__ double sm[];
    const int q = blockIdx.x;
    const int tid = threadIdx.x;
    const int L = p.len[q];
    const int f = p.frame[q];
    const int n = p.n;
    const int* I = p.idx + (size_t)q * p.Lmax;
    if (L <= 0) { if (tid == 0) p.iters[q] = 0; return; }
    const int nb = (L < B) ? L : B;               // block width actually used
    const int k = (p.ncomp < nb) ? p.ncomp : nb;  // get_eigenvectors clamps ncomp to the library size

    // shared layout
    double* X = sm;                         // [L][B]
    double* Y = X + (size_t)p.Lmax * B;     // [L][B]
    double (*Tm)[B + 1] = reinterpret_cast<double (*)[B + 1]>(Y + (size_t)p.Lmax * B);   // [B][B+1]
    double (*Q)[B + 1] = Tm + B;            // [B][B+1]
    double* theta = reinterpret_cast<double*>(Q + B);   // [B]
    double* cnorm = theta + B;              // [B]  column norms^2 of Yr
    double* cres = cnorm + B;               // [B]  residual norms^2
    int* order = reinterpret_cast<int*>(cres + B);      // [B]
    int* Is = order + B;                    // [Lmax]
    __shared__ int done;
    __shared__ double hist;                 // worst residual four iterations ago (thread 0 only)

    for (int i = tid; i < L; i += AT) Is[i] = I[i];
    // start block: hashed pseudo-random entries (well conditioned), orthonormalised below
    for (int e = tid; e < L * B; e += AT) {
        const int i = e / B, c = e % B;
        Y[e] = (c < nb) ? hash_unit((unsigned)i * 131u + 7u, (unsigned)c * 977u + 3u) : 0.0;
    }
    if (tid < B) { cnorm[tid] = 0.0; }
    if (tid == 0) { done = 0; hist = 0.0; }
    __syncthreads();

    int it = 0;
    bool converged = false;
    bool have_ritz = false;
    while (true) {
        // ---- orthonormalise the columns of Y into X: column scaling + Cholesky QR ----------
        if (tid < B) cnorm[tid] = 0.0;
        __syncthreads();
        for (int c = 0; c < nb; ++c) {                 // all lanes take part in every shuffle
            double v = (tid < L) ? Y[tid * B + c] : 0.0;
            v = warp_sum(v * v);
            if ((tid & 31) == 0) atomicAdd(&cnorm[c], v);
        }
        __syncthreads();
        if (tid < B) cres[tid] = (tid < nb && cnorm[tid] > 0.0) ? rsqrt(cnorm[tid]) : 0.0;   // once per column
        __syncthreads();
        for (int e = tid; e < L * B; e += AT) Y[e] *= cres[e % B];
        __syncthreads();
        for (int e = tid; e < B * B; e += AT) {       // S = Yn^T Yn  (four independent chains: DFMA latency 9 cycles)
            const int a = e / B, b = e % B;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            if (a < nb && b < nb) {
                int i = 0;
                for (; i + 3 < L; i += 4) {
                    s0 = fma(Y[i * B + a], Y[i * B + b], s0);
                    s1 = fma(Y[(i + 1) * B + a], Y[(i + 1) * B + b], s1);
                    s2 = fma(Y[(i + 2) * B + a], Y[(i + 2) * B + b], s2);
                    s3 = fma(Y[(i + 3) * B + a], Y[(i + 3) * B + b], s3);
                }
                for (; i < L; ++i) s0 = fma(Y[i * B + a], Y[i * B + b], s0);
            }
            Tm[a][b] = (s0 + s1) + (s2 + s3);
            Q[a][b] = (a == b) ? 1.0 : 0.0;
        }
        __syncthreads();
        if (tid < 32) {
            // S = R^T R by elimination, with U (R^-1 before its column scaling) accumulated alongside: step j only
            // needs 1 / a_jj (one divide, lane-uniform); lane c owns column c of S (rows > j) and of U (rows <= j).
            // Replaces the left-looking factorisation (sqrt + divide per pivot) and the per-row triangular solve with
            // one divide per element (16 dependent divides per row: ~2000 cycles per iteration).
            const int c = tid;
            for (int j = 0; j < nb; ++j) {
                const double ajj = Tm[j][j];
                const double inv = (ajj > 1e-300) ? 1.0 / ajj : 1e300;
                if (c == 0) cnorm[j] = inv;                      // cnorm is free once the column scales are taken
                if (c < nb && c > j) {
                    const double sjc = Tm[j][c] * inv;
                    for (int i = j + 1; i <= c; ++i) Tm[i][c] = fma(-Tm[j][i], sjc, Tm[i][c]);
                    for (int i = 0; i <= j; ++i) Q[i][c] = fma(-Q[i][j], sjc, Q[i][c]);
                }
                __syncwarp();
            }
            if (tid < nb) cnorm[tid] = sqrt(cnorm[tid]);         // 1 / R[j][j]
            __syncwarp();
            for (int e = tid; e < B * B; e += 32) {              // R^-1 = U diag(sqrt(1 / a_jj)), upper triangular
                const int a = e / B, b = e % B;
                Q[a][b] = (a <= b && b < nb) ? Q[a][b] * cnorm[b] : 0.0;
            }
        }
        __syncthreads();
        if (tid < L) {                                 // x = yn R^-1: 16 independent chains, no divide
            double xr[B];
#pragma unroll
            for (int c = 0; c < B; ++c) xr[c] = 0.0;
            for (int m2 = 0; m2 < nb; ++m2) {
                const double v = Y[tid * B + m2];
#pragma unroll
                for (int c = 0; c < B; ++c) xr[c] = fma(v, Q[m2][c], xr[c]);
            }
#pragma unroll
            for (int c = 0; c < B; ++c) X[tid * B + c] = xr[c];
        }
        __syncthreads();
        if (converged || it >= p.max_iter) break;
        ++it;

        // ---- Y = G[I,I] X  (thread = row, coalesced over the mostly-contiguous library indices)
        if (tid < L) {
            double acc[B];
#pragma unroll
            for (int c = 0; c < B; ++c) acc[c] = 0.0;
            const int col = Is[tid];
            for (int j = 0; j < L; ++j) {
                const double g = __ldg(p.G + (size_t)Is[j] * n + col);
                const double* xj = X + j * B;
#pragma unroll
                for (int c = 0; c < B; ++c) acc[c] = fma(g, xj[c], acc[c]);
            }
#pragma unroll
            for (int c = 0; c < B; ++c) Y[tid * B + c] = acc[c];
        }
        __syncthreads();
        // ---- Rayleigh-Ritz: T = X^T Y (symmetrised), eig via one-sided Jacobi
        for (int e = tid; e < B * B; e += AT) {
            const int a = e / B, b = e % B;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            if (a < nb && b < nb) {
                int i = 0;
                for (; i + 3 < L; i += 4) {
                    s0 = fma(X[i * B + a], Y[i * B + b], s0);
                    s1 = fma(X[(i + 1) * B + a], Y[(i + 1) * B + b], s1);
                    s2 = fma(X[(i + 2) * B + a], Y[(i + 2) * B + b], s2);
                    s3 = fma(X[(i + 3) * B + a], Y[(i + 3) * B + b], s3);
                }
                for (; i < L; ++i) s0 = fma(X[i * B + a], Y[i * B + b], s0);
            }
            Q[a][b] = (s0 + s1) + (s2 + s3);
        }
        __syncthreads();
        for (int e = tid; e < B * B; e += AT) {
            const int a = e / B, b = e % B;
            Tm[a][b] = 0.5 * (Q[a][b] + Q[b][a]);
        }
        __syncthreads();
        small_jacobi<B>(Tm, nb, tid);
        if (tid < nb) {                                // theta_j = ||column j||
            double s = 0.0;
            for (int i = 0; i < nb; ++i) s = fma(Tm[i][tid], Tm[i][tid], s);
            theta[tid] = sqrt(s);
        }
        __syncthreads();
        if (tid < nb) {                                // rank by descending theta
            int rk = 0;
            for (int j = 0; j < nb; ++j) rk += (theta[j] > theta[tid] || (theta[j] == theta[tid] && j < tid));
            order[rk] = tid;
        }
        __syncthreads();
        for (int e = tid; e < B * B; e += AT) {        // Q[:, r] = normalised column order[r]
            const int i = e / B, r = e % B;
            if (i < nb && r < nb) {
                const int src = order[r];
                Q[i][r] = theta[src] > 0.0 ? Tm[i][src] / theta[src] : (i == src ? 1.0 : 0.0);
            }
        }
        __syncthreads();
        if (tid < B) { cres[tid] = 0.0; }
        __syncthreads();
        // ---- rotate: Yr = Y Q, Xr = X Q (each thread owns its row); residuals of the leading k pairs
        if (tid < L) {
            double* rows[2] = {Y + tid * B, X + tid * B};
            for (int which = 0; which < 2; ++which) {
                double out[B];
#pragma unroll
                for (int r = 0; r < B; ++r) out[r] = 0.0;
                for (int c = 0; c < nb; ++c) {
                    const double v = rows[which][c];
#pragma unroll
                    for (int r = 0; r < B; ++r) out[r] = fma(v, Q[c][r], out[r]);
                }
#pragma unroll
                for (int r = 0; r < B; ++r) rows[which][r] = out[r];
            }
        }
        for (int r = 0; r < k; ++r) {
            const double d = (tid < L) ? Y[tid * B + r] - theta[order[r]] * X[tid * B + r] : 0.0;
            const double v = warp_sum(d * d);
            if ((tid & 31) == 0) atomicAdd(&cres[r], v);
        }
        __syncthreads();
        have_ritz = true;
        if (tid == 0) {
            double worst = 0.0;
            for (int r = 0; r < k; ++r) worst = fmax(worst, sqrt(cres[r]));
            const double ref = theta[order[k - 1]];
            int dn = (worst <= p.tol * ref) || (nb == L);   // a full-width block is exact after one step
            // Hopeless cases leave early (BASELINE config 3: in 5 of 8 annuli the spectra are flat, every problem ran
            // all max_iter iterations -- 13 ms per annulus -- only to be solved again by the direct solver): the
            // progress over the last 4 iterations predicts the iterations still needed.
            if (!dn && (it & 3) == 0) {
                if (it >= 8 && hist > 0.0) {
                    const double r4 = worst / hist;
                    if (!(r4 < 1.0)) dn = 2;
                    else if (worst > 0.0 && ref > 0.0 &&
                             (double)it + 4.0 * log(p.tol * ref / worst) / log(r4) > (double)p.max_iter) dn = 2;
                }
                hist = worst;
            }
            done = dn;
        }
        __syncthreads();
        if (done == 2) {                                   // left to the direct solver
            if (tid == 0) p.iters[q] = -it;
            return;
        }
        converged = done != 0;
        if (converged || it >= p.max_iter) {
            // X already holds the Ritz vectors (orthonormal up to the residual); finish below
            break;
        }
    }
    (void)have_ritz;

    // ---- weights  w = sum_{j<k} x_j (x_j . g) / theta_j  with g_i = Gt[f][I_i]
    double* coef = cnorm;   // reuse
    if (tid < B) coef[tid] = 0.0;
    __syncthreads();
    {
        const double g = (tid < L) ? __ldg(p.Gt + (size_t)f * n + Is[tid]) : 0.0;
        for (int r = 0; r < k; ++r) {
            const double v = warp_sum((tid < L) ? X[tid * B + r] * g : 0.0);
            if ((tid & 31) == 0) atomicAdd(&coef[r], v);
        }
    }
    __syncthreads();
    if (tid < L) {
        double w = 0.0;
        for (int r = 0; r < k; ++r) {
            const double th = theta[order[r]];
            if (th > 0.0) w = fma(X[tid * B + r], coef[r] / th, w);
        }
        p.W[(size_t)q * n + Is[tid]] = (float)w;
    }
    if (tid == 0) p.iters[q] = converged ? it : -it;
}

// dst[i][c] = src[i][cols[c]]   (segment matrix from the flattened cube)
__global__ void gather_columns_kernel(const float* __restrict__ src, size_t p, const int* __restrict__ cols,
                                      int npx, float* __restrict__ dst) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (c < npx) dst[(size_t)i * npx + c] = src[(size_t)i * p + cols[c]];
}
// dst[i][cols[c]] = src[i][c]
__global__ void scatter_columns_kernel(const float* __restrict__ src, int npx, const int* __restrict__ cols,
                                       size_t p, float* __restrict__ dst) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (c < npx) dst[(size_t)i * p + cols[c]] = src[(size_t)i * npx + c];
}

size_t annular_smem_bytes(int B, int Lmax) {
    return ((size_t)2 * Lmax * B + (size_t)2 * B * (B + 1) + 3 * B) * sizeof(double) +
           ((size_t)B + Lmax) * sizeof(int) + 16;
}

template <int B>
static int launch_annular(const AnnularArgs& a, int nprob, cudaStream_t st) {
    const size_t smem = annular_smem_bytes(B, a.Lmax);
    VB_REQUIRE(smem <= 220 * 1024, "annular: library of %d frames with block %d needs %zu bytes of shared memory",
               a.Lmax, B, smem);
    VB_CHECK_CUDA(cudaFuncSetAttribute(annular_weights_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    annular_weights_kernel<B><<<nprob, AT, smem, st>>>(a);
    VB_CHECK_LAUNCH();
    return 0;
}

int annular_weights(const double* G, const double* Gt, int n, const int* idx, const int* len, const int* frame,
                    int nprob, int Lmax, int ncomp, double tol, int max_iter, float* W, int* iters,
                    cudaStream_t st) {
    VB_REQUIRE(nprob > 0 && n > 0 && Lmax > 0 && ncomp > 0, "annular: empty problem");
    VB_REQUIRE(Lmax <= AT, "annular: libraries larger than %d frames are not supported (max_frames_lib=%d)", AT, Lmax);
    AnnularArgs a{G, Gt ? Gt : G, n, idx, len, frame, Lmax, ncomp, tol > 0 ? tol : 1e-9,
                  max_iter > 0 ? max_iter : 500, W, iters};
    if (ncomp <= 10) return launch_annular<16>(a, nprob, st);
    VB_REQUIRE(ncomp <= 24, "annular: ncomp=%d too large for the batched eigensolver (max 24)", ncomp);
    return launch_annular<32>(a, nprob, st);
}

int gather_columns(const float* src, int n, size_t p, const int* cols, int npx, float* dst, cudaStream_t st) {
    gather_columns_kernel<<<dim3(ceil_div(npx, 256), n), 256, 0, st>>>(src, p, cols, npx, dst);
    VB_CHECK_LAUNCH();
    return 0;
}
int scatter_columns(const float* src, int n, int npx, const int* cols, size_t p, float* dst, cudaStream_t st) {
    scatter_columns_kernel<<<dim3(ceil_div(npx, 256), n), 256, 0, st>>>(src, npx, cols, p, dst);
    VB_CHECK_LAUNCH();
    return 0;
}

}  // namespace vb

// =====================================================================================
// Direct per-problem solver for (nearly) flat spectra
// =====================================================================================
// Subspace iteration converges at lambda_{B+1}/lambda_k per step; in noise-dominated annuli that
// ratio is ~0.97-0.99 and it stalls.  Problems it leaves unconverged are solved directly, one CTA
// each:  Householder tridiagonalisation of G[I,I] (workspace in global memory, L2 resident) ->
// multisection Sturm bisection for the k largest eigenvalues -> simultaneous inverse iteration on
// the tridiagonal (partial-pivoting solves, Gram-Schmidt between rounds) -> back-transformation ->
// projection weights.  Same outputs as annular_weights_kernel.

namespace vb {

__device__ __forceinline__ double block_sum_256(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < AT / 32; ++w) t += red[w];
    return t;
}

// fp64 reciprocal from the 20-bit hardware seed and two Newton steps (4 dependent DFMA): ~70 cycles of latency
// against 126 for the IEEE divide (tools/microbench/fp64_lat.cu); relative error ~1e-16 for NORMAL arguments, which
// is all the Sturm recurrence and the pivots below feed it (|q| >= pivmin >= the smallest normal number).
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// PACKED = true (round 2): the matrix lives in SHARED memory as its column-packed lower triangle (element (r, c),
// r >= c, at c L - c (c - 1) / 2 + r - c: L (L + 1) / 2 doubles, 161 KB at L = 200; the trailing matrix of every
// Householder step is again a packed triangle and reflector j -- column j below the diagonal -- is contiguous),
// the eigenvectors Z stay in shared memory and only the LU factors of the inverse iteration (written once, read once
// per round, off the dependency chain) go to the CTA's global slot.  With the full matrix in a global workspace
// (PACKED = false: libraries too large for shared memory) the Householder steps move 64 MB of trailing-matrix traffic
// per 200-frame problem through L2 / DRAM and the kernel idles on it (ncu, BASELINE config 3: long-scoreboard 20 and
// barrier 18 stalls per issue, 13 % issue-active, 27 ms per 1000 problems).
template <bool PACKED>
__device__ void annular_direct_one(const AnnularArgs& p, int q, double* __restrict__ ws, double* sm) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int L = p.len[q];
    const int f = p.frame[q];
    const int n = p.n, Lmax = p.Lmax;
    const int* I = p.idx + (size_t)q * Lmax;
    if (L <= 0) { if (tid == 0) p.iters[q] = 0; return; }
    const int k = (p.ncomp < L) ? p.ncomp : L;
    double* slot = ws + (size_t)blockIdx.x * Lmax * Lmax;  // global workspace slot of this CTA (Lmax^2 doubles)

    // shared layout (the five vectors hold at least 8 entries: vv / pp double as per-warp scratch of the Gershgorin
    // reduction -- with libraries of fewer than 8 frames they used to overrun into `red`)
    const int Lv = Lmax < 8 ? 8 : Lmax;
    double* d = sm;                       // [Lv]
    double* e = d + Lv;                   // [Lv]   e[j] couples j and j+1
    double* tau = e + Lv;                 // [Lv]
    double* vv = tau + Lv;                // [Lv]
    double* pp = vv + Lv;                 // [Lv]
    double* red = pp + Lv;                // [8]
    double* scal = red + 8;               // [8]  broadcast scalars
    double* lam = scal + 8;               // [32]
    double* blo = lam + 32;               // [32]
    double* bhi = blo + 32;               // [32]
    double* Z = bhi + 32;                 // [ncomp][Lmax]
    // PACKED: A = packed lower triangle in shared memory behind Z, U0..U2 in the global slot (3 k Lmax <= Lmax^2,
    //         checked by the launcher); else: A = full matrix in the global slot, U0..U2 in shared memory
    double* zend = Z + (size_t)p.ncomp * Lmax;
    double* A = PACKED ? zend : slot;
    double* U0 = PACKED ? slot : zend;
    double* U1 = U0 + (size_t)p.ncomp * Lmax;
    double* U2 = U1 + (size_t)p.ncomp * Lmax;
    int* cnt = reinterpret_cast<int*>(PACKED ? A + ((size_t)Lmax * (Lmax + 1) / 2) : U2 + (size_t)p.ncomp * Lmax);   // [AT]
    int* Is = cnt + AT;                   // [Lmax]

    for (int i = tid; i < L; i += AT) Is[i] = I[i];
    __syncthreads();
    if (PACKED) {
        // column c of the triangle = G[I[c..L-1], I[c]] (symmetric: read as the row I[c], ascending indices)
        for (int c = warp; c < L; c += AT / 32) {
            const double* grow = p.G + (size_t)Is[c] * n;
            double* col = A + (size_t)c * L - (size_t)c * (c - 1) / 2 - c;      // col[r] = A(r, c)
            for (int r = c + lane; r < L; r += 32) col[r] = __ldg(grow + Is[r]);
        }
    } else {
        for (int el = tid; el < L * L; el += AT) {
            const int i = el / L, j = el % L;
            A[el] = __ldg(p.G + (size_t)Is[i] * n + Is[j]);
        }
    }
    __syncthreads();

    // ---- Householder tridiagonalisation
    if (PACKED) {
        for (int j = 0; j < L - 1; ++j) {
            const int m = L - 1 - j;
            double* cj = A + (size_t)j * L - (size_t)j * (j - 1) / 2;      // column j: cj[0] = A(j, j), cj[1 + i] = x[i]
            double* x = cj + 1;
            double part = 0.0;
            for (int i = 1 + tid; i < m; i += AT) part = fma(x[i], x[i], part);
            const double sigma = block_sum_256(part, red);
            if (tid == 0) {
                const double alpha = x[0];
                d[j] = cj[0];
                if (sigma == 0.0) {
                    tau[j] = 0.0; e[j] = alpha; scal[0] = 0.0; scal[1] = 0.0;
                } else {
                    const double beta = -copysign(sqrt(alpha * alpha + sigma), alpha);
                    tau[j] = (beta - alpha) / beta;
                    e[j] = beta;
                    scal[0] = tau[j];
                    scal[1] = 1.0 / (alpha - beta);
                }
            }
            __syncthreads();
            const double tj = scal[0];
            if (tj == 0.0) {            // nothing to eliminate: v = e_1
                if (tid == 0) x[0] = 1.0;
                __syncthreads();
                continue;
            }
            const double scale = scal[1];
            for (int i = tid; i < m; i += AT) {
                const double v = (i == 0) ? 1.0 : x[i] * scale;
                vv[i] = v;
                x[i] = v;
            }
            __syncthreads();
            // trailing matrix B (m x m) = packed triangle behind column j: B(r, c) at c m - c (c - 1) / 2 + r - c
            double* B = cj + (m + 1);
            // p = tau * B v: thread i owns row i.  l <= i: B(i, l) sits i - l behind the head of column l (consecutive
            // threads, consecutive words); l > i: B(l, i) inside the thread's own column i.
            if (tid < m) {
                const int i = tid;
                const double* own = B + (size_t)i * m - (size_t)i * (i - 1) / 2 - i;    // own[l] = B(l, i), l >= i
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;   // four chains: a dependent DFMA costs 9 cycles
                int head = i;                       // index of B(i, l) for the running l <= i: cb(l) + i - l
                int l = 0;
                for (; l + 3 < m; l += 4) {
                    const double b0 = (l <= i) ? B[head] : own[l];
                    head += m - 1 - l;
                    const double b1 = (l + 1 <= i) ? B[head] : own[l + 1];
                    head += m - 2 - l;
                    const double b2 = (l + 2 <= i) ? B[head] : own[l + 2];
                    head += m - 3 - l;
                    const double b3 = (l + 3 <= i) ? B[head] : own[l + 3];
                    head += m - 4 - l;
                    a0 = fma(b0, vv[l], a0);
                    a1 = fma(b1, vv[l + 1], a1);
                    a2 = fma(b2, vv[l + 2], a2);
                    a3 = fma(b3, vv[l + 3], a3);
                }
                for (; l < m; ++l) {
                    a0 = fma((l <= i) ? B[head] : own[l], vv[l], a0);
                    head += m - 1 - l;
                }
                pp[i] = tj * ((a0 + a1) + (a2 + a3));
            }
            __syncthreads();
            double pv = (tid < m) ? pp[tid] * vv[tid] : 0.0;
            const double K = -0.5 * tj * block_sum_256(pv, red);
            if (tid < m) pp[tid] = fma(K, vv[tid], pp[tid]);      // w
            __syncthreads();
            // B -= v w^T + w v^T on the triangle, in 32 x 32 tiles dealt round-robin to the warps (row-indexed
            // lanes: conflict-free; the tiles balance the triangular work over the warps)
            {
                const int nrb = (m + 31) >> 5;
                int item = 0;
                for (int rb = 0; rb < nrb; ++rb) {
                    const int r = (rb << 5) + lane;
                    for (int cbk = 0; cbk <= rb; ++cbk, ++item) {
                        if ((item & (AT / 32 - 1)) != warp || r >= m) continue;
                        const double wr = pp[r], vr = vv[r];
                        const int c0 = cbk << 5;
                        const int c1 = min(c0 + 31, r);
                        double* el = B + (size_t)c0 * m - (size_t)c0 * (c0 - 1) / 2 + (r - c0);
                        for (int c = c0; c <= c1; ++c) {
                            *el -= fma(vv[c], wr, pp[c] * vr);
                            el += m - 1 - c;
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (tid == 0) { d[L - 1] = A[(size_t)(L - 1) * L - (size_t)(L - 1) * (L - 2) / 2]; e[L - 1] = 0.0; tau[L - 1] = 0.0; }
        __syncthreads();
    } else {
        // ---- Householder tridiagonalisation (both triangles kept up to date; reflector j stored in row j)
        for (int j = 0; j < L - 1; ++j) {
            const int m = L - 1 - j;
            double* x = A + (size_t)j * L + (j + 1);
            double part = 0.0;
            for (int i = 1 + tid; i < m; i += AT) part = fma(x[i], x[i], part);
            const double sigma = block_sum_256(part, red);
            if (tid == 0) {
                const double alpha = x[0];
                d[j] = A[(size_t)j * L + j];
                if (sigma == 0.0) {
                    tau[j] = 0.0; e[j] = alpha; scal[0] = 0.0; scal[1] = 0.0;
                } else {
                    const double beta = -copysign(sqrt(alpha * alpha + sigma), alpha);
                    tau[j] = (beta - alpha) / beta;
                    e[j] = beta;
                    scal[0] = tau[j];
                    scal[1] = 1.0 / (alpha - beta);
                }
            }
            __syncthreads();
            const double tj = scal[0];
            if (tj == 0.0) {            // nothing to eliminate: v = e_1
                if (tid == 0) x[0] = 1.0;
                __syncthreads();
                continue;
            }
            const double scale = scal[1];
            for (int i = tid; i < m; i += AT) {
                const double v = (i == 0) ? 1.0 : x[i] * scale;
                vv[i] = v;
                x[i] = v;
            }
            __syncthreads();
            // p = tau * A22 v   (A22 symmetric: column i of A22 read as row-major rows l, coalesced over i)
            const double* A22 = A + (size_t)(j + 1) * L + (j + 1);
            if (tid < m) {
                // four independent chains: a dependent DFMA costs 9 cycles on this part (fp64_lat.cu)
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                int l = 0;
                for (; l + 3 < m; l += 4) {
                    a0 = fma(A22[(size_t)l * L + tid], vv[l], a0);
                    a1 = fma(A22[(size_t)(l + 1) * L + tid], vv[l + 1], a1);
                    a2 = fma(A22[(size_t)(l + 2) * L + tid], vv[l + 2], a2);
                    a3 = fma(A22[(size_t)(l + 3) * L + tid], vv[l + 3], a3);
                }
                for (; l < m; ++l) a0 = fma(A22[(size_t)l * L + tid], vv[l], a0);
                pp[tid] = tj * ((a0 + a1) + (a2 + a3));
            }
            __syncthreads();
            double pv = (tid < m) ? pp[tid] * vv[tid] : 0.0;
            const double K = -0.5 * tj * block_sum_256(pv, red);
            if (tid < m) pp[tid] = fma(K, vv[tid], pp[tid]);      // w
            __syncthreads();
            if (tid < m) {
                const double wi = pp[tid], vi = vv[tid];
                double* col = A + (size_t)(j + 1) * L + (j + 1) + tid;
    #pragma unroll 8
                for (int l = 0; l < m; ++l) col[(size_t)l * L] -= vv[l] * wi + pp[l] * vi;
            }
            __syncthreads();
        }
        if (tid == 0) { d[L - 1] = A[(size_t)(L - 1) * L + (L - 1)]; e[L - 1] = 0.0; tau[L - 1] = 0.0; }
        __syncthreads();
    }

    // ---- k largest eigenvalues of the tridiagonal: multisection with Sturm counts
    {
        double gl = 1e300, gh = -1e300, e2max = 0.0;
        for (int i = tid; i < L; i += AT) {
            const double r = ((i > 0) ? fabs(e[i - 1]) : 0.0) + ((i < L - 1) ? fabs(e[i]) : 0.0);
            gl = fmin(gl, d[i] - r); gh = fmax(gh, d[i] + r);
            if (i < L - 1) e2max = fmax(e2max, e[i] * e[i]);
        }
        for (int o = 16; o > 0; o >>= 1) {
            gl = fmin(gl, __shfl_xor_sync(0xffffffffu, gl, o));
            gh = fmax(gh, __shfl_xor_sync(0xffffffffu, gh, o));
            e2max = fmax(e2max, __shfl_xor_sync(0xffffffffu, e2max, o));
        }
        __syncthreads();
        if (lane == 0) { red[warp] = gl; pp[warp] = gh; vv[warp] = e2max; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < AT / 32; ++w) { gl = fmin(gl, red[w]); gh = fmax(gh, pp[w]); e2max = fmax(e2max, vv[w]); }
            const double pivmin = 2.2250738585072014e-308 * fmax(1.0, e2max);
            const double span = gh - gl;
            scal[2] = gl - (1e-3 * span + pivmin);
            scal[3] = gh + (1e-3 * span + pivmin);
            scal[4] = pivmin;
        }
        __syncthreads();
    }
    const double pivmin = scal[4];
    const int P = AT / k;                 // section points per eigenvalue and round
    if (tid < k) { blo[tid] = scal[2]; bhi[tid] = scal[3]; }
    __syncthreads();
    for (int round = 0; round < 64; ++round) {
        const int r = tid / P, s = tid % P;
        int c = 0;
        double xs = 0.0;
        const bool active = r < k;
        if (active) {
            const double lo = blo[r], hi = bhi[r];
            xs = lo + (double)(s + 1) / (double)(P + 1) * (hi - lo);
            double qv = d[0] - xs;
            if (fabs(qv) < pivmin) qv = -pivmin;
            c = qv < 0.0;
            for (int i = 1; i < L; ++i) {
                qv = fma(-e[i - 1] * e[i - 1], fast_rcp(qv), d[i] - xs);
                if (fabs(qv) < pivmin) qv = -pivmin;
                c += qv < 0.0;
            }
        }
        cnt[tid] = c;
        __syncthreads();
        // per eigenvalue: new interval from the P counts (thread r)
        __shared__ int all_done;
        if (tid == 0) all_done = 1;
        __syncthreads();
        if (tid < k) {
            const int a = L - 1 - tid;            // ascending index of the wanted eigenvalue
            double lo = blo[tid], hi = bhi[tid];
            const double w0 = hi - lo;
            double nlo = lo, nhi = hi;
            for (int s2 = 0; s2 < P; ++s2) {
                const double x2 = lo + (double)(s2 + 1) / (double)(P + 1) * w0;
                const int c2 = cnt[tid * P + s2];
                if (c2 <= a) nlo = x2;
                else { nhi = x2; break; }
            }
            blo[tid] = nlo; bhi[tid] = nhi;
            const bool shrunk = (nlo > lo) || (nhi < hi);     // no progress = interval is down to a few ulps
            if (shrunk && nhi - nlo > 4.0 * 2.220446049250313e-16 * fmax(fabs(nlo), fabs(nhi)) + 2.0 * pivmin)
                all_done = 0;
        }
        __syncthreads();
        if (all_done) break;
    }
    if (tid < k) lam[tid] = 0.5 * (blo[tid] + bhi[tid]);
    __syncthreads();

    // ---- simultaneous inverse iteration on the tridiagonal
    for (int el = tid; el < k * L; el += AT) {
        const int r = el / L, i = el % L;
        Z[(size_t)r * Lmax + i] = hash_unit((unsigned)i * 131u + 7u, (unsigned)r * 977u + 3u);
    }
    __syncthreads();
    for (int round = 0; round < 3; ++round) {
        if (tid < k) {
            const int r = tid;
            // factors of row i: UF[(3 i + {0, 1, 2}) K + r] = {1 / pivot, first, second super-diagonal}: the k active
            // threads write and read consecutive words; the reads of the back substitution do not depend on its
            // chain and are fetched PF rows ahead (PACKED: they come from the global slot)
            const int K = p.ncomp;
            double* UF = U0 + r;
            double* b = Z + (size_t)r * Lmax;
            const double lr = lam[r];
            // Gaussian elimination with partial pivoting of (T - lr I); sub-diagonal of row i+1 is e[i]
            double diag = d[0] - lr;              // current pivot-row candidates
            double sup = (L > 1) ? e[0] : 0.0;
            double bi = b[0];
            for (int i = 0; i < L - 1; ++i) {
                const double bn = b[i + 1];
                const double sub = e[i];
                const double nd = d[i + 1] - lr;                    // next row: [sub, nd, e[i+1]]
                const double ns = (i + 2 < L) ? e[i + 1] : 0.0;
                double* uf = UF + (size_t)3 * i * K;
                if (fabs(diag) >= fabs(sub) || fabs(sub) < 1e-300) {
                    const double piv = (fabs(diag) >= 1e-300) ? diag : copysign(1e-300, diag);
                    const double rinv = fast_rcp(piv);
                    const double mlt = sub * rinv;
                    uf[0] = rinv; uf[K] = sup; uf[2 * K] = 0.0;
                    b[i] = bi;
                    bi = fma(-mlt, bi, bn);
                    diag = fma(-mlt, sup, nd);
                    sup = ns;
                } else {
                    const double rinv = fast_rcp(sub);              // |sub| > |diag| >= 0: a normal number here
                    const double mlt = diag * rinv;
                    uf[0] = rinv; uf[K] = nd; uf[2 * K] = ns;
                    b[i] = bn;
                    bi = fma(-mlt, bn, bi);
                    diag = fma(-mlt, nd, sup);
                    sup = -mlt * ns;
                }
            }
            {
                double* uf = UF + (size_t)3 * (L - 1) * K;
                const double piv = (fabs(diag) >= 1e-300) ? diag : copysign(1e-300, diag);
                uf[0] = fast_rcp(piv); uf[K] = 0.0; uf[2 * K] = 0.0;
                b[L - 1] = bi;
            }
            // back substitution  b[i] = (b[i] - u1[i] b[i+1] - u2[i] b[i+2]) / u0[i], running maximum for the scaling
            constexpr int PF = 4;
            double q0[PF], q1[PF], q2[PF];
#pragma unroll
            for (int t = 0; t < PF; ++t) {
                const int i = L - 1 - t;
                const double* uf = UF + (size_t)3 * (i < 0 ? 0 : i) * K;
                q0[t] = uf[0]; q1[t] = uf[K]; q2[t] = uf[2 * K];
            }
            double x1 = 0.0, x2 = 0.0, mx = 0.0;
            for (int i0 = L - 1; i0 >= 0; i0 -= PF) {
                double n0[PF], n1[PF], n2[PF];
#pragma unroll
                for (int t = 0; t < PF; ++t) {
                    const int i = i0 - PF - t;
                    const double* uf = UF + (size_t)3 * (i < 0 ? 0 : i) * K;
                    n0[t] = uf[0]; n1[t] = uf[K]; n2[t] = uf[2 * K];
                }
#pragma unroll
                for (int t = 0; t < PF; ++t) {
                    const int i = i0 - t;
                    if (i >= 0) {
                        const double xi = fma(-q2[t], x2, fma(-q1[t], x1, b[i])) * q0[t];
                        b[i] = xi;
                        mx = fmax(mx, fabs(xi));
                        x2 = x1; x1 = xi;
                    }
                    q0[t] = n0[t]; q1[t] = n1[t]; q2[t] = n2[t];
                }
            }
            // scale to avoid overflow in the next round
            const double inv = (mx > 0.0) ? 1.0 / mx : 1.0;
            for (int i = 0; i < L; ++i) b[i] *= inv;
        }
        __syncthreads();
        if (warp == 0) {                  // modified Gram-Schmidt over the k vectors (one warp, no block syncs)
            for (int r = 0; r < k; ++r) {
                double* zr = Z + (size_t)r * Lmax;
                for (int s2 = 0; s2 < r; ++s2) {
                    const double* zs = Z + (size_t)s2 * Lmax;
                    double dt = 0.0;
                    for (int i = lane; i < L; i += 32) dt = fma(zr[i], zs[i], dt);
                    dt = warp_sum(dt);
                    for (int i = lane; i < L; i += 32) zr[i] -= dt * zs[i];
                    __syncwarp();
                }
                double nn = 0.0;
                for (int i = lane; i < L; i += 32) nn = fma(zr[i], zr[i], nn);
                nn = warp_sum(nn);
                const double inv = (nn > 0.0) ? rsqrt(nn) : 0.0;
                for (int i = lane; i < L; i += 32) zr[i] *= inv;
                __syncwarp();
            }
        }
        __syncthreads();
    }

    // ---- back-transformation  x = H_0 H_1 ... z  (each warp owns whole vectors: no block syncs)
    for (int r = warp; r < k; r += AT / 32) {
        double* z = Z + (size_t)r * Lmax;
        for (int j = L - 2; j >= 0; --j) {
            const double tj = tau[j];
            if (tj == 0.0) continue;
            const int m = L - 1 - j;
            const double* v = PACKED ? A + (size_t)j * L - (size_t)j * (j - 1) / 2 + 1 : A + (size_t)j * L + (j + 1);
            double dt = 0.0;
            for (int i = lane; i < m; i += 32) dt = fma(v[i], z[j + 1 + i], dt);
            dt = warp_sum(dt) * tj;
            for (int i = lane; i < m; i += 32) z[j + 1 + i] -= dt * v[i];
            __syncwarp();
        }
    }
    __syncthreads();

    // ---- ncomp='auto': noise-decay rule of get_eigenvectors (svd.py:640-672) on the library itself.
    // Residual of the library D (L x npx) after m components = sum_{i>m} sigma_i u_i v_i^T, so
    //   mean(res^2) = (trace(G_II) - sum_{i<=m} lambda_i) / (L npx),
    //   mean(res)   = (sum_i r_i - sum_{i<=m} (1.x_i)(x_i.r)) / (L npx)      with r = D 1 (row sums),
    // and np.std(res) = sqrt(mean(res^2) - mean(res)^2): only the leading eigenpairs are needed.
    int kuse = k;
    if (p.rowsum != nullptr) {
        double* ca = blo;                 // 1 . x_r
        double* cb = bhi;                 // x_r . r
        if (tid < 32) { ca[tid] = 0.0; cb[tid] = 0.0; }
        __syncthreads();
        const double rs = (tid < L) ? __ldg(p.rowsum + Is[tid]) : 0.0;
        for (int r = 0; r < k; ++r) {
            const double z = (tid < L) ? Z[(size_t)r * Lmax + tid] : 0.0;
            const double va = warp_sum(z), vb2 = warp_sum(z * rs);
            if (lane == 0) { atomicAdd(&ca[r], va); atomicAdd(&cb[r], vb2); }
        }
        const double srs = block_sum_256(rs, red);
        __syncthreads();
        if (tid == 0) {
            double trace = 0.0;
            for (int i = 0; i < L; ++i) trace += d[i];
            const double tot = (double)L * p.npx;
            const int max_evs = ((double)L < p.npx) ? L : (int)p.npx;
            double s2 = trace, sr = srs, prev = 0.0, decay = 1.0;
            int m = 0, clipped = 0;
            while (decay >= p.noise_tol) {
                ++m;
                if (m <= max_evs) {
                    if (m > k) { clipped = 1; break; }         // the rule wants more eigenpairs than were computed
                    s2 -= lam[m - 1];
                    sr -= ca[m - 1] * cb[m - 1];
                }
                const double mean = sr / tot;
                const double var = s2 / tot - mean * mean;
                const double noise = (var > 0.0) ? sqrt(var) : 0.0;
                if (m > 1) decay = prev - noise;
                prev = noise;
                if (m > max_evs + 1) break;                     // tol <= 0: the reference would not terminate
            }
            if (m > max_evs) m = max_evs;                       // V_big[:ncomp] holds max_evs rows at most
            if (clipped) m = k;
            reinterpret_cast<int*>(scal)[0] = m;
            p.ncomp_out[q] = clipped ? -m : m;
        }
        __syncthreads();
        kuse = reinterpret_cast<int*>(scal)[0];
        __syncthreads();
    }

    // ---- weights  w = sum_r x_r (x_r . g) / lam_r
    double* coef = blo;                   // (bisection intervals are no longer needed)
    if (tid < 32) coef[tid] = 0.0;
    __syncthreads();
    {
        const double g = (tid < L) ? __ldg(p.Gt + (size_t)f * n + Is[tid]) : 0.0;
        for (int r = 0; r < kuse; ++r) {
            const double v = warp_sum((tid < L) ? Z[(size_t)r * Lmax + tid] * g : 0.0);
            if (lane == 0) atomicAdd(&coef[r], v);
        }
    }
    __syncthreads();
    if (tid < L) {
        double w = 0.0;
        for (int r = 0; r < kuse; ++r)
            if (lam[r] > 0.0) w = fma(Z[(size_t)r * Lmax + tid], coef[r] / lam[r], w);
        p.W[(size_t)q * n + Is[tid]] = (float)w;
    }
    if (tid == 0) p.iters[q] = 100000;      // marker: solved directly
}

// Persistent launch: the CTAs work through the problem list, each with ONE workspace slot (Lmax^2 doubles), so the
// workspace is bounded whatever the number of problems.  PACKED = false (matrix in the global slot) is bound by the
// LATENCY of that workspace (ncu, BASELINE config 3: 64 MB of trailing-matrix traffic per 200-frame problem, long
// scoreboard 20 and barrier 18 stalls per issue, 13 % issue-active; two CTAs per SM -- whose 95 MB of workspaces would
// stay L2-resident -- ran 25 % SLOWER than three, i.e. concurrency, not DRAM bandwidth, is what it lacks);
// PACKED = true keeps the matrix in shared memory (one CTA per SM at Lmax = 200, more for smaller libraries).
template <bool PACKED>
__global__ void __launch_bounds__(AT, PACKED ? 1 : 3)
annular_direct_kernel(AnnularArgs p, const int* __restrict__ plist, int nlist, double* __restrict__ ws) {
    extern __shared__ double sm[];
    for (int i = blockIdx.x; i < nlist; i += gridDim.x) {
        annular_direct_one<PACKED>(p, plist[i], ws, sm);
        __syncthreads();
    }
}

int annular_direct_slots() { return 3 * kNumSMs; }      // upper bound of the workspace slots (= CTAs) of one launch

size_t annular_direct_smem_bytes(int k, int Lmax) {
    const int Lv = Lmax < 8 ? 8 : Lmax;
    return ((size_t)5 * Lv + 16 + 96 + (size_t)4 * k * Lmax) * sizeof(double) + ((size_t)AT + Lmax) * sizeof(int) + 16;
}
size_t annular_direct_packed_smem_bytes(int k, int Lmax) {
    const int Lv = Lmax < 8 ? 8 : Lmax;
    return ((size_t)5 * Lv + 16 + 96 + (size_t)k * Lmax + (size_t)Lmax * (Lmax + 1) / 2) * sizeof(double) +
           ((size_t)AT + Lmax) * sizeof(int) + 16;
}

int annular_direct(const AnnularArgs& a, const int* plist, int nlist, double* ws, cudaStream_t st);

int annular_direct_weights(const double* G, const double* Gt, int n, const int* idx, const int* len,
                           const int* frame, int nprob, int Lmax, int ncomp, const int* plist, int nlist, float* W,
                           int* iters, double* ws, cudaStream_t st) {
    VB_REQUIRE(nlist > 0 && nlist <= nprob, "annular_direct: bad problem list");
    AnnularArgs a{G, Gt ? Gt : G, n, idx, len, frame, Lmax, ncomp, 0.0, 0, W, iters};
    return annular_direct(a, plist, nlist, ws, st);
}

// ncomp='auto': every listed problem solved directly with `kmax` eigenpairs, the number used chosen per problem by
// the noise-decay rule (see the kernel).  ncomp_out[q] < 0: the rule asked for more than kmax components.
int annular_auto_weights(const double* G, const double* Gt, int n, const int* idx, const int* len, const int* frame,
                         int nprob, int Lmax, int kmax, const double* rowsum, double npx, double noise_tol,
                         const int* plist, int nlist, float* W, int* iters, int* ncomp_out, double* ws,
                         cudaStream_t st) {
    VB_REQUIRE(nlist > 0 && nlist <= nprob, "annular_auto: bad problem list");
    VB_REQUIRE(rowsum != nullptr && ncomp_out != nullptr && npx >= 1.0, "annular_auto: missing row sums");
    VB_REQUIRE(noise_tol > 0.0, "annular_auto: `tol` must be positive");
    AnnularArgs a{G, Gt ? Gt : G, n, idx, len, frame, Lmax, kmax, 0.0, 0, W, iters};
    a.rowsum = rowsum; a.npx = npx; a.noise_tol = noise_tol; a.ncomp_out = ncomp_out;
    return annular_direct(a, plist, nlist, ws, st);
}

// Solve the problems listed in plist (device, nlist entries) directly.  ws: min(nlist, annular_direct_slots()) * Lmax^2
// doubles.
static int direct_mode() {       // VIP_B200_ANNULAR_DIRECT: 0 = matrix in the global slot (round 1), 1 = packed (default)
    static const int v = [] { const char* e = getenv("VIP_B200_ANNULAR_DIRECT"); return e ? atoi(e) : 1; }();
    return v;
}

int annular_direct(const AnnularArgs& a, const int* plist, int nlist, double* ws, cudaStream_t st) {
    VB_REQUIRE(a.ncomp <= 24 && a.ncomp >= 1, "annular_direct: ncomp must be in 1..24");
    VB_REQUIRE(a.Lmax <= AT, "annular_direct: library too large");
    const size_t smem_p = annular_direct_packed_smem_bytes(a.ncomp, a.Lmax);
    // packed: shared memory must hold the triangle and the slot the 3 k Lmax LU factors
    const bool packed = direct_mode() != 0 && smem_p <= 220 * 1024 && 3 * a.ncomp <= a.Lmax;
    const size_t smem = packed ? smem_p : annular_direct_smem_bytes(a.ncomp, a.Lmax);
    VB_REQUIRE(smem <= 220 * 1024, "annular_direct: %zu bytes of shared memory needed", smem);
    auto kern = packed ? annular_direct_kernel<true> : annular_direct_kernel<false>;
    VB_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    VB_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, AT, smem));
    per_sm = per_sm < 1 ? 1 : (per_sm > 3 ? 3 : per_sm);
    const int slots = per_sm * kNumSMs;
    const int grid = nlist < slots ? nlist : slots;
    kern<<<grid, AT, smem, st>>>(a, plist, nlist, ws);
    VB_CHECK_LAUNCH();
    return 0;
}

}  // namespace vb
