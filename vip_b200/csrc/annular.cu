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
};

// symmetric B x B eigenproblem in shared memory: one-sided Jacobi on the columns of Tm (PSD).
// On exit Tm's columns are mutually orthogonal: column j = theta_j * q_j.
template <int B>
__device__ void small_jacobi(double (*Tm)[B + 1], int nb, int tid) {
    const int warp = tid >> 5, lane = tid & 31, nwarps = AT / 32;
    const int m = (nb + 1) & ~1;   // even number of players
    for (int sweep = 0; sweep < 30; ++sweep) {
        __shared__ int rotated;
        if (tid == 0) rotated = 0;
        __syncthreads();
        for (int r = 0; r < m - 1; ++r) {
            for (int pr = warp; pr < m / 2; pr += nwarps) {
                int a, b;
                const int mm = m - 1;
                if (pr == 0) { a = mm; b = r % mm; } else { a = (r + pr) % mm; b = (r - pr + mm) % mm; }
                if (a < nb && b < nb) {
                    double al = 0, be = 0, ga = 0;
                    for (int i = lane; i < nb; i += 32) {
                        const double x = Tm[i][a], y = Tm[i][b];
                        al = fma(x, x, al); be = fma(y, y, be); ga = fma(x, y, ga);
                    }
                    al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
                    if (al > 0 && be > 0 && fabs(ga) > 1e-15 * sqrt(al) * sqrt(be)) {
                        const double zeta = (be - al) / (2.0 * ga);
                        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                        const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                        for (int i = lane; i < nb; i += 32) {
                            const double x = Tm[i][a], y = Tm[i][b];
                            Tm[i][a] = c * x - s * y;
                            Tm[i][b] = s * x + c * y;
                        }
                        if (lane == 0) rotated = 1;
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

    for (int i = tid; i < L; i += AT) Is[i] = I[i];
    // start block: hashed pseudo-random entries (well conditioned), orthonormalised below
    for (int e = tid; e < L * B; e += AT) {
        const int i = e / B, c = e % B;
        Y[e] = (c < nb) ? hash_unit((unsigned)i * 131u + 7u, (unsigned)c * 977u + 3u) : 0.0;
    }
    if (tid < B) { cnorm[tid] = 0.0; }
    if (tid == 0) done = 0;
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
        for (int e = tid; e < L * B; e += AT) {
            const int c = e % B;
            Y[e] = (c < nb && cnorm[c] > 0.0) ? Y[e] * rsqrt(cnorm[c]) : 0.0;
        }
        __syncthreads();
        for (int e = tid; e < B * B; e += AT) {       // S = Yn^T Yn
            const int a = e / B, b = e % B;
            double s = 0.0;
            if (a < nb && b < nb)
                for (int i = 0; i < L; ++i) s = fma(Y[i * B + a], Y[i * B + b], s);
            Tm[a][b] = s;
        }
        __syncthreads();
        if (tid < 32) {                                // Cholesky S = R^T R (R upper, stored in Q), warp 0
            for (int j = 0; j < nb; ++j) {
                double d = Tm[j][j];
                for (int m2 = 0; m2 < j; ++m2) d -= Q[m2][j] * Q[m2][j];
                d = (d > 1e-300) ? sqrt(d) : 1e-150;
                for (int c = j + tid; c < nb; c += 32) {
                    double v = Tm[j][c];
                    for (int m2 = 0; m2 < j; ++m2) v -= Q[m2][j] * Q[m2][c];
                    Q[j][c] = (c == j) ? d : v / d;
                }
                __syncwarp();
            }
        }
        __syncthreads();
        if (tid < L) {                                 // x^T R = yn^T  (forward substitution per row)
            double xr[B];
#pragma unroll
            for (int c = 0; c < B; ++c) {
                if (c < nb) {
                    double v = Y[tid * B + c];
                    for (int m2 = 0; m2 < c; ++m2) v -= xr[m2] * Q[m2][c];
                    xr[c] = v / Q[c][c];
                } else xr[c] = 0.0;
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
            double s = 0.0;
            if (a < nb && b < nb)
                for (int i = 0; i < L; ++i) s = fma(X[i * B + a], Y[i * B + b], s);
            Q[a][b] = s;
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
            done = (worst <= p.tol * ref) || (nb == L);   // a full-width block is exact after one step
        }
        __syncthreads();
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
