// Symmetric eigensolver for the n x n Gramian in fp64: block one-sided (Hestenes) Jacobi.
//
// Role in the reference: the LAPACK call inside numpy.linalg.svd / eigh
// (src/vip_hci/psfsub/svd.py:450, 470), which numpy always runs in fp64.
//
// One-sided Jacobi on W = G (symmetric PSD): plane rotations applied to column pairs until
// all columns are mutually orthogonal; then W = E diag(lambda), i.e. lambda_j = ||w_j|| and
// e_j = w_j / lambda_j.  Rotations touch two columns only, so disjoint pairs run in parallel.
// Columns are grouped in blocks of WB; a round-robin tournament over block pairs gives
// (NB-1) rounds per sweep, one kernel launch per round, one CTA per block pair.  Each CTA
// stages its 2*WB columns in shared memory and orthogonalises all pairs among them (one warp
// per pair, WB disjoint pairs at a time).  Sweeps stop early through a device-side flag.
#include "common.cuh"
#include <cstdlib>
#include <cstdio>

namespace vb {

struct JacobiState {
    unsigned int rotations;   // rotations applied in the current sweep
    unsigned int converged;   // set when a sweep applied no rotation
    unsigned int sweeps;      // sweeps actually executed
    unsigned int pad;
};

// circle-method pairing: players 0..m-1 (m even), round r in [0, m-1), slot i in [0, m/2)
__device__ __forceinline__ void rr_pair(int m, int r, int i, int& a, int& b) {
    const int mm = m - 1;
    if (i == 0) { a = mm; b = r % mm; }
    else { a = (r + i) % mm; b = (r - i + mm) % mm; }
}

template <int WB>
__global__ void __launch_bounds__(32 * WB)
jacobi_round_kernel(double* __restrict__ W, int n, int nblocks, int round, double tol,
                    JacobiState* __restrict__ state) {
    if (state->converged) return;
    extern __shared__ double cols[];   // [2*WB][n]
    int bi, bj;
    rr_pair(nblocks, round, blockIdx.x, bi, bj);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // global column index of local column c (c < WB: block bi, else block bj); >= n means padding
    auto gcol = [&](int c) { return (c < WB ? bi * WB + c : bj * WB + (c - WB)); };

    // each warp streams whole columns with 8 independent loads in flight per lane
    for (int c = warp; c < 2 * WB; c += WB) {
        const int gc = gcol(c);
        if (gc < n) {
            const double* src = W + (size_t)gc * n;
            double* dst = cols + (size_t)c * n;
#pragma unroll 8
            for (int i = lane; i < n; i += 32) dst[i] = __ldcg(src + i);
        }
    }
    __syncthreads();

    unsigned int nrot = 0;
    for (int r = 0; r < 2 * WB - 1; ++r) {
        int ca, cb;
        rr_pair(2 * WB, r, warp, ca, cb);
        if (gcol(ca) < n && gcol(cb) < n) {
            double* x = cols + (size_t)ca * n;
            double* y = cols + (size_t)cb * n;
            double alpha = 0.0, beta = 0.0, gamma = 0.0, a2 = 0.0, b2 = 0.0, g2 = 0.0;
            int i = lane;
            for (; i + 32 < n; i += 64) {      // two independent accumulator chains
                const double xv = x[i], yv = y[i], xw = x[i + 32], yw = y[i + 32];
                alpha = fma(xv, xv, alpha); beta = fma(yv, yv, beta); gamma = fma(xv, yv, gamma);
                a2 = fma(xw, xw, a2); b2 = fma(yw, yw, b2); g2 = fma(xw, yw, g2);
            }
            if (i < n) {
                const double xv = x[i], yv = y[i];
                alpha = fma(xv, xv, alpha); beta = fma(yv, yv, beta); gamma = fma(xv, yv, gamma);
            }
            alpha = warp_sum(alpha + a2); beta = warp_sum(beta + b2); gamma = warp_sum(gamma + g2);
            if (alpha > 0.0 && beta > 0.0 && fabs(gamma) > tol * sqrt(alpha) * sqrt(beta)) {
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t);
                const double s = c * t;
#pragma unroll 4
                for (int j = lane; j < n; j += 32) {
                    const double xv = x[j], yv = y[j];
                    x[j] = c * xv - s * yv;
                    y[j] = s * xv + c * yv;
                }
                ++nrot;
            }
        }
        __syncthreads();
    }

    for (int c = warp; c < 2 * WB; c += WB) {
        const int gc = gcol(c);
        if (gc < n) {
            double* dst = W + (size_t)gc * n;
            const double* src = cols + (size_t)c * n;
#pragma unroll 8
            for (int i = lane; i < n; i += 32) __stcg(dst + i, src[i]);
        }
    }
    if (lane == 0 && nrot) atomicAdd(&state->rotations, nrot);
}

__global__ void jacobi_sweep_end_kernel(JacobiState* state) {
    if (state->converged) return;
    state->sweeps += 1;
    if (state->rotations == 0) state->converged = 1;
    state->rotations = 0;
}

// lambda_j = ||w_j||, rank by descending lambda, evecs[rank] = w_j / lambda_j (row-major rows)
__global__ void __launch_bounds__(128)
jacobi_norms_kernel(const double* __restrict__ W, int n, double* __restrict__ norms) {
    const int j = blockIdx.x;
    const double* w = W + (size_t)j * n;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s = fma(w[i], w[i], s);
    __shared__ double red[4];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) norms[j] = sqrt(red[0] + red[1] + red[2] + red[3]);
}

__global__ void __launch_bounds__(128)
jacobi_sort_kernel(const double* __restrict__ W, int n, const double* __restrict__ norms,
                   double* __restrict__ evals, double* __restrict__ evecs) {
    const int j = blockIdx.x;
    const double lj = norms[j];
    __shared__ int rank_s;
    if (threadIdx.x == 0) rank_s = 0;
    __syncthreads();
    int cnt = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double li = norms[i];
        cnt += (li > lj || (li == lj && i < j)) ? 1 : 0;
    }
    atomicAdd(&rank_s, cnt);
    __syncthreads();
    const int rank = rank_s;
    if (threadIdx.x == 0) evals[rank] = lj;
    const double inv = lj > 0.0 ? 1.0 / lj : 0.0;
    const double* w = W + (size_t)j * n;
    double* dst = evecs + (size_t)rank * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = w[i] * inv;
}

// Whole eigensolver in ONE launch for small matrices (n <= 128; the (ncomp+10)^2 problems of the randomized SVD
// and its orthonormalisation, small libraries): the matrix lives in shared memory, the 32 warps of a single CTA
// take the disjoint column pairs of a round, and sweeps repeat in-kernel until one applies no rotation.  The
// per-round launches of the block kernel above cost ~450 launches (4 ms) for a 60 x 60 matrix -- 12 % of a
// config-5 step (profiles/r01m_launches_c5.md).  Same rotations, same stopping rule, then norms / ranks / output.
__global__ void __launch_bounds__(1024, 1)
jacobi_small_kernel(const double* __restrict__ G, int n, int max_sweeps, double tol, double* __restrict__ evals,
                    double* __restrict__ evecs, JacobiState* __restrict__ state) {
    extern __shared__ double cols[];            // [n][n]: column c at cols + c*n (G is symmetric)
    __shared__ double norms[128];
    __shared__ unsigned int nrot_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
    for (int e = tid; e < n * n; e += blockDim.x) cols[e] = G[e];
    const int m = (n + 1) & ~1;                 // players of the tournament; an odd n gets one idle slot
    if (tid == 0) nrot_s = 0;
    __syncthreads();
    unsigned int sweeps = 0, conv = (n < 2) ? 1u : 0u;
    for (int sweep = 0; sweep < max_sweeps && !conv; ++sweep) {
        unsigned int nrot = 0;
        for (int r = 0; r < m - 1; ++r) {
            for (int pr = warp; pr < m / 2; pr += nwarps) {
                int ca, cb;
                rr_pair(m, r, pr, ca, cb);
                if (ca < n && cb < n) {
                    double* x = cols + (size_t)ca * n;
                    double* y = cols + (size_t)cb * n;
                    double alpha = 0.0, beta = 0.0, gamma = 0.0;
                    for (int i = lane; i < n; i += 32) {
                        const double xv = x[i], yv = y[i];
                        alpha = fma(xv, xv, alpha); beta = fma(yv, yv, beta); gamma = fma(xv, yv, gamma);
                    }
                    alpha = warp_sum(alpha); beta = warp_sum(beta); gamma = warp_sum(gamma);
                    if (alpha > 0.0 && beta > 0.0 && fabs(gamma) > tol * sqrt(alpha) * sqrt(beta)) {
                        const double zeta = (beta - alpha) / (2.0 * gamma);
                        const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                        const double c = 1.0 / sqrt(1.0 + t * t);
                        const double s = c * t;
                        for (int j = lane; j < n; j += 32) {
                            const double xv = x[j], yv = y[j];
                            x[j] = c * xv - s * yv;
                            y[j] = s * xv + c * yv;
                        }
                        ++nrot;
                    }
                }
            }
            __syncthreads();
        }
        if (lane == 0 && nrot) atomicAdd(&nrot_s, nrot);
        __syncthreads();
        const unsigned int any = nrot_s;
        __syncthreads();
        if (tid == 0) nrot_s = 0;
        ++sweeps;
        if (!any) conv = 1;
    }
    __syncthreads();
    for (int c = warp; c < n; c += nwarps) {
        double sq = 0.0;
        for (int i = lane; i < n; i += 32) sq = fma(cols[(size_t)c * n + i], cols[(size_t)c * n + i], sq);
        sq = warp_sum(sq);
        if (lane == 0) norms[c] = sqrt(sq);
    }
    __syncthreads();
    for (int j = warp; j < n; j += nwarps) {
        const double lj = norms[j];
        int cnt = 0;
        for (int i = lane; i < n; i += 32) {
            const double li = norms[i];
            cnt += (li > lj || (li == lj && i < j)) ? 1 : 0;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) evals[cnt] = lj;
        const double inv = lj > 0.0 ? 1.0 / lj : 0.0;
        for (int i = lane; i < n; i += 32) evecs[(size_t)cnt * n + i] = cols[(size_t)j * n + i] * inv;
    }
    if (tid == 0) {
        state->sweeps = sweeps;
        state->converged = conv;
        state->rotations = 0;
    }
}

size_t eigh_workspace_bytes(int n) {
    return (size_t)n * n * sizeof(double) + (size_t)n * sizeof(double) + 256;
}

template <int WB>
static int jacobi_sweeps(double* W, int n, int max_sweeps, double tol, JacobiState* state, int* launches,
                         cudaStream_t st) {
    int nblocks = ceil_div(n, WB);
    if (nblocks & 1) ++nblocks;
    if (nblocks < 2) nblocks = 2;
    const size_t smem = (size_t)2 * WB * n * sizeof(double);
    VB_REQUIRE(smem <= 220 * 1024, "eigh: n=%d too large for the shared-memory column blocks", n);
    VB_CHECK_CUDA(cudaFuncSetAttribute(jacobi_round_kernel<WB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    int nl = 0;
    for (int sweep = 0; sweep < max_sweeps; ++sweep) {
        for (int r = 0; r < nblocks - 1; ++r) {
            jacobi_round_kernel<WB><<<nblocks / 2, 32 * WB, smem, st>>>(W, n, nblocks, r, tol, state);
            ++nl;
        }
        VB_CHECK_LAUNCH();
        jacobi_sweep_end_kernel<<<1, 1, 0, st>>>(state);
        ++nl;
        // Jacobi needs ~6-10 sweeps; from the 5th on, poll the device flag (one 16-byte D2H + sync
        // per sweep) instead of queueing dozens of no-op launches.
        if (sweep >= 4) {
            JacobiState h;
            VB_CHECK_CUDA(cudaMemcpyAsync(&h, state, sizeof(h), cudaMemcpyDeviceToHost, st));
            VB_CHECK_CUDA(cudaStreamSynchronize(st));
            if (h.converged) break;
        }
    }
    VB_CHECK_LAUNCH();
    *launches += nl;
    return 0;
}

// G: n x n symmetric fp64 (not modified).  evals[n] descending, evecs[n x n] row j = eigenvector j.
// info (host, optional): [0] sweeps executed, [1] converged flag  -- reading it synchronises the stream.
// ---------------------------------------------------------------------------------------------------------------
// Cholesky whitening of a small SPD Gramian:  G = R^T R  (R upper triangular),  Wt = R^-T  (lower triangular), so that
// Wt G Wt^T = I: the rows of Wt Y are orthonormal when G = Y Y^T (CholeskyQR).  One CTA, the matrix in shared memory,
// right-looking factorisation (one barrier pair per column), then one thread per column of the inverse.  Replaces the
// eigen-decomposition in the orthonormalisation passes of the randomized SVD, where a 60 x 60 one-sided Jacobi cost
// 1.5 ms per call (7 calls = the fixed 10 ms that did not shrink with the number of GPUs, profiles/r02k_launches_c5.md).
// A pivot below 1e-30 of the largest diagonal entry marks a numerically dependent row: its row/column of Wt is zeroed
// (that direction is dropped), as the eigen-based whitening did for null eigenvalues.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
chol_whiten_kernel(const double* __restrict__ G, int n, double* __restrict__ Wt) {
    extern __shared__ double cw_smem[];
    double* A = cw_smem;                       // n x (n + 1), upper triangle becomes R
    __shared__ double dmax_s;
    __shared__ int dead[128];
    const int ld = n + 1, tid = threadIdx.x, nt = blockDim.x;
    for (int e = tid; e < n * n; e += nt) {
        const int i = e / n, j = e % n;
        A[i * ld + j] = 0.5 * (G[i * n + j] + G[j * n + i]);
    }
    if (tid < 128) dead[tid] = 0;
    __syncthreads();
    if (tid == 0) {
        double m = 0.0;
        for (int i = 0; i < n; ++i) m = fmax(m, A[i * ld + i]);
        dmax_s = m;
    }
    __syncthreads();
    const double thr = dmax_s * 1e-30;
    for (int j = 0; j < n; ++j) {
        const double ajj = A[j * ld + j];
        const bool ok = ajj > thr && ajj > 0.0;
        const double rd = ok ? rsqrt(ajj) : 0.0;
        __syncthreads();                       // everyone has read the pivot before row j is scaled
        // row j of R:  R[j][c] = A[j][c] / sqrt(ajj)
        for (int c = j + tid; c < n; c += nt) A[j * ld + c] = ok ? A[j * ld + c] * rd : 0.0;
        if (tid == 0) { dead[j] = ok ? 0 : 1; if (!ok) A[j * ld + j] = 1.0; }
        __syncthreads();
        // trailing update: A[i][c] -= R[j][i] R[j][c]  for j < i <= c
        const int m = n - j - 1;
        for (int e = tid; e < m * m; e += nt) {
            const int i = j + 1 + e / m, c = j + 1 + e % m;
            if (c >= i) A[i * ld + c] = fma(-A[j * ld + i], A[j * ld + c], A[i * ld + c]);
        }
    }
    __syncthreads();
    // Wt = R^-T: column q of R^-1 solves R x = e_q (back substitution); Wt[q][i]... we need Wt = (R^-1)^T, i.e.
    // Wt[r][c] = (R^-1)[c][r].  Thread q computes x = R^-1 e_q (non-zero for rows <= q) and writes Wt[q][0..q].
    for (int q = tid; q < n; q += nt) {
        // x_q = 1 / R[q][q]; for i = q-1 .. 0: x_i = -(sum_{m=i+1..q} R[i][m] x_m) / R[i][i]
        double* x = Wt + (size_t)q * n;        // row q of Wt used as the work vector
        for (int i = 0; i < n; ++i) x[i] = 0.0;
        if (!dead[q]) {
            x[q] = 1.0 / A[q * ld + q];
            for (int i = q - 1; i >= 0; --i) {
                double sacc = 0.0;
                for (int mm = i + 1; mm <= q; ++mm) sacc = fma(A[i * ld + mm], x[mm], sacc);
                x[i] = dead[i] ? 0.0 : -sacc / A[i * ld + i];
            }
        }
    }
}

// Wt (n x n fp64, lower triangular) with Wt G Wt^T = I for the SPD matrix G (n <= 128).  Asynchronous on `st`.
int chol_whiten_f64(const double* G, int n, double* Wt, cudaStream_t st) {
    VB_REQUIRE(n >= 1 && n <= 128, "chol_whiten: 1 <= n <= 128");
    const size_t smem = (size_t)n * (n + 1) * sizeof(double);
    static bool attr = false;
    if (!attr) {
        VB_CHECK_CUDA(cudaFuncSetAttribute(chol_whiten_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           128 * 129 * (int)sizeof(double)));
        attr = true;
    }
    chol_whiten_kernel<<<1, 256, smem, st>>>(G, n, Wt);
    VB_CHECK_LAUNCH();
    return 0;
}

int eigh_f64(const double* G, int n, double* evals, double* evecs, int max_sweeps, double tol, void* ws,
             size_t ws_bytes, int* info, int* launches, cudaStream_t st) {
    VB_REQUIRE(n >= 1, "eigh: n must be >= 1");
    VB_REQUIRE(ws_bytes >= eigh_workspace_bytes(n), "eigh: workspace too small");
    if (max_sweeps <= 0) max_sweeps = 30;
    if (tol <= 0) tol = 8.0 * sqrt((double)n) * 1.1102230246251565e-16;
    double* W = reinterpret_cast<double*>(ws);
    double* norms = W + (size_t)n * n;
    JacobiState* state = reinterpret_cast<JacobiState*>(norms + n);
    VB_CHECK_CUDA(cudaMemcpyAsync(W, G, (size_t)n * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    VB_CHECK_CUDA(cudaMemsetAsync(state, 0, sizeof(JacobiState), st));
    int nl = 0;
    static const int use_small = [] { const char* e = getenv("VIP_B200_JACOBI_SMALL"); return e ? atoi(e) : 1; }();
    if (use_small && n <= 128) {
        // single-CTA solver: reads G directly, writes evals/evecs, records {sweeps, converged} in `state`
        const size_t smem = (size_t)n * n * sizeof(double);
        static bool attr = false;
        if (!attr) {
            VB_CHECK_CUDA(cudaFuncSetAttribute(jacobi_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               128 * 128 * (int)sizeof(double)));
            attr = true;
        }
        const int threads = (n <= 32) ? 512 : 1024;
        jacobi_small_kernel<<<1, threads, smem, st>>>(G, n, max_sweeps, tol, evals, evecs, state);
        VB_CHECK_LAUNCH();
        if (launches) *launches = 1;
        if (info) {
            JacobiState h;
            VB_CHECK_CUDA(cudaMemcpyAsync(&h, state, sizeof(h), cudaMemcpyDeviceToHost, st));
            VB_CHECK_CUDA(cudaStreamSynchronize(st));
            info[0] = (int)h.sweeps;
            info[1] = (int)h.converged;
        }
        return 0;
    }
    if (n > 1) {
        int rc;
        // wider blocks = fewer launches; narrower = more CTAs per round and less shared memory
        if (n <= 128) rc = jacobi_sweeps<2>(W, n, max_sweeps, tol, state, &nl, st);
        else if ((size_t)8 * n * sizeof(double) <= 200 * 1024 && n <= 1024)
            rc = jacobi_sweeps<4>(W, n, max_sweeps, tol, state, &nl, st);
        else rc = jacobi_sweeps<2>(W, n, max_sweeps, tol, state, &nl, st);
        if (rc) return rc;
    }
    jacobi_norms_kernel<<<n, 128, 0, st>>>(W, n, norms);
    VB_CHECK_LAUNCH();
    jacobi_sort_kernel<<<n, 128, 0, st>>>(W, n, norms, evals, evecs);
    VB_CHECK_LAUNCH();
    nl += 2;
    if (launches) *launches = nl;
    if (info) {
        JacobiState h;
        VB_CHECK_CUDA(cudaMemcpyAsync(&h, state, sizeof(h), cudaMemcpyDeviceToHost, st));
        VB_CHECK_CUDA(cudaStreamSynchronize(st));
        info[0] = (int)h.sweeps;
        info[1] = (int)h.converged;
    }
    return 0;
}

}  // namespace vb

// =====================================================================================
// Top-k eigenpairs of a large symmetric PSD matrix: fp64 block subspace iteration
// =====================================================================================
// For integer ncomp << n the PCA needs only the leading invariant subspace of G.  Orthogonal
// iteration with Rayleigh-Ritz on a block of B > k vectors costs O(n^2 B) per step (one skinny
// fp64 GEMM) and converges at rate lambda_{B+1}/lambda_k; every step is 6 small launches:
//   Y = G X  ->  T = X^T Y  ->  Ritz (Jacobi on T)  ->  rotate + S = Yr^T Yr + residuals
//            ->  Cholesky/convergence  ->  X = Yr R^-1.
// State lives in global memory (n x B doubles, L2 resident), so any n works.

namespace vb {

struct TopkState {
    int iters;
    int converged;
    double worst;     // max residual / theta_k at the last check
    double prev_worst;   // adaptive Ritz schedule (fused kernel, rr_every = 0): previous check and its iteration
    int prev_it;
    int next_rr;         // iteration index of the next Rayleigh-Ritz step
};


__device__ __forceinline__ double hash_unit2(unsigned int a, unsigned int b) {
    unsigned long long z = ((unsigned long long)a << 32 | b) + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

template <int B>
__global__ void topk_init_kernel(double* __restrict__ Y, int n) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n * B) Y[e] = hash_unit2((unsigned)(e / B) * 131u + 7u, (unsigned)(e % B) * 977u + 3u);
}

// Y[r][:] = sum_j G[r][j] X[j][:]   (CTA: 256/B rows x B columns, one output per thread; n/(256/B) CTAs
// keep ~60 SMs busy on the 500 x 500 problem instead of 16)
template <int B>
__global__ void __launch_bounds__(256)
topk_matvec_kernel(const double* __restrict__ G, int n, const double* __restrict__ X, double* __restrict__ Y,
                   const TopkState* __restrict__ st) {
    if (st->converged) return;
    constexpr int TJ = (B <= 32) ? 128 : 64, RPC = 256 / B;     // 64-wide blocks: 32 KB of X per step
    __shared__ double Gs[RPC][TJ + 1];
    __shared__ double Xs[TJ][B];
    const int r0 = blockIdx.x * RPC;
    const int tr = threadIdx.x / B, tc = threadIdx.x % B;
    double acc0 = 0.0, acc1 = 0.0;
    for (int j0 = 0; j0 < n; j0 += TJ) {
        for (int e = threadIdx.x; e < RPC * TJ; e += 256) {
            const int r = e / TJ, j = e % TJ;
            Gs[r][j] = (r0 + r < n && j0 + j < n) ? G[(size_t)(r0 + r) * n + j0 + j] : 0.0;
        }
        for (int e = threadIdx.x; e < TJ * B; e += 256) {
            const int j = e / B;
            Xs[j][e % B] = (j0 + j < n) ? X[(size_t)(j0 + j) * B + e % B] : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int j = 0; j < TJ; j += 2) {
            acc0 = fma(Gs[tr][j], Xs[j][tc], acc0);
            acc1 = fma(Gs[tr][j + 1], Xs[j + 1][tc], acc1);
        }
        __syncthreads();
    }
    if (r0 + tr < n) Y[(size_t)(r0 + tr) * B + tc] = acc0 + acc1;
}

// T (B x B, atomics; must be zeroed) += P^T Q over this CTA's rows
template <int B>
__global__ void __launch_bounds__(256)
topk_xty_kernel(const double* __restrict__ P, const double* __restrict__ Q, int n, double* __restrict__ T,
                const TopkState* __restrict__ st) {
    if (st->converged) return;
    constexpr int RCH = 16;
    __shared__ double Ps[RCH][B], Qs[RCH][B];
    const int r0 = blockIdx.x * RCH;
    for (int e = threadIdx.x; e < RCH * B; e += 256) {
        const int r = e / B;
        const bool in = r0 + r < n;
        Ps[r][e % B] = in ? P[(size_t)(r0 + r) * B + e % B] : 0.0;
        Qs[r][e % B] = in ? Q[(size_t)(r0 + r) * B + e % B] : 0.0;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < B * B; e += 256) {
        const int a = e / B, b = e % B;
        double s = 0.0;
#pragma unroll 8
        for (int r = 0; r < RCH; ++r) s = fma(Ps[r][a], Qs[r][b], s);
        atomicAdd(&T[e], s);
    }
}

// single CTA: symmetrise T, one-sided Jacobi -> Qm (B x B, columns = Ritz vectors in X coordinates,
// descending theta), theta[B]; zeroes S, res, T for the next kernels
template <int B>
__global__ void __launch_bounds__(16 * B)
topk_ritz_kernel(double* __restrict__ T, double* __restrict__ Qm, double* __restrict__ theta,
                 double* __restrict__ S, double* __restrict__ res, const TopkState* __restrict__ st,
                 double jthr) {
    if (st->converged) return;
    __shared__ double Tm[B][B + 1];
    __shared__ double th[B];
    __shared__ int order[B];
    __shared__ int rotated;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < B * B; e += 16 * B) {
        const int a = e / B, b = e % B;
        Tm[a][b] = 0.5 * (T[a * B + b] + T[b * B + a]);
    }
    __syncthreads();
    for (int sweep = 0; sweep < 40; ++sweep) {
        if (tid == 0) rotated = 0;
        __syncthreads();
        for (int r = 0; r < B - 1; ++r) {
            {
                const int pr = warp;          // one column pair per warp: B/2 warps
                int a, b;
                const int mm = B - 1;
                if (pr == 0) { a = mm; b = r % mm; } else { a = (r + pr) % mm; b = (r - pr + mm) % mm; }
                double al = 0, be = 0, ga = 0;
                for (int i = lane; i < B; i += 32) {
                    const double x = Tm[i][a], y = Tm[i][b];
                    al = fma(x, x, al); be = fma(y, y, be); ga = fma(x, y, ga);
                }
                al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
                if (al > 0 && be > 0 && fabs(ga) > jthr * sqrt(al) * sqrt(be)) {
                    const double zeta = (be - al) / (2.0 * ga);
                    const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                    const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                    for (int i = lane; i < B; i += 32) {
                        const double x = Tm[i][a], y = Tm[i][b];
                        Tm[i][a] = c * x - s * y;
                        Tm[i][b] = s * x + c * y;
                    }
                    if (lane == 0) rotated = 1;
                }
            }
            __syncthreads();
        }
        const int any = rotated;
        __syncthreads();
        if (!any) break;
    }
    if (tid < B) {
        double s = 0.0;
        for (int i = 0; i < B; ++i) s = fma(Tm[i][tid], Tm[i][tid], s);
        th[tid] = sqrt(s);
    }
    __syncthreads();
    if (tid < B) {
        int rk = 0;
        for (int j = 0; j < B; ++j) rk += (th[j] > th[tid] || (th[j] == th[tid] && j < tid));
        order[rk] = tid;
    }
    __syncthreads();
    for (int e = tid; e < B * B; e += 16 * B) {
        const int i = e / B, r = e % B;
        const int src = order[r];
        Qm[e] = th[src] > 0.0 ? Tm[i][src] / th[src] : (i == src ? 1.0 : 0.0);
        S[e] = 0.0;
        T[e] = 0.0;
    }
    if (tid < B) { theta[tid] = th[order[tid]]; res[tid] = 0.0; }
}

// rows per CTA of the rotation kernel: static shared memory holds Q (B x B) and two row slabs
__host__ __device__ constexpr int topk_rot_rows(int B) { return B <= 32 ? 64 : 8; }

// rows: Yr = Y Q, Xr = X Q (in place); S += Yr^T Yr; res[r] += ||Yr[:,r] - theta_r Xr[:,r]||^2
template <int B>
__global__ void __launch_bounds__(256)
topk_rotate_kernel(double* __restrict__ X, double* __restrict__ Y, int n, const double* __restrict__ Qm,
                   const double* __restrict__ theta, double* __restrict__ S, double* __restrict__ res,
                   const TopkState* __restrict__ st) {
    if (st->converged) return;
    constexpr int RCH = topk_rot_rows(B);
    __shared__ double Qs[B][B];
    __shared__ double Ys[RCH][B], Xs[RCH][B];
    const int tid = threadIdx.x, r0 = blockIdx.x * RCH;
    for (int e = tid; e < B * B; e += 256) Qs[e / B][e % B] = Qm[e];
    for (int e = tid; e < RCH * B; e += 256) {
        const int r = e / B;
        const bool in = r0 + r < n;
        Ys[r][e % B] = in ? Y[(size_t)(r0 + r) * B + e % B] : 0.0;
        Xs[r][e % B] = in ? X[(size_t)(r0 + r) * B + e % B] : 0.0;
    }
    __syncthreads();
    // each thread produces entries (row, col) of Yr and Xr: RCH*B entries / 256 threads
    double yr[RCH * B / 256], xr[RCH * B / 256];
#pragma unroll
    for (int q = 0; q < RCH * B / 256; ++q) {
        const int e = tid + q * 256, r = e / B, c = e % B;
        double sy = 0.0, sx = 0.0;
#pragma unroll 8
        for (int m = 0; m < B; ++m) {
            sy = fma(Ys[r][m], Qs[m][c], sy);
            sx = fma(Xs[r][m], Qs[m][c], sx);
        }
        yr[q] = sy; xr[q] = sx;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < RCH * B / 256; ++q) {
        const int e = tid + q * 256, r = e / B, c = e % B;
        Ys[r][c] = yr[q]; Xs[r][c] = xr[q];
        if (r0 + r < n) {
            Y[(size_t)(r0 + r) * B + c] = yr[q];
            X[(size_t)(r0 + r) * B + c] = xr[q];
        }
    }
    __syncthreads();
    for (int e = tid; e < B * B; e += 256) {
        const int a = e / B, b = e % B;
        double s = 0.0;
#pragma unroll 8
        for (int r = 0; r < RCH; ++r) s = fma(Ys[r][a], Ys[r][b], s);
        atomicAdd(&S[e], s);
    }
    if (tid < B) {
        double s = 0.0;
        const double th = theta[tid];
        for (int r = 0; r < RCH; ++r) {
            const double d = Ys[r][tid] - th * Xs[r][tid];
            s = fma(d, d, s);
        }
        atomicAdd(&res[tid], s);
    }
}

// single warp: convergence test on the leading k Ritz pairs; R = chol(D^-1 S D^-1) (upper), dinv = 1/||Yr_c||
template <int B>
__global__ void topk_chol_kernel(double* __restrict__ S, double* __restrict__ res, double* __restrict__ T,
                                 const double* __restrict__ theta, int k, double tol, double* __restrict__ R,
                                 double* __restrict__ dinv, TopkState* __restrict__ st, int check) {
    if (st->converged) return;
    __shared__ double Sn[B][B + 1], dv[B];       // factorised IN PLACE (upper triangle becomes R): one B x B matrix
    const int lane = threadIdx.x;
    if (lane == 0) {
        st->iters += 1;
        if (check) {
            double worst = 0.0;
            for (int r = 0; r < k; ++r) worst = fmax(worst, sqrt(res[r]));
            const double ref = theta[k - 1];
            st->worst = ref > 0.0 ? worst / ref : 0.0;
            if (worst <= tol * ref) st->converged = 1;
        }
    }
    __syncwarp();
    for (int c = lane; c < B; c += 32) dv[c] = S[c * B + c] > 0.0 ? rsqrt(S[c * B + c]) : 0.0;
    __syncwarp();
    for (int e = lane; e < B * B; e += 32) Sn[e / B][e % B] = S[e] * dv[e / B] * dv[e % B];
    __syncwarp();
    for (int j = 0; j < B; ++j) {
        // rows m < j of Sn already hold R; row j still holds the (normalised) Gram entries
        double d = Sn[j][j];
        for (int m = 0; m < j; ++m) d -= Sn[m][j] * Sn[m][j];
        d = (d > 1e-300) ? sqrt(d) : 1e-150;
        __syncwarp();
        for (int c = j + lane; c < B; c += 32) {
            double v = Sn[j][c];
            for (int m = 0; m < j; ++m) v -= Sn[m][j] * Sn[m][c];
            Sn[j][c] = (c == j) ? d : v / d;
        }
        __syncwarp();
    }
    for (int e = lane; e < B * B; e += 32) R[e] = (e / B <= e % B) ? Sn[e / B][e % B] : 0.0;
    for (int c = lane; c < B; c += 32) dinv[c] = dv[c];
    // leave the atomic accumulators clean for the next iteration (RR or plain orthogonalisation)
    for (int e = lane; e < B * B; e += 32) { S[e] = 0.0; T[e] = 0.0; }
    for (int c = lane; c < B; c += 32) res[c] = 0.0;
}

// rows: X = (Yr D^-1) R^-1  (forward substitution per row).  When converged, X (Ritz vectors) is kept.
template <int B>
__global__ void __launch_bounds__(128)
topk_solve_kernel(double* __restrict__ X, const double* __restrict__ Y, int n, const double* __restrict__ R,
                  const double* __restrict__ dinv, const TopkState* __restrict__ st) {
    if (st->converged) return;
    __shared__ double Rs[B][B + 1], ds[B];
    for (int e = threadIdx.x; e < B * B; e += 128) Rs[e / B][e % B] = R[e];
    for (int c = threadIdx.x; c < B; c += 128) ds[c] = dinv[c];
    __syncthreads();
    const int r = blockIdx.x * 128 + threadIdx.x;
    if (r >= n) return;
    double x[B];
#pragma unroll
    for (int c = 0; c < B; ++c) {
        double v = Y[(size_t)r * B + c] * ds[c];
#pragma unroll
        for (int m = 0; m < c; ++m) v -= x[m] * Rs[m][c];
        x[c] = v / Rs[c][c];
    }
#pragma unroll
    for (int c = 0; c < B; ++c) X[(size_t)r * B + c] = x[c];
}

// evals[k], evecs[k x n] (row j = j-th eigenvector) from theta / X
template <int B>
__global__ void topk_output_kernel(const double* __restrict__ X, const double* __restrict__ theta, int n, int k,
                                   double* __restrict__ evals, double* __restrict__ evecs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        for (int j = 0; j < k; ++j) evecs[(size_t)j * n + i] = X[(size_t)i * B + j];
    if (i < k) evals[i] = theta[i];
}

// =====================================================================================
// Fused persistent variant of the subspace iteration (one cooperative launch for all iterations)
// =====================================================================================
// The six kernels above are latency-bound (n x 32 blocks on a 500 x 500 matrix): a launch per phase
// costs more than the phase.  Here ceil(n / RPC) co-resident CTAs keep their RPC rows of X and Y in
// shared memory across the phases of an iteration and meet at a grid barrier between phases:
//   A  Y_rows = G_rows X  (+ partial X^T Y or Y^T Y)      all CTAs
//   B  Rayleigh-Ritz (Jacobi on the B x B matrix)          CTA 0, RR iterations only
//   C  rotate own rows, partial Gram of the rotated block  all CTAs, RR iterations only
//   D  Cholesky of the (column-normalised) Gram, convergence test     CTA 0, warp 0
//   E  own rows  X = Y D^-1 R^-1                           all CTAs
// Rayleigh-Ritz runs every `rr_every`-th iteration; the iteration before it orthonormalises twice
// (CholeskyQR2) so that the Ritz step sees an orthonormal X (a single Cholesky-QR of G X leaves
// 1e-4 of non-orthogonality when lambda_0 / lambda_B ~ 1e6, which stalls the residual test).

__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        const unsigned int target = (epoch + 1u) * gridDim.x;
        const long long t0 = clock64();
        while (*((volatile unsigned int*)bar) < target) {
            if (clock64() - t0 > 4000000000LL) __trap();      // never hang the GPU on a protocol bug
        }
        __threadfence();
    }
    ++epoch;
    __syncthreads();
}

template <int B>
struct FusedSmem {
    static constexpr int RPC = 512 / B, TJ = 64;
    struct PhaseA { double Gs[RPC][TJ + 1]; double Xt[TJ][B]; };
    struct PhaseB { double Tm[B][B + 1]; double th[B]; int order[B]; int rotated; };
    struct PhaseD { double Sn[B][B + 1]; double Rm[B][B + 1]; };
    union U { PhaseA a; PhaseB b; PhaseD d; };
};

template <int B>
__global__ void __launch_bounds__(512, 1)
topk_fused_kernel(const double* __restrict__ G, int n, int k, double tol, int max_iter, double jthr, int rr_every,
                  int chol_mode, int rr0, double* X, double* T, double* S, double* Qm, double* R, double* theta, double* res, double* dinv,
                  TopkState* st, unsigned int* bar) {
    using FS = FusedSmem<B>;
    constexpr int RPC = FS::RPC, TJ = FS::TJ;
    __shared__ typename FS::U u;
    __shared__ double Ys[RPC][B + 1], Xm[RPC][B + 1], Qs[B][B + 1], dv[B], rq[B];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tr = tid / B, tc = tid % B;
    const int r0 = blockIdx.x * RPC, r = r0 + tr;
    unsigned int epoch = 0;

    // partial Gram  out[a][b] += sum_rows P[row][a] Q[row][b]  (atomics into a zeroed B x B accumulator)
    auto partial_gram = [&](const double (*P)[B + 1], const double (*Q)[B + 1], double* out) {
        for (int e = tid; e < B * B; e += 512) {
            const int a = e / B, b = e % B;
            double s = 0.0;
#pragma unroll 8
            for (int q = 0; q < RPC; ++q) s = fma(P[q][a], Q[q][b], s);
            atomicAdd(&out[e], s);
        }
    };
    // phase D on CTA 0 / warp 0
    auto cholesky_phase = [&](int check, int count_iter, int it_now) {
        if (blockIdx.x == 0 && warp == 0) {
            if (lane == 0) {
                if (count_iter) st->iters += 1;
                if (check) {
                    double worst = 0.0;
                    for (int q = 0; q < k; ++q) worst = fmax(worst, sqrt(__ldcg(&res[q])));
                    const double ref = __ldcg(&theta[k - 1]);
                    const double rho = ref > 0.0 ? worst / ref : 0.0;
                    if (worst <= tol * ref) st->converged = 1;
                    else if (rr_every == 0) {
                        // adaptive schedule: convergence is geometric, so two checks give the rate and the number of
                        // iterations still needed; the next Ritz step is placed there (2..8 ahead: the step before
                        // it must be the CholeskyQR2 one) instead of every 4th iteration
                        int step = 4;
                        const double prho = st->prev_worst;
                        const int pit = st->prev_it;
                        double lr = 0.0;                                   // log of the rate per iteration (< 0)
                        if (pit > 0 && prho > 0.0 && rho > 0.0 && rho < prho) {
                            lr = log(rho / prho) / (double)(it_now - pit);
                        } else {
                            // first check: subspace iteration converges like lambda_{B+1} / lambda_k, estimated by
                            // the smallest Ritz value of the block over the k-th
                            const double thb = __ldcg(&theta[B - 1]);
                            if (thb > 0.0 && thb < ref) lr = log(thb / ref);
                        }
                        if (lr < 0.0 && rho > 0.0) {
                            const double need = ceil(log(tol / rho) / lr);
                            step = (need < 2.0) ? 2 : (need > 8.0 ? 8 : (int)need);
                        }
                        st->prev_worst = rho;
                        st->prev_it = it_now;
                        st->next_rr = it_now + step;
                    }
                    st->worst = rho;
                }
            }
            __syncwarp();
            for (int c = lane; c < B; c += 32) {
                const double d = __ldcg(&S[c * B + c]);
                dv[c] = d > 0.0 ? rsqrt(d) : 0.0;
            }
            __syncwarp();
            if (chol_mode == 1) {
                // Right-looking Cholesky, lane c owns column c of the (column-normalised) Gram matrix in shared memory
                // (pitch B + 1: conflict-free): step j scales row j by a reciprocal square root and every lane
                // updates its own column with INDEPENDENT FMAs.  The left-looking loop below recomputes, in every
                // step and on every lane, a length-j dependent chain for the pivot and another for the row, with
                // sqrt + division on the critical path: ~8-17 us per factorisation, 21 factorisations per solve.
                const int c = lane;
                for (int e = lane; e < B * B; e += 32) u.d.Sn[e / B][e % B] = __ldcg(&S[e]) * dv[e / B] * dv[e % B];
                __syncwarp();
#pragma unroll 1
                for (int j = 0; j < B; ++j) {
                    const double ajj = u.d.Sn[j][j];
                    double d, rd;
                    if (ajj > 1e-300) { rd = rsqrt(ajj); d = ajj * rd; } else { d = 1e-150; rd = 1e150; }
                    double rjc = 0.0;
                    if (c < B && c >= j) {
                        rjc = (c == j) ? d : u.d.Sn[j][c] * rd;
                        u.d.Rm[j][c] = rjc;
                    }
                    __syncwarp();
                    if (c < B) {
#pragma unroll 4
                        for (int i = j + 1; i <= c; ++i) u.d.Sn[i][c] = fma(-u.d.Rm[j][i], rjc, u.d.Sn[i][c]);
                    }
                    __syncwarp();
                }
                for (int e = lane; e < B * B; e += 32) {
                    R[e] = (e / B <= e % B) ? u.d.Rm[e / B][e % B] : 0.0;
                    S[e] = 0.0;
                    T[e] = 0.0;
                }
            } else {
            for (int e = lane; e < B * B; e += 32) u.d.Sn[e / B][e % B] = __ldcg(&S[e]) * dv[e / B] * dv[e % B];
            __syncwarp();
            for (int j = 0; j < B; ++j) {
                double d = u.d.Sn[j][j];
                for (int m = 0; m < j; ++m) d -= u.d.Rm[m][j] * u.d.Rm[m][j];
                d = (d > 1e-300) ? sqrt(d) : 1e-150;
                for (int c = j + lane; c < B; c += 32) {
                    double v = u.d.Sn[j][c];
                    for (int m = 0; m < j; ++m) v -= u.d.Rm[m][j] * u.d.Rm[m][c];
                    u.d.Rm[j][c] = (c == j) ? d : v / d;
                }
                __syncwarp();
            }
            for (int e = lane; e < B * B; e += 32) {
                R[e] = (e / B <= e % B) ? u.d.Rm[e / B][e % B] : 0.0;
                S[e] = 0.0;
                T[e] = 0.0;
            }
            }
            for (int c = lane; c < B; c += 32) { dinv[c] = dv[c]; res[c] = 0.0; }
        }
    };
    // phase E: own rows  X = src D^-1 R^-1  (one thread per row), result to global X and to Xm
    auto solve_phase = [&](double (*src)[B + 1]) {
        for (int e = tid; e < B * B; e += 512) Qs[e / B][e % B] = __ldcg(&R[e]);
        if (tid < B) {
            dv[tid] = __ldcg(&dinv[tid]);
            rq[tid] = 1.0 / __ldcg(&R[tid * B + tid]);     // chol_mode 1: reciprocal pivots, no division in the chain
        }
        __syncthreads();
        if (tid < RPC) {
            double x[B];
#pragma unroll
            for (int c = 0; c < B; ++c) {
                double v = src[tid][c] * dv[c];
#pragma unroll
                for (int m = 0; m < c; ++m) v -= x[m] * Qs[m][c];
                x[c] = chol_mode ? v * rq[c] : v / Qs[c][c];
            }
#pragma unroll
            for (int c = 0; c < B; ++c) {
                Xm[tid][c] = x[c];
                if (r0 + tid < n) X[(size_t)(r0 + tid) * B + c] = x[c];
            }
        }
        __syncthreads();
    };

    for (int it = 0; it < max_iter; ++it) {
        bool rr, qr2;
        if (rr_every > 0) {
            const int ph = it % rr_every;
            rr = ph == rr_every - 1;
            qr2 = !rr && ph == rr_every - 2;
        } else {
            // adaptive: first Ritz step at iteration rr0 - 1, then where phase D of the last one placed it (the value
            // was published before the grid barrier that ended that iteration)
            int nxt = __ldcg(&st->next_rr);
            if (nxt < rr0 - 1) nxt = rr0 - 1;
            rr = it == nxt || it == max_iter - 1;
            qr2 = !rr && (it + 1 == nxt || it + 2 == max_iter);
        }

        // ---- A: Y rows = G rows . X
        double acc0 = 0.0, acc1 = 0.0;
        for (int j0 = 0; j0 < n; j0 += TJ) {
            for (int e = tid; e < RPC * TJ; e += 512) {
                const int q = e / TJ, j = e % TJ;
                u.a.Gs[q][j] = (r0 + q < n && j0 + j < n) ? G[(size_t)(r0 + q) * n + j0 + j] : 0.0;
            }
            for (int e = tid; e < TJ * B; e += 512) {
                const int j = e / B;
                u.a.Xt[j][e % B] = (j0 + j < n) ? __ldcg(&X[(size_t)(j0 + j) * B + e % B]) : 0.0;
            }
            __syncthreads();
#pragma unroll 8
            for (int j = 0; j < TJ; j += 2) {
                acc0 = fma(u.a.Gs[tr][j], u.a.Xt[j][tc], acc0);
                acc1 = fma(u.a.Gs[tr][j + 1], u.a.Xt[j + 1][tc], acc1);
            }
            __syncthreads();
        }
        Ys[tr][tc] = (r < n) ? acc0 + acc1 : 0.0;
        if (rr) Xm[tr][tc] = (r < n) ? __ldcg(&X[(size_t)r * B + tc]) : 0.0;
        __syncthreads();
        if (rr) partial_gram(Xm, Ys, T); else partial_gram(Ys, Ys, S);
        grid_barrier(bar, epoch);

        if (rr) {
            // ---- B: Rayleigh-Ritz on CTA 0 (one-sided Jacobi on the columns of T, one column pair per warp)
            if (blockIdx.x == 0) {
                for (int e = tid; e < B * B; e += 512) {
                    const int a = e / B, b = e % B;
                    u.b.Tm[a][b] = 0.5 * (__ldcg(&T[a * B + b]) + __ldcg(&T[b * B + a]));
                }
                __syncthreads();
                for (int sweep = 0; sweep < 40; ++sweep) {
                    if (tid == 0) u.b.rotated = 0;
                    // squared column norms, refreshed once per sweep and updated analytically per rotation
                    // (al' = al - t ga, be' = be + t ga), so a rotation needs ONE shuffle reduction, not three
                    if (tid < B) {
                        double sq = 0.0;
                        for (int i = 0; i < B; ++i) sq = fma(u.b.Tm[i][tid], u.b.Tm[i][tid], sq);
                        u.b.th[tid] = sq;
                    }
                    __syncthreads();
                    for (int rd = 0; rd < B - 1; ++rd) {
                        if (warp < B / 2) {
                            const int pr = warp, mm = B - 1;
                            int a, b;
                            if (pr == 0) { a = mm; b = rd % mm; } else { a = (rd + pr) % mm; b = (rd - pr + mm) % mm; }
                            double ga = 0;
                            for (int i = lane; i < B; i += 32) ga = fma(u.b.Tm[i][a], u.b.Tm[i][b], ga);
                            ga = warp_sum(ga);
                            const double al = u.b.th[a], be = u.b.th[b];
                            if (al > 0 && be > 0 && ga * ga > jthr * jthr * al * be) {
                                // t = tan(theta) = sign(d) 2ga / (|d| + sqrt(d^2 + 4 ga^2)),  d = be - al
                                const double d = be - al, g2 = 2.0 * ga;
                                const double t = copysign(g2, d * g2) / (fabs(d) + sqrt(fma(d, d, g2 * g2)));
                                const double c = rsqrt(fma(t, t, 1.0)), sn = c * t;
                                for (int i = lane; i < B; i += 32) {
                                    const double x = u.b.Tm[i][a], y = u.b.Tm[i][b];
                                    u.b.Tm[i][a] = c * x - sn * y;
                                    u.b.Tm[i][b] = sn * x + c * y;
                                }
                                __syncwarp();          // every lane has read th[a], th[b] (racecheck, round 2)
                                if (lane == 0) {
                                    u.b.th[a] = al - t * ga;
                                    u.b.th[b] = be + t * ga;
                                    u.b.rotated = 1;
                                }
                            }
                        }
                        __syncthreads();
                    }
                    const int any = u.b.rotated;
                    __syncthreads();
                    if (!any) break;
                }
                if (tid < B) {
                    double sq = 0.0;
                    for (int i = 0; i < B; ++i) sq = fma(u.b.Tm[i][tid], u.b.Tm[i][tid], sq);
                    u.b.th[tid] = sqrt(sq);
                }
                __syncthreads();
                if (tid < B) {
                    int rk = 0;
                    for (int j = 0; j < B; ++j)
                        rk += (u.b.th[j] > u.b.th[tid] || (u.b.th[j] == u.b.th[tid] && j < tid));
                    u.b.order[rk] = tid;
                }
                __syncthreads();
                for (int e = tid; e < B * B; e += 512) {
                    const int i = e / B, q = e % B;
                    const int src = u.b.order[q];
                    Qm[e] = u.b.th[src] > 0.0 ? u.b.Tm[i][src] / u.b.th[src] : (i == src ? 1.0 : 0.0);
                }
                if (tid < B) theta[tid] = u.b.th[u.b.order[tid]];
            }
            grid_barrier(bar, epoch);

            // ---- C: rotate own rows, partial Gram of the rotated Y, residuals
            for (int e = tid; e < B * B; e += 512) Qs[e / B][e % B] = __ldcg(&Qm[e]);
            if (tid < B) dv[tid] = __ldcg(&theta[tid]);
            __syncthreads();
            double yr = 0.0, xr = 0.0;
#pragma unroll 8
            for (int m = 0; m < B; ++m) {
                yr = fma(Ys[tr][m], Qs[m][tc], yr);
                xr = fma(Xm[tr][m], Qs[m][tc], xr);
            }
            __syncthreads();
            Ys[tr][tc] = yr;
            Xm[tr][tc] = xr;
            if (r < n) X[(size_t)r * B + tc] = xr;            // Ritz vectors (kept if this step converges)
            __syncthreads();
            partial_gram(Ys, Ys, S);
            if (tid < B) {
                double sq = 0.0;
                const double th = dv[tid];
                for (int q = 0; q < RPC; ++q) {
                    const double d = Ys[q][tid] - th * Xm[q][tid];
                    sq = fma(d, d, sq);
                }
                atomicAdd(&res[tid], sq);
            }
            grid_barrier(bar, epoch);
        }

        // ---- D: Cholesky (+ convergence test after a Ritz step)
        cholesky_phase(rr ? 1 : 0, 1, it);
        grid_barrier(bar, epoch);
        if (rr && __ldcg(&st->converged)) break;

        // ---- E: own rows X = Y D^-1 R^-1
        solve_phase(Ys);
        if (qr2) {
            // second Cholesky-QR pass on the freshly orthogonalised block
            partial_gram(Xm, Xm, S);
            grid_barrier(bar, epoch);
            cholesky_phase(0, 0, it);
            grid_barrier(bar, epoch);
            for (int e = tid; e < RPC * B; e += 512) Ys[e / B][e % B] = Xm[e / B][e % B];
            __syncthreads();
            solve_phase(Ys);
        }
        grid_barrier(bar, epoch);
    }
}

// =====================================================================================
// Fused subspace iteration, second generation (round 2): two grid barriers per plain iteration
// =====================================================================================
// Measured on the kernel above (config 2, n = 500, B = 32): ~35 us per plain iteration and ~150 us per Ritz round,
// all of it latency -- three grid barriers per iteration, a one-warp Cholesky and a one-thread-per-row triangular
// solve on CTA 0 / 16 threads while the grid waits, and a G-tile loop with two block barriers per 64 columns whose
// inner product issues two shared-memory loads per DFMA.  This kernel removes the serial sections instead of
// speeding them up:
//   * 8 rows of G per CTA stay in shared memory for the whole solve (n <= 1184: ceil(n / 8) co-resident CTAs);
//     a thread owns one column of the block and a 1/NKS slice of the inner dimension for ALL 8 rows: one L2 load of
//     X and four broadcast LDS.128 of G per 8 DFMAs, no barrier inside the product;
//   * the B x B Gramians are accumulated with atomics into parity-buffered accumulators (the buffer of the next
//     iteration is cleared by CTA 0 while this one is in use), so EVERY CTA factorises them redundantly right after
//     the barrier -- Cholesky with all 512 threads, explicit triangular inverse, own rows X = Y (D^-1 R^-1) as a
//     small GEMM -- and the barrier between "factorise" and "solve" is gone;
//   * the Rayleigh-Ritz Jacobi and the convergence test / adaptive schedule also run redundantly in every CTA
//     (same inputs, same instruction sequence: identical decisions), which removes the barrier after them.
// Plain iteration: product -> barrier -> factorise + solve -> barrier.  Ritz iteration: one more barrier.
template <int B>
__global__ void __launch_bounds__(512, 1)
topk_fused2_kernel(const double* __restrict__ G, int n, int k, double tol, int max_iter, double jthr, int rr0,
                   double* X, double* acc, double* theta, TopkState* st, unsigned int* bar, int prof) {
    constexpr int RPC = 8, NKS = 512 / B, NOUT = RPC * B;
    constexpr int ACC = 3 * B * B + B;                 // S, T, S2 (second Cholesky-QR pass), residuals
    extern __shared__ __align__(16) double sm2[];
    double* Gs = sm2;                                  // [n][RPC]  own rows of G, transposed
    double* red = Gs + (size_t)n * RPC;                // [NKS][RPC][B] partial products
    __shared__ double Ys[RPC][B + 1], Xm[RPC][B + 1];
    __shared__ double Sn[B][B + 1], Wm[B][B + 1], Qs[B][B + 1];
    __shared__ double dv[B], rdg[B], thv[B], nrm[B], piv[B + 1];
    __shared__ int order[B];
    __shared__ int rotated;
    __shared__ int sh_conv, sh_next_rr, sh_prev_it;
    __shared__ double sh_prev_worst;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tc = tid % B, ks = tid / B;
    const int oq = tid / B, oc = tid % B;              // output element of this thread (tid < NOUT)
    const int r0 = blockIdx.x * RPC;
    unsigned int epoch = 0;
    int it = 0;
    // development aid (VIP_B200_TOPK_PROF=1): phase timestamps of CTA 0 for iteration 2 and the first Ritz iteration
    __shared__ unsigned long long tp[2][12];
    int prof_rr_done = 0;
    auto stamp = [&](int row, int slot) {
        if (prof && blockIdx.x == 0 && tid == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            tp[row][slot] = t;
        }
    };

    for (int e = tid; e < RPC * n; e += 512) {
        const int q = e / n, j = e - q * n;
        Gs[(size_t)j * RPC + q] = (r0 + q < n) ? G[(size_t)(r0 + q) * n + j] : 0.0;
    }
    if (tid < NOUT) Xm[oq][oc] = (r0 + oq < n) ? __ldcg(&X[(size_t)(r0 + oq) * B + oc]) : 0.0;
    if (tid == 0) { sh_conv = 0; sh_next_rr = 0; sh_prev_it = 0; sh_prev_worst = 0.0; }
    __syncthreads();

    // upper triangle (a <= b) of  out += P^T Q  over the RPC own rows; full matrix when `full`
    auto partial_gram = [&](const double (*P)[B + 1], const double (*Q)[B + 1], double* out, bool full) {
        for (int e = tid; e < B * B; e += 512) {
            const int a = e / B, b = e % B;
            if (!full && a > b) continue;
            double s = 0.0;
#pragma unroll
            for (int q = 0; q < RPC; ++q) s = fma(P[q][a], Q[q][b], s);
            atomicAdd(&out[e], s);
        }
    };
    // Wm = D^-1 R^-1 with R^T R = D S D (S: upper triangle of a B x B Gramian in global memory, D = diag(S)^-1/2):
    // right-looking Cholesky on all threads, explicit inverse of the triangular factor (one thread per row)
    auto factorise = [&](const double* S) {
        for (int e = tid; e < B * B; e += 512) {
            const int a = e / B, b = e % B;
            Sn[a][b] = __ldcg(&S[(a <= b) ? e : b * B + a]);
        }
        __syncthreads();
        if (tid < B) dv[tid] = Sn[tid][tid] > 0.0 ? rsqrt(Sn[tid][tid]) : 0.0;
        __syncthreads();
        for (int e = tid; e < B * B; e += 512) {
            const int a = e / B, b = e % B;
            const double v = Sn[a][b] * dv[a] * dv[b];
            Sn[a][b] = v;
            Wm[a][b] = (a == b) ? 1.0 : 0.0;
            if (e == 0) piv[0] = (v > 1e-300) ? 1.0 / v : 1e300;
        }
        __syncthreads();
        // Elimination with the inverse accumulated alongside.  Measured on the B200 (tools/microbench/fp64_lat.cu,
        // profiles/r02v_fp64_lat.txt): DFMA / DMUL / DADD 9 cycles dependent latency, but rsqrt 77, sqrt 100, divide
        // 126 cycles, and only 1 special-function result per 6-9 cycles and SM -- 16 warps computing the same
        // reciprocal are THROUGHPUT bound (the first versions of this phase: 19.7 and 14.4 us of a 29 us iteration).
        // So: step j only needs 1 / a_jj -- the trailing update is  S[i][c] -= S[j][i] S[j][c] / a_jj  and, with U
        // the inverse factor before its column scaling (R^-1 = U diag(a_jj^-1/2)),  U[m][c] -= U[m][j] S[j][c] / a_jj
        // for m <= j < c: at most one FMA per matrix element and step -- and ONE lane computes the reciprocal of the
        // NEXT pivot (look-ahead: two loads, two FMAs, one divide) while the other warps update their elements;
        // the block barrier that ends the step publishes it.
        // Thread mapping (measured with clock64 on the B200: one element visit -- 3 LDS, DMUL, DFMA, STS behind a
        // lane-dependent branch -- costs ~170 cycles, the look-ahead with divide and sqrt 444; serialised in warp 0
        // they made a step 850 cycles): warp 15 does nothing but the look-ahead, warps 0-14 own the elements
        // (one matrix row per warp and visit: the row test is warp-uniform, only the store is predicated).
        constexpr int NUPD = (B * B + 479) / 480;
#pragma unroll 1
        for (int j = 0; j < B; ++j) {
            const double inv = piv[j];
            if (warp == 15) {
                if (lane == 0 && j + 1 < B) {
                    const double sj = Sn[j][j + 1];
                    const double an = fma(-(sj * inv), sj, Sn[j + 1][j + 1]);
                    piv[j + 1] = (an > 1e-300) ? 1.0 / an : 1e300;
                }
            } else {
                double val[NUPD], *dst[NUPD];
                bool act[NUPD];
#pragma unroll
                for (int r = 0; r < NUPD; ++r) {
                    const int e = tid + 480 * r;
                    // threads beyond the matrix read element (0, 0), which no step writes (racecheck: reading another
                    // thread's element, even to discard it, is a hazard), and store nothing
                    const bool valid = e < B * B;
                    const int i = valid ? e / B : 0, c = valid ? e % B : 0;
                    const bool lower = i > j;
                    // the next pivot (j+1, j+1) is left alone: the look-ahead reads its pre-update value in this very
                    // step, and nothing reads it afterwards
                    act[r] = valid && (lower ? (c >= i && !(c == i && i == j + 1)) : (c > j));
                    const double mult = lower ? Sn[j][i] : Wm[i][j];
                    dst[r] = lower ? &Sn[i][c] : &Wm[i][c];
                    val[r] = fma(-(mult * inv), Sn[j][c], *dst[r]);
                }
#pragma unroll
                for (int r = 0; r < NUPD; ++r)
                    if (act[r]) *dst[r] = val[r];
            }
            __syncthreads();
        }
        if (tid < B) rdg[tid] = sqrt(piv[tid]);            // 1 / R[j][j]
        __syncthreads();
        // R^-1 = U diag(sqrt(1 / a_jj)),  W = D^-1 R^-1
        for (int e = tid; e < B * B; e += 512) {
            const int m = e / B, c = e % B;
            Wm[m][c] = (m <= c) ? dv[m] * Wm[m][c] * rdg[c] : 0.0;
        }
        __syncthreads();
    };
    // own rows  X = src . Wm  -> Xm and global X
    auto solve = [&](const double (*src)[B + 1]) {
        double x = 0.0;
        if (tid < NOUT) {
#pragma unroll 8
            for (int m = 0; m < B; ++m) x = fma(src[oq][m], Wm[m][oc], x);     // Wm is upper triangular (zeros below)
        }
        __syncthreads();
        if (tid < NOUT) {
            Xm[oq][oc] = x;
            if (r0 + oq < n) X[(size_t)(r0 + oq) * B + oc] = x;
        }
        __syncthreads();
    };

    for (it = 0; it < max_iter; ++it) {
        int nxt = sh_next_rr;
        if (nxt < rr0 - 1) nxt = rr0 - 1;
        const bool rr = it == nxt || it == max_iter - 1;
        const bool qr2 = !rr && (it + 1 == nxt || it + 2 == max_iter);
        double* A0 = acc + (size_t)(it & 1) * ACC;
        double* S = A0;
        double* T = A0 + B * B;
        double* S2 = A0 + 2 * B * B;
        double* res = A0 + 3 * B * B;
        const int prow = (it == 2 && !rr) ? 0 : (rr && !prof_rr_done) ? 1 : -1;
        if (prow >= 0) stamp(prow, 0);
        if (blockIdx.x == 0) {                          // accumulators of the next iteration (last read before the
            double* A1 = acc + (size_t)((it + 1) & 1) * ACC;      // barrier that ended the previous one)
            for (int e = tid; e < ACC; e += 512) A1[e] = 0.0;
        }

        // ---- A: Y rows = G rows . X   (thread: column tc, inner indices ks, ks + NKS, ...; all RPC rows)
        {
            double a[RPC];
#pragma unroll
            for (int q = 0; q < RPC; ++q) a[q] = 0.0;
#pragma unroll 8
            for (int j = ks; j < n; j += NKS) {
                const double x = __ldcg(&X[(size_t)j * B + tc]);
                const double2* g = reinterpret_cast<const double2*>(Gs + (size_t)j * RPC);
#pragma unroll
                for (int q = 0; q < RPC / 2; ++q) {
                    const double2 gv = g[q];
                    a[2 * q] = fma(gv.x, x, a[2 * q]);
                    a[2 * q + 1] = fma(gv.y, x, a[2 * q + 1]);
                }
            }
#pragma unroll
            for (int q = 0; q < RPC; ++q) red[((size_t)ks * RPC + q) * B + tc] = a[q];
        }
        __syncthreads();
        if (tid < NOUT) {
            double s = 0.0;
#pragma unroll 4
            for (int g = 0; g < NKS; ++g) s += red[((size_t)g * RPC + oq) * B + oc];
            Ys[oq][oc] = s;                             // rows beyond n: their G rows are zero
        }
        __syncthreads();
        if (prow >= 0) stamp(prow, 1);
        if (rr) partial_gram(Xm, Ys, T, true); else partial_gram(Ys, Ys, S, false);
        if (prow >= 0) stamp(prow, 2);
        grid_barrier(bar, epoch);
        if (prow >= 0) stamp(prow, 3);

        if (rr) {
            // ---- B: Rayleigh-Ritz, redundantly in every CTA (one-sided Jacobi on the columns of T)
            for (int e = tid; e < B * B; e += 512) {
                const int a = e / B, b = e % B;
                Sn[a][b] = 0.5 * (__ldcg(&T[a * B + b]) + __ldcg(&T[b * B + a]));
            }
            __syncthreads();
            for (int sweep = 0; sweep < 40; ++sweep) {
                if (tid == 0) rotated = 0;            // set by rotations whose pair was coupled by more than 1e-8
                if (tid < B) {
                    double sq = 0.0;
                    for (int i = 0; i < B; ++i) sq = fma(Sn[i][tid], Sn[i][tid], sq);
                    nrm[tid] = sq;
                }
                __syncthreads();
                for (int rd = 0; rd < B - 1; ++rd) {
                    if (warp < B / 2) {
                        const int pr = warp, mm = B - 1;
                        int a, b;
                        if (pr == 0) { a = mm; b = rd % mm; } else { a = (rd + pr) % mm; b = (rd - pr + mm) % mm; }
                        double ga = 0;
                        for (int i = lane; i < B; i += 32) ga = fma(Sn[i][a], Sn[i][b], ga);
                        ga = warp_sum(ga);
                        const double al = nrm[a], be = nrm[b];
                        if (al > 0 && be > 0 && ga * ga > jthr * jthr * al * be) {
                            // c, s from two rsqrt (77 cycles each) instead of sqrt -> divide -> rsqrt (100 + 126 + 77):
                            // 1/h = rsqrt(d^2 + g^2), cos(2 theta) = |d| / h, c^2 = (1 + cos 2theta) / 2,
                            // s = sign(d) g / (2 h c); no cancellation in either limit (|d| >> |g|: s is a product).
                            const double d = be - al, g2 = 2.0 * ga;
                            const double rh = rsqrt(fma(d, d, g2 * g2));
                            const double c2 = fma(0.5 * fabs(d), rh, 0.5);
                            const double q = rsqrt(c2);
                            const double c = c2 * q;
                            const double sn = copysign(0.5, d) * g2 * rh * q;
                            for (int i = lane; i < B; i += 32) {
                                const double x = Sn[i][a], y = Sn[i][b];
                                Sn[i][a] = c * x - sn * y;
                                Sn[i][b] = sn * x + c * y;
                            }
                            __syncwarp();
                            if (lane == 0) {
                                const double cs2 = 2.0 * c * sn * ga;
                                nrm[a] = fma(c * c, al, fma(sn * sn, be, -cs2));
                                nrm[b] = fma(sn * sn, al, fma(c * c, be, cs2));
                                // cyclic Jacobi converges quadratically: a sweep whose rotations all started from
                                // couplings below 1e-8 leaves them at ~1e-16, so it is the last one -- no extra
                                // sweep (31 rounds, ~10 us) just to observe that nothing is left to rotate
                                if (ga * ga > 1e-16 * al * be) rotated = 1;
                            }
                        }
                    }
                    __syncthreads();
                }
                const int any = rotated;
                __syncthreads();
                if (!any) break;
            }
            if (tid < B) {
                double sq = 0.0;
                for (int i = 0; i < B; ++i) sq = fma(Sn[i][tid], Sn[i][tid], sq);
                nrm[tid] = sqrt(sq);
            }
            __syncthreads();
            if (tid < B) {
                int rk = 0;
                for (int j = 0; j < B; ++j) rk += (nrm[j] > nrm[tid] || (nrm[j] == nrm[tid] && j < tid));
                order[rk] = tid;
            }
            __syncthreads();
            for (int e = tid; e < B * B; e += 512) {
                const int i = e / B, q = e % B;
                const int src = order[q];
                Qs[i][q] = nrm[src] > 0.0 ? Sn[i][src] / nrm[src] : (i == src ? 1.0 : 0.0);
            }
            if (tid < B) {
                thv[tid] = nrm[order[tid]];
                if (blockIdx.x == 0) theta[tid] = thv[tid];
            }
            __syncthreads();

            if (prow >= 0) stamp(prow, 4);
            // ---- C: rotate own rows, partial Gram of the rotated Y, residuals
            double yr = 0.0, xr = 0.0;
            if (tid < NOUT) {
#pragma unroll 8
                for (int m = 0; m < B; ++m) {
                    yr = fma(Ys[oq][m], Qs[m][oc], yr);
                    xr = fma(Xm[oq][m], Qs[m][oc], xr);
                }
            }
            __syncthreads();
            if (tid < NOUT) {
                Ys[oq][oc] = yr;
                Xm[oq][oc] = xr;
                if (r0 + oq < n) X[(size_t)(r0 + oq) * B + oc] = xr;        // Ritz vectors (kept if this step converges)
            }
            __syncthreads();
            partial_gram(Ys, Ys, S, false);
            if (tid < B) {
                double sq = 0.0;
                const double th = thv[tid];
                for (int q = 0; q < RPC; ++q) {
                    const double d = Ys[q][tid] - th * Xm[q][tid];
                    sq = fma(d, d, sq);
                }
                atomicAdd(&res[tid], sq);
            }
            if (prow >= 0) stamp(prow, 5);
            grid_barrier(bar, epoch);
            if (prow >= 0) stamp(prow, 6);

            // convergence test and adaptive schedule: every CTA, from the same global values
            if (tid < B) nrm[tid] = (tid < k) ? sqrt(__ldcg(&res[tid])) : 0.0;      // loads in parallel
            __syncthreads();
            if (tid == 0) {
                double worst = 0.0;
                for (int q = 0; q < k; ++q) worst = fmax(worst, nrm[q]);
                const double ref = thv[k - 1];
                const double rho = ref > 0.0 ? worst / ref : 0.0;
                int conv = 0;
                if (worst <= tol * ref) conv = 1;
                else {
                    int step = 4;
                    const double prho = sh_prev_worst;
                    const int pit = sh_prev_it;
                    double lr = 0.0;
                    if (pit > 0 && prho > 0.0 && rho > 0.0 && rho < prho) {
                        lr = log(rho / prho) / (double)(it - pit);
                    } else {
                        const double thb = thv[B - 1];
                        if (thb > 0.0 && thb < ref) lr = log(thb / ref);
                    }
                    if (lr < 0.0 && rho > 0.0) {
                        const double need = ceil(log(tol / rho) / lr);
                        step = (need < 2.0) ? 2 : (need > 8.0 ? 8 : (int)need);
                    }
                    sh_prev_worst = rho;
                    sh_prev_it = it;
                    sh_next_rr = it + step;
                }
                sh_conv = conv;
                if (blockIdx.x == 0) {
                    st->worst = rho;
                    if (conv) st->converged = 1;
                }
            }
        }
        if (blockIdx.x == 0 && tid == 0) st->iters += 1;
        __syncthreads();
        if (rr && sh_conv) break;

        // ---- D + E: factorise the Gramian, own rows X = Y D^-1 R^-1
        if (prow >= 0) stamp(prow, 7);
        factorise(S);
        if (prow >= 0) stamp(prow, 8);
        solve(Ys);
        if (prow >= 0) stamp(prow, 9);
        if (qr2) {
            // second Cholesky-QR pass on the freshly orthogonalised block (the Ritz step needs an orthonormal X)
            partial_gram(Xm, Xm, S2, false);
            grid_barrier(bar, epoch);
            factorise(S2);
            if (tid < NOUT) Ys[oq][oc] = Xm[oq][oc];
            __syncthreads();
            solve(Ys);
        }
        grid_barrier(bar, epoch);
        if (prow >= 0) {
            stamp(prow, 10);
            if (prow == 1) prof_rr_done = 1;
            if (prof && blockIdx.x == 0 && tid == 0) {
                const unsigned long long* t = tp[prow];
                if (prow == 0)
                    printf("topk_fused2 n=%d B=%d plain it=%d [ns]: product %llu | gram atomics %llu | barrier %llu | "
                           "factorise %llu | solve %llu | barrier %llu | total %llu\n", n, B, it, t[1] - t[0], t[2] - t[1],
                           t[3] - t[2], t[8] - t[7], t[9] - t[8], t[10] - t[9], t[10] - t[0]);
                else
                    printf("topk_fused2 n=%d B=%d ritz it=%d [ns]: product %llu | gram atomics %llu | barrier %llu | "
                           "jacobi %llu | rotate+gram %llu | barrier %llu | check %llu | factorise %llu | solve %llu | "
                           "barrier %llu | total %llu\n", n, B, it, t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3],
                           t[5] - t[4], t[6] - t[5], t[7] - t[6], t[8] - t[7], t[9] - t[8], t[10] - t[9], t[10] - t[0]);
            }
        }
    }
}

size_t topk_fused2_smem_bytes(int n, int B) { return ((size_t)n * 8 + (size_t)512 * 8) * sizeof(double); }

// block width of the subspace solver for k wanted pairs: 16 (k <= 10), 32 (k <= 24: the fused single-launch kernel),
// 64 (k <= 56: per-phase kernels; BASELINE config 5's ncomp = 50 on the exact path, psfsub/svd.py:466-475 takes any k)
int topk_block_width(int k) { return (k <= 10) ? 16 : (k <= 24) ? 32 : 64; }

size_t eigh_topk_workspace_bytes(int n, int B) {
    return ((size_t)2 * n * B + 4 * (size_t)B * B + 4 * B) * sizeof(double) + 512;
}

template <int B>
static int topk_run(const double* G, int n, int k, double tol, int max_iter, double* evals, double* evecs,
                    void* ws, int* info, int* launches, cudaStream_t st, int* async_info) {
    double* X = reinterpret_cast<double*>(ws);
    double* Y = X + (size_t)n * B;
    double* T = Y + (size_t)n * B;
    double* S = T + B * B;
    double* Qm = S + B * B;
    double* R = Qm + B * B;
    double* theta = R + B * B;
    double* res = theta + B;
    double* dinv = res + B;
    TopkState* state = reinterpret_cast<TopkState*>(dinv + 2 * B);
    VB_CHECK_CUDA(cudaMemsetAsync(T, 0, (size_t)(4 * B * B + 4 * B) * sizeof(double) + sizeof(TopkState) + 128, st));
    int nl = 0;
    const int grows = ceil_div(n, 16);
    topk_init_kernel<B><<<ceil_div(n * B, 256), 256, 0, st>>>(Y, n);
    // orthonormalise the start block: S = Y^T Y, Cholesky, solve
    topk_xty_kernel<B><<<grows, 256, 0, st>>>(Y, Y, n, S, state);
    topk_chol_kernel<B><<<1, 32, 0, st>>>(S, res, T, theta, k, tol, R, dinv, state, 0);
    topk_solve_kernel<B><<<ceil_div(n, 128), 128, 0, st>>>(X, Y, n, R, dinv, state);
    nl += 4;
    VB_CHECK_LAUNCH();
    TopkState h{};
    const char* je = getenv("VIP_B200_RITZ_THR");
    double jthr = je ? atof(je) : 1e-15;
    // fused persistent kernel: all iterations in one cooperative launch (grid = ceil(n / rows-per-CTA) <= #SMs)
    const char* fe = getenv("VIP_B200_TOPK_FUSED");
    const int fused_grid = ceil_div(n, 512 / B);
    if constexpr (B <= 32) {
    // second-generation fused kernel (8 rows of G per CTA resident in shared memory, two grid barriers per plain
    // iteration): n <= 8 * #SMs, and the Y block (free after the start-up kernels) holds its accumulators
    const int fused_mode = fe ? atoi(fe) : 2;
    const int grid2 = ceil_div(n, 8);
    const size_t smem2 = topk_fused2_smem_bytes(n, B);
    if (fused_mode >= 2 && grid2 <= kNumSMs && n >= 8 * B && smem2 <= 160 * 1024) {
        const char* r0e = getenv("VIP_B200_TOPK_RR0");
        int rr0 = r0e ? atoi(r0e) : 8;
        if (rr0 < 2) rr0 = 2;
        unsigned int* bar = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(state) + 64);
        double* acc = Y;                                  // 2 x (3 B^2 + B) doubles <= n B for n >= 8 B
        VB_CHECK_CUDA(cudaMemsetAsync(acc, 0, (size_t)2 * (3 * B * B + B) * sizeof(double), st));
        static bool configured = false;
        if (!configured) {
            VB_CHECK_CUDA(cudaFuncSetAttribute(topk_fused2_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               160 * 1024));
            configured = true;
        }
        int fmax = max_iter;
        // rotation threshold |a.b| > thr |a||b|: 1e-15 (first generation) is below the rounding noise of a 32-term
        // fp64 dot product and buys extra sweeps that chase it; B eps is the customary bound
        double jthr2 = je ? jthr : (double)B * 1.1e-16;
        const char* pe = getenv("VIP_B200_TOPK_PROF");
        int prof = pe ? atoi(pe) : 0;
        void* args[] = {(void*)&G, (void*)&n, (void*)&k, (void*)&tol, (void*)&fmax, (void*)&jthr2, (void*)&rr0,
                        (void*)&X, (void*)&acc, (void*)&theta, (void*)&state, (void*)&bar, (void*)&prof};
        const cudaError_t ce = cudaLaunchCooperativeKernel((const void*)topk_fused2_kernel<B>, dim3(grid2), dim3(512),
                                                           args, smem2, st);
        if (ce == cudaSuccess) {
            topk_output_kernel<B><<<ceil_div(n, 256), 256, 0, st>>>(X, theta, n, k, evals, evecs);
            VB_CHECK_LAUNCH();
            if (launches) *launches = nl + 3;
            if (async_info) {
                VB_CHECK_CUDA(cudaMemcpyAsync(async_info, state, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
                return 0;
            }
            VB_CHECK_CUDA(cudaMemcpyAsync(&h, state, sizeof(h), cudaMemcpyDeviceToHost, st));
            VB_CHECK_CUDA(cudaStreamSynchronize(st));
            if (info) { info[0] = h.iters; info[1] = h.converged; }
            return 0;
        }
        (void)cudaGetLastError();     // cooperative launch unavailable: first-generation kernel / per-phase kernels
    }
    if ((fe ? atoi(fe) : 1) && fused_grid <= kNumSMs) {
        const char* re = getenv("VIP_B200_TOPK_RR");
        int rr_every = re ? atoi(re) : 0;        // 0 = adaptive schedule (default; first Ritz step at iteration rr0), 4 = r01n
        if (rr_every < 0) rr_every = 1;
        const char* r0e = getenv("VIP_B200_TOPK_RR0");
        int rr0 = r0e ? atoi(r0e) : 8;
        if (rr0 < 2) rr0 = 2;
        unsigned int* bar = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(state) + 64);
        int fmax = rr_every > 0 ? ceil_div(max_iter, rr_every) * rr_every : max_iter;
        const char* ce_ = getenv("VIP_B200_TOPK_CHOL");
        int chol_mode = ce_ ? atoi(ce_) : 1;          // 1 (default): right-looking Cholesky + reciprocal pivots (phase D/E)
        void* args[] = {(void*)&G, (void*)&n, (void*)&k, (void*)&tol, (void*)&fmax, (void*)&jthr, (void*)&rr_every,
                        (void*)&chol_mode, (void*)&rr0, (void*)&X, (void*)&T, (void*)&S, (void*)&Qm, (void*)&R, (void*)&theta, (void*)&res,
                        (void*)&dinv, (void*)&state, (void*)&bar};
        const cudaError_t ce = cudaLaunchCooperativeKernel((const void*)topk_fused_kernel<B>, dim3(fused_grid),
                                                           dim3(512), args, 0, st);
        if (ce == cudaSuccess) {
            topk_output_kernel<B><<<ceil_div(n, 256), 256, 0, st>>>(X, theta, n, k, evals, evecs);
            VB_CHECK_LAUNCH();
            if (launches) *launches = nl + 2;
            if (async_info) {
                // deferred check: {iters, converged} land in the caller's PINNED buffer when the stream gets
                // there; no host synchronisation, the caller's pipeline keeps running ahead
                VB_CHECK_CUDA(cudaMemcpyAsync(async_info, state, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
                return 0;
            }
            VB_CHECK_CUDA(cudaMemcpyAsync(&h, state, sizeof(h), cudaMemcpyDeviceToHost, st));
            VB_CHECK_CUDA(cudaStreamSynchronize(st));
            if (info) { info[0] = h.iters; info[1] = h.converged; }
            return 0;
        }
        (void)cudaGetLastError();     // cooperative launch unavailable (e.g. MPS): use the per-phase kernels below
    }
    }
    // Rayleigh-Ritz every iteration: without it the columns of G X all tilt towards the dominant
    // eigenvector (lambda_0 / lambda_B ~ 1e5) and the Cholesky-QR of Y^T Y (condition number squared)
    // loses the trailing directions -- measured: no convergence in 400 steps with RR every 4th step.
    for (int it = 0; it < max_iter; ++it) {
        topk_matvec_kernel<B><<<ceil_div(n, 256 / B), 256, 0, st>>>(G, n, X, Y, state);
        topk_xty_kernel<B><<<grows, 256, 0, st>>>(X, Y, n, T, state);
        topk_ritz_kernel<B><<<1, 16 * B, 0, st>>>(T, Qm, theta, S, res, state, jthr);
        topk_rotate_kernel<B><<<ceil_div(n, topk_rot_rows(B)), 256, 0, st>>>(X, Y, n, Qm, theta, S, res, state);
        topk_chol_kernel<B><<<1, 32, 0, st>>>(S, res, T, theta, k, tol, R, dinv, state, 1);
        nl += 2;
        topk_solve_kernel<B><<<ceil_div(n, 128), 128, 0, st>>>(X, Y, n, R, dinv, state);
        nl += 4;
        if ((it >= 7 && (it & 3) == 3) || it == max_iter - 1) {
            VB_CHECK_LAUNCH();
            VB_CHECK_CUDA(cudaMemcpyAsync(&h, state, sizeof(h), cudaMemcpyDeviceToHost, st));
            VB_CHECK_CUDA(cudaStreamSynchronize(st));
            if (h.converged) break;
        }
    }
    topk_output_kernel<B><<<ceil_div(n, 256), 256, 0, st>>>(X, theta, n, k, evals, evecs);
    VB_CHECK_LAUNCH();
    nl += 1;
    if (launches) *launches = nl;
    if (info) { info[0] = h.iters; info[1] = h.converged; }
    if (async_info) { async_info[0] = h.iters; async_info[1] = h.converged; }   // this path synchronised anyway
    return 0;
}

// Leading k eigenpairs of the symmetric PSD matrix G (n x n fp64).  evals[k] descending,
// evecs[k x n] row j = eigenvector j.  info_host: iterations, converged flag.  Synchronises.
int eigh_topk_f64(const double* G, int n, int k, double tol, int max_iter, double* evals, double* evecs,
                  void* ws, size_t ws_bytes, int* info, int* launches, cudaStream_t st, int* async_info) {
    VB_REQUIRE(n >= 1 && k >= 1 && k <= n, "eigh_topk: need 1 <= k <= n");
    if (tol <= 0) tol = 1e-9;    // residual / lambda_k; eigenvector error ~ tol / relative gap
    if (max_iter <= 0) max_iter = 400;   // beyond this the caller is better off with the Jacobi solver
    const int B = topk_block_width(k);
    VB_REQUIRE(k <= 56, "eigh_topk: k=%d too large for the subspace solver (use the Jacobi solver)", k);
    VB_REQUIRE(n >= B, "eigh_topk: n=%d smaller than the block width %d (use the Jacobi solver)", n, B);
    VB_REQUIRE(ws_bytes >= eigh_topk_workspace_bytes(n, B), "eigh_topk: workspace too small");
    if (B == 16) return topk_run<16>(G, n, k, tol, max_iter, evals, evecs, ws, info, launches, st, async_info);
    if (B == 32) return topk_run<32>(G, n, k, tol, max_iter, evals, evecs, ws, info, launches, st, async_info);
    return topk_run<64>(G, n, k, tol, max_iter, evals, evecs, ws, info, launches, st, async_info);
}

}  // namespace vb
