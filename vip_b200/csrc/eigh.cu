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

    for (int c = 0; c < 2 * WB; ++c) {
        const int gc = gcol(c);
        if (gc < n) {
            const double* src = W + (size_t)gc * n;
            for (int i = threadIdx.x; i < n; i += blockDim.x) cols[(size_t)c * n + i] = src[i];
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
            double alpha = 0.0, beta = 0.0, gamma = 0.0;
            for (int i = lane; i < n; i += 32) {
                const double xv = x[i], yv = y[i];
                alpha = fma(xv, xv, alpha);
                beta = fma(yv, yv, beta);
                gamma = fma(xv, yv, gamma);
            }
            alpha = warp_sum(alpha); beta = warp_sum(beta); gamma = warp_sum(gamma);
            if (alpha > 0.0 && beta > 0.0 && fabs(gamma) > tol * sqrt(alpha) * sqrt(beta)) {
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = copysign(1.0, zeta) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t);
                const double s = c * t;
                for (int i = lane; i < n; i += 32) {
                    const double xv = x[i], yv = y[i];
                    x[i] = c * xv - s * yv;
                    y[i] = s * xv + c * yv;
                }
                ++nrot;
            }
        }
        __syncthreads();
    }

    for (int c = 0; c < 2 * WB; ++c) {
        const int gc = gcol(c);
        if (gc < n) {
            double* dst = W + (size_t)gc * n;
            for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = cols[(size_t)c * n + i];
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
