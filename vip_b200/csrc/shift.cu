// Sub-pixel frame shift with the Fourier phase-ramp method ("vip-fft" semantics) for a whole cube.
//
// Replaces the per-frame Python loop  cube_shift -> frame_shift (imlib='vip-fft')
// (reference: src/vip_hci/preproc/recentering.py:257-305, 122-189).
//
// The reference zero-pads each frame by ceil(max|shift|) pixels to an even square plane of N pixels,
// multiplies its 2-d spectrum by exp(-2 pi i (sx fx + sy fy)) (f = fftfreq(N), Nyquist bin = -1/2),
// transforms back, keeps the real part and crops the original window.  The multiplier is separable and
// only the window of the plane is non-zero / needed, so per frame
//      out = Ty X Tx^T  -  (-1)^(r+c) sin(pi sx) sin(pi sy) / N^2 * sum_{r',c'} (-1)^(r'+c') X[r'][c']
// with the real Toeplitz operators  T[m][n] = Re D_N(m - n - s),
//      D_N(t) = (1/N) sum_k exp(2 pi i f_k t),   Re D_N(t) = sin(pi t) cos(pi t / N) / (N sin(pi t / N)).
// The second term is the Nyquist bin: a 1-d shear of a real line is real except for
// i X[N/2] sin(pi s) (-1)^n / N, and the second shear turns the alternating sum of those back into a real
// checkerboard (same bookkeeping as the rotation kernels, csrc/derotate.cu).  This file builds the
// operators and applies the checkerboard term; the two products per frame are batched GEMMs (csrc/gemm.cu).
#include "common.cuh"

namespace vb {

// T[f][m][n] = Re D_N(m - n - s_f), m, n in [0, L); evaluated in fp64, stored in fp32
__global__ void __launch_bounds__(256)
shift_operator_kernel(const double* __restrict__ shift, const int* __restrict__ nplane, int L,
                      float* __restrict__ T) {
    const int f = blockIdx.y, m = blockIdx.x;
    const double s = shift[f];
    const int N = nplane[f];
    const double s_round = rint(s);
    const int s_int = (int)s_round;
    const double s_frac = s - s_round;            // [-0.5, 0.5]
    const double num0 = -sinpi(s_frac);           // sin(pi (e - s_frac)) = (-1)^e * num0
    const double invN = 1.0 / (double)N;
    float* row = T + ((size_t)f * L + m) * L;
    for (int n = threadIdx.x; n < L; n += blockDim.x) {
        const int e = m - n - s_int;
        const double x = (double)e - s_frac;
        double v;
        if (fabs(x) < 1e-9) {
            v = 1.0;
        } else {
            double sn, cs;
            sincospi(x * invN, &sn, &cs);
            v = ((e & 1) ? -num0 : num0) * cs / ((double)N * sn);
        }
        row[n] = (float)v;
    }
}

// kappa[f] = sum_{r,c} (-1)^(r+c) X[f][r][c]   (fp64 accumulation, fixed order)
__global__ void __launch_bounds__(256)
checker_sum_kernel(const float* __restrict__ X, int ny, int nx, double* __restrict__ kappa) {
    __shared__ double red[8];
    const int f = blockIdx.x;
    const float* x = X + (size_t)f * ny * nx;
    double acc = 0.0;
    for (int i = threadIdx.x; i < ny * nx; i += 256) {
        const int r = i / nx, c = i - r * nx;
        const float v = x[i];
        acc += ((r + c) & 1) ? -(double)v : (double)v;
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        kappa[f] = t;
    }
}

// out[f][r][c] -= (-1)^(r+c) coef[f] kappa[f]
__global__ void __launch_bounds__(256)
checker_fix_kernel(float* __restrict__ out, int ny, int nx, const double* __restrict__ coef,
                   const double* __restrict__ kappa) {
    const int f = blockIdx.y;
    const float a = (float)(coef[f] * kappa[f]);
    if (a == 0.f) return;
    float* o = out + (size_t)f * ny * nx;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < ny * nx; i += gridDim.x * 256) {
        const int r = i / nx, c = i - r * nx;
        o[i] -= ((r + c) & 1) ? -a : a;
    }
}

int shift_operators(const double* shift, const int* nplane, int nf, int L, float* T, cudaStream_t st) {
    VB_REQUIRE(nf > 0 && L > 0 && nf <= 65535, "shift_operators: bad sizes nf=%d L=%d", nf, L);
    shift_operator_kernel<<<dim3(L, nf), 256, 0, st>>>(shift, nplane, L, T);
    VB_CHECK_LAUNCH();
    return 0;
}

int checker_correct(const float* X, float* out, int nf, int ny, int nx, const double* coef, double* kappa,
                    cudaStream_t st) {
    VB_REQUIRE(nf > 0 && ny > 0 && nx > 0 && nf <= 65535, "checker_correct: bad sizes");
    checker_sum_kernel<<<nf, 256, 0, st>>>(X, ny, nx, kappa);
    VB_CHECK_LAUNCH();
    int gx = ceil_div(ny * nx, 256 * 8);
    if (gx < 1) gx = 1;
    checker_fix_kernel<<<dim3(gx, nf), 256, 0, st>>>(out, ny, nx, coef, kappa);
    VB_CHECK_LAUNCH();
    return 0;
}

}  // namespace vb
