// S/N of test resolution elements and S/N maps:  vip_hci/metrics/snr_source.py
//   snr()     :321-455  -- flux in a FWHM aperture at the test location against the fluxes of the non-overlapping
//                          apertures at the same separation, small-sample penalty of Mawet et al. 2014
//   snrmap()  :32-204   -- snr() at every pixel of an annulus (the reference forks processes over pixels)
// The reference sums apertures with photutils (CircularAperture + aperture_photometry(method='exact')): the sum over
// pixels of (exact area of circle ∩ unit pixel) x value.  Here: fp64 closed-form overlap areas (inclusion-exclusion
// of quadrant areas, same formula as oracle/vip_oracle.py), one thread per aperture for arbitrary centres and one
// warp per test pixel for the map (lanes stride over the apertures of that pixel's ring, single-pass shifted
// variance).  Pixels entirely inside / outside the circle skip the transcendental path.
#include "common.cuh"

namespace vb {

__device__ __forceinline__ double quadrant_area(double x, double y, double r) {
    x = fmin(x, r);
    y = fmin(y, r);
    if (x * x + y * y <= r * r) return x * y;
    const double xa = sqrt(fmax(r * r - y * y, 0.0));
    const double xm = fmin(x, xa);
    const double hx = 0.5 * (x * sqrt(fmax(r * r - x * x, 0.0)) + r * r * asin(fmin(x / r, 1.0)));
    const double hm = 0.5 * (xm * sqrt(fmax(r * r - xm * xm, 0.0)) + r * r * asin(fmin(xm / r, 1.0)));
    return y * xm + hx - hm;
}
__device__ __forceinline__ double signed_quadrant(double x, double y, double r) {
    const double s = ((x < 0.0) != (y < 0.0)) ? -1.0 : 1.0;
    if (x == 0.0 || y == 0.0) return 0.0;
    return s * quadrant_area(fabs(x), fabs(y), r);
}
// exact area of the unit pixel centred at (dx, dy) relative to the circle centre
__device__ __forceinline__ double pixel_weight(double dx, double dy, double r) {
    const double ax = fabs(dx), ay = fabs(dy);
    const double fx = ax + 0.5, fy = ay + 0.5;                    // farthest corner
    if (fx * fx + fy * fy <= r * r) return 1.0;
    const double nx = fmax(ax - 0.5, 0.0), ny = fmax(ay - 0.5, 0.0);   // nearest point of the pixel
    if (nx * nx + ny * ny >= r * r) return 0.0;
    const double x0 = dx - 0.5, x1 = dx + 0.5, y0 = dy - 0.5, y1 = dy + 0.5;
    double w = signed_quadrant(x1, y1, r) - signed_quadrant(x0, y1, r) - signed_quadrant(x1, y0, r) +
               signed_quadrant(x0, y0, r);
    return fmin(fmax(w, 0.0), 1.0);
}

// sum over pixels of weight x value for one aperture (NaN inside the aperture -> NaN)
__device__ __forceinline__ double aperture_sum(const float* __restrict__ img, int H, int W, double xc, double yc,
                                               double r) {
    int ix0 = (int)floor(xc - r + 0.5), ix1 = (int)ceil(xc + r + 0.5);
    int iy0 = (int)floor(yc - r + 0.5), iy1 = (int)ceil(yc + r + 0.5);
    ix0 = max(ix0, 0); iy0 = max(iy0, 0); ix1 = min(ix1, W); iy1 = min(iy1, H);
    double s = 0.0;
    for (int iy = iy0; iy < iy1; ++iy)
        for (int ix = ix0; ix < ix1; ++ix) {
            const double w = pixel_weight((double)ix - xc, (double)iy - yc, r);
            if (w > 0.0) s = fma(w, (double)__ldg(img + (size_t)iy * W + ix), s);
        }
    return s;
}

__global__ void __launch_bounds__(128)
aperture_sums_kernel(const float* __restrict__ img, int H, int W, const double* __restrict__ xs,
                     const double* __restrict__ ys, int nap, double r, double* __restrict__ out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a < nap) out[a] = aperture_sum(img, H, W, xs[a], ys[a], r);
}

// One warp per test pixel (px[i], py[i]).  Aperture j of the ring: the source offset rotated clockwise by j * angle,
// angle = 2 asin(fwhm / 2 / sep), nap = floor(2 pi / angle) (snr_source.py:265-309; the reference accumulates the
// rotation with a recurrence, here cos/sin(j angle) directly: centres agree to 1e-13 px).  Background = apertures
// 1 .. nap-1 (without 1 and nap-1 under exclude_negative_lobes) of `img`, plus all nap apertures... of `img2` when
// given (array2: every aperture of the second frame, the one at the test location included, snr_source.py:399-405);
// use2alone keeps only those.
__global__ void __launch_bounds__(256)
snr_points_kernel(const float* __restrict__ img, const float* __restrict__ img2, int H, int W,
                  const int* __restrict__ px, const int* __restrict__ py, int npts, double fwhm, double cy, double cx,
                  int excl_lobes, int use2alone, double* __restrict__ snr_out, double* __restrict__ flux_out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= npts) return;
    const double sx = (double)px[warp] - cx, sy = (double)py[warp] - cy;
    const double sep = sqrt(sx * sx + sy * sy);
    const double r = 0.5 * fwhm;
    if (!(sep > r)) {                                   // the reference raises; the map never asks for such pixels
        if (lane == 0) { snr_out[warp] = __longlong_as_double(0x7ff8000000000000LL); if (flux_out) flux_out[warp] = 0.0; }
        return;
    }
    const double angle = 2.0 * asin(r / sep);
    const int nap = (int)floor(2.0 * 3.14159265358979323846 / angle);
    const double fsrc = aperture_sum(img, H, W, (double)px[warp], (double)py[warp], r);   // every lane: no shuffle
    // background fluxes, shifted by the first one for a stable single-pass variance
    double shift = 0.0, sd = 0.0, sd2 = 0.0;
    int cnt = 0;
    bool have_shift = false;
    const int ntot = (img2 != nullptr) ? 2 * nap : nap;
    for (int q0 = 0; q0 < ntot; q0 += 32) {
        const int q = q0 + lane;
        double f = 0.0;
        bool use = false;
        if (q < ntot) {
            const bool second = q >= nap;
            const int j = second ? q - nap : q;
            if (!second) use = (j >= 1) && !use2alone && !(excl_lobes && (j == 1 || j == nap - 1));
            else use = !(excl_lobes && (j == 1 || j == nap - 1));
            if (use) {
                double sj, cj;
                sincos((double)j * angle, &sj, &cj);
                // clockwise (sign = -1): x' = c x + s y, y' = c y - s x
                const double ax = cj * sx + sj * sy + cx, ay = cj * sy - sj * sx + cy;
                f = aperture_sum(second ? img2 : img, H, W, ax, ay, r);
            }
        }
        if (!have_shift) {
            // first used flux of this batch (lowest lane) becomes the shift
            const unsigned m = __ballot_sync(0xffffffffu, use);
            if (m) {
                const int src = __ffs(m) - 1;
                shift = __shfl_sync(0xffffffffu, f, src);
                have_shift = true;
            }
        }
        if (use) { const double d = f - shift; sd += d; sd2 = fma(d, d, sd2); ++cnt; }
    }
    sd = warp_sum(sd);
    sd2 = warp_sum(sd2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) {
        const double n2 = (double)cnt;
        const double mean = shift + sd / n2;
        const double var = (sd2 - sd * sd / n2) / (n2 - 1.0);
        snr_out[warp] = (fsrc - mean) / (sqrt(var) * sqrt(1.0 + 1.0 / n2));
        if (flux_out) flux_out[warp] = fsrc;
    }
}

int aperture_sums(const float* img, int H, int W, const double* xs, const double* ys, int nap, double r, double* out,
                  cudaStream_t st) {
    VB_REQUIRE(H > 0 && W > 0 && nap > 0 && r > 0.0, "aperture_sums: bad arguments");
    aperture_sums_kernel<<<ceil_div(nap, 128), 128, 0, st>>>(img, H, W, xs, ys, nap, r, out);
    VB_CHECK_LAUNCH();
    return 0;
}

int snr_points(const float* img, const float* img2, int H, int W, const int* px, const int* py, int npts, double fwhm,
               double cy, double cx, int excl_lobes, int use2alone, double* snr_out, double* flux_out,
               cudaStream_t st) {
    VB_REQUIRE(H > 0 && W > 0 && npts > 0 && fwhm > 0.0, "snr_points: bad arguments");
    VB_REQUIRE(!(use2alone && img2 == nullptr), "snr_points: use2alone needs the second frame");
    snr_points_kernel<<<ceil_div(npts, 8), 256, 0, st>>>(img, img2, H, W, px, py, npts, fwhm, cy, cx, excl_lobes,
                                                         use2alone, snr_out, flux_out);
    VB_CHECK_LAUNCH();
    return 0;
}

}  // namespace vb
