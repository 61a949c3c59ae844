// Temporal collapse of a derotated cube (n, H*W) -> (H*W): median / mean / sum / max / absmean /
// wmean / trimmean, NaN-aware exactly like numpy's nan* reductions.
//
// Role in the reference: cube_collapse (src/vip_hci/preproc/subsampling.py:30-116), default
// mode 'median' = np.nanmedian(cube, axis=0) (bottleneck is optional and absent here).
//
// Layout: frame-major cube, so for a fixed frame consecutive pixels are contiguous: one thread
// owns one pixel and every load is a fully coalesced 128-byte warp transaction.
// Median = exact radix select on order-preserving uint32 keys, 4 bits per pass (8 passes),
// per-thread 16-bin histograms in shared memory (column layout -> conflict-free); an extra
// pass finds the upper middle element for even counts.  Results are bit-exact w.r.t. numpy
// (selection is exact; the mean of the two middle values is a single fp32 add and halving).
#include "common.cuh"

namespace vb {

enum CollapseMode { kMedian = 0, kMean = 1, kSum = 2, kMax = 3, kAbsMean = 4, kWMean = 5, kTrimMean = 6 };

constexpr int CT = 128;  // threads per CTA

__device__ __forceinline__ unsigned int f2key(float v) {
    const unsigned int u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned int k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// k-th smallest (0-based rank) key among the non-NaN values of this thread's pixel
__device__ __forceinline__ unsigned int radix_select(const float* __restrict__ col, int n, size_t stride,
                                                     unsigned int rank, unsigned int (*hist)[CT]) {
    unsigned int prefix = 0;
    const int tid = threadIdx.x;
#pragma unroll 1
    for (int shift = 28; shift >= 0; shift -= 4) {
#pragma unroll
        for (int b = 0; b < 16; ++b) hist[b][tid] = 0;
        const unsigned int himask = (shift == 28) ? 0u : (0xffffffffu << (shift + 4));
#pragma unroll 4
        for (int i = 0; i < n; ++i) {
            const float v = col[(size_t)i * stride];
            if (v != v) continue;
            const unsigned int key = f2key(v);
            if (((key ^ prefix) & himask) == 0u) hist[(key >> shift) & 15u][tid] += 1;
        }
        unsigned int b = 0;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const unsigned int c = hist[q][tid];
            // first bin whose cumulative count exceeds rank
            if (b == (unsigned)q) {
                if (rank >= c) { rank -= c; b = q + 1; }
            }
        }
        if (b > 15u) b = 15u;  // unreachable when rank < count
        prefix |= b << shift;
    }
    return prefix;
}

__global__ void __launch_bounds__(CT)
collapse_median_kernel(const float* __restrict__ cube, int n, size_t p, float* __restrict__ out) {
    __shared__ unsigned int hist[16][CT];
    const size_t j = (size_t)blockIdx.x * CT + threadIdx.x;
    if (j >= p) return;   // no block-level sync below: histograms are per-thread columns
    const float* col = cube + j;
    unsigned int m = 0;
    for (int i = 0; i < n; ++i) {
        const float v = col[(size_t)i * p];
        m += (v == v) ? 1u : 0u;
    }
    if (m == 0) { out[j] = __uint_as_float(0x7fc00000u); return; }
    const unsigned int r1 = (m - 1) / 2;
    const unsigned int k1 = radix_select(col, n, p, r1, hist);
    const float v1 = key2f(k1);
    if (m & 1u) { out[j] = v1; return; }
    // even count: the next order statistic is v1 again if enough duplicates, else the smallest key > k1
    unsigned int le = 0, mingt = 0xffffffffu;
    for (int i = 0; i < n; ++i) {
        const float v = col[(size_t)i * p];
        if (v != v) continue;
        const unsigned int key = f2key(v);
        if (key <= k1) ++le;
        else if (key < mingt) mingt = key;
    }
    const float v2 = (le >= r1 + 2) ? v1 : key2f(mingt);
    out[j] = (v1 + v2) * 0.5f;
}

// trimmed mean: mean of sorted[k : k+nn] with NaNs sorted last and skipped by the mean
// (subsampling.py:86-101).  sum over ranks [k, k+nn) = via two radix selects and a tie-aware pass.
__global__ void __launch_bounds__(CT)
collapse_trimmean_kernel(const float* __restrict__ cube, int n, size_t p, int k, int nn,
                         float* __restrict__ out) {
    __shared__ unsigned int hist[16][CT];
    const size_t j = (size_t)blockIdx.x * CT + threadIdx.x;
    if (j >= p) return;
    const float* col = cube + j;
    unsigned int m = 0;
    for (int i = 0; i < n; ++i) {
        const float v = col[(size_t)i * p];
        m += (v == v) ? 1u : 0u;
    }
    // ranks [lo, hi) among the valid values (NaNs occupy the top ranks of np.sort)
    const unsigned int lo = (unsigned)k;
    unsigned int hi = (unsigned)(k + nn);
    if (hi > m) hi = m;
    if (lo >= hi) { out[j] = __uint_as_float(0x7fc00000u); return; }
    const unsigned int klo = radix_select(col, n, p, lo, hist);
    const unsigned int khi = radix_select(col, n, p, hi - 1, hist);
    // sum strictly inside (klo, khi) plus the right multiplicity of the two boundary values
    double s = 0.0;
    unsigned int below_lo = 0, eq_lo = 0, below_hi = 0, eq_hi = 0;
    for (int i = 0; i < n; ++i) {
        const float v = col[(size_t)i * p];
        if (v != v) continue;
        const unsigned int key = f2key(v);
        if (key < klo) ++below_lo; else if (key == klo) ++eq_lo;
        if (key < khi) ++below_hi; else if (key == khi) ++eq_hi;
        if (key > klo && key < khi) s += (double)v;
    }
    if (klo == khi) {
        s = (double)key2f(klo) * (double)(hi - lo);
    } else {
        const unsigned int take_lo = below_lo + eq_lo - lo;   // copies of klo with rank >= lo
        const unsigned int take_hi = hi - below_hi;            // copies of khi with rank < hi
        s += (double)key2f(klo) * take_lo + (double)key2f(khi) * take_hi;
    }
    out[j] = (float)(s / (double)(hi - lo));
}

// sequential fp32 accumulation in frame order == numpy's outer-axis reduction order
template <int MODE>
__global__ void __launch_bounds__(256)
collapse_reduce_kernel(const float* __restrict__ cube, int n, size_t p, float* __restrict__ out) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    float acc = (MODE == kMax) ? __uint_as_float(0xff800000u) : 0.f;
    unsigned int cnt = 0;
    for (int i = 0; i < n; ++i) {
        float v = cube[(size_t)i * p + j];
        if (v != v) continue;
        ++cnt;
        if (MODE == kAbsMean) v = fabsf(v);
        if (MODE == kMax) acc = fmaxf(acc, v);
        else acc += v;
    }
    float r;
    if (MODE == kSum) r = acc;
    else if (MODE == kMax) r = cnt ? acc : __uint_as_float(0x7fc00000u);
    else r = cnt ? __fdiv_rn(acc, (float)cnt) : __uint_as_float(0x7fc00000u);
    out[j] = r;
}

// weighted mean: NaN -> 0, inner(w, values) accumulated in fp64, fp64 output
__global__ void __launch_bounds__(256)
collapse_wmean_kernel(const float* __restrict__ cube, int n, size_t p, const double* __restrict__ w,
                      double* __restrict__ out) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    double acc = 0.0;
    for (int i = 0; i < n; ++i) {
        const float v = cube[(size_t)i * p + j];
        if (v == v) acc = fma(w[i], (double)v, acc);
    }
    out[j] = acc;
}

// out: float[p] for every mode except kWMean (double[p]).
int collapse_f32(const float* cube, int n, size_t p, int mode, const double* w, int trim_k, int trim_n,
                 void* out, cudaStream_t st) {
    VB_REQUIRE(n > 0 && p > 0, "collapse: empty cube");
    float* fo = reinterpret_cast<float*>(out);
    const unsigned g256 = (unsigned)ceil_div(p, (size_t)256);
    switch (mode) {
        case kMedian:
            collapse_median_kernel<<<(unsigned)ceil_div(p, (size_t)CT), CT, 0, st>>>(cube, n, p, fo);
            break;
        case kMean:    collapse_reduce_kernel<kMean><<<g256, 256, 0, st>>>(cube, n, p, fo); break;
        case kSum:     collapse_reduce_kernel<kSum><<<g256, 256, 0, st>>>(cube, n, p, fo); break;
        case kMax:     collapse_reduce_kernel<kMax><<<g256, 256, 0, st>>>(cube, n, p, fo); break;
        case kAbsMean: collapse_reduce_kernel<kAbsMean><<<g256, 256, 0, st>>>(cube, n, p, fo); break;
        case kWMean:
            VB_REQUIRE(w != nullptr, "collapse: weights required for wmean");
            collapse_wmean_kernel<<<g256, 256, 0, st>>>(cube, n, p, w, reinterpret_cast<double*>(out));
            break;
        case kTrimMean:
            collapse_trimmean_kernel<<<(unsigned)ceil_div(p, (size_t)CT), CT, 0, st>>>(cube, n, p, trim_k,
                                                                                     trim_n, fo);
            break;
        default:
            VB_REQUIRE(false, "collapse: unknown mode %d", mode);
    }
    VB_CHECK_LAUNCH();
    return 0;
}

}  // namespace vb
