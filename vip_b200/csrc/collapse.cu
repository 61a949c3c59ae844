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
#include <cstdlib>
#include <cstring>

namespace vb {

enum CollapseMode { kMedian = 0, kMean = 1, kSum = 2, kMax = 3, kAbsMean = 4, kWMean = 5, kTrimMean = 6 };

constexpr int CT = 128;  // threads per CTA

__device__ __forceinline__ unsigned int f2key(float v) {
    // negative: ~u, positive: u | 0x80000000 -- as one shift and one three-input logic op
    const unsigned int u = __float_as_uint(v);
    return u ^ ((unsigned int)((int)u >> 31) | 0x80000000u);
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


// ---------------------------------------------------------------------------------------------
// Single-pass median: the n x PXT key tile of a CTA lives in shared memory, so the cube is read from
// HBM exactly once (4 bytes per pixel and frame -- the algorithmic traffic).  SUB adjacent lanes share
// one pixel (lane = pixel * SUB + s owns frames s, s+SUB, ...); every 4-bit radix pass histograms the
// surviving candidates, merges the SUB partial histograms with shuffles, and compacts the candidates
// of the chosen bin in place, so later passes touch ever fewer keys.  Ranks (m-1)/2 and m/2 are
// tracked together; when they fall in different bins the answer is (max of the lower bin, min of
// the upper bin).  NaNs are stored as key 0xFFFFFFFF (above every real key) and counted in pass 0.
// Row stride `stride` (words) satisfies stride = odd * (32 / SUB) mod 32: the 32 lanes of a warp
// (32/SUB pixels x SUB frames) always hit 32 distinct banks.
// ---------------------------------------------------------------------------------------------
template <int SUB>
__device__ __forceinline__ unsigned int group_sum(unsigned int v) {
#pragma unroll
    for (int o = 1; o < SUB; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
template <int SUB>
__device__ __forceinline__ unsigned int group_max(unsigned int v) {
#pragma unroll
    for (int o = 1; o < SUB; o <<= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int SUB>
__device__ __forceinline__ unsigned int group_min(unsigned int v) {
#pragma unroll
    for (int o = 1; o < SUB; o <<= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Load phase shared by the shared-memory median kernels: the n x pxt tile of order-preserving keys
// (NaN -> 0xffffffff, larger than every real key), row pitch `stride` words.  Ends with a CTA barrier.
__device__ __forceinline__ void median_load_tile(const float* __restrict__ cube, int n, size_t p, int pxt, int stride,
                                                 size_t px0, unsigned int* __restrict__ K) {
    const int T = blockDim.x, tid = threadIdx.x;
    // ---- load: every frame row of the tile is one contiguous pxt*4-byte segment
    if ((p & 3) == 0 && (pxt & 3) == 0 && px0 + pxt <= p) {
        const int q = pxt >> 2, total = n * q;
        // 8 independent 128-bit loads in flight per thread (one load at a time leaves HBM idle)
        constexpr int U = 8;
        for (int base = tid; base < total; base += U * T) {
            float4 v[U];
            int row[U], c4[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int idx = base + u * T;
                row[u] = idx / q;
                c4[u] = idx - row[u] * q;
                if (idx < total)
                    v[u] = ld_stream_f4(reinterpret_cast<const float4*>(cube + (size_t)row[u] * p + px0) + c4[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (base + u * T < total) {
                    uint4 k;
                    k.x = (v[u].x == v[u].x) ? f2key(v[u].x) : 0xffffffffu;
                    k.y = (v[u].y == v[u].y) ? f2key(v[u].y) : 0xffffffffu;
                    k.z = (v[u].z == v[u].z) ? f2key(v[u].z) : 0xffffffffu;
                    k.w = (v[u].w == v[u].w) ? f2key(v[u].w) : 0xffffffffu;
                    unsigned int* dst = K + (size_t)row[u] * stride + 4 * c4[u];
                    if ((stride & 3) == 0) *reinterpret_cast<uint4*>(dst) = k;       // rows stay 16-byte aligned
                    else { dst[0] = k.x; dst[1] = k.y; dst[2] = k.z; dst[3] = k.w; }
                }
            }
        }
    } else {
        const int total = n * pxt;
        for (int idx = tid; idx < total; idx += T) {
            const int row = idx / pxt, c = idx - row * pxt;
            unsigned int k = 0xffffffffu;
            if (px0 + c < p) {
                const float v = __ldg(cube + (size_t)row * p + px0 + c);
                if (v == v) k = f2key(v);
            }
            K[(size_t)row * stride + c] = k;
        }
    }
    __syncthreads();

}

template <int SUB>
__global__ void __launch_bounds__(768)
collapse_median_smem_kernel(const float* __restrict__ cube, int n, size_t p, int pxt, int stride,
                            float* __restrict__ out) {
    extern __shared__ unsigned int smem_keys[];
    const int T = blockDim.x, tid = threadIdx.x;
    unsigned int* K = smem_keys;
    // two histogram copies ([16][T] each) so that consecutive keys update independent counters
    unsigned short* __restrict__ H = reinterpret_cast<unsigned short*>(smem_keys + (size_t)n * stride);
    unsigned short* __restrict__ H2 = H + 16 * T;
    const size_t px0 = (size_t)blockIdx.x * pxt;

    median_load_tile(cube, n, p, pxt, stride, px0, K);

    // ---- select
    const int px = tid / SUB, s = tid - px * SUB;
    unsigned int* col = K + px;                       // candidate j of this thread: col[(s + SUB*j) * stride]
    int ncand = (n - s + SUB - 1) / SUB;
    if (ncand < 0) ncand = 0;
    unsigned int r1 = 0, r2 = 0, m = 0, k1 = 0, k2 = 0;
    bool done = false;
#pragma unroll 1
    for (int shift = 28; shift >= 0; shift -= 4) {
#pragma unroll
        for (int b = 0; b < 16; ++b) { H[b * T + tid] = 0; H2[b * T + tid] = 0; }
        unsigned int nnan = 0;
        if (!done) {
            int j = 0;
            for (; j + 1 < ncand; j += 2) {
                const unsigned int ka = col[(size_t)(s + SUB * j) * stride];
                const unsigned int kb = col[(size_t)(s + SUB * (j + 1)) * stride];
                H[((ka >> shift) & 15u) * T + tid] += 1;
                H2[((kb >> shift) & 15u) * T + tid] += 1;
                if (shift == 28) nnan += ((ka == 0xffffffffu) ? 1u : 0u) + ((kb == 0xffffffffu) ? 1u : 0u);
            }
            if (j < ncand) {
                const unsigned int ka = col[(size_t)(s + SUB * j) * stride];
                H[((ka >> shift) & 15u) * T + tid] += 1;
                if (shift == 28) nnan += (ka == 0xffffffffu) ? 1u : 0u;
            }
        }
        unsigned int tot[16];
#pragma unroll
        for (int b = 0; b < 16; ++b)
            tot[b] = group_sum<SUB>((unsigned int)H[b * T + tid] + (unsigned int)H2[b * T + tid]);
        if (shift == 28) {
            m = (unsigned int)n - group_sum<SUB>(nnan);
            if (m == 0) done = true;
            r1 = (m - 1) >> 1;
            r2 = m >> 1;
        }
        unsigned int cum = 0, b1 = 16, b2 = 16, c1 = 0, t1 = 0;
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            if (b1 == 16u && r1 < cum + tot[b]) { b1 = b; c1 = cum; t1 = tot[b]; }
            if (b2 == 16u && r2 < cum + tot[b]) { b2 = b; }
            cum += tot[b];
        }
        // second scan: keep the candidates of bin b1 (compaction in place), max over bin b1, min over bin b2
        unsigned int mx = 0u, mn = 0xffffffffu;
        if (!done) {
            int w = 0, j = 0;
            // two keys per step: both are read before either is written back (w <= j always)
            for (; j + 1 < ncand; j += 2) {
                const unsigned int ka = col[(size_t)(s + SUB * j) * stride];
                const unsigned int kb = col[(size_t)(s + SUB * (j + 1)) * stride];
                const unsigned int da = (ka >> shift) & 15u, db = (kb >> shift) & 15u;
                if (da == b1) { col[(size_t)(s + SUB * w) * stride] = ka; ++w; mx = max(mx, ka); }
                if (da == b2) mn = min(mn, ka);
                if (db == b1) { col[(size_t)(s + SUB * w) * stride] = kb; ++w; mx = max(mx, kb); }
                if (db == b2) mn = min(mn, kb);
            }
            if (j < ncand) {
                const unsigned int ka = col[(size_t)(s + SUB * j) * stride];
                const unsigned int da = (ka >> shift) & 15u;
                if (da == b1) { col[(size_t)(s + SUB * w) * stride] = ka; ++w; mx = max(mx, ka); }
                if (da == b2) mn = min(mn, ka);
            }
            ncand = w;
        }
        mx = group_max<SUB>(mx);
        mn = group_min<SUB>(mn);
        if (!done) {
            if (b1 != b2) { k1 = mx; k2 = mn; done = true; }
            else if (t1 == 1u || shift == 0) { k1 = mx; k2 = mx; done = true; }
            else { r1 -= c1; r2 -= c1; }
        }
        if (__all_sync(0xffffffffu, done)) break;      // warp-uniform: the shuffles above are warp-wide
    }
    if (s == 0 && px < pxt && px0 + px < p) {
        float r;
        if (m == 0) r = __uint_as_float(0x7fc00000u);
        else if (m & 1u) r = key2f(k1);
        else r = (key2f(k1) + key2f(k2)) * 0.5f;
        out[px0 + px] = r;
    }
}

// ---------------------------------------------------------------------------------------------
// Range-adaptive variant (experiment, VIP_B200_MEDIAN_ALGO=range; NOT faster: 0.64 ms vs 0.55 ms at config 2, the
// shared-memory atomics and the two extra scans per round cost more than the saved passes).  The 4-bit radix passes above spend most of their instructions on per-pass
// fixed costs (16 histogram rows to clear and to merge over the lanes, six passes for typical residuals whose
// leading key bits -- sign and high exponent bits -- barely discriminate): 100 instructions per sample
// (ncu r01l).  Here every round first takes min/max of the surviving candidates of a pixel and bins them on
// (key - min) >> sh into 256 bins that span exactly that range (16-bit counters packed in pairs, one
// shared-memory atomic per key); the lanes of the pixel scan the bins, the bin holding the wanted rank is
// compacted in place and becomes the next candidate set.  A round divides the key range by 256, so typical
// pixels finish in two rounds (500 -> a handful -> done) and at most four are ever needed.  Exact selection:
// same results as the radix kernel, bit for bit.
// ---------------------------------------------------------------------------------------------
constexpr int kHistPitch = 136;     // words per pixel: 128 packed counter pairs + 8 (bank spread between pixels)

template <int SUB>
__global__ void __launch_bounds__(768)
collapse_median_range_kernel(const float* __restrict__ cube, int n, size_t p, int pxt, int stride,
                             float* __restrict__ out) {
    extern __shared__ unsigned int smem_keys[];
    const int tid = threadIdx.x;
    unsigned int* K = smem_keys;
    unsigned int* Hs = smem_keys + (size_t)n * stride;
    const size_t px0 = (size_t)blockIdx.x * pxt;
    median_load_tile(cube, n, p, pxt, stride, px0, K);

    constexpr unsigned int FULL = 0xffffffffu;
    constexpr int W = 128 / SUB;                      // histogram words scanned per lane
    const int px = tid / SUB, s = tid - px * SUB;
    unsigned int* col = K + px;                       // candidate j of this thread: col[(s + SUB*j) * stride]
    unsigned int* hist = Hs + (size_t)px * kHistPitch;
    int ncand = (n - s + SUB - 1) / SUB;
    if (ncand < 0) ncand = 0;
    unsigned int r1 = 0, r2 = 0, m = 0, k1 = 0, k2 = 0;
    bool done = false;
#pragma unroll 1
    for (int round = 0; round < 6; ++round) {
        // (1) range of the candidates (NaN keys only exist in round 0: they never fall into a selected bin)
        unsigned int kmin = FULL, kmax = 0u, nnan = 0u;
        if (!done) {
#pragma unroll 4
            for (int j = 0; j < ncand; ++j) {
                const unsigned int k = col[(size_t)(s + SUB * j) * stride];
                if (k != FULL) { kmin = min(kmin, k); kmax = max(kmax, k); }
                else ++nnan;
            }
        }
        kmin = group_min<SUB>(kmin);
        kmax = group_max<SUB>(kmax);
        if (round == 0) {
            m = (unsigned int)n - group_sum<SUB>(nnan);
            if (m == 0) done = true;
            r1 = (m - 1) >> 1;
            r2 = m >> 1;
        }
        if (!done && kmin == kmax) { k1 = kmin; k2 = kmin; done = true; }
        if (__all_sync(FULL, done)) break;
        int sh = 24 - __clz(kmax - kmin);             // (32 - clz) - 8: (kmax - kmin) >> sh < 256
        if (sh < 0) sh = 0;
        // (2) clear, (3) fill the 256 bins of this pixel
#pragma unroll
        for (int w = 0; w < W; ++w) hist[s * W + w] = 0u;
        __syncwarp();
        if (!done) {
#pragma unroll 4
            for (int j = 0; j < ncand; ++j) {
                const unsigned int k = col[(size_t)(s + SUB * j) * stride];
                if (k != FULL) {
                    const unsigned int d = (k - kmin) >> sh;
                    atomicAdd(&hist[d >> 1], 1u << ((d & 1u) << 4));
                }
            }
        }
        __syncwarp();
        // (4) bins of the two wanted ranks: lane s scans bins [2 W s, 2 W (s+1))
        unsigned int hw[W];
        unsigned int loc = 0u;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            hw[w] = hist[s * W + w];
            loc += (hw[w] & 0xffffu) + (hw[w] >> 16);
        }
        unsigned int incl = loc;
#pragma unroll
        for (int o = 1; o < SUB; o <<= 1) {
            const unsigned int t = __shfl_up_sync(FULL, incl, o, SUB);
            if (s >= o) incl += t;
        }
        const unsigned int excl = incl - loc;
        unsigned int pb = 0u, pc = 0u;                // owner lanes fill: pb = b1 | b2 << 8 | flags, pc = c1 | t1 << 16
        if (!done) {
            const bool own1 = r1 >= excl && r1 < incl, own2 = r2 >= excl && r2 < incl;
            if (own1 || own2) {
                unsigned int cum = excl;
#pragma unroll
                for (int w = 0; w < W; ++w) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const unsigned int c = h ? (hw[w] >> 16) : (hw[w] & 0xffffu);
                        const unsigned int bin = (unsigned int)(2 * (s * W + w) + h);
                        if (own1 && r1 >= cum && r1 < cum + c) { pb |= bin; pc = cum | (c << 16); }
                        if (own2 && r2 >= cum && r2 < cum + c) { pb |= bin << 8; }
                        cum += c;
                    }
                }
            }
        }
        pb = group_sum<SUB>(pb);                      // exactly one lane contributes each field
        pc = group_sum<SUB>(pc);
        const unsigned int b1 = pb & 0xffu, b2 = (pb >> 8) & 0xffu, c1 = pc & 0xffffu, t1 = pc >> 16;
        // (5) keep bin b1 (compaction in place), max over bin b1, min over bin b2
        unsigned int mx = 0u, mn = FULL;
        if (!done) {
            int w = 0;
#pragma unroll 4
            for (int j = 0; j < ncand; ++j) {
                const unsigned int k = col[(size_t)(s + SUB * j) * stride];
                if (k == FULL) continue;
                const unsigned int d = (k - kmin) >> sh;
                if (d == b1) { col[(size_t)(s + SUB * w) * stride] = k; ++w; mx = max(mx, k); }
                if (d == b2) mn = min(mn, k);
            }
            ncand = w;
        }
        mx = group_max<SUB>(mx);
        mn = group_min<SUB>(mn);
        if (!done) {
            if (b1 != b2) { k1 = mx; k2 = mn; done = true; }
            else if (t1 == 1u || sh == 0) { k1 = mx; k2 = mx; done = true; }
            else { r1 -= c1; r2 -= c1; }
        }
        __syncwarp();
        if (__all_sync(FULL, done)) break;
    }
    if (s == 0 && px < pxt && px0 + px < p) {
        float r;
        if (m == 0) r = __uint_as_float(0x7fc00000u);
        else if (m & 1u) r = key2f(k1);
        else r = (key2f(k1) + key2f(k2)) * 0.5f;
        out[px0 + px] = r;
    }
}

// ---------------------------------------------------------------------------------------------
// Warp-per-pixel median (round 2, default for 64 <= n <= 1024): keys of a pixel live in REGISTERS.
// The radix kernels above spend ~100 instructions per sample (ncu r01l/r02k: issue-bound, 14 % of the HBM rate):
// every 4-bit pass re-reads the surviving keys from shared memory, updates 16-bit histogram cells with
// read-modify-write sequences and merges 16 bins over the lanes, and the leading key bits (sign, high exponent) hardly
// discriminate, so about three passes touch every key.  Here a warp owns one pixel: lane l holds frames l, l + 32, ...
// (KPL registers), read once from the shared-memory tile, and the two middle ranks are BRACKETED by counting:
//   * a 32-key sample (one per lane) is sorted across the warp (bitonic, shuffles); its median is the first pivot and
//     its order statistics guide the next ones: idx += (r - c) * 32 / m, overshooting by 1..3 sample ranks so that
//     the target gets bracketed from both sides;
//   * a pass counts the keys below the pivot (compare + predicated add per key, one REDUX for the warp) and moves the
//     lower or the upper end of the bracket [lo, hi) -- counts (clo, chi) are exact, so the search never loses the
//     ranks; once both ends come from counts the pivot is interpolated between them in VALUE space;
//   * when at most 32 keys are left inside the bracket they are compacted (ballot + popc) into one key per lane,
//     sorted, and ranks r1 - clo, r2 - clo are read off with a shuffle;
//   * exactness does not depend on the guesses: every pivot lies strictly inside (lo, hi) so the bracket shrinks on
//     every pass, every third pass after the sixth halves it in KEY space (<= 32 such passes to one key), heavy ties
//     end at hi - lo == 1 or at the first pass (which also counts keys <= pivot when the sample has ties).
// Pass counts of this logic on 500-sample columns (tools/median_bracket_model.py, the same code in numpy checked bit
// for bit against np.nanmedian): 3.4 on Gaussian / heavy-tailed / outlier-ridden data, 2 with 60 % zeros.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned int warp_sort32(unsigned int v, int lane) {
#pragma unroll
    for (int kk = 2; kk <= 32; kk <<= 1) {
#pragma unroll
        for (int jj = kk >> 1; jj > 0; jj >>= 1) {
            const unsigned int o = __shfl_xor_sync(0xffffffffu, v, jj);
            const bool keep_min = ((lane & jj) == 0) == ((lane & kk) == 0);
            v = keep_min ? min(v, o) : max(v, o);
        }
    }
    return v;
}

// c += (k < pv) as compare + predicated add (what the loop is meant to cost: ncu r02z showed the C++ form compiled to
// 4.75 instructions per key -- the compiler also packed every compare result into a bit mask for later reuse)
__device__ __forceinline__ void count_below(unsigned int& c, unsigned int k, unsigned int pv) {
    asm("{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %1, %2;\n\t@q add.u32 %0, %0, 1;\n\t}" : "+r"(c) : "r"(k), "r"(pv));
}

template <int KPL, int PXT>
__global__ void __launch_bounds__(256, (KPL <= 16) ? 3 : 2)
collapse_median_warp_kernel(const float* __restrict__ cube, int n, size_t p, float* __restrict__ out) {
    constexpr int STRIDE = PXT + 1, ROWS = KPL * 32, Q = PXT / 4, RSTEP = 256 / Q;
    constexpr unsigned int PAD = 0xffffffffu, FULL = 0xffffffffu;
    extern __shared__ unsigned int smem_keys[];
    unsigned int* K = smem_keys;                                       // [ROWS][STRIDE], rows >= n hold PAD
    unsigned int* cand_all = smem_keys + (size_t)ROWS * STRIDE;        // [8 warps][32]
    int* nanflag = reinterpret_cast<int*>(cand_all + 8 * 32);         // [PXT]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t px0 = (size_t)blockIdx.x * PXT;

    if (tid < PXT) nanflag[tid] = 0;
    for (int e = tid; e < (ROWS - n) * PXT; e += 256) K[(size_t)(n + e / PXT) * STRIDE + e % PXT] = PAD;
    if ((p & 3) == 0 && px0 + PXT <= p && (reinterpret_cast<size_t>(cube) & 15) == 0) {
        // fast path: thread = (row group r, float4 column c4); rows r, r + RSTEP, ...; 8 loads in flight
        __syncthreads();
        const int c4 = tid % Q, r = tid / Q;
        const float* src = cube + px0 + 4 * c4;
        bool anynan = false;
        for (int row0 = r; row0 < n; row0 += 8 * RSTEP) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int row = row0 + u * RSTEP;
                if (row < n) v[u] = ld_stream_f4(reinterpret_cast<const float4*>(src + (size_t)row * p));
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int row = row0 + u * RSTEP;
                if (row < n) {
                    unsigned int* dst = K + (size_t)row * STRIDE + 4 * c4;
                    const bool n0 = v[u].x != v[u].x, n1 = v[u].y != v[u].y, n2 = v[u].z != v[u].z,
                               n3 = v[u].w != v[u].w;
                    dst[0] = n0 ? PAD : f2key(v[u].x);
                    dst[1] = n1 ? PAD : f2key(v[u].y);
                    dst[2] = n2 ? PAD : f2key(v[u].z);
                    dst[3] = n3 ? PAD : f2key(v[u].w);
                    anynan |= n0 | n1 | n2 | n3;
                }
            }
        }
        if (anynan) {                                 // conservative: the flag only selects the NaN-counting path
#pragma unroll
            for (int t = 0; t < 4; ++t) nanflag[4 * c4 + t] = 1;
        }
        __syncthreads();
    } else {
        __syncthreads();
        if (tid < PXT) nanflag[tid] = 1;                              // generic path: count NaNs for every pixel
        median_load_tile(cube, n, p, PXT, STRIDE, px0, K);
    }

    unsigned int* cand = cand_all + warp * 32;
    for (int px = warp; px < PXT; px += 8) {
        if (px0 + px >= p) break;
        const unsigned int* col = K + px + (size_t)lane * STRIDE;
        unsigned int k[KPL];
#pragma unroll
        for (int j = 0; j < KPL; ++j) k[j] = col[(size_t)(32 * j) * STRIDE];
        auto count_lt = [&](unsigned int pv) {
            unsigned int c0 = 0, c1 = 0;
#pragma unroll
            for (int j = 0; j < KPL; j += 2) { count_below(c0, k[j], pv); count_below(c1, k[j + 1], pv); }
            return __reduce_add_sync(FULL, c0 + c1);
        };
        unsigned int m = (unsigned int)n;
        if (nanflag[px]) {                                 // NaNs in this pixel (border of the rotated frames)
            unsigned int nn = 0;
#pragma unroll
            for (int j = 0; j < KPL; ++j) nn += (k[j] == PAD && lane + 32 * j < n) ? 1u : 0u;
            m -= __reduce_add_sync(FULL, nn);
        }
        float result;
        if (m == 0) {
            result = __uint_as_float(0x7fc00000u);
        } else {
            const unsigned int r1 = (m - 1) >> 1, r2 = m >> 1;
            const unsigned int s = warp_sort32(k[0], lane);            // n >= 64: row `lane` is a real frame
            const int ns = __popc(__ballot_sync(FULL, s != PAD));
            unsigned int lo = 0u, hi = PAD, clo = 0u, chi = m, k1 = 0u, k2 = 0u;
            bool have_lo = false, have_hi = false, first = true;
            int idx = ns >> 1, it = 0, guard = 0;
            unsigned int pv = (ns > 0) ? __shfl_sync(FULL, s, idx) : (PAD >> 1);
            for (;;) {
                if (++guard > 400) __trap();                            // impossible by construction: fail loudly
                const unsigned int c = count_lt(pv);
                if (first && ns > 1) {
                    const unsigned int sa = __shfl_sync(FULL, s, min(idx + 1, ns - 1));
                    const unsigned int sb = __shfl_sync(FULL, s, max(idx - 1, 0));
                    if (sa == pv || sb == pv) {                         // ties around the sample median
                        const unsigned int cle = count_lt(pv + 1u);
                        if (c <= r1 && r2 < cle) { k1 = k2 = pv; break; }
                    }
                }
                first = false;
                if (c <= r1) { lo = pv; clo = c; have_lo = true; }
                else if (c > r2) { hi = pv; chi = c; have_hi = true; }
                else {
                    // c == r2 == r1 + 1: the pivot separates the two middle ranks
                    // (compared with pv - 1 on purpose: written as `k < pv` the assembler keeps all KPL compare results
                    // of every counting pass alive in a bit mask for this rarely taken block -- 2 extra instructions
                    // per key and pass, ncu r02z)
                    unsigned int a = 0u, b = PAD;
                    const unsigned int pm = pv - 1u;                    // pv > 0: c >= 1 keys lie below it
#pragma unroll
                    for (int j = 0; j < KPL; ++j) {
                        if (k[j] <= pm) a = max(a, k[j]); else b = min(b, k[j]);
                    }
                    k1 = __reduce_max_sync(FULL, a);
                    k2 = __reduce_min_sync(FULL, b);
                    break;
                }
                if (chi - clo <= 32u) {
                    // the keys of [lo, hi): per-lane count, exclusive scan over the lanes, compacted into one key per
                    // lane, sorted; the two ranks are read off with a shuffle
                    const unsigned int width = hi - lo;
                    unsigned int cnt = 0;
#pragma unroll
                    for (int j = 0; j < KPL; ++j) count_below(cnt, k[j] - lo, width);
                    unsigned int off = cnt;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const unsigned int o = __shfl_up_sync(FULL, off, d);
                        if (lane >= d) off += o;
                    }
                    off -= cnt;
                    if (cnt) {
#pragma unroll
                        for (int j = 0; j < KPL; ++j)
                            if ((k[j] - lo) < width) cand[off++] = k[j];
                    }
                    __syncwarp();
                    unsigned int v = ((unsigned int)lane < chi - clo) ? cand[lane] : PAD;
                    __syncwarp();
                    v = warp_sort32(v, lane);
                    k1 = __shfl_sync(FULL, v, (int)(r1 - clo));
                    k2 = __shfl_sync(FULL, v, (int)(r2 - clo));
                    break;
                }
                if (hi - lo <= 1u) { k1 = k2 = lo; break; }              // every key left equals lo
                ++it;
                const bool use_mid = it >= 6 && (it % 3) == 0;
                unsigned int pn = 0u;
                bool ok = false;
                if (!use_mid && !(have_lo && have_hi) && ns > 0) {
                    const int up = c <= r1;
                    const int d = up ? (int)(((r1 - c) * (unsigned int)ns) / m)
                                     : -(int)(((c - r1) * (unsigned int)ns) / m);
                    const int g = min(it, 3);
                    idx = min(max(idx + (up ? d + g : d - g), 0), ns - 1);
                    pn = __shfl_sync(FULL, s, idx);
                    ok = true;
                } else if (!use_mid && have_lo && have_hi) {
                    const float flo = key2f(lo), fhi = key2f(hi);
                    float t = ((float)r1 + 0.5f - (float)clo) / (float)(chi - clo);
                    t = fminf(fmaxf(t, 0.15f), 0.85f);
                    const float pf = flo + (fhi - flo) * t;
                    ok = pf == pf;
                    pn = f2key(pf);
                }
                pv = (ok && pn > lo && pn < hi) ? pn : lo + ((hi - lo) >> 1);
            }
            const float a = key2f(k1), b = key2f(k2);
            result = (m & 1u) ? a : (a + b) * 0.5f;
        }
        if (lane == 0) out[px0 + px] = result;
    }
}

struct MedianCfg { int sub, pxt, stride, threads; size_t smem; };

// Pick (SUB, pixel tile, padded stride) maximising resident threads per SM; returns false when n is too
// large for a shared-memory tile (the multi-pass kernel handles those).
static bool pick_median_cfg(int n, MedianCfg* best, int range_variant) {
    const size_t smem_max = 227 * 1024 - 1024;
    // lanes per pixel: the per-pass histogram merge costs 16*log2(SUB) shuffles per lane, the scan
    // n/SUB keys -- keep the scan the larger part (about 64 keys per lane) unless n forces more lanes
    int sub0 = 4;
    while (sub0 < 32 && n / sub0 > 64) sub0 <<= 1;
    // development override: VIP_B200_MEDIAN_CFG="sub,pxt"
    if (const char* e = getenv("VIP_B200_MEDIAN_CFG")) {
        int sub = 0, pxt = 0;
        if (sscanf(e, "%d,%d", &sub, &pxt) == 2 && (sub == 4 || sub == 8 || sub == 16 || sub == 32) && pxt > 0) {
            const int g = 32 / sub;
            const int stride = ((pxt / g) & 1) ? pxt : pxt + g;
            const int threads = pxt * sub;
            const size_t smem = (size_t)n * stride * 4 +
                                (range_variant ? (size_t)pxt * kHistPitch * 4 : (size_t)64 * threads);
            if (pxt % g == 0 && threads <= 768 && (threads & 31) == 0 && smem <= smem_max) {
                *best = MedianCfg{sub, pxt, stride, threads, smem};
                return true;
            }
        }
    }
    double best_score = 0.0;
    for (int sub = sub0; sub <= 32; sub <<= 1) {
        const int g = 32 / sub;
        for (int pxt = 96; pxt >= 8; pxt -= 8) {
            const int stride = ((pxt / g) & 1) ? pxt : pxt + g;
            const int threads = pxt * sub;
            if (threads > 768 || (threads & 31)) continue;
            const size_t smem = (size_t)n * stride * 4 +
                                (range_variant ? (size_t)pxt * kHistPitch * 4 : (size_t)64 * threads);
            if (smem > smem_max) continue;
            int ctas = (int)((227 * 1024) / (smem + 1024));
            if (ctas * threads > 2048) ctas = 2048 / threads;
            if (ctas > 4) ctas = 4;
            // resident pixels per SM decide the throughput; >= 2 CTAs/SM lets one CTA load while another selects
            double score = (double)ctas * pxt;
            if (ctas >= 2) score *= 1.25;
            if (pxt < 16) score *= 0.7;
            if (score > best_score) { best_score = score; *best = MedianCfg{sub, pxt, stride, threads, smem}; }
        }
        if (best_score > 0.0) break;      // the smallest SUB that fits wins
    }
    return best_score > 0.0;
}

static int launch_median_smem(const float* cube, int n, size_t p, float* out, const MedianCfg& c, int range_variant,
                              cudaStream_t st) {
    const unsigned grid = (unsigned)ceil_div(p, (size_t)c.pxt);
#define VB_MEDIAN_CASE(S)                                                                              \
    case S: {                                                                                          \
        static bool configured = false;                                                                \
        if (!configured) {                                                                             \
            VB_CHECK_CUDA(cudaFuncSetAttribute(collapse_median_smem_kernel<S>,                         \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024))); \
            VB_CHECK_CUDA(cudaFuncSetAttribute(collapse_median_range_kernel<S>,                        \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024))); \
            configured = true;                                                                         \
        }                                                                                              \
        if (range_variant)                                                                             \
            collapse_median_range_kernel<S><<<grid, c.threads, c.smem, st>>>(cube, n, p, c.pxt, c.stride, out); \
        else                                                                                           \
            collapse_median_smem_kernel<S><<<grid, c.threads, c.smem, st>>>(cube, n, p, c.pxt, c.stride, out);  \
        break;                                                                                         \
    }
    switch (c.sub) {
        VB_MEDIAN_CASE(4)
        VB_MEDIAN_CASE(8)
        VB_MEDIAN_CASE(16)
        VB_MEDIAN_CASE(32)
        default: VB_REQUIRE(false, "collapse: bad median configuration");
    }
#undef VB_MEDIAN_CASE
    VB_CHECK_LAUNCH();
    return 0;
}

static int launch_median_warp(const float* cube, int n, size_t p, float* out, cudaStream_t st) {
    // pixel tile: 32 pixels (128-byte row segments, KPL <= 16) or 16 (n > 512: 32 keys per lane); the row pitch
    // PXT + 1 is odd = conflict-free column reads (lane l reads row l + 32 j: bank (l + px) mod 32); the tile is
    // padded to KPL * 32 rows so that key loads need no bounds test
#define VB_MEDIAN_WARP_CASE(KPL, PXT)                                                                         \
    {                                                                                                         \
        const size_t smem = (size_t)(KPL * 32) * (PXT + 1) * 4 + 8 * 32 * 4 + PXT * 4;                        \
        static bool configured = false;                                                                       \
        if (!configured) {                                                                                    \
            VB_CHECK_CUDA(cudaFuncSetAttribute(collapse_median_warp_kernel<KPL, PXT>,                         \
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));      \
            VB_CHECK_CUDA(cudaFuncSetAttribute(collapse_median_warp_kernel<KPL, PXT>,                         \
                                               cudaFuncAttributePreferredSharedMemoryCarveout, 100));         \
            configured = true;                                                                                \
        }                                                                                                     \
        collapse_median_warp_kernel<KPL, PXT><<<(unsigned)ceil_div(p, (size_t)PXT), 256, smem, st>>>(cube, n, p, out); \
    }
    if (n <= 128) VB_MEDIAN_WARP_CASE(4, 32)
    else if (n <= 256) VB_MEDIAN_WARP_CASE(8, 32)
    else if (n <= 512) VB_MEDIAN_WARP_CASE(16, 32)
    else VB_MEDIAN_WARP_CASE(32, 16)
#undef VB_MEDIAN_WARP_CASE
    VB_CHECK_LAUNCH();
    return 0;
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
        case kMedian: {
            MedianCfg cfg;
            const char* e = getenv("VIP_B200_MEDIAN_MULTIPASS");
            // default: 4-bit radix kernel; VIP_B200_MEDIAN_ALGO=range selects the range-adaptive 256-bin rounds
            // (measured slower on config 2: 0.64 vs 0.55 ms, profiles/r01o_ab.md -- kept as a tested experiment)
            const char* a = getenv("VIP_B200_MEDIAN_ALGO");
            const int range_variant = (a && strcmp(a, "range") == 0) ? 1 : 0;
            // default for 64 <= n <= 1024: warp-per-pixel bracket search on register-resident keys
            // (VIP_B200_MEDIAN_ALGO=radix keeps the 4-bit radix kernel)
            if (!(e && atoi(e)) && !(a && (strcmp(a, "radix") == 0 || strcmp(a, "range") == 0)) && n >= 64 &&
                n <= 1024)
                return launch_median_warp(cube, n, p, fo, st);
            if (!(e && atoi(e)) && pick_median_cfg(n, &cfg, range_variant))
                return launch_median_smem(cube, n, p, fo, cfg, range_variant, st);
            collapse_median_kernel<<<(unsigned)ceil_div(p, (size_t)CT), CT, 0, st>>>(cube, n, p, fo);
            break;
        }
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
