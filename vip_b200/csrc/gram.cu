// Gramian of the frames x pixels matrix, accumulated to fp64 (CUDA-core path).
//
// Role in the reference: the O(n^2 p) part of the PCA decomposition -- numpy's fp64 SVD of
// M^T for svd_mode='lapack' (src/vip_hci/psfsub/svd.py:466-475) or  C = M.M^T  for 'eigen'
// (svd.py:447-450).  numpy always decomposes in fp64 and a Gramian squares the condition
// number (sigma_0/sigma_k ~ 1e3 on halo-dominated ADI cubes), so G is accumulated in fp64:
// the fp32 inputs are widened in registers, products are exact, every K-chunk of a tile is
// reduced with fp64 atomics.  (A first version with fp32 FMA chains of 4096 + fp64 chunk sums
// measured 1.1e-4 parity on the GPU -- over the 1e-4 budget; fp64 measured <2e-5.)
// Optional mean deflation (M = 1 m^T + D, G = D D^T + rank-2 terms in fp64) is kept for
// callers that want it; it is not needed for accuracy any more.
#include "common.cuh"
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

namespace vb {

constexpr int GT = 128;   // output tile (GT x GT)
constexpr int GK = 16;    // k-slab per smem stage
constexpr int GLD = GT + 4;

// temporal mean per pixel (fp64 accumulate -> fp32), one thread per pixel, coalesced over p
__global__ void colmean_kernel(const float* __restrict__ A, int n, size_t p, float* __restrict__ mean) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= p) return;
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += (double)A[(size_t)i * p + j];
    mean[j] = (float)(s / n);
}

// C[na x nb] (fp64, atomics) += (A - 1 m^T)[rows] . (B - 1 m^T)[rows]^T over the K-chunk of this CTA.
// tiles: list of (ti, tj) output tiles; blockIdx.y = K-chunk.
__global__ void __launch_bounds__(256)
gram_tile_kernel(const float* __restrict__ A, const float* __restrict__ B, int na, int nb, size_t p,
                 size_t ld, const float* __restrict__ mean, double* __restrict__ C, int ldc,
                 const int2* __restrict__ tiles, int kchunk) {   // p = K extent, ld = row stride
    __shared__ __align__(16) float As[2][GK][GLD];
    __shared__ __align__(16) float Bs[2][GK][GLD];
    const int2 tile = tiles[blockIdx.x];
    const int row0 = tile.x * GT, col0 = tile.y * GT;
    const size_t k0 = (size_t)blockIdx.y * kchunk;
    const size_t k1 = (k0 + kchunk < p) ? k0 + kchunk : p;
    const int tid = threadIdx.x;
    const int lk = tid & 15, lr = tid >> 4;   // loader: k offset, first row
    const int ty = tid >> 4, tx = tid & 15;   // compute: 8x8 micro-tile at (ty*8, tx*8)

    // fp64 accumulators: fp32 x fp32 products are exact in fp64, so G carries only the final
    // 1e-16 rounding -- B200 issues DFMA at half the FFMA rate, which this path can afford
    double acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;

    float ra[8], rb[8];
    auto gload = [&](size_t kk) {
        const size_t k = kk + lk;
        const bool kin = k < k1;
        const float m = (mean != nullptr && kin) ? __ldg(mean + k) : 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = lr + 16 * i;
            ra[i] = (kin && row0 + r < na) ? __ldg(A + (size_t)(row0 + r) * ld + k) - m : 0.f;
            rb[i] = (kin && col0 + r < nb) ? __ldg(B + (size_t)(col0 + r) * ld + k) - m : 0.f;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            As[buf][lk][lr + 16 * i] = ra[i];
            Bs[buf][lk][lr + 16 * i] = rb[i];
        }
    };

    gload(k0);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (size_t kk = k0; kk < k1; kk += GK) {
        const bool more = kk + GK < k1;
        if (more) gload(kk + GK);
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 8 + 4]);
            const double av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const double bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            sstore(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = row0 + ty * 8 + i;
        if (r >= na) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = col0 + tx * 8 + j;
            if (c < nb) atomicAdd(C + (size_t)r * ldc + c, acc[i][j]);
        }
    }
}

// Second-generation tile kernel (default; VIP_B200_GRAM_V1=1 selects the kernel above).
// The first kernel keeps fp32 operands in shared memory and widens them in every thread that uses them: 16
// F2F.F64.F32 per 64 DFMA per k-step, and on B200 the conversion pipe is as narrow as the fp64 pipe -- the ncu
// launch list of the config-5 slice (profiles/r01m_launches_c5.md) shows it at 31 % of the DFMA peak.  Here every
// operand is widened ONCE, when it is stored to shared memory (16 conversions per thread per 16-deep slab), and
// the inner loop is LDS.128 + DFMA only.  The B micro-tile is interleaved (column pairs tx*2 + 32 h) so that the
// sixteen 16-byte reads of a half warp are contiguous (no bank conflicts); the A micro-tile is a broadcast.
// TA = rows of the A tile: 128 for square Gramian tiles, 64 for skinny products (the (ncomp+10) x n sketches of
// the randomized SVD and the n x ncomp projection coefficients, which would waste half of a 128-row tile).
// trans = 1 writes C^T (used to put the skinny operand on the A side whichever argument it is).
template <int TA>
__global__ void __launch_bounds__(256, (TA == 64) ? 2 : 1)
gram_tile2_kernel(const float* __restrict__ A, const float* __restrict__ B, int na, int nb, size_t p,
                  size_t ld, const float* __restrict__ mean, double* __restrict__ C, int ldc, int trans,
                  const int2* __restrict__ tiles, int kchunk) {
    constexpr int MI = TA / 16;                 // micro-tile rows per thread
    constexpr int LDA = TA + 2, LDB = GT + 2;   // even pitches keep the 16-byte alignment of the double2 reads
    extern __shared__ __align__(16) double gsm[];
    double* As = gsm;                           // [2][GK][LDA]
    double* Bs = gsm + 2 * GK * LDA;            // [2][GK][LDB]
    const int2 tile = tiles[blockIdx.x];
    const int row0 = tile.x * TA, col0 = tile.y * GT;
    const size_t k0 = (size_t)blockIdx.y * kchunk;
    const size_t k1 = (k0 + kchunk < p) ? k0 + kchunk : p;
    const int tid = threadIdx.x;
    const int lk = tid & 15, lr = tid >> 4;     // loader: k offset, first row
    const int ty = tid >> 4, tx = tid & 15;     // compute: rows ty*MI.., column pairs tx*2 + 32 h

    double acc[MI][8];
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;

    float ra[MI], rb[8];
    auto gload = [&](size_t kk) {
        const size_t k = kk + lk;
        const bool kin = k < k1;
        const float m = (mean != nullptr && kin) ? __ldg(mean + k) : 0.f;
#pragma unroll
        for (int i = 0; i < MI; ++i) {
            const int r = lr + 16 * i;
            ra[i] = (kin && row0 + r < na) ? __ldg(A + (size_t)(row0 + r) * ld + k) - m : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = lr + 16 * i;
            rb[i] = (kin && col0 + r < nb) ? __ldg(B + (size_t)(col0 + r) * ld + k) - m : 0.f;
        }
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < MI; ++i) As[(buf * GK + lk) * LDA + lr + 16 * i] = (double)ra[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) Bs[(buf * GK + lk) * LDB + lr + 16 * i] = (double)rb[i];
    };

    gload(k0);
    sstore(0);
    __syncthreads();
    int buf = 0;
    for (size_t kk = k0; kk < k1; kk += GK) {
        const bool more = kk + GK < k1;
        if (more) gload(kk + GK);
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            const double* arow = As + (buf * GK + k) * LDA + ty * MI;
            const double* brow = Bs + (buf * GK + k) * LDB + tx * 2;
            double av[MI], bv[8];
#pragma unroll
            for (int h = 0; h < MI / 2; ++h) {
                const double2 a2 = *reinterpret_cast<const double2*>(arow + 2 * h);
                av[2 * h] = a2.x; av[2 * h + 1] = a2.y;
            }
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                const double2 b2 = *reinterpret_cast<const double2*>(brow + 32 * h);
                bv[2 * h] = b2.x; bv[2 * h + 1] = b2.y;
            }
#pragma unroll
            for (int i = 0; i < MI; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            sstore(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int i = 0; i < MI; ++i) {
        const int r = row0 + ty * MI + i;
        if (r >= na) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = col0 + (j >> 1) * 32 + tx * 2 + (j & 1);
            if (c < nb) atomicAdd(trans ? C + (size_t)c * ldc + r : C + (size_t)r * ldc + c, acc[i][j]);
        }
    }
}

template <int TA>
static size_t gram_tile2_smem() { return (size_t)2 * GK * ((TA + 2) + (GT + 2)) * sizeof(double); }

static bool gram_use_v1() {
    static const int v1 = [] { const char* e = getenv("VIP_B200_GRAM_V1"); return (e && atoi(e)) ? 1 : 0; }();
    return v1 != 0;
}

// one launch of the tile kernel (ta = 128 or 64 rows per A tile; tiles were listed for that height)
static int launch_gram_tiles(int ta, const float* A, const float* B, int na, int nb, size_t p, size_t ld,
                             const float* mean, double* C, int ldc, int trans, const int2* tiles, int ntiles,
                             unsigned nchunks, int kchunk, cudaStream_t st) {
    if (ta == 128 && !trans && gram_use_v1()) {
        gram_tile_kernel<<<dim3(ntiles, nchunks), 256, 0, st>>>(A, B, na, nb, p, ld, mean, C, ldc, tiles, kchunk);
    } else if (ta == 128) {
        static bool attr = false;
        if (!attr) {
            VB_CHECK_CUDA(cudaFuncSetAttribute(gram_tile2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)gram_tile2_smem<128>()));
            attr = true;
        }
        gram_tile2_kernel<128><<<dim3(ntiles, nchunks), 256, gram_tile2_smem<128>(), st>>>(
            A, B, na, nb, p, ld, mean, C, ldc, trans, tiles, kchunk);
    } else {
        static bool attr = false;
        if (!attr) {
            VB_CHECK_CUDA(cudaFuncSetAttribute(gram_tile2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               (int)gram_tile2_smem<64>()));
            attr = true;
        }
        gram_tile2_kernel<64><<<dim3(ntiles, nchunks), 256, gram_tile2_smem<64>(), st>>>(
            A, B, na, nb, p, ld, mean, C, ldc, trans, tiles, kchunk);
    }
    VB_CHECK_LAUNCH();
    return 0;
}

// Dm[i] = sum_k (A[i,k] - m[k]) m[k]   (blocks 0..n-1),   mm = sum_k m[k]^2  (block n)
__global__ void __launch_bounds__(256)
deflate_terms_kernel(const float* __restrict__ A, int n, size_t p, const float* __restrict__ mean,
                     double* __restrict__ Dm, double* __restrict__ mm) {
    const int i = blockIdx.x;
    double s = 0.0;
    if (i < n) {
        const float* row = A + (size_t)i * p;
        for (size_t k = threadIdx.x; k < p; k += blockDim.x) {
            const float m = mean[k];
            s += (double)(row[k] - m) * (double)m;
        }
    } else {
        for (size_t k = threadIdx.x; k < p; k += blockDim.x) {
            const double m = mean[k];
            s += m * m;
        }
    }
    __shared__ double red[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += red[w];
        if (i < n) Dm[i] = t; else *mm = t;
    }
}

// G = sym(Gd upper tiles) [+ deflation terms]
__global__ void gram_assemble_kernel(const double* __restrict__ Gd, int n, const double* __restrict__ Dm,
                                     const double* __restrict__ mm, double* __restrict__ G) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= n) return;
    // only tiles with tile_row <= tile_col were accumulated; mirror element-wise so that G is exactly
    // symmetric (the two triangles of a diagonal tile are accumulated in different orders)
    double v = (i <= j) ? Gd[(size_t)i * n + j] : Gd[(size_t)j * n + i];
    if (Dm != nullptr) v += Dm[i] + Dm[j] + *mm;
    G[(size_t)i * n + j] = v;
}

int gram_assemble(const double* Gd, int n, const double* Dm, const double* mm, double* G, cudaStream_t st) {
    gram_assemble_kernel<<<dim3(ceil_div(n, 128), n), 128, 0, st>>>(Gd, n, Dm, mm, G);
    VB_CHECK_LAUNCH();
    return 0;
}

// tensor-core path (gram_tc.cu)
size_t gram_tc_workspace_bytes(int n, size_t p);
bool gram_tc_eligible(int n, size_t p);
int gram_tc_begin(int n, size_t p, void* ws, size_t ws_bytes, int* ntiles_out, cudaStream_t st);
int gram_tc_accumulate(const float* A, int n, size_t p, size_t ld, size_t c0, size_t c1, void* ws, int ntiles,
                       int* launches, cudaStream_t st);
int gram_tc_finish(int n, size_t p, void* ws, double* G, int* launches, cudaStream_t st);
int gram_tc_f32(const float* A, int n, size_t p, double* G, void* ws, size_t ws_bytes, int* launches,
                cudaStream_t st);

// Split-K chunk so that (#tiles x #chunks) fills whole waves of the 148 SMs (1 CTA/SM: 250 registers)
static int pick_kchunk(size_t p, int ntiles, int target) {
    const int base = (int)ceil_div(p, (size_t)target);
    int best = base;
    double best_eff = 0.0;
    for (int c = base; c <= base + 64; ++c) {
        const long long total = (long long)ntiles * c;
        const long long waves = (total + kNumSMs - 1) / kNumSMs;
        const double eff = (double)total / (double)(waves * kNumSMs);
        if (eff > best_eff + 1e-9) { best_eff = eff; best = c; }
        if (eff > 0.985) break;
    }
    return ceil_div((int)ceil_div(p, (size_t)best), GK) * GK;
}

size_t gram_workspace_bytes(int n, size_t p) {
    if (gram_tc_eligible(n, p)) return gram_tc_workspace_bytes(n, p);
    const int nt = ceil_div(n, GT);
    size_t b = 0;
    b += ((p * sizeof(float) + 255) / 256) * 256;                  // mean
    b += (((size_t)n * n * sizeof(double) + 255) / 256) * 256;     // Gd
    b += (((size_t)n + 1) * sizeof(double) + 255) / 256 * 256;     // Dm, mm
    b += ((size_t)nt * nt * sizeof(int2) + 255) / 256 * 256;       // tile list
    return b;
}

// G (n x n fp64, row-major) = A A^T, optionally through the mean-deflated formulation.
int gram_f32(const float* A, int n, size_t p, int deflate, double* G, void* ws, size_t ws_bytes,
             int kchunk, int* launches, cudaStream_t st) {
    VB_REQUIRE(n > 0 && p > 0, "gram: empty matrix");
    VB_REQUIRE(ws_bytes >= gram_workspace_bytes(n, p), "gram: workspace too small");
    if (gram_tc_eligible(n, p) && kchunk <= 0) return gram_tc_f32(A, n, p, G, ws, ws_bytes, launches, st);
    {
        const int nt0 = ceil_div(n, GT);
        if (kchunk <= 0) kchunk = pick_kchunk(p, nt0 * (nt0 + 1) / 2, 4096);
    }
    kchunk = ceil_div(kchunk, GK) * GK;
    char* w = reinterpret_cast<char*>(ws);
    float* mean = reinterpret_cast<float*>(w);
    w += ((p * sizeof(float) + 255) / 256) * 256;
    double* Gd = reinterpret_cast<double*>(w);
    w += (((size_t)n * n * sizeof(double) + 255) / 256) * 256;
    double* Dm = reinterpret_cast<double*>(w);
    double* mm = Dm + n;
    w += (((size_t)n + 1) * sizeof(double) + 255) / 256 * 256;
    int2* tiles = reinterpret_cast<int2*>(w);

    const int nt = ceil_div(n, GT);
    VB_REQUIRE(nt <= 512, "gram: n=%d too large (max %d)", n, 512 * GT);
    std::vector<int2> htiles;
    for (int i = 0; i < nt; ++i)
        for (int j = i; j < nt; ++j) htiles.push_back(make_int2(i, j));
    const int ntiles = (int)htiles.size();
    // pageable source: the runtime stages the data before returning, so the vector may go out of scope
    VB_CHECK_CUDA(cudaMemcpyAsync(tiles, htiles.data(), ntiles * sizeof(int2), cudaMemcpyHostToDevice, st));
    VB_CHECK_CUDA(cudaMemsetAsync(Gd, 0, (size_t)n * n * sizeof(double), st));
    int nl = 0;
    if (deflate) {
        colmean_kernel<<<(unsigned)ceil_div(p, (size_t)256), 256, 0, st>>>(A, n, p, mean);
        VB_CHECK_LAUNCH();
        deflate_terms_kernel<<<n + 1, 256, 0, st>>>(A, n, p, mean, Dm, mm);
        VB_CHECK_LAUNCH();
        nl += 2;
    }
    const unsigned nchunks = (unsigned)ceil_div(p, (size_t)kchunk);
    VB_REQUIRE(nchunks <= 65535, "gram: too many K chunks");
    if (int rc = launch_gram_tiles(128, A, A, n, n, p, p, deflate ? mean : nullptr, Gd, n, 0, tiles, ntiles, nchunks,
                                   kchunk, st)) return rc;
    gram_assemble_kernel<<<dim3(ceil_div(n, 128), n), 128, 0, st>>>(Gd, n, deflate ? Dm : nullptr, mm, G);
    VB_CHECK_LAUNCH();
    nl += 2;
    if (launches) *launches = nl;
    return 0;
}

// ---- pageable host arrays: multi-threaded staging through a pinned double buffer ------------------------------
// cudaMemcpy2DAsync from PAGEABLE memory is staged by the driver on one thread (~10 GB/s: 50 ms for the 524 MB cube of
// config 2, against 9.5 ms of PCIe time).  A drop-in caller passes a plain numpy array, so the upload path copies each
// pixel slab into pinned memory with a few host threads (memcpy bandwidth adds up across cores) while the previous
// slab is in flight on the DMA engine.
struct Stager {
    void* buf[2] = {nullptr, nullptr};
    size_t cap = 0;
    cudaEvent_t done[2];
    bool ev_ready = false;
};

static bool host_is_pinned(const void* ptr) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, ptr) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

static int stager_threads() {
    static const int v = [] {
        const char* e = getenv("VIP_B200_STAGE_THREADS");
        int t = e ? atoi(e) : 8;
        const int hw = (int)std::thread::hardware_concurrency();
        if (hw > 0 && t > hw) t = hw;
        return t < 1 ? 1 : t;
    }();
    return v;
}

// rows [0, n) x columns [c0, c1) of the row-major host matrix (row pitch p floats) -> contiguous dst (n x (c1-c0))
static void stage_slab(const float* host, int n, size_t p, size_t c0, size_t c1, float* dst) {
    const int nt = stager_threads();
    const size_t w = c1 - c0;
    auto work = [&](int t) {
        const int r0 = (int)((long long)n * t / nt), r1 = (int)((long long)n * (t + 1) / nt);
        for (int r = r0; r < r1; ++r) memcpy(dst + (size_t)r * w, host + (size_t)r * p + c0, w * sizeof(float));
    };
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
}

// Contiguous host -> device copy of `nbytes` on stream `st`, staged through a pinned double buffer by several host
// threads when the source is pageable (plain numpy arrays): ~3-4x the rate of the driver's own single-threaded
// staging.  Returns after the last chunk has been QUEUED (the pinned buffers belong to the library); the caller's
// source may be reused as soon as the call returns.  Pinned sources take one plain cudaMemcpyAsync.
int memcpy_h2d_staged(void* dst, const void* src_host, size_t nbytes, cudaStream_t st) {
    if (nbytes == 0) return 0;
    if (host_is_pinned(src_host) || getenv("VIP_B200_NO_STAGING") != nullptr) {
        VB_CHECK_CUDA(cudaMemcpyAsync(dst, src_host, nbytes, cudaMemcpyHostToDevice, st));
        return 0;
    }
    static Stager stagers[64];
    static std::mutex mu;
    int dev = 0;
    VB_CHECK_CUDA(cudaGetDevice(&dev));
    VB_REQUIRE(dev >= 0 && dev < 64, "memcpy_h2d_staged: device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(mu);              // one staged copy per process at a time (shared buffers)
    Stager& sg = stagers[dev];
    const size_t chunk = (size_t)32 << 20;
    if (!sg.ev_ready) {
        for (int i = 0; i < 2; ++i) VB_CHECK_CUDA(cudaEventCreateWithFlags(&sg.done[i], cudaEventDisableTiming));
        sg.ev_ready = true;
    }
    if (sg.cap < chunk) {
        for (int i = 0; i < 2; ++i) {
            if (sg.buf[i]) cudaFreeHost(sg.buf[i]);
            sg.buf[i] = nullptr;
            if (cudaHostAlloc(&sg.buf[i], chunk, cudaHostAllocDefault) != cudaSuccess) {
                (void)cudaGetLastError();
                sg.cap = 0;
                VB_CHECK_CUDA(cudaMemcpyAsync(dst, src_host, nbytes, cudaMemcpyHostToDevice, st));
                return 0;
            }
        }
        sg.cap = chunk;
    }
    const int nt = stager_threads();
    const char* src = reinterpret_cast<const char*>(src_host);
    char* d = reinterpret_cast<char*>(dst);
    int s = 0;
    for (size_t off = 0; off < nbytes; off += chunk, ++s) {
        const size_t len = (nbytes - off < chunk) ? nbytes - off : chunk;
        const int b = s & 1;
        if (s >= 2) VB_CHECK_CUDA(cudaEventSynchronize(sg.done[b]));
        char* sb = reinterpret_cast<char*>(sg.buf[b]);
        auto work = [&](int t) {
            const size_t a0 = len * t / nt, a1 = len * (t + 1) / nt;
            memcpy(sb + a0, src + off + a0, a1 - a0);
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
        VB_CHECK_CUDA(cudaMemcpyAsync(d + off, sb, len, cudaMemcpyHostToDevice, st));
        VB_CHECK_CUDA(cudaEventRecord(sg.done[b], st));
    }
    // the pinned buffers must not be refilled by the next call before these DMAs are done
    for (int b = 0; b < 2 && b < s; ++b) VB_CHECK_CUDA(cudaEventSynchronize(sg.done[b]));
    return 0;
}

// Strided rows of a host array -> contiguous device buffer (height rows of width bytes, source pitch spitch): the
// same multi-threaded pinned staging as memcpy_h2d_staged, the worker threads gathering the rows.  Used to upload the
// frames [f0, f1) of every spectral channel of a (z, n, H, W) IFS cube chunk by chunk on a copy stream while the
// previous chunk is being processed (psfsub/sdi.py).  Returns after the last DMA has completed.
int memcpy2d_h2d_staged(void* dst, const void* src_host, size_t spitch, size_t width, size_t height, cudaStream_t st) {
    if (width == 0 || height == 0) return 0;
    if (spitch == width) return memcpy_h2d_staged(dst, src_host, width * height, st);
    if (host_is_pinned(src_host) || getenv("VIP_B200_NO_STAGING") != nullptr) {
        VB_CHECK_CUDA(cudaMemcpy2DAsync(dst, width, src_host, spitch, width, height, cudaMemcpyHostToDevice, st));
        VB_CHECK_CUDA(cudaStreamSynchronize(st));
        return 0;
    }
    static Stager stagers[64];
    static std::mutex mu;
    int dev = 0;
    VB_CHECK_CUDA(cudaGetDevice(&dev));
    VB_REQUIRE(dev >= 0 && dev < 64, "memcpy2d_h2d_staged: device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(mu);
    Stager& sg = stagers[dev];
    const size_t chunk = (size_t)32 << 20;
    if (!sg.ev_ready) {
        for (int i = 0; i < 2; ++i) VB_CHECK_CUDA(cudaEventCreateWithFlags(&sg.done[i], cudaEventDisableTiming));
        sg.ev_ready = true;
    }
    if (sg.cap < chunk) {
        for (int i = 0; i < 2; ++i) {
            if (sg.buf[i]) cudaFreeHost(sg.buf[i]);
            sg.buf[i] = nullptr;
            if (cudaHostAlloc(&sg.buf[i], chunk, cudaHostAllocDefault) != cudaSuccess) {
                (void)cudaGetLastError();
                sg.cap = 0;
                VB_CHECK_CUDA(cudaMemcpy2DAsync(dst, width, src_host, spitch, width, height, cudaMemcpyHostToDevice, st));
                VB_CHECK_CUDA(cudaStreamSynchronize(st));
                return 0;
            }
        }
        sg.cap = chunk;
    }
    const int nt = stager_threads();
    const char* src = reinterpret_cast<const char*>(src_host);
    char* d = reinterpret_cast<char*>(dst);
    const size_t nbytes = width * height;
    int s = 0;
    for (size_t off = 0; off < nbytes; off += chunk, ++s) {
        const size_t len = (nbytes - off < chunk) ? nbytes - off : chunk;
        const int b = s & 1;
        if (s >= 2) VB_CHECK_CUDA(cudaEventSynchronize(sg.done[b]));
        char* sb = reinterpret_cast<char*>(sg.buf[b]);
        auto work = [&](int t) {
            size_t a = off + len * t / nt;                      // destination-contiguous byte range [a, a1)
            const size_t a1 = off + len * (t + 1) / nt;
            while (a < a1) {
                const size_t row = a / width, col = a % width;
                size_t m = width - col;
                if (m > a1 - a) m = a1 - a;
                memcpy(sb + (a - off), src + row * spitch + col, m);
                a += m;
            }
        };
        std::vector<std::thread> th;
        for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
        VB_CHECK_CUDA(cudaMemcpyAsync(d + off, sb, len, cudaMemcpyHostToDevice, st));
        VB_CHECK_CUDA(cudaEventRecord(sg.done[b], st));
    }
    for (int b = 0; b < 2 && b < s; ++b) VB_CHECK_CUDA(cudaEventSynchronize(sg.done[b]));
    return 0;
}

// one pixel slab to the device on `copy_stream`: direct strided DMA from pinned memory, or staged (buffer s % 2)
static int upload_slab(const float* host, int n, size_t p, size_t c0, size_t c1, float* M, Stager* staged, int s,
                       cudaStream_t copy_stream) {
    const size_t w = c1 - c0;
    if (staged == nullptr) {
        VB_CHECK_CUDA(cudaMemcpy2DAsync(M + c0, p * sizeof(float), host + c0, p * sizeof(float), w * sizeof(float), n,
                                        cudaMemcpyHostToDevice, copy_stream));
        return 0;
    }
    const int b = s & 1;
    if (s >= 2) VB_CHECK_CUDA(cudaEventSynchronize(staged->done[b]));      // the DMA that last used this buffer
    float* sb = reinterpret_cast<float*>(staged->buf[b]);
    stage_slab(host, n, p, c0, c1, sb);
    VB_CHECK_CUDA(cudaMemcpy2DAsync(M + c0, p * sizeof(float), sb, w * sizeof(float), w * sizeof(float), n,
                                    cudaMemcpyHostToDevice, copy_stream));
    VB_CHECK_CUDA(cudaEventRecord(staged->done[b], copy_stream));
    return 0;
}

// Host cube -> device matrix M (n x p) AND G = M M^T, pipelined: the cube is uploaded in `nslabs` pixel
// slabs (strided 2-D DMAs on a private copy stream) and the Gramian of each slab is accumulated on
// `st` as soon as that slab has landed, so the fp64 SYRK hides behind the PCIe transfer.
// `host` should be pinned (page-locked) memory for the copies to be truly asynchronous.
int upload_gram_f32(const float* host, int n, size_t p, float* M, double* G, void* ws, size_t ws_bytes,
                    int nslabs, int* launches, cudaStream_t st) {
    VB_REQUIRE(n > 0 && p > 0, "upload_gram: empty matrix");
    VB_REQUIRE(ws_bytes >= gram_workspace_bytes(n, p), "upload_gram: workspace too small");
    // one private copy stream + event set PER DEVICE (a process may drive several GPUs), created under a lock
    struct CopyCtx { cudaStream_t stream = nullptr; cudaEvent_t ev[64]; cudaEvent_t start; bool ready = false; };
    static CopyCtx ctxs[64];
    static std::mutex ctx_mutex;
    int dev = 0;
    VB_CHECK_CUDA(cudaGetDevice(&dev));
    VB_REQUIRE(dev >= 0 && dev < 64, "upload_gram: device index %d out of range", dev);
    CopyCtx& cx = ctxs[dev];
    {
        std::lock_guard<std::mutex> lock(ctx_mutex);
        if (!cx.ready) {
            VB_CHECK_CUDA(cudaStreamCreateWithFlags(&cx.stream, cudaStreamNonBlocking));
            for (int i = 0; i < 64; ++i) VB_CHECK_CUDA(cudaEventCreateWithFlags(&cx.ev[i], cudaEventDisableTiming));
            VB_CHECK_CUDA(cudaEventCreateWithFlags(&cx.start, cudaEventDisableTiming));
            cx.ready = true;
        }
    }
    cudaStream_t copy_stream = cx.stream;
    cudaEvent_t* ev = cx.ev;
    cudaEvent_t start_ev = cx.start;
    if (nslabs <= 0) nslabs = 8;
    if (nslabs > 64) nslabs = 64;
    // pageable source: staged through this device's pinned double buffer (see Stager above)
    static Stager stagers[64];
    Stager* staged = nullptr;
    if (!host_is_pinned(host) && stager_threads() > 0 && getenv("VIP_B200_NO_STAGING") == nullptr) {
        staged = &stagers[dev];
        const size_t slab_w = ceil_div(ceil_div(p, (size_t)nslabs), (size_t)64) * 64 + 128;
        const size_t need = (size_t)n * slab_w * sizeof(float);
        std::lock_guard<std::mutex> lock(ctx_mutex);
        if (!staged->ev_ready) {
            for (int i = 0; i < 2; ++i) VB_CHECK_CUDA(cudaEventCreateWithFlags(&staged->done[i], cudaEventDisableTiming));
            staged->ev_ready = true;
        }
        if (staged->cap < need) {
            for (int i = 0; i < 2; ++i) {
                if (staged->buf[i]) cudaFreeHost(staged->buf[i]);
                staged->buf[i] = nullptr;
                if (cudaHostAlloc(&staged->buf[i], need, cudaHostAllocDefault) != cudaSuccess) {
                    (void)cudaGetLastError();
                    staged = nullptr;            // no pinned memory to spare: let the driver stage the copy
                    break;
                }
            }
            if (staged) staged->cap = need;
            else stagers[dev].cap = 0;
        }
    }
    if (gram_tc_eligible(n, p)) {
        // tensor-core SYRK per slab (split pass + tcgen05 kernel), same copy/compute pipeline
        int ntiles = 0, nl = 0;
        if (int rc = gram_tc_begin(n, p, ws, ws_bytes, &ntiles, st)) return rc;
        VB_CHECK_CUDA(cudaEventRecord(start_ev, st));
        VB_CHECK_CUDA(cudaStreamWaitEvent(copy_stream, start_ev, 0));
        const size_t slab = ceil_div(ceil_div(p, (size_t)nslabs), (size_t)64) * 64;
        int s = 0;
        for (size_t c0 = 0; c0 < p; c0 += slab, ++s) {
            const size_t c1 = (c0 + slab < p) ? c0 + slab : p;
            if (int rc = upload_slab(host, n, p, c0, c1, M, staged, s, copy_stream)) return rc;
            VB_CHECK_CUDA(cudaEventRecord(ev[s], copy_stream));
            VB_CHECK_CUDA(cudaStreamWaitEvent(st, ev[s], 0));
            if (int rc = gram_tc_accumulate(M, n, p, p, c0, c1, ws, ntiles, &nl, st)) return rc;
        }
        if (int rc = gram_tc_finish(n, p, ws, G, &nl, st)) return rc;
        if (launches) *launches = nl;
        return 0;
    }
    char* w = reinterpret_cast<char*>(ws);
    w += ((p * sizeof(float) + 255) / 256) * 256;                     // (mean slot, unused)
    double* Gd = reinterpret_cast<double*>(w);
    w += (((size_t)n * n * sizeof(double) + 255) / 256) * 256;
    w += (((size_t)n + 1) * sizeof(double) + 255) / 256 * 256;
    int2* tiles = reinterpret_cast<int2*>(w);
    const int nt = ceil_div(n, GT);
    VB_REQUIRE(nt <= 512, "upload_gram: n=%d too large (max %d)", n, 512 * GT);
    std::vector<int2> htiles;
    for (int i = 0; i < nt; ++i)
        for (int j = i; j < nt; ++j) htiles.push_back(make_int2(i, j));
    const int ntiles = (int)htiles.size();
    VB_CHECK_CUDA(cudaMemcpyAsync(tiles, htiles.data(), ntiles * sizeof(int2), cudaMemcpyHostToDevice, st));
    VB_CHECK_CUDA(cudaMemsetAsync(Gd, 0, (size_t)n * n * sizeof(double), st));
    // the copy stream must not start overwriting M before earlier work on `st` is done with it
    VB_CHECK_CUDA(cudaEventRecord(start_ev, st));
    VB_CHECK_CUDA(cudaStreamWaitEvent(copy_stream, start_ev, 0));
    const size_t slab = ceil_div(ceil_div(p, (size_t)nslabs), (size_t)GK) * GK;
    int nl = 0, s = 0;
    for (size_t c0 = 0; c0 < p; c0 += slab, ++s) {
        const size_t c1 = (c0 + slab < p) ? c0 + slab : p;
        if (int rc = upload_slab(host, n, p, c0, c1, M, staged, s, copy_stream)) return rc;
        VB_CHECK_CUDA(cudaEventRecord(ev[s], copy_stream));
        VB_CHECK_CUDA(cudaStreamWaitEvent(st, ev[s], 0));
        const int kchunk = 1024;
        const unsigned nchunks = (unsigned)ceil_div(c1 - c0, (size_t)kchunk);
        if (int rc = launch_gram_tiles(128, M + c0, M + c0, n, n, c1 - c0, p, nullptr, Gd, n, 0, tiles, ntiles,
                                       nchunks, kchunk, st)) return rc;
        ++nl;
    }
    gram_assemble_kernel<<<dim3(ceil_div(n, 128), n), 128, 0, st>>>(Gd, n, nullptr, nullptr, G);
    VB_CHECK_LAUNCH();
    if (launches) *launches = nl + 1;
    return 0;
}

// C (na x nb fp64, row-major, overwritten) = A B^T.  The operand with the fewer padded rows goes on the A side of
// the tile kernel (64-row tiles when that saves work): sketches (l x n), coefficients (n x k).
int cross_gram_f32(const float* A, int na, const float* B, int nb, size_t p, double* C, void* ws,
                   size_t ws_bytes, int kchunk, cudaStream_t st) {
    VB_REQUIRE(na > 0 && nb > 0 && p > 0, "cross_gram: empty problem");
    auto padded = [](int rows, int t) { return ceil_div(rows, t) * t; };
    // candidate layouts: (A side, tile height); cost = padded rows of A side x padded (128) rows of the B side
    int trans = 0, ta = 128;
    if (!gram_use_v1()) {
        long long best = (long long)padded(na, 128) * padded(nb, 128);
        const long long c64 = (long long)padded(na, 64) * padded(nb, 128);
        const long long t64 = (long long)padded(nb, 64) * padded(na, 128);
        if (c64 < best) { best = c64; ta = 64; trans = 0; }
        if (t64 < best) { best = t64; ta = 64; trans = 1; }
    }
    const float* Pa = trans ? B : A;
    const float* Pb = trans ? A : B;
    const int ra = trans ? nb : na, rb = trans ? na : nb;
    const int nta = ceil_div(ra, ta), ntb = ceil_div(rb, GT);
    VB_REQUIRE((size_t)nta * ntb <= 65535, "cross_gram: too many tiles");
    VB_REQUIRE(ws_bytes >= (size_t)nta * ntb * sizeof(int2), "cross_gram: workspace too small");
    std::vector<int2> htiles;
    for (int i = 0; i < nta; ++i)
        for (int j = 0; j < ntb; ++j) htiles.push_back(make_int2(i, j));
    const int ntiles = (int)htiles.size();
    if (kchunk <= 0) {
        // fill whole waves of the resident CTA slots (2 per SM for 64-row tiles, 1 otherwise) with ~4096-deep chunks
        const int slots = kNumSMs * ((ta == 64) ? 2 : 1);
        const long long want = (long long)ntiles * (long long)ceil_div(p, (size_t)4096);
        const long long waves = (want + slots - 1) / slots;
        long long nch = (waves * slots) / ntiles;
        if (nch < 1) nch = 1;
        if (nch > 65535) nch = 65535;
        kchunk = (int)ceil_div(p, (size_t)nch);
    }
    kchunk = ceil_div(kchunk, GK) * GK;
    int2* tiles = reinterpret_cast<int2*>(ws);
    VB_CHECK_CUDA(cudaMemcpyAsync(tiles, htiles.data(), ntiles * sizeof(int2), cudaMemcpyHostToDevice, st));
    VB_CHECK_CUDA(cudaMemsetAsync(C, 0, (size_t)na * nb * sizeof(double), st));
    const unsigned nchunks = (unsigned)ceil_div(p, (size_t)kchunk);
    VB_REQUIRE(nchunks <= 65535, "cross_gram: too many K chunks");
    return launch_gram_tiles(ta, Pa, Pb, ra, rb, p, p, nullptr, C, nb, trans, tiles, ntiles, nchunks, kchunk, st);
}

}  // namespace vb
