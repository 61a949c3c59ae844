// Gramian on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), fp32-exact products.
//
// Role in the reference: the O(n^2 p) part of the PCA decomposition -- numpy's fp64 SVD of M^T for
// svd_mode='lapack' (src/vip_hci/psfsub/svd.py:466-475) or C = M.M^T for 'eigen' (svd.py:447-450).
//
// Numerics.  numpy decomposes in fp64 and a Gramian squares the condition number, so a TF32/BF16
// single pass is not an option ("fp32 parity through a Gramian").  Instead:
//   1. the temporal mean m[k] of every pixel is removed (M = 1 m^T + D); the rank-2 terms
//      D m 1^T + 1 (D m)^T + (m.m) 1 1^T are accumulated in fp64 on the CUDA cores (cheap, O(n p));
//   2. D (fp32) is split into three bf16 planes  D = D1 + D2 + D3  (8+8+8 mantissa bits: exact);
//   3. D D^T is formed from the six leading cross products D1D1, D1D2, D2D1, D1D3, D3D1, D2D2
//      (each bf16 x bf16 product is exact in the fp32 accumulator; the dropped terms are < 2^-24
//      relative) by tcgen05.mma with the accumulator in TMEM;
//   4. every K-chunk (128 pixels: measured 3e-8 of sqrt(G_ii G_jj) on config 2, vs 1.5e-7 at 512) the TMEM accumulator is drained with tcgen05.ld and added to
//      fp64 registers, so fp32 accumulation error never spans more than one chunk; the two TMEM
//      accumulators are double-buffered so the drain overlaps the next chunk's MMAs.
//
// Kernel shape: one CTA per (128 x 128 output tile, K-split); warp 0 = TMA producer, warp 1 = TMEM
// allocator + single-thread MMA issuer, warps 2..9 = epilogue (two warps per TMEM lane quarter, 64
// columns each -> 64 fp64 accumulators per thread).  Operands are K-major, staged by 3-D TMA
// (k, row, plane) into 128-/64-byte-swizzled shared memory.  Results are reduced over K-splits with
// fp64 atomics into the upper-triangular tile workspace that gram_assemble_kernel mirrors.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include <vector>

namespace vb {

// gram.cu: G = sym(upper 128 x 128 tiles of Gd) [+ Dm[i] + Dm[j] + mm]
int gram_assemble(const double* Gd, int n, const double* Dm, const double* mm, double* G, cudaStream_t st);

namespace tc {

constexpr int TM = 128;                 // UMMA M  (rows of the output tile)
constexpr int TN = 128;                 // UMMA N  (columns of the output tile)
constexpr int UK = 16;                  // UMMA K for 16-bit operands
constexpr int NEPI = 8;                 // epilogue warps
constexpr int NTHREADS = 64 + 32 * NEPI;
constexpr int TMEM_COLS = 2 * TN;       // two fp32 accumulators

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug must abort the kernel (trap) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if (clock64() - t0 > 4000000000LL) __trap();     // ~2 s at 1.9 GHz
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1,
                                            int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"((unsigned long long)tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, both operands K-major bf16, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread = lane)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor of a K-major bf16 tile whose rows are one swizzle span (SW bytes) wide:
// 8-row core groups SW*8 bytes apart (SBO), LBO unused for swizzled K-major, descriptor version 1 (sm_100)
template <int SW>
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    constexpr uint64_t layout = (SW == 128) ? 2 : (SW == 64) ? 4 : 6;
    return (uint64_t)((addr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((SW * 8) >> 4) << 32) | (1ull << 46) |
           (layout << 61);
}
// instruction descriptor: D fp32, A/B bf16, both K-major, M x N
constexpr uint32_t kInstrDesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) |
                                ((uint32_t)(TM >> 4) << 24);

template <int BK, int NPIECE>
struct Cfg {
    static constexpr int SW = BK * 2;                       // swizzle span = row bytes
    static constexpr int PIECE_BYTES = TM * BK * 2;
    static constexpr int STAGE_BYTES = 2 * NPIECE * PIECE_BYTES;
    static constexpr int STAGES = (200 * 1024 / STAGE_BYTES) < 8 ? (200 * 1024 / STAGE_BYTES) : 8;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;
    static constexpr int NPROD = (NPIECE == 3) ? 6 : (NPIECE == 2) ? 3 : 1;
    static_assert(BK == 32 || BK == 64, "BK must be one 64- or 128-byte swizzle span");
    static_assert(STAGES >= 2, "pipeline needs two stages");
};

// C (fp64, atomics) += A-tile . B-tile^T over the K-blocks [kb0, kb1) of this CTA's split.
//   tmA/tmB: 3-D maps (k, row, piece) over the bf16 planes; tiles[]: (row tile, column tile).
template <int BK, int NPIECE>
__global__ void __launch_bounds__(NTHREADS, 1)
gram_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const int2* __restrict__ tiles, int kb_total, int chunk_kb, int same_ab, double* __restrict__ C,
                 int ldc, int na, int nb) {
    using K = Cfg<BK, NPIECE>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[K::STAGES], bar_empty[K::STAGES], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int2 tile = tiles[blockIdx.x];
    const int row0 = tile.x * TM, col0 = tile.y * TN;
    const bool diag = same_ab && tile.x == tile.y;
    // K-blocks of this split
    const int nsplit = gridDim.y, ks = blockIdx.y;
    const int base = kb_total / nsplit, rem = kb_total % nsplit;
    const int kb0 = ks * base + (ks < rem ? ks : rem);
    const int nkb = base + (ks < rem ? 1 : 0);
    const int nchunks = (nkb + chunk_kb - 1) / chunk_kb;

    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < K::STAGES; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 1);
            mbar_init(smem_u32(&bar_empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bar_tfull[b]), 1);
            mbar_init(smem_u32(&bar_tempty[b]), NEPI);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_base_s)), "r"((uint32_t)TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0 && nkb > 0) {
            const uint32_t stage_tx = (diag ? NPIECE : 2 * NPIECE) * K::PIECE_BYTES;
            for (int i = 0; i < nkb; ++i) {
                const int s = i % K::STAGES;
                const uint32_t ph = (i / K::STAGES) & 1;
                mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
                const uint32_t full = smem_u32(&bar_full[s]);
                mbar_expect_tx(full, stage_tx);
                const uint32_t sa = smem0 + s * K::STAGE_BYTES;
                const int kcoord = (kb0 + i) * BK;
#pragma unroll
                for (int pc = 0; pc < NPIECE; ++pc) tma_load_3d(sa + pc * K::PIECE_BYTES, &tmA, full, kcoord, row0, pc);
                if (!diag) {
#pragma unroll
                    for (int pc = 0; pc < NPIECE; ++pc)
                        tma_load_3d(sa + (NPIECE + pc) * K::PIECE_BYTES, &tmB, full, kcoord, col0, pc);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            constexpr int PA[6] = {0, 0, 1, 0, 2, 1};
            constexpr int PB[6] = {0, 1, 0, 2, 0, 1};
            int i = 0;
            for (int c = 0; c < nchunks; ++c) {
                const int buf = c & 1;
                mbar_wait(smem_u32(&bar_tempty[buf]), ((c >> 1) & 1) ^ 1);
                fence_after();
                const uint32_t tacc = tmem_base + buf * TN;
                const int iend = (i + chunk_kb < nkb) ? i + chunk_kb : nkb;
                bool first = true;
                for (; i < iend; ++i) {
                    const int s = i % K::STAGES;
                    const uint32_t ph = (i / K::STAGES) & 1;
                    mbar_wait(smem_u32(&bar_full[s]), ph);
                    fence_after();
                    const uint32_t sa = smem0 + s * K::STAGE_BYTES;
                    const uint32_t sb = diag ? sa : sa + NPIECE * K::PIECE_BYTES;
#pragma unroll
                    for (int pr = 0; pr < K::NPROD; ++pr) {
                        const uint64_t ad = smem_desc<K::SW>(sa + PA[pr] * K::PIECE_BYTES);
                        const uint64_t bd = smem_desc<K::SW>(sb + PB[pr] * K::PIECE_BYTES);
#pragma unroll
                        for (int kk = 0; kk < BK / UK; ++kk) {
                            // advance 32 bytes (16 bf16) inside the swizzle span: +2 in 16-byte units
                            umma_bf16(tacc, ad + 2 * kk, bd + 2 * kk, kInstrDesc, first ? 0u : 1u);
                            first = false;
                        }
                    }
                    umma_commit(smem_u32(&bar_empty[s]));       // frees the smem stage when the MMAs retire
                }
                umma_commit(smem_u32(&bar_tfull[buf]));         // accumulator of this chunk is complete
            }
        }
    } else {
        // ===== epilogue: drain TMEM chunks into fp64 registers =====
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;       // which 64 columns
        double acc[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = 0.0;
        for (int c = 0; c < nchunks; ++c) {
            const int buf = c & 1;
            mbar_wait(smem_u32(&bar_tfull[buf]), (c >> 1) & 1);
            fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * TN + half * 64;
            uint32_t v[32];
            tmem_ld32(taddr, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] += (double)__uint_as_float(v[j]);
            tmem_ld32(taddr + 32, v);
            tmem_ld_wait();
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[buf]));
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[32 + j] += (double)__uint_as_float(v[j]);
        }
        const int r = row0 + q * 32 + lane;
        if (r < na && nchunks > 0) {
            double* dst = C + (size_t)r * ldc + col0 + half * 64;
#pragma unroll
            for (int j = 0; j < 64; ++j)
                if (col0 + half * 64 + j < nb) atomicAdd(dst + j, acc[j]);
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ---- pre-pass: D = A - 1 m^T split into bf16 planes; Dm[i] += sum_k D[i,k] m[k] (fp64) ----------------
// grid (ceil(w / 2048), n), 256 threads, 8 consecutive pixels per thread.
template <int NPIECE>
__global__ void __launch_bounds__(256)
split_planes_kernel(const float* __restrict__ A, int n, size_t w, size_t ld, const float* __restrict__ mean,
                    __nv_bfloat16* __restrict__ planes, size_t ldp, size_t plane_stride,
                    double* __restrict__ Dm) {
    const int row = blockIdx.y;
    const size_t k0 = ((size_t)blockIdx.x * 256 + threadIdx.x) * 8;
    double dm = 0.0;
    if (k0 < w) {
        float d[8];
        const float* src = A + (size_t)row * ld + k0;
        const bool vec = (k0 + 8 <= w) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
        if (vec) {
            const float4 a0 = ld_stream_f4(reinterpret_cast<const float4*>(src));
            const float4 a1 = ld_stream_f4(reinterpret_cast<const float4*>(src) + 1);
            d[0] = a0.x; d[1] = a0.y; d[2] = a0.z; d[3] = a0.w;
            d[4] = a1.x; d[5] = a1.y; d[6] = a1.z; d[7] = a1.w;
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = (k0 + j < w) ? __ldg(src + j) : 0.f;
        }
        if (mean != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float m = (k0 + j < w) ? __ldg(mean + k0 + j) : 0.f;
                d[j] -= m;
                dm += (double)d[j] * (double)m;
            }
        }
        __align__(16) __nv_bfloat16 pc[NPIECE][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float r = d[j];
#pragma unroll
            for (int q = 0; q < NPIECE; ++q) {
                const __nv_bfloat16 h = __float2bfloat16_rn(r);
                pc[q][j] = h;
                r -= __bfloat162float(h);        // exact: the remainder fits the fp32 mantissa
            }
        }
        __nv_bfloat16* dst = planes + (size_t)row * ldp + k0;
#pragma unroll
        for (int q = 0; q < NPIECE; ++q)
            *reinterpret_cast<uint4*>(dst + q * plane_stride) = *reinterpret_cast<const uint4*>(pc[q]);
    }
    if (mean != nullptr) {
        __shared__ double red[8];
        dm = warp_sum(dm);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dm;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int i = 0; i < 8; ++i) t += red[i];
            atomicAdd(Dm + row, t);
        }
    }
}

// temporal mean of every pixel of the slab (fp64 accumulate -> fp32) and mm += sum_k m[k]^2
__global__ void __launch_bounds__(256)
slab_mean_kernel(const float* __restrict__ A, int n, size_t w, size_t ld, float* __restrict__ mean,
                 double* __restrict__ mm) {
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double sq = 0.0;
    if (j < w) {
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += (double)__ldg(A + (size_t)i * ld + j);
        const float m = (float)(s / n);
        mean[j] = m;
        sq = (double)m * (double)m;
    }
    __shared__ double red[8];
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[i];
        atomicAdd(mm, t);
    }
}


// =====================================================================================
// Batched GEMM  C[b] = A[b % amod] . B[b]^T  on the same tcgen05 pipeline (round 2, second half)
// =====================================================================================
// Role in the reference: the two-sided separable resampling  Re(L X L^T)  that replaces the FFT zoom of scale_fft
// (preproc/rescaling.py:1114-1217) for every (ADI frame, channel) of an IFS cube -- BASELINE config 4 spends 83 % of
// its kernel time there on the CUDA-core GEMM of csrc/gemm.cu (profiles/r04c_launches_c4.md).  Both operands are
// K-major bf16x3 planes (three-way error-free split of fp32, six cross products: fp32-grade products, fp32
// accumulation over K <= 896 in TMEM); operators depend on the channel only (row block b % amod of the A planes),
// frames on the batch index (row block b of the B planes).  One CTA per 128 x 128 output tile: warp 0 = TMA
// producer, warp 1 = MMA issuer, warps 2..9 = epilogue.  The epilogue transposes the accumulator through the (by
// then idle) pipeline buffers so that every store instruction writes one contiguous row segment, and writes either
// fp32 C or -- OUT_PLANES -- the bf16x3 planes of C directly, in the layout the NEXT product reads as its K-major B
// operand: rows r < msplit at row r, rows r >= msplit at row r - msplit shifted by N columns (the real and imaginary
// halves of  [Lr; Li] X^T  side by side).
struct GemmTcArgs {
    int M, N, K;                 // per-batch extents (A: M x K, B: N x K)
    int amod, nper;              // operator index = b % amod; nper = batch / amod (batches per operator)
    int ntj;                     // column tiles per batch
    float* C; long long ldc, strideC;                          // fp32 output (OUT_PLANES = false)
    __nv_bfloat16* P; long long ldp, plane_stride; int msplit; // plane output (OUT_PLANES = true)
};

// Accuracy: the tensor core adds every MMA result to the fp32 accumulator with ONE truncating rounding (measured:
// 0.5 ulp of systematic loss per MMA instruction on same-sign data, 1e-6 relative after 36 instructions), so
//   * the leading product D1.D1 and the five small cross products (2^-8 ... 2^-16 of it) go to SEPARATE TMEM
//     accumulators -- the corrections are then rounded relative to their own size, not to the sum's;
//   * every CHUNK_KB K-blocks (128 values of k) both accumulators are drained and added to fp32 registers with
//     round-to-nearest CUDA-core adds; the two accumulator pairs are double-buffered so the drain of chunk c overlaps
//     the MMAs of chunk c + 1 (same protocol as gram_umma_kernel).
constexpr int GEMM_CHUNK_KB = 2;
constexpr int GEMM_TMEM_COLS = 4 * TN;      // (hi, lo) x 2 buffers

template <bool OUT_PLANES>
__global__ void __launch_bounds__(NTHREADS, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmTcArgs g) {
    constexpr int BK = 64, NPIECE = 3;
    using K = Cfg<BK, NPIECE>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[K::STAGES], bar_empty[K::STAGES], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // consecutive CTAs share the operator (L2 reuse of its planes): y -> b = (y % nper) * amod + y / nper
    const int y = blockIdx.y;
    const int b = (g.nper > 0) ? (y % g.nper) * g.amod + y / g.nper : y;
    const int ti = blockIdx.x / g.ntj, tj = blockIdx.x % g.ntj;
    const int rowA0 = (g.amod > 0 ? b % g.amod : b) * g.M + ti * TM;
    const int rowB0 = b * g.N + tj * TN;
    const int nkb = (g.K + BK - 1) / BK;
    const int nchunks = (nkb + GEMM_CHUNK_KB - 1) / GEMM_CHUNK_KB;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;

    if (threadIdx.x == 0) {
        for (int s = 0; s < K::STAGES; ++s) {
            mbar_init(smem_u32(&bar_full[s]), 1);
            mbar_init(smem_u32(&bar_empty[s]), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bar_tfull[i]), 1);
            mbar_init(smem_u32(&bar_tempty[i]), NEPI);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_base_s)), "r"((uint32_t)GEMM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % K::STAGES;
                const uint32_t ph = (i / K::STAGES) & 1;
                mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
                const uint32_t full = smem_u32(&bar_full[s]);
                mbar_expect_tx(full, 2 * NPIECE * K::PIECE_BYTES);
                const uint32_t sa = smem0 + s * K::STAGE_BYTES;
#pragma unroll
                for (int pc = 0; pc < NPIECE; ++pc) tma_load_3d(sa + pc * K::PIECE_BYTES, &tmA, full, i * BK, rowA0, pc);
#pragma unroll
                for (int pc = 0; pc < NPIECE; ++pc)
                    tma_load_3d(sa + (NPIECE + pc) * K::PIECE_BYTES, &tmB, full, i * BK, rowB0, pc);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr int PA[6] = {0, 0, 1, 0, 2, 1};
            constexpr int PB[6] = {0, 1, 0, 2, 0, 1};
            int i = 0;
            for (int c = 0; c < nchunks; ++c) {
                const int buf = c & 1;
                mbar_wait(smem_u32(&bar_tempty[buf]), ((c >> 1) & 1) ^ 1);
                fence_after();
                const uint32_t thi = tmem_base + buf * 2 * TN, tlo = thi + TN;
                const int iend = (i + GEMM_CHUNK_KB < nkb) ? i + GEMM_CHUNK_KB : nkb;
                bool first_hi = true, first_lo = true;
                for (; i < iend; ++i) {
                    const int s = i % K::STAGES;
                    const uint32_t ph = (i / K::STAGES) & 1;
                    mbar_wait(smem_u32(&bar_full[s]), ph);
                    fence_after();
                    const uint32_t sa = smem0 + s * K::STAGE_BYTES;
                    const uint32_t sb = sa + NPIECE * K::PIECE_BYTES;
#pragma unroll
                    for (int pr = 0; pr < 6; ++pr) {
                        const uint64_t ad = smem_desc<K::SW>(sa + PA[pr] * K::PIECE_BYTES);
                        const uint64_t bd = smem_desc<K::SW>(sb + PB[pr] * K::PIECE_BYTES);
#pragma unroll
                        for (int kk = 0; kk < BK / UK; ++kk) {
                            if (pr == 0) { umma_bf16(thi, ad + 2 * kk, bd + 2 * kk, kInstrDesc, first_hi ? 0u : 1u); first_hi = false; }
                            else { umma_bf16(tlo, ad + 2 * kk, bd + 2 * kk, kInstrDesc, first_lo ? 0u : 1u); first_lo = false; }
                        }
                    }
                    umma_commit(smem_u32(&bar_empty[s]));
                }
                umma_commit(smem_u32(&bar_tfull[buf]));
            }
        }
    } else {
        // ===== epilogue: drain the chunks into fp32 registers, then registers -> shared (transpose) -> global =====
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;       // which 64 columns
        constexpr int PITCH = 66;               // floats per staged row (even: float2 reads; 2-way conflicts on the fill)
        float acc[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) acc[j] = 0.f;
        for (int c = 0; c < nchunks; ++c) {
            const int buf = c & 1;
            mbar_wait(smem_u32(&bar_tfull[buf]), (c >> 1) & 1);
            fence_after();
            const uint32_t thi = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 2 * TN + half * 64;
            uint32_t vh[32], vl[32];
            tmem_ld32(thi, vh);
            tmem_ld32(thi + TN, vl);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] += __uint_as_float(vh[j]) + __uint_as_float(vl[j]);
            tmem_ld32(thi + 32, vh);
            tmem_ld32(thi + TN + 32, vl);
            tmem_ld_wait();
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[buf]));
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[32 + j] += __uint_as_float(vh[j]) + __uint_as_float(vl[j]);
        }
        // every MMA has retired (the last tfull) and every TMA load was consumed: the pipeline buffers are free
        float* stage = reinterpret_cast<float*>(smem_raw + (smem0 - smem_u32(smem_raw))) + (size_t)(warp - 2) * 32 * PITCH;
#pragma unroll
        for (int j = 0; j < 64; ++j) stage[lane * PITCH + j] = acc[j];
        __syncwarp();
        const int col = tj * TN + half * 64 + 2 * lane;      // this lane's two columns of every row
        if (col < g.N) {                                     // N is even: col + 1 < N as well
            for (int rr = 0; rr < 32; ++rr) {
                const int r = ti * TM + q * 32 + rr;
                if (r >= g.M) break;
                const float2 val = *reinterpret_cast<const float2*>(stage + rr * PITCH + 2 * lane);
                if (!OUT_PLANES) {
                    *reinterpret_cast<float2*>(g.C + (size_t)b * g.strideC + (size_t)r * g.ldc + col) = val;
                } else {
                    const int r_lo = (r >= g.msplit) ? r - g.msplit : r;
                    const int shift = (r >= g.msplit) ? g.N : 0;
                    __nv_bfloat16* dst = g.P + ((size_t)b * g.msplit + r_lo) * g.ldp + shift + col;
                    float x0 = val.x, x1 = val.y;
#pragma unroll
                    for (int pc = 0; pc < NPIECE; ++pc) {
                        const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                        x0 -= __bfloat162float(h0);
                        x1 -= __bfloat162float(h1);
                        __nv_bfloat162 h;
                        h.x = h0; h.y = h1;
                        *reinterpret_cast<__nv_bfloat162*>(dst + (size_t)pc * g.plane_stride) = h;
                    }
                }
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 1) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)GEMM_TMEM_COLS) : "memory");
    }
}

// fp32 rows (row stride ldx) -> three bf16 planes (row stride ldp, a multiple of 8), 8 consecutive elements per thread
__global__ void __launch_bounds__(256)
split3_rows_kernel(const float* __restrict__ X, long long rows, int Kc, long long ldx,
                   __nv_bfloat16* __restrict__ planes, long long ldp, long long plane_stride, int tpr) {
    const int rpb = 256 / tpr;                              // rows per block
    const long long row = (long long)blockIdx.x * rpb + threadIdx.x / tpr;
    const int k0 = (threadIdx.x % tpr) * 8;
    if (row >= rows || threadIdx.x / tpr >= rpb || k0 >= Kc) return;
    const float* src = X + row * ldx + k0;
    float d[8];
    if (k0 + 8 <= Kc && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(src));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
        d[0] = a0.x; d[1] = a0.y; d[2] = a0.z; d[3] = a0.w;
        d[4] = a1.x; d[5] = a1.y; d[6] = a1.z; d[7] = a1.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = (k0 + j < Kc) ? __ldg(src + j) : 0.f;
    }
    __align__(16) __nv_bfloat16 pc[3][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        float r = d[j];
#pragma unroll
        for (int qq = 0; qq < 3; ++qq) {
            const __nv_bfloat16 h = __float2bfloat16_rn(r);
            pc[qq][j] = h;
            r -= __bfloat162float(h);
        }
    }
    __nv_bfloat16* dst = planes + row * ldp + k0;
#pragma unroll
    for (int qq = 0; qq < 3; ++qq)
        *reinterpret_cast<uint4*>(dst + qq * plane_stride) = *reinterpret_cast<const uint4*>(pc[qq]);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 3-D map (k, row, piece) over bf16 planes; box = BK x 128 x 1, swizzle = BK*2 bytes, OOB reads zero
static int make_map(CUtensorMap* tm, const __nv_bfloat16* planes, size_t w, int rows, int npiece, size_t ldp,
                    size_t plane_stride, int BK) {
    EncodeTiledFn fn = encode_fn();
    VB_REQUIRE(fn != nullptr, "gram_tc: cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)rows, (cuuint64_t)npiece};
    const cuuint64_t strides[2] = {(cuuint64_t)ldp * 2, (cuuint64_t)plane_stride * 2};
    const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)TM, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)planes, dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE,
                          BK == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VB_REQUIRE(r == CUDA_SUCCESS, "gram_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// K-splits so that (#tiles x #splits) fills whole waves of the 148 SMs, each split keeping >= 16 K-blocks
static int pick_splits(int ntiles, int kb_total) {
    const int smax = (kb_total / 16 > 0) ? kb_total / 16 : 1;
    int s = (ntiles <= kNumSMs) ? kNumSMs / ntiles : 1;     // one wave when the tiles alone do not fill it
    if (ntiles > kNumSMs) {
        // several waves of tiles: split K just enough to make the last wave nearly full
        double best_eff = 0.0;
        for (int c = 1; c <= 8; ++c) {
            const long long total = (long long)ntiles * c;
            const long long waves = (total + kNumSMs - 1) / kNumSMs;
            const double eff = (double)total / (double)(waves * kNumSMs);
            if (eff > best_eff + 0.03) { best_eff = eff; s = c; }
        }
    }
    return s < smax ? s : smax;
}

template <int BK, int NPIECE>
static int launch_umma(const CUtensorMap& tmA, const CUtensorMap& tmB, const int2* tiles, int ntiles, size_t w,
                       int same_ab, double* C, int ldc, int na, int nb, cudaStream_t st) {
    using K = Cfg<BK, NPIECE>;
    static bool configured = false;
    if (!configured) {
        VB_CHECK_CUDA(cudaFuncSetAttribute(gram_umma_kernel<BK, NPIECE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           K::SMEM_BYTES));
        configured = true;
    }
    const int kb_total = (int)ceil_div(w, (size_t)BK);
    const int nsplit = pick_splits(ntiles, kb_total);
    const int chunk_kb = env_int("VIP_B200_GRAM_CHUNK", 128) / BK;
    gram_umma_kernel<BK, NPIECE><<<dim3(ntiles, nsplit), NTHREADS, K::SMEM_BYTES, st>>>(
        tmA, tmB, tiles, kb_total, chunk_kb > 0 ? chunk_kb : 1, same_ab, C, ldc, na, nb);
    VB_CHECK_LAUNCH();
    return 0;
}

}  // namespace tc

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------

struct GramTcPlan {
    int n;
    size_t wmax;        // slab width (pixels)
    size_t ldp;         // plane row stride (elements)
    int npiece;
    size_t off_mean, off_Gd, off_Dm, off_tiles, off_planes, total;
};

static size_t align256(size_t b) { return (b + 255) / 256 * 256; }

static GramTcPlan gram_tc_plan(int n, size_t p) {
    GramTcPlan pl;
    pl.n = n;
    pl.npiece = tc::env_int("VIP_B200_GRAM_PIECES", 3) == 2 ? 2 : 3;
    // slab width: bounded plane workspace (default 1 GiB; config 2 fits in one slab); at least 8192
    // pixels, multiple of 64.  (VIP_B200_GRAM_SLAB_MB=96 keeps a slab's planes L2-resident between the
    // split pass and the MMA kernel at the price of more launches.)
    const size_t budget = (size_t)tc::env_int("VIP_B200_GRAM_SLAB_MB", 1024) << 20;
    size_t w = budget / ((size_t)pl.npiece * 2 * (size_t)n);
    if (w < 8192) w = 8192;
    w = w / 64 * 64;
    if (w > p) w = (p + 63) / 64 * 64;
    pl.wmax = w;
    pl.ldp = w;
    const int nt = ceil_div(n, tc::TM);
    size_t o = 0;
    pl.off_mean = o;   o += align256(w * sizeof(float));
    pl.off_Gd = o;     o += align256((size_t)n * n * sizeof(double));
    pl.off_Dm = o;     o += align256(((size_t)n + 1) * sizeof(double));
    pl.off_tiles = o;  o += align256((size_t)nt * nt * sizeof(int2));
    pl.off_planes = (o + 1023) / 1024 * 1024;
    o = pl.off_planes + (size_t)pl.npiece * n * pl.ldp * 2;
    pl.total = o;
    return pl;
}

size_t gram_tc_workspace_bytes(int n, size_t p) { return gram_tc_plan(n, p).total; }

// the tensor-core path pays off (and its tolerances were validated) for large problems only
bool gram_tc_eligible(int n, size_t p) {
    if (tc::env_int("VIP_B200_GRAM_TC", 1) == 0) return false;
    return n >= 32 && n <= 512 * tc::TM && (size_t)n * p >= ((size_t)1 << 22);
}

// zero the accumulators and upload the upper-triangular tile list
int gram_tc_begin(int n, size_t p, void* ws, size_t ws_bytes, int* ntiles_out, cudaStream_t st) {
    const GramTcPlan pl = gram_tc_plan(n, p);
    VB_REQUIRE(ws_bytes >= pl.total, "gram_tc: workspace too small (%zu < %zu)", ws_bytes, pl.total);
    char* w = reinterpret_cast<char*>(ws);
    const int nt = ceil_div(n, tc::TM);
    VB_REQUIRE(nt <= 512, "gram_tc: n=%d too large (max %d)", n, 512 * tc::TM);
    std::vector<int2> htiles;
    for (int i = 0; i < nt; ++i)
        for (int j = i; j < nt; ++j) htiles.push_back(make_int2(i, j));
    const int ntiles = (int)htiles.size();
    // pageable source: the runtime stages the data before returning, so the vector may go out of scope
    VB_CHECK_CUDA(cudaMemcpyAsync(w + pl.off_tiles, htiles.data(), ntiles * sizeof(int2), cudaMemcpyHostToDevice, st));
    VB_CHECK_CUDA(cudaMemsetAsync(w + pl.off_Gd, 0, (size_t)n * n * sizeof(double), st));
    VB_CHECK_CUDA(cudaMemsetAsync(w + pl.off_Dm, 0, ((size_t)n + 1) * sizeof(double), st));
    *ntiles_out = ntiles;
    return 0;
}

// accumulate the columns [c0, c1) of A (row stride ld) into the workspace accumulators
int gram_tc_accumulate(const float* A, int n, size_t p, size_t ld, size_t c0, size_t c1, void* ws, int ntiles,
                       int* launches, cudaStream_t st) {
    const GramTcPlan pl = gram_tc_plan(n, p);
    char* wsb = reinterpret_cast<char*>(ws);
    float* mean = reinterpret_cast<float*>(wsb + pl.off_mean);
    double* Gd = reinterpret_cast<double*>(wsb + pl.off_Gd);
    double* Dm = reinterpret_cast<double*>(wsb + pl.off_Dm);
    const int2* tiles = reinterpret_cast<const int2*>(wsb + pl.off_tiles);
    __nv_bfloat16* planes = reinterpret_cast<__nv_bfloat16*>(wsb + pl.off_planes);
    const size_t plane_stride = (size_t)n * pl.ldp;
    const int BK = tc::env_int("VIP_B200_GRAM_BK", 64) == 32 ? 32 : 64;
    int nl = 0;
    for (size_t s0 = c0; s0 < c1; s0 += pl.wmax) {
        const size_t w = (c1 - s0 < pl.wmax) ? c1 - s0 : pl.wmax;
        const float* As = A + s0;
        tc::slab_mean_kernel<<<(unsigned)ceil_div(w, (size_t)256), 256, 0, st>>>(As, n, w, ld, mean, Dm + n);
        VB_CHECK_LAUNCH();
        const dim3 sg((unsigned)ceil_div(w, (size_t)2048), n);
        if (pl.npiece == 3)
            tc::split_planes_kernel<3><<<sg, 256, 0, st>>>(As, n, w, ld, mean, planes, pl.ldp, plane_stride, Dm);
        else
            tc::split_planes_kernel<2><<<sg, 256, 0, st>>>(As, n, w, ld, mean, planes, pl.ldp, plane_stride, Dm);
        VB_CHECK_LAUNCH();
        CUtensorMap tm;
        if (int rc = tc::make_map(&tm, planes, w, n, pl.npiece, pl.ldp, plane_stride, BK)) return rc;
        int rc;
        if (BK == 64 && pl.npiece == 3) rc = tc::launch_umma<64, 3>(tm, tm, tiles, ntiles, w, 1, Gd, n, n, n, st);
        else if (BK == 64) rc = tc::launch_umma<64, 2>(tm, tm, tiles, ntiles, w, 1, Gd, n, n, n, st);
        else if (pl.npiece == 3) rc = tc::launch_umma<32, 3>(tm, tm, tiles, ntiles, w, 1, Gd, n, n, n, st);
        else rc = tc::launch_umma<32, 2>(tm, tm, tiles, ntiles, w, 1, Gd, n, n, n, st);
        if (rc) return rc;
        nl += 3;
    }
    if (launches) *launches += nl;
    return 0;
}

// G = sym(upper tiles) + deflation terms
int gram_tc_finish(int n, size_t p, void* ws, double* G, int* launches, cudaStream_t st) {
    const GramTcPlan pl = gram_tc_plan(n, p);
    char* wsb = reinterpret_cast<char*>(ws);
    const double* Gd = reinterpret_cast<const double*>(wsb + pl.off_Gd);
    const double* Dm = reinterpret_cast<const double*>(wsb + pl.off_Dm);
    if (int rc = gram_assemble(Gd, n, Dm, Dm + n, G, st)) return rc;
    if (launches) *launches += 1;
    return 0;
}

int gram_tc_f32(const float* A, int n, size_t p, double* G, void* ws, size_t ws_bytes, int* launches,
                cudaStream_t st) {
    VB_REQUIRE(n > 0 && p > 0, "gram_tc: empty matrix");
    int ntiles = 0, nl = 0;
    if (int rc = gram_tc_begin(n, p, ws, ws_bytes, &ntiles, st)) return rc;
    if (int rc = gram_tc_accumulate(A, n, p, p, 0, p, ws, ntiles, &nl, st)) return rc;
    if (int rc = gram_tc_finish(n, p, ws, G, &nl, st)) return rc;
    if (launches) *launches = nl;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// batched bf16x3 GEMM on tcgen05 (SDI rescaling operators)
// ------------------------------------------------------------------------------------------------
int split3_bf16(const float* X, long long rows, int K, long long ldx, void* planes, long long ldp,
                long long plane_stride, cudaStream_t st) {
    VB_REQUIRE(rows > 0 && K > 0 && K <= 2048, "split3: bad extents (rows=%lld, K=%d; K <= 2048)", rows, K);
    VB_REQUIRE(ldp % 8 == 0 && ldp >= (K + 7) / 8 * 8 && plane_stride % 8 == 0, "split3: plane pitch must be a multiple of 8 >= K");
    VB_REQUIRE((reinterpret_cast<uintptr_t>(planes) & 15) == 0, "split3: planes must be 16-byte aligned");
    int tpr = ((K + 7) / 8 + 31) / 32 * 32;
    if (tpr > 256) tpr = 256;
    const int rpb = 256 / tpr;
    const long long blocks = (rows + rpb - 1) / rpb;
    VB_REQUIRE(blocks < (1ll << 31), "split3: too many rows");
    tc::split3_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(X, rows, K, ldx, reinterpret_cast<__nv_bfloat16*>(planes),
                                                             ldp, plane_stride, tpr);
    VB_CHECK_LAUNCH();
    return 0;
}

// C[b] (M x N) = A[b % amod] (M x K) . B[b] (N x K)^T for b < batch.  A planes: (amod * M) rows, B planes: (batch * N)
// rows, both three bf16 planes of pitch ldpA / ldpB.  Output: fp32 C (ldc, strideC) when planesC == nullptr, else the
// bf16x3 planes of C with rows r >= msplit folded to row r - msplit at column offset N (pitch ldpC >= 2 N).
int gemm_bf16x3_tc(const void* planesA, long long ldpA, long long strideA, int amod, const void* planesB,
                   long long ldpB, long long strideB, int M, int N, int K, int batch, float* C, long long ldc,
                   long long strideC, void* planesC, long long ldpC, long long strideCp, int msplit,
                   cudaStream_t st) {
    VB_REQUIRE(M > 0 && N > 0 && K > 0 && batch > 0 && amod > 0, "gemm_tc: empty problem");
    VB_REQUIRE(batch <= 65535, "gemm_tc: at most 65535 batches per call (got %d)", batch);
    VB_REQUIRE(N % 2 == 0, "gemm_tc: N must be even (got %d)", N);
    VB_REQUIRE(ldpA % 8 == 0 && ldpB % 8 == 0 && strideA % 8 == 0 && strideB % 8 == 0, "gemm_tc: plane pitches must be multiples of 8");
    VB_REQUIRE((long long)amod * M < (1ll << 31) && (long long)batch * N < (1ll << 31), "gemm_tc: too many rows");
    const bool out_planes = planesC != nullptr;
    if (out_planes) {
        VB_REQUIRE(msplit > 0 && msplit <= M && ldpC % 2 == 0 && strideCp % 2 == 0 && ldpC >= (M > msplit ? 2 : 1) * N,
                   "gemm_tc: bad plane output layout");
        VB_REQUIRE((reinterpret_cast<uintptr_t>(planesC) & 3) == 0, "gemm_tc: output planes must be 4-byte aligned");
    } else {
        VB_REQUIRE(C != nullptr && ldc % 2 == 0 && strideC % 2 == 0 && (reinterpret_cast<uintptr_t>(C) & 7) == 0,
                   "gemm_tc: fp32 output needs even ldc / strideC and an 8-byte aligned base");
    }
    CUtensorMap tmA, tmB;
    if (int rc = tc::make_map(&tmA, reinterpret_cast<const __nv_bfloat16*>(planesA), (size_t)K, amod * M, 3,
                              (size_t)ldpA, (size_t)strideA, 64)) return rc;
    if (int rc = tc::make_map(&tmB, reinterpret_cast<const __nv_bfloat16*>(planesB), (size_t)K, batch * N, 3,
                              (size_t)ldpB, (size_t)strideB, 64)) return rc;
    using Kc = tc::Cfg<64, 3>;
    static bool configured = false;
    if (!configured) {
        VB_CHECK_CUDA(cudaFuncSetAttribute(tc::gemm_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Kc::SMEM_BYTES));
        VB_CHECK_CUDA(cudaFuncSetAttribute(tc::gemm_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           Kc::SMEM_BYTES));
        configured = true;
    }
    tc::GemmTcArgs g;
    g.M = M; g.N = N; g.K = K;
    g.amod = amod;
    g.nper = (batch % amod == 0) ? batch / amod : 0;
    g.ntj = ceil_div(N, tc::TN);
    g.C = C; g.ldc = ldc; g.strideC = strideC;
    g.P = reinterpret_cast<__nv_bfloat16*>(planesC); g.ldp = ldpC; g.plane_stride = strideCp; g.msplit = msplit;
    const dim3 grid(ceil_div(M, tc::TM) * g.ntj, batch);
    if (out_planes) tc::gemm_umma_kernel<true><<<grid, tc::NTHREADS, Kc::SMEM_BYTES, st>>>(tmA, tmB, g);
    else tc::gemm_umma_kernel<false><<<grid, tc::NTHREADS, Kc::SMEM_BYTES, st>>>(tmA, tmB, g);
    VB_CHECK_LAUNCH();
    return 0;
}

}  // namespace vb
