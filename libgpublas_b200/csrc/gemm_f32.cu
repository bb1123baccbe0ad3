// gemm_f32.cu -- SGEMM for sm_100a on the 5th-generation tensor cores (tcgen05.mma kind::tf32) with
// the 3xTF32 split that keeps FP32 accuracy, replacing the reference's forward to cublasSgemm
// (blas_level3/gemm.cc:46-83, :143-160).  north_star (1): "SGEMM held to FP32 accuracy via FFMA or
// 3xTF32 split emulation".
//
//  1. split pass (HBM-bound, ~1 ms at n=16384):  a = hi + lo with hi = RN_tf32(a), lo = RN_tf32(a - hi)
//     (a - hi is exact in fp32).  Both operands are written k-contiguous ("K-major"), so the
//     transpose/no-transpose cases collapse into one GEMM kernel: Asplit[2][m][kpad], Bsplit[2][n][kpad].
//  2. GEMM: CTA tile 128 x BN (BN = 256 with BK = 16 and four 48 KiB stages when the grid still fills the SMs, else
//     BN = 128 with BK = 32 and three 64 KiB stages; one swizzle row of 64 or 128 bytes per operand row).  Warp 0 = TMA producer
//     (one 3-D box per operand per stage brings hi and lo planes together), warp 1 = MMA issuer
//     (one thread; per 8-wide k step three tcgen05.mma into the same TMEM accumulator:
//     lo*hi, hi*lo, hi*hi -- small terms first), warps 2..9 = epilogue (tcgen05.ld 32x32b.x64: lane
//     quarter x column half; each thread owns one row of the tile, so a warp stores 32 consecutive
//     rows of a C column = 128 B).
//     smem ring: mbarrier full[] armed by TMA bytes, empty[] released by tcgen05.commit.
//  The dropped lo*lo term is O(2^-22) relative; products are exact in the tensor core, accumulation
//  is fp32 in TMEM.
#include "common.cuh"
#include "kernels.h"
#include "gemm_generic.cuh"
#include "runtime.h"
#include <cstdio>
#include <cstdlib>

namespace b200 {

constexpr int SG_BM = 128;
constexpr int SG_THREADS = 64 + 256;      // TMA warp, MMA warp, 8 epilogue warps (lane quarter x column half)
constexpr int SG_CHUNK_K = 128;
constexpr int SG_DEFAULT_WIDE_CFG = 2;     // large problems: 128x256, BK=16, 4 stages (232 TFLOP/s at n=16384 vs 203 for 128x128)           // k per TMEM accumulation chunk (see the epilogue comment)
// Tile configurations <BN, BK, STAGES>: BK = 32 floats is one 128-byte swizzle row, BK = 16 one 64-byte swizzle row.
//   <128, 32, 3>  64 KiB stages; the 128-wide N keeps the MMA shared-memory read rate at ~120 B/clk (the limit is 128)
//   <256, 16, 4>  48 KiB stages; N = 256 halves the A re-reads per flop (88 B/clk) and raises flop/byte of L2 traffic
//   <256, 32, 2>  96 KiB stages (two only)

struct SgemmParams {
    int m, n, k;
    float alpha, beta;
    float* C; int64_t ldc;
    int mask;
    int tiles_m, tiles_n;
    const int* nonfinite;     // set by the split pass when an operand holds Inf/NaN: this kernel stands down, the FFMA tile kernel runs
    // CGEMM runs on the same kernel as a real product of doubled size (see cgemm_tf32x3): C is then the interleaved (re,im)
    // matrix viewed as 2m x n floats, rows 2i / 2i+1 = real / imaginary part of complex row i, and beta is complex.
    int cplx; float beta_im;
};

// ------------------------------------------------------------------ split pass
__device__ __forceinline__ float tf32_rn(float a) {
    // round-to-nearest (ties away) onto the 10-bit tf32 mantissa; inf/nan pass through; a finite value whose rounding
    // would carry into the exponent 0xFF (|a| >= 0x7F7FF000, e.g. FLT_MAX) is truncated instead, so hi stays finite
    uint32_t b = __float_as_uint(a);
    if ((b & 0x7f800000u) != 0x7f800000u) {
        const uint32_t r = b + 0x1000u;
        if ((r & 0x7f800000u) != 0x7f800000u) b = r;
    }
    return __uint_as_float(b & 0xffffe000u);
}
// Inf/NaN operands cannot go through the split: hi = Inf meets the other operand's lo = 0 in the hi*lo term and Inf*0 = NaN
// where sgemm returns Inf.  The split pass raises a flag instead; the tensor-core kernel then returns at once and the FFMA
// tile kernel queued behind it (which otherwise returns at once) computes the product with IEEE semantics.
__device__ __forceinline__ void split_store(float a, float* hi, float* lo, int* nonfinite) {
    if ((__float_as_uint(a) & 0x7f800000u) == 0x7f800000u) *nonfinite = 1;
    const float h = tf32_rn(a);
    *hi = h;
    *lo = tf32_rn(a - h);
}
// source already k-contiguous: element (kk, r) at src[kk + r*ld]
__global__ void __launch_bounds__(256) split_kmajor_kernel(int rows, int k, const float* __restrict__ src, int64_t ld,
                                                           float* __restrict__ hi, float* __restrict__ lo, int64_t kpad, int* nonfinite) {
    const int kk = blockIdx.x * 256 + threadIdx.x;
    if (kk >= k) return;
    for (int r = blockIdx.y; r < rows; r += gridDim.y)
        split_store(__ldg(src + kk + (int64_t)r * ld), hi + (int64_t)r * kpad + kk, lo + (int64_t)r * kpad + kk, nonfinite);
}
// source row-contiguous: element (r, kk) at src[r + kk*ld]  ->  dst[r*kpad + kk]   (32x32 smem transpose)
__global__ void __launch_bounds__(256) split_transpose_kernel(int rows, int k, const float* __restrict__ src, int64_t ld,
                                                              float* __restrict__ hi, float* __restrict__ lo, int64_t kpad, int* nonfinite) {
    __shared__ float t[32][33];
    const int r0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int kk = k0 + ty + 8 * j, r = r0 + tx;
        t[ty + 8 * j][tx] = (r < rows && kk < k) ? __ldg(src + r + (int64_t)kk * ld) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int r = r0 + ty + 8 * j, kk = k0 + tx;
        if (r < rows && kk < k) split_store(t[tx][ty + 8 * j], hi + (int64_t)r * kpad + kk, lo + (int64_t)r * kpad + kk, nonfinite);
    }
}

// ------------------------------------------------------------------ tcgen05 helpers
template <int ROW_BYTES> __device__ __forceinline__ uint64_t umma_desc_kmajor(uint32_t saddr) {
    // K-major operand tile whose rows are ROW_BYTES long (= the TMA swizzle span): 8-row groups ROW_BYTES*8 apart (SBO),
    // LBO unused (=1), descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B / 4 = SWIZZLE_64B
    static_assert(ROW_BYTES == 128 || ROW_BYTES == 64, "swizzle span");
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)((ROW_BYTES * 8) >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)(ROW_BYTES == 128 ? 2 : 4) << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
        "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr));
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}

// Bounded mbarrier wait: a protocol error (wrong expect_tx byte count, lost commit) traps after ~seconds
// instead of hanging the GPU until the watchdog; costs one add per failed poll.
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 28)) { printf("b200blas: sgemm mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
    }
}
#define mbar_wait mbar_wait_bounded

// instruction descriptor: D=F32, A=B=TF32, both K-major, M=128, N=BN
template <int BN> __host__ __device__ constexpr uint32_t sg_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(SG_BM >> 4) << 24);
}

template <int BN, int SG_BK, int SG_STAGES>
__global__ void __launch_bounds__(SG_THREADS, 1)
sgemm_tf32x3_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const SgemmParams p) {
    constexpr int A_PLANE = SG_BM * SG_BK * 4, B_PLANE = BN * SG_BK * 4;     // one (hi or lo) tile
    constexpr int STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;
    constexpr uint32_t TMEM_COLS = 2 * BN;                                   // two ping-pong accumulators
    constexpr int SG_CHUNK_STAGES = SG_CHUNK_K / SG_BK;
    constexpr int ROW_BYTES = SG_BK * 4;
    constexpr int HALF = BN / 2;                                             // columns per epilogue warp

    int tile_m, tile_n;
    {
        constexpr int BAND = 8;
        int t = blockIdx.x;
        int band = t / (BAND * p.tiles_m);
        int r = t - band * (BAND * p.tiles_m);
        int bw = min(BAND, p.tiles_n - band * BAND);
        tile_m = r / bw;
        tile_n = band * BAND + (r - tile_m * bw);
    }
    const int m0 = tile_m * SG_BM, n0 = tile_n * BN;
    if (*p.nonfinite) return;                      // Inf/NaN in an operand: the FFMA tile kernel behind this launch does the work
    {
        const int rlo = p.cplx ? (m0 >> 1) : m0, rhi = p.cplx ? ((m0 + SG_BM - 1) >> 1) : (m0 + SG_BM - 1);    // rows in C's own (complex) index space
        if (p.mask == MASK_LOWER && rhi < n0) return;
        if (p.mask == MASK_UPPER && n0 + BN - 1 < rlo) return;
    }

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = (uint64_t*)(smem + SG_STAGES * STAGE_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + SG_STAGES);
    const uint32_t accfull0 = smem_u32(bars + 2 * SG_STAGES), accempty0 = smem_u32(bars + 2 * SG_STAGES + 2);
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * SG_STAGES + 4);
    const uint32_t smem_base = smem_u32(smem);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ktiles = (p.k + SG_BK - 1) / SG_BK;
    const int nchunks = (ktiles + SG_CHUNK_STAGES - 1) / SG_CHUNK_STAGES;

    if (tid == 0) {
        for (int s = 0; s < SG_STAGES; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int b = 0; b < 2; b++) { mbar_init(accfull0 + 8 * b, 1); mbar_init(accempty0 + 8 * b, 256); }
        mbar_fence_init();
        tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapB);
    }
    if (warp == 2) {   // TMEM allocation is warp-collective; the same warp frees it at the end
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_acc = *tmem_slot;

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kt = 0; kt < ktiles; kt++) {
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                const uint32_t fb = full0 + 8 * stage;
                const uint32_t sA = smem_base + stage * STAGE_BYTES, sB = sA + 2 * A_PLANE;
                mbar_expect_tx(fb, STAGE_BYTES);
                tma_load_3d(sA, &mapA, kt * SG_BK, m0, 0, fb);     // box {32 k, 128 rows, 2 planes}: hi then lo
                tma_load_3d(sB, &mapB, kt * SG_BK, n0, 0, fb);
                if (++stage == SG_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // =========================== MMA issuer (one thread) ===========================
        if (lane == 0) {
            constexpr uint32_t idesc = sg_idesc<BN>();
            int stage = 0; uint32_t phase = 0;
            for (int c = 0; c < nchunks; c++) {
                const int buf = c & 1;
                mbar_wait(accempty0 + 8 * buf, ((c >> 1) & 1) ^ 1);     // epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tacc = tmem_acc + (uint32_t)(buf * BN);
                const int kt0 = c * SG_CHUNK_STAGES, kt1 = min(ktiles, kt0 + SG_CHUNK_STAGES);
                for (int kt = kt0; kt < kt1; kt++) {
                    mbar_wait(full0 + 8 * stage, phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t sA = smem_base + stage * STAGE_BYTES, sB = sA + 2 * A_PLANE;
                    const uint64_t ahi = umma_desc_kmajor<ROW_BYTES>(sA), alo = umma_desc_kmajor<ROW_BYTES>(sA + A_PLANE);
                    const uint64_t bhi = umma_desc_kmajor<ROW_BYTES>(sB), blo = umma_desc_kmajor<ROW_BYTES>(sB + B_PLANE);
#pragma unroll
                    for (int ks = 0; ks < SG_BK / 8; ks++) {
                        const uint64_t adv = (uint64_t)((ks * 32) >> 4);   // 8 tf32 = 32 bytes along k inside the swizzle row
                        umma_tf32(tacc, alo + adv, bhi + adv, idesc, (kt > kt0 || ks > 0) ? 1u : 0u);   // first MMA of a chunk overwrites
                        umma_tf32(tacc, ahi + adv, blo + adv, idesc, 1);
                        umma_tf32(tacc, ahi + adv, bhi + adv, idesc, 1);
                    }
                    umma_commit(empty0 + 8 * stage);      // frees the smem slot when these MMAs have read it
                    if (++stage == SG_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(accfull0 + 8 * buf);          // this chunk's partial product is complete
            }
        }
        __syncwarp();
    } else {
        // =========================== epilogue warps 2..9 ===========================
        // The tensor core accumulates in fp32 with truncation, a bias that grows linearly with the number of
        // accumulation steps.  So the k loop is cut into chunks of SG_CHUNK_STAGES*32; each chunk is summed in
        // TMEM from zero and folded into per-thread register accumulators here with round-to-nearest FADDs,
        // while the MMA warp is already filling the other TMEM accumulator.
        const int q = warp & 3;                       // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;             // which HALF columns of the tile
        float acc[HALF];
#pragma unroll
        for (int j = 0; j < HALF; j++) acc[j] = 0.f;
        for (int c = 0; c < nchunks; c++) {
            const int buf = c & 1;
            mbar_wait(accfull0 + 8 * buf, (c >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tsrc = tmem_acc + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * BN + half * HALF);
            if constexpr (HALF == 64) {
                uint32_t v[64];
                tmem_ld64(tsrc, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(accempty0 + 8 * buf);     // values are in registers: the MMA warp may overwrite the buffer
#pragma unroll
                for (int j = 0; j < 64; j++) acc[j] += __uint_as_float(v[j]);
            } else {                                  // 128 columns per warp: 32 at a time keeps the staging registers low
#pragma unroll
                for (int c0 = 0; c0 < HALF; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tsrc + (uint32_t)c0, v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; j++) acc[c0 + j] += __uint_as_float(v[j]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(accempty0 + 8 * buf);
            }
        }
        const int64_t row = (int64_t)m0 + 32 * q + lane;
        const bool beta0 = p.beta == 0.f && p.beta_im == 0.f;
        const int64_t crow = p.cplx ? (row >> 1) : row;          // row in C's own index space (triangle masks are defined there)
        const bool imag_row = p.cplx && (row & 1);
#pragma unroll
        for (int j = 0; j < HALF; j++) {
            const int64_t col = (int64_t)n0 + half * HALF + j;
            const bool ok = row < p.m && col < p.n && tri_keep(p.mask, crow, col);
            float* cp = p.C + row + col * p.ldc;
            float r = p.alpha * acc[j];
            if (!beta0) {
                const float old = ok ? *cp : 0.f;
                if (p.cplx) {
                    // complex beta: (re, im) of one element sit in adjacent lanes (rows 2i, 2i+1 of the float view)
                    const float other = __shfl_xor_sync(0xffffffffu, old, 1);
                    r += imag_row ? fmaf(p.beta, old, p.beta_im * other) : fmaf(p.beta, old, -p.beta_im * other);
                } else r = fmaf(p.beta, old, r);
            }
            if (ok) *cp = r;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
static bool make_map_split(CUtensorMap* map, const float* base, int64_t rows, int64_t k, int64_t kpad, int box_rows, int bk) {
    cuuint64_t gdim[3] = {(cuuint64_t)k, (cuuint64_t)rows, 2};
    cuuint64_t gstride[2] = {(cuuint64_t)kpad * 4, (cuuint64_t)rows * (cuuint64_t)kpad * 4};
    cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)box_rows, 2}, estr[3] = {1, 1, 1};
    return encode_tensor_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, gdim, gstride, box, estr,
                             bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

template <int BN, int BK, int STAGES>
static bool launch_sg(cudaStream_t s, const float* as, const float* bs, int64_t kpad, SgemmParams p) {
    constexpr int SMEM = STAGES * (2 * SG_BM + 2 * BN) * BK * 4 + (2 * STAGES + 6) * 8 + 1024;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    CUtensorMap ma, mb;
    memset(&ma, 0, sizeof ma); memset(&mb, 0, sizeof mb);
    if (!make_map_split(&ma, as, p.m, p.k, kpad, SG_BM, BK) || !make_map_split(&mb, bs, p.n, p.k, kpad, BN, BK)) return false;
    set_max_dynamic_smem((const void*)sgemm_tf32x3_kernel<BN, BK, STAGES>, SMEM);
    p.tiles_m = (p.m + SG_BM - 1) / SG_BM;
    p.tiles_n = (p.n + BN - 1) / BN;
    sgemm_tf32x3_kernel<BN, BK, STAGES><<<p.tiles_m * p.tiles_n, SG_THREADS, SMEM, s>>>(ma, mb, p);
    return true;
}

static bool sgemm_tf32x3(cudaStream_t s, int oa, int ob, int m, int n, int k, float alpha, const float* A, int64_t lda,
                         const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int mask) {
    if (!tma_available()) return false;
    const int64_t kpad = ((int64_t)k + 3) / 4 * 4;          // 16-byte row pitch for TMA
    float* as = (float*)ws_alloc((size_t)2 * m * kpad * 4);
    float* bs = (float*)ws_alloc((size_t)2 * n * kpad * 4);
    int* nonfinite = (int*)ws_alloc(256);
    B200_CUDA(cudaMemsetAsync(nonfinite, 0, 4, s));
    // op(A) is m x k: 'N' stores it m-contiguous (transpose needed), 'T'/'C' store it k-contiguous
    if (oa == 0) split_transpose_kernel<<<dim3((m + 31) / 32, (k + 31) / 32), 256, 0, s>>>(m, k, A, lda, as, as + (int64_t)m * kpad, kpad, nonfinite);
    else split_kmajor_kernel<<<dim3((k + 255) / 256, m < 65535 ? m : 65535), 256, 0, s>>>(m, k, A, lda, as, as + (int64_t)m * kpad, kpad, nonfinite);
    // op(B) is k x n: 'N' stores it k-contiguous per column, 'T'/'C' n-contiguous
    if (ob == 0) split_kmajor_kernel<<<dim3((k + 255) / 256, n < 65535 ? n : 65535), 256, 0, s>>>(n, k, B, ldb, bs, bs + (int64_t)n * kpad, kpad, nonfinite);
    else split_transpose_kernel<<<dim3((n + 31) / 32, (k + 31) / 32), 256, 0, s>>>(n, k, B, ldb, bs, bs + (int64_t)n * kpad, kpad, nonfinite);
    SgemmParams p;
    p.m = m; p.n = n; p.k = k; p.alpha = alpha; p.beta = beta; p.C = C; p.ldc = ldc; p.mask = mask; p.tiles_m = p.tiles_n = 0; p.nonfinite = nonfinite; p.cplx = 0; p.beta_im = 0.f;
    // tile configuration: option sgemm_cfg=<n> / B200BLAS_SGEMM_CFG override (0: 128x128 BK32 x3, 1: 128x256 BK32 x2, 2: 128x256 BK16 x4)
    static const int cfg_env = getenv("B200BLAS_SGEMM_CFG") ? atoi(getenv("B200BLAS_SGEMM_CFG")) : -1;
    const int64_t sms = sm_count() > 0 ? sm_count() : 148;
    int cfg = g_opts.sgemm_cfg >= 0 ? g_opts.sgemm_cfg : cfg_env;
    if (cfg < 0) cfg = ((int64_t)((m + 127) / 128) * ((n + 255) / 256) >= sms) ? SG_DEFAULT_WIDE_CFG : 0;
    bool ok;
    if (cfg == 1) ok = launch_sg<256, 32, 2>(s, as, bs, kpad, p);
    else if (cfg == 2) ok = launch_sg<256, 16, 4>(s, as, bs, kpad, p);
    else ok = launch_sg<128, 32, 3>(s, as, bs, kpad, p);
    if (!ok) return false;
    gemm_generic_launch<float>(s, "NTC"[oa], "NTC"[ob], m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask, nonfinite);   // runs only if the flag is set
    last_variant = VAR_TF32X3_TCGEN05;
    return true;
}

void sgemm_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, float alpha, const float* A, int64_t lda,
               const float* B, int64_t ldb, float beta, float* C, int64_t ldc, int mask) {
    if (m <= 0 || n <= 0) return;
    if (alpha == 0.f || k <= 0) { scale_matrix<float>(s, m, n, beta, C, ldc, mask); last_variant = VAR_SCALE_ONLY; return; }
    // size-based variant selector: the tensor-core path pays a split pass over A and B, so it is used
    // when the product is large enough to amortise it
    int variant = force_variant;
    if (variant == VAR_NONE) variant = ((double)m * n * k >= 256.0 * 256.0 * 256.0 && m >= 64 && n >= 64) ? VAR_TF32X3_TCGEN05 : VAR_GENERIC_TILE;
    if (variant == VAR_TF32X3_TCGEN05 && sgemm_tf32x3(s, op_code(ta), op_code(tb), m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask)) return;
    gemm_generic_launch<float>(s, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask);
}

// ------------------------------------------------------------------ CGEMM on the same tensor-core kernel
// (reference blas_level3/gemm.cc:181-198 forwards to cublasCgemm; round 1 ran CGEMM on the FFMA tile kernel.)
// A complex product is a real product of doubled size on the INTERLEAVED result: with C viewed as a 2m x n float matrix
// (rows 2i / 2i+1 = re / im of complex row i, leading dimension 2*ldc)
//     C'' = A'' * B'',   A'' (2m x 2k): row 2i = [ Re a(i,:) | -Im a(i,:) ],  row 2i+1 = [ Im a(i,:) | Re a(i,:) ],
//                        B'' (2k x n) :  column j = [ Re b(:,j) ; Im b(:,j) ],      a = alpha * op(A), b = op(B)
// -- 8mnk real flops, the same count as the complex product, all of them on tcgen05 (3xTF32 split per real operand).
// The split pass builds A'' and B'' k-contiguous straight from the interleaved operands (transposes, conjugations and the
// complex alpha folded in); the GEMM kernel is sgemm_tf32x3_kernel with a complex-beta epilogue (SgemmParams::cplx).
// src element (r, kk) of op(X): KMAJOR -> src[kk + r*ld] (k contiguous), else src[r + kk*ld] (r contiguous, 32x32 smem transpose)
template <bool IS_A>
__device__ __forceinline__ void csplit_emit(cuFloatComplex v, bool conj, bool scale, cuFloatComplex alpha, int64_t r, int kk, int k, float* hi, float* lo,
                                            int64_t kpad, int* nonfinite) {
    if (conj) v.y = -v.y;
    if (scale) v = make_cuFloatComplex(alpha.x * v.x - alpha.y * v.y, alpha.x * v.y + alpha.y * v.x);
    if (IS_A) {
        const int64_t b0 = (2 * r) * kpad, b1 = (2 * r + 1) * kpad;
        split_store(v.x, hi + b0 + kk, lo + b0 + kk, nonfinite);
        split_store(-v.y, hi + b0 + k + kk, lo + b0 + k + kk, nonfinite);
        split_store(v.y, hi + b1 + kk, lo + b1 + kk, nonfinite);
        split_store(v.x, hi + b1 + k + kk, lo + b1 + k + kk, nonfinite);
    } else {
        const int64_t b0 = r * kpad;
        split_store(v.x, hi + b0 + kk, lo + b0 + kk, nonfinite);
        split_store(v.y, hi + b0 + k + kk, lo + b0 + k + kk, nonfinite);
    }
}
template <bool IS_A>
__global__ void __launch_bounds__(256) csplit_kmajor_kernel(int rows, int k, const cuFloatComplex* __restrict__ src, int64_t ld, bool conj, bool scale,
                                                            cuFloatComplex alpha, float* __restrict__ hi, float* __restrict__ lo, int64_t kpad, int* nonfinite) {
    const int kk = blockIdx.x * 256 + threadIdx.x;
    if (kk >= k) return;
    for (int r = blockIdx.y; r < rows; r += gridDim.y)
        csplit_emit<IS_A>(__ldg(src + kk + (int64_t)r * ld), conj, scale, alpha, r, kk, k, hi, lo, kpad, nonfinite);
}
template <bool IS_A>
__global__ void __launch_bounds__(256) csplit_transpose_kernel(int rows, int k, const cuFloatComplex* __restrict__ src, int64_t ld, bool conj, bool scale,
                                                               cuFloatComplex alpha, float* __restrict__ hi, float* __restrict__ lo, int64_t kpad, int* nonfinite) {
    __shared__ cuFloatComplex t[32][33];
    const int r0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int kk = k0 + ty + 8 * j, r = r0 + tx;
        t[ty + 8 * j][tx] = (r < rows && kk < k) ? __ldg(src + r + (int64_t)kk * ld) : make_cuFloatComplex(0.f, 0.f);
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int r = r0 + ty + 8 * j, kk = k0 + tx;
        if (r < rows && kk < k) csplit_emit<IS_A>(t[tx][ty + 8 * j], conj, scale, alpha, r, kk, k, hi, lo, kpad, nonfinite);
    }
}

static bool cgemm_tf32x3(cudaStream_t s, int oa, int ob, int m, int n, int k, cuFloatComplex alpha, const cuFloatComplex* A, int64_t lda,
                         const cuFloatComplex* B, int64_t ldb, cuFloatComplex beta, cuFloatComplex* C, int64_t ldc, int mask) {
    if (!tma_available() || (int64_t)2 * m > 0x7fffffff || (int64_t)2 * k > 0x7fffffff) return false;
    const int M2 = 2 * m, K2 = 2 * k;
    const int64_t kpad = ((int64_t)K2 + 3) / 4 * 4;
    float* as = (float*)ws_alloc((size_t)2 * M2 * kpad * 4);
    float* bs = (float*)ws_alloc((size_t)2 * n * kpad * 4);
    int* nonfinite = (int*)ws_alloc(256);
    B200_CUDA(cudaMemsetAsync(nonfinite, 0, 4, s));
    const bool scale = !(alpha.x == 1.f && alpha.y == 0.f);
    float* alo = as + (int64_t)M2 * kpad; float* blo = bs + (int64_t)n * kpad;
    // op(A) is m x k: 'N' stores it row(m)-contiguous (transpose), 'T'/'C' k-contiguous
    if (oa == 0) csplit_transpose_kernel<true><<<dim3((m + 31) / 32, (k + 31) / 32), 256, 0, s>>>(m, k, A, lda, false, scale, alpha, as, alo, kpad, nonfinite);
    else csplit_kmajor_kernel<true><<<dim3((k + 255) / 256, m < 65535 ? m : 65535), 256, 0, s>>>(m, k, A, lda, oa == 2, scale, alpha, as, alo, kpad, nonfinite);
    // op(B) is k x n: 'N' is k-contiguous per column, 'T'/'C' n-contiguous
    if (ob == 0) csplit_kmajor_kernel<false><<<dim3((k + 255) / 256, n < 65535 ? n : 65535), 256, 0, s>>>(n, k, B, ldb, false, false, alpha, bs, blo, kpad, nonfinite);
    else csplit_transpose_kernel<false><<<dim3((n + 31) / 32, (k + 31) / 32), 256, 0, s>>>(n, k, B, ldb, ob == 2, false, alpha, bs, blo, kpad, nonfinite);
    SgemmParams p;
    p.m = M2; p.n = n; p.k = K2; p.alpha = 1.f; p.beta = beta.x; p.beta_im = beta.y; p.C = (float*)C; p.ldc = 2 * ldc; p.mask = mask;
    p.tiles_m = p.tiles_n = 0; p.nonfinite = nonfinite; p.cplx = 1;
    static const int cfg_env = getenv("B200BLAS_SGEMM_CFG") ? atoi(getenv("B200BLAS_SGEMM_CFG")) : -1;
    const int64_t sms = sm_count() > 0 ? sm_count() : 148;
    int cfg = g_opts.sgemm_cfg >= 0 ? g_opts.sgemm_cfg : cfg_env;
    if (cfg < 0) cfg = ((int64_t)((M2 + 127) / 128) * ((n + 255) / 256) >= sms) ? SG_DEFAULT_WIDE_CFG : 0;
    bool ok;
    if (cfg == 1) ok = launch_sg<256, 32, 2>(s, as, bs, kpad, p);
    else if (cfg == 2) ok = launch_sg<256, 16, 4>(s, as, bs, kpad, p);
    else ok = launch_sg<128, 32, 3>(s, as, bs, kpad, p);
    if (!ok) return false;
    gemm_generic_launch<cuFloatComplex>(s, "NTC"[oa], "NTC"[ob], m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask, nonfinite);   // runs only if the flag is set
    last_variant = VAR_TF32X3_TCGEN05;
    return true;
}

void cgemm_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, cuFloatComplex alpha, const cuFloatComplex* A,
               int64_t lda, const cuFloatComplex* B, int64_t ldb, cuFloatComplex beta, cuFloatComplex* C, int64_t ldc,
               int mask) {
    if (m <= 0 || n <= 0) return;
    if (num<cuFloatComplex>::is_zero(alpha) || k <= 0) { scale_matrix<cuFloatComplex>(s, m, n, beta, C, ldc, mask); last_variant = VAR_SCALE_ONLY; return; }
    // same size rule as SGEMM (in real flops a complex product is 4x a real one of the same shape)
    int variant = force_variant;
    if (variant == VAR_NONE) variant = ((double)m * n * k >= 160.0 * 160.0 * 160.0 && m >= 32 && n >= 64) ? VAR_TF32X3_TCGEN05 : VAR_GENERIC_TILE;
    if (variant == VAR_TF32X3_TCGEN05 && cgemm_tf32x3(s, op_code(ta), op_code(tb), m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask)) return;
    gemm_generic_launch<cuFloatComplex>(s, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask);
}

#undef mbar_wait
}  // namespace b200
