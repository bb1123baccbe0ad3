// gemm_f64.cu -- DGEMM for sm_100a on the FP64 tensor pipe (DMMA), replacing the reference's
// forward to cublasDgemm (blas_level3/gemm.cc:46-83, :162-179).
//
// Design (DESIGN.md section "DGEMM"):
//  * CTA tile 128x128, BK=16 doubles per stage, 6-stage shared-memory ring (6 x 32 KiB).
//  * Warp-specialised: warp-group 0 is the producer (one elected thread issues TMA
//    cp.async.bulk.tensor loads completing on mbarriers; or, for operands TMA cannot address --
//    odd lda / 8-byte-aligned base -- all 128 producer threads stage tiles with LDG+STS into the
//    SAME swizzled layout), warp-groups 1-2 are 8 consumer warps (2 x 4), each owning a 64x32
//    register tile = 32 DMMA.8x8x4 accumulators (128 registers), fed by conflict-free LDS.64.
//  * setmaxnreg moves registers from the producer group (40) to the consumers (232).
//  * Shared tiles use the TMA 128-byte swizzle; the k index each lane feeds to DMMA is a
//    permutation (k = {0,3,12,15}[lane&3] ^ {0,1,4,5}[step]) chosen so that fragment loads are
//    bank-conflict-free for both the m/n-contiguous ("COL") and k-contiguous ("KC") tile layouts,
//    i.e. for all four transpose combinations, without ever transposing data.
//  * Epilogue applies alpha/beta straight from registers (8 consecutive rows per column segment),
//    optionally restricted to a triangle (SYRK) -- C traffic is <1% of the k loop at the sizes
//    that matter.
#include "common.cuh"
#include "kernels.h"
#include "gemm_generic.cuh"
#include "runtime.h"
#include <cstdlib>
#include <cstdio>

namespace b200 {

constexpr int DG_BK = 16;
constexpr int DG_STAGES = 6;
enum { LAY_COL = 0, LAY_KC = 1 };

// byte offset of element (r, kk) inside one operand tile; r = row of op(A) (or column of op(B))
// within the tile, kk = k index within the stage.  Must match CU_TENSOR_MAP_SWIZZLE_128B.
template <int LAY> __host__ __device__ __forceinline__ uint32_t tile_off(int r, int kk) {
    if (LAY == LAY_COL) {   // 16-row slabs of [k][16 contiguous r], 2 KiB each
        int in = r & 15;
        return (uint32_t)((r >> 4) * 2048 + kk * 128 + ((((in >> 1) ^ (kk & 7))) << 4) + (in & 1) * 8);
    } else {                // [r][16 contiguous k], 128 B per row
        return (uint32_t)(r * 128 + ((((kk >> 1) ^ (r & 7))) << 4) + (kk & 1) * 8);
    }
}

struct DgemmParams {
    int m, n, k;
    double alpha, beta;
    const double* A; int64_t lda;
    const double* B; int64_t ldb;
    double* C; int64_t ldc;      // C-in (read when beta != 0)
    double* D; int64_t ldd;      // output; == C for plain BLAS, a peer-mapped tile for partitioned calls
    int mask;
    int tiles_m, tiles_n;
    // Optional panel-readiness flags (partitioned multi-GPU GEMM): rows [g*a_group, (g+1)*a_group) of op(A) may be read
    // only once aflags[g] >= flag_epoch, columns [h*b_group, (h+1)*b_group) of op(B) once bflags[h] >= flag_epoch.
    // The flags are written by the home GPU's copy engine over NVLink after the corresponding panel piece has landed,
    // in the order the tile schedule consumes them, so ONE launch overlaps transfer and compute.
    const uint32_t* aflags; const uint32_t* bflags; int a_group, b_group; uint32_t flag_epoch;
};

__device__ __forceinline__ void wait_flag(const uint32_t* flag, uint32_t epoch) {
    uint32_t v;
    for (uint32_t spins = 0;; spins++) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v >= epoch) break;
        __nanosleep(200);
        // bounded (~10 s): a panel piece that never arrives is a protocol error -- trap instead of hanging the GPU
        if (spins > (1u << 25)) { printf("b200blas: dgemm panel flag timeout (block %d, flag %p = %u, epoch %u)\n", blockIdx.x, flag, v, epoch); __trap(); }
    }
    asm volatile("fence.proxy.async;" ::: "memory");   // order the async-proxy (TMA) reads after the acquire
}

// MB*NB >= 32 (the 128x128 tile): one CTA per SM, registers moved from the producer group to the consumers.
// Smaller tiles (64x64, 64x32; used when a 128x128 tiling would leave most SMs without a tile): the consumers
// need few registers, so two CTAs share an SM and setmaxnreg is skipped.
template <int MB, int NB, int LAYA, int LAYB, bool USE_TMA>
__global__ void __launch_bounds__(384, (MB * NB >= 32) ? 1 : 2)
dgemm_dmma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const DgemmParams p) {
    constexpr int BM = 16 * MB, BN = 32 * NB;          // 2 consumer warps along m, 4 along n
    constexpr int A_BYTES = BM * DG_BK * 8, B_BYTES = BN * DG_BK * 8;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;

    // ---- tile coordinates: bands of 16 tile-columns, walked down the rows, so the ~148 CTAs in
    // flight share A row-panels and B column-panels in L2 ----
    int tile_m, tile_n;
    {
        constexpr int BAND = 16;
        int t = blockIdx.x;
        int band = t / (BAND * p.tiles_m);
        int r = t - band * (BAND * p.tiles_m);
        int bw = min(BAND, p.tiles_n - band * BAND);
        tile_m = r / bw;
        tile_n = band * BAND + (r - tile_m * bw);
    }
    const int m0 = tile_m * BM, n0 = tile_n * BN;
    if (p.mask == MASK_LOWER && m0 + BM - 1 < n0) return;
    if (p.mask == MASK_UPPER && n0 + BN - 1 < m0) return;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // dynamic smem base is only guaranteed 16-B aligned: round up to the 1 KiB swizzle atom
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space
    uint64_t* bars = (uint64_t*)(smem + DG_STAGES * STAGE_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + DG_STAGES);
    const uint32_t smem_base = smem_u32(smem);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ktiles = (p.k + DG_BK - 1) / DG_BK;

    if (tid == 0) {
        for (int s = 0; s < DG_STAGES; s++) {
            mbar_init(full0 + 8 * s, USE_TMA ? 1 : 128);
            mbar_init(empty0 + 8 * s, 256);   // one arrival per consumer thread
        }
        mbar_fence_init();
        if (USE_TMA) { tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapB); }
    }
    __syncthreads();

    if (warp < 4) {
        // =========================== producer warp-group ===========================
        if (MB * NB >= 32) setmaxnreg_dec<40>();
        if (USE_TMA) {
            if (tid == 0) {
                int stage = 0; uint32_t phase = 0;
                if (p.aflags) wait_flag(p.aflags + m0 / p.a_group, p.flag_epoch);   // group sizes are multiples of the tile
                if (p.bflags) wait_flag(p.bflags + n0 / p.b_group, p.flag_epoch);
                for (int kt = 0; kt < ktiles; kt++) {
                    mbar_wait(empty0 + 8 * stage, phase ^ 1);
                    const uint32_t fb = full0 + 8 * stage;
                    const uint32_t sA = smem_base + stage * STAGE_BYTES, sB = sA + A_BYTES;
                    mbar_expect_tx(fb, STAGE_BYTES);
                    const int k0 = kt * DG_BK;
                    if (LAYA == LAY_COL) {
#pragma unroll
                        for (int sl = 0; sl < BM / 16; sl++) tma_load_2d(sA + sl * 2048, &mapA, m0 + sl * 16, k0, fb);
                    } else {
                        tma_load_2d(sA, &mapA, k0, m0, fb);
                    }
                    if (LAYB == LAY_COL) {
#pragma unroll
                        for (int sl = 0; sl < BN / 16; sl++) tma_load_2d(sB + sl * 2048, &mapB, n0 + sl * 16, k0, fb);
                    } else {
                        tma_load_2d(sB, &mapB, k0, n0, fb);
                    }
                    if (++stage == DG_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        } else {
            // LDG staging for operands TMA cannot describe (lda odd / base not 16-B aligned)
            int stage = 0; uint32_t phase = 0;
            if (p.aflags) wait_flag(p.aflags + m0 / p.a_group, p.flag_epoch);
            if (p.bflags) wait_flag(p.bflags + n0 / p.b_group, p.flag_epoch);
            for (int kt = 0; kt < ktiles; kt++) {
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                uint8_t* sA = smem + stage * STAGE_BYTES;
                uint8_t* sB = sA + A_BYTES;
                const int k0 = kt * DG_BK;
#pragma unroll 4
                for (int it = 0; it < BM * DG_BK / 128; it++) {
                    int idx = tid + it * 128, r, kk;
                    if (LAYA == LAY_COL) { r = idx % BM; kk = idx / BM; } else { kk = idx % DG_BK; r = idx / DG_BK; }
                    double v = 0.0;
                    if (m0 + r < p.m && k0 + kk < p.k)
                        v = (LAYA == LAY_COL) ? __ldg(p.A + (int64_t)(m0 + r) + (int64_t)(k0 + kk) * p.lda)
                                              : __ldg(p.A + (int64_t)(k0 + kk) + (int64_t)(m0 + r) * p.lda);
                    *(double*)(sA + tile_off<LAYA>(r, kk)) = v;
                }
#pragma unroll 4
                for (int it = 0; it < BN * DG_BK / 128; it++) {
                    int idx = tid + it * 128, r, kk;
                    if (LAYB == LAY_COL) { r = idx % BN; kk = idx / BN; } else { kk = idx % DG_BK; r = idx / DG_BK; }
                    double v = 0.0;
                    if (n0 + r < p.n && k0 + kk < p.k)
                        v = (LAYB == LAY_COL) ? __ldg(p.B + (int64_t)(n0 + r) + (int64_t)(k0 + kk) * p.ldb)
                                              : __ldg(p.B + (int64_t)(k0 + kk) + (int64_t)(n0 + r) * p.ldb);
                    *(double*)(sB + tile_off<LAYB>(r, kk)) = v;
                }
                mbar_arrive(full0 + 8 * stage);   // release: this thread's stores are visible to waiters
                if (++stage == DG_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // =============================== consumer warps ===============================
    if (MB * NB >= 32) setmaxnreg_inc<232>();
    const int cw = warp - 4;
    const int wm0 = (cw & 1) * (8 * MB), wn0 = (cw >> 1) * (8 * NB);
    const int g = lane >> 2, tig = lane & 3;
    const int kb = (tig & 1) * 3 + (tig >> 1) * 12;   // {0,3,12,15}

    // per-lane fragment offsets for the 4 k-steps of a stage; row-block parity only matters for COL
    uint32_t offA[2][4], offB[2][4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int kk = kb ^ ((s & 1) | ((s >> 1) << 2));   // ^ {0,1,4,5}
        offA[0][s] = tile_off<LAYA>(wm0 + g, kk);
        offA[1][s] = tile_off<LAYA>(wm0 + 8 + g, kk);
        offB[0][s] = A_BYTES + tile_off<LAYB>(wn0 + g, kk);
        offB[1][s] = A_BYTES + tile_off<LAYB>(wn0 + 8 + g, kk);
    }
    constexpr int PAIR_STRIDE = 2048;   // two 8-row blocks: one COL slab, or 16 KC rows x 128 B

    double acc[MB][NB][2];
#pragma unroll
    for (int i = 0; i < MB; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    // beta != 0: pull this tile of C into L2 now (no registers, no waiting), so that the epilogue's reads -- which
    // cannot be overlapped with the k loop because the accumulators fill the register file -- hit L2 instead of
    // HBM.  Matters for short k (panel updates with k = 1-2 K: the epilogue is ~10% of the tile time).
    if (p.beta != 0.0) {
        const int ct = tid - 128;                              // 0..255
        for (int idx = ct; idx < BN * (BM / 16); idx += 256) {
            const int col = idx / (BM / 16), seg = idx - col * (BM / 16);
            const int64_t r = (int64_t)m0 + seg * 16, c = (int64_t)n0 + col;
            if (r < p.m && c < p.n) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.C + r + c * p.ldc));
        }
    }

    {
        int stage = 0; uint32_t phase = 0;
        for (int kt = 0; kt < ktiles; kt++) {
            mbar_wait(full0 + 8 * stage, phase);
            const uint8_t* sS = smem + stage * STAGE_BYTES;
#pragma unroll
            for (int s = 0; s < 4; s++) {
                double a[MB], b[NB];
#pragma unroll
                for (int i = 0; i < MB; i++) a[i] = *(const double*)(sS + offA[i & 1][s] + (i >> 1) * PAIR_STRIDE);
#pragma unroll
                for (int j = 0; j < NB; j++) b[j] = *(const double*)(sS + offB[j & 1][s] + (j >> 1) * PAIR_STRIDE);
#pragma unroll
                for (int i = 0; i < MB; i++)
#pragma unroll
                    for (int j = 0; j < NB; j++) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
            // Releasing a stage to the TMA producer: the consumer's shared-memory reads are generic-proxy accesses, the
            // refill is an async-proxy write, and the PTX memory model orders the two only through a proxy fence --
            // an mbarrier arrive alone is NOT enough.  Without this fence two CTAs sharing an SM (the small tiles)
            // produced sporadic stale 8x8 blocks in 12-60% of launches (profiles/r01b_small_tile_race.txt); with one
            // CTA per SM it never showed, but the hazard is the same.  Every lane fences and arrives for its own reads.
            if (USE_TMA) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(empty0 + 8 * stage);
            if (++stage == DG_STAGES) { stage = 0; phase ^= 1; }
        }
    }

    // ---- epilogue: C = alpha*acc + beta*C on the kept region ----
    // Loads of the old C are issued 2*MB at a time BEFORE any dependent store: C and D may alias, so the compiler must
    // keep program order between a store and the next load, and a load->fma->store chain per element serialises the
    // memory latency (measured: +22 us per 128x128 tile, 2.5 ms on a 16384^2 C).
    const bool beta0 = (p.beta == 0.0);
#pragma unroll
    for (int j = 0; j < NB; j++) {
        double old[2][MB];
        bool ok[2][MB];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int64_t col = n0 + wn0 + 8 * j + 2 * tig + c;
            const double* cp = p.C + col * p.ldc;
#pragma unroll
            for (int i = 0; i < MB; i++) {
                const int64_t row = m0 + wm0 + 8 * i + g;
                ok[c][i] = col < p.n && row < p.m && tri_keep(p.mask, row, col);
                old[c][i] = (!beta0 && ok[c][i]) ? cp[row] : 0.0;
            }
        }
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int64_t col = n0 + wn0 + 8 * j + 2 * tig + c;
            double* dp = p.D + col * p.ldd;
#pragma unroll
            for (int i = 0; i < MB; i++) {
                const int64_t row = m0 + wm0 + 8 * i + g;
                if (!ok[c][i]) continue;
                double v = p.alpha * acc[i][j][c];
                if (!beta0) v = fma(p.beta, old[c][i], v);
                dp[row] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
static bool make_map_f64(CUtensorMap* map, const double* base, int layout, int64_t rows /*m or n extent*/,
                         int64_t kext, int64_t ld, int tile_rows) {
    // COL: dim0 = rows (contiguous), dim1 = k ; KC: dim0 = k (contiguous), dim1 = rows
    cuuint64_t gdim[2], gstride[1];
    cuuint32_t box[2], estr[2] = {1, 1};
    if (layout == LAY_COL) { gdim[0] = rows; gdim[1] = kext; box[0] = 16; box[1] = DG_BK; }
    else                   { gdim[0] = kext; gdim[1] = rows; box[0] = DG_BK; box[1] = tile_rows; }
    gstride[0] = (cuuint64_t)ld * 8;
    return encode_tensor_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, gdim, gstride, box, estr,
                             CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int MB, int NB, int LAYA, int LAYB, bool USE_TMA>
static void launch_dmma(cudaStream_t s, const CUtensorMap& ma, const CUtensorMap& mb, const DgemmParams& p) {
    constexpr int BM = 16 * MB, BN = 32 * NB;
    constexpr int SMEM0 = DG_STAGES * (BM + BN) * DG_BK * 8 + 2 * DG_STAGES * 8 + 1024;
    auto kern = dgemm_dmma_kernel<MB, NB, LAYA, LAYB, USE_TMA>;
    static const int extra = getenv("B200BLAS_DBG_EXTRA_SMEM") ? atoi(getenv("B200BLAS_DBG_EXTRA_SMEM")) : 0;   // experiment: force 1 CTA/SM
    const int SMEM = SMEM0 + (MB * NB < 32 ? extra : 0);
    set_max_dynamic_smem((const void*)kern, SMEM);
    kern<<<p.tiles_m * p.tiles_n, 384, SMEM, s>>>(ma, mb, p);
}

template <int MB, int NB>
static void dgemm_dmma_dispatch(cudaStream_t s, bool nota, bool notb, bool tma, DgemmParams p) {
    constexpr int BM = 16 * MB, BN = 32 * NB;
    p.tiles_m = (p.m + BM - 1) / BM;
    p.tiles_n = (p.n + BN - 1) / BN;
    CUtensorMap ma, mb;
    memset(&ma, 0, sizeof ma); memset(&mb, 0, sizeof mb);
    const int la = nota ? LAY_COL : LAY_KC, lb = notb ? LAY_KC : LAY_COL;
    if (tma) {
        tma = make_map_f64(&ma, p.A, la, p.m, p.k, p.lda, BM) && make_map_f64(&mb, p.B, lb, p.n, p.k, p.ldb, BN);
    }
#define B200_DL(LA, LB)                                                      \
    do {                                                                     \
        if (tma) launch_dmma<MB, NB, LA, LB, true>(s, ma, mb, p);            \
        else     launch_dmma<MB, NB, LA, LB, false>(s, ma, mb, p);           \
    } while (0)
    if (la == LAY_COL && lb == LAY_KC) B200_DL(LAY_COL, LAY_KC);
    else if (la == LAY_KC && lb == LAY_KC) B200_DL(LAY_KC, LAY_KC);
    else if (la == LAY_COL && lb == LAY_COL) B200_DL(LAY_COL, LAY_COL);
    else B200_DL(LAY_KC, LAY_COL);
#undef B200_DL
    last_variant = tma ? VAR_DMMA_TMA : VAR_DMMA_LDG;
}

void dgemm_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, double alpha, const double* A, int64_t lda,
               const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int mask) {
    dgemm_out_dev(s, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, C, ldc, mask);
}

static thread_local const uint32_t* t_aflags = nullptr;
static thread_local const uint32_t* t_bflags = nullptr;
static thread_local int t_a_group = 0, t_b_group = 0;
static thread_local uint32_t t_flag_epoch = 0;
// The next dgemm_out_dev call of this thread polls the panel flags (see DgemmParams::aflags); group sizes must be
// multiples of the 128-wide CTA tile.
void dgemm_set_panel_flags(const uint32_t* aflags, int a_group, const uint32_t* bflags, int b_group, uint32_t epoch) {
    t_aflags = aflags; t_bflags = bflags; t_a_group = a_group; t_b_group = b_group; t_flag_epoch = epoch;
}

// D := alpha*op(A)*op(B) + beta*C with D possibly distinct from C (D may be a peer-mapped pointer: the
// epilogue then stores the tile over NVLink -- the fused compute + C-return of the partitioned GEMM).
void dgemm_out_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, double alpha, const double* A, int64_t lda,
                   const double* B, int64_t ldb, double beta, const double* Cin, int64_t ldc, double* D, int64_t ldd, int mask) {
    double* C = const_cast<double*>(Cin);
    if (m <= 0 || n <= 0) return;
    const bool separate_out = (D != C);
    if ((alpha == 0.0 || k <= 0) && !separate_out) {
        scale_matrix<double>(s, m, n, beta, C, ldc, mask);
        last_variant = VAR_SCALE_ONLY;
        return;
    }
    const bool nota = op_code(ta) == 0, notb = op_code(tb) == 0;
    // ---- size-based variant selector (north_star (3); replaces the reference's CPU cutoff,
    // gemm.cc:129-141): tiny products stay on the generic tile kernel, everything else runs on
    // the DMMA pipeline; TMA staging when both operands are TMA-addressable. ----
    int variant = force_variant;
    if (variant == VAR_NONE) {
        const double work = (double)m * n * k;
        variant = (work < 32.0 * 32.0 * 32.0) ? VAR_GENERIC_TILE : VAR_DMMA_TMA;
    }
    if (separate_out || t_aflags || t_bflags) variant = (variant == VAR_DMMA_LDG) ? VAR_DMMA_LDG : VAR_DMMA_TMA;   // only the DMMA kernel has a D operand / flags
    if (variant == VAR_GENERIC_TILE) {
        gemm_generic_launch<double>(s, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask);
        return;
    }
    const bool tma_ok = variant != VAR_DMMA_LDG && tma_available() &&
                        ((uintptr_t)A % 16 == 0) && ((uintptr_t)B % 16 == 0) && (lda % 2 == 0) && (ldb % 2 == 0) &&
                        lda * 8 < ((int64_t)1 << 40) && ldb * 8 < ((int64_t)1 << 40);
    DgemmParams p;
    p.m = m; p.n = n; p.k = k; p.alpha = alpha; p.beta = beta;
    p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc; p.D = D; p.ldd = ldd; p.mask = mask;
    p.tiles_m = p.tiles_n = 0;
    p.aflags = t_aflags; p.bflags = t_bflags; p.a_group = t_a_group; p.b_group = t_b_group; p.flag_epoch = t_flag_epoch;
    const bool flagged = t_aflags || t_bflags;
    t_aflags = t_bflags = nullptr;
    // Tile shape by available parallelism: a CTA streams its whole k range at one SM's DMMA rate (2.1 us per
    // 16-deep step of a 128x128 tile), so when a 128x128 tiling yields fewer tiles than SMs -- panel updates,
    // triangular-solve leaves, Cholesky diagonal blocks -- smaller tiles finish sooner although each is less efficient.
    const int64_t sms = sm_count() > 0 ? sm_count() : 148;
    auto ntiles = [&](int bm, int bn) { return (int64_t)((m + bm - 1) / bm) * ((n + bn - 1) / bn); };
    static const int dbg_tiles = getenv("B200BLAS_DBG_TILES") ? atoi(getenv("B200BLAS_DBG_TILES")) : 0;   // 1: 128x128 only, 2: 64x64 only, 3: 64x32 only
    const bool small_tma = tma_ok;
    if (dbg_tiles == 2) { dgemm_dmma_dispatch<4, 2>(s, nota, notb, small_tma, p); return; }
    if (dbg_tiles == 3) { dgemm_dmma_dispatch<4, 1>(s, nota, notb, small_tma, p); return; }
    // variant=dmma_tma (forced) means the TMA kernel proper, i.e. the 128x128 tile
    // a dimension of at most 64 (the k = 64 updates of the triangular recursions) would leave half of every 128-wide tile empty
    // ... and a short k loop (<= smallk) is dominated by per-CTA fill and epilogue, which two 64x64 CTAs per SM overlap
    // (measured on the k <= 256 updates of DTRSM 30720 x 2048: 5.47 -> 5.22 ms); only while the 128x128 tiling has few waves
    static const int smallk = getenv("B200BLAS_DGEMM_SMALLK") ? atoi(getenv("B200BLAS_DGEMM_SMALLK")) : 256;
    const bool narrow = ((m <= 64 || n <= 64) || (k <= smallk && ntiles(128, 128) < 4 * sms)) && ntiles(64, 64) >= sms;
    if (dbg_tiles == 1 || force_variant == VAR_DMMA_TMA || flagged || (ntiles(128, 128) >= sms && !narrow)) dgemm_dmma_dispatch<8, 4>(s, nota, notb, tma_ok, p);
    else if (ntiles(64, 64) >= sms) dgemm_dmma_dispatch<4, 2>(s, nota, notb, small_tma, p);
    else dgemm_dmma_dispatch<4, 1>(s, nota, notb, small_tma, p);
}

}  // namespace b200
