// gemm_z.cu -- ZGEMM for sm_100a on the FP64 tensor pipe (DMMA), replacing the reference's forward
// to cublasZgemm (blas_level3/gemm.cc:46-83, :200-217).
//
// Same pipeline as gemm_f64.cu (TMA producer warp-group, 8 DMMA consumer warps, mbarrier ring,
// setmaxnreg), on interleaved (re,im) operands used IN PLACE -- no planar split, no extra pass:
//  * a complex matrix is addressed by TMA as a real matrix whose contiguous dimension is doubled,
//    so one 128-byte swizzle row holds 8 complex numbers; a stage is 8 complex k (BK=8);
//  * every fragment load is one LDS.128 (re,im) -- lane (g,t) takes k = 2t+s, which makes the
//    quarter-warp's eight 16-byte chunks hit eight different bank groups in both tile layouts;
//  * one complex 8x8x4 block product is four DMMA.8x8x4:  Cre += Are*Bre; Cre += (-Aim)*Bim;
//    Cim += Are*Bim; Cim += Aim*Bre.  Conjugation ('C') flips the sign of the imaginary fragment
//    as it is loaded (an integer XOR on the high word, off the FP64 pipe);
//  * CTA tile 64 x 128 complex, warp tile 32 x 32 complex = 64 DMMA accumulators (128 registers).
// Algorithmic flops 8mnk = 4 real products of 2mnk: exactly what the four DMMA streams issue, no padding work.
#include "common.cuh"
#include "kernels.h"
#include "gemm_generic.cuh"
#include "runtime.h"

namespace b200 {

constexpr int ZG_BK = 8;        // complex k per stage
constexpr int ZG_STAGES = 6;
enum { ZLAY_COL = 0, ZLAY_KC = 1 };

// byte offset of complex element (r, kk) in an operand tile (CU_TENSOR_MAP_SWIZZLE_128B pattern)
template <int LAY> __host__ __device__ __forceinline__ uint32_t ztile_off(int r, int kk) {
    if (LAY == ZLAY_COL)   // slabs of 8 complex rows: [kk][8 rows x 16 B], 1 KiB each
        return (uint32_t)((r >> 3) * 1024 + kk * 128 + (((r & 7) ^ kk) << 4));
    else                   // [r][8 complex k], 128 B per row
        return (uint32_t)(r * 128 + ((kk ^ (r & 7)) << 4));
}

struct ZgemmParams {
    int m, n, k;
    cuDoubleComplex alpha, beta;
    const cuDoubleComplex* A; int64_t lda;
    const cuDoubleComplex* B; int64_t ldb;
    cuDoubleComplex* C; int64_t ldc;
    int mask;
    int tiles_m, tiles_n;
};

__device__ __forceinline__ double2 lds_c64(const uint8_t* p) { return *reinterpret_cast<const double2*>(p); }

template <int MB, int NB, int LAYA, int LAYB, bool CONJA, bool CONJB>
__global__ void __launch_bounds__(384, 1)
zgemm_dmma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                  const ZgemmParams p) {
    constexpr int BM = 16 * MB, BN = 32 * NB;          // 2 consumer warps along m, 4 along n
    constexpr int A_BYTES = BM * ZG_BK * 16, B_BYTES = BN * ZG_BK * 16;
    constexpr int STAGE_BYTES = A_BYTES + B_BYTES;

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
    const int m0 = tile_m * BM, n0 = tile_n * BN;
    if (p.mask == MASK_LOWER && m0 + BM - 1 < n0) return;
    if (p.mask == MASK_UPPER && n0 + BN - 1 < m0) return;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = (uint64_t*)(smem + ZG_STAGES * STAGE_BYTES);
    const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + ZG_STAGES);
    const uint32_t smem_base = smem_u32(smem);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ktiles = (p.k + ZG_BK - 1) / ZG_BK;

    if (tid == 0) {
        for (int s = 0; s < ZG_STAGES; s++) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, 256);
        }
        mbar_fence_init();
        tma_prefetch_desc(&mapA); tma_prefetch_desc(&mapB);
    }
    __syncthreads();

    if (warp < 4) {
        // =========================== producer warp-group ===========================
        setmaxnreg_dec<40>();
        if (tid == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int kt = 0; kt < ktiles; kt++) {
                mbar_wait(empty0 + 8 * stage, phase ^ 1);
                const uint32_t fb = full0 + 8 * stage;
                const uint32_t sA = smem_base + stage * STAGE_BYTES, sB = sA + A_BYTES;
                mbar_expect_tx(fb, STAGE_BYTES);
                const int k0 = kt * ZG_BK;
                // coordinates are in doubles along the contiguous dimension (2 per complex)
                if (LAYA == ZLAY_COL) {
#pragma unroll
                    for (int sl = 0; sl < BM / 8; sl++) tma_load_2d(sA + sl * 1024, &mapA, 2 * (m0 + sl * 8), k0, fb);
                } else {
                    tma_load_2d(sA, &mapA, 2 * k0, m0, fb);
                }
                if (LAYB == ZLAY_COL) {
#pragma unroll
                    for (int sl = 0; sl < BN / 8; sl++) tma_load_2d(sB + sl * 1024, &mapB, 2 * (n0 + sl * 8), k0, fb);
                } else {
                    tma_load_2d(sB, &mapB, 2 * k0, n0, fb);
                }
                if (++stage == ZG_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }

    // =============================== consumer warps ===============================
    setmaxnreg_inc<232>();
    const int cw = warp - 4;
    const int wm0 = (cw & 1) * (8 * MB), wn0 = (cw >> 1) * (8 * NB);
    const int g = lane >> 2, tig = lane & 3;

    uint32_t offA[2], offB[2];
#pragma unroll
    for (int s = 0; s < 2; s++) {
        const int kk = 2 * tig + s;
        offA[s] = ztile_off<LAYA>(wm0 + g, kk);
        offB[s] = A_BYTES + ztile_off<LAYB>(wn0 + g, kk);
    }
    constexpr int BLK_STRIDE = 1024;   // next 8-row block: one COL slab, or 8 KC rows x 128 B

    double cre[MB][NB][2], cim[MB][NB][2];
#pragma unroll
    for (int i = 0; i < MB; i++)
#pragma unroll
        for (int j = 0; j < NB; j++) cre[i][j][0] = cre[i][j][1] = cim[i][j][0] = cim[i][j][1] = 0.0;

    {
        int stage = 0; uint32_t phase = 0;
        for (int kt = 0; kt < ktiles; kt++) {
            mbar_wait(full0 + 8 * stage, phase);
            const uint8_t* sS = smem + stage * STAGE_BYTES;
#pragma unroll
            for (int s = 0; s < 2; s++) {
                double2 a[MB], b[NB];
                double nai[MB];
#pragma unroll
                for (int i = 0; i < MB; i++) {
                    a[i] = lds_c64(sS + offA[s] + i * BLK_STRIDE);
                    if (CONJA) a[i].y = -a[i].y;
                    nai[i] = -a[i].y;
                }
#pragma unroll
                for (int j = 0; j < NB; j++) {
                    b[j] = lds_c64(sS + offB[s] + j * BLK_STRIDE);
                    if (CONJB) b[j].y = -b[j].y;
                }
#pragma unroll
                for (int i = 0; i < MB; i++)
#pragma unroll
                    for (int j = 0; j < NB; j++) {
                        dmma884(cre[i][j][0], cre[i][j][1], a[i].x, b[j].x);
                        dmma884(cim[i][j][0], cim[i][j][1], a[i].x, b[j].y);
                        dmma884(cre[i][j][0], cre[i][j][1], nai[i], b[j].y);
                        dmma884(cim[i][j][0], cim[i][j][1], a[i].y, b[j].x);
                    }
            }
            // generic-proxy reads of this stage must be fenced against the async-proxy (TMA) refill before the stage is
            // released (see gemm_f64.cu); every lane fences and arrives for its own reads
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(empty0 + 8 * stage);
            if (++stage == ZG_STAGES) { stage = 0; phase ^= 1; }
        }
    }

    // ---- epilogue: C = alpha*acc + beta*C on the kept region (old values loaded MB at a time before the stores
    // that may alias them, see gemm_f64.cu) ----
    const bool beta0 = (p.beta.x == 0.0 && p.beta.y == 0.0);
#pragma unroll
    for (int j = 0; j < NB; j++) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int64_t col = n0 + wn0 + 8 * j + 2 * tig + c;
            cuDoubleComplex* cp = p.C + col * p.ldc;
            double2 old[MB];
            bool ok[MB];
#pragma unroll
            for (int i = 0; i < MB; i++) {
                const int64_t row = m0 + wm0 + 8 * i + g;
                ok[i] = col < p.n && row < p.m && tri_keep(p.mask, row, col);
                old[i] = (!beta0 && ok[i]) ? *reinterpret_cast<const double2*>(cp + row) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int i = 0; i < MB; i++) {
                const int64_t row = m0 + wm0 + 8 * i + g;
                if (!ok[i]) continue;
                const double xr = cre[i][j][c], xi = cim[i][j][c];
                double vr = fma(p.alpha.x, xr, -(p.alpha.y * xi));
                double vi = fma(p.alpha.x, xi, p.alpha.y * xr);
                if (!beta0) {
                    vr += fma(p.beta.x, old[i].x, -(p.beta.y * old[i].y));
                    vi += fma(p.beta.x, old[i].y, p.beta.y * old[i].x);
                }
                *reinterpret_cast<double2*>(cp + row) = make_double2(vr, vi);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
static bool make_map_c64(CUtensorMap* map, const cuDoubleComplex* base, int layout, int64_t rows, int64_t kext, int64_t ld,
                         int tile_rows) {
    cuuint64_t gdim[2], gstride[1];
    cuuint32_t box[2], estr[2] = {1, 1};
    if (layout == ZLAY_COL) { gdim[0] = 2 * rows; gdim[1] = kext; box[0] = 16; box[1] = ZG_BK; }
    else                    { gdim[0] = 2 * kext; gdim[1] = rows; box[0] = 16; box[1] = tile_rows; }
    gstride[0] = (cuuint64_t)ld * 16;
    return encode_tensor_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, gdim, gstride, box, estr,
                             CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int MB, int NB, int LAYA, int LAYB, bool CA, bool CB>
static void launch_z(cudaStream_t s, const CUtensorMap& ma, const CUtensorMap& mb, const ZgemmParams& p) {
    constexpr int BM = 16 * MB, BN = 32 * NB;
    constexpr int SMEM = ZG_STAGES * (BM + BN) * ZG_BK * 16 + 2 * ZG_STAGES * 8 + 1024;
    auto kern = zgemm_dmma_kernel<MB, NB, LAYA, LAYB, CA, CB>;
    set_max_dynamic_smem((const void*)kern, SMEM);
    kern<<<p.tiles_m * p.tiles_n, 384, SMEM, s>>>(ma, mb, p);
}

// returns false when the operands cannot be described to TMA (caller falls back to the generic tile kernel)
static bool zgemm_dmma(cudaStream_t s, int oa, int ob, ZgemmParams p) {
    constexpr int MB = 4, NB = 4, BM = 16 * MB, BN = 32 * NB;
    if (!tma_available() || ((uintptr_t)p.A % 16) || ((uintptr_t)p.B % 16) || p.lda * 16 >= ((int64_t)1 << 40) ||
        p.ldb * 16 >= ((int64_t)1 << 40))
        return false;
    p.tiles_m = (p.m + BM - 1) / BM;
    p.tiles_n = (p.n + BN - 1) / BN;
    CUtensorMap ma, mb;
    memset(&ma, 0, sizeof ma); memset(&mb, 0, sizeof mb);
    const int la = oa == 0 ? ZLAY_COL : ZLAY_KC, lb = ob == 0 ? ZLAY_KC : ZLAY_COL;
    if (!make_map_c64(&ma, p.A, la, p.m, p.k, p.lda, BM) || !make_map_c64(&mb, p.B, lb, p.n, p.k, p.ldb, BN)) return false;
    const bool ca = oa == 2, cb = ob == 2;
#define B200_ZL(LA, LB)                                                                  \
    do {                                                                                 \
        if (ca && cb) launch_z<MB, NB, LA, LB, true, true>(s, ma, mb, p);                \
        else if (ca)  launch_z<MB, NB, LA, LB, true, false>(s, ma, mb, p);               \
        else if (cb)  launch_z<MB, NB, LA, LB, false, true>(s, ma, mb, p);               \
        else          launch_z<MB, NB, LA, LB, false, false>(s, ma, mb, p);              \
    } while (0)
    // conjugation only exists for transposed operands: 'N' layouts are instantiated without it
    if (la == ZLAY_COL && lb == ZLAY_KC) launch_z<MB, NB, ZLAY_COL, ZLAY_KC, false, false>(s, ma, mb, p);
    else if (la == ZLAY_KC && lb == ZLAY_KC) { if (ca) launch_z<MB, NB, ZLAY_KC, ZLAY_KC, true, false>(s, ma, mb, p); else launch_z<MB, NB, ZLAY_KC, ZLAY_KC, false, false>(s, ma, mb, p); }
    else if (la == ZLAY_COL && lb == ZLAY_COL) { if (cb) launch_z<MB, NB, ZLAY_COL, ZLAY_COL, false, true>(s, ma, mb, p); else launch_z<MB, NB, ZLAY_COL, ZLAY_COL, false, false>(s, ma, mb, p); }
    else B200_ZL(ZLAY_KC, ZLAY_COL);
#undef B200_ZL
    last_variant = VAR_DMMA_TMA;
    return true;
}

void zgemm_dev(cudaStream_t s, char ta, char tb, int m, int n, int k, cuDoubleComplex alpha, const cuDoubleComplex* A,
               int64_t lda, const cuDoubleComplex* B, int64_t ldb, cuDoubleComplex beta, cuDoubleComplex* C,
               int64_t ldc, int mask) {
    if (m <= 0 || n <= 0) return;
    if (num<cuDoubleComplex>::is_zero(alpha) || k <= 0) { scale_matrix<cuDoubleComplex>(s, m, n, beta, C, ldc, mask); last_variant = VAR_SCALE_ONLY; return; }
    int variant = force_variant;
    if (variant == VAR_NONE) variant = ((double)m * n * k < 24.0 * 24.0 * 24.0) ? VAR_GENERIC_TILE : VAR_DMMA_TMA;
    if (variant == VAR_DMMA_TMA || variant == VAR_DMMA_LDG) {
        ZgemmParams p;
        p.m = m; p.n = n; p.k = k; p.alpha = alpha; p.beta = beta;
        p.A = A; p.lda = lda; p.B = B; p.ldb = ldb; p.C = C; p.ldc = ldc; p.mask = mask; p.tiles_m = p.tiles_n = 0;
        if (zgemm_dmma(s, op_code(ta), op_code(tb), p)) return;
    }
    gemm_generic_launch<cuDoubleComplex>(s, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, mask);
}

}  // namespace b200
