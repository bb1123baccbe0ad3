// multi_gemm.h -- Level-3 calls partitioned over the GPUs of one box from inside the interposed symbol (multi_gemm.cu).
#pragma once
#include <stdint.h>
#include <vector>

namespace b200 {

// One hop of one piece of an operand panel: kind 0 = rows [off, off+len) of grid row gidx's block of op(A), 1 = columns of grid
// column gidx's block of op(B); src = slot that forwards it (-1: the origin -- home HBM or host memory), dst = receiving slot.
struct MgHop { int kind, gidx, piece; int64_t off, len; int src, dst; };
void mg_grid(int ndev, int* P, int* Q);
void mg_block_range(int64_t total, int parts, int idx, int64_t* lo, int64_t* hi);
int64_t mg_a_group(int64_t tile_rows, int pieces = 16);
int64_t mg_b_group();
std::vector<MgHop> mg_plan(int ndev, int64_t m, int64_t n, bool host_source, int a_pieces = 16);

struct MgStats { unsigned long long calls, devices, origin_bytes, forward_bytes, hops; };
extern MgStats g_mg_stats;

// false: the call is not partitioned (devices < 2, too small, mixed residency, no peer access) and takes the 1-GPU path
template <typename T>
bool multi_gemm(char ta, char tb, int m, int n, int k, T alpha, const T* a, int64_t lda, const T* b, int64_t ldb, T beta, T* c, int64_t ldc);

// multi_level3.cu: ?syrk_ on equal-area strips of the triangle, ?trsm_ / ?trmm_ on blocks of independent right-hand sides.  Same
// contract: false = not partitioned.  (uplo, trans, side, diag already normalised to upper case; trans 'N' | 'T' | 'C'.)
template <typename T>
bool multi_syrk(char uplo, char trans, int n, int k, T alpha, const T* a, int64_t lda, T beta, T* c, int64_t ldc);
template <typename T>
bool multi_trxm(bool solve, char side, char uplo, char trans, char diag, int m, int n, T alpha, const T* a, int64_t lda, T* b, int64_t ldb);
void ml3_strips(int64_t n, int ndev, int64_t* bounds);                       // ndev + 1 strip boundaries of equal referenced area
std::vector<MgHop> ml3_syrk_plan(int ndev, int64_t n, bool host_source);     // row pieces of op(A): gidx = strip, off = global row
std::vector<MgHop> ml3_tri_plan(int ndev, int64_t na, bool host_source);     // column groups of the triangle: off = first column

// Blocked Cholesky workload (lower, in place, `a` device-accessible on the home GPU) over ndev devices; returns LAPACK info.
int multi_cholesky_lower(int n, double* a, int64_t lda, int nb, int ndev);

}  // namespace b200
