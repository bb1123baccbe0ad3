// fortran_l3.cu -- Fortran-ABI Level-3 entry points: netlib argument checks and quick returns on
// the host (reference blas_level3/*.cc `xxx_check`), operand residency, device kernel, synchronous
// return.  No size cutoff to a CPU BLAS (reference gemm.cc:129-141 gemm_perf_check): small
// problems pick a small-tile GPU variant instead.
#include "abi_common.h"
#include "staged_gemm.cuh"
#include "staged_level3.cuh"
#include "multi_gemm.h"
#include <type_traits>
#include "../../include/b200blas.h"

using namespace b200;

namespace {

template <typename T> struct GemmDev;
template <> struct GemmDev<float> { static constexpr auto fn = sgemm_dev; };
template <> struct GemmDev<double> { static constexpr auto fn = dgemm_dev; };
template <> struct GemmDev<cuFloatComplex> { static constexpr auto fn = cgemm_dev; };
template <> struct GemmDev<cuDoubleComplex> { static constexpr auto fn = zgemm_dev; };

// reference gemm.cc:87-127 (gemm_check) + :46-83 (_b2c_gemm)
template <typename T>
void gemm_entry(const char* name, const char* transa, const char* transb, const int* m, const int* n, const int* k,
                const T* alpha, const T* a, const int* lda, const T* b, const int* ldb, const T* beta, T* c,
                const int* ldc) {
    const bool nota = lsame(transa, 'N'), notb = lsame(transb, 'N');
    const int nrowa = nota ? *m : *k, nrowb = notb ? *k : *n;
    int info = 0;
    if (!nota && !lsame(transa, 'C') && !lsame(transa, 'T')) info = 1;
    else if (!notb && !lsame(transb, 'C') && !lsame(transb, 'T')) info = 2;
    else if (*m < 0) info = 3;
    else if (*n < 0) info = 4;
    else if (*k < 0) info = 5;
    else if (*lda < imax(1, nrowa)) info = 8;
    else if (*ldb < imax(1, nrowb)) info = 10;
    else if (*ldc < imax(1, *m)) info = 13;
    if (info) { call_xerbla(name, info); return; }
    if (*m == 0 || *n == 0 || ((is0(*alpha) || *k == 0) && is1(*beta))) return;

    CallScope scope(name);
    const bool scale_only = is0(*alpha) || *k == 0;
    const char ta = nota ? 'N' : (lsame(transa, 'T') ? 'T' : 'C');
    const char tb = notb ? 'N' : (lsame(transb, 'T') ? 'T' : 'C');
    // devices=<n>: large products are 2-D tile-partitioned over the GPUs of the box right here, behind the symbol (multi_gemm.cu)
    if (!scale_only && g_opts.devices > 1 && multi_gemm<T>(ta, tb, *m, *n, *k, *alpha, a, (int64_t)*lda, b, (int64_t)*ldb, *beta, c, (int64_t)*ldc)) {
        log_exec(name, "%c%c m=%d n=%d k=%d lda=%d ldb=%d ldc=%d (partitioned over %d devices)", ta, tb, *m, *n, *k, *lda, *ldb, *ldc, g_opts.devices);
        return;
    }
    // tracked managed operands the CPU has just filled: range-by-range migration with the multiply behind it (staged_gemm.cuh)
    if (!scale_only && gemm_first_touch<T>(GemmDev<T>::fn, ta, tb, *m, *n, *k, *alpha, a, (int64_t)*lda, b, (int64_t)*ldb, *beta, c, (int64_t)*ldc)) {
        log_exec(name, "%c%c m=%d n=%d k=%d lda=%d ldb=%d ldc=%d (first-touch migration under the multiply)", ta, tb, *m, *n, *k, *lda, *ldb, *ldc);
        return;
    }
    // large host-resident operands: chunked staging overlapped with the multiply (staged_gemm.cuh)
    if (!scale_only && gemm_pipelined<T>(GemmDev<T>::fn, ta, tb, *m, *n, *k, *alpha, a, (int64_t)*lda, b, (int64_t)*ldb, *beta, c, (int64_t)*ldc)) {
        log_exec(name, "%c%c m=%d n=%d k=%d lda=%d ldb=%d ldc=%d (pipelined staging)", ta, tb, *m, *n, *k, *lda, *ldb, *ldc);
        return;
    }
    // A is lda x (nota ? k : m), B is ldb x (notb ? n : k)  (reference compute_size, gemm.cc:36-44,
    // done in 64-bit here: the reference's int byte counts overflow at n=16384)
    Operand oa(scale_only ? nullptr : a, nrowa, nota ? *k : *m, *lda, sizeof(T), ACC_IN);
    Operand ob(scale_only ? nullptr : b, nrowb, notb ? *n : *k, *ldb, sizeof(T), ACC_IN);
    Operand oc(c, *m, *n, *ldc, sizeof(T), is0(*beta) ? ACC_OUT : ACC_INOUT);
    GemmDev<T>::fn(current_stream(), ta, tb, *m, *n, *k, *alpha, (const T*)oa.dev(), oa.ld(), (const T*)ob.dev(), ob.ld(),
                   *beta, (T*)oc.dev(), oc.ld(), MASK_FULL);
    oc.release();
    log_exec(name, "%c%c m=%d n=%d k=%d lda=%d ldb=%d ldc=%d", ta, tb, *m, *n, *k, *lda, *ldb, *ldc);
}

// reference syrk.cc:78-146 (syrk_check: info 1,2,3,4,7,10) + :43-76.  The reference's alpha==0 host
// loops are replaced by the device scale kernel (and its upper/beta!=0 loop, syrk.cc:124-128, which
// scales the wrong triangle, is not reproduced).  Complex SYRK rejects 'C' like netlib ZSYRK.
template <typename T>
void syrk_entry(const char* name, bool cplx, const char* uplo, const char* trans, const int* n, const int* k, const T* alpha,
                const T* a, const int* lda, const T* beta, T* c, const int* ldc) {
    const bool upper = lsame(uplo, 'U'), nota = lsame(trans, 'N');
    const int nrowa = nota ? *n : *k;
    int info = 0;
    if (!upper && !lsame(uplo, 'L')) info = 1;
    else if (!nota && !lsame(trans, 'T') && (cplx || !lsame(trans, 'C'))) info = 2;
    else if (*n < 0) info = 3;
    else if (*k < 0) info = 4;
    else if (*lda < imax(1, nrowa)) info = 7;
    else if (*ldc < imax(1, *n)) info = 10;
    if (info) { call_xerbla(name, info); return; }
    if (*n == 0 || ((is0(*alpha) || *k == 0) && is1(*beta))) return;
    CallScope scope(name);
    const bool scale_only = is0(*alpha) || *k == 0;
    // devices=<n>: equal-area strips of the triangle, one masked GEMM per device (multi_level3.cu)
    if (!scale_only && g_opts.devices > 1 && multi_syrk<T>(upper ? 'U' : 'L', nota ? 'N' : 'T', *n, *k, *alpha, a, (int64_t)*lda, *beta, c, (int64_t)*ldc)) {
        log_exec(name, "%c%c n=%d k=%d lda=%d ldc=%d (partitioned over %d devices)", upper ? 'U' : 'L', nota ? 'N' : 'T', *n, *k, *lda, *ldc, g_opts.devices);
        return;
    }
    // large host-resident operands: k-chunks of A in, finished trapezoids of C out, under the multiply (staged_level3.cuh)
    if (!scale_only && syrk_pipelined<T>(upper ? 'U' : 'L', nota ? 'N' : 'T', *n, *k, *alpha, a, (int64_t)*lda, *beta, c, (int64_t)*ldc)) {
        log_exec(name, "%c%c n=%d k=%d lda=%d ldc=%d (pipelined staging)", upper ? 'U' : 'L', nota ? 'N' : 'T', *n, *k, *lda, *ldc);
        return;
    }
    Operand oa(scale_only ? nullptr : a, nrowa, nota ? *k : *n, *lda, sizeof(T), ACC_IN);
    // C is read even when beta == 0: the unreferenced triangle must survive a staged round trip
    Operand oc(c, *n, *n, *ldc, sizeof(T), ACC_INOUT);
    syrk_dev<T>(current_stream(), upper ? 'U' : 'L', nota ? 'N' : 'T', *n, *k, *alpha, (const T*)oa.dev(), oa.ld(), *beta, (T*)oc.dev(), oc.ld());
    oc.release();
    log_exec(name, "%c%c n=%d k=%d lda=%d ldc=%d", upper ? 'U' : 'L', nota ? 'N' : 'T', *n, *k, *lda, *ldc);
}

// reference trsm.cc:76-132 / trmm.cc:81-137 (info 1,2,3,4,5,6,9,11) + trsm.cc:40-73 / trmm.cc:42-79
template <typename T>
void tr_entry(const char* name, bool solve, const char* side, const char* uplo, const char* transa, const char* diag, const int* m,
              const int* n, const T* alpha, const T* a, const int* lda, T* b, const int* ldb) {
    const bool lside = lsame(side, 'L'), upper = lsame(uplo, 'U');
    const int nrowa = lside ? *m : *n;
    int info = 0;
    if (!lside && !lsame(side, 'R')) info = 1;
    else if (!upper && !lsame(uplo, 'L')) info = 2;
    else if (!lsame(transa, 'N') && !lsame(transa, 'T') && !lsame(transa, 'C')) info = 3;
    else if (!lsame(diag, 'U') && !lsame(diag, 'N')) info = 4;
    else if (*m < 0) info = 5;
    else if (*n < 0) info = 6;
    else if (*lda < imax(1, nrowa)) info = 9;
    else if (*ldb < imax(1, *m)) info = 11;
    if (info) { call_xerbla(name, info); return; }
    if (*m == 0 || *n == 0) return;
    CallScope scope(name);
    const char t = lsame(transa, 'N') ? 'N' : (lsame(transa, 'T') ? 'T' : 'C');
    // devices=<n>: independent right-hand sides, one block per device (multi_level3.cu)
    if (g_opts.devices > 1 && multi_trxm<T>(solve, lside ? 'L' : 'R', upper ? 'U' : 'L', t, lsame(diag, 'U') ? 'U' : 'N', *m, *n, *alpha, a, (int64_t)*lda, b, (int64_t)*ldb)) {
        log_exec(name, "%c%c%c%c m=%d n=%d lda=%d ldb=%d (partitioned over %d devices)", lside ? 'L' : 'R', upper ? 'U' : 'L', t, lsame(diag, 'U') ? 'U' : 'N', *m, *n, *lda, *ldb, g_opts.devices);
        return;
    }
    // large host-resident B: blocks of right-hand sides travel in and out under the solve / product (staged_level3.cuh)
    if (trxm_pipelined<T>(solve, lside ? 'L' : 'R', upper ? 'U' : 'L', t, lsame(diag, 'U') ? 'U' : 'N', *m, *n, *alpha, a, (int64_t)*lda, b, (int64_t)*ldb)) {
        log_exec(name, "%c%c%c%c m=%d n=%d lda=%d ldb=%d (pipelined staging)", lside ? 'L' : 'R', upper ? 'U' : 'L', t, lsame(diag, 'U') ? 'U' : 'N', *m, *n, *lda, *ldb);
        return;
    }
    Operand oa(is0(*alpha) ? nullptr : a, nrowa, nrowa, *lda, sizeof(T), ACC_IN);
    Operand ob(b, *m, *n, *ldb, sizeof(T), is0(*alpha) ? ACC_OUT : ACC_INOUT);
    if (solve) trsm_dev<T>(current_stream(), lside ? 'L' : 'R', upper ? 'U' : 'L', t, lsame(diag, 'U') ? 'U' : 'N', *m, *n, *alpha, (const T*)oa.dev(), oa.ld(), (T*)ob.dev(), ob.ld());
    else trmm_dev<T>(current_stream(), lside ? 'L' : 'R', upper ? 'U' : 'L', t, lsame(diag, 'U') ? 'U' : 'N', *m, *n, *alpha, (const T*)oa.dev(), oa.ld(), (T*)ob.dev(), ob.ld());
    ob.release();
    log_exec(name, "%c%c%c%c m=%d n=%d lda=%d ldb=%d", lside ? 'L' : 'R', upper ? 'U' : 'L', t, lsame(diag, 'U') ? 'U' : 'N', *m, *n, *lda, *ldb);
}

// reference symm.cc:70-112 / hemm.cc:68-110 (info 1,2,3,4,7,9,12)
template <typename T>
void symm_entry(const char* name, bool herm, const char* side, const char* uplo, const int* m, const int* n, const T* alpha, const T* a,
                const int* lda, const T* b, const int* ldb, const T* beta, T* c, const int* ldc) {
    const bool lside = lsame(side, 'L'), upper = lsame(uplo, 'U');
    const int nrowa = lside ? *m : *n;
    int info = 0;
    if (!lside && !lsame(side, 'R')) info = 1;
    else if (!upper && !lsame(uplo, 'L')) info = 2;
    else if (*m < 0) info = 3;
    else if (*n < 0) info = 4;
    else if (*lda < imax(1, nrowa)) info = 7;
    else if (*ldb < imax(1, *m)) info = 9;
    else if (*ldc < imax(1, *m)) info = 12;
    if (info) { call_xerbla(name, info); return; }
    if (*m == 0 || *n == 0 || (is0(*alpha) && is1(*beta))) return;
    CallScope scope(name);
    const bool scale_only = is0(*alpha);
    Operand oa(scale_only ? nullptr : a, nrowa, nrowa, *lda, sizeof(T), ACC_IN);
    Operand ob(scale_only ? nullptr : b, *m, *n, *ldb, sizeof(T), ACC_IN);
    Operand oc(c, *m, *n, *ldc, sizeof(T), is0(*beta) ? ACC_OUT : ACC_INOUT);
    symm_dev<T>(current_stream(), herm, lside ? 'L' : 'R', upper ? 'U' : 'L', *m, *n, *alpha, (const T*)oa.dev(), oa.ld(), (const T*)ob.dev(), ob.ld(),
                *beta, (T*)oc.dev(), oc.ld());
    oc.release();
    log_exec(name, "%c%c m=%d n=%d lda=%d ldb=%d ldc=%d", lside ? 'L' : 'R', upper ? 'U' : 'L', *m, *n, *lda, *ldb, *ldc);
}

// reference syr2k.cc:67-118 / her2k.cc:72-124 (info 1,2,3,4,7,9,12).  herm: trans in {N,C}, real beta (RB = real type).
template <typename T, typename RB>
void r2k_entry(const char* name, bool herm, bool cplx, const char* uplo, const char* trans, const int* n, const int* k, const T* alpha, const T* a,
               const int* lda, const T* b, const int* ldb, const RB* beta, T* c, const int* ldc) {
    const bool upper = lsame(uplo, 'U'), nota = lsame(trans, 'N');
    const int nrowa = nota ? *n : *k;
    int info = 0;
    const bool trans_ok = nota || (herm ? lsame(trans, 'C') : (lsame(trans, 'T') || (!cplx && lsame(trans, 'C'))));
    if (!upper && !lsame(uplo, 'L')) info = 1;
    else if (!trans_ok) info = 2;
    else if (*n < 0) info = 3;
    else if (*k < 0) info = 4;
    else if (*lda < imax(1, nrowa)) info = 7;
    else if (*ldb < imax(1, nrowa)) info = 9;
    else if (*ldc < imax(1, *n)) info = 12;
    if (info) { call_xerbla(name, info); return; }
    if (*n == 0 || ((is0(*alpha) || *k == 0) && is1(*beta))) return;
    CallScope scope(name);
    const bool scale_only = is0(*alpha) || *k == 0;
    Operand oa(scale_only ? nullptr : a, nrowa, nota ? *k : *n, *lda, sizeof(T), ACC_IN);
    Operand ob(scale_only ? nullptr : b, nrowa, nota ? *k : *n, *ldb, sizeof(T), ACC_IN);
    Operand oc(c, *n, *n, *ldc, sizeof(T), ACC_INOUT);      // the unreferenced triangle must survive a staged round trip
    if constexpr (std::is_same<T, RB>::value)
        syr2k_dev<T>(current_stream(), upper ? 'U' : 'L', nota ? 'N' : 'T', *n, *k, *alpha, (const T*)oa.dev(), oa.ld(), (const T*)ob.dev(), ob.ld(), *beta,
                     (T*)oc.dev(), oc.ld());
    else
        her2k_dev<T, RB>(current_stream(), upper ? 'U' : 'L', nota ? 'N' : 'C', *n, *k, *alpha, (const T*)oa.dev(), oa.ld(), (const T*)ob.dev(), ob.ld(),
                         *beta, (T*)oc.dev(), oc.ld());
    oc.release();
    log_exec(name, "%c%c n=%d k=%d lda=%d ldb=%d ldc=%d", upper ? 'U' : 'L', nota ? 'N' : 'T', *n, *k, *lda, *ldb, *ldc);
}

// reference herk.cc:64-121 (info 1,2,3,4,7,10; real alpha and beta; trans in {N,C})
template <typename T, typename RB>
void herk_entry(const char* name, const char* uplo, const char* trans, const int* n, const int* k, const RB* alpha, const T* a, const int* lda,
                const RB* beta, T* c, const int* ldc) {
    const bool upper = lsame(uplo, 'U'), nota = lsame(trans, 'N');
    const int nrowa = nota ? *n : *k;
    int info = 0;
    if (!upper && !lsame(uplo, 'L')) info = 1;
    else if (!nota && !lsame(trans, 'C')) info = 2;
    else if (*n < 0) info = 3;
    else if (*k < 0) info = 4;
    else if (*lda < imax(1, nrowa)) info = 7;
    else if (*ldc < imax(1, *n)) info = 10;
    if (info) { call_xerbla(name, info); return; }
    if (*n == 0 || ((*alpha == RB(0) || *k == 0) && *beta == RB(1))) return;
    CallScope scope(name);
    const bool scale_only = *alpha == RB(0) || *k == 0;
    Operand oa(scale_only ? nullptr : a, nrowa, nota ? *k : *n, *lda, sizeof(T), ACC_IN);
    Operand oc(c, *n, *n, *ldc, sizeof(T), ACC_INOUT);
    herk_dev<T, RB>(current_stream(), upper ? 'U' : 'L', nota ? 'N' : 'C', *n, *k, *alpha, (const T*)oa.dev(), oa.ld(), *beta, (T*)oc.dev(), oc.ld());
    oc.release();
    log_exec(name, "%c%c n=%d k=%d lda=%d ldc=%d", upper ? 'U' : 'L', nota ? 'N' : 'C', *n, *k, *lda, *ldc);
}

}  // namespace

extern "C" {

// D := alpha*op(A)*op(B) + beta*C on DEVICE pointers, D distinct from C allowed (and D may be a CUDA-IPC
// peer mapping).  Building block of the partitioned multi-GPU GEMM; arguments by value, no staging.
void b200blas_dgemm_out(char transa, char transb, int m, int n, int k, double alpha, const double* a, long long lda,
                        const double* b, long long ldb, double beta, const double* c, long long ldc, double* d, long long ldd) {
    CallScope scope;
    dgemm_out_dev(current_stream(), transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, d, ldd, MASK_FULL);
}

// Same, for a rank of the partitioned GEMM whose operand panels are still arriving: a CTA reads its tile's rows of
// op(A) only after aflags[row / a_group] >= epoch and its columns of op(B) only after bflags[col / b_group] >= epoch
// (flags live in this device's memory and are written remotely by the home GPU's copy engine).
void b200blas_dgemm_out_flagged(char transa, char transb, int m, int n, int k, double alpha, const double* a, long long lda,
                                const double* b, long long ldb, double beta, const double* c, long long ldc, double* d, long long ldd,
                                const unsigned* aflags, int a_group, const unsigned* bflags, int b_group, unsigned epoch) {
    CallScope scope;
    if ((aflags && (a_group <= 0 || a_group % 128)) || (bflags && (b_group <= 0 || b_group % 128)))
        fatal("b200blas_dgemm_out_flagged", __FILE__, __LINE__, "flag group sizes must be positive multiples of 128");
    dgemm_set_panel_flags(aflags, a_group, bflags, b_group, epoch);
    dgemm_out_dev(current_stream(), transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc, d, ldd, MASK_FULL);
}

// ---- partitioned-call introspection (tests, bench) ----
// The hop list a partitioned call of this shape would issue, 7 ints per hop: kind, grid index, piece, offset, length, src, dst.
// Returns the number of hops (may exceed cap: only cap hops are written).  Pure host logic, touches no device.
int b200blas_mg_plan(int ndev, long long m, long long n, int host_source, int* out, int cap) {
    const std::vector<MgHop> plan = mg_plan(ndev, m, n, host_source != 0);
    int w = 0;
    for (const MgHop& h : plan) {
        if (w < cap) {
            int* o = out + 7 * w;
            o[0] = h.kind; o[1] = h.gidx; o[2] = h.piece; o[3] = (int)h.off; o[4] = (int)h.len; o[5] = h.src; o[6] = h.dst;
        }
        w++;
    }
    return w;
}
// The same for ?syrk_ (which = 0: row pieces of op(A), n = order of C) and ?trsm_/?trmm_ (which = 1: column groups of the triangle, n = its order)
int b200blas_ml3_plan(int which, int ndev, long long n, int host_source, int* out, int cap) {
    const std::vector<MgHop> plan = which == 0 ? ml3_syrk_plan(ndev, n, host_source != 0) : ml3_tri_plan(ndev, n, host_source != 0);
    int w = 0;
    for (const MgHop& h : plan) {
        if (w < cap) {
            int* o = out + 7 * w;
            o[0] = h.kind; o[1] = h.gidx; o[2] = h.piece; o[3] = (int)h.off; o[4] = (int)h.len; o[5] = h.src; o[6] = h.dst;
        }
        w++;
    }
    return w;
}
// strip boundaries of a partitioned ?syrk_: out[0..ndev]
void b200blas_ml3_strips(long long n, int ndev, long long* out) {
    int64_t b[kMaxDevices + 1];
    if (ndev < 1 || ndev > kMaxDevices) return;
    ml3_strips(n, ndev, b);
    for (int i = 0; i <= ndev; i++) out[i] = b[i];
}
void b200blas_mg_geometry(int ndev, long long m, long long n, int slot, long long* out4) {
    int P, Q; mg_grid(ndev, &P, &Q);
    int64_t r0, r1, c0, c1;
    mg_block_range(m, P, slot / Q, &r0, &r1); mg_block_range(n, Q, slot % Q, &c0, &c1);
    out4[0] = r0; out4[1] = r1; out4[2] = c0; out4[3] = c1;
}
// Blocked right-looking Cholesky workload (BASELINE.json configs[3]) in one call: lower factor in place, a on the device /
// managed; block size nb; runs on `devices=<n>` GPUs (1: look-ahead on one GPU).  Returns LAPACK info.
int b200blas_cholesky_lower(int n, double* a, long long lda, int nb) {
    if (n < 0) return -1;
    if (!a && n > 0) return -2;
    if (lda < (n > 1 ? n : 1)) return -3;
    CallScope scope("cholesky_lower");
    const Residency r = classify(a);
    if (r == RES_HOST_PINNED || r == RES_HOST_PAGEABLE) {
        Operand oa(a, n, n, lda, sizeof(double), ACC_INOUT);
        const int info = multi_cholesky_lower(n, (double*)oa.dev(), oa.ld(), nb, 1);
        oa.release();
        return info;
    }
    if (r == RES_MANAGED) make_resident(a, (size_t)((int64_t)(n - 1) * lda + n) * 8, current_stream());
    return multi_cholesky_lower(n, a, lda, nb, g_opts.devices);
}
void b200blas_mg_stats(unsigned long long* out5) {
    out5[0] = g_mg_stats.calls; out5[1] = g_mg_stats.devices; out5[2] = g_mg_stats.origin_bytes; out5[3] = g_mg_stats.forward_bytes; out5[4] = g_mg_stats.hops;
}

void sgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const float* alpha,
            const float* a, const int* lda, const float* b, const int* ldb, const float* beta, float* c, const int* ldc) {
    gemm_entry<float>("sgemm_", transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
void dgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const double* alpha,
            const double* a, const int* lda, const double* b, const int* ldb, const double* beta, double* c, const int* ldc) {
    gemm_entry<double>("dgemm_", transa, transb, m, n, k, alpha, a, lda, b, ldb, beta, c, ldc);
}
void cgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const b200_c32* alpha,
            const b200_c32* a, const int* lda, const b200_c32* b, const int* ldb, const b200_c32* beta, b200_c32* c, const int* ldc) {
    gemm_entry<cuFloatComplex>("cgemm_", transa, transb, m, n, k, (const cuFloatComplex*)alpha, (const cuFloatComplex*)a, lda,
                               (const cuFloatComplex*)b, ldb, (const cuFloatComplex*)beta, (cuFloatComplex*)c, ldc);
}
void zgemm_(const char* transa, const char* transb, const int* m, const int* n, const int* k, const b200_c64* alpha,
            const b200_c64* a, const int* lda, const b200_c64* b, const int* ldb, const b200_c64* beta, b200_c64* c, const int* ldc) {
    gemm_entry<cuDoubleComplex>("zgemm_", transa, transb, m, n, k, (const cuDoubleComplex*)alpha, (const cuDoubleComplex*)a, lda,
                                (const cuDoubleComplex*)b, ldb, (const cuDoubleComplex*)beta, (cuDoubleComplex*)c, ldc);
}

#define B200_SYRK(P, T, CT, CPLX)                                                                                       \
    void P##syrk_(const char* uplo, const char* trans, const int* n, const int* k, const T* alpha, const T* a, const int* lda, \
                  const T* beta, T* c, const int* ldc) {                                                                \
        syrk_entry<CT>(#P "syrk_", CPLX, uplo, trans, n, k, (const CT*)alpha, (const CT*)a, lda, (const CT*)beta, (CT*)c, ldc); \
    }
B200_SYRK(s, float, float, false)
B200_SYRK(d, double, double, false)
B200_SYRK(c, b200_c32, cuFloatComplex, true)
B200_SYRK(z, b200_c64, cuDoubleComplex, true)
#define B200_TR(P, T, CT)                                                                                               \
    void P##trsm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n, \
                  const T* alpha, const T* a, const int* lda, T* b, const int* ldb) {                                   \
        tr_entry<CT>(#P "trsm_", true, side, uplo, transa, diag, m, n, (const CT*)alpha, (const CT*)a, lda, (CT*)b, ldb); \
    }                                                                                                                   \
    void P##trmm_(const char* side, const char* uplo, const char* transa, const char* diag, const int* m, const int* n, \
                  const T* alpha, const T* a, const int* lda, T* b, const int* ldb) {                                   \
        tr_entry<CT>(#P "trmm_", false, side, uplo, transa, diag, m, n, (const CT*)alpha, (const CT*)a, lda, (CT*)b, ldb); \
    }
B200_TR(s, float, float)
B200_TR(d, double, double)
B200_TR(c, b200_c32, cuFloatComplex)
B200_TR(z, b200_c64, cuDoubleComplex)

#define B200_SYMM(P, NAME, T, CT, HERM)                                                                                 \
    void P##NAME##_(const char* side, const char* uplo, const int* m, const int* n, const T* alpha, const T* a, const int* lda, \
                    const T* b, const int* ldb, const T* beta, T* c, const int* ldc) {                                  \
        symm_entry<CT>(#P #NAME "_", HERM, side, uplo, m, n, (const CT*)alpha, (const CT*)a, lda, (const CT*)b, ldb, (const CT*)beta, (CT*)c, ldc); \
    }
B200_SYMM(s, symm, float, float, false)
B200_SYMM(d, symm, double, double, false)
B200_SYMM(c, symm, b200_c32, cuFloatComplex, false)
B200_SYMM(z, symm, b200_c64, cuDoubleComplex, false)
B200_SYMM(c, hemm, b200_c32, cuFloatComplex, true)
B200_SYMM(z, hemm, b200_c64, cuDoubleComplex, true)
#define B200_SYR2K(P, T, CT, CPLX)                                                                                      \
    void P##syr2k_(const char* uplo, const char* trans, const int* n, const int* k, const T* alpha, const T* a, const int* lda, \
                   const T* b, const int* ldb, const T* beta, T* c, const int* ldc) {                                   \
        r2k_entry<CT, CT>(#P "syr2k_", false, CPLX, uplo, trans, n, k, (const CT*)alpha, (const CT*)a, lda, (const CT*)b, ldb, (const CT*)beta, (CT*)c, ldc); \
    }
B200_SYR2K(s, float, float, false)
B200_SYR2K(d, double, double, false)
B200_SYR2K(c, b200_c32, cuFloatComplex, true)
B200_SYR2K(z, b200_c64, cuDoubleComplex, true)
void cher2k_(const char* uplo, const char* trans, const int* n, const int* k, const b200_c32* alpha, const b200_c32* a, const int* lda,
             const b200_c32* b, const int* ldb, const float* beta, b200_c32* c, const int* ldc) {
    r2k_entry<cuFloatComplex, float>("cher2k_", true, true, uplo, trans, n, k, (const cuFloatComplex*)alpha, (const cuFloatComplex*)a, lda, (const cuFloatComplex*)b, ldb, beta, (cuFloatComplex*)c, ldc);
}
void zher2k_(const char* uplo, const char* trans, const int* n, const int* k, const b200_c64* alpha, const b200_c64* a, const int* lda,
             const b200_c64* b, const int* ldb, const double* beta, b200_c64* c, const int* ldc) {
    r2k_entry<cuDoubleComplex, double>("zher2k_", true, true, uplo, trans, n, k, (const cuDoubleComplex*)alpha, (const cuDoubleComplex*)a, lda, (const cuDoubleComplex*)b, ldb, beta, (cuDoubleComplex*)c, ldc);
}
void cherk_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha, const b200_c32* a, const int* lda, const float* beta,
            b200_c32* c, const int* ldc) {
    herk_entry<cuFloatComplex, float>("cherk_", uplo, trans, n, k, alpha, (const cuFloatComplex*)a, lda, beta, (cuFloatComplex*)c, ldc);
}
void zherk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha, const b200_c64* a, const int* lda, const double* beta,
            b200_c64* c, const int* ldc) {
    herk_entry<cuDoubleComplex, double>("zherk_", uplo, trans, n, k, alpha, (const cuDoubleComplex*)a, lda, beta, (cuDoubleComplex*)c, ldc);
}

}  // extern "C"
