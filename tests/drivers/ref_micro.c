/* ref_micro.c -- the reference's other micro-tests (tests/c/copy.c, dsdot.c, rot.c, gbmv.c, trmv.c, trsm.c, hemm.c; the
 * harness is tests/c/test.c) restated as ONE driver.  Each of them allocates its operands with calloc, fills them with a
 * closed-form pattern, calls one cblas_* routine and dumps the result for eyeballing -- no expected values, no assertion
 * (SURVEY.md section 4).  Here the same fills and the same calls; the result goes to a binary file so a test can compare
 * the run against the CPU BLAS (plain) with the run under LD_PRELOAD=libb200blas.so, and where the pattern has a closed
 * form it is checked in place:
 *     copy  (copy.c:26-32)   x[i] = i+1;                               cblas_scopy          => y == x
 *     dsdot (dsdot.c:26-33)  x[i] = y[i] = i+1;                        cblas_dsdot          => n(n+1)(2n+1)/6
 *     rot   (rot.c:27-36)    x[i] = y[i] = i + (i+1)I, c = s = 1;      cblas_csrot          => x' = 2x, y' = 0
 *     gbmv  (gbmv.c:36-50)   x[i] = y[i] = i, ones on a kl = ku = 2 band, alpha = beta = 1;  cblas_sgbmv (ColMajor, NoTrans)
 *     trmv  (trmv.c:24-32)   upper A(r,c) = (r n + c) mod (n^2/10), x[i] = i;               cblas_strmv (Upper, NoTrans, NonUnit)
 *     trsm  (trsm.c:21-31)   upper A(r,c) = (r m + c) mod 10, B(r,c) = (r n + c) mod 10;    cblas_dtrsm (Left, Upper, NoTrans, NonUnit)
 *                            -- the reference's fill puts zeros on A's diagonal (its own FIXME: "segfault when m > 100");
 *                               here the diagonal gets +10 so the system is solvable and the two runs can be compared
 *     gemm  (gemm.c:29-43)   the ORIGINAL single-precision form of BASELINE config 1: A(r,c) = r, B(r,c) = c, alpha = 1, beta = 0;
 *                            cblas_sgemm (ColMajor, NoTrans, NoTrans) => C(r,c) = k r c, exact in f32 while n^3 < 2^24 (n <= 256)
 *     hemm  (hemm.c:27-41)   Hermitian A(r,c) = c + rI (c > r), A(r,r) = r, B(r,c) = c + (r mod 10)I; cblas_chemm (Left, Upper)
 * Usage: ref_micro <test> <n> <outfile>      prints: RESULT test=<..> n=<..> ns=<..> closed_form_err=<..|nan> tracked=<0|1>
 */
#define _GNU_SOURCE
#include <complex.h>
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

enum { ColMajor = 102, NoTrans = 111, Upper = 121, NonUnit = 131, Left = 141 };
extern void cblas_scopy(int, const float*, int, float*, int);
extern double cblas_dsdot(int, const float*, int, const float*, int);
extern void csrot_(const int*, void*, const int*, void*, const int*, const float*, const float*);
extern void cblas_sgbmv(int, int, int, int, int, int, float, const float*, int, const float*, int, float, float*, int);
extern void cblas_strmv(int, int, int, int, int, const float*, int, float*, int);
extern void cblas_dtrsm(int, int, int, int, int, int, int, double, const double*, int, double*, int);
extern void cblas_sgemm(int, int, int, int, int, int, float, const float*, int, const float*, int, float, float*, int);
extern void cblas_chemm(int, int, int, int, int, const void*, const void*, int, const void*, int, const void*, void*, int);

static double now_ns(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC_RAW, &ts);
    return ts.tv_sec * 1e9 + ts.tv_nsec;
}
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

int main(int argc, char** argv) {
    if (argc < 4) { fprintf(stderr, "usage: %s <copy|dsdot|rot|gbmv|trmv|trsm|gemm|hemm> <n> <outfile>\n", argv[0]); return 2; }
    const char* test = argv[1];
    const int n = atoi(argv[2]);
    int (*is_tracked)(const void*) = (int (*)(const void*))dlsym(RTLD_DEFAULT, "b200blas_is_tracked");
    const void* res = NULL; size_t res_bytes = 0;
    const void* biggest = NULL;
    double err = NAN, t0 = 0, t1 = 0, scalar = 0;
    if (!strcmp(test, "copy")) {
        float* x = calloc(n, sizeof *x); float* y = calloc(n, sizeof *y);
        for (int i = 0; i < n; i++) x[i] = i + 1;
        t0 = now_ns(); cblas_scopy(n, x, 1, y, 1); t1 = now_ns();
        err = 0; for (int i = 0; i < n; i++) err = fmax(err, fabs((double)y[i] - (i + 1)));
        res = y; res_bytes = n * sizeof *y; biggest = x;
    } else if (!strcmp(test, "dsdot")) {
        float* x = calloc(n, sizeof *x); float* y = calloc(n, sizeof *y);
        for (int i = 0; i < n; i++) x[i] = y[i] = i + 1;
        t0 = now_ns(); scalar = cblas_dsdot(n, x, 1, y, 1); t1 = now_ns();
        err = fabs(scalar - (double)n * (n + 1.0) * (2.0 * n + 1.0) / 6.0);
        res = &scalar; res_bytes = sizeof scalar; biggest = x;
    } else if (!strcmp(test, "rot")) {
        float complex* x = calloc(n, sizeof *x); float complex* y = calloc(n, sizeof *y);
        for (int i = 0; i < n; i++) x[i] = y[i] = i + (i + 1) * I;
        /* rot.c calls cblas_csrot, which OpenBLAS 0.3.15 (the CPU BLAS of this image) does not export: bound at run time, so the
         * plain run uses the Fortran symbol and the run under LD_PRELOAD the interposer's cblas_csrot */
        void (*cb)(int, void*, int, void*, int, float, float) = (void (*)(int, void*, int, void*, int, float, float))dlsym(RTLD_DEFAULT, "cblas_csrot");
        const int one = 1; const float c = 1.f, sn = 1.f;
        t0 = now_ns();
        if (cb) cb(n, x, 1, y, 1, c, sn); else csrot_(&n, x, &one, y, &one, &c, &sn);
        t1 = now_ns();
        err = 0; for (int i = 0; i < n; i++) err = fmax(err, fmax(cabs(x[i] - 2 * (i + (i + 1) * I)), cabs(y[i])));
        res = x; res_bytes = n * sizeof *x; biggest = x;
    } else if (!strcmp(test, "gbmv")) {
        const int m = n, kl = 2, ku = 2, lda = m;
        float* A = calloc(m, n * sizeof *A); float* x = calloc(n, sizeof *x); float* y = calloc(n, sizeof *y);
        for (int i = 0; i < n; i++) x[i] = y[i] = i;
        for (int row = 0; row < m; ++row)
            for (int col = imax(0, row - kl); col < imin(n, row + ku + 1); col++) A[(kl - row + col) + row * lda] = 1;
        t0 = now_ns(); cblas_sgbmv(ColMajor, NoTrans, m, n, kl, ku, 1.f, A, n, x, 1, 1.f, y, 1); t1 = now_ns();
        res = y; res_bytes = n * sizeof *y; biggest = A;
    } else if (!strcmp(test, "trmv")) {
        float* A = calloc(n, n * sizeof *A); float* x = calloc(n, sizeof *x);
        const long mod = imax(1, (int)(((long)n * n) / 10));
        for (int row = 0; row < n; ++row)
            for (int col = row; col < n; ++col) A[(size_t)col * n + row] = ((long)row * n + col) % mod;
        for (int i = 0; i < n; i++) x[i] = i;
        t0 = now_ns(); cblas_strmv(ColMajor, Upper, NoTrans, NonUnit, n, A, n, x, 1); t1 = now_ns();
        res = x; res_bytes = n * sizeof *x; biggest = A;
    } else if (!strcmp(test, "trsm")) {
        const int m = n, nrhs = imax(1, n / 2);
        double* A = calloc(m, m * sizeof *A); double* B = calloc(m, nrhs * sizeof *B);
        for (int row = 0; row < m; ++row)
            for (int col = row; col < m; ++col) A[(size_t)col * m + row] = ((long)row * m + col) % 10 + (row == col ? 10 : 0);
        for (int row = 0; row < m; ++row)
            for (int col = 0; col < nrhs; ++col) B[(size_t)col * m + row] = ((long)row * nrhs + col) % 10;
        t0 = now_ns(); cblas_dtrsm(ColMajor, Left, Upper, NoTrans, NonUnit, m, nrhs, 1.0, A, m, B, m); t1 = now_ns();
        res = B; res_bytes = (size_t)m * nrhs * sizeof *B; biggest = A;
    } else if (!strcmp(test, "gemm")) {
        const int m = n, k = n;
        float* A = calloc(m, k * sizeof *A); float* B = calloc(k, n * sizeof *B); float* C = calloc(m, n * sizeof *C);
        for (int row = 0; row < m; ++row)
            for (int col = 0; col < k; ++col) A[(size_t)col * m + row] = (row * k + col) / k;
        for (int row = 0; row < k; ++row)
            for (int col = 0; col < n; ++col) B[(size_t)col * k + row] = (row * n + col) % n;
        t0 = now_ns(); cblas_sgemm(ColMajor, NoTrans, NoTrans, m, n, k, 1.f, A, m, B, k, 0.f, C, m); t1 = now_ns();
        err = 0;
        for (int col = 0; col < n; ++col)
            for (int row = 0; row < m; ++row) err = fmax(err, fabs((double)C[(size_t)col * m + row] - (double)k * row * col));
        res = C; res_bytes = (size_t)m * n * sizeof *C; biggest = A;
    } else if (!strcmp(test, "hemm")) {
        const int m = n;
        float complex* A = calloc(m, m * sizeof *A); float complex* B = calloc(m, n * sizeof *B); float complex* C = calloc(m, n * sizeof *C);
        for (int row = 0; row < m; ++row)
            for (int col = row; col < m; ++col) {
                if (col == row) A[(size_t)col * m + row] = col;
                else { A[(size_t)col * m + row] = col + row * I; A[(size_t)row * m + col] = col - row * I; }
            }
        for (int row = 0; row < m; ++row)
            for (int col = 0; col < n; ++col) B[(size_t)col * m + row] = ((long)row * n + col) % n + (row % 10) * I;
        const float complex alpha = 1, beta = 0;
        t0 = now_ns(); cblas_chemm(ColMajor, Left, Upper, m, n, &alpha, A, m, B, m, &beta, C, m); t1 = now_ns();
        res = C; res_bytes = (size_t)m * n * sizeof *C; biggest = A;
    } else { fprintf(stderr, "unknown test %s\n", test); return 2; }
    FILE* f = fopen(argv[3], "wb");
    if (!f || fwrite(res, 1, res_bytes, f) != res_bytes) { perror("write"); return 1; }
    fclose(f);
    printf("RESULT test=%s n=%d ns=%.0f closed_form_err=%g tracked=%d\n", test, n, t1 - t0, err, (is_tracked && biggest) ? is_tracked(biggest) : 0);
    return 0;
}
