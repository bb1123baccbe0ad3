/* gemm_fixture.c -- the reference's micro-benchmark tests/c/gemm.c + tests/c/test.c restated for
 * BASELINE config 1 (DGEMM m=n=k=N through the interposer), in f64 through BOTH symbol families:
 *   - calloc'd column-major matrices (so the allocation tracker sees them), A[i + j*m] = i,
 *     B[i + j*k] = j  (gemm.c:29-35), alpha = 1, beta = 0  =>  C[i,j] = k*i*j exactly;
 *   - 11 rounds, the first discarded (test.c:60-65), CLOCK_MONOTONIC_RAW around the call only
 *     (test.c:29-37).
 * Run plain  -> symbols bind to the CPU BLAS it is linked with (reference result + CPU timing);
 * run under LD_PRELOAD=libb200blas.so -> same symbols bind to the GPU library.
 * Prints one machine-readable line:  RESULT api=<f77|cblas> n=<N> avg_ns=<..> max_abs_err=<..> checksum=<..>
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

extern void dgemm_(const char*, const char*, const int*, const int*, const int*, const double*, const double*, const int*,
                   const double*, const int*, const double*, double*, const int*);
extern void cblas_dgemm(int, int, int, int, int, int, double, const double*, int, const double*, int, double, double*, int);

static double now_ns(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC_RAW, &ts);
    return ts.tv_sec * 1e9 + ts.tv_nsec;
}

int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 1024;
    int rounds = argc > 2 ? atoi(argv[2]) : 11;
    for (int api = 0; api < 2; api++) {
        double total = 0, maxerr = 0, checksum = 0;
        for (int r = 0; r < rounds; r++) {
            int m = n, k = n;
            double* A = calloc((size_t)m, (size_t)k * sizeof *A);
            double* B = calloc((size_t)k, (size_t)n * sizeof *B);
            double* C = calloc((size_t)m, (size_t)n * sizeof *C);
            if (!A || !B || !C) { perror("calloc"); return 1; }
            for (int row = 0; row < m; ++row)
                for (int col = 0; col < k; ++col) A[(size_t)col * m + row] = (double)((row * (long)k + col) / k);
            for (int row = 0; row < k; ++row)
                for (int col = 0; col < n; ++col) B[(size_t)col * k + row] = (double)((row * (long)n + col) % n);
            const double alpha = 1, beta = 0;
            double t0 = now_ns();
            if (api == 0) dgemm_("N", "N", &m, &n, &k, &alpha, A, &m, B, &k, &beta, C, &m);
            else cblas_dgemm(102, 111, 111, m, n, k, alpha, A, m, B, k, beta, C, m);
            double t1 = now_ns();
            if (r > 0 || rounds == 1) total += t1 - t0;
            if (r == rounds - 1) {
                for (int j = 0; j < n; j++)
                    for (int i = 0; i < m; i++) {
                        double want = (double)k * i * j, e = fabs(C[(size_t)j * m + i] - want);
                        if (e > maxerr) maxerr = e;
                        checksum += C[(size_t)j * m + i];
                    }
            }
            free(A); free(B); free(C);
        }
        printf("RESULT api=%s n=%d avg_ns=%.0f max_abs_err=%g checksum=%.17g\n", api ? "cblas" : "f77", n,
               total / (rounds > 1 ? rounds - 1 : 1), maxerr, checksum);
    }
    /* when the interposer is loaded, report its hit/miss counters (reference statistics.csv) */
    struct { unsigned long long v[9]; } st;
    void (*get)(void*) = (void (*)(void*))dlsym(RTLD_DEFAULT, "b200blas_get_stats");
    if (get) {
        get(&st);
        printf("STATS hits=%llu misses=%llu calls=%llu h2d=%llu d2h=%llu managed_allocs=%llu managed_frees=%llu\n", st.v[0], st.v[1], st.v[2],
               st.v[3], st.v[4], st.v[6], st.v[7]);
    } else {
        printf("STATS none (CPU BLAS)\n");
    }
    return 0;
}
