/* dgemm_big.c -- BASELINE config 2 as an unmodified C program: DGEMM m=n=k=N (default 16384) on calloc'd matrices through the
 * Fortran symbol.  Plain run: the CPU BLAS it is linked with.  Under LD_PRELOAD=libb200blas.so the calloc'd operands are tracked
 * managed blocks and dgemm_ runs on the GPU; with BLAS2CUDA_OPTIONS=devices=8 the same binary, unchanged, is partitioned over
 * the 8 GPUs of the box (the reference's multi-GPU comparator is a preload as well: tests/c/nvblas.conf:6-9).
 * Usage: dgemm_big [N] [rounds]
 * Prints: RESULT n=<N> rounds=<R> first_ms=<..> best_ms=<..> median_ms=<..> best_tflops=<..> max_rel_err=<..>   (error of 8 sampled entries
 * against long-double dot products of the same operands) */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

extern void dgemm_(const char*, const char*, const int*, const int*, const int*, const double*, const double*, const int*,
                   const double*, const int*, const double*, double*, const int*);
static double now_ms(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC_RAW, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
static int cmp(const void* a, const void* b) { double x = *(const double*)a, y = *(const double*)b; return x < y ? -1 : x > y; }

int main(int argc, char** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 16384, rounds = argc > 2 ? atoi(argv[2]) : 6;
    const size_t nn = (size_t)n * n;
    double *A = calloc(nn, sizeof *A), *B = calloc(nn, sizeof *B), *C = calloc(nn, sizeof *C);
    if (!A || !B || !C) { perror("calloc"); return 1; }
    unsigned long long s = 2;
    for (size_t i = 0; i < nn; i++) {
        s = s * 6364136223846793005ull + 1442695040888963407ull; A[i] = (double)(s >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
        s = s * 6364136223846793005ull + 1442695040888963407ull; B[i] = (double)(s >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
    }
    const double one = 1.0, zero = 0.0;
    double t[64];
    for (int r = 0; r < rounds && r < 64; r++) {
        double t0 = now_ms();
        dgemm_("N", "N", &n, &n, &n, &one, A, &n, B, &n, &zero, C, &n);
        t[r] = now_ms() - t0;
    }
    double maxrel = 0;
    for (int q = 0; q < 8; q++) {
        const size_t i = (size_t)((q * 2654435761u) % (unsigned)n), j = (size_t)((q * 40503u + 7u) % (unsigned)n);
        long double acc = 0, mag = 0;
        for (int k = 0; k < n; k++) { long double p = (long double)A[i + (size_t)k * n] * B[k + j * (size_t)n]; acc += p; mag += fabsl(p); }
        double rel = (double)(fabsl((long double)C[i + j * (size_t)n] - acc) / mag);
        if (rel > maxrel) maxrel = rel;
    }
    const double first = t[0];
    qsort(t + (rounds > 1), (size_t)(rounds > 1 ? rounds - 1 : 1), sizeof t[0], cmp);      /* steady state: everything after the first call */
    const double best = t[rounds > 1], med = t[(rounds > 1) + (rounds > 1 ? (rounds - 1) / 2 : 0)];
    printf("RESULT n=%d rounds=%d first_ms=%.3f best_ms=%.3f median_ms=%.3f best_tflops=%.2f max_rel_err=%.3g\n", n, rounds, first, best, med,
           2.0 * n * (double)n * n / best / 1e9, maxrel);
    free(A); free(B); free(C);
    return maxrel < 1e-13 ? 0 : 1;
}
