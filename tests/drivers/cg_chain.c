/* cg_chain.c -- BASELINE config 3: a CG-style Level-1/2 chain on calloc'd (tracked -> managed)
 * buffers: q = A p (dgemv_), a = rho / ddot_(p,q), x += a p, r -= a q (daxpy_), rho' = ddot_(r,r),
 * p = r + (rho'/rho) p (dscal_ + daxpy_), ||r|| (dnrm2_), idamax_(r).  A is symmetric diagonally dominant
 * so the iteration converges; every BLAS call goes through the PLT, so the same binary runs on the CPU
 * BLAS (plain) and on libb200blas.so (LD_PRELOAD).
 * Prints: RESULT n=<n> iters=<k> rnorm=<..> xsum=<..> imax=<..> avg_iter_ns=<..> first_iter_ns=<..> steady_iter_ns=<..>
 * (the first iteration pays the one-off migration of the CPU-initialised managed matrix; steady = the rest)
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

extern void dgemv_(const char*, const int*, const int*, const double*, const double*, const int*, const double*, const int*,
                   const double*, double*, const int*);
extern double ddot_(const int*, const double*, const int*, const double*, const int*);
extern double dnrm2_(const int*, const double*, const int*);
extern void daxpy_(const int*, const double*, const double*, const int*, double*, const int*);
extern void dscal_(const int*, const double*, double*, const int*);
extern void dcopy_(const int*, const double*, const int*, double*, const int*);
extern int idamax_(const int*, const double*, const int*);

static unsigned long long sm_state;
static double u01(void) {   /* splitmix64 -> U(-1,1) */
    unsigned long long z = (sm_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
}
static double now_ns(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC_RAW, &ts); return ts.tv_sec * 1e9 + ts.tv_nsec; }

int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : 2048, iters = argc > 2 ? atoi(argv[2]) : 20, one = 1;
    double* A = calloc((size_t)n * n, sizeof *A);
    double *x = calloc(n, sizeof *x), *r = calloc(n, sizeof *r), *p = calloc(n, sizeof *p), *q = calloc(n, sizeof *q);
    if (!A || !x || !r || !p || !q) { perror("calloc"); return 1; }
    sm_state = 5;
    for (int j = 0; j < n; j++)
        for (int i = 0; i <= j; i++) { double v = u01(); A[(size_t)j * n + i] = v; A[(size_t)i * n + j] = v; }
    for (int i = 0; i < n; i++) A[(size_t)i * n + i] = n;          /* diagonally dominant => SPD */
    sm_state = 6;
    for (int i = 0; i < n; i++) r[i] = u01();                      /* b; x0 = 0 => r0 = b */
    dcopy_(&n, r, &one, p, &one);
    double rho = ddot_(&n, r, &one, r, &one), t0 = now_ns();
    const double d1 = 1.0, d0 = 0.0;
    double t_first = 0;
    for (int it = 0; it < iters; it++) {
        if (it == 1) t_first = now_ns();
        dgemv_("N", &n, &n, &d1, A, &n, p, &one, &d0, q, &one);
        double a = rho / ddot_(&n, p, &one, q, &one), ma = -a;
        daxpy_(&n, &a, p, &one, x, &one);
        daxpy_(&n, &ma, q, &one, r, &one);
        double rho2 = ddot_(&n, r, &one, r, &one), b = rho2 / rho;
        dscal_(&n, &b, p, &one);
        daxpy_(&n, &d1, r, &one, p, &one);
        rho = rho2;
    }
    double t1 = now_ns();
    double rn = dnrm2_(&n, r, &one);
    int imax = idamax_(&n, x, &one);
    double xs = 0;
    for (int i = 0; i < n; i++) xs += x[i];
    if (iters < 2) t_first = t1;
    printf("RESULT n=%d iters=%d rnorm=%.6e xsum=%.15g imax=%d avg_iter_ns=%.0f first_iter_ns=%.0f steady_iter_ns=%.0f\n", n, iters, rn, xs, imax,
           (t1 - t0) / iters, t_first - t0, iters > 1 ? (t1 - t_first) / (iters - 1) : 0.0);
    struct { unsigned long long v[9]; } st;
    void (*get)(void*) = (void (*)(void*))dlsym(RTLD_DEFAULT, "b200blas_get_stats");
    if (get) { get(&st); printf("STATS hits=%llu misses=%llu calls=%llu h2d=%llu d2h=%llu prefetch=%llu managed_allocs=%llu\n", st.v[0], st.v[1], st.v[2], st.v[3], st.v[4], st.v[5], st.v[6]); }
    else printf("STATS none (CPU BLAS)\n");
    free(A); free(x); free(r); free(p); free(q);
    return 0;
}
