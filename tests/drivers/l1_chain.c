/* l1_chain.c -- BASELINE config 3, Level-1 half: DDOT / DAXPY / DNRM2 / IDAMAX chained on calloc'd
 * (tracked -> managed under LD_PRELOAD=libb200blas.so) vectors that the CPU initialises, so the first call of
 * each vector pays the host->device migration and every later call must find it resident.
 * Prints per routine the average ns/call over the steady-state iterations (the first iteration is reported
 * separately), a checksum line RESULT for parity with the CPU BLAS run of the same binary, and -- when the
 * library is loaded -- where the driver last placed each vector.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

extern double ddot_(const int*, const double*, const int*, const double*, const int*);
extern double dnrm2_(const int*, const double*, const int*);
extern void daxpy_(const int*, const double*, const double*, const int*, double*, const int*);
extern int idamax_(const int*, const double*, const int*);

static double now_ns(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC_RAW, &ts); return ts.tv_sec * 1e9 + ts.tv_nsec; }

int main(int argc, char** argv) {
    int n = argc > 1 ? atoi(argv[1]) : (1 << 20), iters = argc > 2 ? atoi(argv[2]) : 10, one = 1;
    int namax = argc > 3 ? atoi(argv[3]) : 0;       /* optional: IDAMAX on its own vector of this length (BASELINE config 5: 2^28) */
    double *x = calloc(n, sizeof *x), *y = calloc(n, sizeof *y);
    double *z = namax > 0 ? calloc((size_t)namax, sizeof *z) : NULL;
    if (!x || !y || (namax > 0 && !z)) { perror("calloc"); return 1; }
    unsigned long long s = 7;
    for (int i = 0; i < n; i++) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        x[i] = (double)(s >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        y[i] = (double)(s >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0;
    }
    x[n / 3] = 3.0; x[n / 2] = -3.0;                 /* planted tie: idamax must return the first */
    if (z) {
        for (int i = 0; i < namax; i++) { s = s * 6364136223846793005ull + 1442695040888963407ull; z[i] = (double)(s >> 11) * (1.0 / 9007199254740992.0) * 2.0 - 1.0; }
        z[namax / 3] = 3.0; z[namax / 2] = -3.0; z[namax - 5] = 3.0;
    }
    const double* ax = z ? z : x; int an = z ? namax : n;
    double t[4] = {0, 0, 0, 0}, first[4] = {0, 0, 0, 0}, acc = 0, nrm = 0;
    int imax = 0;
    const double a = 1e-3, ma = -1e-3;
    for (int it = 0; it < iters; it++) {
        double t0 = now_ns(); double d = ddot_(&n, x, &one, y, &one);
        double t1 = now_ns(); daxpy_(&n, (it & 1) ? &ma : &a, x, &one, y, &one);
        double t2 = now_ns(); nrm = dnrm2_(&n, y, &one);
        double t3 = now_ns(); imax = idamax_(&an, ax, &one);
        double t4 = now_ns();
        acc += d * 1e-9;
        double dt[4] = {t1 - t0, t2 - t1, t3 - t2, t4 - t3};
        for (int k = 0; k < 4; k++) { if (it == 0) first[k] = dt[k]; else t[k] += dt[k]; }
    }
    int ss = iters > 1 ? iters - 1 : 1;
    printf("RESULT n=%d iters=%d acc=%.12e nrm=%.12e imax=%d\n", n, iters, acc, nrm, imax);
    printf("TIMES first_ns ddot=%.0f daxpy=%.0f dnrm2=%.0f idamax=%.0f steady_ns ddot=%.0f daxpy=%.0f dnrm2=%.0f idamax=%.0f\n",
           first[0], first[1], first[2], first[3], t[0] / ss, t[1] / ss, t[2] / ss, t[3] / ss);
    printf("GBS steady ddot=%.1f daxpy=%.1f dnrm2=%.1f idamax=%.1f\n", 16.0 * n / (t[0] / ss), 24.0 * n / (t[1] / ss), 8.0 * n / (t[2] / ss),
           8.0 * an / (t[3] / ss));
    int (*where)(const void*, size_t) = (int (*)(const void*, size_t))dlsym(RTLD_DEFAULT, "b200blas_residency");
    struct { unsigned long long v[9]; } st;
    void (*get)(void*) = (void (*)(void*))dlsym(RTLD_DEFAULT, "b200blas_get_stats");
    if (where && get) {
        get(&st);
        printf("RESIDENCY x=%d y=%d (device ordinal the range was last prefetched to; -1 = host)\n", where(x, (size_t)n * 8), where(y, (size_t)n * 8));
        printf("STATS hits=%llu misses=%llu calls=%llu h2d=%llu d2h=%llu prefetch=%llu managed_allocs=%llu\n", st.v[0], st.v[1], st.v[2], st.v[3], st.v[4], st.v[5], st.v[6]);
    } else printf("STATS none (CPU BLAS)\n");
    free(x); free(y); free(z);
    return 0;
}
