// Peak probes for B200 (sm_100a): FP64 DFMA, DMMA shapes, FFMA, HBM read, and cuBLAS context numbers.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_peaks probe_peaks.cu -lcublas
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#define CK(x) do{cudaError_t ck_err_=(x); if(ck_err_!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(ck_err_),__LINE__); exit(1);} }while(0)

__global__ void k_dfma(double* out, int iters, double a, double b) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma(float* out, int iters, float a, float b) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma884(double* out, int iters, double a, double b) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = threadIdx.x; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1},{%2},{%3},{%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma1684(double* out, int iters, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5},{%6},{%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a), "d"(b), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma1688(double* out, int iters, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7},{%8,%9},{%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NACC>
__global__ void k_dmma16816(double* out, int iters, double a, double b) {
  double c[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = threadIdx.x; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < NACC; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3},{%4,%5,%6,%7,%8,%9,%10,%11},{%12,%13,%14,%15},{%0,%1,%2,%3};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                   : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a), "d"(b), "d"(a));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_read(const double2* __restrict__ x, size_t n2, double* out) {
  double s = 0;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n2; i += 4 * stride) {
    double2 a = x[i], b = x[i + stride], c = x[i + 2 * stride], d = x[i + 3 * stride];
    s += a.x + a.y + b.x + b.y + c.x + c.y + d.x + d.y;
  }
  for (; i < n2; i += stride) { double2 a = x[i]; s += a.x + a.y; }
  if (s == 1.2345e-300) out[0] = s;
}

template <class F> float timeit(F f, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  CK(cudaGetLastError());
  return best;
}

int main(int argc, char** argv) {
  int quick = argc > 1;
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sms=%d cc=%d.%d clock=%d kHz mem=%zu MiB l2=%d\n", p.name, p.multiProcessorCount, p.major, p.minor, p.clockRate, p.totalGlobalMem >> 20, p.l2CacheSize);
  int v; cudaDeviceGetAttribute(&v, cudaDevAttrConcurrentManagedAccess, 0); printf("concurrentManagedAccess=%d\n", v);
  cudaDeviceGetAttribute(&v, cudaDevAttrPageableMemoryAccess, 0); printf("pageableMemoryAccess=%d\n", v);
  cudaDeviceGetAttribute(&v, cudaDevAttrPageableMemoryAccessUsesHostPageTables, 0); printf("pageableUsesHostPT=%d\n", v);
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, 0); printf("smemOptin=%d\n", v);
  int nsm = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
  int iters = 20000;
  for (int wpb : {4, 8, 16, 32}) {
    int thr = wpb * 32, blocks = nsm;  // one block per SM
    float ms = timeit([&] { k_dfma<<<blocks, thr>>>(out, iters, 1.0000001, 1e-9); });
    printf("DFMA    warps/SM=%2d : %.2f TFLOP/s\n", wpb, 2.0 * 16 * iters * (double)thr * blocks / ms / 1e9);
  }
  for (int wpb : {8, 32}) {
    int thr = wpb * 32, blocks = nsm;
    float ms = timeit([&] { k_ffma<<<blocks, thr>>>((float*)out, iters, 1.0000001f, 1e-9f); });
    printf("FFMA    warps/SM=%2d : %.2f TFLOP/s\n", wpb, 2.0 * 16 * iters * (double)thr * blocks / ms / 1e9);
  }
  int it2 = 4000;
  for (int wpb : {4, 8, 16}) {
    int thr = wpb * 32, blocks = nsm;
    double w = (double)wpb * blocks * it2;
    float ms;
    ms = timeit([&] { k_dmma884<8><<<blocks, thr>>>(out, it2, 1.0000001, 1e-9); });
    printf("DMMA m8n8k4   x8  warps/SM=%2d : %.2f TFLOP/s\n", wpb, 2.0 * 8 * 8 * 4 * 8 * w / ms / 1e9);
    ms = timeit([&] { k_dmma1684<8><<<blocks, thr>>>(out, it2, 1.0000001, 1e-9); });
    printf("DMMA m16n8k4  x8  warps/SM=%2d : %.2f TFLOP/s\n", wpb, 2.0 * 16 * 8 * 4 * 8 * w / ms / 1e9);
    ms = timeit([&] { k_dmma1688<8><<<blocks, thr>>>(out, it2, 1.0000001, 1e-9); });
    printf("DMMA m16n8k8  x8  warps/SM=%2d : %.2f TFLOP/s\n", wpb, 2.0 * 16 * 8 * 8 * 8 * w / ms / 1e9);
    ms = timeit([&] { k_dmma16816<8><<<blocks, thr>>>(out, it2, 1.0000001, 1e-9); });
    printf("DMMA m16n8k16 x8  warps/SM=%2d : %.2f TFLOP/s\n", wpb, 2.0 * 16 * 8 * 16 * 8 * w / ms / 1e9);
    ms = timeit([&] { k_dmma16816<2><<<blocks, thr>>>(out, it2, 1.0000001, 1e-9); });
    printf("DMMA m16n8k16 x2  warps/SM=%2d : %.2f TFLOP/s\n", wpb, 2.0 * 16 * 8 * 16 * 2 * w / ms / 1e9);
    if (wpb <= 8) ms = timeit([&] { k_dmma1688<16><<<blocks, thr>>>(out, it2, 1.0000001, 1e-9); });
    if (wpb <= 8) printf("DMMA m16n8k8  x16 warps/SM=%2d : %.2f TFLOP/s\n", wpb, 2.0 * 16 * 8 * 8 * 16 * w / ms / 1e9);
  }
  // HBM read bandwidth
  {
    size_t n = (size_t)1 << 29;  // 4 GiB of doubles
    double* x; CK(cudaMalloc(&x, n * 8)); CK(cudaMemset(x, 0, n * 8));
    for (int bps : {4, 8, 16}) {
      float ms = timeit([&] { k_read<<<nsm * bps, 256>>>((const double2*)x, n / 2, out); });
      printf("HBM read 4GiB blocks/SM=%2d: %.1f GB/s\n", bps, n * 8.0 / ms / 1e6);
    }
    double* y; CK(cudaMalloc(&y, n * 8));
    float ms = timeit([&] { CK(cudaMemcpyAsync(y, x, n * 8, cudaMemcpyDeviceToDevice)); });
    printf("cudaMemcpy D2D 4GiB: %.1f GB/s (r+w)\n", 2.0 * n * 8.0 / ms / 1e6);
    cudaFree(x); cudaFree(y);
  }
  // cuBLAS context
  cublasHandle_t h; { cublasStatus_t st = cublasCreate(&h); printf("cublasCreate status=%d\n", (int)st); int ver; cublasGetVersion(h,&ver); printf("cublas version %d\n", ver);} 
#define CB(x) do{cublasStatus_t st_=(x); if(st_!=CUBLAS_STATUS_SUCCESS){printf("cublas error %d line %d\n",(int)st_,__LINE__);} }while(0)
  for (int n : {4096, 8192, 16384}) {
    if (quick && n > 8192) break;
    size_t e = (size_t)n * n;
    double *A, *B, *C; CK(cudaMalloc(&A, e * 8)); CK(cudaMalloc(&B, e * 8)); CK(cudaMalloc(&C, e * 8));
    CK(cudaMemset(A, 0, e * 8)); CK(cudaMemset(B, 0, e * 8)); CK(cudaMemset(C, 0, e * 8));
    // fill with nonzero pattern
    std::vector<double> hA(e); for (size_t i = 0; i < e; i++) hA[i] = ((i * 2654435761u) % 2001) / 1000.0 - 1.0;
    CK(cudaMemcpy(A, hA.data(), e * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(B, hA.data(), e * 8, cudaMemcpyHostToDevice));
    double al = 1, be = 0;
    float ms = timeit([&] { CB(cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n)); }, 3);
    printf("cuBLAS DGEMM NN n=%d: %.3f ms %.2f TFLOP/s\n", n, ms, 2.0 * n * (double)n * n / ms / 1e9);
    ms = timeit([&] { CB(cublasDgemm(h, CUBLAS_OP_T, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n)); }, 3);
    printf("cuBLAS DGEMM TN n=%d: %.3f ms %.2f TFLOP/s\n", n, ms, 2.0 * n * (double)n * n / ms / 1e9);
    float* fA = (float*)A; float* fB = (float*)B; float* fC = (float*)C; float fal = 1, fbe = 0;
    std::vector<float> hF(e); for (size_t i = 0; i < e; i++) hF[i] = (float)hA[i];
    CK(cudaMemcpy(fA, hF.data(), e * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(fB, hF.data(), e * 4, cudaMemcpyHostToDevice));
    ms = timeit([&] { CB(cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &fal, fA, n, fB, n, &fbe, fC, n)); }, 3);
    printf("cuBLAS SGEMM NN n=%d: %.3f ms %.2f TFLOP/s\n", n, ms, 2.0 * n * (double)n * n / ms / 1e9);
    if (n <= 8192) {
      cuDoubleComplex za = {0.7, -0.9}, zb = {1.3, -1.1};
      int nz = n / 2 * 2;  // reuse buffers: complex n/... keep n but buffers hold e doubles = e/2 complex -> use n' = n/sqrt2; simpler: alloc
      cuDoubleComplex *ZA, *ZB, *ZC; CK(cudaMalloc(&ZA, e * 16)); CK(cudaMalloc(&ZB, e * 16)); CK(cudaMalloc(&ZC, e * 16));
      CK(cudaMemcpy(ZA, hA.data(), e * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(((double*)ZA) + e, hA.data(), e * 8, cudaMemcpyHostToDevice));
      CK(cudaMemcpy(ZB, ZA, e * 16, cudaMemcpyDeviceToDevice)); CK(cudaMemset(ZC, 0, e * 16));
      ms = timeit([&] { CB(cublasZgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, nz, nz, nz, &za, ZA, nz, ZB, nz, &zb, ZC, nz)); }, 3);
      printf("cuBLAS ZGEMM NN n=%d: %.3f ms %.2f TFLOP/s (8mnk)\n", n, ms, 8.0 * n * (double)n * n / ms / 1e9);
      cudaFree(ZA); cudaFree(ZB); cudaFree(ZC);
    }
    cudaFree(A); cudaFree(B); cudaFree(C);
  }
  {
    int n = 32768; size_t e = (size_t)n * n;
    double *A, *x, *y; CK(cudaMalloc(&A, e * 8)); CK(cudaMalloc(&x, n * 8)); CK(cudaMalloc(&y, n * 8));
    CK(cudaMemset(A, 0, e * 8)); CK(cudaMemset(x, 0, n * 8)); CK(cudaMemset(y, 0, n * 8));
    double al = 1, be = 0;
    float ms = timeit([&] { CB(cublasDgemv(h, CUBLAS_OP_N, n, n, &al, A, n, x, 1, &be, y, 1)); });
    printf("cuBLAS DGEMV N n=%d: %.3f ms %.1f GB/s\n", n, ms, 8.0 * (e + 3.0 * n) / ms / 1e6);
    ms = timeit([&] { CB(cublasDgemv(h, CUBLAS_OP_T, n, n, &al, A, n, x, 1, &be, y, 1)); });
    printf("cuBLAS DGEMV T n=%d: %.3f ms %.1f GB/s\n", n, ms, 8.0 * (e + 3.0 * n) / ms / 1e6);
    size_t nv = (size_t)1 << 26; double r; int idx;
    double* X = A; double* Y = A + nv;
    ms = timeit([&] { CB(cublasDdot(h, (int)nv, X, 1, Y, 1, &r)); });
    printf("cuBLAS DDOT n=2^26: %.3f ms %.1f GB/s\n", ms, 16.0 * nv / ms / 1e6);
    ms = timeit([&] { CB(cublasDnrm2(h, (int)nv, X, 1, &r)); });
    printf("cuBLAS DNRM2 n=2^26: %.3f ms %.1f GB/s\n", ms, 8.0 * nv / ms / 1e6);
    ms = timeit([&] { CB(cublasDaxpy(h, (int)nv, &al, X, 1, Y, 1)); });
    printf("cuBLAS DAXPY n=2^26: %.3f ms %.1f GB/s\n", ms, 24.0 * nv / ms / 1e6);
    size_t nb = (size_t)1 << 28;
    ms = timeit([&] { CB(cublasIdamax(h, (int)nb, X, 1, &idx)); });
    printf("cuBLAS IDAMAX n=2^28: %.3f ms %.1f GB/s\n", ms, 8.0 * nb / ms / 1e6);
  }
  // sustained DFMA for ~3 s to see clocks
  {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    int launches = 60;
    for (int i = 0; i < launches; i++) k_dmma16816<8><<<nsm, 256>>>(out, 40000, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("sustained DMMA m16n8k16 8 warps/SM %.0f ms: %.2f TFLOP/s\n", ms, 2.0 * 16 * 8 * 16 * 8 * 8.0 * nsm * 40000 * launches / ms / 1e9);
    cudaEventRecord(e0);
    for (int i = 0; i < launches; i++) k_dfma<<<nsm, 1024>>>(out, 100000, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    printf("sustained DFMA 32 warps/SM %.0f ms: %.2f TFLOP/s\n", ms, 2.0 * 16 * 100000 * 1024.0 * nsm * launches / ms / 1e9);
  }
  return 0;
}
