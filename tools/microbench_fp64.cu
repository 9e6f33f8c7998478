// FP64 ceilings on B200 (sm_100a): DFMA, DMMA (m8n8k4 / m16n8k8), exp(double),
// sqrt/div latency, pinned PCIe copies.  Output: one JSON object on stdout.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench_fp64 microbench_fp64.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double *c, const double *a, const double *b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int ILP>
__global__ void k_dfma(double *out, int iters, double a, double b) {
  double acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += acc[i];
  if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void k_dmma884(double *out, int iters, double a, double b) {
  double c0[ILP], c1[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x; c1[i] = i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma884(c0[i], c1[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
  if (s == 123.456) out[0] = s;
}

template <int ILP>
__global__ void k_dmma1688(double *out, int iters, double a, double b) {
  double c[ILP][4];
  double av[4] = {a, a, b, b}, bv[2] = {b, a};
#pragma unroll
  for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) dmma1688(c[i], av, bv);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  if (s == 123.456) out[0] = s;
}

__global__ void k_exp(double *out, int iters, double x0) {
  double x = x0 - 1e-3 * threadIdx.x, s = 0;
  for (int it = 0; it < iters; it++) {
    s += exp(x); x -= 1e-4;
    s += exp(x * 0.5); s += exp(x * 0.25); s += exp(x * 0.125);
  }
  if (s == 123.456) out[0] = s;
}

// dependent chain latencies measured with clock64 by a single warp
__global__ void k_lat(long long *out, double seed) {
  double v = seed;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; i++) v = sqrt(v + 1.0);
  long long t1 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; i++) v = 1.0 / (v + 1.0);
  long long t2 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; i++) v = fma(v, 1.0000001, 1e-9);
  long long t3 = clock64();
  double c0 = v, c1 = v;
#pragma unroll 1
  for (int i = 0; i < 256; i++) dmma884(c0, c1, 1e-3, 1e-3);
  long long t4 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; i++) v = rsqrt(v + 1.0);
  long long t5 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; i++) v = __shfl_sync(0xffffffffu, v, (threadIdx.x + 1) & 31);
  long long t6 = clock64();
  if (threadIdx.x == 0) {
    out[0] = (t1 - t0); out[1] = (t2 - t1); out[2] = (t3 - t2); out[3] = (t4 - t3);
    out[4] = (t5 - t4); out[5] = (t6 - t5);
    out[6] = (long long)(v + c0 + c1);
  }
}

template <typename F>
static float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  f(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int sms = p.multiProcessorCount;
  double *d; CK(cudaMalloc(&d, 1 << 20));
  long long *dl; CK(cudaMalloc(&dl, 64 * 8));
  const int iters = 4096;
  printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);

  // DFMA: blocks x threads sweep
  for (int wpb : {4, 8, 16}) {
    int grid = sms * (wpb == 16 ? 2 : 4);
    float ms = time_ms([&] { k_dfma<8><<<grid, wpb * 32>>>(d, iters, 1.0000001, 1e-9); });
    double fl = 2.0 * 8 * iters * (double)grid * wpb * 32;
    printf(", \"dfma_tflops_w%d\": %.3f", wpb, fl / ms * 1e-9);
  }
  for (int wpb : {4, 8, 16}) {
    int grid = sms * (wpb == 16 ? 2 : 4);
    float ms = time_ms([&] { k_dmma884<8><<<grid, wpb * 32>>>(d, iters, 1e-3, 1e-3); });
    double fl = 2.0 * 256 * 8 * iters * (double)grid * wpb;
    printf(", \"dmma884_tflops_w%d\": %.3f", wpb, fl / ms * 1e-9);
  }
  {
    int wpb = 4, grid = sms;  // 1 warp per SMSP, ILP 8: issue-limited view
    float ms = time_ms([&] { k_dmma884<8><<<grid, wpb * 32>>>(d, iters, 1e-3, 1e-3); });
    double fl = 2.0 * 256 * 8 * iters * (double)grid * wpb;
    printf(", \"dmma884_tflops_1warp_per_smsp_ilp8\": %.3f", fl / ms * 1e-9);
    ms = time_ms([&] { k_dmma884<2><<<grid, wpb * 32>>>(d, iters, 1e-3, 1e-3); });
    fl = 2.0 * 256 * 2 * iters * (double)grid * wpb;
    printf(", \"dmma884_tflops_1warp_per_smsp_ilp2\": %.3f", fl / ms * 1e-9);
    ms = time_ms([&] { k_dfma<8><<<grid, wpb * 32>>>(d, iters, 1.0000001, 1e-9); });
    fl = 2.0 * 8 * iters * (double)grid * wpb * 32;
    printf(", \"dfma_tflops_1warp_per_smsp_ilp8\": %.3f", fl / ms * 1e-9);
  }
  for (int wpb : {4, 8, 16}) {
    int grid = sms * (wpb == 16 ? 2 : 4);
    float ms = time_ms([&] { k_dmma1688<4><<<grid, wpb * 32>>>(d, iters, 1e-3, 1e-3); });
    double fl = 2.0 * 16 * 8 * 8 * 4 * iters * (double)grid * wpb;
    printf(", \"dmma1688_tflops_w%d\": %.3f", wpb, fl / ms * 1e-9);
  }
  {
    int wpb = 8, grid = sms * 4;
    float ms = time_ms([&] { k_exp<<<grid, wpb * 32>>>(d, 1024, -0.5); });
    double n = 4.0 * 1024 * (double)grid * wpb * 32;
    printf(", \"exp_gexp_per_s\": %.3f", n / ms * 1e-6);
  }
  {
    k_lat<<<1, 32>>>(dl, 2.0); CK(cudaDeviceSynchronize());
    long long h[8]; CK(cudaMemcpy(h, dl, 56, cudaMemcpyDeviceToHost));
    printf(", \"lat_cycles\": {\"sqrt_add\": %.1f, \"div_add\": %.1f, \"dfma\": %.1f, \"dmma884\": %.1f, \"rsqrt_add\": %.1f, \"shfl64\": %.1f}",
           h[0] / 256.0, h[1] / 256.0, h[2] / 256.0, h[3] / 256.0, h[4] / 256.0, h[5] / 256.0);
  }
  {
    size_t nb = 256u << 20;
    void *hp, *dp; CK(cudaMallocHost(&hp, nb)); CK(cudaMalloc(&dp, nb));
    float ms = time_ms([&] { CK(cudaMemcpyAsync(dp, hp, nb, cudaMemcpyHostToDevice)); }, 3);
    printf(", \"h2d_pinned_gbs\": %.2f", nb / ms * 1e-6);
    ms = time_ms([&] { CK(cudaMemcpyAsync(hp, dp, nb, cudaMemcpyDeviceToHost)); }, 3);
    printf(", \"d2h_pinned_gbs\": %.2f", nb / ms * 1e-6);
    cudaFreeHost(hp); cudaFree(dp);
  }
  int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
  printf(", \"clock_khz_attr\": %d}\n", clk);
  return 0;
}
