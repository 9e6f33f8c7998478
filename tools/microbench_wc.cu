// Duplex PCIe rate as a function of how the bytes of one end-to-end step are cut into copies: H2D in `up` pieces over
// `ups` streams while D2H runs in `dn` pieces on another stream (and: default vs write-combined pinned memory).
// nvcc -O2 -o tools/microbench_wc tools/microbench_wc.cu && tools/microbench_wc
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
static void run(void* d_in, const void* h_in, void* h_out, const void* d_out, size_t bytes, int up, int ups, int dn, const char* tag) {
  cudaStream_t s[4], b; for (auto& x : s) cudaStreamCreate(&x); cudaStreamCreate(&b);
  cudaEvent_t e0, e1, f1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&f1);
  const int reps = 6; const size_t ub = bytes / up, db = dn ? bytes / dn : 0;
  cudaDeviceSynchronize();
  cudaEventRecord(e0, b);
  for (int r = 0; r < reps; ++r) {
    for (int p = 0; p < (up > dn ? up : dn); ++p) {
      if (p < up) cudaMemcpyAsync((char*)d_in + p * ub, (const char*)h_in + p * ub, ub, cudaMemcpyHostToDevice, s[p % ups]);
      if (p < dn) cudaMemcpyAsync((char*)h_out + p * db, (const char*)d_out + p * db, db, cudaMemcpyDeviceToHost, b);
    }
  }
  cudaDeviceSynchronize();
  cudaEventRecord(e1, b); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("{\"%s\": {\"h2d_pieces\": %d, \"h2d_streams\": %d, \"d2h_pieces\": %d, \"ms_per_144MB_each_way\": %.3f, \"GBps_each_way\": %.1f}}\n",
         tag, up, ups, dn, ms / reps, bytes * reps / (ms * 1e-3) / 1e9);
  for (auto& x : s) cudaStreamDestroy(x); cudaStreamDestroy(b);
}
// the same bytes as `pieces` 2-D copies: `rows` rows each, the rows of one copy a third (half) of the buffer apart --
// how one chunk of three input arrays (two output arrays) that live in ONE pinned block travels as a single copy
static void run2d(void* d_in, const void* h_in, void* h_out, const void* d_out, size_t bytes, int pieces, const char* tag) {
  cudaStream_t a, b; cudaStreamCreate(&a); cudaStreamCreate(&b);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 6;
  const size_t up_pitch = bytes / 3, up_w = up_pitch / pieces, dn_pitch = bytes / 2, dn_w = dn_pitch / pieces;
  cudaDeviceSynchronize();
  cudaEventRecord(e0, b);
  for (int r = 0; r < reps; ++r)
    for (int p = 0; p < pieces; ++p) {
      cudaMemcpy2DAsync((char*)d_in + p * up_w, up_pitch, (const char*)h_in + p * up_w, up_pitch, up_w, 3, cudaMemcpyHostToDevice, a);
      cudaMemcpy2DAsync((char*)h_out + p * dn_w, dn_pitch, (const char*)d_out + p * dn_w, dn_pitch, dn_w, 2, cudaMemcpyDeviceToHost, b);
    }
  cudaDeviceSynchronize();
  cudaEventRecord(e1, b); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("{\"%s\": {\"copies_each_way\": %d, \"ms_per_144MB_each_way\": %.3f, \"GBps_each_way\": %.1f}}\n", tag, pieces, ms / reps,
         bytes * reps / (ms * 1e-3) / 1e9);
}
int main() {
  const size_t bytes = 144u << 20;
  void *h_def, *h_wc, *h_out, *d_in, *d_out;
  cudaHostAlloc(&h_def, bytes, cudaHostAllocDefault); cudaHostAlloc(&h_wc, bytes, cudaHostAllocWriteCombined);
  cudaHostAlloc(&h_out, bytes, cudaHostAllocDefault);
  memset(h_def, 1, bytes); memset(h_wc, 1, bytes);
  cudaMalloc(&d_in, bytes); cudaMalloc(&d_out, bytes); cudaMemset(d_out, 1, bytes);
  run(d_in, h_def, h_out, d_out, bytes, 1, 1, 1, "one copy each way");
  run(d_in, h_wc, h_out, d_out, bytes, 1, 1, 1, "one copy each way, write-combined source");
  run(d_in, h_def, h_out, d_out, bytes, 52, 1, 52, "52 + 52 copies (what a 13-chunk step issues)");
  run(d_in, h_def, h_out, d_out, bytes, 39, 1, 13, "39 up, 13 down");
  run(d_in, h_def, h_out, d_out, bytes, 39, 3, 13, "39 up on 3 streams, 13 down");
  run(d_in, h_def, h_out, d_out, bytes, 13, 1, 13, "13 up, 13 down");
  run(d_in, h_def, h_out, d_out, bytes, 8, 1, 8, "8 up, 8 down");
  run(d_in, h_def, h_out, d_out, bytes, 52, 1, 0, "52 up alone");
  run2d(d_in, h_def, h_out, d_out, bytes, 13, "13 two-dimensional copies each way (3 rows up, 2 rows down)");
  run2d(d_in, h_def, h_out, d_out, bytes, 9, "9 two-dimensional copies each way");
  return 0;
}
