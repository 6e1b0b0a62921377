// Dependent-issue latencies of the fp64 instructions on the serial LM section (one warp, one CTA): cycles per op.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_latency tools/micro/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double x0, double a, double b) {
  double x = x0;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) x = fma(x, a, b);
  }
  long long t1 = clock64();
  double y = x0 + 1.0;
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) y = 1.0 / (y + 1.0);
  }
  long long t2 = clock64();
  float f = float(x0);
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f = fmaf(f, float(a), float(b));
  }
  long long t3 = clock64();
  double z = x0 + 2.0;
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) z = sqrt(z + 1.0);
  }
  long long t4 = clock64();
  // two independent DFMA chains interleaved (ILP 2)
  double p = x0, q = x0 + 0.5;
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 16; ++j) { p = fma(p, a, b); q = fma(q, b, a); }
  }
  long long t5 = clock64();
  if (threadIdx.x == 0) {
    out[0] = x + y + double(f) + z + p + q;
    cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4;
  }
}
int main() {
  double* d; long long* c;
  cudaMalloc(&d, 8); cudaMalloc(&c, 64);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(d, c, 0.5, 0.999, 0.001);
  long long h[5];
  cudaMemcpy(h, c, 40, cudaMemcpyDeviceToHost);
  printf("DFMA dependent: %.1f cycles/op | 1/(y+1) fp64: %.1f cycles/op | FFMA dependent: %.1f | sqrt(z+1) fp64: %.1f | 2 interleaved DFMA chains: %.1f cycles per pair\n",
         h[0] / 1024.0, h[1] / 256.0, h[2] / 1024.0, h[3] / 256.0, h[4] / 1024.0);
  return 0;
}
