// Micro-benchmark (experiment, not product): 4x4 f32 footprint gathers of edge-like point lists, (a) 16 LDG from a pitch-linear
// image as the solve kernel does, (b) 4 tld4 (tex2Dgather) from a CUDA array.  One 512-thread CTA per image, R sweeps each.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tex_gather tex_gather.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
constexpr int W = 640, H = 480, PAD = 4, PITCH = 648;
__global__ void __launch_bounds__(512, 1) k_ldg(const float* __restrict__ img, const uint2* __restrict__ pts, const int* __restrict__ npts, int cap, int reps, float* sink) {
  const float* base = img + size_t(blockIdx.x) * PITCH * (H + 2 * PAD) + (PAD - 1) * PITCH + (PAD - 1);
  const uint2* p = pts + size_t(blockIdx.x) * cap;
  const int n = npts[blockIdx.x];
  float s = 0.f;
  for (int r = 0; r < reps; ++r)
    for (int j = threadIdx.x; j < n; j += 512) {
      const uint2 q = __ldg(p + j);
      const int iu = min(max(int(q.x & 0xffff) + (r & 3) - 1, -3), W + 1), iv = min(max(int(q.x >> 16) + ((r >> 2) & 1), -3), H + 1);
      const int off = iv * PITCH + iu;
      const float* p0 = base + off; const float* p1 = p0 + PITCH; const float* p2 = p1 + PITCH; const float* p3 = p2 + PITCH;
      s += __ldg(p0) + __ldg(p0 + 1) + __ldg(p0 + 2) + __ldg(p0 + 3) + __ldg(p1) + __ldg(p1 + 1) + __ldg(p1 + 2) + __ldg(p1 + 3) +
           __ldg(p2) + __ldg(p2 + 1) + __ldg(p2 + 2) + __ldg(p2 + 3) + __ldg(p3) + __ldg(p3 + 1) + __ldg(p3 + 2) + __ldg(p3 + 3);
    }
  sink[blockIdx.x * 512 + threadIdx.x] = s;
}
__global__ void __launch_bounds__(512, 1) k_tex(const cudaTextureObject_t* __restrict__ tex, const uint2* __restrict__ pts, const int* __restrict__ npts, int cap, int reps, float* sink) {
  const cudaTextureObject_t t = tex[blockIdx.x];
  const uint2* p = pts + size_t(blockIdx.x) * cap;
  const int n = npts[blockIdx.x];
  float s = 0.f;
  for (int r = 0; r < reps; ++r)
    for (int j = threadIdx.x; j < n; j += 512) {
      const uint2 q = __ldg(p + j);
      const int iu = int(q.x & 0xffff) + (r & 3) - 1, iv = int(q.x >> 16) + ((r >> 2) & 1);
      // quad (cols iu-1, iu; rows iv-1, iv) has its bilinear centre at (iu, iv) in unnormalised texel-centre coordinates
      const float x0 = float(iu), y0 = float(iv);
      const float4 a = tex2Dgather<float4>(t, x0, y0, 0), b = tex2Dgather<float4>(t, x0 + 2.f, y0, 0);
      const float4 c = tex2Dgather<float4>(t, x0, y0 + 2.f, 0), d = tex2Dgather<float4>(t, x0 + 2.f, y0 + 2.f, 0);
      s += a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w + d.x + d.y + d.z + d.w;
    }
  sink[blockIdx.x * 512 + threadIdx.x] = s;
}
int main(int argc, char** argv) {
  const int N = argc > 1 ? atoi(argv[1]) : 296, reps = argc > 2 ? atoi(argv[2]) : 8;
  const int cap = 40000;
  std::vector<float> himg(size_t(PITCH) * (H + 2 * PAD));
  for (size_t i = 0; i < himg.size(); ++i) himg[i] = float(i % 1000) * 0.001f;
  float* dimg; CK(cudaMalloc(&dimg, himg.size() * 4 * N));
  for (int i = 0; i < N; ++i) CK(cudaMemcpy(dimg + himg.size() * i, himg.data(), himg.size() * 4, cudaMemcpyHostToDevice));
  // edge-like points: horizontal runs (mean length 3) at ~9 % density, row-major order
  std::vector<uint2> hp(size_t(cap) * N); std::vector<int> hn(N);
  srand(1);
  for (int i = 0; i < N; ++i) {
    int n = 0;
    for (int y = 2; y < H - 2 && n < cap - 8; ++y)
      for (int x = 2; x < W - 2 && n < cap - 8;) {
        if (rand() % 100 < 3) { int len = 1 + rand() % 5; for (int k = 0; k < len && x < W - 2; ++k, ++x) hp[size_t(i) * cap + n++] = make_uint2(unsigned(x) | (unsigned(y) << 16), 0); }
        else ++x;
      }
    hn[i] = n;
  }
  uint2* dp; int* dn; float* sink;
  CK(cudaMalloc(&dp, hp.size() * 8)); CK(cudaMemcpy(dp, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&dn, N * 4)); CK(cudaMemcpy(dn, hn.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMalloc(&sink, size_t(N) * 512 * 4));
  std::vector<cudaTextureObject_t> ht(N);
  cudaChannelFormatDesc cd = cudaCreateChannelDesc<float>();
  for (int i = 0; i < N; ++i) {
    cudaArray_t arr; CK(cudaMallocArray(&arr, &cd, W, H, cudaArrayTextureGather));
    CK(cudaMemcpy2DToArray(arr, 0, 0, himg.data() + PAD * PITCH + PAD, PITCH * 4, W * 4, H, cudaMemcpyHostToDevice));
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeArray; rd.res.array.array = arr;
    cudaTextureDesc td = {}; td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint; td.readMode = cudaReadModeElementType; td.normalizedCoords = 0;
    CK(cudaCreateTextureObject(&ht[i], &rd, &td, nullptr));
  }
  cudaTextureObject_t* dt; CK(cudaMalloc(&dt, N * sizeof(cudaTextureObject_t))); CK(cudaMemcpy(dt, ht.data(), N * sizeof(cudaTextureObject_t), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double tot = 0; for (int i = 0; i < N; ++i) tot += hn[i];
  for (int mode = 0; mode < 2; ++mode) {
    for (int it = 0; it < 3; ++it) {
      CK(cudaEventRecord(e0));
      if (mode == 0) k_ldg<<<N, 512>>>(dimg, dp, dn, cap, reps, sink); else k_tex<<<N, 512>>>(dt, dp, dn, cap, reps, sink);
      CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("%s: %d images x %d sweeps, %.0f points/image: %.3f ms, %.2f G point-gathers/s\n", mode ? "tld4 x4" : "ldg x16", N, reps, tot / N, ms, tot * reps / ms * 1e-6);
    }
  }
  // correctness spot check of the gather addressing: texel values are index-derived
  return 0;
}
