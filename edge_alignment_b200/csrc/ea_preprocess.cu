// Frame preprocessing kernels (sm_100a), batched over n frames.
//
// Replaces the OpenCV call sequences of
//   get_aX                  standalone/utils.cpp:201-281  (GaussianBlur, cvtColor, Laplacian, convertScaleAbs,
//                                                          threshold, Z>0 test, row-major ordered compaction)
//   get_distance_transform  standalone/utils.cpp:38-83    (same gradient, medianBlur 3, distanceTransform(L2,3),
//                                                          normalize MINMAX)
// Integer stages are bit-exact restatements of OpenCV's portable code paths (pinned in tests/golden).
// HBM-bound byte/integer work: coalesced loads, shared-memory staging, no tensor cores.
#include "ea_internal.h"

namespace {

__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * n - 2 - i;
  return i;
}

// ---- pyramid: cv::resize(0.5) INTER_LINEAR on BGR (== 2x2 mean, +2 >> 2), INTER_NEAREST on depth ---------
__global__ void __launch_bounds__(256) k_pyr_down(const uint8_t* __restrict__ src_bgr, const uint16_t* __restrict__ src_depth,
                                                  size_t src_frame_stride_px, const int32_t* __restrict__ src_slots,
                                                  uint8_t* __restrict__ dst_bgr, uint16_t* __restrict__ dst_depth,
                                                  const int32_t* __restrict__ dst_slots, int sw, int sh) {
  const int dw = sw >> 1, dh = sh >> 1;
  const int f = blockIdx.y;
  const size_t sbase = (src_slots ? size_t(src_slots[f]) : size_t(f)) * src_frame_stride_px;
  const size_t dbase = size_t(dst_slots[f]) * size_t(dw) * dh;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < dw * dh; i += gridDim.x * blockDim.x) {
    const int x = i % dw, y = i / dw;
    const uint8_t* p = src_bgr + (sbase + size_t(2 * y) * sw + 2 * x) * 3;
    const uint8_t* q = p + size_t(sw) * 3;
    uint8_t* d = dst_bgr + (dbase + i) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) d[c] = uint8_t((p[c] + p[3 + c] + q[c] + q[3 + c] + 2) >> 2);
    if (src_depth) dst_depth[dbase + i] = src_depth[sbase + size_t(2 * y) * sw + 2 * x];
  }
}

// ---- fused GaussianBlur3 -> RGB2GRAY -> Laplacian3 -> |.| saturate -> threshold -> bit mask ----------------
// Block = 32x8 pixels; one ballot word per warp row.
#define ET_W 32
#define ET_H 8
__global__ void __launch_bounds__(256) k_edge_mask(const uint8_t* __restrict__ bgr, const uint16_t* __restrict__ depth,
                                                   size_t frame_stride_px, const int32_t* __restrict__ src_slots,
                                                   const int32_t* __restrict__ dst_slots, uint32_t* __restrict__ edge_bits,
                                                   uint32_t* __restrict__ ref_bits, int w, int h, int words, int thresh) {
  __shared__ uint8_t tile[ET_H + 4][(ET_W + 4) * 3 + 4];
  __shared__ uint8_t gray[ET_H + 2][ET_W + 2 + 2];
  const int f = blockIdx.z;
  const size_t sbase = (src_slots ? size_t(src_slots[f]) : size_t(f)) * frame_stride_px;
  const int x0 = blockIdx.x * ET_W, y0 = blockIdx.y * ET_H;
  const int tid = threadIdx.y * ET_W + threadIdx.x;
  // stage the BGR tile with a 2-pixel halo; outside the image map through REFLECT_101 (the composition of the
  // blur's and the Laplacian's border handling reduces to this because both kernels are symmetric)
  for (int i = tid; i < (ET_H + 4) * (ET_W + 4) * 3; i += 256) {
    const int row = i / ((ET_W + 4) * 3), b = i % ((ET_W + 4) * 3);
    const int px = b / 3, c = b % 3;
    const int gx = reflect101(x0 - 2 + px, w), gy = reflect101(y0 - 2 + row, h);
    tile[row][b] = bgr[(sbase + size_t(gy) * w + gx) * 3 + c];
  }
  __syncthreads();
  for (int i = tid; i < (ET_H + 2) * (ET_W + 2); i += 256) {
    const int gyr = i / (ET_W + 2), gxr = i % (ET_W + 2);  // gray position == tile position (gxr+1, gyr+1)
    int ch[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint8_t* r0 = &tile[gyr][gxr * 3 + c];
      const uint8_t* r1 = &tile[gyr + 1][gxr * 3 + c];
      const uint8_t* r2 = &tile[gyr + 2][gxr * 3 + c];
      const int s = (r0[0] + 2 * r0[3] + r0[6]) + 2 * (r1[0] + 2 * r1[3] + r1[6]) + (r2[0] + 2 * r2[3] + r2[6]);
      ch[c] = (s + 8) >> 4;                       // GaussianBlur 3x3 sigma 0: [1 2 1]/4 separable, one rounding
    }
    gray[gyr][gxr] = uint8_t((ch[0] * 9798 + ch[1] * 19235 + ch[2] * 3735 + 16384) >> 15);  // RGB2GRAY on BGR data
  }
  __syncthreads();
  const int lx = threadIdx.x, ly = threadIdx.y, x = x0 + lx, y = y0 + ly;
  bool edge = false, valid = false;
  if (x < w && y < h) {
    const int gx = lx + 1, gy = ly + 1;
    int v = 2 * (gray[gy - 1][gx - 1] + gray[gy - 1][gx + 1] + gray[gy + 1][gx - 1] + gray[gy + 1][gx + 1]) - 8 * gray[gy][gx];
    v = v < 0 ? -v : v;                             // Laplacian ksize 3 + convertScaleAbs
    edge = min(v, 255) > thresh;
    if (depth) valid = depth[sbase + size_t(y) * w + x] > 0;
  }
  const unsigned eb = __ballot_sync(0xffffffffu, edge);
  const unsigned rb = __ballot_sync(0xffffffffu, edge && valid);
  if (lx == 0 && y < h) {
    const size_t o = (size_t(dst_slots[f]) * h + y) * words + blockIdx.x;
    edge_bits[o] = eb;
    if (ref_bits) ref_bits[o] = rb;
  }
}

// ---- ordered compaction: one CTA per frame, row-major rank == reference's loop order (utils.cpp:268-280) ----
#define CP_THREADS 1024
__global__ void __launch_bounds__(CP_THREADS) k_compact(const uint32_t* __restrict__ ref_bits, const uint16_t* __restrict__ depth,
                                                        size_t frame_stride_px, const int32_t* __restrict__ src_slots,
                                                        const int32_t* __restrict__ dst_slots, float4* __restrict__ pts,
                                                        int* __restrict__ n_pts, int level, int* __restrict__ overflow,
                                                        int w, int h, int words, int cap) {
  __shared__ int warp_sums[CP_THREADS / 32];
  __shared__ int total_s;
  const int f = blockIdx.x, slot = dst_slots[f];
  const uint32_t* bits = ref_bits + size_t(slot) * h * words;
  const uint16_t* dep = depth + (src_slots ? size_t(src_slots[f]) : size_t(f)) * frame_stride_px;
  float4* out = pts + size_t(slot) * cap;
  const int nw = h * words;
  const int per = (nw + CP_THREADS - 1) / CP_THREADS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int b = tid * per, e = min(b + per, nw);
  int cnt = 0;
  for (int i = b; i < e; ++i) cnt += __popc(bits[i]);
  int incl = cnt;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = warp_sums[lane];
    int s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += t; }
    warp_sums[lane] = s - v;
    if (lane == 31) total_s = s;
  }
  __syncthreads();
  int rank = warp_sums[warp] + incl - cnt;
  for (int i = b; i < e; ++i) {
    unsigned m = bits[i];
    const int y = i / words, xb = (i % words) * 32;
    while (m) {
      const int bit = __ffs(m) - 1;
      m &= m - 1;
      const int x = xb + bit;
      if (rank < cap) out[rank] = make_float4(float(x), float(y), float(dep[size_t(y) * w + x]), 1.0f);
      ++rank;
    }
  }
  if (tid == 0) {
    n_pts[slot * EA_MAX_LEVELS + level] = min(total_s, cap);
    if (total_s > cap) atomicExch(overflow, 1);
  }
}

// ---- 3x3 median on the binary mask, bit-parallel -----------------------------------------------------------
// a, b, c: bit fields of rows y-1, y, y+1 where bit k+1 <-> pixel x0+k (bit 0 is x0-1).  Returns the field of
// "at least 5 of the 9 neighbours are edge" aligned so that bit k <-> pixel x0+k.
__device__ __forceinline__ unsigned median3_bits(unsigned a, unsigned b, unsigned c) {
  const unsigned s0 = a ^ b ^ c, s1 = (a & b) | (c & (a ^ b));  // per-column vertical count = s0 + 2 s1
  const unsigned A0 = s0, B0 = s0 >> 1, C0 = s0 >> 2, A1 = s1, B1 = s1 >> 1, C1 = s1 >> 2;
  const unsigned lo = A0 ^ B0 ^ C0, carry = (A0 & B0) | (C0 & (A0 ^ B0));
  // m = A1 + B1 + C1 + carry ; total = lo + 2 m >= 5  <=>  m >= 3 or (m == 2 and lo)
  const unsigned t = A1 ^ B1, u = A1 & B1, v = C1 ^ carry, wv = C1 & carry;
  const unsigned m0 = t ^ v, tv = t & v;
  const unsigned ge1 = u | wv | tv, eq2 = u & wv;          // number of "twos": >=1, ==2
  const unsigned m_ge3 = eq2 | (ge1 & m0);
  const unsigned m_eq2 = ge1 & ~eq2 & ~m0;
  return m_ge3 | (m_eq2 & lo);
}

// bits [x_start-1, x_start+P] of a mask row as a field (bit 0 <-> x_start-1), with BORDER_REPLICATE in x
__device__ __forceinline__ unsigned row_field(const uint32_t* __restrict__ row, int words, int w, int x_start, int nbits) {
  unsigned f = 0;
  // fast path: fully interior
  const int xs = x_start - 1;
  if (xs >= 0 && xs + nbits <= w) {
    const int wi = xs >> 5, sh = xs & 31;
    const unsigned lo = row[wi], hi = (wi + 1 < words) ? row[wi + 1] : 0u;
    f = __funnelshift_r(lo, hi, sh);
  } else {
    for (int k = 0; k < nbits; ++k) {
      int x = min(max(xs + k, 0), w - 1);
      f |= ((row[x >> 5] >> (x & 31)) & 1u) << k;
    }
  }
  return nbits >= 32 ? f : (f & ((1u << nbits) - 1u));
}

// ---- chamfer 3x3 distance transform, OpenCV's fixed-point two-pass recurrence (distanceTransform_3x3) -------
// One warp per frame; rows are sequential, each row is a lane-blocked min-plus scan.  Integer min/+ is
// associative, so the scan reproduces the raster recurrence bit for bit.
#define DT_HV 62587u        // cvRound(0.955f  * 65536)
#define DT_DG 89738u        // cvRound(1.3693f * 65536)
#define DT_INF 0x3FFFFFFF   // internal "no seed yet"; written out as OpenCV's DIST_MAX
#define DT_DISTMAX (0xFFFFFFFFu - DT_DG)

template <int P>
__global__ void __launch_bounds__(32) k_chamfer_dt(const uint32_t* __restrict__ edge_bits, const int32_t* __restrict__ dst_slots,
                                                   float* __restrict__ dt, unsigned* __restrict__ minmax, int level, int w, int h,
                                                   int words, int use_median) {
  extern __shared__ int smem[];
  int* bufA = smem;            // [w + 2], index x+1
  int* bufB = smem + (w + 2);
  const int lane = threadIdx.x;
  const int slot = dst_slots[blockIdx.x];
  const uint32_t* bits = edge_bits + size_t(slot) * h * words;
  int* gi = reinterpret_cast<int*>(dt + size_t(slot) * w * h);
  float* gf = dt + size_t(slot) * w * h;
  const int chunk = 32 * P, n_chunks = (w + chunk - 1) / chunk;
  for (int i = lane; i < w + 2; i += 32) { bufA[i] = DT_INF; bufB[i] = DT_INF; }
  __syncwarp();
  int* prev = bufA;
  int* cur = bufB;
  // ---------------- forward pass ----------------
  for (int y = 0; y < h; ++y) {
    const uint32_t* r0 = bits + size_t(max(y - 1, 0)) * words;
    const uint32_t* r1 = bits + size_t(y) * words;
    const uint32_t* r2 = bits + size_t(min(y + 1, h - 1)) * words;
    int carry = DT_INF;  // value of the pixel left of the chunk
    for (int c = 0; c < n_chunks; ++c) {
      const int xl = c * chunk + lane * P;   // first pixel of this lane
      unsigned ef = 0;
      if (xl < w) {
        if (use_median) ef = median3_bits(row_field(r0, words, w, xl, P + 2), row_field(r1, words, w, xl, P + 2), row_field(r2, words, w, xl, P + 2));
        else ef = row_field(r1, words, w, xl, P + 2) >> 1;
      }
      int d[P];
      int run = DT_INF;
      int pl = (xl < w) ? prev[xl] : DT_INF, pc = (xl < w) ? prev[xl + 1] : DT_INF;   // prev[x-1], prev[x] (index shift 1)
#pragma unroll
      for (int k = 0; k < P; ++k) {
        const int x = xl + k;
        const int pr = (x < w) ? prev[x + 2] : DT_INF;
        int cval = min(min(pl, pr) + int(DT_DG), pc + int(DT_HV));
        if ((ef >> k) & 1u) cval = 0;
        run = min(cval, run + int(DT_HV));
        d[k] = run;
        pl = pc; pc = pr;
      }
      // carry chain across lanes: out_i = min(L_i, out_{i-1} + HV*P)
      const int step = int(DT_HV) * P;
      int e = min(run, DT_INF) - step * lane;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, e, o); if (lane >= o) e = min(e, t); }
      // include the chunk carry (acts like a lane -1 with out = carry)
      const int with_carry = min(e + step * lane, carry + step * (lane + 1));
      int cin = __shfl_up_sync(0xffffffffu, with_carry, 1);
      if (lane == 0) cin = carry;
      cin = min(cin, DT_INF);
#pragma unroll
      for (int k = 0; k < P; ++k) {
        const int x = xl + k;
        const int v = min(min(d[k], cin + int(DT_HV) * (k + 1)), DT_INF);
        if (x < w) cur[x + 1] = v;
      }
      carry = min(__shfl_sync(0xffffffffu, with_carry, 31), DT_INF);
    }
    __syncwarp();
    int* grow = gi + size_t(y) * w;
    for (int i = lane; i < w; i += 32) grow[i] = cur[i + 1];
    int* t = prev; prev = cur; cur = t;
    __syncwarp();
  }
  // ---------------- backward pass ----------------
  for (int i = lane; i < w + 2; i += 32) prev[i] = DT_INF;   // "row below the image"
  __syncwarp();
  unsigned vmax = 0, vmin = 0xFFFFFFFFu;
  for (int y = h - 1; y >= 0; --y) {
    const int* grow = gi + size_t(y) * w;
    for (int i = lane; i < w; i += 32) cur[i + 1] = grow[i];
    __syncwarp();
    int carry = DT_INF;  // value of the pixel right of the chunk
    for (int c = n_chunks - 1; c >= 0; --c) {
      // mirrored lane order: lane 0 owns the right-most block of the chunk
      const int xr = c * chunk + (31 - lane) * P + (P - 1);  // right-most pixel of this lane
      int d[P];
      int run = DT_INF;
      // below[x+1], below[x]; the buffers hold pixels -1..w at indices 0..w+1 (borders stay INF)
      int pr = (xr + 2 <= w + 1) ? prev[xr + 2] : DT_INF, pc = (xr + 1 <= w + 1) ? prev[xr + 1] : DT_INF;
#pragma unroll
      for (int k = 0; k < P; ++k) {
        const int x = xr - k;
        const int pl = (x <= w + 1) ? prev[x] : DT_INF;   // below[x-1]
        const int t0 = (x < w) ? cur[x + 1] : DT_INF;
        const int cval = min(t0, min(min(pl, pr) + int(DT_DG), pc + int(DT_HV)));
        run = min(cval, run + int(DT_HV));
        d[k] = run;
        pr = pc; pc = pl;
      }
      const int step = int(DT_HV) * P;
      int e = min(run, DT_INF) - step * lane;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, e, o); if (lane >= o) e = min(e, t); }
      const int with_carry = min(e + step * lane, carry + step * (lane + 1));
      int cin = __shfl_up_sync(0xffffffffu, with_carry, 1);
      if (lane == 0) cin = carry;
      cin = min(cin, DT_INF);
#pragma unroll
      for (int k = 0; k < P; ++k) {
        const int x = xr - k;
        const int v = min(min(d[k], cin + int(DT_HV) * (k + 1)), DT_INF);
        if (x < w && x >= 0) cur[x + 1] = v;
      }
      carry = min(__shfl_sync(0xffffffffu, with_carry, 31), DT_INF);
    }
    __syncwarp();
    float* frow = gf + size_t(y) * w;
    for (int i = lane; i < w; i += 32) {
      const int v = cur[i + 1];
      const unsigned t0 = (v >= DT_INF) ? DT_DISTMAX : unsigned(v);
      vmax = max(vmax, t0); vmin = min(vmin, t0);
      frow[i] = float(t0) * (1.0f / 65536.0f);
    }
    int* t = prev; prev = cur; cur = t;
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmax = max(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    vmin = min(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
  }
  if (lane == 0) { minmax[(slot * EA_MAX_LEVELS + level) * 2] = vmin; minmax[(slot * EA_MAX_LEVELS + level) * 2 + 1] = vmax; }
}

// ---- cv::normalize(NORM_MINMAX, alpha=0, beta) on CV_32F -----------------------------------------------------
__global__ void __launch_bounds__(256) k_dt_normalize(float* __restrict__ dt, const int32_t* __restrict__ dst_slots,
                                                      const unsigned* __restrict__ minmax, int level, int npx, double beta) {
  const int slot = dst_slots[blockIdx.y];
  const float mn = float(minmax[(slot * EA_MAX_LEVELS + level) * 2]) * (1.0f / 65536.0f);
  const float mx = float(minmax[(slot * EA_MAX_LEVELS + level) * 2 + 1]) * (1.0f / 65536.0f);
  const double range = double(mx) - double(mn);
  const double scale = beta * (range > 2.220446049250313e-16 ? 1.0 / range : 0.0);
  const float fs = float(scale), fb = float(0.0 - double(mn) * scale);
  float* d = dt + size_t(slot) * npx;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += gridDim.x * blockDim.x) d[i] = __fadd_rn(__fmul_rn(d[i], fs), fb);
}

__global__ void __launch_bounds__(256) k_unpack_mask(const uint32_t* __restrict__ bits, int w, int h, int words, int median,
                                                     uint8_t* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w * h; i += gridDim.x * blockDim.x) {
    const int x = i % w, y = i / w;
    int edge;
    if (!median) edge = (bits[size_t(y) * words + (x >> 5)] >> (x & 31)) & 1;
    else {
      int cnt = 0;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int xx = min(max(x + dx, 0), w - 1), yy = min(max(y + dy, 0), h - 1);
          cnt += (bits[size_t(yy) * words + (xx >> 5)] >> (xx & 31)) & 1;
        }
      edge = cnt >= 5;
    }
    out[i] = edge ? 0 : 255;   // reference convention: B == 0 on edges (utils.cpp:60-70)
  }
}

template <int P>
cudaError_t launch_dt(const EaPrepArgs& A, int l, cudaStream_t stream) {
  const EaPrepLevel& L = A.lv[l];
  const size_t smem = size_t(L.w + 2) * 2 * sizeof(int);
  if (smem > 48 * 1024) cudaFuncSetAttribute(k_chamfer_dt<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  k_chamfer_dt<P><<<A.n, 32, smem, stream>>>(L.edge_bits, A.slots, L.dt, A.dt_minmax, l, L.w, L.h, L.words, A.use_median);
  return cudaGetLastError();
}

}  // namespace

cudaError_t ea_launch_preprocess(const EaPrepArgs& A, int sm_count, cudaStream_t stream, int* launches) {
  int nl = 0;
  cudaError_t err = cudaSuccess;
  const bool want_ref = (A.roles & EA_ROLE_REF) != 0, want_now = (A.roles & EA_ROLE_NOW) != 0;
  if (want_ref && !A.in_depth) return cudaErrorInvalidValue;
  const size_t px0 = size_t(A.lv[0].w) * A.lv[0].h;
  for (int l = 0; l < A.n_levels; ++l) {
    const EaPrepLevel& L = A.lv[l];
    const size_t pxl = size_t(L.w) * L.h;
    if (l > 0) {
      const EaPrepLevel& S = A.lv[l - 1];
      const uint8_t* sb = (l == 1) ? A.in_bgr : S.bgr;
      const uint16_t* sd = want_ref ? ((l == 1) ? A.in_depth : S.depth) : nullptr;
      const size_t sstride = (l == 1) ? px0 : size_t(S.w) * S.h;
      dim3 grid(unsigned((pxl + 255) / 256), unsigned(A.n));
      k_pyr_down<<<grid, 256, 0, stream>>>(sb, sd, sstride, (l == 1) ? nullptr : A.slots, L.bgr, L.depth, A.slots, S.w, S.h);
      ++nl;
    }
    const uint8_t* b = (l == 0) ? A.in_bgr : L.bgr;
    const uint16_t* d = want_ref ? ((l == 0) ? A.in_depth : L.depth) : nullptr;
    const int32_t* ss = (l == 0) ? nullptr : A.slots;
    dim3 grid(unsigned((L.w + ET_W - 1) / ET_W), unsigned((L.h + ET_H - 1) / ET_H), unsigned(A.n));
    k_edge_mask<<<grid, dim3(ET_W, ET_H), 0, stream>>>(b, d, pxl, ss, A.slots, L.edge_bits, want_ref ? L.ref_bits : nullptr, L.w, L.h, L.words, A.grad_threshold);
    ++nl;
    if (want_ref) {
      k_compact<<<A.n, CP_THREADS, 0, stream>>>(L.ref_bits, d, pxl, ss, A.slots, L.pts, A.n_pts, l, A.overflow, L.w, L.h, L.words, L.cap);
      ++nl;
    }
    if (want_now) {
      if (L.w <= 160) err = launch_dt<5>(A, l, stream);
      else if (L.w <= 352) err = launch_dt<11>(A, l, stream);
      else err = launch_dt<21>(A, l, stream);
      ++nl;
      if (err != cudaSuccess) return err;
      if (A.dt_normalize != EA_NORM_NONE) {
        int bx = int((pxl + 255) / 256);
        const int cap = sm_count * 8;
        if (bx > cap) bx = cap;
        k_dt_normalize<<<dim3(unsigned(bx), unsigned(A.n)), 256, 0, stream>>>(L.dt, A.slots, A.dt_minmax, l, int(pxl), A.dt_normalize == EA_NORM_255 ? 255.0 : 1.0);
        ++nl;
      }
    }
  }
  if (launches) *launches = nl;
  return cudaGetLastError();
}

cudaError_t ea_launch_unpack_mask(const uint32_t* bits, int w, int h, int words, int median, uint8_t* out, cudaStream_t stream) {
  k_unpack_mask<<<(w * h + 255) / 256, 256, 0, stream>>>(bits, w, h, words, median, out);
  return cudaGetLastError();
}
