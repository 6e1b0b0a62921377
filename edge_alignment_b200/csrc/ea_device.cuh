// Device-side data structures and the per-point evaluation shared by every kernel of the
// pose-solve path.  sm_100a only.
//
// Per-point math replaces, fused:
//   EAResidue::operator()<Jet<double,7>>      standalone/utils.h:48-80
//   ceres::BiCubicInterpolator::Evaluate      (Ceres, cubic_interpolation.h; see oracle/)
//   AutoDiff + QuaternionParameterization     standalone/utils.h:87, standalone_edge_align.cpp:277
//   LossFunction + Corrector                  standalone_edge_align.cpp:272
// Precision plan (DESIGN.md "Numerics"): pose transform + projection + fractional pixel offsets
// in fp64 (B200 keeps a 1:2 fp64 pipe), bicubic + analytic 1x6 Jacobian in fp32, normal
// equations reduced by warp butterflies in fp32 and accumulated in fp64.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ea_cabi.h"

#define EA_WARP 32
#define EA_NSUM 32  // butterfly slots: 0..20 upper-tri J^T J, 21..26 J^T r, 27 #failed, 28..31 unused
#define EA_SUMS 29  // reduced totals: [0..20] H upper-tri, [21..26] b, [27] #failed, [28] cost

struct EaLevelDesc {   // one per (slot, level); lives in device memory
  const float4* pts;   // point stream (EA_POINTS_PIXEL: {u,v,raw depth,1}; EA_POINTS_XYZ: {X,Y,Z,1})
  const int* n_pts;    // device-resident count (written by the compaction kernel)
  const float* dt;     // normalised distance transform [h][w] f32
  int w, h;
  int pts_mode, pad;
};

struct EaLevelGeom {   // per level, identical for every slot of a frameset (kernel parameter => constant bank)
  double fx, fy, cx, cy;
  double inv_fx, inv_fy;
  int w, h;
};

struct EaPose {  // current rotation (Eigen un-normalised quaternion formula) and translation
  double R[9];
  double t[3];
};

// Eigen::Quaternion::toRotationMatrix(), as called at standalone/utils.h:51-53 (no normalisation).
__device__ __forceinline__ void ea_pose_from_q(const double* x7, EaPose& P) {
  const double qw = x7[0], qx = x7[1], qy = x7[2], qz = x7[3];
  const double tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  const double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  const double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  P.R[0] = 1.0 - (tyy + tzz); P.R[1] = txy - twz;         P.R[2] = txz + twy;
  P.R[3] = txy + twz;         P.R[4] = 1.0 - (txx + tzz); P.R[5] = tyz - twx;
  P.R[6] = txz - twy;         P.R[7] = tyz + twx;         P.R[8] = 1.0 - (txx + tyy);
  P.t[0] = x7[4]; P.t[1] = x7[5]; P.t[2] = x7[6];
}

struct EaPointEval {
  float f, dfdu, dfdv;     // bicubic value and gradient (grid row == u, col == v: SEA:258)
  float yx, yy, yz;        // R * X
  float px, py, iz;        // transformed point (x', y') and 1/z'
  bool fail;               // |z'| < 0.01  (utils.h:70-73)
};

// ceres CubicHermiteSpline<1> (Catmull-Rom), fp32.
__device__ __forceinline__ void ea_cubic(float p0, float p1, float p2, float p3, float x, float& f, float& dfdx) {
  const float a = 0.5f * (-p0 + 3.0f * p1 - 3.0f * p2 + p3);
  const float b = 0.5f * (2.0f * p0 - 5.0f * p1 + 4.0f * p2 - p3);
  const float c = 0.5f * (-p0 + p2);
  f = fmaf(x, fmaf(x, fmaf(x, a, b), c), p1);
  dfdx = fmaf(x, fmaf(3.0f * a, x, 2.0f * b), c);
}
__device__ __forceinline__ float ea_cubic_val(float p0, float p1, float p2, float p3, float x) {
  const float a = 0.5f * (-p0 + 3.0f * p1 - 3.0f * p2 + p3);
  const float b = 0.5f * (2.0f * p0 - 5.0f * p1 + 4.0f * p2 - p3);
  const float c = 0.5f * (-p0 + p2);
  return fmaf(x, fmaf(x, fmaf(x, a, b), c), p1);
}

// Warp, project, bicubic lookup for one edge point.
//   ref: intrinsics of the level the point list was extracted at (back-projection, utils.cpp:235-237)
//   now: intrinsics of the distance transform's level (projection, utils.h:74-75)
template <bool XYZ>
__device__ __forceinline__ void ea_point_eval(const float4 p, const EaLevelGeom& ref, const EaLevelGeom& now,
                                              double inv_depth_scale, const EaPose& P, const float* __restrict__ dt,
                                              EaPointEval& o) {
  double X, Y, Z;
  if (XYZ) {
    X = double(p.x); Y = double(p.y); Z = double(p.z);
  } else {
    Z = double(p.z) * inv_depth_scale;
    X = (double(p.x) - ref.cx) * Z * ref.inv_fx;
    Y = (double(p.y) - ref.cy) * Z * ref.inv_fy;
  }
  const double yx = P.R[0] * X + P.R[1] * Y + P.R[2] * Z;
  const double yy = P.R[3] * X + P.R[4] * Y + P.R[5] * Z;
  const double yz = P.R[6] * X + P.R[7] * Y + P.R[8] * Z;
  const double px = yx + P.t[0], py = yy + P.t[1], pz = yz + P.t[2];
  o.fail = (pz < 0.01) && (pz > -0.01);
  const double iz = 1.0 / pz;
  const double u = now.fx * px * iz + now.cx;
  const double v = now.fy * py * iz + now.cy;
  // floor + fractional offsets in fp64, then hand fp32 to the interpolator
  double fu = floor(u), fv = floor(v);
  const int W = now.w, H = now.h;
  // keep the 4x4 footprint arithmetic inside int range: anything this far out clamps to one edge texel
  fu = fmin(fmax(fu, -4.0), double(W + 4));
  fv = fmin(fmax(fv, -4.0), double(H + 4));
  const int iu = int(fu), iv = int(fv);
  float du = float(u - fu), dv = float(v - fv);
  du = fminf(fmaxf(du, 0.0f), 1.0f);
  dv = fminf(fmaxf(dv, 0.0f), 1.0f);
  // Grid2D::GetValue clamp-to-edge
  const int x0 = min(max(iu - 1, 0), W - 1), x1 = min(max(iu, 0), W - 1), x2 = min(max(iu + 1, 0), W - 1),
            x3 = min(max(iu + 2, 0), W - 1);
  const float* r0 = dt + size_t(min(max(iv - 1, 0), H - 1)) * W;
  const float* r1 = dt + size_t(min(max(iv, 0), H - 1)) * W;
  const float* r2 = dt + size_t(min(max(iv + 1, 0), H - 1)) * W;
  const float* r3 = dt + size_t(min(max(iv + 2, 0), H - 1)) * W;
  // 16 texel gather (L1/L2-resident); issue all loads before use
  const float p00 = __ldg(r0 + x0), p01 = __ldg(r0 + x1), p02 = __ldg(r0 + x2), p03 = __ldg(r0 + x3);
  const float p10 = __ldg(r1 + x0), p11 = __ldg(r1 + x1), p12 = __ldg(r1 + x2), p13 = __ldg(r1 + x3);
  const float p20 = __ldg(r2 + x0), p21 = __ldg(r2 + x1), p22 = __ldg(r2 + x2), p23 = __ldg(r2 + x3);
  const float p30 = __ldg(r3 + x0), p31 = __ldg(r3 + x1), p32 = __ldg(r3 + x2), p33 = __ldg(r3 + x3);
  // BiCubicInterpolator::Evaluate: for each grid row (== image column x_k) spline along c (== image y),
  // then spline the four results along r (== image x).
  float f0, f1, f2, f3, d0, d1, d2, d3;
  ea_cubic(p00, p10, p20, p30, dv, f0, d0);
  ea_cubic(p01, p11, p21, p31, dv, f1, d1);
  ea_cubic(p02, p12, p22, p32, dv, f2, d2);
  ea_cubic(p03, p13, p23, p33, dv, f3, d3);
  ea_cubic(f0, f1, f2, f3, du, o.f, o.dfdu);
  o.dfdv = ea_cubic_val(d0, d1, d2, d3, du);
  o.yx = float(yx); o.yy = float(yy); o.yz = float(yz);
  o.px = float(px); o.py = float(py); o.iz = float(iz);
}

// Loss (ceres/loss_function.cc) + Corrector (rho'' <= 0 for all three => scale by sqrt(rho')) in fp32.
// Returns sqrt(rho'), writes rho(s).
__device__ __forceinline__ float ea_loss_eval(int type, float a, float r, float& rho0) {
  const float s = r * r;
  if (type == EA_LOSS_CAUCHY) {
    const float b = a * a, c = 1.0f / b;
    const float sum = fmaf(s, c, 1.0f);
    rho0 = b * log1pf(s * c);
    return rsqrtf(sum);
  } else if (type == EA_LOSS_HUBER) {
    const float b = a * a;
    if (s > b) {
      const float ar = fabsf(r);
      rho0 = 2.0f * a * ar - b;
      return sqrtf(a / ar);
    }
    rho0 = s;
    return 1.0f;
  }
  rho0 = s;
  return 1.0f;
}

// Analytic local Jacobian (collapsed closed form of AutoDiff x QuaternionParameterization,
// SURVEY.md A.3): g = dr/dp', J = [ 2 (R X) x g | g ], then robust re-weighting.
__device__ __forceinline__ void ea_jacobian(const EaPointEval& e, const EaLevelGeom& now, float w, float J[6]) {
  const float g0 = e.dfdu * float(now.fx) * e.iz;
  const float g1 = e.dfdv * float(now.fy) * e.iz;
  const float g2 = -(g0 * e.px + g1 * e.py) * e.iz;
  J[0] = 2.0f * (e.yy * g2 - e.yz * g1) * w;
  J[1] = 2.0f * (e.yz * g0 - e.yx * g2) * w;
  J[2] = 2.0f * (e.yx * g1 - e.yy * g0) * w;
  J[3] = g0 * w; J[4] = g1 * w; J[5] = g2 * w;
}

// Transposing warp reduction: every lane holds 32 partial sums v[0..31]; afterwards lane L holds in v[0]
// the warp-wide sum of slot L.  31 shuffles instead of 32 x 5.
template <int N>
__device__ __forceinline__ void ea_bfly_stage(float (&v)[EA_NSUM], bool upper, int bit) {
#pragma unroll
  for (int k = 0; k < N / 2; ++k) {
    const float send = upper ? v[k] : v[k + N / 2];
    const float keep = upper ? v[k + N / 2] : v[k];
    v[k] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
  }
}
__device__ __forceinline__ float ea_warp_transpose_reduce(float (&v)[EA_NSUM], int lane) {
  ea_bfly_stage<32>(v, lane & 16, 16);
  ea_bfly_stage<16>(v, lane & 8, 8);
  ea_bfly_stage<8>(v, lane & 4, 4);
  ea_bfly_stage<4>(v, lane & 2, 2);
  ea_bfly_stage<2>(v, lane & 1, 1);
  return v[0];
}
__device__ __forceinline__ double ea_warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// accumulate one point's outer products into the 32 fp32 slots
__device__ __forceinline__ void ea_accumulate(float (&acc)[EA_NSUM], const float J[6], float rw) {
  int k = 0;
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int c = a; c < 6; ++c) { acc[k] = fmaf(J[a], J[c], acc[k]); ++k; }
#pragma unroll
  for (int a = 0; a < 6; ++a) acc[21 + a] = fmaf(J[a], rw, acc[21 + a]);
}
