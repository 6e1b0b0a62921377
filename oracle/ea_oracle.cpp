// =============================================================================
// ea_oracle.cpp -- CPU fp64 restatement of edge_alignment's pose-solve hot path.
//
// THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, the smoke() check
// in __graft_entry__.py and bench.py's cpu_baseline / --impl reference legs may
// load it.  The product path (edge_alignment_b200/csrc) never links or calls it.
//
// What it restates (reference file:line, relative to kuwt/edge_alignment):
//   * standalone/utils.cpp:38-83    get_distance_transform   (blur->gray->Laplacian->
//                                    threshold->median->chamfer DT->min-max normalise)
//   * standalone/utils.cpp:201-281  get_aX                   (edge select + back-projection,
//                                    row-major ordered compaction)
//   * standalone/utils.h:38-99      EAResidue::operator()<T> (warp, project, bicubic lookup)
//   * standalone/standalone_edge_align.cpp:256-301           (problem assembly: stride,
//                                    per-block loss, quaternion parameterisation, solve)
//   * standalone/PoseManipUtils.cpp:3-27                     (q wxyz / t marshalling)
// and the third-party semantics those lines call into, which are NOT vendored in
// the reference and NOT installed in this image (Ceres Solver <2.2 -- README run
// used 1.12.0; OpenCV 3; Eigen 3).  Their published algorithms are restated here:
//   * ceres/cubic_interpolation.h   CubicHermiteSpline, Grid2D, BiCubicInterpolator
//   * ceres/jet.h + autodiff        forward-mode Jet<double,7>
//   * ceres/loss_function.cc, corrector.cc   Trivial/Cauchy/Huber + Corrector
//   * ceres/local_parameterization.cc        QuaternionParameterization
//   * ceres/trust_region_minimizer.cc, levenberg_marquardt_strategy.cc,
//     dense_qr_solver.cc                      LM trust region, Householder QR on [J;D]
//   * Eigen Quaternion::toRotationMatrix      (un-normalised formula)
//   * OpenCV imgproc: GaussianBlur 3x3, cvtColor RGB2GRAY (15-bit), Laplacian k=3,
//     convertScaleAbs, medianBlur 3, distanceTransform 3x3 (fixed-point two pass),
//     normalize MINMAX, resize 0.5 (INTER_LINEAR u8 / INTER_NEAREST u16)
//
// PINNING STATUS
//   * Preprocessing: PINNED bit-exactly against genuine OpenCV (cv2 4.13, IPP
//     disabled so the portable code path runs) on the reference's five bundled
//     frames -- tests/golden/make_golden.py generates, tests/test_oracle_*.py check.
//   * Solver side: PARITY UNPINNED at the Ceres boundary.  The reference ships no
//     tests or golden vectors and Ceres cannot be built here.  Independent pins
//     used instead: Jet Jacobian == closed form == central differences; survey
//     anchor values; scipy least_squares optimum; README loose
//     check (standalone/README.md:26-71).
// =============================================================================
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

// -----------------------------------------------------------------------------
// OpenCV imgproc restatements (integer exact)
// -----------------------------------------------------------------------------
inline int reflect101(int i, int n) {  // cv::BORDER_REFLECT_101 (BORDER_DEFAULT)
  if (n == 1) return 0;
  while (i < 0 || i >= n) {
    if (i < 0) i = -i;
    else i = 2 * n - 2 - i;
  }
  return i;
}
inline int replicate(int i, int n) { return i < 0 ? 0 : (i >= n ? n - 1 : i); }

// cv::GaussianBlur(src, dst, Size(3,3), 0, 0, BORDER_DEFAULT) on 8UC3
// (standalone/utils.cpp:50,215): separable [1 2 1]/4, fixed point, one final
// rounding => (sum w_ij p_ij + 8) >> 4.
void gaussian3_u8c3(const uint8_t* src, int w, int h, uint8_t* dst) {
  static const int k[3] = {1, 2, 1};
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      for (int c = 0; c < 3; ++c) {
        int acc = 0;
        for (int dy = -1; dy <= 1; ++dy) {
          int yy = reflect101(y + dy, h);
          for (int dx = -1; dx <= 1; ++dx) {
            int xx = reflect101(x + dx, w);
            acc += k[dy + 1] * k[dx + 1] * src[(size_t(yy) * w + xx) * 3 + c];
          }
        }
        dst[(size_t(y) * w + x) * 3 + c] = uint8_t((acc + 8) >> 4);
      }
}

// cv::blur(src, dst, Size(3,3)) on 8UC3 (standalone/utils.cpp:88): box / 9, rounded.
void box3_u8c3(const uint8_t* src, int w, int h, uint8_t* dst) {
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      for (int c = 0; c < 3; ++c) {
        int acc = 0;
        for (int dy = -1; dy <= 1; ++dy) {
          int yy = reflect101(y + dy, h);
          for (int dx = -1; dx <= 1; ++dx)
            acc += src[(size_t(yy) * w + reflect101(x + dx, w)) * 3 + c];
        }
        dst[(size_t(y) * w + x) * 3 + c] = uint8_t(std::lrint(acc * (1.0 / 9.0)));
      }
}

// cv::cvtColor(CV_RGB2GRAY) applied to imread's BGR data (standalone/utils.cpp:51,216):
// channel 0 gets the "R" weight.  OpenCV 4.x integer form, shift 15.
void rgb2gray_u8(const uint8_t* src, int w, int h, uint8_t* dst) {
  for (size_t i = 0, n = size_t(w) * h; i < n; ++i)
    dst[i] = uint8_t((src[3 * i] * 9798 + src[3 * i + 1] * 19235 + src[3 * i + 2] * 3735 + 16384) >> 15);
}

// cv::Laplacian(CV_16S, ksize=3) + cv::convertScaleAbs (standalone/utils.cpp:53-55,218-220):
// kernel [2 0 2; 0 -8 0; 2 0 2], REFLECT_101, then min(|x|,255).
void laplacian3_abs_u8(const uint8_t* g, int w, int h, uint8_t* dst) {
  for (int y = 0; y < h; ++y) {
    int ym = reflect101(y - 1, h), yp = reflect101(y + 1, h);
    for (int x = 0; x < w; ++x) {
      int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
      int v = 2 * (g[size_t(ym) * w + xm] + g[size_t(ym) * w + xp] + g[size_t(yp) * w + xm] + g[size_t(yp) * w + xp]) -
              8 * g[size_t(y) * w + x];
      v = v < 0 ? -v : v;
      dst[size_t(y) * w + x] = uint8_t(v > 255 ? 255 : v);
    }
  }
}

// cv::medianBlur(src, dst, 3) on 8UC1, BORDER_REPLICATE (standalone/utils.cpp:75).
void median3_u8(const uint8_t* src, int w, int h, uint8_t* dst) {
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      uint8_t v[9];
      int k = 0;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) v[k++] = src[size_t(replicate(y + dy, h)) * w + replicate(x + dx, w)];
      std::nth_element(v, v + 4, v + 9);
      dst[size_t(y) * w + x] = v[4];
    }
}

// cv::distanceTransform(src, dst, DIST_L2, 3) -- OpenCV's portable
// distanceTransform_3x3: 16.16 fixed point two-pass chamfer, a=0.955 b=1.3693
// (standalone/utils.cpp:80).  Distance to the nearest ZERO pixel.
void chamfer3_dt(const uint8_t* src, int w, int h, float* dst) {
  const unsigned HV = unsigned(std::lrint(0.955f * 65536.0));
  const unsigned DG = unsigned(std::lrint(1.3693f * 65536.0));
  const unsigned DIST_MAX = UINT_MAX - DG;
  const float scale = 1.f / 65536.f;
  const int step = w + 2;
  std::vector<unsigned> temp(size_t(step) * (h + 2), DIST_MAX);
  for (int i = 0; i < h; ++i) {
    unsigned* t = temp.data() + size_t(i + 1) * step + 1;
    const uint8_t* s = src + size_t(i) * w;
    for (int j = 0; j < w; ++j) {
      if (!s[j]) t[j] = 0;
      else {
        unsigned t0 = t[j - step - 1] + DG, v = t[j - step] + HV;
        if (t0 > v) t0 = v;
        v = t[j - step + 1] + DG; if (t0 > v) t0 = v;
        v = t[j - 1] + HV;        if (t0 > v) t0 = v;
        t[j] = t0 > DIST_MAX ? DIST_MAX : t0;
      }
    }
  }
  for (int i = h - 1; i >= 0; --i) {
    unsigned* t = temp.data() + size_t(i + 1) * step + 1;
    float* d = dst + size_t(i) * w;
    for (int j = w - 1; j >= 0; --j) {
      unsigned t0 = t[j];
      if (t0 > HV) {
        unsigned v = t[j + step + 1] + DG; if (t0 > v) t0 = v;
        v = t[j + step] + HV;              if (t0 > v) t0 = v;
        v = t[j + step - 1] + DG;          if (t0 > v) t0 = v;
        v = t[j + 1] + HV;                 if (t0 > v) t0 = v;
        t[j] = t0;
      }
      t0 = t0 > DIST_MAX ? DIST_MAX : t0;
      d[j] = float(t0) * scale;
    }
  }
}

// cv::normalize(src, dst, alpha, beta, NORM_MINMAX) on CV_32F (standalone/utils.cpp:81):
// scale=(beta-alpha)/(max-min), shift=alpha-min*scale in double; dst=src*float(scale)+float(shift).
void normalize_minmax(float* d, size_t n, double alpha, double beta) {
  float mn = d[0], mx = d[0];
  for (size_t i = 1; i < n; ++i) { mn = std::min(mn, d[i]); mx = std::max(mx, d[i]); }
  double scale = (beta - alpha) * ((double(mx) - double(mn)) > DBL_EPSILON ? 1.0 / (double(mx) - double(mn)) : 0.0);
  double shift = alpha - double(mn) * scale;
  float fs = float(scale), fb = float(shift);
  for (size_t i = 0; i < n; ++i) d[i] = d[i] * fs + fb;
}

// cv::resize(src, dst, Size(), 0.5, 0.5) INTER_LINEAR on 8UC3 (src/ea.cpp:38 precedent):
// exact 2x decimation => (a+b+c+d+2)>>2.
void half_linear_u8c3(const uint8_t* s, int w, int h, uint8_t* d) {
  int w2 = w / 2, h2 = h / 2;
  for (int y = 0; y < h2; ++y)
    for (int x = 0; x < w2; ++x)
      for (int c = 0; c < 3; ++c) {
        const uint8_t* p = s + (size_t(2 * y) * w + 2 * x) * 3 + c;
        d[(size_t(y) * w2 + x) * 3 + c] = uint8_t((p[0] + p[3] + p[size_t(w) * 3] + p[size_t(w) * 3 + 3] + 2) >> 2);
      }
}
// cv::resize(..., INTER_NEAREST) on 16UC1: dst(x,y)=src(2x,2y) (pyramid extension).
void half_nearest_u16(const uint16_t* s, int w, int h, uint16_t* d) {
  int w2 = w / 2, h2 = h / 2;
  for (int y = 0; y < h2; ++y)
    for (int x = 0; x < w2; ++x) d[size_t(y) * w2 + x] = s[size_t(2 * y) * w + 2 * x];
}

// cv::Canny(src, dst, t1, t2, 3, L2gradient) -- OpenCV's portable algorithm (canny.cpp): Sobel 3x3 with
// BORDER_REPLICATE per channel, per pixel the channel with the largest magnitude (first on ties), non-maximum
// suppression with the fixed-point tan(22.5 deg) test, hysteresis over 8-neighbours.  cn = 1 or 3.
// Call sites: standalone/utils.cpp:94 (gray, 30/90, L1), src/SolveEA.cpp:46,102 (colour, 150/100 swapped, L2).
void canny_u8(const uint8_t* src, int w, int h, int cn, double t1, double t2, bool l2, uint8_t* dst) {
  if (t1 > t2) std::swap(t1, t2);
  if (l2) {
    t1 = std::min(32767.0, t1); t2 = std::min(32767.0, t2);
    if (t1 > 0) t1 *= t1;
    if (t2 > 0) t2 *= t2;
  }
  const int low = int(std::floor(t1)), high = int(std::floor(t2));
  std::vector<int> mag(size_t(w + 2) * (h + 2), 0);
  std::vector<short> gx(size_t(w) * h), gy(size_t(w) * h);
  auto px = [&](int x, int y, int c) { return int(src[(size_t(replicate(y, h)) * w + replicate(x, w)) * cn + c]); };
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      int best = -1, bdx = 0, bdy = 0;
      for (int c = 0; c < cn; ++c) {
        const int dx = (px(x + 1, y - 1, c) + 2 * px(x + 1, y, c) + px(x + 1, y + 1, c)) - (px(x - 1, y - 1, c) + 2 * px(x - 1, y, c) + px(x - 1, y + 1, c));
        const int dy = (px(x - 1, y + 1, c) + 2 * px(x, y + 1, c) + px(x + 1, y + 1, c)) - (px(x - 1, y - 1, c) + 2 * px(x, y - 1, c) + px(x + 1, y - 1, c));
        const int m = l2 ? dx * dx + dy * dy : std::abs(dx) + std::abs(dy);
        if (m > best) { best = m; bdx = dx; bdy = dy; }
      }
      mag[size_t(y + 1) * (w + 2) + x + 1] = best; gx[size_t(y) * w + x] = short(bdx); gy[size_t(y) * w + x] = short(bdy);
    }
  // map: 1 = not an edge, 0 = candidate, 2 = edge; 1-pixel border of 1s
  std::vector<uint8_t> map(size_t(w + 2) * (h + 2), 1);
  std::vector<int> stack;
  const int TG22 = 13573, ms = w + 2;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const int* M = &mag[size_t(y + 1) * ms + x + 1];
      const int m = M[0];
      bool is_max = false;
      if (m > low) {
        const int xs = gx[size_t(y) * w + x], ys = gy[size_t(y) * w + x];
        const int ax = std::abs(xs), ay = std::abs(ys) << 15;
        const int tg22x = ax * TG22;
        if (ay < tg22x) is_max = (m > M[-1] && m >= M[1]);
        else {
          const int tg67x = tg22x + (ax << 16);
          if (ay > tg67x) is_max = (m > M[-ms] && m >= M[ms]);
          else { const int sg = (xs ^ ys) < 0 ? -1 : 1; is_max = (m > M[-ms - sg] && m > M[ms + sg]); }
        }
      }
      uint8_t& c = map[size_t(y + 1) * ms + x + 1];
      if (is_max) { if (m > high) { c = 2; stack.push_back(int(&c - map.data())); } else c = 0; }
      else c = 1;
    }
  while (!stack.empty()) {
    const int i = stack.back(); stack.pop_back();
    const int nb[8] = {-ms - 1, -ms, -ms + 1, -1, 1, ms - 1, ms, ms + 1};
    for (int k = 0; k < 8; ++k) if (map[i + nb[k]] == 0) { map[i + nb[k]] = 2; stack.push_back(i + nb[k]); }
  }
  for (int y = 0; y < h; ++y) for (int x = 0; x < w; ++x) dst[size_t(y) * w + x] = map[size_t(y + 1) * ms + x + 1] == 2 ? 255 : 0;
}

// cv::distanceTransform(src, dst, DIST_L2, DIST_MASK_PRECISE) -- OpenCV's trueDistTrans (distransform.cpp): a column pass
// (vertical distance to the nearest zero pixel; "infinite" when a column has none) and a row pass taking the lower
// envelope of the parabolas.  All quantities are integers < 2^24, so the envelope value is the
// exact integer minimum and the result is sqrtf of it.   Call site: src/SolveEA.cpp:108.
void exact_edt(const uint8_t* src, int w, int h, float* dst) {
  const int INF = 1 << 20;                 // "no zero pixel in this column"
  std::vector<int> g(size_t(w) * h);
  for (int x = 0; x < w; ++x) {
    int dist = INF;
    for (int y = h - 1; y >= 0; --y) { dist = src[size_t(y) * w + x] == 0 ? 0 : std::min(dist + 1, INF); g[size_t(y) * w + x] = dist; }
    dist = INF;
    for (int y = 0; y < h; ++y) { dist = src[size_t(y) * w + x] == 0 ? 0 : std::min(dist + 1, INF); g[size_t(y) * w + x] = std::min(g[size_t(y) * w + x], dist); }
  }
  for (int y = 0; y < h; ++y) {
    const int* gr = &g[size_t(y) * w];
    for (int q = 0; q < w; ++q) {
      long long best = -1;
      for (int p = 0; p < w; ++p) {
        if (gr[p] >= INF) continue;
        const long long v = (long long)(q - p) * (q - p) + (long long)gr[p] * gr[p];
        if (best < 0 || v < best) best = v;
      }
      dst[size_t(y) * w + q] = best < 0 ? 65536.0f : std::sqrt(float(best));   // OpenCV's value when the image has no zero pixel
    }
  }
}

// gradient magnitude image shared by get_aX and get_distance_transform
void laplacian_edge_strength(const uint8_t* bgr, int w, int h, uint8_t* lap8) {
  std::vector<uint8_t> blur(size_t(w) * h * 3), gray(size_t(w) * h);
  gaussian3_u8c3(bgr, w, h, blur.data());
  rgb2gray_u8(blur.data(), w, h, gray.data());
  laplacian3_abs_u8(gray.data(), w, h, lap8);
}

// -----------------------------------------------------------------------------
// Ceres restatements
// -----------------------------------------------------------------------------
template <int N>
struct Jet {  // ceres/jet.h
  double a;
  double v[N];
  Jet() : a(0) { for (int i = 0; i < N; ++i) v[i] = 0; }
  Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; }  // NOLINT
  Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0; v[k] = 1.0; }
};
template <int N> Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> Jet<N> operator-(const Jet<N>& f) { Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> Jet<N> operator/(const Jet<N>& f, const Jet<N>& g) {
  // ceres/jet.h: h = f/g ; dh = (df - h dg)/g
  Jet<N> h; const double gi = 1.0 / g.a; h.a = f.a * gi;
  for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - h.a * g.v[i]) * gi;
  return h;
}
template <int N> Jet<N> operator+(const Jet<N>& f, double s) { Jet<N> h = f; h.a += s; return h; }
template <int N> Jet<N> operator+(double s, const Jet<N>& f) { return f + s; }
template <int N> Jet<N> operator-(const Jet<N>& f, double s) { Jet<N> h = f; h.a -= s; return h; }
template <int N> Jet<N> operator-(double s, const Jet<N>& f) { Jet<N> h = -f; h.a += s; return h; }
template <int N> Jet<N> operator*(const Jet<N>& f, double s) { Jet<N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
template <int N> Jet<N> operator*(double s, const Jet<N>& f) { return f * s; }
template <int N> bool operator<(const Jet<N>& f, const Jet<N>& g) { return f.a < g.a; }
template <int N> bool operator>(const Jet<N>& f, const Jet<N>& g) { return f.a > g.a; }
inline double scalar_part(double x) { return x; }
template <int N> double scalar_part(const Jet<N>& x) { return x.a; }

// ceres/cubic_interpolation.h : CubicHermiteSpline<1> (Catmull-Rom)
inline void cubic_hermite(double p0, double p1, double p2, double p3, double x, double* f, double* dfdx) {
  const double a = 0.5 * (-p0 + 3.0 * p1 - 3.0 * p2 + p3);
  const double b = 0.5 * (2.0 * p0 - 5.0 * p1 + 4.0 * p2 - p3);
  const double c = 0.5 * (-p0 + p2);
  const double d = p1;
  if (f) *f = d + x * (c + x * (b + x * a));
  if (dfdx) *dfdx = c + x * (2.0 * b + 3.0 * a * x);
}

// ceres::Grid2D<double,1> (row-major, clamp-to-edge) built as in
// standalone_edge_align.cpp:258 over Eigen's COLUMN-major copy of the DT:
//   Grid2D(e_disTrans.data(), 0, cols, 0, rows)  => grid row == image x (u),
//   grid col == image y (v), value(r=x,c=y) == DT[y][x].
// We keep the DT row-major [y][x] (f32 widened on read, like cv2eigen) and index it
// accordingly -- same numbers, no transpose copy.
struct DtGrid {
  const float* dt; int w, h;
  inline double get(int r /*x*/, int c /*y*/) const {
    int xi = std::min(std::max(0, r), w - 1);
    int yi = std::min(std::max(0, c), h - 1);
    return double(dt[size_t(yi) * w + xi]);
  }
};

// ceres::BiCubicInterpolator::Evaluate(r, c, f, dfdr, dfdc)
inline void bicubic_eval(const DtGrid& g, double r, double c, double* f, double* dfdr, double* dfdc) {
  const int row = int(std::floor(r));
  const int col = int(std::floor(c));
  double fk[4], dfk[4];
  for (int k = 0; k < 4; ++k) {
    double p0 = g.get(row - 1 + k, col - 1), p1 = g.get(row - 1 + k, col), p2 = g.get(row - 1 + k, col + 1),
           p3 = g.get(row - 1 + k, col + 2);
    cubic_hermite(p0, p1, p2, p3, c - col, &fk[k], &dfk[k]);
  }
  cubic_hermite(fk[0], fk[1], fk[2], fk[3], r - row, f, dfdr);
  if (dfdc) cubic_hermite(dfk[0], dfk[1], dfk[2], dfk[3], r - row, dfdc, nullptr);
}
inline void interp(const DtGrid& g, const double& r, const double& c, double* f) { bicubic_eval(g, r, c, f, nullptr, nullptr); }
template <int N> void interp(const DtGrid& g, const Jet<N>& r, const Jet<N>& c, Jet<N>* f) {
  double fv, dr, dc;
  bicubic_eval(g, r.a, c.a, &fv, &dr, &dc);
  f->a = fv;
  for (int i = 0; i < N; ++i) f->v[i] = dr * r.v[i] + dc * c.v[i];
}

// standalone/utils.h:38-99  EAResidue (functor restated; T = double or Jet<7>)
struct EAResidue {
  double fx, fy, cx, cy, X, Y, Z;
  const DtGrid* grid;
  template <typename T>
  bool operator()(const T* quat, const T* t, T* residue) const {
    // Eigen::Quaternion<T>(w,x,y,z).toRotationMatrix()  (utils.h:51-53; un-normalised)
    const T qw = quat[0], qx = quat[1], qy = quat[2], qz = quat[3];
    const T tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
    const T twx = tx * qw, twy = ty * qw, twz = tz * qw;
    const T txx = tx * qx, txy = ty * qx, txz = tz * qx;
    const T tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    T R[3][3];
    R[0][0] = 1.0 - (tyy + tzz); R[0][1] = txy - twz;         R[0][2] = txz + twy;
    R[1][0] = txy + twz;         R[1][1] = 1.0 - (txx + tzz); R[1][2] = tyz - twx;
    R[2][0] = txz - twy;         R[2][1] = tyz + twx;         R[2][2] = 1.0 - (txx + tyy);
    // b_X = b_T_a * (X,Y,Z,1)  (utils.h:60-67)
    T bX[3];
    for (int i = 0; i < 3; ++i) bX[i] = R[i][0] * X + R[i][1] * Y + R[i][2] * Z + t[i];
    // utils.h:70-73
    if (scalar_part(bX[2]) < 0.01 && scalar_part(bX[2]) > -0.01) return false;
    T u = fx * bX[0] / bX[2] + cx;  // utils.h:74
    T v = fy * bX[1] / bX[2] + cy;  // utils.h:75
    interp(*grid, u, v, residue);   // utils.h:77
    return true;
  }
};

// standalone/utils.h:101-421  EAResidueEx (radial-tangential distortion, :140-149), EAResidueSecondCam (rigid rig,
// :244-256: b_T_a_SecCam = T_1to2 * b_T_a * T_1to2_inv) and EAResidueSecondCamEx (both) -- one functor, same arithmetic
// order as the reference's three classes; use_rig / use_dist select the variant.
struct EAResidueVariant {
  double fx, fy, cx, cy, X, Y, Z;
  const DtGrid* grid;
  bool use_dist; double k1, k2, p1, p2, k3;          // functor argument order (k1,k2,p1,p2,k3), utils.h:106
  bool use_rig; const double* T21; const double* T12;   // row-major 4x4: trans_1to2, trans_1to2_inv
  template <typename T>
  bool operator()(const T* quat, const T* t, T* residue) const {
    const T qw = quat[0], qx = quat[1], qy = quat[2], qz = quat[3];
    const T tx = 2.0 * qx, ty = 2.0 * qy, tz = 2.0 * qz;
    const T twx = tx * qw, twy = ty * qw, twz = tz * qw;
    const T txx = tx * qx, txy = ty * qx, txz = tz * qx;
    const T tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    T M[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) M[i][j] = T(0.0);
    M[0][0] = 1.0 - (tyy + tzz); M[0][1] = txy - twz;         M[0][2] = txz + twy;
    M[1][0] = txy + twz;         M[1][1] = 1.0 - (txx + tzz); M[1][2] = tyz - twx;
    M[2][0] = txz - twy;         M[2][1] = tyz + twx;         M[2][2] = 1.0 - (txx + tyy);
    M[0][3] = t[0]; M[1][3] = t[1]; M[2][3] = t[2]; M[3][3] = T(1.0);
    if (use_rig) {   // Eigen evaluates (A * B) * C
      T AB[4][4], ABC[4][4];
      for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { T s = T(0.0); for (int k = 0; k < 4; ++k) s = s + M[k][j] * T21[i * 4 + k]; AB[i][j] = s; }
      for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { T s = T(0.0); for (int k = 0; k < 4; ++k) s = s + AB[i][k] * T12[k * 4 + j]; ABC[i][j] = s; }
      for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) M[i][j] = ABC[i][j];
    }
    T bX[3];
    for (int i = 0; i < 3; ++i) bX[i] = M[i][0] * X + M[i][1] * Y + M[i][2] * Z + M[i][3];
    if (scalar_part(bX[2]) < 0.01 && scalar_part(bX[2]) > -0.01) return false;
    T u, v;
    if (use_dist) {
      T x = bX[0] / bX[2], y = bX[1] / bX[2];
      T r2 = x * x + y * y, r4 = r2 * r2, r6 = r4 * r2;
      T radial = 1.0 + k1 * r2 + k2 * r4 + k3 * r6;
      T dx = x * radial + 2.0 * p1 * x * y + p2 * (r2 + 2.0 * x * x);
      T dy = y * radial + 2.0 * p2 * x * y + p1 * (r2 + 2.0 * y * y);
      u = fx * dx + cx; v = fy * dy + cy;
    } else {
      u = fx * bX[0] / bX[2] + cx; v = fy * bX[1] / bX[2] + cy;
    }
    interp(*grid, u, v, residue);
    return true;
  }
};

enum { LOSS_TRIVIAL = 0, LOSS_CAUCHY = 1, LOSS_HUBER = 2 };
// ceres/loss_function.cc
inline void loss_eval(int type, double a, double s, double rho[3]) {
  if (type == LOSS_CAUCHY) {
    const double b = a * a, c = 1.0 / b;
    const double sum = 1.0 + s * c, inv = 1.0 / sum;
    rho[0] = b * std::log(sum);
    rho[1] = std::max(DBL_MIN, inv);
    rho[2] = -c * (inv * inv);
  } else if (type == LOSS_HUBER) {
    const double b = a * a;
    if (s > b) {
      const double r = std::sqrt(s);
      rho[0] = 2.0 * a * r - b;
      rho[1] = std::max(DBL_MIN, a / r);
      rho[2] = -rho[1] / (2.0 * s);
    } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
  } else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
}

// ceres::QuaternionParameterization
inline void quat_product(const double z[4], const double w[4], double zw[4]) {
  zw[0] = z[0] * w[0] - z[1] * w[1] - z[2] * w[2] - z[3] * w[3];
  zw[1] = z[0] * w[1] + z[1] * w[0] + z[2] * w[3] - z[3] * w[2];
  zw[2] = z[0] * w[2] - z[1] * w[3] + z[2] * w[0] + z[3] * w[1];
  zw[3] = z[0] * w[3] + z[1] * w[2] - z[2] * w[1] + z[3] * w[0];
}
inline void quat_plus(const double x[4], const double d[3], double out[4]) {
  const double n = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  if (n > 0.0) {
    const double s = std::sin(n) / n;
    double qd[4] = {std::cos(n), s * d[0], s * d[1], s * d[2]};
    quat_product(qd, x, out);
  } else { for (int i = 0; i < 4; ++i) out[i] = x[i]; }
}
inline void quat_local_jacobian(const double x[4], double P[12] /*4x3 row-major*/) {
  P[0] = -x[1]; P[1]  = -x[2]; P[2]  = -x[3];
  P[3] =  x[0]; P[4]  =  x[3]; P[5]  = -x[2];
  P[6] = -x[3]; P[7]  =  x[0]; P[8]  =  x[1];
  P[9] =  x[2]; P[10] = -x[1]; P[11] =  x[0];
}
// Plus on the 7-vector (q block with QuaternionParameterization, t block identity)
inline void pose_plus(const double x[7], const double d[6], double out[7]) {
  quat_plus(x, d, out);
  for (int i = 0; i < 3; ++i) out[4 + i] = x[4 + i] + d[3 + i];
}

struct Options {  // ceres::Solver::Options defaults; SEA:282-284 only sets DENSE_QR + stdout
  int max_num_iterations = 50;
  double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
  double initial_radius = 1e4, max_radius = 1e16, min_radius = 1e-32;
  double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
  int jacobi_scaling = 1, max_consecutive_invalid = 5;
  int loss_type = LOSS_CAUCHY; double loss_scale = 1.0;
  int strategy = 0;   // 0 LEVENBERG_MARQUARDT (Ceres default), 1 DOGLEG / TRADITIONAL_DOGLEG (src/SolveEA.cpp:192)
};

enum Termination { TERM_CONVERGENCE_GRADIENT = 1, TERM_CONVERGENCE_FUNCTION = 2, TERM_CONVERGENCE_PARAMETER = 3,
                   TERM_CONVERGENCE_MIN_RADIUS = 4, TERM_NO_CONVERGENCE = 5, TERM_FAILURE_EVAL_X0 = 6,
                   TERM_FAILURE_INVALID_STEPS = 7 };

struct View {          // one camera's residual blocks (standalone_edge_align.cpp:791-803 adds one loop per camera)
  const double* pts; int n_total; int stride;
  double fx, fy, cx, cy;
  DtGrid grid;
  bool use_dist = false; double dist[5] = {0, 0, 0, 0, 0};
  bool use_rig = false; double T21[16], T12[16];
  int n_res() const { return (n_total + stride - 1) / stride; }
};
struct Problem {
  std::vector<View> views;
  Problem() {}
  Problem(const double* pts, int n_total, int stride, double fx, double fy, double cx, double cy, DtGrid g) {
    View v; v.pts = pts; v.n_total = n_total; v.stride = stride; v.fx = fx; v.fy = fy; v.cx = cx; v.cy = cy; v.grid = g;
    views.push_back(v);
  }
  int n_res() const { int n = 0; for (const View& v : views) n += v.n_res(); return n; }
  const View& locate(int bi, int* local) const {
    size_t k = 0;
    while (bi >= views[k].n_res()) { bi -= views[k].n_res(); ++k; }
    *local = bi;
    return views[k];
  }
};

// One residual block: autodiff via Jet<7>, local-parameterisation product, loss+corrector.
// Returns false if the functor returns false (|z|<0.01).
inline bool eval_block(const Problem& p, int bi, const double x[7], const Options& o, double* cost, double* r_out,
                       double* J_out /*6 or null*/, double* raw_r /*unrobustified, or null*/) {
  int li;
  const View& V = p.locate(bi, &li);
  const double* P3 = V.pts + size_t(li) * V.stride * 3;
  const bool plain = !V.use_dist && !V.use_rig;
  EAResidue f{V.fx, V.fy, V.cx, V.cy, P3[0], P3[1], P3[2], &V.grid};
  EAResidueVariant fv{V.fx, V.fy, V.cx, V.cy, P3[0], P3[1], P3[2], &V.grid, V.use_dist, V.dist[0], V.dist[1], V.dist[2], V.dist[3], V.dist[4],
                      V.use_rig, V.T21, V.T12};
  double r, Jq[4] = {0, 0, 0, 0}, Jt[3] = {0, 0, 0};
  if (J_out) {
    Jet<7> q[4], t[3], res;
    for (int i = 0; i < 4; ++i) q[i] = Jet<7>(x[i], i);
    for (int i = 0; i < 3; ++i) t[i] = Jet<7>(x[4 + i], 4 + i);
    if (!(plain ? f(q, t, &res) : fv(q, t, &res))) return false;
    r = res.a;
    for (int i = 0; i < 4; ++i) Jq[i] = res.v[i];
    for (int i = 0; i < 3; ++i) Jt[i] = res.v[4 + i];
  } else {
    if (!(plain ? f(x, x + 4, &r) : fv(x, x + 4, &r))) return false;
  }
  if (raw_r) *raw_r = r;
  const double sq = r * r;
  double rho[3];
  loss_eval(o.loss_type, o.loss_scale, sq, rho);
  *cost = 0.5 * rho[0];
  // Corrector: rho''<=0 for all three losses => scale by sqrt(rho')
  const double s1 = std::sqrt(rho[1]);
  if (J_out) {
    double P[12];
    quat_local_jacobian(x, P);
    for (int j = 0; j < 3; ++j) J_out[j] = (Jq[0] * P[j] + Jq[1] * P[3 + j] + Jq[2] * P[6 + j] + Jq[3] * P[9 + j]) * s1;
    for (int j = 0; j < 3; ++j) J_out[3 + j] = Jt[j] * s1;
  }
  *r_out = r * s1;
  return true;
}

// Evaluate the whole problem.  J (n x 6 row-major) optional.
inline bool eval_problem(const Problem& p, const double x[7], const Options& o, double* cost, std::vector<double>* r,
                         std::vector<double>* J) {
  const int n = p.n_res();
  double c = 0;
  if (r) r->assign(n, 0.0);
  if (J) J->assign(size_t(n) * 6, 0.0);
  for (int i = 0; i < n; ++i) {
    double ci, ri;
    if (!eval_block(p, i, x, o, &ci, &ri, J ? J->data() + size_t(i) * 6 : nullptr, nullptr)) return false;
    c += ci;
    if (r) (*r)[i] = ri;
  }
  *cost = c;
  return true;
}

// Householder QR least squares: min || A y - b ||, A is m x 6 row-major (DenseQRSolver on [J; D]).
inline bool householder_lstsq6(std::vector<double>& A, std::vector<double>& b, int m, double y[6]) {
  const int n = 6;
  for (int k = 0; k < n; ++k) {
    double norm = 0;
    for (int i = k; i < m; ++i) norm += A[size_t(i) * n + k] * A[size_t(i) * n + k];
    norm = std::sqrt(norm);
    if (norm == 0.0) return false;
    const double alpha = A[size_t(k) * n + k] > 0 ? -norm : norm;
    // v = x - alpha e1
    A[size_t(k) * n + k] -= alpha;
    double vnorm2 = 0;
    for (int i = k; i < m; ++i) vnorm2 += A[size_t(i) * n + k] * A[size_t(i) * n + k];
    if (vnorm2 == 0.0) return false;
    for (int j = k + 1; j < n; ++j) {
      double dot = 0;
      for (int i = k; i < m; ++i) dot += A[size_t(i) * n + k] * A[size_t(i) * n + j];
      const double f = 2.0 * dot / vnorm2;
      for (int i = k; i < m; ++i) A[size_t(i) * n + j] -= f * A[size_t(i) * n + k];
    }
    double dot = 0;
    for (int i = k; i < m; ++i) dot += A[size_t(i) * n + k] * b[i];
    const double f = 2.0 * dot / vnorm2;
    for (int i = k; i < m; ++i) b[i] -= f * A[size_t(i) * n + k];
    // store R's diagonal in place of v's head: keep v below, remember alpha separately
    // (we only need R for back substitution: R_kk = alpha, R_kj already in A[k][j])
    A[size_t(k) * n + k] = alpha;
    for (int i = k + 1; i < m; ++i) A[size_t(i) * n + k] = 0.0;
  }
  for (int k = n - 1; k >= 0; --k) {
    double s = b[k];
    for (int j = k + 1; j < n; ++j) s -= A[size_t(k) * n + j] * y[j];
    y[k] = s / A[size_t(k) * n + k];
    if (!std::isfinite(y[k])) return false;
  }
  return true;
}

struct IterRecord { double cost, cost_change, gradient_max_norm, step_norm, relative_decrease, radius; int accepted; };

struct SolveSummary {
  int termination, iterations, accepted, rejected, n_residuals;
  double initial_cost, final_cost;
  int jac_evals, res_evals;  // counts of full residual+Jacobian passes and residual-only passes
};

// ceres::Solve with TRUST_REGION / LEVENBERG_MARQUARDT / DENSE_QR (trust_region_minimizer.cc).
SolveSummary lm_solve(const Problem& p, double x[7], const Options& o, IterRecord* trace, int trace_cap, int* trace_n) {
  SolveSummary S{};
  const int n = p.n_res();
  S.n_residuals = n;
  std::vector<double> r, J, Js, A, rhs;
  double cost;
  int tn = 0;
  auto push = [&](const IterRecord& rec) { if (trace && tn < trace_cap) trace[tn] = rec; ++tn; };
  auto finish = [&](int term) { S.termination = term; S.final_cost = cost; if (trace_n) *trace_n = std::min(tn, trace_cap); return S; };
  // IterationZero
  cost = 0;
  if (!eval_problem(p, x, o, &cost, &r, &J)) { S.initial_cost = 0; return finish(TERM_FAILURE_EVAL_X0); }
  S.jac_evals++;
  S.initial_cost = cost;
  double g[6], scale[6];
  auto gradient = [&]() { for (int j = 0; j < 6; ++j) g[j] = 0; for (int i = 0; i < n; ++i) for (int j = 0; j < 6; ++j) g[j] += J[size_t(i) * 6 + j] * r[i]; };
  auto grad_max_norm = [&]() {
    double ng[6], xp[7]; for (int j = 0; j < 6; ++j) ng[j] = -g[j];
    pose_plus(x, ng, xp);
    double m = 0; for (int i = 0; i < 7; ++i) m = std::max(m, std::fabs(x[i] - xp[i]));
    return m;
  };
  gradient();
  for (int j = 0; j < 6; ++j) {
    double s = 0; for (int i = 0; i < n; ++i) s += J[size_t(i) * 6 + j] * J[size_t(i) * 6 + j];
    scale[j] = o.jacobi_scaling ? 1.0 / (1.0 + std::sqrt(s)) : 1.0;
  }
  auto scale_J = [&]() { Js = J; for (int i = 0; i < n; ++i) for (int j = 0; j < 6; ++j) Js[size_t(i) * 6 + j] *= scale[j]; };
  scale_J();
  double gmax = grad_max_norm();
  push({cost, 0, gmax, 0, 0, o.initial_radius, 1});
  if (gmax <= o.gradient_tolerance) return finish(TERM_CONVERGENCE_GRADIENT);

  double radius = o.initial_radius, decrease_factor = 2.0, diag[6];
  bool reuse_diagonal = false;
  bool dl_reuse = false; double dl_mu = 1e-8, dl_alpha = 0, dl_step_norm = 0, dl_grad[6], dl_gn[6];   // DoglegStrategy state
  int invalid_run = 0, iter = 0;
  for (;;) {
    if (iter >= o.max_num_iterations) return finish(TERM_NO_CONVERGENCE);
    if (radius < o.min_radius) return finish(TERM_CONVERGENCE_MIN_RADIUS);
    ++iter;
    S.iterations = iter;
    double y[6], step[6];
    bool ok;
    if (o.strategy == 0) {
      // LevenbergMarquardtStrategy::ComputeStep
      if (!reuse_diagonal)
        for (int j = 0; j < 6; ++j) {
          double s = 0; for (int i = 0; i < n; ++i) s += Js[size_t(i) * 6 + j] * Js[size_t(i) * 6 + j];
          diag[j] = std::min(std::max(s, o.min_lm_diagonal), o.max_lm_diagonal);
        }
      A.assign(size_t(n + 6) * 6, 0.0);
      rhs.assign(n + 6, 0.0);
      std::copy(Js.begin(), Js.end(), A.begin());
      std::copy(r.begin(), r.end(), rhs.begin());
      for (int j = 0; j < 6; ++j) A[size_t(n + j) * 6 + j] = std::sqrt(diag[j] / radius);
      ok = householder_lstsq6(A, rhs, n + 6, y);
      reuse_diagonal = true;
      if (ok) for (int j = 0; j < 6; ++j) step[j] = -y[j];
    } else {
      // DoglegStrategy::ComputeStep, TRADITIONAL_DOGLEG (ceres/dogleg_strategy.cc)
      ok = true;
      if (!dl_reuse) {
        dl_reuse = true;
        for (int j = 0; j < 6; ++j) {
          double s = 0; for (int i = 0; i < n; ++i) s += Js[size_t(i) * 6 + j] * Js[size_t(i) * 6 + j];
          diag[j] = std::sqrt(std::min(std::max(s, o.min_lm_diagonal), o.max_lm_diagonal));
        }
        // gradient_ = (J^T r) / D ; Cauchy step length alpha = |g|^2 / |J (g / D)|^2
        for (int j = 0; j < 6; ++j) { double gj = 0; for (int i = 0; i < n; ++i) gj += Js[size_t(i) * 6 + j] * r[i]; dl_grad[j] = gj / diag[j]; }
        double g2 = 0, Jg2 = 0;
        for (int j = 0; j < 6; ++j) g2 += dl_grad[j] * dl_grad[j];
        for (int i = 0; i < n; ++i) { double m = 0; for (int j = 0; j < 6; ++j) m += Js[size_t(i) * 6 + j] * (dl_grad[j] / diag[j]); Jg2 += m * m; }
        dl_alpha = g2 / Jg2;
        // Gauss-Newton step with a small regularisation mu * D^2, raised until the solve succeeds
        ok = false;
        while (dl_mu < 1.0) {
          A.assign(size_t(n + 6) * 6, 0.0); rhs.assign(n + 6, 0.0);
          std::copy(Js.begin(), Js.end(), A.begin()); std::copy(r.begin(), r.end(), rhs.begin());
          for (int j = 0; j < 6; ++j) A[size_t(n + j) * 6 + j] = diag[j] * std::sqrt(dl_mu);
          if (householder_lstsq6(A, rhs, n + 6, y)) { ok = true; break; }
          dl_mu *= 10.0;
        }
        if (ok) for (int j = 0; j < 6; ++j) dl_gn[j] = -y[j] * diag[j];
      }
      if (ok) {
        double gn_norm = 0, g_norm = 0, gdotgn = 0;
        for (int j = 0; j < 6; ++j) { gn_norm += dl_gn[j] * dl_gn[j]; g_norm += dl_grad[j] * dl_grad[j]; gdotgn += dl_grad[j] * dl_gn[j]; }
        gn_norm = std::sqrt(gn_norm); g_norm = std::sqrt(g_norm);
        double ds[6];
        if (gn_norm <= radius) { for (int j = 0; j < 6; ++j) ds[j] = dl_gn[j]; dl_step_norm = gn_norm; }
        else if (g_norm * dl_alpha >= radius) { for (int j = 0; j < 6; ++j) ds[j] = -(radius / g_norm) * dl_grad[j]; dl_step_norm = radius; }
        else {
          const double b_dot_a = -dl_alpha * gdotgn;
          const double a2 = std::pow(dl_alpha * g_norm, 2.0);
          const double bma2 = a2 - 2 * b_dot_a + gn_norm * gn_norm;
          const double c = b_dot_a - a2;
          const double d = std::sqrt(c * c + bma2 * (radius * radius - a2));
          const double beta = (c <= 0) ? (d - c) / bma2 : (radius * radius - a2) / (d + c);
          double nn = 0;
          for (int j = 0; j < 6; ++j) { ds[j] = (-dl_alpha * (1.0 - beta)) * dl_grad[j] + beta * dl_gn[j]; nn += ds[j] * ds[j]; }
          dl_step_norm = std::sqrt(nn);
        }
        for (int j = 0; j < 6; ++j) step[j] = ds[j] / diag[j];
      }
    }
    double model_cost_change = 0;
    if (ok) {
      // model_cost_change = -(J step)^T (r + J step / 2)
      for (int i = 0; i < n; ++i) {
        double m = 0; for (int j = 0; j < 6; ++j) m += Js[size_t(i) * 6 + j] * step[j];
        model_cost_change -= m * (r[i] + 0.5 * m);
      }
    }
    if (!ok || !(model_cost_change > 0.0)) {  // HandleInvalidStep
      if (++invalid_run >= o.max_consecutive_invalid) return finish(TERM_FAILURE_INVALID_STEPS);
      if (o.strategy == 0) { radius *= 0.5; reuse_diagonal = false; }
      else { dl_mu *= 10.0; dl_reuse = false; }          // DoglegStrategy::StepIsInvalid
      push({cost, 0, gmax, 0, 0, radius, 0});
      S.rejected++;
      continue;
    }
    invalid_run = 0;
    double delta[6], xc[7], cand_cost;
    for (int j = 0; j < 6; ++j) delta[j] = step[j] * scale[j];
    pose_plus(x, delta, xc);
    bool evok = eval_problem(p, xc, o, &cand_cost, nullptr, nullptr);
    S.res_evals++;
    if (!evok) cand_cost = std::numeric_limits<double>::max();
    // ParameterToleranceReached
    double sn = 0, xn = 0;
    for (int i = 0; i < 7; ++i) { sn += (x[i] - xc[i]) * (x[i] - xc[i]); xn += x[i] * x[i]; }
    sn = std::sqrt(sn); xn = std::sqrt(xn);
    if (sn <= o.parameter_tolerance * (xn + o.parameter_tolerance)) return finish(TERM_CONVERGENCE_PARAMETER);
    // FunctionToleranceReached (before acceptance: x is NOT updated)
    const double cost_change = cost - cand_cost;
    if (std::fabs(cost_change) <= o.function_tolerance * cost) {
      push({cost, cost_change, gmax, sn, 0, radius, 0});
      return finish(TERM_CONVERGENCE_FUNCTION);
    }
    const double rel = evok ? cost_change / model_cost_change : -std::numeric_limits<double>::max();
    if (rel > o.min_relative_decrease) {  // HandleSuccessfulStep
      for (int i = 0; i < 7; ++i) x[i] = xc[i];
      eval_problem(p, x, o, &cost, &r, &J);
      S.jac_evals++;
      gradient(); scale_J();
      gmax = grad_max_norm();
      if (o.strategy == 0) {
        radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rel - 1.0, 3));
        radius = std::min(o.max_radius, radius);
        decrease_factor = 2.0; reuse_diagonal = false;
      } else {                                             // DoglegStrategy::StepAccepted
        if (rel < 0.25) radius *= 0.5;
        if (rel > 0.75) radius = std::max(radius, 3.0 * dl_step_norm);
        dl_mu = std::max(1e-8, 2.0 * dl_mu / 10.0); dl_reuse = false;
      }
      S.accepted++;
      push({cost, cost_change, gmax, sn, rel, radius, 1});
      if (gmax <= o.gradient_tolerance) return finish(TERM_CONVERGENCE_GRADIENT);
    } else {  // HandleUnsuccessfulStep
      if (o.strategy == 0) { radius = radius / decrease_factor; decrease_factor *= 2.0; reuse_diagonal = true; }
      else { radius *= 0.5; dl_reuse = true; }           // DoglegStrategy::StepRejected
      S.rejected++;
      push({cost, cost_change, gmax, sn, rel, radius, 0});
    }
  }
}

struct Level {
  int w, h; double fx, fy, cx, cy;
  std::vector<double> pts;  // xyz
  std::vector<float> dt;
};

void build_ref_points(const uint8_t* bgr, const uint16_t* depth, int w, int h, double fx, double fy, double cx, double cy,
                      double zscale, int thresh, std::vector<double>& pts, std::vector<int>* uvd) {
  // standalone/utils.cpp:201-281 get_aX
  std::vector<uint8_t> lap(size_t(w) * h);
  laplacian_edge_strength(bgr, w, h, lap.data());
  pts.clear();
  for (int v = 0; v < h; ++v)
    for (int u = 0; u < w; ++u) {
      double Z = double(depth[size_t(v) * w + u]) / zscale;
      if (lap[size_t(v) * w + u] > thresh && Z > 0) {
        pts.push_back((u - cx) * Z / fx);
        pts.push_back((v - cy) * Z / fy);
        pts.push_back(Z);
        if (uvd) { uvd->push_back(u); uvd->push_back(v); uvd->push_back(depth[size_t(v) * w + u]); }
      }
    }
}

void build_now_dt(const uint8_t* bgr, int w, int h, int thresh, int use_median, int norm_mode, std::vector<float>& dt,
                  uint8_t* mask_out) {
  // standalone/utils.cpp:38-83 get_distance_transform
  std::vector<uint8_t> lap(size_t(w) * h), B(size_t(w) * h), Bf(size_t(w) * h);
  laplacian_edge_strength(bgr, w, h, lap.data());
  for (size_t i = 0; i < B.size(); ++i) B[i] = lap[i] > thresh ? 0 : 255;
  if (use_median) median3_u8(B.data(), w, h, Bf.data()); else Bf = B;
  if (mask_out) std::memcpy(mask_out, Bf.data(), Bf.size());
  dt.resize(size_t(w) * h);
  chamfer3_dt(Bf.data(), w, h, dt.data());
  if (norm_mode == 1) normalize_minmax(dt.data(), dt.size(), 0.0, 1.0);
  else if (norm_mode == 2) normalize_minmax(dt.data(), dt.size(), 0.0, 255.0);
}

}  // namespace

// =============================================================================
// C ABI for ctypes (tests / bench cpu_baseline only)
// =============================================================================
extern "C" {

struct eo_options {
  int max_num_iterations; double function_tolerance, gradient_tolerance, parameter_tolerance;
  double initial_radius, max_radius, min_radius, min_relative_decrease, min_lm_diagonal, max_lm_diagonal;
  int jacobi_scaling, max_consecutive_invalid, loss_type; double loss_scale;
  int strategy, pad;
};
static Options to_opts(const eo_options* e) {
  Options o;
  if (!e) return o;
  o.max_num_iterations = e->max_num_iterations; o.function_tolerance = e->function_tolerance;
  o.gradient_tolerance = e->gradient_tolerance; o.parameter_tolerance = e->parameter_tolerance;
  o.initial_radius = e->initial_radius; o.max_radius = e->max_radius; o.min_radius = e->min_radius;
  o.min_relative_decrease = e->min_relative_decrease; o.min_lm_diagonal = e->min_lm_diagonal;
  o.max_lm_diagonal = e->max_lm_diagonal; o.jacobi_scaling = e->jacobi_scaling;
  o.max_consecutive_invalid = e->max_consecutive_invalid; o.loss_type = e->loss_type; o.loss_scale = e->loss_scale;
  o.strategy = e->strategy;
  return o;
}
void eo_options_default(eo_options* e) {
  Options o;
  e->max_num_iterations = o.max_num_iterations; e->function_tolerance = o.function_tolerance;
  e->gradient_tolerance = o.gradient_tolerance; e->parameter_tolerance = o.parameter_tolerance;
  e->initial_radius = o.initial_radius; e->max_radius = o.max_radius; e->min_radius = o.min_radius;
  e->min_relative_decrease = o.min_relative_decrease; e->min_lm_diagonal = o.min_lm_diagonal;
  e->max_lm_diagonal = o.max_lm_diagonal; e->jacobi_scaling = o.jacobi_scaling;
  e->max_consecutive_invalid = o.max_consecutive_invalid; e->loss_type = o.loss_type; e->loss_scale = o.loss_scale;
  e->strategy = o.strategy; e->pad = 0;
}

// ---- image stages (exposed one by one so each can be pinned against cv2) ----
void eo_gaussian3_u8c3(const uint8_t* s, int w, int h, uint8_t* d) { gaussian3_u8c3(s, w, h, d); }
void eo_box3_u8c3(const uint8_t* s, int w, int h, uint8_t* d) { box3_u8c3(s, w, h, d); }
void eo_rgb2gray(const uint8_t* s, int w, int h, uint8_t* d) { rgb2gray_u8(s, w, h, d); }
void eo_laplacian3_abs(const uint8_t* s, int w, int h, uint8_t* d) { laplacian3_abs_u8(s, w, h, d); }
void eo_median3(const uint8_t* s, int w, int h, uint8_t* d) { median3_u8(s, w, h, d); }
void eo_chamfer3_dt(const uint8_t* s, int w, int h, float* d) { chamfer3_dt(s, w, h, d); }
void eo_normalize_minmax(float* d, int n, double a, double b) { normalize_minmax(d, size_t(n), a, b); }
void eo_half_linear_u8c3(const uint8_t* s, int w, int h, uint8_t* d) { half_linear_u8c3(s, w, h, d); }
void eo_half_nearest_u16(const uint16_t* s, int w, int h, uint16_t* d) { half_nearest_u16(s, w, h, d); }

void eo_canny(const uint8_t* s, int w, int h, int cn, double t1, double t2, int l2, uint8_t* d) { canny_u8(s, w, h, cn, t1, t2, l2 != 0, d); }
void eo_exact_edt(const uint8_t* s, int w, int h, float* d) { exact_edt(s, w, h, d); }
// get_aX (utils.cpp:201-281): returns N; fills xyz (3N doubles) and uvd (3N ints) up to cap points.
int eo_get_aX(const uint8_t* bgr, const uint16_t* depth, int w, int h, double fx, double fy, double cx, double cy,
              double zscale, int thresh, double* xyz, int* uvd, int cap) {
  std::vector<double> pts; std::vector<int> q;
  build_ref_points(bgr, depth, w, h, fx, fy, cx, cy, zscale, thresh, pts, &q);
  int n = int(pts.size() / 3);
  int m = std::min(n, cap);
  if (xyz) std::memcpy(xyz, pts.data(), size_t(m) * 3 * sizeof(double));
  if (uvd) std::memcpy(uvd, q.data(), size_t(m) * 3 * sizeof(int));
  return n;
}
// get_distance_transform (utils.cpp:38-83). norm_mode: 0 none, 1 [0,1], 2 [0,255]. mask optional (post-median, 0=edge).
void eo_get_distance_transform(const uint8_t* bgr, int w, int h, int thresh, int use_median, int norm_mode, float* dt,
                               uint8_t* mask) {
  std::vector<float> d;
  build_now_dt(bgr, w, h, thresh, use_median, norm_mode, d, mask);
  std::memcpy(dt, d.data(), d.size() * sizeof(float));
}

// ---- single evaluation -------------------------------------------------------
// pts: xyz triplets (n_total), stride selects every stride-th (SEA:267).
// Outputs (each optional): residuals[n] (robustified), raw[n] (plain DT lookups),
// J[n*6] (local, robustified), sums[28] = {cost, b[6]=J^T r, H upper-tri 21 row-major}.
// Returns 0 ok, 1 if any functor returned false.
int eo_eval(const double* pts, int n_total, int stride, const float* dt, int w, int h, double fx, double fy, double cx,
            double cy, const double* pose7, const eo_options* eo, double* residuals, double* raw, double* J,
            double* sums) {
  Options o = to_opts(eo);
  Problem p(pts, n_total, stride, fx, fy, cx, cy, DtGrid{dt, w, h});
  const int n = p.n_res();
  double cost = 0, b[6] = {0}, H[21] = {0};
  for (int i = 0; i < n; ++i) {
    double ci, ri, Ji[6], rr;
    if (!eval_block(p, i, pose7, o, &ci, &ri, Ji, &rr)) return 1;
    cost += ci;
    if (residuals) residuals[i] = ri;
    if (raw) raw[i] = rr;
    if (J) std::memcpy(J + size_t(i) * 6, Ji, sizeof(Ji));
    int k = 0;
    for (int a = 0; a < 6; ++a) { b[a] += Ji[a] * ri; for (int c = a; c < 6; ++c) H[k++] += Ji[a] * Ji[c]; }
  }
  if (sums) { sums[0] = cost; std::memcpy(sums + 1, b, sizeof(b)); std::memcpy(sums + 7, H, sizeof(H)); }
  return 0;
}

// Residual-only evaluation through the plain-double functor path (what Ceres uses for
// candidate cost): residuals raw (unrobustified).  Also returns dfdu, dfdv per point.
int eo_eval_raw(const double* pts, int n_total, int stride, const float* dt, int w, int h, double fx, double fy,
                double cx, double cy, const double* pose7, double* raw, double* uv) {
  Problem p(pts, n_total, stride, fx, fy, cx, cy, DtGrid{dt, w, h});
  const int n = p.n_res();
  for (int i = 0; i < n; ++i) {
    const double* P3 = pts + size_t(i) * stride * 3;
    EAResidue f{fx, fy, cx, cy, P3[0], P3[1], P3[2], &p.views[0].grid};
    double r;
    if (!f(pose7, pose7 + 4, &r)) return 1;
    raw[i] = r;
    if (uv) {
      // recompute u,v for diagnostics (same arithmetic as the functor)
      Jet<1> dummy; (void)dummy;
      const double qw = pose7[0], qx = pose7[1], qy = pose7[2], qz = pose7[3];
      const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz, twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx,
                   txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
      double bx = (1 - (tyy + tzz)) * P3[0] + (txy - twz) * P3[1] + (txz + twy) * P3[2] + pose7[4];
      double by = (txy + twz) * P3[0] + (1 - (txx + tzz)) * P3[1] + (tyz - twx) * P3[2] + pose7[5];
      double bz = (txz - twy) * P3[0] + (tyz + twx) * P3[1] + (1 - (txx + tyy)) * P3[2] + pose7[6];
      uv[2 * i] = fx * bx / bz + cx; uv[2 * i + 1] = fy * by / bz + cy;
    }
  }
  return 0;
}

struct eo_summary {
  int termination, iterations, accepted, rejected, n_residuals, jac_evals, res_evals, pad;
  double initial_cost, final_cost;
};
static void fill(eo_summary* s, const SolveSummary& S) {
  if (!s) return;
  s->termination = S.termination; s->iterations = S.iterations; s->accepted = S.accepted; s->rejected = S.rejected;
  s->n_residuals = S.n_residuals; s->jac_evals = S.jac_evals; s->res_evals = S.res_evals; s->pad = 0;
  s->initial_cost = S.initial_cost; s->final_cost = S.final_cost;
}

// ceres::Solve restatement on prepared inputs.  pose7 in/out (q wxyz, t).  trace: 7 doubles per record
// {cost, cost_change, gradient_max_norm, step_norm, relative_decrease, radius, accepted}.
int eo_solve(const double* pts, int n_total, int stride, const float* dt, int w, int h, double fx, double fy, double cx,
             double cy, double* pose7, const eo_options* eo, eo_summary* summary, double* trace, int trace_cap,
             int* trace_n) {
  Options o = to_opts(eo);
  Problem p(pts, n_total, stride, fx, fy, cx, cy, DtGrid{dt, w, h});
  std::vector<IterRecord> tr(trace_cap > 0 ? trace_cap : 0);
  int tn = 0;
  SolveSummary S = lm_solve(p, pose7, o, tr.data(), trace_cap, &tn);
  for (int i = 0; i < tn && trace; ++i) {
    double* d = trace + size_t(i) * 7;
    d[0] = tr[i].cost; d[1] = tr[i].cost_change; d[2] = tr[i].gradient_max_norm; d[3] = tr[i].step_norm;
    d[4] = tr[i].relative_decrease; d[5] = tr[i].radius; d[6] = tr[i].accepted;
  }
  if (trace_n) *trace_n = tn;
  fill(summary, S);
  return 0;
}

// ---- multi-camera problems: residual variants of standalone/utils.h:101-421 -----------------------------------
// One eo_view per camera; all views constrain the same pose (standalone_edge_align.cpp:791-803, 3204-3219).
struct eo_view {
  const double* pts; int n_total, stride; const float* dt; int w, h;
  double fx, fy, cx, cy;
  int use_dist; double dist[5];       // k1,k2,p1,p2,k3
  int use_rig; double T21[16], T12[16];  // trans_1to2, trans_1to2_inv (row-major 4x4)
};
static Problem views_problem(const eo_view* v, int n_views) {
  Problem p;
  for (int i = 0; i < n_views; ++i) {
    View V; V.pts = v[i].pts; V.n_total = v[i].n_total; V.stride = v[i].stride; V.fx = v[i].fx; V.fy = v[i].fy; V.cx = v[i].cx; V.cy = v[i].cy;
    V.grid = DtGrid{v[i].dt, v[i].w, v[i].h};
    V.use_dist = v[i].use_dist != 0; for (int k = 0; k < 5; ++k) V.dist[k] = v[i].dist[k];
    V.use_rig = v[i].use_rig != 0; for (int k = 0; k < 16; ++k) { V.T21[k] = v[i].T21[k]; V.T12[k] = v[i].T12[k]; }
    p.views.push_back(V);
  }
  return p;
}
int eo_eval_views(const eo_view* views, int n_views, const double* pose7, const eo_options* eo, double* residuals, double* raw,
                  double* J, double* sums) {
  Options o = to_opts(eo);
  Problem p = views_problem(views, n_views);
  const int n = p.n_res();
  double cost = 0, b[6] = {0}, H[21] = {0};
  for (int i = 0; i < n; ++i) {
    double ci, ri, Ji[6], rr;
    if (!eval_block(p, i, pose7, o, &ci, &ri, Ji, &rr)) return 1;
    cost += ci;
    if (residuals) residuals[i] = ri;
    if (raw) raw[i] = rr;
    if (J) std::memcpy(J + size_t(i) * 6, Ji, sizeof(Ji));
    int k = 0;
    for (int a = 0; a < 6; ++a) { b[a] += Ji[a] * ri; for (int c = a; c < 6; ++c) H[k++] += Ji[a] * Ji[c]; }
  }
  if (sums) { sums[0] = cost; std::memcpy(sums + 1, b, sizeof(b)); std::memcpy(sums + 7, H, sizeof(H)); }
  return 0;
}
int eo_solve_views(const eo_view* views, int n_views, double* pose7, const eo_options* eo, eo_summary* summary) {
  Options o = to_opts(eo);
  Problem p = views_problem(views, n_views);
  SolveSummary S = lm_solve(p, pose7, o, nullptr, 0, nullptr);
  fill(summary, S);
  return 0;
}

// ---- full pair: preprocess both frames (pyramid) + coarse-to-fine solve -------
// Pyramid semantics are an EXTENSION (the reference is single level; DESIGN.md "Pyramid"):
// level l image = l successive cv::resize(0.5) of BGR (INTER_LINEAR) and depth (INTER_NEAREST);
// K_l: f/2^l, c_l=(c+0.5)/2^l-0.5; each level runs the same LM from the coarser level's pose.
struct eo_pair_cfg {
  int w, h, n_levels, stride, thresh, use_median, norm_mode, pad;
  double fx, fy, cx, cy, zscale;
};
static void build_levels(const uint8_t* bgr, const uint16_t* depth, const eo_pair_cfg& c, bool want_pts, bool want_dt,
                         std::vector<Level>& L) {
  L.resize(c.n_levels);
  std::vector<uint8_t> img(bgr, bgr + size_t(c.w) * c.h * 3), tmp;
  std::vector<uint16_t> dep, dtmp;
  if (depth) dep.assign(depth, depth + size_t(c.w) * c.h);
  int w = c.w, h = c.h;
  for (int l = 0; l < c.n_levels; ++l) {
    const double s = 1.0 / double(1 << l);
    L[l].w = w; L[l].h = h; L[l].fx = c.fx * s; L[l].fy = c.fy * s;
    L[l].cx = (c.cx + 0.5) * s - 0.5; L[l].cy = (c.cy + 0.5) * s - 0.5;
    if (want_pts && depth) build_ref_points(img.data(), dep.data(), w, h, L[l].fx, L[l].fy, L[l].cx, L[l].cy, c.zscale, c.thresh, L[l].pts, nullptr);
    if (want_dt) build_now_dt(img.data(), w, h, c.thresh, c.use_median, c.norm_mode, L[l].dt, nullptr);
    if (l + 1 < c.n_levels) {
      tmp.resize(size_t(w / 2) * (h / 2) * 3);
      half_linear_u8c3(img.data(), w, h, tmp.data());
      img.swap(tmp);
      if (depth) { dtmp.resize(size_t(w / 2) * (h / 2)); half_nearest_u16(dep.data(), w, h, dtmp.data()); dep.swap(dtmp); }
      w /= 2; h /= 2;
    }
  }
}
static void solve_levels(const std::vector<Level>& R, const std::vector<Level>& N, const eo_pair_cfg& c, const Options& o,
                         double* pose7, eo_summary* summaries) {
  for (int l = c.n_levels - 1; l >= 0; --l) {
    Problem p(R[l].pts.data(), int(R[l].pts.size() / 3), c.stride, N[l].fx, N[l].fy, N[l].cx, N[l].cy,
              DtGrid{N[l].dt.data(), N[l].w, N[l].h});
    SolveSummary S{};
    if (p.n_res() > 0) S = lm_solve(p, pose7, o, nullptr, 0, nullptr);
    if (summaries) fill(summaries + l, S);
  }
}
int eo_align_pair(const uint8_t* ref_bgr, const uint16_t* ref_depth, const uint8_t* now_bgr, const eo_pair_cfg* cfg,
                  const eo_options* eo, double* pose7, eo_summary* summaries /*n_levels*/) {
  Options o = to_opts(eo);
  std::vector<Level> R, N;
  build_levels(ref_bgr, ref_depth, *cfg, true, false, R);
  build_levels(now_bgr, nullptr, *cfg, false, true, N);
  solve_levels(R, N, *cfg, o, pose7, summaries);
  return 0;
}

// ---- multi-threaded batch of independent pairs (CPU baseline timing) ----------
// Frames are contiguous: bgr[n][h][w][3], depth[n][h][w].  Pair i aligns ref_idx[i] -> now_idx[i]
// from poses[i*7..] (in/out).  Returns wall seconds for the whole batch.  include_preprocess: 1 = each pair
// preprocesses both of its frames (pair-total), 0 = solve-only timing (preprocessing done before the clock).
double eo_align_batch(const uint8_t* bgr, const uint16_t* depth, int n_pairs, const int* ref_idx, const int* now_idx,
                      const eo_pair_cfg* cfg, const eo_options* eo, double* poses, eo_summary* summaries, int n_threads,
                      int include_preprocess) {
  Options o = to_opts(eo);
  const size_t fsz = size_t(cfg->w) * cfg->h;
  std::vector<std::vector<Level>> R(n_pairs), N(n_pairs);
  auto prep = [&](int i) {
    build_levels(bgr + fsz * 3 * ref_idx[i], depth + fsz * ref_idx[i], *cfg, true, false, R[i]);
    build_levels(bgr + fsz * 3 * now_idx[i], nullptr, *cfg, false, true, N[i]);
  };
  auto run = [&](bool do_prep, bool do_solve) {
    std::atomic<int> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < std::max(1, n_threads); ++t)
      th.emplace_back([&]() {
        for (int i; (i = next.fetch_add(1)) < n_pairs;) {
          if (do_prep) prep(i);
          if (do_solve) solve_levels(R[i], N[i], *cfg, o, poses + size_t(i) * 7, summaries ? summaries + size_t(i) * cfg->n_levels : nullptr);
        }
      });
    for (auto& t : th) t.join();
  };
  if (!include_preprocess) run(true, false);
  auto t0 = std::chrono::steady_clock::now();
  if (include_preprocess) run(true, true); else run(false, true);
  auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

int eo_hardware_threads() { return int(std::thread::hardware_concurrency()); }

// pose helpers (standalone/PoseManipUtils.cpp:3-27 semantics, Eigen conventions)
void eo_quat_to_matrix(const double* pose7, double* T16 /*row-major 4x4*/) {
  const double qw = pose7[0], qx = pose7[1], qy = pose7[2], qz = pose7[3];
  const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz, twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx,
               txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  double R[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T16[i * 4 + j] = R[i * 3 + j]; T16[i * 4 + 3] = pose7[4 + i]; }
  T16[12] = T16[13] = T16[14] = 0; T16[15] = 1;
}
void eo_quat_plus(const double* x7, const double* d6, double* out7) { pose_plus(x7, d6, out7); }

}  // extern "C"
