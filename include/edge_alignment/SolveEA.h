// class SolveEA -- drop-in for the reference facade (include/SolveEA.h:36-71, src/SolveEA.cpp): same method names,
// argument meaning and call order (src/ea.cpp:184-199):
//     SolveEA* ea = new SolveEA();  ea->setRefFrame(ref_im, ref_depth);  ea->setNowFrame(now_im, now_depth);
//     ea->setAsCERESProblem();
// Everything behind it runs on the GPU through the C ABI (ea_cabi.h).  Additive accessors return what the reference
// only printed (SolveEA.cpp:204-213): getPose(), getSummary(), setK().
//
// Defaults = the literals of src/SolveEA.cpp: half-TUM intrinsics (:12-23), cv::Canny(rgb, 150, 100, 3, true) on the
// colour image (:46,102), exact Euclidean distance transform of the inverted edge map (:108) normalised to [0,255]
// (:109), every edge point (:163), NULL loss (:171), 25 iterations (:185), depth in metres as CV_32F with Z==0 -> 1.0
// (:68-69), start at identity (:130-131), TRUST_REGION with traditional DOGLEG (:191-192).  One repair of a reference
// defect: the distance field is sampled as one channel with the standalone's (row == u, col == v)
// convention (the reference wraps it in a 2-channel grid, :152).  useStandalonePipeline() switches to the edge detector /
// DT / loss of standalone/utils.cpp + edge_align_test1.
#pragma once
#include <cmath>
#include <cstdio>

#include "Frame.h"

class SolveEA {
 public:
  SolveEA() {
    ea_frame_params_default(&fp_);
    fp_.width = 320; fp_.height = 240;                       // src/ea.cpp:38 feeds half-resolution frames
    fp_.fx = .5 * 525.0; fp_.fy = .5 * 525.0; fp_.cx = .5 * 319.5; fp_.cy = .5 * 239.5;   // SolveEA.cpp:15-18
    fp_.depth_scale = 5000.0; fp_.dt_normalize = EA_NORM_255;                             // SolveEA.cpp:109
    fp_.edge_detector = EA_EDGE_CANNY_COLOR; fp_.canny_low = 150.0; fp_.canny_high = 100.0; fp_.canny_l2 = 1;   // :46
    fp_.dt_kind = EA_DT_EXACT;                                                            // SolveEA.cpp:108
    fp_.max_points = 0;
    fp_.depth_type = EA_DEPTH_F32; fp_.zero_depth_to_one = 1;                             // SolveEA.cpp:27,68-69
    ea_solve_params_default(&sp_);
    sp_.point_stride = 1; sp_.loss_type = EA_LOSS_TRIVIAL; sp_.max_num_iterations = 25;    // SolveEA.cpp:163,171,185
    sp_.trust_region_strategy = EA_STRATEGY_DOGLEG;                                                      // DOGLEG, SolveEA.cpp:192
    init_pose_[0] = 1; for (int i = 1; i < 7; ++i) init_pose_[i] = 0;                     // SolveEA.cpp:130-131
    for (int i = 0; i < 7; ++i) pose_[i] = init_pose_[i];
  }

  // "TODO : Write a function to set K" (SolveEA.cpp:12)
  void setK(double fx, double fy, double cx, double cy) { fp_.fx = fx; fp_.fy = fy; fp_.cx = cx; fp_.cy = cy; dirty_ = true; }
  void setImageSize(int width, int height) { fp_.width = width; fp_.height = height; dirty_ = true; }
  ea_frame_params& frameParams() { dirty_ = true; return fp_; }
  ea_solve_params& solveParams() { return sp_; }
  // SolveEA.cpp:69 keeps edge pixels without depth by placing them at Z = 1.0 (default); false = drop them like
  // the standalone get_aX does (utils.cpp:258 "Z > 0")
  void setZeroDepthToOne(bool on) { fp_.zero_depth_to_one = on ? 1 : 0; dirty_ = true; }
  // the working pipeline of standalone/: Laplacian>35 edges (+median for the DT), chamfer-3 DT in [0,1], Cauchy(1),
  // every 30th point, 50 iterations, edge pixels without depth dropped  (utils.cpp:38-83,201-281; SEA:256-286)
  void useStandalonePipeline() {
    fp_.edge_detector = EA_EDGE_LAPLACIAN; fp_.dt_kind = EA_DT_CHAMFER3; fp_.dt_normalize = EA_NORM_01; fp_.use_median = 1;
    sp_.point_stride = 30; sp_.loss_type = EA_LOSS_CAUCHY; sp_.loss_scale = 1.0; sp_.max_num_iterations = 50;
    sp_.trust_region_strategy = EA_STRATEGY_LM;
    fp_.zero_depth_to_one = 0; fp_.depth_type = EA_DEPTH_U16; dirty_ = true;
  }

  template <class MatT> void setRefFrame(const MatT& rgb, const MatT& depth) {           // SolveEA.cpp:29-82
    ensure(rgb);
    ref_.set(rgb, &depth, EA_ROLE_REF);
    have_ref_ = true;
  }
  template <class MatT> void setNowFrame(const MatT& rgb, const MatT& /*depth: stored but unused, SolveEA.cpp:86-119*/) {
    ensure(rgb);
    now_.set(rgb, static_cast<const MatT*>(nullptr), EA_ROLE_NOW);
    have_now_ = true;
  }

  // Builds and solves the problem (SolveEA.cpp:124-216).  EVERY call starts from q=(1,0,0,0), t=0 like the reference
  // (SolveEA.cpp:130-131) -- or from the pose given to setInitialPose -- never from the previous call's result, unless
  // setWarmStart(true) asked for that.  Throws if a frame is missing (the reference left that unchecked, SolveEA.cpp:122).
  void setAsCERESProblem() {
    if (!have_ref_ || !have_now_) throw std::runtime_error("SolveEA::setAsCERESProblem: call setRefFrame and setNowFrame first");
    if (!warm_start_ || !solved_once_) for (int i = 0; i < 7; ++i) pose_[i] = init_pose_[i];
    solved_once_ = true;
    const int32_t zero = 0;
    summaries_.assign(size_t(fp_.n_levels), ea_summary{});
    ea::check(ea_solve_batch(ref_.context(), 1, ref_.handle(), &zero, now_.handle(), &zero, pose_, &sp_, summaries_.data()), "ea_solve_batch");
  }

  void setInitialPose(const double q_wxyz[4], const double t[3]) {
    for (int i = 0; i < 4; ++i) init_pose_[i] = q_wxyz[i];
    for (int i = 0; i < 3; ++i) init_pose_[4 + i] = t[i];
    solved_once_ = false;
  }
  // opt-in (not in the reference): later setAsCERESProblem calls start from the previous result (frame-to-frame tracking)
  void setWarmStart(bool on) { warm_start_ = on; }
  void getPose(double q_wxyz[4], double t[3]) const { for (int i = 0; i < 4; ++i) q_wxyz[i] = pose_[i]; for (int i = 0; i < 3; ++i) t[i] = pose_[4 + i]; }
  const std::vector<ea_summary>& getSummary() const { return summaries_; }

  // _verify3dPts (SolveEA.cpp:275-306) re-stated without the imshow: fraction of the reference edge points that
  // project inside the now image at identity.
  double _verify3dPts() const {
    std::vector<float> p = ref_.edgePoints(0);
    const size_t n = p.size() / 4;
    size_t inside = 0;
    for (size_t i = 0; i < n; ++i) inside += (p[4 * i] >= 0 && p[4 * i] < fp_.width && p[4 * i + 1] >= 0 && p[4 * i + 1] < fp_.height);
    return n ? double(inside) / double(n) : 0.0;
  }

  // _sampleCERESProblem (SolveEA.cpp:218-272): the reference solves a random 34x27 linear least-squares toy with
  // HuberLoss(0.1), 25 iterations, DENSE_QR "to ensure it works" -- a link check of the solver.  The equivalent check here
  // drives the production device path (padded distance field, point stream, fused evaluation, on-device LM) on a small
  // synthetic problem with a known answer: points on the edges of a rectangle, displaced by a known small pose, against that
  // rectangle's exact distance field, HuberLoss(0.1), 25 iterations.  Prints what the reference prints (initial / final cost,
  // problem sizes, brief report); throws if the solver does not reduce the cost by 1e3 and recover the pose.
  void _sampleCERESProblem() {
    ea_context* ctx = ea::default_context();
    const int W = 64, H = 48;
    ea_frame_params fp; ea_frame_params_default(&fp);
    fp.width = W; fp.height = H; fp.fx = 60.0; fp.fy = 60.0; fp.cx = 31.5; fp.cy = 23.5; fp.max_points = 512; fp.dt_normalize = EA_NORM_NONE;
    Frame toy(fp, ctx);
    std::vector<float> dt(size_t(W) * H);
    const int x0 = 14, x1 = 50, y0 = 10, y1 = 38;
    auto seg = [](double p, double a, double b) { return p < a ? a - p : (p > b ? p - b : 0.0); };
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        double d = 1e9;
        for (int k = 0; k < 2; ++k) {   // distance to the two vertical and two horizontal sides
          const double dx = std::abs(x - (k ? x1 : x0)), oy = seg(y, y0, y1);
          const double dy = std::abs(y - (k ? y1 : y0)), ox = seg(x, x0, x1);
          d = std::min(d, std::min(std::sqrt(dx * dx + oy * oy), std::sqrt(dy * dy + ox * ox)));
        }
        dt[size_t(y) * W + x] = float(d / 16.0);
      }
    ea::check(ea_frameset_set_dt(toy.handle(), 0, 0, dt.data()), "set_dt");
    // truth: a small rotation about z plus a translation; the stored points are the rectangle's edge points moved by its inverse
    const double th = 0.02, c = std::cos(th), s = std::sin(th), tx = 0.03, ty = -0.02, tz = 0.01;
    std::vector<float> p4;
    for (int i = 0; i < 96; ++i) {
      const int side = i % 4, k = i / 4;
      const double u = side < 2 ? (side ? x1 : x0) : x0 + 1.5 * k, v = side < 2 ? y0 + 1.1 * k : (side == 2 ? y0 : y1);
      const double Z = 1.5 + 0.01 * (i % 7);
      const double X = (u - fp.cx) * Z / fp.fx - tx, Y = (v - fp.cy) * Z / fp.fy - ty, Zc = Z - tz;   // p' - t
      const double a[3] = {c * X + s * Y, -s * X + c * Y, Zc};                                        // R^T (p' - t)
      p4.push_back(float(a[0])); p4.push_back(float(a[1])); p4.push_back(float(a[2])); p4.push_back(1.0f);
    }
    const int n = int(p4.size() / 4);
    ea::check(ea_frameset_set_points(toy.handle(), 0, 0, p4.data(), n, EA_POINTS_XYZ), "set_points");
    ea_solve_params sp; ea_solve_params_default(&sp);
    sp.point_stride = 1; sp.loss_type = EA_LOSS_HUBER; sp.loss_scale = 0.1; sp.max_num_iterations = 25;   // SolveEA.cpp:233,263
    double x[7] = {1, 0, 0, 0, 0, 0, 0};
    const int32_t zero = 0;
    ea_summary sm{};
    std::printf("init values of x_init=[1, 0, 0, 0, 0, 0, 0]';\n");
    ea::check(ea_solve_batch(ctx, 1, toy.handle(), &zero, toy.handle(), &zero, x, &sp, &sm), "ea_solve_batch");
    std::printf("cost : %g\nNumParameterBlocks 2\nNumParameters 7\nNumResidualBlocks %d\nNumResiduals %d\n", sm.initial_cost, sm.n_residuals, sm.n_residuals);
    std::printf("Ceres-equivalent device solver report: iterations %d, initial cost %.6e, final cost %.6e, termination %d\n",
                sm.iterations, sm.initial_cost, sm.final_cost, sm.termination);
    std::printf("final value of x_final=[%g, %g, %g, %g, %g, %g, %g]';\n", x[0], x[1], x[2], x[3], x[4], x[5], x[6]);
    sample_ok_ = sm.n_residuals == n && sm.final_cost < 1e-3 * sm.initial_cost && std::abs(x[3] - std::sin(th / 2)) < 2e-3 &&
                 std::abs(x[4] - tx) < 5e-3 && std::abs(x[5] - ty) < 5e-3;
    if (!sample_ok_) throw std::runtime_error("SolveEA::_sampleCERESProblem: the device solver did not solve the sample problem");
  }
  bool sampleProblemOk() const { return sample_ok_; }

  Frame& refFrame() { return ref_; }
  Frame& nowFrame() { return now_; }

 private:
  template <class MatT> void ensure(const MatT& rgb) {
    if (rgb.cols != fp_.width || rgb.rows != fp_.height) { fp_.width = rgb.cols; fp_.height = rgb.rows; dirty_ = true; }
    if (fp_.max_points <= 0 || dirty_) fp_.max_points = fp_.width * fp_.height;   // Z==0 -> 1 keeps every edge pixel
    if (dirty_ || !ref_.valid()) { ref_.init(fp_); now_.init(fp_); have_ref_ = have_now_ = false; dirty_ = false; }
  }
  ea_frame_params fp_{};
  ea_solve_params sp_{};
  Frame ref_, now_;
  bool have_ref_ = false, have_now_ = false, dirty_ = true;
  bool warm_start_ = false, solved_once_ = false, sample_ok_ = false;
  double pose_[7], init_pose_[7];
  std::vector<ea_summary> summaries_;
};
