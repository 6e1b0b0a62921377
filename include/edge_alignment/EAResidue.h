// class EAResidue -- the per-point residual functor of the reference (standalone/utils.h:38-99; ROS twin
// include/EAResidue.h:69-126), kept as a HOST-SIDE PROBE: same constructor shape (fx,fy,cx,cy, X,Y,Z, distance field)
// and the same operator()(q, t, residue) contract (returns false when |z'| < 0.01, utils.h:70-73).  The Ceres
// BiCubicInterpolator argument becomes a Frame holding the now-frame's distance transform; the evaluation runs the
// production device function on a one-point list.  Meant for tests and for probing single points, not for speed.
//
// The projection uses the now Frame's own intrinsics (that is where the distance field lives); the constructor's
// fx, fy, cx, cy must therefore equal them -- as they do at every call site of the reference (one K per problem,
// standalone_edge_align.cpp:152,271) -- and the constructor throws std::invalid_argument otherwise instead of silently
// evaluating for the wrong camera.
#pragma once
#include <cmath>

#include "Frame.h"

class EAResidue {
 public:
  EAResidue(double fx, double fy, double cx, double cy, double a_Xx, double a_Xy, double a_Xz, const Frame& now_frame)
      : X_(a_Xx), Y_(a_Xy), Z_(a_Xz), now_(&now_frame) {
    ea_frame_params p = now_frame.params();
    auto differs = [](double a, double b) { return std::abs(a - b) > 1e-9 * (1.0 + std::abs(b)); };
    if (differs(fx, p.fx) || differs(fy, p.fy) || differs(cx, p.cx) || differs(cy, p.cy))
      throw std::invalid_argument("EAResidue: fx, fy, cx, cy differ from the now Frame's intrinsics (the Frame's K is the one used)");
    p.fx = fx; p.fy = fy; p.cx = cx; p.cy = cy; p.max_points = 64;
    pt_.init(p, now_frame.context());
    const float p4[4] = {float(a_Xx), float(a_Xy), float(a_Xz), 1.0f};
    ea::check(ea_frameset_set_points(pt_.handle(), 0, 0, p4, 1, EA_POINTS_XYZ), "set_points");
    ea_solve_params_default(&sp_);
    sp_.point_stride = 1; sp_.loss_type = EA_LOSS_TRIVIAL;
  }
  // standalone/utils.h:82-92: the reference returns a ceres::CostFunction* owning a new EAResidue (AutoDiff<EAResidue,1,4,3>);
  // without Ceres the functor itself is the cost function: caller owns the returned object.
  static EAResidue* Create(const double fx, const double fy, const double cx, const double cy, const double a_Xx,
                           const double a_Xy, const double a_Xz, const Frame& now_frame) {
    return new EAResidue(fx, fy, cx, cy, a_Xx, a_Xy, a_Xz, now_frame);
  }
  // residue[0] = DT(u, v); optional jacobian[6] = d residue / d (half-angle rotation, translation)
  bool operator()(const double* quat_wxyz, const double* t, double* residue, double* jacobian6 = nullptr) const {
    double pose[7] = {quat_wxyz[0], quat_wxyz[1], quat_wxyz[2], quat_wxyz[3], t[0], t[1], t[2]};
    int n = 0, failed = 0;
    double raw = 0, J[6];
    ea::check(ea_eval(now_->context(), pt_.handle(), 0, now_->handle(), 0, 0, pose, &sp_, &n, &raw, nullptr, J, nullptr, &failed), "ea_eval");
    if (failed) return false;
    residue[0] = raw;
    if (jacobian6) for (int i = 0; i < 6; ++i) jacobian6[i] = J[i];
    return true;
  }

 private:
  double X_, Y_, Z_;
  const Frame* now_;
  mutable Frame pt_;
  ea_solve_params sp_{};
};
