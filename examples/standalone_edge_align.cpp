// The reference's standalone program flow (standalone/standalone_edge_align.cpp:109-305, edge_align_test1) on the B200 path:
// read an RGB-D frame A and an RGB frame B from files, align, print the initial / final guess in the reference's format,
// write the overlay images and the edge point cloud it writes.  No OpenCV, Ceres or Eigen: include/edge_alignment/io.h
// for the files, include/ea_cabi.h for everything else.
//
//   g++ -std=c++17 -I include examples/standalone_edge_align.cpp -L edge_alignment_b200 -l:libea_b200.so -lz \
//       -Wl,-rpath,$PWD/edge_alignment_b200 -o standalone_edge_align
//   ./standalone_edge_align rgb/1.png depth/1.png rgb/3.png [out_dir]
#include <ea_cabi.h>
#include <edge_alignment/io.h>

#include <cstdio>
#include <string>
#include <vector>

#define CHECK(call)                                                                      \
  do {                                                                                   \
    if ((call) != EA_OK) { std::fprintf(stderr, "%s\n  -> %s\n", #call, ea_last_error()); return 1; } \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 4) { std::fprintf(stderr, "usage: %s rgbA.png depthA.png rgbB.png [out_dir]\n", argv[0]); return 2; }
  const std::string out_dir = argc > 4 ? argv[4] : ".";
  try {
    ea::io::Image imA = ea::io::imread(argv[1]);                                   // SEA:118
    ea::io::Image imA_depth = ea::io::imread(argv[2], ea::io::READ_ANYDEPTH);      // SEA:119
    ea::io::Image imB = ea::io::imread(argv[3]);                                   // SEA:122
    if (imA_depth.type != ea::U16C1 || imA.type != ea::U8C3 || imB.type != ea::U8C3) { std::fprintf(stderr, "unexpected image types\n"); return 1; }
    std::printf("imA %dx%d, imA_depth %dx%d (16 bit), imB %dx%d\n", imA.cols, imA.rows, imA_depth.cols, imA_depth.rows, imB.cols, imB.rows);
    // the uploads below read width x height pixels from every buffer: all three images must have that size
    if (imA_depth.cols != imA.cols || imA_depth.rows != imA.rows || imB.cols != imA.cols || imB.rows != imA.rows) {
      std::fprintf(stderr, "the depth image and image B must have image A's size (%dx%d)\n", imA.cols, imA.rows);
      return 1;
    }

    ea_context* ctx = nullptr;
    CHECK(ea_create(0, &ctx));
    ea_frame_params fp;
    ea_frame_params_default(&fp);            // K = 525/525/319.5/239.5, depth/5000, threshold 35, median 3, DT in [0,1]: SEA:145-160, utils.cpp:38-83,201-281
    fp.width = imA.cols; fp.height = imA.rows;
    ea_solve_params sp;
    ea_solve_params_default(&sp);            // every 30th point, CauchyLoss(1), Ceres defaults: SEA:265-286
    ea_frameset* fs = nullptr;
    CHECK(ea_frameset_create(ctx, &fp, 2, &fs));
    const int32_t slotA = 0, slotB = 1;
    CHECK(ea_frameset_preprocess_host(fs, 1, &slotA, imA.pixels.data(), imA_depth.pixels.data(), EA_ROLE_REF));   // get_aX
    CHECK(ea_frameset_preprocess_host(fs, 1, &slotB, imB.pixels.data(), nullptr, EA_ROLE_NOW));                   // get_distance_transform
    CHECK(ea_sync(ctx));
    int n_pts = 0;
    CHECK(ea_frameset_get_num_points(fs, slotA, 0, &n_pts));
    std::printf("%d pts out of %d have a large gradient\n", n_pts, imA.cols * imA.rows);   // utils.cpp:262

    // the edge points as 3-D points (a_X) for the overlay and the .obj the reference writes (SEA:162-179)
    std::vector<float> uvd(size_t(n_pts) * 4), xyz(size_t(n_pts) * 3);
    int got = 0;
    CHECK(ea_frameset_get_points(fs, slotA, 0, uvd.data(), n_pts, &got));
    for (int i = 0; i < n_pts; ++i) {
      const double Z = uvd[4 * i + 2] / fp.depth_scale;
      xyz[3 * i] = float((uvd[4 * i] - fp.cx) * Z / fp.fx); xyz[3 * i + 1] = float((uvd[4 * i + 1] - fp.cy) * Z / fp.fy); xyz[3 * i + 2] = float(Z);
    }
    ea::io::save_point_cloud_obj(out_dir + "/sceneEdgeCloud.obj", xyz.data(), size_t(n_pts));

    double pose[7] = {1, 0, 0, 0, 0, 0, 0};   // b_T_a_optvar = Identity: SEA:240
    const double K4[4] = {fp.fx, fp.fy, fp.cx, fp.cy};
    std::vector<double> uv;
    ea::io::reproject(xyz.data(), size_t(n_pts), 3, pose, K4, uv);
    ea::io::imwrite_png(out_dir + "/initial.png", ea::io::overlay(imB.view(), uv).view());   // SEA:236-238
    std::printf("Initial Guess : %s\n", ea::io::prettyprint_pose(pose).c_str());             // SEA:242

    ea_summary summary;
    CHECK(ea_solve_batch(ctx, 1, fs, &slotA, fs, &slotB, pose, &sp, &summary));               // SEA:256-286
    std::printf("Residual blocks %d, iterations %d, cost %.6e -> %.6e, termination %d\n", summary.n_residuals, summary.iterations,
                summary.initial_cost, summary.final_cost, summary.termination);
    std::printf("Final Guess : %s\n", ea::io::prettyprint_pose(pose).c_str());               // SEA:301
    std::printf("pose %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", pose[0], pose[1], pose[2], pose[3], pose[4], pose[5], pose[6]);
    ea::io::reproject(xyz.data(), size_t(n_pts), 3, pose, K4, uv);
    ea::io::imwrite_png(out_dir + "/final.png", ea::io::overlay(imB.view(), uv).view());      // SEA:303

    ea_frameset_destroy(fs);
    ea_destroy(ctx);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
