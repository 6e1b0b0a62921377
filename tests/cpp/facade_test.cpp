// Compiled and run by tests/test_gpu_facade.py: drives the reference's call sequence (src/ea.cpp:184-199) through the
// C++ facade and prints the pose so the Python test can compare it with the C-ABI / oracle result.
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>

#include "edge_alignment/EAResidue.h"
#include "edge_alignment/SolveEA.h"

static std::vector<unsigned char> read_file(const char* path, size_t n) {
  std::vector<unsigned char> b(n);
  FILE* f = fopen(path, "rb");
  if (!f || fread(b.data(), 1, n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
  fclose(f);
  return b;
}

int main(int argc, char** argv) {
  if (argc < 6) { fprintf(stderr, "usage: facade_test ref.bgr ref.depth16 now.bgr W H\n"); return 2; }
  const int W = atoi(argv[4]), H = atoi(argv[5]);
  auto ref_bgr = read_file(argv[1], size_t(W) * H * 3);
  auto ref_d = read_file(argv[2], size_t(W) * H * 2);
  auto now_bgr = read_file(argv[3], size_t(W) * H * 3);
  ea::Mat ref_im(H, W, ea::U8C3, ref_bgr.data()), ref_depth(H, W, ea::U16C1, ref_d.data());
  ea::Mat now_im(H, W, ea::U8C3, now_bgr.data()), now_depth;
  try {
    SolveEA* ea = new SolveEA();                    // src/ea.cpp:184
    ea->setK(525.0, 525.0, 319.5, 239.5);           // full-resolution TUM intrinsics (standalone_edge_align.cpp:152)
    ea->useStandalonePipeline();                    // edge_align_test1 settings so the oracle's golden results apply
    ea->setRefFrame(ref_im, ref_depth);             // src/ea.cpp:186
    ea->setNowFrame(now_im, now_depth);             // src/ea.cpp:188
    printf("inside %.6f\n", ea->_verify3dPts());    // src/ea.cpp:193
    ea->setAsCERESProblem();                        // src/ea.cpp:196
    double q[4], t[3];
    ea->getPose(q, t);
    const ea_summary& s = ea->getSummary()[0];
    printf("pose %.15g %.15g %.15g %.15g %.15g %.15g %.15g\n", q[0], q[1], q[2], q[3], t[0], t[1], t[2]);
    printf("summary %d %d %d %.15g %.15g\n", s.termination, s.iterations, s.n_residuals, s.initial_cost, s.final_cost);
    printf("npts %d\n", ea->refFrame().numEdgePoints());
    // EAResidue probe: first reference point at identity
    std::vector<float> p = ea->refFrame().edgePoints();
    const double Z = p[2] / 5000.0 /* standalone pipeline keeps raw u16 depth */, X = (p[0] - 319.5) * Z / 525.0, Y = (p[1] - 239.5) * Z / 525.0;
    EAResidue res(525.0, 525.0, 319.5, 239.5, X, Y, Z, ea->nowFrame());
    const double qi[4] = {1, 0, 0, 0}, ti[3] = {0, 0, 0};
    double r = -1, J[6];
    const bool ok = res(qi, ti, &r, J);
    printf("residue %d %.9g %.9g %.9g\n", int(ok), r, J[3], J[4]);
    // standalone/utils.h:82-92: EAResidue::Create(...) -- same evaluation through the factory; caller owns the object
    {
      EAResidue* made = EAResidue::Create(525.0, 525.0, 319.5, 239.5, X, Y, Z, ea->nowFrame());
      double r2 = -2;
      const bool ok2 = (*made)(qi, ti, &r2);
      printf("created %d %.9g\n", int(ok2), r2);
      delete made;
      bool threw = false;
      try { EAResidue wrong(500.0, 525.0, 319.5, 239.5, X, Y, Z, ea->nowFrame()); } catch (const std::invalid_argument&) { threw = true; }
      printf("wrongK %d\n", int(threw));
    }
    // a second setAsCERESProblem starts from identity again (src/SolveEA.cpp:130-131), not from the first result
    {
      ea->setAsCERESProblem();
      double q2[4], t2[3];
      ea->getPose(q2, t2);
      bool same = ea->getSummary()[0].iterations == s.iterations;   // (s aliases the refreshed summary: compare poses too)
      for (int i = 0; i < 4; ++i) same = same && q2[i] == q[i];
      for (int i = 0; i < 3; ++i) same = same && t2[i] == t[i];
      printf("restart %d\n", int(same));
    }
    // include/SolveEA.h:47 / src/SolveEA.cpp:218-272: the solver self-check
    ea->_sampleCERESProblem();
    printf("sample %d\n", int(ea->sampleProblemOk()));
    // the reference's own (ROS) flavour: Canny colour + exact DT [0,255] + every point + no loss + 25 iterations
    {
      SolveEA ros;
      ros.setK(525.0, 525.0, 319.5, 239.5);
      ros.setZeroDepthToOne(false);
      ros.setRefFrame(ref_im, ref_depth);
      ros.setNowFrame(now_im, now_depth);
      ros.setAsCERESProblem();
      double rq[4], rt[3];
      ros.getPose(rq, rt);
      const ea_summary& rs = ros.getSummary()[0];
      printf("rospose %.15g %.15g %.15g %.15g %.15g %.15g %.15g\n", rq[0], rq[1], rq[2], rq[3], rt[0], rt[1], rt[2]);
      printf("rossummary %d %d %d %.15g %.15g\n", rs.termination, rs.iterations, rs.n_residuals, rs.initial_cost, rs.final_cost);
    }
    // calling order is enforced (the reference left it unchecked, SolveEA.cpp:122)
    SolveEA fresh;
    try { fresh.setAsCERESProblem(); printf("order unchecked\n"); } catch (const std::exception&) { printf("order checked\n"); }
    delete ea;
  } catch (const std::exception& e) { fprintf(stderr, "exception: %s\n", e.what()); return 1; }
  return 0;
}
