// Command-line driver for include/edge_alignment/io.h, used by tests/test_io_formats.py (CPU only).
#include <edge_alignment/io.h>

#include <cstdlib>
#include <iostream>

using namespace ea;

static void dump(const std::string& path, const void* p, size_t n) {
  std::ofstream f(path, std::ios::binary);
  f.write(static_cast<const char*>(p), std::streamsize(n));
}

int main(int argc, char** argv) {
  try {
    const std::string cmd = argc > 1 ? argv[1] : "";
    if (cmd == "img" && argc == 5) {            // img <file> <flags> <out.raw>
      io::Image im = io::imread(argv[2], std::atoi(argv[3]));
      dump(argv[4], im.pixels.data(), im.pixels.size());
      std::cout << im.rows << " " << im.cols << " " << im.type << "\n";
    } else if (cmd == "repng" && argc == 5) {   // repng <file> <flags> <out.png>
      io::Image im = io::imread(argv[2], std::atoi(argv[3]));
      io::imwrite_png(argv[4], im.view());
    } else if (cmd == "ply" && argc == 4) {     // ply <file> <out.raw>
      std::vector<float> xyz = io::read_ply_xyz(argv[2]);
      dump(argv[3], xyz.data(), xyz.size() * 4);
      std::cout << xyz.size() / 3 << "\n";
    } else if (cmd == "obj" && argc == 6) {     // obj <in.raw> <n> <stride> <out.obj>
      std::vector<unsigned char> raw = io::detail::slurp(argv[2]);
      return io::save_point_cloud_obj(argv[5], reinterpret_cast<const float*>(raw.data()), size_t(std::atol(argv[3])), size_t(std::atol(argv[4])));
    } else if (cmd == "pose" && argc == 9) {    // pose qw qx qy qz tx ty tz
      double p[7], T[16], back[7];
      for (int i = 0; i < 7; ++i) p[i] = std::atof(argv[2 + i]);
      io::pose_to_matrix(p, T);
      io::matrix_to_pose(T, back);
      std::cout << io::prettyprint_pose(p) << "\n";
      std::cout.precision(17);
      for (int i = 0; i < 16; ++i) std::cout << T[i] << (i == 15 ? "\n" : " ");
      for (int i = 0; i < 7; ++i) std::cout << back[i] << (i == 6 ? "\n" : " ");
    } else if (cmd == "overlay" && argc == 17) {   // overlay <img> <xyz.raw> <n> <7 pose> <4 K> <out.png>
      io::Image im = io::imread(argv[2]);
      std::vector<unsigned char> raw = io::detail::slurp(argv[3]);
      double p[7], K[4];
      for (int i = 0; i < 7; ++i) p[i] = std::atof(argv[5 + i]);
      for (int i = 0; i < 4; ++i) K[i] = std::atof(argv[12 + i - 0]);
      std::vector<double> uv;
      io::reproject(reinterpret_cast<const float*>(raw.data()), size_t(std::atol(argv[4])), 3, p, K, uv);
      io::Image out = io::overlay(im.view(), uv);
      io::imwrite_png(argv[argc - 1], out.view());
    } else {
      std::cerr << "usage: io_test img|repng|ply|obj|pose|overlay ...\n";
      return 2;
    }
  } catch (const std::exception& e) {
    std::cerr << "error: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
