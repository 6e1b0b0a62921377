// Minimal image view used by the facade when OpenCV headers are not available.  Any type with the same members
// (cv::Mat in particular: rows, cols, data, step, type(), channels()) can be passed to SolveEA / Frame directly.
#pragma once
#include <cstddef>
#include <cstdint>

namespace ea {

enum MatType { U8C3 = 16 /*CV_8UC3*/, U8C1 = 0 /*CV_8UC1*/, U16C1 = 2 /*CV_16UC1*/, F32C1 = 5 /*CV_32FC1*/ };

struct Mat {
  int rows = 0, cols = 0;
  unsigned char* data = nullptr;
  size_t step = 0;  // bytes per row
  int type_ = U8C3;
  Mat() = default;
  Mat(int r, int c, int t, void* d, size_t s = 0) : rows(r), cols(c), data(static_cast<unsigned char*>(d)), type_(t) {
    const size_t elem = (t == U8C3) ? 3 : (t == U8C1 ? 1 : (t == U16C1 ? 2 : 4));
    step = s ? s : elem * size_t(c);
  }
  int type() const { return type_; }
};

}  // namespace ea
