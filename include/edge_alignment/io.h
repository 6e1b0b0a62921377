// File formats and console helpers either side of the pose-solve path -- host-side, header-only,
// zlib is the only dependency (link with -lz).  They replace what the reference's callers take from OpenCV / tinyply:
//   cv::imread of 8-bit colour PNG / BMP and 16-bit depth PNG      standalone_edge_align.cpp:118-122, 317-321, 876-886
//   read_ply_to_float3 (tinyply: vertex x, y, z)                    standalone_edge_align.cpp:2172-2220
//   SavePointCloudToObj ("v x y z" lines, no trailing newline)      standalone_edge_align.cpp:27-106
//   reproject + s_overlay (paint projected points red, imwrite)     standalone/utils.cpp:497-510, 594-608
//   PoseManipUtils::raw_to_eigenmat / eigenmat_to_raw / R2ypr / prettyprintMatrix4d   standalone/PoseManipUtils.cpp:3-27, 82-174
// Nothing here is on the measured path; the decoders favour clarity over speed.
#pragma once
#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "Mat.h"

namespace ea {
namespace io {

// An image that owns its pixels; view() is what SolveEA / Frame / the C ABI take.
struct Image {
  int rows = 0, cols = 0, type = U8C3;
  std::vector<unsigned char> pixels;
  Mat view() { return Mat(rows, cols, type, pixels.data()); }
  bool empty() const { return pixels.empty(); }
  size_t elem_size() const { return type == U8C3 ? 3 : (type == U8C1 ? 1 : (type == U16C1 ? 2 : 4)); }
};

enum ReadFlags { READ_COLOR = 0 /* 8-bit BGR, like cv::imread(path) */, READ_ANYDEPTH = 1 /* keep 16-bit gray: CV_LOAD_IMAGE_ANYDEPTH */,
                 READ_GRAYSCALE = 2 /* 8-bit gray: CV_LOAD_IMAGE_GRAYSCALE */ };

namespace detail {
inline std::vector<unsigned char> slurp(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}
inline uint32_t be32(const unsigned char* p) { return (uint32_t(p[0]) << 24) | (uint32_t(p[1]) << 16) | (uint32_t(p[2]) << 8) | p[3]; }
inline uint32_t le32(const unsigned char* p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }
inline uint16_t le16(const unsigned char* p) { return uint16_t(p[0] | (p[1] << 8)); }
inline int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
// OpenCV's 8-bit BGR -> gray (cvtColor COLOR_BGR2GRAY: 14-bit fixed point, what imread(GRAYSCALE) of a colour file gives to ~1 LSB)
inline unsigned char bgr2gray(const unsigned char* p) { return (unsigned char)((p[0] * 1868 + p[1] * 9617 + p[2] * 4899 + 8192) >> 14); }

// raster of decoded samples -> Image under the cv::imread conventions
inline void finish(Image& out, int w, int h, int channels /*1, 3 (RGB) or 4 (RGBA)*/, int bits, const std::vector<unsigned char>& raw, int flags) {
  out.rows = h; out.cols = w;
  const size_t n = size_t(w) * h;
  if (channels == 1 && bits == 16) {
    if (flags & READ_ANYDEPTH) {
      out.type = U16C1; out.pixels.resize(n * 2);
      for (size_t i = 0; i < n; ++i) { const uint16_t v = uint16_t((raw[2 * i] << 8) | raw[2 * i + 1]); std::memcpy(&out.pixels[2 * i], &v, 2); }
      return;
    }
    std::vector<unsigned char> g(n);
    for (size_t i = 0; i < n; ++i) g[i] = raw[2 * i];   // 16 -> 8 bit: high byte, as OpenCV
    if (flags & READ_GRAYSCALE) { out.type = U8C1; out.pixels.swap(g); return; }
    out.type = U8C3; out.pixels.resize(n * 3);
    for (size_t i = 0; i < n; ++i) out.pixels[3 * i] = out.pixels[3 * i + 1] = out.pixels[3 * i + 2] = g[i];
    return;
  }
  if (bits != 8) throw std::runtime_error("unsupported bit depth");
  if (channels == 1) {
    if (flags & (READ_GRAYSCALE | READ_ANYDEPTH)) { out.type = U8C1; out.pixels.assign(raw.begin(), raw.begin() + n); return; }
    out.type = U8C3; out.pixels.resize(n * 3);
    for (size_t i = 0; i < n; ++i) out.pixels[3 * i] = out.pixels[3 * i + 1] = out.pixels[3 * i + 2] = raw[i];
    return;
  }
  std::vector<unsigned char> bgr(n * 3);
  for (size_t i = 0; i < n; ++i) { const unsigned char* s = &raw[i * channels]; bgr[3 * i] = s[2]; bgr[3 * i + 1] = s[1]; bgr[3 * i + 2] = s[0]; }
  if (flags & READ_GRAYSCALE) {
    out.type = U8C1; out.pixels.resize(n);
    for (size_t i = 0; i < n; ++i) out.pixels[i] = bgr2gray(&bgr[3 * i]);
    return;
  }
  out.type = U8C3; out.pixels.swap(bgr);
}
}  // namespace detail

// PNG: colour types 0 (gray 8/16), 2 (RGB 8), 3 (palette 8), 6 (RGBA 8); non-interlaced.
inline void decode_png(const std::vector<unsigned char>& file, Image& out, int flags) {
  using namespace detail;
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  if (file.size() < 33 || std::memcmp(file.data(), sig, 8) != 0) throw std::runtime_error("not a PNG file");
  int w = 0, h = 0, bits = 0, ctype = -1, interlace = 0;
  std::vector<unsigned char> idat, palette;
  size_t pos = 8;
  while (pos + 12 <= file.size()) {
    const uint32_t len = be32(&file[pos]);
    const char* tag = reinterpret_cast<const char*>(&file[pos + 4]);
    if (pos + 12 + len > file.size()) throw std::runtime_error("truncated PNG chunk");
    const unsigned char* body = &file[pos + 8];
    if (!std::memcmp(tag, "IHDR", 4)) { w = int(be32(body)); h = int(be32(body + 4)); bits = body[8]; ctype = body[9]; interlace = body[12]; }
    else if (!std::memcmp(tag, "PLTE", 4)) palette.assign(body, body + len);
    else if (!std::memcmp(tag, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
    else if (!std::memcmp(tag, "IEND", 4)) break;
    pos += 12 + len;
  }
  if (w <= 0 || h <= 0 || interlace) throw std::runtime_error("unsupported PNG (size / interlace)");
  if (w > 32768 || h > 32768) throw std::runtime_error("PNG dimensions above 32768 are refused (unchecked IHDR sizes would size the buffers)");
  if (ctype < 0 || idat.empty()) throw std::runtime_error("PNG without IHDR / IDAT");
  int channels;
  switch (ctype) { case 0: channels = 1; break; case 2: channels = 3; break; case 3: channels = 1; break; case 6: channels = 4; break;
                   default: throw std::runtime_error("unsupported PNG colour type"); }
  if (!((bits == 8) || (bits == 16 && ctype == 0))) throw std::runtime_error("unsupported PNG bit depth");
  const size_t bpp = size_t(channels) * bits / 8, stride = size_t(w) * bpp;
  std::vector<unsigned char> raw((stride + 1) * h);
  uLongf got = uLongf(raw.size());
  if (uncompress(raw.data(), &got, idat.data(), uLong(idat.size())) != Z_OK || got != raw.size()) throw std::runtime_error("PNG inflate failed");
  std::vector<unsigned char> img(stride * h), zero(stride, 0);
  for (int y = 0; y < h; ++y) {
    const unsigned char* src = &raw[(stride + 1) * y];
    unsigned char* dst = &img[stride * y];
    const unsigned char* up = y ? &img[stride * (y - 1)] : zero.data();
    const int filter = src[0];
    ++src;
    for (size_t x = 0; x < stride; ++x) {
      const int a = x >= bpp ? dst[x - bpp] : 0, b = up[x], c = x >= bpp ? up[x - bpp] : 0;
      int v = src[x];
      switch (filter) { case 0: break; case 1: v += a; break; case 2: v += b; break; case 3: v += (a + b) >> 1; break; case 4: v += paeth(a, b, c); break;
                        default: throw std::runtime_error("bad PNG filter"); }
      dst[x] = (unsigned char)v;
    }
  }
  if (ctype == 3) {
    std::vector<unsigned char> rgb(size_t(w) * h * 3);
    for (size_t i = 0; i < size_t(w) * h; ++i) {
      const size_t k = size_t(img[i]) * 3;
      if (k + 2 >= palette.size()) throw std::runtime_error("PNG palette index out of range");
      rgb[3 * i] = palette[k]; rgb[3 * i + 1] = palette[k + 1]; rgb[3 * i + 2] = palette[k + 2];
    }
    finish(out, w, h, 3, 8, rgb, flags);
    return;
  }
  finish(out, w, h, channels, bits, img, flags);
}

// BMP: uncompressed 24-bit, 32-bit and 8-bit palettised, bottom-up or top-down.
inline void decode_bmp(const std::vector<unsigned char>& file, Image& out, int flags) {
  using namespace detail;
  if (file.size() < 54 || file[0] != 'B' || file[1] != 'M') throw std::runtime_error("not a BMP file");
  const uint32_t off = le32(&file[10]), hdr = le32(&file[14]);
  const int w = int(le32(&file[18]));
  int h = int(le32(&file[22]));
  const int bpp = le16(&file[28]);
  const uint32_t comp = le32(&file[30]);
  const bool top_down = h < 0;
  if (top_down) h = -h;
  if (w <= 0 || h <= 0 || (comp != 0 && !(comp == 3 && bpp == 32)) || !(bpp == 24 || bpp == 32 || bpp == 8)) throw std::runtime_error("unsupported BMP");
  const size_t row = ((size_t(w) * bpp + 31) / 32) * 4;
  if (w > 32768 || h > 32768) throw std::runtime_error("BMP dimensions above 32768 are refused");
  if (size_t(off) > file.size() || size_t(off) + row * size_t(h) > file.size()) throw std::runtime_error("truncated BMP");
  // 8-bit files index a 256-entry BGRA palette that sits between the headers and the pixel data: all of it must be there
  if (bpp == 8 && (size_t(hdr) < 12 || size_t(14) + hdr + 4 * 256 > size_t(off) || size_t(14) + hdr + 4 * 256 > file.size()))
    throw std::runtime_error("BMP palette is missing or truncated");
  const unsigned char* pal = file.data() + (bpp == 8 ? 14 + size_t(hdr) : 0);
  std::vector<unsigned char> rgb(size_t(w) * h * 3);
  bool gray_palette = (bpp == 8);
  for (int y = 0; y < h; ++y) {
    const unsigned char* src = &file[off + row * (top_down ? y : (h - 1 - y))];
    for (int x = 0; x < w; ++x) {
      unsigned char b, g, r;
      if (bpp == 8) { const unsigned char* e = pal + 4 * src[x]; b = e[0]; g = e[1]; r = e[2]; if (!(b == g && g == r)) gray_palette = false; }
      else { const unsigned char* e = src + size_t(x) * (bpp / 8); b = e[0]; g = e[1]; r = e[2]; }
      unsigned char* d = &rgb[(size_t(y) * w + x) * 3];
      d[0] = r; d[1] = g; d[2] = b;
    }
  }
  if (gray_palette && (flags & READ_GRAYSCALE)) {   // a gray palette file read as gray keeps its values exactly
    std::vector<unsigned char> g(size_t(w) * h);
    for (size_t i = 0; i < g.size(); ++i) g[i] = rgb[3 * i];
    finish(out, w, h, 1, 8, g, flags);
    return;
  }
  finish(out, w, h, 3, 8, rgb, flags);
}

// cv::imread replacement for the formats the reference's callers use.
inline Image imread(const std::string& path, int flags = READ_COLOR) {
  const std::vector<unsigned char> file = detail::slurp(path);
  Image out;
  if (file.size() >= 2 && file[0] == 'B' && file[1] == 'M') decode_bmp(file, out, flags);
  else decode_png(file, out, flags);
  return out;
}

// cv::imwrite(".png") for 8-bit gray / BGR and 16-bit gray images (filter 0, zlib default compression).
inline void imwrite_png(const std::string& path, const Mat& im) {
  const int t = im.type();
  if (!(t == U8C3 || t == U8C1 || t == U16C1)) throw std::runtime_error("imwrite_png: unsupported type");
  const int ch = (t == U8C3) ? 3 : 1, bits = (t == U16C1) ? 16 : 8;
  const size_t stride = size_t(im.cols) * ch * bits / 8;
  std::vector<unsigned char> raw((stride + 1) * im.rows);
  for (int y = 0; y < im.rows; ++y) {
    unsigned char* d = &raw[(stride + 1) * y];
    const unsigned char* s = im.data + im.step * y;
    *d++ = 0;
    if (t == U8C3) for (int x = 0; x < im.cols; ++x) { d[3 * x] = s[3 * x + 2]; d[3 * x + 1] = s[3 * x + 1]; d[3 * x + 2] = s[3 * x]; }
    else if (t == U16C1) for (int x = 0; x < im.cols; ++x) { d[2 * x] = s[2 * x + 1]; d[2 * x + 1] = s[2 * x]; }
    else std::memcpy(d, s, stride);
  }
  uLongf clen = compressBound(uLong(raw.size()));
  std::vector<unsigned char> comp(clen);
  if (compress(comp.data(), &clen, raw.data(), uLong(raw.size())) != Z_OK) throw std::runtime_error("deflate failed");
  std::ofstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot write " + path);
  auto chunk = [&](const char* tag, const unsigned char* body, uint32_t len) {
    unsigned char h[8] = {(unsigned char)(len >> 24), (unsigned char)(len >> 16), (unsigned char)(len >> 8), (unsigned char)len,
                          (unsigned char)tag[0], (unsigned char)tag[1], (unsigned char)tag[2], (unsigned char)tag[3]};
    uLong crc = crc32(0L, h + 4, 4);
    if (len) crc = crc32(crc, body, len);
    const unsigned char c[4] = {(unsigned char)(crc >> 24), (unsigned char)(crc >> 16), (unsigned char)(crc >> 8), (unsigned char)crc};
    f.write(reinterpret_cast<const char*>(h), 8);
    if (len) f.write(reinterpret_cast<const char*>(body), len);
    f.write(reinterpret_cast<const char*>(c), 4);
  };
  static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  f.write(reinterpret_cast<const char*>(sig), 8);
  unsigned char ihdr[13] = {0};
  const uint32_t w = uint32_t(im.cols), h = uint32_t(im.rows);
  ihdr[0] = (unsigned char)(w >> 24); ihdr[1] = (unsigned char)(w >> 16); ihdr[2] = (unsigned char)(w >> 8); ihdr[3] = (unsigned char)w;
  ihdr[4] = (unsigned char)(h >> 24); ihdr[5] = (unsigned char)(h >> 16); ihdr[6] = (unsigned char)(h >> 8); ihdr[7] = (unsigned char)h;
  ihdr[8] = (unsigned char)bits; ihdr[9] = (unsigned char)(ch == 3 ? 2 : 0);
  chunk("IHDR", ihdr, 13);
  chunk("IDAT", comp.data(), uint32_t(clen));
  chunk("IEND", nullptr, 0);
}

// PLY vertex positions (read_ply_to_float3): ascii or binary_little_endian; x, y, z may be float or double and may be
// mixed with other scalar properties.  Returns xyz triplets.
inline std::vector<float> read_ply_xyz(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("cannot open " + path);
  std::string line;
  if (!std::getline(f, line) || line.substr(0, 3) != "ply") throw std::runtime_error("not a PLY file");
  struct Prop { std::string name; int size; bool is_float, is_signed; };
  std::vector<Prop> props;
  size_t n_vertex = 0;
  bool ascii = true, in_vertex = false, vertex_first = true, seen_element = false;
  auto type_of = [](const std::string& t, Prop& p) {
    if (t == "float" || t == "float32") { p.size = 4; p.is_float = true; }
    else if (t == "double" || t == "float64") { p.size = 8; p.is_float = true; }
    else if (t == "char" || t == "int8") { p.size = 1; p.is_float = false; p.is_signed = true; }
    else if (t == "uchar" || t == "uint8") { p.size = 1; p.is_float = false; p.is_signed = false; }
    else if (t == "short" || t == "int16") { p.size = 2; p.is_float = false; p.is_signed = true; }
    else if (t == "ushort" || t == "uint16") { p.size = 2; p.is_float = false; p.is_signed = false; }
    else if (t == "int" || t == "int32") { p.size = 4; p.is_float = false; p.is_signed = true; }
    else if (t == "uint" || t == "uint32") { p.size = 4; p.is_float = false; p.is_signed = false; }
    else throw std::runtime_error("PLY: unknown property type " + t);
  };
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::istringstream ss(line);
    std::string key;
    ss >> key;
    if (key == "format") { std::string fmt; ss >> fmt; if (fmt == "ascii") ascii = true; else if (fmt == "binary_little_endian") ascii = false; else throw std::runtime_error("PLY: unsupported format " + fmt); }
    else if (key == "element") { std::string name; size_t n; ss >> name >> n; in_vertex = (name == "vertex"); if (in_vertex) { n_vertex = n; vertex_first = !seen_element; } seen_element = true; }
    else if (key == "property" && in_vertex) { std::string t, name; ss >> t; if (t == "list") throw std::runtime_error("PLY: list property in the vertex element"); ss >> name; Prop p{name, 0, false, false}; type_of(t, p); props.push_back(p); }
    else if (key == "end_header") break;
  }
  if (!vertex_first) throw std::runtime_error("PLY: the vertex element must come first");
  int ix = -1, iy = -1, iz = -1;
  for (size_t i = 0; i < props.size(); ++i) { if (props[i].name == "x") ix = int(i); if (props[i].name == "y") iy = int(i); if (props[i].name == "z") iz = int(i); }
  if (ix < 0 || iy < 0 || iz < 0) throw std::runtime_error("PLY: vertex element has no x, y, z");
  std::vector<float> xyz(n_vertex * 3);
  std::vector<double> vals(props.size());
  for (size_t v = 0; v < n_vertex; ++v) {
    if (ascii) {
      for (size_t i = 0; i < props.size(); ++i) if (!(f >> vals[i])) throw std::runtime_error("PLY: truncated vertex data");
    } else {
      for (size_t i = 0; i < props.size(); ++i) {
        unsigned char b[8] = {0};
        f.read(reinterpret_cast<char*>(b), props[i].size);
        if (!f) throw std::runtime_error("PLY: truncated vertex data");
        const Prop& p = props[i];
        if (p.is_float) { if (p.size == 4) { float t; std::memcpy(&t, b, 4); vals[i] = t; } else { double t; std::memcpy(&t, b, 8); vals[i] = t; } }
        else if (p.size == 1) vals[i] = p.is_signed ? double(int8_t(b[0])) : double(b[0]);
        else if (p.size == 2) { uint16_t t; std::memcpy(&t, b, 2); vals[i] = p.is_signed ? double(int16_t(t)) : double(t); }
        else { uint32_t t; std::memcpy(&t, b, 4); vals[i] = p.is_signed ? double(int32_t(t)) : double(t); }
      }
    }
    xyz[3 * v] = float(vals[ix]); xyz[3 * v + 1] = float(vals[iy]); xyz[3 * v + 2] = float(vals[iz]);
  }
  return xyz;
}

// SavePointCloudToObj for float data: one "v x y z" line per point, values printed like an ostream prints a float, no
// newline after the last line.  stride = floats per point in `xyz` (3, or 4 for homogeneous lists).
inline int save_point_cloud_obj(const std::string& path, const float* xyz, size_t n, size_t stride = 3) {
  if (!xyz || n == 0) { std::printf("no points\n"); return -1; }
  if (path.empty()) { std::printf("path empty\n"); return -1; }
  std::ofstream f(path.c_str());
  if (!f) return -1;
  for (size_t i = 0; i < n; ++i) {
    f << "v " << xyz[i * stride] << ' ' << xyz[i * stride + 1] << ' ' << xyz[i * stride + 2];
    if (i + 1 < n) f << '\n';
  }
  return 0;
}

// ---- poses (pose7 = {qw, qx, qy, qz, tx, ty, tz}; T row-major 4x4) -------------------------------------------------
inline void pose_to_matrix(const double* pose7, double* T16) {   // PoseManipUtils::raw_to_eigenmat
  const double w = pose7[0], x = pose7[1], y = pose7[2], z = pose7[3];
  const double n = 1.0 / std::sqrt(w * w + x * x + y * y + z * z);
  const double qw = w * n, qx = x * n, qy = y * n, qz = z * n;
  const double R[9] = {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw),
                       2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw),
                       2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)};
  for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T16[4 * r + c] = R[3 * r + c]; T16[4 * r + 3] = pose7[4 + r]; }
  T16[12] = T16[13] = T16[14] = 0.0; T16[15] = 1.0;
}
inline void matrix_to_pose(const double* T16, double* pose7) {   // PoseManipUtils::eigenmat_to_raw (Eigen's Quaterniond(Matrix3d))
  const double m00 = T16[0], m01 = T16[1], m02 = T16[2], m10 = T16[4], m11 = T16[5], m12 = T16[6], m20 = T16[8], m21 = T16[9], m22 = T16[10];
  const double tr = m00 + m11 + m22;
  double w, x, y, z;
  if (tr > 0.0) { double t = std::sqrt(tr + 1.0); w = 0.5 * t; t = 0.5 / t; x = (m21 - m12) * t; y = (m02 - m20) * t; z = (m10 - m01) * t; }
  else if (m00 >= m11 && m00 >= m22) { double t = std::sqrt(m00 - m11 - m22 + 1.0); x = 0.5 * t; t = 0.5 / t; w = (m21 - m12) * t; y = (m10 + m01) * t; z = (m20 + m02) * t; }
  else if (m11 >= m22) { double t = std::sqrt(m11 - m22 - m00 + 1.0); y = 0.5 * t; t = 0.5 / t; w = (m02 - m20) * t; z = (m21 + m12) * t; x = (m01 + m10) * t; }
  else { double t = std::sqrt(m22 - m00 - m11 + 1.0); z = 0.5 * t; t = 0.5 / t; w = (m10 - m01) * t; x = (m02 + m20) * t; y = (m12 + m21) * t; }
  pose7[0] = w; pose7[1] = x; pose7[2] = y; pose7[3] = z; pose7[4] = T16[3]; pose7[5] = T16[7]; pose7[6] = T16[11];
}
inline void matrix_to_ypr(const double* T16, double* ypr_deg) {   // PoseManipUtils::R2ypr, degrees
  const double n0 = T16[0], n1 = T16[4], n2 = T16[8], o0 = T16[1], o1 = T16[5], a0 = T16[2], a1 = T16[6];
  const double y = std::atan2(n1, n0);
  const double p = std::atan2(-n2, n0 * std::cos(y) + n1 * std::sin(y));
  const double r = std::atan2(a0 * std::sin(y) - a1 * std::cos(y), -o0 * std::sin(y) + o1 * std::cos(y));
  const double k = 180.0 / M_PI;
  ypr_deg[0] = y * k; ypr_deg[1] = p * k; ypr_deg[2] = r * k;
}
inline std::string prettyprint_pose(const double* pose7) {   // PoseManipUtils::prettyprintMatrix4d
  double T[16], ypr[3];
  pose_to_matrix(pose7, T);
  matrix_to_ypr(T, ypr);
  char tmp[200];
  std::snprintf(tmp, sizeof tmp, ":YPR=(%4.2f,%4.2f,%4.2f)  :TxTyTz=(%4.2f,%4.2f,%4.2f)", ypr[0], ypr[1], ypr[2], T[3], T[7], T[11]);
  return std::string(tmp);
}

// ---- visual check: reproject + s_overlay ------------------------------------------------------------------------------
// uv[2 i], uv[2 i + 1] = K * (b_T_a * X_i) / z.  K = {fx, fy, cx, cy}.
inline void reproject(const float* xyz, size_t n, size_t stride, const double* pose7, const double* K4, std::vector<double>& uv) {
  double T[16];
  pose_to_matrix(pose7, T);
  uv.resize(2 * n);
  for (size_t i = 0; i < n; ++i) {
    const double X = xyz[i * stride], Y = xyz[i * stride + 1], Z = xyz[i * stride + 2];
    const double x = T[0] * X + T[1] * Y + T[2] * Z + T[3], y = T[4] * X + T[5] * Y + T[6] * Z + T[7], z = T[8] * X + T[9] * Y + T[10] * Z + T[11];
    uv[2 * i] = K4[0] * (x / z) + K4[2];
    uv[2 * i + 1] = K4[1] * (y / z) + K4[3];
  }
}
// Paints the projected points red on a copy of a BGR image.  Unlike the reference (utils.cpp:603) points outside the image are skipped.
inline Image overlay(const Mat& bgr, const std::vector<double>& uv) {
  if (bgr.type() != U8C3) throw std::runtime_error("overlay needs an 8-bit BGR image");
  Image out;
  out.rows = bgr.rows; out.cols = bgr.cols; out.type = U8C3; out.pixels.resize(size_t(bgr.rows) * bgr.cols * 3);
  for (int y = 0; y < bgr.rows; ++y) std::memcpy(&out.pixels[size_t(y) * bgr.cols * 3], bgr.data + bgr.step * y, size_t(bgr.cols) * 3);
  for (size_t i = 0; i + 1 < uv.size(); i += 2) {
    if (!(uv[i] >= 0.0 && uv[i + 1] >= 0.0)) continue;
    const long u = long(uv[i]), v = long(uv[i + 1]);      // truncation, as at<>(double, double) converts
    if (u >= bgr.cols || v >= bgr.rows) continue;
    unsigned char* p = &out.pixels[(size_t(v) * bgr.cols + size_t(u)) * 3];
    p[0] = 0; p[1] = 0; p[2] = 255;
  }
  return out;
}

}  // namespace io
}  // namespace ea
