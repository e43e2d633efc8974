// C++ host of the B200 ray-caster (see tvk_host.h).  Pure host code: every device
// operation is a call into the C ABI of include/tvk_cuda.h.
#include "tvk_host.h"
#include "tvk_uvf.h"
#include <cuda_runtime_api.h>   // only cudaHostAlloc/cudaFreeHost for the pinned out-of-core store
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <fstream>

namespace tvk {

// ---------------------------------------------------------------------------- math
float Vec3::length() const { return std::sqrt(x * x + y * y + z * z); }
Vec3 Vec3::normalized() const { const float l = length(); return Vec3(x / l, y / l, z / l); }

Matrix4::Matrix4() { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0f : 0.0f; }

Matrix4 Matrix4::operator*(const Matrix4& o) const {      // Vectors.h:936-946
  Matrix4 r;
  for (int x = 0; x < 16; x += 4)
    for (int y = 0; y < 4; ++y)
      r.m[x + y] = m[0 + x] * o.m[0 + y] + m[1 + x] * o.m[4 + y] + m[2 + x] * o.m[8 + y] + m[3 + x] * o.m[12 + y];
  return r;
}

Matrix4 Matrix4::transposed() const {
  Matrix4 r;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[4 * i + j] = m[4 * j + i];
  return r;
}

// General 4x4 inverse by cofactors (the reference expands the same adjugate in
// fp32, Vectors.h:1042-1140; we accumulate in double and round once).
Matrix4 Matrix4::inverse() const {
  double a[16], inv[16];
  for (int i = 0; i < 16; ++i) a[i] = m[i];
  inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
  inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
  inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
  inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
  inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
  inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
  inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
  inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
  inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
  inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
  inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
  inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
  inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
  inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
  inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
  inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
  const double det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
  Matrix4 r;
  for (int i = 0; i < 16; ++i) r.m[i] = (float)(inv[i] / det);
  return r;
}

Matrix4 Matrix4::Translation(float x, float y, float z) { Matrix4 r; r.m[12] = x; r.m[13] = y; r.m[14] = z; return r; }
Matrix4 Matrix4::Scaling(float x, float y, float z) { Matrix4 r; r.m[0] = x; r.m[5] = y; r.m[10] = z; return r; }
Matrix4 Matrix4::RotationX(double a) {                      // Vectors.h:984-992
  Matrix4 r; const float c = (float)std::cos(a), s = (float)std::sin(a);
  r.m[5] = c; r.m[6] = s; r.m[9] = -s; r.m[10] = c; return r;
}
Matrix4 Matrix4::RotationY(double a) {
  Matrix4 r; const float c = (float)std::cos(a), s = (float)std::sin(a);
  r.m[0] = c; r.m[2] = -s; r.m[8] = s; r.m[10] = c; return r;
}
Matrix4 Matrix4::RotationZ(double a) {
  Matrix4 r; const float c = (float)std::cos(a), s = (float)std::sin(a);
  r.m[0] = c; r.m[1] = s; r.m[4] = -s; r.m[5] = c; return r;
}
Matrix4 Matrix4::LookAt(const Vec3& eye, const Vec3& at, const Vec3& up) {    // Vectors.h:1250-1265
  Vec3 F = at - eye, U = up;
  Vec3 S = F.cross(U);
  U = S.cross(F);
  F = F.normalized(); U = U.normalized(); S = S.normalized();
  Matrix4 r;
  r.m[0] = S.x; r.m[4] = S.y; r.m[8] = S.z; r.m[12] = -S.dot(eye);
  r.m[1] = U.x; r.m[5] = U.y; r.m[9] = U.z; r.m[13] = -U.dot(eye);
  r.m[2] = -F.x; r.m[6] = -F.y; r.m[10] = -F.z; r.m[14] = F.dot(eye);
  r.m[3] = 0; r.m[7] = 0; r.m[11] = 0; r.m[15] = 1;
  return r;
}
Matrix4 Matrix4::Perspective(float fovy, float aspect, float n, float f) {    // Vectors.h:1267-1277
  fovy = fovy * float(3.14159265358979323846 / 180.0);
  const float cotan = float(1.0 / std::tan(double(fovy) / 2.0));
  Matrix4 r;
  for (int i = 0; i < 16; ++i) r.m[i] = 0;
  r.m[0] = cotan / aspect; r.m[5] = cotan;
  r.m[10] = -(f + n) / (f - n); r.m[14] = -2.0f * (f * n) / (f - n);
  r.m[11] = -1.0f;
  return r;
}
Matrix4 Matrix4::PerspectiveOffCenter(float left, float right, float bottom, float top, float n, float f) {   // Vectors.h:1286-1291
  Matrix4 r;
  for (int i = 0; i < 16; ++i) r.m[i] = 0;
  r.m[0] = 2.0f * n / (right - left); r.m[8] = (right + left) / (right - left);
  r.m[5] = 2.0f * n / (top - bottom); r.m[9] = (top + bottom) / (top - bottom);
  r.m[10] = -(f + n) / (f - n); r.m[14] = -2.0f * (f * n) / (f - n);
  r.m[11] = -1.0f;
  return r;
}
Matrix4 Matrix4::Ortho(float left, float right, float bottom, float top, float n, float f) {   // Vectors.h:1279-1284
  Matrix4 r;
  r.m[0] = 2.0f / (right - left); r.m[12] = -(right + left) / (right - left);
  r.m[5] = 2.0f / (top - bottom); r.m[13] = -(top + bottom) / (top - bottom);
  r.m[10] = -2.0f / (f - n); r.m[14] = -(f + n) / (f - n);
  return r;
}
void Matrix4::StereoLookAtAndProjection(const Vec3& eye, const Vec3& at, const Vec3& up, float fovy, float aspect, float n, float f,
                                        float focalLength, float eyeDist, Matrix4 view[2], Matrix4 proj[2]) {
  const float radians = float(3.14159265358979323846 / 180.0) * fovy / 2;
  const float wd2 = n * float(std::tan(radians));
  const float nfdl = n / focalLength;
  const float shift = eyeDist * nfdl;
  const float top = wd2, bottom = -wd2;
  proj[0] = PerspectiveOffCenter(-aspect * wd2 + shift, aspect * wd2 + shift, bottom, top, n, f);
  proj[1] = PerspectiveOffCenter(-aspect * wd2 - shift, aspect * wd2 - shift, bottom, top, n, f);
  const Matrix4 look = LookAt(eye, at, up);
  view[0] = Translation(eyeDist, 0.0f, 0.0f) * look;
  view[1] = Translation(-eyeDist, 0.0f, 0.0f) * look;
}
void Matrix4::transform(const float v[4], float out[4]) const {
  for (int j = 0; j < 4; ++j) out[j] = v[0] * m[j] + v[1] * m[4 + j] + v[2] * m[8 + j] + v[3] * m[12 + j];
}

// ---------------------------------------------------------------------------- TransferFunction1D
void TransferFunction1D::Set(const float* rgba, size_t n) { m_color.assign(rgba, rgba + 4 * n); ComputeNonZeroLimits(); }

bool TransferFunction1D::Load(std::istream& f) {                              // TransferFunction1D.cpp:279-295
  uint32_t n = 0;
  f >> n;
  if (!f || n == 0 || n > (1u << 24)) return false;
  m_color.assign(4 * (size_t)n, 0.0f);
  for (size_t i = 0; i < 4 * (size_t)n; ++i) if (!(f >> m_color[i])) return false;
  ComputeNonZeroLimits();
  return true;
}
bool TransferFunction1D::Save(std::ostream& f) const {                        // :298-309 ("r g b a " + endl per entry)
  f << GetSize() << std::endl;
  for (size_t i = 0; i < GetSize(); ++i) {
    for (int j = 0; j < 4; ++j) f << m_color[4 * i + j] << " ";
    f << std::endl;
  }
  return (bool)f;
}
bool TransferFunction1D::Load(const std::string& filename) {
  std::ifstream f(filename.c_str());
  return f.is_open() && Load(f);
}
bool TransferFunction1D::Save(const std::string& filename) const {
  std::ofstream f(filename.c_str());
  return f.is_open() && Save(f);
}

// ---------------------------------------------------------------------------- TransferFunction2D
static bool LoadSwatch(std::istream& f, TFPolygon& p) {                       // TFPolygon::Load, TransferFunction2D.cpp:511-543
  uint32_t n = 0;
  f >> p.radial;
  f >> n;
  if (!f || n > (1u << 20)) return false;
  p.points.resize(n);
  for (auto& q : p.points) f >> q[0] >> q[1];
  f >> p.gradientCoords[0][0] >> p.gradientCoords[0][1];
  f >> p.gradientCoords[1][0] >> p.gradientCoords[1][1];
  f >> n;
  if (!f || n > (1u << 20)) return false;
  p.stops.resize(n);
  for (auto& s : p.stops) { f >> s.first; for (int j = 0; j < 4; ++j) f >> s.second[j]; }
  return !f.fail();
}
static void SaveSwatch(std::ostream& f, const TFPolygon& p) {                 // TFPolygon::Save, :545-568
  f << p.radial << std::endl;
  f << uint32_t(p.points.size()) << std::endl;
  for (const auto& q : p.points) { f << q[0] << " " << q[1] << " "; f << std::endl; }
  f << p.gradientCoords[0][0] << " " << p.gradientCoords[0][1] << " ";
  f << p.gradientCoords[1][0] << " " << p.gradientCoords[1][1];
  f << std::endl;
  f << p.stops.size() << std::endl;
  for (const auto& s : p.stops) {
    f << s.first << "  ";
    for (int j = 0; j < 4; ++j) f << s.second[j] << " ";
    f << std::endl;
  }
}
bool TransferFunction2D::LoadRest(std::istream& f, bool allowNoSwatch) {
  uint32_t n = 0;
  f >> n;
  if (!f || n > (1u << 20)) return false;
  m_swatches.assign(n, TFPolygon());
  if (n == 0 && !allowNoSwatch) return false;                                 // "No swatches?!  Ridiculous 2D TFqn, bailing..." (:181-184)
  for (auto& s : m_swatches) if (!LoadSwatch(f, s)) return false;
  return true;
}
bool TransferFunction2D::Load(const std::string& filename) {
  std::ifstream f(filename.c_str());
  if (!f.is_open()) return false;
  f >> m_w >> m_h;
  if (!f || !m_w || !m_h) return false;
  if (!m_trans1d.Load(f)) return false;
  m_trans1d.Resample(m_w);
  return LoadRest(f, false);
}
bool TransferFunction2D::Load(const std::string& filename, size_t targetW, size_t targetH) {
  std::ifstream f(filename.c_str());
  if (!f.is_open() || !targetW || !targetH) return false;
  m_w = targetW; m_h = targetH;
  size_t fw = 0, fh = 0;
  f >> fw >> fh;                                                              // "ignore the size in the file" (:109-111)
  if (!f) return false;
  if (!m_trans1d.Load(f)) return false;
  m_trans1d.Resample(targetW);
  return LoadRest(f, true);
}
bool TransferFunction2D::Save(const std::string& filename) const {
  std::ofstream f(filename.c_str());
  if (!f.is_open()) return false;
  f << m_w << " " << m_h << std::endl;
  m_trans1d.Save(f);
  f << m_swatches.size() << std::endl;
  for (const auto& s : m_swatches) SaveSwatch(f, s);
  return (bool)f;
}

void TransferFunction2D::Render(std::vector<uint8_t>& rgba, double limits[4]) const {
  // GetRenderSize (TransferFunction2D.h:136-144): a 2:1 canvas of at most 4096 pixels, scaled to the table afterwards
  const size_t cw0 = std::min<size_t>(m_w, 4096), ch0 = std::min<size_t>(m_h, 4096);
  const size_t W = cw0 > ch0 ? cw0 : ch0 * 2, H = cw0 > ch0 ? size_t(cw0 / 2.0) : ch0;
  std::vector<float> canvas(4 * W * std::max<size_t>(H, 1), 0.0f);              // straight (non-premultiplied) RGBA in [0,1]
  // the 1-D function, one pixel per entry, stretched over the canvas (Update1DTrans :399-424 + drawImage :311)
  TransferFunction1D t1 = m_trans1d;
  const size_t n1 = std::min<size_t>(m_w, t1.GetSize());
  std::vector<uint8_t> row;
  t1.GetByteArray(row);
  for (size_t y = 0; y < H; ++y)
    for (size_t x = 0; x < W && n1; ++x) {
      const size_t i = std::min<size_t>(n1 - 1, x * n1 / W);
      for (int c = 0; c < 4; ++c) canvas[4 * (y * W + x) + c] = row[4 * i + c] / 255.0f;
    }
  for (const TFPolygon& sw : m_swatches) {
    const size_t np = sw.points.size();
    if (np < 3 || sw.stops.empty()) continue;
    // Normalized2Offscreen (:288-291): int(coord * size)
    std::vector<std::array<int, 2>> pt(np);
    for (size_t j = 0; j < np; ++j) pt[j] = {int(sw.points[j][0] * int(W)), int(sw.points[j][1] * int(H))};
    const int g0[2] = {int(sw.gradientCoords[0][0] * int(W)), int(sw.gradientCoords[0][1] * int(H))};
    const int g1[2] = {int(sw.gradientCoords[1][0] * int(W)), int(sw.gradientCoords[1][1] * int(H))};
    const double dx = g1[0] - g0[0], dy = g1[1] - g0[1], len2 = dx * dx + dy * dy, rad = std::sqrt(len2);
    // stops as Qt keeps them: sorted by position, colours quantised to 8 bits (QColor(int(c*255)...), :333-339)
    std::vector<std::pair<float, std::array<float, 4>>> st = sw.stops;
    std::stable_sort(st.begin(), st.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    for (auto& s : st) for (int c = 0; c < 4; ++c) s.second[c] = float(std::min(255, std::max(0, int(s.second[c] * 255)))) / 255.0f;
    auto colour = [&](double t, float out[4]) {
      t = std::min(1.0, std::max(0.0, t));                                     // pad spread
      if (t <= st.front().first) { for (int c = 0; c < 4; ++c) out[c] = st.front().second[c]; return; }
      if (t >= st.back().first) { for (int c = 0; c < 4; ++c) out[c] = st.back().second[c]; return; }
      for (size_t k = 1; k < st.size(); ++k)
        if (t <= st[k].first) {
          const double span = st[k].first - st[k - 1].first, u = span > 0 ? (t - st[k - 1].first) / span : 1.0;
          for (int c = 0; c < 4; ++c) out[c] = float(st[k - 1].second[c] * (1.0 - u) + st[k].second[c] * u);
          return;
        }
    };
    int ymin = INT32_MAX, ymax = INT32_MIN;
    for (const auto& q : pt) { ymin = std::min(ymin, q[1]); ymax = std::max(ymax, q[1]); }
    for (int y = std::max(ymin, 0); y < std::min<int>(ymax, (int)H); ++y) {
      const double cy = y + 0.5;
      std::vector<double> xs;                                                  // odd-even fill: crossings of the scan line through pixel centres
      for (size_t j = 0; j < np; ++j) {
        const auto& a = pt[j]; const auto& b = pt[(j + 1) % np];
        if ((a[1] <= cy) != (b[1] <= cy)) xs.push_back(a[0] + (cy - a[1]) * double(b[0] - a[0]) / double(b[1] - a[1]));
      }
      std::sort(xs.begin(), xs.end());
      for (size_t k = 0; k + 1 < xs.size(); k += 2) {
        const int x0 = std::max(0, (int)std::ceil(xs[k] - 0.5)), x1 = std::min<int>((int)W, (int)std::ceil(xs[k + 1] - 0.5));
        for (int x = x0; x < x1; ++x) {
          const double px = x + 0.5, py = cy;
          double t;
          if (sw.radial) t = rad > 0 ? std::sqrt((px - g0[0]) * (px - g0[0]) + (py - g0[1]) * (py - g0[1])) / rad : 1.0;
          else t = len2 > 0 ? ((px - g0[0]) * dx + (py - g0[1]) * dy) / len2 : 0.0;
          float src[4];
          colour(t, src);
          float* d = &canvas[4 * ((size_t)y * W + x)];
          const float sa = src[3], da = d[3], oa = sa + da * (1.0f - sa);          // source-over on straight alpha
          for (int c = 0; c < 3; ++c) d[c] = oa > 0 ? (src[c] * sa + d[c] * da * (1.0f - sa)) / oa : 0.0f;
          d[3] = oa;
        }
      }
    }
  }
  // scaled(m_iSize) (:355, fast transformation = nearest) + 8-bit RGBA
  rgba.assign(4 * m_w * m_h, 0);
  uint64_t lim[4] = {m_w, 0, m_h, 0};
  for (size_t y = 0; y < m_h; ++y)
    for (size_t x = 0; x < m_w; ++x) {
      const size_t sx = std::min(W - 1, x * W / m_w), sy = std::min(std::max<size_t>(H, 1) - 1, y * H / m_h);
      for (int c = 0; c < 4; ++c) rgba[4 * (y * m_w + x) + c] = (uint8_t)std::lround(std::min(1.0f, std::max(0.0f, canvas[4 * (sy * W + sx) + c])) * 255.0f);
      if (rgba[4 * (y * m_w + x) + 3]) {
        lim[0] = std::min<uint64_t>(lim[0], x); lim[1] = std::max<uint64_t>(lim[1], x);
        lim[2] = std::min<uint64_t>(lim[2], y); lim[3] = std::max<uint64_t>(lim[3], y);
      }
    }
  for (int i = 0; i < 4; ++i) limits[i] = double(lim[i]);
}
static float Smoothstep(float x) { return 3 * x * x - 2 * x * x * x; }
void TransferFunction1D::SetStdFunction(float c, float g, int comp) {        // TransferFunction1D.cpp:81-112
  const size_t n = GetSize();
  if (!n) return;
  const size_t ic = size_t((n - 1) * float(std::min<float>(std::max<float>(0, c), 1)));
  const size_t ig = size_t((n - 1) * float(std::min<float>(std::max<float>(0, g), 1)));
  const size_t start = (ig / 2 > ic) ? 0 : ic - (ig / 2);
  const size_t end = (ig / 2 + ic > n) ? n : ic + (ig / 2);
  for (size_t i = 0; i < start; ++i) m_color[4 * i + comp] = 0;
  for (size_t i = start; i < end; ++i) m_color[4 * i + comp] = Smoothstep(float(i - ic + (ig / 2)) / float(ig));
  for (size_t i = end; i < n; ++i) m_color[4 * i + comp] = 1;
}
void TransferFunction1D::SetStdFunction(float c, float g) {
  for (int k = 0; k < 4; ++k) SetStdFunction(c, g, k);
  ComputeNonZeroLimits();
}
void TransferFunction1D::GetByteArray(std::vector<uint8_t>& out, uint8_t used) const {   // :311-328 (truncates)
  out.resize(m_color.size());
  for (size_t i = 0; i < m_color.size(); ++i)
    out[i] = (uint8_t)(std::max(0.0f, std::min(m_color[i], 1.0f)) * used);
}
void TransferFunction1D::Resample(size_t iTargetSize) {                      // TransferFunction1D.cpp:198-231
  size_t iSourceSize = GetSize();
  if (iTargetSize == iSourceSize) return;
  std::vector<float> tmp(4 * iTargetSize, 0.0f);
  const size_t n = iTargetSize;
  if (iTargetSize < iSourceSize) {
    // the reference shrinks BOTH counters as it goes, so that the bucket width adapts to what is left
    size_t iFrom = 0;
    for (size_t i = 0; i < n; ++i) {
      const size_t iTo = iFrom + iSourceSize / iTargetSize;
      for (int c = 0; c < 4; ++c) {
        float acc = 0.0f;
        for (size_t j = iFrom; j < iTo; ++j) acc += m_color[4 * j + c];
        tmp[4 * i + c] = acc / float(iTo - iFrom);
      }
      iTargetSize -= 1;
      iSourceSize -= iTo - iFrom;
      iFrom = iTo;
    }
  } else {
    for (size_t i = 0; i < n; ++i) {
      const float fPos = float(i) * float(iSourceSize - 1) / float(iTargetSize);
      const size_t iFloor = size_t(std::floor(fPos));
      const size_t iCeil = std::min(iFloor + 1, iSourceSize - 1);
      const float fInterp = fPos - float(iFloor);
      for (int c = 0; c < 4; ++c) tmp[4 * i + c] = m_color[4 * iFloor + c] * (1 - fInterp) + m_color[4 * iCeil + c] * fInterp;
    }
  }
  m_color.swap(tmp);
  ComputeNonZeroLimits();
}

void ResampleHistogram1D(const std::vector<uint64_t>& in, size_t target, std::vector<uint32_t>& out) {   // uvfDataset.cpp:747-800
  out.assign(target, 0u);
  if (!target) return;
  if (target == in.size()) { for (size_t i = 0; i < target; ++i) out[i] = uint32_t(in[i]); return; }
  const float sampleFactor = static_cast<float>(in.size()) / static_cast<float>(target);
  float accWeight = 0.0f, currWeight = 1.0f, accValue = 0.0f;
  size_t j = 0;
  bool bLast = false;
  for (size_t i = 0; i < in.size(); i++) {
    if (bLast) {
      out[j] = uint32_t(accValue);
      currWeight = 1.0f - currWeight;
      j++;
      i--;
      bLast = false;
      accValue = 0;
      accWeight = 0;
    } else {
      if (sampleFactor - accWeight > 1) currWeight = 1.0f;
      else { currWeight = sampleFactor - accWeight; bLast = true; }
    }
    accValue += static_cast<float>(in[i]) * currWeight;
    accWeight += currWeight;
    if (j == target - 1) break;
  }
}

void TransferFunction1D::ComputeNonZeroLimits() {                                        // :362-370
  m_lo = GetSize(); m_hi = 0;
  for (size_t i = 0; i < GetSize(); ++i)
    if (m_color[4 * i + 3] != 0) { m_lo = std::min<uint64_t>(m_lo, i); m_hi = i; }
}

// ---------------------------------------------------------------------------- BrickedVolume
uint32_t BrickedVolume::LargestSingleBrickLOD() const {
  for (uint32_t l = 0; l < info.lod_count; ++l)
    if ((uint64_t)info.lod_bricks[l][0] * info.lod_bricks[l][1] * info.lod_bricks[l][2] == 1) return l;
  return info.lod_count - 1;
}
void BrickedVolume::MaxUsedBrickSizes(uint32_t out[3]) const {
  out[0] = out[1] = out[2] = 0;
  for (size_t i = 0; i < sizes.size() / 3; ++i) for (int a = 0; a < 3; ++a) out[a] = std::max(out[a], sizes[3 * i + a]);
}
uint64_t BrickedVolume::TocIndex(uint32_t x, uint32_t y, uint32_t z, uint32_t lod) const {
  return info.lod_offset[lod] + x + (uint64_t)y * info.lod_bricks[lod][0] + (uint64_t)z * info.lod_bricks[lod][0] * info.lod_bricks[lod][1];
}
void BrickedVolume::Range(double& lo, double& hi) const {
  const uint64_t n0 = (uint64_t)info.lod_bricks[0][0] * info.lod_bricks[0][1] * info.lod_bricks[0][2];
  lo = minmax[0]; hi = minmax[1];
  for (uint64_t i = 1; i < n0; ++i) { lo = std::min(lo, minmax[2 * i]); hi = std::max(hi, minmax[2 * i + 1]); }
}

// ---------------------------------------------------------------------------- BrickPool
void BrickPool::ComputePoolCapacity(const BrickedVolume& v, uint64_t gpuMem, uint32_t max3d, uint32_t cap[3]) {
  uint32_t bs[3];
  v.MaxUsedBrickSizes(bs);
  ComputePoolCapacity(bs, v.bpv, v.info.total_bricks, gpuMem, max3d, cap);
}

void BrickPool::ComputePoolCapacity(const uint32_t bs[3], uint32_t bpv, uint64_t totalBricks, uint64_t gpuMem, uint32_t max3d,
                                    uint32_t cap[3]) {
  // the layout the GPU memory allows (GPUMemMan.cpp:780-803)
  const uint64_t maxVoxels = gpuMem / bpv;
  const uint64_t r3 = (uint64_t)std::pow((double)maxVoxels, 1.0 / 3.0);
  uint64_t gx = uint64_t(((float)r3 / bs[0]) + 0.5f) * bs[0];
  if (gx > max3d) gx = (max3d / bs[0]) * bs[0];
  if (gx < bs[0]) gx = bs[0];
  uint64_t gy = ((maxVoxels / (gx * gx)) / bs[1]) * bs[1];
  if (gy > max3d) gy = (max3d / bs[1]) * bs[1];
  if (gy < bs[1]) gy = bs[1];
  uint64_t gz = ((maxVoxels / (gx * gy)) / bs[2]) * bs[2];
  if (gz > max3d) gz = (max3d / bs[2]) * bs[2];
  if (gz < bs[2]) gz = bs[2];
  // the layout the dataset needs (:805-826)
  const uint64_t nb = totalBricks;                         // GetTotalBrickCount: all LODs of the file
  const uint64_t r3b = (uint64_t)std::pow((double)nb, 1.0 / 3.0);
  uint64_t dx = (uint64_t)bs[0] * std::max<uint64_t>(r3b, 1);
  if (dx > max3d) dx = (max3d / bs[0]) * bs[0];
  uint64_t dy = bs[1] * (uint64_t)std::ceil(float(nb) / ((dx / bs[0]) * (dx / bs[0])));
  if (dy > max3d) dy = (max3d / bs[1]) * bs[1];
  uint64_t dz = bs[2] * (uint64_t)std::ceil(float(nb) / ((dx / bs[0]) * (dy / bs[1])));
  if (dz > max3d) dz = (max3d / bs[2]) * bs[2];
  const bool useDataset = (dx * dy * dz) < (gx * gy * gz);
  cap[0] = (uint32_t)((useDataset ? dx : gx) / bs[0]);
  cap[1] = (uint32_t)((useDataset ? dy : gy) / bs[1]);
  cap[2] = (uint32_t)((useDataset ? dz : gz) / bs[2]);
}

int BrickPool::Create(const uint32_t capacity[3], uint32_t sampler) {
  tvk_pool_desc d{};
  m_vol->MaxUsedBrickSizes(m_maxBrick);
  for (int a = 0; a < 3; ++a) { d.volume_size[a] = m_vol->domain[a]; d.max_brick_size[a] = m_maxBrick[a]; d.pool_capacity[a] = m_cap[a] = capacity[a]; }
  d.overlap = m_vol->overlap;
  d.bytes_per_voxel = m_vol->bpv;
  d.lod_count = m_lodCount = m_vol->LargestSingleBrickLOD() + 1;
  d.sampler = sampler;
  int rc = tvk_pool_create(m_ctx, &d);
  if (rc) return rc;
  rc = tvk_pool_total_bricks(m_ctx, &m_total);
  if (rc) return rc;
  // GetBrickLayout / LoD offset table (GLVolumePool.cpp:109-117,222-230)
  m_lodOffset.assign(m_lodCount, 0);
  m_lodBricks.assign(3 * m_lodCount, 0);
  uint32_t off = 0;
  for (uint32_t l = 0; l < m_lodCount; ++l) {
    for (int a = 0; a < 3; ++a) {
      const uint32_t inner = m_maxBrick[a] - 2 * m_vol->overlap;
      const uint32_t base = (uint32_t)std::ceil((double)m_vol->domain[a] / inner);
      m_lodBricks[3 * l + a] = (uint32_t)std::ceil((double)base / (1u << l));
    }
    m_lodOffset[l] = off;
    off += m_lodBricks[3 * l] * m_lodBricks[3 * l + 1] * m_lodBricks[3 * l + 2];
  }
  // m_vMinMaxScalar: dataset min/max in pool index order (:232-239)
  m_minmax.assign(2 * (size_t)m_total, 0.0);
  for (uint32_t i = 0; i < m_total; ++i) {
    uint32_t c[4];
    BrickCoords(i, c);
    if (c[0] >= m_vol->info.lod_bricks[c[3]][0] || c[1] >= m_vol->info.lod_bricks[c[3]][1] || c[2] >= m_vol->info.lod_bricks[c[3]][2])
      return TVK_ERR_UNSUPPORTED;      // pool layout and octree layout disagree (reference would mis-index)
    const uint64_t t = m_vol->TocIndex(c[0], c[1], c[2], c[3]);
    m_minmax[2 * i] = m_vol->minmax[2 * t];
    m_minmax[2 * i + 1] = m_vol->minmax[2 * t + 1];
  }
  rc = tvk_visibility_set_minmax(m_ctx, m_minmax.data(), nullptr, m_total);
  if (rc) return rc;
  const uint32_t nslots = capacity[0] * capacity[1] * capacity[2];
  m_slots.assign(nslots, PoolSlot());
  for (uint32_t i = 0; i < nslots; ++i) m_slots[i].pos = i;
  m_order = ORDER_TOUCHED; m_head = 0; m_movable = 0;
  m_time = 2; m_insertPos = 0; m_uploadedBytes = 0;
  return TVK_OK;
}

uint32_t BrickPool::BrickIndex(uint32_t x, uint32_t y, uint32_t z, uint32_t lod) const {
  return m_lodOffset[lod] + x + y * m_lodBricks[3 * lod] + z * m_lodBricks[3 * lod] * m_lodBricks[3 * lod + 1];
}
void BrickPool::BrickCoords(uint32_t index, uint32_t c[4]) const {
  uint32_t lod = (uint32_t)(std::upper_bound(m_lodOffset.begin(), m_lodOffset.end(), index) - m_lodOffset.begin()) - 1;
  const uint32_t bx = m_lodBricks[3 * lod], by = m_lodBricks[3 * lod + 1];
  index -= m_lodOffset[lod];
  c[0] = index % bx; c[1] = (index % (bx * by)) / bx; c[2] = index / (bx * by); c[3] = lod;
}
bool BrickPool::IsBrickResident(uint32_t x, uint32_t y, uint32_t z, uint32_t lod) const {
  const int32_t id = (int32_t)BrickIndex(x, y, z, lod);
  for (const PoolSlot& s : m_slots) if (s.brick == id) return true;
  return false;
}

void BrickPool::PrepareForPaging() {
  // std::sort by m_iTimeOfCreation (:946-959); stable here so that ties are deterministic.  After a paging pass the
  // order is [the m_insertPos slots just filled, ascending][the rest, ascending]: moving the new ones to the back IS the
  // stable sort in the common case (no slot flagged empty or restored in between).  The pinned coarsest brick (time
  // UINT64_MAX) sits at the very end and stays there.  The slots in front of it are kept as a RING (SlotAt): moving the
  // first m_insertPos of them to the back is advancing its head -- no slot moves (rotating 48k slots took 0.3 ms per
  // sub-frame, in the gap between a march and the upload it asked for).  Anything else -- a slot flagged empty or restored,
  // a reset, the pinned brick arriving -- puts the ring back at the vector's start and sorts.
  auto by_time = [](const PoolSlot& a, const PoolSlot& b) { return a.time < b.time; };
  if (m_order == ORDER_SORTED) { m_insertPos = 0; return; }
  const size_t movable = m_slots.size() - ((!m_slots.empty() && m_slots.back().time == UINT64_MAX) ? 1 : 0);
  if (m_order == ORDER_INSERTED && movable == m_movable && m_insertPos <= movable) {
    if (movable) m_head = (m_head + m_insertPos) % movable;
  } else {
    if (m_head && m_head < m_movable) std::rotate(m_slots.begin(), m_slots.begin() + m_head, m_slots.begin() + m_movable);
    m_head = 0;
    m_movable = movable;
    if (m_insertPos > 0 && m_insertPos < movable) std::rotate(m_slots.begin(), m_slots.begin() + m_insertPos, m_slots.begin() + movable);
    if (!std::is_sorted(m_slots.begin(), m_slots.end(), by_time)) std::stable_sort(m_slots.begin(), m_slots.end(), by_time);
  }
  m_insertPos = 0;
  m_order = ORDER_SORTED;
}

int BrickPool::Reset() {
  // GLVolumePool::PH_Reset (:261-284): every slot unused again, page table all MISSING, then UploadFirstBrick; the
  // device pool itself is kept (slots are overwritten as bricks arrive)
  // (slot positions are a permutation of 0..n-1: writing them out in order IS the sort by position)
  for (size_t i = 0; i < m_slots.size(); ++i) { m_slots[i] = PoolSlot(); m_slots[i].pos = (uint32_t)i; }
  m_time = 2; m_insertPos = 0; m_uploadedBytes = 0; m_order = ORDER_TOUCHED; m_head = 0; m_movable = 0;
  int rc = tvk_pool_clear_metadata(m_ctx);                  // page table all MISSING, on the stream (no host wait)
  if (rc) return rc;
  return UploadFirstBrick(false);
}

int BrickPool::UploadFirstBrick(bool wait) {
  // coarsest brick -> last slot, time = UINT64_MAX so that it is never evicted (:763-766)
  const uint32_t idx = m_lodOffset[m_lodCount - 1];
  uint32_t c[4];
  BrickCoords(idx, c);
  const uint64_t t = m_vol->TocIndex(c[0], c[1], c[2], c[3]);
  if (m_head && m_head < m_movable) std::rotate(m_slots.begin(), m_slots.begin() + m_head, m_slots.begin() + m_movable);   // ring back to the start
  m_head = 0;
  PoolSlot& s = m_slots[m_slots.size() - 1];
  s.brick = (int32_t)idx; s.time = UINT64_MAX; m_order = ORDER_TOUCHED;
  const void* src = m_vol->host_store ? (const void*)(m_vol->host_store + m_vol->offsets[t])
                                      : (const void*)((const uint8_t*)m_vol->device_store + m_vol->offsets[t]);
  int rc = m_vol->lz4_store ? tvk_pool_upload_bricks_compressed(m_ctx, m_vol->store_codec, 1, &s.pos, &src, &m_vol->comp_len[t], &m_vol->sizes[3 * t])
                            : tvk_pool_upload_bricks(m_ctx, 1, &s.pos, &src, &m_vol->sizes[3 * t], m_vol->host_store ? 2 : 1);   // 2: the pinned store is one allocation
  if (rc) return rc;
  rc = wait ? tvk_pool_upload_wait(m_ctx) : tvk_pool_upload_fence(m_ctx);    // fence: the next march is ordered behind it
  if (rc) return rc;
  const uint32_t val = s.pos + TVK_BI_FLAG_COUNT;
  return tvk_pool_patch_metadata(m_ctx, 1, &idx, &val);
}

int BrickPool::UploadBricks(const std::vector<uint32_t>& xyzl, uint32_t& paged, bool wait) {
  paged = 0;
  const size_t n = xyzl.size() / 4;
  m_lastMs[0] = m_lastMs[1] = m_lastMs[2] = m_lastMs[3] = 0.0;
  if (!n) return TVK_OK;
  typedef std::chrono::steady_clock clk;
  auto ms_since = [](clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); };
  clk::time_point t0 = clk::now();
  PrepareForPaging();
  m_lastMs[0] = ms_since(t0);
  t0 = clk::now();
  std::vector<uint32_t> slot, sizes, pidx, pval, clen;
  std::vector<const void*> src;
  for (size_t i = 0; i < n; ++i) {
    // "in this frame we already replaced all bricks (except the single low-res brick)" (:768-778)
    if (m_insertPos >= m_slots.size() - 1) break;
    const uint32_t* b = &xyzl[4 * i];
    if (b[3] >= m_lodCount || b[0] >= m_lodBricks[3 * b[3]] || b[1] >= m_lodBricks[3 * b[3] + 1] || b[2] >= m_lodBricks[3 * b[3] + 2])
      return TVK_ERR_INVALID;
    const uint32_t idx = BrickIndex(b[0], b[1], b[2], b[3]);
    PoolSlot& s = SlotAt(m_insertPos);
    if (s.ContainsVisibleBrick()) { pidx.push_back((uint32_t)s.brick); pval.push_back(TVK_BI_MISSING); }   // evicted
    s.brick = (int32_t)idx;
    s.time = m_time++;
    if (m_order == ORDER_SORTED) m_order = ORDER_INSERTED;
    pidx.push_back(idx); pval.push_back(s.pos + TVK_BI_FLAG_COUNT);
    const uint64_t t = m_vol->TocIndex(b[0], b[1], b[2], b[3]);
    slot.push_back(s.pos);
    src.push_back(m_vol->host_store ? (const void*)(m_vol->host_store + m_vol->offsets[t])
                                    : (const void*)((const uint8_t*)m_vol->device_store + m_vol->offsets[t]));
    for (int a = 0; a < 3; ++a) sizes.push_back(m_vol->sizes[3 * t + a]);
    if (m_vol->lz4_store) clen.push_back(m_vol->comp_len[t]);
    m_uploadedBytes += (uint64_t)m_vol->sizes[3 * t] * m_vol->sizes[3 * t + 1] * m_vol->sizes[3 * t + 2] * m_vol->bpv;
    ++m_insertPos;
    ++paged;
  }
  m_lastMs[1] = ms_since(t0);
  t0 = clk::now();
  // compressed store: the LZ4 bytes cross PCIe and the device inflates them (blocks until placed)
  int rc = m_vol->lz4_store
               ? tvk_pool_upload_bricks_compressed(m_ctx, m_vol->store_codec, (uint32_t)slot.size(), slot.data(), src.data(), clen.data(), sizes.data())
               : tvk_pool_upload_bricks(m_ctx, (uint32_t)slot.size(), slot.data(), src.data(), sizes.data(), m_vol->host_store ? 2 : 1);
  if (rc) return rc;
  rc = wait ? tvk_pool_upload_wait(m_ctx) : tvk_pool_upload_fence(m_ctx);
  if (rc) return rc;
  m_lastMs[2] = ms_since(t0);
  t0 = clk::now();
  // a slot can be evicted and re-filled within one pass only across passes, so later patches win
  rc = tvk_pool_patch_metadata(m_ctx, (uint32_t)pidx.size(), pidx.data(), pval.data());
  m_lastMs[3] = ms_since(t0);
  // The sort the NEXT pass starts with, done now: the transfer just issued keeps the device busy, whereas at the start of
  // the next pass the device would sit idle behind it.  Same order either way.
  if (!rc && !wait) PrepareForPaging();
  return rc;
}

bool BrickPool::ContainsData(const tvk_visibility& v, uint32_t i) const {     // :963-989
  const double lo = m_minmax[2 * i], hi = m_minmax[2 * i + 1];
  switch (v.render_mode) {
    case TVK_RM_1DTRANS: return v.smax >= lo && v.smin <= hi;
    case TVK_RM_2DTRANS: return (v.smax >= lo && v.smin <= hi) && (v.gmax >= -DBL_MAX && v.gmin <= DBL_MAX);
    default: return v.smin >= lo && v.smin <= hi;
  }
}

int BrickPool::RecomputeVisibility(const tvk_visibility& vis, uint32_t counts[4]) {
  // RecomputeVisibilityForBrickPool (:992-1018): slot bookkeeping on the host; the resident
  // entries are (re)written so that a brick flagged EMPTY earlier gets its slot back
  std::vector<uint32_t> pidx, pval;
  for (PoolSlot& s : m_slots) {
    if (!s.WasEverUsed()) continue;
    const bool cd = ContainsData(vis, (uint32_t)s.brick), had = s.ContainsVisibleBrick();
    if (cd) { if (!had) s.Restore(); pidx.push_back((uint32_t)s.brick); pval.push_back(s.pos + TVK_BI_FLAG_COUNT); }
    else { if (had) s.FlagEmpty(); pidx.push_back((uint32_t)s.brick); pval.push_back(TVK_BI_EMPTY); }
  }
  m_order = ORDER_TOUCHED;
  int rc = tvk_pool_patch_metadata(m_ctx, (uint32_t)pidx.size(), pidx.data(), pval.data());
  if (rc) return rc;
  // RecomputeVisibilityForOctree (:1021-1359) on the device
  return tvk_visibility_recompute(m_ctx, &vis, counts);
}

static void free_volume(tvk_ctx* ctx, BrickedVolume*& v) {
  if (!v) return;
  if (ctx) tvk_device_synchronize(ctx);          // an upload in flight may still read the pinned store
  if (v->classic_store) cudaFree(v->classic_store);
  if (v->host_store) cudaFreeHost(v->host_store);
  delete v;
  v = nullptr;
}

// ---------------------------------------------------------------------------- CudaGridLeaper
CudaGridLeaper::CudaGridLeaper() : m_tf1d(256) { m_tf1d.SetStdFunction(); }
CudaGridLeaper::~CudaGridLeaper() { Cleanup(); }

bool CudaGridLeaper::Fail(const std::string& what, int rc) {
  m_err = what + " (status " + std::to_string(rc) + (m_ctx ? std::string("): ") + tvk_last_error(m_ctx) : ")");
  return false;
}

bool CudaGridLeaper::Initialize(int device) {
  if (m_ctx) return true;
  const int rc = tvk_create(device, &m_ctx);
  if (rc) { m_ctx = nullptr; return Fail("tvk_create", rc); }
  return true;
}

void CudaGridLeaper::Cleanup() {
  delete m_pool; m_pool = nullptr;
  free_volume(m_ctx, m_vol);
  free_volume(m_ctx, m_streamVol);
  if (m_ctx) { tvk_destroy(m_ctx); m_ctx = nullptr; }
}

// ---------------------------------------------------------------------------- sort-last members
bool CudaGridLeaper::SortLastConfigure(uint32_t members, uint32_t member, const uint32_t g[3], const uint32_t lo[3], const uint32_t hi[3],
                                       const uint32_t start[3], double globalMaxValue) {
  if (!m_vol) { m_err = "sort-last: load the box first"; return false; }
  if (!members || member >= members) { m_err = "sort-last: bad member index"; return false; }
  for (int a = 0; a < 3; ++a)
    if (!(start[a] <= lo[a] && lo[a] < hi[a] && hi[a] <= start[a] + m_vol->domain[a] && start[a] + m_vol->domain[a] <= g[a])) {
      m_err = "sort-last: the owned box must lie inside the loaded box, and that inside the global volume"; return false;
    }
  m_sl = SortLast();
  m_sl.on = true; m_sl.members = members; m_sl.member = member; m_sl.maxValue = globalMaxValue;
  m_visValid = false;
  for (int a = 0; a < 3; ++a) { m_sl.global[a] = g[a]; m_sl.lo[a] = lo[a]; m_sl.hi[a] = hi[a]; m_sl.start[a] = start[a]; }
  ScheduleCompleteRedraw();
  return true;
}

bool CudaGridLeaper::SortLastExchangeCreate(void* handle64) {
  if (!m_sl.on || !m_w || !m_h) { m_err = "sort-last: configure and resize first"; return false; }
  const int rc = tvk_exchange_create(m_ctx, m_sl.members, m_sl.member, m_w, m_h, handle64);
  if (rc) return Fail("tvk_exchange_create", rc);
  m_sl.exchange = false;
  return true;
}

bool CudaGridLeaper::SortLastExchangeConnect(const void* handles, void* const* bases) {
  if (!m_sl.on) { m_err = "sort-last: not configured"; return false; }
  const int rc = tvk_exchange_connect(m_ctx, handles, bases);
  if (rc) return Fail("tvk_exchange_connect", rc);
  m_sl.exchange = true;
  return true;
}

bool CudaGridLeaper::SortLastBeginFrame() {
  if (!m_sl.on || !m_sl.exchange) { m_err = "sort-last: exchange not connected"; return false; }
  m_sl.frame++;
  return true;
}

bool CudaGridLeaper::SortLastComposite(const uint32_t* order) {
  if (!m_sl.on || !m_sl.exchange || !m_sl.frame || !order) { m_err = "sort-last: nothing to composite"; return false; }
  const int rc = tvk_exchange_composite(m_ctx, order, m_sl.frame);
  return rc ? Fail("tvk_exchange_composite", rc) : true;
}

// sortlast.depth_order: along every axis a ray moves monotonically, so sorting the boxes by the per-axis distances of their
// intervals from the eye (a tuple, lexicographic) is a linear extension of the visibility order of a regular grid of boxes
void CudaGridLeaper::SortLastDepthOrder(const uint32_t* b, uint32_t* order) {
  float ext[3], mx = 0;
  for (int a = 0; a < 3; ++a) { ext[a] = float(m_sl.global[a]) * (m_vol ? m_vol->scale[a] : 1.0f); mx = std::max(mx, ext[a]); }
  for (int a = 0; a < 3; ++a) ext[a] /= mx;
  const Matrix4 mv = m_rotation * m_translation * Matrix4::LookAt(m_eye, m_at, m_up);
  const Matrix4 emm = mv.inverse() * Matrix4::Scaling(1.0f / ext[0], 1.0f / ext[1], 1.0f / ext[2]) * Matrix4::Translation(0.5f, 0.5f, 0.5f);
  const float o4[4] = {0, 0, 0, 1};
  float eye[4];
  emm.transform(o4, eye);                                 // the eye in global normalised coordinates
  struct Key { double k[3]; uint32_t r; };
  std::vector<Key> keys(m_sl.members);
  for (uint32_t r = 0; r < m_sl.members; ++r) {
    keys[r].r = r;
    for (int a = 0; a < 3; ++a) {
      const double lo = double(b[6 * r + a]) / m_sl.global[a], hi = double(b[6 * r + 3 + a]) / m_sl.global[a];
      keys[r].k[a] = std::max(std::max(lo - double(eye[a]), 0.0), double(eye[a]) - hi);
    }
  }
  std::sort(keys.begin(), keys.end(), [](const Key& x, const Key& y) {
    for (int a = 0; a < 3; ++a) if (x.k[a] != y.k[a]) return x.k[a] < y.k[a];
    return x.r < y.r;
  });
  for (uint32_t i = 0; i < m_sl.members; ++i) order[i] = keys[i].r;
}

bool CudaGridLeaper::SortLastReadFrameRGBA8(uint8_t* out, bool flipRows) {
  if (!m_sl.on || !m_sl.exchange || !m_sl.frame || m_sl.member != 0) { m_err = "sort-last: the finished frame lands on member 0"; return false; }
  const int rc = tvk_exchange_read_final_rgba8(m_ctx, m_sl.frame, out, flipRows ? 1 : 0);
  return rc ? Fail("tvk_exchange_read_final_rgba8", rc) : true;
}

bool CudaGridLeaper::LoadStreamedBegin(const uint32_t size[3], uint32_t bpv, const float scale[3], uint32_t brickSize, uint32_t overlap,
                                       bool median) {
  if (!m_ctx) { m_err = "renderer not initialised"; return false; }
  delete m_pool; m_pool = nullptr;
  free_volume(m_ctx, m_vol);
  if (m_streamVol) free_volume(m_ctx, m_streamVol);
  tvk_build_release(m_ctx);
  tvk_build_desc& d = m_streamDesc;
  d = tvk_build_desc{};
  for (int a = 0; a < 3; ++a) d.volume_size[a] = size[a];
  d.brick_size = brickSize; d.overlap = overlap; d.bytes_per_voxel = bpv; d.use_median = median ? 1 : 0;
  BrickedVolume* v = new BrickedVolume();
  for (int a = 0; a < 3; ++a) { v->domain[a] = size[a]; v->scale[a] = scale ? scale[a] : 1.0f; }
  v->bpv = bpv; v->brick_size = brickSize; v->overlap = overlap;
  int rc = tvk_build_plan(&d, &v->info);
  if (rc) { delete v; return Fail("tvk_build_plan", rc); }
  // the store's size is the sum of the 256-byte aligned bricks: learn it from a ToC-only begin (no store yet)
  const uint64_t n = v->info.total_bricks;
  const uint32_t core = brickSize - 2 * overlap;
  uint64_t bytes = 0;
  for (uint32_t l = 0; l < v->info.lod_count; ++l)
    for (uint32_t z = 0; z < v->info.lod_bricks[l][2]; ++z)
      for (uint32_t y = 0; y < v->info.lod_bricks[l][1]; ++y)
        for (uint32_t x = 0; x < v->info.lod_bricks[l][0]; ++x) {
          const uint32_t c[3] = {x, y, z};
          uint64_t vox = 1;
          for (int a = 0; a < 3; ++a) {
            const bool last = c[a] >= v->info.lod_bricks[l][a] - 1;
            const uint32_t rem = v->info.lod_size[l][a] % core;
            vox *= (last && rem) ? (2 * overlap + rem) : brickSize;
          }
          bytes += (vox * bpv + 255) / 256 * 256;
        }
  v->store_bytes = bytes;
  if (cudaHostAlloc((void**)&v->host_store, bytes, cudaHostAllocDefault) != cudaSuccess) {
    (void)cudaGetLastError(); delete v; m_err = "cudaHostAlloc failed for the out-of-core brick store"; return false;
  }
  rc = tvk_build_stream_begin(m_ctx, &d, v->host_store, &v->info);
  if (rc || v->info.total_bytes != bytes) {
    cudaFreeHost(v->host_store); delete v;
    if (!rc) { m_err = "store size mismatch"; return false; }
    return Fail("tvk_build_stream_begin", rc);
  }
  (void)n;
  m_streamVol = v;
  m_lastBuildMs = 0.0f;
  return true;
}

bool CudaGridLeaper::LoadStreamedRows(uint32_t bz0, uint32_t bz1, uint32_t& z0, uint32_t& rows) const {
  return m_streamVol && tvk_build_stream_rows(&m_streamDesc, bz0, bz1, &z0, &rows) == TVK_OK;
}

bool CudaGridLeaper::LoadStreamedSlab(const void* slab, uint32_t z0, uint32_t rows, uint32_t bz0, uint32_t bz1) {
  if (!m_streamVol) { m_err = "LoadStreamedSlab without LoadStreamedBegin"; return false; }
  const int rc = tvk_build_stream_slab(m_ctx, slab, z0, rows, bz0, bz1);
  return rc ? Fail("tvk_build_stream_slab", rc) : true;
}

bool CudaGridLeaper::LoadStreamedEnd(uint64_t gpuMemBytes) {
  if (!m_streamVol) { m_err = "LoadStreamedEnd without LoadStreamedBegin"; return false; }
  BrickedVolume* v = m_streamVol;
  int rc = tvk_build_stream_end(m_ctx, &v->info);
  if (rc) return Fail("tvk_build_stream_end", rc);
  tvk_device_synchronize(m_ctx);                         // every brick has landed in the pinned store
  const uint64_t n = v->info.total_bricks;
  v->offsets.resize(n); v->sizes.resize(3 * n); v->minmax.resize(2 * n);
  rc = tvk_build_get_toc(m_ctx, v->offsets.data(), v->sizes.data(), n);
  if (!rc) rc = tvk_build_get_minmax(m_ctx, v->minmax.data(), n);
  if (rc) return Fail("tvk_build_get_*", rc);
  v->hist1d.resize(1u << (8 * v->bpv));
  uint32_t hn = 0;
  rc = tvk_build_hist1d(m_ctx, v->hist1d.data(), (uint32_t)v->hist1d.size(), &hn);
  if (rc) return Fail("tvk_build_hist1d", rc);
  v->hist1d.resize(hn);
  tvk_build_last_ms(m_ctx, &m_lastBuildMs);
  tvk_build_release(m_ctx);                              // (the store is caller-owned: it stays)
  m_streamVol = nullptr;
  const uint32_t tfSize = std::max<uint32_t>(1, std::min<uint32_t>(hn, 4096));
  return FinishLoad(v, gpuMemBytes, tfSize);
}

bool CudaGridLeaper::LoadVolumeFile(const std::string& path, uint32_t brickSize, uint32_t overlap, bool median, bool outOfCore,
                                    uint64_t gpuMemBytes, bool computeHist2D) {
  if (!m_ctx) { m_err = "renderer not initialised"; return false; }
  tvk_raw_desc d;
  char err[512] = {0};
  if (tvk_raw_probe(path.c_str(), &d, err, sizeof(err))) { m_err = err; return false; }
  if (d.component_count != 1) { m_err = "multi-component (colour) volumes are not supported"; return false; }
  for (int a = 0; a < 3; ++a) if (d.size[a] == 0 || d.size[a] > 0xFFFFFFFFull) { m_err = "bad volume size"; return false; }
  const uint64_t n = d.size[0] * d.size[1] * d.size[2];
  const uint32_t bytes = d.component_bits / 8;
  std::vector<uint8_t> host((size_t)(n * bytes));
  if (tvk_raw_read(&d, host.data(), host.size(), err, sizeof(err))) { m_err = err; return false; }
  // element type of RAWConverter::ConvertRAWDataset's dispatcher (RAWConverter.cpp:196-352)
  uint32_t type;
  if (d.is_float) type = bytes == 4 ? TVK_DT_F32 : TVK_DT_F64;
  else if (bytes == 1) type = d.is_signed ? TVK_DT_I8 : TVK_DT_U8;
  else if (bytes == 2) type = d.is_signed ? TVK_DT_I16 : TVK_DT_U16;
  else if (bytes == 4) type = d.is_signed ? TVK_DT_I32 : TVK_DT_U32;
  else type = d.is_signed ? TVK_DT_I64 : TVK_DT_U64;
  void* d_q = nullptr;
  if (cudaMalloc(&d_q, (size_t)(2 * n + 16)) != cudaSuccess) { (void)cudaGetLastError(); m_err = "out of device memory for the quantised volume"; return false; }
  std::vector<uint64_t> hist(4096);
  tvk_quantize_result qr{};
  int rc = tvk_quantize(m_ctx, type, n, host.data(), 0, 0, d_q, hist.data(), &qr);
  if (rc) { cudaFree(d_q); return Fail("tvk_quantize", rc); }
  const uint32_t size[3] = {(uint32_t)d.size[0], (uint32_t)d.size[1], (uint32_t)d.size[2]};
  bool ok;
  if (qr.generated) {
    std::vector<uint8_t>().swap(host);
    ok = LoadRawVolume(d_q, size, qr.out_bits / 8, d.aspect, brickSize, overlap, median, outOfCore, gpuMemBytes, computeHist2D, true);
  } else {
    ok = LoadRawVolume(host.data(), size, bytes, d.aspect, brickSize, overlap, median, outOfCore, gpuMemBytes, computeHist2D, false);
  }
  // (an in-core build keeps reading the flat volume for the 2-D histogram only inside LoadRawVolume: safe to free now)
  tvk_device_synchronize(m_ctx);
  tvk_build_forget_raw(m_ctx);
  cudaFree(d_q);
  return ok;
}

bool CudaGridLeaper::LoadRawVolume(const void* raw, const uint32_t size[3], uint32_t bpv, const float scale[3],
                                   uint32_t brickSize, uint32_t overlap, bool median, bool outOfCore,
                                   uint64_t gpuMemBytes, bool computeHist2D, bool rawOnDevice) {
  if (!m_ctx) { m_err = "renderer not initialised"; return false; }
  delete m_pool; m_pool = nullptr;
  free_volume(m_ctx, m_vol);
  tvk_build_desc d{};
  for (int a = 0; a < 3; ++a) d.volume_size[a] = size[a];
  d.brick_size = brickSize; d.overlap = overlap; d.bytes_per_voxel = bpv; d.use_median = median ? 1 : 0;
  BrickedVolume* v = new BrickedVolume();
  int rc = tvk_build_octree(m_ctx, &d, raw, rawOnDevice ? 1 : 0, &v->info);
  if (rc) { delete v; return Fail("tvk_build_octree", rc); }
  for (int a = 0; a < 3; ++a) { v->domain[a] = size[a]; v->scale[a] = scale ? scale[a] : 1.0f; }
  v->bpv = bpv; v->brick_size = brickSize; v->overlap = overlap;
  const uint64_t n = v->info.total_bricks;
  v->offsets.resize(n); v->sizes.resize(3 * n); v->minmax.resize(2 * n);
  rc = tvk_build_get_toc(m_ctx, v->offsets.data(), v->sizes.data(), n);
  if (!rc) rc = tvk_build_get_minmax(m_ctx, v->minmax.data(), n);
  if (rc) { delete v; return Fail("tvk_build_get_*", rc); }
  v->hist1d.resize(1u << (8 * bpv));
  uint32_t hn = 0;
  rc = tvk_build_hist1d(m_ctx, v->hist1d.data(), (uint32_t)v->hist1d.size(), &hn);
  if (rc) { delete v; return Fail("tvk_build_hist1d", rc); }
  v->hist1d.resize(hn);
  // RAWConverter caps the 1-D histogram (= TF size) at 4096 bins (Quantize.h:429-578)
  uint32_t tfSize = std::max<uint32_t>(1, std::min<uint32_t>(hn, 4096));
  if (computeHist2D) {
    double lo, hi;
    v->Range(lo, hi);
    v->hist2d.resize((size_t)tfSize * 256);
    rc = tvk_build_hist2d(m_ctx, tfSize, hi, v->hist2d.data(), &v->max_grad_magnitude);
    if (rc) { delete v; return Fail("tvk_build_hist2d", rc); }
    v->hist2d_x = tfSize; v->hist2d_y = 256;
  }
  v->store_bytes = v->info.total_bytes;
  if (outOfCore) {
    if (cudaHostAlloc((void**)&v->host_store, v->store_bytes, cudaHostAllocDefault) != cudaSuccess) {
      delete v; m_err = "cudaHostAlloc failed for the out-of-core brick store"; return false;
    }
    rc = tvk_build_get_bricks(m_ctx, v->host_store, 0, v->store_bytes);
    if (rc) { cudaFreeHost(v->host_store); delete v; return Fail("tvk_build_get_bricks", rc); }
    tvk_build_release(m_ctx);
  } else {
    rc = tvk_build_store_device_ptr(m_ctx, &v->device_store);
    if (rc) { delete v; return Fail("tvk_build_store_device_ptr", rc); }
  }
  return FinishLoad(v, gpuMemBytes, tfSize);
}

// pool set-up + AbstrRenderer::LoadFile tail, shared by the raw-volume and the UVF loaders
bool CudaGridLeaper::FinishLoad(BrickedVolume* v, uint64_t gpuMemBytes, uint32_t tfSize) {
  const uint32_t bpv = v->bpv;
  int rc;
  m_vol = v;
  // pool sizing: dataset-sized unless a GPU memory budget is given (GPUMemMan.cpp:766-844)
  uint32_t cap[3];
  uint32_t bs[3];
  v->MaxUsedBrickSizes(bs);
  const uint64_t need = (uint64_t)bs[0] * bs[1] * bs[2] * bpv * (v->info.total_bricks + 8);
  BrickPool::ComputePoolCapacity(*v, gpuMemBytes ? gpuMemBytes : need, 16384, cap);
  m_pool = new BrickPool(m_ctx, v);
  uint32_t sampler = m_sampler;
  if (sampler == kSamplerAuto) {
    // the layered-array pool read with four-texel gathers: exact filter arithmetic, two fetches per sample and 1x the
    // memory (the quad-packed pool needs 4x and turns memory-latency bound on large volumes)
    sampler = TVK_SAMPLER_GATHER;
  }
  m_activeSampler = sampler;
  rc = m_pool->Create(cap, sampler);
  if (rc) return Fail("pool create", rc);
  rc = m_pool->UploadFirstBrick();
  if (rc) return Fail("UploadFirstBrick", rc);
  // AbstrRenderer::LoadFile tail (:245-261): default iso value = half the range; TF sized like the histogram
  double lo, hi;
  v->Range(lo, hi);
  m_isoValue = float(hi - lo) / 2.0f;
  m_cvIsoValue = float(m_isoValue + m_isoValue / 2.0);       // AbstrRenderer.cpp:261
  m_tf1d.Resize(tfSize);
  m_tf1d.SetStdFunction();
  m_visValid = false;
  Changed1DTrans();
  ScheduleCompleteRedraw();
  return true;
}

// renderer.loadDataset(file.uvf): UVFDataset::Open (uvfDataset.cpp:83-216) through the reader of tvk_uvf.h.
// The bricks are read (and inflated) once into pinned host memory, which is the pager's source.
bool CudaGridLeaper::LoadDataset(const std::string& uvfFile, uint64_t gpuMemBytes, bool verifyChecksum, bool keepCompressed) {
  if (!m_ctx) { m_err = "renderer not initialised"; return false; }
  char err[256] = {0};
  tvk_uvf* u = nullptr;
  if (tvk_uvf_open_timestep(uvfFile.c_str(), verifyChecksum ? 1 : 0, (uint32_t)m_loadTimestep, &u, err, sizeof(err))) { m_err = std::string("UVF open failed: ") + err; return false; }
  tvk_uvf_info ui;
  tvk_uvf_get_info(u, &ui);
  // what SetTimestep needs to come back for another time step of the same file
  m_file = uvfFile; m_fileGpuMem = gpuMemBytes; m_fileVerify = verifyChecksum; m_fileKeepCompressed = keepCompressed;
  m_fileRebrick = m_rebrickSize; m_fileRebrickMinMax = m_rebrickMinMax;
  m_timesteps = std::max<uint32_t>(1u, ui.timesteps); m_timestep = m_loadTimestep;
  if (ui.toc_block < 0 && ui.rdb_block >= 0 && !ui.rdb_signed) {
    // legacy RasterDataBlock file (e.g. DemoData/c60.uvf: one 128^3 brick): stitch the finest level back together
    // and re-brick it for the ray-guided path, as IOManager::LoadRebrickedDataset / DynamicBrickingDS do
    // (IOManager.cpp:1281-1320, IO/DynamicBrickingDS.cpp) -- here by running the K2 builder on the volume
    const uint32_t size[3] = {(uint32_t)ui.rdb_domain[0], (uint32_t)ui.rdb_domain[1], (uint32_t)ui.rdb_domain[2]};
    const uint32_t bpv = ui.rdb_bits / 8;
    std::vector<uint8_t> flat((size_t)size[0] * size[1] * size[2] * bpv);
    const int rc = tvk_uvf_read_rdb_volume(u, flat.data(), flat.size());
    const std::string why = rc ? tvk_uvf_last_error(u) : "";
    tvk_uvf_close(u);
    if (rc) { m_err = "RasterDataBlock read failed: " + why; return false; }
    const float scale[3] = {1, 1, 1};
    return LoadRawVolume(flat.data(), size, bpv, scale, 64, 2, false, false, gpuMemBytes, true);
  }
  if (ui.toc_block < 0 || (ui.vol.component_type != TVK_UVF_CT_UINT8 && ui.vol.component_type != TVK_UVF_CT_UINT16) ||
      ui.maxmin_bricks != ui.total_bricks) {
    tvk_uvf_close(u);
    m_err = "UVF file has no 8/16-bit unsigned TOC block with a matching max/min block (legacy RasterDataBlock files need rebricking)";
    return false;
  }
  delete m_pool; m_pool = nullptr;
  free_volume(m_ctx, m_vol);
  tvk_build_release(m_ctx);
  BrickedVolume* v = new BrickedVolume();
  v->bpv = ui.vol.component_type == TVK_UVF_CT_UINT16 ? 2 : 1;
  v->brick_size = (uint32_t)ui.vol.brick_size[0]; v->overlap = ui.vol.overlap;
  tvk_build_desc d{};
  for (int a = 0; a < 3; ++a) { v->domain[a] = d.volume_size[a] = (uint32_t)ui.vol.volume_size[a]; v->scale[a] = (float)ui.vol.aspect[a]; }
  d.brick_size = v->brick_size; d.overlap = v->overlap; d.bytes_per_voxel = v->bpv;
  tvk_build_plan(&d, &v->info);
  const uint64_t n = ui.total_bricks;
  v->offsets.resize(n); v->sizes.resize(3 * n); v->minmax.resize(2 * n);
  tvk_uvf_get_toc(u, nullptr, nullptr, nullptr, v->sizes.data(), n);
  uint64_t off = 0;
  for (uint64_t i = 0; i < n; ++i) {          // same 256-byte brick alignment as the builder's store
    v->offsets[i] = off;
    off += ((uint64_t)v->sizes[3 * i] * v->sizes[3 * i + 1] * v->sizes[3 * i + 2] * v->bpv + 255) / 256 * 256;
  }
  v->store_bytes = v->info.total_bytes = off;
  std::vector<double> mm4(4 * n);
  tvk_uvf_get_minmax(u, mm4.data(), n);
  for (uint64_t i = 0; i < n; ++i) { v->minmax[2 * i] = mm4[4 * i]; v->minmax[2 * i + 1] = mm4[4 * i + 1]; }
  v->hist1d.resize(ui.hist1d_size);
  if (ui.hist1d_size) tvk_uvf_get_hist1d(u, v->hist1d.data(), ui.hist1d_size);
  v->max_grad_magnitude = ui.max_grad_magnitude;
  if (ui.hist2d_x && ui.hist2d_y) {                        // kept so that SaveUVF writes the file's own 2-D histogram back
    v->hist2d.resize((size_t)(ui.hist2d_x * ui.hist2d_y));
    if (tvk_uvf_get_hist2d(u, v->hist2d.data(), v->hist2d.size()) == TVK_OK) { v->hist2d_x = ui.hist2d_x; v->hist2d_y = ui.hist2d_y; }
    else v->hist2d.clear();
  }
  if (keepCompressed) {
    // all bricks LZ4, or all zlib? then the store keeps what the file holds and the device inflates (ExtendedOctree.cpp:313-360
    // does it on the host, brick by brick, before the upload)
    std::vector<uint64_t> flen(n);
    std::vector<uint32_t> comp(n);
    bool all = n > 0 && tvk_uvf_get_toc(u, nullptr, flen.data(), comp.data(), nullptr, n) == TVK_OK;
    const uint32_t codec = all ? comp[0] : 0u;
    all = all && (codec == TVK_UVF_COMP_LZ4 || codec == TVK_UVF_COMP_ZLIB);
    for (uint64_t i = 0; i < n && all; ++i) all = comp[i] == codec && flen[i] <= 0xFFFFFFFFull;
    if (all) {
      v->lz4_store = true;                                 // (the flag's name is older than the zlib decoder: "kept compressed")
      v->store_codec = codec;
      v->comp_len.resize(n);
      uint64_t coff = 0;
      for (uint64_t i = 0; i < n; ++i) { v->offsets[i] = coff; v->comp_len[i] = (uint32_t)flen[i]; coff += flen[i]; }
      v->store_bytes = coff;                               // (info.total_bytes stays the inflated size)
    }
  }
  if (cudaHostAlloc((void**)&v->host_store, v->store_bytes, cudaHostAllocDefault) != cudaSuccess) {
    tvk_uvf_close(u); delete v; m_err = "cudaHostAlloc failed for the brick store"; return false;
  }
  const int rc = v->lz4_store ? tvk_uvf_read_bricks_raw(u, 0, n, v->host_store, v->offsets.data())
                              : tvk_uvf_read_bricks(u, 0, n, v->host_store, v->offsets.data());
  if (rc) { m_err = std::string("UVF brick read failed: ") + tvk_uvf_last_error(u); tvk_uvf_close(u); cudaFreeHost(v->host_store); delete v; return false; }
  tvk_uvf_close(u);
  if (m_rebrickSize) {                                     // LoadRebricked: present the file with smaller bricks
    BrickedVolume* r = Rebrick(v, m_rebrickSize, m_rebrickMinMax);
    cudaFreeHost(v->host_store);
    delete v;
    if (!r) return false;
    v = r;
  }
  const uint32_t tfSize = std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)v->hist1d.size(), 4096));
  return FinishLoad(v, gpuMemBytes, tfSize);
}

// AbstrRenderer::LoadRebricked (AbstrRenderer.cpp:266-273) -> IOManager::LoadRebrickedDataset (IOManager.cpp:1281-1320) ->
// DynamicBrickingDS.  The reference wraps the source dataset and cuts target bricks one GetBrick() at a time; here the
// whole target store is cut by one kernel that reads the source bricks straight out of the pinned host store
// (tvk_rebrick), and the result is an ordinary paged volume.  Because an ExtendedOctree pyramid always runs down to
// 1x1x1, the target has the layout a direct build with the small brick size would have (single-brick coarsest level
// included), so both renderers take it as it is.
bool CudaGridLeaper::LoadRebricked(const std::string& uvfFile, uint32_t brickSize, int minmaxMode, uint64_t gpuMemBytes, bool verifyChecksum) {
  if (minmaxMode < TVK_MM_SOURCE || minmaxMode > TVK_MM_DYNAMIC) { m_err = "minmaxType too large"; return false; }   // IOManager.cpp:1290-1292
  if (!brickSize) { m_err = "null brick size"; return false; }
  m_rebrickSize = brickSize; m_rebrickMinMax = minmaxMode;
  const bool ok = LoadDataset(uvfFile, gpuMemBytes, verifyChecksum, false);
  m_rebrickSize = 0;
  return ok;
}

BrickedVolume* CudaGridLeaper::Rebrick(const BrickedVolume* src, uint32_t brickSize, int minmaxMode) {
  tvk_rebrick_desc d{};
  d.lod_count = src->info.lod_count;
  for (uint32_t l = 0; l < d.lod_count; ++l) for (int a = 0; a < 3; ++a) d.lod_size[l][a] = src->info.lod_size[l][a];
  for (int a = 0; a < 3; ++a) { d.src_brick[a] = src->brick_size; d.tgt_brick[a] = brickSize; }
  d.overlap = src->overlap; d.bytes_per_voxel = src->bpv;
  tvk_rebrick_info ri;
  if (tvk_rebrick_plan(&d, &ri) != TVK_OK) {
    m_err = "target brick size (" + std::to_string(brickSize) + ") without ghost voxels is not a divisor of the source brick size (" +
            std::to_string(src->brick_size) + ") without ghost voxels, or the overlap is not 2";
    return nullptr;
  }
  BrickedVolume* v = new BrickedVolume();
  for (int a = 0; a < 3; ++a) { v->domain[a] = src->domain[a]; v->scale[a] = src->scale[a]; }
  v->bpv = src->bpv; v->brick_size = ri.tgt_brick[0]; v->overlap = src->overlap;
  tvk_build_desc bd{};
  for (int a = 0; a < 3; ++a) bd.volume_size[a] = v->domain[a];
  bd.brick_size = v->brick_size; bd.overlap = v->overlap; bd.bytes_per_voxel = v->bpv;
  tvk_build_plan(&bd, &v->info);
  const uint64_t n = ri.total_bricks;
  if (v->info.total_bricks != n || v->info.lod_count != d.lod_count) { delete v; m_err = "re-bricked layout differs from the octree layout of the target brick size"; return nullptr; }
  v->offsets.resize(n); v->sizes.resize(3 * n); v->minmax.resize(2 * n);
  std::vector<uint64_t> sidx(n);
  tvk_rebrick_toc(&d, v->offsets.data(), v->sizes.data(), sidx.data(), nullptr, n);
  v->store_bytes = v->info.total_bytes = ri.total_bytes;
  v->hist1d = src->hist1d; v->max_grad_magnitude = src->max_grad_magnitude;
  v->hist2d = src->hist2d; v->hist2d_x = src->hist2d_x; v->hist2d_y = src->hist2d_y;
  void* d_dst = nullptr;
  if (cudaMalloc(&d_dst, ri.total_bytes) != cudaSuccess) { delete v; m_err = "re-bricking: no device memory for the target store"; return nullptr; }
  cudaMemset(d_dst, 0, ri.total_bytes);                     // (alignment padding between bricks)
  std::vector<double> mm(2 * n);
  int rc = tvk_rebrick(m_ctx, &d, src->host_store, src->offsets.data(), src->sizes.data(), src->info.total_bricks, d_dst, mm.data());
  if (rc == TVK_OK && cudaHostAlloc((void**)&v->host_store, v->store_bytes, cudaHostAllocDefault) != cudaSuccess) rc = TVK_ERR_NOMEM;
  if (rc == TVK_OK && cudaMemcpy(v->host_store, d_dst, v->store_bytes, cudaMemcpyDeviceToHost) != cudaSuccess) rc = TVK_ERR_CUDA;
  cudaFree(d_dst);
  if (rc) { Fail("tvk_rebrick", rc); if (v->host_store) cudaFreeHost(v->host_store); delete v; return nullptr; }
  // MaxMinForKey (DynamicBrickingDS.cpp:862-876)
  for (uint64_t i = 0; i < n; ++i) {
    if (minmaxMode == TVK_MM_SOURCE) { v->minmax[2 * i] = src->minmax[2 * sidx[i]]; v->minmax[2 * i + 1] = src->minmax[2 * sidx[i] + 1]; }
    else { v->minmax[2 * i] = mm[2 * i]; v->minmax[2 * i + 1] = mm[2 * i + 1]; }
  }
  return v;
}

bool CudaGridLeaper::PrepareHostVolume(BrickedVolume* v) {
  if (!v || (v->bpv != 1 && v->bpv != 2)) { m_err = "PrepareHostVolume: 8/16-bit unsigned volumes only"; return false; }
  tvk_build_desc d{};
  for (int a = 0; a < 3; ++a) d.volume_size[a] = v->domain[a];
  d.brick_size = v->brick_size; d.overlap = v->overlap; d.bytes_per_voxel = v->bpv;
  const int rc = tvk_build_plan(&d, &v->info);
  if (rc) return Fail("tvk_build_plan", rc);
  const uint64_t n = v->info.total_bricks;
  if (v->sizes.size() != 3 * n || v->minmax.size() != 2 * n) { m_err = "PrepareHostVolume: sizes / minmax do not match the octree's brick count"; return false; }
  v->offsets.resize(n);
  uint64_t off = 0;
  for (uint64_t i = 0; i < n; ++i) {          // same 256-byte brick alignment as the builder's store
    v->offsets[i] = off;
    off += ((uint64_t)v->sizes[3 * i] * v->sizes[3 * i + 1] * v->sizes[3 * i + 2] * v->bpv + 255) / 256 * 256;
  }
  v->store_bytes = v->info.total_bytes = off;
  if (cudaHostAlloc((void**)&v->host_store, v->store_bytes, cudaHostAllocDefault) != cudaSuccess) {
    v->host_store = nullptr; m_err = "cudaHostAlloc failed for the brick store"; return false;
  }
  return true;
}

bool CudaGridLeaper::AdoptVolume(BrickedVolume* v, uint64_t gpuMemBytes) {
  if (!m_ctx) { m_err = "renderer not initialised"; return false; }
  if (!v || !v->host_store) { m_err = "AdoptVolume: the volume has no host brick store (PrepareHostVolume first)"; return false; }
  delete m_pool; m_pool = nullptr;
  if (m_vol && m_vol != v) free_volume(m_ctx, m_vol);
  m_vol = nullptr;
  tvk_build_release(m_ctx);
  const uint32_t tfSize = std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)v->hist1d.size(), 4096));
  return FinishLoad(v, gpuMemBytes, tfSize);
}

// writes the loaded volume as a UVF v5 file stock ImageVis3D opens (tvk_uvf_write)
bool CudaGridLeaper::SaveUVF(const std::string& uvfFile, uint32_t compression, uint32_t level) {
  if (!m_ctx || !m_vol) { m_err = "no dataset"; return false; }
  if (m_vol->lz4_store) { m_err = "the dataset was loaded with its bricks kept compressed; reload it with loadDataset to save a copy"; return false; }
  const BrickedVolume& v = *m_vol;
  std::vector<uint8_t> tmp;
  const uint8_t* store = v.host_store;
  if (!store) {
    tmp.resize(v.store_bytes);
    const int rc = tvk_build_get_bricks(m_ctx, tmp.data(), 0, v.store_bytes);
    if (rc) return Fail("tvk_build_get_bricks", rc);
    store = tmp.data();
  }
  const uint32_t tfSize = std::max<uint32_t>(1, std::min<uint32_t>((uint32_t)v.hist1d.size(), 4096));
  std::vector<uint64_t> h2;
  float mg = v.max_grad_magnitude;
  if (!v.hist2d.empty()) {                    // the histogram the file (or the load) came with
    h2 = v.hist2d;
  } else if (!v.host_store) {                 // the builder state is alive: the 2-D histogram comes from it
    double lo, hi;
    v.Range(lo, hi);
    h2.resize((size_t)tfSize * 256);
    const int rc = tvk_build_hist2d(m_ctx, tfSize, hi, h2.data(), &mg);
    if (rc) return Fail("tvk_build_hist2d", rc);
  }
  tvk_uvf_write_desc w{};
  w.vol.component_type = v.bpv == 2 ? TVK_UVF_CT_UINT16 : TVK_UVF_CT_UINT8;
  w.vol.component_count = 1;
  for (int a = 0; a < 3; ++a) { w.vol.volume_size[a] = v.domain[a]; w.vol.aspect[a] = v.scale[a]; w.vol.brick_size[a] = v.brick_size; }
  w.vol.overlap = v.overlap; w.vol.compression = compression; w.vol.compression_level = level;
  w.total_bricks = v.info.total_bricks;
  w.brick_offsets = v.offsets.data(); w.brick_sizes = v.sizes.data(); w.store = store; w.minmax = v.minmax.data();
  w.hist1d = v.hist1d.data(); w.hist1d_size = v.hist1d.size();
  if (!h2.empty()) { w.hist2d = h2.data(); w.hist2d_x = v.hist2d.empty() ? tfSize : v.hist2d_x; w.hist2d_y = v.hist2d.empty() ? 256 : v.hist2d_y; }
  w.max_grad_magnitude = mg;
  const char* keys[3] = {"Source Endianness", "Source Type", "Source Bitwidth"};
  const char* vals[3] = {"little", "integer", v.bpv == 2 ? "16" : "8"};
  w.kv_keys = keys; w.kv_values = vals; w.kv_count = 3;
  char err[256] = {0};
  if (tvk_uvf_write(uvfFile.c_str(), &w, err, sizeof(err))) { m_err = err; return false; }
  return true;
}

void CudaGridLeaper::Resize(uint32_t w, uint32_t h) { m_w = w; m_h = h; ScheduleCompleteRedraw(); }

bool CudaGridLeaper::SetRendermode(int mode) {
  if (mode < RM_1DTRANS || mode > RM_ISOSURFACE) { m_err = "invalid render mode"; return false; }
  if (mode == RM_2DTRANS && !m_has2d) { m_err = "no 2D transfer function set"; return false; }
  // the device holds ONE table (texture unit 1 of the reference: 1D or 2D, GLGridLeaper.cpp:754-870): re-upload on a switch
  if (m_ctx && mode != m_mode) {
    if (mode == RM_2DTRANS) {
      const int rc = Upload2DTable();
      if (rc) return Fail("tvk_set_tf2d", rc);
    } else if (m_mode == RM_2DTRANS && m_tf1d.GetSize()) {
      std::vector<uint8_t> bytes;
      m_tf1d.GetByteArray(bytes);
      const int rc = tvk_set_tf1d(m_ctx, bytes.data(), (uint32_t)m_tf1d.GetSize());
      if (rc) return Fail("tvk_set_tf1d", rc);
    }
  }
  m_mode = mode;
  RecomputeBrickVisibility(false);
  ScheduleCompleteRedraw();
  return true;
}
void CudaGridLeaper::SetUseLighting(bool b) { m_lighting = b; ScheduleCompleteRedraw(); }
void CudaGridLeaper::SetSampleRateModifier(float f) { m_sampleRate = f; ScheduleCompleteRedraw(); }
void CudaGridLeaper::SetFoV(float f) { m_fov = f; ScheduleCompleteRedraw(); }
void CudaGridLeaper::SetViewPos(const Vec3& p) { m_at = m_at + (p - m_eye); m_eye = p; ScheduleCompleteRedraw(); }   // AbstrRenderer.cpp:1587-1591
void CudaGridLeaper::SetViewDir(const Vec3& d) { m_at = d + m_eye; ScheduleCompleteRedraw(); }                       // :1603-1606
void CudaGridLeaper::SetUpDir(const Vec3& u) { m_up = u; ScheduleCompleteRedraw(); }
void CudaGridLeaper::ResetViewPos() { m_at = Vec3(0, 0, 0); m_eye = Vec3(0, 0, 1.6f); ScheduleCompleteRedraw(); }
void CudaGridLeaper::ResetViewDir() { m_at = Vec3(0, 0, 0); ScheduleCompleteRedraw(); }
void CudaGridLeaper::ResetUpDir() { m_up = Vec3(0, 1, 0); ScheduleCompleteRedraw(); }
void CudaGridLeaper::SetViewParameters(float angle, float znear, float zfar, const Vec3& eye, const Vec3& ref, const Vec3& up) {
  m_fov = angle; m_znear = znear; m_zfar = zfar; m_eye = eye; m_at = ref; m_up = up;
  ScheduleCompleteRedraw();
}
void CudaGridLeaper::SetRotation4x4(const Matrix4& m) { m_rotation = m; ScheduleCompleteRedraw(); }
void CudaGridLeaper::SetTranslation4x4(const Matrix4& m) { m_translation = m; ScheduleCompleteRedraw(); }
void CudaGridLeaper::SetIsoValue(float v) { m_isoValue = v; RecomputeBrickVisibility(false); ScheduleCompleteRedraw(); }
void CudaGridLeaper::SetIsoValueRelative(float rel) {           // AbstrRenderer.cpp:395-410
  double lo = 0, hi = 0;
  if (m_vol) m_vol->Range(lo, hi);
  SetIsoValue(float(lo) + rel * float(hi - lo));
}
void CudaGridLeaper::SetIsosurfaceColor(const Vec3& c) { m_isoColor = c; ScheduleCompleteRedraw(); }
void CudaGridLeaper::SetLightColors(const float a[4], const float d[4], const float s[4], const Vec3& dir) {
  memcpy(m_ambient, a, 16); memcpy(m_diffuse, d, 16); memcpy(m_specular, s, 16); m_lightDir = dir;
  ScheduleCompleteRedraw();
}

bool CudaGridLeaper::Set1DTrans(const std::vector<uint8_t>& rgba) {          // GLRenderer::Set1DTrans
  if (rgba.size() % 4 || rgba.empty()) { m_err = "1D transfer function must be RGBA8"; return false; }
  std::vector<float> f(rgba.size());
  for (size_t i = 0; i < rgba.size(); ++i) f[i] = rgba[i] / 255.0f;
  m_tf1d.Set(f.data(), rgba.size() / 4);
  if (m_mode != RM_2DTRANS) {
    int rc = tvk_set_tf1d(m_ctx, rgba.data(), (uint32_t)(rgba.size() / 4));
    if (rc) return Fail("tvk_set_tf1d", rc);
  }
  RecomputeBrickVisibility(false);
  ScheduleCompleteRedraw();
  return true;
}

void CudaGridLeaper::Changed1DTrans() {                                      // GLGridLeaper.cpp:632-635
  if (!m_ctx) return;
  std::vector<uint8_t> bytes;
  m_tf1d.ComputeNonZeroLimits();
  m_tf1d.GetByteArray(bytes);
  if (m_mode != RM_2DTRANS || !m_has2d) tvk_set_tf1d(m_ctx, bytes.data(), (uint32_t)m_tf1d.GetSize());
  RecomputeBrickVisibility(false);
  ScheduleCompleteRedraw();
}

// GLRenderer::Render2DView (GLRenderer.cpp:1057-1176), slice and MIP-preview path: the whole target is ONE 2-D render
// region showing the finest LOD that is a single brick.  windowMode / sliceIndex / useMIP / flipView are the fields of
// RenderRegion2D (RenderRegion.h:53-140); the result is the frame (GetFrameFloat / GetFrameRGBA8 / CaptureSingleFrame).
bool CudaGridLeaper::ComputeSliceParams(int windowMode, uint64_t sliceIndex, bool useMIP, bool flipX, bool flipY, tvk_slice_params& p, uint64_t& toc) const {
  if (!m_vol || !m_w || !m_h || windowMode < TVK_WM_SAGITTAL || windowMode > TVK_WM_CORONAL) return false;
  const BrickedVolume& v = *m_vol;
  memset(&p, 0, sizeof(p));
  p.width = m_w; p.height = m_h; p.window_mode = (uint32_t)windowMode; p.use_mip = useMIP ? 1u : 0u;
  const uint32_t lod = v.LargestSingleBrickLOD();          // "the LOD level with just one brick" (:1095-1105)
  toc = v.TocIndex(0, 0, 0, lod);
  double aspect[3];
  for (int a = 0; a < 3; ++a) {
    p.voxels[a] = v.sizes[3 * toc + a];                    // vVoxelCount (= vRealVoxelCount: no power-of-two padding)
    // vMinCoords = (vVoxelCount - domain(lod) + 1) / (2 * vRealVoxelCount); vMaxCoords = vVoxelCount / vRealVoxelCount - vMinCoords (:1141-1142)
    p.min_coords[a] = (float(p.voxels[a]) - float(v.info.lod_size[lod][a]) + 1.0f) / (2.0f * float(p.voxels[a]));
    p.max_coords[a] = (float(p.voxels[a]) / float(p.voxels[a])) - p.min_coords[a];
    aspect[a] = double(v.scale[a]) * double(v.domain[a]);  // GetScale() * GetDomainSize() (:1144-1145)
  }
  p.bytes_per_voxel = v.bpv;
  // vWinAspectRatio = 1 / region size, normalised by its largest component (:1147-1149)
  double win[2] = {1.0 / double(m_w), 1.0 / double(m_h)};
  const double wm = std::max(win[0], win[1]);
  win[0] /= wm; win[1] /= wm;
  // RenderSlice (:922-1014): which volume axes the quad shows, and what flipView swaps
  const int ah = windowMode == TVK_WM_SAGITTAL ? 1 : 0, av = windowMode == TVK_WM_CORONAL ? 1 : 2;
  if (flipX) std::swap(p.min_coords[ah], p.max_coords[ah]);
  if (flipY) std::swap(p.min_coords[av], p.max_coords[av]);
  double q[2] = {aspect[ah] * win[0], aspect[av] * win[1]};
  const double qm = std::max(q[0], q[1]);
  p.quad[0] = float(q[0] / qm); p.quad[1] = float(q[1] / qm);
  p.slice = float(double(sliceIndex) / double(v.domain[windowMode]));        // sliceDir == windowMode (:1151, :1170-1171)
  double lo, hi;
  v.Range(lo, hi);
  const double maxValue = (lo > hi) ? double(m_tf1d.GetSize()) : hi;
  p.trans_scale = float(((1u << (8 * v.bpv)) - 1) / maxValue);
  p.gradient_scale = (v.max_grad_magnitude == 0) ? 1.0f : 1.0f / v.max_grad_magnitude;
  p.render_mode = (!useMIP && m_mode == RM_2DTRANS && m_has2d) ? TVK_RM_2DTRANS : TVK_RM_1DTRANS;   // `default:` = the 1-D slice shader (:1072-1079)
  return true;
}

bool CudaGridLeaper::Render2DView(int windowMode, uint64_t sliceIndex, bool useMIP, bool flipX, bool flipY) {
  if (!m_ctx || !m_vol) { m_err = "Render2DView() without dataset"; return false; }
  tvk_slice_params p;
  uint64_t toc = 0;
  if (!ComputeSliceParams(windowMode, sliceIndex, useMIP, flipX, flipY, p, toc)) { m_err = "bad 2-D view (window mode, or no render target)"; return false; }
  if (sliceIndex >= m_vol->domain[windowMode]) { m_err = "slice index beyond the domain"; return false; }
  const void* brick = m_vol->host_store ? (const void*)(m_vol->host_store + m_vol->offsets[toc])
                                        : (const void*)((const uint8_t*)m_vol->device_store + m_vol->offsets[toc]);
  if (m_vol->lz4_store) { m_err = "2-D views need an uncompressed brick store"; return false; }
  // the context holds ONE table: the slice shaders of the 1-D path (and Transfer-MIP) want the 1-D function
  const bool swapTF = p.render_mode == TVK_RM_1DTRANS && m_mode == RM_2DTRANS && m_has2d;
  std::vector<uint8_t> bytes;
  if (swapTF) { m_tf1d.GetByteArray(bytes); const int rc = tvk_set_tf1d(m_ctx, bytes.data(), (uint32_t)m_tf1d.GetSize()); if (rc) return Fail("tvk_set_tf1d", rc); }
  int rc = tvk_render_slice(m_ctx, &p, brick);
  if (swapTF) { const int r2 = Upload2DTable(); if (!rc && r2) rc = r2; }
  if (rc) return Fail("tvk_render_slice", rc);
  m_blank = true; m_converged = false;                     // the 3-D frame has to be redrawn from scratch afterwards
  return true;
}

// The full-resolution MIP of the capture mode (GLRenderer::Render2DView :1182-1250 with m_eRendererTarget == RT_CAPTURE):
// AbstrRenderer::PlanHQMIPFrame (:1214-1245) picks the LOD and lists EVERY brick of it (a 2-D region needs them all,
// RegionNeedsBrick :879-882), GLRenderer::RenderHQMIPPreLoop (:1256-1283) builds m_maMIPRotation, GLRaycaster's Pre/InLoop
// (GLRaycaster.cpp:481-529) set the model-view and fRayStepsize.  region.rotation of the 2-D region is the identity here.
bool CudaGridLeaper::PlanHQMIP(int windowMode, bool flipX, bool flipY, float angleDeg, bool ortho, bool mipLOD, tvk_mip_params& p,
                               std::vector<tvk_classic_brick>& bricks, uint32_t& lod) {
  if (!m_vol || !m_w || !m_h || windowMode < TVK_WM_SAGITTAL || windowMode > TVK_WM_CORONAL) { m_err = "bad MIP view (window mode, dataset or render target)"; return false; }
  const BrickedVolume& v = *m_vol;
  // PlanHQMIPFrame: halve until the smallest domain side is below the largest window side, then one level back
  uint64_t vc[3] = {v.domain[0], v.domain[1], v.domain[2]};
  uint64_t l = 0;
  if (mipLOD)
    while (std::min(vc[0], std::min(vc[1], vc[2])) >= std::max<uint64_t>(m_w, m_h)) { for (int a = 0; a < 3; ++a) vc[a] /= 2; ++l; }
  if (l > 0) l = std::min<uint64_t>(v.info.lod_count - 1, l - 1);
  lod = (uint32_t)l;
  const double PI = 3.141592653589793238462643383;
  Matrix4 rotDir, flipXm, flipYm;
  if (windowMode == TVK_WM_SAGITTAL) rotDir = Matrix4::RotationX(-PI / 2.0) * Matrix4::RotationY(-PI / 2.0);
  else if (windowMode == TVK_WM_AXIAL) rotDir = Matrix4::RotationX(-PI / 2.0);
  if (flipX) flipYm = Matrix4::Scaling(-1, 1, 1);
  if (flipY) flipXm = Matrix4::Scaling(1, -1, 1);
  const Matrix4 mipRot = rotDir * Matrix4() * flipXm * flipYm * Matrix4::RotationY(PI * double(angleDeg) / 180.0);
  ClassicPlanOverride ov;
  ov.allBricks = true;
  const float aspect = float(m_w) / float(m_h);
  if (ortho) {
    // GLRenderer.cpp:1183-1197
    double wa[2] = {1.0 / double(m_w), 1.0 / double(m_h)};
    const double wm = std::max(wa[0], wa[1]);
    wa[0] /= wm; wa[1] /= wm;
    const float r2 = (wa[0] < wa[1]) ? std::max(1.0f, 1.414213f * float(wa[0] / wa[1])) : 1.414213f;
    ov.proj = Matrix4::Ortho(-0.5f * r2 / float(wa[0]), +0.5f * r2 / float(wa[0]), -0.5f * r2 / float(wa[1]), +0.5f * r2 / float(wa[1]), -100.0f, 100.0f);
    ov.mv = mipRot;
  } else {
    ov.proj = Matrix4::Perspective(m_fov, aspect, m_znear, m_zfar);
    ov.mv = mipRot * Matrix4::LookAt(m_eye, m_at, m_up);
  }
  tvk_classic_params cp;
  uint32_t lodOut = 0;
  if (!PlanClassicFrame(cp, bricks, lodOut, (int)lod, nullptr, 0, &ov)) return false;
  memset(&p, 0, sizeof(p));
  p.width = m_w; p.height = m_h;
  memcpy(p.inv_model_view, cp.inv_model_view, 64);
  memcpy(p.inv_projection, cp.inv_projection, 64);
  p.bytes_per_voxel = v.bpv;
  p.trans_scale = cp.trans_scale;
  return true;
}

bool CudaGridLeaper::RenderHQMIP(int windowMode, bool flipX, bool flipY, float angleDeg, bool ortho, bool mipLOD) {
  if (!m_ctx || !m_vol) { m_err = "RenderHQMIP() without dataset"; return false; }
  const void* store = ClassicStore();
  if (!store) return false;
  tvk_mip_params p;
  std::vector<tvk_classic_brick> bricks;
  uint32_t lod = 0;
  if (!PlanHQMIP(windowMode, flipX, flipY, angleDeg, ortho, mipLOD, p, bricks, lod)) return false;
  const bool swapTF = m_mode == RM_2DTRANS && m_has2d;      // the context holds ONE table; the MIP wants the 1-D function
  std::vector<uint8_t> bytes;
  if (swapTF) { m_tf1d.GetByteArray(bytes); const int rc = tvk_set_tf1d(m_ctx, bytes.data(), (uint32_t)m_tf1d.GetSize()); if (rc) return Fail("tvk_set_tf1d", rc); }
  int rc = tvk_raycast_mip(m_ctx, &p, bricks.data(), (uint32_t)bricks.size(), store);
  if (swapTF) { const int r2 = Upload2DTable(); if (!rc && r2) rc = r2; }
  if (rc) return Fail("tvk_raycast_mip", rc);
  m_blank = true; m_converged = false;                     // the 3-D frame has to be redrawn from scratch afterwards
  return true;
}

bool CudaGridLeaper::Load2DTrans(const std::string& filename) {
  // AbstrRenderer::LoadFile sizes the 2-D function like the 1-D one: value bins x 256 gradient bins (AbstrRenderer.cpp:236-244)
  TransferFunction2D tf;
  if (!tf.Load(filename, std::max<size_t>(1, m_tf1d.GetSize()), 256)) { m_err = "could not load 2D transfer function " + filename; return false; }
  std::vector<uint8_t> table;
  double nz[4];
  tf.Render(table, nz);
  return Set2DTrans(table.data(), (uint32_t)tf.Width(), (uint32_t)tf.Height(), nz);
}

bool CudaGridLeaper::Set2DTrans(const uint8_t* rgba, uint32_t w, uint32_t h, double nz[4]) {
  // the swatch rasterisation of TransferFunction2D needs Qt (TransferFunction2D.cpp:294-361);
  // the renderer takes the rasterised RGBA8 table and its non-zero limits instead
  if (!rgba || !w || !h) { m_err = "bad 2D transfer function"; return false; }
  m_tf2dW = w; m_tf2dH = h;
  if (m_mode == RM_2DTRANS && m_ctx) {
    // in 2-D mode the table goes straight to the context, whose page-locked staging block then IS the host copy
    // (Upload2DTable); one 4 MB copy per edit instead of two
    int rc = tvk_set_tf2d(m_ctx, rgba, w, h);
    if (rc) return Fail("tvk_set_tf2d", rc);
    m_tf2dBytes.clear(); m_tf2dInCtx = true;
  } else {                                                 // uploaded when the mode is switched to 2D
    m_tf2dBytes.assign(rgba, rgba + (size_t)w * h * 4); m_tf2dInCtx = false;
  }
  m_has2d = true;
  for (int i = 0; i < 4; ++i) m_tf2dLimits[i] = nz[i];
  if (m_mode == RM_2DTRANS) RecomputeBrickVisibility(false);
  ScheduleCompleteRedraw();
  return true;
}

int CudaGridLeaper::Upload2DTable() {                                        // the 2-D table back onto the device (mode switch, slice view)
  if (m_tf2dInCtx) {
    const uint8_t* p = nullptr;
    uint32_t w = 0, h = 0;
    const int rc = tvk_tf2d_staging(m_ctx, &p, &w, &h);
    return rc ? rc : tvk_set_tf2d(m_ctx, p, w, h);
  }
  const int rc = tvk_set_tf2d(m_ctx, m_tf2dBytes.data(), m_tf2dW, m_tf2dH);
  if (!rc) { m_tf2dBytes.clear(); m_tf2dBytes.shrink_to_fit(); m_tf2dInCtx = true; }
  return rc;
}

bool CudaGridLeaper::RecomputeBrickVisibility(bool force) {                  // GLGridLeaper.cpp:647-688
  if (!m_pool || !m_vol) return false;
  double lo, hi;
  m_vol->Range(lo, hi);
  double maxValue = (lo > hi) ? double(m_tf1d.GetSize()) : hi;
  if (m_sl.on && m_sl.maxValue > 0) maxValue = m_sl.maxValue;          // sort-last: the table spans the global value range
  const double rescale = maxValue / double(m_tf1d.GetSize() - 1);
  tvk_visibility v{};
  v.render_mode = (uint32_t)m_mode;
  if (m_mode == RM_1DTRANS) { v.smin = double(m_tf1d.NonZeroMin()) * rescale; v.smax = double(m_tf1d.NonZeroMax()) * rescale; }
  else if (m_mode == RM_2DTRANS) {
    v.smin = m_tf2dLimits[0] * rescale; v.smax = m_tf2dLimits[1] * rescale; v.gmin = m_tf2dLimits[2]; v.gmax = m_tf2dLimits[3];
  } else v.smin = m_isoValue;
  // VisibilityState::NeedsUpdate (VisibilityState.h:15-50)
  if (!force && m_visValid && v.render_mode == m_lastVis.render_mode && v.smin == m_lastVis.smin && v.smax == m_lastVis.smax &&
      v.gmin == m_lastVis.gmin && v.gmax == m_lastVis.gmax)
    return true;
  const int rc = m_pool->RecomputeVisibility(v, nullptr);   // the counts are not used here: no read-back, no host wait
  if (rc) return Fail("RecomputeVisibility", rc);
  m_lastVis = v; m_visValid = true;
  return true;
}

bool CudaGridLeaper::RecalcVisibility() { const bool ok = RecomputeBrickVisibility(true); ScheduleCompleteRedraw(); return ok; }

void CudaGridLeaper::ClearWorkingSet() {                                     // GLVolumePool::PH_Reset (:261-284)
  if (!m_pool || !m_vol) return;
  const int rc = m_pool->Reset();
  if (rc) Fail("BrickPool::Reset", rc);
  m_visValid = false;
  RecomputeBrickVisibility(true);
  ScheduleCompleteRedraw();
}

bool CudaGridLeaper::ComputeRaycastParams(tvk_raycast_params& p, int eye) {
  if (!m_vol || !m_w || !m_h) { m_err = "no dataset / render target"; return false; }
  memset(&p, 0, sizeof(p));
  p.width = m_w; p.height = m_h;
  // GetVolumeAABB (AbstrRenderer.cpp:1102-1109)
  // (sort-last: the volume the camera sees is the GLOBAL one; this member holds a box of it)
  float dom[3], ext[3], mx = 0;
  for (int a = 0; a < 3; ++a) { dom[a] = m_sl.on ? float(m_sl.global[a]) : float(m_vol->domain[a]); ext[a] = dom[a] * m_vol->scale[a]; mx = std::max(mx, ext[a]); }
  for (int a = 0; a < 3; ++a) ext[a] /= mx;
  // GLRenderer::ComputeViewAndProjection (:888-920) and Paint (:627)
  const float aspect = float(m_w) / float(m_h);
  Matrix4 view = Matrix4::LookAt(m_eye, m_at, m_up);
  Matrix4 proj = Matrix4::Perspective(m_fov, aspect, m_znear, m_zfar);
  if (m_stereo) {
    Matrix4 views[2], projs[2];
    Matrix4::StereoLookAtAndProjection(m_eye, m_at, m_up, m_fov, aspect, m_znear, m_zfar, m_stereoFocalLength, m_stereoEyeDist, views, projs);
    view = views[eye ? 1 : 0]; proj = projs[eye ? 1 : 0];
  }
  const Matrix4 mv = m_rotation * m_translation * view;
  // ComputeEyeToModelMatrix (GLGridLeaper.cpp:560-573), centre = 0
  Matrix4 emm = mv.inverse() * Matrix4::Translation(0, 0, 0) * Matrix4::Scaling(1.0f / ext[0], 1.0f / ext[1], 1.0f / ext[2]) *
                Matrix4::Translation(0.5f, 0.5f, 0.5f);
  const Matrix4 emmGlobal = emm;                        // lighting lives in the global volume's model space
  if (m_sl.on) {
    // global normalised coordinates -> this box's: scale by global / local size, shift by the box's start
    const float ld[3] = {float(m_vol->domain[0]), float(m_vol->domain[1]), float(m_vol->domain[2])};
    emm = emm * Matrix4::Scaling(dom[0] / ld[0], dom[1] / ld[1], dom[2] / ld[2]) *
          Matrix4::Translation(-float(m_sl.start[0]) / ld[0], -float(m_sl.start[1]) / ld[1], -float(m_sl.start[2]) / ld[2]);
    for (int a = 0; a < 3; ++a) {
      p.model_pos_scale_m1[a] = float(double(ld[a]) / double(dom[a]) - 1.0);
      p.model_pos_offset[a] = float(double(m_sl.start[a]) / double(dom[a]));
    }
  }
  memcpy(p.eye_to_model, emm.m, 64);
  memcpy(p.model_to_eye, emm.inverse().m, 64);
  memcpy(p.model_view_it, mv.inverse().transposed().m, 64);
  memcpy(p.inv_projection, proj.inverse().m, 64);
  // CullingLOD::SetScreenParams (CullingLOD.cpp:65-66), screen space error 1
  p.lod_factor = 2.0f * (float)tan(m_fov * ((3.1416f / 180.0f) / 2.0f)) * 1.0f / float(m_h);
  float wse = 0;
  for (int a = 0; a < 3; ++a) wse = std::max(wse, ext[a] / dom[a]);   // GLVolumePool.cpp:801 (sort-last: of the global volume)
  p.level_zero_wse = wse;
  p.sample_rate_modifier = m_sampleRate;
  // CalculateScaling (GLRenderer.cpp:1813-1819) / MaxValue (AbstrRenderer.cpp:860-866)
  double lo, hi;
  m_vol->Range(lo, hi);
  double maxValue = (lo > hi) ? double(m_tf1d.GetSize()) : hi;
  if (m_sl.on && m_sl.maxValue > 0) maxValue = m_sl.maxValue;          // the transfer function spans the GLOBAL value range
  const uint32_t iMax = (1u << (8 * m_vol->bpv)) - 1;
  p.trans_scale = float(iMax / maxValue);
  p.gradient_scale = (m_vol->max_grad_magnitude == 0) ? 1.0f : 1.0f / m_vol->max_grad_magnitude;
  p.isoval = float(m_isoValue / double(1u << (8 * m_vol->bpv)));       // GetNormalizedIsovalue
  p.render_mode = (uint32_t)m_mode;
  p.lighting = (m_lighting && m_mode != RM_ISOSURFACE) ? 1u : 0u;
  float smin = std::min(m_vol->scale[0], std::min(m_vol->scale[1], m_vol->scale[2]));
  for (int a = 0; a < 3; ++a) {
    p.light_ambient[a] = m_ambient[a] * m_ambient[3];
    p.light_diffuse[a] = m_diffuse[a] * m_diffuse[3];
    p.light_specular[a] = m_specular[a] * m_specular[3];
    p.domain_scale[a] = 1.0f / (m_vol->scale[a] / smin);
  }
  const float l4[4] = {m_lightDir.x, m_lightDir.y, m_lightDir.z, 0.0f}, e4[4] = {m_eye.x, m_eye.y, m_eye.z, 1.0f};
  float o[4];
  emmGlobal.transform(l4, o);
  const Vec3 ml = Vec3(o[0], o[1], o[2]).normalized();
  p.model_light_dir[0] = ml.x; p.model_light_dir[1] = ml.y; p.model_light_dir[2] = ml.z;
  emmGlobal.transform(e4, o);                           // sic: the world-space eye through eye->model (:735)
  p.model_eye_pos[0] = o[0]; p.model_eye_pos[1] = o[1]; p.model_eye_pos[2] = o[2];
  p.iso_color[0] = m_isoColor.x; p.iso_color[1] = m_isoColor.y; p.iso_color[2] = m_isoColor.z;
  p.compose_light_dir[0] = m_lightDir.x; p.compose_light_dir[1] = m_lightDir.y; p.compose_light_dir[2] = m_lightDir.z;
  p.new_view = m_blank ? 1u : 0u;
  p.missing_strategy = m_strategy; p.rehash_count = m_rehash; p.hash_table_size = m_hashSize;
  p.clip_enabled = 0;
  if (m_clipEnabled) {
    // GLGridLeaper::FillBBoxVBO (:515-526): the world-space plane goes back to model space with
    // PLANE::operator*(inverse(rotation*translation)) = plane4 * (rotation*translation)^T, all four components
    // divided by |normal| (Vectors.h:1460-1477); Clipper::BoxPlane keeps n.p + d <= 0.  Model space is the centred
    // box of size `ext`; the kernel wants the plane over normalised [0,1]^3 coordinates.
    const Matrix4 rt = (m_rotation * m_translation).transposed();
    float q[4];
    rt.transform(m_clipPlane, q);
    const float len = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    if (len > 0.0f) {
      for (int a = 0; a < 4; ++a) q[a] /= len;
      p.clip_enabled = 1;
      for (int a = 0; a < 3; ++a) p.clip_plane[a] = q[a] * ext[a];
      p.clip_plane[3] = q[3] - 0.5f * (q[0] * ext[0] + q[1] * ext[1] + q[2] * ext[2]);
    }
  }
  for (int a = 0; a < 3; ++a) { p.box_min[a] = 0.0f; p.box_max[a] = 1.0f; }
  if (m_sl.on)
    for (int a = 0; a < 3; ++a) {
      p.box_min[a] = float((double(m_sl.lo[a]) - double(m_sl.start[a])) / double(m_vol->domain[a]));
      p.box_max[a] = float((double(m_sl.hi[a]) - double(m_sl.start[a])) / double(m_vol->domain[a]));
    }
  p.count_samples = 1;
  return true;
}

void CudaGridLeaper::SetClipPlane(const float plane[4]) { memcpy(m_clipPlane, plane, 16); if (m_clipEnabled) ScheduleCompleteRedraw(); }
void CudaGridLeaper::EnableClipPlane(bool on) { if (on != m_clipEnabled) { m_clipEnabled = on; ScheduleCompleteRedraw(); } }
// AbstrRenderer::SetTimestep (AbstrRenderer.cpp:1345-1361): time steps of a time-dependent UVF are separate octrees; the
// renderer swaps the brick store for the requested step's (read from the file again) and starts with an empty working
// set.  Transfer functions, iso value, mode and view are the renderer's and stay.
bool CudaGridLeaper::SetTimestep(size_t t) {
  if (t == m_timestep) return true;
  if (m_file.empty() || t >= m_timesteps) { m_err = "time step " + std::to_string(t) + " requested, the dataset has " + std::to_string(m_timesteps); return false; }
  const TransferFunction1D tf = m_tf1d;
  const float iso = m_isoValue;
  m_loadTimestep = t;
  m_rebrickSize = m_fileRebrick; m_rebrickMinMax = m_fileRebrickMinMax;
  const std::string file = m_file;
  const bool ok = LoadDataset(file, m_fileGpuMem, m_fileVerify, m_fileKeepCompressed);
  m_rebrickSize = 0;
  m_loadTimestep = 0;
  if (!ok) return false;
  m_tf1d = tf; m_isoValue = iso;
  Changed1DTrans();
  if (m_has2d && m_mode == RM_2DTRANS) { const int rc = Upload2DTable(); if (rc) return Fail("tvk_set_tf2d", rc); }
  m_visValid = false;
  ScheduleCompleteRedraw();
  return true;
}
void CudaGridLeaper::SetLODLimits(uint32_t lo, uint32_t hi) { m_lodLimits[0] = lo; m_lodLimits[1] = hi; }
void CudaGridLeaper::SetBGColors(const float top[3], const float bottom[3]) { memcpy(m_bg[0], top, 12); memcpy(m_bg[1], bottom, 12); }

// ---------------------------------------------------------------------------------------------------------
// Classic per-brick ray-caster (EVolumeRendererType OPENGL_RAYCASTER): AbstrRenderer::PlanFrame /
// ComputeMinLODForCurrentView / BuildSubFrameBrickList / RegionNeedsBrick / ContainsData / brick_distance
// (AbstrRenderer.cpp:1125-1212,789-803,999-1100,874-920,953-997,808-841), CullingLOD::Update / IsVisible /
// GetLODLevel (CullingLOD.cpp:89-160), the brick metadata of UVFDataset::ComputeMetadataTOC / GetTextCoords
// (uvfDataset.cpp:242-288,1747-1774) and the per-brick uniforms of GLRaycaster::SetBrickDepShaderVars /
// ComputeEyeToTextureMatrix (GLRaycaster.cpp:239-300,589-612).  The final sub-frame of a view (LOD =
// m_iMinLODForCurrentView, or 0 for RT_CAPTURE) is rendered in one launch of tvk_raycast_classic.
bool CudaGridLeaper::PlanClassicFrame(tvk_classic_params& p, std::vector<tvk_classic_brick>& bricks, uint32_t& lodOut, int forceLod,
                                      uint32_t* minLodOut, int eye, const ClassicPlanOverride* ov) {
  if (!m_vol || !m_w || !m_h) { m_err = "no dataset / render target"; return false; }
  const BrickedVolume& v = *m_vol;
  memset(&p, 0, sizeof(p));
  bricks.clear();
  const float aspect = float(m_w) / float(m_h);
  // Stereo (AbstrRenderer::PlanFrame :1196-1201, BuildLeftEyeSubFrameBrickList :843-858, GLRaycaster::Render3DInLoop :389):
  // LOD, frustum culling and the brick set come from eye 0's matrices; the other eye draws the SAME bricks in its own
  // front-to-back order with its own model-view and projection.
  Matrix4 view = Matrix4::LookAt(m_eye, m_at, m_up), proj = Matrix4::Perspective(m_fov, aspect, m_znear, m_zfar);
  Matrix4 viewE = view, projE = proj;
  if (m_stereo) {
    Matrix4 views[2], projs[2];
    Matrix4::StereoLookAtAndProjection(m_eye, m_at, m_up, m_fov, aspect, m_znear, m_zfar, m_stereoFocalLength, m_stereoEyeDist, views, projs);
    view = views[0]; proj = projs[0];
    viewE = views[eye ? 1 : 0]; projE = projs[eye ? 1 : 0];
  }
  if (ov) projE = proj = ov->proj;                          // (the full-resolution MIP brings its own matrices)
  const Matrix4 mv = ov ? ov->mv : m_rotation * m_translation * view;
  const Matrix4 mvE = ov ? ov->mv : m_rotation * m_translation * viewE;
  const Matrix4 invMv = mvE.inverse();                      // everything drawn uses the eye's own matrices
  float dom0[3], ext[3], mx = 0;
  for (int a = 0; a < 3; ++a) { dom0[a] = float(v.domain[a]); ext[a] = dom0[a] * v.scale[a]; mx = std::max(mx, ext[a]); }
  for (int a = 0; a < 3; ++a) ext[a] /= mx;
  // ComputeMinLODForCurrentView
  float wse = 0;
  for (int a = 0; a < 3; ++a) wse = std::max(wse, ext[a] / dom0[a]);
  const float c4[4] = {0, 0, 0, 1};
  float o4[4];
  mv.transform(c4, o4);
  const float fz = std::max(m_znear, -o4[2]);
  const float lodFactor = 2.0f * (float)tan(m_fov * ((3.1416f / 180.0f) / 2.0f)) * 1.0f / float(m_h);
  const int level = (int)floor(log(lodFactor * fz / wse) / log(2.0f));
  const int lodCount = (int)v.info.lod_count;
  const int minLod = std::min(std::max(level, (int)m_lodLimits[1]), lodCount - 1);
  if (minLodOut) *minLodOut = (uint32_t)minLod;
  // GLRaycaster::SetBrickDepShaderVars (:253-254): a "decrease sampling rate now" pass halves the rate (m_fSampleDecFactor = 2)
  const float effRate = ov ? m_sampleRate : m_sampleRate / (m_decRateNow ? 2.0f : 1.0f);
  const uint32_t lod = forceLod >= 0 ? (uint32_t)std::min(forceLod, lodCount - 1) : ((m_target == RT_CAPTURE) ? 0u : (uint32_t)minLod);
  lodOut = lod;
  // CullingLOD::Update: planes from modelView * projection
  const Matrix4 mvp = mv * proj;
  float planes[6][4];
  const int cols[6] = {0, 0, 1, 1, 2, 2};
  const float sgn[6] = {-1, 1, -1, 1, 1, -1};
  for (int k = 0; k < 6; ++k)
    for (int r = 0; r < 4; ++r) planes[k][r] = sgn[k] * mvp.m[4 * r + cols[k]] + mvp.m[4 * r + 3];
  // per-LOD aspect (ExtendedOctree::ComputeMetadata, ExtendedOctree.cpp:188-243)
  double asp[3] = {1, 1, 1};
  {
    uint64_t s[3] = {v.domain[0], v.domain[1], v.domain[2]};
    for (uint32_t l = 1; l <= lod; ++l) {
      for (int a = 0; a < 3; ++a)
        if (s[a] > 1) {
          const uint64_t n = (s[a] + 1) / 2;
          asp[a] *= (s[a] % 2) ? (double)((float)s[a] / (float)n) : 2.0;
          s[a] = n;
        }
      const double m = std::max(asp[0], std::max(asp[1], asp[2]));
      for (int a = 0; a < 3; ++a) asp[a] /= m;
    }
  }
  const uint32_t* nb = v.info.lod_bricks[lod];
  float dl[3], nds[3], maxval = 0;
  for (int a = 0; a < 3; ++a) { dl[a] = float(v.info.lod_size[lod][a]); nds[a] = dl[a] * float(asp[a]); maxval = std::max(maxval, nds[a]); }
  for (int a = 0; a < 3; ++a) nds[a] /= maxval;
  float vscale[3], corr = 0, dmax = std::max(dom0[0], std::max(dom0[1], dom0[2]));
  for (int a = 0; a < 3; ++a) corr = std::max(corr, v.scale[a] * dom0[a] / dmax);
  for (int a = 0; a < 3; ++a) vscale[a] = v.scale[a] / corr;
  const uint32_t core = v.brick_size - 2 * v.overlap;
  std::vector<float> extAxis[3], corner[3];
  for (int a = 0; a < 3; ++a) {
    float acc = 0;
    for (uint32_t c = 0; c < nb[a]; ++c) {
      const bool last = c >= nb[a] - 1;
      const uint32_t rem = v.info.lod_size[lod][a] % core;
      const uint32_t size = (last && rem) ? 2 * v.overlap + rem : v.brick_size;
      const float e = float(size - 2 * v.overlap) * float(asp[a]) / maxval;
      extAxis[a].push_back(e);
      corner[a].push_back(acc);
      acc += e;
    }
  }
  // ContainsData (1D): the TF's non-zero index range scaled to data values
  double lo, hi;
  v.Range(lo, hi);
  const double maxValue = (lo > hi) ? double(m_tf1d.GetSize()) : hi;
  const double resc = maxValue / double(m_tf1d.GetSize() - 1);
  // AbstrRenderer::ContainsData (:953-995): 1D / 2D scalar limits (a TOC file's gradient limits are -/+DBL_MAX, so
  // the 2D gradient test passes for every brick), isosurface: isovalue <= max
  const double smin = (m_mode == RM_2DTRANS) ? m_tf2dLimits[0] * resc : double(m_tf1d.NonZeroMin()) * resc;
  const double smax = (m_mode == RM_2DTRANS) ? m_tf2dLimits[1] * resc : double(m_tf1d.NonZeroMax()) * resc;
  const Matrix4 world = m_rotation * m_translation;
  struct Item { float dist; uint64_t idx; tvk_classic_brick b; };
  std::vector<Item> items;
  for (uint32_t x = 0; x < nb[0]; ++x)
    for (uint32_t y = 0; y < nb[1]; ++y)
      for (uint32_t z = 0; z < nb[2]; ++z) {
        const uint64_t idx = v.info.lod_offset[lod] + x + (uint64_t)y * nb[0] + (uint64_t)z * nb[0] * nb[1];
        const uint32_t bc[3] = {x, y, z};
        float be[3], bcn[3];
        for (int a = 0; a < 3; ++a) {
          const float e = extAxis[a][bc[a]];
          const float cen = (corner[a][bc[a]] + e / 2.0f) - nds[a] * 0.5f;
          be[a] = e * vscale[a]; bcn[a] = cen * vscale[a];
        }
        bool vis = true;
        for (int k = 0; k < 6 && vis && !(ov && ov->allBricks); ++k) {
          const float* pl = planes[k];
          if (pl[0] * bcn[0] + pl[1] * bcn[1] + pl[2] * bcn[2] + pl[3] <=
              -(0.5f * be[0] * fabsf(pl[0]) + 0.5f * be[1] * fabsf(pl[1]) + 0.5f * be[2] * fabsf(pl[2]))) vis = false;
        }
        if (!vis) continue;
        if (m_clipEnabled && !(ov && ov->allBricks)) {     // AbstrRenderer::Clipped (:923-948): all eight world-space corners beyond the plane
          bool all = true;
          for (int k = 0; k < 8 && all; ++k) {
            const float pt[4] = {bcn[0] + ((k & 4) ? 0.5f : -0.5f) * be[0], bcn[1] + ((k & 2) ? 0.5f : -0.5f) * be[1],
                                 bcn[2] + ((k & 1) ? 0.5f : -0.5f) * be[2], 1.0f};
            float w4[4];
            world.transform(pt, w4);
            const float wx = w4[0] / w4[3], wy = w4[1] / w4[3], wz = w4[2] / w4[3];
            if (!(m_clipPlane[0] * wx + m_clipPlane[1] * wy + m_clipPlane[2] * wz + m_clipPlane[3] >= 0)) all = false;
          }
          if (all) continue;
        }
        const bool contains = (m_mode == RM_ISOSURFACE) ? (m_isoValue <= v.minmax[2 * idx + 1])
                            : (m_mode == RM_2DTRANS)  ? (smax >= v.minmax[2 * idx] && smin <= v.minmax[2 * idx + 1] &&
                                                         m_tf2dLimits[3] >= -DBL_MAX && m_tf2dLimits[2] <= DBL_MAX)
                                                      : (smax >= v.minmax[2 * idx] && smin <= v.minmax[2 * idx + 1]);
        if (!contains && !(ov && ov->allBricks)) continue;                                // bIsEmpty: never drawn
        Item it;
        memset(&it.b, 0, sizeof(it.b));
        float tmin[3], tmax[3], rs = FLT_MAX;
        for (int a = 0; a < 3; ++a) {
          const float n = float(v.sizes[3 * idx + a]);
          tmin[a] = float(v.overlap) / n;
          tmax[a] = (1.0f - tmin[a]) * float(asp[a]);
          rs = std::min(rs, be[a] * (1.0f / n) * 0.5f * 1.0f / effRate);
          it.b.center[a] = bcn[a]; it.b.extension[a] = be[a]; it.b.voxels[a] = v.sizes[3 * idx + a];
        }
        it.b.ray_step = rs;
        it.b.store_offset = v.offsets[idx];
        float dist = FLT_MAX;
        for (int sx = -1; sx <= 1; sx += 2) for (int sy = -1; sy <= 1; sy += 2) for (int sz = -1; sz <= 1; sz += 2) {
          const float pt[4] = {bcn[0] + float(sx) * be[0] * 0.4999f, bcn[1] + float(sy) * be[1] * 0.4999f, bcn[2] + float(sz) * be[2] * 0.4999f, 1.0f};
          float e4[4];
          mvE.transform(pt, e4);
          dist = std::min(dist, sqrtf(e4[0] * e4[0] + e4[1] * e4[1] + e4[2] * e4[2]));
        }
        it.dist = dist; it.idx = idx;
        // ComputeEyeToTextureMatrix(p1 = max point, t1 = max coords, p2 = min point, t2 = min coords)
        float p1[3], p2[3];
        for (int a = 0; a < 3; ++a) { p1[a] = bcn[a] + be[a] / 2.0f; p2[a] = bcn[a] - be[a] / 2.0f; }
        const Matrix4 tm = invMv * Matrix4::Translation(-p1[0], -p1[1], -p1[2]) *
                           Matrix4::Scaling((tmin[0] - tmax[0]) / (p2[0] - p1[0]), (tmin[1] - tmax[1]) / (p2[1] - p1[1]), (tmin[2] - tmax[2]) / (p2[2] - p1[2])) *
                           Matrix4::Translation(tmax[0], tmax[1], tmax[2]);
        memcpy(it.b.eye_to_tex, tm.m, 64);
        items.push_back(it);
      }
  std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.dist < b.dist || (a.dist == b.dist && a.idx < b.idx); });
  for (const Item& it : items) bricks.push_back(it.b);
  p.width = m_w; p.height = m_h;
  memcpy(p.inv_model_view, invMv.m, 64);
  const Matrix4 nmT = invMv.transposed();
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) p.normal_matrix[3 * r + c] = nmT.m[4 * r + c];
  memcpy(p.inv_projection, projE.inverse().m, 64);
  float ratio = 0;
  for (int a = 0; a < 3; ++a) ratio = std::max(ratio, dom0[a] / dl[a]);
  p.step_scale = 1.0f / effRate * ratio;
  const uint32_t iMax = (1u << (8 * v.bpv)) - 1;
  p.trans_scale = float(iMax / maxValue);
  p.lighting = m_lighting ? 1u : 0u;
  for (int a = 0; a < 3; ++a) {
    p.light_ambient[a] = m_ambient[a] * m_ambient[3];
    p.light_diffuse[a] = m_diffuse[a] * m_diffuse[3];
    p.light_specular[a] = m_specular[a] * m_specular[3];
    p.domain_scale[a] = 1.0f / v.scale[a];
  }
  p.light_dir[0] = m_lightDir.x; p.light_dir[1] = m_lightDir.y; p.light_dir[2] = m_lightDir.z;
  p.bytes_per_voxel = v.bpv;
  p.render_mode = (uint32_t)m_mode;
  p.gradient_scale = (v.max_grad_magnitude == 0) ? 1.0f : 1.0f / v.max_grad_magnitude;   // GLRenderer.cpp:1840-1841
  p.isoval = float(m_isoValue / double(1u << (8 * v.bpv)));                                // GetNormalizedIsovalue
  p.proj_param[0] = m_zfar / (m_zfar - m_znear);                                           // CullingLOD::GetDepthScaleParams
  p.proj_param[1] = m_zfar * m_znear / (m_znear - m_zfar);
  p.iso_color[0] = m_isoColor.x; p.iso_color[1] = m_isoColor.y; p.iso_color[2] = m_isoColor.z;
  if (m_clipEnabled) {
    // RenderBox (GLRaycaster.cpp:327-337): world-space plane back to model space, PLANE::operator* divides all four
    // components by |normal| (Vectors.h:1460-1477), then the normal is normalised once more (a no-op)
    const Matrix4 rt = world.transposed();
    float q[4];
    rt.transform(m_clipPlane, q);
    const float len = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    if (len > 0.0f) {
      p.clip_enabled = 1;
      for (int a = 0; a < 4; ++a) p.clip_plane[a] = q[a] / len;
    }
  }
  return true;
}

// AbstrRenderer::ComputeMaxLODForCurrentView (AbstrRenderer.cpp:681-787): at which (coarse) LOD offset a new view starts,
// from how long the first two sub-frames of the previous view took.
void CudaGridLeaper::ComputeMaxLODForCurrentView(tvkr_lod_state& s) {
  if (s.target != RT_CAPTURE && s.msec_passed[0] >= 0.0f) {
    if (s.msec_passed[0] > s.max_ms_per_frame) {
      // too slow: after three such views, start one LOD coarser; at the coarsest, fall back on resolution / sampling rate
      if (s.lod_not_ok_counter < 3) {
        s.lod_not_ok_counter++;
      } else {
        s.lod_not_ok_counter = 0;
        const uint64_t skip = std::max<uint64_t>(1, s.perf_lod_skip) - 1;
        if (s.perf_lod_skip != skip) {
          s.msec_passed[0] = s.msec_passed[1];
          s.perf_lod_skip = skip;
        } else if (s.low_res_intermediate) {
          if (!(s.decrease_sampling_rate && s.decrease_screen_res)) {
            if (!s.decrease_screen_res) s.decrease_screen_res = 1;
            else s.decrease_sampling_rate = 1;
          }
        }
      }
    } else {
      // the last view finished its sub-frame, and its second sub-frame was fast enough: step the quality up
      if (s.brick_list_size == s.bricks_rendered && s.msec_passed[1] >= 0.0f && s.msec_passed[1] <= s.max_ms_per_frame) {
        s.lod_not_ok_counter = 0;
        if (s.decrease_sampling_rate || s.decrease_screen_res) {
          if (s.decrease_sampling_rate) s.decrease_sampling_rate = 0;
          else if (s.decrease_screen_res) s.decrease_screen_res = 0;
        } else {
          const uint64_t skip = std::min<uint64_t>(s.max_lod_index - s.min_lod_for_view, s.perf_lod_skip + 1);
          if (s.perf_lod_skip != skip) s.perf_lod_skip = skip;
        }
      }
    }
    s.start_lod_offset = std::max(s.min_lod_for_view, s.max_lod_index - s.perf_lod_skip);
  } else if (s.target != RT_INTERACTIVE) {
    s.start_lod_offset = s.min_lod_for_view;
  } else {
    s.start_lod_offset = s.max_lod_index;          // the very first frame: take it easy
  }
  s.start_lod_offset = std::min(s.start_lod_offset, static_cast<uint64_t>(s.max_lod_index - s.lod_limit_x));
  s.current_lod_offset = s.start_lod_offset;
  s.msec_passed[0] = s.msec_passed[1] = -1.0f;     // RestartTimers
}

bool CudaGridLeaper::PaintClassic() {
  if (!m_ctx || !m_vol) { m_err = "Paint() without dataset"; return false; }
  const void* store = ClassicStore();
  if (!store) return false;
  tvk_classic_params p;
  std::vector<tvk_classic_brick> bricks;
  uint32_t lod = 0, minLod = 0;
  return PaintClassicWith(store, p, bricks, lod, minLod);
}

const void* CudaGridLeaper::ClassicStore() {
  const void* store = m_vol->device_store;
  if (!store) {
    // datasets loaded from a UVF file (or out of core) keep their bricks in pinned host memory: the classic path reads
    // whole bricks from a device-resident store, so the store is copied up once when it fits
    if (!m_vol->classic_store) {
      size_t freeB = 0, totalB = 0;
      if (!m_vol->host_store || m_vol->lz4_store || cudaMemGetInfo(&freeB, &totalB) != cudaSuccess || m_vol->store_bytes + (1ull << 30) > freeB) {
        m_err = "the classic ray-caster needs the brick store in device memory and it does not fit (use the ray-guided renderer)";
        return nullptr;
      }
      if (cudaMalloc(&m_vol->classic_store, m_vol->store_bytes) != cudaSuccess ||
          cudaMemcpy(m_vol->classic_store, m_vol->host_store, m_vol->store_bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
        if (m_vol->classic_store) { cudaFree(m_vol->classic_store); m_vol->classic_store = nullptr; }
        (void)cudaGetLastError();
        m_err = "could not copy the brick store to the device for the classic ray-caster";
        return nullptr;
      }
    }
    store = m_vol->classic_store;
  }
  return store;
}

bool CudaGridLeaper::PaintClassicWith(const void* store, tvk_classic_params& p, std::vector<tvk_classic_brick>& bricks, uint32_t& lod, uint32_t& minLod) {
  // ---- AbstrRenderer::PlanFrame (:1125-1212).  One Paint() renders one whole sub-frame = the brick list of one LOD (the
  // reference's RT_HEADLESS time slice: GLRenderer::Render3DView, GLRenderer.cpp:2662-2675, stops early only for
  // interactive targets; a list that one kernel launch walks has nothing to slice).
  tvkr_lod_state& L = m_ladder;
  L.target = (int32_t)m_target;
  L.max_lod_index = m_vol->info.lod_count - 1;
  L.lod_limit_x = std::min<uint64_t>(m_lodLimits[0], L.max_lod_index);
  if (m_blank) {
    if (!PlanClassicFrame(p, bricks, lod, -1, &minLod)) return false;       // ComputeMinLODForCurrentView
    L.min_lod_for_view = minLod;
    ComputeMaxLODForCurrentView(L);
    m_decRateNow = L.decrease_sampling_rate != 0;
    m_decResNow = L.decrease_screen_res != 0;
    m_perf.frames += 1;
  } else if (m_decRateNow || m_decResNow) {
    m_decRateNow = m_decResNow = false;                                     // the low-quality pass is done: same LOD, full quality
  } else if (L.current_lod_offset > L.min_lod_for_view) {
    L.current_lod_offset--;                                                 // the sub-frame is complete: next finer LOD
  }
  m_currentLOD = (m_target == RT_CAPTURE) ? 0u : (uint32_t)std::min<uint64_t>(L.current_lod_offset, L.max_lod_index);
  if (!PlanClassicFrame(p, bricks, lod, (int)m_currentLOD)) return false;   // BuildSubFrameBrickList
  L.brick_list_size = bricks.size();
  L.bricks_rendered = 0;
  // ClearView: isosurface mode with the lens (GLRaycaster::Render3DInLoop's second pass + the ClearView composition)
  const bool clearView = m_cv && m_mode == RM_ISOSURFACE;
  tvk_clearview_params cv;
  if (clearView && !ComputeClearViewParams(cv, 0)) return false;
  int rc = clearView ? tvk_raycast_classic_cv(m_ctx, &p, &cv, bricks.data(), (uint32_t)bricks.size(), store)
                     : tvk_raycast_classic(m_ctx, &p, bricks.data(), (uint32_t)bricks.size(), store);
  if (rc) return Fail("tvk_raycast_classic", rc);
  m_cvFrame = m_mode == RM_ISOSURFACE;                   // the first surface's hit buffer is readable (Pick, ClearView focus)
  if (m_stereo) {
    // the other eye: the same bricks in its own order, into its own picture; then GLRenderer::EndFrame's composition
    tvk_classic_params pr;
    std::vector<tvk_classic_brick> br;
    uint32_t lodR = 0;
    if (!PlanClassicFrame(pr, br, lodR, (int)m_currentLOD, nullptr, 1)) return false;
    tvk_clearview_params cvr;
    if (clearView && !ComputeClearViewParams(cvr, 1)) return false;
    rc = tvk_select_eye(m_ctx, 1);
    if (!rc) rc = clearView ? tvk_raycast_classic_cv(m_ctx, &pr, &cvr, br.data(), (uint32_t)br.size(), store)
                            : tvk_raycast_classic(m_ctx, &pr, br.data(), (uint32_t)br.size(), store);
    const int rc2 = tvk_select_eye(m_ctx, 0);
    if (rc || rc2) return Fail("tvk_raycast_classic (right eye)", rc ? rc : rc2);
    if (!ComposeStereo()) return false;
  }
  L.bricks_rendered = bricks.size();
  m_blank = false;
  m_perf.subframes += 1;
  tvk_stats st;
  float ms = 0.0f;
  if (tvk_get_stats(m_ctx, &st, 0) == TVK_OK) { ms = st.last_raycast_ms; m_perf.raycast_ms += ms; m_perf.samples = (double)st.samples; }
  if (m_msOverride >= 0.0f) ms = m_msOverride;
  // AbstrRenderer::CompletedASubframe (:640-669): the first sub-frame's time tells whether the start LOD is affordable,
  // the second one's whether a finer start would be
  const bool first = (L.current_lod_offset == L.start_lod_offset) && (!L.decrease_screen_res || m_decResNow) &&
                     (!L.decrease_sampling_rate || m_decRateNow);
  const bool second = !first && (L.current_lod_offset == L.start_lod_offset ||
                                 (L.current_lod_offset + 1 == L.start_lod_offset && !(L.decrease_screen_res || L.decrease_sampling_rate)));
  if (first) L.msec_passed[0] = ms;
  else if (second) L.msec_passed[1] = ms;
  m_converged = (m_target == RT_CAPTURE) || (L.current_lod_offset <= L.min_lod_for_view && !m_decRateNow && !m_decResNow);
  return true;
}

double CudaGridLeaper::PerfQuery(int counter) {
  if (counter < 0 || counter >= PERF_END) return 0.0;
  const double v = m_perfTable[counter];
  m_perfTable[counter] = 0.0;
  return v;
}

bool CudaGridLeaper::Paint() {                                               // GLGridLeaper::Render3DRegion (:914-1154)
  if (m_classic) return PaintClassic();
  if (!m_ctx || !m_vol || !m_pool) { m_err = "Paint() without dataset"; return false; }
  if (!m_w || !m_h) { m_err = "Paint() on a zero-sized target"; return false; }
  if (m_stereo && m_sl.on) { m_err = "stereo is not available for a sort-last member"; return false; }
  typedef std::chrono::steady_clock clk;
  auto ms_since = [](clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); };
  const clk::time_point tPaint = clk::now();
  if (!m_visValid) RecomputeBrickVisibility(true);
  tvk_raycast_params p;
  if (!ComputeRaycastParams(p)) return false;
  if (m_blank) { m_perf.frames += 1; }
  if (m_sl.on && m_sl.exchange) {
    // every sub-frame of the displayed frame sends its pixels (tagged with the frame number) to the stripe owners; the
    // owners blend when the application says every member has converged (SortLastComposite)
    if (!m_sl.frame) { m_err = "sort-last: Paint() before SortLastBeginFrame()"; return false; }
    const int xr = tvk_exchange_begin_frame(m_ctx, m_sl.frame);
    if (xr) return Fail("tvk_exchange_begin_frame", xr);
  }
  // the march is queued behind whatever upload the previous sub-frame left in flight (tvk_pool_upload_fence)
  int rc = tvk_raycast_subframe(m_ctx, &p);
  if (rc) return Fail("tvk_raycast_subframe", rc);
  if (m_stereo) {
    // the second eye (GLGridLeaper.cpp:950-958): its own frame state, the same pool and request table
    tvk_raycast_params pr;
    if (!ComputeRaycastParams(pr, 1)) return false;
    rc = tvk_select_eye(m_ctx, 1);
    if (!rc) rc = tvk_raycast_subframe(m_ctx, &pr);
    const int rc2 = tvk_select_eye(m_ctx, 0);
    if (rc || rc2) return Fail("tvk_raycast_subframe (right eye)", rc ? rc : rc2);
  }
  m_blank = false;
  m_perf.subframes += 1;
  m_perfTable[PERF_SUBFRAMES] += 1;
  // ONE host synchronisation per sub-frame: request table and counters come back together
  clk::time_point t0 = clk::now();
  std::vector<uint32_t> req(4 * (size_t)m_hashSize);
  uint32_t n = 0;
  rc = tvk_read_requests(m_ctx, req.data(), m_hashSize, &n);
  if (rc) return Fail("tvk_read_requests", rc);
  req.resize(4 * (size_t)n);
  const double tRead = ms_since(t0);
  tvk_stats st;
  double marchMs = 0.0;
  if (tvk_get_stats(m_ctx, &st, 0) == TVK_OK) { marchMs = st.last_raycast_ms; m_perf.raycast_ms += marchMs; m_perf.samples = (double)st.samples; }
  m_perfTable[PERF_RAYCAST] += marchMs;
  m_perfTable[PERF_READ_HTABLE] += std::max(0.0, tRead - marchMs);          // the wait covers the march itself
  if (n) {
    uint32_t paged = 0;
    t0 = clk::now();
    const uint64_t before = m_pool->UploadedBytes();
    // the upload is issued and ordered before the next sub-frame; the host goes on (Paint() returns while PCIe works)
    rc = m_pool->UploadBricks(req, paged, false);
    if (rc) return Fail("UploadBricks", rc);
    m_perf.paged_bricks += paged;
    m_perf.uploaded_bytes = (double)m_pool->UploadedBytes();
    m_perfTable[PERF_UPLOAD_BRICKS] += ms_since(t0);
    m_perfTable[PERF_POOL_SORT] += m_pool->LastSortMs();
    m_perfTable[PERF_POOL_GET_BRICK] += m_pool->LastGetBrickMs();
    m_perfTable[PERF_POOL_UPLOAD_BRICK] += m_pool->LastUploadMs();
    m_perfTable[PERF_POOL_UPLOAD_METADATA] += m_pool->LastMetadataMs();
    m_perfTable[PERF_POOL_UPLOADED_MEM] += (double)(m_pool->UploadedBytes() - before);
    m_perfTable[PERF_EO_BRICKS] += paged;
  }
  m_converged = (n == 0);
  if (m_converged && m_stereo && !ComposeStereo()) return false;        // GLRenderer::EndFrame (:757-812)
  m_perfTable[PERF_RENDER] += ms_since(tPaint);
  return true;
}

// the model-view matrix of an eye (rotation * translation * view, GLRenderer.cpp:627-629)
Matrix4 CudaGridLeaper::ModelView(int eye) const {
  Matrix4 view = Matrix4::LookAt(m_eye, m_at, m_up);
  if (m_stereo) {
    Matrix4 views[2], projs[2];
    Matrix4::StereoLookAtAndProjection(m_eye, m_at, m_up, m_fov, float(m_w) / float(m_h), m_znear, m_zfar, m_stereoFocalLength, m_stereoEyeDist, views, projs);
    view = views[eye ? 1 : 0];
  }
  return m_rotation * m_translation * view;
}

bool CudaGridLeaper::ComputeClearViewParams(tvk_clearview_params& cv, int eye) {
  if (!m_vol || !m_w || !m_h) { m_err = "no dataset / render target"; return false; }
  memset(&cv, 0, sizeof(cv));
  cv.cv_isoval = float(double(m_cvIsoValue) / double(1u << (8 * m_vol->bpv)));          // GetNormalizedCVIsovalue
  cv.cv_color[0] = m_cvColor.x; cv.cv_color[1] = m_cvColor.y; cv.cv_color[2] = m_cvColor.z;
  cv.size = m_cvSize; cv.context_scale = m_cvContextScale; cv.border_scale = m_cvBorderScale;
  float t[4];
  ModelView(eye).transform(m_cvPos, t);                    // transPos = m_vCVPos * modelView (GLRenderer.cpp:2789)
  cv.pick_pos[0] = t[0]; cv.pick_pos[1] = t[1]; cv.pick_pos[2] = t[2];
  return true;
}

// AbstrRenderer::SetCVFocusPos (:1354-1361) + GLRenderer::CVFocusHasChanged (GLRenderer.cpp:2833-2851): the first surface's hit
// under the mouse goes back to world space through the inverse model-view; no hit there sends the lens to "nirvana"
bool CudaGridLeaper::SetCVFocusPos(int x, int y) {
  if (!m_ctx || !m_w || !m_h) { m_err = "no render target"; return false; }
  if (x < 0 || y < 0 || x >= (int)m_w || y >= (int)m_h) { m_err = "focus position outside the window"; return false; }
  float hit[4] = {0, 0, 0, 0};
  // (the reference reads row m_vWinSize.y - mouse.y; the last row is clamped here)
  const uint32_t row = (uint32_t)std::min<int>((int)m_h - y, (int)m_h - 1);
  if (!m_cvFrame || tvk_read_iso_hit(m_ctx, (uint32_t)x, row, hit) != TVK_OK) { m_err = "no isosurface frame of the classic renderer to pick from"; return false; }
  if (hit[3] != 0.0f) {
    const float h4[4] = {hit[0], hit[1], hit[2], 1.0f};
    ModelView(0).inverse().transform(h4, m_cvPos);
  } else {
    m_cvPos[0] = m_cvPos[1] = m_cvPos[2] = 10000000.0f; m_cvPos[3] = 0.0f;
  }
  CVChanged();
  return true;
}

// GLRenderer::Pick (GLRenderer.cpp:2856-2872): the eye-space position of the isosurface under a window position (y from the top)
bool CudaGridLeaper::Pick(int x, int y, float out[3]) {
  if (m_mode != RM_ISOSURFACE) { m_err = "Can only determine pick locations in isosurface rendering mode."; return false; }
  if (!m_ctx || x < 0 || y < 0 || x >= (int)m_w || y >= (int)m_h) { m_err = "pick position outside the window"; return false; }
  float hit[4] = {0, 0, 0, 0};
  const uint32_t row = (uint32_t)std::min<int>((int)m_h - y, (int)m_h - 1);
  if (tvk_read_iso_hit(m_ctx, (uint32_t)x, row, hit) != TVK_OK) { m_err = "picking reads the hit buffer of the last isosurface frame: paint first"; return false; }
  if (hit[3] == 0.0f) { m_err = "No intersection."; return false; }
  out[0] = hit[0]; out[1] = hit[1]; out[2] = hit[2];
  return true;
}

bool CudaGridLeaper::ComposeStereo() {
  if (!m_ctx) return false;
  const int rc = tvk_compose_stereo(m_ctx, (uint32_t)m_stereoMode, m_stereoEyeSwap ? 1u : 0u, 0.5f, m_afId);
  return rc ? Fail("tvk_compose_stereo", rc) : true;
}
bool CudaGridLeaper::SetStereoMode(int mode) {
  if (mode < TVK_SM_RB || mode > TVK_SM_AF) { m_err = "bad stereo mode"; return false; }
  m_stereoMode = mode;
  if (m_stereo) ScheduleCompleteRedraw();
  return true;
}
void CudaGridLeaper::SetStereoEyeSwap(bool swap) { m_stereoEyeSwap = swap; if (m_stereo) ScheduleCompleteRedraw(); }
bool CudaGridLeaper::ToggleStereoFrame() {                            // AbstrRenderer.cpp:1582-1585: recompose only
  m_afId = 1 - m_afId;
  if (m_stereo && !m_blank && (m_converged || m_classic)) return ComposeStereo();
  return true;
}

bool CudaGridLeaper::CheckForRedraw() const { return m_w && m_h && (!m_converged || m_blank); }

bool CudaGridLeaper::PaintUntilConverged(uint32_t maxSub, uint32_t& subframes) {
  subframes = 0;
  do {
    if (!Paint()) return false;
    ++subframes;
  } while (!m_converged && subframes < maxSub);
  return m_converged;
}

bool CudaGridLeaper::GetFrameFloat(std::vector<float>& out) {
  if (!m_ctx || !m_w || !m_h) return false;
  out.resize((size_t)m_w * m_h * 4);
  const int rc = tvk_read_image(m_ctx, out.data());
  return rc ? Fail("tvk_read_image", rc) : true;
}

bool CudaGridLeaper::GetFrameRGBA8(std::vector<uint8_t>& out) {
  if (!m_ctx || !m_w || !m_h) return false;
  out.resize((size_t)m_w * m_h * 4);
  return GetFrameRGBA8(out.data());
}

bool CudaGridLeaper::GetFrameRGBA8(uint8_t* out) {                           // width*height*4 bytes, caller-owned
  if (!m_ctx || !m_w || !m_h || !out) return false;
  // glReadPixels(GL_RGBA, GL_UNSIGNED_BYTE): clamp, scale, round to nearest; then the vertical flip of
  // FrameCapture::SaveImage (FrameCapture.h:90-98) -- both done on the device, a quarter of the bytes come back
  const int rc = tvk_read_image_rgba8(m_ctx, out, 1);
  return rc ? Fail("tvk_read_image_rgba8", rc) : true;
}

// minimal, dependency-free PNG writer (stored deflate blocks)
static uint32_t crc32_buf(uint32_t crc, const uint8_t* d, size_t n) {
  static uint32_t table[256];
  static bool init = false;
  if (!init) {
    for (uint32_t i = 0; i < 256; ++i) { uint32_t c = i; for (int k = 0; k < 8; ++k) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1; table[i] = c; }
    init = true;
  }
  crc = ~crc;
  for (size_t i = 0; i < n; ++i) crc = table[(crc ^ d[i]) & 0xFF] ^ (crc >> 8);
  return ~crc;
}
static void put_be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
static void png_chunk(std::vector<uint8_t>& out, const char* type, const std::vector<uint8_t>& data) {
  put_be32(out, (uint32_t)data.size());
  std::vector<uint8_t> td(type, type + 4);
  td.insert(td.end(), data.begin(), data.end());
  out.insert(out.end(), td.begin(), td.end());
  put_be32(out, crc32_buf(0, td.data(), td.size()));
}
static bool write_png(const std::string& fn, uint32_t w, uint32_t h, const std::vector<uint8_t>& rgba) {
  std::vector<uint8_t> raw;
  raw.reserve((size_t)h * (w * 4 + 1));
  for (uint32_t y = 0; y < h; ++y) { raw.push_back(0); raw.insert(raw.end(), rgba.begin() + (size_t)y * w * 4, rgba.begin() + (size_t)(y + 1) * w * 4); }
  std::vector<uint8_t> z;
  z.push_back(0x78); z.push_back(0x01);
  uint32_t a = 1, b = 0;
  for (size_t off = 0; off < raw.size();) {
    const size_t n = std::min<size_t>(65535, raw.size() - off);
    z.push_back(off + n == raw.size() ? 1 : 0);
    z.push_back(n & 0xFF); z.push_back(n >> 8); z.push_back(~n & 0xFF); z.push_back((~n >> 8) & 0xFF);
    z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
    for (size_t i = 0; i < n; ++i) { a = (a + raw[off + i]) % 65521; b = (b + a) % 65521; }
    off += n;
  }
  put_be32(z, (b << 16) | a);
  std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', '\r', '\n', 0x1a, '\n'};
  std::vector<uint8_t> ihdr;
  put_be32(ihdr, w); put_be32(ihdr, h);
  ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
  png_chunk(out, "IHDR", ihdr);
  png_chunk(out, "IDAT", z);
  png_chunk(out, "IEND", std::vector<uint8_t>());
  FILE* f = fopen(fn.c_str(), "wb");
  if (!f) return false;
  const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
  fclose(f);
  return ok;
}

bool CudaGridLeaper::CaptureSingleFrame(const std::string& filename, bool preserveTransparency) {
  std::vector<uint8_t> img;
  if (!GetFrameRGBA8(img)) return false;
  // FrameCapture.h:129-156: with transparency the colour is de-multiplied, without it alpha = 255
  for (size_t i = 0; i < img.size(); i += 4) {
    if (preserveTransparency) {
      const uint8_t a = img[i + 3];
      for (int k = 0; k < 3; ++k) img[i + k] = a ? (uint8_t)std::min(255, int(0.5f + img[i + k] / (a / 255.0f))) : 0;
    } else img[i + 3] = 255;
  }
  if (!write_png(filename, m_w, m_h, img)) { m_err = "cannot write " + filename; return false; }
  return true;
}

}  // namespace tvk

// ============================================================================ C wrapper (include/tvk_renderer.h)
#include "tvk_renderer.h"
using tvk::CudaGridLeaper;
using tvk::Matrix4;
using tvk::Vec3;

#define R(h) (reinterpret_cast<CudaGridLeaper*>(h))

extern "C" {
tvkr_renderer* tvkr_create(int device) {
  CudaGridLeaper* r = new CudaGridLeaper();
  if (!r->Initialize(device)) { delete r; return nullptr; }
  return reinterpret_cast<tvkr_renderer*>(r);
}
void tvkr_destroy(tvkr_renderer* h) { delete R(h); }
const char* tvkr_last_error(tvkr_renderer* h) { return h ? R(h)->LastError() : "null renderer"; }
int tvkr_load_raw_volume(tvkr_renderer* h, const void* raw, const uint32_t size[3], uint32_t bpv, const float scale[3],
                         uint32_t brick, uint32_t overlap, int median, int out_of_core, uint64_t gpu_mem_bytes, int hist2d) {
  return h && R(h)->LoadRawVolume(raw, size, bpv, scale, brick, overlap, median != 0, out_of_core != 0, gpu_mem_bytes, hist2d != 0) ? 0 : 1;
}
int tvkr_load_dataset(tvkr_renderer* h, const char* uvf_file, uint64_t gpu_mem_bytes, int verify_checksum) {
  return h && uvf_file && R(h)->LoadDataset(uvf_file, gpu_mem_bytes, verify_checksum != 0) ? 0 : 1;
}
int tvkr_sortlast_configure(tvkr_renderer* h, uint32_t members, uint32_t member, const uint32_t global_size[3], const uint32_t lo[3],
                            const uint32_t hi[3], const uint32_t start[3], double global_max_value) {
  return h && global_size && lo && hi && start && R(h)->SortLastConfigure(members, member, global_size, lo, hi, start, global_max_value) ? 0 : 1;
}
int tvkr_sortlast_exchange_create(tvkr_renderer* h, void* handle64) { return h && R(h)->SortLastExchangeCreate(handle64) ? 0 : 1; }
int tvkr_sortlast_exchange_connect(tvkr_renderer* h, const void* handles, void* const* bases) { return h && R(h)->SortLastExchangeConnect(handles, bases) ? 0 : 1; }
int tvkr_sortlast_exchange_base(tvkr_renderer* h, uint32_t member, void** base) {
  if (!h || !base) return 1;
  return tvk_exchange_peer_base(R(h)->Context(), member, base) == TVK_OK ? 0 : 1;
}
int tvkr_sortlast_begin_frame(tvkr_renderer* h) { return h && R(h)->SortLastBeginFrame() ? 0 : 1; }
int tvkr_sortlast_composite(tvkr_renderer* h, const uint32_t* order) { return h && R(h)->SortLastComposite(order) ? 0 : 1; }
int tvkr_sortlast_depth_order(tvkr_renderer* h, const uint32_t* boxes_lo_hi, uint32_t* order) {
  if (!h || !boxes_lo_hi || !order || !R(h)->SortLastActive()) return 1;
  R(h)->SortLastDepthOrder(boxes_lo_hi, order);
  return 0;
}
int tvkr_sortlast_read_frame_rgba8(tvkr_renderer* h, uint8_t* out, int flip_rows) { return h && out && R(h)->SortLastReadFrameRGBA8(out, flip_rows != 0) ? 0 : 1; }
int tvkr_load_streamed_begin(tvkr_renderer* h, const uint32_t size[3], uint32_t bpv, const float scale[3], uint32_t brick, uint32_t overlap, int median) {
  return h && size && R(h)->LoadStreamedBegin(size, bpv, scale, brick, overlap, median != 0) ? 0 : 1;
}
int tvkr_load_streamed_rows(tvkr_renderer* h, uint32_t bz0, uint32_t bz1, uint32_t* z0, uint32_t* rows) {
  return h && z0 && rows && R(h)->LoadStreamedRows(bz0, bz1, *z0, *rows) ? 0 : 1;
}
int tvkr_load_streamed_slab(tvkr_renderer* h, const void* device_slab, uint32_t z0, uint32_t rows, uint32_t bz0, uint32_t bz1) {
  return h && device_slab && R(h)->LoadStreamedSlab(device_slab, z0, rows, bz0, bz1) ? 0 : 1;
}
int tvkr_load_streamed_end(tvkr_renderer* h, uint64_t gpu_mem_bytes, float* builder_kernel_ms) {
  if (!h || !R(h)->LoadStreamedEnd(gpu_mem_bytes)) return 1;
  if (builder_kernel_ms) *builder_kernel_ms = R(h)->LastBuildMs();
  return 0;
}
int tvkr_load_volume_file(tvkr_renderer* h, const char* path, uint32_t brick, uint32_t overlap, int median, int out_of_core,
                          uint64_t gpu_mem_bytes, int hist2d) {
  return h && path && R(h)->LoadVolumeFile(path, brick, overlap, median != 0, out_of_core != 0, gpu_mem_bytes, hist2d != 0) ? 0 : 1;
}
int tvkr_load_rebricked(tvkr_renderer* h, const char* uvf_file, uint32_t brick_size, int minmax_mode, uint64_t gpu_mem_bytes, int verify_checksum) {
  return h && uvf_file && R(h)->LoadRebricked(uvf_file, brick_size, minmax_mode, gpu_mem_bytes, verify_checksum != 0) ? 0 : 1;
}
int tvkr_load_dataset_compressed(tvkr_renderer* h, const char* uvf_file, uint64_t gpu_mem_bytes, int verify_checksum) {
  return h && uvf_file && R(h)->LoadDataset(uvf_file, gpu_mem_bytes, verify_checksum != 0, true) ? 0 : 1;
}
int tvkr_store_is_compressed(tvkr_renderer* h) { return h && R(h)->Volume() && R(h)->Volume()->lz4_store ? 1 : 0; }
int tvkr_save_uvf(tvkr_renderer* h, const char* uvf_file, uint32_t compression, uint32_t level) {
  return h && uvf_file && R(h)->SaveUVF(uvf_file, compression, level) ? 0 : 1;
}
int tvkr_set_renderer_type(tvkr_renderer* h, int type) { if (!h || type < 0 || type > 1) return 1; R(h)->SetClassic(type == 1); return 0; }
int tvkr_plan_classic_frame(tvkr_renderer* h, tvk_classic_params* p, tvk_classic_brick* bricks, uint32_t cap, uint32_t* n, uint32_t* lod) {
  if (!h || !p || !n || !lod) return 1;
  std::vector<tvk_classic_brick> b;
  if (!R(h)->PlanClassicFrame(*p, b, *lod)) return 1;
  *n = (uint32_t)b.size();
  if (bricks) { if (b.size() > cap) return 1; memcpy(bricks, b.data(), sizeof(tvk_classic_brick) * b.size()); }
  return 0;
}
int tvkr_set_clip_plane(tvkr_renderer* h, const float plane[4]) { if (!h || !plane) return 1; R(h)->SetClipPlane(plane); return 0; }
int tvkr_enable_clip_plane(tvkr_renderer* h, int on) { if (!h) return 1; R(h)->EnableClipPlane(on != 0); return 0; }
int tvkr_pick(tvkr_renderer* h, int x, int y, float out[3]) { return h && out && R(h)->Pick(x, y, out) ? 0 : 1; }
int tvkr_set_clearview(tvkr_renderer* h, int on) { if (!h) return 1; R(h)->SetCV(on != 0); return 0; }
int tvkr_get_clearview(tvkr_renderer* h) { return h && R(h)->GetCV() ? 1 : 0; }
int tvkr_set_cv_iso_value(tvkr_renderer* h, float v) { if (!h) return 1; R(h)->SetCVIsoValue(v); return 0; }
int tvkr_get_cv_iso_value(tvkr_renderer* h, float* v) { if (!h || !v) return 1; *v = R(h)->GetCVIsoValue(); return 0; }
int tvkr_set_cv_color(tvkr_renderer* h, const float c[3]) { if (!h || !c) return 1; R(h)->SetCVColor(Vec3(c[0], c[1], c[2])); return 0; }
int tvkr_set_cv_size(tvkr_renderer* h, float f) { if (!h) return 1; R(h)->SetCVSize(f); return 0; }
int tvkr_set_cv_context_scale(tvkr_renderer* h, float f) { if (!h) return 1; R(h)->SetCVContextScale(f); return 0; }
int tvkr_set_cv_border_scale(tvkr_renderer* h, float f) { if (!h) return 1; R(h)->SetCVBorderScale(f); return 0; }
int tvkr_set_cv_focus_pos(tvkr_renderer* h, int x, int y) { return h && R(h)->SetCVFocusPos(x, y) ? 0 : 1; }
int tvkr_set_cv_focus_pos_fvec(tvkr_renderer* h, const float p[4]) { if (!h || !p) return 1; R(h)->SetCVFocusPosFVec(p); return 0; }
int tvkr_get_clearview_params(tvkr_renderer* h, int eye, tvk_clearview_params* out) {
  return h && out && (eye == 0 || eye == 1) && R(h)->ComputeClearViewParams(*out, eye) ? 0 : 1;
}
int tvkr_set_stereo(tvkr_renderer* h, int on) { if (!h) return 1; R(h)->SetStereo(on != 0); return 0; }
int tvkr_get_stereo(tvkr_renderer* h) { return h && R(h)->GetStereo() ? 1 : 0; }
int tvkr_set_stereo_eye_dist(tvkr_renderer* h, float d) { if (!h) return 1; R(h)->SetStereoEyeDist(d); return 0; }
int tvkr_set_stereo_focal_length(tvkr_renderer* h, float f) { if (!h || !(f > 0.0f)) return 1; R(h)->SetStereoFocalLength(f); return 0; }
int tvkr_set_stereo_mode(tvkr_renderer* h, int mode) { return h && R(h)->SetStereoMode(mode) ? 0 : 1; }
int tvkr_set_stereo_eye_swap(tvkr_renderer* h, int swap) { if (!h) return 1; R(h)->SetStereoEyeSwap(swap != 0); return 0; }
int tvkr_init_stereo_frame(tvkr_renderer* h) { if (!h) return 1; R(h)->InitStereoFrame(); return 0; }
int tvkr_toggle_stereo_frame(tvkr_renderer* h) { return h && R(h)->ToggleStereoFrame() ? 0 : 1; }
int tvkr_get_raycast_params_eye(tvkr_renderer* h, int eye, tvk_raycast_params* p) { return h && p && (eye == 0 || eye == 1) && R(h)->ComputeRaycastParams(*p, eye) ? 0 : 1; }
int tvkr_plan_classic_frame_eye(tvkr_renderer* h, int eye, tvk_classic_params* p, tvk_classic_brick* bricks, uint32_t cap, uint32_t* n, uint32_t* lod) {
  if (!h || !p || !n || (eye != 0 && eye != 1)) return 1;
  std::vector<tvk_classic_brick> b;
  uint32_t l = 0;
  if (!R(h)->PlanClassicFrame(*p, b, l, -1, nullptr, eye)) return 1;
  *n = (uint32_t)b.size();
  if (lod) *lod = l;
  if (bricks) { if (cap < b.size()) return 1; memcpy(bricks, b.data(), b.size() * sizeof(tvk_classic_brick)); }
  return 0;
}
int tvkr_get_eye_frame_float(tvkr_renderer* h, int eye, float* out) {
  return h && out && (eye == 0 || eye == 1) && R(h)->Context() && tvk_read_eye_image(R(h)->Context(), (uint32_t)eye, out) == TVK_OK ? 0 : 1;
}
int tvkr_stereo_matrices(const float eye[3], const float at[3], const float up[3], float fovy, float aspect, float znear, float zfar,
                         float focal_length, float eye_dist, float out[64]) {
  if (!eye || !at || !up || !out) return 1;
  Matrix4 v[2], p[2];
  Matrix4::StereoLookAtAndProjection(Vec3(eye[0], eye[1], eye[2]), Vec3(at[0], at[1], at[2]), Vec3(up[0], up[1], up[2]), fovy, aspect, znear, zfar,
                                     focal_length, eye_dist, v, p);
  memcpy(out, v[0].m, 64); memcpy(out + 16, v[1].m, 64); memcpy(out + 32, p[0].m, 64); memcpy(out + 48, p[1].m, 64);
  return 0;
}
int tvkr_set_light_colors(tvkr_renderer* h, const float ambient[4], const float diffuse[4], const float specular[4], const float dir[3]) {
  if (!h || !ambient || !diffuse || !specular || !dir) return 1;
  R(h)->SetLightColors(ambient, diffuse, specular, Vec3(dir[0], dir[1], dir[2]));
  return 0;
}
int tvkr_get_view(tvkr_renderer* h, float pos[3], float dir[3], float up[3]) {
  if (!h) return 1;
  const Vec3 e = R(h)->GetViewPos(), d = R(h)->GetViewDir(), u = R(h)->GetUpDir();
  if (pos) { pos[0] = e.x; pos[1] = e.y; pos[2] = e.z; }
  if (dir) { dir[0] = d.x; dir[1] = d.y; dir[2] = d.z; }
  if (up) { up[0] = u.x; up[1] = u.y; up[2] = u.z; }
  return 0;
}
int tvkr_reset_view(tvkr_renderer* h) { if (!h) return 1; R(h)->ResetViewPos(); R(h)->ResetViewDir(); R(h)->ResetUpDir(); return 0; }
int tvkr_set_timestep(tvkr_renderer* h, uint64_t t) { return h && R(h)->SetTimestep((size_t)t) ? 0 : 1; }
int tvkr_get_timesteps(tvkr_renderer* h, uint64_t* count, uint64_t* current) {
  if (!h) return 1;
  if (count) *count = R(h)->Timesteps();
  if (current) *current = R(h)->Timestep();
  return 0;
}
int tvkr_set_lod_limits(tvkr_renderer* h, uint32_t lo, uint32_t hi) { if (!h) return 1; R(h)->SetLODLimits(lo, hi); return 0; }
int tvkr_set_bg_colors(tvkr_renderer* h, const float top[3], const float bottom[3]) { if (!h || !top || !bottom) return 1; R(h)->SetBGColors(top, bottom); return 0; }
int tvkr_set_blend_precision(tvkr_renderer* h, int) { return h ? 0 : 1; }
int tvkr_resize(tvkr_renderer* h, uint32_t w, uint32_t hh) { if (!h) return 1; R(h)->Resize(w, hh); return 0; }
int tvkr_set_render_mode(tvkr_renderer* h, int mode) { return h && R(h)->SetRendermode(mode) ? 0 : 1; }
int tvkr_set_lighting_enabled(tvkr_renderer* h, int on) { if (!h) return 1; R(h)->SetUseLighting(on != 0); return 0; }
int tvkr_set_sample_rate_modifier(tvkr_renderer* h, float f) { if (!h) return 1; R(h)->SetSampleRateModifier(f); return 0; }
int tvkr_set_fov(tvkr_renderer* h, float f) { if (!h) return 1; R(h)->SetFoV(f); return 0; }
int tvkr_set_view_pos(tvkr_renderer* h, const float p[3]) { if (!h) return 1; R(h)->SetViewPos(Vec3(p[0], p[1], p[2])); return 0; }
int tvkr_set_view_dir(tvkr_renderer* h, const float p[3]) { if (!h) return 1; R(h)->SetViewDir(Vec3(p[0], p[1], p[2])); return 0; }
int tvkr_set_up_dir(tvkr_renderer* h, const float p[3]) { if (!h) return 1; R(h)->SetUpDir(Vec3(p[0], p[1], p[2])); return 0; }
int tvkr_set_view_parameters(tvkr_renderer* h, float angle, float znear, float zfar, const float eye[3], const float ref[3], const float up[3]) {
  if (!h) return 1;
  R(h)->SetViewParameters(angle, znear, zfar, Vec3(eye[0], eye[1], eye[2]), Vec3(ref[0], ref[1], ref[2]), Vec3(up[0], up[1], up[2]));
  return 0;
}
int tvkr_set_rotation4x4(tvkr_renderer* h, const float m[16]) { if (!h) return 1; Matrix4 t; memcpy(t.m, m, 64); R(h)->SetRotation4x4(t); return 0; }
int tvkr_set_translation4x4(tvkr_renderer* h, const float m[16]) { if (!h) return 1; Matrix4 t; memcpy(t.m, m, 64); R(h)->SetTranslation4x4(t); return 0; }
int tvkr_set_iso_value(tvkr_renderer* h, float v) { if (!h) return 1; R(h)->SetIsoValue(v); return 0; }
int tvkr_set_iso_value_relative(tvkr_renderer* h, float v) { if (!h) return 1; R(h)->SetIsoValueRelative(v); return 0; }
int tvkr_set_isosurface_color(tvkr_renderer* h, const float c[3]) { if (!h) return 1; R(h)->SetIsosurfaceColor(Vec3(c[0], c[1], c[2])); return 0; }
int tvkr_set_1d_trans(tvkr_renderer* h, const uint8_t* rgba, uint32_t n) {
  if (!h || !rgba) return 1;
  return R(h)->Set1DTrans(std::vector<uint8_t>(rgba, rgba + 4 * (size_t)n)) ? 0 : 1;
}
int tvkr_set_1d_trans_float(tvkr_renderer* h, const float* rgba, uint32_t n) {
  if (!h || !rgba) return 1;
  R(h)->Get1DTrans().Set(rgba, n);
  R(h)->Changed1DTrans();
  return 0;
}
int tvkr_load_1d_trans(tvkr_renderer* h, const char* filename) {
  if (!h || !filename || !R(h)->Get1DTrans().Load(filename)) return 1;
  R(h)->Changed1DTrans();
  return 0;
}
int tvkr_render_2d_view(tvkr_renderer* h, int window_mode, uint64_t slice_index, int use_mip, int flip_x, int flip_y) {
  return h && R(h)->Render2DView(window_mode, slice_index, use_mip != 0, flip_x != 0, flip_y != 0) ? 0 : 1;
}
int tvkr_render_hq_mip(tvkr_renderer* h, int window_mode, int flip_x, int flip_y, float angle_deg, int ortho, int mip_lod) {
  return h && R(h)->RenderHQMIP(window_mode, flip_x != 0, flip_y != 0, angle_deg, ortho != 0, mip_lod != 0) ? 0 : 1;
}
int tvkr_plan_hq_mip(tvkr_renderer* h, int window_mode, int flip_x, int flip_y, float angle_deg, int ortho, int mip_lod, tvk_mip_params* p,
                     tvk_classic_brick* bricks, uint32_t cap, uint32_t* n, uint32_t* lod) {
  if (!h || !p || !n) return 1;
  std::vector<tvk_classic_brick> b;
  uint32_t l = 0;
  if (!R(h)->PlanHQMIP(window_mode, flip_x != 0, flip_y != 0, angle_deg, ortho != 0, mip_lod != 0, *p, b, l)) return 1;
  *n = (uint32_t)b.size();
  if (lod) *lod = l;
  if (bricks) { if (cap < b.size()) return 1; memcpy(bricks, b.data(), b.size() * sizeof(tvk_classic_brick)); }
  return 0;
}
int tvkr_get_slice_params(tvkr_renderer* h, int window_mode, uint64_t slice_index, int use_mip, int flip_x, int flip_y, tvk_slice_params* out) {
  uint64_t toc = 0;
  return h && out && R(h)->ComputeSliceParams(window_mode, slice_index, use_mip != 0, flip_x != 0, flip_y != 0, *out, toc) ? 0 : 1;
}
int tvkr_load_2d_trans(tvkr_renderer* h, const char* filename) { return h && filename && R(h)->Load2DTrans(filename) ? 0 : 1; }
int tvkr_tf2d_render(const char* filename, uint32_t w, uint32_t hh, uint8_t* rgba, double nz_limits[4], const char* save_as) {
  if (!filename || !rgba || !nz_limits || !w || !hh) return 1;
  tvk::TransferFunction2D tf;
  if (!tf.Load(filename, w, hh)) return 1;
  std::vector<uint8_t> t;
  tf.Render(t, nz_limits);
  memcpy(rgba, t.data(), t.size());
  if (save_as && *save_as) {                                                  // Load(file) keeps the file's own size: the round trip Save writes
    tvk::TransferFunction2D own;
    if (!own.Load(filename) || !own.Save(save_as)) return 1;
  }
  return 0;
}
int tvkr_set_2d_trans(tvkr_renderer* h, const uint8_t* rgba, uint32_t w, uint32_t hh, const double nz_limits[4]) {
  double nz[4] = {nz_limits[0], nz_limits[1], nz_limits[2], nz_limits[3]};
  return h && R(h)->Set2DTrans(rgba, w, hh, nz) ? 0 : 1;
}
int tvkr_set_state(tvkr_renderer* h, uint32_t strategy, uint32_t hash_size, uint32_t rehash, uint32_t sampler) {
  if (!h || hash_size == 0 || rehash == 0 || strategy > TVK_MBS_SKIP_TWO_LEVELS || (sampler > TVK_SAMPLER_GATHER && sampler != tvk::kSamplerAuto)) return 1;
  R(h)->SetMissingBrickStrategy(strategy); R(h)->SetHashTable(hash_size, rehash); R(h)->SetSampler(sampler);
  return 0;
}
int tvkr_paint(tvkr_renderer* h) { return h && R(h)->Paint() ? 0 : 1; }
int tvkr_check_for_redraw(tvkr_renderer* h) { return h && R(h)->CheckForRedraw() ? 1 : 0; }
int tvkr_converged(tvkr_renderer* h) { return h && R(h)->Converged() ? 1 : 0; }
int tvkr_paint_until_converged(tvkr_renderer* h, uint32_t max_subframes, uint32_t* subframes) {
  uint32_t s = 0;
  const bool ok = h && R(h)->PaintUntilConverged(max_subframes, s);
  if (subframes) *subframes = s;
  return ok ? 0 : 1;
}
int tvkr_clear_working_set(tvkr_renderer* h) { if (!h) return 1; R(h)->ClearWorkingSet(); return 0; }
int tvkr_recalc_visibility(tvkr_renderer* h) { return h && R(h)->RecalcVisibility() ? 0 : 1; }
int tvkr_get_frame_rgba8(tvkr_renderer* h, uint8_t* out) {
  return (h && out && R(h)->GetFrameRGBA8(out)) ? 0 : 1;
}
int tvkr_get_frame_float(tvkr_renderer* h, float* out) {
  if (!h || !out) return 1;
  return tvk_read_image(R(h)->Context(), out) == TVK_OK ? 0 : 1;
}
int tvkr_capture_single_frame(tvkr_renderer* h, const char* filename, int preserve_transparency) {
  return h && filename && R(h)->CaptureSingleFrame(filename, preserve_transparency != 0) ? 0 : 1;
}
int tvkr_get_raycast_params(tvkr_renderer* h, tvk_raycast_params* p) { return h && p && R(h)->ComputeRaycastParams(*p) ? 0 : 1; }
int tvkr_get_dataset_info(tvkr_renderer* h, tvkr_dataset_info* o) {
  if (!h || !o || !R(h)->Volume()) return 1;
  const tvk::BrickedVolume* v = R(h)->Volume();
  memset(o, 0, sizeof(*o));
  for (int a = 0; a < 3; ++a) { o->domain[a] = v->domain[a]; o->pool_capacity[a] = R(h)->Pool()->Capacity()[a]; }
  o->lod_count = R(h)->Pool()->LodCount();
  o->total_bricks = R(h)->Pool()->TotalBricks();
  o->tf_size = (uint32_t)R(h)->Get1DTrans().GetSize();
  v->Range(o->range[0], o->range[1]);
  o->max_grad_magnitude = v->max_grad_magnitude;
  o->iso_value = R(h)->GetIsoValue();
  o->out_of_core = v->host_store ? 1 : 0;
  return 0;
}
int tvkr_get_pool_slots(tvkr_renderer* h, int32_t* brick, uint64_t* time, uint32_t cap, uint32_t* n) {
  if (!h || !R(h)->Pool() || !n) return 1;
  const std::vector<tvk::PoolSlot>& s = R(h)->Pool()->Slots();
  *n = (uint32_t)s.size();
  if (cap < s.size()) return 1;
  // reported in POSITION order (the vector itself is kept in LRU order after a paging pass)
  for (const tvk::PoolSlot& p : s) { if (brick) brick[p.pos] = p.brick; if (time) time[p.pos] = p.time; }
  return 0;
}
int tvkr_get_perf(tvkr_renderer* h, double out[6]) {
  if (!h || !out) return 1;
  const tvk::PerfCounters& p = R(h)->Perf();
  out[0] = p.subframes; out[1] = p.raycast_ms; out[2] = p.paged_bricks; out[3] = p.uploaded_bytes; out[4] = p.samples; out[5] = p.frames;
  return 0;
}
int tvkr_compute_max_lod_for_view(tvkr_lod_state* s) {
  if (!s) return -1;
  tvk::CudaGridLeaper::ComputeMaxLODForCurrentView(*s);
  return 0;
}
int tvkr_set_renderer_target(tvkr_renderer* h, int target) {
  if (!h || target < tvk::RT_INTERACTIVE || target > tvk::RT_HEADLESS) return 1;
  R(h)->SetRendererTarget(target);
  return 0;
}
int tvkr_set_max_ms_per_frame(tvkr_renderer* h, float ms) { if (!h) return 1; R(h)->SetMaxMSPerFrame(ms); return 0; }
int tvkr_override_subframe_ms(tvkr_renderer* h, float ms) { if (!h) return 1; R(h)->OverrideSubframeMs(ms); return 0; }
int tvkr_get_lod_state(tvkr_renderer* h, tvkr_lod_state* out, uint32_t* lod) {
  if (!h || !out) return 1;
  *out = R(h)->LodState();
  if (lod) *lod = R(h)->CurrentLOD();
  return 0;
}
int tvkr_tf1d_resample(const float* rgba, uint32_t n, uint32_t target, float* out) {
  if (!rgba || !out || !n || !target) return -1;
  tvk::TransferFunction1D tf(n);
  tf.Set(rgba, n);
  tf.Resample(target);
  memcpy(out, tf.ColorData().data(), sizeof(float) * 4 * target);
  return 0;
}
int tvkr_get_histogram1d(tvkr_renderer* h, uint32_t* out, uint32_t cap, uint32_t* n) {
  // Dataset::Get1DHistogram as the TF editor reads it: min(file bins, 4096, 2^bits) bins (uvfDataset.cpp:755-757)
  if (!h || !n || !R(h)->Volume()) return 1;
  const tvk::BrickedVolume& v = *R(h)->Volume();
  const size_t target = std::min<size_t>(v.hist1d.size(), std::min<size_t>(4096, size_t(1) << (8 * v.bpv)));
  *n = (uint32_t)target;
  if (!out) return 0;
  if (cap < target) return 1;
  std::vector<uint32_t> o;
  tvk::ResampleHistogram1D(v.hist1d, target, o);
  memcpy(out, o.data(), sizeof(uint32_t) * target);
  return 0;
}
int tvkr_resample_hist1d(const uint64_t* bins, uint64_t n, uint32_t target, uint32_t* out) {
  if (!bins || !out || !target) return -1;
  std::vector<uint64_t> in(bins, bins + n);
  std::vector<uint32_t> o;
  tvk::ResampleHistogram1D(in, target, o);
  memcpy(out, o.data(), sizeof(uint32_t) * target);
  return 0;
}
int tvkr_compute_pool_capacity(const uint32_t max_brick[3], uint32_t bytes_per_voxel, uint64_t total_bricks, uint64_t gpu_mem_bytes,
                               uint32_t max_3d_texture, uint32_t capacity[3]) {
  if (!max_brick || !capacity || !bytes_per_voxel || !max_brick[0] || !max_brick[1] || !max_brick[2] || !max_3d_texture) return -1;
  tvk::BrickPool::ComputePoolCapacity(max_brick, bytes_per_voxel, total_bricks, gpu_mem_bytes, max_3d_texture, capacity);
  return 0;
}
double tvkr_perf_query(tvkr_renderer* h, int counter) { return h ? R(h)->PerfQuery(counter) : 0.0; }
tvk_ctx* tvkr_context(tvkr_renderer* h) { return h ? R(h)->Context() : nullptr; }
}
