// TEST INFRASTRUCTURE ONLY.  A small C++17 rendition of the GLSL 4.20 value types, built-ins and
// texture/image objects that the reference's ray-casting shaders use, so that THE REFERENCE'S OWN
// SHADER TEXT (Tuvok/Shaders/*.glsl and the strings GLVolumePool/GLHashTable::GetShaderFragment
// generate) can be compiled by g++ and executed per pixel on the CPU.  oracle/glsl/build_glslref.py
// reads those sources where they lie under /root/reference, rewrites only the GLSL storage syntax
// (uniform / in / out / layout qualifiers, array constructors, float literal suffixes) and compiles
// the result together with this header into oracle/_ref/libglslref.so.  Nothing of the reference is
// copied into the repository; the shader logic that runs is the reference's.
//
// Numerics: every GLSL float operation is one IEEE fp32 operation in source order (the unit is
// compiled with -ffp-contract=off, literals get an `f` suffix); texture(sampler3D) is GL_LINEAR /
// CLAMP_TO_EDGE evaluated as the OpenGL 4.2 specification writes it (section 3.9.11: texel
// coordinates u - 0.5, weights frac(), fp32), on normalised-fixed-point texels (value / (2^n - 1));
// the transfer-function textures are GL_NEAREST / CLAMP_TO_EDGE RGBA8 (GPUMemMan.cpp:421-427);
// uint(float) of a negative or NaN value gives 0 (NVIDIA).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>

namespace glsl {

typedef unsigned int uint;

template <class T, int N> struct vec;
template <class T, int N, int... I> struct swz;

// ---------------------------------------------------------------- traits
template <class A> struct vt { static constexpr int n = 0; typedef A t; };
template <class T, int N> struct vt<vec<T, N>> { static constexpr int n = N; typedef T t; };
template <class T, int N, int... I> struct vt<swz<T, N, I...>> { static constexpr int n = (int)sizeof...(I); typedef T t; };
template <class A> constexpr bool is_vec = vt<A>::n > 0;
template <class A> constexpr bool is_num = std::is_arithmetic<A>::value;

// ---------------------------------------------------------------- swizzle proxies (live inside vec's union)
template <class T, int N, int... I> struct swz {
  T d[N];
  static constexpr int M = (int)sizeof...(I);
  vec<T, M> get() const { const T t[M] = {d[I]...}; vec<T, M> r; for (int k = 0; k < M; ++k) r.d[k] = t[k]; return r; }
  operator vec<T, M>() const { return get(); }
  void set(const vec<T, M>& v) { const int idx[M] = {I...}; for (int k = 0; k < M; ++k) d[idx[k]] = v.d[k]; }
  swz& operator=(const swz& o) { set(o.get()); return *this; }
  swz& operator=(const vec<T, M>& v) { set(v); return *this; }
  template <int K, int... J> swz& operator=(const swz<T, K, J...>& o) { set(o.get()); return *this; }
  template <class B> swz& operator+=(const B& b) { set(get() + b); return *this; }
  template <class B> swz& operator-=(const B& b) { set(get() - b); return *this; }
  template <class B> swz& operator*=(const B& b) { set(get() * b); return *this; }
  template <class B> swz& operator/=(const B& b) { set(get() / b); return *this; }
};

// ---------------------------------------------------------------- component gathering for constructors
template <class T> struct gather {
  T v[16]; int n = 0;
  template <class A> typename std::enable_if<is_num<A>>::type add(const A& a) { v[n++] = (T)a; }
  template <class U, int K> void add(const vec<U, K>& a) { for (int i = 0; i < K; ++i) v[n++] = (T)a.d[i]; }
  template <class U, int K, int... J> void add(const swz<U, K, J...>& a) { add(a.get()); }
};

#define GLSL_VEC_COMMON(N)                                                                              \
  vec() { for (int i = 0; i < N; ++i) d[i] = T(0); }                                                    \
  vec(const vec& o) { for (int i = 0; i < N; ++i) d[i] = o.d[i]; }                                      \
  vec& operator=(const vec& o) { for (int i = 0; i < N; ++i) d[i] = o.d[i]; return *this; }             \
  template <class... A> explicit vec(const A&... a) {                                                   \
    gather<T> g; (g.add(a), ...);                                                                       \
    if (g.n == 1) for (int i = 0; i < N; ++i) d[i] = g.v[0];                                            \
    else for (int i = 0; i < N; ++i) d[i] = g.v[i];                                                     \
  }                                                                                                     \
  T& operator[](int i) { return d[i]; }                                                                 \
  const T& operator[](int i) const { return d[i]; }                                                     \
  template <class B> vec& operator+=(const B& b) { *this = *this + b; return *this; }                   \
  template <class B> vec& operator-=(const B& b) { *this = *this - b; return *this; }                   \
  template <class B> vec& operator*=(const B& b) { *this = *this * b; return *this; }                   \
  template <class B> vec& operator/=(const B& b) { *this = *this / b; return *this; }

template <class T> struct vec<T, 2> {
  union {
    T d[2];
    struct { T x, y; };
    struct { T r, g; };
    swz<T, 2, 0, 1> xy; swz<T, 2, 1, 0> yx; swz<T, 2, 0, 1> rg;
  };
  GLSL_VEC_COMMON(2)
};
template <class T> struct vec<T, 3> {
  union {
    T d[3];
    struct { T x, y, z; };
    struct { T r, g, b; };
    swz<T, 3, 0, 1> xy; swz<T, 3, 1, 2> yz; swz<T, 3, 0, 2> xz;
    swz<T, 3, 0, 1, 2> xyz; swz<T, 3, 0, 1, 2> rgb; swz<T, 3, 2, 1, 0> zyx; swz<T, 3, 2, 1, 0> bgr;
  };
  GLSL_VEC_COMMON(3)
};
template <class T> struct vec<T, 4> {
  union {
    T d[4];
    struct { T x, y, z, w; };
    struct { T r, g, b, a; };
    swz<T, 4, 0, 1> xy; swz<T, 4, 2, 3> zw; swz<T, 4, 0, 1> rg;
    swz<T, 4, 0, 1, 2> xyz; swz<T, 4, 0, 1, 2> rgb; swz<T, 4, 1, 2, 3> yzw;
    swz<T, 4, 0, 1, 2, 3> xyzw; swz<T, 4, 0, 1, 2, 3> rgba;
    swz<T, 4, 0, 0, 0, 0> xxxx; swz<T, 4, 0, 3> xw;          // (read-only uses in MIP-slice-FS.glsl / Transfer-MIP-FS.glsl)
  };
  GLSL_VEC_COMMON(4)
};

typedef vec<float, 2> vec2;  typedef vec<float, 3> vec3;  typedef vec<float, 4> vec4;
typedef vec<int, 2> ivec2;   typedef vec<int, 3> ivec3;   typedef vec<int, 4> ivec4;
typedef vec<uint, 2> uvec2;  typedef vec<uint, 3> uvec3;  typedef vec<uint, 4> uvec4;
typedef vec<bool, 2> bvec2;  typedef vec<bool, 3> bvec3;  typedef vec<bool, 4> bvec4;

template <class T, int N> inline const vec<T, N>& val(const vec<T, N>& a) { return a; }
template <class T, int N, int... I> inline vec<T, (int)sizeof...(I)> val(const swz<T, N, I...>& a) { return a.get(); }

// ---------------------------------------------------------------- component-wise arithmetic
// vector (.) vector of the same size, vector (.) scalar and scalar (.) vector; a scalar operand takes the
// vector's component type (GLSL's int -> uint / int -> float implicit conversions).
template <class A, class B, class F, typename std::enable_if<is_vec<A> && is_vec<B>, int>::type = 0>
inline auto zip(const A& a, const B& b, F f) {
  static_assert(vt<A>::n == vt<B>::n, "vector size mismatch");
  typedef typename vt<A>::t T;
  const auto va = val(a); const auto vb = val(b);
  vec<T, vt<A>::n> r;
  for (int i = 0; i < vt<A>::n; ++i) r.d[i] = f(va.d[i], (T)vb.d[i]);
  return r;
}
template <class A, class B, class F, typename std::enable_if<is_vec<A> && is_num<B>, int>::type = 0>
inline auto zip(const A& a, const B& b, F f) {
  typedef typename vt<A>::t T;
  const auto va = val(a);
  vec<T, vt<A>::n> r;
  for (int i = 0; i < vt<A>::n; ++i) r.d[i] = f(va.d[i], (T)b);
  return r;
}
template <class A, class B, class F, typename std::enable_if<is_num<A> && is_vec<B>, int>::type = 0>
inline auto zip(const A& a, const B& b, F f) {
  typedef typename vt<B>::t T;
  const auto vb = val(b);
  vec<T, vt<B>::n> r;
  for (int i = 0; i < vt<B>::n; ++i) r.d[i] = f((T)a, vb.d[i]);
  return r;
}
#define GLSL_BINOP(OP)                                                                                        \
  template <class A, class B, typename std::enable_if<(is_vec<A> && (is_vec<B> || is_num<B>)) || (is_num<A> && is_vec<B>), int>::type = 0> \
  inline auto operator OP(const A& a, const B& b) { return zip(a, b, [](auto x, auto y) { return (decltype(x))(x OP y); }); }
GLSL_BINOP(+) GLSL_BINOP(-) GLSL_BINOP(*) GLSL_BINOP(/) GLSL_BINOP(%)
#undef GLSL_BINOP
template <class A, typename std::enable_if<is_vec<A>, int>::type = 0>
inline auto operator-(const A& a) { auto r = val(a); for (int i = 0; i < vt<A>::n; ++i) r.d[i] = -r.d[i]; return vec<typename vt<A>::t, vt<A>::n>(r); }
template <class A, class B, typename std::enable_if<is_vec<A> && is_vec<B>, int>::type = 0>
inline bool operator==(const A& a, const B& b) {
  const auto va = val(a); const auto vb = val(b);
  for (int i = 0; i < vt<A>::n; ++i) if (!(va.d[i] == vb.d[i])) return false;
  return true;
}
template <class A, class B, typename std::enable_if<is_vec<A> && is_vec<B>, int>::type = 0>
inline bool operator!=(const A& a, const B& b) { return !(a == b); }

// ---------------------------------------------------------------- scalar built-ins
// GLSL's implicit conversions for two-argument built-ins: float wins, then uint, then int
template <class A, class B> struct common2 {
  typedef typename std::conditional<std::is_floating_point<A>::value || std::is_floating_point<B>::value, float,
          typename std::conditional<std::is_unsigned<A>::value || std::is_unsigned<B>::value, uint, int>::type>::type type;
};
template <class A, class B, typename std::enable_if<is_num<A> && is_num<B>, int>::type = 0>
inline typename common2<A, B>::type min(A a, B b) { typedef typename common2<A, B>::type T; return (T)b < (T)a ? (T)b : (T)a; }
template <class A, class B, typename std::enable_if<is_num<A> && is_num<B>, int>::type = 0>
inline typename common2<A, B>::type max(A a, B b) { typedef typename common2<A, B>::type T; return (T)a < (T)b ? (T)b : (T)a; }
inline float abs(float a) { return std::fabs(a); }
inline int abs(int a) { return a < 0 ? -a : a; }
inline float floor(float a) { return std::floor(a); }
inline float ceil(float a) { return std::ceil(a); }
inline float fract(float a) { return a - std::floor(a); }
inline float sqrt(float a) { return std::sqrt(a); }
inline float inversesqrt(float a) { return 1.0f / std::sqrt(a); }
inline float pow(float a, float b) { return std::pow(a, b); }
inline float exp2(float a) { return std::exp2(a); }
inline float log2(float a) { return std::log2(a); }
inline float exp(float a) { return std::exp(a); }
inline float log(float a) { return std::log(a); }
inline float clamp(float v, float lo, float hi) { return min(max(v, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }      // x*(1-a) + y*a, as the spec writes it
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float sign(float a) { return a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f); }
// conversions the shaders spell as constructors
inline uint to_uint(float f) { if (!(f > 0.0f)) return 0u; if (f >= 4294967296.0f) return 0xFFFFFFFFu; return (uint)f; }
inline uint to_uint(double f) { return to_uint((float)f); }
template <class A, typename std::enable_if<std::is_integral<A>::value, int>::type = 0> inline uint to_uint(A a) { return (uint)a; }

// ---------------------------------------------------------------- vector built-ins
#define GLSL_MAP1(NAME)                                                                                         \
  template <class A, typename std::enable_if<is_vec<A>, int>::type = 0> inline auto NAME(const A& a) {          \
    auto r = val(a); vec<typename vt<A>::t, vt<A>::n> o; for (int i = 0; i < vt<A>::n; ++i) o.d[i] = NAME(r.d[i]); return o; }
GLSL_MAP1(abs) GLSL_MAP1(floor) GLSL_MAP1(ceil) GLSL_MAP1(fract) GLSL_MAP1(sqrt) GLSL_MAP1(sign)
#undef GLSL_MAP1
template <class A, class B, typename std::enable_if<is_vec<A> && (is_vec<B> || is_num<B>), int>::type = 0>
inline auto min(const A& a, const B& b) { return zip(a, b, [](auto x, auto y) { return y < x ? y : x; }); }
template <class A, class B, typename std::enable_if<is_vec<A> && (is_vec<B> || is_num<B>), int>::type = 0>
inline auto max(const A& a, const B& b) { return zip(a, b, [](auto x, auto y) { return x < y ? y : x; }); }
template <class A, class L, class H, typename std::enable_if<is_vec<A>, int>::type = 0>
inline auto clamp(const A& v, const L& lo, const H& hi) { return min(max(v, lo), hi); }
template <class A, class B, class C, typename std::enable_if<is_vec<A> && is_vec<B>, int>::type = 0>
inline auto mix(const A& a, const B& b, const C& t) { return a * (1.0f - t) + b * t; }
template <class E, class A, typename std::enable_if<is_vec<A> && (is_num<E> || is_vec<E>), int>::type = 0>
inline auto step(const E& edge, const A& x) { return zip(x, edge, [](auto xv, auto e) { return xv < e ? (decltype(xv))0 : (decltype(xv))1; }); }
template <class A, class B, typename std::enable_if<is_vec<A> && is_vec<B>, int>::type = 0>
inline float dot(const A& a, const B& b) {
  const auto va = val(a); const auto vb = val(b);
  float s = va.d[0] * vb.d[0];
  for (int i = 1; i < vt<A>::n; ++i) s = s + va.d[i] * vb.d[i];
  return s;
}
template <class A, typename std::enable_if<is_vec<A>, int>::type = 0> inline float length(const A& a) { return std::sqrt(dot(a, a)); }
inline float length(float a) { return std::fabs(a); }
template <class A, class B, typename std::enable_if<is_vec<A> && is_vec<B>, int>::type = 0>
inline float distance(const A& a, const B& b) { return length(a - b); }
template <class A, typename std::enable_if<is_vec<A>, int>::type = 0> inline auto normalize(const A& a) { return a / length(a); }
template <class A, class B, typename std::enable_if<is_vec<A> && is_vec<B>, int>::type = 0>
inline auto reflect(const A& i, const B& n) { return i - 2.0f * dot(n, i) * n; }
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }

// ---------------------------------------------------------------- mat4 (column major, as glUniformMatrix4fv(transpose = GL_FALSE) loads 16 floats)
struct mat4 {
  float m[16];
  mat4() { std::memset(m, 0, sizeof(m)); }
  explicit mat4(const float* p) { std::memcpy(m, p, sizeof(m)); }
};
typedef mat4 mat4x4;
inline vec4 operator*(const mat4& M, const vec4& v) {
  vec4 r;
  for (int row = 0; row < 4; ++row)
    r.d[row] = ((M.m[row] * v.d[0] + M.m[4 + row] * v.d[1]) + M.m[8 + row] * v.d[2]) + M.m[12 + row] * v.d[3];
  return r;
}

// ---------------------------------------------------------------- textures and images
struct sampler1D {                       // RGBA8, GL_NEAREST, GL_CLAMP_TO_EDGE (the 1-D transfer function)
  const uint8_t* rgba = nullptr; int w = 0;
};
struct sampler2D {                       // RGBA32F render-target texture (texelFetch) or RGBA8 NEAREST table (texture)
  const float* f = nullptr; const uint8_t* rgba = nullptr; int w = 0, h = 0;
};
struct sampler3D {                       // single-channel UNORM8/16, GL_LINEAR, GL_CLAMP_TO_EDGE (the brick pool)
  const void* data = nullptr; int bpv = 1; int sx = 0, sy = 0, sz = 0;
};
struct usampler3D {                      // R32UI, texelFetch only (pool metadata)
  const uint32_t* data = nullptr; int sx = 0, sy = 0, sz = 0;
};
struct uimage1D { uint32_t* data = nullptr; int n = 0; };
struct uimage2D { uint32_t* data = nullptr; int w = 0, h = 0; };

inline int nearest_texel(float s, int n) {          // GL 4.2 3.9.11: i = floor(u), clamped to [0, n-1]
  const float u = s * (float)n;
  int i = (int)std::floor(u);
  return i < 0 ? 0 : (i > n - 1 ? n - 1 : i);
}
inline vec4 rgba8(const uint8_t* p) { return vec4((float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, (float)p[3] / 255.0f); }
inline vec4 texture(const sampler1D& s, float c) { return rgba8(s.rgba + 4 * (size_t)nearest_texel(c, s.w)); }
inline vec4 texture(const sampler2D& s, const vec2& c) {
  const int i = nearest_texel(c.x, s.w), j = nearest_texel(c.y, s.h);
  if (s.rgba) return rgba8(s.rgba + 4 * ((size_t)j * s.w + i));
  const float* p = s.f + 4 * ((size_t)j * s.w + i);
  return vec4(p[0], p[1], p[2], p[3]);
}
inline vec4 texture2D(const sampler2D& s, const vec2& c) { return texture(s, c); }
inline vec4 texelFetch(const sampler2D& s, const ivec2& p, int) {
  const float* q = s.f + 4 * ((size_t)p.y * s.w + p.x);
  return vec4(q[0], q[1], q[2], q[3]);
}
inline uvec4 texelFetch(const usampler3D& s, const ivec3& p, int) {
  return uvec4(s.data[(size_t)p.x + (size_t)p.y * s.sx + (size_t)p.z * s.sx * s.sy], 0u, 0u, 1u);
}
inline float texel3(const sampler3D& s, int x, int y, int z) {
  x = x < 0 ? 0 : (x > s.sx - 1 ? s.sx - 1 : x);
  y = y < 0 ? 0 : (y > s.sy - 1 ? s.sy - 1 : y);
  z = z < 0 ? 0 : (z > s.sz - 1 ? s.sz - 1 : z);
  const size_t i = (size_t)x + (size_t)y * s.sx + (size_t)z * s.sx * s.sy;
  return s.bpv == 1 ? (float)((const uint8_t*)s.data)[i] / 255.0f : (float)((const uint16_t*)s.data)[i] / 65535.0f;
}
inline vec4 texture(const sampler3D& s, const vec3& c) {          // GL 4.2 3.9.11, TEXTURE_MIN/MAG_FILTER = LINEAR
  const float u = c.x * (float)s.sx - 0.5f, v = c.y * (float)s.sy - 0.5f, w = c.z * (float)s.sz - 0.5f;
  const float fu = std::floor(u), fv = std::floor(v), fw = std::floor(w);
  const int i0 = (int)fu, j0 = (int)fv, k0 = (int)fw;
  const float a = u - fu, b = v - fv, g = w - fw;
  // tau = (1-a)(1-b)(1-g) t000 + a(1-b)(1-g) t100 + ... (equation 3.27)
  const float t =
      (1.0f - a) * (1.0f - b) * (1.0f - g) * texel3(s, i0, j0, k0) + a * (1.0f - b) * (1.0f - g) * texel3(s, i0 + 1, j0, k0) +
      (1.0f - a) * b * (1.0f - g) * texel3(s, i0, j0 + 1, k0) + a * b * (1.0f - g) * texel3(s, i0 + 1, j0 + 1, k0) +
      (1.0f - a) * (1.0f - b) * g * texel3(s, i0, j0, k0 + 1) + a * (1.0f - b) * g * texel3(s, i0 + 1, j0, k0 + 1) +
      (1.0f - a) * b * g * texel3(s, i0, j0 + 1, k0 + 1) + a * b * g * texel3(s, i0 + 1, j0 + 1, k0 + 1);
  return vec4(t, t, t, 1.0f);            // GL_LUMINANCE / GL_RED: the shaders read .r
}
inline uint imageAtomicCompSwap(uimage1D& img, int p, uint compare, uint data) {
  const uint old = img.data[p];
  if (old == compare) img.data[p] = data;
  return old;
}
inline uint imageAtomicCompSwap(uimage2D& img, const ivec2& p, uint compare, uint data) {
  uint32_t& cell = img.data[(size_t)p.y * img.w + p.x];
  const uint old = cell;
  if (old == compare) cell = data;
  return old;
}

struct discard_exception {};

}  // namespace glsl
