// TEST INFRASTRUCTURE ONLY (oracle build). Never linked into the product.
//
// Minimal stand-in for the un-vendored third-party `glm` (version unpinned by the
// reference: it arrives through TinyEngine 1.7, SimpleWindErosion.cpp:1,16) so that the
// reference hot-path headers compile headless, unmodified. Only what these call sites
// use is provided:
//   source/wind.hpp:6,13,30,56,61,70,75,79,83,87,91,100,104
//   source/world.hpp:52,54,83-92,103,107,112
//   source/cellpool.h:52,204-220,248-254,304-317,343-352,381-384,432-445
//   source/include/math.h:11-21
//
// Semantics restated from glm 0.9.9's scalar (non-SIMD) code paths -- an ASSUMPTION that
// cannot be verified in this container (glm is absent); it is cross-checked only by the
// survey's independent probe fingerprints:
//   dot(a,b)      : t = a*b ; (t.x + t.y) + t.z            (left-to-right)
//   length(v)     : sqrt(dot(v,v))
//   normalize(v)  : v * (1 / sqrt(dot(v,v)))                (inversesqrt = 1/sqrt)
//   cross(x,y)    : (x.y*y.z - y.y*x.z, x.z*y.x - y.z*x.x, x.x*y.y - y.x*x.y)
//   mix(x,y,a)    : x*(1-a) + y*a
//   round(v)      : std::round per component (half away from zero)
//   vecN<float> -> vecN<int> : C truncation; all vec conversions implicit
//   v *= U scalar : each component *= static_cast<T>(scalar)
#pragma once
#include <cmath>
#include <cstddef>
#include <type_traits>

namespace glm {

template <typename T> struct tvec2 {
  T x, y;
  constexpr tvec2() : x(0), y(0) {}
  template <typename A> requires std::is_arithmetic_v<A>
  constexpr tvec2(A s) : x(static_cast<T>(s)), y(static_cast<T>(s)) {}
  template <typename A, typename B>
  constexpr tvec2(A a, B b) : x(static_cast<T>(a)), y(static_cast<T>(b)) {}
  template <typename U>
  constexpr tvec2(const tvec2<U>& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}

  tvec2& operator+=(const tvec2& o) { x += o.x; y += o.y; return *this; }
  tvec2& operator-=(const tvec2& o) { x -= o.x; y -= o.y; return *this; }
  tvec2& operator/=(const tvec2& o) { x /= o.x; y /= o.y; return *this; }
  tvec2& operator*=(const tvec2& o) { x *= o.x; y *= o.y; return *this; }
};

template <typename T> struct tvec3 {
  T x, y, z;
  constexpr tvec3() : x(0), y(0), z(0) {}
  template <typename A> requires std::is_arithmetic_v<A>
  constexpr tvec3(A s) : x(static_cast<T>(s)), y(static_cast<T>(s)), z(static_cast<T>(s)) {}
  template <typename A, typename B, typename C>
  constexpr tvec3(A a, B b, C c)
      : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)) {}
  template <typename U>
  constexpr tvec3(const tvec3<U>& v)
      : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}

  tvec3& operator+=(const tvec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
  tvec3& operator-=(const tvec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
  template <typename U> tvec3& operator*=(U s) {
    x *= static_cast<T>(s); y *= static_cast<T>(s); z *= static_cast<T>(s);
    return *this;
  }
};

typedef tvec2<float> vec2;
typedef tvec2<int> ivec2;
typedef tvec3<float> vec3;
typedef tvec3<int> ivec3;

// ---- vec2 arithmetic -------------------------------------------------------
template <typename T> tvec2<T> operator+(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x + b.x, a.y + b.y); }
template <typename T> tvec2<T> operator-(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x - b.x, a.y - b.y); }
template <typename T> tvec2<T> operator*(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x * b.x, a.y * b.y); }
template <typename T> tvec2<T> operator/(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x / b.x, a.y / b.y); }
template <typename T> tvec2<T> operator*(T s, const tvec2<T>& a) { return tvec2<T>(s * a.x, s * a.y); }
template <typename T> tvec2<T> operator*(const tvec2<T>& a, T s) { return tvec2<T>(a.x * s, a.y * s); }
template <typename T> tvec2<T> operator/(const tvec2<T>& a, T s) { return tvec2<T>(a.x / s, a.y / s); }
// mixed int-vector (+) float-vector as used by `vec2 p = ivec2 + ...` never occurs; the
// reference only mixes via conversion on assignment.

// ---- vec3 arithmetic -------------------------------------------------------
template <typename T> tvec3<T> operator+(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> tvec3<T> operator-(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> tvec3<T> operator*(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <typename T> tvec3<T> operator*(T s, const tvec3<T>& a) { return tvec3<T>(s * a.x, s * a.y, s * a.z); }
template <typename T> tvec3<T> operator*(const tvec3<T>& a, T s) { return tvec3<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> tvec3<T> operator/(const tvec3<T>& a, T s) { return tvec3<T>(a.x / s, a.y / s, a.z / s); }
template <typename T> tvec3<T> operator-(const tvec3<T>& a, T s) { return tvec3<T>(a.x - s, a.y - s, a.z - s); }
template <typename T> tvec3<T> operator+(const tvec3<T>& a, T s) { return tvec3<T>(a.x + s, a.y + s, a.z + s); }

// ---- functions -------------------------------------------------------------
inline float dot(const vec2& a, const vec2& b) { vec2 t = a * b; return t.x + t.y; }
inline float dot(const vec3& a, const vec3& b) { vec3 t = a * b; return t.x + t.y + t.z; }
inline float length(const vec2& v) { return std::sqrt(dot(v, v)); }
inline float length(const vec3& v) { return std::sqrt(dot(v, v)); }
inline vec3 normalize(const vec3& v) { return v * (1.0f / std::sqrt(dot(v, v))); }
inline vec3 cross(const vec3& x, const vec3& y) {
  return vec3(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
inline vec3 mix(const vec3& x, const vec3& y, float a) { return x * (1.0f - a) + y * a; }
inline vec2 round(const vec2& v) { return vec2(std::round(v.x), std::round(v.y)); }

}  // namespace glm
