// Minimal stand-in for the subset of GLM that the reference hot path uses.
//
// TEST INFRASTRUCTURE ONLY (oracle/): this header exists so that the UNMODIFIED
// reference sources under /root/reference/src can be compiled in this image,
// where GLM (a vcpkg dependency of the reference, vcpkg.json:5-27, pinned only
// through builtin-baseline 73e9c8e7...) is not installed and cannot be fetched.
// It is written from GLM's documented/public semantics, not copied from GLM:
//   min(x,y)      = (y < x) ? y : x            max(x,y) = (x < y) ? y : x
//   sign(x)       = (0 < x) - (x < 0)
//   dot(a,b)      = a.x*b.x + a.y*b.y + a.z*b.z (left to right)
//   normalize(v)  = v * inversesqrt(dot(v,v)),  inversesqrt(x) = 1/sqrt(x)
//   cross(a,b)    = (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y)
//   mix(x,y,a)    = x*(1-a) + y*a
//   clamp(x,lo,hi)= min(max(x,lo),hi)
//   smoothstep    : t = clamp((x-e0)/(e1-e0),0,1); t*t*(3-2t)
//   vec -> ivec conversion truncates toward zero (static_cast per component)
// Because the real GLM is absent, parity against "the reference" means parity
// against the reference sources compiled with THIS shim.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstddef>

#ifdef __CUDACC__
#define GLM_FN __host__ __device__ inline
#else
#define GLM_FN inline
#endif
#define GLM_CFN GLM_FN constexpr

namespace glm {

template <typename T> struct tvec2;
template <typename T> struct tvec3;
template <typename T> struct tvec4;

template <typename T> struct tvec2 {
	union { T x, r, s; };
	union { T y, g, t; };
	tvec2() = default;
	tvec2(const tvec2&) = default;
	tvec2& operator=(const tvec2&) = default;
	template <typename A> GLM_CFN explicit tvec2(A a) : x(static_cast<T>(a)), y(static_cast<T>(a)) {}
	template <typename A, typename B> GLM_CFN tvec2(A a, B b) : x(static_cast<T>(a)), y(static_cast<T>(b)) {}
	template <typename U> GLM_CFN tvec2(const tvec2<U>& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
	GLM_CFN T& operator[](int i) { return i == 0 ? x : y; }
	GLM_CFN const T& operator[](int i) const { return i == 0 ? x : y; }
	template <typename U> GLM_CFN tvec2& operator+=(const tvec2<U>& v) { x += static_cast<T>(v.x); y += static_cast<T>(v.y); return *this; }
	template <typename U> GLM_CFN tvec2& operator-=(const tvec2<U>& v) { x -= static_cast<T>(v.x); y -= static_cast<T>(v.y); return *this; }
	template <typename U> GLM_CFN tvec2& operator*=(const tvec2<U>& v) { x *= static_cast<T>(v.x); y *= static_cast<T>(v.y); return *this; }
	template <typename U> GLM_CFN tvec2& operator*=(U s) { x *= static_cast<T>(s); y *= static_cast<T>(s); return *this; }
	template <typename U> GLM_CFN tvec2& operator/=(U s) { x /= static_cast<T>(s); y /= static_cast<T>(s); return *this; }
};

template <typename T> struct tvec3 {
	union { T x, r, s; };
	union { T y, g, t; };
	union { T z, b, p; };
	tvec3() = default;
	tvec3(const tvec3&) = default;
	tvec3& operator=(const tvec3&) = default;
	template <typename A> GLM_CFN explicit tvec3(A a) : x(static_cast<T>(a)), y(static_cast<T>(a)), z(static_cast<T>(a)) {}
	template <typename A, typename B, typename C> GLM_CFN tvec3(A a, B b, C c) : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)) {}
	template <typename U> GLM_CFN tvec3(const tvec3<U>& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}
	GLM_CFN T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
	GLM_CFN const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
	template <typename U> GLM_CFN tvec3& operator+=(const tvec3<U>& v) { x += static_cast<T>(v.x); y += static_cast<T>(v.y); z += static_cast<T>(v.z); return *this; }
	template <typename U> GLM_CFN tvec3& operator-=(const tvec3<U>& v) { x -= static_cast<T>(v.x); y -= static_cast<T>(v.y); z -= static_cast<T>(v.z); return *this; }
	template <typename U> GLM_CFN tvec3& operator*=(const tvec3<U>& v) { x *= static_cast<T>(v.x); y *= static_cast<T>(v.y); z *= static_cast<T>(v.z); return *this; }
	template <typename U> GLM_CFN tvec3& operator/=(const tvec3<U>& v) { x /= static_cast<T>(v.x); y /= static_cast<T>(v.y); z /= static_cast<T>(v.z); return *this; }
	template <typename U> GLM_CFN tvec3& operator*=(U s) { x *= static_cast<T>(s); y *= static_cast<T>(s); z *= static_cast<T>(s); return *this; }
	template <typename U> GLM_CFN tvec3& operator/=(U s) { x /= static_cast<T>(s); y /= static_cast<T>(s); z /= static_cast<T>(s); return *this; }
};

template <typename T> struct tvec4 {
	union { T x, r, s; };
	union { T y, g, t; };
	union { T z, b, p; };
	union { T w, a, q; };
	tvec4() = default;
	tvec4(const tvec4&) = default;
	tvec4& operator=(const tvec4&) = default;
	template <typename A> GLM_CFN explicit tvec4(A v) : x(static_cast<T>(v)), y(static_cast<T>(v)), z(static_cast<T>(v)), w(static_cast<T>(v)) {}
	template <typename A, typename B, typename C, typename D> GLM_CFN tvec4(A a_, B b_, C c_, D d_) : x(static_cast<T>(a_)), y(static_cast<T>(b_)), z(static_cast<T>(c_)), w(static_cast<T>(d_)) {}
	GLM_CFN T& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
	GLM_CFN const T& operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};

typedef tvec2<float> vec2;
typedef tvec3<float> vec3;
typedef tvec4<float> vec4;
typedef tvec2<int> ivec2;
typedef tvec3<int> ivec3;
typedef tvec4<int> ivec4;
typedef tvec3<unsigned int> uvec3;

// ---- vec2 operators
template <typename T> GLM_CFN tvec2<T> operator+(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x + b.x, a.y + b.y); }
template <typename T> GLM_CFN tvec2<T> operator-(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x - b.x, a.y - b.y); }
template <typename T> GLM_CFN tvec2<T> operator*(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x * b.x, a.y * b.y); }
template <typename T> GLM_CFN tvec2<T> operator/(const tvec2<T>& a, const tvec2<T>& b) { return tvec2<T>(a.x / b.x, a.y / b.y); }
template <typename T> GLM_CFN tvec2<T> operator*(const tvec2<T>& a, T s) { return tvec2<T>(a.x * s, a.y * s); }
template <typename T> GLM_CFN tvec2<T> operator*(T s, const tvec2<T>& a) { return tvec2<T>(s * a.x, s * a.y); }
template <typename T> GLM_CFN tvec2<T> operator/(const tvec2<T>& a, T s) { return tvec2<T>(a.x / s, a.y / s); }
template <typename T> GLM_CFN tvec2<T> operator-(const tvec2<T>& a) { return tvec2<T>(-a.x, -a.y); }
template <typename T> GLM_CFN bool operator==(const tvec2<T>& a, const tvec2<T>& b) { return a.x == b.x && a.y == b.y; }
template <typename T> GLM_CFN bool operator!=(const tvec2<T>& a, const tvec2<T>& b) { return !(a == b); }

// ---- vec3 operators
template <typename T> GLM_CFN tvec3<T> operator+(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> GLM_CFN tvec3<T> operator-(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> GLM_CFN tvec3<T> operator*(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <typename T> GLM_CFN tvec3<T> operator/(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x / b.x, a.y / b.y, a.z / b.z); }
template <typename T> GLM_CFN tvec3<T> operator%(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(a.x % b.x, a.y % b.y, a.z % b.z); }
template <typename T> GLM_CFN tvec3<T> operator+(const tvec3<T>& a, T s) { return tvec3<T>(a.x + s, a.y + s, a.z + s); }
template <typename T> GLM_CFN tvec3<T> operator+(T s, const tvec3<T>& a) { return tvec3<T>(s + a.x, s + a.y, s + a.z); }
template <typename T> GLM_CFN tvec3<T> operator-(const tvec3<T>& a, T s) { return tvec3<T>(a.x - s, a.y - s, a.z - s); }
template <typename T> GLM_CFN tvec3<T> operator-(T s, const tvec3<T>& a) { return tvec3<T>(s - a.x, s - a.y, s - a.z); }
template <typename T> GLM_CFN tvec3<T> operator*(const tvec3<T>& a, T s) { return tvec3<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> GLM_CFN tvec3<T> operator*(T s, const tvec3<T>& a) { return tvec3<T>(s * a.x, s * a.y, s * a.z); }
template <typename T> GLM_CFN tvec3<T> operator/(const tvec3<T>& a, T s) { return tvec3<T>(a.x / s, a.y / s, a.z / s); }
template <typename T> GLM_CFN tvec3<T> operator/(T s, const tvec3<T>& a) { return tvec3<T>(s / a.x, s / a.y, s / a.z); }
template <typename T> GLM_CFN tvec3<T> operator%(const tvec3<T>& a, T s) { return tvec3<T>(a.x % s, a.y % s, a.z % s); }
template <typename T> GLM_CFN tvec3<T> operator-(const tvec3<T>& a) { return tvec3<T>(-a.x, -a.y, -a.z); }
template <typename T> GLM_CFN bool operator==(const tvec3<T>& a, const tvec3<T>& b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
template <typename T> GLM_CFN bool operator!=(const tvec3<T>& a, const tvec3<T>& b) { return !(a == b); }

// ---- vec4 operators
template <typename T> GLM_CFN tvec4<T> operator+(const tvec4<T>& a, const tvec4<T>& b) { return tvec4<T>(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
template <typename T> GLM_CFN tvec4<T> operator*(const tvec4<T>& a, T s) { return tvec4<T>(a.x * s, a.y * s, a.z * s, a.w * s); }
template <typename T> GLM_CFN tvec4<T> operator/(const tvec4<T>& a, T s) { return tvec4<T>(a.x / s, a.y / s, a.z / s, a.w / s); }

// ---- scalar functions
template <typename T> GLM_CFN T min(T x, T y) { return (y < x) ? y : x; }
template <typename T> GLM_CFN T max(T x, T y) { return (x < y) ? y : x; }
template <typename T> GLM_CFN T clamp(T x, T lo, T hi) { return min(max(x, lo), hi); }
template <typename T> GLM_CFN T sign(T x) { return static_cast<T>(static_cast<T>(0) < x) - static_cast<T>(x < static_cast<T>(0)); }
template <typename T> GLM_CFN T mix(T x, T y, T a) { return x * (static_cast<T>(1) - a) + y * a; }
GLM_FN float inversesqrt(float x) { return 1.0f / sqrtf(x); }
GLM_FN float smoothstep(float e0, float e1, float x) {
	const float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
	return t * t * (3.0f - 2.0f * t);
}
GLM_FN float abs(float x) { return fabsf(x); }
GLM_FN float trunc(float x) { return truncf(x); }
GLM_FN float floor(float x) { return floorf(x); }
GLM_FN float pow(float x, float y) { return powf(x, y); }
GLM_FN float exp(float x) { return expf(x); }

// ---- vector functions
template <typename T> GLM_CFN tvec3<T> min(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
template <typename T> GLM_CFN tvec3<T> max(const tvec3<T>& a, const tvec3<T>& b) { return tvec3<T>(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
template <typename T> GLM_CFN tvec3<T> sign(const tvec3<T>& a) { return tvec3<T>(sign(a.x), sign(a.y), sign(a.z)); }
GLM_FN vec3 abs(const vec3& a) { return vec3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
GLM_FN vec3 trunc(const vec3& a) { return vec3(truncf(a.x), truncf(a.y), truncf(a.z)); }
GLM_FN vec3 floor(const vec3& a) { return vec3(floorf(a.x), floorf(a.y), floorf(a.z)); }
GLM_FN vec3 exp(const vec3& a) { return vec3(expf(a.x), expf(a.y), expf(a.z)); }
GLM_FN vec3 pow(const vec3& a, const vec3& b) { return vec3(powf(a.x, b.x), powf(a.y, b.y), powf(a.z, b.z)); }
GLM_FN vec4 pow(const vec4& a, const vec4& b) { return vec4(powf(a.x, b.x), powf(a.y, b.y), powf(a.z, b.z), powf(a.w, b.w)); }
GLM_CFN float dot(const vec2& a, const vec2& b) { return a.x * b.x + a.y * b.y; }
GLM_CFN float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GLM_CFN vec3 cross(const vec3& a, const vec3& b) {
	return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
GLM_FN vec3 normalize(const vec3& v) { return v * inversesqrt(dot(v, v)); }
GLM_FN vec2 normalize(const vec2& v) { return v * inversesqrt(dot(v, v)); }
GLM_CFN vec3 mix(const vec3& x, const vec3& y, float a) { return x * (1.0f - a) + y * a; }
GLM_CFN vec3 clamp(const vec3& x, float lo, float hi) { return vec3(clamp(x.x, lo, hi), clamp(x.y, lo, hi), clamp(x.z, lo, hi)); }

} // namespace glm
