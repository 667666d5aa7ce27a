// glm.hpp — minimal stand-in for the subset of GLM the engine's path-tracing interfaces use
// (include/resource/scene.h, include/gfx/path_integrator.h, src/engine/resource/scene.cpp:186-360,632-646,
// src/engine/gfx/path_integrator.cpp:136-161 in the reference).  The reference vendors the real GLM under
// external/AssetCore; this image has no copy that travels to the GPU box, so the shim ships the few types and
// functions it needs, written from GLM's documented conventions: column-major matrices (m[col][row]),
// right-handed, OpenGL clip space (-1..1) for perspective(), quaternions stored x,y,z,w and constructed
// (w, x, y, z).  When the real GLM is first on the include path it is used instead of this file.
#pragma once
#include <cmath>
#include <cstdint>

namespace glm
{
template <typename T>
struct tvec2
{
    T x, y;
    tvec2() : x(0), y(0) {}
    explicit tvec2(T s) : x(s), y(s) {}
    tvec2(T a, T b) : x(a), y(b) {}
    template <typename U>
    tvec2(const tvec2<U>& o) : x(T(o.x)), y(T(o.y)) {}
    T&       operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
};
template <typename T>
struct tvec3
{
    T x, y, z;
    tvec3() : x(0), y(0), z(0) {}
    explicit tvec3(T s) : x(s), y(s), z(s) {}
    tvec3(T a, T b, T c) : x(a), y(b), z(c) {}
    T&       operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    tvec3&   operator+=(const tvec3& o)
    {
        x += o.x, y += o.y, z += o.z;
        return *this;
    }
};
template <typename T>
struct tvec4
{
    union
    {
        struct
        {
            T x, y, z, w;
        };
        struct
        {
            T r, g, b, a;
        };
    };
    tvec4() : x(0), y(0), z(0), w(0) {}
    explicit tvec4(T s) : x(s), y(s), z(s), w(s) {}
    tvec4(T a_, T b_, T c_, T d_) : x(a_), y(b_), z(c_), w(d_) {}
    tvec4(const tvec3<T>& v, T d_) : x(v.x), y(v.y), z(v.z), w(d_) {}
    tvec4(T a_, const tvec3<T>& v) : x(a_), y(v.x), z(v.y), w(v.z) {}
    template <typename U>
    tvec4(const tvec4<U>& o) : x(T(o.x)), y(T(o.y)), z(T(o.z)), w(T(o.w)) {}
    T&       operator[](int i) { return (&x)[i]; }
    const T& operator[](int i) const { return (&x)[i]; }
    operator tvec3<T>() const { return tvec3<T>(x, y, z); } // scene.cpp:240 returns mat4 * vec4 as a vec3
};
using vec2  = tvec2<float>;
using vec3  = tvec3<float>;
using vec4  = tvec4<float>;
using ivec2 = tvec2<int32_t>;
using ivec4 = tvec4<int32_t>;
using uvec2 = tvec2<uint32_t>;
using uvec4 = tvec4<uint32_t>;

inline vec3  operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3  operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3  operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3  operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3  operator*(float s, const vec3& a) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3  operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3  operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec4  operator+(const vec4& a, const vec4& b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4  operator*(const vec4& a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3  cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float length(const vec3& a) { return std::sqrt(dot(a, a)); }
inline vec3  normalize(const vec3& a) { return a * (1.0f / length(a)); }
inline vec3  pow(const vec3& a, const vec3& e) { return vec3(std::pow(a.x, e.x), std::pow(a.y, e.y), std::pow(a.z, e.z)); }
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }
inline float degrees(float rad) { return rad * 57.295779513082320876798154814105f; }

struct mat4
{
    vec4 c[4]; // columns
    mat4() : mat4(1.0f) {}
    explicit mat4(float d)
    {
        c[0] = vec4(d, 0, 0, 0), c[1] = vec4(0, d, 0, 0), c[2] = vec4(0, 0, d, 0), c[3] = vec4(0, 0, 0, d);
    }
    mat4(const vec4& a, const vec4& b, const vec4& cc, const vec4& d)
    {
        c[0] = a, c[1] = b, c[2] = cc, c[3] = d;
    }
    vec4&       operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline vec4 operator*(const mat4& m, const vec4& v)
{
    // GLM order: (col0*x + col1*y) + (col2*z + col3*w)
    vec4 r;
    for (int i = 0; i < 4; i++) r[i] = (m[0][i] * v.x + m[1][i] * v.y) + (m[2][i] * v.z + m[3][i] * v.w);
    return r;
}
inline mat4 operator*(const mat4& a, const mat4& b)
{
    mat4 r(0.0f);
    for (int j = 0; j < 4; j++)
        for (int i = 0; i < 4; i++) r[j][i] = a[0][i] * b[j][0] + a[1][i] * b[j][1] + a[2][i] * b[j][2] + a[3][i] * b[j][3];
    return r;
}
inline mat4 transpose(const mat4& m)
{
    mat4 r(0.0f);
    for (int j = 0; j < 4; j++)
        for (int i = 0; i < 4; i++) r[j][i] = m[i][j];
    return r;
}
inline mat4 translate(const mat4& m, const vec3& v)
{
    mat4 r = m;
    r[3]   = m[0] * v.x + m[1] * v.y + m[2] * v.z + m[3];
    return r;
}
inline mat4 scale(const mat4& m, const vec3& v)
{
    mat4 r = m;
    r[0] = m[0] * v.x, r[1] = m[1] * v.y, r[2] = m[2] * v.z;
    return r;
}
// general 4x4 inverse by cofactors (2x2 sub-determinants of the lower two rows, then of the upper two)
inline mat4 inverse(const mat4& m)
{
    const float a00 = m[0][0], a01 = m[0][1], a02 = m[0][2], a03 = m[0][3];
    const float a10 = m[1][0], a11 = m[1][1], a12 = m[1][2], a13 = m[1][3];
    const float a20 = m[2][0], a21 = m[2][1], a22 = m[2][2], a23 = m[2][3];
    const float a30 = m[3][0], a31 = m[3][1], a32 = m[3][2], a33 = m[3][3];
    const float b00 = a00 * a11 - a01 * a10, b01 = a00 * a12 - a02 * a10, b02 = a00 * a13 - a03 * a10;
    const float b03 = a01 * a12 - a02 * a11, b04 = a01 * a13 - a03 * a11, b05 = a02 * a13 - a03 * a12;
    const float b06 = a20 * a31 - a21 * a30, b07 = a20 * a32 - a22 * a30, b08 = a20 * a33 - a23 * a30;
    const float b09 = a21 * a32 - a22 * a31, b10 = a21 * a33 - a23 * a31, b11 = a22 * a33 - a23 * a32;
    const float det = b00 * b11 - b01 * b10 + b02 * b09 + b03 * b08 - b04 * b07 + b05 * b06;
    const float id  = 1.0f / det;
    mat4        r(0.0f);
    r[0][0] = (a11 * b11 - a12 * b10 + a13 * b09) * id;
    r[0][1] = (a02 * b10 - a01 * b11 - a03 * b09) * id;
    r[0][2] = (a31 * b05 - a32 * b04 + a33 * b03) * id;
    r[0][3] = (a22 * b04 - a21 * b05 - a23 * b03) * id;
    r[1][0] = (a12 * b08 - a10 * b11 - a13 * b07) * id;
    r[1][1] = (a00 * b11 - a02 * b08 + a03 * b07) * id;
    r[1][2] = (a32 * b02 - a30 * b05 - a33 * b01) * id;
    r[1][3] = (a20 * b05 - a22 * b02 + a23 * b01) * id;
    r[2][0] = (a10 * b10 - a11 * b08 + a13 * b06) * id;
    r[2][1] = (a01 * b08 - a00 * b10 - a03 * b06) * id;
    r[2][2] = (a30 * b04 - a31 * b02 + a33 * b00) * id;
    r[2][3] = (a21 * b02 - a20 * b04 - a23 * b00) * id;
    r[3][0] = (a11 * b07 - a10 * b09 - a12 * b06) * id;
    r[3][1] = (a00 * b09 - a01 * b07 + a02 * b06) * id;
    r[3][2] = (a31 * b01 - a30 * b03 - a32 * b00) * id;
    r[3][3] = (a20 * b03 - a21 * b01 + a22 * b00) * id;
    return r;
}
// right-handed, clip z in [-1, 1] (no GLM_FORCE_* is defined by the reference's build)
inline mat4 perspective(float fovy, float aspect, float z_near, float z_far)
{
    const float t = std::tan(fovy / 2.0f);
    mat4        r(0.0f);
    r[0][0] = 1.0f / (aspect * t);
    r[1][1] = 1.0f / t;
    r[2][2] = -(z_far + z_near) / (z_far - z_near);
    r[2][3] = -1.0f;
    r[3][2] = -(2.0f * z_far * z_near) / (z_far - z_near);
    return r;
}
template <typename M>
inline const float* value_ptr(const M& m)
{
    return &m[0][0];
}
} // namespace glm
