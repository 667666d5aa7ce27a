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
inline vec4 operator*(const vec4& a, const vec4& b) { return vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline vec4 operator-(const vec4& a, const vec4& b) { return vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
// 4x4 inverse and determinant with GLM 0.9.9's operation order (detail/func_matrix.inl, compute_inverse<4,4> /
// compute_determinant<4,4>; the reference vendors GLM at b3f8772 under external/AssetCore/external/glm): the path's
// view_proj_inverse (path_integrator.cpp:152) and the parent-relative transforms (scene.cpp:319-321) go through it, and
// the tables must come out bit for bit (tests/test_ref_scene.py runs the same inputs through the reference's GLM)
inline mat4 inverse(const mat4& m)
{
    const float Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3], Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3], Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
    const float Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3], Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3], Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
    const float Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2], Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2], Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
    const float Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3], Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3], Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
    const float Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2], Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2], Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
    const float Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1], Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1], Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
    const vec4  Fac0(Coef00, Coef00, Coef02, Coef03), Fac1(Coef04, Coef04, Coef06, Coef07), Fac2(Coef08, Coef08, Coef10, Coef11);
    const vec4  Fac3(Coef12, Coef12, Coef14, Coef15), Fac4(Coef16, Coef16, Coef18, Coef19), Fac5(Coef20, Coef20, Coef22, Coef23);
    const vec4  Vec0(m[1][0], m[0][0], m[0][0], m[0][0]), Vec1(m[1][1], m[0][1], m[0][1], m[0][1]), Vec2(m[1][2], m[0][2], m[0][2], m[0][2]), Vec3(m[1][3], m[0][3], m[0][3], m[0][3]);
    const vec4  Inv0(Vec1 * Fac0 - Vec2 * Fac1 + Vec3 * Fac2), Inv1(Vec0 * Fac0 - Vec2 * Fac3 + Vec3 * Fac4), Inv2(Vec0 * Fac1 - Vec1 * Fac3 + Vec3 * Fac5), Inv3(Vec0 * Fac2 - Vec1 * Fac4 + Vec2 * Fac5);
    const vec4  SignA(+1.0f, -1.0f, +1.0f, -1.0f), SignB(-1.0f, +1.0f, -1.0f, +1.0f);
    const mat4  Inverse(Inv0 * SignA, Inv1 * SignB, Inv2 * SignA, Inv3 * SignB);
    const vec4  Row0(Inverse[0][0], Inverse[1][0], Inverse[2][0], Inverse[3][0]);
    const vec4  Dot0(m[0] * Row0);
    const float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
    const float OneOverDeterminant = 1.0f / Dot1;
    return mat4(Inverse[0] * OneOverDeterminant, Inverse[1] * OneOverDeterminant, Inverse[2] * OneOverDeterminant, Inverse[3] * OneOverDeterminant);
}
inline float determinant(const mat4& m)
{
    const float SubFactor00 = m[2][2] * m[3][3] - m[3][2] * m[2][3], SubFactor01 = m[2][1] * m[3][3] - m[3][1] * m[2][3], SubFactor02 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
    const float SubFactor03 = m[2][0] * m[3][3] - m[3][0] * m[2][3], SubFactor04 = m[2][0] * m[3][2] - m[3][0] * m[2][2], SubFactor05 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
    const vec4  DetCof(+(m[1][1] * SubFactor00 - m[1][2] * SubFactor01 + m[1][3] * SubFactor02), -(m[1][0] * SubFactor00 - m[1][2] * SubFactor03 + m[1][3] * SubFactor04),
                      +(m[1][0] * SubFactor01 - m[1][1] * SubFactor03 + m[1][3] * SubFactor05), -(m[1][0] * SubFactor02 - m[1][1] * SubFactor04 + m[1][2] * SubFactor05));
    return m[0][0] * DetCof[0] + m[0][1] * DetCof[1] + m[0][2] * DetCof[2] + m[0][3] * DetCof[3];
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
