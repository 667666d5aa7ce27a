// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
//
// Minimal GLSL-style vector math for the CPU restatement of the reference's shaders.
// Operation order is fixed here (and compiled with -ffp-contract=off) because GLSL leaves
// it to the driver; the CUDA path states the same order independently in its own headers:
//   dot(a,b)      = a.x*b.x + a.y*b.y + a.z*b.z            (left to right)
//   length(v)     = sqrt(dot(v,v))
//   normalize(v)  = v * (1 / length(v))                     (one divide, three multiplies)
//   reflect(I,N)  = I - 2*dot(N,I)*N
//   mix(a,b,t)    = a*(1-t) + b*t
//   M*v (mat4)    = c0*v.x + c1*v.y + c2*v.z + c3*v.w       (left to right)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc
{
struct vec2
{
    float x, y;
};
struct vec3
{
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(float a) : x(a), y(a), z(a) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float  operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
struct vec4
{
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    vec3  xyz() const { return vec3(x, y, z); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};

inline vec3 operator+(vec3 a, vec3 b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline vec3 operator-(vec3 a, vec3 b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline vec3 operator*(vec3 a, vec3 b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline vec3 operator/(vec3 a, vec3 b) { return vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline vec3 operator*(vec3 a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
inline vec3 operator*(float s, vec3 a) { return vec3(s * a.x, s * a.y, s * a.z); }
inline vec3 operator/(vec3 a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
inline vec3 operator-(vec3 a) { return vec3(-a.x, -a.y, -a.z); }
inline vec3& operator+=(vec3& a, vec3 b)
{
    a = a + b;
    return a;
}
inline vec3& operator*=(vec3& a, float s)
{
    a = a * s;
    return a;
}
inline vec3& operator/=(vec3& a, float s)
{
    a = a / s;
    return a;
}
inline vec4 operator+(vec4 a, vec4 b) { return vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline vec4 operator*(vec4 a, float s) { return vec4(a.x * s, a.y * s, a.z * s, a.w * s); }
inline vec4 operator/(vec4 a, float s) { return vec4(a.x / s, a.y / s, a.z / s, a.w / s); }

inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3  cross(vec3 a, vec3 b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline vec3  normalize(vec3 a) { return a * (1.0f / length(a)); }
inline vec3  reflect(vec3 I, vec3 N) { return I - N * (2.0f * dot(N, I)); }
inline float clampf(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline vec3  mix(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
inline vec3  vmin(vec3 a, vec3 b) { return vec3(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)); }
inline float smoothstep(float e0, float e1, float x)
{
    float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}

// column-major 4x4 (glm::mat4 / GLSL mat4): c[col][row]
struct mat4
{
    float c[4][4];
};
inline mat4 mat4_from(const float* m)
{
    mat4 r;
    std::memcpy(r.c, m, 64);
    return r;
}
inline vec4 mul(const mat4& m, vec4 v)
{
    vec4 r;
    r.x = m.c[0][0] * v.x + m.c[1][0] * v.y + m.c[2][0] * v.z + m.c[3][0] * v.w;
    r.y = m.c[0][1] * v.x + m.c[1][1] * v.y + m.c[2][1] * v.z + m.c[3][1] * v.w;
    r.z = m.c[0][2] * v.x + m.c[1][2] * v.y + m.c[2][2] * v.z + m.c[3][2] * v.w;
    r.w = m.c[0][3] * v.x + m.c[1][3] * v.y + m.c[2][3] * v.z + m.c[3][3] * v.w;
    return r;
}
// mat3(m) * v
inline vec3 mul3(const mat4& m, vec3 v)
{
    vec3 r;
    r.x = m.c[0][0] * v.x + m.c[1][0] * v.y + m.c[2][0] * v.z;
    r.y = m.c[0][1] * v.x + m.c[1][1] * v.y + m.c[2][1] * v.z;
    r.z = m.c[0][2] * v.x + m.c[1][2] * v.y + m.c[2][2] * v.z;
    return r;
}
// mat3 with columns x,y,z times v
inline vec3 mul_cols(vec3 cx, vec3 cy, vec3 cz, vec3 v)
{
    vec3 r;
    r.x = cx.x * v.x + cy.x * v.y + cz.x * v.z;
    r.y = cx.y * v.x + cy.y * v.y + cz.y * v.z;
    r.z = cx.z * v.x + cy.z * v.y + cz.z * v.z;
    return r;
}

inline uint32_t float_bits(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline float bits_float(uint32_t u)
{
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
} // namespace orc
