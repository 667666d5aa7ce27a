// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
//
// glsl_compat.h — just enough of the GLSL 4.60 type system and built-in library, written as C++17, for g++ to
// compile the REFERENCE'S OWN SHADER FILES (src/engine/shader/*.glsl, *.frag — read from /root/reference at build
// time by gen.py, never copied into the repository) as ordinary functions.  This is the "reference run here" that
// pins the CPU restatement in ../helios_oracle.cpp: same inputs through both, outputs must be bit-identical.
//
// What GLSL leaves to the implementation is fixed here exactly as in ../glsl_math.h (the restatement's maths),
// so that any difference between the two paths is a difference in the shader LOGIC, never in a library routine:
//   dot(a,b)      = a.x*b.x + a.y*b.y + a.z*b.z  (left to right)      length(v) = sqrt(dot(v,v))
//   normalize(v)  = v * (1 / length(v))                                reflect(I,N) = I - N*(2*dot(N,I))
//   mix(a,b,t)    = a*(1-t) + b*t                                      clamp(x,a,b) = fmin(fmax(x,a),b)
//   M*v           = c0*v.x + c1*v.y + ... (left to right)              min/max      = fmin/fmax (NaN-ignoring)
//   pow/exp/sin/cos/acos/sqrt/floor = the C library's float overloads; compiled with -ffp-contract=off.
// Swizzles are union members (proxy objects), which is what lets `v.normal.xyz = m * v.normal.xyz;` compile as is.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace glsl
{
typedef unsigned int uint;

template <class V, class T, int N, int A, int B>
struct Swz2
{
    T d[N];
    operator V() const { return V(d[A], d[B]); }
    Swz2& operator=(const V& v)
    {
        d[A] = v.x, d[B] = v.y;
        return *this;
    }
    Swz2& operator=(const Swz2& o) { return *this = (V)o; }
};
template <class V, class T, int N, int A, int B, int C>
struct Swz3
{
    T d[N];
    operator V() const { return V(d[A], d[B], d[C]); }
    Swz3& operator=(const V& v)
    {
        d[A] = v.x, d[B] = v.y, d[C] = v.z;
        return *this;
    }
    Swz3& operator=(const Swz3& o) { return *this = (V)o; }
};

// ---- float vectors ------------------------------------------------------------------------------------------
struct uvec2;
struct ivec2;
struct vec2
{
    union
    {
        struct
        {
            float x, y;
        };
        struct
        {
            float r, g;
        };
        Swz2<vec2, float, 2, 0, 1> xy;
    };
    vec2() : x(0), y(0) {}
    vec2(const vec2& o) : x(o.x), y(o.y) {}
    vec2& operator=(const vec2& o)
    {
        x = o.x, y = o.y;
        return *this;
    }
    explicit vec2(float a) : x(a), y(a) {}
    vec2(float a, float b) : x(a), y(b) {}
    explicit vec2(const uvec2& u);
    explicit vec2(const ivec2& u);
};
struct vec3
{
    union
    {
        struct
        {
            float x, y, z;
        };
        struct
        {
            float r, g, b;
        };
        Swz2<vec2, float, 3, 0, 1>    xy;
        Swz2<vec2, float, 3, 1, 2>    yz;
        Swz3<vec3, float, 3, 0, 1, 2> xyz;
        Swz3<vec3, float, 3, 0, 1, 2> rgb;
    };
    vec3() : x(0), y(0), z(0) {}
    vec3(const vec3& o) : x(o.x), y(o.y), z(o.z) {}
    vec3& operator=(const vec3& o)
    {
        x = o.x, y = o.y, z = o.z;
        return *this;
    }
    explicit vec3(float a) : x(a), y(a), z(a) {}
    explicit vec3(double a) : x((float)a), y((float)a), z((float)a) {}
    explicit vec3(int a) : x((float)a), y((float)a), z((float)a) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    vec3(int a, int b, int c) : x((float)a), y((float)b), z((float)c) {}
    float  operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
struct alignas(16) vec4
{
    union
    {
        struct
        {
            float x, y, z, w;
        };
        struct
        {
            float r, g, b, a;
        };
        Swz2<vec2, float, 4, 0, 1>    xy;
        Swz2<vec2, float, 4, 2, 3>    zw;
        Swz3<vec3, float, 4, 0, 1, 2> xyz;
        Swz3<vec3, float, 4, 0, 1, 2> rgb;
        Swz3<vec3, float, 4, 1, 2, 3> yzw;
    };
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(const vec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
    vec4& operator=(const vec4& o)
    {
        x = o.x, y = o.y, z = o.z, w = o.w;
        return *this;
    }
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(const vec3& v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    vec4(const vec2& v, float c, float d) : x(v.x), y(v.y), z(c), w(d) {}
    float  operator[](int i) const { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : (i == 2 ? z : w)); }
};

// ---- integer vectors (only what the shaders touch) ----------------------------------------------------------
struct uvec2
{
    union
    {
        struct
        {
            uint x, y;
        };
        Swz2<uvec2, uint, 2, 0, 1> xy;
    };
    uvec2() : x(0), y(0) {}
    uvec2(const uvec2& o) : x(o.x), y(o.y) {}
    uvec2& operator=(const uvec2& o)
    {
        x = o.x, y = o.y;
        return *this;
    }
    uvec2(uint a, uint b) : x(a), y(b) {}
};
struct uvec3
{
    uint x, y, z;
    uvec3() : x(0), y(0), z(0) {}
    uvec3(uint a, uint b, uint c) : x(a), y(b), z(c) {}
};
struct alignas(16) uvec4
{
    union
    {
        struct
        {
            uint x, y, z, w;
        };
        Swz2<uvec2, uint, 4, 0, 1> xy;
        Swz2<uvec2, uint, 4, 2, 3> zw;
    };
    uvec4() : x(0), y(0), z(0), w(0) {}
    uvec4(const uvec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
    uvec4& operator=(const uvec4& o)
    {
        x = o.x, y = o.y, z = o.z, w = o.w;
        return *this;
    }
};
struct ivec2
{
    int x, y;
    ivec2() : x(0), y(0) {}
    ivec2(int a, int b) : x(a), y(b) {}
    explicit ivec2(const uvec2& u) : x((int)u.x), y((int)u.y) {}
};
struct alignas(16) ivec4
{
    union
    {
        struct
        {
            int x, y, z, w;
        };
        Swz2<ivec2, int, 4, 0, 1> xy;
        Swz2<ivec2, int, 4, 2, 3> zw;
    };
    ivec4() : x(0), y(0), z(0), w(0) {}
    ivec4(const ivec4& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}
    ivec4& operator=(const ivec4& o)
    {
        x = o.x, y = o.y, z = o.z, w = o.w;
        return *this;
    }
};
inline vec2::vec2(const uvec2& u) : x((float)u.x), y((float)u.y) {}
inline vec2::vec2(const ivec2& u) : x((float)u.x), y((float)u.y) {}

// ---- arithmetic ---------------------------------------------------------------------------------------------
#define GLSL_VEC_OPS(V, EXPR_VV, EXPR_VS, EXPR_SV)                                      \
    inline V operator+(const V& a, const V& b) { return EXPR_VV(+); }                   \
    inline V operator-(const V& a, const V& b) { return EXPR_VV(-); }                   \
    inline V operator*(const V& a, const V& b) { return EXPR_VV(*); }                   \
    inline V operator/(const V& a, const V& b) { return EXPR_VV(/); }                   \
    inline V operator+(const V& a, float s) { return EXPR_VS(+); }                      \
    inline V operator-(const V& a, float s) { return EXPR_VS(-); }                      \
    inline V operator*(const V& a, float s) { return EXPR_VS(*); }                      \
    inline V operator/(const V& a, float s) { return EXPR_VS(/); }                      \
    inline V operator+(float s, const V& a) { return EXPR_SV(+); }                      \
    inline V operator-(float s, const V& a) { return EXPR_SV(-); }                      \
    inline V operator*(float s, const V& a) { return EXPR_SV(*); }                      \
    inline V operator/(float s, const V& a) { return EXPR_SV(/); }                      \
    inline V& operator+=(V& a, const V& b) { return a = a + b; }                        \
    inline V& operator-=(V& a, const V& b) { return a = a - b; }                        \
    inline V& operator*=(V& a, const V& b) { return a = a * b; }                        \
    inline V& operator/=(V& a, const V& b) { return a = a / b; }                        \
    inline V& operator*=(V& a, float s) { return a = a * s; }                           \
    inline V& operator/=(V& a, float s) { return a = a / s; }

#define V2_VV(op) vec2(a.x op b.x, a.y op b.y)
#define V2_VS(op) vec2(a.x op s, a.y op s)
#define V2_SV(op) vec2(s op a.x, s op a.y)
#define V3_VV(op) vec3(a.x op b.x, a.y op b.y, a.z op b.z)
#define V3_VS(op) vec3(a.x op s, a.y op s, a.z op s)
#define V3_SV(op) vec3(s op a.x, s op a.y, s op a.z)
#define V4_VV(op) vec4(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w)
#define V4_VS(op) vec4(a.x op s, a.y op s, a.z op s, a.w op s)
#define V4_SV(op) vec4(s op a.x, s op a.y, s op a.z, s op a.w)
GLSL_VEC_OPS(vec2, V2_VV, V2_VS, V2_SV)
GLSL_VEC_OPS(vec3, V3_VV, V3_VS, V3_SV)
GLSL_VEC_OPS(vec4, V4_VV, V4_VS, V4_SV)
inline vec2 operator-(const vec2& a) { return vec2(-a.x, -a.y); }
inline vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
inline vec4 operator-(const vec4& a) { return vec4(-a.x, -a.y, -a.z, -a.w); }

// ---- matrices (column major, m[col][row]) --------------------------------------------------------------------
struct alignas(16) mat4
{
    float c[4][4];
};
struct mat3
{
    vec3 c[3];
    mat3() {}
    mat3(const vec3& x, const vec3& y, const vec3& z)
    {
        c[0] = x, c[1] = y, c[2] = z;
    }
    explicit mat3(const mat4& m)
    {
        for (int k = 0; k < 3; k++) c[k] = vec3(m.c[k][0], m.c[k][1], m.c[k][2]);
    }
};
inline vec4 operator*(const mat4& m, const vec4& v)
{
    vec4 r;
    r.x = m.c[0][0] * v.x + m.c[1][0] * v.y + m.c[2][0] * v.z + m.c[3][0] * v.w;
    r.y = m.c[0][1] * v.x + m.c[1][1] * v.y + m.c[2][1] * v.z + m.c[3][1] * v.w;
    r.z = m.c[0][2] * v.x + m.c[1][2] * v.y + m.c[2][2] * v.z + m.c[3][2] * v.w;
    r.w = m.c[0][3] * v.x + m.c[1][3] * v.y + m.c[2][3] * v.z + m.c[3][3] * v.w;
    return r;
}
inline vec3 operator*(const mat3& m, const vec3& v)
{
    vec3 r;
    r.x = m.c[0].x * v.x + m.c[1].x * v.y + m.c[2].x * v.z;
    r.y = m.c[0].y * v.x + m.c[1].y * v.y + m.c[2].y * v.z;
    r.z = m.c[0].z * v.x + m.c[1].z * v.y + m.c[2].z * v.z;
    return r;
}

// ---- built-in functions --------------------------------------------------------------------------------------
inline float sqrt(float x) { return std::sqrt(x); }
inline float sin(float x) { return std::sin(x); }
inline float cos(float x) { return std::cos(x); }
inline float acos(float x) { return std::acos(x); }
inline float exp(float x) { return std::exp(x); }
inline float pow(float x, float y) { return std::pow(x, y); }
inline float floor(float x) { return std::floor(x); }
inline float abs(float x) { return std::fabs(x); }
inline bool  isnan(float x) { return x != x; }
inline float min(float a, float b) { return std::fmin(a, b); }
inline float max(float a, float b) { return std::fmax(a, b); }
inline float clamp(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float smoothstep(float e0, float e1, float x)
{
    float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline float uintBitsToFloat(uint u)
{
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
inline vec2  min(const vec2& a, const vec2& b) { return vec2(min(a.x, b.x), min(a.y, b.y)); }
inline vec2  max(const vec2& a, const vec2& b) { return vec2(max(a.x, b.x), max(a.y, b.y)); }
inline vec3  min(const vec3& a, const vec3& b) { return vec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline vec3  max(const vec3& a, const vec3& b) { return vec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline vec3  clamp(const vec3& v, float lo, float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
inline vec3  mix(const vec3& a, const vec3& b, float t) { return a * (1.0f - t) + b * t; }
inline vec3  pow(const vec3& a, const vec3& e) { return vec3(pow(a.x, e.x), pow(a.y, e.y), pow(a.z, e.z)); }
inline vec3  exp(const vec3& a) { return vec3(exp(a.x), exp(a.y), exp(a.z)); }
inline float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3  cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float length(const vec3& a) { return sqrt(dot(a, a)); }
inline vec3  normalize(const vec3& a) { return a * (1.0f / length(a)); }
inline vec3  reflect(const vec3& I, const vec3& N) { return I - N * (2.0f * dot(N, I)); }

// ---- opaque resource types; the stage wrappers route them to the "driver" (ref_abi.h) -------------------------
struct accelerationStructureEXT
{
};
struct samplerCube
{
};
struct sampler2D
{
    int index;
};
struct sampler2D_array
{
    sampler2D operator[](int i) const { return sampler2D { i }; }
    sampler2D operator[](uint i) const { return sampler2D { (int)i }; }
};
struct image2D
{
    float* data;
    int    width, height;
};
template <class T>
inline T nonuniformEXT(T v)
{
    return v;
}
static const uint gl_RayFlagsOpaqueEXT               = 1u;
static const uint gl_RayFlagsTerminateOnFirstHitEXT = 4u;
inline vec4 imageLoad(const image2D& im, const ivec2& p)
{
    const float* q = im.data + ((size_t)p.y * im.width + p.x) * 4;
    return vec4(q[0], q[1], q[2], q[3]);
}
inline void imageStore(const image2D& im, const ivec2& p, const vec4& v)
{
    float* q = im.data + ((size_t)p.y * im.width + p.x) * 4;
    q[0] = v.x, q[1] = v.y, q[2] = v.z, q[3] = v.w;
}
} // namespace glsl
