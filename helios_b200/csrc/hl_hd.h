// hl_hd.h — host/device portability layer and fp32 vector math for the device-logic headers.
//
// The device logic of every stage lives in HL_HD inline functions so the SAME source is compiled
//   (a) by nvcc for sm_100a (--fmad=false: every a*b+c below is two IEEE roundings; explicit
//       hl_fma() marks the places where a fused multiply-add is wanted), and
//   (b) by g++ (-ffp-contract=off) inside tests/emul, a kernel-logic emulator used to debug the
//       stages in a container that has no GPU.  The emulator is a debugging aid under tests/; the
//       product library contains no CPU execution path.
//
// GLSL semantics restated (the reference's shaders leave the evaluation order to the driver; this is
// the order this implementation fixes — the CPU oracle states the same order independently):
//   dot      = a.x*b.x + a.y*b.y + a.z*b.z      normalize(v) = v * (1 / sqrt(dot(v,v)))
//   reflect  = I - N * (2*dot(N,I))             mix(a,b,t)   = a*(1-t) + b*t
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define HL_HD __host__ __device__ __forceinline__
#define HL_DEVICE_CODE 1
#else
#define HL_HD inline
#endif

namespace hl
{
struct f2
{
    float x, y;
};
struct f3
{
    float x, y, z;
};
struct f4
{
    float x, y, z, w;
};
struct alignas(8) u2 // (8-byte aligned: one 64-bit shared / local memory access per traversal stack entry)
{
    uint32_t x, y;
};

HL_HD f3 mk3(float x, float y, float z)
{
    f3 r;
    r.x = x, r.y = y, r.z = z;
    return r;
}
HL_HD f3 mk3(float s) { return mk3(s, s, s); }
HL_HD f3 mk3(const float* p) { return mk3(p[0], p[1], p[2]); }
HL_HD f4 mk4(float x, float y, float z, float w)
{
    f4 r;
    r.x = x, r.y = y, r.z = z, r.w = w;
    return r;
}
HL_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
HL_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
HL_HD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
HL_HD f3 operator/(f3 a, f3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
HL_HD f3 operator*(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
HL_HD f3 operator*(float s, f3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
HL_HD f3 operator/(f3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
HL_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
HL_HD f4 operator+(f4 a, f4 b) { return mk4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
HL_HD f4 operator*(f4 a, float s) { return mk4(a.x * s, a.y * s, a.z * s, a.w * s); }

HL_HD float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
HL_HD f3    cross(f3 a, f3 b) { return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
HL_HD float length(f3 a) { return sqrtf(dot(a, a)); }
HL_HD f3    normalize(f3 a) { return a * (1.0f / length(a)); }
HL_HD f3    reflect(f3 I, f3 N) { return I - N * (2.0f * dot(N, I)); }
HL_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
HL_HD float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
HL_HD f3    mix3(f3 a, f3 b, float t) { return a * (1.0f - t) + b * t; }
HL_HD float smoothstepf(float e0, float e1, float x)
{
    float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
HL_HD float max3f(f3 a) { return fmaxf(a.x, fmaxf(a.y, a.z)); }
HL_HD bool  is_black(f3 c) { return c.x == 0.0f && c.y == 0.0f && c.z == 0.0f; }

// column-major mat4 (m[col*4+row]) times vec4 / mat3(m) times vec3, summed left to right
HL_HD f4 mat4_mul(const float* m, f4 v)
{
    f4 r;
    r.x = m[0] * v.x + m[4] * v.y + m[8] * v.z + m[12] * v.w;
    r.y = m[1] * v.x + m[5] * v.y + m[9] * v.z + m[13] * v.w;
    r.z = m[2] * v.x + m[6] * v.y + m[10] * v.z + m[14] * v.w;
    r.w = m[3] * v.x + m[7] * v.y + m[11] * v.z + m[15] * v.w;
    return r;
}
HL_HD f3 mat4_mul_point_xyz(const float* m, f3 v, float w)
{
    f3 r;
    r.x = m[0] * v.x + m[4] * v.y + m[8] * v.z + m[12] * w;
    r.y = m[1] * v.x + m[5] * v.y + m[9] * v.z + m[13] * w;
    r.z = m[2] * v.x + m[6] * v.y + m[10] * v.z + m[14] * w;
    return r;
}
HL_HD f3 mat3_mul(const float* m, f3 v)
{
    f3 r;
    r.x = m[0] * v.x + m[4] * v.y + m[8] * v.z;
    r.y = m[1] * v.x + m[5] * v.y + m[9] * v.z;
    r.z = m[2] * v.x + m[6] * v.y + m[10] * v.z;
    return r;
}
// mat3(cx, cy, cz) * v
HL_HD f3 cols_mul(f3 cx, f3 cy, f3 cz, f3 v)
{
    f3 r;
    r.x = cx.x * v.x + cy.x * v.y + cz.x * v.z;
    r.y = cx.y * v.x + cy.y * v.y + cz.y * v.z;
    r.z = cx.z * v.x + cy.z * v.y + cz.z * v.z;
    return r;
}

// ---- bit helpers -------------------------------------------------------------------------------
HL_HD uint32_t f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
HL_HD float u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
HL_HD float hl_fma(float a, float b, float c) { return fmaf(a, b, c); }
HL_HD int   hl_popc(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}
// index of the highest set bit (x != 0)
HL_HD int hl_bfind(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)x);
#else
    return 31 - __builtin_clz(x);
#endif
}
HL_HD int hl_clz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __clzll((long long)x);
#else
    return x ? __builtin_clzll(x) : 64;
#endif
}
HL_HD uint32_t hl_byte(uint32_t v, int i) { return (v >> (8 * i)) & 0xFFu; }
// PRMT: result byte k = byte number (sel >> 4k) & 7 of the 8-byte pool {a.b0..a.b3, b.b0..b.b3}
// (selector nibbles are always < 8 here, so the sign-replication mode of the instruction is never used)
HL_HD uint32_t hl_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t pool = ((uint64_t)b << 32) | a;
    uint32_t       r    = 0;
    for (int k = 0; k < 4; k++) r |= (uint32_t)((pool >> (8 * ((sel >> (4 * k)) & 7u))) & 0xFFu) << (8 * k);
    return r;
#endif
}
HL_HD float    hl_inf() { return u2f(0x7f800000u); }
} // namespace hl
