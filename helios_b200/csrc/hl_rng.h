// hl_rng.h — per-path random numbers: xoroshiro64* seeded by a Wang hash of (pixel, frame).
// Restates src/engine/shader/random.glsl:11-50 and sampling.glsl:6-36 of the reference.
#pragma once
#include "hl_hd.h"

namespace hl
{
struct Rng
{
    uint32_t x, y;
};
HL_HD uint32_t rotl32(uint32_t v, uint32_t k) { return (v << k) | (v >> (32u - k)); }
HL_HD uint32_t rng_next(Rng& r) // random.glsl:17-26
{
    const uint32_t out = r.x * 0x9e3779bbu;
    r.y ^= r.x;
    r.x = rotl32(r.x, 26) ^ r.y ^ (r.y << 9);
    r.y = rotl32(r.y, 13);
    return out;
}
HL_HD uint32_t wang_hash(uint32_t s) // random.glsl:30-38
{
    s = (s ^ 61u) ^ (s >> 16);
    s *= 9u;
    s ^= s >> 4;
    s *= 0x27d4eb2du;
    s ^= s >> 15;
    return s;
}
HL_HD Rng rng_seed(uint32_t px, uint32_t py, uint32_t frame) // random.glsl:40-50
{
    Rng r;
    r.x = wang_hash((px << 16) | py);
    r.y = wang_hash(frame);
    rng_next(r); // warm-up draw is discarded
    return r;
}
HL_HD float rand01(Rng& r) { return u2f(0x3f800000u | (rng_next(r) >> 9)) - 1.0f; }     // sampling.glsl:6-10
HL_HD uint32_t rand_below(Rng& r, uint32_t n) { return (uint32_t)floorf(rand01(r) * (float)n); } // sampling.glsl:12-16

// orthonormal basis around z (sampling.glsl:28-36): columns x, y, z
HL_HD void basis_around(f3 z, f3& x, f3& y)
{
    const bool steep = fabsf(dot(z, mk3(0.0f, 1.0f, 0.0f))) > 0.99f;
    const f3   ref   = steep ? mk3(0.0f, 0.0f, 1.0f) : mk3(0.0f, 1.0f, 0.0f);
    x                = normalize(cross(ref, z));
    y                = cross(z, x);
}
} // namespace hl
