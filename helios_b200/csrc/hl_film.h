// hl_film.h — the resolve stage: per-sample clamp + progressive blend (path_trace_rgen.glsl:217-248) fused with
// exposure, ACES / Reinhard tone mapping and gamma (tone_map.frag:20-51), and the Hosek-Wilkie sky texel
// (procedural_sky.frag:48-75).
#pragma once
#include "hl_hd.h"

namespace hl
{
// rgen:219-247: clamp the sample to RADIANCE_CLAMP_COLOR = 1, then running mean with 1/num_frames
// (frame 0 is stored as-is and is overwritten by frame 1: SURVEY A.8-1)
HL_HD f3 accumulate_running_mean(f3 L, f3 prev, uint32_t num_frames)
{
    const f3 c = mk3(fminf(L.x, 1.0f), fminf(L.y, 1.0f), fminf(L.z, 1.0f));
    if (num_frames == 0) return c;
    return prev + (c - prev) / (float)num_frames;
}
HL_HD f3 accumulate_sum(f3 L, f3 prev)
{
    return prev + mk3(fminf(L.x, 1.0f), fminf(L.y, 1.0f), fminf(L.z, 1.0f));
}
HL_HD float aces_curve(float x) // tone_map.frag:20-28
{
    const float a = 2.51f, b = 0.03f, c = 2.43f, d = 0.59f, e = 0.14f;
    return clampf((x * (a * x + b)) / (x * (c * x + d) + e), 0.0f, 1.0f);
}
// tone_map.frag:35-51 + UNORM8 store; returns one 8-bit channel
HL_HD uint32_t tone_map_channel(float c, float exposure, int op)
{
    c *= exposure;
    if (op == 0)
        c = aces_curve(c);
    else if (op == 1)
        c = c / (1.0f + c);
    c = powf(c, 1.0f / 2.2f);
    if (c != c) return 0u;
    return (uint32_t)(int)(clampf(c, 0.0f, 1.0f) * 255.0f + 0.5f);
}
HL_HD uint32_t tone_map_rgba8(f3 c, float exposure, int op)
{
    return tone_map_channel(c.x, exposure, op) | (tone_map_channel(c.y, exposure, op) << 8) | (tone_map_channel(c.z, exposure, op) << 16) | 0xFF000000u;
}

// direction of a cube-map texel centre (Vulkan face order +X,-X,+Y,-Y,+Z,-Z; s,t in texel units)
HL_HD f3 cube_texel_direction(int face, uint32_t i, uint32_t j, uint32_t size)
{
    const float sc = 2.0f * (((float)i + 0.5f) / (float)size) - 1.0f;
    const float tc = 2.0f * (((float)j + 0.5f) / (float)size) - 1.0f;
    f3          p;
    switch (face)
    {
        case 0: p = mk3(1.0f, -tc, -sc); break;
        case 1: p = mk3(-1.0f, -tc, sc); break;
        case 2: p = mk3(sc, 1.0f, tc); break;
        case 3: p = mk3(sc, -1.0f, -tc); break;
        case 4: p = mk3(sc, -tc, 1.0f); break;
        default: p = mk3(-sc, -tc, -1.0f); break;
    }
    return normalize(p);
}
HL_HD f3 pow3(f3 a, float e) { return mk3(powf(a.x, e), powf(a.y, e), powf(a.z, e)); }
HL_HD f3 exp3(f3 a) { return mk3(expf(a.x), expf(a.y), expf(a.z)); }
// procedural_sky.frag:48-64; cf = A,B,C,D,E,F,G,H,I,Z (vec4 each)
HL_HD f3 hosek_wilkie_radiance(const float* cf, f3 v, f3 sun)
{
    const f3    A = mk3(cf + 0), B = mk3(cf + 4), C = mk3(cf + 8), D = mk3(cf + 12), E = mk3(cf + 16);
    const f3    F = mk3(cf + 20), G = mk3(cf + 24), H = mk3(cf + 28), I = mk3(cf + 32), Z = mk3(cf + 36);
    const float cos_theta = clampf(v.y, 0.0f, 1.0f);
    const float cos_gamma = clampf(dot(v, sun), 0.0f, 1.0f);
    const float gamma     = acosf(cos_gamma);
    const f3    chi = mk3(1.0f + cos_gamma * cos_gamma) / pow3(mk3(1.0f) + H * H - 2.0f * cos_gamma * H, 1.5f);
    const f3    r   = (mk3(1.0f) + A * exp3(B / (cos_theta + 0.01f))) * (C + D * exp3(E * gamma) + F * (cos_gamma * cos_gamma) + G * chi + I * sqrtf(cos_theta));
    return Z * r;
}
} // namespace hl
