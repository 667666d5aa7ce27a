// hl_tex.h — software texture filtering for the path: textureLod(sampler2D, uv, 0) and
// texture(samplerCube, dir) (path_trace_rchit.glsl:196-251, path_trace_rahit.glsl:162-168,
// path_trace_rmiss.glsl:60, path_trace_rchit.glsl:370).  The reference leaves filtering to the Vulkan
// driver (sampler state gfx/vk.cpp:3557-3587: LINEAR, REPEAT); this implementation fixes: level 0 only,
// fp32 bilinear weights, texel centres at +0.5, 8-bit decode through a host-built LUT, cube faces
// selected with the Vulkan major-axis table and filtered inside the face with clamp-to-edge.
#pragma once
#include "hl_scene.h"

namespace hl
{
HL_HD f4 texel_load(const TexView& t, const float* lut8, int x, int y)
{
    const size_t i = (size_t)y * t.w + (size_t)x;
    if (t.format == HL_TEX_RGBA32F)
    {
        const float* p = (const float*)t.texels + i * 4;
        return mk4(p[0], p[1], p[2], p[3]);
    }
    const uint32_t px  = ((const uint32_t*)t.texels)[i];
    const float*   lut = lut8 + (t.format == HL_TEX_RGBA8_SRGB ? 256 : (t.format == HL_TEX_RGBA8_SNORM ? 512 : 0));
    const float*   la  = t.format == HL_TEX_RGBA8_SRGB ? lut8 : lut; // sRGB alpha is linear
    return mk4(lut[px & 0xFF], lut[(px >> 8) & 0xFF], lut[(px >> 16) & 0xFF], la[px >> 24]);
}
HL_HD f4 bilinear(f4 t00, f4 t10, f4 t01, f4 t11, float fx, float fy)
{
    const f4 a = t00 * (1.0f - fx) + t10 * fx;
    const f4 b = t01 * (1.0f - fx) + t11 * fx;
    return a * (1.0f - fy) + b * fy;
}
// REPEAT addressing of the bilinear footprint.  After u -= floor(u), u is in [0, 1] (1.0 when a tiny negative u rounds
// up), so x = u * W - 0.5 lies in [-0.5, W - 0.5] and floor(x) in [-1, W - 1]: the wrap is one conditional add, not
// an integer modulo (four of those were ~100 instructions per sample).
struct TexFootprint
{
    int   ix0, iy0, ix1, iy1;
    float fx, fy;
};
HL_HD TexFootprint texture_footprint(const TexView& t, float u, float v)
{
    TexFootprint f;
    u -= floorf(u);
    v -= floorf(v);
    const float x = u * (float)t.w - 0.5f, y = v * (float)t.h - 0.5f;
    const float x0 = floorf(x), y0 = floorf(y);
    f.fx = x - x0, f.fy = y - y0;
    const int W = (int)t.w, H = (int)t.h;
    f.ix0 = (int)x0, f.iy0 = (int)y0;
    if (f.ix0 < 0) f.ix0 += W;
    if (f.iy0 < 0) f.iy0 += H;
    f.ix1 = f.ix0 + 1 == W ? 0 : f.ix0 + 1;
    f.iy1 = f.iy0 + 1 == H ? 0 : f.iy0 + 1;
    return f;
}
HL_HD f4 sample_texture_lod0(const SceneView& s, int index, float u, float v)
{
    if (!(fabsf(u) < 1e30f) || !(fabsf(v) < 1e30f)) return mk4(0.0f, 0.0f, 0.0f, 0.0f);
    const TexView      t = s.textures[index];
    const TexFootprint f = texture_footprint(t, u, v);
    return bilinear(texel_load(t, s.lut8, f.ix0, f.iy0), texel_load(t, s.lut8, f.ix1, f.iy0), texel_load(t, s.lut8, f.ix0, f.iy1), texel_load(t, s.lut8, f.ix1, f.iy1), f.fx, f.fy);
}
// .w of sample_texture_lod0, bit for bit (the channels filter independently): what the any-hit shader needs
// (path_trace_rahit.glsl:162-188 reads only albedo.a)
HL_HD float texel_alpha(const TexView& t, const float* lut8, int x, int y)
{
    const size_t i = (size_t)y * t.w + (size_t)x;
    if (t.format == HL_TEX_RGBA32F) return ((const float*)t.texels)[i * 4 + 3];
    const uint32_t px = ((const uint32_t*)t.texels)[i];
    return lut8[(t.format == HL_TEX_RGBA8_SNORM ? 512 : 0) + (px >> 24)]; // sRGB alpha is linear (UNORM table)
}
HL_HD float sample_texture_alpha_lod0(const SceneView& s, int index, float u, float v)
{
    if (!(fabsf(u) < 1e30f) || !(fabsf(v) < 1e30f)) return 0.0f;
    const TexView      t = s.textures[index];
    const TexFootprint f = texture_footprint(t, u, v);
    const float a = texel_alpha(t, s.lut8, f.ix0, f.iy0) * (1.0f - f.fx) + texel_alpha(t, s.lut8, f.ix1, f.iy0) * f.fx;
    const float b = texel_alpha(t, s.lut8, f.ix0, f.iy1) * (1.0f - f.fx) + texel_alpha(t, s.lut8, f.ix1, f.iy1) * f.fx;
    return a * (1.0f - f.fy) + b * f.fy;
}

HL_HD f3 sample_environment(const EnvView& e, f3 r)
{
    if (e.size == 0) return mk3(0.0f);
    const float ax = fabsf(r.x), ay = fabsf(r.y), az = fabsf(r.z);
    int         face;
    float       sc, tc, ma;
    if (az >= ax && az >= ay)
    {
        const bool pos = r.z >= 0.0f;
        face = pos ? 4 : 5, sc = pos ? r.x : -r.x, tc = -r.y, ma = az;
    }
    else if (ay >= ax)
    {
        const bool pos = r.y >= 0.0f;
        face = pos ? 2 : 3, sc = r.x, tc = pos ? r.z : -r.z, ma = ay;
    }
    else
    {
        const bool pos = r.x >= 0.0f;
        face = pos ? 0 : 1, sc = pos ? -r.z : r.z, tc = -r.y, ma = ax;
    }
    if (!(ma > 0.0f) || !(ma < 1e30f)) return mk3(0.0f);
    const float s = 0.5f * (sc / ma + 1.0f), t = 0.5f * (tc / ma + 1.0f);
    if (!(s >= 0.0f && s <= 1.0f && t >= 0.0f && t <= 1.0f)) return mk3(0.0f);
    const int   N = (int)e.size;
    const float x = s * (float)N - 0.5f, y = t * (float)N - 0.5f;
    const float x0 = floorf(x), y0 = floorf(y);
    const float fx = x - x0, fy = y - y0;
    int         ix0 = (int)x0, iy0 = (int)y0;
    int         ix1 = ix0 + 1, iy1 = iy0 + 1;
    ix0 = ix0 < 0 ? 0 : (ix0 > N - 1 ? N - 1 : ix0);
    iy0 = iy0 < 0 ? 0 : (iy0 > N - 1 ? N - 1 : iy0);
    ix1 = ix1 > N - 1 ? N - 1 : ix1;
    iy1 = iy1 > N - 1 ? N - 1 : iy1;
    const f4* base = e.faces + (size_t)face * N * N;
    const f4  c    = bilinear(base[(size_t)iy0 * N + ix0], base[(size_t)iy0 * N + ix1], base[(size_t)iy1 * N + ix0], base[(size_t)iy1 * N + ix1], fx, fy);
    return mk3(c.x, c.y, c.z);
}
} // namespace hl
