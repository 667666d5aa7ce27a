// ORACLE — TEST INFRASTRUCTURE ONLY.  The reference's tone_map.frag and procedural_sky.frag compiled as C++
// (see gen.py), one fragment per call.
#include "stage_common.h"
namespace glsl
{
namespace tone_map
{
static thread_local const float* t_src; // RGBA32F colour attachment bound as samplerColor
static thread_local int          t_w, t_h;
// the pass draws one full-screen triangle at the attachment's resolution: every fragment samples its own texel
// centre, so the filter mode does not matter and a fetch of texel floor(uv * size) is exact
inline vec4 texture(sampler2D, vec2 uv)
{
    int x = (int)floor(uv.x * (float)t_w), y = (int)floor(uv.y * (float)t_h);
    x = x < 0 ? 0 : (x >= t_w ? t_w - 1 : x), y = y < 0 ? 0 : (y >= t_h ? t_h - 1 : y);
    const float* q = t_src + ((size_t)y * t_w + x) * 4;
    return vec4(q[0], q[1], q[2], q[3]);
}
#define main glsl_main
#include "tone_map.frag.inc"
#undef main
} // namespace tone_map
namespace procedural_sky
{
#define main glsl_main
#include "procedural_sky.frag.inc"
#undef main
} // namespace procedural_sky
} // namespace glsl

// out4 = outFragColor of the fragment whose centre maps to texel (x, y) of src
extern "C" void ref_tone_map_fragment(const float* src, int w, int h, int x, int y, float exposure, uint32_t op, float* out4)
{
    using namespace glsl::tone_map;
    t_src = src, t_w = w, t_h = h;
    u_PushConstants.exposure          = exposure;
    u_PushConstants.tone_map_operator = op;
    inUV                              = glsl::vec2(((float)x + 0.5f) / (float)w, ((float)y + 0.5f) / (float)h);
    glsl_main();
    out4[0] = outFragColor.x, out4[1] = outFragColor.y, out4[2] = outFragColor.z, out4[3] = outFragColor.w;
}
// ubo40 = the HosekWilkieUBO (A..I, Z as vec4); pos = interpolated cube position of the fragment
extern "C" void ref_sky_bind(const float* ubo40, const float* sun_dir3)
{
    using namespace glsl::procedural_sky;
    std::memcpy((void*)&u_PerFrameUBO, ubo40, 160);
    u_PushConstants.direction = glsl::vec3(sun_dir3[0], sun_dir3[1], sun_dir3[2]);
}
extern "C" void ref_sky_fragment(const float* pos3, float* out4)
{
    using namespace glsl::procedural_sky;
    FS_IN_Position = glsl::vec3(pos3[0], pos3[1], pos3[2]);
    glsl_main();
    out4[0] = FS_OUT_Color.x, out4[1] = FS_OUT_Color.y, out4[2] = FS_OUT_Color.z, out4[3] = FS_OUT_Color.w;
}
