// ORACLE — TEST INFRASTRUCTURE ONLY.  random.glsl / sampling.glsl / brdf.glsl of the reference compiled as C++
// (see gen.py) behind the same unit hooks the restatement exports (or_rng_*, or_next_floats, or_*_uber), so the
// known-answer tests can put the two side by side function by function.
#include "stage_common.h"
namespace glsl
{
namespace units
{
#include "brdf.glsl.inc"
static SurfaceProperties make_surface(const float* n, float roughness, float metallic, const float* albedo)
{
    SurfaceProperties p;
    p.normal    = vec3(n[0], n[1], n[2]);
    p.albedo    = vec4(albedo[0], albedo[1], albedo[2], 1.0f);
    p.roughness = max(roughness, MIN_ROUGHNESS);                  // rchit:273
    p.metallic  = metallic;
    p.F0        = mix(vec3(0.03f), p.albedo.xyz, p.metallic);     // rchit:275
    p.alpha     = p.roughness * p.roughness;
    p.alpha2    = p.alpha * p.alpha;
    return p;
}
} // namespace units
} // namespace glsl
using namespace glsl::units;
using glsl::vec3;

extern "C" uint32_t ref_rng_hash(uint32_t s) { return rng_hash(s); }
extern "C" void     ref_rng_sequence(uint32_t sx, uint32_t sy, uint32_t n, uint32_t* out_results, uint32_t* out_state)
{
    RNG r;
    r.s = glsl::uvec2(sx, sy);
    for (uint32_t i = 0; i < n; i++)
    {
        out_results[i]       = rng_next(r);
        out_state[2 * i]     = r.s.x;
        out_state[2 * i + 1] = r.s.y;
    }
}
extern "C" void ref_rng_init(uint32_t x, uint32_t y, uint32_t frame, uint32_t* out2)
{
    RNG r   = rng_init(glsl::uvec2(x, y), frame);
    out2[0] = r.s.x, out2[1] = r.s.y;
}
extern "C" void ref_next_floats(uint32_t sx, uint32_t sy, uint32_t n, float* out)
{
    RNG r;
    r.s = glsl::uvec2(sx, sy);
    for (uint32_t i = 0; i < n; i++) out[i] = next_float(r);
}
extern "C" void ref_next_uints(uint32_t sx, uint32_t sy, uint32_t nmax, uint32_t n, uint32_t* out)
{
    RNG r;
    r.s = glsl::uvec2(sx, sy);
    for (uint32_t i = 0; i < n; i++) out[i] = next_uint(r, nmax);
}
// out4 = brdf.rgb, pdf
extern "C" void ref_evaluate_uber(const float* n, const float* wo, const float* wi, float roughness, float metallic, const float* albedo, float* out4)
{
    SurfaceProperties p  = make_surface(n, roughness, metallic, albedo);
    vec3              Wo = vec3(wo[0], wo[1], wo[2]), Wi = vec3(wi[0], wi[1], wi[2]);
    vec3              Wh = normalize(Wo + Wi);
    vec3              f  = evaluate_uber(p, Wo, Wh, Wi);
    out4[0] = f.x, out4[1] = f.y, out4[2] = f.z, out4[3] = pdf_uber(p, Wo, Wh, Wi);
}
// out7 = brdf.rgb, Wi.xyz, pdf
extern "C" void ref_sample_uber(const float* n, const float* wo, float roughness, float metallic, const float* albedo, uint32_t sx, uint32_t sy, float* out7)
{
    SurfaceProperties p  = make_surface(n, roughness, metallic, albedo);
    vec3              Wo = vec3(wo[0], wo[1], wo[2]), Wi;
    float             pdf;
    RNG               r;
    r.s    = glsl::uvec2(sx, sy);
    vec3 f = sample_uber(p, Wo, r, Wi, pdf);
    out7[0] = f.x, out7[1] = f.y, out7[2] = f.z, out7[3] = Wi.x, out7[4] = Wi.y, out7[5] = Wi.z, out7[6] = pdf;
}
extern "C" void ref_sample_cosine_lobe(const float* n, float r0, float r1, float* out3)
{
    vec3 w  = sample_cosine_lobe(vec3(n[0], n[1], n[2]), glsl::vec2(r0, r1));
    out3[0] = w.x, out3[1] = w.y, out3[2] = w.z;
}
