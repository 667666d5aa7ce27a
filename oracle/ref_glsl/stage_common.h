// ORACLE — TEST INFRASTRUCTURE ONLY.
// stage_common.h — included by every stage_*.cpp before the generated shader text: resource built-ins routed to
// the driver, and the two <cmath> macros the reference's common.glsl redefines.
#pragma once
#include "glsl_compat.h"
#include "ref_abi.h"
#undef M_PI
#undef INFINITY

namespace glsl
{
inline vec4 textureLod(sampler2D s, vec2 uv, float /*lod: the path only ever passes 0*/)
{
    vec4 r;
    ref_drv_texture2d(s.index, uv.x, uv.y, &r.x);
    return r;
}
inline vec4 texture(samplerCube, vec3 d)
{
    vec4 r;
    ref_drv_texture_cube(&d.x, &r.x);
    return r;
}
inline void glsl_trace(uint flags, uint sbt_offset, uint miss_index, vec3 o, float tmin, vec3 d, float tmax, void* payload)
{
    ref_drv_trace(flags, sbt_offset, miss_index, &o.x, tmin, &d.x, tmax, payload);
}
} // namespace glsl
// traceRayEXT(as, flags, cullMask, sbtRecordOffset, sbtRecordStride, missIndex, origin, tmin, dir, tmax, payloadLocation);
// GLSL_PAYLOAD_AT maps the payload LOCATION to the stage's variable declared with that layout(location = N).
#define traceRayEXT(as, flags, mask, sbt_off, sbt_stride, miss, o, tmin, d, tmax, loc) glsl_trace(flags, sbt_off, miss, o, tmin, d, tmax, GLSL_PAYLOAD_AT(loc))
