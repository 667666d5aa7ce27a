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

#if defined(RAY_DEBUG_VIEW)
// Storage buffers are shared memory: the generated `static DebugRayDrawArgs_t DebugRayDrawArgs;` of each stage becomes the
// declaration of an accessor (`static DebugRayDrawArgs_t (*glsl_debug_draw_args());`) that GLSL_DEBUG_BLOCKS defines right
// after the shader text, and every use (`DebugRayDrawArgs.count`) goes to the driver's one block.
#define DebugRayDrawArgs (*glsl_debug_draw_args())
#define DebugRayVertexBuffer (*glsl_debug_vertex_buffer())
#define GLSL_DEBUG_BLOCKS                                                                                                  \
    static DebugRayDrawArgs_t*     glsl_debug_draw_args() { return (DebugRayDrawArgs_t*)ref_drv_debug_draw_args(); }       \
    static DebugRayVertexBuffer_t* glsl_debug_vertex_buffer() { return (DebugRayVertexBuffer_t*)ref_drv_debug_vertex_block(); } \
    static_assert(sizeof(DebugRayDrawArgs_t) == sizeof(RefDebugDrawArgs) && sizeof(DebugRayVertexBuffer_t) == sizeof(RefDebugVertexBlock) && sizeof(DebugRayVertex) == 32, "set 5 layout");
namespace glsl
{
inline uint atomicAdd(uint& mem, uint data) // single-threaded launches only (ref_gather_debug_rays)
{
    const uint old = mem;
    mem += data;
    return old;
}
} // namespace glsl
#else
#define GLSL_DEBUG_BLOCKS
#endif
