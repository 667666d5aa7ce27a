// ORACLE — TEST INFRASTRUCTURE ONLY.  The reference's path_trace_rgen.glsl compiled as C++ (see gen.py).
#include "stage_common.h"
namespace glsl
{
namespace rgen
{
static thread_local uvec3 gl_LaunchIDEXT;
#define GLSL_PAYLOAD_AT(loc) ((void*)&p_PathTracePayload) /* layout(location = 0), rgen:116 */
#define main glsl_main
#include "path_trace_rgen.glsl.inc"
#undef main
GLSL_DEBUG_BLOCKS
static_assert(sizeof(PathTraceConsts) == 192 && sizeof(Instance) == 144 && sizeof(Vertex) == 80 && sizeof(Material) == 80 && sizeof(Light) == 64, "std430 layout");
} // namespace rgen
} // namespace glsl

extern "C" void ref_rgen_bind(const RefBindings* b)
{
    using namespace glsl::rgen;
    std::memcpy((void*)&u_PathTraceConsts, b->push_constants, sizeof(u_PathTraceConsts));
#if !defined(RAY_DEBUG_VIEW) // the ray-debug pipeline binds the debug buffers as set 5 instead of the two images (rgen:64-93)
    i_PreviousColor = glsl::image2D { const_cast<float*>(b->previous_color), b->width, b->height };
    i_CurrentColor  = glsl::image2D { b->current_color, b->width, b->height };
#endif
}
extern "C" void ref_rgen_invoke(uint32_t launch_x, uint32_t launch_y)
{
    using namespace glsl::rgen;
    gl_LaunchIDEXT = glsl::uvec3(launch_x, launch_y, 0);
    glsl_main();
}
