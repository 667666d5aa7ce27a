// ORACLE — TEST INFRASTRUCTURE ONLY.  The reference's path_trace_rmiss.glsl and the two visibility stages
// (path_trace_shadow.rchit / .rmiss) compiled as C++ (see gen.py).
#include "stage_common.h"
namespace glsl
{
namespace rmiss
{
static thread_local vec3 gl_WorldRayDirectionEXT;
static thread_local vec3 gl_WorldRayOriginEXT; // the two below are only read by the RAY_DEBUG_VIEW block (rmiss:40-58)
static thread_local float gl_RayTmaxEXT;
#define main glsl_main
#include "path_trace_rmiss.glsl.inc"
#undef main
GLSL_DEBUG_BLOCKS
} // namespace rmiss
namespace shadow_rchit
{
#define main glsl_main
#include "path_trace_shadow.rchit.inc"
#undef main
} // namespace shadow_rchit
namespace shadow_rmiss
{
#define main glsl_main
#include "path_trace_shadow.rmiss.inc"
#undef main
} // namespace shadow_rmiss
} // namespace glsl

extern "C" void ref_rmiss_invoke(void* payload, const RefRay* ray)
{
    using namespace glsl::rmiss;
    p_PathTracePayload      = *(const PathTracePayload*)payload;
    gl_WorldRayDirectionEXT = glsl::vec3(ray->direction[0], ray->direction[1], ray->direction[2]);
    gl_WorldRayOriginEXT    = glsl::vec3(ray->origin[0], ray->origin[1], ray->origin[2]);
    gl_RayTmaxEXT           = ray->tmax;
    glsl_main();
    *(PathTracePayload*)payload = p_PathTracePayload;
}
extern "C" void ref_shadow_rchit_invoke(void* payload)
{
    using namespace glsl::shadow_rchit;
    p_Visibility = *(const bool*)payload;
    glsl_main();
    *(bool*)payload = p_Visibility;
}
extern "C" void ref_shadow_rmiss_invoke(void* payload)
{
    using namespace glsl::shadow_rmiss;
    p_Visibility = *(const bool*)payload;
    glsl_main();
    *(bool*)payload = p_Visibility;
}
