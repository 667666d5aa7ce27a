// ORACLE — TEST INFRASTRUCTURE ONLY.  The reference's path_trace_rahit.glsl compiled as C++ (see gen.py).
#include "stage_common.h"
namespace glsl
{
namespace rahit
{
static thread_local int  gl_InstanceCustomIndexEXT, gl_GeometryIndexEXT, gl_PrimitiveID;
static thread_local bool t_ignored;
#define ignoreIntersectionEXT  \
    {                          \
        t_ignored = true;      \
        return;                \
    }
#define main glsl_main
#include "path_trace_rahit.glsl.inc"
#undef main
GLSL_DEBUG_BLOCKS
} // namespace rahit
} // namespace glsl

extern "C" void ref_rahit_bind(const RefBindings* b)
{
    using namespace glsl::rahit;
    std::memcpy((void*)&u_PathTraceConsts, b->push_constants, sizeof(u_PathTraceConsts));
    Materials.data = (Material*)b->materials;
    Instances.data = (Instance*)b->instances;
    Lights.data    = (Light*)b->lights;
    Vertices       = (VertexBuffer*)b->vertices;
    Indices        = (IndexBuffer*)b->indices;
    SubmeshInfo    = (SubmeshInfoBuffer*)b->submesh_info;
}
extern "C" int ref_rahit_invoke(const RefHit* hit)
{
    using namespace glsl::rahit;
    b_HitAttribs              = glsl::vec2(hit->u, hit->v);
    gl_InstanceCustomIndexEXT = (int)hit->instance;
    gl_GeometryIndexEXT       = (int)hit->geometry;
    gl_PrimitiveID            = (int)hit->primitive;
    t_ignored                 = false;
    glsl_main();
    return t_ignored ? 1 : 0;
}
