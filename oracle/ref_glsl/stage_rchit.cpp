// ORACLE — TEST INFRASTRUCTURE ONLY.  The reference's path_trace_rchit.glsl compiled as C++ (see gen.py).
#include "stage_common.h"
namespace glsl
{
namespace rchit
{
static thread_local int  gl_InstanceCustomIndexEXT, gl_GeometryIndexEXT, gl_PrimitiveID;
static thread_local vec3 gl_WorldRayDirectionEXT;
static thread_local vec3 gl_WorldRayOriginEXT; // the two below are only read by the RAY_DEBUG_VIEW blocks (rchit:548-567)
static thread_local float gl_HitTEXT;
/* layout(location = 1) p_IndirectPayload, layout(location = 2) p_Visibility, rchit:123-125 */
#define GLSL_PAYLOAD_AT(loc) ((loc) == 1 ? (void*)&p_IndirectPayload : (void*)&p_Visibility)
#define main glsl_main
#include "path_trace_rchit.glsl.inc"
#undef main
GLSL_DEBUG_BLOCKS
static_assert(sizeof(PathTraceConsts) == 192 && sizeof(Instance) == 144 && sizeof(Vertex) == 80 && sizeof(Material) == 80 && sizeof(Light) == 64, "std430 layout");
} // namespace rchit
} // namespace glsl

extern "C" void ref_rchit_bind(const RefBindings* b)
{
    using namespace glsl::rchit;
    std::memcpy((void*)&u_PathTraceConsts, b->push_constants, sizeof(u_PathTraceConsts));
    Materials.data = (Material*)b->materials;
    Instances.data = (Instance*)b->instances;
    Lights.data    = (Light*)b->lights;
    Vertices       = (VertexBuffer*)b->vertices;
    Indices        = (IndexBuffer*)b->indices;
    SubmeshInfo    = (SubmeshInfoBuffer*)b->submesh_info;
}
// One closest-hit invocation.  The stage recurses through traceRayEXT (rchit:523), and in GLSL every invocation
// owns its payload / attribute / built-in variables: the caller's are saved here and restored before the
// incoming payload (rayPayloadInEXT aliases the caller's variable) is written back.
extern "C" void ref_rchit_invoke(void* payload, const RefHit* hit, const RefRay* ray)
{
    using namespace glsl::rchit;
    const PathTracePayload s_in = p_PathTracePayload, s_indirect = p_IndirectPayload;
    const bool             s_vis  = p_Visibility;
    const glsl::vec2       s_attr = b_HitAttribs;
    const int              s_i = gl_InstanceCustomIndexEXT, s_g = gl_GeometryIndexEXT, s_p = gl_PrimitiveID;
    const glsl::vec3       s_d = gl_WorldRayDirectionEXT, s_o = gl_WorldRayOriginEXT;
    const float            s_t = gl_HitTEXT;

    p_PathTracePayload        = *(const PathTracePayload*)payload;
    b_HitAttribs              = glsl::vec2(hit->u, hit->v);
    gl_InstanceCustomIndexEXT = (int)hit->instance; // scene.cpp:1254: instanceCustomIndex = mesh node index
    gl_GeometryIndexEXT       = (int)hit->geometry;
    gl_PrimitiveID            = (int)hit->primitive;
    gl_WorldRayDirectionEXT   = glsl::vec3(ray->direction[0], ray->direction[1], ray->direction[2]);
    gl_WorldRayOriginEXT      = glsl::vec3(ray->origin[0], ray->origin[1], ray->origin[2]);
    gl_HitTEXT                = hit->t;
    glsl_main();
    const PathTracePayload out = p_PathTracePayload;

    p_PathTracePayload = s_in, p_IndirectPayload = s_indirect, p_Visibility = s_vis, b_HitAttribs = s_attr;
    gl_InstanceCustomIndexEXT = s_i, gl_GeometryIndexEXT = s_g, gl_PrimitiveID = s_p, gl_WorldRayDirectionEXT = s_d, gl_WorldRayOriginEXT = s_o, gl_HitTEXT = s_t;
    *(PathTracePayload*)payload = out;
}
