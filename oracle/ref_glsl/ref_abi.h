// ORACLE — TEST INFRASTRUCTURE ONLY.
// ref_abi.h — glue between the reference's shader stages compiled as C++ (stage_*.cpp) and the "driver"
// (runtime.cpp): what Vulkan provides the shaders at run time — descriptor bindings, traceRayEXT with its shader
// binding table (path_integrator.cpp:221-225 of the reference: hit group 0 = path_trace.rchit + rahit, hit group 1 =
// path_trace_shadow.rchit + rahit, miss 0 = path_trace.rmiss, miss 1 = path_trace_shadow.rmiss), texture units.
#pragma once
#include <cstdint>

extern "C"
{
struct RefHit
{
    float    t, u, v;
    uint32_t instance, geometry, primitive;
};
struct RefRay
{
    float origin[3], tmin, direction[3], tmax;
};
// descriptor sets 0..6 + push constants of the path-trace pipeline (path_trace_rchit.glsl:8-111)
struct RefBindings
{
    const void*  materials;      // set 0 binding 0: Material[]
    const void*  instances;      // set 0 binding 1: Instance[]
    const void*  lights;         // set 0 binding 2: Light[]
    void* const* vertices;       // set 1: per mesh Vertex[]
    void* const* indices;        // set 2: per mesh uint[]
    void* const* submesh_info;   // set 3: per instance uvec2[]
    const float* previous_color; // set 5: RGBA32F
    float*       current_color;  // set 6: RGBA32F
    int          width, height;
    const void*  push_constants; // 192 bytes
};

// driver services (runtime.cpp; backed by the oracle's scene, traversal and texture units — the parts the
// reference leaves to the Vulkan implementation)
void ref_drv_texture2d(int index, float u, float v, float* out4);
void ref_drv_texture_cube(const float* dir3, float* out4);
void ref_drv_trace(uint32_t flags, uint32_t sbt_offset, uint32_t miss_index, const float* origin3, float tmin, const float* dir3, float tmax, void* payload);

// RAY_DEBUG_VIEW builds: descriptor set 5 (DebugRayVertexBuffer / DebugRayDrawArgs, rchit:64-77) is ONE pair of storage
// buffers shared by every stage; the generated per-stage declarations are routed to these two blocks (stage_common.h)
struct RefDebugVertexBlock
{
    void* vertices; // DebugRayVertex[] (2 x vec4)
};
struct RefDebugDrawArgs
{
    uint32_t count, instance_count, first, base_instance;
};
RefDebugVertexBlock* ref_drv_debug_vertex_block();
RefDebugDrawArgs*    ref_drv_debug_draw_args();

// stage entry points (stage_*.cpp)
void ref_rgen_bind(const RefBindings* b);
void ref_rgen_invoke(uint32_t launch_x, uint32_t launch_y);
void ref_rchit_bind(const RefBindings* b);
void ref_rchit_invoke(void* payload, const RefHit* hit, const RefRay* ray);
void ref_rahit_bind(const RefBindings* b);
int  ref_rahit_invoke(const RefHit* hit); // 1 = ignoreIntersectionEXT was executed
void ref_rmiss_invoke(void* payload, const RefRay* ray);
void ref_shadow_rchit_invoke(void* payload);
void ref_shadow_rmiss_invoke(void* payload);
}
