// hl_scene.h — device-side view of the scene: the reference's shader ABI tables (descriptor sets 0-4,
// path_trace_rgen.glsl:8-60) plus the acceleration structures that replace the driver's TLAS/BLAS.
#pragma once
#include "../../include/helios_b200.h"
#include "hl_hd.h"

namespace hl
{
// 8-wide compressed BVH node, 80 bytes = five 16-byte vector loads.
//   n0: origin p.xyz (fp32), exponents e.xyz (u8), imask (u8: slot holds an internal child)
//   n1: child_base (u32), leaf_base (u32), meta[8] (u8: 0 empty | 001sssss internal, slot 24+i |
//       cccooooo leaf, c = unary count, o = offset from leaf_base)
//   n2: qlo.x[8], qlo.y[8]   n3: qlo.z[8], qhi.x[8]   n4: qhi.y[8], qhi.z[8]    (u8 grid coordinates)
// child box = p + q * 2^(e-127...) — see wide_node_decode_scale().
struct WideNode
{
    float    px, py, pz;
    uint8_t  ex, ey, ez, imask;
    uint32_t child_base, leaf_base;
    uint8_t  meta[8];
    uint8_t  qlox[8], qloy[8];
    uint8_t  qloz[8], qhix[8];
    uint8_t  qhiy[8], qhiz[8];
};
static_assert(sizeof(WideNode) == 80, "WideNode must be 80 bytes");

// 48-byte leaf triangle: p0, e1 = p1 - p0, e2 = p2 - p0 (fp32, computed once exactly as the traversal
// would) + identity: primitive index inside its geometry, geometry (submesh) index, opaque flag.
struct LeafTri
{
    float    p0x, p0y, p0z;
    uint32_t prim;
    float    e1x, e1y, e1z;
    uint32_t geom_flags; // bit 31 = VK_GEOMETRY_OPAQUE_BIT, bits 0..30 geometry index
    float    e2x, e2y, e2z;
    uint32_t pad;
};
static_assert(sizeof(LeafTri) == 48, "LeafTri must be 48 bytes");

// Any-hit record of one leaf triangle (same index as its LeafTri; only meshes that have a non-opaque submesh carry them):
// the three texture coordinates the any-hit shader interpolates (path_trace_rahit.glsl:121-160 fetches index -> vertex ->
// tex_coord for that), copied bit for bit at build time so that the alpha test is one 32-byte load instead of six gathers.
struct AlphaTri
{
    float u0, v0, u1, v1, u2, v2;
    float pad[2];
};
static_assert(sizeof(AlphaTri) == 32, "AlphaTri must be 32 bytes");
// per instance: where the any-hit stage finds its data
struct InstAlpha
{
    const AlphaTri* alpha = nullptr; // the instance's mesh's records (nullptr: every geometry of the mesh is opaque)
    const LeafTri*  tris  = nullptr; // that mesh's leaf array (record index = leaf pointer - tris)
    uint32_t        info_base = 0;   // first entry of the instance in SceneView::geom_alpha (= submesh_offset[instance])
    uint32_t        pad[3] = { 0, 0, 0 };
};
static_assert(sizeof(InstAlpha) == 32, "InstAlpha must be 32 bytes");
// per (instance, geometry): what fetch_albedo(...).a needs (path_trace_rahit.glsl:162-172): albedo texture or the constant alpha
struct GeomAlpha
{
    int32_t texture; // material.texture_indices0.x; -1 = none
    float   alpha;   // material.albedo.a
};

struct MeshView
{
    const hl_vertex* vertices;
    const uint32_t*  indices;
    const WideNode*  nodes; // BLAS, root = node 0
    const LeafTri*   tris;
    uint32_t         n_tris;
    uint32_t         n_submeshes;
};

struct TexView
{
    const void* texels;
    uint32_t    w, h;
    int32_t     format;
    uint32_t    pad;
};

struct EnvView
{
    // 6 faces of (size + 2)^2 texels: every face carries a one-texel border holding the texels of the adjoining faces
    // (corners: the mean of the three that exist), filled once per upload / sky bake (hl_tex.h cube_pad_texel), so that
    // the seamless bilinear lookup Vulkan prescribes for cube maps is one branch-free 2 x 2 fetch
    const f4* faces;
    uint32_t  size; // 0 = black default cube map
};

struct SceneView
{
    const hl_material* materials;
    const hl_instance* instances;
    const float*       inst_inv;       // 12 floats per instance: world->object 3x4, row-major
    const uint32_t*    submesh_info;   // flattened (prim offset, material) pairs
    const uint32_t*    submesh_offset; // per instance: first pair index in submesh_info
    const hl_light*    lights;
    const MeshView*    meshes;
    const TexView*     textures;
    const float*       lut8; // 3 x 256: unorm, srgb, snorm decode tables
    EnvView            env;
    const WideNode*    tlas_nodes; // leaves reference tlas_leaf[]
    const uint32_t*    tlas_leaf;  // instance index per TLAS leaf slot
    uint32_t           n_instances;
    uint32_t           n_lights;
    uint32_t           single_identity; // 1: one instance with an identity transform -> BLAS traversed directly
    uint32_t           pad0 = 0;
    const uint32_t*    inst_identity = nullptr; // bit i: instance i has an identity model matrix (object space = world space: entering and
                                                // leaving it needs no ray transform and no new reciprocal direction); nullptr = none
    const InstAlpha*   inst_alpha = nullptr; // [n_instances] any-hit records (nullptr: the any-hit stage walks the vertex tables)
    const GeomAlpha*   geom_alpha = nullptr; // [sum of submesh counts] parallel to submesh_info
};

struct Hit
{
    float    t, u, v;
    uint32_t instance, geometry, primitive; // 0xFFFFFFFF = miss
};

#define HL_MISS 0xFFFFFFFFu
#define HL_RAY_OPAQUE 1u    /* gl_RayFlagsOpaqueEXT: the any-hit stage never runs */
#define HL_RAY_TERMINATE 2u /* gl_RayFlagsTerminateOnFirstHitEXT */
} // namespace hl
