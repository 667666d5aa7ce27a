// hl_bvh.h — software replacement for traceRayEXT (call sites path_trace_rgen.glsl:205,
// path_trace_rchit.glsl:438 and :523): two-level traversal of 8-wide compressed BVHs with the alpha-test
// any-hit stage (path_trace_rahit.glsl:174-188) resolved inside the loop.
//
// Semantics fixed here (Vulkan leaves them to the driver):
//   * a hit needs tmin < t < tmax (strict) and Moeller-Trumbore u >= 0, v >= 0, u + v <= 1, det != 0,
//     evaluated in object space with the un-normalised transformed direction (t is space independent);
//   * closest hit = min t, ties -> min (instance, geometry, primitive): independent of traversal order;
//   * node tests are conservative (never cull a triangle the fp32 triangle test would accept).
// Node layout and the octant-ordered traversal follow the 8-wide compressed BVH of Ylitie, Karras and
// Laine (HPG 2017); the code is written from the paper's description.
#pragma once
#include "hl_scene.h"
#include "hl_tex.h"

namespace hl
{
#ifndef HL_STACK_FAST
#define HL_STACK_FAST 12 /* entries kept in fast (shared) memory per ray */
#endif
#define HL_STACK_SPILL 52 /* further entries in thread-local memory; total depth 64 */

// traversal stack: the first HL_STACK_FAST entries live in `fast` (shared memory on the GPU,
// interleaved with `stride` so that a warp's accesses are conflict-free), deeper entries spill.
struct TravStack
{
    u2* fast;
    int stride;
    int sp;
    u2  spill[HL_STACK_SPILL];
    HL_HD void push(u2 e)
    {
        if (sp < HL_STACK_FAST)
            fast[sp * stride] = e;
        else if (sp - HL_STACK_FAST < HL_STACK_SPILL)
            spill[sp - HL_STACK_FAST] = e;
        sp++;
    }
    HL_HD u2 pop()
    {
        sp--;
        if (sp < HL_STACK_FAST) return fast[sp * stride];
        if (sp - HL_STACK_FAST < HL_STACK_SPILL) return spill[sp - HL_STACK_FAST];
        u2 z;
        z.x = 0, z.y = 0;
        return z;
    }
};

struct RayCtx
{
    f3       o, d, idir;
    uint32_t octinv; // 3 bits: bit2 = x, bit1 = y, bit0 = z; set when the direction component is >= 0
    float    tmin;
};

HL_HD float safe_rcp_dir(float d)
{
    // |d| < 1e-20 (including +-0) -> +-1e20 keeps the slab test finite; NaN stays NaN
    const float lim = 1e-20f;
    if (fabsf(d) < lim) return (f2u(d) >> 31) ? -1e20f : 1e20f;
    return 1.0f / d;
}
HL_HD RayCtx make_ray_ctx(f3 o, f3 d, float tmin)
{
    RayCtx r;
    r.o = o, r.d = d, r.tmin = tmin;
    r.idir   = mk3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
    r.octinv = (r.idir.x < 0.0f ? 0u : 4u) | (r.idir.y < 0.0f ? 0u : 2u) | (r.idir.z < 0.0f ? 0u : 1u);
    return r;
}

// Tests the 8 quantised child boxes of one node; returns the hit mask: bits 24..31 = internal children in
// octant-permuted order (highest bit = visit first), bits 0..23 = leaf primitives (offset from leaf_base).
HL_HD uint32_t intersect_children(const WideNode& n, const RayCtx& r, float tbest)
{
    const float adjx = u2f((uint32_t)n.ex << 23) * r.idir.x;
    const float adjy = u2f((uint32_t)n.ey << 23) * r.idir.y;
    const float adjz = u2f((uint32_t)n.ez << 23) * r.idir.z;
    const float orgx = (n.px - r.o.x) * r.idir.x;
    const float orgy = (n.py - r.o.y) * r.idir.y;
    const float orgz = (n.pz - r.o.z) * r.idir.z;
    // conservative slack (absolute, in t): covers the rounding of org/adj/fma and the fact that the fp32
    // triangle test can accept rays that miss the exact box by a few ulp of the ray-box distance
    const float slack = 1.9073486e-6f /* 2^-19 */ *
                        (fmaxf(fabsf(orgx), fmaxf(fabsf(orgy), fabsf(orgz))) + 255.0f * fmaxf(fabsf(adjx), fmaxf(fabsf(adjy), fabsf(adjz))));
    const float tlo_bound = r.tmin - slack;
    const float thi_bound = tbest + slack;
    const bool  nx = r.idir.x < 0.0f, ny = r.idir.y < 0.0f, nz = r.idir.z < 0.0f;
    uint32_t    hitmask = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 8; j++)
    {
        const uint32_t meta = n.meta[j];
        if (meta == 0) continue;
        const float qnx = (float)(nx ? n.qhix[j] : n.qlox[j]), qfx = (float)(nx ? n.qlox[j] : n.qhix[j]);
        const float qny = (float)(ny ? n.qhiy[j] : n.qloy[j]), qfy = (float)(ny ? n.qloy[j] : n.qhiy[j]);
        const float qnz = (float)(nz ? n.qhiz[j] : n.qloz[j]), qfz = (float)(nz ? n.qloz[j] : n.qhiz[j]);
        const float tnear = fmaxf(fmaxf(hl_fma(qnx, adjx, orgx), hl_fma(qny, adjy, orgy)), fmaxf(hl_fma(qnz, adjz, orgz), tlo_bound));
        const float tfar  = fminf(fminf(hl_fma(qfx, adjx, orgx), hl_fma(qfy, adjy, orgy)), fminf(hl_fma(qfz, adjz, orgz), thi_bound));
        if (tnear <= tfar + slack)
        {
            const bool     inner = (meta & 0x18u) == 0x18u;
            const uint32_t bit   = inner ? ((meta ^ r.octinv) & 0x1Fu) : (meta & 0x1Fu);
            hitmask |= (meta >> 5) << bit;
        }
    }
    return hitmask;
}

// Generic octant-ordered traversal of one wide BVH.  leaf(index, tbest) handles one leaf primitive,
// may shrink tbest, and returns true to terminate the whole query (TerminateOnFirstHit).
template <class Leaf>
HL_HD bool traverse_wide(const WideNode* nodes, const RayCtx& r, float& tbest, TravStack& st, Leaf& leaf)
{
    const int sp0 = st.sp;
    u2        ngroup;
    ngroup.x = 0, ngroup.y = 0x80000000u;
    for (;;)
    {
        u2 tgroup;
        tgroup.x = 0, tgroup.y = 0;
        if (ngroup.y > 0x00FFFFFFu)
        {
            const uint32_t hits = ngroup.y;
            const int      bit  = hl_bfind(hits);
            const uint32_t base = ngroup.x;
            ngroup.y &= ~(1u << bit);
            if (ngroup.y > 0x00FFFFFFu) st.push(ngroup);
            const uint32_t  slot = (uint32_t)(bit - 24) ^ r.octinv;
            const uint32_t  rel  = (uint32_t)hl_popc((hits & 0xFFu) & ~(0xFFFFFFFFu << slot));
            const WideNode& n    = nodes[base + rel];
            const uint32_t  mask = intersect_children(n, r, tbest);
            ngroup.x = n.child_base, ngroup.y = (mask & 0xFF000000u) | n.imask;
            tgroup.x = n.leaf_base, tgroup.y = mask & 0x00FFFFFFu;
        }
        while (tgroup.y)
        {
            const int i = hl_bfind(tgroup.y);
            tgroup.y &= ~(1u << i);
            if (leaf(tgroup.x + (uint32_t)i, tbest))
            {
                st.sp = sp0;
                return true;
            }
        }
        if (ngroup.y <= 0x00FFFFFFu)
        {
            if (st.sp == sp0) break;
            ngroup = st.pop();
        }
    }
    return false;
}

// path_trace_rahit.glsl:174-188: true when the candidate intersection is ignored (albedo alpha < 0.1)
HL_HD bool any_hit_ignores(const SceneView& s, uint32_t inst, uint32_t geom, uint32_t prim, float bu, float bv)
{
    const hl_instance& I    = s.instances[inst];
    const MeshView&    m    = s.meshes[I.mesh_index];
    const uint32_t*    info = s.submesh_info + 2 * (size_t)(s.submesh_offset[inst] + geom); // fetch_hit_info
    const uint32_t     pid  = prim + info[0];
    const hl_material& mat  = s.materials[info[1]];
    if (mat.texture_indices0[0] == -1) return mat.albedo[3] < 0.1f;
    const float* t0 = m.vertices[m.indices[3 * (size_t)pid + 0]].tex_coord;
    const float* t1 = m.vertices[m.indices[3 * (size_t)pid + 1]].tex_coord;
    const float* t2 = m.vertices[m.indices[3 * (size_t)pid + 2]].tex_coord;
    const float  b0 = 1.0f - bu - bv;
    const float  tu = t0[0] * b0 + t1[0] * bu + t2[0] * bv;
    const float  tv = t0[1] * b0 + t1[1] * bu + t2[1] * bv;
    return sample_texture_lod0(s, mat.texture_indices0[0], tu, tv).w < 0.1f;
}

struct TriLeaf
{
    const SceneView* s;
    const LeafTri*   tris;
    f3               o, d; // object space
    float            tmin, tmax;
    uint32_t         inst, flags;
    Hit*             best;
    HL_HD bool       operator()(uint32_t index, float& tbest)
    {
        const LeafTri tr   = tris[index];
        const f3      e1   = mk3(tr.e1x, tr.e1y, tr.e1z);
        const f3      e2   = mk3(tr.e2x, tr.e2y, tr.e2z);
        const f3      pvec = cross(d, e2);
        const float   det  = dot(e1, pvec);
        if (det == 0.0f || det != det) return false;
        const float inv  = 1.0f / det;
        const f3    tvec = o - mk3(tr.p0x, tr.p0y, tr.p0z);
        const float u    = dot(tvec, pvec) * inv;
        if (!(u >= 0.0f && u <= 1.0f)) return false;
        const f3    qvec = cross(tvec, e1);
        const float v    = dot(d, qvec) * inv;
        if (!(v >= 0.0f && u + v <= 1.0f)) return false;
        const float t = dot(e2, qvec) * inv;
        if (!(t > tmin && t < tmax)) return false;
        const uint32_t geom = tr.geom_flags & 0x7FFFFFFFu;
        Hit&           b    = *best;
        if (!(t < b.t))
        {
            if (t > b.t) return false;
            // equal t: lexicographic (instance, geometry, primitive)
            if (inst != b.instance)
            {
                if (inst > b.instance) return false;
            }
            else if (geom != b.geometry)
            {
                if (geom > b.geometry) return false;
            }
            else if (tr.prim >= b.primitive)
                return false;
        }
        if (!(flags & HL_RAY_OPAQUE) && !(tr.geom_flags >> 31) && any_hit_ignores(*s, inst, geom, tr.prim, u, v)) return false;
        b.t = t, b.u = u, b.v = v, b.instance = inst, b.geometry = geom, b.primitive = tr.prim;
        tbest = t;
        return (flags & HL_RAY_TERMINATE) != 0;
    }
};

struct InstLeaf
{
    const SceneView* s;
    f3               o, d; // world space
    float            tmin, tmax;
    uint32_t         flags;
    Hit*             best;
    TravStack*       st;
    HL_HD bool       operator()(uint32_t index, float& tbest)
    {
        const uint32_t inst = s->tlas_leaf[index];
        const float*   m    = s->inst_inv + 12 * (size_t)inst;
        TriLeaf        leaf;
        leaf.o.x = (m[0] * o.x + m[1] * o.y + m[2] * o.z) + m[3];
        leaf.o.y = (m[4] * o.x + m[5] * o.y + m[6] * o.z) + m[7];
        leaf.o.z = (m[8] * o.x + m[9] * o.y + m[10] * o.z) + m[11];
        leaf.d.x = m[0] * d.x + m[1] * d.y + m[2] * d.z;
        leaf.d.y = m[4] * d.x + m[5] * d.y + m[6] * d.z;
        leaf.d.z = m[8] * d.x + m[9] * d.y + m[10] * d.z;
        const MeshView& mesh = s->meshes[s->instances[inst].mesh_index];
        if (mesh.n_tris == 0) return false;
        leaf.s = s, leaf.tris = mesh.tris, leaf.tmin = tmin, leaf.tmax = tmax, leaf.inst = inst, leaf.flags = flags, leaf.best = best;
        const RayCtx r = make_ray_ctx(leaf.o, leaf.d, tmin);
        return traverse_wide(mesh.nodes, r, tbest, *st, leaf);
    }
};

// traceRayEXT: fills `best` (instance == HL_MISS when nothing was hit)
HL_HD void trace_ray(const SceneView& s, f3 o, float tmin, f3 d, float tmax, uint32_t flags, Hit& best, TravStack& st)
{
    best.t = tmax, best.u = 0.0f, best.v = 0.0f;
    best.instance = best.geometry = best.primitive = HL_MISS;
    float tbest   = tmax;
    st.sp         = 0;
    if (s.n_instances == 0) return;
    if (s.single_identity)
    {
        const MeshView& mesh = s.meshes[s.instances[0].mesh_index];
        if (mesh.n_tris == 0) return;
        TriLeaf leaf;
        leaf.s = &s, leaf.tris = mesh.tris, leaf.o = o, leaf.d = d, leaf.tmin = tmin, leaf.tmax = tmax, leaf.inst = 0, leaf.flags = flags, leaf.best = &best;
        const RayCtx r = make_ray_ctx(o, d, tmin);
        traverse_wide(mesh.nodes, r, tbest, st, leaf);
        return;
    }
    InstLeaf leaf;
    leaf.s = &s, leaf.o = o, leaf.d = d, leaf.tmin = tmin, leaf.tmax = tmax, leaf.flags = flags, leaf.best = &best, leaf.st = &st;
    const RayCtx r = make_ray_ctx(o, d, tmin);
    traverse_wide(s.tlas_nodes, r, tbest, st, leaf);
}
} // namespace hl
