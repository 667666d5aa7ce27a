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
//
// Execution shape: ONE step function per ray for both levels (an instance entry pushes a sentinel and
// switches the ray to object space; popping the sentinel switches back).  The callers loop on a warp vote, so
// the 32 rays of a warp re-converge at the top of every step instead of drifting into private instruction
// streams, and the persistent trace kernels refill finished lanes with new rays between steps.
#pragma once
#include "hl_scene.h"
#include "hl_tex.h"

namespace hl
{
#if defined(HL_TRAVERSAL_STATS) && !defined(__CUDA_ARCH__)
// emulator-only instrumentation (tests/emul): nodes visited / leaf primitives tested per query
struct TraversalStats
{
    uint64_t nodes = 0, leaves = 0;
};
inline TraversalStats& traversal_stats()
{
    static thread_local TraversalStats s;
    return s;
}
inline bool& emul_force_postpone()
{
    static bool f = false;
    return f;
}
#define HL_STAT_NODE() (traversal_stats().nodes++)
#define HL_STAT_LEAF() (traversal_stats().leaves++)
#else
#define HL_STAT_NODE() ((void)0)
#define HL_STAT_LEAF() ((void)0)
#endif

#if defined(__CUDA_ARCH__)
#define HL_WARP_BALLOT(p) __ballot_sync(0xFFFFFFFFu, (p))
#else
#define HL_WARP_BALLOT(p) ((p) ? 1u : 0u)
#endif

#ifndef HL_STACK_FAST
#define HL_STACK_FAST 12 /* entries kept in fast (shared) memory per ray */
#endif
#ifndef HL_STACK_SPILL
#define HL_STACK_SPILL 52 /* further entries in thread-local memory; total depth 64 */
#endif

// Stack overflow is counted, never silent: hl_get_counters fails loudly while the count is non-zero.  An entry that
// does not fit is dropped (that subtree is lost for this ray — the result is flagged as untrustworthy, but nothing
// dropped is ever popped, so the instance sentinel cannot be confused with a lost entry and no index leaves its array).
#if defined(__CUDACC__)
static __device__ unsigned long long g_trav_overflow; // per translation unit; hl_wavefront.cu owns the trace kernels
#endif
#if !defined(__CUDA_ARCH__)
inline unsigned long long& emul_trav_overflow()
{
    static unsigned long long n = 0;
    return n;
}
#endif
HL_HD void note_stack_overflow()
{
#if defined(__CUDA_ARCH__)
    atomicAdd(&g_trav_overflow, 1ull);
#else
    emul_trav_overflow()++;
#endif
}
// traversal stack: the first HL_STACK_FAST entries live in fast memory, deeper entries spill to thread-local memory.
// GPU: the fast part is ONE block-wide array in shared memory (entry k of thread t at [k * HL_TRACE_BLOCK + t]: a warp's
// accesses are conflict-free), named directly so that every access is one LDS.64 / STS.64, and the struct holds nothing but
// scalars and a pointer to the caller's spill array — it stays in registers.  (Round 2, from the SASS: with the spill array
// INSIDE the struct the whole struct lived in local memory, and every push / pop re-loaded the stack pointer, the base
// pointer and the stride from there and went through generic 32-bit loads.)
#ifndef HL_TRACE_BLOCK
#define HL_TRACE_BLOCK 128 /* threads per block of every kernel that traces */
#endif
#if defined(__CUDACC__)
static __shared__ u2 g_trav_fast[HL_STACK_FAST * HL_TRACE_BLOCK];
#endif
struct TravStack
{
#if defined(__CUDA_ARCH__)
    uint32_t fast; // shared-space byte address of this thread's entry 0 (entries HL_TRACE_BLOCK * 8 bytes apart)
#else
    u2* fast; // host (emulator): HL_STACK_FAST entries
#endif
    int      sp;
    uint8_t* pairs; // GPU: HL_COOP_TABLE bytes of shared memory per WARP (cooperative triangle phase, coop_triangles below)
    u2*      spill; // HL_STACK_SPILL entries of the caller's thread-local memory
    HL_HD void init(u2* spill_mem, u2* host_fast)
    {
#if defined(__CUDA_ARCH__)
        fast = (uint32_t)__cvta_generic_to_shared(&g_trav_fast[threadIdx.x]);
        (void)host_fast;
#else
        fast = host_fast;
#endif
        sp = 0, spill = spill_mem, pairs = nullptr;
    }
    HL_HD bool has_room(int n) const { return sp + n <= HL_STACK_FAST + HL_STACK_SPILL; }
    HL_HD void push(u2 e)
    {
        if (sp < HL_STACK_FAST)
        {
#if defined(__CUDA_ARCH__)
            // (volatile keeps the stack's accesses in program order among themselves; nothing else touches this memory)
            asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(fast + (uint32_t)sp * (HL_TRACE_BLOCK * 8u)), "r"(e.x), "r"(e.y));
#else
            fast[sp] = e;
#endif
        }
        else if (sp - HL_STACK_FAST < HL_STACK_SPILL)
            spill[sp - HL_STACK_FAST] = e;
        else
        {
            note_stack_overflow();
            return;
        }
        sp++;
    }
    HL_HD u2 pop()
    {
        sp--;
        if (sp < HL_STACK_FAST)
        {
#if defined(__CUDA_ARCH__)
            u2 e;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e.x), "=r"(e.y) : "r"(fast + (uint32_t)sp * (HL_TRACE_BLOCK * 8u)));
            return e;
#else
            return fast[sp];
#endif
        }
        return spill[sp - HL_STACK_FAST];
    }
};

struct RayCtx
{
    f3       o, d, idir;
    uint32_t octinv; // 3 bits: bit2 = x, bit1 = y, bit0 = z; set when the direction component is >= 0
};

HL_HD float safe_rcp_dir(float d)
{
    // |d| < 1e-20 (including +-0) -> +-1e20 keeps the slab test finite; NaN stays NaN
    const float lim = 1e-20f;
    if (fabsf(d) < lim) return (f2u(d) >> 31) ? -1e20f : 1e20f;
#if defined(__CUDA_ARCH__)
    // one MUFU.RCP (relative error 2^-23) instead of the ~10-instruction IEEE division: the reciprocal direction only
    // feeds the conservative node tests, whose slack (2^-16 of the distance, intersect_children) is 100 x wider; hit
    // parameters never see it.  (ncu, foliage scene: the division was 7 % of k_extend's warp instructions at 5.6 lanes.)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
    return r;
#else
    return 1.0f / d;
#endif
}
HL_HD RayCtx make_ray_ctx(f3 o, f3 d)
{
    RayCtx r;
    r.o = o, r.d = d;
    r.idir   = mk3(safe_rcp_dir(d.x), safe_rcp_dir(d.y), safe_rcp_dir(d.z));
    r.octinv = (r.idir.x < 0.0f ? 0u : 4u) | (r.idir.y < 0.0f ? 0u : 2u) | (r.idir.z < 0.0f ? 0u : 1u);
    return r;
}

struct U4
{
    uint32_t x, y, z, w;
};
HL_HD U4 load_u4(const void* p)
{
#if defined(__CUDA_ARCH__)
    const uint4 v = __ldg((const uint4*)p);
    U4          r;
    r.x = v.x, r.y = v.y, r.z = v.z, r.w = v.w;
    return r;
#else
    U4 r;
    memcpy(&r, p, 16);
    return r;
#endif
}
// Quantised plane byte -> t in two instructions: one PRMT drops byte i of `word` into mantissa bits 8..15 of the axis'
// scale 2^(e-127) (`E` = e << 23, the node's exponent byte in place), giving F = 2^(e-127) * (1 + b * 2^-15) exactly, and
// one FMA evaluates F * A + B with A = idir * 2^15 and B = (org -+ slack) - 2^(e-127) * A, i.e. b * 2^(e-127) * idir + org
// -+ slack.  The only new error is the rounding of B (<= 2^-24 |org| + 2^-9 |adj|), which the slack below absorbs.  A flat
// axis (e = 0, all bytes 0) gives F = 0 and t = B = org -+ slack.  (Round 2: the PRMT used to merge the byte into the
// CONSTANT 1.0f; with two immediates — selector and constant — ptxas kept the selector in a register and re-materialised it
// before nearly every PRMT: 43 of the 311 instructions of a node visit were such moves.  With the exponent word as the
// second source the selector is the only immediate.)
HL_HD float plane_t(uint32_t word, int i, uint32_t E, float A, float B) { return hl_fma(u2f(hl_prmt(word, E, 0x7604u | ((uint32_t)i << 4))), A, B); }

// Tests the 8 quantised child boxes of one node (five 16-byte loads); returns the hit mask: bits 24..31 =
// internal children in octant-permuted order (highest bit = visit first), bits 0..23 = leaf primitives.
HL_HD uint32_t intersect_children(const WideNode* node, const RayCtx& r, float tmin, float tbest, uint32_t& child_base, uint32_t& leaf_base, uint32_t& imask)
{
    const U4 n0 = load_u4((const char*)node + 0);
    const U4 n1 = load_u4((const char*)node + 16);
    const U4 n2 = load_u4((const char*)node + 32);
    const U4 n3 = load_u4((const char*)node + 48);
    const U4 n4 = load_u4((const char*)node + 64);
    child_base  = n1.x, leaf_base = n1.y, imask = n0.w >> 24;
    const uint32_t Ex = (n0.w << 23) & 0x7F800000u, Ey = (n0.w << 15) & 0x7F800000u, Ez = (n0.w << 7) & 0x7F800000u;
    const float    adjx = u2f(Ex) * r.idir.x, adjy = u2f(Ey) * r.idir.y, adjz = u2f(Ez) * r.idir.z;
    const float dx = u2f(n0.x) - r.o.x, dy = u2f(n0.y) - r.o.y, dz = u2f(n0.z) - r.o.z;
    const float orgx = dx * r.idir.x, orgy = dy * r.idir.y, orgz = dz * r.idir.z;
    // conservative per-axis slack (in t): covers the rounding of org/adj/B/fma and the fact that the fp32
    // triangle test can accept rays that miss the exact triangle (hence its box) by a few ulp of the DISTANCE
    // between ray origin and triangle — the distance over all axes (D below), not the one along the axis being
    // tested: a ray that runs almost parallel to an axis-aligned wall is close to the box in that axis and far
    // away in the others (measured with tools/tree_invariance.py: with a slack relative to |org| alone, 1 of
    // 2.6e8 secondary rays of the city scene lost an edge hit, u + v = 0.99999, in one tree and not in another).
    // The slack is folded into the fma addend: near planes use org - s, far planes org + s.  (Scaled per axis
    // by |idir|, not a common maximum in t: an axis with a near-zero direction component has a huge |idir| and
    // would otherwise open every box of the tree.)  2^-16 * 192 |adj| = 2^-8.4 |adj| > 2^-9 |adj| (rounding of
    // B) + the 2^-19 * 255 |adj| of the exact-byte formulation.
    // The factor on D: the fp32 Moeller-Trumbore test places the hit with an in-plane error of about
    // k / cos(theta) * 2^-24 * D (k a handful of roundings, theta the angle between ray and triangle normal), so a
    // grazing ray can be accepted while passing the triangle's box at that distance.  2^-19 D covered k / cos(theta)
    // <= 32 and lost one edge hit (u + v = 0.99995 on a wall seen at 107 units, grazing) of 5e7 rays of the 4K city
    // frame in the LBVH tree but not in the SAH tree (tests/test_gpu_fullsize.py found it); 2^-16 D covers
    // k / cos(theta) <= 256, i.e. rays within 0.5 degrees of the plane, and inflates a box by 1.5e-5 of its distance.
#ifndef HL_NODE_SLACK
#define HL_NODE_SLACK 1.5258789e-5f /* 2^-16 */
#endif
    const float C  = HL_NODE_SLACK;
    const float D  = fmaxf(fmaxf(fabsf(dx), fabsf(dy)), fabsf(dz));
    const float sx = C * (fabsf(r.idir.x) * D + 192.0f * fabsf(adjx));
    const float sy = C * (fabsf(r.idir.y) * D + 192.0f * fabsf(adjy));
    const float sz = C * (fabsf(r.idir.z) * D + 192.0f * fabsf(adjz));
    const float Ax = adjx * 32768.0f, Ay = adjy * 32768.0f, Az = adjz * 32768.0f;             // 2^(e-127) * A of plane_t
    const float Rx = r.idir.x * 32768.0f, Ry = r.idir.y * 32768.0f, Rz = r.idir.z * 32768.0f; // A of plane_t
    const float Bnx = (orgx - sx) - Ax, Bfx = (orgx + sx) - Ax;
    const float Bny = (orgy - sy) - Ay, Bfy = (orgy + sy) - Ay;
    const float Bnz = (orgz - sz) - Az, Bfz = (orgz + sz) - Az;
    const bool  nx = r.idir.x < 0.0f, ny = r.idir.y < 0.0f, nz = r.idir.z < 0.0f;
    const uint32_t oct4 = r.octinv * 0x01010101u;
    uint32_t       hitmask = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int h = 0; h < 2; h++)
    {
        // words holding children 4h..4h+3 of each plane set
        const uint32_t meta4 = h ? n1.w : n1.z;
        const uint32_t lox = h ? n2.y : n2.x, loy = h ? n2.w : n2.z;
        const uint32_t loz = h ? n3.y : n3.x, hix = h ? n3.w : n3.z;
        const uint32_t hiy = h ? n4.y : n4.x, hiz = h ? n4.w : n4.z;
        const uint32_t nearx = nx ? hix : lox, farx = nx ? lox : hix;
        const uint32_t neary = ny ? hiy : loy, fary = ny ? loy : hiy;
        const uint32_t nearz = nz ? hiz : loz, farz = nz ? loz : hiz;
        // four children at a time: inner children (meta & 0x18 == 0x18) take the octant-permuted bit index
        const uint32_t inner4 = ((meta4 & (meta4 << 1)) & 0x10101010u) >> 4; // 0x01 per inner child
        const uint32_t bit4   = (meta4 ^ (oct4 & (inner4 * 0xFFu))) & 0x1F1F1F1Fu;
        const uint32_t cnt4   = (meta4 >> 5) & 0x07070707u; // 0 for an empty slot: contributes nothing
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 4; j++)
        {
            const float tnear = fmaxf(fmaxf(plane_t(nearx, j, Ex, Rx, Bnx), plane_t(neary, j, Ey, Ry, Bny)), fmaxf(plane_t(nearz, j, Ez, Rz, Bnz), tmin));
            const float tfar  = fminf(fminf(plane_t(farx, j, Ex, Rx, Bfx), plane_t(fary, j, Ey, Ry, Bfy)), fminf(plane_t(farz, j, Ez, Rz, Bfz), tbest));
            if (tnear <= tfar) hitmask |= hl_prmt(cnt4, 0u, 0x4440u | (uint32_t)j) << hl_prmt(bit4, 0u, 0x4440u | (uint32_t)j);
        }
    }
    return hitmask;
}

// path_trace_rahit.glsl:174-188: true when the candidate intersection is ignored (albedo alpha < 0.1).  `tri` = the leaf
// record of the candidate.  With any-hit records (SceneView::inst_alpha, built by hl_scene_set_tables) the texture coordinates
// come from the AlphaTri beside the leaf and the material's alpha source from the per-geometry table: the same values the
// walk instance -> mesh -> index -> vertex -> tex_coord and submesh_info -> material reads, in two loads instead of nine.
HL_HD bool any_hit_ignores(const SceneView& s, uint32_t inst, uint32_t geom, uint32_t prim, float bu, float bv, const LeafTri* tri)
{
    const float b0 = 1.0f - bu - bv;
    if (s.inst_alpha)
    {
        const InstAlpha ia = s.inst_alpha[inst];
        if (ia.alpha)
        {
            const GeomAlpha ga = s.geom_alpha[ia.info_base + geom];
            if (ga.texture == -1) return ga.alpha < 0.1f;
            const AlphaTri& a  = ia.alpha[tri - ia.tris];
            const float     tu = a.u0 * b0 + a.u1 * bu + a.u2 * bv;
            const float     tv = a.v0 * b0 + a.v1 * bu + a.v2 * bv;
            return sample_texture_alpha_lod0(s, ga.texture, tu, tv) < 0.1f;
        }
    }
    const hl_instance& I    = s.instances[inst];
    const MeshView&    m    = s.meshes[I.mesh_index];
    const uint32_t*    info = s.submesh_info + 2 * (size_t)(s.submesh_offset[inst] + geom); // fetch_hit_info
    const uint32_t     pid  = prim + info[0];
    const hl_material& mat  = s.materials[info[1]];
    if (mat.texture_indices0[0] == -1) return mat.albedo[3] < 0.1f;
    const float* t0 = m.vertices[m.indices[3 * (size_t)pid + 0]].tex_coord;
    const float* t1 = m.vertices[m.indices[3 * (size_t)pid + 1]].tex_coord;
    const float* t2 = m.vertices[m.indices[3 * (size_t)pid + 2]].tex_coord;
    const float  tu = t0[0] * b0 + t1[0] * bu + t2[0] * bv;
    const float  tv = t0[1] * b0 + t1[1] * bu + t2[1] * bv;
    return sample_texture_alpha_lod0(s, mat.texture_indices0[0], tu, tv) < 0.1f;
}

// Moeller-Trumbore against one 48-byte leaf record; updates `best` per the closest-hit / tie rule.
// Returns true when the candidate was accepted.
HL_HD bool test_leaf_triangle(const SceneView& s, const LeafTri* tri, f3 o, f3 d, float tmin, float tmax, uint32_t inst, uint32_t flags, Hit& b)
{
    const U4    a = load_u4((const char*)tri + 0), e1w = load_u4((const char*)tri + 16), e2w = load_u4((const char*)tri + 32);
    const f3    e1   = mk3(u2f(e1w.x), u2f(e1w.y), u2f(e1w.z));
    const f3    e2   = mk3(u2f(e2w.x), u2f(e2w.y), u2f(e2w.z));
    const f3    pvec = cross(d, e2);
    const float det  = dot(e1, pvec);
    if (det == 0.0f || det != det) return false;
    const float inv  = 1.0f / det;
    const f3    tvec = o - mk3(u2f(a.x), u2f(a.y), u2f(a.z));
    const float u    = dot(tvec, pvec) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    const f3    qvec = cross(tvec, e1);
    const float v    = dot(d, qvec) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return false;
    const float t = dot(e2, qvec) * inv;
    if (!(t > tmin && t < tmax)) return false;
    const uint32_t prim = a.w, geom = e1w.w & 0x7FFFFFFFu;
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
        else if (prim >= b.primitive)
            return false;
    }
    if (!(flags & HL_RAY_OPAQUE) && !(e1w.w >> 31) && any_hit_ignores(s, inst, geom, prim, u, v, tri)) return false;
    b.t = t, b.u = u, b.v = v, b.instance = inst, b.geometry = geom, b.primitive = prim;
    return true;
}

// An instance whose model matrix is the identity needs no ray transform: (1*x + 0*y + 0*z) + 0 = x bit for bit — as long as
// no component is NaN / infinite (0 * inf = NaN in the real transform) and no direction component is a signed zero (the sum
// of the 0 * d terms decides the sign of a zero result); such rays take the general path, so results never depend on this.
HL_HD bool instance_passthrough(const SceneView& s, uint32_t id, f3 o, f3 d)
{
    if (!s.inst_identity || !((s.inst_identity[id >> 5] >> (id & 31u)) & 1u)) return false;
    const float big = 3.0e38f;
    return d.x != 0.0f && d.y != 0.0f && d.z != 0.0f && fabsf(d.x) < big && fabsf(d.y) < big && fabsf(d.z) < big && fabsf(o.x) < big && fabsf(o.y) < big && fabsf(o.z) < big;
}

// ---- traceRayEXT as a resumable state machine ---------------------------------------------------------
// trav_begin / trav_busy / trav_step_warp: one step = (pop a stack entry if nothing is current, visit one
// node) followed by up to HL_TRI_PER_STEP triangle tests (or one instance entry at the top level).  The
// persistent trace kernels call trav_step_warp in a warp-convergent loop and hand finished lanes a new ray
// between steps; trace_ray() below runs one query per lane to completion (tail kernel, generic trace entry
// point, emulator).
#ifndef HL_TRI_PER_STEP
#define HL_TRI_PER_STEP 2
#endif
#ifndef HL_PREFETCH_NEXT_NODE
#define HL_PREFETCH_NEXT_NODE 0
#endif
struct Trav
{
    RayCtx          r;    // ray in the space of the tree being traversed
    f3              o, d; // world-space ray
    float           tmin, tmax;
    uint32_t        flags;
    const WideNode* nodes;
    const LeafTri*  tris;
    uint32_t        inst; // HL_MISS = top level
    u2              ngroup, tgroup;
    Hit             best; // instance == HL_MISS when nothing was hit
};

// where every query of a scene starts: the top-level tree, or — a scene of one identity instance — that instance's mesh.
// Scene-wide, so the persistent kernels look it up once per thread instead of once per ray (three dependent loads).
struct TravStart
{
    const WideNode* nodes;
    const LeafTri*  tris;
    uint32_t        inst, ngroup_y; // ngroup_y == 0: nothing to traverse
};
HL_HD TravStart trav_start(const SceneView& s)
{
    TravStart b;
    b.nodes = s.tlas_nodes, b.tris = nullptr, b.inst = HL_MISS, b.ngroup_y = 0;
    if (s.n_instances != 0)
    {
        if (s.single_identity)
        {
            const MeshView& mesh = s.meshes[s.instances[0].mesh_index];
            if (mesh.n_tris != 0) b.nodes = mesh.nodes, b.tris = mesh.tris, b.inst = 0, b.ngroup_y = 0x80000000u;
        }
        else
            b.ngroup_y = 0x80000000u;
    }
    return b;
}
HL_HD void trav_begin(const TravStart& b, Trav& t, TravStack& st, bool active, f3 o, float tmin, f3 d, float tmax, uint32_t flags)
{
    t.best.t = tmax, t.best.u = 0.0f, t.best.v = 0.0f;
    t.best.instance = t.best.geometry = t.best.primitive = HL_MISS;
    st.sp = 0;
    t.ngroup.x = 0, t.ngroup.y = active ? b.ngroup_y : 0u, t.tgroup.x = 0, t.tgroup.y = 0;
    t.o = o, t.d = d, t.tmin = tmin, t.tmax = tmax, t.flags = flags;
    t.r     = make_ray_ctx(o, d);
    t.nodes = b.nodes, t.tris = b.tris, t.inst = b.inst;
}
HL_HD void trav_begin(const SceneView& s, Trav& t, TravStack& st, bool active, f3 o, float tmin, f3 d, float tmax, uint32_t flags)
{
    trav_begin(trav_start(s), t, st, active, o, tmin, d, tmax, flags);
}
HL_HD bool trav_busy(const Trav& t, const TravStack& st) { return t.ngroup.y > 0x00FFFFFFu || t.tgroup.y != 0 || st.sp > 0; }

// phase 1 of a step: nothing current -> pop one stack entry; then, if a node group is current, visit its
// nearest pending child.  (Pop and visit share a step so that a lane never sits out the other lanes' node
// test just to fetch its next entry.)
HL_HD void trav_step_nodes(const SceneView& s, Trav& t, TravStack& st)
{
    if (t.tgroup.y != 0) return;
    if (t.ngroup.y <= 0x00FFFFFFu)
    {
        if (st.sp <= 0) return;
        const u2 e = st.pop();
        if (e.y == 0)
        {
            // sentinel: leave the instance, back to world space and the top-level tree (an identity instance never left it)
            if (!instance_passthrough(s, t.inst, t.o, t.d)) t.r = make_ray_ctx(t.o, t.d);
            t.nodes = s.tlas_nodes, t.tris = nullptr, t.inst = HL_MISS;
            return;
        }
        if (e.y <= 0x00FFFFFFu)
        {
            t.tgroup = e;
            return;
        }
        t.ngroup = e;
    }
    // visit the nearest pending child of the current node group
    const uint32_t hits = t.ngroup.y;
    const int      bit  = hl_bfind(hits);
    const uint32_t base = t.ngroup.x;
    t.ngroup.y &= ~(1u << bit);
    if (t.ngroup.y > 0x00FFFFFFu) st.push(t.ngroup);
    const uint32_t slot = (uint32_t)(bit - 24) ^ t.r.octinv;
    const uint32_t rel  = (uint32_t)hl_popc((hits & 0xFFu) & ~(0xFFFFFFFFu << slot));
    HL_STAT_NODE();
    uint32_t       cb, lb, im;
    const uint32_t mask = intersect_children(t.nodes + (base + rel), t.r, t.tmin, t.best.t, cb, lb, im);
    t.ngroup.x = cb, t.ngroup.y = (mask & 0xFF000000u) | im;
    t.tgroup.x = lb, t.tgroup.y = mask & 0x00FFFFFFu;
#if defined(__CUDA_ARCH__) && HL_PREFETCH_NEXT_NODE
    // the node this lane visits next (the nearest child that was hit) is known now, one step — and possibly a triangle phase —
    // before its five loads are issued: ask for its lines
    if (t.ngroup.y > 0x00FFFFFFu)
    {
        const uint32_t h2 = t.ngroup.y;
        const uint32_t s2 = (uint32_t)(hl_bfind(h2) - 24) ^ t.r.octinv;
        const uint32_t r2 = (uint32_t)hl_popc((h2 & 0xFFu) & ~(0xFFFFFFFFu << s2));
        const char*    pn = (const char*)(t.nodes + (cb + r2));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pn));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pn + 64));
    }
#endif
}
// triangle postponing (after Ylitie et al.): a lane that has a leaf group AND inner children pending may
// put the leaf group on the stack and keep descending, so that triangle tests run when many lanes have one
HL_HD bool trav_can_postpone(const Trav& t) { return t.inst != HL_MISS && t.ngroup.y > 0x00FFFFFFu; }
HL_HD void trav_postpone(Trav& t, TravStack& st)
{
    // below the node group: the nearer inner children are visited first, then these triangles
    st.push(t.tgroup);
    t.tgroup.y = 0;
}
// phase 2 of a step: up to HL_TRI_PER_STEP triangle tests of the current leaf group (bottom level), or one
// instance entry (top level)
HL_HD void trav_step_leaves(const SceneView& s, Trav& t, TravStack& st)
{
    if (t.tgroup.y == 0) return;
    if (t.inst != HL_MISS)
    {
        bool done = false;
#if HL_TRI_PER_STEP < 24
        for (int k = 0; k < HL_TRI_PER_STEP && t.tgroup.y; k++)
#else
        while (t.tgroup.y)
#endif
        {
            const int i = hl_bfind(t.tgroup.y);
            t.tgroup.y &= ~(1u << i);
            HL_STAT_LEAF();
            if (test_leaf_triangle(s, t.tris + (t.tgroup.x + (uint32_t)i), t.r.o, t.r.d, t.tmin, t.tmax, t.inst, t.flags, t.best) && (t.flags & HL_RAY_TERMINATE))
            {
                done = true;
                break;
            }
        }
        if (done) t.ngroup.y = 0, t.tgroup.y = 0, st.sp = 0;
    }
    else
    {
        // top level: enter ONE instance; the rest of the leaf group and the node group wait on the stack
        const int i = hl_bfind(t.tgroup.y);
        t.tgroup.y &= ~(1u << i);
        HL_STAT_LEAF();
        const uint32_t  id   = s.tlas_leaf[t.tgroup.x + (uint32_t)i];
        const MeshView& mesh = s.meshes[s.instances[id].mesh_index];
        if (mesh.n_tris != 0 && !st.has_room(3))
            note_stack_overflow(); // the instance is skipped: its sentinel must never be the entry that gets dropped
        else if (mesh.n_tris != 0)
        {
            if (t.tgroup.y) st.push(t.tgroup);
            if (t.ngroup.y > 0x00FFFFFFu) st.push(t.ngroup);
            u2 sentinel;
            sentinel.x = 0xFFFFFFFFu, sentinel.y = 0;
            st.push(sentinel);
            if (instance_passthrough(s, id, t.o, t.d))
                t.r.o = t.o + mk3(0.0f); // (the transform's "+ translation" turns a -0 origin component into +0)
            else
            {
                const float* m = s.inst_inv + 12 * (size_t)id;
                const f3     o = t.o, d = t.d;
                f3           oo, od;
                oo.x = (m[0] * o.x + m[1] * o.y + m[2] * o.z) + m[3];
                oo.y = (m[4] * o.x + m[5] * o.y + m[6] * o.z) + m[7];
                oo.z = (m[8] * o.x + m[9] * o.y + m[10] * o.z) + m[11];
                od.x = m[0] * d.x + m[1] * d.y + m[2] * d.z;
                od.y = m[4] * d.x + m[5] * d.y + m[6] * d.z;
                od.z = m[8] * d.x + m[9] * d.y + m[10] * d.z;
                t.r = make_ray_ctx(oo, od);
            }
            t.nodes = mesh.nodes, t.tris = mesh.tris, t.inst = id;
            t.ngroup.x = 0, t.ngroup.y = 0x80000000u, t.tgroup.y = 0;
        }
    }
}

// ---- warp-cooperative triangle phase (GPU) ---------------------------------------------------------------------
// After a node visit a few lanes of the warp hold a handful of leaf triangles each while the others hold none: run
// lane by lane, the ~90-instruction triangle test issues at 5-6 of 32 lanes (ncu, round 1: 20 % of k_extend's warp
// instructions).  Here the warp pools its pending (ray, triangle) pairs instead — every owner lane contributes up to
// HL_COOP_K triangles of its current leaf group — and lane p of the warp tests pair p with the OWNER's ray, fetched
// through shuffles; the owner then takes the closest accepted result of its own pairs.  Same arithmetic on the same
// operands (test_leaf_candidate is the code test_leaf_triangle runs), closest hit + tie rule are order independent and
// the any-hit decision is per candidate, so the result is bit-identical to the lane-by-lane phase.
// Pair order: round j holds the j-th triangle of every owner that has one, owners in lane order; slot of (owner, j) =
// (pairs in rounds < j) + rank of the owner in round j.  Owners publish "lane | j << 5" at their slots in a small
// per-warp table in shared memory; results travel back through shuffles.
#ifndef HL_COOP_LEAVES
#define HL_COOP_LEAVES 0 /* measured (tools/tune_trace.py set "coop", round 2): images bit-identical, but 1.55 -> 1.76 ms/frame on configs[1], 6.2 -> 7.0 on configs[2], 9.2 -> 10.8 on configs[3]: the 13 shuffles + merge cost as much as the two lane-by-lane tests they replace and the kernel goes from 72 to 96 registers */
#endif
#ifndef HL_COOP_K
#define HL_COOP_K 4
#endif
#define HL_COOP_TABLE (32 * HL_COOP_K)
#if defined(__CUDA_ARCH__)
// the candidate half of test_leaf_triangle: Moeller-Trumbore + range test; no comparison with the best hit so far
__device__ __forceinline__ bool test_leaf_candidate(const LeafTri* tri, f3 o, f3 d, float tmin, float tmax, float& t, float& u, float& v, uint32_t& prim, uint32_t& geom_flags)
{
    const U4    a = load_u4((const char*)tri + 0), e1w = load_u4((const char*)tri + 16), e2w = load_u4((const char*)tri + 32);
    const f3    e1   = mk3(u2f(e1w.x), u2f(e1w.y), u2f(e1w.z));
    const f3    e2   = mk3(u2f(e2w.x), u2f(e2w.y), u2f(e2w.z));
    const f3    pvec = cross(d, e2);
    const float det  = dot(e1, pvec);
    prim = a.w, geom_flags = e1w.w;
    if (det == 0.0f || det != det) return false;
    const float inv  = 1.0f / det;
    const f3    tvec = o - mk3(u2f(a.x), u2f(a.y), u2f(a.z));
    u                = dot(tvec, pvec) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    const f3 qvec = cross(tvec, e1);
    v             = dot(d, qvec) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return false;
    t = dot(e2, qvec) * inv;
    return t > tmin && t < tmax;
}
// `mine`: this lane holds a bottom-level leaf group (all 32 lanes call; at least one has mine == true)
__device__ __forceinline__ void coop_triangles(const SceneView& s, Trav& t, TravStack& st, bool mine)
{
    const unsigned FULL = 0xFFFFFFFFu;
    unsigned       lane;
    asm("mov.u32 %0, %%laneid;" : "=r"(lane));
    const uint32_t lt = (1u << lane) - 1u;
    // up to HL_COOP_K triangles of my leaf group: their bit indices, one byte each (nearest-first order = highest bit first)
    uint32_t rem = mine ? t.tgroup.y : 0u, bits = 0u;
    int      cnt = 0;
#pragma unroll
    for (int k = 0; k < HL_COOP_K; k++)
        if (rem)
        {
            const int i = hl_bfind(rem);
            rem &= ~(1u << i), bits |= (uint32_t)i << (8 * k), cnt++;
        }
    uint32_t slot[HL_COOP_K];
    uint32_t total = 0;
#pragma unroll
    for (int k = 0; k < HL_COOP_K; k++)
    {
        const uint32_t m = __ballot_sync(FULL, cnt > k);
        slot[k]          = total + (uint32_t)__popc(m & lt);
        if (cnt > k) st.pairs[slot[k]] = (uint8_t)(lane | ((uint32_t)k << 5));
        total += (uint32_t)__popc(m);
    }
    __syncwarp();
    float    bt = hl_inf(); // closest accepted candidate among my pairs so far, and where its record sits
    uint32_t bsrc = 0, bbase = 0xFFFFFFFFu;
    bool     tie = false;
    for (uint32_t base = 0; base < total; base += 32u)
    {
        const uint32_t p   = base + lane;
        const bool     act = p < total;
        const uint32_t e = act ? (uint32_t)st.pairs[p] : 0u, owner = e & 31u, k = e >> 5;
        // the owner's query
        const uint32_t obits = __shfl_sync(FULL, bits, owner), obase = __shfl_sync(FULL, t.tgroup.x, owner), oinst = __shfl_sync(FULL, t.inst, owner), oflags = __shfl_sync(FULL, t.flags, owner);
        const unsigned long long otris = __shfl_sync(FULL, (unsigned long long)t.tris, owner);
        const f3 oo = mk3(__shfl_sync(FULL, t.r.o.x, owner), __shfl_sync(FULL, t.r.o.y, owner), __shfl_sync(FULL, t.r.o.z, owner));
        const f3 od = mk3(__shfl_sync(FULL, t.r.d.x, owner), __shfl_sync(FULL, t.r.d.y, owner), __shfl_sync(FULL, t.r.d.z, owner));
        const float otmin = __shfl_sync(FULL, t.tmin, owner), otmax = __shfl_sync(FULL, t.tmax, owner), obest = __shfl_sync(FULL, t.best.t, owner);
        float    ct = hl_inf(), cu = 0.0f, cv = 0.0f;
        uint32_t cprim = 0, cgf = 0;
        if (act)
        {
            const uint32_t i = (obits >> (8 * k)) & 0xFFu;
            HL_STAT_LEAF();
            float tt, uu, vv;
            bool  ok = test_leaf_candidate((const LeafTri*)otris + (obase + i), oo, od, otmin, otmax, tt, uu, vv, cprim, cgf);
            // a candidate behind the owner's best hit cannot win (equal t: the owner applies the tie rule)
            ok = ok && !(tt > obest);
            if (ok && !(oflags & HL_RAY_OPAQUE) && !(cgf >> 31) && any_hit_ignores(s, oinst, cgf & 0x7FFFFFFFu, cprim, uu, vv, (const LeafTri*)otris + (obase + i))) ok = false;
            if (ok) ct = tt, cu = uu, cv = vv;
        }
        // owners: the closest accepted candidate of my pairs that sit in this batch
#pragma unroll
        for (int j = 0; j < HL_COOP_K; j++)
        {
            const float tj = __shfl_sync(FULL, ct, slot[j] & 31u);
            if (j < cnt && slot[j] - base < 32u)
            {
                if (tj == bt && tj < hl_inf()) tie = true;
                if (tj < bt) bt = tj, bsrc = slot[j] & 31u, bbase = base;
            }
        }
        // fetch the winner's record while this batch's registers are live (lanes whose winner sits in an earlier batch keep theirs)
        const bool     fresh = bbase == base;
        const float    wu = __shfl_sync(FULL, cu, bsrc), wv = __shfl_sync(FULL, cv, bsrc);
        const uint32_t wp = __shfl_sync(FULL, cprim, bsrc), wg = __shfl_sync(FULL, cgf, bsrc) & 0x7FFFFFFFu;
        if (__any_sync(FULL, tie))
        {
            // two of an owner's candidates at the same distance (coincident triangles): resolve by the tie rule over all of them
#pragma unroll
            for (int j = 0; j < HL_COOP_K; j++)
            {
                const uint32_t src = slot[j] & 31u;
                const float    tj = __shfl_sync(FULL, ct, src), uj = __shfl_sync(FULL, cu, src), vj = __shfl_sync(FULL, cv, src);
                const uint32_t pj = __shfl_sync(FULL, cprim, src), gj = __shfl_sync(FULL, cgf, src) & 0x7FFFFFFFu;
                if (tie && j < cnt && slot[j] - base < 32u && tj < hl_inf())
                {
                    Hit& b = t.best;
                    bool take = tj < b.t;
                    if (!take && tj == b.t) take = t.inst != b.instance ? t.inst < b.instance : (gj != b.geometry ? gj < b.geometry : pj < b.primitive);
                    if (take) b.t = tj, b.u = uj, b.v = vj, b.instance = t.inst, b.geometry = gj, b.primitive = pj;
                }
            }
        }
        if (fresh && !tie)
        {
            Hit& b = t.best;
            bool take = bt < b.t;
            if (!take && bt == b.t) take = t.inst != b.instance ? t.inst < b.instance : (wg != b.geometry ? wg < b.geometry : wp < b.primitive);
            if (take) b.t = bt, b.u = wu, b.v = wv, b.instance = t.inst, b.geometry = wg, b.primitive = wp;
        }
        if (tie) bt = t.best.t, tie = false, bbase = 0xFFFFFFFFu; // (the slow path folded everything into best)
    }
    if (mine)
    {
        t.tgroup.y = rem;
        if ((t.flags & HL_RAY_TERMINATE) && bt < hl_inf()) t.ngroup.y = 0, t.tgroup.y = 0, st.sp = 0; // gl_RayFlagsTerminateOnFirstHitEXT
    }
    __syncwarp(); // the table is rewritten by the next call
}
#endif

// One step of every lane of the warp; `busy_mask` = ballot of the lanes that hold an unfinished query (all
// 32 lanes must call this).  The leaf phase runs when at least HL_TRI_MIN_LANES lanes (or every busy lane)
// have a leaf group; otherwise lanes that can postpone do so and the others keep theirs for the next step
// (their count only grows, so the phase is eventually taken).  Scheduling only: which step tests which
// triangle never changes the result (closest hit + tie rule are order independent).
#ifndef HL_TRI_MIN_LANES
#define HL_TRI_MIN_LANES 1
#endif
HL_HD void trav_step_warp(const SceneView& s, Trav& t, TravStack& st, bool busy, uint32_t busy_mask)
{
    if (busy) trav_step_nodes(s, t, st);
    const bool want = busy && t.tgroup.y != 0;
#if defined(__CUDA_ARCH__) && HL_COOP_LEAVES
    (void)busy_mask;
    // top level: one instance entry per lane and step, as before; bottom level: the warp's triangles pooled
    const bool top = want && t.inst == HL_MISS;
    if (top) trav_step_leaves(s, t, st);
    const bool mine = want && !top;
    if (__ballot_sync(0xFFFFFFFFu, mine)) coop_triangles(s, t, st, mine);
#else
#if defined(__CUDA_ARCH__)
    const uint32_t wm = __ballot_sync(0xFFFFFFFFu, want);
    if (wm == 0) return;
    const int  need = min((int)HL_TRI_MIN_LANES, __popc(busy_mask));
    const bool run  = __popc(wm) >= need;
#elif defined(HL_TRAVERSAL_STATS)
    const bool run = !emul_force_postpone(); // emulator: can exercise the postponing path on every opportunity
    (void)busy_mask;
#else
    const bool run = true;
    (void)busy_mask;
#endif
    if (!want) return;
    if (run)
        trav_step_leaves(s, t, st);
    else if (trav_can_postpone(t))
        trav_postpone(t, st);
#if !defined(__CUDA_ARCH__)
    else
        trav_step_leaves(s, t, st); // a single host lane cannot wait for company
#endif
#endif
}

// traceRayEXT, one query per lane to completion: fills `best` (instance == HL_MISS when nothing was hit).
// `active` = false runs an empty query (GPU lanes without a ray still take part in the warp votes: the loop
// condition is a warp vote, so the 32 rays of a warp re-converge at the top of every step).
HL_HD void trace_ray(const SceneView& s, bool active, f3 o, float tmin, f3 d, float tmax, uint32_t flags, Hit& best, TravStack& st)
{
    Trav t;
    trav_begin(s, t, st, active, o, tmin, d, tmax, flags);
    for (;;)
    {
        const bool     busy = trav_busy(t, st);
        const uint32_t bm   = HL_WARP_BALLOT(busy);
        if (bm == 0) break;
        trav_step_warp(s, t, st, busy, bm);
    }
    best = t.best;
}
} // namespace hl
