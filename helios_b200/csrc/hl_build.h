// hl_build.h — per-element logic of the GPU BVH builder that replaces vkCmdBuildAccelerationStructuresKHR
// (reference call sites: src/engine/gfx/vk.cpp:3207-3226 for the per-mesh BLAS, src/engine/gfx/renderer.cpp:147-168
// for the TLAS).  Pipeline: primitive boxes -> 63-bit Morton keys -> radix sort -> binary radix tree (Karras 2012)
// -> bottom-up box fit + SAH cost tables -> SAH-optimal collapse (dynamic programme) into 8-wide nodes with
// octant-ordered child slots and 8-bit quantised child boxes (Ylitie et al. 2017).  Each function handles ONE element so that the CUDA kernels are
// thin loops over thread ids (and the tests/emul harness can run the same logic sequentially).
#pragma once
#include "hl_scene.h"

namespace hl
{
#if defined(__CUDA_ARCH__)
HL_HD uint32_t hl_atomic_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
#else
inline uint32_t hl_atomic_add(uint32_t* p, uint32_t v)
{
    uint32_t o = *p;
    *p += v;
    return o;
}
#endif

struct Box
{
    float lo[3], hi[3];
};

// ---- Morton keys ----------------------------------------------------------------------------------
HL_HD uint64_t spread21(uint64_t v)
{
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x1F00000000FFFFull;
    v = (v | (v << 16)) & 0x1F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
// centroid of `b` relative to the scene box -> 21 bits per axis, x in the most significant position
HL_HD uint64_t morton_key(const Box& b, const Box& scene)
{
    uint64_t q[3];
    for (int a = 0; a < 3; a++)
    {
        const float ext = scene.hi[a] - scene.lo[a];
        const float c   = 0.5f * (b.lo[a] + b.hi[a]);
        float       t   = ext > 0.0f ? (c - scene.lo[a]) / ext : 0.0f;
        t               = fminf(fmaxf(t, 0.0f), 1.0f);
        uint32_t v      = (uint32_t)(t * 2097151.0f);
        q[a]            = v > 2097151u ? 2097151u : v;
    }
    return (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
}

// ---- binary radix tree (Karras, HPG 2012) ----------------------------------------------------------
// n leaves, n-1 internal nodes.  Node ids: internal i in [0, n-1), leaf j is id (n-1)+j.
struct BinaryTree
{
    uint32_t  n;      // leaves
    uint32_t* left;   // [n-1]
    uint32_t* right;  // [n-1]
    uint32_t* first;  // [n-1] covered leaf range
    uint32_t* last;   // [n-1]
    uint32_t* parent; // [2n-1]
    Box*      box;    // [2n-1]
    uint32_t* visits; // [n-1] arrival counters for the bottom-up pass
    float*    cost;   // [(2n-1) * 7] SAH cost table of the collapse DP (see sah_node_costs)
    uint32_t* dec;    // [n-1] packed decisions of the collapse DP
    float     c_prim; // SAH cost of one leaf primitive test relative to one wide-node visit
};
HL_HD int key_delta(const uint64_t* keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + (31 - hl_bfind((uint32_t)(i ^ j))); // tie-break on the index: 64 + clz32(i ^ j)
    return hl_clz64(a ^ b);
}
HL_HD void radix_tree_node(const uint64_t* keys, BinaryTree& t, int i)
{
    const int n  = (int)t.n;
    const int d  = key_delta(keys, n, i, i + 1) > key_delta(keys, n, i, i - 1) ? 1 : -1;
    const int dmin = key_delta(keys, n, i, i - d);
    int       lmax = 2;
    while (key_delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int s = lmax / 2; s >= 1; s /= 2)
        if (key_delta(keys, n, i, i + (l + s) * d) > dmin) l += s;
    const int j     = i + l * d;
    const int dnode = key_delta(keys, n, i, j);
    int       split = 0;
    int       step  = l;
    do
    {
        step = (step + 1) >> 1;
        if (key_delta(keys, n, i, i + (split + step) * d) > dnode) split += step;
    } while (step > 1);
    const int gamma = i + split * d + (d < 0 ? -1 : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const uint32_t L = (lo == gamma) ? (uint32_t)(n - 1 + gamma) : (uint32_t)gamma;
    const uint32_t R = (hi == gamma + 1) ? (uint32_t)(n - 1 + gamma + 1) : (uint32_t)(gamma + 1);
    t.left[i] = L, t.right[i] = R, t.first[i] = (uint32_t)lo, t.last[i] = (uint32_t)hi;
    t.parent[L] = (uint32_t)i, t.parent[R] = (uint32_t)i;
    if (i == 0) t.parent[0] = 0xFFFFFFFFu;
}
// boxes written by other SMs during the bottom-up fit must not be served from a stale L1 line
HL_HD Box load_box_coherent(const Box* p)
{
#if defined(__CUDA_ARCH__)
    Box          r;
    const float* f = (const float*)p;
    r.lo[0] = __ldcg(f + 0), r.lo[1] = __ldcg(f + 1), r.lo[2] = __ldcg(f + 2);
    r.hi[0] = __ldcg(f + 3), r.hi[1] = __ldcg(f + 4), r.hi[2] = __ldcg(f + 5);
    return r;
#else
    return *p;
#endif
}
HL_HD Box box_union(const Box& a, const Box& b)
{
    Box r;
    for (int k = 0; k < 3; k++) r.lo[k] = fminf(a.lo[k], b.lo[k]), r.hi[k] = fmaxf(a.hi[k], b.hi[k]);
    return r;
}
HL_HD float box_half_area(const Box& b)
{
    const float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
    return dx * dy + dy * dz + dz * dx;
}
#define HL_MAX_LEAF_PRIMS 3u
HL_HD uint32_t subtree_prims(const BinaryTree& t, uint32_t node)
{
    return node >= t.n - 1 ? 1u : t.last[node] - t.first[node] + 1u;
}
HL_HD uint32_t subtree_first(const BinaryTree& t, uint32_t node) { return node >= t.n - 1 ? node - (t.n - 1) : t.first[node]; }

// ---- SAH-optimal collapse, bottom-up half (dynamic programme of Ylitie, Karras, Laine, HPG 2017 §4, written
// from the paper's recurrences).  C(n, i), i = 1..7 = cheapest SAH cost of representing binary subtree n with
// at most i child slots of one wide node:
//   C(n, 1) = min( leaf:  A_n * P_n * c_prim           (P_n <= 3 primitives),
//                  inner: A_n * c_node + D(n, 8) )      n becomes a wide node of its own
//   C(n, i) = min( D(n, i), C(n, i-1) )                 i = 2..7
//   D(n, j) = min over 0 < k < j of C(left, k) + C(right, j - k)
// dec[n] packs what achieved each minimum, 4 bits per i (0 = leaf, 15 = inner, k = split giving the left
// child k slots) and the k of D(n, 8) in bits 28..30; collapse_one() replays it top-down.
#define HL_SAH_C_NODE 1.0f
#ifndef HL_SAH_C_PRIM_TRIANGLE
#define HL_SAH_C_PRIM_TRIANGLE 0.35f /* measured: ~90 SASS instructions per triangle test vs ~250 per node visit */
#endif
#define HL_SAH_C_PRIM_INSTANCE 16.0f /* an instance entry costs a whole bottom-level traversal */
HL_HD float load_f32_coherent(const float* p)
{
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}
HL_HD void sah_leaf_costs(BinaryTree& t, uint32_t node, float area)
{
    for (int i = 0; i < 7; i++) t.cost[(size_t)node * 7 + i] = area * t.c_prim;
}
HL_HD void sah_node_costs(BinaryTree& t, uint32_t node, float area)
{
    float cl[8], cr[8];
    const float inf = hl_inf();
    cl[0] = cr[0] = inf;
    for (int i = 0; i < 7; i++) cl[i + 1] = load_f32_coherent(t.cost + (size_t)t.left[node] * 7 + i), cr[i + 1] = load_f32_coherent(t.cost + (size_t)t.right[node] * 7 + i);
    float    D[9];
    uint32_t K[9];
    for (int j = 2; j <= 8; j++)
    {
        float    best = inf;
        uint32_t kb   = 1;
        for (int k = (j - 7 > 1 ? j - 7 : 1); k < j && k <= 7; k++)
        {
            const float c = cl[k] + cr[j - k];
            if (c < best) best = c, kb = (uint32_t)k;
        }
        D[j] = best, K[j] = kb;
    }
    const uint32_t P       = subtree_prims(t, node);
    const float    c_leaf  = P <= HL_MAX_LEAF_PRIMS ? area * (float)P * t.c_prim : inf;
    const float    c_inner = area * HL_SAH_C_NODE + D[8];
    float          c       = c_leaf <= c_inner ? c_leaf : c_inner;
    uint32_t       d       = c_leaf <= c_inner ? 0u : 15u;
    uint32_t       packed  = d | (K[8] << 28);
    t.cost[(size_t)node * 7] = c;
    for (int i = 2; i <= 7; i++)
    {
        if (D[i] < c) c = D[i], d = K[i];
        packed |= d << (4 * (i - 1));
        t.cost[(size_t)node * 7 + (i - 1)] = c;
    }
    t.dec[node] = packed;
}

// bottom-up fit: called once per leaf (after its box is written); the second arrival at a node continues.
// On the GPU the caller issues __threadfence() between the writes and the counter increment.
template <class Fence>
HL_HD void fit_from_leaf(BinaryTree& t, uint32_t leaf, Fence fence)
{
    sah_leaf_costs(t, (t.n - 1) + leaf, box_half_area(t.box[(t.n - 1) + leaf]));
    uint32_t node = t.parent[(t.n - 1) + leaf];
    while (node != 0xFFFFFFFFu)
    {
        fence();
        if (hl_atomic_add(&t.visits[node], 1u) == 0u) return;
        fence();
        const Box b = box_union(load_box_coherent(&t.box[t.left[node]]), load_box_coherent(&t.box[t.right[node]]));
        t.box[node] = b;
        sah_node_costs(t, node, box_half_area(b));
        node = t.parent[node];
    }
}

// ---- collapse to 8-wide, top-down half -------------------------------------------------------------

// smallest biased exponent e with 255 * 2^(e-127) >= extent (0 for a flat axis)
HL_HD uint32_t quant_exponent(float extent)
{
    if (!(extent > 0.0f)) return 0u;
    const float    s    = extent / 255.0f;
    const uint32_t bits = f2u(s);
    uint32_t       e    = (bits >> 23) & 0xFFu;
    if (bits & 0x7FFFFFu) e++;
    if (e == 0) e = 1;
    while (e < 254u && !(255.0f * u2f(e << 23) >= extent)) e++;
    return e;
}
// conservative 8-bit grid coordinates of [clo, chi] relative to origin p with cell size 2^(e-127)
HL_HD void quantize_axis(float p, uint32_t e, float clo, float chi, uint8_t& qlo, uint8_t& qhi)
{
    if (e == 0)
    {
        qlo = 0, qhi = 0;
        return;
    }
    const float s  = u2f(e << 23);
    float       fl = floorf((clo - p) / s), fh = ceilf((chi - p) / s);
    fl = fminf(fmaxf(fl, 0.0f), 255.0f), fh = fminf(fmaxf(fh, 0.0f), 255.0f);
    int l = (int)fl, h = (int)fh;
    while (l > 0 && p + (float)l * s > clo) l--;
    while (h < 255 && p + (float)h * s < chi) h++;
    qlo = (uint8_t)l, qhi = (uint8_t)h;
}

struct CollapseTask
{
    uint32_t wide, bnode;
};
struct WideOut
{
    WideNode* nodes;
    uint32_t* node_counter; // next free wide node
    uint32_t* leaf_counter; // next free leaf slot
};
// A task is published to the work queue with ONE aligned 8-byte store, so a consumer polling the slot sees
// either the empty pattern (all ones) or the whole task.
HL_HD void publish_task(CollapseTask* slot, CollapseTask t)
{
#if defined(__CUDA_ARCH__)
    *(volatile unsigned long long*)slot = ((unsigned long long)t.bnode << 32) | t.wide;
#else
    *slot = t;
#endif
}
// Builds wide node `task.wide` from binary subtree `task.bnode`; appends one task per internal child to the
// queue `next` (slots handed out by `next_count`) and returns how many it appended.
// leaf_writer(dst_leaf_index, sorted_leaf_position) stores one leaf primitive record.
template <class LeafWriter>
HL_HD uint32_t collapse_one(const BinaryTree& t, CollapseTask task, WideOut out, CollapseTask* next, uint32_t* next_count, LeafWriter& leaf_writer)
{
    uint32_t ch[8];
    uint32_t ch_inner = 0; // bit k: child k becomes a wide node of its own
    int      nch = 0;
    const uint32_t leaf0 = t.n - 1;
    if (task.bnode >= leaf0 || subtree_prims(t, task.bnode) <= HL_MAX_LEAF_PRIMS)
        ch[nch++] = task.bnode; // tiny tree: the root's only child is one leaf
    else
    {
        // replay the DP: (node, slots) pairs; at most 8 are ever pending
        uint32_t sn[8], ss[8];
        int      sp = 0;
        const uint32_t k8 = (t.dec[task.bnode] >> 28) & 7u;
        sn[sp] = t.right[task.bnode], ss[sp++] = 8u - k8;
        sn[sp] = t.left[task.bnode], ss[sp++] = k8;
        while (sp > 0 && nch < 8)
        {
            sp--;
            const uint32_t m = sn[sp], i = ss[sp];
            uint32_t       d = 0;
            if (m < leaf0) d = (t.dec[m] >> (4 * ((i < 1 ? 1 : i) - 1))) & 15u;
            if (m >= leaf0 || d == 0u)
                ch[nch++] = m;
            else if (d == 15u)
                ch_inner |= 1u << nch, ch[nch++] = m;
            else
            {
                sn[sp] = t.right[m], ss[sp++] = i - d;
                sn[sp] = t.left[m], ss[sp++] = d;
            }
        }
    }
    const Box nb = t.box[task.bnode];
    // octant-ordered slot assignment: greedy minimum of cost(child, slot) = (centroid_c - centroid_node) . dir(slot)
    float cx[8], cy[8], cz[8];
    for (int k = 0; k < nch; k++)
    {
        const Box& b = t.box[ch[k]];
        cx[k] = (b.lo[0] + b.hi[0]) - (nb.lo[0] + nb.hi[0]);
        cy[k] = (b.lo[1] + b.hi[1]) - (nb.lo[1] + nb.hi[1]);
        cz[k] = (b.lo[2] + b.hi[2]) - (nb.lo[2] + nb.hi[2]);
    }
    int      slot_child[8];
    uint32_t child_done = 0, slot_done = 0;
    for (int s = 0; s < 8; s++) slot_child[s] = -1;
    // cost(child, slot) once, then 8 rounds of "cheapest remaining pair"
    float cost[64];
    for (int k = 0; k < nch; k++)
        for (int s = 0; s < 8; s++) cost[k * 8 + s] = ((s & 4) ? -cx[k] : cx[k]) + ((s & 2) ? -cy[k] : cy[k]) + ((s & 1) ? -cz[k] : cz[k]);
    for (int it = 0; it < nch; it++)
    {
        float bc = 3.0e38f;
        int   bk = -1, bs = -1;
        for (int k = 0; k < nch; k++)
        {
            if (child_done & (1u << k)) continue;
            for (int s = 0; s < 8; s++)
            {
                if (slot_done & (1u << s)) continue;
                const float c = cost[k * 8 + s];
                if (c < bc) bc = c, bk = k, bs = s;
            }
        }
        if (bk < 0)
        { // NaN boxes: fall back to the first free pair
            for (int k = 0; k < nch && bk < 0; k++)
                if (!(child_done & (1u << k))) bk = k;
            for (int s = 0; s < 8 && bs < 0; s++)
                if (!(slot_done & (1u << s))) bs = s;
        }
        slot_child[bs] = bk;
        child_done |= 1u << bk, slot_done |= 1u << bs;
    }
    WideNode w;
    w.px = nb.lo[0], w.py = nb.lo[1], w.pz = nb.lo[2];
    const uint32_t ex = quant_exponent(nb.hi[0] - nb.lo[0]), ey = quant_exponent(nb.hi[1] - nb.lo[1]), ez = quant_exponent(nb.hi[2] - nb.lo[2]);
    w.ex = (uint8_t)ex, w.ey = (uint8_t)ey, w.ez = (uint8_t)ez;
    uint32_t n_inner = 0, n_leafprims = 0, imask = 0;
    for (int s = 0; s < 8; s++)
    {
        if (slot_child[s] < 0) continue;
        const uint32_t c = ch[slot_child[s]];
        if (ch_inner & (1u << slot_child[s]))
            n_inner++, imask |= 1u << s;
        else
            n_leafprims += subtree_prims(t, c);
    }
    w.imask      = (uint8_t)imask;
    w.child_base = n_inner ? hl_atomic_add(out.node_counter, n_inner) : 0u;
    w.leaf_base  = n_leafprims ? hl_atomic_add(out.leaf_counter, n_leafprims) : 0u;
    uint32_t inner_rank = 0, leaf_off = 0;
    uint32_t first_task = n_inner ? hl_atomic_add(next_count, n_inner) : 0u;
    for (int s = 0; s < 8; s++)
    {
        if (slot_child[s] < 0)
        {
            w.meta[s] = 0;
            w.qlox[s] = w.qloy[s] = w.qloz[s] = 255;
            w.qhix[s] = w.qhiy[s] = w.qhiz[s] = 0;
            continue;
        }
        const uint32_t c = ch[slot_child[s]];
        const Box&     b = t.box[c];
        quantize_axis(w.px, ex, b.lo[0], b.hi[0], w.qlox[s], w.qhix[s]);
        quantize_axis(w.py, ey, b.lo[1], b.hi[1], w.qloy[s], w.qhiy[s]);
        quantize_axis(w.pz, ez, b.lo[2], b.hi[2], w.qloz[s], w.qhiz[s]);
        if (imask & (1u << s))
        {
            w.meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
            CollapseTask nt;
            nt.wide = w.child_base + inner_rank, nt.bnode = c;
            publish_task(next + (first_task + inner_rank), nt);
            inner_rank++;
        }
        else
        {
            const uint32_t cnt = subtree_prims(t, c), f = subtree_first(t, c);
            w.meta[s]          = (uint8_t)((((1u << cnt) - 1u) << 5) | leaf_off);
            for (uint32_t k = 0; k < cnt; k++) leaf_writer(w.leaf_base + leaf_off + k, f + k);
            leaf_off += cnt;
        }
    }
    out.nodes[task.wide] = w;
    return n_inner;
}

// geometry lookup for a flat triangle index: tri_start[g] <= f < tri_start[g+1]
HL_HD uint32_t find_geometry(const uint32_t* tri_start, uint32_t n_geom, uint32_t f)
{
    uint32_t lo = 0, hi = n_geom;
    while (hi - lo > 1)
    {
        const uint32_t mid = (lo + hi) / 2;
        if (tri_start[mid] <= f)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}
struct TriLeafWriter
{
    const hl_vertex*  vertices;
    const uint32_t*   indices;
    const hl_submesh* submeshes;
    const uint32_t*   tri_start; // [n_geom + 1] prefix sum of triangle counts
    uint32_t          n_geom;
    const uint32_t*   sorted_prim; // flat triangle index per sorted leaf position
    LeafTri*          tris;
    HL_HD void        operator()(uint32_t dst, uint32_t sorted_pos) const
    {
        const uint32_t f = sorted_prim[sorted_pos];
        const uint32_t g = find_geometry(tri_start, n_geom, f);
        const uint32_t p = f - tri_start[g];
        const size_t   b = (size_t)submeshes[g].base_index + 3 * (size_t)p;
        const float*   a = vertices[indices[b + 0]].position;
        const float*   c1 = vertices[indices[b + 1]].position;
        const float*   c2 = vertices[indices[b + 2]].position;
        LeafTri        r;
        r.p0x = a[0], r.p0y = a[1], r.p0z = a[2];
        r.e1x = c1[0] - a[0], r.e1y = c1[1] - a[1], r.e1z = c1[2] - a[2];
        r.e2x = c2[0] - a[0], r.e2y = c2[1] - a[1], r.e2z = c2[2] - a[2];
        r.prim       = p;
        r.geom_flags = g | (submeshes[g].opaque ? 0x80000000u : 0u);
        r.pad        = 0;
        tris[dst]    = r;
    }
};
HL_HD Box triangle_box(const hl_vertex* vertices, const uint32_t* indices, const hl_submesh* submeshes, const uint32_t* tri_start, uint32_t n_geom, uint32_t f)
{
    const uint32_t g = find_geometry(tri_start, n_geom, f);
    const size_t   b = (size_t)submeshes[g].base_index + 3 * (size_t)(f - tri_start[g]);
    const float*   p0 = vertices[indices[b + 0]].position;
    const float*   p1 = vertices[indices[b + 1]].position;
    const float*   p2 = vertices[indices[b + 2]].position;
    Box            r;
    for (int k = 0; k < 3; k++) r.lo[k] = fminf(p0[k], fminf(p1[k], p2[k])), r.hi[k] = fmaxf(p0[k], fmaxf(p1[k], p2[k]));
    return r;
}
// collapse only records which sorted primitive lands in which leaf slot; a second, fully parallel pass writes
// the records (keeps the vertex fetches out of the collapse's parent -> child dependency chain)
struct DeferredLeafWriter
{
    uint32_t*  leaf_pos;
    HL_HD void operator()(uint32_t dst, uint32_t sorted_pos) const { leaf_pos[dst] = sorted_pos; }
};
struct InstLeafWriter
{
    const uint32_t* sorted_prim;
    uint32_t*       leaf;
    HL_HD void      operator()(uint32_t dst, uint32_t sorted_pos) const { leaf[dst] = sorted_prim[sorted_pos]; }
};
} // namespace hl
