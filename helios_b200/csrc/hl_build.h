// hl_build.h — per-element logic of the GPU BVH builder that replaces vkCmdBuildAccelerationStructuresKHR
// (reference call sites: src/engine/gfx/vk.cpp:3207-3226 for the per-mesh BLAS, src/engine/gfx/renderer.cpp:147-168
// for the TLAS).  Pipeline: primitive boxes -> 63-bit Morton keys -> radix sort -> binary radix tree (Karras 2012)
// -> bottom-up box fit + SAH cost tables -> SAH-optimal collapse (dynamic programme) into 8-wide nodes with
// octant-ordered child slots and 8-bit quantised child boxes (Ylitie et al. 2017).  Each function handles ONE element so that the CUDA kernels are
// thin loops over thread ids (and the tests/emul harness can run the same logic sequentially).
#pragma once
#include "hl_scene.h"
#if defined(__CUDACC__)
#include <cooperative_groups.h>
#include <cooperative_groups/scan.h>
#endif

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
// the recurrences for one node from its children's rows cl[1..7], cr[1..7] ([0] unused); P = primitives below the node.
// Writes the node's row and decisions and returns the row in out[0..6].
HL_HD void sah_node_costs_rows(BinaryTree& t, uint32_t node, float area, uint32_t P, const float* cl, const float* cr, float* out)
{
    const float inf = hl_inf();
    float       D[9];
    uint32_t    K[9];
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
    const float    c_leaf  = P <= HL_MAX_LEAF_PRIMS ? area * (float)P * t.c_prim : inf;
    const float    c_inner = area * HL_SAH_C_NODE + D[8];
    // (the decision must be valid whatever the costs are: with coordinates around 1e20 the areas overflow to +inf, and
    //  inf <= inf would make a leaf of a subtree that holds more than HL_MAX_LEAF_PRIMS primitives)
    const bool     leaf    = P <= HL_MAX_LEAF_PRIMS && c_leaf <= c_inner;
    float          c       = leaf ? c_leaf : c_inner;
    uint32_t       d       = leaf ? 0u : 15u;
    uint32_t       packed  = d | (K[8] << 28);
    t.cost[(size_t)node * 7] = c, out[0] = c;
    for (int i = 2; i <= 7; i++)
    {
        if (D[i] < c) c = D[i], d = K[i];
        packed |= d << (4 * (i - 1));
        t.cost[(size_t)node * 7 + (i - 1)] = c, out[i - 1] = c;
    }
    t.dec[node] = packed;
}
HL_HD void sah_node_costs(BinaryTree& t, uint32_t node, float area)
{
    float cl[8], cr[8], out[7];
    cl[0] = cr[0] = hl_inf();
    for (int i = 0; i < 7; i++) cl[i + 1] = load_f32_coherent(t.cost + (size_t)t.left[node] * 7 + i), cr[i + 1] = load_f32_coherent(t.cost + (size_t)t.right[node] * 7 + i);
    sah_node_costs_rows(t, node, area, subtree_prims(t, node), cl, cr, out);
}

// bottom-up fit: called once per leaf (after its box is written); the second arrival at a node continues.
// On the GPU the caller issues __threadfence() between the writes and the counter increment.
// `stop_above` != 0: nodes that cover more than stop_above primitives are left alone — when the upper levels are going to be
// re-split (top_* below), everything above the cut is re-linked and re-fitted by top_refit_from_cluster anyway, so the
// first fit only has to produce the boxes and cost tables of the cluster subtrees.
template <class Fence>
HL_HD void fit_from_leaf(BinaryTree& t, uint32_t leaf, Fence fence, uint32_t stop_above = 0u)
{
    sah_leaf_costs(t, (t.n - 1) + leaf, box_half_area(t.box[(t.n - 1) + leaf]));
    uint32_t node = t.parent[(t.n - 1) + leaf];
    while (node != 0xFFFFFFFFu)
    {
        if (stop_above && subtree_prims(t, node) > stop_above) return;
        fence();
        if (hl_atomic_add(&t.visits[node], 1u) == 0u) return;
        fence();
        const Box b = box_union(load_box_coherent(&t.box[t.left[node]]), load_box_coherent(&t.box[t.right[node]]));
        t.box[node] = b;
        sah_node_costs(t, node, box_half_area(b));
        node = t.parent[node];
    }
}

// ---- binned-SAH re-split of the upper levels ------------------------------------------------------------
// The LBVH above a cut is rebuilt top-down with the binned surface-area heuristic (16 bins per axis; the
// HLBVH scheme of Garanzha, Pantaleoni, McAllister, HPG 2011, written from the paper's description):
//   * clusters = the maximal radix-tree subtrees with <= C primitives (they keep their Morton-built interior);
//   * the K-1 internal nodes above the cut are re-linked, level by level: every level runs three data-parallel
//     phases — BIN (each cluster adds its box to the 3 x 16 bins of the node it currently sits in), SPLIT (each
//     node sweeps its bins for the cheapest area(L)·prims(L) + area(R)·prims(R) plane, CHOOSE; then takes a binary
//     node id from the free list, links itself to its parent and creates its two children in the next level's
//     array, COMMIT) and ASSIGN (each cluster moves to the child on its side of the plane and grows that child's centroid
//     bounds).  Nothing is physically partitioned: a cluster only carries the index of its current node.
//   * a node with <= HL_TOP_SMALL clusters leaves the level loop: its clusters register in a per-node list during
//     the next BIN phase, and after the last level one thread per such node builds the whole subtree with exact
//     (sorted sweep) SAH splits — no bins, and the long tail of nearly-empty levels disappears;
//   * nodes without a usable plane (all centroids equal), nodes that found no free bin slot and levels beyond
//     HL_TOP_SAH_LEVELS split by arrival order (first half left), which bounds the number of levels;
//   * a node holding one cluster is not materialised: the cluster links itself to the parent.
// Afterwards top_refit_from_cluster() recomputes boxes, primitive counts and the collapse cost tables of the
// re-linked nodes bottom-up.  Tree shape never changes a traversal RESULT (closest hit + tie rule are order
// independent), so the arrival-order splits may differ from run to run.
#define HL_TOP_BINS 16
#ifndef HL_DEFAULT_SAH_CLUSTER
#define HL_DEFAULT_SAH_CLUSTER 2 /* primitives per cluster below the cut (triangle trees; instance trees use 1) */
#endif
#define HL_TOP_MAX_LEVELS 160
#define HL_TOP_SAH_LEVELS 120
#define HL_TOP_DONE 0xFFFFFFFFu
#define HL_TOP_MODE_BINNED 0u
#define HL_TOP_MODE_ARRIVAL 1u
#define HL_TOP_MODE_SINGLE 2u
#define HL_TOP_MODE_SMALL 3u
#define HL_TOP_MODE_TREELET 4u
#ifndef HL_TOP_SMALL
#define HL_TOP_SMALL 8u /* nodes with at most this many clusters are finished by one thread each with exact SAH sweeps (16: that kernel alone took 0.8 of 5.3 ms at 1M triangles, one level fewer in the level loop) */
#endif
#define HL_ORD_POS_INF 0xFF800000u /* f2ord(+inf) */
#define HL_ORD_NEG_INF 0x007FFFFFu /* f2ord(-inf) */

// order-preserving float <-> uint32 map, so that float min / max can be done with integer atomics
HL_HD uint32_t f2ord(float f)
{
    const uint32_t u = f2u(f);
    return (u >> 31) ? ~u : (u | 0x80000000u);
}
HL_HD float ord2f(uint32_t o) { return u2f((o >> 31) ? (o & 0x7FFFFFFFu) : ~o); }

#if defined(__CUDA_ARCH__)
HL_HD uint32_t hl_atomic_min(uint32_t* p, uint32_t v) { return atomicMin(p, v); }
HL_HD uint32_t hl_atomic_max(uint32_t* p, uint32_t v) { return atomicMax(p, v); }
HL_HD uint32_t hl_load_cg(const uint32_t* p) { return __ldcg(p); }
#else
inline uint32_t hl_atomic_min(uint32_t* p, uint32_t v)
{
    const uint32_t o = *p;
    if (v < o) *p = v;
    return o;
}
inline uint32_t hl_atomic_max(uint32_t* p, uint32_t v)
{
    const uint32_t o = *p;
    if (v > o) *p = v;
    return o;
}
inline uint32_t hl_load_cg(const uint32_t* p) { return *p; }
#endif

// counter += amount, returns the old value.  The SPLIT phase hands out node ids, child slots, bin slots and
// small-node records through a few global counters; on the GPU the lanes of a warp that allocate from the same
// counter together (same amount) are served by ONE atomic.
HL_HD uint32_t hl_alloc(uint32_t* counter, uint32_t amount)
{
#if defined(__CUDA_ARCH__)
    const unsigned act = __activemask();
    int            same_c, same_a;
    __match_all_sync(act, (unsigned long long)counter, &same_c);
    __match_all_sync(act, amount, &same_a);
    if (same_c && same_a)
    {
        unsigned lane;
        asm("mov.u32 %0, %%laneid;" : "=r"(lane));
        const int leader = __ffs(act) - 1;
        uint32_t  base   = 0;
        if ((int)lane == leader) base = atomicAdd(counter, amount * (uint32_t)__popc(act));
        base = __shfl_sync(act, base, leader);
        return base + amount * (uint32_t)__popc(act & ((1u << lane) - 1u));
    }
#endif
    return hl_atomic_add(counter, amount);
}

// counter += amount (amounts differ per lane, 0 allowed), returns the old value.  On the GPU the lanes that arrive here
// together are served by ONE atomic (prefix sum over the coalesced group): the collapse hands out wide-node, leaf and queue
// slots from three global counters, and one atomic per lane on the same address is what bounded that kernel.
HL_HD uint32_t hl_alloc_var(uint32_t* counter, uint32_t amount)
{
#if defined(__CUDA_ARCH__)
    namespace cg = cooperative_groups;
    const cg::coalesced_group g    = cg::coalesced_threads();
    const uint32_t            incl = cg::inclusive_scan(g, amount);
    uint32_t                  base = 0;
    if (g.thread_rank() == g.size() - 1 && incl) base = atomicAdd(counter, incl);
    base = g.shfl(base, g.size() - 1);
    return base + incl - amount;
#else
    return hl_atomic_add(counter, amount);
#endif
}

struct TopBin
{
    uint32_t lo[3], hi[3]; // union of the cluster boxes, f2ord encoding
    uint32_t prims, clusters;
}; // 32 B
struct TopNode
{
    uint32_t cb_lo[3], cb_hi[3]; // bounds of the cluster centroids (f2ord), grown by the ASSIGN phase of the level above
    uint32_t count;              // clusters in the node
    uint32_t link;               // (binary node id of the parent << 1) | side; 0xFFFFFFFF = root
    uint32_t mode;
    uint32_t bins;     // BINNED: bin slot of this node in the level's bin array; SMALL: index of its TopSmall record
    uint32_t split;    // BINNED, after SPLIT: axis | (last bin of the left side << 2)
    uint32_t n_left;   // clusters that go left
    uint32_t child;    // index of the left child in the next level's node array (right = + 1)
    uint32_t arrivals; // ARRIVAL: ticket counter
    uint32_t prims;    // primitives in the node (an upper bound below an ARRIVAL node); only kept when treelets are on
    uint32_t p_left;   // BINNED, after SPLIT: primitives left of the plane
}; // 64 B
struct TopSmall
{
    uint32_t count, link, list_base, arrivals; // arrivals: ticket counter of the registering clusters
}; // 16 B
// two-level re-split (hl_builder.cu k_treelets): a node of at most `treelet_prims` primitives leaves the level loop as a
// TREELET — its clusters register in `tlist` during the next BIN phase, and one thread block re-splits the fine clusters below
// them in shared memory.  A treelet of `count` clusters owns count - 1 ids of the free list (what the level loop would have used).
struct TopTreelet
{
    uint32_t count, link, list_base, id_base, arrivals, root; // root: binary node id of the finished treelet (written by k_treelets)
}; // 24 B
// what the per-level passes need of a cluster, in cluster order (written once by top_seed_cluster): the BIN and ASSIGN
// phases stream 32-byte records instead of gathering box + range of a scattered radix-tree node per cluster and level
struct TopCluster
{
    float    lo[3], hi[3];
    uint32_t prims, pad;
};
struct TopBuild
{
    uint32_t        cluster_prims; // C
    uint32_t        k_cap;         // capacity of the per-cluster / per-level arrays
    uint32_t        bins_cap;      // bin slots per level parity
    const uint32_t* n_clusters;    // K (device resident)
    const uint32_t* cluster;       // [K] binary node id of each cluster root
    const uint32_t* free_nodes;    // [K - 1] internal nodes above the cut, the root (id 0) first
    uint32_t*       cnode;         // [K] index of the level node each cluster sits in, HL_TOP_DONE once linked
    TopCluster*     crec;          // [K] box + primitive count of each cluster
    TopNode*        level[2];      // node arrays of the even / odd levels
    TopBin*         bins[2];       // bin arrays of the even / odd levels, 3 * HL_TOP_BINS per slot
    TopSmall*       small;         // [K / 2 + 1] nodes with 2..HL_TOP_SMALL clusters, finished after the level loop
    uint32_t*       small_count;
    uint32_t*       list;          // [(K / 2 + 1) * HL_TOP_SMALL] their cluster lists, HL_TOP_SMALL entries per record
    uint32_t*       level_count;   // [HL_TOP_MAX_LEVELS + 1] nodes per level
    uint32_t*       bins_used;     // [HL_TOP_MAX_LEVELS + 1] bin slots handed out per level
    uint32_t*       free_next;     // next unused entry of free_nodes
    uint32_t        treelet_prims = 0; // 0: no treelets (the level loop runs down to SMALL / SINGLE nodes)
    TopTreelet*     treelet       = nullptr; // [K]
    uint32_t*       treelet_count = nullptr;
    uint32_t*       tlist         = nullptr; // [K] cluster indices, `count` entries per treelet
    uint32_t*       tlist_used    = nullptr;
};

HL_HD uint32_t subtree_prims_cg(const BinaryTree& t, uint32_t node) { return node >= t.n - 1 ? 1u : hl_load_cg(t.last + node) - hl_load_cg(t.first + node) + 1u; }
// the two predicates that define the cut (evaluated on the radix tree before anything is re-linked)
HL_HD bool top_is_cluster_root(const BinaryTree& t, uint32_t m, uint32_t C) { return m != 0u && subtree_prims(t, m) <= C && subtree_prims(t, t.parent[m]) > C; }
HL_HD bool top_is_upper_node(const BinaryTree& t, uint32_t m, uint32_t C) { return m < t.n - 1 && subtree_prims(t, m) > C; }

HL_HD void top_cluster_centroid(const TopCluster& r, float c[3])
{
    for (int k = 0; k < 3; k++) c[k] = 0.5f * (r.lo[k] + r.hi[k]);
}
// bin of centroid coordinate c inside [lo, hi]; -1 when the axis has no extent
HL_HD int top_bin_of(float c, float lo, float hi)
{
    const float ext = hi - lo;
    if (!(ext > 0.0f)) return -1;
    int b = (int)((c - lo) / ext * (float)HL_TOP_BINS);
    return b < 0 ? 0 : (b >= HL_TOP_BINS ? HL_TOP_BINS - 1 : b);
}
// HL_TOP_ONE_AXIS_BELOW: nodes with fewer clusters than this are binned along the longest axis of their centroid bounds
// only (a third of the atomics of the BIN phase); 0 = every node sweeps all three axes.  Default: every node.  Measured
// (profiles/r02h_one_axis_binning.log): build 4.04 -> 3.45 ms at 1M triangles, 20.2 -> 17.4 ms at 10M; SAH cost of the
// terrain 11.20 -> 11.35 with the frame time inside the run-to-run noise (1.52-1.56 ms), and the city's frame 8.65 ->
// 7.98 ms (its box-shaped building meshes get more regular trees from longest-axis planes than from the 3-axis SAH pick).
#ifndef HL_TOP_ONE_AXIS_BELOW
#define HL_TOP_ONE_AXIS_BELOW 0xFFFFFFFFu
#endif
HL_HD int top_bin_axis(const TopNode& N)
{
    if (HL_TOP_ONE_AXIS_BELOW == 0u || hl_load_cg(&N.count) >= HL_TOP_ONE_AXIS_BELOW) return -1;
    float e[3];
    for (int k = 0; k < 3; k++) e[k] = ord2f(hl_load_cg(&N.cb_hi[k])) - ord2f(hl_load_cg(&N.cb_lo[k]));
    return e[0] >= e[1] && e[0] >= e[2] ? 0 : (e[1] >= e[2] ? 1 : 2);
}
// dst.lo = min(dst.lo, vlo), dst.hi = max(dst.hi, vhi), optional counters += (prims, 1).  On the GPU the lanes
// of a warp that arrive here together with the same destination (the common case near the root: clusters are
// in Morton order, so neighbours share node and bin) are combined with warp reductions into one set of atomics
// per distinct destination.
HL_HD void top_merge(uint32_t* lo, uint32_t* hi, uint32_t* counters, const uint32_t vlo[3], const uint32_t vhi[3], uint32_t prims)
{
#if defined(__CUDA_ARCH__)
    // lanes with the same destination form a group; every group reduces on its own (disjoint masks)
    const unsigned grp = __match_any_sync(__activemask(), (unsigned long long)lo);
    if (grp & (grp - 1u))
    {
        uint32_t rl[3], rh[3];
        for (int k = 0; k < 3; k++) rl[k] = __reduce_min_sync(grp, vlo[k]), rh[k] = __reduce_max_sync(grp, vhi[k]);
        const uint32_t rp = counters ? __reduce_add_sync(grp, prims) : 0u;
        unsigned       lane;
        asm("mov.u32 %0, %%laneid;" : "=r"(lane));
        if (lane == (unsigned)(__ffs(grp) - 1))
        {
            for (int k = 0; k < 3; k++) atomicMin(lo + k, rl[k]), atomicMax(hi + k, rh[k]);
            if (counters) atomicAdd(counters, rp), atomicAdd(counters + 1, (uint32_t)__popc(grp));
        }
        return;
    }
#endif
    for (int k = 0; k < 3; k++) hl_atomic_min(lo + k, vlo[k]), hl_atomic_max(hi + k, vhi[k]);
    if (counters) hl_atomic_add(counters, prims), hl_atomic_add(counters + 1, 1u);
}
HL_HD TopBin load_bin_coherent(const TopBin* p)
{
#if defined(__CUDA_ARCH__)
    const uint4 a = __ldcg((const uint4*)p), b = __ldcg((const uint4*)p + 1);
    TopBin      r;
    r.lo[0] = a.x, r.lo[1] = a.y, r.lo[2] = a.z, r.hi[0] = a.w, r.hi[1] = b.x, r.hi[2] = b.y, r.prims = b.z, r.clusters = b.w;
    return r;
#else
    return *p;
#endif
}
HL_HD void top_clear_bin(TopBin* b)
{
#if defined(__CUDA_ARCH__)
    ((uint4*)b)[0] = make_uint4(HL_ORD_POS_INF, HL_ORD_POS_INF, HL_ORD_POS_INF, HL_ORD_NEG_INF);
    ((uint4*)b)[1] = make_uint4(HL_ORD_NEG_INF, HL_ORD_NEG_INF, 0u, 0u);
#else
    for (int k = 0; k < 3; k++) b->lo[k] = HL_ORD_POS_INF, b->hi[k] = HL_ORD_NEG_INF;
    b->prims = 0, b->clusters = 0;
#endif
}
// mode of a node that holds `count` clusters at `level`; BINNED nodes get a bin slot (or fall back when none is left)
HL_HD void top_init_node(TopBuild& tb, uint32_t level, TopNode& n, uint32_t count, uint32_t link, uint32_t prims = 0xFFFFFFFFu)
{
    for (int k = 0; k < 3; k++) n.cb_lo[k] = HL_ORD_POS_INF, n.cb_hi[k] = HL_ORD_NEG_INF;
    n.count = count, n.link = link, n.bins = 0, n.split = 0, n.n_left = 0, n.child = 0, n.arrivals = 0;
    n.prims = prims, n.p_left = 0;
    if (tb.treelet_prims && (prims <= tb.treelet_prims || (uint64_t)count * tb.cluster_prims <= tb.treelet_prims))
    {
        // (prims is exact below BINNED nodes and an upper bound below ARRIVAL nodes; count * cluster size bounds it as well)
        n.mode = HL_TOP_MODE_TREELET, n.bins = hl_alloc(tb.treelet_count, 1u);
        TopTreelet& r = tb.treelet[n.bins];
        r.count = count, r.link = link, r.list_base = hl_alloc(tb.tlist_used, count), r.id_base = count > 1u ? hl_alloc(tb.free_next, count - 1u) : 0u;
        r.arrivals = 0u, r.root = 0xFFFFFFFFu;
        return;
    }
    n.mode = count == 1u ? HL_TOP_MODE_SINGLE : HL_TOP_MODE_ARRIVAL;
    if (count >= 2u && count <= HL_TOP_SMALL)
    {
        n.mode = HL_TOP_MODE_SMALL, n.bins = hl_alloc(tb.small_count, 1u);
        TopSmall& r = tb.small[n.bins];
        r.count = count, r.link = link, r.list_base = n.bins * HL_TOP_SMALL, r.arrivals = 0u;
    }
    else if (count > HL_TOP_SMALL && level < HL_TOP_SAH_LEVELS)
    {
        const uint32_t slot = hl_alloc(tb.bins_used + level, 1u);
        if (slot < tb.bins_cap)
        {
            n.mode = HL_TOP_MODE_BINNED, n.bins = slot; // (the caller's phase clears the bins)
        }
    }
}
// level 0: one node over all K clusters (run by ONE thread before top_seed_cluster)
HL_HD void top_begin(TopBuild& tb, uint32_t K, uint32_t prims = 0xFFFFFFFFu)
{
    tb.level_count[0] = 1u;
    top_init_node(tb, 0u, tb.level[0][0], K, 0xFFFFFFFFu, prims);
    for (uint32_t b = 0; b < 3u * HL_TOP_BINS; b++) top_clear_bin(tb.bins[0] + b);
}
// every cluster starts in the root node and contributes its centroid to the root's centroid bounds
HL_HD void top_seed_cluster(const BinaryTree& t, TopBuild& tb, uint32_t i)
{
    tb.cnode[i] = 0u;
    TopCluster     r;
    const uint32_t m0 = tb.cluster[i];
    for (int k = 0; k < 3; k++) r.lo[k] = t.box[m0].lo[k], r.hi[k] = t.box[m0].hi[k];
    r.prims = subtree_prims(t, m0), r.pad = 0u;
    tb.crec[i] = r;
    float c[3];
    top_cluster_centroid(r, c);
    uint32_t oc[3];
    for (int k = 0; k < 3; k++) oc[k] = f2ord(c[k]);
    TopNode& root = tb.level[0][0];
    top_merge(root.cb_lo, root.cb_hi, nullptr, oc, oc, 0u);
}
HL_HD void top_link(BinaryTree& t, uint32_t link, uint32_t node)
{
    if (link == 0xFFFFFFFFu)
    {
        t.parent[node] = 0xFFFFFFFFu;
        return;
    }
    const uint32_t p = link >> 1;
    if (link & 1u)
        t.right[p] = node;
    else
        t.left[p] = node;
    t.parent[node] = p;
}
// BIN phase, one cluster
HL_HD void top_bin_cluster(BinaryTree& t, TopBuild& tb, uint32_t level, uint32_t i)
{
    const uint32_t nd = tb.cnode[i];
    if (nd == HL_TOP_DONE) return;
    TopNode&       N    = tb.level[level & 1u][nd];
    const uint32_t mode = hl_load_cg(&N.mode);
    const uint32_t m    = tb.cluster[i];
    if (mode == HL_TOP_MODE_SINGLE)
    {
        top_link(t, hl_load_cg(&N.link), m);
        tb.cnode[i] = HL_TOP_DONE;
        return;
    }
    if (mode == HL_TOP_MODE_SMALL)
    {
        // register with the node's record; top_small_node() links the cluster after the level loop
        TopSmall& r = tb.small[hl_load_cg(&N.bins)];
        tb.list[hl_load_cg(&r.list_base) + hl_atomic_add(&r.arrivals, 1u)] = i;
        tb.cnode[i] = HL_TOP_DONE;
        return;
    }
    if (mode == HL_TOP_MODE_TREELET)
    {
        TopTreelet& r = tb.treelet[hl_load_cg(&N.bins)];
        tb.tlist[hl_load_cg(&r.list_base) + hl_atomic_add(&r.arrivals, 1u)] = i;
        tb.cnode[i] = HL_TOP_DONE;
        return;
    }
    if (mode != HL_TOP_MODE_BINNED) return;
    const TopCluster b = tb.crec[i];
    float            c[3];
    top_cluster_centroid(b, c);
    uint32_t olo[3], ohi[3];
    for (int k = 0; k < 3; k++) olo[k] = f2ord(b.lo[k]), ohi[k] = f2ord(b.hi[k]);
    const uint32_t prims = b.prims;
    TopBin*        bins  = tb.bins[level & 1u] + (size_t)hl_load_cg(&N.bins) * (3 * HL_TOP_BINS);
    const int      only  = top_bin_axis(N);
    for (int ax = 0; ax < 3; ax++)
    {
        if (only >= 0 && ax != only) continue;
        const int bi = top_bin_of(c[ax], ord2f(hl_load_cg(&N.cb_lo[ax])), ord2f(hl_load_cg(&N.cb_hi[ax])));
        if (bi < 0) continue;
        TopBin& B = bins[ax * HL_TOP_BINS + bi];
        top_merge(B.lo, B.hi, &B.prims, olo, ohi, prims);
    }
}
// A node with 2..HL_TOP_SMALL clusters: the whole subtree, built by one thread with exact SAH splits (clusters
// sorted along each axis, every position between two neighbours is a candidate).
// the number of binary node ids top_small_node() takes from free_nodes[free_base ...]
HL_HD uint32_t top_small_node_ids(const TopBuild& tb, uint32_t record) { return hl_load_cg(&tb.small[record].count) - 1u; }
HL_HD void top_small_node(BinaryTree& t, TopBuild& tb, uint32_t record, uint32_t free_base)
{
    const uint32_t count = hl_load_cg(&tb.small[record].count), link = hl_load_cg(&tb.small[record].link), list_base = hl_load_cg(&tb.small[record].list_base);
    uint32_t       ids[HL_TOP_SMALL], pr[HL_TOP_SMALL];
    Box            bx[HL_TOP_SMALL];
    uint8_t        ord[HL_TOP_SMALL]; // permutation of the clusters; sub-ranges of it are the nodes being built
    // the clusters registered in arrival order (an atomic ticket): sort them by cluster index first, so that the subtree —
    // ties between equal centroids included — is a function of the SET of clusters, not of the scheduling of one run
    uint32_t idx[HL_TOP_SMALL];
    for (uint32_t k = 0; k < count; k++)
    {
        const uint32_t v = hl_load_cg(tb.list + list_base + k);
        int            j = (int)k - 1;
        while (j >= 0 && idx[j] > v) idx[j + 1] = idx[j], j--;
        idx[j + 1] = v;
    }
    for (uint32_t k = 0; k < count; k++)
    {
        ids[k] = tb.cluster[idx[k]];
        bx[k] = t.box[ids[k]], pr[k] = subtree_prims(t, ids[k]), ord[k] = (uint8_t)k;
    }
    uint32_t       used = 0;
    uint8_t        slo[HL_TOP_SMALL], shi[HL_TOP_SMALL];
    uint32_t       slink[HL_TOP_SMALL];
    int            sp = 0;
    slo[0] = 0, shi[0] = (uint8_t)count, slink[0] = link, sp = 1;
    while (sp > 0)
    {
        sp--;
        const int      lo = slo[sp], hi = shi[sp];
        const uint32_t lk = slink[sp];
        if (hi - lo == 1)
        {
            top_link(t, lk, ids[ord[lo]]);
            continue;
        }
        const uint32_t self = tb.free_nodes[free_base + used++];
        top_link(t, lk, self);
        t.visits[self] = 0u;
        float best = hl_inf();
        int   bax = 2, bk = (lo + hi) / 2;
        for (int pass = 0; pass < 4; pass++)
        {
            // passes 0..2 evaluate the axes; pass 3 restores the order of the winning axis (unless it was sorted last)
            const int ax = pass < 3 ? pass : bax;
            if (pass == 3 && bax == 2) break;
            for (int a = lo + 1; a < hi; a++)
            {
                const uint8_t m  = ord[a];
                const float   cm = bx[m].lo[ax] + bx[m].hi[ax];
                int           b  = a - 1;
                while (b >= lo && bx[ord[b]].lo[ax] + bx[ord[b]].hi[ax] > cm) ord[b + 1] = ord[b], b--;
                ord[b + 1] = m;
            }
            if (pass == 3) break;
            float    rarea[HL_TOP_SMALL];
            uint32_t rprims[HL_TOP_SMALL];
            Box      acc = bx[ord[hi - 1]];
            uint32_t p   = pr[ord[hi - 1]];
            rarea[hi - 1] = box_half_area(acc), rprims[hi - 1] = p;
            for (int k = hi - 2; k > lo; k--)
            {
                acc = box_union(acc, bx[ord[k]]), p += pr[ord[k]];
                rarea[k] = box_half_area(acc), rprims[k] = p;
            }
            acc = bx[ord[lo]], p = pr[ord[lo]];
            for (int k = lo + 1; k < hi; k++)
            {
                // left = [lo, k), right = [k, hi)
                const float cost = box_half_area(acc) * (float)p + rarea[k] * (float)rprims[k];
                if (cost < best) best = cost, bax = ax, bk = k;
                acc = box_union(acc, bx[ord[k]]), p += pr[ord[k]];
            }
        }
        slo[sp] = (uint8_t)bk, shi[sp] = (uint8_t)hi, slink[sp] = (self << 1) | 1u, sp++;
        slo[sp] = (uint8_t)lo, shi[sp] = (uint8_t)bk, slink[sp] = self << 1, sp++;
    }
}
// SPLIT phase, step 1 (CHOOSE), one BINNED node of `level`: sweep the bins for the cheapest plane and record it in
// the node (split, n_left), or turn the node into an ARRIVAL node when no plane separates its clusters.  This is
// the one-thread form (emulator); the CUDA builder runs the same selection with one warp per node, a bin per
// lane and shuffle scans (hl_builder.cu: top_choose_node_warp).
HL_HD void top_choose_node(TopBuild& tb, uint32_t level, uint32_t j)
{
    TopNode& N = tb.level[level & 1u][j];
    if (hl_load_cg(&N.mode) != HL_TOP_MODE_BINNED) return;
    const TopBin* bins = tb.bins[level & 1u] + (size_t)hl_load_cg(&N.bins) * (3 * HL_TOP_BINS);
    float         best = hl_inf();
    uint32_t      split = 0u, n_left = 0u, p_left = 0u;
    for (int ax = 0; ax < 3; ax++)
    {
        // suffix unions, then a prefix sweep over the 15 candidate planes (an empty bin is the identity of the union)
        Box      bb[HL_TOP_BINS];
        uint32_t bp[HL_TOP_BINS], bc[HL_TOP_BINS];
        for (int b = 0; b < HL_TOP_BINS; b++)
        {
            const TopBin B = load_bin_coherent(bins + ax * HL_TOP_BINS + b);
            for (int k = 0; k < 3; k++) bb[b].lo[k] = ord2f(B.lo[k]), bb[b].hi[k] = ord2f(B.hi[k]);
            bp[b] = B.prims, bc[b] = B.clusters;
        }
        float    rarea[HL_TOP_BINS];
        uint32_t rp[HL_TOP_BINS], rc[HL_TOP_BINS];
        Box      acc = bb[HL_TOP_BINS - 1];
        uint32_t p = bp[HL_TOP_BINS - 1], c = bc[HL_TOP_BINS - 1];
        rarea[HL_TOP_BINS - 1] = box_half_area(acc), rp[HL_TOP_BINS - 1] = p, rc[HL_TOP_BINS - 1] = c;
        for (int b = HL_TOP_BINS - 2; b > 0; b--)
        {
            acc = box_union(acc, bb[b]), p += bp[b], c += bc[b];
            rarea[b] = box_half_area(acc), rp[b] = p, rc[b] = c;
        }
        acc = bb[0], p = bp[0], c = bc[0];
        for (int b = 0; b < HL_TOP_BINS - 1; b++)
        {
            // left = bins 0..b, right = bins b+1..15
            if (c != 0u && rc[b + 1] != 0u)
            {
                const float cost = box_half_area(acc) * (float)p + rarea[b + 1] * (float)rp[b + 1];
                if (cost < best) best = cost, split = (uint32_t)ax | ((uint32_t)b << 2), n_left = c, p_left = p;
            }
            acc = box_union(acc, bb[b + 1]), p += bp[b + 1], c += bc[b + 1];
        }
    }
    if (best < hl_inf())
        N.split = split, N.n_left = n_left, N.p_left = p_left;
    else
        N.mode = HL_TOP_MODE_ARRIVAL;
}
// SPLIT phase, step 2 (COMMIT), one node of `level` (after every node of the level has chosen): take a binary node
// id, link it to the parent, create the two children in the next level's array.  Their bins are cleared by
// top_clear_bin() calls spread over all threads during the ASSIGN phase.
HL_HD void top_commit_node(BinaryTree& t, TopBuild& tb, uint32_t level, uint32_t j)
{
    // (fields written by other thread blocks are read past the L1)
    TopNode&       N    = tb.level[level & 1u][j];
    const uint32_t mode = hl_load_cg(&N.mode), count = hl_load_cg(&N.count);
    if (mode == HL_TOP_MODE_SINGLE || mode == HL_TOP_MODE_SMALL || mode == HL_TOP_MODE_TREELET) return;
    const uint32_t self = tb.free_nodes[hl_alloc(tb.free_next, 1u)];
    top_link(t, hl_load_cg(&N.link), self);
    t.visits[self]       = 0u;
    const uint32_t n_left = mode == HL_TOP_MODE_BINNED ? hl_load_cg(&N.n_left) : count / 2u;
    const uint32_t child  = hl_alloc(tb.level_count + (level + 1u), 2u);
    N.n_left = n_left, N.child = child, N.arrivals = 0u;
    TopNode* next = tb.level[(level + 1u) & 1u];
    const uint32_t prims = hl_load_cg(&N.prims), p_left = hl_load_cg(&N.p_left);
    const bool     exact = mode == HL_TOP_MODE_BINNED && prims != 0xFFFFFFFFu && p_left <= prims;
    top_init_node(tb, level + 1u, next[child], n_left, self << 1, exact ? p_left : prims);
    top_init_node(tb, level + 1u, next[child + 1u], count - n_left, (self << 1) | 1u, exact ? prims - p_left : prims);
}
// bin entries of `level` that must be empty before its BIN phase
HL_HD uint32_t top_bins_to_clear(const TopBuild& tb, uint32_t level)
{
    const uint32_t used = hl_load_cg(tb.bins_used + level);
    return (used < tb.bins_cap ? used : tb.bins_cap) * (3u * HL_TOP_BINS);
}
// ASSIGN phase, one cluster
HL_HD void top_assign_cluster(const BinaryTree& t, TopBuild& tb, uint32_t level, uint32_t i)
{
    const uint32_t nd = tb.cnode[i];
    if (nd == HL_TOP_DONE) return;
    TopNode& N = tb.level[level & 1u][nd];
    float    c[3];
    top_cluster_centroid(tb.crec[i], c);
    uint32_t side;
    if (hl_load_cg(&N.mode) == HL_TOP_MODE_BINNED)
    {
        const uint32_t split = hl_load_cg(&N.split);
        const int      ax    = (int)(split & 3u);
        side                 = top_bin_of(c[ax], ord2f(hl_load_cg(&N.cb_lo[ax])), ord2f(hl_load_cg(&N.cb_hi[ax]))) > (int)(split >> 2) ? 1u : 0u;
    }
    else
        side = hl_atomic_add(&N.arrivals, 1u) >= hl_load_cg(&N.n_left) ? 1u : 0u;
    const uint32_t child = hl_load_cg(&N.child) + side;
    tb.cnode[i]          = child;
    TopNode& Cn          = tb.level[(level + 1u) & 1u][child];
    uint32_t oc[3];
    for (int k = 0; k < 3; k++) oc[k] = f2ord(c[k]);
    top_merge(Cn.cb_lo, Cn.cb_hi, nullptr, oc, oc, 0u);
}
// bottom-up over the re-linked nodes, started once per cluster (own launch, after the re-split): box, primitive
// count (kept in `last` with first = 0; ranges above the cut are no longer contiguous in Morton order, so the
// stored count is floored at HL_MAX_LEAF_PRIMS + 1 — such a node can never be emitted as a leaf) and the
// collapse cost table.
template <class Fence>
HL_HD void top_refit_from_cluster(BinaryTree& t, uint32_t cluster_root, Fence fence)
{
    uint32_t node = t.parent[cluster_root];
    while (node != 0xFFFFFFFFu)
    {
        fence();
        if (hl_atomic_add(&t.visits[node], 1u) == 0u) return;
        fence();
        const uint32_t l = t.left[node], r = t.right[node];
        const Box      b = box_union(load_box_coherent(&t.box[l]), load_box_coherent(&t.box[r]));
        t.box[node]      = b;
        const uint32_t p = subtree_prims_cg(t, l) + subtree_prims_cg(t, r);
        t.first[node] = 0u, t.last[node] = (p > HL_MAX_LEAF_PRIMS ? p : HL_MAX_LEAF_PRIMS + 1u) - 1u;
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
HL_HD uint32_t collapse_one(const BinaryTree& t, CollapseTask task, WideOut out, CollapseTask* next, uint32_t* next_count, LeafWriter& leaf_writer, uint32_t* outstanding = nullptr)
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
            uint32_t       d = 0, lm = 0, rm = 0;
            if (m < leaf0)
            {
                // (decisions and links together: one round trip per binary level instead of two)
                const uint32_t dw = t.dec[m];
                lm = t.left[m], rm = t.right[m];
                d  = (dw >> (4 * ((i < 1 ? 1 : i) - 1))) & 15u;
            }
            if (m >= leaf0 || d == 0u)
                ch[nch++] = m;
            else if (d == 15u)
                ch_inner |= 1u << nch, ch[nch++] = m;
            else
            {
                sn[sp] = rm, ss[sp++] = i - d;
                sn[sp] = lm, ss[sp++] = d;
            }
        }
    }
    const Box nb = t.box[task.bnode];
    // octant-ordered slot assignment: greedy minimum of cost(child, slot) = (centroid_c - centroid_node) . dir(slot)
    // the children's boxes and leaf ranges, loaded ONCE and all at the same time (unrolled: eight independent gathers in flight —
    // the kernel waits on memory, ncu: 46 long-scoreboard stalls per issued instruction when they were loaded one per loop trip,
    // twice)
    Box      cb[8];
    uint32_t cprims[8], cfirst[8];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 8; k++)
        if (k < nch)
        {
            const uint32_t c = ch[k];
            cb[k] = t.box[c];
            const bool     lf = c >= leaf0;
            const uint32_t f = lf ? c - leaf0 : t.first[c], l = lf ? c - leaf0 : t.last[c];
            cfirst[k] = f, cprims[k] = l - f + 1u;
        }
    float cx[8], cy[8], cz[8];
    for (int k = 0; k < nch; k++)
    {
        const Box& b = cb[k];
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
        // (the remaining children x the free slots, both in ascending order: bit scans instead of 64 tests per round)
        for (uint32_t km = ~child_done & ((1u << nch) - 1u); km; km &= km - 1u)
        {
            const int k = hl_bfind(km & (0u - km));
            for (uint32_t sm = ~slot_done & 0xFFu; sm; sm &= sm - 1u)
            {
                const int   s = hl_bfind(sm & (0u - sm));
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
        if (ch_inner & (1u << slot_child[s]))
            n_inner++, imask |= 1u << s;
        else
            n_leafprims += cprims[slot_child[s]];
    }
    w.imask      = (uint8_t)imask;
    // (every lane allocates, also with amount 0: the lanes of a warp that build a node together share one atomic per counter)
    w.child_base = hl_alloc_var(out.node_counter, n_inner);
    w.leaf_base  = hl_alloc_var(out.leaf_counter, n_leafprims);
    uint32_t inner_rank = 0, leaf_off = 0;
    uint32_t first_task = hl_alloc_var(next_count, n_inner);
    if (!n_inner) w.child_base = 0u;
    if (!n_leafprims) w.leaf_base = 0u;
    // the children's tasks first: the threads polling for them start while this one still quantises boxes and writes leaves.
    // `outstanding` (persistent GPU launch: published - finished tasks) grows BEFORE a child can be seen, so that a child which
    // finishes before its parent does never brings the count to zero early.
    if (outstanding)
    {
        hl_alloc_var(outstanding, n_inner);
#if defined(__CUDA_ARCH__)
        __threadfence(); // the count is in place before any of the tasks below can be seen
#endif
    }
    for (int s = 0; s < 8; s++)
        if (slot_child[s] >= 0 && (imask & (1u << s)))
        {
            CollapseTask nt;
            nt.wide = w.child_base + inner_rank, nt.bnode = ch[slot_child[s]];
            publish_task(next + (first_task + inner_rank), nt);
            inner_rank++;
        }
    for (int s = 0; s < 8; s++)
    {
        if (slot_child[s] < 0)
        {
            w.meta[s] = 0;
            w.qlox[s] = w.qloy[s] = w.qloz[s] = 255;
            w.qhix[s] = w.qhiy[s] = w.qhiz[s] = 0;
            continue;
        }
        const Box& b = cb[slot_child[s]];
        quantize_axis(w.px, ex, b.lo[0], b.hi[0], w.qlox[s], w.qhix[s]);
        quantize_axis(w.py, ey, b.lo[1], b.hi[1], w.qloy[s], w.qhiy[s]);
        quantize_axis(w.pz, ez, b.lo[2], b.hi[2], w.qloz[s], w.qhiz[s]);
        if (imask & (1u << s))
        {
            w.meta[s] = (uint8_t)((1u << 5) | (24u + (uint32_t)s));
        }
        else
        {
            const uint32_t cnt = cprims[slot_child[s]], f = cfirst[slot_child[s]];
            w.meta[s]          = (uint8_t)((((1u << cnt) - 1u) << 5) | leaf_off);
            for (uint32_t k = 0; k < cnt; k++) leaf_writer(w.leaf_base + leaf_off + k, f + k);
            leaf_off += cnt;
        }
    }
    out.nodes[task.wide] = w;
    return n_inner;
}

// ---- instance-tree refit (moving instances) ---------------------------------------------------------------------------
// The reference creates its top-level structure with ALLOW_UPDATE (src/engine/resource/scene.cpp:797) and re-issues the build
// whenever a transform changes (src/engine/gfx/renderer.cpp:147-168).  Here the instance tree keeps its topology and only
// the boxes move: refit_node_box() recomputes the box of one wide node from its children (instance boxes for leaf slots,
// the children's boxes — from the previous sweep — for inner slots), refit_requantize() rewrites the node's origin,
// exponents and the 8-bit child planes from those boxes with the builder's own conservative rounding.  Slot assignment
// (the octant order chosen at build time) stays: it affects visit order only, never a result.
HL_HD Box box_empty()
{
    Box b;
    for (int k = 0; k < 3; k++) b.lo[k] = 3.0e38f, b.hi[k] = -3.0e38f;
    return b;
}
HL_HD uint32_t wide_inner_rank(const WideNode& w, int s) { return (uint32_t)hl_popc((uint32_t)w.imask & ((1u << s) - 1u)); }
// box of child slot s of an instance-tree node (slot must not be empty)
HL_HD Box refit_child_box(const WideNode& w, int s, const uint32_t* inst_leaf, const Box* inst_boxes, const Box* node_box)
{
    if (w.imask & (1u << s)) return load_box_coherent(&node_box[w.child_base + wide_inner_rank(w, s)]);
    const uint32_t meta = w.meta[s], off = meta & 31u;
    Box            b    = box_empty();
    for (uint32_t k = 0; k < 3u; k++)
        if ((meta >> 5) & (1u << k)) b = box_union(b, inst_boxes[inst_leaf[w.leaf_base + off + k]]);
    return b;
}
HL_HD Box refit_node_box(const WideNode& w, const uint32_t* inst_leaf, const Box* inst_boxes, const Box* node_box)
{
    Box b = box_empty();
    for (int s = 0; s < 8; s++)
        if (w.meta[s]) b = box_union(b, refit_child_box(w, s, inst_leaf, inst_boxes, node_box));
    return b;
}
HL_HD void refit_requantize(WideNode& w, const Box& nb, const uint32_t* inst_leaf, const Box* inst_boxes, const Box* node_box)
{
    w.px = nb.lo[0], w.py = nb.lo[1], w.pz = nb.lo[2];
    const uint32_t ex = quant_exponent(nb.hi[0] - nb.lo[0]), ey = quant_exponent(nb.hi[1] - nb.lo[1]), ez = quant_exponent(nb.hi[2] - nb.lo[2]);
    w.ex = (uint8_t)ex, w.ey = (uint8_t)ey, w.ez = (uint8_t)ez;
    for (int s = 0; s < 8; s++)
    {
        if (!w.meta[s]) continue;
        const Box b = refit_child_box(w, s, inst_leaf, inst_boxes, node_box);
        quantize_axis(w.px, ex, b.lo[0], b.hi[0], w.qlox[s], w.qhix[s]);
        quantize_axis(w.py, ey, b.lo[1], b.hi[1], w.qloy[s], w.qhiy[s]);
        quantize_axis(w.pz, ez, b.lo[2], b.hi[2], w.qloz[s], w.qhiz[s]);
    }
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
    AlphaTri*         alpha = nullptr; // any-hit records beside the leaves (meshes with a non-opaque submesh), or nullptr
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
        if (alpha)
        {
            const float* t0 = vertices[indices[b + 0]].tex_coord;
            const float* t1 = vertices[indices[b + 1]].tex_coord;
            const float* t2 = vertices[indices[b + 2]].tex_coord;
            AlphaTri     q;
            q.u0 = t0[0], q.v0 = t0[1], q.u1 = t1[0], q.v1 = t1[1], q.u2 = t2[0], q.v2 = t2[1], q.pad[0] = q.pad[1] = 0.0f;
            alpha[dst] = q;
        }
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
