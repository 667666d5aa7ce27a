// hl_builder.cu — GPU BVH builder (replaces vkCmdBuildAccelerationStructuresKHR: per-mesh BLAS build
// src/engine/gfx/vk.cpp:3207-3226, TLAS build src/engine/gfx/renderer.cpp:147-168 in the reference).
// Kernels are thin per-thread loops around the element functions in hl_build.h.
//
// Pass list and algorithmic bytes per triangle (N triangles, fp32 boxes of 24 B, 64-bit Morton keys):
//   tri_boxes      12 idx + 36 pos read, 24 written                      72
//   box_reduce     24 read                                               24
//   morton         24 read, 8 + 4 written                                36
//   radix sort     8 passes x (12 read + 12 written)  (CUB, 63 key bits) 192
//   radix_tree     ~2 x 8 key reads (cached), 24 written                 40
//   fit (fine)     24 read (sorted box) + 24 written + 28 cost row per leaf, the same per node inside a fine cluster        ~230
//   re-split       two levels (k_treelets below): level loop over the coarse cut (K_A ~ N / 20 clusters: 32 B records, ~12 levels:
//                  a few MB per level, bound by its grid barriers), then per treelet of <= 512 primitives ONE pass over its fine
//                  clusters: ~4 links + 24 box + 28 cost row read per cluster, and per re-linked node 3 links + 24 box + 8 range +
//                  28 cost row + 4 decisions written (all levels of the treelet run in shared memory)                            ~150
//                  (instance trees / HL_NO_TREELETS: the one-level re-split of round 1 — clusters of <= C primitives through a
//                  level loop of BIN / CHOOSE / COMMIT / ASSIGN passes, K x (24 box + 4 node id) read twice + ~27 atomics per level)
//   collapse       ~48 read (boxes) + 8 decisions + 0.12 x 80 node + 48 leaf + 48 src + 2 x 8 queue   ~180
//   total                                                              ~ 930 B / triangle
#include "hl_internal.h"
#include <cooperative_groups/reduce.h>
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>

namespace hl
{
__global__ void k_tri_boxes(const hl_vertex* v, const uint32_t* idx, const hl_submesh* subs, const uint32_t* tri_start, uint32_t n_geom, uint32_t n, Box* out)
{
    for (uint32_t f = blockIdx.x * blockDim.x + threadIdx.x; f < n; f += gridDim.x * blockDim.x) out[f] = triangle_box(v, idx, subs, tri_start, n_geom, f);
}

__global__ void k_box_reduce(const Box* in, uint32_t n, Box* out)
{
    __shared__ Box sm[256];
    Box            acc;
    for (int k = 0; k < 3; k++) acc.lo[k] = 3.0e38f, acc.hi[k] = -3.0e38f;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc = box_union(acc, in[i]);
    sm[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1)
    {
        if ((int)threadIdx.x < s) sm[threadIdx.x] = box_union(sm[threadIdx.x], sm[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = sm[0];
}

__global__ void k_morton(const Box* boxes, const Box* scene, uint32_t n, uint64_t* keys, uint32_t* vals)
{
    const Box sb = *scene;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        keys[i] = morton_key(boxes[i], sb);
        vals[i] = i;
    }
}

__global__ void k_radix_tree(const uint64_t* keys, BinaryTree t)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i + 1 < t.n; i += gridDim.x * blockDim.x) radix_tree_node(keys, t, (int)i);
}

struct DeviceFence
{
    __device__ void operator()() const { __threadfence(); }
};
__global__ void k_fit(BinaryTree t, const Box* prim_boxes, const uint32_t* sorted, uint32_t stop_above)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < t.n; j += gridDim.x * blockDim.x)
    {
        t.box[(t.n - 1) + j] = prim_boxes[sorted[j]];
        fit_from_leaf(t, j, DeviceFence(), stop_above);
    }
}

// ---- binned-SAH re-split of the upper levels (hl_build.h: top_*): one persistent launch, the three phases of
// every level separated by a grid-wide barrier (the launch is cooperative, so all blocks are resident).
struct IsClusterRoot
{
    BinaryTree t;
    uint32_t   C;
    __device__ bool operator()(uint32_t m) const { return top_is_cluster_root(t, m, C); }
};
struct IsUpperNode
{
    BinaryTree t;
    uint32_t   C;
    __device__ bool operator()(uint32_t m) const { return top_is_upper_node(t, m, C); }
};
// bar[0] = arrivals, bar[1] = generation
__device__ __forceinline__ void grid_barrier(uint32_t* bar)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        volatile uint32_t* gen = bar + 1;
        const uint32_t     g   = *gen;
        __threadfence();
        if (atomicAdd(bar, 1u) == gridDim.x - 1u)
        {
            bar[0] = 0u;
            __threadfence();
            atomicAdd(bar + 1, 1u);
        }
        else
            while (*gen == g) __nanosleep(40);
        __threadfence();
    }
    __syncthreads();
}
__device__ __forceinline__ void top_trace(unsigned long long* trace, uint32_t& n)
{
    if (trace && blockIdx.x == 0 && threadIdx.x == 0 && n < 1000u)
    {
        unsigned long long ns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
        trace[n++] = ns;
    }
}
// CHOOSE for one BINNED node by one warp (same selection as top_choose_node, hl_build.h): lane b < 16 holds bin b
// of the current axis, inclusive prefix / suffix unions by shuffle scans, candidate plane b = prefix(b) | suffix(b+1),
// warp arg-min over the 3 x 15 candidates.  All in registers.
__device__ __forceinline__ void top_choose_node_warp(TopBuild& tb, uint32_t level, uint32_t j, uint32_t lane)
{
    const unsigned FULL = 0xFFFFFFFFu;
    TopNode&       N    = tb.level[level & 1u][j];
    if (__ldcg(&N.mode) != HL_TOP_MODE_BINNED) return; // warp-uniform
    const TopBin* bins = tb.bins[level & 1u] + (size_t)__ldcg(&N.bins) * (3 * HL_TOP_BINS);
    float         best = hl_inf();
    uint32_t      split = 0xFFFFFFFFu, n_left = 0u, p_left = 0u;
    for (int ax = 0; ax < 3; ax++)
    {
        float    lo[3], hi[3];
        uint32_t p = 0u, c = 0u;
        for (int k = 0; k < 3; k++) lo[k] = hl_inf(), hi[k] = -hl_inf();
        if (lane < HL_TOP_BINS)
        {
            const TopBin B = load_bin_coherent(bins + ax * HL_TOP_BINS + lane);
            for (int k = 0; k < 3; k++) lo[k] = ord2f(B.lo[k]), hi[k] = ord2f(B.hi[k]);
            p = B.prims, c = B.clusters;
        }
        float    plo[3] = { lo[0], lo[1], lo[2] }, phi[3] = { hi[0], hi[1], hi[2] }, slo[3] = { lo[0], lo[1], lo[2] }, shi[3] = { hi[0], hi[1], hi[2] };
        uint32_t pp = p, pc = c, sp = p, sc = c;
#pragma unroll
        for (int d = 1; d < HL_TOP_BINS; d <<= 1)
        {
            // lanes >= 16 hold the identity, so the suffix scan may pull from them without a range check below lane 16
            const bool up = (int)lane >= d, dn = lane + d < 32u;
#pragma unroll
            for (int k = 0; k < 3; k++)
            {
                const float a = __shfl_up_sync(FULL, plo[k], d), b = __shfl_up_sync(FULL, phi[k], d);
                const float e = __shfl_down_sync(FULL, slo[k], d), f = __shfl_down_sync(FULL, shi[k], d);
                if (up) plo[k] = fminf(plo[k], a), phi[k] = fmaxf(phi[k], b);
                if (dn) slo[k] = fminf(slo[k], e), shi[k] = fmaxf(shi[k], f);
            }
            const uint32_t a = __shfl_up_sync(FULL, pp, d), b = __shfl_up_sync(FULL, pc, d);
            const uint32_t e = __shfl_down_sync(FULL, sp, d), f = __shfl_down_sync(FULL, sc, d);
            if (up) pp += a, pc += b;
            if (dn) sp += e, sc += f;
        }
        // suffix of the bins right of this lane
        const float    sarea = (shi[0] - slo[0]) * (shi[1] - slo[1]) + (shi[1] - slo[1]) * (shi[2] - slo[2]) + (shi[2] - slo[2]) * (shi[0] - slo[0]);
        const float    rarea = __shfl_down_sync(FULL, sarea, 1);
        const uint32_t rp = __shfl_down_sync(FULL, sp, 1), rc = __shfl_down_sync(FULL, sc, 1);
        if (lane < HL_TOP_BINS - 1u && pc != 0u && rc != 0u)
        {
            const float parea = (phi[0] - plo[0]) * (phi[1] - plo[1]) + (phi[1] - plo[1]) * (phi[2] - plo[2]) + (phi[2] - plo[2]) * (phi[0] - plo[0]);
            const float cost  = parea * (float)pp + rarea * (float)rp;
            if (cost < best) best = cost, split = (uint32_t)ax | (lane << 2), n_left = pc, p_left = pp;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        const float    ob = __shfl_xor_sync(FULL, best, d);
        const uint32_t os = __shfl_xor_sync(FULL, split, d), on = __shfl_xor_sync(FULL, n_left, d), op = __shfl_xor_sync(FULL, p_left, d);
        if (ob < best || (ob == best && os < split)) best = ob, split = os, n_left = on, p_left = op;
    }
    if (lane == 0u)
    {
        if (best < hl_inf())
            N.split = split, N.n_left = n_left, N.p_left = p_left;
        else
            N.mode = HL_TOP_MODE_ARRIVAL;
    }
}
// The first levels hold a handful of nodes, and every cluster of the tree adds itself to their bins: with global atomics
// the BIN and ASSIGN phases of those levels serialise on a few hundred addresses (traced at 1M triangles: 60-180 us per
// level for work that moves 18 MB).  While a level has at most HL_TOP_SMEM_NODES nodes every block accumulates into its own
// copy of the bins / of the children's centroid bounds in shared memory and flushes each non-empty entry once.
#define HL_TOP_SMEM_NODES 16u
__device__ __forceinline__ void top_bin_level_private(BinaryTree& t, TopBuild& tb, uint32_t level, uint32_t K, uint32_t nodes, TopBin* sb, uint32_t tid, uint32_t nthr)
{
    const uint32_t nb = nodes * 3u * HL_TOP_BINS;
    for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x) top_clear_bin(sb + b);
    __syncthreads();
    for (uint32_t i = tid; i < K; i += nthr)
    {
        const uint32_t nd = tb.cnode[i];
        if (nd == HL_TOP_DONE) continue;
        TopNode& N = tb.level[level & 1u][nd];
        if (__ldcg(&N.mode) != HL_TOP_MODE_BINNED)
        {
            top_bin_cluster(t, tb, level, i); // SINGLE / SMALL / ARRIVAL: the generic path (no bins involved)
            continue;
        }
        const TopCluster b = tb.crec[i];
        float            c[3];
        top_cluster_centroid(b, c);
        const uint32_t prims = b.prims;
        TopBin*        bins  = sb + nd * (3u * HL_TOP_BINS);
        const int      only  = top_bin_axis(N);
        for (int ax = 0; ax < 3; ax++)
        {
            if (only >= 0 && ax != only) continue;
            const int bi = top_bin_of(c[ax], ord2f(__ldcg(&N.cb_lo[ax])), ord2f(__ldcg(&N.cb_hi[ax])));
            if (bi < 0) continue;
            TopBin& B = bins[ax * HL_TOP_BINS + bi];
            for (int k = 0; k < 3; k++) atomicMin(&B.lo[k], f2ord(b.lo[k])), atomicMax(&B.hi[k], f2ord(b.hi[k]));
            atomicAdd(&B.prims, prims), atomicAdd(&B.clusters, 1u);
        }
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nb; b += blockDim.x)
    {
        const TopBin& S = sb[b];
        if (S.clusters == 0u) continue;
        const TopNode& N = tb.level[level & 1u][b / (3u * HL_TOP_BINS)];
        TopBin&        G = tb.bins[level & 1u][(size_t)__ldcg(&N.bins) * (3u * HL_TOP_BINS) + b % (3u * HL_TOP_BINS)];
        for (int k = 0; k < 3; k++) atomicMin(&G.lo[k], S.lo[k]), atomicMax(&G.hi[k], S.hi[k]);
        atomicAdd(&G.prims, S.prims), atomicAdd(&G.clusters, S.clusters);
    }
    __syncthreads();
}
__device__ __forceinline__ void top_assign_level_private(const BinaryTree& t, TopBuild& tb, uint32_t level, uint32_t K, uint32_t next_nodes, uint32_t* scb, uint32_t tid, uint32_t nthr)
{
    for (uint32_t k = threadIdx.x; k < next_nodes * 6u; k += blockDim.x) scb[k] = (k % 6u) < 3u ? HL_ORD_POS_INF : HL_ORD_NEG_INF;
    __syncthreads();
    for (uint32_t i = tid; i < K; i += nthr)
    {
        const uint32_t nd = tb.cnode[i];
        if (nd == HL_TOP_DONE) continue;
        TopNode& N = tb.level[level & 1u][nd];
        float    c[3];
        top_cluster_centroid(tb.crec[i], c);
        uint32_t side;
        if (__ldcg(&N.mode) == HL_TOP_MODE_BINNED)
        {
            const uint32_t split = __ldcg(&N.split);
            const int      ax    = (int)(split & 3u);
            side                 = top_bin_of(c[ax], ord2f(__ldcg(&N.cb_lo[ax])), ord2f(__ldcg(&N.cb_hi[ax]))) > (int)(split >> 2) ? 1u : 0u;
        }
        else
            side = atomicAdd(&N.arrivals, 1u) >= __ldcg(&N.n_left) ? 1u : 0u;
        const uint32_t child = __ldcg(&N.child) + side;
        tb.cnode[i]          = child;
        for (int k = 0; k < 3; k++) atomicMin(&scb[child * 6u + k], f2ord(c[k])), atomicMax(&scb[child * 6u + 3u + k], f2ord(c[k]));
    }
    __syncthreads();
    TopNode* next = tb.level[(level + 1u) & 1u];
    for (uint32_t k = threadIdx.x; k < next_nodes * 6u; k += blockDim.x)
    {
        const uint32_t v = scb[k], child = k / 6u, a = k % 6u;
        if (a < 3u)
        {
            if (v != HL_ORD_POS_INF) atomicMin(&next[child].cb_lo[a], v);
        }
        else if (v != HL_ORD_NEG_INF)
            atomicMax(&next[child].cb_hi[a - 3u], v);
    }
    __syncthreads();
}
// `trace` (debug, HL_TOP_TRACE=1): global timer after every phase, read back and printed by the host
__global__ void __launch_bounds__(256, 4) k_top_build(BinaryTree t, TopBuild tb, uint32_t* bar, unsigned long long* trace)
{
    __shared__ TopBin s_bins[HL_TOP_SMEM_NODES * 3u * HL_TOP_BINS]; // 24 KB; the ASSIGN phase reuses it for 2 x 16 x 6 words
    uint32_t       ntrace = 1;
    const uint32_t K      = *tb.n_clusters;
    if (K < 2u || K > tb.k_cap) return; // (uniform) cut too fine for the scratch arrays: keep the radix tree
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    const uint32_t lane = threadIdx.x & 31u, warp = tid >> 5, nwarps = nthr >> 5;
    top_trace(trace, ntrace);
    if (tid == 0) top_begin(tb, K, t.n);
    grid_barrier(bar);
    for (uint32_t i = tid; i < K; i += nthr) top_seed_cluster(t, tb, i);
    grid_barrier(bar);
    top_trace(trace, ntrace);
    for (uint32_t level = 0; level + 1u < HL_TOP_MAX_LEVELS; level++)
    {
        const uint32_t nodes = __ldcg(tb.level_count + level);
        if (nodes <= HL_TOP_SMEM_NODES)
            top_bin_level_private(t, tb, level, K, nodes, s_bins, tid, nthr);
        else
            for (uint32_t i = tid; i < K; i += nthr) top_bin_cluster(t, tb, level, i);
        grid_barrier(bar);
        top_trace(trace, ntrace);
        // CHOOSE + COMMIT of a node by the same warp (a node's commit only needs its own choice): one grid barrier less per level
        for (uint32_t j = warp; j < nodes; j += nwarps)
        {
            top_choose_node_warp(tb, level, j, lane);
            __syncwarp();
            if (lane == 0u) top_commit_node(t, tb, level, j);
        }
        grid_barrier(bar);
        top_trace(trace, ntrace);
        if (__ldcg(tb.level_count + level + 1u) == 0u) break; // only single / small nodes were left
        TopBin* next_bins = tb.bins[(level + 1u) & 1u];
        for (uint32_t b = tid, nb = top_bins_to_clear(tb, level + 1u); b < nb; b += nthr) top_clear_bin(next_bins + b);
        const uint32_t next_nodes = __ldcg(tb.level_count + level + 1u);
        if (next_nodes <= 2u * HL_TOP_SMEM_NODES)
            top_assign_level_private(t, tb, level, K, next_nodes, (uint32_t*)s_bins, tid, nthr);
        else
            for (uint32_t i = tid; i < K; i += nthr) top_assign_cluster(t, tb, level, i);
        grid_barrier(bar);
        top_trace(trace, ntrace);
    }
    if (trace && tid == 0) trace[0] = ntrace;
}
// the nodes that left the level loop with 2..HL_TOP_SMALL clusters: one thread each (own launch with few threads
// per SM, so that the per-thread work arrays stay in the L1).  Their binary node ids: one allocation per warp, a
// prefix sum over the lanes' needs.
__global__ void __launch_bounds__(128) k_top_small(BinaryTree t, TopBuild tb)
{
    const uint32_t K = *tb.n_clusters;
    if (K < 2u || K > tb.k_cap) return;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    const uint32_t small = *tb.small_count;
    const uint32_t lane  = threadIdx.x & 31u;
    for (uint32_t r0 = tid - lane; r0 < small; r0 += nthr)
    {
        const uint32_t r    = r0 + lane;
        const uint32_t need = r < small ? top_small_node_ids(tb, r) : 0u;
        uint32_t       incl = need;
        for (int d = 1; d < 32; d <<= 1)
        {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((int)lane >= d) incl += v;
        }
        uint32_t base = 0;
        if (lane == 31u) base = atomicAdd(tb.free_next, incl);
        base = __shfl_sync(0xFFFFFFFFu, base, 31);
        if (r < small) top_small_node(t, tb, r, base + incl - need);
    }
}
__global__ void k_top_refit(BinaryTree t, TopBuild tb)
{
    // (no bail-out here: k_fit stopped at the cut, so the nodes above it are fitted by this pass whether k_top_build re-linked
    //  them or — cut too fine for the scratch arrays — left the radix tree's topology in place)
    const uint32_t K = *tb.n_clusters;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < K; i += gridDim.x * blockDim.x) top_refit_from_cluster(t, tb.cluster[i], DeviceFence());
}

// ---- two-level re-split (round 2): treelets built inside one thread block ---------------------------------------------
// The level-synchronous re-split above moves every cluster through global memory once per level and phase (BIN / ASSIGN:
// ~8 L2 atomics per cluster, 4 grid barriers per level, ~20 levels): 2.2 of the 3.5 ms of a 1M-triangle build.  Here the
// radix tree is cut TWICE.  The level loop runs over the COARSE cut (subtrees of <= C_A primitives, a few ten thousand of
// them) and stops at nodes of <= HL_TREELET_PRIMS primitives ("treelets", hl_build.h TopTreelet).  Every treelet is then
// re-split down to the FINE cut (clusters of <= C primitives) by ONE block entirely in shared memory: the fine clusters below
// the treelet's coarse subtrees are loaded once (boxes, primitive counts, node ids), the binned-SAH recursion runs level by
// level over index ranges of a permutation — one warp per range: centroid bounds by warp reductions, 16 bins along the longest
// axis in shared memory, plane choice by the shuffle scans of top_choose_node_warp, stable partition by ballot prefix —,
// ranges of <= HL_TOP_SMALL clusters are finished by one thread each with exact SAH sweeps (as top_small_node does), the
// re-linked nodes reuse the treelet's own node ids (the free-list ids the level loop did not use + the radix-tree nodes
// between the two cuts) and are fitted (box, primitive count, collapse cost table) bottom-up by the same block, level by
// level, without atomics or fences.  Same planes as the one-level re-split except that, above the treelets, a plane cannot cut
// through a coarse subtree.  Everything is a function of the input: items in (cluster index, Morton) order, node ids by
// (level, range index), ties by (axis, bin) — the tree is the same run to run.
#ifndef HL_TREELET_PRIMS
#define HL_TREELET_PRIMS 512u /* 44 KB of shared memory per block of 128 threads: 5 treelets in flight per SM.  Measured (tools/tune_trace.py set "treelet"): 256 / 512 / 1024 primitives give SAH 123.7 / 122.7 / 122.5 on the 5M-triangle foliage mesh (one-level re-split: 120.6) and the same frame times within the run-to-run noise */
#endif
#ifndef HL_TREELET_THREADS
#define HL_TREELET_THREADS 128
#endif
#define HL_TREELET_WARPS (HL_TREELET_THREADS / 32)
#define HL_TREELET_PER_THREAD (HL_TREELET_PRIMS / HL_TREELET_THREADS)
#ifndef HL_TREELET_TINY
#define HL_TREELET_TINY 3u   /* ranges of 2..3 clusters are finished by one thread (every partition is evaluated, from registers) */
#endif
#ifndef HL_TREELET_EXACT
#define HL_TREELET_EXACT 8u  /* ranges of at most this many clusters (<= 8: lane = 8 axis + candidate): exact SAH over all three axes by one warp, a candidate plane per lane (the rule of top_small_node); larger ones: 16 bins along the longest axis */
#endif
#ifndef HL_TREELET_ROWS
#define HL_TREELET_ROWS 160u /* warp-built nodes whose children's rows are kept in shared memory (a node per > HL_TOP_SMALL clusters: ~60 in a balanced treelet); a treelet with more of them — a long chain of lopsided splits — is fitted through global memory */
#endif
// first fit, fine clusters only: the thread of a cluster root walks its subtree in post-order (stackless: parent links) and
// writes leaf boxes, boxes and cost tables — no arrival counters, no fences (the atomic bottom-up pass took 46 of the 80 ms of
// a 50M-triangle build)
__global__ void k_fit_fine(BinaryTree t, const Box* prim_boxes, const uint32_t* sorted, uint32_t C)
{
    const uint32_t leaf0 = t.n - 1;
    // one thread per LEAF (its parent chain is read with neighbouring threads: coalesced; a thread per node re-read five
    // scattered words per node to find the cluster roots — 5.8 ms at 50M triangles): the thread of a cluster's first leaf fits it
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < t.n; j += gridDim.x * blockDim.x)
    {
        uint32_t m = leaf0 + j;
        for (;;)
        {
            const uint32_t p = t.parent[m];
            if (p == 0xFFFFFFFFu || subtree_prims(t, p) > C) break;
            m = p;
        }
        if (subtree_first(t, m) != j) continue;
        uint32_t node = m;
        while (node < leaf0) node = t.left[node];
        for (;;)
        {
            if (node >= leaf0)
            {
                const Box b = prim_boxes[sorted[node - leaf0]];
                t.box[node] = b;
                sah_leaf_costs(t, node, box_half_area(b));
            }
            if (node == m) break;
            const uint32_t p = t.parent[node];
            if (t.left[p] == node)
            {
                node = t.right[p];
                while (node < leaf0) node = t.left[node];
                continue;
            }
            const Box b = box_union(t.box[t.left[p]], t.box[t.right[p]]);
            t.box[p]    = b;
            sah_node_costs(t, p, box_half_area(b));
            node = p;
        }
    }
}
// boxes of the coarse subtrees' roots (what the level loop bins): union of the leaf boxes k_fit_fine wrote.  The nodes between
// the two cuts are not fitted here — the treelets re-link them.
__global__ void k_coarse_boxes(BinaryTree t, const uint32_t* cluster, const uint32_t* n_clusters, uint32_t C)
{
    const uint32_t K = *n_clusters, leaf0 = t.n - 1;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < K; i += gridDim.x * blockDim.x)
    {
        const uint32_t R = cluster[i];
        if (R >= leaf0 || subtree_prims(t, R) <= C) continue; // fitted by k_fit_fine
        const uint32_t f = t.first[R], l = t.last[R];
        Box            b = t.box[leaf0 + f];
        for (uint32_t j = f + 1u; j <= l; j++) b = box_union(b, t.box[leaf0 + j]);
        t.box[R] = b;
    }
}
struct TreeletSmem
{
    float    lo[3][HL_TREELET_PRIMS], hi[3][HL_TREELET_PRIMS]; // item boxes
    uint32_t prims[HL_TREELET_PRIMS];                          // item primitive counts
    uint32_t node[HL_TREELET_PRIMS];                           // item = fine cluster: its binary node id
    uint32_t ids[HL_TREELET_PRIMS];                            // node ids the treelet may use; [0] becomes its root
    uint16_t perm[2][HL_TREELET_PRIMS];
    uint16_t seg_start[2][HL_TREELET_PRIMS / 2 + 2], seg_cnt[2][HL_TREELET_PRIMS / 2 + 2];
    uint32_t seg_link[2][HL_TREELET_PRIMS / 2 + 2];
    uint16_t seg_left[HL_TREELET_PRIMS / 2 + 2];   // items that go left, per range of the current level
    uint16_t level_base[HL_TREELET_PRIMS + 2];     // ids[level_base[L] ...] = the nodes created at level L
    uint16_t small_start[HL_TREELET_PRIMS / 2 + 2], small_cnt[HL_TREELET_PRIMS / 2 + 2]; // ranges finished by one thread
    uint32_t small_link[HL_TREELET_PRIMS / 2 + 2];
    uint8_t  small_buf[HL_TREELET_PRIMS / 2 + 2];  // which perm buffer holds the range
    uint16_t small_parent[HL_TREELET_PRIMS / 2 + 2]; // local index (position in ids) of the warp-built node above the range
    uint16_t seg_parent[2][HL_TREELET_PRIMS / 2 + 2];
    // bottom-up in shared memory: the cost-table rows of the two children of every warp-built node (filled by whoever finishes
    // the child: the scan stage for single clusters, treelet_small for small ranges, the level pass for warp-built nodes)
    union
    {
        struct
        {
            uint32_t sub_root[HL_TREELET_PRIMS];    // coarse subtrees of the treelet (sorted by cluster index): root node,
            uint32_t sub_off[HL_TREELET_PRIMS + 1]; //   first position in the concatenated leaf ranges (dead once the items are collected)
        };
        float child_row[HL_TREELET_ROWS][2][7];
    };
    float    node_area[HL_TREELET_ROWS];
    uint32_t node_prims[HL_TREELET_ROWS];
    uint16_t node_parent[HL_TREELET_ROWS]; // local index of the parent | side << 15; 0xFFFF = the treelet's root
    alignas(16) TopBin bins[HL_TREELET_WARPS][HL_TOP_BINS]; // (top_clear_bin stores 16 bytes at a time)
    uint32_t scan[HL_TREELET_WARPS + 1];
    uint32_t n_items, n_ids, n_next;
};
// exclusive prefix sum over one value per thread of the block (blockDim.x == HL_TREELET_THREADS); returns the total in `total`
__device__ __forceinline__ uint32_t treelet_block_scan(uint32_t v, uint32_t* warp_sums, uint32_t& total)
{
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t       incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((int)lane >= d) incl += o;
    }
    __syncthreads(); // (warp_sums may still be read by the previous call)
    if (lane == 31u) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t base = 0, sum = 0;
#pragma unroll
    for (uint32_t w = 0; w < HL_TREELET_WARPS; w++)
    {
        const uint32_t x = warp_sums[w];
        if (w < warp) base += x;
        sum += x;
    }
    total = sum;
    return base + incl - v;
}
// plane choice of one range by one warp from the 16 bins of ONE axis in shared memory: same selection as top_choose_node_warp
__device__ __forceinline__ bool treelet_choose(const TopBin* bins, uint32_t lane, uint32_t& split_bin, uint32_t& n_left)
{
    const unsigned FULL = 0xFFFFFFFFu;
    float          best = hl_inf();
    split_bin = 0xFFFFFFFFu, n_left = 0u;
    float    lo[3], hi[3];
    uint32_t p = 0u, c = 0u;
    for (int k = 0; k < 3; k++) lo[k] = hl_inf(), hi[k] = -hl_inf();
    if (lane < HL_TOP_BINS)
    {
        const TopBin B = bins[lane];
        for (int k = 0; k < 3; k++) lo[k] = ord2f(B.lo[k]), hi[k] = ord2f(B.hi[k]);
        p = B.prims, c = B.clusters;
    }
    float    plo[3] = { lo[0], lo[1], lo[2] }, phi[3] = { hi[0], hi[1], hi[2] }, slo[3] = { lo[0], lo[1], lo[2] }, shi[3] = { hi[0], hi[1], hi[2] };
    uint32_t pp = p, pc = c, sp = p, sc = c;
#pragma unroll
    for (int d = 1; d < HL_TOP_BINS; d <<= 1)
    {
        const bool up = (int)lane >= d, dn = lane + d < 32u; // (lanes >= 16 hold the identity)
#pragma unroll
        for (int k = 0; k < 3; k++)
        {
            const float a = __shfl_up_sync(FULL, plo[k], d), b = __shfl_up_sync(FULL, phi[k], d);
            const float e = __shfl_down_sync(FULL, slo[k], d), f = __shfl_down_sync(FULL, shi[k], d);
            if (up) plo[k] = fminf(plo[k], a), phi[k] = fmaxf(phi[k], b);
            if (dn) slo[k] = fminf(slo[k], e), shi[k] = fmaxf(shi[k], f);
        }
        const uint32_t a = __shfl_up_sync(FULL, pp, d), b = __shfl_up_sync(FULL, pc, d);
        const uint32_t e = __shfl_down_sync(FULL, sp, d), f = __shfl_down_sync(FULL, sc, d);
        if (up) pp += a, pc += b;
        if (dn) sp += e, sc += f;
    }
    const float    sarea = (shi[0] - slo[0]) * (shi[1] - slo[1]) + (shi[1] - slo[1]) * (shi[2] - slo[2]) + (shi[2] - slo[2]) * (shi[0] - slo[0]);
    const float    rarea = __shfl_down_sync(FULL, sarea, 1);
    const uint32_t rp = __shfl_down_sync(FULL, sp, 1), rc = __shfl_down_sync(FULL, sc, 1);
    if (lane < HL_TOP_BINS - 1u && pc != 0u && rc != 0u)
    {
        const float parea = (phi[0] - plo[0]) * (phi[1] - plo[1]) + (phi[1] - plo[1]) * (phi[2] - plo[2]) + (phi[2] - plo[2]) * (phi[0] - plo[0]);
        best = parea * (float)pp + rarea * (float)rp, split_bin = lane, n_left = pc;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        const float    ob = __shfl_xor_sync(FULL, best, d);
        const uint32_t os = __shfl_xor_sync(FULL, split_bin, d), on = __shfl_xor_sync(FULL, n_left, d);
        if (ob < best || (ob == best && os < split_bin)) best = ob, split_bin = os, n_left = on;
    }
    return best < hl_inf();
}
// a range of 2 or 3 items, by ONE thread, from registers: for three items every partition (one item against the other two) is
// evaluated — a superset of the sorted-sweep candidates of top_small_node.  Links the nodes, writes their boxes / counts / cost
// tables; `up` (may be null) receives the row of the range's root.
struct TinyItem
{
    Box      b;
    uint32_t prims, node;
    float    row[8]; // [1..7]
};
__device__ __forceinline__ TinyItem tiny_load(const BinaryTree& t, const TreeletSmem& S, uint32_t it)
{
    TinyItem r;
    for (int q = 0; q < 3; q++) r.b.lo[q] = S.lo[q][it], r.b.hi[q] = S.hi[q][it];
    r.prims = S.prims[it], r.node = S.node[it];
    r.row[0] = hl_inf();
    const float* c = t.cost + (size_t)r.node * 7; // (k_fit_fine's row)
    for (int i = 0; i < 7; i++) r.row[i + 1] = c[i];
    return r;
}
// node `self` over children x (left) and y (right), both finished; returns the node as an item
__device__ __forceinline__ TinyItem tiny_join(BinaryTree& t, uint32_t self, uint32_t link, const TinyItem& x, const TinyItem& y)
{
    TinyItem r;
    top_link(t, link, self);
    top_link(t, self << 1, x.node);
    top_link(t, (self << 1) | 1u, y.node);
    r.b = box_union(x.b, y.b);
    const uint32_t p = x.prims + y.prims;
    r.prims = p > HL_MAX_LEAF_PRIMS ? p : HL_MAX_LEAF_PRIMS + 1u, r.node = self;
    t.box[self]   = r.b;
    t.first[self] = 0u, t.last[self] = r.prims - 1u;
    r.row[0] = hl_inf();
    sah_node_costs_rows(t, self, box_half_area(r.b), r.prims, x.row, y.row, r.row + 1);
    return r;
}
__device__ void treelet_tiny(BinaryTree& t, const TreeletSmem& S, const uint16_t* pin, uint32_t count, uint32_t link, uint32_t id0, float* up)
{
    const TinyItem a = tiny_load(t, S, pin[0]), b = tiny_load(t, S, pin[1]);
    TinyItem       root;
    if (count == 2u)
        root = tiny_join(t, S.ids[id0], link, a, b);
    else
    {
        const TinyItem c = tiny_load(t, S, pin[2]);
        // item k alone against the other two
        const float c0 = box_half_area(a.b) * (float)a.prims + box_half_area(box_union(b.b, c.b)) * (float)(b.prims + c.prims);
        const float c1 = box_half_area(b.b) * (float)b.prims + box_half_area(box_union(a.b, c.b)) * (float)(a.prims + c.prims);
        const float c2 = box_half_area(c.b) * (float)c.prims + box_half_area(box_union(a.b, b.b)) * (float)(a.prims + b.prims);
        const int   k  = c0 <= c1 && c0 <= c2 ? 0 : (c1 <= c2 ? 1 : 2);
        const TinyItem& one = k == 0 ? a : (k == 1 ? b : c);
        const TinyItem& p0  = k == 0 ? b : a;
        const TinyItem& p1  = k == 2 ? b : c;
        const uint32_t  self = S.ids[id0], pair = S.ids[id0 + 1u];
        // (the pair's link is written before the root's own: top_link only touches the child's parent entry and the parent's slot)
        const TinyItem two = tiny_join(t, pair, (self << 1) | 1u, p0, p1);
        root               = tiny_join(t, self, link, one, two);
    }
    if (up)
        for (int i = 0; i < 7; i++) up[i] = root.row[i + 1];
}
// exact SAH split of a range of <= 8 items by one warp: lane 8 a + j evaluates, along axis a, the plane that puts the items
// before item j — in (centroid, position) order — on the left; the cheapest of the 3 x (count - 1) planes wins, ties by (axis,
// position).  Returns false when no plane separates the items (only when no centroid compares: non-finite boxes).  right = the
// item at position `lane` goes right.
__device__ __forceinline__ bool treelet_exact(const TreeletSmem& S, const uint16_t* pin, uint32_t cnt, uint32_t lane, bool& right, uint32_t& n_left)
{
    const unsigned FULL = 0xFFFFFFFFu;
    const uint32_t a = lane >> 3, j = lane & 7u;
    const bool     cand = a < 3u && j < cnt;
    const uint32_t ax = cand ? a : 0u, me = pin[cand ? j : 0u];
    const float    cj = S.lo[ax][me] + S.hi[ax][me];
    float          llo[3], lhi[3], rlo[3], rhi[3];
    uint32_t       lp = 0u, rp = 0u, lc = 0u;
    for (int q = 0; q < 3; q++) llo[q] = rlo[q] = hl_inf(), lhi[q] = rhi[q] = -hl_inf();
    for (uint32_t i = 0; i < cnt; i++)
    {
        const uint32_t it = pin[i];
        const float    ci = S.lo[ax][it] + S.hi[ax][it];
        const bool     left = ci < cj || (ci == cj && i < j);
        const uint32_t pr = S.prims[it];
#pragma unroll
        for (int q = 0; q < 3; q++)
        {
            const float lo = S.lo[q][it], hi = S.hi[q][it];
            llo[q] = left ? fminf(llo[q], lo) : llo[q], lhi[q] = left ? fmaxf(lhi[q], hi) : lhi[q];
            rlo[q] = left ? rlo[q] : fminf(rlo[q], lo), rhi[q] = left ? rhi[q] : fmaxf(rhi[q], hi);
        }
        lp += left ? pr : 0u, rp += left ? 0u : pr, lc += left ? 1u : 0u;
    }
    float    best = hl_inf();
    uint32_t key  = 0xFFFFFFFFu; // axis << 8 | position
    if (cand && lc != 0u)
    {
        const float la = (lhi[0] - llo[0]) * (lhi[1] - llo[1]) + (lhi[1] - llo[1]) * (lhi[2] - llo[2]) + (lhi[2] - llo[2]) * (lhi[0] - llo[0]);
        const float ra = (rhi[0] - rlo[0]) * (rhi[1] - rlo[1]) + (rhi[1] - rlo[1]) * (rhi[2] - rlo[2]) + (rhi[2] - rlo[2]) * (rhi[0] - rlo[0]);
        const float c  = la * (float)lp + ra * (float)rp;
        if (c < hl_inf()) best = c, key = (a << 8) | j;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        const float    ob = __shfl_xor_sync(FULL, best, d);
        const uint32_t ok = __shfl_xor_sync(FULL, key, d);
        if (ob < best || (ob == best && ok < key)) best = ob, key = ok;
    }
    if (key == 0xFFFFFFFFu) return false;
    const uint32_t wa = key >> 8, wj = key & 0xFFu;
    const bool     in = lane < cnt;
    const uint32_t it = pin[in ? lane : 0u], cut_it = pin[wj];
    const float    mine = S.lo[wa][it] + S.hi[wa][it], cut = S.lo[wa][cut_it] + S.hi[wa][cut_it];
    right  = in && !(mine < cut || (mine == cut && lane < wj));
    n_left = (uint32_t)__popc(__ballot_sync(FULL, in && !right));
    return true;
}
// err: set to 1 when a treelet does not fit the shared arrays (cannot happen while the primitive counts of the level loop are
// upper bounds; checked by the host)
// cursor: the next treelet to hand out (zeroed by the host): treelets hold 2 .. 512 primitives, so the blocks take them one at a
// time instead of a fixed stride
__global__ void __launch_bounds__(HL_TREELET_THREADS) k_treelets(BinaryTree t, TopBuild tb, uint32_t C, uint32_t* err, uint32_t* cursor)
{
    extern __shared__ __align__(16) unsigned char treelet_raw[];
    TreeletSmem&   S     = *(TreeletSmem*)treelet_raw;
    const uint32_t tid   = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t leaf0 = t.n - 1;
    const uint32_t count = *tb.treelet_count;
    for (;;)
    {
        __syncthreads(); // the previous treelet is finished with the shared arrays (and with n_next)
        if (tid == 0) S.n_next = atomicAdd(cursor, 1u);
        __syncthreads();
        const uint32_t ti = S.n_next;
        if (ti >= count) break;
        TopTreelet&    rec = tb.treelet[ti];
        const uint32_t m   = rec.count, link = rec.link;
        // ---- the treelet's coarse subtrees, sorted by cluster index (they registered in arrival order)
        {
            const uint32_t* list = tb.tlist + rec.list_base;
            uint32_t        sizes = 0; // (this thread's subtrees: the sum of their sizes)
            for (uint32_t k = tid; k < m; k += HL_TREELET_THREADS)
            {
                const uint32_t v = list[k];
                uint32_t       rank = 0;
                for (uint32_t j = 0; j < m; j++) rank += list[j] < v ? 1u : 0u;
                S.sub_root[rank] = tb.cluster[v];
            }
            __syncthreads();
            // exclusive scan of the subtree sizes, blocked: thread tid owns entries [tid * per, tid * per + per)
            const uint32_t per = (m + HL_TREELET_THREADS - 1u) / HL_TREELET_THREADS;
            for (uint32_t k = tid * per; k < min(m, tid * per + per); k++) sizes += subtree_prims(t, S.sub_root[k]);
            uint32_t total;
            uint32_t at = treelet_block_scan(sizes, S.scan, total);
            for (uint32_t k = tid * per; k < min(m, tid * per + per); k++) S.sub_off[k] = at, at += subtree_prims(t, S.sub_root[k]);
            if (tid == 0) S.sub_off[m] = total;
            __syncthreads();
        }
        const uint32_t P = S.sub_off[m];
        if (P > HL_TREELET_PRIMS)
        {
            if (tid == 0) *err = 1u;
            continue; // (uniform)
        }
        // ---- collect, in (subtree, Morton) order: the fine clusters (items) and the node ids between the two cuts
        {
            uint32_t root_of[HL_TREELET_PER_THREAD], flag[HL_TREELET_PER_THREAD], idf[HL_TREELET_PER_THREAD];
            uint32_t mine = 0, mine_ids = 0;
#pragma unroll
            for (uint32_t k = 0; k < HL_TREELET_PER_THREAD; k++)
            {
                // (thread tid owns positions 4 tid .. 4 tid + 3: a blocked arrangement keeps the scan order = position order)
                const uint32_t j = tid * HL_TREELET_PER_THREAD + k;
                flag[k] = 0u, root_of[k] = 0u, idf[k] = 0xFFFFFFFFu;
                if (j < P)
                {
                    uint32_t a = 0, b = m; // subtree s with sub_off[s] <= j < sub_off[s + 1]
                    while (b - a > 1u)
                    {
                        const uint32_t c = (a + b) >> 1;
                        if (S.sub_off[c] <= j) a = c; else b = c;
                    }
                    const uint32_t Rs = S.sub_root[a], off = j - S.sub_off[a];
                    const uint32_t fs = subtree_first(t, Rs), ls = fs + subtree_prims(t, Rs) - 1u;
                    uint32_t       mm = leaf0 + fs + off;
                    while (mm != Rs) // (never above the subtree's root: the nodes up there may be re-linked by another block right now)
                    {
                        const uint32_t pm = t.parent[mm];
                        if (subtree_prims(t, pm) > C) break;
                        mm = pm;
                    }
                    if (subtree_first(t, mm) == fs + off)
                    {
                        flag[k] = 1u, root_of[k] = mm;
                        // the node ids between the two cuts: one per pair of adjacent clusters of a subtree — their lowest common
                        // ancestor, i.e. the first ancestor of the left one that reaches further right (all inside the subtree:
                        // radix-tree ranges nobody else touches)
                        const uint32_t last_m = fs + off + subtree_prims(t, mm) - 1u;
                        uint32_t       x = mm;
                        while (x != Rs)
                        {
                            x = t.parent[x];
                            if (t.last[x] > last_m)
                            {
                                idf[k] = x;
                                break;
                            }
                        }
                    }
                    (void)ls;
                }
                mine += flag[k], mine_ids += idf[k] != 0xFFFFFFFFu ? 1u : 0u;
            }
            uint32_t total;
            uint32_t at = treelet_block_scan(mine, S.scan, total);
#pragma unroll
            for (uint32_t k = 0; k < HL_TREELET_PER_THREAD; k++)
                if (flag[k])
                {
                    const uint32_t mm = root_of[k];
                    const Box      b  = t.box[mm];
                    for (int a = 0; a < 3; a++) S.lo[a][at] = b.lo[a], S.hi[a][at] = b.hi[a];
                    S.prims[at] = subtree_prims(t, mm), S.node[at] = mm, S.perm[0][at] = (uint16_t)at;
                    at++;
                }
            if (tid == 0) S.n_items = total;
            const uint32_t own = m - 1u; // ids of the free list
            at = own + treelet_block_scan(mine_ids, S.scan, total);
#pragma unroll
            for (uint32_t k = 0; k < HL_TREELET_PER_THREAD; k++)
                if (idf[k] != 0xFFFFFFFFu) S.ids[at++] = idf[k];
            for (uint32_t k = tid; k < own; k += HL_TREELET_THREADS) S.ids[k] = tb.free_nodes[rec.id_base + k];
            if (tid == 0) S.n_ids = own + total, S.seg_start[0][0] = 0, S.seg_link[0][0] = link, S.seg_parent[0][0] = 0xFFFFu, S.level_base[0] = 0;
            __syncthreads();
            // a treelet of ONE coarse subtree keeps that subtree's root as its root (the whole tree's root is node 0)
            if (own == 0u)
                for (uint32_t k = tid; k < total; k += HL_TREELET_THREADS)
                    if (S.ids[k] == S.sub_root[0] && k != 0u) S.ids[k] = S.ids[0], S.ids[0] = S.sub_root[0]; // (at most one k matches)
        }
        __syncthreads();
        const uint32_t n_items = S.n_items;
        if (S.n_ids + 1u != n_items)
        {
            if (tid == 0) *err = 2u;
            continue; // (uniform)
        }
        if (n_items == 1u)
        {
            if (tid == 0) top_link(t, link, S.node[0]), rec.root = S.node[0];
            continue; // (uniform)
        }
        if (tid == 0) rec.root = S.ids[0];
        // ---- top-down: one level per iteration, one warp per range; ranges of <= HL_TREELET_TINY items are set aside
        uint32_t nseg = 0u, nsmall = 0u, level = 0u, base = 0u; // base = ids consumed by the levels above
        uint32_t cur  = 0u;
        if (n_items <= HL_TREELET_TINY)
        {
            if (tid == 0) S.small_start[0] = 0, S.small_cnt[0] = (uint16_t)n_items, S.small_link[0] = link, S.small_buf[0] = 0, S.small_parent[0] = 0xFFFFu;
            nsmall = 1u;
        }
        else
        {
            if (tid == 0) S.seg_cnt[0][0] = (uint16_t)n_items;
            nseg = 1u;
        }
        __syncthreads();
        while (nseg)
        {
            for (uint32_t k = warp; k < nseg; k += HL_TREELET_WARPS)
            {
                const uint32_t  s0 = S.seg_start[cur][k], cnt = S.seg_cnt[cur][k];
                const uint16_t* pin = S.perm[cur] + s0;
                uint16_t*       pout = S.perm[cur ^ 1u] + s0;
                // centroid bounds, and the node's own box and primitive count
                float    clo[3] = { hl_inf(), hl_inf(), hl_inf() }, chi[3] = { -hl_inf(), -hl_inf(), -hl_inf() };
                float    blo[3] = { hl_inf(), hl_inf(), hl_inf() }, bhi[3] = { -hl_inf(), -hl_inf(), -hl_inf() };
                uint32_t psum = 0u;
                for (uint32_t i = lane; i < cnt; i += 32u)
                {
                    const uint32_t it = pin[i];
                    for (int a = 0; a < 3; a++)
                    {
                        const float l = S.lo[a][it], h = S.hi[a][it], c = 0.5f * (l + h);
                        clo[a] = fminf(clo[a], c), chi[a] = fmaxf(chi[a], c);
                        blo[a] = fminf(blo[a], l), bhi[a] = fmaxf(bhi[a], h);
                    }
                    psum += S.prims[it];
                }
                float cbl[3], cbh[3];
                Box   nb;
                for (int a = 0; a < 3; a++)
                {
                    cbl[a] = ord2f(__reduce_min_sync(0xFFFFFFFFu, f2ord(clo[a]))), cbh[a] = ord2f(__reduce_max_sync(0xFFFFFFFFu, f2ord(chi[a])));
                    nb.lo[a] = ord2f(__reduce_min_sync(0xFFFFFFFFu, f2ord(blo[a]))), nb.hi[a] = ord2f(__reduce_max_sync(0xFFFFFFFFu, f2ord(bhi[a])));
                }
                psum = __reduce_add_sync(0xFFFFFFFFu, psum);
                uint32_t n_left = 0u;
                if (cnt <= HL_TREELET_EXACT)
                {
                    bool       right  = false;
                    const bool planar = treelet_exact(S, pin, cnt, lane, right, n_left);
                    if (!planar) n_left = cnt / 2u, right = lane >= n_left; // (cannot happen for finite boxes)
                    const bool     in = lane < cnt;
                    const uint32_t mr = __ballot_sync(0xFFFFFFFFu, in && right), ml = __ballot_sync(0xFFFFFFFFu, in && !right);
                    const uint32_t lt = (1u << lane) - 1u;
                    if (in) pout[right ? n_left + (uint32_t)__popc(mr & lt) : (uint32_t)__popc(ml & lt)] = pin[lane];
                }
                else
                {
                    const float e0 = cbh[0] - cbl[0], e1 = cbh[1] - cbl[1], e2 = cbh[2] - cbl[2];
                    const int   ax = e0 >= e1 && e0 >= e2 ? 0 : (e1 >= e2 ? 1 : 2); // longest axis of the centroid bounds (top_bin_axis)
                    const float al = ax == 0 ? cbl[0] : (ax == 1 ? cbl[1] : cbl[2]), ah = ax == 0 ? cbh[0] : (ax == 1 ? cbh[1] : cbh[2]);
                    TopBin*     bins = S.bins[warp];
                    if (lane < HL_TOP_BINS) top_clear_bin(bins + lane);
                    __syncwarp();
                    for (uint32_t i = lane; i < cnt; i += 32u)
                    {
                        const uint32_t it = pin[i];
                        const int      bi = top_bin_of(0.5f * (S.lo[ax][it] + S.hi[ax][it]), al, ah);
                        if (bi < 0) continue;
                        TopBin& B = bins[bi];
                        for (int q = 0; q < 3; q++) atomicMin(&B.lo[q], f2ord(S.lo[q][it])), atomicMax(&B.hi[q], f2ord(S.hi[q][it]));
                        atomicAdd(&B.prims, S.prims[it]), atomicAdd(&B.clusters, 1u);
                    }
                    __syncwarp();
                    uint32_t   sbin;
                    const bool planar = treelet_choose(bins, lane, sbin, n_left);
                    __syncwarp();
                    if (!planar) n_left = cnt / 2u; // all centroids equal: first half left
                    uint32_t nl = 0u, nr = 0u;
                    for (uint32_t i0 = 0; i0 < cnt; i0 += 32u)
                    {
                        const uint32_t i  = i0 + lane;
                        const bool     in = i < cnt;
                        uint32_t       it = 0u;
                        bool           right = false;
                        if (in)
                        {
                            it    = pin[i];
                            right = planar ? top_bin_of(0.5f * (S.lo[ax][it] + S.hi[ax][it]), al, ah) > (int)sbin : i >= n_left;
                        }
                        const uint32_t mr = __ballot_sync(0xFFFFFFFFu, in && right), ml = __ballot_sync(0xFFFFFFFFu, in && !right);
                        const uint32_t lt = (1u << lane) - 1u;
                        if (in) pout[right ? n_left + nr + (uint32_t)__popc(mr & lt) : nl + (uint32_t)__popc(ml & lt)] = (uint16_t)it;
                        nl += (uint32_t)__popc(ml), nr += (uint32_t)__popc(mr);
                    }
                }
                if (lane == 0u)
                {
                    const uint32_t self = S.ids[base + k], idx = base + k;
                    const uint32_t pf   = psum > HL_MAX_LEAF_PRIMS ? psum : HL_MAX_LEAF_PRIMS + 1u;
                    S.seg_left[k] = (uint16_t)n_left;
                    top_link(t, S.seg_link[cur][k], self);
                    t.box[self]   = nb;
                    t.first[self] = 0u, t.last[self] = pf - 1u;
                    if (idx < HL_TREELET_ROWS) S.node_area[idx] = box_half_area(nb), S.node_prims[idx] = pf, S.node_parent[idx] = (uint16_t)(S.seg_parent[cur][k] == 0xFFFFu ? 0xFFFFu : (S.seg_parent[cur][k] | ((S.seg_link[cur][k] & 1u) << 15)));
                }
            }
            __syncthreads();
            // the ranges of the next level and the small ranges (exclusive scans over the ranges of this one); single items link themselves
            uint32_t total_next = 0u;
            for (uint32_t k0 = 0; k0 < nseg; k0 += HL_TREELET_THREADS)
            {
                const uint32_t k = k0 + tid;
                uint32_t       s0 = 0, cnt = 0, nl = 0, self = 0, need = 0, need_small = 0;
                if (k < nseg)
                {
                    s0 = S.seg_start[cur][k], cnt = S.seg_cnt[cur][k], nl = S.seg_left[k], self = S.ids[base + k];
                    need       = (nl > HL_TREELET_TINY ? 1u : 0u) + (cnt - nl > HL_TREELET_TINY ? 1u : 0u);
                    need_small = (nl >= 2u && nl <= HL_TREELET_TINY ? 1u : 0u) + (cnt - nl >= 2u && cnt - nl <= HL_TREELET_TINY ? 1u : 0u);
                }
                uint32_t tot, tot_small;
                uint32_t at  = total_next + treelet_block_scan(need, S.scan, tot);
                uint32_t ats = nsmall + treelet_block_scan(need_small, S.scan, tot_small);
                if (k < nseg)
                {
                    const uint16_t* pnew = S.perm[cur ^ 1u];
                    for (uint32_t side = 0; side < 2u; side++)
                    {
                        const uint32_t cs = side ? s0 + nl : s0, cc = side ? cnt - nl : nl, lk = (self << 1) | side;
                        if (cc == 1u)
                        {
                            const uint32_t item_node = S.node[pnew[cs]];
                            top_link(t, lk, item_node);
                            if (base + k < HL_TREELET_ROWS)
                                for (int i = 0; i < 7; i++) S.child_row[base + k][side][i] = t.cost[(size_t)item_node * 7 + i]; // (k_fit_fine's row)
                        }
                        else if (cc <= HL_TREELET_TINY)
                            S.small_start[ats] = (uint16_t)cs, S.small_cnt[ats] = (uint16_t)cc, S.small_link[ats] = lk, S.small_buf[ats] = (uint8_t)(cur ^ 1u), S.small_parent[ats] = (uint16_t)(base + k), ats++;
                        else
                            S.seg_start[cur ^ 1u][at] = (uint16_t)cs, S.seg_cnt[cur ^ 1u][at] = (uint16_t)cc, S.seg_link[cur ^ 1u][at] = lk, S.seg_parent[cur ^ 1u][at] = (uint16_t)(base + k), at++;
                    }
                }
                total_next += tot, nsmall += tot_small;
            }
            base += nseg, level++;
            if (tid == 0) S.level_base[level] = (uint16_t)base;
            nseg = total_next, cur ^= 1u;
            __syncthreads();
        }
        // ---- the small ranges: ids by an exclusive scan over (count - 1), one thread per range (build + fit).  Range k goes to
        // lane k / warps of warp k % warps: the ranges are spread over ALL warps of the block (the threads run long, divergent
        // instruction streams: with ranges 0..31 on warp 0 the other warps waited at the barrier below for 40 % of the kernel)
        {
            uint32_t run = base;
            for (uint32_t k0 = 0; k0 < nsmall; k0 += HL_TREELET_THREADS)
            {
                const uint32_t k    = k0 + tid;
                const uint32_t need = k < nsmall ? (uint32_t)S.small_cnt[k] - 1u : 0u;
                uint32_t       tot;
                const uint32_t at = run + treelet_block_scan(need, S.scan, tot);
                if (k < nsmall) S.seg_start[0][k] = (uint16_t)at; // (the range lists are free now: seg_start[0] holds the id offsets)
                run += tot;
            }
            __syncthreads();
            for (uint32_t k0 = 0; k0 < nsmall; k0 += HL_TREELET_THREADS)
            {
                const uint32_t k = k0 + lane * HL_TREELET_WARPS + warp;
                if (k < nsmall)
                {
                    const uint32_t par = S.small_parent[k];
                    float*         up  = par < HL_TREELET_ROWS ? S.child_row[par][S.small_link[k] & 1u] : nullptr;
                    treelet_tiny(t, S, S.perm[S.small_buf[k]] + S.small_start[k], S.small_cnt[k], S.small_link[k], S.seg_start[0][k], up);
                }
            }
        }
        __syncthreads();
        // ---- bottom-up over the levels the warps created (boxes and counts were written on the way down): the collapse cost tables,
        // from the children's rows in shared memory; through global memory when the treelet has more warp-built nodes than rows
        if (base <= HL_TREELET_ROWS)
        {
            for (uint32_t L = level; L-- > 0u;)
            {
                const uint32_t b0 = S.level_base[L], b1 = S.level_base[L + 1u];
                for (uint32_t k = b0 + tid; k < b1; k += HL_TREELET_THREADS)
                {
                    float cl[8], cr[8], out[7];
                    cl[0] = cr[0] = hl_inf();
                    for (int i = 0; i < 7; i++) cl[i + 1] = S.child_row[k][0][i], cr[i + 1] = S.child_row[k][1][i];
                    sah_node_costs_rows(t, S.ids[k], S.node_area[k], S.node_prims[k], cl, cr, out);
                    const uint32_t par = S.node_parent[k];
                    if (par != 0xFFFFu)
                        for (int i = 0; i < 7; i++) S.child_row[par & 0x7FFFu][par >> 15][i] = out[i];
                }
                __syncthreads();
            }
        }
        else
        {
            for (uint32_t L = level; L-- > 0u;)
            {
                const uint32_t b0 = S.level_base[L], b1 = S.level_base[L + 1u];
                for (uint32_t k = b0 + tid; k < b1; k += HL_TREELET_THREADS)
                {
                    const uint32_t node = S.ids[k];
                    sah_node_costs(t, node, box_half_area(load_box_coherent(&t.box[node]))); // (children's rows past the L1 as well: load_f32_coherent)
                }
                __syncthreads();
            }
        }
    }
}
// the nodes of the level loop above the treelets: bottom-up from every treelet root
__global__ void k_top_refit_treelets(BinaryTree t, TopBuild tb)
{
    const uint32_t count = *tb.treelet_count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        const uint32_t root = tb.treelet[i].root;
        if (root != 0xFFFFFFFFu) top_refit_from_cluster(t, root, DeviceFence());
    }
}

// Collapse as ONE persistent launch over a device-side work queue (no host round trip per tree level):
// queue[0] holds the root task; a thread takes the next queue index, polls the slot until its task is published
// (a parent publishes one task per internal child while it writes its own node) and processes it.
// ctl[0] = node counter, ctl[1] = leaf counter, ctl[2] = queue tail, ctl[3] = queue head,
// ctl[4] = outstanding tasks (published - finished), ctl[5] = done flag.
template <class LeafWriter>
__global__ void k_write_leaves(LeafWriter writer, const uint32_t* leaf_pos, uint32_t n)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) writer(i, leaf_pos[i]);
}

__global__ void __launch_bounds__(128) k_collapse(BinaryTree t, WideOut out, CollapseTask* queue, uint32_t capacity, uint32_t* ctl, DeferredLeafWriter writer)
{
    volatile uint32_t* vdone = ctl + 5;
    uint32_t           idx   = 0;
    bool               have  = false;
    unsigned           nap   = 64u; // ns between two looks at an empty slot: doubles up to ~4 us (every resident thread polls: at a
                                    // fixed 64 ns the polls alone asked more of the L2 than it delivers, and the working threads'
                                    // loads queued behind them — ncu: 46 long-scoreboard stalls per issued instruction)
    for (;;)
    {
        if (!have) idx = hl_alloc_var(ctl + 3, 1u), have = true, nap = 64u; // one atomic per group of lanes that need a new queue index
        bool ready = false;
        CollapseTask task;
        task.wide = task.bnode = 0xFFFFFFFFu;
        if (idx < capacity)
        {
            const unsigned long long raw = *(volatile unsigned long long*)(queue + idx);
            task.wide = (uint32_t)raw, task.bnode = (uint32_t)(raw >> 32);
            ready = raw != ~0ull;
        }
        if (ready)
        {
            collapse_one(t, task, out, queue, ctl + 2, writer, ctl + 4); // (adds the tasks it publishes to ctl[4] before publishing them)
            have = false;
            // outstanding -= 1 per finished task (summed over the lanes that finished together: one atomic); the update that brings
            // it to zero ends the launch: every published task was counted before it became visible, so zero means none is left.
            {
                namespace cg = cooperative_groups;
                const cg::coalesced_group g = cg::coalesced_threads();
                if (g.thread_rank() == 0 && atomicSub(ctl + 4, g.size()) == g.size())
                {
                    __threadfence();
                    *vdone = 1u;
                }
            }
        }
        else if (*vdone)
            break;
        else
        {
            __nanosleep(nap); // waiting lanes yield the issue slots to the lanes of the warp that hold a task
            nap = nap < 4096u ? nap * 2u : nap;
        }
    }
}

static inline int grid_for(uint32_t n, int block, int cap) { return (int)std::min<uint64_t>(((uint64_t)n + block - 1) / block, (uint64_t)cap); }

// Generic build over `n` primitive boxes already on the device.  make_writer(sorted_prim, leaf_buffer) returns the
// leaf-record writer; leaf_bytes = size of one leaf record.
template <class MakeWriter>
static void build_wide_device(hl_context_t* ctx, const Box* d_boxes, uint32_t n, size_t leaf_bytes, WideBVHDev& out, MakeWriter make_writer)
{
    cudaStream_t st = ctx->stream;
    out.n_nodes = out.n_leaves = out.n_binary = 0;
    out.ms_build = 0.0f;
    for (int k = 0; k < 3; k++) out.root.lo[k] = out.root.hi[k] = 0.0f;
    if (n == 0)
    {
        out.nodes.alloc(sizeof(WideNode));
        out.leaves.alloc(leaf_bytes);
        return;
    }
    // ---- all allocations first: the timed region below contains device work only
    const int    cap = ctx->sm_count * 8;
    const int    rb  = grid_for(n, 256, 1024);
    const size_t ni  = n > 1 ? n - 1 : 1;
    ScratchBuf   partial, scene, keys_a, keys_b, vals_a, vals_b, cub_tmp, tree_u32, tree_box, tree_cost, big_nodes, queue, ctr, leaf_pos;
    partial.alloc(sizeof(Box) * rb, st);
    scene.alloc(sizeof(Box), st);
    keys_a.alloc(8ull * n, st), keys_b.alloc(8ull * n, st), vals_a.alloc(4ull * n, st), vals_b.alloc(4ull * n, st);
    cub::DoubleBuffer<uint64_t> dk(keys_a.as<uint64_t>(), keys_b.as<uint64_t>());
    cub::DoubleBuffer<uint32_t> dv(vals_a.as<uint32_t>(), vals_b.as<uint32_t>());
    size_t                      tmp_bytes = 0;
    HL_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, (int)n, 0, 63, st));
    cub_tmp.alloc(tmp_bytes, st);
    tree_u32.alloc(4ull * (ni * 5 + 2ull * n), st);
    tree_box.alloc(sizeof(Box) * (2ull * n), st);
    tree_cost.alloc(4ull * (7ull * 2 * n + ni), st);
    const uint32_t capacity = n + 1u; // at most one wide node per primitive
    big_nodes.alloc(sizeof(WideNode) * (size_t)n, st);
    queue.alloc(sizeof(CollapseTask) * (size_t)capacity, st);
    leaf_pos.alloc(4ull * n, st);
    ctr.alloc(32, st);
    out.leaves.alloc(leaf_bytes * (size_t)n);
    // binned-SAH re-split of the upper levels: cluster size C (triangle trees: the context option, doubled until the
    // cut has at most ~4M clusters; instance trees: single instances), scratch sized for k_cap clusters
    const bool tri_tree = leaf_bytes == sizeof(LeafTri);
    uint32_t   C        = ctx->sah_cluster == 0 ? 0u : (tri_tree ? ctx->sah_cluster : 1u);
    // two-level re-split (k_treelets): the fine cut stays at C whatever the size — nothing is kept per fine cluster —, the
    // level loop runs over the coarse cut C_A and stops at treelets.  HL_NO_TREELETS=1 (debug / A-B runs): the one-level re-split.
    // Triangle trees only: an instance tree has a few thousand large, overlapping boxes — planes that cannot cut through a coarse
    // subtree of 32 instances cost the city scene 7 % of its frame time, and its build takes 0.1 ms either way.
    const bool two_level = tri_tree && C != 0 && C <= 8u && n >= 2 && n > C && getenv("HL_NO_TREELETS") == nullptr;
    while (!two_level && C && tri_tree && n / C > (4u << 20)) C *= 2;
    const bool     resplit  = C != 0 && n >= 2 && n > C;
    // coarse cut of the two-level re-split; at most HL_TREELET_PRIMS / HL_TOP_SMALL, so that a node of <= HL_TOP_SMALL coarse
    // subtrees is always a treelet (top_init_node) and the level loop never produces SMALL nodes
    uint32_t       C_A      = std::min<uint32_t>(n > (4u << 20) ? 64u : 32u, HL_TREELET_PRIMS / HL_TOP_SMALL);
    if (const char* e = getenv("HL_COARSE_CUT")) C_A = std::min<uint32_t>(std::max(2 * C, (uint32_t)atoi(e)), HL_TREELET_PRIMS / HL_TOP_SMALL); // (tuning runs)
    const uint32_t C_top    = two_level ? C_A : C;                    // cluster size of the level-synchronous kernels
    const bool     resplit_top = resplit && n > C_top;                // (a tree of at most C_A primitives is one treelet)
    const uint32_t k_cap    = resplit_top ? (uint32_t)std::min<uint64_t>(n, 4ull * n / C_top + 1024) : (two_level ? 1u : 0u);
    const uint32_t bins_cap = k_cap / (HL_TOP_SMALL + 1u) + 2u;
    const size_t   top_ctl_words = 8 + 2 * (HL_TOP_MAX_LEVELS + 1) + 1; // (+ the treelet cursor)
    ScratchBuf     top_crec;
    ScratchBuf     top_clusters, top_free, top_cnode, top_level[2], top_bins[2], top_list, top_small, top_ctl, top_tmp, top_trace_buf, top_treelets, top_tlist;
    size_t         top_tmp_bytes = 0;
    int            top_grid      = 0;
    if (resplit)
    {
        top_ctl.alloc(4ull * top_ctl_words, st);
        size_t a = 0, b = 0;
        thrust::counting_iterator<uint32_t> ids(0u);
        top_clusters.alloc(4ull * n, st), top_free.alloc(4ull * n, st);
        if (resplit_top)
        {
            top_cnode.alloc(4ull * k_cap, st), top_crec.alloc(sizeof(TopCluster) * (size_t)k_cap, st);
            for (int p = 0; p < 2; p++)
            {
                top_level[p].alloc(sizeof(TopNode) * ((size_t)k_cap + 2), st);
                top_bins[p].alloc(sizeof(TopBin) * (size_t)bins_cap * (3 * HL_TOP_BINS), st);
            }
            top_list.alloc(4ull * ((size_t)k_cap / 2 + 1) * HL_TOP_SMALL, st), top_small.alloc(sizeof(TopSmall) * ((size_t)k_cap / 2 + 1), st);
            if (getenv("HL_TOP_TRACE")) top_trace_buf.alloc(8000, st);
            HL_CUDA(cub::DeviceSelect::If(nullptr, a, ids, top_clusters.as<uint32_t>(), top_ctl.as<uint32_t>(), (int)(2 * n - 1), IsClusterRoot { BinaryTree(), C_top }, st));
            HL_CUDA(cub::DeviceSelect::If(nullptr, b, ids, top_free.as<uint32_t>(), top_ctl.as<uint32_t>() + 1, (int)(n - 1), IsUpperNode { BinaryTree(), C_top }, st));
            int per_sm = 0;
            HL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_top_build, 256, 0));
            // blocks: a thread per ~4 clusters of the expected cut, between one and four blocks per SM — every grid barrier costs
            // one atomic arrival per block on one address, and the level loop of the two-level re-split is nothing but barriers
            top_grid = ctx->sm_count * std::max(1, std::min(per_sm, 4));
            top_grid = std::max(1, std::min<int>(top_grid, std::max<int>(ctx->sm_count, (int)((k_cap + 1023u) / 1024u))));
            top_grid = std::max(1, std::min<int>(top_grid, (int)((k_cap + 255u) / 256u)));
            if (const char* e = getenv("HL_TOP_GRID")) top_grid = std::max(1, std::min(top_grid, atoi(e))); // (tuning runs)
        }
        if (two_level)
        {
            top_treelets.alloc(sizeof(TopTreelet) * ((size_t)k_cap + 1), st), top_tlist.alloc(4ull * ((size_t)k_cap + 1), st);
            // (per device, so not cached in a static: a process may hold contexts on several GPUs)
            HL_CUDA(cudaFuncSetAttribute(k_treelets, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TreeletSmem)));
        }
        top_tmp_bytes = std::max(a, b);
        top_tmp.alloc(std::max<size_t>(top_tmp_bytes, 16), st);
    }

    cudaEvent_t e0, e1, e2, e3;
    HL_CUDA(cudaEventCreate(&e0));
    HL_CUDA(cudaEventCreate(&e1));
    HL_CUDA(cudaEventCreate(&e2));
    HL_CUDA(cudaEventCreate(&e3));
    HL_CUDA(cudaEventRecord(e0, st));
    // scene box
    k_box_reduce<<<rb, 256, 0, st>>>(d_boxes, n, partial.as<Box>());
    k_box_reduce<<<1, 256, 0, st>>>(partial.as<Box>(), (uint32_t)rb, scene.as<Box>());
    ctx->launches += 2;
    // Morton keys + sort
    k_morton<<<grid_for(n, 256, cap), 256, 0, st>>>(d_boxes, scene.as<Box>(), n, keys_a.as<uint64_t>(), vals_a.as<uint32_t>());
    ctx->launches++;
    HL_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp_bytes, dk, dv, (int)n, 0, 63, st));
    ctx->launches += 10;
    const uint64_t* keys   = dk.Current();
    const uint32_t* sorted = dv.Current();
    // binary radix tree
    BinaryTree t;
    t.cost   = tree_cost.as<float>();
    t.dec    = (uint32_t*)(t.cost + 7ull * 2 * n);
    t.c_prim = leaf_bytes == sizeof(LeafTri) ? HL_SAH_C_PRIM_TRIANGLE : HL_SAH_C_PRIM_INSTANCE;
    t.n      = n;
    t.left   = tree_u32.as<uint32_t>();
    t.right  = t.left + ni;
    t.first  = t.right + ni;
    t.last   = t.first + ni;
    t.visits = t.last + ni;
    t.parent = t.visits + ni;
    t.box    = tree_box.as<Box>();
    HL_CUDA(cudaMemsetAsync(t.visits, 0, 4ull * ni, st));
    HL_CUDA(cudaMemsetAsync(t.parent, 0xFF, 4ull * 2 * n, st));
    if (n > 1)
    {
        k_radix_tree<<<grid_for(n - 1, 256, cap), 256, 0, st>>>(keys, t);
        ctx->launches++;
    }
    if (two_level)
    {
        k_fit_fine<<<grid_for(n, 256, cap), 256, 0, st>>>(t, d_boxes, sorted, C); // boxes + cost tables inside the fine clusters
        ctx->launches++;
    }
    else
    {
        k_fit<<<grid_for(n, 256, cap), 256, 0, st>>>(t, d_boxes, sorted, resplit ? C : 0u); // with a re-split: cluster subtrees only, k_top_refit fits the rest
        ctx->launches++;
    }
    if (resplit)
    {
        uint32_t* ctl = top_ctl.as<uint32_t>(); // [0] K, [1] upper nodes, [2] next free node, [3] treelets, [4..5] barrier, [6] small nodes, [7] treelet error, [8..] per-level counters; behind them: tlist entries used
        HL_CUDA(cudaMemsetAsync(ctl, 0, 4ull * top_ctl_words, st));
        TopBuild tb;
        tb.cluster_prims = C_top, tb.k_cap = k_cap, tb.bins_cap = bins_cap;
        tb.n_clusters = ctl, tb.cluster = top_clusters.as<uint32_t>(), tb.free_nodes = top_free.as<uint32_t>(), tb.cnode = top_cnode.as<uint32_t>(), tb.crec = top_crec.as<TopCluster>();
        for (int p = 0; p < 2; p++) tb.level[p] = top_level[p].as<TopNode>(), tb.bins[p] = top_bins[p].as<TopBin>();
        tb.free_next   = ctl + 2;
        tb.list = top_list.as<uint32_t>(), tb.small = top_small.as<TopSmall>(), tb.small_count = ctl + 6;
        tb.level_count = ctl + 8, tb.bins_used = tb.level_count + (HL_TOP_MAX_LEVELS + 1);
        if (two_level)
        {
            tb.treelet_prims = HL_TREELET_PRIMS, tb.treelet = top_treelets.as<TopTreelet>(), tb.treelet_count = ctl + 3, tb.tlist = top_tlist.as<uint32_t>();
            tb.tlist_used = tb.bins_used + HL_TOP_MAX_LEVELS; // (the last per-level slot: levels stop at HL_TOP_MAX_LEVELS - 1)
        }
        if (resplit_top)
        {
            thrust::counting_iterator<uint32_t> ids(0u);
            HL_CUDA(cub::DeviceSelect::If(top_tmp.p, top_tmp_bytes, ids, top_clusters.as<uint32_t>(), ctl, (int)(2 * n - 1), IsClusterRoot { t, C_top }, st));
            HL_CUDA(cub::DeviceSelect::If(top_tmp.p, top_tmp_bytes, ids, top_free.as<uint32_t>(), ctl + 1, (int)(n - 1), IsUpperNode { t, C_top }, st));
            if (two_level)
            {
                k_coarse_boxes<<<grid_for(k_cap, 256, cap), 256, 0, st>>>(t, top_clusters.as<uint32_t>(), ctl, C);
                ctx->launches++;
            }
            uint32_t*           bar    = ctl + 4;
            unsigned long long* trace  = top_trace_buf.p ? top_trace_buf.as<unsigned long long>() : nullptr;
            void*               args[] = { (void*)&t, (void*)&tb, (void*)&bar, (void*)&trace };
            HL_CUDA(cudaLaunchCooperativeKernel((const void*)k_top_build, dim3((unsigned)top_grid), dim3(256), args, 0, st));
            ctx->launches += 5;
            if (!two_level)
            {
                k_top_small<<<grid_for(k_cap / 2 + 1, 128, ctx->sm_count * 16), 128, 0, st>>>(t, tb);
                k_top_refit<<<grid_for(n, 256, cap), 256, 0, st>>>(t, tb);
                ctx->launches += 2;
            }
            if (trace)
            {
                std::vector<unsigned long long> h(1000);
                std::vector<uint32_t>           hc(top_ctl_words);
                HL_CUDA(cudaMemcpyAsync(h.data(), trace, 8000, cudaMemcpyDeviceToHost, st));
                HL_CUDA(cudaMemcpyAsync(hc.data(), ctl, 4 * top_ctl_words, cudaMemcpyDeviceToHost, st));
                HL_CUDA(cudaStreamSynchronize(st));
                fprintf(stderr, "[top re-split] n %u C %u K %u small %u treelets %u grid %d: phase us:", n, C_top, hc[0], hc[6], hc[3], top_grid);
                for (uint32_t k = 2; k < h[0] && k < 1000; k++) fprintf(stderr, " %.1f", (double)(h[k] - h[k - 1]) * 1e-3);
                fprintf(stderr, "\n  nodes per level:");
                for (uint32_t k = 0; k < HL_TOP_MAX_LEVELS && hc[8 + k]; k++) fprintf(stderr, " %u", hc[8 + k]);
                fprintf(stderr, "\n");
            }
        }
        else if (two_level)
        {
            // a tree of at most C_A primitives: ONE treelet whose only coarse subtree is the radix tree's root
            const TopTreelet whole = { 1u, 0xFFFFFFFFu, 0u, 0u, 0u, 0xFFFFFFFFu };
            const uint32_t   zero = 0u, one = 1u;
            HL_CUDA(cudaMemcpyAsync(top_treelets.p, &whole, sizeof(whole), cudaMemcpyHostToDevice, st));
            HL_CUDA(cudaMemcpyAsync(top_tlist.p, &zero, 4, cudaMemcpyHostToDevice, st));
            HL_CUDA(cudaMemcpyAsync(top_clusters.p, &zero, 4, cudaMemcpyHostToDevice, st));
            HL_CUDA(cudaMemcpyAsync(ctl + 3, &one, 4, cudaMemcpyHostToDevice, st));
        }
        if (two_level)
        {
            k_treelets<<<ctx->sm_count * 5, HL_TREELET_THREADS, sizeof(TreeletSmem), st>>>(t, tb, C, ctl + 7, ctl + (top_ctl_words - 1));
            ctx->launches++;
            if (resplit_top)
            {
                k_top_refit_treelets<<<grid_for(k_cap, 256, cap), 256, 0, st>>>(t, tb);
                ctx->launches++;
            }
        }
    }
    // collapse: one persistent launch over a device-side work queue, then the leaf records in a parallel pass
    HL_CUDA(cudaMemsetAsync(queue.p, 0xFF, sizeof(CollapseTask) * (size_t)capacity, st));
    uint32_t h_ctr[8] = { 1u, 0u, 1u, 0u, 1u, 0u, 0u, 0u }; // nodes, leaves, tail, head, outstanding, done
    HL_CUDA(cudaMemcpyAsync(ctr.p, h_ctr, 32, cudaMemcpyHostToDevice, st));
    CollapseTask root_task;
    root_task.wide = 0, root_task.bnode = 0;
    HL_CUDA(cudaMemcpyAsync(queue.p, &root_task, sizeof(root_task), cudaMemcpyHostToDevice, st));
    WideOut wo;
    wo.nodes = big_nodes.as<WideNode>(), wo.node_counter = ctr.as<uint32_t>(), wo.leaf_counter = ctr.as<uint32_t>() + 1;
    DeferredLeafWriter deferred;
    deferred.leaf_pos = leaf_pos.as<uint32_t>();
    k_collapse<<<ctx->sm_count * 10, 128, 0, st>>>(t, wo, queue.as<CollapseTask>(), capacity, ctr.as<uint32_t>(), deferred); // (48 registers: 10 blocks per SM are resident; the kernel waits on memory)
    k_write_leaves<<<grid_for(n, 256, cap), 256, 0, st>>>(make_writer(sorted, out.leaves.p), leaf_pos.as<uint32_t>(), n);
    ctx->launches += 2;
    HL_CUDA(cudaEventRecord(e1, st));
    HL_CUDA(cudaMemcpyAsync(h_ctr, ctr.p, 8, cudaMemcpyDeviceToHost, st));
    HL_CUDA(cudaMemcpyAsync(&out.root, t.box, sizeof(Box), cudaMemcpyDeviceToHost, st));
    uint32_t treelet_err = 0;
    if (two_level) HL_CUDA(cudaMemcpyAsync(&treelet_err, top_ctl.as<uint32_t>() + 7, 4, cudaMemcpyDeviceToHost, st));
    HL_CUDA(cudaMemcpyAsync(&out.sah_cost, t.cost, 4, cudaMemcpyDeviceToHost, st));
    HL_CUDA(cudaStreamSynchronize(st));
    if (treelet_err) throw CudaError(HL_ERR_STATE, treelet_err == 1u ? "BVH build: a treelet exceeds the shared-memory arrays" : "BVH build: treelet node ids and clusters do not add up");
    out.n_nodes = h_ctr[0], out.n_leaves = h_ctr[1], out.n_binary = 2 * n - 1;
    {
        const float a = box_half_area(out.root);
        out.sah_cost  = a > 0.0f ? out.sah_cost / a : 0.0f; // C(root, 1) / A_root: expected cost per ray hitting the root box
    }
    out.nodes.alloc(sizeof(WideNode) * (size_t)out.n_nodes);
    HL_CUDA(cudaEventRecord(e2, st));
    HL_CUDA(cudaMemcpyAsync(out.nodes.p, big_nodes.p, sizeof(WideNode) * (size_t)out.n_nodes, cudaMemcpyDeviceToDevice, st));
    HL_CUDA(cudaEventRecord(e3, st));
    HL_CUDA(cudaEventSynchronize(e3));
    float ms_a = 0.0f, ms_b = 0.0f;
    HL_CUDA(cudaEventElapsedTime(&ms_a, e0, e1));
    HL_CUDA(cudaEventElapsedTime(&ms_b, e2, e3));
    out.ms_build = ms_a + ms_b;
    cudaEventDestroy(e0), cudaEventDestroy(e1), cudaEventDestroy(e2), cudaEventDestroy(e3);
}

void build_mesh_bvh(hl_context_t* ctx, hl_mesh_t* mesh)
{
    t_scratch_pool           = ctx->scratch_pool;
    cudaStream_t          st = ctx->stream;
    const uint32_t        ng = (uint32_t)mesh->subs.size();
    std::vector<uint32_t> tri_start(ng + 1, 0);
    for (uint32_t g = 0; g < ng; g++) tri_start[g + 1] = tri_start[g] + mesh->subs[g].index_count / 3;
    const uint32_t n = tri_start[ng];
    mesh->tri_start.upload(tri_start.data(), 4ull * (ng + 1), st);
    mesh->submeshes.upload(mesh->subs.data(), sizeof(hl_submesh) * (size_t)ng, st);
    ScratchBuf boxes;
    boxes.alloc(sizeof(Box) * (size_t)std::max(n, 1u), st);
    cudaEvent_t e0, e1;
    HL_CUDA(cudaEventCreate(&e0));
    HL_CUDA(cudaEventCreate(&e1));
    HL_CUDA(cudaEventRecord(e0, st));
    if (n)
    {
        k_tri_boxes<<<grid_for(n, 256, ctx->sm_count * 8), 256, 0, st>>>(mesh->vertices.as<hl_vertex>(), mesh->indices.as<uint32_t>(), mesh->submeshes.as<hl_submesh>(), mesh->tri_start.as<uint32_t>(), ng, n, boxes.as<Box>());
        ctx->launches++;
    }
    HL_CUDA(cudaEventRecord(e1, st));
    const hl_vertex*  dv  = mesh->vertices.as<hl_vertex>();
    const uint32_t*   di  = mesh->indices.as<uint32_t>();
    const hl_submesh* dsm = mesh->submeshes.as<hl_submesh>();
    const uint32_t*   dts = mesh->tri_start.as<uint32_t>();
    bool              any_hit = false; // a geometry without VK_GEOMETRY_OPAQUE_BIT runs the any-hit stage: give its triangles alpha records
    for (const hl_submesh& sm : mesh->subs) any_hit = any_hit || !sm.opaque;
    if (any_hit && n) mesh->alpha.alloc(sizeof(AlphaTri) * (size_t)n);
    AlphaTri* da = any_hit && n ? mesh->alpha.as<AlphaTri>() : nullptr;
    build_wide_device(ctx, boxes.as<Box>(), n, sizeof(LeafTri), mesh->bvh, [=](const uint32_t* sorted, void* leaves) {
        TriLeafWriter w;
        w.vertices = dv, w.indices = di, w.submeshes = dsm, w.tri_start = dts, w.n_geom = ng, w.sorted_prim = sorted, w.tris = (LeafTri*)leaves, w.alpha = da;
        return w;
    });
    float ms0 = 0.0f;
    HL_CUDA(cudaEventSynchronize(e1));
    HL_CUDA(cudaEventElapsedTime(&ms0, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    mesh->stats.triangles       = n;
    mesh->stats.wide_nodes      = mesh->bvh.n_nodes;
    mesh->stats.binary_nodes    = mesh->bvh.n_binary;
    mesh->stats.ms_build        = mesh->bvh.ms_build + ms0;
    mesh->stats.sah_cost        = mesh->bvh.sah_cost;
    mesh->stats.bytes_nodes     = sizeof(WideNode) * (uint64_t)mesh->bvh.n_nodes;
    mesh->stats.bytes_triangles = sizeof(LeafTri) * (uint64_t)mesh->bvh.n_leaves;
}

// Instance-tree refit in ONE block (at most 1024 instances -> at most ~1024 wide nodes): sweeps "box of every node from its
// children's boxes of the previous sweep" until nothing changes — the number of sweeps is the depth of the tree — then
// requantises every node.  node_box = scratch, one Box per wide node.
__global__ void __launch_bounds__(256) k_tlas_refit(WideNode* nodes, uint32_t n_nodes, const uint32_t* inst_leaf, const Box* inst_boxes, Box* node_box)
{
    __shared__ int changed;
    for (uint32_t i = threadIdx.x; i < n_nodes; i += blockDim.x) node_box[i] = box_empty();
    __syncthreads();
    for (int sweep = 0; sweep < 64; sweep++)
    {
        if (threadIdx.x == 0) changed = 0;
        __syncthreads();
        Box      mine[4]; // up to 4 nodes per thread (1024 / 256)
        uint32_t k = 0;
        for (uint32_t i = threadIdx.x; i < n_nodes && k < 4u; i += blockDim.x, k++) mine[k] = refit_node_box(nodes[i], inst_leaf, inst_boxes, node_box);
        __syncthreads();
        k = 0;
        for (uint32_t i = threadIdx.x; i < n_nodes && k < 4u; i += blockDim.x, k++)
        {
            bool diff = false;
            for (int a = 0; a < 3; a++) diff = diff || mine[k].lo[a] != node_box[i].lo[a] || mine[k].hi[a] != node_box[i].hi[a];
            if (diff) node_box[i] = mine[k], changed = 1;
        }
        __syncthreads();
        if (!changed) break;
        __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < n_nodes; i += blockDim.x)
    {
        WideNode w = nodes[i];
        refit_requantize(w, node_box[i], inst_leaf, inst_boxes, node_box);
        nodes[i] = w;
    }
}

void refit_tlas(hl_context_t* ctx, const std::vector<Box>& instance_boxes)
{
    cudaStream_t st = ctx->stream;
    t_scratch_pool  = ctx->scratch_pool;
    if (ctx->tlas.n_nodes == 0 || ctx->tlas.n_nodes > 4u * 256u) throw CudaError(HL_ERR_STATE, "refit_tlas: no instance tree (or one larger than the refit kernel handles)");
    ScratchBuf boxes, node_box;
    boxes.alloc(sizeof(Box) * instance_boxes.size(), st);
    node_box.alloc(sizeof(Box) * (size_t)ctx->tlas.n_nodes, st);
    HL_CUDA(cudaMemcpyAsync(boxes.p, instance_boxes.data(), sizeof(Box) * instance_boxes.size(), cudaMemcpyHostToDevice, st));
    cudaEvent_t e0, e1;
    HL_CUDA(cudaEventCreate(&e0));
    HL_CUDA(cudaEventCreate(&e1));
    HL_CUDA(cudaEventRecord(e0, st));
    k_tlas_refit<<<1, 256, 0, st>>>(ctx->tlas.nodes.as<WideNode>(), ctx->tlas.n_nodes, ctx->tlas.leaves.as<uint32_t>(), boxes.as<Box>(), node_box.as<Box>());
    ctx->launches++;
    HL_CUDA(cudaEventRecord(e1, st));
    HL_CUDA(cudaMemcpyAsync(&ctx->tlas.root, node_box.p, sizeof(Box), cudaMemcpyDeviceToHost, st));
    HL_CUDA(cudaStreamSynchronize(st)); // (the host vector must outlive the copy)
    HL_CUDA(cudaEventElapsedTime(&ctx->tlas.ms_build, e0, e1));
    cudaEventDestroy(e0), cudaEventDestroy(e1);
}

void build_tlas(hl_context_t* ctx, const std::vector<Box>& instance_boxes)
{
    t_scratch_pool = ctx->scratch_pool;
    DevBuf boxes;
    boxes.upload(instance_boxes.data(), sizeof(Box) * instance_boxes.size(), ctx->stream);
    build_wide_device(ctx, boxes.as<Box>(), (uint32_t)instance_boxes.size(), sizeof(uint32_t), ctx->tlas, [](const uint32_t* sorted, void* leaves) {
        InstLeafWriter w;
        w.sorted_prim = sorted, w.leaf = (uint32_t*)leaves;
        return w;
    });
}
} // namespace hl
