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
//   fit            24 read (sorted box) + 2 x 24 written + 48 read + 2 x 28 cost-table rows written, 56 read   232
//   top re-split   clusters of <= C primitives (K ~ N / C..N): 2 selects over the node ids (8), then per level
//                  K x (24 box + 4 node id) read twice + 21 + 6 atomics per cluster still above a small node;
//                  ~log2(K / 16) + imbalance levels; refit of the K - 1 re-linked nodes (as `fit`)       ~150 (C = 2)
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
    uint32_t      split = 0xFFFFFFFFu, n_left = 0u;
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
            if (cost < best) best = cost, split = (uint32_t)ax | (lane << 2), n_left = pc;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1)
    {
        const float    ob = __shfl_xor_sync(FULL, best, d);
        const uint32_t os = __shfl_xor_sync(FULL, split, d), on = __shfl_xor_sync(FULL, n_left, d);
        if (ob < best || (ob == best && os < split)) best = ob, split = os, n_left = on;
    }
    if (lane == 0u)
    {
        if (best < hl_inf())
            N.split = split, N.n_left = n_left;
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
    if (tid == 0) top_begin(tb, K);
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
        for (uint32_t j = warp; j < nodes; j += nwarps) top_choose_node_warp(tb, level, j, lane);
        grid_barrier(bar);
        top_trace(trace, ntrace);
        for (uint32_t j = tid; j < nodes; j += nthr) top_commit_node(t, tb, level, j);
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
    for (;;)
    {
        if (!have) idx = hl_alloc_var(ctl + 3, 1u), have = true; // one atomic per group of lanes that need a new queue index
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
            const uint32_t spawned = collapse_one(t, task, out, queue, ctl + 2, writer);
            have                   = false;
            // outstanding += spawned - 1 (summed over the lanes that finished a task together: one atomic); the update that brings
            // it to zero ends the launch.  Every finished task was published before, so zero is only reached at the very end.
            {
                namespace cg = cooperative_groups;
                const cg::coalesced_group g     = cg::coalesced_threads();
                const uint32_t            delta = cg::reduce(g, spawned - 1u, cg::plus<uint32_t>());
                if (g.thread_rank() == 0 && atomicAdd(ctl + 4, delta) + delta == 0u)
                {
                    __threadfence();
                    *vdone = 1u;
                }
            }
        }
        else if (*vdone)
            break;
        else
            __nanosleep(64); // waiting lanes yield the issue slots to the lanes of the warp that hold a task
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
    while (C && tri_tree && n / C > (4u << 20)) C *= 2;
    const bool     resplit  = C != 0 && n >= 2 && n > C;
    const uint32_t k_cap    = resplit ? (uint32_t)std::min<uint64_t>(n, 4ull * n / C + 1024) : 0u;
    const uint32_t bins_cap = k_cap / (HL_TOP_SMALL + 1u) + 2u;
    const size_t   top_ctl_words = 8 + 2 * (HL_TOP_MAX_LEVELS + 1);
    ScratchBuf     top_crec;
    ScratchBuf     top_clusters, top_free, top_cnode, top_level[2], top_bins[2], top_list, top_small, top_ctl, top_tmp, top_trace_buf;
    size_t         top_tmp_bytes = 0;
    int            top_grid      = 0;
    if (resplit)
    {
        top_clusters.alloc(4ull * n, st), top_free.alloc(4ull * n, st), top_cnode.alloc(4ull * k_cap, st), top_crec.alloc(sizeof(TopCluster) * (size_t)k_cap, st);
        for (int p = 0; p < 2; p++)
        {
            top_level[p].alloc(sizeof(TopNode) * ((size_t)k_cap + 2), st);
            top_bins[p].alloc(sizeof(TopBin) * (size_t)bins_cap * (3 * HL_TOP_BINS), st);
        }
        top_list.alloc(4ull * ((size_t)k_cap / 2 + 1) * HL_TOP_SMALL, st), top_small.alloc(sizeof(TopSmall) * ((size_t)k_cap / 2 + 1), st);
        top_ctl.alloc(4ull * top_ctl_words, st);
        if (getenv("HL_TOP_TRACE")) top_trace_buf.alloc(8000, st);
        size_t a = 0, b = 0;
        thrust::counting_iterator<uint32_t> ids(0u);
        HL_CUDA(cub::DeviceSelect::If(nullptr, a, ids, top_clusters.as<uint32_t>(), top_ctl.as<uint32_t>(), (int)(2 * n - 1), IsClusterRoot { BinaryTree(), C }, st));
        HL_CUDA(cub::DeviceSelect::If(nullptr, b, ids, top_free.as<uint32_t>(), top_ctl.as<uint32_t>() + 1, (int)(n - 1), IsUpperNode { BinaryTree(), C }, st));
        top_tmp_bytes = std::max(a, b);
        top_tmp.alloc(top_tmp_bytes, st);
        int per_sm = 0;
        HL_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_top_build, 256, 0));
        top_grid = ctx->sm_count * std::max(1, std::min(per_sm, 4));
        top_grid = std::max(1, std::min<int>(top_grid, (int)((k_cap + 255u) / 256u)));
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
    k_fit<<<grid_for(n, 256, cap), 256, 0, st>>>(t, d_boxes, sorted, resplit ? C : 0u); // with a re-split: cluster subtrees only, k_top_refit fits the rest
    ctx->launches++;
    if (resplit)
    {
        uint32_t* ctl = top_ctl.as<uint32_t>(); // [0] K, [1] upper nodes, [2] next free node, [4..5] barrier, [6] small nodes, [8..] per-level counters
        HL_CUDA(cudaMemsetAsync(ctl, 0, 4ull * top_ctl_words, st));
        thrust::counting_iterator<uint32_t> ids(0u);
        HL_CUDA(cub::DeviceSelect::If(top_tmp.p, top_tmp_bytes, ids, top_clusters.as<uint32_t>(), ctl, (int)(2 * n - 1), IsClusterRoot { t, C }, st));
        HL_CUDA(cub::DeviceSelect::If(top_tmp.p, top_tmp_bytes, ids, top_free.as<uint32_t>(), ctl + 1, (int)(n - 1), IsUpperNode { t, C }, st));
        TopBuild tb;
        tb.cluster_prims = C, tb.k_cap = k_cap, tb.bins_cap = bins_cap;
        tb.n_clusters = ctl, tb.cluster = top_clusters.as<uint32_t>(), tb.free_nodes = top_free.as<uint32_t>(), tb.cnode = top_cnode.as<uint32_t>(), tb.crec = top_crec.as<TopCluster>();
        for (int p = 0; p < 2; p++) tb.level[p] = top_level[p].as<TopNode>(), tb.bins[p] = top_bins[p].as<TopBin>();
        tb.free_next   = ctl + 2;
        tb.list = top_list.as<uint32_t>(), tb.small = top_small.as<TopSmall>(), tb.small_count = ctl + 6;
        tb.level_count = ctl + 8, tb.bins_used = tb.level_count + (HL_TOP_MAX_LEVELS + 1);
        uint32_t*           bar    = ctl + 4;
        unsigned long long* trace  = top_trace_buf.p ? top_trace_buf.as<unsigned long long>() : nullptr;
        void*               args[] = { (void*)&t, (void*)&tb, (void*)&bar, (void*)&trace };
        HL_CUDA(cudaLaunchCooperativeKernel((const void*)k_top_build, dim3((unsigned)top_grid), dim3(256), args, 0, st));
        k_top_small<<<grid_for(k_cap / 2 + 1, 128, ctx->sm_count * 16), 128, 0, st>>>(t, tb);
        k_top_refit<<<grid_for(n, 256, cap), 256, 0, st>>>(t, tb);
        ctx->launches += 9;
        if (trace)
        {
            std::vector<unsigned long long> h(1000);
            std::vector<uint32_t>           hc(top_ctl_words);
            HL_CUDA(cudaMemcpyAsync(h.data(), trace, 8000, cudaMemcpyDeviceToHost, st));
            HL_CUDA(cudaMemcpyAsync(hc.data(), ctl, 4 * top_ctl_words, cudaMemcpyDeviceToHost, st));
            HL_CUDA(cudaStreamSynchronize(st));
            fprintf(stderr, "[top re-split] n %u C %u K %u small %u grid %d: phase us:", n, C, hc[0], hc[6], top_grid);
            for (uint32_t k = 2; k < h[0] && k < 1000; k++) fprintf(stderr, " %.1f", (double)(h[k] - h[k - 1]) * 1e-3);
            fprintf(stderr, "\n  nodes per level:");
            for (uint32_t k = 0; k < HL_TOP_MAX_LEVELS && hc[8 + k]; k++) fprintf(stderr, " %u", hc[8 + k]);
            fprintf(stderr, "\n");
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
    k_collapse<<<cap, 128, 0, st>>>(t, wo, queue.as<CollapseTask>(), capacity, ctr.as<uint32_t>(), deferred);
    k_write_leaves<<<grid_for(n, 256, cap), 256, 0, st>>>(make_writer(sorted, out.leaves.p), leaf_pos.as<uint32_t>(), n);
    ctx->launches += 2;
    HL_CUDA(cudaEventRecord(e1, st));
    HL_CUDA(cudaMemcpyAsync(h_ctr, ctr.p, 8, cudaMemcpyDeviceToHost, st));
    HL_CUDA(cudaMemcpyAsync(&out.root, t.box, sizeof(Box), cudaMemcpyDeviceToHost, st));
    HL_CUDA(cudaMemcpyAsync(&out.sah_cost, t.cost, 4, cudaMemcpyDeviceToHost, st));
    HL_CUDA(cudaStreamSynchronize(st));
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
