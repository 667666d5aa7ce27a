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
//   collapse       ~48 read (boxes) + 8 decisions + 0.12 x 80 node + 48 leaf + 48 src + 2 x 8 queue   ~180
//   total                                                              ~ 780 B / triangle
#include "hl_internal.h"
#include <cub/cub.cuh>

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
__global__ void k_fit(BinaryTree t, const Box* prim_boxes, const uint32_t* sorted)
{
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < t.n; j += gridDim.x * blockDim.x)
    {
        t.box[(t.n - 1) + j] = prim_boxes[sorted[j]];
        fit_from_leaf(t, j, DeviceFence());
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
    for (;;)
    {
        if (!have) idx = atomicAdd(ctl + 3, 1u), have = true;
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
            // outstanding += spawned - 1; the thread that brings it to zero ends the launch
            if (atomicAdd(ctl + 4, spawned - 1u) + spawned - 1u == 0u)
            {
                __threadfence();
                *vdone = 1u;
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
    k_fit<<<grid_for(n, 256, cap), 256, 0, st>>>(t, d_boxes, sorted);
    ctx->launches++;
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
    build_wide_device(ctx, boxes.as<Box>(), n, sizeof(LeafTri), mesh->bvh, [=](const uint32_t* sorted, void* leaves) {
        TriLeafWriter w;
        w.vertices = dv, w.indices = di, w.submeshes = dsm, w.tri_start = dts, w.n_geom = ng, w.sorted_prim = sorted, w.tris = (LeafTri*)leaves;
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

void build_tlas(hl_context_t* ctx, const std::vector<Box>& instance_boxes)
{
    DevBuf boxes;
    boxes.upload(instance_boxes.data(), sizeof(Box) * instance_boxes.size(), ctx->stream);
    build_wide_device(ctx, boxes.as<Box>(), (uint32_t)instance_boxes.size(), sizeof(uint32_t), ctx->tlas, [](const uint32_t* sorted, void* leaves) {
        InstLeafWriter w;
        w.sorted_prim = sorted, w.leaf = (uint32_t*)leaves;
        return w;
    });
}
} // namespace hl
