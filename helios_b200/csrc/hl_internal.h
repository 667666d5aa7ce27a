// hl_internal.h — host-side internals of libhelios_b200.so (context, device buffers, launch prototypes).
#pragma once
#include "hl_build.h"
#include "hl_scene.h"
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>
#include <vector>

namespace hl
{
struct CudaError : std::runtime_error
{
    int status;
    CudaError(int st, const std::string& m) : std::runtime_error(m), status(st) {}
};
#define HL_CUDA(call)                                                                                                              \
    do                                                                                                                             \
    {                                                                                                                              \
        cudaError_t e_ = (call);                                                                                                   \
        if (e_ != cudaSuccess)                                                                                                     \
            throw hl::CudaError(e_ == cudaErrorMemoryAllocation ? HL_ERR_OUT_OF_MEMORY : HL_ERR_CUDA,                              \
                                std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

struct DevBuf
{
    void*  p     = nullptr;
    size_t bytes = 0;
    DevBuf() {}
    DevBuf(const DevBuf&)            = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr, bytes = 0;
    }
    void alloc(size_t n)
    {
        if (n <= bytes && p) return;
        release();
        if (n == 0) n = 16;
        HL_CUDA(cudaMalloc(&p, n));
        bytes = n;
    }
    void upload(const void* src, size_t n, cudaStream_t s)
    {
        alloc(n);
        if (n) HL_CUDA(cudaMemcpyAsync(p, src, n, cudaMemcpyHostToDevice, s));
    }
    template <class T>
    T* as() const { return (T*)p; }
};

// build scratch: stream-ordered allocation from the context's PRIVATE memory pool (hl_context_create; the builder entry
// points select it for the calling thread), so a second build reuses the first build's memory and no other
// cudaMallocAsync user of the process is affected
inline thread_local cudaMemPool_t t_scratch_pool = nullptr;
struct ScratchBuf
{
    void*        p = nullptr;
    cudaStream_t s = nullptr;
    ScratchBuf() {}
    ScratchBuf(const ScratchBuf&)            = delete;
    ScratchBuf& operator=(const ScratchBuf&) = delete;
    ~ScratchBuf()
    {
        if (p) cudaFreeAsync(p, s);
    }
    void alloc(size_t n, cudaStream_t stream)
    {
        if (p) cudaFreeAsync(p, s), p = nullptr;
        s = stream;
        if (t_scratch_pool)
            HL_CUDA(cudaMallocFromPoolAsync(&p, n ? n : 16, t_scratch_pool, stream));
        else
            HL_CUDA(cudaMallocAsync(&p, n ? n : 16, stream));
    }
    template <class T>
    T* as() const { return (T*)p; }
};

struct WideBVHDev
{
    DevBuf   nodes, leaves;
    uint32_t n_nodes = 0, n_leaves = 0, n_binary = 0;
    Box      root;
    float    ms_build = 0.0f;
    float    sah_cost = 0.0f; // C(root, 1) / A(root) of the collapse DP (hl_build.h)
};

#define HL_MAX_BOUNCES 64

// device counter block layout (uint32 indices), zeroed at the start of every frame; uint64 totals[2] follow
#define CTR_EXT_COUNT 0                    /* [HL_MAX_BOUNCES + 1] extension-queue sizes per bounce */
#define CTR_SH_COUNT (CTR_EXT_COUNT + 65)  /* [HL_MAX_BOUNCES] shadow-queue sizes per bounce */
#define CTR_EXT_FETCH (CTR_SH_COUNT + 64)  /* [HL_MAX_BOUNCES] work cursors of the persistent extend launches */
#define CTR_SH_FETCH (CTR_EXT_FETCH + 64)  /* [HL_MAX_BOUNCES] work cursors of the persistent connect launches */
#define CTR_TAIL_FETCH (CTR_SH_FETCH + 64)
#define CTR_TAIL_EXT (CTR_TAIL_FETCH + 1)
#define CTR_TAIL_SH (CTR_TAIL_FETCH + 2)
#define CTR_TAIL_DONE (CTR_TAIL_FETCH + 3)
#define CTR_U32_TOTAL (CTR_TAIL_FETCH + 4)
#define CTR_TOTALS_OFFSET ((CTR_U32_TOTAL * 4 + 7) / 8 * 8)
#define CTR_BYTES (CTR_TOTALS_OFFSET + 16)

} // namespace hl

struct hl_mesh_t
{
    hl::DevBuf              vertices, indices, submeshes, tri_start;
    hl::DevBuf              alpha; // AlphaTri per leaf (only when a submesh is not opaque)
    std::vector<hl_submesh> subs;
    uint32_t                n_vertices = 0, n_indices = 0;
    hl::WideBVHDev          bvh;
    hl_build_stats          stats {};
};

// Wavefront state of ONE frame in flight.  The context keeps two: frame f runs on slot f & 1 with its own CUDA
// stream, so the latency-bound late bounces of frame f overlap the throughput-bound first bounces of frame f + 1
// (the frames share nothing but the read-only scene; their resolve passes are chained by events).
// frames in flight: consecutive frames rotate through ctx->n_slots wavefront state slots / streams (hl_wavefront.cu,
// HL_OPT_FRAMES_IN_FLIGHT); slots beyond n_slots stay unallocated
#define HL_MAX_WAVE_SLOTS 8
#ifndef HL_DEFAULT_WAVE_SLOTS
#define HL_DEFAULT_WAVE_SLOTS 4
#endif
struct hl_wave_slot
{
    hl::DevBuf state_a, state_b;   // per path: (T.xyz, rng.x), (L.xyz, rng.y)
    hl::DevBuf ext_o[2], ext_d[2]; // extension queue: (o.xyz, path), (d.xyz, -)
    hl::DevBuf hit_a, hit_b;       // (t,u,v,prim), (instance, geometry)
    hl::DevBuf sh_o, sh_d, sh_c;   // shadow queue: (o.xyz, path), (d.xyz, tmax), (contribution.xyz, -)
    hl::DevBuf rgba8;              // tone-map target of the fused resolve pass (per slot: an asynchronous read-back of
                                   // frame f must not race with frame f + 1's resolve)
    hl::DevBuf counters;           // uint32: ext_count[65], sh_count[64], fetch_ext[64], fetch_sh[64], tail; then uint64 totals[2]
    cudaStream_t stream   = nullptr;
    cudaEvent_t  resolved = nullptr; // recorded after the slot's last resolve pass
    bool         pending  = false;   // frames were issued on `stream` since the last join with the main stream
    // asynchronous read-back (hl_render_frame_readback): the device->host copy of `rgba8` runs on its own stream, so the
    // slot's next frame can start tracing behind it; only that frame's resolve pass (which overwrites rgba8) waits for it
    cudaStream_t copy_stream  = nullptr;
    cudaEvent_t  image_ready  = nullptr, copy_done = nullptr;
    bool         copy_pending = false;
    uint8_t*     copy_host    = nullptr; // destination of that copy: a later read-back into the same memory is ordered behind it
    // the bounce loop of a frame (tail / extend / shade / connect per bounce: ~30 launches with arguments that only change
    // with the scene tables or the integrator settings) as an instantiated CUDA graph; rebuilt when `graph_key` changes
    cudaGraphExec_t graph_exec     = nullptr;
    uint32_t        graph_launches = 0; // kernel nodes in it
    struct GraphKey
    {
        hl::SceneView view;
        uint32_t      num_lights, max_ray_bounces, bounces, tail_start, tail_threshold, W, H;
        float         shadow_ray_bias;
    } graph_key;
};

struct hl_context_t
{
    int          device = 0;
    cudaStream_t stream = nullptr;
    uint32_t     W = 0, H = 0;
    int          sm_count = 148;
    cudaMemPool_t scratch_pool = nullptr; // private pool of the builder's scratch (ScratchBuf)
    std::string  err;
    uint64_t     launches = 0;
    bool         profiling = false;
    int          accum_mode = HL_ACCUM_RUNNING_MEAN;
    uint32_t     tail_start = 4, tail_threshold = 98304; // see k_tail (hl_wavefront.cu); tuned with 4 frames in flight (tools/gpu/slot_sweep.sh)
    uint32_t     sah_cluster = HL_DEFAULT_SAH_CLUSTER;   // binned-SAH re-split of the BVHs' upper levels (hl_build.h); 0 = off

    // resources
    std::vector<hl_mesh_t*>  meshes;
    std::vector<hl::DevBuf*> textures;
    std::vector<hl::TexView> tex_views;
    hl::DevBuf               env_faces;  // 6 * size^2 texels as uploaded / baked (hl_envmap_read returns these)
    hl::DevBuf               env_padded; // the same with the one-texel seamless border: what the kernels sample
    uint32_t                 env_size = 0;
    // scene tables
    hl::DevBuf      materials, instances, inst_inv, submesh_info, submesh_offset, lights, mesh_views, tex_views_dev, lut8, inst_alpha, geom_alpha, inst_identity;
    hl::WideBVHDev  tlas;
    std::vector<hl_instance> h_instances; // the table as uploaded (mesh_index in context order): hl_scene_update_instances edits transforms in place
    std::vector<hl_mesh_t*>  h_inst_mesh; // mesh of every instance
    hl::SceneView   view {};
    bool            scene_ready = false;
    // film + wavefront state
    hl::DevBuf   accum, rgba8;          // rgba8: target of the stand-alone tone-map pass (hl_tonemap)
    void*        rgba8_cur = nullptr;   // the RGBA8 image written last (ctx->rgba8 or a slot's)
    hl_wave_slot slot[HL_MAX_WAVE_SLOTS];
    int          n_slots     = HL_DEFAULT_WAVE_SLOTS;
    uint64_t     frame_seq   = 0;       // frames issued; frame f uses slot[f % n_slots]
    int          pipeline    = 1;       // 0: every frame on the main stream (also forced while profiling)
    int          use_graphs  = 1;       // pipelined frames replay their bounce loop from a CUDA graph (HL_OPT_CUDA_GRAPH)
    cudaEvent_t  main_ev     = nullptr; // orders work enqueued on the main stream before the next frame
    size_t       queue_capacity = 0;
    // profiling
    cudaEvent_t ev[2 + 5 * HL_MAX_BOUNCES + 4] {};
    hl_bounce_profile bounce_prof[HL_MAX_BOUNCES] {};
    uint32_t    bounce_prof_n = 0;
    uint64_t    prof_tail_ext = 0, prof_tail_sh = 0;
    bool        ev_ready = false;
    hl_counters last {};
    cudaEvent_t user_ev[8] {};
    bool        user_ev_ready = false;
    uint64_t    frames = 0;
    // multi-GPU (hl_comm.cu): NCCL communicator (opaque) and / or the peer-memory group this context belongs to
    void*       comm = nullptr;
    int         comm_nranks = 1, comm_rank = 0;
    bool        comm_p2p = false;                               // bound by hl_comm_init_all with peer access between all members
    cudaEvent_t comm_ready_ev = nullptr, comm_done_ev = nullptr; // "my image is complete" / "my slice is written"
    uint64_t    sum_samples = 0;                                // HL_ACCUM_SUM: full-frame launches since the last clear
};

namespace hl
{
// hl_builder.cu
void build_mesh_bvh(hl_context_t* ctx, hl_mesh_t* mesh);
void build_tlas(hl_context_t* ctx, const std::vector<Box>& instance_boxes);
void refit_tlas(hl_context_t* ctx, const std::vector<Box>& instance_boxes); // same topology, new boxes (hl_scene_update_instances)
// hl_wavefront.cu
void wavefront_alloc(hl_context_t* ctx);
void wavefront_release(hl_context_t* ctx);
void env_pad(hl_context_t* ctx); // env_faces -> env_padded + ctx->view.env
void wavefront_set_slots(hl_context_t* ctx, int n_slots); // HL_OPT_FRAMES_IN_FLIGHT
void wavefront_join(hl_context_t* ctx); // main stream waits for every frame in flight
struct ResolveOptions
{
    bool     tone_map = false; // also produce the RGBA8 image (fused into the resolve pass for full-frame launches)
    float    exposure = 1.0f;
    int      op       = 0;
    uint8_t* host     = nullptr; // asynchronous read-back of that image on the frame's stream
};
void wavefront_render_frame(hl_context_t* ctx, const hl_push_constants& pc, uint32_t lw, uint32_t lh, const ResolveOptions& opt = ResolveOptions());
void wavefront_primary_hits(hl_context_t* ctx, const hl_push_constants& pc);
void wavefront_output_buffer(hl_context_t* ctx, const hl_push_constants& pc, int which, float4* d_out); // debug output buffers
void wavefront_debug_rays(hl_context_t* ctx, const hl_push_constants& pc, uint32_t n, float4* d_verts, uint32_t capacity, uint32_t* d_count); // ray debug view
uint32_t wavefront_path_pixel(hl_context_t* ctx, uint32_t i); // path index -> row-major pixel index of a full-frame launch
void wavefront_trace_rays(hl_context_t* ctx, const float* d_rays, uint32_t n, uint32_t flags, void* d_hits);
uint64_t trav_overflow_count(hl_context_t* ctx, bool reset);
void film_clear(hl_context_t* ctx);
void film_tonemap(hl_context_t* ctx, float exposure, int op, float scale);
void sky_bake(hl_context_t* ctx, const float* coeffs40, const float* sun3, uint32_t size);
// hl_comm.cu
void comm_release(hl_context_t* ctx);
} // namespace hl
