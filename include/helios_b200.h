/*
 * helios_b200.h — C ABI of the B200-native path-tracing hot path.
 *
 * This is the drop-in boundary for the path-trace pass of diharaw/helios (reference
 * paths below are relative to the reference checkout).  Everything the reference hands
 * to the Vulkan driver on that path (vertex/index buffers, the Material / Instance /
 * Light storage buffers, the per-instance submesh table, textures, the environment cube
 * map and the 192-byte push-constant block) crosses this boundary as plain pointers and
 * sizes; acceleration-structure build, traversal, shading, accumulation and tone
 * mapping happen behind it in hand-written sm_100a CUDA kernels.
 *
 * Conventions
 *   - every entry point is extern "C", returns an hl_status (0 = HL_OK); no exception
 *     crosses the boundary; hl_last_error() gives the message of the last failure;
 *   - host input arrays are copied during the call (the caller keeps ownership);
 *     output pointers are caller-allocated HOST memory unless the name says _device;
 *   - one host thread per context; calls on one context are not re-entrant;
 *   - there is NO CPU fallback: without a CUDA device hl_context_create fails.
 *
 * The POD structs are the reference's shader ABI, byte for byte:
 *   hl_vertex          = Vertex          include/resource/mesh.h:10-17, common.glsl:45-52
 *   hl_material        = MaterialData    src/engine/resource/scene.cpp:25-32, common.glsl:80-87
 *   hl_light           = LightData       src/engine/resource/scene.cpp:36-42, common.glsl:89-95
 *   hl_instance        = InstanceData    src/engine/resource/scene.cpp:46-52, common.glsl:104-109
 *   hl_push_constants  = PushConstants   src/engine/gfx/path_integrator.cpp:11-28, path_trace_rgen.glsl:93-110
 */
#ifndef HELIOS_B200_H
#define HELIOS_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define HL_API __attribute__((visibility("default")))
#else
#define HL_API
#endif

/* ------------------------------------------------------------------ status codes */
typedef int hl_status;
#define HL_OK 0
#define HL_ERR_INVALID_ARGUMENT 1
#define HL_ERR_NO_DEVICE 2
#define HL_ERR_CUDA 3
#define HL_ERR_OUT_OF_MEMORY 4
#define HL_ERR_LIMIT 5      /* a reference limit was exceeded (include/resource/scene.h:14-17) */
#define HL_ERR_STATE 6      /* call made in the wrong state (e.g. render before hl_scene_commit) */

/* reference limits, include/resource/scene.h:14-17 */
#define HL_MAX_SCENE_MESH_INSTANCE_COUNT 1024
#define HL_MAX_SCENE_LIGHT_COUNT 100000
#define HL_MAX_SCENE_MATERIAL_COUNT 4096
#define HL_MAX_SCENE_MATERIAL_TEXTURE_COUNT (HL_MAX_SCENE_MATERIAL_COUNT * 4)

/* light types, common.glsl:11-15 */
#define HL_LIGHT_DIRECTIONAL 0
#define HL_LIGHT_SPOT 1
#define HL_LIGHT_POINT 2
#define HL_LIGHT_ENVIRONMENT_MAP 3
#define HL_LIGHT_AREA 4

/* tone map operators, tone_map.frag:3-4 */
#define HL_TONE_MAP_ACES 0
#define HL_TONE_MAP_REINHARD 1

/* texel formats (level 0 is the only level the path samples: textureLod(..., 0.0)) */
#define HL_TEX_RGBA8_UNORM 0
#define HL_TEX_RGBA8_SRGB 1
#define HL_TEX_RGBA8_SNORM 2 /* the reference's 8-bit non-sRGB quirk, core/resource_manager.cpp:32-36 */
#define HL_TEX_RGBA32F 3

/* accumulation modes */
#define HL_ACCUM_RUNNING_MEAN 0 /* path_trace_rgen.glsl:219-247, the reference's progressive blend */
#define HL_ACCUM_SUM 1          /* per-GPU partial sums for samples-per-pixel sharding */

/* ------------------------------------------------------------------ shader-ABI PODs */
typedef struct hl_vertex
{
    float position[4]; /* w = submesh index (core/resource_manager.cpp:467-473) */
    float tex_coord[4];
    float normal[4];
    float tangent[4];
    float bitangent[4];
} hl_vertex; /* 80 B */

typedef struct hl_material
{
    int32_t texture_indices0[4]; /* x albedo, y normal, z roughness, w metallic; -1 = none */
    int32_t texture_indices1[4]; /* x emissive, z roughness channel, w metallic channel */
    float   albedo[4];
    float   emissive[4];
    float   roughness_metallic[4];
} hl_material; /* 80 B */

typedef struct hl_light
{
    float light_data0[4];
    float light_data1[4];
    float light_data2[4];
    float light_data3[4];
} hl_light; /* 64 B */

typedef struct hl_instance
{
    float    model_matrix[16];  /* column-major (glm::mat4) */
    float    normal_matrix[16]; /* column-major; global transform without scale */
    uint32_t mesh_index;
    float    padding[3];
} hl_instance; /* 144 B */

typedef struct hl_push_constants
{
    float    view_proj_inverse[16]; /* column-major */
    float    camera_pos[4];
    float    up_direction[4];
    float    right_direction[4];
    float    focal_plane[4];
    int32_t  ray_debug_pixel_coord[4];
    uint32_t launch_id_size[4]; /* tile x, tile y, W, H */
    float    accumulation;
    uint32_t num_lights;
    uint32_t num_frames;
    uint32_t debug_vis;
    uint32_t max_ray_bounces;
    float    shadow_ray_bias;
    float    focal_length;
    float    aperture_radius;
} hl_push_constants; /* 192 B */

/* One BLAS geometry per submesh: include/resource/mesh.h:19-29 + src/engine/resource/mesh.cpp:66-103.
 * opaque = (material type == MATERIAL_OPAQUE && !alpha_tested)  -> VK_GEOMETRY_OPAQUE_BIT_KHR  */
typedef struct hl_submesh
{
    uint32_t base_index;  /* first index of the submesh in the mesh's index buffer */
    uint32_t index_count; /* 3 * triangle count */
    uint32_t vertex_count;
    uint32_t opaque;
} hl_submesh;

typedef struct hl_counters
{
    uint64_t extension_rays; /* rays traced by the extend stage (primary + indirect) */
    uint64_t shadow_rays;    /* rays traced by the connect stage */
    uint64_t frames;         /* hl_render_frame calls since the last reset */
    float    ms_generate;    /* device time of the last frame, per stage (CUDA events)  */
    float    ms_extend;
    float    ms_shade;
    float    ms_connect;
    float    ms_resolve;
    float    ms_frame;
} hl_counters;

typedef struct hl_build_stats
{
    uint32_t triangles;
    uint32_t wide_nodes;
    uint32_t binary_nodes;
    float    ms_build;     /* device time of the whole BLAS build (CUDA events) */
    float    sah_cost;     /* SAH cost of the wide BVH per unit root area (node visit = 1, triangle test = 0.35) */
    uint64_t bytes_nodes;
    uint64_t bytes_triangles;
} hl_build_stats;

typedef struct hl_context_t* hl_context;
typedef struct hl_mesh_t*    hl_mesh;

/* ------------------------------------------------------------------ context */
/* replaces vk::Backend::create + Renderer::create_output_images (src/engine/gfx/renderer.cpp:1483-1503):
 * allocates the accumulation image (RGBA32F, W x H), the RGBA8 tone-map target and the ray queues. */
HL_API hl_status hl_context_create(int device_ordinal, uint32_t width, uint32_t height, hl_context* out_ctx);
HL_API hl_status hl_context_destroy(hl_context ctx);
/* Renderer::on_window_resize (src/engine/gfx/renderer.cpp:715-726) */
HL_API hl_status hl_context_resize(hl_context ctx, uint32_t width, uint32_t height);
HL_API const char* hl_last_error(hl_context ctx /* may be NULL: last create error */);
HL_API const char* hl_version(void);

/* ------------------------------------------------------------------ resources */
/* replaces Mesh::create + BatchUploader::build_blas (src/engine/resource/mesh.cpp:27-116,
 * src/engine/gfx/vk.cpp:3160-3235): uploads VBO/IBO and builds the bottom-level BVH on the GPU. */
HL_API hl_status hl_mesh_create(hl_context ctx, const hl_vertex* vertices, uint32_t n_vertices,
                                const uint32_t* indices, uint32_t n_indices,
                                const hl_submesh* submeshes, uint32_t n_submeshes, hl_mesh* out_mesh);
HL_API hl_status hl_mesh_destroy(hl_context ctx, hl_mesh mesh);
HL_API hl_status hl_mesh_build_stats(hl_context ctx, hl_mesh mesh, hl_build_stats* out);

/* replaces Texture2D::create (include/resource/texture.h:39); returns the texture-array index
 * the Material table refers to (descriptor set 4, path_trace_rgen.glsl:58). */
HL_API hl_status hl_texture2d_create(hl_context ctx, int format, uint32_t width, uint32_t height,
                                     const void* level0_texels, int32_t* out_index);
HL_API hl_status hl_textures_clear(hl_context ctx);

/* environment cube map (set 0 binding 4): six faces +X,-X,+Y,-Y,+Z,-Z, each size*size RGBA32F,
 * row 0 = t 0.  size == 0 restores the reference's black default cube map (vk.cpp:3589-3612). */
HL_API hl_status hl_envmap_set(hl_context ctx, uint32_t size, const float* rgba32f_faces);
/* replaces HosekWilkieSkyModel::update's GPU bake (hosek_wilkie_sky_model.cpp:709-763,
 * procedural_sky.frag:48-75): coeffs = A,B,C,D,E,F,G,H,I,Z as vec4 (the 160-byte UBO). */
HL_API hl_status hl_sky_update(hl_context ctx, const float coeffs[40], const float sun_direction[3]);
HL_API hl_status hl_envmap_read(hl_context ctx, float* rgba32f_faces /* 6*size*size*4 */, uint32_t* out_size);

/* replaces Scene::create_gpu_resources + the TLAS build (src/engine/resource/scene.cpp:915-1311,
 * src/engine/gfx/renderer.cpp:113-181).  instance i uses meshes[i] (== the mesh whose global index is
 * instances[i].mesh_index) and submesh_info[i] = n_submeshes(mesh) x (base_index/3, material index). */
HL_API hl_status hl_scene_set_tables(hl_context ctx, const hl_material* materials, uint32_t n_materials,
                                     const hl_instance* instances, const hl_mesh* meshes,
                                     const uint32_t* const* submesh_info, uint32_t n_instances,
                                     const hl_light* lights, uint32_t n_lights);

/* Moving instances: new model / normal matrices for the n_instances instances of the installed tables (same order; each
 * instance keeps its mesh and submesh table — mesh_index is ignored; materials, lights and textures stay).  Replaces the
 * reference's per-change top-level rebuild (the structure is created ALLOW_UPDATE, src/engine/resource/scene.cpp:797, and
 * rebuilt in src/engine/gfx/renderer.cpp:147-168) by a REFIT of the instance tree: topology kept, node boxes recomputed
 * and requantised on the GPU in one launch; world->object transforms are recomputed.  Results are those of a rebuild
 * (closest hit + tie rule do not depend on the tree).  Returns HL_ERR_INVALID_ARGUMENT when the count differs. */
HL_API hl_status hl_scene_update_instances(hl_context ctx, const hl_instance* instances, uint32_t n_instances);

/* ------------------------------------------------------------------ the hot path */
/* replaces PathIntegrator::launch_rays -> vkCmdTraceRaysKHR (path_integrator.cpp:125-200): one sample
 * per pixel over the launch rectangle [tile, tile + (launch_w, launch_h)) clipped to (W, H), blended
 * into the accumulation image exactly as path_trace_rgen.glsl:217-248.  launch_w/h = 0 -> full frame.
 * Asynchronous on the context's stream. */
HL_API hl_status hl_render_frame(hl_context ctx, const hl_push_constants* pc, uint32_t launch_w, uint32_t launch_h);
/* The same frame with the progressive blend and the tone map fused into ONE resolve pass (reads the frame's radiance
 * and the previous accumulation, writes RGBA32F + RGBA8): Renderer::render = PathIntegrator::render + tone_map
 * (renderer.cpp:225-330, :369-428) in a single call.  In HL_ACCUM_SUM mode (full-frame launches only) the image shown
 * is this GPU's own sum / n, n = full-frame launches since hl_accum_clear — the rank-local progressive preview; the
 * final picture of a sharded render comes from the reduction + hl_tonemap(sample_scale = 1 / total).  Asynchronous. */
HL_API hl_status hl_render_frame_tonemapped(hl_context ctx, const hl_push_constants* pc, uint32_t launch_w, uint32_t launch_h,
                                            float exposure, int tone_map_operator);
/* The same, plus an asynchronous device->host copy of the frame's RGBA8 image into rgba8_host (W*H*4 bytes, pinned memory
 * for a truly asynchronous copy) on the frame's own stream: the read-back of frame f overlaps the rendering of frame
 * f + 1 (the reference's save path also trails the frame, renderer.cpp:637-711).  The image is complete after
 * hl_synchronize(); the caller must not reuse rgba8_host before that: with N frames in flight (HL_OPT_FRAMES_IN_FLIGHT,
 * default 4) a ring of N host buffers is safe, because frame f + N runs on frame f's stream, behind its copy.  Full-frame
 * or tiled launches. */
HL_API hl_status hl_render_frame_readback(hl_context ctx, const hl_push_constants* pc, uint32_t launch_w, uint32_t launch_h,
                                          float exposure, int tone_map_operator, uint8_t* rgba8_host);
/* copies the RGBA8 target written by the last hl_tonemap / hl_render_frame_tonemapped to host memory (synchronises) */
HL_API hl_status hl_read_rgba8(hl_context ctx, uint8_t* rgba8_host);
/* Renderer::render's restart branch (renderer.cpp:212-223): clears the accumulation image. */
HL_API hl_status hl_accum_clear(hl_context ctx);
HL_API hl_status hl_set_accum_mode(hl_context ctx, int mode);
/* parity hook (BASELINE config 5): closest hit of the primary ray of every pixel; arrays of W*H,
 * miss = 0xFFFFFFFF ids and t = +inf.  Any pointer may be NULL. */
HL_API hl_status hl_trace_primary_ids(hl_context ctx, const hl_push_constants* pc, uint32_t* instance,
                                      uint32_t* geometry, uint32_t* primitive, float* t, float* u, float* v);
/* debug output buffers: Renderer::set_current_output_buffer (include/gfx/renderer.h:25-33, src/engine/gfx/renderer.cpp:
 * 459-547 + debug_visualization.frag:144-161 in the reference, a rasterised view of one material channel).  Here:
 * the channel of the surface hit by each pixel's primary ray of `pc` (camera rays as in hl_render_frame), RGBA32F,
 * row 0 = v 0; pixels that see no surface are (0,0,0,1).  Normals are encoded n * 0.5 + 0.5 as in the reference. */
#define HL_OUTPUT_BUFFER_ALBEDO 0
#define HL_OUTPUT_BUFFER_NORMALS 1
#define HL_OUTPUT_BUFFER_ROUGHNESS 2
#define HL_OUTPUT_BUFFER_METALLIC 3
#define HL_OUTPUT_BUFFER_EMISSIVE 4
HL_API hl_status hl_render_output_buffer(hl_context ctx, const hl_push_constants* pc, int output_buffer, float* rgba32f_host);
/* ray debug view: PathIntegrator::gather_debug_rays (src/engine/gfx/path_integrator.cpp:88-104) + the RAY_DEBUG_VIEW variant
 * of the pipeline (:259-307; path_trace_rgen.glsl:137-147,193-195, path_trace_rchit.glsl:500-514,548-567,
 * path_trace_rmiss.glsl:40-58) + the vertex / draw-argument buffers of Renderer::create_ray_debug_buffers
 * (src/engine/gfx/renderer.cpp:1473-1479).  num_debug_rays paths (launch ids (tile x + i, tile y)) start through pixel
 * pc->ray_debug_pixel_coord.xy (extent .zw), are never ended by Russian roulette, and every ray after the primary one
 * leaves a line segment = two vertices (origin; hit point, or origin + direction * 10000 on a miss) in the path's colour.
 * Writes min(*vertex_count, max_vertices) vertices (segment order is unspecified, as in the reference: atomicAdd);
 * *vertex_count is the full count (the reference's DebugRayDrawArgs.count; its buffer holds MAX_DEBUG_RAY_DRAW_COUNT * 2
 * = 2048 vertices, include/gfx/renderer.h:9).  Synchronous. */
typedef struct hl_debug_ray_vertex
{
    float position[4];
    float color[4];
} hl_debug_ray_vertex; /* 32 B, DebugRayVertex common.glsl:54-58 */
HL_API hl_status hl_gather_debug_rays(hl_context ctx, const hl_push_constants* pc, uint32_t num_debug_rays,
                                      hl_debug_ray_vertex* vertices_host, uint32_t max_vertices, uint32_t* vertex_count);
/* generic closest-hit / visibility query on caller-supplied rays (8 floats per ray: o.xyz, tmin, d.xyz, tmax);
 * flags: bit0 = opaque (skip any-hit), bit1 = terminate on first hit. hit = 6 x 4 B per ray:
 * t,u,v (float) instance,geometry,primitive (u32). */
HL_API hl_status hl_trace_rays(hl_context ctx, const float* rays, uint32_t n_rays, uint32_t flags, void* hits);
/* replaces Renderer::tone_map (renderer.cpp:369-428, tone_map.frag): RGBA8, row 0 = top of the image
 * (the reference's negative-height viewport).  sample_scale multiplies the accumulation value first
 * (1 for the running mean; 1/count for HL_ACCUM_SUM).  rgba8_host may be NULL (device-only). */
HL_API hl_status hl_tonemap(hl_context ctx, float exposure, int tone_map_operator, float sample_scale, uint8_t* rgba8_host);
HL_API hl_status hl_read_accum(hl_context ctx, float* rgba32f_host);
HL_API hl_status hl_write_accum(hl_context ctx, const float* rgba32f_host);
/* device pointer of the accumulation image (W*H*4 floats) for NCCL reduction by the host layer */
HL_API hl_status hl_accum_device_ptr(hl_context ctx, void** out_ptr);
HL_API hl_status hl_synchronize(hl_context ctx);
/* fails with HL_ERR_LIMIT when a traversal ran out of stack since the last hl_reset_counters (hl_counters is still filled):
 * results of such rays are not trustworthy (hl_bvh.h TravStack; depth 64 entries) */
HL_API hl_status hl_get_counters(hl_context ctx, hl_counters* out);
/* per-bounce breakdown of the last frame rendered with hl_set_profiling(1): queue sizes and CUDA-event times of each
 * stage launch.  ms_extend brackets k_extend alone (the tail kernel has its own column).  tail_* = rays traced by the
 * tail kernel (all bounces it finished).  Returns the number of bounces in *n_bounces (<= capacity are written). */
typedef struct hl_bounce_profile
{
    uint32_t extension_rays; /* size of the extension queue this bounce's k_extend launch found (0: empty launch) */
    uint32_t shadow_rays;
    float    ms_tail, ms_extend, ms_shade, ms_connect;
} hl_bounce_profile;
HL_API hl_status hl_get_bounce_profile(hl_context ctx, hl_bounce_profile* out, uint32_t capacity, uint32_t* n_bounces,
                                       uint64_t* tail_extension_rays, uint64_t* tail_shadow_rays);
HL_API hl_status hl_reset_counters(hl_context ctx);
/* per-stage CUDA-event timing on/off (off by default: no events inside the frame) */
HL_API hl_status hl_set_profiling(hl_context ctx, int enabled);
/* CUDA-event stopwatch on the context's stream (slots 0..7): the library launches on its own stream, which
 * torch.cuda.Event does not observe */
HL_API hl_status hl_event_record(hl_context ctx, int slot);
HL_API hl_status hl_event_elapsed_ms(hl_context ctx, int slot_begin, int slot_end, float* out_ms); /* synchronises on slot_end */
/* tuning knobs of the wavefront scheduler (results do not depend on them) */
#define HL_OPT_TAIL_THRESHOLD 1 /* queue size at or below which the late bounces are finished by one per-path kernel; 0 = never */
#define HL_OPT_TAIL_START 2     /* first bounce at which that switch may happen (>= 1) */
#define HL_OPT_PIPELINE 3       /* 1 (default): consecutive frames rotate through several wavefront state slots / streams so the
                                   sparse late bounces of frame f overlap the first bounces of the frames after it (blends stay
                                   in frame order); 0: one frame at a time */
#define HL_OPT_FRAMES_IN_FLIGHT 5 /* number of those slots, 1..8 (default 4; the reference keeps up to 3 frames in flight,
                                   include/gfx/vk.h:66); each costs 172 B per pixel of device memory */
#define HL_OPT_CUDA_GRAPH 6     /* 1 (default): the bounce loop of a pipelined frame is replayed from a CUDA graph (one launch instead
                                   of ~30; captured per frame slot, re-captured when scene tables or integrator settings change) */
/* builder knob (applies to meshes / scene tables created afterwards; changes the tree, never a traversal result) */
#define HL_OPT_SAH_CLUSTER 4    /* binned-SAH re-split of the LBVH above a cut: primitives per cluster below the cut
                                   (default 2; larger = faster build, coarser refinement); 0 = plain LBVH topology.  Triangle
                                   meshes with 1..8 primitives per cluster are re-split in two levels (a level loop over coarse
                                   subtrees, then one thread block per treelet of <= 512 primitives; environment variable
                                   HL_NO_TREELETS=1: the one-level re-split, for A/B runs); larger settings and instance trees
                                   take the one-level re-split */
HL_API hl_status hl_set_option(hl_context ctx, int option, int64_t value);
/* number of kernels launched by this library on this context since creation */
HL_API hl_status hl_kernel_launches(hl_context ctx, uint64_t* out);

/* ------------------------------------------------------------------ multi-GPU (SURVEY.md 8e / 8b hl_multi_gpu_reduce)
 * Samples per pixel are sharded across the GPUs of one box: the scene is replicated, rank g renders frame indices
 * g+1, g+1+G, ... into its own HL_ACCUM_SUM image, and ONE reduction of the W*H*4-float accumulation images combines
 * them; 1 / samples goes into the tone-map pass (sample_scale).  The reference has no multi-GPU path: what is kept is
 * its blend (path_trace_rgen.glsl:219-247) — sum / count equals that running mean up to fp32 rounding.
 *
 * (1) one process per GPU (torchrun / MPI style): rank 0 calls hl_comm_unique_id and hands the HL_COMM_ID_BYTES bytes to
 *     the other ranks by any means; every rank calls hl_comm_init_rank (collective, blocking), then hl_accum_all_reduce /
 *     hl_accum_reduce — NCCL on the context's stream, asynchronous like hl_render_frame, ordered behind the frames in
 *     flight.  libnccl.so.2 is opened at run time (no link-time dependency); HL_ERR_STATE when it cannot be loaded.
 * (2) all GPUs in one process (helios_headless --gpus N): hl_comm_init_all binds n contexts (rank = position in the
 *     array; contexts may also share a device), hl_multi_gpu_reduce sums their images into the root's, and
 *     hl_multi_gpu_resolve does that plus tone map plus read-back.  With peer access between all members both run as ONE
 *     kernel per GPU over peer memory (each GPU reduces 1/n of the image in rank order and writes the fp32 sum and the
 *     tone-mapped RGBA8 pixels straight into the root's images; CUDA events order it, nothing blocks the host);
 *     otherwise they fall back to ncclReduce + the root's tone-map pass. */
#define HL_COMM_ID_BYTES 128
HL_API hl_status hl_comm_unique_id(uint8_t* id /* HL_COMM_ID_BYTES */);
HL_API hl_status hl_comm_init_rank(hl_context ctx, const uint8_t* id, int n_ranks, int rank);
HL_API hl_status hl_comm_init_all(hl_context* ctxs, int n);
HL_API hl_status hl_comm_destroy(hl_context ctx);
HL_API const char* hl_comm_last_error(void); /* message of the last failed hl_comm_* / hl_multi_gpu_* call of this thread */
/* accumulation image <- sum over ranks, on every rank / on `root` only (other ranks keep their own image) */
HL_API hl_status hl_accum_all_reduce(hl_context ctx);
HL_API hl_status hl_accum_reduce(hl_context ctx, int root);
/* single-process form: root's accumulation image <- sum over the n contexts; asynchronous on the contexts' streams */
HL_API hl_status hl_multi_gpu_reduce(hl_context* ctxs, int n, int root);
/* the same, fused with Renderer::tone_map (tone_map.frag) of sum * sample_scale into the root's RGBA8 image; rgba8_host may be
 * NULL (device only: hl_read_rgba8(root) fetches it later), otherwise the call copies the image there and synchronises */
HL_API hl_status hl_multi_gpu_resolve(hl_context* ctxs, int n, int root, float exposure, int tone_map_operator, float sample_scale,
                                      uint8_t* rgba8_host);

#ifdef __cplusplus
}
#endif
#endif /* HELIOS_B200_H */
