// hl_api.cu — the extern "C" boundary declared in include/helios_b200.h.  No exception crosses it: every entry
// point catches, stores the message in the context and returns an hl_status.  There is no CPU path: without
// a CUDA device hl_context_create fails with HL_ERR_NO_DEVICE.
#include "hl_internal.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>

using namespace hl;

static thread_local std::string g_create_error;

// HL_TRY makes the main stream wait for the frames in flight (hl_wave_slot) before the call enqueues anything;
// HL_TRY_FRAME is for the frame entry points themselves, which order their own streams.
#define HL_TRY_FRAME(ctx_)                                \
    hl_context_t* c_ = (ctx_);                            \
    if (!c_) return HL_ERR_INVALID_ARGUMENT;              \
    try                                                   \
    {                                                     \
        HL_CUDA(cudaSetDevice(c_->device));
#define HL_TRY(ctx_)   \
    HL_TRY_FRAME(ctx_) \
    hl::wavefront_join(c_);
#define HL_CATCH                                          \
    }                                                     \
    catch (const hl::CudaError& e)                        \
    {                                                     \
        c_->err = e.what();                               \
        return e.status;                                  \
    }                                                     \
    catch (const std::bad_alloc&)                         \
    {                                                     \
        c_->err = "host out of memory";                   \
        return HL_ERR_OUT_OF_MEMORY;                      \
    }                                                     \
    catch (const std::exception& e)                       \
    {                                                     \
        c_->err = e.what();                               \
        return HL_ERR_CUDA;                               \
    }                                                     \
    return HL_OK;
#define HL_FAIL(code, msg)       \
    do                           \
    {                            \
        c_->err = (msg);         \
        return (code);           \
    } while (0)

extern "C" {

const char* hl_version(void) { return "helios_b200 0.1 (sm_100a)"; }

const char* hl_last_error(hl_context ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

hl_status hl_context_create(int device_ordinal, uint32_t width, uint32_t height, hl_context* out_ctx)
{
    if (!out_ctx || width == 0 || height == 0)
    {
        g_create_error = "hl_context_create: invalid argument";
        return HL_ERR_INVALID_ARGUMENT;
    }
    *out_ctx  = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device_ordinal < 0 || device_ordinal >= count)
    {
        cudaGetLastError();
        g_create_error = "hl_context_create: no CUDA device (this library has no CPU path)";
        return HL_ERR_NO_DEVICE;
    }
    hl_context_t* c = new hl_context_t();
    try
    {
        c->device = device_ordinal;
        HL_CUDA(cudaSetDevice(device_ordinal));
        HL_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        HL_CUDA(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device_ordinal));
        {
            // build scratch comes from a PRIVATE stream-ordered pool (ScratchBuf): freed scratch up to 1 GiB stays in it for
            // the next build, the rest goes back to the driver; the device's default pool (other cudaMallocAsync users of
            // the process: torch, NCCL) is left alone
            cudaMemPoolProps props;
            memset(&props, 0, sizeof(props));
            props.allocType = cudaMemAllocationTypePinned, props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice, props.location.id = device_ordinal;
            HL_CUDA(cudaMemPoolCreate(&c->scratch_pool, &props));
            uint64_t keep = 1ull << 30;
            HL_CUDA(cudaMemPoolSetAttribute(c->scratch_pool, cudaMemPoolAttrReleaseThreshold, &keep));
        }
        c->W = width, c->H = height;
        // 8-bit decode tables (unorm, srgb, snorm), computed in double like the oracle's
        float lut[768];
        for (int i = 0; i < 256; i++)
        {
            const double v = i / 255.0;
            lut[i]         = (float)v;
            lut[256 + i]   = (float)(v <= 0.04045 ? v / 12.92 : std::pow((v + 0.055) / 1.055, 2.4));
            const int sn   = (int8_t)(uint8_t)i;
            lut[512 + i]   = (float)std::max(-1.0, sn / 127.0);
        }
        c->lut8.upload(lut, sizeof(lut), c->stream);
        wavefront_alloc(c);
        HL_CUDA(cudaStreamSynchronize(c->stream));
    }
    catch (const std::exception& e)
    {
        g_create_error = e.what();
        delete c;
        return HL_ERR_CUDA;
    }
    *out_ctx = c;
    return HL_OK;
}

hl_status hl_context_destroy(hl_context ctx)
{
    if (!ctx) return HL_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    hl::comm_release(ctx);
    for (auto m : ctx->meshes) delete m;
    for (auto t : ctx->textures) delete t;
    if (ctx->ev_ready)
        for (auto& e : ctx->ev) cudaEventDestroy(e);
    if (ctx->user_ev_ready)
        for (auto& e : ctx->user_ev) cudaEventDestroy(e);
    for (hl_wave_slot& w : ctx->slot)
    {
        if (w.stream) cudaStreamSynchronize(w.stream);
        if (w.graph_exec) cudaGraphExecDestroy(w.graph_exec);
        if (w.stream) cudaStreamDestroy(w.stream);
        if (w.copy_stream) cudaStreamSynchronize(w.copy_stream), cudaStreamDestroy(w.copy_stream);
        if (w.image_ready) cudaEventDestroy(w.image_ready);
        if (w.copy_done) cudaEventDestroy(w.copy_done);
        if (w.resolved) cudaEventDestroy(w.resolved);
    }
    if (ctx->main_ev) cudaEventDestroy(ctx->main_ev);
    cudaStreamDestroy(ctx->stream);
    cudaMemPool_t pool = ctx->scratch_pool;
    delete ctx; // (frees the device buffers)
    if (pool) cudaMemPoolDestroy(pool);
    return HL_OK;
}

hl_status hl_context_resize(hl_context ctx, uint32_t width, uint32_t height)
{
    HL_TRY(ctx)
    if (width == 0 || height == 0) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_context_resize: zero extent");
    HL_CUDA(cudaStreamSynchronize(c_->stream));
    c_->W = width, c_->H = height;
    wavefront_release(c_);
    wavefront_alloc(c_);
    HL_CATCH
}

hl_status hl_mesh_create(hl_context ctx, const hl_vertex* vertices, uint32_t n_vertices, const uint32_t* indices, uint32_t n_indices, const hl_submesh* submeshes, uint32_t n_submeshes,
                         hl_mesh* out_mesh)
{
    HL_TRY(ctx)
    if (!out_mesh || (!vertices && n_vertices) || (!indices && n_indices) || (!submeshes && n_submeshes)) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_mesh_create: null argument");
    *out_mesh = nullptr;
    uint64_t tris = 0;
    for (uint32_t g = 0; g < n_submeshes; g++)
    {
        if ((uint64_t)submeshes[g].base_index + submeshes[g].index_count > n_indices) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_mesh_create: submesh index range exceeds the index buffer");
        tris += submeshes[g].index_count / 3;
    }
    if (tris > 0x7FFFFFFFull) HL_FAIL(HL_ERR_LIMIT, "hl_mesh_create: more than 2^31 triangles");
    for (uint32_t i = 0; i < n_indices; i++)
        if (indices[i] >= n_vertices) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_mesh_create: index out of range");
    hl_mesh_t* m = new hl_mesh_t();
    try
    {
        m->subs.assign(submeshes, submeshes + n_submeshes);
        m->n_vertices = n_vertices, m->n_indices = n_indices;
        m->vertices.upload(vertices, sizeof(hl_vertex) * (size_t)n_vertices, c_->stream);
        m->indices.upload(indices, 4ull * n_indices, c_->stream);
        build_mesh_bvh(c_, m);
    }
    catch (...)
    {
        delete m;
        throw;
    }
    c_->meshes.push_back(m);
    c_->scene_ready = false;
    *out_mesh       = m;
    HL_CATCH
}

hl_status hl_mesh_destroy(hl_context ctx, hl_mesh mesh)
{
    HL_TRY(ctx)
    auto it = std::find(c_->meshes.begin(), c_->meshes.end(), mesh);
    if (it == c_->meshes.end()) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_mesh_destroy: unknown mesh");
    HL_CUDA(cudaStreamSynchronize(c_->stream));
    delete *it;
    c_->meshes.erase(it);
    c_->scene_ready = false;
    HL_CATCH
}

hl_status hl_mesh_build_stats(hl_context ctx, hl_mesh mesh, hl_build_stats* out)
{
    HL_TRY(ctx)
    if (!mesh || !out) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_mesh_build_stats: null argument");
    *out = mesh->stats;
    HL_CATCH
}

hl_status hl_texture2d_create(hl_context ctx, int format, uint32_t width, uint32_t height, const void* texels, int32_t* out_index)
{
    HL_TRY(ctx)
    if (!texels || !out_index || width == 0 || height == 0 || format < 0 || format > 3) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_texture2d_create: invalid argument");
    if (c_->textures.size() >= HL_MAX_SCENE_MATERIAL_TEXTURE_COUNT) HL_FAIL(HL_ERR_LIMIT, "hl_texture2d_create: MAX_SCENE_MATERIAL_TEXTURE_COUNT exceeded");
    const uint64_t bytes = (uint64_t)width * height * (format == HL_TEX_RGBA32F ? 16 : 4);
    if (width > 65536u || height > 65536u || bytes > (1ull << 34)) HL_FAIL(HL_ERR_LIMIT, "hl_texture2d_create: extent above 65536 or level above 16 GiB");
    std::unique_ptr<DevBuf> b(new DevBuf());
    b->upload(texels, (size_t)bytes, c_->stream);
    HL_CUDA(cudaStreamSynchronize(c_->stream));
    TexView v;
    v.texels = b->p, v.w = width, v.h = height, v.format = format, v.pad = 0;
    c_->tex_views.reserve(c_->tex_views.size() + 1); // the two vectors change together or not at all
    c_->textures.push_back(b.get());
    b.release();
    c_->tex_views.push_back(v);
    *out_index      = (int32_t)c_->tex_views.size() - 1;
    c_->scene_ready = false;
    HL_CATCH
}

hl_status hl_textures_clear(hl_context ctx)
{
    HL_TRY(ctx)
    HL_CUDA(cudaStreamSynchronize(c_->stream));
    for (auto t : c_->textures) delete t;
    c_->textures.clear(), c_->tex_views.clear();
    c_->scene_ready = false;
    HL_CATCH
}

hl_status hl_envmap_set(hl_context ctx, uint32_t size, const float* faces)
{
    HL_TRY(ctx)
    if (size && !faces) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_envmap_set: null faces");
    c_->env_size = size;
    if (size)
    {
        c_->env_faces.upload(faces, (size_t)6 * size * size * 16, c_->stream);
        HL_CUDA(cudaStreamSynchronize(c_->stream));
    }
    env_pad(c_);
    HL_CATCH
}

hl_status hl_sky_update(hl_context ctx, const float coeffs[40], const float sun_direction[3])
{
    HL_TRY(ctx)
    if (!coeffs || !sun_direction) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_sky_update: null argument");
    sky_bake(c_, coeffs, sun_direction, 512); // SKY_CUBEMAP_SIZE, hosek_wilkie_sky_model.cpp:17
    env_pad(c_);
    HL_CATCH
}

hl_status hl_envmap_read(hl_context ctx, float* faces, uint32_t* out_size)
{
    HL_TRY(ctx)
    if (out_size) *out_size = c_->env_size;
    if (faces && c_->env_size)
    {
        HL_CUDA(cudaMemcpyAsync(faces, c_->env_faces.p, (size_t)6 * c_->env_size * c_->env_size * 16, cudaMemcpyDeviceToHost, c_->stream));
        HL_CUDA(cudaStreamSynchronize(c_->stream));
    }
    HL_CATCH
}

// world -> object 3x4 (row-major) from the column-major model matrix: double precision, rounded once
static void affine_inverse(const float* m, float* out)
{
    double a00 = m[0], a01 = m[4], a02 = m[8], t0 = m[12];
    double a10 = m[1], a11 = m[5], a12 = m[9], t1 = m[13];
    double a20 = m[2], a21 = m[6], a22 = m[10], t2 = m[14];
    double c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
    double det = a00 * c00 + a01 * c01 + a02 * c02;
    double id  = 1.0 / det;
    double i00 = c00 * id, i01 = (a02 * a21 - a01 * a22) * id, i02 = (a01 * a12 - a02 * a11) * id;
    double i10 = c01 * id, i11 = (a00 * a22 - a02 * a20) * id, i12 = (a02 * a10 - a00 * a12) * id;
    double i20 = c02 * id, i21 = (a01 * a20 - a00 * a21) * id, i22 = (a00 * a11 - a01 * a10) * id;
    out[0] = (float)i00, out[1] = (float)i01, out[2] = (float)i02, out[3] = (float)(-(i00 * t0 + i01 * t1 + i02 * t2));
    out[4] = (float)i10, out[5] = (float)i11, out[6] = (float)i12, out[7] = (float)(-(i10 * t0 + i11 * t1 + i12 * t2));
    out[8] = (float)i20, out[9] = (float)i21, out[10] = (float)i22, out[11] = (float)(-(i20 * t0 + i21 * t1 + i22 * t2));
}

// world box of an instance: the 8 corners of the mesh's object-space root box through the model matrix, in double, padded
static Box instance_world_box(const float* M, const Box& rb)
{
    double lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
    for (int cidx = 0; cidx < 8; cidx++)
    {
        const double x = (cidx & 1) ? rb.hi[0] : rb.lo[0], y = (cidx & 2) ? rb.hi[1] : rb.lo[1], z = (cidx & 4) ? rb.hi[2] : rb.lo[2];
        for (int a = 0; a < 3; a++)
        {
            const double w = (double)M[a] * x + (double)M[4 + a] * y + (double)M[8 + a] * z + (double)M[12 + a];
            lo[a] = std::min(lo[a], w), hi[a] = std::max(hi[a], w);
        }
    }
    Box out;
    for (int a = 0; a < 3; a++)
    {
        const double pad = 1e-5 * (std::fabs(lo[a]) + std::fabs(hi[a]) + (hi[a] - lo[a]));
        out.lo[a] = (float)(lo[a] - pad), out.hi[a] = (float)(hi[a] + pad);
    }
    return out;
}
static const float kIdentity16[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };

// bit i set: instance i's model matrix is the identity (hl_bvh.h skips the ray transform for it)
static void upload_identity_bits(hl_context_t* c, const std::vector<hl_instance>& inst)
{
    std::vector<uint32_t> bits((inst.size() + 31) / 32 + 1, 0u);
    for (size_t i = 0; i < inst.size(); i++)
        if (memcmp(inst[i].model_matrix, kIdentity16, 64) == 0) bits[i >> 5] |= 1u << (i & 31);
    c->inst_identity.upload(bits.data(), 4 * bits.size(), c->stream);
    HL_CUDA(cudaStreamSynchronize(c->stream));
    c->view.inst_identity = c->inst_identity.as<uint32_t>();
}

hl_status hl_scene_set_tables(hl_context ctx, const hl_material* materials, uint32_t n_materials, const hl_instance* instances, const hl_mesh* meshes,
                              const uint32_t* const* submesh_info, uint32_t n_instances, const hl_light* lights, uint32_t n_lights)
{
    HL_TRY(ctx)
    if ((!materials && n_materials) || (n_instances && (!instances || !meshes || !submesh_info)) || (!lights && n_lights)) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_scene_set_tables: null argument");
    if (n_instances > HL_MAX_SCENE_MESH_INSTANCE_COUNT) HL_FAIL(HL_ERR_LIMIT, "hl_scene_set_tables: MAX_SCENE_MESH_INSTANCE_COUNT (1024) exceeded");
    if (n_materials > HL_MAX_SCENE_MATERIAL_COUNT) HL_FAIL(HL_ERR_LIMIT, "hl_scene_set_tables: MAX_SCENE_MATERIAL_COUNT (4096) exceeded");
    if (n_lights > HL_MAX_SCENE_LIGHT_COUNT) HL_FAIL(HL_ERR_LIMIT, "hl_scene_set_tables: MAX_SCENE_LIGHT_COUNT (100000) exceeded");
    // every index a shader follows is checked here: a bad one would be a device out-of-bounds read, and the resulting
    // cudaErrorIllegalAddress is sticky for the whole process (load_surface, any_hit_ignores, sample_light: hl_shade.h, hl_bvh.h)
    const int32_t n_tex = (int32_t)c_->tex_views.size();
    for (uint32_t m = 0; m < n_materials; m++)
    {
        const int32_t ti[5] = { materials[m].texture_indices0[0], materials[m].texture_indices0[1], materials[m].texture_indices0[2], materials[m].texture_indices0[3], materials[m].texture_indices1[0] };
        for (int k = 0; k < 5; k++)
            if (ti[k] < -1 || ti[k] >= n_tex) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_scene_set_tables: material " + std::to_string(m) + " refers to texture " + std::to_string(ti[k]) + " (" + std::to_string(n_tex) + " textures exist)");
    }
    for (uint32_t l = 0; l < n_lights; l++)
    {
        const hl_light& L = lights[l];
        if (!(L.light_data0[0] == (float)HL_LIGHT_AREA)) continue;
        // sample_light reads: instance = data0.y, material = data0.z, first primitive = data0.w, primitive count = data1.z
        const float f[4] = { L.light_data0[1], L.light_data0[2], L.light_data0[3], L.light_data1[2] };
        for (int k = 0; k < 4; k++)
            if (!(f[k] >= 0.0f && f[k] < 2147483648.0f)) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_scene_set_tables: area light " + std::to_string(l) + " has a negative or non-finite index field");
        const uint32_t li = (uint32_t)f[0], lm = (uint32_t)f[1], lp = (uint32_t)f[2], lc = (uint32_t)f[3];
        if (li >= n_instances) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_scene_set_tables: area light " + std::to_string(l) + " refers to instance " + std::to_string(li));
        if (lm >= n_materials) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_scene_set_tables: area light " + std::to_string(l) + " refers to material " + std::to_string(lm));
        auto it = std::find(c_->meshes.begin(), c_->meshes.end(), meshes[li]);
        if (it == c_->meshes.end()) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_scene_set_tables: instance refers to an unknown mesh");
        if ((uint64_t)lp + std::max(lc, 1u) > (uint64_t)(*it)->n_indices / 3) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_scene_set_tables: area light " + std::to_string(l) + " primitive range exceeds its mesh");
    }
    cudaStream_t st = c_->stream;
    HL_CUDA(cudaStreamSynchronize(st)); // Scene::create_gpu_resources calls wait_idle (scene.cpp:929)
    // mesh table in context order; instance.mesh_index is rewritten to that order
    std::vector<MeshView> views(c_->meshes.size());
    for (size_t k = 0; k < c_->meshes.size(); k++)
    {
        hl_mesh_t* m = c_->meshes[k];
        MeshView&  v = views[k];
        v.vertices = m->vertices.as<hl_vertex>(), v.indices = m->indices.as<uint32_t>();
        v.nodes = m->bvh.nodes.as<WideNode>(), v.tris = m->bvh.leaves.as<LeafTri>();
        v.n_tris = m->bvh.n_leaves, v.n_submeshes = (uint32_t)m->subs.size();
    }
    std::vector<hl_instance> inst(instances, instances + n_instances);
    std::vector<float>       inv((size_t)n_instances * 12);
    std::vector<uint32_t>    info, offset(n_instances);
    std::vector<InstAlpha>   ialpha(n_instances);
    std::vector<GeomAlpha>   galpha;
    std::vector<Box>         boxes(n_instances);
    bool                     identity = n_instances == 1;
    const float*             I16      = kIdentity16;
    for (uint32_t i = 0; i < n_instances; i++)
    {
        auto it = std::find(c_->meshes.begin(), c_->meshes.end(), meshes[i]);
        if (it == c_->meshes.end()) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_scene_set_tables: instance refers to an unknown mesh");
        hl_mesh_t* m       = *it;
        inst[i].mesh_index = (uint32_t)(it - c_->meshes.begin());
        offset[i]          = (uint32_t)(info.size() / 2);
        for (size_t g = 0; g < m->subs.size(); g++)
        {
            if (submesh_info[i][2 * g + 1] >= n_materials) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_scene_set_tables: material index out of range");
            info.push_back(submesh_info[i][2 * g]), info.push_back(submesh_info[i][2 * g + 1]);
            const hl_material& mat = materials[submesh_info[i][2 * g + 1]];
            GeomAlpha          ga;
            ga.texture = mat.texture_indices0[0], ga.alpha = mat.albedo[3];
            galpha.push_back(ga);
        }
        ialpha[i].alpha = m->alpha.p && m->bvh.n_leaves ? m->alpha.as<AlphaTri>() : nullptr, ialpha[i].tris = m->bvh.leaves.as<LeafTri>(), ialpha[i].info_base = offset[i];
        affine_inverse(inst[i].model_matrix, &inv[(size_t)i * 12]);
        if (memcmp(inst[i].model_matrix, I16, 64) != 0) identity = false;
        boxes[i] = instance_world_box(inst[i].model_matrix, m->bvh.root);
    }
    c_->h_instances = inst;
    c_->h_inst_mesh.resize(n_instances);
    for (uint32_t i = 0; i < n_instances; i++) c_->h_inst_mesh[i] = c_->meshes[inst[i].mesh_index];
    c_->materials.upload(materials, sizeof(hl_material) * (size_t)n_materials, st);
    c_->instances.upload(inst.data(), sizeof(hl_instance) * (size_t)n_instances, st);
    c_->inst_inv.upload(inv.data(), 4 * inv.size(), st);
    c_->submesh_info.upload(info.data(), 4 * info.size(), st);
    c_->submesh_offset.upload(offset.data(), 4 * offset.size(), st);
    c_->lights.upload(lights, sizeof(hl_light) * (size_t)n_lights, st);
    c_->inst_alpha.upload(ialpha.data(), sizeof(InstAlpha) * ialpha.size(), st);
    c_->geom_alpha.upload(galpha.data(), sizeof(GeomAlpha) * galpha.size(), st);
    c_->mesh_views.upload(views.data(), sizeof(MeshView) * views.size(), st);
    c_->tex_views_dev.upload(c_->tex_views.data(), sizeof(TexView) * c_->tex_views.size(), st);
    HL_CUDA(cudaStreamSynchronize(st));
    build_tlas(c_, boxes);
    SceneView& v = c_->view;
    v.materials = c_->materials.as<hl_material>(), v.instances = c_->instances.as<hl_instance>(), v.inst_inv = c_->inst_inv.as<float>();
    v.submesh_info = c_->submesh_info.as<uint32_t>(), v.submesh_offset = c_->submesh_offset.as<uint32_t>(), v.lights = c_->lights.as<hl_light>();
    v.meshes = c_->mesh_views.as<MeshView>(), v.textures = c_->tex_views_dev.as<TexView>(), v.lut8 = c_->lut8.as<float>();
    v.env.faces = c_->env_size ? c_->env_padded.as<f4>() : nullptr, v.env.size = c_->env_size;
    v.tlas_nodes = c_->tlas.nodes.as<WideNode>(), v.tlas_leaf = c_->tlas.leaves.as<uint32_t>();
    v.n_instances = n_instances, v.n_lights = n_lights, v.single_identity = identity ? 1u : 0u;
    v.inst_alpha = c_->inst_alpha.as<InstAlpha>(), v.geom_alpha = c_->geom_alpha.as<GeomAlpha>();
    upload_identity_bits(c_, inst);
    c_->scene_ready = true;
    HL_CATCH
}

hl_status hl_scene_update_instances(hl_context ctx, const hl_instance* instances, uint32_t n_instances)
{
    HL_TRY(ctx)
    if (!instances && n_instances) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_scene_update_instances: null argument");
    if (!c_->scene_ready) HL_FAIL(HL_ERR_STATE, "hl_scene_update_instances: hl_scene_set_tables has not been called since the last resource change");
    if (n_instances != c_->h_instances.size()) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_scene_update_instances: the instance count differs from the installed tables (use hl_scene_set_tables)");
    if (n_instances == 0) return HL_OK;
    cudaStream_t       st = c_->stream;
    std::vector<float> inv((size_t)n_instances * 12);
    std::vector<Box>   boxes(n_instances);
    bool               identity = n_instances == 1;
    for (uint32_t i = 0; i < n_instances; i++)
    {
        hl_instance& I = c_->h_instances[i];
        memcpy(I.model_matrix, instances[i].model_matrix, 64), memcpy(I.normal_matrix, instances[i].normal_matrix, 64); // the instance keeps its mesh
        affine_inverse(I.model_matrix, &inv[(size_t)i * 12]);
        if (memcmp(I.model_matrix, kIdentity16, 64) != 0) identity = false;
        boxes[i] = instance_world_box(I.model_matrix, c_->h_inst_mesh[i]->bvh.root);
    }
    c_->instances.upload(c_->h_instances.data(), sizeof(hl_instance) * (size_t)n_instances, st);
    c_->inst_inv.upload(inv.data(), 4 * inv.size(), st);
    HL_CUDA(cudaStreamSynchronize(st)); // (the staging vectors go out of scope)
    if (n_instances == 1 || c_->tlas.n_nodes > 1024u)
        build_tlas(c_, boxes); // nothing to keep / larger than the one-block refit handles
    else
        refit_tlas(c_, boxes);
    c_->view.tlas_nodes = c_->tlas.nodes.as<WideNode>(), c_->view.tlas_leaf = c_->tlas.leaves.as<uint32_t>();
    c_->view.single_identity = identity ? 1u : 0u;
    upload_identity_bits(c_, c_->h_instances);
    HL_CATCH
}

static bool clip_launch(hl_context_t* c, const hl_push_constants* pc, uint32_t& lw, uint32_t& lh)
{
    const uint32_t W = pc->launch_id_size[2], H = pc->launch_id_size[3];
    if (W != c->W || H != c->H) return false;
    const uint32_t tx = pc->launch_id_size[0], ty = pc->launch_id_size[1];
    if (lw == 0) lw = W;
    if (lh == 0) lh = H;
    lw = tx >= W ? 0 : std::min(lw, W - tx); // pixels with launch_id >= (W,H) do nothing (rgen:185)
    lh = ty >= H ? 0 : std::min(lh, H - ty);
    return true;
}

hl_status hl_render_frame(hl_context ctx, const hl_push_constants* pc, uint32_t launch_w, uint32_t launch_h)
{
    HL_TRY_FRAME(ctx)
    if (!pc) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_render_frame: null push constants");
    if (!c_->scene_ready) HL_FAIL(HL_ERR_STATE, "hl_render_frame: hl_scene_set_tables has not been called since the last resource change");
    if (!clip_launch(c_, pc, launch_w, launch_h)) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_render_frame: launch_id_size.zw differs from the context extent");
    if (pc->max_ray_bounces > HL_MAX_BOUNCES) HL_FAIL(HL_ERR_LIMIT, "hl_render_frame: max_ray_bounces > 64");
    wavefront_render_frame(c_, *pc, launch_w, launch_h);
    HL_CUDA(cudaGetLastError());
    HL_CATCH
}

hl_status hl_render_frame_tonemapped(hl_context ctx, const hl_push_constants* pc, uint32_t launch_w, uint32_t launch_h, float exposure, int op)
{
    HL_TRY_FRAME(ctx)
    if (!pc) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_render_frame_tonemapped: null push constants");
    if (!c_->scene_ready) HL_FAIL(HL_ERR_STATE, "hl_render_frame_tonemapped: hl_scene_set_tables has not been called since the last resource change");
    if (op != HL_TONE_MAP_ACES && op != HL_TONE_MAP_REINHARD) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_render_frame_tonemapped: unknown tone map operator");
    if (!clip_launch(c_, pc, launch_w, launch_h)) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_render_frame_tonemapped: launch_id_size.zw differs from the context extent");
    if (pc->max_ray_bounces > HL_MAX_BOUNCES) HL_FAIL(HL_ERR_LIMIT, "hl_render_frame_tonemapped: max_ray_bounces > 64");
    if (c_->accum_mode == HL_ACCUM_SUM && (launch_w != c_->W || launch_h != c_->H)) HL_FAIL(HL_ERR_STATE, "hl_render_frame_tonemapped: HL_ACCUM_SUM takes full-frame launches only (tiles: hl_render_frame + hl_tonemap)");
    ResolveOptions opt;
    opt.tone_map = true, opt.exposure = exposure, opt.op = op;
    wavefront_render_frame(c_, *pc, launch_w, launch_h, opt);
    HL_CUDA(cudaGetLastError());
    HL_CATCH
}

hl_status hl_render_frame_readback(hl_context ctx, const hl_push_constants* pc, uint32_t launch_w, uint32_t launch_h, float exposure, int op, uint8_t* rgba8_host)
{
    HL_TRY_FRAME(ctx)
    if (!pc || !rgba8_host) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_render_frame_readback: null argument");
    if (!c_->scene_ready) HL_FAIL(HL_ERR_STATE, "hl_render_frame_readback: hl_scene_set_tables has not been called since the last resource change");
    if (op != HL_TONE_MAP_ACES && op != HL_TONE_MAP_REINHARD) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_render_frame_readback: unknown tone map operator");
    if (!clip_launch(c_, pc, launch_w, launch_h)) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_render_frame_readback: launch_id_size.zw differs from the context extent");
    if (pc->max_ray_bounces > HL_MAX_BOUNCES) HL_FAIL(HL_ERR_LIMIT, "hl_render_frame_readback: max_ray_bounces > 64");
    if (c_->accum_mode == HL_ACCUM_SUM && (launch_w != c_->W || launch_h != c_->H)) HL_FAIL(HL_ERR_STATE, "hl_render_frame_readback: HL_ACCUM_SUM takes full-frame launches only (tiles: hl_render_frame + hl_tonemap)");
    ResolveOptions opt;
    opt.tone_map = true, opt.exposure = exposure, opt.op = op, opt.host = rgba8_host;
    wavefront_render_frame(c_, *pc, launch_w, launch_h, opt);
    HL_CUDA(cudaGetLastError());
    HL_CATCH
}

hl_status hl_read_rgba8(hl_context ctx, uint8_t* rgba8_host)
{
    HL_TRY(ctx)
    if (!rgba8_host) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_read_rgba8: null pointer");
    HL_CUDA(cudaMemcpyAsync(rgba8_host, c_->rgba8_cur ? c_->rgba8_cur : c_->rgba8.p, (size_t)c_->W * c_->H * 4, cudaMemcpyDeviceToHost, c_->stream));
    HL_CUDA(cudaStreamSynchronize(c_->stream));
    HL_CATCH
}

hl_status hl_accum_clear(hl_context ctx)
{
    HL_TRY(ctx)
    film_clear(c_);
    HL_CATCH
}

hl_status hl_set_accum_mode(hl_context ctx, int mode)
{
    HL_TRY(ctx)
    if (mode != HL_ACCUM_RUNNING_MEAN && mode != HL_ACCUM_SUM) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_set_accum_mode: unknown mode");
    c_->accum_mode = mode;
    HL_CATCH
}

hl_status hl_trace_primary_ids(hl_context ctx, const hl_push_constants* pc, uint32_t* instance, uint32_t* geometry, uint32_t* primitive, float* t, float* u, float* v)
{
    HL_TRY(ctx)
    if (!pc) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_trace_primary_ids: null push constants");
    if (!c_->scene_ready) HL_FAIL(HL_ERR_STATE, "hl_trace_primary_ids: scene tables not set");
    if (pc->launch_id_size[2] != c_->W || pc->launch_id_size[3] != c_->H) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_trace_primary_ids: extent mismatch");
    wavefront_primary_hits(c_, *pc);
    if (!instance && !geometry && !primitive && !t && !u && !v) return HL_OK; // device-only: results stay in the hit buffers
    const size_t       n = (size_t)c_->W * c_->H;
    std::vector<float> ha(n * 4);
    std::vector<uint32_t> hb(n * 2);
    HL_CUDA(cudaMemcpyAsync(ha.data(), c_->slot[0].hit_a.p, n * 16, cudaMemcpyDeviceToHost, c_->stream));
    HL_CUDA(cudaMemcpyAsync(hb.data(), c_->slot[0].hit_b.p, n * 8, cudaMemcpyDeviceToHost, c_->stream));
    HL_CUDA(cudaStreamSynchronize(c_->stream));
    for (size_t i = 0; i < n; i++)
    {
        const bool   hit = hb[2 * i] != HL_MISS;
        const size_t p   = wavefront_path_pixel(c_, (uint32_t)i); // the hit buffers are in path order (8 x 4 pixel tiles per warp)
        if (instance) instance[p] = hb[2 * i];
        if (geometry) geometry[p] = hb[2 * i + 1];
        if (primitive) memcpy(&primitive[p], &ha[4 * i + 3], 4);
        if (t) t[p] = hit ? ha[4 * i] : INFINITY;
        if (u) u[p] = hit ? ha[4 * i + 1] : 0.0f;
        if (v) v[p] = hit ? ha[4 * i + 2] : 0.0f;
    }
    HL_CATCH
}

hl_status hl_render_output_buffer(hl_context ctx, const hl_push_constants* pc, int output_buffer, float* rgba32f_host)
{
    HL_TRY(ctx)
    if (!pc || !rgba32f_host) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_render_output_buffer: null argument");
    if (output_buffer < HL_OUTPUT_BUFFER_ALBEDO || output_buffer > HL_OUTPUT_BUFFER_EMISSIVE) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_render_output_buffer: unknown output buffer");
    if (!c_->scene_ready) HL_FAIL(HL_ERR_STATE, "hl_render_output_buffer: scene tables not set");
    if (pc->launch_id_size[2] != c_->W || pc->launch_id_size[3] != c_->H) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_render_output_buffer: extent mismatch");
    wavefront_join(c_);
    const size_t n = (size_t)c_->W * c_->H;
    DevBuf       out;
    out.alloc(n * 16);
    wavefront_output_buffer(c_, *pc, output_buffer, out.as<float4>());
    HL_CUDA(cudaMemcpyAsync(rgba32f_host, out.p, n * 16, cudaMemcpyDeviceToHost, c_->stream));
    HL_CUDA(cudaStreamSynchronize(c_->stream));
    HL_CATCH
}

hl_status hl_gather_debug_rays(hl_context ctx, const hl_push_constants* pc, uint32_t num_debug_rays, hl_debug_ray_vertex* vertices_host, uint32_t max_vertices, uint32_t* vertex_count)
{
    HL_TRY(ctx)
    if (!pc || !vertex_count || (!vertices_host && max_vertices)) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_gather_debug_rays: null argument");
    if (!c_->scene_ready) HL_FAIL(HL_ERR_STATE, "hl_gather_debug_rays: scene tables not set");
    *vertex_count = 0;
    if (num_debug_rays == 0) return HL_OK;
    DevBuf verts, count;
    verts.alloc((size_t)max_vertices * sizeof(hl_debug_ray_vertex));
    count.alloc(4);
    HL_CUDA(cudaMemsetAsync(count.p, 0, 4, c_->stream));
    wavefront_debug_rays(c_, *pc, num_debug_rays, verts.as<float4>(), max_vertices, count.as<uint32_t>());
    uint32_t n = 0;
    HL_CUDA(cudaMemcpyAsync(&n, count.p, 4, cudaMemcpyDeviceToHost, c_->stream));
    HL_CUDA(cudaStreamSynchronize(c_->stream));
    *vertex_count = n;
    const uint32_t m = n < max_vertices ? n : max_vertices;
    if (m) HL_CUDA(cudaMemcpy(vertices_host, verts.p, (size_t)m * sizeof(hl_debug_ray_vertex), cudaMemcpyDeviceToHost));
    HL_CATCH
}

hl_status hl_trace_rays(hl_context ctx, const float* rays, uint32_t n_rays, uint32_t flags, void* hits)
{
    HL_TRY(ctx)
    if ((!rays || !hits) && n_rays) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_trace_rays: null argument");
    if (!c_->scene_ready) HL_FAIL(HL_ERR_STATE, "hl_trace_rays: scene tables not set");
    DevBuf dr, dh;
    dr.upload(rays, (size_t)n_rays * 32, c_->stream);
    dh.alloc((size_t)n_rays * 24);
    wavefront_trace_rays(c_, dr.as<float>(), n_rays, flags, dh.p);
    if (n_rays) HL_CUDA(cudaMemcpyAsync(hits, dh.p, (size_t)n_rays * 24, cudaMemcpyDeviceToHost, c_->stream));
    HL_CUDA(cudaStreamSynchronize(c_->stream));
    HL_CATCH
}

hl_status hl_tonemap(hl_context ctx, float exposure, int op, float sample_scale, uint8_t* rgba8_host)
{
    HL_TRY(ctx)
    film_tonemap(c_, exposure, op, sample_scale);
    if (rgba8_host)
    {
        HL_CUDA(cudaMemcpyAsync(rgba8_host, c_->rgba8.p, (size_t)c_->W * c_->H * 4, cudaMemcpyDeviceToHost, c_->stream));
        HL_CUDA(cudaStreamSynchronize(c_->stream));
    }
    HL_CATCH
}

hl_status hl_read_accum(hl_context ctx, float* out)
{
    HL_TRY(ctx)
    if (!out) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_read_accum: null pointer");
    HL_CUDA(cudaMemcpyAsync(out, c_->accum.p, (size_t)c_->W * c_->H * 16, cudaMemcpyDeviceToHost, c_->stream));
    HL_CUDA(cudaStreamSynchronize(c_->stream));
    HL_CATCH
}

hl_status hl_write_accum(hl_context ctx, const float* in)
{
    HL_TRY(ctx)
    if (!in) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_write_accum: null pointer");
    HL_CUDA(cudaMemcpyAsync(c_->accum.p, in, (size_t)c_->W * c_->H * 16, cudaMemcpyHostToDevice, c_->stream));
    HL_CUDA(cudaStreamSynchronize(c_->stream));
    HL_CATCH
}

hl_status hl_accum_device_ptr(hl_context ctx, void** out_ptr)
{
    HL_TRY(ctx)
    if (!out_ptr) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_accum_device_ptr: null pointer");
    *out_ptr = c_->accum.p;
    HL_CATCH
}

hl_status hl_synchronize(hl_context ctx)
{
    HL_TRY(ctx)
    HL_CUDA(cudaStreamSynchronize(c_->stream));
    HL_CUDA(cudaGetLastError());
    HL_CATCH
}

hl_status hl_get_counters(hl_context ctx, hl_counters* out)
{
    HL_TRY(ctx)
    if (!out) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_get_counters: null pointer");
    uint64_t totals[HL_MAX_WAVE_SLOTS][2];
    for (int k = 0; k < c_->n_slots; k++) HL_CUDA(cudaMemcpyAsync(totals[k], (char*)c_->slot[k].counters.p + CTR_TOTALS_OFFSET, 16, cudaMemcpyDeviceToHost, c_->stream));
    HL_CUDA(cudaStreamSynchronize(c_->stream));
    *out                = c_->last;
    out->extension_rays = out->shadow_rays = 0, out->frames = c_->frames;
    for (int k = 0; k < c_->n_slots; k++) out->extension_rays += totals[k][0], out->shadow_rays += totals[k][1];
    const uint64_t lost = trav_overflow_count(c_, false);
    if (lost) HL_FAIL(HL_ERR_LIMIT, "hl_get_counters: " + std::to_string(lost) + " traversal-stack overflows since the last hl_reset_counters (BVH deeper than the 64-entry stack): results are incomplete");
    HL_CATCH
}

hl_status hl_get_bounce_profile(hl_context ctx, hl_bounce_profile* out, uint32_t capacity, uint32_t* n_bounces, uint64_t* tail_ext, uint64_t* tail_sh)
{
    HL_TRY(ctx)
    if (!n_bounces || (!out && capacity)) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_get_bounce_profile: null pointer");
    *n_bounces = c_->bounce_prof_n;
    for (uint32_t b = 0; b < c_->bounce_prof_n && b < capacity; b++) out[b] = c_->bounce_prof[b];
    if (tail_ext) *tail_ext = c_->prof_tail_ext;
    if (tail_sh) *tail_sh = c_->prof_tail_sh;
    HL_CATCH
}

hl_status hl_reset_counters(hl_context ctx)
{
    HL_TRY(ctx)
    for (int k = 0; k < c_->n_slots; k++) HL_CUDA(cudaMemsetAsync((char*)c_->slot[k].counters.p + CTR_TOTALS_OFFSET, 0, 16, c_->stream));
    trav_overflow_count(c_, true);
    c_->frames = 0;
    HL_CATCH
}

hl_status hl_set_profiling(hl_context ctx, int enabled)
{
    HL_TRY(ctx)
    c_->profiling = enabled != 0;
    HL_CATCH
}

hl_status hl_event_record(hl_context ctx, int slot)
{
    HL_TRY(ctx)
    if (slot < 0 || slot >= 8) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_event_record: slot out of range");
    if (!c_->user_ev_ready)
    {
        for (auto& e : c_->user_ev) HL_CUDA(cudaEventCreate(&e));
        c_->user_ev_ready = true;
    }
    HL_CUDA(cudaEventRecord(c_->user_ev[slot], c_->stream));
    HL_CATCH
}

hl_status hl_event_elapsed_ms(hl_context ctx, int a, int b, float* out_ms)
{
    HL_TRY(ctx)
    if (a < 0 || a >= 8 || b < 0 || b >= 8 || !out_ms || !c_->user_ev_ready) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_event_elapsed_ms: invalid argument");
    HL_CUDA(cudaEventSynchronize(c_->user_ev[b]));
    HL_CUDA(cudaEventElapsedTime(out_ms, c_->user_ev[a], c_->user_ev[b]));
    HL_CATCH
}

hl_status hl_set_option(hl_context ctx, int option, int64_t value)
{
    HL_TRY(ctx)
    if (option == HL_OPT_TAIL_THRESHOLD)
        c_->tail_threshold = (uint32_t)std::max<int64_t>(0, std::min<int64_t>(value, 0x7FFFFFFF));
    else if (option == HL_OPT_TAIL_START)
        c_->tail_start = (uint32_t)std::max<int64_t>(1, value);
    else if (option == HL_OPT_PIPELINE)
        c_->pipeline = value != 0;
    else if (option == HL_OPT_FRAMES_IN_FLIGHT)
    {
        if (value < 1 || value > HL_MAX_WAVE_SLOTS) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_set_option: frames in flight must be 1..8");
        // (HL_TRY joined the frames in flight) ray totals of slots that go out of use move into slot 0
        wavefront_set_slots(c_, (int)value);
    }
    else if (option == HL_OPT_CUDA_GRAPH)
        c_->use_graphs = value != 0;
    else if (option == HL_OPT_SAH_CLUSTER)
        c_->sah_cluster = (uint32_t)std::max<int64_t>(0, std::min<int64_t>(value, 1 << 20));
    else
        HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_set_option: unknown option");
    HL_CATCH
}

hl_status hl_kernel_launches(hl_context ctx, uint64_t* out)
{
    HL_TRY(ctx)
    if (!out) HL_FAIL(HL_ERR_INVALID_ARGUMENT, "hl_kernel_launches: null pointer");
    *out = c_->launches;
    HL_CATCH
}

} // extern "C"
