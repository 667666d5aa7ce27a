// tests/emul/emul.cpp — kernel-logic emulator (DEBUG HARNESS, tests only).
//
// The build container has no GPU.  The per-element device logic of every stage lives in the HL_HD headers
// under helios_b200/csrc/ (hl_build.h, hl_bvh.h, hl_shade.h, hl_camera.h, hl_film.h); the CUDA kernels are thin
// loops over thread ids around those functions.  This file compiles the SAME headers with g++ and drives them
// with plain loops in wavefront order (generate -> extend -> shade -> connect -> resolve), so that the builder,
// the traversal and the shading can be checked against the oracle before GPU time is spent.  It is never
// loaded by the helios_b200 package and is not a fallback: the product path is the CUDA library only.
#define HL_TRAVERSAL_STATS 1
#include "../../helios_b200/csrc/hl_build.h"
#include "../../helios_b200/csrc/hl_bvh.h"
#include "../../helios_b200/csrc/hl_camera.h"
#include "../../helios_b200/csrc/hl_debug.h"
#include "../../helios_b200/csrc/hl_film.h"
#include "../../helios_b200/csrc/hl_shade.h"
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <numeric>
#include <vector>

using namespace hl;
#define EM_API extern "C" __attribute__((visibility("default")))

struct WideBVH
{
    std::vector<WideNode> nodes;
    std::vector<LeafTri>  tris;
    std::vector<uint32_t> inst_leaf;
    Box                   root;
    uint32_t              n_binary = 0;
    float                 sah      = 0.0f; // C(root, 1) / A(root) of the collapse DP
};

// ---- binned-SAH re-split of the LBVH's upper levels: sequential driver of the per-element phase functions of
// hl_build.h (the CUDA builder runs the same functions from one persistent kernel, hl_builder.cu k_top_build).
// cluster size 0 = off.
static int g_sah_top_cluster = HL_DEFAULT_SAH_CLUSTER;
EM_API void em_set_sah_top(int cluster_prims) { g_sah_top_cluster = cluster_prims; }
static uint32_t g_top_levels = 0;
EM_API uint32_t em_last_top_levels() { return g_top_levels; }
static void refine_top(BinaryTree& t, uint32_t C)
{
    g_top_levels = 0;
    if (t.n < 2 || t.n <= C) return;
    std::vector<uint32_t> clusters, upper;
    for (uint32_t m = 0; m < 2 * t.n - 1; m++)
        if (top_is_cluster_root(t, m, C)) clusters.push_back(m);
    for (uint32_t m = 0; m + 1 < t.n; m++)
        if (top_is_upper_node(t, m, C)) upper.push_back(m);
    const uint32_t K = (uint32_t)clusters.size();
    if (K < 2 || upper.size() != K - 1 || upper[0] != 0u) abort();
    std::vector<TopCluster> crec(K);
    std::vector<uint32_t> cnode(K), level_count(HL_TOP_MAX_LEVELS + 1, 0), bins_used(HL_TOP_MAX_LEVELS + 1, 0);
    std::vector<TopNode>  lv[2] = { std::vector<TopNode>(K + 2), std::vector<TopNode>(K + 2) };
    const uint32_t        bins_cap = K / (HL_TOP_SMALL + 1) + 2;
    std::vector<uint32_t> list((size_t)(K / 2 + 1) * HL_TOP_SMALL);
    std::vector<TopSmall> small(K / 2 + 1);
    uint32_t              small_count = 0;
    std::vector<TopBin>   bins[2]  = { std::vector<TopBin>((size_t)bins_cap * 3 * HL_TOP_BINS), std::vector<TopBin>((size_t)bins_cap * 3 * HL_TOP_BINS) };
    uint32_t free_next = 0, kdev = K;
    TopBuild tb;
    tb.cluster_prims = C, tb.k_cap = K, tb.bins_cap = bins_cap, tb.n_clusters = &kdev, tb.cluster = clusters.data(), tb.free_nodes = upper.data();
    tb.cnode = cnode.data(), tb.crec = crec.data(), tb.level[0] = lv[0].data(), tb.level[1] = lv[1].data(), tb.bins[0] = bins[0].data(), tb.bins[1] = bins[1].data();
    tb.list = list.data(), tb.small = small.data(), tb.small_count = &small_count;
    tb.level_count = level_count.data(), tb.bins_used = bins_used.data(), tb.free_next = &free_next;
    top_begin(tb, K);
    for (uint32_t i = 0; i < K; i++) top_seed_cluster(t, tb, i);
    for (uint32_t L = 0;; L++)
    {
        if (L + 1 >= HL_TOP_MAX_LEVELS) abort();
        for (uint32_t i = 0; i < K; i++) top_bin_cluster(t, tb, L, i);
        for (uint32_t j = 0; j < level_count[L]; j++) top_choose_node(tb, L, j);
        for (uint32_t j = 0; j < level_count[L]; j++) top_commit_node(t, tb, L, j);
        for (uint32_t b = 0, nb = top_bins_to_clear(tb, L + 1); b < nb; b++) top_clear_bin(tb.bins[(L + 1) & 1u] + b);
        g_top_levels = L + 1;
        if (level_count[L + 1] == 0) break;
        for (uint32_t i = 0; i < K; i++) top_assign_cluster(t, tb, L, i);
    }
    for (uint32_t r = 0; r < small_count; r++)
    {
        const uint32_t ids = top_small_node_ids(tb, r);
        top_small_node(t, tb, r, free_next);
        free_next += ids;
    }
    if (free_next != K - 1) abort();
    for (uint32_t i = 0; i < K; i++)
        if (cnode[i] != HL_TOP_DONE) abort();
    for (uint32_t i = 0; i < K; i++) top_refit_from_cluster(t, clusters[i], [] {});
}

template <class MakeWriter>
static void build_wide(const std::vector<Box>& prim, WideBVH& out, MakeWriter make_writer, bool tri_leaves)
{
    const uint32_t n = (uint32_t)prim.size();
    out.nodes.clear();
    if (n == 0) return;
    Box scene = prim[0];
    for (auto& b : prim) scene = box_union(scene, b);
    std::vector<uint64_t> keys(n);
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; i++) keys[i] = morton_key(prim[i], scene);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    std::vector<uint64_t> skeys(n);
    for (uint32_t i = 0; i < n; i++) skeys[i] = keys[order[i]];
    std::vector<uint32_t> left(n), right(n), first(n), last(n), parent(2 * n, 0xFFFFFFFFu), visits(n, 0);
    std::vector<Box>      box(2 * n);
    std::vector<float>    cost((size_t)2 * n * 7);
    std::vector<uint32_t> dec(n, 0);
    BinaryTree            t;
    t.cost = cost.data(), t.dec = dec.data(), t.c_prim = tri_leaves ? HL_SAH_C_PRIM_TRIANGLE : HL_SAH_C_PRIM_INSTANCE;
    t.n = n, t.left = left.data(), t.right = right.data(), t.first = first.data(), t.last = last.data();
    t.parent = parent.data(), t.box = box.data(), t.visits = visits.data();
    for (int i = 0; i + 1 < (int)n; i++) radix_tree_node(skeys.data(), t, i);
    for (uint32_t j = 0; j < n; j++) box[(n - 1) + j] = prim[order[j]];
    for (uint32_t j = 0; j < n; j++) fit_from_leaf(t, j, [] {});
    if (g_sah_top_cluster > 0) refine_top(t, tri_leaves ? (uint32_t)g_sah_top_cluster : 1u);
    out.root     = box[0];
    out.n_binary = 2 * n - 1;
    out.sah      = cost[0] / box_half_area(box[0]);
    out.nodes.resize(n);
    if (tri_leaves)
        out.tris.resize(n);
    else
        out.inst_leaf.resize(n);
    uint32_t node_counter = 1, leaf_counter = 0;
    WideOut  wo { out.nodes.data(), &node_counter, &leaf_counter };
    std::vector<CollapseTask> cur(1), next(n);
    cur[0].wide = 0, cur[0].bnode = (n == 1) ? 0u : 0u;
    auto writer = make_writer(order.data());
    while (!cur.empty())
    {
        uint32_t nc = 0;
        for (auto& task : cur) collapse_one(t, task, wo, next.data(), &nc, writer);
        cur.assign(next.begin(), next.begin() + nc);
    }
    out.nodes.resize(node_counter);
}

struct EmMesh
{
    std::vector<hl_vertex>  v;
    std::vector<uint32_t>   idx;
    std::vector<hl_submesh> subs;
    std::vector<uint32_t>   tri_start;
    std::vector<AlphaTri>   alpha; // any-hit records beside the leaves (meshes with a non-opaque submesh), as hl_builder.cu allocates them
    WideBVH                 bvh;
};
struct EmTexture
{
    int                  format;
    uint32_t             w, h;
    std::vector<uint8_t> data;
};
struct EmScene
{
    std::vector<EmMesh*>     meshes;
    std::vector<EmTexture>   textures;
    std::vector<f4>          env;
    uint32_t                 env_size = 0;
    std::vector<hl_material> materials;
    std::vector<hl_instance> instances;
    std::vector<hl_light>    lights;
    std::vector<uint32_t>    submesh_info, submesh_offset;
    std::vector<float>       inst_inv;
    std::vector<uint32_t>    inst_identity;
    std::vector<InstAlpha>   inst_alpha;
    std::vector<GeomAlpha>   geom_alpha;
    std::vector<MeshView>    mesh_views;
    std::vector<TexView>     tex_views;
    std::vector<float>       lut8;
    WideBVH                  tlas;
    SceneView                view;
    ~EmScene()
    {
        for (auto m : meshes) delete m;
    }
};

// identical formula to the product library (hl_api.cu): world -> object 3x4 in double, rounded once
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

EM_API EmScene* em_scene_new()
{
    EmScene* s = new EmScene();
    s->lut8.resize(768);
    for (int i = 0; i < 256; i++)
    {
        double c        = i / 255.0;
        s->lut8[i]      = (float)c;
        s->lut8[256 + i] = (float)(c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4));
        int sn          = (int8_t)(uint8_t)i;
        s->lut8[512 + i] = (float)std::max(-1.0, sn / 127.0);
    }
    return s;
}
EM_API void em_scene_free(EmScene* s) { delete s; }
EM_API int  em_scene_add_mesh(EmScene* s, const hl_vertex* v, uint32_t nv, const uint32_t* idx, uint32_t ni, const hl_submesh* subs, uint32_t ns)
{
    EmMesh* m = new EmMesh();
    m->v.assign(v, v + nv);
    m->idx.assign(idx, idx + ni);
    m->subs.assign(subs, subs + ns);
    m->tri_start.resize(ns + 1);
    m->tri_start[0] = 0;
    for (uint32_t g = 0; g < ns; g++) m->tri_start[g + 1] = m->tri_start[g] + subs[g].index_count / 3;
    const uint32_t   ntri = m->tri_start[ns];
    std::vector<Box> prim(ntri);
    for (uint32_t f = 0; f < ntri; f++) prim[f] = triangle_box(m->v.data(), m->idx.data(), m->subs.data(), m->tri_start.data(), ns, f);
    bool any_hit = false;
    for (const hl_submesh& sm : m->subs) any_hit = any_hit || !sm.opaque;
    if (any_hit) m->alpha.resize(ntri);
    build_wide(prim, m->bvh, [&](const uint32_t* order) {
        TriLeafWriter w;
        w.vertices = m->v.data(), w.indices = m->idx.data(), w.submeshes = m->subs.data(), w.tri_start = m->tri_start.data();
        w.n_geom = ns, w.sorted_prim = order, w.tris = m->bvh.tris.data(), w.alpha = any_hit && ntri ? m->alpha.data() : nullptr;
        return w;
    }, true);
    s->meshes.push_back(m);
    return (int)s->meshes.size() - 1;
}
EM_API int em_scene_add_texture(EmScene* s, int format, uint32_t w, uint32_t h, const void* data)
{
    EmTexture t;
    t.format = format, t.w = w, t.h = h;
    size_t n = (size_t)w * h * (format == 3 ? 16 : 4);
    t.data.assign((const uint8_t*)data, (const uint8_t*)data + n);
    s->textures.push_back(std::move(t));
    return (int)s->textures.size() - 1;
}
EM_API void em_scene_set_envmap(EmScene* s, uint32_t size, const float* faces)
{
    s->env_size = size;
    const uint32_t P = size + 2; // bordered faces (hl_tex.h cube_pad_texel), as env_pad() builds them on the device
    s->env.resize((size_t)6 * P * P);
    for (int face = 0; face < 6; face++)
        for (uint32_t y = 0; y < P; y++)
            for (uint32_t x = 0; x < P; x++) s->env[((size_t)face * P + y) * P + x] = cube_pad_texel((const f4*)faces, (int)size, face, (int)x - 1, (int)y - 1);
}
EM_API void em_sky_bake(const float* cf40, const float* sun, uint32_t size, float* out)
{
    for (int face = 0; face < 6; face++)
        for (uint32_t j = 0; j < size; j++)
            for (uint32_t i = 0; i < size; i++)
            {
                f3     c = hosek_wilkie_radiance(cf40, cube_texel_direction(face, i, j, size), mk3(sun));
                float* o = out + (((size_t)face * size + j) * size + i) * 4;
                o[0] = c.x, o[1] = c.y, o[2] = c.z, o[3] = 1.0f;
            }
}
EM_API void em_scene_set_tables(EmScene* s, const hl_material* mats, uint32_t nm, const hl_instance* inst, const uint32_t* const* submesh_info, uint32_t ni, const hl_light* lights, uint32_t nl)
{
    s->materials.assign(mats, mats + nm);
    s->instances.assign(inst, inst + ni);
    s->lights.assign(lights, lights + nl);
    s->submesh_info.clear();
    s->submesh_offset.resize(ni);
    s->inst_inv.resize((size_t)ni * 12);
    std::vector<Box> iboxes(ni);
    bool             identity = ni == 1;
    for (uint32_t i = 0; i < ni; i++)
    {
        const EmMesh& m      = *s->meshes[inst[i].mesh_index];
        s->submesh_offset[i] = (uint32_t)(s->submesh_info.size() / 2);
        s->submesh_info.insert(s->submesh_info.end(), submesh_info[i], submesh_info[i] + 2 * m.subs.size());
        affine_inverse(inst[i].model_matrix, &s->inst_inv[(size_t)i * 12]);
        static const float I[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
        if (memcmp(inst[i].model_matrix, I, 64) != 0) identity = false;
        double       lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
        const float* M     = inst[i].model_matrix;
        const Box&   rb    = m.bvh.root;
        for (int c = 0; c < 8; c++)
        {
            double x = (c & 1) ? rb.hi[0] : rb.lo[0], y = (c & 2) ? rb.hi[1] : rb.lo[1], z = (c & 4) ? rb.hi[2] : rb.lo[2];
            for (int a = 0; a < 3; a++)
            {
                double w = (double)M[a] * x + (double)M[4 + a] * y + (double)M[8 + a] * z + (double)M[12 + a];
                lo[a] = std::min(lo[a], w), hi[a] = std::max(hi[a], w);
            }
        }
        for (int a = 0; a < 3; a++)
        {
            double pad      = 1e-5 * (std::fabs(lo[a]) + std::fabs(hi[a]) + (hi[a] - lo[a]));
            iboxes[i].lo[a] = (float)(lo[a] - pad), iboxes[i].hi[a] = (float)(hi[a] + pad);
        }
    }
    build_wide(iboxes, s->tlas, [&](const uint32_t* order) {
        InstLeafWriter w;
        w.sorted_prim = order, w.leaf = s->tlas.inst_leaf.data();
        return w;
    }, false);
    s->mesh_views.resize(s->meshes.size());
    for (size_t k = 0; k < s->meshes.size(); k++)
    {
        EmMesh&   m = *s->meshes[k];
        MeshView& v = s->mesh_views[k];
        v.vertices = m.v.data(), v.indices = m.idx.data(), v.nodes = m.bvh.nodes.data(), v.tris = m.bvh.tris.data();
        v.n_tris = (uint32_t)m.bvh.tris.size(), v.n_submeshes = (uint32_t)m.subs.size();
    }
    s->tex_views.resize(s->textures.size());
    for (size_t k = 0; k < s->textures.size(); k++)
    {
        s->tex_views[k].texels = s->textures[k].data.data();
        s->tex_views[k].w = s->textures[k].w, s->tex_views[k].h = s->textures[k].h, s->tex_views[k].format = s->textures[k].format;
    }
    SceneView& v = s->view;
    v.materials = s->materials.data(), v.instances = s->instances.data(), v.inst_inv = s->inst_inv.data();
    v.submesh_info = s->submesh_info.data(), v.submesh_offset = s->submesh_offset.data(), v.lights = s->lights.data();
    v.meshes = s->mesh_views.data(), v.textures = s->tex_views.data(), v.lut8 = s->lut8.data();
    v.env.faces = s->env.data(), v.env.size = s->env_size;
    v.tlas_nodes = s->tlas.nodes.data(), v.tlas_leaf = s->tlas.inst_leaf.data();
    v.n_instances = ni, v.n_lights = nl, v.single_identity = identity ? 1u : 0u;
    // any-hit records, as hl_scene_set_tables builds them
    s->inst_alpha.assign(ni, InstAlpha());
    s->geom_alpha.clear();
    for (uint32_t i = 0; i < ni; i++)
    {
        const EmMesh& m = *s->meshes[inst[i].mesh_index];
        for (size_t g = 0; g < m.subs.size(); g++)
        {
            const hl_material& mat = mats[submesh_info[i][2 * g + 1]];
            GeomAlpha          ga;
            ga.texture = mat.texture_indices0[0], ga.alpha = mat.albedo[3];
            s->geom_alpha.push_back(ga);
        }
        s->inst_alpha[i].alpha = m.alpha.empty() ? nullptr : m.alpha.data(), s->inst_alpha[i].tris = m.bvh.tris.data(), s->inst_alpha[i].info_base = s->submesh_offset[i];
    }
    v.inst_alpha = s->inst_alpha.data(), v.geom_alpha = s->geom_alpha.data();
    s->inst_identity.assign((ni + 31) / 32 + 1, 0u);
    for (uint32_t i = 0; i < ni; i++)
    {
        static const float I[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
        if (memcmp(inst[i].model_matrix, I, 64) == 0) s->inst_identity[i >> 5] |= 1u << (i & 31);
    }
    v.inst_identity = s->inst_identity.data();
}
EM_API void em_scene_force_two_level(EmScene* s) { s->view.single_identity = 0; }
// hl_scene_update_instances on the host: new transforms, instance tree REFITTED with the product's per-node functions
// (hl_build.h refit_node_box / refit_requantize; the CUDA kernel k_tlas_refit runs the same sweeps in one block)
EM_API int em_scene_update_instances(EmScene* s, const hl_instance* inst, uint32_t ni)
{
    if (ni != s->instances.size() || s->tlas.nodes.empty()) return 1;
    std::vector<Box> iboxes(ni);
    bool             identity = ni == 1;
    for (uint32_t i = 0; i < ni; i++)
    {
        memcpy(s->instances[i].model_matrix, inst[i].model_matrix, 64), memcpy(s->instances[i].normal_matrix, inst[i].normal_matrix, 64);
        affine_inverse(inst[i].model_matrix, &s->inst_inv[(size_t)i * 12]);
        static const float I[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
        if (memcmp(inst[i].model_matrix, I, 64) != 0) identity = false;
        const EmMesh& m     = *s->meshes[s->instances[i].mesh_index];
        double        lo[3] = { 1e300, 1e300, 1e300 }, hi[3] = { -1e300, -1e300, -1e300 };
        const float*  M     = inst[i].model_matrix;
        const Box&    rb    = m.bvh.root;
        for (int c = 0; c < 8; c++)
        {
            double x = (c & 1) ? rb.hi[0] : rb.lo[0], y = (c & 2) ? rb.hi[1] : rb.lo[1], z = (c & 4) ? rb.hi[2] : rb.lo[2];
            for (int a = 0; a < 3; a++)
            {
                double w = (double)M[a] * x + (double)M[4 + a] * y + (double)M[8 + a] * z + (double)M[12 + a];
                lo[a] = std::min(lo[a], w), hi[a] = std::max(hi[a], w);
            }
        }
        for (int a = 0; a < 3; a++)
        {
            double pad      = 1e-5 * (std::fabs(lo[a]) + std::fabs(hi[a]) + (hi[a] - lo[a]));
            iboxes[i].lo[a] = (float)(lo[a] - pad), iboxes[i].hi[a] = (float)(hi[a] + pad);
        }
    }
    std::vector<WideNode>& nodes = s->tlas.nodes;
    std::vector<Box>       nb(nodes.size(), box_empty());
    int                    sweeps = 0;
    for (; sweeps < 64; sweeps++)
    {
        std::vector<Box> next(nodes.size());
        for (size_t i = 0; i < nodes.size(); i++) next[i] = refit_node_box(nodes[i], s->tlas.inst_leaf.data(), iboxes.data(), nb.data());
        if (memcmp(next.data(), nb.data(), sizeof(Box) * nb.size()) == 0) break;
        nb = next;
    }
    for (size_t i = 0; i < nodes.size(); i++) refit_requantize(nodes[i], nb[i], s->tlas.inst_leaf.data(), iboxes.data(), nb.data());
    s->view.single_identity = identity ? 1u : 0u;
    s->inst_identity.assign((ni + 31) / 32 + 1, 0u);
    for (uint32_t i = 0; i < ni; i++)
    {
        static const float I[16] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1 };
        if (memcmp(inst[i].model_matrix, I, 64) == 0) s->inst_identity[i >> 5] |= 1u << (i & 31);
    }
    s->view.inst_identity = s->inst_identity.data();
    return 0;
}

static void trace_one(const SceneView& s, f3 o, float tmin, f3 d, float tmax, uint32_t flags, Hit& h)
{
    u2        fast[HL_STACK_FAST], spill[HL_STACK_SPILL];
    TravStack st;
    st.init(spill, fast);
    trace_ray(s, true, o, tmin, d, tmax, flags, h, st);
}

EM_API void em_trace_primary_ids(const EmScene* s, const hl_push_constants* pc, uint32_t* inst, uint32_t* geom, uint32_t* prim, float* t, float* u, float* v)
{
    const uint32_t W = pc->launch_id_size[2], H = pc->launch_id_size[3];
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t y = 0; y < (int64_t)H; y++)
        for (uint32_t x = 0; x < W; x++)
        {
            Rng rng = rng_seed(x, (uint32_t)y, pc->num_frames);
            f3  o, d;
            primary_ray(*pc, x, (uint32_t)y, rng, o, d);
            Hit h;
            trace_one(s->view, o, 0.001f, d, 10000.0f, 0, h);
            size_t i = (size_t)y * W + x;
            const bool hit = h.instance != HL_MISS;
            inst[i] = h.instance, geom[i] = h.geometry, prim[i] = h.primitive;
            t[i] = hit ? h.t : hl_inf(), u[i] = hit ? h.u : 0.0f, v[i] = hit ? h.v : 0.0f;
        }
}
struct EmDebugOut
{
    float*   verts;
    uint32_t count = 0, capacity;
    uint32_t alloc2()
    {
        const uint32_t k = count;
        count += 2;
        return k;
    }
    void put(uint32_t k, f3 p, f3 c)
    {
        if (k >= capacity) return;
        float* q = verts + (size_t)k * 8;
        q[0] = p.x, q[1] = p.y, q[2] = p.z, q[3] = 1.0f, q[4] = c.x, q[5] = c.y, q[6] = c.z, q[7] = 1.0f;
    }
};
EM_API uint32_t em_gather_debug_rays(const EmScene* s, const hl_push_constants* pc, uint32_t n, float* out, uint32_t max_vertices)
{
    EmDebugOut o;
    o.verts = out, o.capacity = max_vertices;
    for (uint32_t i = 0; i < n; i++)
    {
        u2        fast[HL_STACK_FAST], spill[HL_STACK_SPILL];
        TravStack st;
        st.init(spill, fast);
        debug_ray_path(s->view, *pc, i, true, st, o);
    }
    return o.count;
}
EM_API void em_output_buffer(const EmScene* s, const hl_push_constants* pc, int which, float* out)
{
    const uint32_t W = pc->launch_id_size[2], H = pc->launch_id_size[3];
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t y = 0; y < (int64_t)H; y++)
        for (uint32_t x = 0; x < W; x++)
        {
            Rng rng = rng_seed(x, (uint32_t)y, pc->num_frames);
            f3  o, d;
            primary_ray(*pc, x, (uint32_t)y, rng, o, d);
            Hit h;
            trace_one(s->view, o, 0.001f, d, 10000.0f, 0, h);
            const f4 c = output_buffer_value(s->view, h, which);
            float*   q = out + ((size_t)y * W + x) * 4;
            q[0] = c.x, q[1] = c.y, q[2] = c.z, q[3] = c.w;
        }
}
EM_API void em_trace_rays(const EmScene* s, const float* rays, uint32_t n, uint32_t flags, void* hits)
{
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < (int64_t)n; i++)
    {
        const float* r = rays + i * 8;
        Hit          h;
        trace_one(s->view, mk3(r[0], r[1], r[2]), r[3], mk3(r[4], r[5], r[6]), r[7], flags, h);
        float*    o  = (float*)hits + i * 6;
        uint32_t* ou = (uint32_t*)o;
        const bool hit = h.instance != HL_MISS;
        o[0] = hit ? h.t : hl_inf(), o[1] = hit ? h.u : 0.0f, o[2] = hit ? h.v : 0.0f;
        ou[3] = h.instance, ou[4] = h.geometry, ou[5] = h.primitive;
    }
}

// optional per-ray traversal log for em_render_frame: [16 bounces][cap] = nodes | leaves << 16
static uint32_t* g_node_log     = nullptr;
static size_t    g_node_log_cap = 0;
EM_API void      em_set_force_postpone(int on) { emul_force_postpone() = on != 0; }
EM_API void      em_set_node_log(uint32_t* buf, size_t cap) { g_node_log = buf, g_node_log_cap = cap; }

// one launch in wavefront order; accum is updated in place
EM_API void em_render_frame(const EmScene* sc, const hl_push_constants* pcp, uint32_t lw, uint32_t lh, float* accum, uint64_t* counters, int accum_mode)
{
    const hl_push_constants& pc = *pcp;
    const SceneView&         s  = sc->view;
    const uint32_t           W = pc.launch_id_size[2], H = pc.launch_id_size[3];
    if (lw == 0) lw = W;
    if (lh == 0) lh = H;
    const size_t n = (size_t)lw * lh;
    struct Path
    {
        f3       T, L;
        Rng      rng;
        uint32_t px, py;
        bool     valid;
    };
    struct ExtRay
    {
        f3       o, d;
        uint32_t path;
    };
    struct ShRay
    {
        f3       o, d, c;
        float    tmax;
        uint32_t path;
    };
    std::vector<Path>   paths(n);
    std::vector<ExtRay> q, qn;
    std::vector<Hit>    hits;
    q.reserve(n);
    for (size_t i = 0; i < n; i++) // generate
    {
        Path& p = paths[i];
        p.px = pc.launch_id_size[0] + (uint32_t)(i % lw), p.py = pc.launch_id_size[1] + (uint32_t)(i / lw);
        p.valid = p.px < W && p.py < H;
        if (!p.valid) continue;
        p.T = mk3(1.0f), p.L = mk3(0.0f);
        p.rng = rng_seed(p.px, p.py, pc.num_frames);
        ExtRay r;
        primary_ray(pc, p.px, p.py, p.rng, r.o, r.d);
        r.path = (uint32_t)i;
        q.push_back(r);
    }
    ShadeParams prm;
    prm.num_lights = pc.num_lights, prm.max_ray_bounces = pc.max_ray_bounces, prm.shadow_ray_bias = pc.shadow_ray_bias;
    for (uint32_t depth = 0; depth < std::max(1u, pc.max_ray_bounces) && !q.empty(); depth++)
    {
        const uint32_t ext_flags = depth == 0 ? 0u : HL_RAY_OPAQUE;
        const float    ext_tmin  = depth == 0 ? 0.001f : 0.0001f;
        hits.resize(q.size());
#pragma omp parallel for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)q.size(); i++) // extend
        {
            TraversalStats& ts = traversal_stats();
            ts.nodes = ts.leaves = 0;
            trace_one(s, q[i].o, ext_tmin, q[i].d, 10000.0f, ext_flags, hits[i]);
            if (g_node_log && depth < 16 && (size_t)i < g_node_log_cap) g_node_log[depth * g_node_log_cap + i] = (uint32_t)ts.nodes | ((uint32_t)ts.leaves << 16);
        }
        if (counters) counters[0] += q.size();
        std::vector<ShadeResult> res(q.size());
        std::vector<uint8_t>     is_hit(q.size());
#pragma omp parallel for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)q.size(); i++) // shade
        {
            Path& p = paths[q[i].path];
            is_hit[i] = hits[i].instance != HL_MISS;
            if (!is_hit[i])
            {
                p.L = p.L + shade_miss(s, depth, q[i].d, p.T);
                continue;
            }
            shade_hit(s, prm, depth, q[i].d, hits[i], p.T, p.rng, res[i]);
            p.L = p.L + res[i].emitted;
            if (res[i].continues) p.T = res[i].T;
        }
        std::vector<ShRay> sq;
        qn.clear();
        for (size_t i = 0; i < q.size(); i++) // compact
        {
            if (!is_hit[i]) continue;
            if (res[i].has_shadow)
            {
                ShRay r;
                r.o = res[i].shadow_o, r.d = res[i].shadow_d, r.c = res[i].direct, r.tmax = res[i].shadow_tmax, r.path = q[i].path;
                sq.push_back(r);
            }
            if (res[i].continues)
            {
                ExtRay r;
                r.o = res[i].next_o, r.d = res[i].next_d, r.path = q[i].path;
                qn.push_back(r);
            }
        }
        const uint32_t sh_flags = depth == 0 ? 0u : (HL_RAY_OPAQUE | HL_RAY_TERMINATE);
#pragma omp parallel for schedule(dynamic, 256)
        for (int64_t i = 0; i < (int64_t)sq.size(); i++) // connect
        {
            Hit h;
            trace_one(s, sq[i].o, 0.0001f, sq[i].d, sq[i].tmax, sh_flags, h);
            if (h.instance == HL_MISS) paths[sq[i].path].L = paths[sq[i].path].L + sq[i].c;
        }
        if (counters) counters[1] += sq.size();
        q.swap(qn);
    }
    for (size_t i = 0; i < n; i++) // resolve
    {
        const Path& p = paths[i];
        if (!p.valid) continue;
        float*   a    = accum + ((size_t)p.py * W + p.px) * 4;
        const f3 prev = mk3(a[0], a[1], a[2]);
        const f3 c    = accum_mode == HL_ACCUM_SUM ? accumulate_sum(p.L, prev) : accumulate_running_mean(p.L, prev, pc.num_frames);
        a[0] = c.x, a[1] = c.y, a[2] = c.z, a[3] = 1.0f;
    }
}
// per-query traversal statistics over a ray batch: out[0] = node visits, out[1] = leaf primitive tests,
// out[2] = max node visits of a single ray
EM_API void em_traversal_stats(const EmScene* s, const float* rays, uint32_t n, uint32_t flags, uint64_t* out)
{
    uint64_t nodes = 0, leaves = 0, worst = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : nodes, leaves) reduction(max : worst)
    for (int64_t i = 0; i < (int64_t)n; i++)
    {
        const float* r = rays + i * 8;
        Hit          h;
        TraversalStats& st = traversal_stats();
        st.nodes = st.leaves = 0;
        trace_one(s->view, mk3(r[0], r[1], r[2]), r[3], mk3(r[4], r[5], r[6]), r[7], flags, h);
        nodes += st.nodes, leaves += st.leaves;
        worst = std::max<uint64_t>(worst, st.nodes);
    }
    out[0] = nodes, out[1] = leaves, out[2] = worst;
}
// traversal-stack overflows counted so far (hl_bvh.h note_stack_overflow; the counter is not thread safe on the host: use with one thread)
EM_API uint64_t em_stack_overflows(int reset)
{
    const uint64_t n = emul_trav_overflow();
    if (reset) emul_trav_overflow() = 0;
    return n;
}
EM_API void em_tonemap(const float* accum, uint32_t W, uint32_t H, float exposure, int op, float scale, uint8_t* out)
{
    for (uint32_t r = 0; r < H; r++)
        for (uint32_t x = 0; x < W; x++)
        {
            const float* a = accum + ((size_t)(H - 1 - r) * W + x) * 4;
            ((uint32_t*)out)[(size_t)r * W + x] = tone_map_rgba8(mk3(a[0] * scale, a[1] * scale, a[2] * scale), exposure, op);
        }
}
EM_API float em_mesh_sah(const EmScene* s, int mesh) { return s->meshes[mesh]->bvh.sah; }
EM_API void em_mesh_stats(const EmScene* s, int mesh, uint32_t* out3)
{
    out3[0] = (uint32_t)s->meshes[mesh]->bvh.tris.size();
    out3[1] = (uint32_t)s->meshes[mesh]->bvh.nodes.size();
    out3[2] = s->meshes[mesh]->bvh.n_binary;
}
