// oracle/ref_ast/ref_ast_tool.cpp — TEST INFRASTRUCTURE (checker for SURVEY.md 8 row f1, never shipped or
// measured).  Built by oracle/oracle.py:build_ref_ast() together with the REFERENCE'S OWN AssetCore sources,
// compiled where they lie under /root/reference (loader.cpp, filesystem.cpp, mesh / material / scene exporters);
// outputs go to oracle/_ref/.  Two jobs:
//   ref_ast_tool dump (image|mesh|material|scene) <file> [strip-prefix]   load with the reference's loader, print
//                                                                          the canonical text of dump_format.h
//   ref_ast_tool write-fixtures <dir>     fill ast::Mesh / ast::Material / ast::Scene values and let the
//                                         reference's exporters write them: files "as the reference writes them"
// The image exporter needs nvtt / cmft and is not built; export_image / import_image are stubbed (the material
// exporter only calls them for textures that do not exist yet — the fixture textures are written beforehand by
// helios_b200/ast_io.py and checked by the reference's loader through `dump image`).
#include "dump_format.h"
#include <exporter/image_exporter.h>
#include <exporter/material_exporter.h>
#include <exporter/mesh_exporter.h>
#include <exporter/scene_exporter.h>
#include <loader/loader.h>
#include <cmath>

namespace ast
{
bool import_image(Image&, const std::string&, const PixelType&, int) { return false; }
bool export_image(Image&, const ImageExportOptions&) { return false; }
} // namespace ast

static void dump_node(FILE* f, const std::shared_ptr<ast::SceneNode>& n, int depth, const std::string& prefix)
{
    if (!n)
    {
        std::fprintf(f, "%*snode null\n", depth, "");
        return;
    }
    std::fprintf(f, "%*snode type=%d name=\"%s\" children=%zu", depth, "", (int)n->type, n->name.c_str(), n->children.size());
    if (auto t = std::dynamic_pointer_cast<ast::TransformNode>(n))
        dumpfmt::vec(f, "position", &t->position.x, 3), dumpfmt::vec(f, "rotation", &t->rotation.x, 3), dumpfmt::vec(f, "scale", &t->scale.x, 3);
    if (auto m = std::dynamic_pointer_cast<ast::MeshNode>(n)) std::fprintf(f, " mesh=\"%s\" material_override=\"%s\" casts_shadow=%d", m->mesh.c_str(), m->material_override.c_str(), (int)m->casts_shadow);
    if (auto l = std::dynamic_pointer_cast<ast::DirectionalLightNode>(n))
        dumpfmt::vec(f, "color", &l->color.x, 3), dumpfmt::vec(f, "intensity", &l->intensity, 1), dumpfmt::vec(f, "radius", &l->radius, 1), std::fprintf(f, " casts_shadows=%d", (int)l->casts_shadows);
    if (auto l = std::dynamic_pointer_cast<ast::SpotLightNode>(n))
        dumpfmt::vec(f, "color", &l->color.x, 3), dumpfmt::vec(f, "intensity", &l->intensity, 1), dumpfmt::vec(f, "radius", &l->radius, 1), std::fprintf(f, " casts_shadows=%d", (int)l->casts_shadows),
            dumpfmt::vec(f, "inner", &l->inner_cone_angle, 1), dumpfmt::vec(f, "outer", &l->outer_cone_angle, 1);
    if (auto l = std::dynamic_pointer_cast<ast::PointLightNode>(n))
        dumpfmt::vec(f, "color", &l->color.x, 3), dumpfmt::vec(f, "intensity", &l->intensity, 1), dumpfmt::vec(f, "radius", &l->radius, 1), std::fprintf(f, " casts_shadows=%d", (int)l->casts_shadows);
    if (auto c = std::dynamic_pointer_cast<ast::CameraNode>(n)) dumpfmt::vec(f, "near", &c->near_plane, 1), dumpfmt::vec(f, "far", &c->far_plane, 1), dumpfmt::vec(f, "fov", &c->fov, 1);
    if (auto i = std::dynamic_pointer_cast<ast::IBLNode>(n)) std::fprintf(f, " image=\"%s\"", i->image.c_str());
    std::fprintf(f, "\n");
    for (auto& c : n->children) dump_node(f, c, depth + 1, prefix);
}

static int dump(const std::string& kind, const std::string& path, const std::string& prefix)
{
    if (kind == "image")
    {
        ast::Image img;
        if (!ast::load_image(path, img)) return std::printf("load failed\n"), 0;
        dumpfmt::image_header(stdout, img.name.c_str(), img.components, img.mip_slices, img.array_slices, (int)img.type, (int)img.compression);
        for (int a = 0; a < img.array_slices; a++)
            for (int m = 0; m < img.mip_slices; m++) dumpfmt::image_level(stdout, a, m, img.data[a][m].width, img.data[a][m].height, img.data[a][m].data, img.data[a][m].size);
    }
    else if (kind == "mesh")
    {
        ast::Mesh mesh;
        if (!ast::load_mesh(path, mesh)) return std::printf("load failed\n"), 0;
        dumpfmt::mesh(stdout, mesh, prefix);
    }
    else if (kind == "material")
    {
        ast::Material m;
        if (!ast::load_material(path, m)) return std::printf("load failed\n"), 0;
        dumpfmt::material(stdout, m, prefix);
    }
    else if (kind == "scene")
    {
        ast::Scene s;
        if (!ast::load_scene(path, s)) return std::printf("load failed\n"), 0;
        std::printf("scene name=\"%s\"\n", s.name.c_str());
        dump_node(stdout, s.scene_graph, 0, prefix);
    }
    else
        return 2;
    return 0;
}

// ---- fixtures written by the reference's exporters ------------------------------------------------------------
static ast::Material make_material(const std::string& name, bool alpha, float r, float g, float b, float emissive, float metallic, float roughness, const std::string& albedo_source_texture)
{
    ast::Material m;
    m.name = name, m.double_sided = false, m.alpha_mask = alpha, m.material_type = alpha ? ast::MATERIAL_TRANSPARENT : ast::MATERIAL_OPAQUE, m.shading_model = ast::SHADING_MODEL_STANDARD;
    ast::MaterialProperty p;
    p.type = ast::PROPERTY_ALBEDO, p.vec4_value[0] = r, p.vec4_value[1] = g, p.vec4_value[2] = b, p.vec4_value[3] = 1.0f;
    m.properties.push_back(p);
    p.type = ast::PROPERTY_EMISSIVE, p.vec4_value[0] = p.vec4_value[1] = p.vec4_value[2] = emissive, p.vec4_value[3] = 0.0f;
    m.properties.push_back(p); // written with THREE values by the exporter, which the loader then rejects
    p.type = ast::PROPERTY_METALLIC, p.float_value = metallic;
    m.properties.push_back(p);
    p.type = ast::PROPERTY_ROUGHNESS, p.float_value = roughness;
    m.properties.push_back(p);
    if (!albedo_source_texture.empty())
    {
        ast::Texture t;
        t.type = ast::TEXTURE_ALBEDO, t.path = albedo_source_texture, t.srgb = true, t.channel_index = 0;
        m.textures.push_back(t);
        t.type = ast::TEXTURE_ROUGHNESS, t.srgb = false, t.channel_index = 1;
        m.textures.push_back(t);
    }
    return m;
}

static int write_fixtures(const std::string& dir)
{
    // a mesh of two submeshes: a unit quad (2 triangles) and a pyramid (4 triangles), deterministic attributes
    ast::Mesh mesh;
    mesh.name = "fixture_mesh";
    auto add_vertex = [&](float x, float y, float z, float u, float v) {
        ast::Vertex vt;
        vt.position = glm::vec3(x, y, z), vt.tex_coord = glm::vec2(u, v);
        vt.normal = glm::normalize(glm::vec3(0.25f * x, 1.0f, 0.5f * z)), vt.tangent = glm::vec3(1.0f, 0.0f, 0.0f), vt.bitangent = glm::vec3(0.0f, 0.0f, 1.0f);
        mesh.vertices.push_back(vt);
    };
    add_vertex(-1, 0, -1, 0, 0), add_vertex(1, 0, -1, 1, 0), add_vertex(1, 0, 1, 1, 1), add_vertex(-1, 0, 1, 0, 1);
    add_vertex(-0.5f, 0, -0.5f, 0, 0), add_vertex(0.5f, 0, -0.5f, 1, 0), add_vertex(0.5f, 0, 0.5f, 1, 1), add_vertex(-0.5f, 0, 0.5f, 0, 1), add_vertex(0, 0.75f, 0, 0.5f, 0.5f);
    const uint32_t idx[] = { 0, 1, 2, 0, 2, 3, 4, 5, 8, 5, 6, 8, 6, 7, 8, 7, 4, 8 };
    mesh.indices.assign(idx, idx + 18);
    ast::SubMesh s0 {}, s1 {};
    s0.material_index = 0, s0.index_count = 6, s0.vertex_count = 4, s0.base_vertex = 0, s0.base_index = 0, s0.max_extents = glm::vec3(1, 0, 1), s0.min_extents = glm::vec3(-1, 0, -1);
    std::strcpy(s0.name, "quad");
    s1.material_index = 1, s1.index_count = 12, s1.vertex_count = 5, s1.base_vertex = 0, s1.base_index = 6, s1.max_extents = glm::vec3(0.5f, 0.75f, 0.5f), s1.min_extents = glm::vec3(-0.5f, 0, -0.5f);
    std::strcpy(s1.name, "pyramid");
    mesh.submeshes.push_back(s0), mesh.submeshes.push_back(s1);
    mesh.max_extents = glm::vec3(1, 0.75f, 1), mesh.min_extents = glm::vec3(-1, 0, -1);
    // "<dir>/texture/checker.ast" exists already (written by ast_io.py), so the exporter only references it
    mesh.materials.push_back(make_material("fixture_floor", false, 0.8f, 0.7f, 0.6f, 0.0f, 0.0f, 0.9f, dir + "/source/checker.png"));
    mesh.materials.push_back(make_material("fixture_glow", true, 0.1f, 0.2f, 0.3f, 4.0f, 0.25f, 0.5f, ""));
    ast::MeshExportOption mo;
    mo.output_root_folder_path = dir, mo.use_compression = false;
    if (!ast::export_mesh(mesh, mo)) return 1;

    auto root  = std::make_shared<ast::TransformNode>();
    root->type = ast::SCENE_NODE_ROOT, root->name = "root", root->position = glm::vec3(0), root->rotation = glm::vec3(0), root->scale = glm::vec3(1);
    auto mn  = std::make_shared<ast::MeshNode>();
    mn->type = ast::SCENE_NODE_MESH, mn->name = "mesh_a", mn->mesh = "mesh/fixture_mesh.ast", mn->material_override = "", mn->casts_shadow = true;
    mn->position = glm::vec3(0.5f, -0.25f, 2.0f), mn->rotation = glm::vec3(10.0f, 20.0f, 30.0f), mn->scale = glm::vec3(1.0f, 2.0f, 0.5f);
    auto mn2  = std::make_shared<ast::MeshNode>();
    mn2->type = ast::SCENE_NODE_MESH, mn2->name = "mesh_b", mn2->mesh = "mesh/fixture_mesh.ast", mn2->material_override = "material/fixture_glow.json", mn2->casts_shadow = false;
    mn2->position = glm::vec3(-3.0f, 0.0f, 0.0f), mn2->rotation = glm::vec3(0.0f, -45.0f, 0.0f), mn2->scale = glm::vec3(1.0f, 1.0f, 0.0f);
    mn->children.push_back(mn2);
    auto cam  = std::make_shared<ast::CameraNode>();
    cam->type = ast::SCENE_NODE_CAMERA, cam->name = "camera", cam->near_plane = 0.1f, cam->far_plane = 500.0f, cam->fov = 55.0f;
    cam->position = glm::vec3(0.0f, 1.5f, 6.0f), cam->rotation = glm::vec3(-12.0f, 5.0f, 0.0f), cam->scale = glm::vec3(1);
    auto dl  = std::make_shared<ast::DirectionalLightNode>();
    dl->type = ast::SCENE_NODE_DIRECTIONAL_LIGHT, dl->name = "sun", dl->color = glm::vec3(1.0f, 0.9f, 0.8f), dl->intensity = 3.0f, dl->radius = 0.05f, dl->casts_shadows = true;
    dl->position = glm::vec3(0), dl->rotation = glm::vec3(50.0f, 30.0f, 0.0f), dl->scale = glm::vec3(1);
    auto sl  = std::make_shared<ast::SpotLightNode>();
    sl->type = ast::SCENE_NODE_SPOT_LIGHT, sl->name = "spot", sl->color = glm::vec3(0.2f, 0.4f, 1.0f), sl->intensity = 40.0f, sl->radius = 0.1f, sl->casts_shadows = true, sl->inner_cone_angle = 20.0f, sl->outer_cone_angle = 35.0f;
    sl->position = glm::vec3(2.0f, 3.0f, 1.0f), sl->rotation = glm::vec3(90.0f, 0.0f, 0.0f), sl->scale = glm::vec3(1);
    auto pl  = std::make_shared<ast::PointLightNode>();
    pl->type = ast::SCENE_NODE_POINT_LIGHT, pl->name = "bulb", pl->color = glm::vec3(1.0f, 0.5f, 0.25f), pl->intensity = 15.0f, pl->radius = 0.2f, pl->casts_shadows = false;
    pl->position = glm::vec3(-2.0f, 2.0f, -1.0f), pl->rotation = glm::vec3(0), pl->scale = glm::vec3(1);
    auto ibl  = std::make_shared<ast::IBLNode>();
    ibl->type = ast::SCENE_NODE_IBL, ibl->name = "sky", ibl->image = "texture/env.ast";
    root->children = { mn, cam, dl, sl, pl, ibl };
    ast::Scene scene;
    scene.name = "fixture_scene", scene.scene_graph = root;
    return ast::export_scene(scene, dir + "/scene/fixture_scene.json") ? 0 : 1;
}

int main(int argc, char** argv)
{
    const std::string cmd = argc > 1 ? argv[1] : "";
    if (cmd == "dump" && argc >= 4) return dump(argv[2], argv[3], argc > 4 ? argv[4] : "");
    if (cmd == "write-fixtures" && argc >= 3) return write_fixtures(argv[2]);
    std::fprintf(stderr, "usage: ref_ast_tool dump (image|mesh|material|scene) <file> [strip-prefix] | write-fixtures <dir>\n");
    return 2;
}
