// tests/ast/my_ast_dump.cpp — TEST TOOL: loads an AssetCore file with helios_b200's loader (shim: loader/loader.h)
// and prints the canonical text of oracle/ref_ast/dump_format.h, for comparison with the reference loader's dump
// (oracle/_ref/ref_ast_tool) and with the committed golden dumps (tests/golden/ast/*.dump).
//   my_ast_dump (image|mesh|material|scene) <file> [strip-prefix]
//   my_ast_dump texels <image file> <srgb 0|1> <slice> <out.bin>   level 0 as ResourceManager uploads it: prints "format w h", writes the texels
//   my_ast_dump matrix px py pz rx ry rz sx sy sz                   the node's local matrix (recompose_matrix_from_components), 16 bit patterns
#include "../../oracle/ref_ast/dump_format.h"
#include <core/resource_manager.h>
#include <loader/loader.h>
#include <cstdlib>

static void dump_node(FILE* f, const std::shared_ptr<ast::SceneNode>& n, int depth)
{
    if (!n)
    {
        std::fprintf(f, "%*snode null\n", depth, "");
        return;
    }
    std::fprintf(f, "%*snode type=%d name=\"%s\" children=%zu", depth, "", (int)n->type, n->name.c_str(), n->children.size());
    const int t = n->type;
    if (t != ast::SCENE_NODE_IBL && t != ast::SCENE_NODE_CUSTOM) dumpfmt::vec(f, "position", n->position, 3), dumpfmt::vec(f, "rotation", n->rotation, 3), dumpfmt::vec(f, "scale", n->scale, 3);
    if (t == ast::SCENE_NODE_MESH) std::fprintf(f, " mesh=\"%s\" material_override=\"%s\" casts_shadow=%d", n->mesh.c_str(), n->material_override.c_str(), (int)n->casts_shadow);
    if (t == ast::SCENE_NODE_DIRECTIONAL_LIGHT || t == ast::SCENE_NODE_SPOT_LIGHT || t == ast::SCENE_NODE_POINT_LIGHT)
        dumpfmt::vec(f, "color", n->color, 3), dumpfmt::vec(f, "intensity", &n->intensity, 1), dumpfmt::vec(f, "radius", &n->radius, 1), std::fprintf(f, " casts_shadows=%d", (int)n->casts_shadows);
    if (t == ast::SCENE_NODE_SPOT_LIGHT) dumpfmt::vec(f, "inner", &n->inner_cone_angle, 1), dumpfmt::vec(f, "outer", &n->outer_cone_angle, 1);
    if (t == ast::SCENE_NODE_CAMERA) dumpfmt::vec(f, "near", &n->near_plane, 1), dumpfmt::vec(f, "far", &n->far_plane, 1), dumpfmt::vec(f, "fov", &n->fov, 1);
    if (t == ast::SCENE_NODE_IBL) std::fprintf(f, " image=\"%s\"", n->image.c_str());
    std::fprintf(f, "\n");
    for (auto& c : n->children) dump_node(f, c, depth + 1);
}

int main(int argc, char** argv)
{
    if (argc < 3) return std::fprintf(stderr, "usage: my_ast_dump (image|mesh|material|scene) <file> [strip-prefix]\n"), 2;
    const std::string kind = argv[1], path = argv[2], prefix = argc > 3 ? argv[3] : "";
    if (kind == "texels" && argc >= 6)
    {
        ast::Image img;
        if (!ast::load_image(path, img)) return std::printf("load failed\n"), 0;
        int                  fmt = 0;
        uint32_t             w = 0, h = 0;
        std::vector<uint8_t> texels;
        if (!helios::convert_image_level0(img, std::atoi(argv[4]), std::atoi(argv[3]) != 0, fmt, w, h, texels)) return std::printf("no format\n"), 0;
        std::printf("%d %u %u\n", fmt, w, h);
        FILE* f = std::fopen(argv[5], "wb");
        std::fwrite(texels.data(), 1, texels.size(), f);
        std::fclose(f);
        return 0;
    }
    if (kind == "matrix" && argc >= 11)
    {
        float v[9];
        for (int i = 0; i < 9; i++) v[i] = std::strtof(argv[2 + i], nullptr);
        const glm::mat4 m = helios::recompose_matrix_from_components(v, v + 3, v + 6);
        for (int i = 0; i < 16; i++) std::printf("%08x%s", dumpfmt::bits(((const float*)&m)[i]), i == 15 ? "\n" : " ");
        return 0;
    }
    if (kind == "image")
    {
        ast::Image img;
        if (!ast::load_image(path, img)) return std::printf("load failed\n"), 0;
        dumpfmt::image_header(stdout, img.name, img.components, img.mip_slices, img.array_slices, (int)img.type, (int)img.compression);
        for (int a = 0; a < img.array_slices; a++)
            for (int m = 0; m < img.mip_slices; m++)
            {
                const ast::Image::Level& L = img.data[(size_t)a][(size_t)m];
                dumpfmt::image_level(stdout, a, m, (int)L.width, (int)L.height, L.bytes.data(), L.bytes.size());
            }
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
        dump_node(stdout, s.scene_graph, 0);
    }
    else
        return 2;
    return 0;
}
