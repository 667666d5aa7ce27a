// oracle/ref_ast/dump_format.h — TEST INFRASTRUCTURE.  Canonical text form of loaded AssetCore assets, shared by
// the two dump tools so that "reference loader" and "helios_b200 loader" outputs can be compared byte for byte:
//   oracle/ref_ast/ref_ast_tool.cpp   links the REFERENCE's external/AssetCore/src/loader/loader.cpp
//   tests/ast/my_ast_dump.cpp         links helios_b200/shim (loader/loader.h)
// Floats are printed as their bit patterns; bulk payloads as length + FNV-1a hash; paths with a caller-given
// prefix removed (fixtures live in temporary directories).  The material / mesh printers are templates: both
// data models use the same member names for these types.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

namespace dumpfmt
{
inline uint64_t fnv1a(const void* p, size_t n, uint64_t h = 1469598103934665603ull)
{
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
    return h;
}
inline uint32_t bits(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline std::string strip(const std::string& s, const std::string& prefix) { return !prefix.empty() && s.compare(0, prefix.size(), prefix) == 0 ? s.substr(prefix.size()) : s; }
inline void        vec(FILE* f, const char* key, const float* v, int n)
{
    std::fprintf(f, " %s=", key);
    for (int i = 0; i < n; i++) std::fprintf(f, "%s%08x", i ? "," : "", bits(v[i]));
}
inline std::string cstr(const char* p, size_t n)
{
    size_t len = 0;
    while (len < n && p[len]) len++;
    return std::string(p, len);
}

template <class Material>
void material(FILE* f, const Material& m, const std::string& prefix, const char* indent = "")
{
    std::fprintf(f, "%smaterial name=\"%s\" double_sided=%d alpha_mask=%d type=%d shading=%d textures=%zu properties=%zu\n", indent, m.name.c_str(), (int)m.double_sided, (int)m.alpha_mask, (int)m.material_type,
                 (int)m.shading_model, m.textures.size(), m.properties.size());
    for (const auto& t : m.textures) std::fprintf(f, "%s texture type=%d path=\"%s\" srgb=%d channel=%d\n", indent, (int)t.type, strip(t.path, prefix).c_str(), (int)t.srgb, (int)t.channel_index);
    for (const auto& p : m.properties)
    {
        std::fprintf(f, "%s property type=%d", indent, (int)p.type);
        if ((int)p.type <= 1)
            vec(f, "value", p.vec4_value, 4);
        else
            vec(f, "value", &p.float_value, 1);
        std::fprintf(f, "\n");
    }
}

template <class Mesh>
void mesh(FILE* f, const Mesh& m, const std::string& prefix)
{
    std::fprintf(f, "mesh name=\"%s\" vertices=%zu skeletal=%zu indices=%zu submeshes=%zu materials=%zu", m.name.c_str(), m.vertices.size(), m.skeletal_vertices.size(), m.indices.size(), m.submeshes.size(),
                 m.materials.size());
    vec(f, "max", (const float*)&m.max_extents, 3), vec(f, "min", (const float*)&m.min_extents, 3);
    std::fprintf(f, "\n vertices bytes=%zu fnv=%016llx\n", m.vertices.size() * sizeof(m.vertices[0]), (unsigned long long)fnv1a(m.vertices.data(), m.vertices.size() * sizeof(m.vertices[0])));
    std::fprintf(f, " indices bytes=%zu fnv=%016llx\n", m.indices.size() * 4, (unsigned long long)fnv1a(m.indices.data(), m.indices.size() * 4));
    for (size_t i = 0; i < m.submeshes.size(); i++)
    {
        const auto& s = m.submeshes[i];
        std::fprintf(f, " submesh %zu material_index=%u index_count=%u vertex_count=%u base_vertex=%u base_index=%u", i, s.material_index, s.index_count, s.vertex_count, s.base_vertex, s.base_index);
        vec(f, "max", (const float*)&s.max_extents, 3), vec(f, "min", (const float*)&s.min_extents, 3);
        std::fprintf(f, " name=\"%s\"\n", cstr(s.name, sizeof(s.name)).c_str());
    }
    for (size_t i = 0; i < m.material_paths.size(); i++) std::fprintf(f, " material_path %zu \"%s\"\n", i, strip(m.material_paths[i], prefix).c_str());
    for (size_t i = 0; i < m.materials.size(); i++) material(f, m.materials[i], prefix, " ");
}

inline void image_header(FILE* f, const std::string& name, int components, int mips, int slices, int type, int compression)
{
    std::fprintf(f, "image name=\"%s\" components=%d mips=%d slices=%d type=%d compression=%d\n", name.c_str(), components, mips, slices, type, compression);
}
inline void image_level(FILE* f, int a, int m, int w, int h, const void* data, size_t size)
{
    std::fprintf(f, " level slice=%d mip=%d width=%d height=%d bytes=%zu fnv=%016llx\n", a, m, w, h, size, (unsigned long long)fnv1a(data, size));
}
} // namespace dumpfmt
