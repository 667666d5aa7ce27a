// scene.cpp — scene graph, render-state gathering and the Material / Instance / Light table build
// (reference: src/engine/resource/scene.cpp; line references are to that file).
#include <resource/material.h>
#include <resource/mesh.h>
#include <resource/scene.h>
#include <resource/texture.h>
#include <cmath>
#include <cstring>
#include <unordered_set>

namespace helios
{
static uint32_t g_node_counter = 0;

// light types of the shader ABI (common.glsl:11-15)
enum LightType
{
    LIGHT_DIRECTIONAL,
    LIGHT_SPOT,
    LIGHT_POINT,
    LIGHT_ENVIRONMENT_MAP,
    LIGHT_AREA
};

// ---- Node ----------------------------------------------------------------------------------------
Node::Node(const NodeType& type, const std::string& name) : m_type(type), m_name(name), m_id(g_node_counter++) {}
Node::~Node() {}

void Node::add_child(Node::Ptr child)
{
    m_is_heirarchy_dirty = true;
    child->m_parent      = this;
    m_children.push_back(child);
}
Node::Ptr Node::find_child(const std::string& name)
{
    for (auto& child : m_children)
    {
        if (child->m_name == name) return child;
        if (auto found = child->find_child(name)) return found;
    }
    return nullptr;
}
Node::Ptr Node::find_child(const NodeType& type)
{
    for (auto& child : m_children)
    {
        if (child->type() == type) return child;
        if (auto found = child->find_child(type)) return found;
    }
    return nullptr;
}
void Node::remove_child(const std::string& name)
{
    m_is_heirarchy_dirty = true;
    for (size_t i = 0; i < m_children.size(); i++)
        if (m_children[i]->m_name == name)
        {
            m_children.erase(m_children.begin() + (long)i);
            break;
        }
}
// :149-159 — a dirty hierarchy anywhere below marks the frame as a hierarchy update
void Node::update_children(RenderState& render_state)
{
    if (m_is_heirarchy_dirty)
    {
        render_state.m_scene_state = SCENE_STATE_HIERARCHY_UPDATED;
        m_is_heirarchy_dirty       = false;
    }
    for (auto& child : m_children) child->update(render_state);
}
void Node::mark_transforms_as_dirty()
{
    m_is_transform_dirty = true;
    for (auto& child : m_children) child->mark_transforms_as_dirty();
}

// ---- TransformNode -------------------------------------------------------------------------------
TransformNode::TransformNode(const NodeType& type, const std::string& name) : Node(type, name) {}
TransformNode::~TransformNode() {}

// :186-202 — M = T * R * S; note that a transform change also reports HIERARCHY_UPDATED
void TransformNode::update(RenderState& render_state)
{
    if (!m_is_transform_dirty) return;
    const glm::mat4 R = glm::mat4_cast(m_orientation);
    const glm::mat4 S = glm::scale(glm::mat4(1.0f), m_scale);
    const glm::mat4 T = glm::translate(glm::mat4(1.0f), m_position);
    m_prev_model_matrix          = m_model_matrix;
    m_model_matrix_without_scale = T * R;
    m_model_matrix               = m_model_matrix_without_scale * S;
    render_state.m_scene_state   = SCENE_STATE_HIERARCHY_UPDATED;
    m_is_transform_dirty         = false;
}
glm::vec3 TransformNode::forward() { return m_orientation * glm::vec3(0.0f, 0.0f, 1.0f); }
glm::vec3 TransformNode::up() { return m_orientation * glm::vec3(0.0f, 1.0f, 0.0f); }
glm::vec3 TransformNode::left() { return m_orientation * glm::vec3(1.0f, 0.0f, 0.0f); }
glm::vec3 TransformNode::local_position() { return m_position; }
// :234-270 — only the direct parent's matrix is applied (the reference does not walk further up)
glm::vec3 TransformNode::global_position()
{
    if (auto* p = dynamic_cast<TransformNode*>(m_parent)) return p->m_model_matrix_without_scale * glm::vec4(m_position, 1.0f);
    return m_position;
}
glm::mat4 TransformNode::global_transform()
{
    if (auto* p = dynamic_cast<TransformNode*>(m_parent)) return p->m_model_matrix_without_scale * m_model_matrix;
    return m_model_matrix;
}
glm::mat4 TransformNode::global_transform_without_scale()
{
    if (auto* p = dynamic_cast<TransformNode*>(m_parent)) return p->m_model_matrix_without_scale * m_model_matrix_without_scale;
    return m_model_matrix_without_scale;
}
glm::mat4 TransformNode::local_transform() { return m_model_matrix; }
glm::mat4 TransformNode::normal_matrix() { return global_transform_without_scale(); }
glm::quat TransformNode::orientation() { return m_orientation; }
glm::vec3 TransformNode::scale() { return m_scale; }

void TransformNode::set_from_local_transform(const glm::mat4& transform)
{
    mark_transforms_as_dirty();
    glm::vec3 skew;
    glm::vec4 persp;
    glm::decompose(transform, m_scale, m_orientation, m_position, skew, persp);
}
void TransformNode::set_from_global_transform(const glm::mat4& transform)
{
    mark_transforms_as_dirty();
    glm::mat4 local = transform;
    if (auto* p = dynamic_cast<TransformNode*>(m_parent)) local = glm::inverse(p->m_model_matrix_without_scale) * transform;
    glm::vec3 skew;
    glm::vec4 persp;
    glm::decompose(local, m_scale, m_orientation, m_position, skew, persp);
}
void TransformNode::set_orientation(const glm::quat& q)
{
    mark_transforms_as_dirty();
    m_orientation = q;
}
static void euler_quats(const glm::vec3& e, glm::quat& pitch, glm::quat& yaw, glm::quat& roll)
{
    pitch = glm::quat(glm::vec3(glm::radians(e.x), 0.0f, 0.0f));
    yaw   = glm::quat(glm::vec3(0.0f, glm::radians(e.y), 0.0f));
    roll  = glm::quat(glm::vec3(0.0f, 0.0f, glm::radians(e.z)));
}
void TransformNode::set_orientation_from_euler_yxz(const glm::vec3& e)
{
    mark_transforms_as_dirty();
    glm::quat p, y, r;
    euler_quats(e, p, y, r);
    m_orientation = y * p * r;
}
void TransformNode::set_orientation_from_euler_xyz(const glm::vec3& e)
{
    mark_transforms_as_dirty();
    glm::quat p, y, r;
    euler_quats(e, p, y, r);
    m_orientation = p * y * r;
}
void TransformNode::set_position(const glm::vec3& position)
{
    mark_transforms_as_dirty();
    m_position = position;
}
void TransformNode::set_scale(const glm::vec3& scale)
{
    mark_transforms_as_dirty();
    m_scale = scale;
}
void TransformNode::move(const glm::vec3& displacement)
{
    mark_transforms_as_dirty();
    m_position += displacement;
}
void TransformNode::rotate_euler_yxz(const glm::vec3& e)
{
    mark_transforms_as_dirty();
    glm::quat p, y, r;
    euler_quats(e, p, y, r);
    m_orientation = m_orientation * (y * p * r);
}
void TransformNode::rotate_euler_xyz(const glm::vec3& e)
{
    mark_transforms_as_dirty();
    glm::quat p, y, r;
    euler_quats(e, p, y, r);
    m_orientation = m_orientation * (p * y * r);
}

// ---- concrete nodes ------------------------------------------------------------------------------
RootNode::RootNode(const std::string& name) : TransformNode(NODE_ROOT, name) {}
RootNode::~RootNode() {}
void RootNode::update(RenderState& render_state)
{
    if (!m_is_enabled) return;
    TransformNode::update(render_state);
    update_children(render_state);
}

MeshNode::MeshNode(const std::string& name) : TransformNode(NODE_MESH, name) {}
MeshNode::~MeshNode() {}
void MeshNode::update(RenderState& render_state)
{
    if (!m_is_enabled) return;
    TransformNode::update(render_state);
    if (m_mesh) render_state.m_meshes.push_back(this);
    update_children(render_state);
}
void MeshNode::set_mesh(std::shared_ptr<Mesh> mesh)
{
    m_mesh = mesh;
    m_material_indices.assign(mesh ? mesh->sub_meshes().size() : 0, glm::uvec2(0, 0));
    m_is_heirarchy_dirty = true; // a new mesh must reach the tables (the viewer calls Scene::force_update)
}
void MeshNode::set_material_override(std::shared_ptr<Material> material_override) { m_material_override = material_override; }

DirectionalLightNode::DirectionalLightNode(const std::string& name) : TransformNode(NODE_DIRECTIONAL_LIGHT, name), LightParameters(0.1f) {}
DirectionalLightNode::~DirectionalLightNode() {}
void DirectionalLightNode::update(RenderState& render_state)
{
    if (!m_is_enabled) return;
    TransformNode::update(render_state);
    render_state.m_directional_lights.push_back(this);
    update_children(render_state);
}

SpotLightNode::SpotLightNode(const std::string& name) : TransformNode(NODE_SPOT_LIGHT, name), LightParameters(5.0f) {}
SpotLightNode::~SpotLightNode() {}
void SpotLightNode::update(RenderState& render_state)
{
    if (!m_is_enabled) return;
    TransformNode::update(render_state);
    render_state.m_spot_lights.push_back(this);
    update_children(render_state);
}

PointLightNode::PointLightNode(const std::string& name) : TransformNode(NODE_POINT_LIGHT, name), LightParameters(5.0f) {}
PointLightNode::~PointLightNode() {}
void PointLightNode::update(RenderState& render_state)
{
    if (!m_is_enabled) return;
    TransformNode::update(render_state);
    render_state.m_point_lights.push_back(this);
    update_children(render_state);
}

CameraNode::CameraNode(const std::string& name) : TransformNode(NODE_CAMERA, name) {}
CameraNode::~CameraNode() {}
// :632-646
void CameraNode::update(RenderState& render_state)
{
    if (!m_is_enabled) return;
    TransformNode::update(render_state);
    m_projection_matrix = glm::perspective(glm::radians(m_fov), float(render_state.viewport_width()) / float(render_state.viewport_height()), m_near_plane, m_far_plane);
    m_view_matrix       = glm::inverse(global_transform_without_scale());
    if (!render_state.m_camera) render_state.m_camera = this;
    update_children(render_state);
}
glm::vec3 CameraNode::camera_forward() { return -forward(); }
glm::vec3 CameraNode::camera_left() { return -left(); }

IBLNode::IBLNode(const std::string& name) : Node(NODE_IBL, name) {}
IBLNode::~IBLNode() {}
void IBLNode::update(RenderState& render_state)
{
    if (!m_is_enabled) return;
    if (!render_state.m_ibl_environment_map) render_state.m_ibl_environment_map = this;
    update_children(render_state);
}
void IBLNode::set_image(std::shared_ptr<TextureCube> image)
{
    m_image              = image;
    m_is_heirarchy_dirty = true;
}

// ---- RenderState ---------------------------------------------------------------------------------
RenderState::RenderState()
{
    m_meshes.reserve(MAX_SCENE_MESH_INSTANCE_COUNT);
}
RenderState::~RenderState() {}
void RenderState::clear()
{
    m_meshes.clear();
    m_directional_lights.clear();
    m_spot_lights.clear();
    m_point_lights.clear();
    m_camera              = nullptr;
    m_ibl_environment_map = nullptr;
    m_cmd_buffer          = nullptr;
    m_scene               = nullptr;
    m_num_lights          = 0;
    m_scene_state         = SCENE_STATE_READY;
}
void RenderState::setup(uint32_t width, uint32_t height, vk::CommandBuffer::Ptr cmd_buffer)
{
    clear();
    m_viewport_width  = width;
    m_viewport_height = height;
    m_cmd_buffer      = cmd_buffer;
}

// ---- Scene ---------------------------------------------------------------------------------------
Scene::Ptr Scene::create(vk::Backend::Ptr backend, const std::string& name, Node::Ptr root, const std::string& path) { return std::shared_ptr<Scene>(new Scene(backend, name, root, path)); }
Scene::Scene(vk::Backend::Ptr backend, const std::string& name, Node::Ptr root, const std::string& path) : vk::Object(backend), m_root(root), m_backend(backend), m_name(name), m_path(path)
{
    m_sky_model = std::unique_ptr<HosekWilkieSkyModel>(new HosekWilkieSkyModel(backend));
}
Scene::~Scene() {}

void      Scene::set_root_node(Node::Ptr node) { m_root = node; }
Node::Ptr Scene::root_node() { return m_root; }
Node::Ptr Scene::find_node(const std::string& name)
{
    if (!m_root) return nullptr;
    if (m_root->name() == name) return m_root;
    return m_root->find_child(name);
}
CameraNode::Ptr Scene::find_camera()
{
    if (!m_root) return nullptr;
    if (m_root->type() == NODE_CAMERA) return std::dynamic_pointer_cast<CameraNode>(m_root);
    return std::dynamic_pointer_cast<CameraNode>(m_root->find_child(NODE_CAMERA));
}

// :875-911
void Scene::update(RenderState& render_state)
{
    render_state.m_scene = this;
    if (!m_root) return;
    m_root->update(render_state);
    render_state.m_num_lights = m_num_area_lights + (uint32_t)render_state.m_directional_lights.size() + (uint32_t)render_state.m_spot_lights.size() + (uint32_t)render_state.m_point_lights.size();
    const bool has_ibl = render_state.ibl_environment_map() && render_state.ibl_environment_map()->image();
    if (has_ibl)
        render_state.m_num_lights++;
    else if (render_state.m_directional_lights.size() > 0)
    {
        render_state.m_num_lights++;
        // the reference re-fits and re-bakes the sky every frame; the result only depends on the direction, so the
        // shim skips the device bake when the direction did not change (identical output)
        const glm::vec3 sun = -render_state.m_directional_lights[0]->forward();
        if (!m_sky_valid || sun.x != m_last_sun_direction.x || sun.y != m_last_sun_direction.y || sun.z != m_last_sun_direction.z)
        {
            m_sky_model->update(render_state.cmd_buffer(), sun);
            m_last_sun_direction = sun, m_sky_valid = true, m_env_source_id = 0xFFFFFFFFu;
        }
    }
    if (m_force_update)
    {
        render_state.m_scene_state = SCENE_STATE_HIERARCHY_UPDATED;
        m_force_update             = false;
    }
    create_gpu_resources(render_state);
    // m_num_area_lights is known only after the tables were rebuilt (the reference has the same one-frame lag
    // on the first frame, hidden by its restart); recompute so that the count handed to the integrator is exact
    render_state.m_num_lights = (uint32_t)m_tables.lights.size();
}

static void copy_mat4(float dst[16], const glm::mat4& m) { std::memcpy(dst, glm::value_ptr(m), 64); }

// :915-1311
void Scene::create_gpu_resources(RenderState& render_state)
{
    if (render_state.m_scene_state == SCENE_STATE_READY) return;
    auto backend = m_backend.lock();
    if (!backend)
    {
        HELIOS_LOG_FATAL("Scene::update: the backend was destroyed before the scene");
        throw std::runtime_error("Scene::update: the backend was destroyed before the scene");
    }
    if (render_state.m_meshes.size() > MAX_SCENE_MESH_INSTANCE_COUNT)
    {
        HELIOS_LOG_FATAL("Scene::update: more than MAX_SCENE_MESH_INSTANCE_COUNT mesh instances");
        throw std::runtime_error("Scene::update: more than MAX_SCENE_MESH_INSTANCE_COUNT mesh instances");
    }
    SceneTables& T = m_tables;
    m_previous_tables = T; // what the device holds (when m_tables_installed): a change of transforms only becomes a refit
    T.materials.clear(), T.instances.clear(), T.lights.clear(), T.submesh_info.clear();
    m_num_area_lights = 0;
    m_global_mesh_indices.clear();
    std::unordered_set<uint32_t>           processed_meshes, processed_materials, processed_textures;
    std::unordered_map<uint32_t, uint32_t> global_material_indices;
    std::vector<std::shared_ptr<Texture2D>> texture_array; // descriptor set 4, in first-use order
    uint32_t                               mesh_index_counter = 0;

    // a texture takes the next slot of the global array the FIRST time any material uses it; a later material
    // using the same texture keeps -1 in its table row (reference behaviour, :958-1066, reproduced as is)
    auto claim = [&](const std::shared_ptr<Texture2D>& tex, int32_t& slot) {
        if (processed_textures.find(tex->id()) != processed_textures.end()) return false;
        processed_textures.insert(tex->id());
        slot = (int32_t)texture_array.size();
        texture_array.push_back(tex);
        return true;
    };

    for (size_t mesh_node_idx = 0; mesh_node_idx < render_state.m_meshes.size(); mesh_node_idx++)
    {
        MeshNode*   mesh_node = render_state.m_meshes[mesh_node_idx];
        auto        mesh      = mesh_node->mesh();
        const auto& materials = mesh->materials();
        const auto& submeshes = mesh->sub_meshes();
        if (processed_meshes.find(mesh->id()) == processed_meshes.end())
        {
            processed_meshes.insert(mesh->id());
            m_global_mesh_indices[mesh->id()] = mesh_index_counter++;
            for (size_t i = 0; i < submeshes.size(); i++)
            {
                const SubMesh& submesh  = submeshes[i];
                auto           material = materials[submesh.mat_idx];
                if (mesh_node->material_override()) material = mesh_node->material_override();
                if (processed_materials.find(material->id()) == processed_materials.end())
                {
                    processed_materials.insert(material->id());
                    if (T.materials.size() >= MAX_SCENE_MATERIAL_COUNT) throw std::runtime_error("Scene::update: more than MAX_SCENE_MATERIAL_COUNT materials");
                    hl_material md;
                    for (int k = 0; k < 4; k++) md.texture_indices0[k] = -1, md.texture_indices1[k] = -1, md.albedo[k] = 0.0f, md.emissive[k] = 0.0f, md.roughness_metallic[k] = 0.0f;
                    if (material->albedo_texture())
                        claim(material->albedo_texture(), md.texture_indices0[0]);
                    else
                    {
                        // constant albedo is authored in sRGB: pow(rgb, 2.2), alpha untouched (:973-977)
                        const glm::vec4 a   = material->albedo_value();
                        const glm::vec3 lin = glm::pow(glm::vec3(a.x, a.y, a.z), glm::vec3(2.2f));
                        md.albedo[0] = lin.x, md.albedo[1] = lin.y, md.albedo[2] = lin.z, md.albedo[3] = a.w;
                    }
                    if (material->normal_texture()) claim(material->normal_texture(), md.texture_indices0[1]);
                    if (material->roughness_texture())
                    {
                        if (claim(material->roughness_texture(), md.texture_indices0[2])) md.texture_indices1[2] = material->roughness_texture_info().array_index;
                    }
                    else
                        md.roughness_metallic[0] = material->roughness_value();
                    if (material->metallic_texture())
                    {
                        if (claim(material->metallic_texture(), md.texture_indices0[3])) md.texture_indices1[3] = material->metallic_texture_info().array_index;
                    }
                    else
                        md.roughness_metallic[1] = material->metallic_value();
                    if (material->emissive_texture())
                        claim(material->emissive_texture(), md.texture_indices1[0]);
                    else
                    {
                        const glm::vec4 e = material->emissive_value();
                        md.emissive[0] = e.x, md.emissive[1] = e.y, md.emissive[2] = e.z, md.emissive[3] = e.w;
                    }
                    global_material_indices[material->id()] = (uint32_t)T.materials.size();
                    T.materials.push_back(md);
                }
                // area lights: one per emissive submesh of the FIRST node that uses a mesh (:1087-1096)
                if (material->is_emissive())
                {
                    m_num_area_lights++;
                    hl_light L {};
                    L.light_data0[0] = float(LIGHT_AREA), L.light_data0[1] = float(mesh_node_idx), L.light_data0[2] = float(global_material_indices[material->id()]);
                    L.light_data0[3] = float(submesh.base_index / 3);
                    L.light_data1[0] = float(submesh.index_count / 3);
                    T.lights.push_back(L);
                }
            }
        }
        // (primitive offset, material index) per submesh of this instance (:1106-1119)
        auto& pairs = mesh_node->material_indices_buffer();
        pairs.resize(submeshes.size());
        for (size_t i = 0; i < submeshes.size(); i++)
        {
            auto material = materials[submeshes[i].mat_idx];
            if (mesh_node->material_override()) material = mesh_node->material_override();
            pairs[i] = glm::uvec2(submeshes[i].base_index / 3, global_material_indices[material->id()]);
        }
        T.submesh_info.push_back(pairs);
        // instance row (:1243-1250)
        hl_instance inst {};
        copy_mat4(inst.model_matrix, mesh_node->global_transform());
        copy_mat4(inst.normal_matrix, mesh_node->normal_matrix());
        inst.mesh_index = m_global_mesh_indices[mesh->id()];
        T.instances.push_back(inst);
    }
    // light list order: area..., environment, directional..., point..., spot... (:1253-1306)
    const bool has_ibl = render_state.ibl_environment_map() && render_state.ibl_environment_map()->image();
    if (has_ibl || render_state.m_directional_lights.size() > 0)
    {
        hl_light L {};
        L.light_data0[0] = float(LIGHT_ENVIRONMENT_MAP);
        T.lights.push_back(L);
    }
    for (auto* light : render_state.m_directional_lights)
    {
        hl_light        L {};
        const glm::vec3 c = light->color(), f = light->forward();
        L.light_data0[0] = float(LIGHT_DIRECTIONAL), L.light_data0[1] = c.x, L.light_data0[2] = c.y, L.light_data0[3] = c.z;
        L.light_data1[0] = f.x, L.light_data1[1] = f.y, L.light_data1[2] = f.z, L.light_data1[3] = light->intensity();
        L.light_data2[3] = light->radius();
        T.lights.push_back(L);
    }
    for (auto* light : render_state.m_point_lights)
    {
        hl_light        L {};
        const glm::vec3 c = light->color(), p = light->global_position();
        L.light_data0[0] = float(LIGHT_POINT), L.light_data0[1] = c.x, L.light_data0[2] = c.y, L.light_data0[3] = c.z;
        L.light_data1[3] = light->intensity();
        L.light_data2[0] = p.x, L.light_data2[1] = p.y, L.light_data2[2] = p.z, L.light_data2[3] = light->radius();
        T.lights.push_back(L);
    }
    for (auto* light : render_state.m_spot_lights)
    {
        hl_light        L {};
        const glm::vec3 c = light->color(), f = light->forward(), p = light->global_position();
        L.light_data0[0] = float(LIGHT_SPOT), L.light_data0[1] = c.x, L.light_data0[2] = c.y, L.light_data0[3] = c.z;
        L.light_data1[0] = f.x, L.light_data1[1] = f.y, L.light_data1[2] = f.z, L.light_data1[3] = light->intensity();
        L.light_data2[0] = p.x, L.light_data2[1] = p.y, L.light_data2[2] = p.z, L.light_data2[3] = light->radius();
        L.light_data3[0] = cosf(glm::radians(light->inner_cone_angle())), L.light_data3[1] = cosf(glm::radians(light->outer_cone_angle()));
        T.lights.push_back(L);
    }
    if (T.lights.size() > MAX_SCENE_LIGHT_COUNT) throw std::runtime_error("Scene::update: more than MAX_SCENE_LIGHT_COUNT lights");
    T.num_textures = (uint32_t)texture_array.size();
    m_texture_array_ids.clear();
    for (auto& tex : texture_array) m_texture_array_ids.push_back(tex->id());

    if (!backend->has_device()) return; // host-only inspection of the tables
    hl_context ctx = backend->context();
    std::vector<hl_mesh> handles(T.instances.size());
    for (size_t i = 0; i < T.instances.size(); i++) handles[i] = render_state.m_meshes[i]->mesh()->acceleration_structure();
    std::vector<uint32_t> texture_ids;
    for (auto& tex : texture_array) texture_ids.push_back(tex->id());
    m_texture_array_ids = texture_ids;
    // Moving mesh nodes (the common interactive edit): same materials, lights, textures, meshes and submesh tables, only the
    // instances' matrices differ -> hl_scene_update_instances refits the instance tree instead of rebuilding everything.
    // (the reference rebuilds its TLAS here although it was created ALLOW_UPDATE, scene.cpp:797, renderer.cpp:147-168)
    {
        const SceneTables& P    = m_previous_tables;
        auto               same = [](const void* a, const void* b, size_t n) { return n == 0 || std::memcmp(a, b, n) == 0; };
        auto               same_info = [&](const std::vector<std::vector<glm::uvec2>>& a, const std::vector<std::vector<glm::uvec2>>& b) {
            if (a.size() != b.size()) return false;
            for (size_t i = 0; i < a.size(); i++)
                if (a[i].size() != b[i].size() || !same(a[i].data(), b[i].data(), sizeof(glm::uvec2) * a[i].size())) return false;
            return true;
        };
        bool               only_transforms = m_tables_installed && !T.instances.empty() && P.instances.size() == T.instances.size() && P.materials.size() == T.materials.size() &&
                               P.lights.size() == T.lights.size() && same_info(P.submesh_info, T.submesh_info) && handles == m_installed_meshes && texture_ids == m_installed_textures &&
                               same(P.materials.data(), T.materials.data(), sizeof(hl_material) * T.materials.size()) && same(P.lights.data(), T.lights.data(), sizeof(hl_light) * T.lights.size());
        const bool env_unchanged = (has_ibl && m_env_source_id == render_state.ibl_environment_map()->image()->id()) || (!has_ibl && !render_state.m_directional_lights.empty() && m_sky_valid) ||
                                   (!has_ibl && render_state.m_directional_lights.empty() && m_env_source_id == 0xFFFFFFFFu && !m_sky_valid && m_black_env_installed);
        for (size_t i = 0; only_transforms && i < T.instances.size(); i++) only_transforms = P.instances[i].mesh_index == T.instances[i].mesh_index;
        if (only_transforms && env_unchanged)
        {
            backend->check(hl_scene_update_instances(ctx, T.instances.data(), (uint32_t)T.instances.size()), "hl_scene_update_instances");
            return;
        }
    }
    m_tables_installed = false;
    // textures: slot i of the device array = texture_array[i]
    backend->check(hl_textures_clear(ctx), "hl_textures_clear");
    for (auto& tex : texture_array)
    {
        int32_t slot = -1;
        backend->check(hl_texture2d_create(ctx, tex->format(), tex->width(), tex->height(), tex->texels().data(), &slot), "hl_texture2d_create");
    }
    // environment: IBL image, else the sky baked by HosekWilkieSkyModel::update, else the black default cube (:1122-1135)
    if (has_ibl)
    {
        auto cube = render_state.ibl_environment_map()->image();
        if (m_env_source_id != cube->id())
        {
            backend->check(hl_envmap_set(ctx, cube->size(), cube->faces().data()), "hl_envmap_set");
            m_env_source_id = cube->id(), m_sky_valid = false;
        }
    }
    else if (render_state.m_directional_lights.empty())
    {
        backend->check(hl_envmap_set(ctx, 0, nullptr), "hl_envmap_set");
        m_env_source_id = 0xFFFFFFFFu, m_sky_valid = false, m_black_env_installed = true;
    }
    if (has_ibl || !render_state.m_directional_lights.empty()) m_black_env_installed = false;
    std::vector<const uint32_t*> info(T.instances.size());
    for (size_t i = 0; i < T.instances.size(); i++) info[i] = reinterpret_cast<const uint32_t*>(T.submesh_info[i].data());
    backend->check(hl_scene_set_tables(ctx, T.materials.data(), (uint32_t)T.materials.size(), T.instances.data(), handles.data(), info.data(), (uint32_t)T.instances.size(), T.lights.data(),
                                       (uint32_t)T.lights.size()),
                   "hl_scene_set_tables");
    m_tables_installed = true, m_installed_meshes = handles, m_installed_textures = texture_ids;
}
} // namespace helios
