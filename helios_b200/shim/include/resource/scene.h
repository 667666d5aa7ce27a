// resource/scene.h — the scene graph the path reads (reference: include/resource/scene.h:11-468,
// src/engine/resource/scene.cpp).  Node / TransformNode / RootNode / MeshNode / DirectionalLightNode /
// SpotLightNode / PointLightNode / CameraNode / IBLNode / RenderState / Scene keep the reference's public
// surface.  Scene::update(RenderState&) gathers the render state, runs the Hosek-Wilkie fit when a directional
// light drives the sky, and — on a hierarchy or transform change — rebuilds the Material / Instance / Light
// tables and the per-instance (primitive offset, material) pairs with the reference's rules
// (scene.cpp:915-1311, quirks included) and hands them to hl_scene_set_tables, which also builds the top
// level BVH (the reference does that in Renderer::render, renderer.cpp:113-181).
#pragma once
#include <gfx/hosek_wilkie_sky_model.h>
#include <gfx/vk.h>
#include <glm.hpp>
#include <gtc/quaternion.hpp>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace helios
{
#define MAX_SCENE_MESH_INSTANCE_COUNT 1024
#define MAX_SCENE_LIGHT_COUNT 100000
#define MAX_SCENE_MATERIAL_COUNT 4096
#define MAX_SCENE_MATERIAL_TEXTURE_COUNT (MAX_SCENE_MATERIAL_COUNT * 4)

class Scene;
class Mesh;
class Material;
class TextureCube;
class RenderState;

enum NodeType
{
    NODE_MESH,
    NODE_CAMERA,
    NODE_DIRECTIONAL_LIGHT,
    NODE_SPOT_LIGHT,
    NODE_POINT_LIGHT,
    NODE_IBL,
    NODE_ROOT
};

class Node
{
public:
    using Ptr = std::shared_ptr<Node>;
    friend class Scene;

    Node(const NodeType& type, const std::string& name);
    virtual ~Node();

    virtual void update(RenderState& render_state) = 0;

    void add_child(Node::Ptr child);
    Node::Ptr find_child(const std::string& name);
    Node::Ptr find_child(const NodeType& type);
    void remove_child(const std::string& name);

    bool is_enabled() { return m_is_enabled; }
    bool is_transform_dirty() { return m_is_transform_dirty; }
    void enable() { m_is_enabled = true; }
    void disable() { m_is_enabled = false; }
    const std::vector<std::shared_ptr<Node>>& children() { return m_children; }
    std::string name() { return m_name; }
    Node* parent() { return m_parent; }
    uint32_t id() { return m_id; }
    NodeType type() { return m_type; }

protected:
    void update_children(RenderState& render_state);
    void mark_transforms_as_dirty();

    NodeType m_type;
    bool m_is_enabled = true;
    bool m_is_transform_dirty = true;
    bool m_is_heirarchy_dirty = true;
    std::string m_name;
    Node* m_parent = nullptr;
    std::vector<std::shared_ptr<Node>> m_children;
    uint32_t m_id = 0;
};

class TransformNode : public Node
{
public:
    using Ptr = std::shared_ptr<TransformNode>;

    TransformNode(const NodeType& type, const std::string& name);
    ~TransformNode();

    void update(RenderState& render_state) override;

    glm::vec3 forward();
    glm::vec3 up();
    glm::vec3 left();
    glm::vec3 local_position();
    glm::vec3 global_position();
    glm::mat4 global_transform();
    glm::mat4 global_transform_without_scale();
    glm::mat4 local_transform();
    glm::mat4 normal_matrix();
    glm::quat orientation();
    glm::vec3 scale();
    void set_from_local_transform(const glm::mat4& transform);
    void set_from_global_transform(const glm::mat4& transform);
    void set_orientation(const glm::quat& q);
    void set_orientation_from_euler_yxz(const glm::vec3& e);
    void set_orientation_from_euler_xyz(const glm::vec3& e);
    void set_position(const glm::vec3& position);
    void set_scale(const glm::vec3& scale);
    void move(const glm::vec3& displacement);
    void rotate_euler_yxz(const glm::vec3& e);
    void rotate_euler_xyz(const glm::vec3& e);

protected:
    glm::vec3 m_position = glm::vec3(0.0f);
    glm::quat m_orientation = glm::quat(glm::vec3(0.0f));
    glm::vec3 m_scale = glm::vec3(1.0f);
    glm::mat4 m_prev_model_matrix = glm::mat4(1.0f);
    glm::mat4 m_model_matrix = glm::mat4(1.0f);
    glm::mat4 m_model_matrix_without_scale = glm::mat4(1.0f);
};

class RootNode : public TransformNode
{
public:
    using Ptr = std::shared_ptr<RootNode>;
    RootNode(const std::string& name);
    ~RootNode();
    void update(RenderState& render_state) override;
};

class MeshNode : public TransformNode
{
public:
    using Ptr = std::shared_ptr<MeshNode>;
    MeshNode(const std::string& name);
    ~MeshNode();
    void update(RenderState& render_state) override;

    void set_mesh(std::shared_ptr<Mesh> mesh);
    void set_material_override(std::shared_ptr<Material> material_override);
    std::shared_ptr<Mesh> mesh() { return m_mesh; }
    std::shared_ptr<Material> material_override() { return m_material_override; }
    // (primitive offset, global material index) per submesh: descriptor set 3 of the reference
    std::vector<glm::uvec2>& material_indices_buffer() { return m_material_indices; }

private:
    std::shared_ptr<Mesh> m_mesh;
    std::shared_ptr<Material> m_material_override;
    std::vector<glm::uvec2> m_material_indices;
};

// colour / intensity / radius shared by the three punctual light nodes (the reference repeats these six accessors in each
// class, include/resource/scene.h:154-232; same names and semantics here, one definition)
class LightParameters
{
public:
    void set_color(const glm::vec3& color) { m_color = color; }
    void set_intensity(const float& intensity) { m_intensity = intensity; }
    void set_radius(const float& r) { m_radius = r; }
    glm::vec3 color() { return m_color; }
    float intensity() { return m_intensity; }
    float radius() { return m_radius; }

protected:
    explicit LightParameters(float default_radius) : m_radius(default_radius) {}
    glm::vec3 m_color = glm::vec3(1.0f);
    float m_intensity = 1.0f;
    float m_radius;
};

class DirectionalLightNode : public TransformNode, public LightParameters
{
public:
    using Ptr = std::shared_ptr<DirectionalLightNode>;
    DirectionalLightNode(const std::string& name);
    ~DirectionalLightNode();
    void update(RenderState& render_state) override;
};

class SpotLightNode : public TransformNode, public LightParameters
{
public:
    using Ptr = std::shared_ptr<SpotLightNode>;
    SpotLightNode(const std::string& name);
    ~SpotLightNode();
    void update(RenderState& render_state) override;

    void set_inner_cone_angle(const float& cone_angle) { m_inner_cone_angle = cone_angle; }
    void set_outer_cone_angle(const float& cone_angle) { m_outer_cone_angle = cone_angle; }
    float inner_cone_angle() { return m_inner_cone_angle; }
    float outer_cone_angle() { return m_outer_cone_angle; }

private:
    float m_inner_cone_angle = 40.0f;
    float m_outer_cone_angle = 50.0f;
};

class PointLightNode : public TransformNode, public LightParameters
{
public:
    using Ptr = std::shared_ptr<PointLightNode>;
    PointLightNode(const std::string& name);
    ~PointLightNode();
    void update(RenderState& render_state) override;
};

class CameraNode : public TransformNode
{
public:
    using Ptr = std::shared_ptr<CameraNode>;
    CameraNode(const std::string& name);
    ~CameraNode();
    void update(RenderState& render_state) override;

    glm::vec3 camera_forward();
    glm::vec3 camera_left();
    void set_near_plane(const float& near_plane) { m_near_plane = near_plane; }
    void set_far_plane(const float& far_plane) { m_far_plane = far_plane; }
    void set_fov(const float& fov) { m_fov = fov; }
    void set_focal_length(const float& focal_length) { m_focal_length = focal_length; }
    void set_aperture_radius(const float& aperture_radius) { m_aperture_radius = aperture_radius; }
    float near_plane() { return m_near_plane; }
    float far_plane() { return m_far_plane; }
    float fov() { return m_fov; }
    float focal_length() { return m_focal_length; }
    float aperture_radius() { return m_aperture_radius; }
    glm::mat4 view_matrix() { return m_view_matrix; }
    glm::mat4 projection_matrix() { return m_projection_matrix; }

private:
    float m_near_plane = 1.0f;
    float m_far_plane = 1000.0f;
    float m_fov = 60.0f;
    float m_focal_length = 8.0f;
    float m_aperture_radius = 0.1f;
    glm::mat4 m_view_matrix = glm::mat4(1.0f);
    glm::mat4 m_projection_matrix = glm::mat4(1.0f);
};

class IBLNode : public Node
{
public:
    using Ptr = std::shared_ptr<IBLNode>;
    IBLNode(const std::string& name);
    ~IBLNode();
    void update(RenderState& render_state) override;

    void set_image(std::shared_ptr<TextureCube> image);
    std::shared_ptr<TextureCube> image() { return m_image; }

private:
    std::shared_ptr<TextureCube> m_image;
};

enum SceneState
{
    SCENE_STATE_READY,
    SCENE_STATE_HIERARCHY_UPDATED,
    SCENE_STATE_TRANSFORMS_UPDATED
};

class RenderState
{
public:
    friend class Node;
    friend class TransformNode;
    friend class MeshNode;
    friend class DirectionalLightNode;
    friend class SpotLightNode;
    friend class PointLightNode;
    friend class CameraNode;
    friend class IBLNode;
    friend class Scene;
    friend class Renderer;

    RenderState();
    ~RenderState();

    void clear();
    void setup(uint32_t width, uint32_t height, vk::CommandBuffer::Ptr cmd_buffer);

    const std::vector<MeshNode*>& meshes() { return m_meshes; }
    const std::vector<DirectionalLightNode*>& directional_lights() { return m_directional_lights; }
    const std::vector<SpotLightNode*>& spot_lights() { return m_spot_lights; }
    const std::vector<PointLightNode*>& point_lights() { return m_point_lights; }
    CameraNode* camera() { return m_camera; }
    IBLNode* ibl_environment_map() { return m_ibl_environment_map; }
    SceneState scene_state() { return m_scene_state; }
    Scene* scene() { return m_scene; }
    uint32_t viewport_width() { return m_viewport_width; }
    uint32_t viewport_height() { return m_viewport_height; }
    uint32_t num_lights() { return m_num_lights; }
    vk::CommandBuffer::Ptr cmd_buffer() { return m_cmd_buffer; }

private:
    std::vector<MeshNode*> m_meshes;
    std::vector<DirectionalLightNode*> m_directional_lights;
    std::vector<SpotLightNode*> m_spot_lights;
    std::vector<PointLightNode*> m_point_lights;
    CameraNode* m_camera = nullptr;
    IBLNode* m_ibl_environment_map = nullptr;
    SceneState m_scene_state = SCENE_STATE_READY;
    Scene* m_scene = nullptr;
    uint32_t m_viewport_width = 0;
    uint32_t m_viewport_height = 0;
    uint32_t m_num_lights = 0;
    vk::CommandBuffer::Ptr m_cmd_buffer;
};

// the tables of the last Scene::update, as handed to hl_scene_set_tables (host copies; tools and tests read them)
struct SceneTables
{
    std::vector<hl_material> materials;
    std::vector<hl_instance> instances;
    std::vector<hl_light> lights;
    std::vector<std::vector<glm::uvec2>> submesh_info; // per instance
    uint32_t num_textures = 0;
};

class Scene : public vk::Object
{
public:
    using Ptr = std::shared_ptr<Scene>;

    static Scene::Ptr create(vk::Backend::Ptr backend, const std::string& name, Node::Ptr root = nullptr, const std::string& path = "");
    ~Scene();

    void update(RenderState& render_state);
    void set_root_node(Node::Ptr node);
    Node::Ptr root_node();
    Node::Ptr find_node(const std::string& name);
    CameraNode::Ptr find_camera();

    void set_name(const std::string& name) { m_name = name; }
    std::string name() { return m_name; }
    std::string path() { return m_path; }
    void force_update() { m_force_update = true; }
    HosekWilkieSkyModel* sky_model() { return m_sky_model.get(); }
    const SceneTables& tables() const { return m_tables; }
    // Texture2D ids in the order of the device texture array (descriptor set 4 of the reference), as of the last table build
    const std::vector<uint32_t>& texture_array_ids() const { return m_texture_array_ids; }

private:
    Scene(vk::Backend::Ptr backend, const std::string& name, Node::Ptr root = nullptr, const std::string& path = "");
    void create_gpu_resources(RenderState& render_state);

    Node::Ptr m_root;
    std::unordered_map<uint32_t, uint32_t> m_global_material_indices;
    std::unordered_map<uint32_t, uint32_t> m_global_mesh_indices;
    uint32_t m_num_area_lights = 0;
    std::unique_ptr<HosekWilkieSkyModel> m_sky_model;
    std::weak_ptr<vk::Backend> m_backend;
    std::string m_name;
    std::string m_path;
    bool m_force_update = false;
    SceneTables m_tables;
    // what hl_scene_set_tables installed last (create_gpu_resources turns a transform-only change into hl_scene_update_instances)
    SceneTables m_previous_tables;
    bool m_tables_installed = false, m_black_env_installed = false;
    std::vector<hl_mesh> m_installed_meshes;
    std::vector<uint32_t> m_installed_textures;
    std::vector<uint32_t> m_texture_array_ids;
    // state of the device-side copy
    glm::vec3 m_last_sun_direction = glm::vec3(0.0f);
    bool m_sky_valid = false;
    uint32_t m_env_source_id = 0xFFFFFFFFu; // TextureCube id currently uploaded (0xFFFFFFFF = none)
};
} // namespace helios
