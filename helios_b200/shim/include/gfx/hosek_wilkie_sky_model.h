// gfx/hosek_wilkie_sky_model.h — HosekWilkieSkyModel (reference: include/gfx/hosek_wilkie_sky_model.h,
// src/engine/gfx/hosek_wilkie_sky_model.cpp:41-94,658-763).  update() does the reference's double-precision
// coefficient fit on the host (turbidity 4, albedo 0.1, sun-direction luminance normalised to 1.15) and hands
// the ten vec4 (A..I, Z) to hl_sky_update, which bakes the 6 x 512^2 RGBA32F cube map on the GPU
// (procedural_sky.frag:48-75).  The RGB dataset (Hosek & Wilkie v1.4a, 3 x 1080 + 3 x 120 doubles) is read
// from helios_b200/data/hosek_rgb_v1_4a.f64; $HELIOS_B200_DATA overrides the directory.
#pragma once
#include <gfx/vk.h>
#include <glm.hpp>
#include <vector>

namespace helios
{
class HosekWilkieSkyModel
{
public:
    HosekWilkieSkyModel(vk::Backend::Ptr backend);
    ~HosekWilkieSkyModel();

    // direction = towards the sun (= -DirectionalLightNode::forward(), scene.cpp:893)
    void update(vk::CommandBuffer::Ptr cmd_buf, glm::vec3 direction);

    // the uniform block of the last update: A,B,C,D,E,F,G,H,I,Z as vec4 (w = 0)
    const float* coefficients() const { return m_block; }
    // host-only evaluation of the fit (no device call); out40 receives the same block
    void evaluate_coefficients(glm::vec3 direction, float out40[40]);

private:
    void load_dataset();

    std::weak_ptr<vk::Backend> m_backend;
    std::vector<double> m_rgb_dataset; // 3600 doubles
    float m_block[40];
    float m_sun_luminance_target = 1.15f;
    float m_ground_albedo = 0.1f;
    float m_turbidity = 4.0f;
};
} // namespace helios
