// ORACLE — TEST INFRASTRUCTURE ONLY.  The host (C++) half of the reference's Hosek-Wilkie sky model — the spline
// evaluation and the coefficient fit of HosekWilkieSkyModel::update (gfx/hosek_wilkie_sky_model.cpp:41-75, 662-686)
// — compiled from the reference's own text (cut out by gen.py at build time) against the reference's own
// dataset header and the glm it vendors.  Pins or_sky_coeffs() and helios_b200/data/hosek_rgb_v1_4a.f64.
#include <glm.hpp>
#include <algorithm>
#define _USE_MATH_DEFINES
#include <math.h>
#include <cstddef>

#include <gfx/hosek_data_rgb.inl>

namespace helios
{
#include "hosek_host_functions.inc"

// the data members update() works on (include/gfx/hosek_wilkie_sky_model.h:32-36 of the reference)
struct SkyHost
{
    float     m_normalized_sun_y, m_albedo, m_turbidity;
    glm::vec3 A, B, C, D, E, F, G, H, I;
    glm::vec3 Z;
    void      update_coefficients(glm::vec3 direction)
    {
#include "hosek_host_update.inc"
    }
};
} // namespace helios

// out40 = the HosekWilkieUBO: A..I, Z as vec4 with w = 0 (hosek_wilkie_sky_model.cpp:688-699)
extern "C" void ref_sky_coeffs(const float* direction, float turbidity, float albedo, float normalized_sun_y, float* out40)
{
    helios::SkyHost h;
    h.m_normalized_sun_y = normalized_sun_y, h.m_albedo = albedo, h.m_turbidity = turbidity;
    h.update_coefficients(glm::vec3(direction[0], direction[1], direction[2]));
    const glm::vec3* v[10] = { &h.A, &h.B, &h.C, &h.D, &h.E, &h.F, &h.G, &h.H, &h.I, &h.Z };
    for (int k = 0; k < 10; k++) out40[4 * k] = v[k]->x, out40[4 * k + 1] = v[k]->y, out40[4 * k + 2] = v[k]->z, out40[4 * k + 3] = 0.0f;
}
