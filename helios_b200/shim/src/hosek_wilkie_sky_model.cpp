// hosek_wilkie_sky_model.cpp — host half of the Hosek-Wilkie sky (reference:
// src/engine/gfx/hosek_wilkie_sky_model.cpp:41-94 spline / evaluate / radiance, :658-686 update).
// Same arithmetic as helios_b200/sky.py (float where the reference uses float, double for the spline).
#include <gfx/hosek_wilkie_sky_model.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <dlfcn.h>

namespace helios
{
namespace
{
// quintic Bezier over 6 control points `stride` apart
double spline(const double* s, int stride, double v)
{
    const double w[6] = { 1, 5, 10, 10, 5, 1 };
    double       r    = 0.0;
    for (int k = 0; k < 6; k++) r += w[k] * std::pow(1.0 - v, 5 - k) * std::pow(v, k) * s[k * stride];
    return r;
}
double evaluate(const double* ds, int stride, float turbidity, float albedo, float sun_theta)
{
    const float  e  = std::max(0.0f, 1.0f - sun_theta / (3.14159265358979323846f / 2.0f));
    const double k  = std::pow(e, 1.0f / 3.0f); // float pow, as in the reference
    const int    t0 = std::min(std::max((int)turbidity, 1), 10);
    const int    t1 = std::min(t0 + 1, 10);
    const double tk = std::min(std::max(turbidity - (float)t0, 0.0f), 1.0f);
    const double* a0 = ds;
    const double* a1 = ds + stride * 6 * 10;
    const double a0t0 = spline(a0 + stride * 6 * (t0 - 1), stride, k);
    const double a1t0 = spline(a1 + stride * 6 * (t0 - 1), stride, k);
    const double a0t1 = spline(a0 + stride * 6 * (t1 - 1), stride, k);
    const double a1t1 = spline(a1 + stride * 6 * (t1 - 1), stride, k);
    const double al = albedo;
    return a0t0 * (1 - al) * (1 - tk) + a1t0 * al * (1 - tk) + a0t1 * (1 - al) * tk + a1t1 * al * tk;
}
// sky radiance along the sun direction, per channel (hosek_wilkie_sky_model.cpp:69-75)
float hosek(float cos_theta, float gamma, float cos_gamma, const float cf[10][3], int ch)
{
    const float A = cf[0][ch], B = cf[1][ch], C = cf[2][ch], D = cf[3][ch], E = cf[4][ch], F = cf[5][ch], G = cf[6][ch], H = cf[7][ch], I = cf[8][ch];
    const float chi = (1.0f + cos_gamma * cos_gamma) / std::pow(1.0f + H * H - 2.0f * cos_gamma * H, 1.5f);
    return (1.0f + A * std::exp(B / (cos_theta + 0.01f))) * (C + D * std::exp(E * gamma) + F * (cos_gamma * cos_gamma) + G * chi + I * std::sqrt(std::max(0.0f, cos_theta)));
}
std::string data_directory()
{
    if (const char* env = std::getenv("HELIOS_B200_DATA")) return env;
    Dl_info info;
    if (dladdr((const void*)&data_directory, &info) && info.dli_fname)
    {
        std::string p(info.dli_fname);
        const size_t slash = p.find_last_of('/');
        return (slash == std::string::npos ? std::string(".") : p.substr(0, slash)) + "/data";
    }
    return "data";
}
} // namespace

HosekWilkieSkyModel::HosekWilkieSkyModel(vk::Backend::Ptr backend) : m_backend(backend)
{
    for (float& c : m_block) c = 0.0f;
}
HosekWilkieSkyModel::~HosekWilkieSkyModel() {}

void HosekWilkieSkyModel::load_dataset()
{
    if (!m_rgb_dataset.empty()) return;
    const std::string path = data_directory() + "/hosek_rgb_v1_4a.f64";
    FILE*             f    = std::fopen(path.c_str(), "rb");
    if (!f)
    {
        const std::string msg = "HosekWilkieSkyModel: cannot open " + path + " (set HELIOS_B200_DATA)";
        HELIOS_LOG_FATAL(msg);
        throw std::runtime_error(msg);
    }
    m_rgb_dataset.resize(3600);
    const size_t n = std::fread(m_rgb_dataset.data(), sizeof(double), 3600, f);
    std::fclose(f);
    if (n != 3600)
    {
        m_rgb_dataset.clear();
        const std::string msg = "HosekWilkieSkyModel: " + path + " is truncated";
        HELIOS_LOG_FATAL(msg);
        throw std::runtime_error(msg);
    }
}

void HosekWilkieSkyModel::evaluate_coefficients(glm::vec3 direction, float out40[40])
{
    load_dataset();
    const float sun_theta = std::acos(std::min(std::max(direction.y, 0.0f), 1.0f));
    float       cf[10][3];
    for (int i = 0; i < 3; i++)
    {
        const double* rgb = m_rgb_dataset.data() + 1080 * i;
        const double* rad = m_rgb_dataset.data() + 3240 + 120 * i;
        for (int k = 0; k < 7; k++) cf[k][i] = (float)evaluate(rgb + k, 9, m_turbidity, m_ground_albedo, sun_theta);
        // H and I are stored swapped in the dataset (:674-676)
        cf[7][i] = (float)evaluate(rgb + 8, 9, m_turbidity, m_ground_albedo, sun_theta);
        cf[8][i] = (float)evaluate(rgb + 7, 9, m_turbidity, m_ground_albedo, sun_theta);
        cf[9][i] = (float)evaluate(rad, 1, m_turbidity, m_ground_albedo, sun_theta);
    }
    if (m_sun_luminance_target != 0.0f)
    {
        float S[3];
        for (int i = 0; i < 3; i++) S[i] = hosek(std::cos(sun_theta), 0.0f, 1.0f, cf, i) * cf[9][i];
        const float lum = S[0] * 0.2126f + S[1] * 0.7152f + S[2] * 0.0722f;
        for (int i = 0; i < 3; i++) cf[9][i] = cf[9][i] / lum, cf[9][i] = cf[9][i] * m_sun_luminance_target;
    }
    for (int k = 0; k < 10; k++)
    {
        out40[4 * k + 0] = cf[k][0], out40[4 * k + 1] = cf[k][1], out40[4 * k + 2] = cf[k][2], out40[4 * k + 3] = 0.0f;
    }
}

void HosekWilkieSkyModel::update(vk::CommandBuffer::Ptr cmd_buf, glm::vec3 direction)
{
    (void)cmd_buf;
    evaluate_coefficients(direction, m_block);
    auto backend = m_backend.lock();
    if (backend && backend->has_device())
    {
        const float sun[3] = { direction.x, direction.y, direction.z };
        backend->check(hl_sky_update(backend->context(), m_block, sun), "hl_sky_update");
    }
}
} // namespace helios
