// ORACLE — TEST INFRASTRUCTURE ONLY.  glm_pin: every GLM function the path's host code goes through (scene.cpp:186-360,
// 632-646, path_integrator.cpp:136-161), evaluated on seeded random inputs; the result bits go to stdout.  Compiled twice:
// against the GLM the reference vendors (external/AssetCore/external/glm, 0.9.9 @ b3f8772; oracle/Makefile target ref_scene)
// and against the stand-in the C++ host layer ships (helios_b200/shim/include/glm.hpp).  The two outputs must be
// byte-identical (tests/test_ref_scene.py); the reference build's output is committed as tests/golden/glm_pin.bin.
#include <glm.hpp>
#include <gtc/quaternion.hpp>
#include <gtx/matrix_decompose.hpp>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>

static uint32_t g_state = 12345u;
static float    rnd(float lo, float hi)
{
    g_state = g_state * 1664525u + 1013904223u;
    return lo + (hi - lo) * (float(g_state >> 8) / 16777216.0f);
}
static void out(const float* p, int n) { std::fwrite(p, 4, (size_t)n, stdout); }
static void out_mat(const glm::mat4& m)
{
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++)
        {
            const float v = m[c][r];
            out(&v, 1);
        }
}
static void out3(const glm::vec3& v)
{
    const float a[3] = { v.x, v.y, v.z };
    out(a, 3);
}
static void out4(const glm::vec4& v)
{
    const float a[4] = { v.x, v.y, v.z, v.w };
    out(a, 4);
}
static void outq(const glm::quat& q)
{
    const float a[4] = { q.w, q.x, q.y, q.z };
    out(a, 4);
}
static glm::quat random_unit_quat()
{
    float w = rnd(-1, 1), x = rnd(-1, 1), y = rnd(-1, 1), z = rnd(-1, 1);
    const float l = std::sqrt(w * w + x * x + y * y + z * z);
    if (!(l > 1e-3f)) return glm::quat(1, 0, 0, 0);
    return glm::quat(w / l, x / l, y / l, z / l);
}

int main(int argc, char** argv)
{
    const int iterations = argc > 1 ? std::atoi(argv[1]) : 300;
    for (int it = 0; it < iterations; it++)
    {
        const glm::quat q = random_unit_quat(), q2 = random_unit_quat();
        const glm::vec3 p(rnd(-50, 50), rnd(-5, 30), rnd(-50, 50)), sc(rnd(0.2f, 3.0f), rnd(0.2f, 3.0f), rnd(0.2f, 3.0f)), v(rnd(-2, 2), rnd(-2, 2), rnd(-2, 2));
        // TransformNode::update
        const glm::mat4 R = glm::mat4_cast(q), S = glm::scale(glm::mat4(1.0f), sc), T = glm::translate(glm::mat4(1.0f), p);
        const glm::mat4 TR = T * R, M = TR * S;
        out_mat(R), out_mat(S), out_mat(T), out_mat(TR), out_mat(M);
        // forward / up / left, global_position
        out3(q * glm::vec3(0.0f, 0.0f, 1.0f)), out3(q * v), outq(q * q2);
        out4(TR * glm::vec4(v, 1.0f));
        // set_from_global_transform under a parent
        const glm::mat4 parent = glm::translate(glm::mat4(1.0f), glm::vec3(rnd(-5, 5), rnd(-5, 5), rnd(-5, 5))) * glm::mat4_cast(q2);
        const glm::mat4 inv_parent = glm::inverse(parent), local = inv_parent * M;
        out_mat(inv_parent), out_mat(local);
        glm::vec3 ds, dt, dk;
        glm::quat dq;
        glm::vec4 dp;
        const bool ok = glm::decompose(it % 7 == 0 ? local : M, ds, dq, dt, dk, dp);
        const float okf = ok ? 1.0f : 0.0f;
        out(&okf, 1);
        if (ok) out3(ds), outq(dq), out3(dt);
        // negative-determinant basis (mirrored instance)
        if (it % 11 == 0)
        {
            glm::mat4 mir = M;
            mir[0]        = mir[0] * -1.0f;
            if (glm::decompose(mir, ds, dq, dt, dk, dp)) out3(ds), outq(dq), out3(dt);
        }
        // CameraNode::update + PathIntegrator::launch_rays
        const float     fov = rnd(20, 100), aspect = rnd(0.5f, 2.5f), zn = rnd(0.05f, 2.0f), zf = rnd(50, 2000);
        const glm::mat4 proj = glm::perspective(glm::radians(fov), aspect, zn, zf), view = glm::inverse(TR);
        out_mat(proj), out_mat(view), out_mat(glm::inverse(proj * view)), out_mat(glm::transpose(M));
        // light rows / material rows
        const float c = cosf(glm::radians(fov * 0.5f));
        out(&c, 1);
        out3(glm::pow(glm::vec3(rnd(0, 1), rnd(0, 1), rnd(0, 1)), glm::vec3(2.2f)));
        out3(glm::normalize(v)), out3(glm::cross(v, p));
        const float d = glm::dot(v, p), l = glm::length(v);
        out(&d, 1), out(&l, 1);
    }
    return 0;
}
