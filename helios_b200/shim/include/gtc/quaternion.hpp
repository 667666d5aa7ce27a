// gtc/quaternion.hpp — quaternion subset of the GLM stand-in (see ../glm.hpp)
#pragma once
#include <glm.hpp>

namespace glm
{
struct quat
{
    float x, y, z, w;
    quat() : x(0), y(0), z(0), w(1) {}
    quat(float w_, float x_, float y_, float z_) : x(x_), y(y_), z(z_), w(w_) {}
    // from Euler angles (pitch, yaw, roll) in radians
    explicit quat(const vec3& e)
    {
        const vec3 c(std::cos(e.x * 0.5f), std::cos(e.y * 0.5f), std::cos(e.z * 0.5f));
        const vec3 s(std::sin(e.x * 0.5f), std::sin(e.y * 0.5f), std::sin(e.z * 0.5f));
        w = c.x * c.y * c.z + s.x * s.y * s.z;
        x = s.x * c.y * c.z - c.x * s.y * s.z;
        y = c.x * s.y * c.z + s.x * c.y * s.z;
        z = c.x * c.y * s.z - s.x * s.y * c.z;
    }
};
inline quat operator*(const quat& p, const quat& q)
{
    return quat(p.w * q.w - p.x * q.x - p.y * q.y - p.z * q.z, p.w * q.x + p.x * q.w + p.y * q.z - p.z * q.y, p.w * q.y + p.y * q.w + p.z * q.x - p.x * q.z,
                p.w * q.z + p.z * q.w + p.x * q.y - p.y * q.x);
}
inline vec3 operator*(const quat& q, const vec3& v)
{
    const vec3 qv(q.x, q.y, q.z);
    const vec3 uv  = cross(qv, v);
    const vec3 uuv = cross(qv, uv);
    return v + ((uv * q.w) + uuv) * 2.0f;
}
inline quat normalize(const quat& q)
{
    const float l = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    if (l <= 0.0f) return quat();
    const float i = 1.0f / l;
    return quat(q.w * i, q.x * i, q.y * i, q.z * i);
}
inline mat4 mat4_cast(const quat& q)
{
    mat4        r(1.0f);
    const float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z, qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z, qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
    r[0][0] = 1.0f - 2.0f * (qyy + qzz), r[0][1] = 2.0f * (qxy + qwz), r[0][2] = 2.0f * (qxz - qwy);
    r[1][0] = 2.0f * (qxy - qwz), r[1][1] = 1.0f - 2.0f * (qxx + qzz), r[1][2] = 2.0f * (qyz + qwx);
    r[2][0] = 2.0f * (qxz + qwy), r[2][1] = 2.0f * (qyz - qwx), r[2][2] = 1.0f - 2.0f * (qxx + qyy);
    return r;
}
// rotation matrix (orthonormal upper 3x3) -> quaternion, largest-component branch
inline quat quat_cast(const mat4& m)
{
    const float fx = m[0][0] - m[1][1] - m[2][2], fy = m[1][1] - m[0][0] - m[2][2], fz = m[2][2] - m[0][0] - m[1][1], fw = m[0][0] + m[1][1] + m[2][2];
    int         big  = 0;
    float       best = fw;
    if (fx > best) best = fx, big = 1;
    if (fy > best) best = fy, big = 2;
    if (fz > best) best = fz, big = 3;
    const float v = std::sqrt(best + 1.0f) * 0.5f, mult = 0.25f / v;
    switch (big)
    {
        case 0: return quat(v, (m[1][2] - m[2][1]) * mult, (m[2][0] - m[0][2]) * mult, (m[0][1] - m[1][0]) * mult);
        case 1: return quat((m[1][2] - m[2][1]) * mult, v, (m[0][1] + m[1][0]) * mult, (m[2][0] + m[0][2]) * mult);
        case 2: return quat((m[2][0] - m[0][2]) * mult, (m[0][1] + m[1][0]) * mult, v, (m[1][2] + m[2][1]) * mult);
        default: return quat((m[0][1] - m[1][0]) * mult, (m[2][0] + m[0][2]) * mult, (m[1][2] + m[2][1]) * mult, v);
    }
}
// affine TRS decomposition (no skew / perspective): what TransformNode::set_from_*_transform needs
// (scene.cpp:298-323 calls glm::decompose and discards skew and perspective)
inline bool decompose(const mat4& m, vec3& scale, quat& orientation, vec3& translation, vec3& skew, vec4& perspective)
{
    translation = vec3(m[3][0], m[3][1], m[3][2]);
    vec3 c0(m[0][0], m[0][1], m[0][2]), c1(m[1][0], m[1][1], m[1][2]), c2(m[2][0], m[2][1], m[2][2]);
    scale = vec3(length(c0), length(c1), length(c2));
    if (scale.x == 0.0f || scale.y == 0.0f || scale.z == 0.0f) return false;
    if (dot(c0, cross(c1, c2)) < 0.0f) scale = vec3(-scale.x, -scale.y, -scale.z); // mirrored basis
    c0 = c0 / scale.x, c1 = c1 / scale.y, c2 = c2 / scale.z;
    mat4 r(1.0f);
    r[0] = vec4(c0, 0.0f), r[1] = vec4(c1, 0.0f), r[2] = vec4(c2, 0.0f);
    orientation = quat_cast(r);
    skew        = vec3(0.0f);
    perspective = vec4(0.0f, 0.0f, 0.0f, 1.0f);
    return true;
}
} // namespace glm
