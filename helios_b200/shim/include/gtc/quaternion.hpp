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
// glm::decompose (gtx/matrix_decompose.inl of GLM 0.9.9, the version the reference vendors; itself after WebCore's
// TransformationMatrix) with its operation order: normalise by m[3][3], Gram-Schmidt over the three basis rows with the
// shear factors, flip check, then the quaternion straight from the orthonormal rows.  TransformNode::set_from_*_transform
// (scene.cpp:298-323) feeds every mesh node's matrix through it, so the instance table only comes out bit for bit with the
// same arithmetic.  The perspective partition is solved as GLM does when the bottom row is not (0, 0, 0, w).
inline bool decompose(const mat4& model, vec3& scale, quat& orientation, vec3& translation, vec3& skew, vec4& perspective)
{
    const float eps = 1.1920928955078125e-7f; // epsilon<float>()
    mat4        L   = model;
    if (std::fabs(L[3][3]) < eps) return false;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) L[i][j] /= L[3][3];
    mat4 P = L;
    for (int i = 0; i < 3; i++) P[i][3] = 0.0f;
    P[3][3] = 1.0f;
    if (std::fabs(determinant(P)) < eps) return false;
    if (!(std::fabs(L[0][3]) < eps) || !(std::fabs(L[1][3]) < eps) || !(std::fabs(L[2][3]) < eps))
    {
        const vec4 rhs(L[0][3], L[1][3], L[2][3], L[3][3]);
        perspective = transpose(inverse(P)) * rhs;
        L[0][3] = L[1][3] = L[2][3] = 0.0f;
        L[3][3]                     = 1.0f;
    }
    else
        perspective = vec4(0.0f, 0.0f, 0.0f, 1.0f);
    translation = vec3(L[3][0], L[3][1], L[3][2]);
    vec3 row[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) row[i][j] = L[i][j];
    auto unit    = [](const vec3& v) { return v * 1.0f / length(v); };                                     // detail::scale(v, 1)
    auto combine = [](const vec3& a, const vec3& b, float as, float bs) { return (a * as) + (b * bs); }; // detail::combine
    scale.x = length(row[0]);
    row[0]  = unit(row[0]);
    skew.z  = dot(row[0], row[1]);
    row[1]  = combine(row[1], row[0], 1.0f, -skew.z);
    scale.y = length(row[1]);
    row[1]  = unit(row[1]);
    skew.z /= scale.y;
    skew.y = dot(row[0], row[2]);
    row[2] = combine(row[2], row[0], 1.0f, -skew.y);
    skew.x = dot(row[1], row[2]);
    row[2] = combine(row[2], row[1], 1.0f, -skew.x);
    scale.z = length(row[2]);
    row[2]  = unit(row[2]);
    skew.y /= scale.z;
    skew.x /= scale.z;
    if (dot(row[0], cross(row[1], row[2])) < 0.0f)
        for (int i = 0; i < 3; i++) scale[i] *= -1.0f, row[i] = row[i] * -1.0f;
    float       o[4]; // x, y, z, w
    const float trace = row[0].x + row[1].y + row[2].z;
    if (trace > 0.0f)
    {
        float root = std::sqrt(trace + 1.0f);
        o[3]       = 0.5f * root;
        root       = 0.5f / root;
        o[0] = root * (row[1].z - row[2].y), o[1] = root * (row[2].x - row[0].z), o[2] = root * (row[0].y - row[1].x);
    }
    else
    {
        static const int next[3] = { 1, 2, 0 };
        int              i       = 0;
        if (row[1].y > row[0].x) i = 1;
        if (row[2].z > row[i][i]) i = 2;
        const int j = next[i], k = next[j];
        float     root = std::sqrt(row[i][i] - row[j][j] - row[k][k] + 1.0f);
        o[i]           = 0.5f * root;
        root           = 0.5f / root;
        o[j] = root * (row[i][j] + row[j][i]), o[k] = root * (row[i][k] + row[k][i]), o[3] = root * (row[j][k] - row[k][j]);
    }
    orientation = quat(o[3], o[0], o[1], o[2]);
    return true;
}
} // namespace glm
