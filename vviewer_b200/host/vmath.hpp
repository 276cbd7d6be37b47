/*
 * Minimal linear algebra for the host-side scene description (replaces the glm subset the reference's
 * scene model uses: src/lib/vengine/math/Transform.cpp, src/lib/vengine/core/Camera.cpp).
 * Conventions are glm's: column-major mat4 (m[col][row]), right-handed, depth range [0,1]
 * (GLM_FORCE_DEPTH_ZERO_TO_ONE, Camera.hpp:4-5), quaternion (w,x,y,z), angles in radians.
 */
#pragma once
#include <cmath>
#include <cstring>
#include <algorithm>

namespace vengine {
namespace vm {

struct vec2 {
    float x = 0, y = 0;
    vec2() {}
    vec2(float a, float b) : x(a), y(b) {}
};
struct vec3 {
    float x = 0, y = 0, z = 0;
    vec3() {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct vec4 {
    float x = 0, y = 0, z = 0, w = 0;
    vec4() {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    vec4(vec3 v, float d) : x(v.x), y(v.y), z(v.z), w(d) {}
    /* glm-style colour aliases */
    float &r() { return x; }
    float &g() { return y; }
    float &b() { return z; }
    float &a() { return w; }
    float &operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return a * s; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator/(vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(vec3 a) { return a / length(a); }
inline vec3 min3(vec3 a, vec3 b) { return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)}; }
inline vec3 max3(vec3 a, vec3 b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }

struct quat {
    float w = 1, x = 0, y = 0, z = 0;
    quat() {}
    quat(float W, float X, float Y, float Z) : w(W), x(X), y(Y), z(Z) {}
    /* glm::quat(vec3 eulerAngles): pitch = x, yaw = y, roll = z */
    explicit quat(vec3 e) {
        vec3 c(std::cos(e.x * 0.5f), std::cos(e.y * 0.5f), std::cos(e.z * 0.5f));
        vec3 s(std::sin(e.x * 0.5f), std::sin(e.y * 0.5f), std::sin(e.z * 0.5f));
        w = c.x * c.y * c.z + s.x * s.y * s.z;
        x = s.x * c.y * c.z - c.x * s.y * s.z;
        y = c.x * s.y * c.z + s.x * c.y * s.z;
        z = c.x * c.y * s.z - s.x * s.y * c.z;
    }
};
inline quat operator*(quat p, quat q) {
    return quat(p.w * q.w - p.x * q.x - p.y * q.y - p.z * q.z, p.w * q.x + p.x * q.w + p.y * q.z - p.z * q.y,
                p.w * q.y + p.y * q.w + p.z * q.x - p.x * q.z, p.w * q.z + p.z * q.w + p.x * q.y - p.y * q.x);
}
/* glm::rotate(quat, vec3) == q * v */
inline vec3 rotate(quat q, vec3 v) {
    vec3 qv(q.x, q.y, q.z);
    vec3 uv = cross(qv, v);
    vec3 uuv = cross(qv, uv);
    return v + ((uv * q.w) + uuv) * 2.0f;
}
inline quat angleAxis(float angle, vec3 axis) {
    float s = std::sin(angle * 0.5f);
    return quat(std::cos(angle * 0.5f), axis.x * s, axis.y * s, axis.z * s);
}

struct mat4 {
    float m[4][4]; /* m[col][row] */
    mat4() { std::memset(m, 0, sizeof(m)); }
    explicit mat4(float d) {
        std::memset(m, 0, sizeof(m));
        m[0][0] = m[1][1] = m[2][2] = m[3][3] = d;
    }
    float *operator[](int c) { return m[c]; }
    const float *operator[](int c) const { return m[c]; }
    const float *data() const { return &m[0][0]; }
};
inline mat4 operator*(const mat4 &a, const mat4 &b) {
    mat4 r;
    for (int c = 0; c < 4; c++)
        for (int rr = 0; rr < 4; rr++) {
            float s = 0;
            for (int k = 0; k < 4; k++) s += a.m[k][rr] * b.m[c][k];
            r.m[c][rr] = s;
        }
    return r;
}
inline vec4 operator*(const mat4 &a, vec4 v) {
    vec4 r;
    for (int rr = 0; rr < 4; rr++) r[rr] = a.m[0][rr] * v.x + a.m[1][rr] * v.y + a.m[2][rr] * v.z + a.m[3][rr] * v.w;
    return r;
}
inline mat4 translate(vec3 t) {
    mat4 r(1.0f);
    r.m[3][0] = t.x;
    r.m[3][1] = t.y;
    r.m[3][2] = t.z;
    return r;
}
inline mat4 scale(vec3 s) {
    mat4 r(1.0f);
    r.m[0][0] = s.x;
    r.m[1][1] = s.y;
    r.m[2][2] = s.z;
    return r;
}
inline mat4 toMat4(quat q) {
    mat4 r(1.0f);
    float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z, qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z, qwx = q.w * q.x,
          qwy = q.w * q.y, qwz = q.w * q.z;
    r.m[0][0] = 1 - 2 * (qyy + qzz);
    r.m[0][1] = 2 * (qxy + qwz);
    r.m[0][2] = 2 * (qxz - qwy);
    r.m[1][0] = 2 * (qxy - qwz);
    r.m[1][1] = 1 - 2 * (qxx + qzz);
    r.m[1][2] = 2 * (qyz + qwx);
    r.m[2][0] = 2 * (qxz + qwy);
    r.m[2][1] = 2 * (qyz - qwx);
    r.m[2][2] = 1 - 2 * (qxx + qyy);
    return r;
}
/* glm::quat_cast of a pure rotation matrix */
inline quat quat_cast(const mat4 &m) {
    float fourXSquaredMinus1 = m.m[0][0] - m.m[1][1] - m.m[2][2];
    float fourYSquaredMinus1 = m.m[1][1] - m.m[0][0] - m.m[2][2];
    float fourZSquaredMinus1 = m.m[2][2] - m.m[0][0] - m.m[1][1];
    float fourWSquaredMinus1 = m.m[0][0] + m.m[1][1] + m.m[2][2];
    int biggestIndex = 0;
    float big = fourWSquaredMinus1;
    if (fourXSquaredMinus1 > big) { big = fourXSquaredMinus1; biggestIndex = 1; }
    if (fourYSquaredMinus1 > big) { big = fourYSquaredMinus1; biggestIndex = 2; }
    if (fourZSquaredMinus1 > big) { big = fourZSquaredMinus1; biggestIndex = 3; }
    float biggestVal = std::sqrt(big + 1.0f) * 0.5f;
    float mult = 0.25f / biggestVal;
    switch (biggestIndex) {
        case 0: return quat(biggestVal, (m.m[1][2] - m.m[2][1]) * mult, (m.m[2][0] - m.m[0][2]) * mult, (m.m[0][1] - m.m[1][0]) * mult);
        case 1: return quat((m.m[1][2] - m.m[2][1]) * mult, biggestVal, (m.m[0][1] + m.m[1][0]) * mult, (m.m[2][0] + m.m[0][2]) * mult);
        case 2: return quat((m.m[2][0] - m.m[0][2]) * mult, (m.m[0][1] + m.m[1][0]) * mult, biggestVal, (m.m[1][2] + m.m[2][1]) * mult);
        default: return quat((m.m[0][1] - m.m[1][0]) * mult, (m.m[2][0] + m.m[0][2]) * mult, (m.m[1][2] + m.m[2][1]) * mult, biggestVal);
    }
}
/* glm::lookAt (RH) */
inline mat4 lookAt(vec3 eye, vec3 center, vec3 up) {
    vec3 f = normalize(center - eye);
    vec3 s = normalize(cross(f, up));
    vec3 u = cross(s, f);
    mat4 r(1.0f);
    r.m[0][0] = s.x; r.m[1][0] = s.y; r.m[2][0] = s.z;
    r.m[0][1] = u.x; r.m[1][1] = u.y; r.m[2][1] = u.z;
    r.m[0][2] = -f.x; r.m[1][2] = -f.y; r.m[2][2] = -f.z;
    r.m[3][0] = -dot(s, eye);
    r.m[3][1] = -dot(u, eye);
    r.m[3][2] = dot(f, eye);
    return r;
}
/* glm::perspective, RH, depth zero-to-one */
inline mat4 perspective(float fovy, float aspect, float zNear, float zFar) {
    float t = std::tan(fovy / 2.0f);
    mat4 r;
    r.m[0][0] = 1.0f / (aspect * t);
    r.m[1][1] = 1.0f / t;
    r.m[2][2] = zFar / (zNear - zFar);
    r.m[2][3] = -1.0f;
    r.m[3][2] = -(zFar * zNear) / (zFar - zNear);
    return r;
}
/* glm::ortho, RH, depth zero-to-one */
inline mat4 ortho(float left, float right, float bottom, float top, float zNear, float zFar) {
    mat4 r(1.0f);
    r.m[0][0] = 2.0f / (right - left);
    r.m[1][1] = 2.0f / (top - bottom);
    r.m[2][2] = -1.0f / (zFar - zNear);
    r.m[3][0] = -(right + left) / (right - left);
    r.m[3][1] = -(top + bottom) / (top - bottom);
    r.m[3][2] = -zNear / (zFar - zNear);
    return r;
}
/* general 4x4 inverse (cofactor expansion, same algebra as glm::inverse) */
inline mat4 inverse(const mat4 &M) {
    const float *a = &M.m[0][0];
    float inv[16];
    inv[0] = a[5] * a[10] * a[15] - a[5] * a[11] * a[14] - a[9] * a[6] * a[15] + a[9] * a[7] * a[14] + a[13] * a[6] * a[11] - a[13] * a[7] * a[10];
    inv[4] = -a[4] * a[10] * a[15] + a[4] * a[11] * a[14] + a[8] * a[6] * a[15] - a[8] * a[7] * a[14] - a[12] * a[6] * a[11] + a[12] * a[7] * a[10];
    inv[8] = a[4] * a[9] * a[15] - a[4] * a[11] * a[13] - a[8] * a[5] * a[15] + a[8] * a[7] * a[13] + a[12] * a[5] * a[11] - a[12] * a[7] * a[9];
    inv[12] = -a[4] * a[9] * a[14] + a[4] * a[10] * a[13] + a[8] * a[5] * a[14] - a[8] * a[6] * a[13] - a[12] * a[5] * a[10] + a[12] * a[6] * a[9];
    inv[1] = -a[1] * a[10] * a[15] + a[1] * a[11] * a[14] + a[9] * a[2] * a[15] - a[9] * a[3] * a[14] - a[13] * a[2] * a[11] + a[13] * a[3] * a[10];
    inv[5] = a[0] * a[10] * a[15] - a[0] * a[11] * a[14] - a[8] * a[2] * a[15] + a[8] * a[3] * a[14] + a[12] * a[2] * a[11] - a[12] * a[3] * a[10];
    inv[9] = -a[0] * a[9] * a[15] + a[0] * a[11] * a[13] + a[8] * a[1] * a[15] - a[8] * a[3] * a[13] - a[12] * a[1] * a[11] + a[12] * a[3] * a[9];
    inv[13] = a[0] * a[9] * a[14] - a[0] * a[10] * a[13] - a[8] * a[1] * a[14] + a[8] * a[2] * a[13] + a[12] * a[1] * a[10] - a[12] * a[2] * a[9];
    inv[2] = a[1] * a[6] * a[15] - a[1] * a[7] * a[14] - a[5] * a[2] * a[15] + a[5] * a[3] * a[14] + a[13] * a[2] * a[7] - a[13] * a[3] * a[6];
    inv[6] = -a[0] * a[6] * a[15] + a[0] * a[7] * a[14] + a[4] * a[2] * a[15] - a[4] * a[3] * a[14] - a[12] * a[2] * a[7] + a[12] * a[3] * a[6];
    inv[10] = a[0] * a[5] * a[15] - a[0] * a[7] * a[13] - a[4] * a[1] * a[15] + a[4] * a[3] * a[13] + a[12] * a[1] * a[7] - a[12] * a[3] * a[5];
    inv[14] = -a[0] * a[5] * a[14] + a[0] * a[6] * a[13] + a[4] * a[1] * a[14] - a[4] * a[2] * a[13] - a[12] * a[1] * a[6] + a[12] * a[2] * a[5];
    inv[3] = -a[1] * a[6] * a[11] + a[1] * a[7] * a[10] + a[5] * a[2] * a[11] - a[5] * a[3] * a[10] - a[9] * a[2] * a[7] + a[9] * a[3] * a[6];
    inv[7] = a[0] * a[6] * a[11] - a[0] * a[7] * a[10] - a[4] * a[2] * a[11] + a[4] * a[3] * a[10] + a[8] * a[2] * a[7] - a[8] * a[3] * a[6];
    inv[11] = -a[0] * a[5] * a[11] + a[0] * a[7] * a[9] + a[4] * a[1] * a[11] - a[4] * a[3] * a[9] - a[8] * a[1] * a[7] + a[8] * a[3] * a[5];
    inv[15] = a[0] * a[5] * a[10] - a[0] * a[6] * a[9] - a[4] * a[1] * a[10] + a[4] * a[2] * a[9] + a[8] * a[1] * a[6] - a[8] * a[2] * a[5];
    float det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    float id = 1.0f / det;
    mat4 r;
    float *o = &r.m[0][0];
    for (int i = 0; i < 16; i++) o[i] = inv[i] * id;
    return r;
}

}  // namespace vm
}  // namespace vengine
