/*
 * ORACLE — test infrastructure only. Nothing under oracle/ is part of the product; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Minimal vector / matrix helpers for the CPU restatement (GLSL-flavoured names).
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

namespace orc {

struct vec2 {
    float x, y;
};
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
    float &operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator/(vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline vec3 operator/(vec3 a, vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline vec3 &operator+=(vec3 &a, vec3 b) {
    a = a + b;
    return a;
}
inline vec3 &operator*=(vec3 &a, vec3 b) {
    a = a * b;
    return a;
}
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(vec3 a) { return a / length(a); }
inline vec3 vmax(vec3 a, vec3 b) { return {std::max(a.x, b.x), std::max(a.y, b.y), std::max(a.z, b.z)}; }
inline vec3 vmin(vec3 a, vec3 b) { return {std::min(a.x, b.x), std::min(a.y, b.y), std::min(a.z, b.z)}; }
inline vec3 vexp(vec3 a) { return {std::exp(a.x), std::exp(a.y), std::exp(a.z)}; }
inline float clampf(float v, float lo, float hi) { return std::min(std::max(v, lo), hi); }
inline vec3 vclamp(vec3 a, float lo, float hi) { return {clampf(a.x, lo, hi), clampf(a.y, lo, hi), clampf(a.z, lo, hi)}; }
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 vmix(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 reflect(vec3 I, vec3 N) { return I - N * (2.0f * dot(N, I)); }
inline float max3(vec3 v) { return std::max(std::max(v.x, v.y), v.z); }

/* column-major 4x4, m[c*4+r] like glm */
struct mat4 {
    float m[16];
    float at(int r, int c) const { return m[c * 4 + r]; }
};
inline mat4 mat4_from(const float *p) {
    mat4 r;
    std::memcpy(r.m, p, sizeof(r.m));
    return r;
}
inline vec3 xform_point(const mat4 &M, vec3 p) {
    /* fixed evaluation order (no contraction; oracle is built with -ffp-contract=off) so that the
     * world-space triangles are bit-identical to the device flatten kernel */
    return {((M.at(0, 0) * p.x + M.at(0, 1) * p.y) + M.at(0, 2) * p.z) + M.at(0, 3),
            ((M.at(1, 0) * p.x + M.at(1, 1) * p.y) + M.at(1, 2) * p.z) + M.at(1, 3),
            ((M.at(2, 0) * p.x + M.at(2, 1) * p.y) + M.at(2, 2) * p.z) + M.at(2, 3)};
}
inline vec3 xform_dir(const mat4 &M, vec3 d) {
    return {(M.at(0, 0) * d.x + M.at(0, 1) * d.y) + M.at(0, 2) * d.z, (M.at(1, 0) * d.x + M.at(1, 1) * d.y) + M.at(1, 2) * d.z,
            (M.at(2, 0) * d.x + M.at(2, 1) * d.y) + M.at(2, 2) * d.z};
}
/* 3x3 helpers on the upper-left block, row-major float[9] */
struct mat3 {
    float a[9];
    float at(int r, int c) const { return a[r * 3 + c]; }
};
inline mat3 upper3(const mat4 &M) {
    mat3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.a[i * 3 + j] = M.at(i, j);
    return r;
}
inline mat3 inverse3(const mat3 &A) {
    float a = A.at(0, 0), b = A.at(0, 1), c = A.at(0, 2);
    float d = A.at(1, 0), e = A.at(1, 1), f = A.at(1, 2);
    float g = A.at(2, 0), h = A.at(2, 1), i = A.at(2, 2);
    float co00 = e * i - f * h, co01 = -(d * i - f * g), co02 = d * h - e * g;
    float det = a * co00 + b * co01 + c * co02;
    float id = 1.0f / det;
    mat3 r;
    r.a[0] = co00 * id;
    r.a[1] = -(b * i - c * h) * id;
    r.a[2] = (b * f - c * e) * id;
    r.a[3] = co01 * id;
    r.a[4] = (a * i - c * g) * id;
    r.a[5] = -(a * f - c * d) * id;
    r.a[6] = co02 * id;
    r.a[7] = -(a * h - b * g) * id;
    r.a[8] = (a * e - b * d) * id;
    return r;
}
inline vec3 mul3(const mat3 &A, vec3 v) {
    return {A.a[0] * v.x + A.a[1] * v.y + A.a[2] * v.z, A.a[3] * v.x + A.a[4] * v.y + A.a[5] * v.z,
            A.a[6] * v.x + A.a[7] * v.y + A.a[8] * v.z};
}
inline vec3 mul3_transposed(const mat3 &A, vec3 v) {
    return {A.a[0] * v.x + A.a[3] * v.y + A.a[6] * v.z, A.a[1] * v.x + A.a[4] * v.y + A.a[7] * v.z,
            A.a[2] * v.x + A.a[5] * v.y + A.a[8] * v.z};
}

}  // namespace orc
