// oracle/ref_shim/glm/glm.hpp -- TEST INFRASTRUCTURE.
//
// Minimal stand-in for the GLM headers that the reference rasterizer includes
// (third-party/diff-gaussian-rasterization-w-depth/.gitmodules:1-3 names GLM as
// an un-vendored submodule; it is absent from /root/reference and from this
// image).  Only what forward.cu / backward.cu / rasterizer_impl.cu use is
// provided: vec3, vec4, column-major mat3, dot/length/transpose/max and the
// arithmetic operators, each evaluated in the order GLM documents
// (dot = x*x' + y*y' + z*z'; (A*B)[j][i] = A[0][i]*B[j][0] + A[1][i]*B[j][1] + A[2][i]*B[j][2]).
// Written from GLM's public interface; no GLM source is reproduced.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define GLM_HD __host__ __device__ inline

namespace glm {

struct vec3 {
    float x, y, z;
    GLM_HD vec3() : x(0.f), y(0.f), z(0.f) {}
    GLM_HD explicit vec3(float s) : x(s), y(s), z(s) {}
    GLM_HD vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    GLM_HD float& operator[](int i) { return (&x)[i]; }
    GLM_HD const float& operator[](int i) const { return (&x)[i]; }
    GLM_HD vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
    GLM_HD vec3& operator+=(float s) { x += s; y += s; z += s; return *this; }
    GLM_HD vec3& operator-=(const vec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
    GLM_HD vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
};

struct vec4 {
    float x, y, z, w;
    GLM_HD vec4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
    GLM_HD vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    GLM_HD float& operator[](int i) { return (&x)[i]; }
    GLM_HD const float& operator[](int i) const { return (&x)[i]; }
};

GLM_HD vec3 operator+(const vec3& a, const vec3& b) { return vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
GLM_HD vec3 operator-(const vec3& a, const vec3& b) { return vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
GLM_HD vec3 operator*(const vec3& a, const vec3& b) { return vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
GLM_HD vec3 operator-(const vec3& a) { return vec3(-a.x, -a.y, -a.z); }
GLM_HD vec3 operator*(const vec3& a, float s) { return vec3(a.x * s, a.y * s, a.z * s); }
GLM_HD vec3 operator*(float s, const vec3& a) { return vec3(s * a.x, s * a.y, s * a.z); }
GLM_HD vec3 operator/(const vec3& a, float s) { return vec3(a.x / s, a.y / s, a.z / s); }
GLM_HD vec3 operator+(const vec3& a, float s) { return vec3(a.x + s, a.y + s, a.z + s); }
GLM_HD vec3 operator-(const vec3& a, float s) { return vec3(a.x - s, a.y - s, a.z - s); }

GLM_HD float dot(const vec3& a, const vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
GLM_HD float dot(const vec4& a, const vec4& b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }
GLM_HD float length(const vec3& a) { return sqrtf(dot(a, a)); }
GLM_HD float length(const vec4& a) { return sqrtf(dot(a, a)); }
GLM_HD vec3 max(const vec3& a, float s) { return vec3(fmaxf(a.x, s), fmaxf(a.y, s), fmaxf(a.z, s)); }

// Column-major 3x3: m[c] is column c, m[c][r] is row r of column c.
struct mat3 {
    vec3 c[3];
    GLM_HD mat3() {}
    GLM_HD explicit mat3(float d) { c[0] = vec3(d, 0.f, 0.f); c[1] = vec3(0.f, d, 0.f); c[2] = vec3(0.f, 0.f, d); }
    GLM_HD mat3(float x0, float y0, float z0, float x1, float y1, float z1, float x2, float y2, float z2)
    { c[0] = vec3(x0, y0, z0); c[1] = vec3(x1, y1, z1); c[2] = vec3(x2, y2, z2); }
    GLM_HD vec3& operator[](int i) { return c[i]; }
    GLM_HD const vec3& operator[](int i) const { return c[i]; }
};

GLM_HD mat3 transpose(const mat3& m)
{
    return mat3(m[0][0], m[1][0], m[2][0], m[0][1], m[1][1], m[2][1], m[0][2], m[1][2], m[2][2]);
}

GLM_HD mat3 operator*(const mat3& a, const mat3& b)
{
    mat3 r;
    for (int j = 0; j < 3; ++j)
        for (int i = 0; i < 3; ++i)
            r[j][i] = a[0][i] * b[j][0] + a[1][i] * b[j][1] + a[2][i] * b[j][2];
    return r;
}
GLM_HD vec3 operator*(const mat3& m, const vec3& v)
{
    return vec3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z,
                m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
                m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
}
GLM_HD mat3 operator*(float s, const mat3& m) { mat3 r; r[0] = s * m[0]; r[1] = s * m[1]; r[2] = s * m[2]; return r; }
GLM_HD mat3 operator*(const mat3& m, float s) { mat3 r; r[0] = m[0] * s; r[1] = m[1] * s; r[2] = m[2] * s; return r; }

}  // namespace glm
