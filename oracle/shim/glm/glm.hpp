// Stand-in for the subset of GLM that /root/reference/src uses (TEST INFRASTRUCTURE ONLY; see Eigen/Dense shim).
// Column-major mat4 (m[col][row]) exactly as GLM; formulas are the textbook closed forms GLM documents.
#ifndef RT795_ORACLE_SHIM_GLM
#define RT795_ORACLE_SHIM_GLM
#include <cmath>
#include <initializer_list>

namespace glm {

struct vec4;
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    vec3(float a, float b, float c) : x(a), y(b), z(c) {}
    vec3(const vec4& v);  // implicit truncation (src/Helper.cpp:94-95)
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
    vec3 operator*(float s) const { return vec3(x * s, y * s, z * s); }
};
struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    vec4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    float& operator[](int i) { return (&x)[i]; }
    const float& operator[](int i) const { return (&x)[i]; }
};
inline vec3::vec3(const vec4& v) : x(v.x), y(v.y), z(v.z) {}

struct mat4 {
    vec4 c[4];  // columns
    mat4() { for (int i = 0; i < 4; i++) c[i] = vec4(0, 0, 0, 0); }
    explicit mat4(double d) { for (int i = 0; i < 4; i++) { c[i] = vec4(0, 0, 0, 0); c[i][i] = (float)d; } }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
    vec4 operator*(const vec4& v) const {
        vec4 r;
        for (int row = 0; row < 4; row++) r[row] = c[0][row] * v.x + c[1][row] * v.y + c[2][row] * v.z + c[3][row] * v.w;
        return r;
    }
    mat4 operator*(const mat4& o) const {
        mat4 r;
        for (int col = 0; col < 4; col++) r.c[col] = (*this) * o.c[col];
        return r;
    }
};

inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }

inline mat4 translate(const mat4& m, const vec3& v) {
    mat4 r = m;
    for (int row = 0; row < 4; row++) r.c[3][row] = m.c[0][row] * v.x + m.c[1][row] * v.y + m.c[2][row] * v.z + m.c[3][row];
    return r;
}
inline mat4 scale(const mat4& m, const vec3& v) {
    mat4 r;
    for (int row = 0; row < 4; row++) {
        r.c[0][row] = m.c[0][row] * v.x; r.c[1][row] = m.c[1][row] * v.y;
        r.c[2][row] = m.c[2][row] * v.z; r.c[3][row] = m.c[3][row];
    }
    return r;
}
inline mat4 rotate(const mat4& m, float angle, const vec3& axisIn) {
    float co = std::cos(angle), si = std::sin(angle);
    float len = std::sqrt(axisIn.x * axisIn.x + axisIn.y * axisIn.y + axisIn.z * axisIn.z);
    vec3 a(axisIn.x / len, axisIn.y / len, axisIn.z / len);
    vec3 t(a.x * (1 - co), a.y * (1 - co), a.z * (1 - co));
    float R[3][3];  // R[col][row]
    R[0][0] = co + t.x * a.x;        R[0][1] = t.x * a.y + si * a.z;  R[0][2] = t.x * a.z - si * a.y;
    R[1][0] = t.y * a.x - si * a.z;  R[1][1] = co + t.y * a.y;        R[1][2] = t.y * a.z + si * a.x;
    R[2][0] = t.z * a.x + si * a.y;  R[2][1] = t.z * a.y - si * a.x;  R[2][2] = co + t.z * a.z;
    mat4 r;
    for (int col = 0; col < 3; col++)
        for (int row = 0; row < 4; row++)
            r.c[col][row] = m.c[0][row] * R[col][0] + m.c[1][row] * R[col][1] + m.c[2][row] * R[col][2];
    r.c[3] = m.c[3];
    return r;
}
inline mat4 inverse(const mat4& M) {
    // general 4x4 inverse by cofactors
    float a[16], inv[16];
    for (int col = 0; col < 4; col++) for (int row = 0; row < 4; row++) a[col * 4 + row] = M.c[col][row];
    inv[0] = a[5]*a[10]*a[15] - a[5]*a[11]*a[14] - a[9]*a[6]*a[15] + a[9]*a[7]*a[14] + a[13]*a[6]*a[11] - a[13]*a[7]*a[10];
    inv[4] = -a[4]*a[10]*a[15] + a[4]*a[11]*a[14] + a[8]*a[6]*a[15] - a[8]*a[7]*a[14] - a[12]*a[6]*a[11] + a[12]*a[7]*a[10];
    inv[8] = a[4]*a[9]*a[15] - a[4]*a[11]*a[13] - a[8]*a[5]*a[15] + a[8]*a[7]*a[13] + a[12]*a[5]*a[11] - a[12]*a[7]*a[9];
    inv[12] = -a[4]*a[9]*a[14] + a[4]*a[10]*a[13] + a[8]*a[5]*a[14] - a[8]*a[6]*a[13] - a[12]*a[5]*a[10] + a[12]*a[6]*a[9];
    inv[1] = -a[1]*a[10]*a[15] + a[1]*a[11]*a[14] + a[9]*a[2]*a[15] - a[9]*a[3]*a[14] - a[13]*a[2]*a[11] + a[13]*a[3]*a[10];
    inv[5] = a[0]*a[10]*a[15] - a[0]*a[11]*a[14] - a[8]*a[2]*a[15] + a[8]*a[3]*a[14] + a[12]*a[2]*a[11] - a[12]*a[3]*a[10];
    inv[9] = -a[0]*a[9]*a[15] + a[0]*a[11]*a[13] + a[8]*a[1]*a[15] - a[8]*a[3]*a[13] - a[12]*a[1]*a[11] + a[12]*a[3]*a[9];
    inv[13] = a[0]*a[9]*a[14] - a[0]*a[10]*a[13] - a[8]*a[1]*a[14] + a[8]*a[2]*a[13] + a[12]*a[1]*a[10] - a[12]*a[2]*a[9];
    inv[2] = a[1]*a[6]*a[15] - a[1]*a[7]*a[14] - a[5]*a[2]*a[15] + a[5]*a[3]*a[14] + a[13]*a[2]*a[7] - a[13]*a[3]*a[6];
    inv[6] = -a[0]*a[6]*a[15] + a[0]*a[7]*a[14] + a[4]*a[2]*a[15] - a[4]*a[3]*a[14] - a[12]*a[2]*a[7] + a[12]*a[3]*a[6];
    inv[10] = a[0]*a[5]*a[15] - a[0]*a[7]*a[13] - a[4]*a[1]*a[15] + a[4]*a[3]*a[13] + a[12]*a[1]*a[7] - a[12]*a[3]*a[5];
    inv[14] = -a[0]*a[5]*a[14] + a[0]*a[6]*a[13] + a[4]*a[1]*a[14] - a[4]*a[2]*a[13] - a[12]*a[1]*a[6] + a[12]*a[2]*a[5];
    inv[3] = -a[1]*a[6]*a[11] + a[1]*a[7]*a[10] + a[5]*a[2]*a[11] - a[5]*a[3]*a[10] - a[9]*a[2]*a[7] + a[9]*a[3]*a[6];
    inv[7] = a[0]*a[6]*a[11] - a[0]*a[7]*a[10] - a[4]*a[2]*a[11] + a[4]*a[3]*a[10] + a[8]*a[2]*a[7] - a[8]*a[3]*a[6];
    inv[11] = -a[0]*a[5]*a[11] + a[0]*a[7]*a[9] + a[4]*a[1]*a[11] - a[4]*a[3]*a[9] - a[8]*a[1]*a[7] + a[8]*a[3]*a[5];
    inv[15] = a[0]*a[5]*a[10] - a[0]*a[6]*a[9] - a[4]*a[1]*a[10] + a[4]*a[2]*a[9] + a[8]*a[1]*a[6] - a[8]*a[2]*a[5];
    float det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    float id = 1.0f / det;
    mat4 r;
    for (int col = 0; col < 4; col++) for (int row = 0; row < 4; row++) r.c[col][row] = inv[col * 4 + row] * id;
    return r;
}
inline mat4 transpose(const mat4& M) {
    mat4 r;
    for (int col = 0; col < 4; col++) for (int row = 0; row < 4; row++) r.c[col][row] = M.c[row][col];
    return r;
}
inline mat4 inverseTranspose(const mat4& M) { return transpose(inverse(M)); }

}  // namespace glm
#endif
