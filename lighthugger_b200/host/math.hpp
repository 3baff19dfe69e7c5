// Small column-major float math used by the host-side scene preparation.
// Mirrors what the reference does on the host with glm (lookAt, inverse,
// matrix products: src/main.cpp:735-755, src/shared_cpu_gpu.h:96,
// src/projection.h:3-20).  Host-only: the device and the oracle consume the
// resulting Uniforms/Instance bytes, so no bit-parity constraint applies here.
#pragma once
#include <cmath>
#include <cstdint>

#include "lh_shared_cpu_gpu.h"

namespace lh::host {

struct Vec3 {
    float x = 0, y = 0, z = 0;
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(Vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline float length(Vec3 a) { return std::sqrt(dot(a, a)); }
inline Vec3 normalize(Vec3 a) {   // glm's: v * inversesqrt(dot(v, v)) (the reference host builds its view matrix with glm::lookAt)
    float inv = 1.0f / std::sqrt(dot(a, a));
    return {a.x * inv, a.y * inv, a.z * inv};
}

inline LhMat4 identity4() {
    LhMat4 r{};
    for (int i = 0; i < 4; i++) r.m[i][i] = 1.0f;
    return r;
}

// r = a * b (column vectors: (a*b)*v == a*(b*v))
inline LhMat4 mul(const LhMat4& a, const LhMat4& b) {
    LhMat4 r{};
    for (int c = 0; c < 4; c++)
        for (int row = 0; row < 4; row++) {
            float s = 0.0f;
            for (int k = 0; k < 4; k++) s += a.m[k][row] * b.m[c][k];
            r.m[c][row] = s;
        }
    return r;
}

inline LhMat4 translation(Vec3 t) {
    LhMat4 r = identity4();
    r.m[3][0] = t.x;
    r.m[3][1] = t.y;
    r.m[3][2] = t.z;
    return r;
}

inline LhMat4 scaling(Vec3 s) {
    LhMat4 r = identity4();
    r.m[0][0] = s.x;
    r.m[1][1] = s.y;
    r.m[2][2] = s.z;
    return r;
}

// unit quaternion (x, y, z, w) -> rotation
inline LhMat4 rotation(const float q[4]) {
    float x = q[0], y = q[1], z = q[2], w = q[3];
    LhMat4 r = identity4();
    r.m[0][0] = 1 - 2 * (y * y + z * z);
    r.m[0][1] = 2 * (x * y + z * w);
    r.m[0][2] = 2 * (x * z - y * w);
    r.m[1][0] = 2 * (x * y - z * w);
    r.m[1][1] = 1 - 2 * (x * x + z * z);
    r.m[1][2] = 2 * (y * z + x * w);
    r.m[2][0] = 2 * (x * z + y * w);
    r.m[2][1] = 2 * (y * z - x * w);
    r.m[2][2] = 1 - 2 * (x * x + y * y);
    return r;
}

// General 4x4 inverse by cofactors, evaluated in double and rounded once.
inline LhMat4 inverse(const LhMat4& in) {
    double a[16], inv[16];
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) a[c * 4 + r] = in.m[c][r];
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
    double det = a[0] * inv[0] + a[1] * inv[4] + a[2] * inv[8] + a[3] * inv[12];
    double idet = 1.0 / det;
    LhMat4 out{};
    for (int c = 0; c < 4; c++)
        for (int r = 0; r < 4; r++) out.m[c][r] = (float)(inv[c * 4 + r] * idet);
    return out;
}

// Right-handed look-at, camera looks down -Z (glm::lookAt semantics, src/main.cpp:735-739).
inline LhMat4 look_at(Vec3 eye, Vec3 center, Vec3 up) {
    Vec3 f = normalize(center - eye);
    Vec3 s = normalize(cross(f, up));
    Vec3 u = cross(s, f);
    LhMat4 r = identity4();
    r.m[0][0] = s.x; r.m[1][0] = s.y; r.m[2][0] = s.z;
    r.m[0][1] = u.x; r.m[1][1] = u.y; r.m[2][1] = u.z;
    r.m[0][2] = -f.x; r.m[1][2] = -f.y; r.m[2][2] = -f.z;
    r.m[3][0] = -dot(s, eye);
    r.m[3][1] = -dot(u, eye);
    r.m[3][2] = dot(f, eye);
    return r;
}

// Infinite reverse-Z perspective with flipped Y (src/projection.h:3-20):
// clip.w = -z_view, clip.z = near, depth = near / w.
inline LhMat4 infinite_reverse_z_perspective(float fov_rad, float width, float height, float near_plane) {
    float focal = 1.0f / std::tan(fov_rad / 2.0f);
    LhMat4 r{};
    r.m[0][0] = focal / (width / height);
    r.m[1][1] = -focal;
    r.m[2][3] = -1.0f;
    r.m[3][2] = near_plane;
    return r;
}

// splitmix64, used as a counter-based generator: value k of stream `seed`
// is mix(seed + (k+1) * golden).
inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
inline float u01(uint64_t seed, uint64_t counter) {
    uint64_t v = splitmix64(seed + counter * 0x9E3779B97F4A7C15ull);
    return (float)(v >> 40) * (1.0f / 16777216.0f);  // 24 random bits in [0,1)
}

}  // namespace lh::host
