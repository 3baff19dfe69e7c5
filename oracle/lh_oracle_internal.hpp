// Internals shared by the oracle translation units.  TEST INFRASTRUCTURE
// (see lh_oracle.h).  Holds the pinned vector arithmetic and the pinned
// rasterisation model:
//
//   * clip against w - z >= 0 only (for the infinite reverse-Z projection the
//     Vulkan clip volume 0 <= z_c <= w_c reduces to w >= near; x/y are handled
//     by the scissor), intersections computed from the inside vertex towards
//     the outside one: t = d_in / (d_in - d_out), P = I + t * (O - I);
//   * ndc = clip.xy / clip.w, depth = clip.z / clip.w,
//     screen = ndc * (size/2) + size/2, snapped to 8 sub-pixel bits with
//     round-half-even and clamped to +-2^29;
//   * signed doubled area A2 = (x1-x0)(y2-y0) - (y1-y0)(x2-x0) on the snapped
//     integers; Vulkan's area is -A2/2, so front-facing (CCW front) <=> A2 < 0;
//   * samples at pixel centres, 64-bit integer edge functions, top-left rule:
//     an edge with direction d owns its boundary iff d.y < 0 or
//     (d.y == 0 and d.x > 0) once the triangle is oriented A2 > 0;
//   * depth interpolated linearly in screen space with one reciprocal per
//     triangle: z = z0 + (float(w1) * (z1 - z0) + float(w2) * (z2 - z0)) * (1 / float(A2)).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <utility>

#include "lh_oracle.h"

namespace lho {

struct V3 {
    float x, y, z;
};
struct V4 {
    float x, y, z, w;
};

inline V4 mat4_mul(const LhMat4& m, V4 v) {
    V4 r;
    r.x = (m.m[0][0] * v.x + m.m[1][0] * v.y) + (m.m[2][0] * v.z + m.m[3][0] * v.w);
    r.y = (m.m[0][1] * v.x + m.m[1][1] * v.y) + (m.m[2][1] * v.z + m.m[3][1] * v.w);
    r.z = (m.m[0][2] * v.x + m.m[1][2] * v.y) + (m.m[2][2] * v.z + m.m[3][2] * v.w);
    r.w = (m.m[0][3] * v.x + m.m[1][3] * v.y) + (m.m[2][3] * v.z + m.m[3][3] * v.w);
    return r;
}
inline LhMat4 mat4_mat4(const LhMat4& a, const LhMat4& b) {
    LhMat4 r;
    for (int c = 0; c < 4; c++) {
        V4 col = mat4_mul(a, V4{b.m[c][0], b.m[c][1], b.m[c][2], b.m[c][3]});
        r.m[c][0] = col.x; r.m[c][1] = col.y; r.m[c][2] = col.z; r.m[c][3] = col.w;
    }
    return r;
}
inline V3 mat3_mul(const LhMat3& m, V3 v) {
    V3 r;
    r.x = (m.m[0][0] * v.x + m.m[1][0] * v.y) + m.m[2][0] * v.z;
    r.y = (m.m[0][1] * v.x + m.m[1][1] * v.y) + m.m[2][1] * v.z;
    r.z = (m.m[0][2] * v.x + m.m[1][2] * v.y) + m.m[2][2] * v.z;
    return r;
}
inline float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float length3(V3 a) { return sqrtf(dot3(a, a)); }
// glm's normalize (detail/func_geometric.inl:82-91): v * inversesqrt(dot(v, v)) with inversesqrt(x) = 1 / sqrt(x)
inline V3 normalize3(V3 a) {
    float inv = 1.0f / sqrtf(dot3(a, a));
    return V3{a.x * inv, a.y * inv, a.z * inv};
}
inline V3 cross3(V3 a, V3 b) {
    return V3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y};
}

enum { CULL_NONE = 0, CULL_BACK = 1, CULL_FRONT = 2 };

// ---- shared helpers implemented in lh_oracle.cpp ----
unsigned thread_count(const LhoConfig* cfg);
void parallel_for(uint64_t count, unsigned threads, const std::function<void(uint64_t, uint64_t)>& body);
uint64_t* sector_counter(const LhUniforms* u, const LhoConfig* cfg, uint32_t cascade);
LhPrefixSumValue* sector_values(const LhUniforms* u, const LhoConfig* cfg, uint32_t cascade);
void transform_meshlet_vertices(const LhInstance* inst, const LhMeshInfo* mi, const LhMeshlet* m, const LhMat4& matrix,
                                V4* clip_out, V3* world_out);

// ---- textures (lh_oracle_resolve.cpp) ----
// bilinear, level `lod` (clamped), repeat addressing; returns linear RGBA
void sample_texture_lod(const LhoTexture& t, float u, float v, float lod, float out[4]);

// textureGrad through repeat_sampler (trilinear, anisotropy 16), textureLod(samplerCube, ., 0), and one tap of
// sampler2DArrayShadow + shadowmap_comparison_sampler; also what the reference's GLSL is given as its texture unit
// when it runs on the host (oracle/ref_shim/)
void sample_texture_grad(const LhoTexture& t, const float uv[2], const float dx[2], const float dy[2], float out[4]);
void sample_skybox(const LhoConfig* cfg, V3 d, float out[3]);
float sample_shadow(const float* layer, uint32_t n, float u, float v, float ref);
void sample_lut(const float* lut_rgb, uint32_t n, const float uvw[3], float out[3]);   // display_transform_lut + clamp_sampler

struct AlphaTest {
    bool enabled;
    const LhoTexture* texture;
};
inline AlphaTest make_alpha_test(const LhoConfig* cfg, const LhMeshInfo* mi, bool alpha) {
    AlphaTest a{false, nullptr};
    if (alpha && cfg->textures && mi->base_color_texture_index < cfg->num_textures) {
        a.enabled = true;
        a.texture = &cfg->textures[mi->base_color_texture_index];
    }
    return a;
}
// uv = float(u16) * texture_scale + texture_offset, rasterization.glsl:78-81
inline void load_triangle_uvs(const LhMeshInfo* mi, const LhMeshlet* m, uint32_t i0, uint32_t i1, uint32_t i2, float uv[3][2]) {
    const uint16_t* uvs = (const uint16_t*)(uintptr_t)mi->uvs;
    bool wide = (mi->flags & LH_MESH_INFO_FLAGS_32_BIT_INDICES) != 0;
    uint32_t loc[3] = {i0, i1, i2};
    for (int k = 0; k < 3; k++) {
        uint32_t idx = wide ? ((const uint32_t*)(uintptr_t)mi->indices)[m->index_offset + loc[k]]
                            : ((const uint16_t*)(uintptr_t)mi->indices)[m->index_offset + loc[k]];
        uv[k][0] = (float)uvs[2 * idx + 0] * mi->texture_scale[0] + mi->texture_offset[0];
        uv[k][1] = (float)uvs[2 * idx + 1] * mi->texture_scale[1] + mi->texture_offset[1];
    }
}

// ---- pinned rasteriser ----

struct ClipVert {
    float x, y, z, w, u, v;
};
struct SnapVert {
    int32_t x, y;
    float z, w, u, v;
};

inline bool project_snap(const ClipVert& c, float half_w, float half_h, SnapVert* out) {
    float ndc_x = c.x / c.w, ndc_y = c.y / c.w, z = c.z / c.w;
    float sx = ndc_x * half_w + half_w;
    float sy = ndc_y * half_h + half_h;
    float fx = sx * 256.0f, fy = sy * 256.0f;
    if (!(std::isfinite(fx) && std::isfinite(fy) && std::isfinite(z))) return false;
    const float lim = 536870912.0f;  // 2^29
    fx = fminf(fmaxf(fx, -lim), lim);
    fy = fminf(fmaxf(fy, -lim), lim);
    out->x = (int32_t)nearbyintf(fx);
    out->y = (int32_t)nearbyintf(fy);
    out->z = z;
    out->w = c.w;
    out->u = c.u;
    out->v = c.v;
    return true;
}

inline bool top_left(int64_t dx, int64_t dy) { return dy < 0 || (dy == 0 && dx > 0); }

template <class Emit>
void raster_snapped(SnapVert a, SnapVert b, SnapVert c, int cull, uint32_t width, uint32_t height, const AlphaTest& at,
                    Emit&& emit) {
    int64_t A2 = (int64_t)(b.x - a.x) * (c.y - a.y) - (int64_t)(b.y - a.y) * (c.x - a.x);
    if (A2 == 0) return;
    bool front = A2 < 0;
    if (cull == CULL_BACK && !front) return;
    if (cull == CULL_FRONT && front) return;
    if (A2 < 0) {
        std::swap(b, c);
        A2 = -A2;
    }
    int32_t minx = std::min(a.x, std::min(b.x, c.x)), maxx = std::max(a.x, std::max(b.x, c.x));
    int32_t miny = std::min(a.y, std::min(b.y, c.y)), maxy = std::max(a.y, std::max(b.y, c.y));
    int32_t px0 = std::max<int32_t>(0, (minx - 128 + 255) >> 8), px1 = std::min<int32_t>((int32_t)width - 1, (maxx - 128) >> 8);
    int32_t py0 = std::max<int32_t>(0, (miny - 128 + 255) >> 8), py1 = std::min<int32_t>((int32_t)height - 1, (maxy - 128) >> 8);
    if (px0 > px1 || py0 > py1) return;
    int64_t e0x = (int64_t)c.x - b.x, e0y = (int64_t)c.y - b.y;  // edge b->c, weight of a
    int64_t e1x = (int64_t)a.x - c.x, e1y = (int64_t)a.y - c.y;  // edge c->a, weight of b
    int64_t e2x = (int64_t)b.x - a.x, e2y = (int64_t)b.y - a.y;  // edge a->b, weight of c
    int64_t bias0 = top_left(e0x, e0y) ? 0 : -1;
    int64_t bias1 = top_left(e1x, e1y) ? 0 : -1;
    int64_t bias2 = top_left(e2x, e2y) ? 0 : -1;
    float fA = (float)A2;
    float inv_area = 1.0f / fA;
    float dz1 = b.z - a.z, dz2 = c.z - a.z;
    for (int32_t py = py0; py <= py1; py++) {
        int64_t cy = (int64_t)py * 256 + 128;
        for (int32_t px = px0; px <= px1; px++) {
            int64_t cx = (int64_t)px * 256 + 128;
            int64_t w0 = e0x * (cy - b.y) - e0y * (cx - b.x);
            int64_t w1 = e1x * (cy - c.y) - e1y * (cx - c.x);
            int64_t w2 = e2x * (cy - a.y) - e2y * (cx - a.x);
            if (((w0 + bias0) | (w1 + bias1) | (w2 + bias2)) < 0) continue;
            if (at.enabled) {
                float b1 = (float)w1 / fA, b2 = (float)w2 / fA;
                // perspective-correct uv, then the alpha test of
                // visbuffer_alpha_clip_pixel (rasterization.glsl:90-97), sampled at mip 0
                float b0 = (float)w0 / fA;
                float q0 = b0 / a.w, q1 = b1 / b.w, q2 = b2 / c.w;
                float den = (q0 + q1) + q2;
                float tu = ((q0 * a.u + q1 * b.u) + q2 * c.u) / den;
                float tv = ((q0 * a.v + q1 * b.v) + q2 * c.v) / den;
                float rgba[4];
                sample_texture_lod(*at.texture, tu, tv, 0.0f, rgba);
                if (rgba[3] < 0.5f) continue;
            }
            float z = a.z + ((float)w1 * dz1 + (float)w2 * dz2) * inv_area;
            emit(px, py, z);
        }
    }
}

// tri: clip-space vertices in index order; uv per vertex (used only when at.enabled)
template <class Emit>
void raster_clip_triangle(const V4 tri[3], const float uv[3][2], int cull, uint32_t width, uint32_t height, bool clip_near,
                          const AlphaTest& at, Emit&& emit) {
    ClipVert v[3];
    for (int k = 0; k < 3; k++) v[k] = ClipVert{tri[k].x, tri[k].y, tri[k].z, tri[k].w, uv[k][0], uv[k][1]};
    float half_w = (float)width * 0.5f, half_h = (float)height * 0.5f;
    bool in[3] = {true, true, true};
    float d[3] = {0, 0, 0};
    int n_in = 3;
    if (clip_near) {
        n_in = 0;
        for (int k = 0; k < 3; k++) {
            d[k] = v[k].w - v[k].z;
            in[k] = d[k] >= 0.0f;
            n_in += in[k] ? 1 : 0;
        }
    }
    if (n_in == 0) return;
    if (n_in == 3) {
        SnapVert s[3];
        for (int k = 0; k < 3; k++)
            if (!project_snap(v[k], half_w, half_h, &s[k])) return;
        raster_snapped(s[0], s[1], s[2], cull, width, height, at, emit);
        return;
    }
    ClipVert poly[4];
    int n = 0;
    for (int e = 0; e < 3; e++) {
        int ia = e, ib = (e + 1) % 3;
        if (in[ia]) poly[n++] = v[ia];
        if (in[ia] != in[ib]) {
            int ii = in[ia] ? ia : ib, io = in[ia] ? ib : ia;
            float t = d[ii] / (d[ii] - d[io]);
            const ClipVert& I = v[ii];
            const ClipVert& O = v[io];
            poly[n++] = ClipVert{I.x + t * (O.x - I.x), I.y + t * (O.y - I.y), I.z + t * (O.z - I.z),
                                 I.w + t * (O.w - I.w), I.u + t * (O.u - I.u), I.v + t * (O.v - I.v)};
        }
    }
    SnapVert s[4];
    for (int k = 0; k < n; k++)
        if (!project_snap(poly[k], half_w, half_h, &s[k])) return;
    raster_snapped(s[0], s[1], s[2], cull, width, height, at, emit);
    if (n == 4) raster_snapped(s[0], s[2], s[3], cull, width, height, at, emit);
}

}  // namespace lho
