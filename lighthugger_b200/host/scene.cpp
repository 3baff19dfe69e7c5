// Host-side scene preparation: procedural meshes -> quantised vertex streams
// -> meshlets (+ bounds) -> MeshInfo / Instance / Uniforms in the layouts of
// include/lh_shared_cpu_gpu.h.  See include/lh_scene.h for the reference
// call sites each piece stands in for.
//
// The meshlet builder here is this project's own (a greedy fill over
// patch-ordered grid triangles, 64 vertices / 124 triangles, output format
// identical to what src/resources/meshlets.cpp:90-139 packs).  The bounds
// (sphere + normal cone) follow the published meshoptimizer v0.20 algorithm
// (meshopt_computeClusterBounds) so that Meshlet.cone_* / bounding_sphere mean
// exactly what the reference's culling shaders expect; tests check them
// bit-for-bit against the reference's own build when oracle/_ref is present.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <vector>

#include "lh_scene.h"
#include "math.hpp"
#ifdef LH_HAVE_MESHOPTIMIZER
#include <meshoptimizer.h>
#endif

using namespace lh::host;

namespace {

constexpr uint32_t kPatchQuads = 7;  // (7+1)^2 = 64 unique vertices, 98 triangles per patch

struct Mesh {
    std::vector<uint16_t> positions;  // x3
    std::vector<int8_t> normals;      // x3
    std::vector<uint16_t> uvs;        // x2
    std::vector<uint16_t> idx16;
    std::vector<uint32_t> idx32;
    std::vector<uint8_t> micro;
    std::vector<LhMeshlet> meshlets;
    std::vector<uint32_t> triangles;  // the triangle list the meshlets were built from (3 vertex indices each)
    uint64_t num_triangles = 0;
    LhMeshInfo info{};
    float dq_scale[3] = {1, 1, 1};
    float dq_offset[3] = {0, 0, 0};
    bool wide_indices = false;
    int shares_geometry_of = -1;      // >= 0: a material variant of that mesh: own MeshInfo, the other mesh's stream buffers
};

struct InstanceRecord {
    int mesh;
    LhMat4 transform;
};

struct FloatMesh {
    std::vector<float> pos, nrm, uv;  // xyz, xyz, uv
    std::vector<uint32_t> tris;       // patch-ordered triangle list
};

// Emits a (segments_u x segments_v) quad grid through `eval`, triangles ordered
// patch by patch so that the greedy meshlet fill produces compact meshlets.
void emit_grid(FloatMesh& fm, uint32_t su, uint32_t sv, bool wrap_u, bool wrap_v, bool flip,
               const std::function<void(float, float, float*, float*)>& eval) {
    uint32_t vu = wrap_u ? su : su + 1;
    uint32_t vv = wrap_v ? sv : sv + 1;
    uint32_t base = (uint32_t)(fm.pos.size() / 3);
    for (uint32_t j = 0; j < vv; j++)
        for (uint32_t i = 0; i < vu; i++) {
            float fu = (float)i / (float)su, fv = (float)j / (float)sv;
            float p[3], n[3];
            eval(fu, fv, p, n);
            fm.pos.insert(fm.pos.end(), p, p + 3);
            fm.nrm.insert(fm.nrm.end(), n, n + 3);
            fm.uv.push_back(fu);
            fm.uv.push_back(fv);
        }
    auto vid = [&](uint32_t i, uint32_t j) { return base + (j % vv) * vu + (i % vu); };
    for (uint32_t pj = 0; pj < sv; pj += kPatchQuads)
        for (uint32_t pi = 0; pi < su; pi += kPatchQuads)
            for (uint32_t j = pj; j < std::min(pj + kPatchQuads, sv); j++)
                for (uint32_t i = pi; i < std::min(pi + kPatchQuads, su); i++) {
                    uint32_t a = vid(i, j), b = vid(i + 1, j), c = vid(i + 1, j + 1), d = vid(i, j + 1);
                    if (!flip) {
                        uint32_t t[6] = {a, b, c, a, c, d};
                        fm.tris.insert(fm.tris.end(), t, t + 6);
                    } else {
                        uint32_t t[6] = {a, d, c, a, c, b};
                        fm.tris.insert(fm.tris.end(), t, t + 6);
                    }
                }
}

void make_float_mesh(const LhMeshDesc& d, FloatMesh& fm) {
    const float two_pi = 6.28318530717958647692f;
    uint32_t su = std::max(3u, d.segments_u), sv = std::max(3u, d.segments_v);
    switch (d.kind) {
    case LH_MESH_TORUS: {
        const float R = 1.0f, r = 0.35f;
        emit_grid(fm, su, sv, true, true, /*flip=*/true, [&](float fu, float fv, float* p, float* n) {
            float u = fu * two_pi, v = fv * two_pi;
            float cu = std::cos(u), sn_u = std::sin(u), cv = std::cos(v), sn_v = std::sin(v);
            p[0] = (R + r * cv) * cu; p[1] = r * sn_v; p[2] = (R + r * cv) * sn_u;
            n[0] = cv * cu; n[1] = sn_v; n[2] = cv * sn_u;
        });
        break;
    }
    case LH_MESH_SPHERE: {
        emit_grid(fm, su, sv, true, false, /*flip=*/false, [&](float fu, float fv, float* p, float* n) {
            float phi = fu * two_pi, theta = fv * (two_pi * 0.5f);
            float st = std::sin(theta), ct = std::cos(theta);
            n[0] = st * std::cos(phi); n[1] = ct; n[2] = st * std::sin(phi);
            p[0] = n[0]; p[1] = n[1]; p[2] = n[2];
        });
        break;
    }
    case LH_MESH_PLANE: {
        emit_grid(fm, su, sv, false, false, /*flip=*/true, [&](float fu, float fv, float* p, float* n) {
            p[0] = fu - 0.5f; p[1] = 0.0f; p[2] = fv - 0.5f;
            n[0] = 0; n[1] = 1; n[2] = 0;
        });
        break;
    }
    case LH_MESH_BOX:
    default: {
        // six faces; face f has outward normal N, tangent frame (A, B) with A x B = N
        const float F[6][9] = {
            {1, 0, 0, 0, 1, 0, 0, 0, 1},  {-1, 0, 0, 0, 0, 1, 0, 1, 0}, {0, 1, 0, 0, 0, 1, 1, 0, 0},
            {0, -1, 0, 1, 0, 0, 0, 0, 1}, {0, 0, 1, 1, 0, 0, 0, 1, 0},  {0, 0, -1, 0, 1, 0, 1, 0, 0}};
        for (int f = 0; f < 6; f++) {
            const float* N = F[f];
            const float* A = F[f] + 3;
            const float* B = F[f] + 6;
            emit_grid(fm, su, sv, false, false, /*flip=*/false, [&](float fu, float fv, float* p, float* n) {
                for (int k = 0; k < 3; k++) {
                    p[k] = 0.5f * N[k] + (fu - 0.5f) * A[k] + (fv - 0.5f) * B[k];
                    n[k] = N[k];
                }
            });
        }
        break;
    }
    }
}

// Ritter-style sphere as used for MeshInfo.bounding_sphere and inside the
// meshlet bounds (src/resources/bounding_sphere.h:3-74, meshoptimizer
// clusterizer.cpp computeBoundingSphere): extremal pair along the axis of
// largest spread, then grow to include every point in input order.
void bounding_sphere(float result[4], const float* pts, size_t count) {
    size_t lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (size_t i = 0; i < count; i++)
        for (int a = 0; a < 3; a++) {
            if (pts[3 * i + a] < pts[3 * lo[a] + a]) lo[a] = i;
            if (pts[3 * i + a] > pts[3 * hi[a] + a]) hi[a] = i;
        }
    float best = 0;
    int axis = 0;
    for (int a = 0; a < 3; a++) {
        const float* p = pts + 3 * lo[a];
        const float* q = pts + 3 * hi[a];
        float d2 = (q[0] - p[0]) * (q[0] - p[0]) + (q[1] - p[1]) * (q[1] - p[1]) + (q[2] - p[2]) * (q[2] - p[2]);
        if (d2 > best) {
            best = d2;
            axis = a;
        }
    }
    const float* p = pts + 3 * lo[axis];
    const float* q = pts + 3 * hi[axis];
    float c[3] = {(p[0] + q[0]) / 2, (p[1] + q[1]) / 2, (p[2] + q[2]) / 2};
    float radius = sqrtf(best) / 2;
    for (size_t i = 0; i < count; i++) {
        const float* s = pts + 3 * i;
        float d2 = (s[0] - c[0]) * (s[0] - c[0]) + (s[1] - c[1]) * (s[1] - c[1]) + (s[2] - c[2]) * (s[2] - c[2]);
        if (d2 > radius * radius) {
            float d = sqrtf(d2);
            float k = 0.5f + (radius / d) / 2;
            c[0] = c[0] * k + s[0] * (1 - k);
            c[1] = c[1] * k + s[1] * (1 - k);
            c[2] = c[2] * k + s[2] * (1 - k);
            radius = (radius + d) / 2;
        }
    }
    result[0] = c[0]; result[1] = c[1]; result[2] = c[2]; result[3] = radius;
}

// Sphere + normal cone of one meshlet (meshoptimizer v0.20
// meshopt_computeMeshletBounds / meshopt_computeClusterBounds semantics;
// called by the reference at src/resources/meshlets.cpp:69-86).
void meshlet_bounds(const uint32_t* verts, const uint8_t* tris, uint32_t tri_count, const float* positions,
                    LhMeshlet* out) {
    float normals[LH_MAX_MESHLET_TRIANGLES][3];
    float corners[LH_MAX_MESHLET_TRIANGLES * 3][3];
    size_t n = 0;
    for (uint32_t t = 0; t < tri_count; t++) {
        const float* p0 = positions + 3 * verts[tris[3 * t + 0]];
        const float* p1 = positions + 3 * verts[tris[3 * t + 1]];
        const float* p2 = positions + 3 * verts[tris[3 * t + 2]];
        float e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
        float e2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
        float nx = e1[1] * e2[2] - e1[2] * e2[1];
        float ny = e1[2] * e2[0] - e1[0] * e2[2];
        float nz = e1[0] * e2[1] - e1[1] * e2[0];
        float area = sqrtf(nx * nx + ny * ny + nz * nz);
        if (area == 0.f) continue;  // degenerate triangles carry no normal
        normals[n][0] = nx / area; normals[n][1] = ny / area; normals[n][2] = nz / area;
        memcpy(corners[3 * n + 0], p0, 12);
        memcpy(corners[3 * n + 1], p1, 12);
        memcpy(corners[3 * n + 2], p2, 12);
        n++;
    }
    memset(out->cone_apex, 0, sizeof(out->cone_apex));
    memset(out->cone_axis, 0, sizeof(out->cone_axis));
    out->cone_cutoff = 0;
    memset(out->bounding_sphere, 0, sizeof(out->bounding_sphere));
    if (n == 0) return;  // all-zero bounds: trivially rejected by the cone test

    float ps[4], ns[4];
    bounding_sphere(ps, &corners[0][0], n * 3);
    bounding_sphere(ns, &normals[0][0], n);
    float axis[3] = {ns[0], ns[1], ns[2]};
    float alen = sqrtf(axis[0] * axis[0] + axis[1] * axis[1] + axis[2] * axis[2]);
    float inv = alen == 0.f ? 0.f : 1.f / alen;
    axis[0] *= inv; axis[1] *= inv; axis[2] *= inv;

    float mindp = 1.f;
    for (size_t i = 0; i < n; i++) {
        float dp = normals[i][0] * axis[0] + normals[i][1] * axis[1] + normals[i][2] * axis[2];
        mindp = dp < mindp ? dp : mindp;
    }
    out->bounding_sphere[0] = ps[0]; out->bounding_sphere[1] = ps[1];
    out->bounding_sphere[2] = ps[2]; out->bounding_sphere[3] = ps[3];
    if (mindp <= 0.1f) {  // cone wider than ~168 degrees: never culled
        out->cone_cutoff = 1;
        return;
    }
    float maxt = 0;
    for (size_t i = 0; i < n; i++) {
        float cx = ps[0] - corners[3 * i][0], cy = ps[1] - corners[3 * i][1], cz = ps[2] - corners[3 * i][2];
        float dc = cx * normals[i][0] + cy * normals[i][1] + cz * normals[i][2];
        float dn = axis[0] * normals[i][0] + axis[1] * normals[i][1] + axis[2] * normals[i][2];
        float t = dc / dn;
        maxt = t > maxt ? t : maxt;
    }
    out->cone_apex[0] = ps[0] - axis[0] * maxt;
    out->cone_apex[1] = ps[1] - axis[1] * maxt;
    out->cone_apex[2] = ps[2] - axis[2] * maxt;
    out->cone_axis[0] = axis[0]; out->cone_axis[1] = axis[1]; out->cone_axis[2] = axis[2];
    out->cone_cutoff = sqrtf(1 - mindp * mindp);
}

#ifdef LH_HAVE_MESHOPTIMIZER
// The reference's build_meshlets (src/resources/meshlets.cpp:3-140): meshopt_buildMeshlets of the vendored
// meshoptimizer v0.20 with 64 vertices / 124 triangles / cone weight 0.25, then the same packing into Meshlet records,
// vertex indices and 4-byte-padded micro-index runs.  The library is compiled from where the reference vendors it
// (Makefile: external/meshoptimizer/src/clusterizer.cpp); tests compare the result byte for byte with the reference's
// own function (oracle/_ref/libref_sceneprep.so) and with tests/golden/torus_*_ref.npz.
bool build_meshlets_meshopt(Mesh& mesh, const std::vector<uint32_t>& tris, const std::vector<float>& raw_positions, size_t num_vertices) {
    if (tris.empty()) return true;
    const size_t max_vertices = LH_MAX_MESHLET_UNIQUE_VERTICES, max_triangles = LH_MAX_MESHLET_TRIANGLES;
    const float cone_weight = 0.25f;
    const size_t max_meshlets = meshopt_buildMeshletsBound(tris.size(), max_vertices, max_triangles);
    std::vector<meshopt_Meshlet> meshlets(max_meshlets);
    std::vector<unsigned int> vertex_indices(max_meshlets * max_vertices);
    std::vector<unsigned char> micro(max_meshlets * max_triangles * 3);
    const size_t count = meshopt_buildMeshlets(meshlets.data(), vertex_indices.data(), micro.data(), tris.data(), tris.size(), raw_positions.data(),
                                               num_vertices, sizeof(float) * 3, max_vertices, max_triangles, cone_weight);
    if (count == 0) return true;
    const meshopt_Meshlet& last = meshlets[count - 1];
    vertex_indices.resize(last.vertex_offset + last.vertex_count);
    micro.resize(last.triangle_offset + ((last.triangle_count * 3 + 3) & ~3u));
    mesh.meshlets.resize(count);
    for (size_t i = 0; i < count; i++) {
        const meshopt_Meshlet& m = meshlets[i];
        LhMeshlet out{};
        meshlet_bounds(&vertex_indices[m.vertex_offset], &micro[m.triangle_offset], m.triangle_count, raw_positions.data(), &out);
        out.triangle_offset = m.triangle_offset;
        out.index_offset = m.vertex_offset;
        out.triangle_count = (uint8_t)m.triangle_count;
        out.index_count = (uint8_t)m.vertex_count;
        mesh.meshlets[i] = out;
    }
    mesh.micro.assign(micro.begin(), micro.end());
    mesh.wide_indices = num_vertices > 65536;
    if (mesh.wide_indices) {
        mesh.idx32.assign(vertex_indices.begin(), vertex_indices.end());
    } else {
        mesh.idx16.resize(vertex_indices.size());
        for (size_t i = 0; i < vertex_indices.size(); i++) mesh.idx16[i] = (uint16_t)vertex_indices[i];
    }
    return true;
}
#endif

bool use_meshoptimizer() {
#ifdef LH_HAVE_MESHOPTIMIZER
    const char* e = getenv("LH_MESHLET_BUILDER");   // "greedy": the built-in fallback even where meshoptimizer is linked
    return !(e && strcmp(e, "greedy") == 0);
#else
    return false;
#endif
}

// Fallback where the reference's vendored meshoptimizer is not available at build time.
// Greedy meshlet fill: walk the triangle list in order, start a new meshlet
// whenever adding the triangle would exceed 64 unique vertices or 124
// triangles.  Output = Meshlet[] + vertex-index buffer + micro-indices with
// each meshlet's run padded to 4 bytes, as the reference packs them.
void build_meshlets(Mesh& mesh, const std::vector<uint32_t>& tris, const std::vector<float>& raw_positions,
                    size_t num_vertices) {
#ifdef LH_HAVE_MESHOPTIMIZER
    if (use_meshoptimizer()) {
        build_meshlets_meshopt(mesh, tris, raw_positions, num_vertices);
        return;
    }
#endif
    std::vector<int16_t> local(num_vertices, -1);
    std::vector<uint32_t> vertex_indices;  // concatenated per-meshlet vertex lists
    std::vector<uint32_t> cur_verts;
    std::vector<uint8_t> cur_tris;
    uint32_t index_offset = 0;

    auto flush = [&]() {
        if (cur_tris.empty()) return;
        LhMeshlet m{};
        meshlet_bounds(cur_verts.data(), cur_tris.data(), (uint32_t)(cur_tris.size() / 3), raw_positions.data(), &m);
        m.triangle_offset = (uint32_t)mesh.micro.size();
        m.index_offset = index_offset;
        m.triangle_count = (uint8_t)(cur_tris.size() / 3);
        m.index_count = (uint8_t)cur_verts.size();
        mesh.meshlets.push_back(m);
        mesh.micro.insert(mesh.micro.end(), cur_tris.begin(), cur_tris.end());
        while (mesh.micro.size() & 3) mesh.micro.push_back(0);
        vertex_indices.insert(vertex_indices.end(), cur_verts.begin(), cur_verts.end());
        index_offset += (uint32_t)cur_verts.size();
        for (uint32_t v : cur_verts) local[v] = -1;
        cur_verts.clear();
        cur_tris.clear();
    };

    for (size_t t = 0; t + 2 < tris.size(); t += 3) {
        uint32_t v[3] = {tris[t], tris[t + 1], tris[t + 2]};
        uint32_t fresh = 0;
        for (int k = 0; k < 3; k++) {
            bool seen = local[v[k]] >= 0;
            for (int q = 0; q < k; q++) seen = seen || v[q] == v[k];
            fresh += seen ? 0 : 1;
        }
        if (cur_verts.size() + fresh > LH_MAX_MESHLET_UNIQUE_VERTICES ||
            cur_tris.size() / 3 + 1 > LH_MAX_MESHLET_TRIANGLES)
            flush();
        for (int k = 0; k < 3; k++) {
            if (local[v[k]] < 0) {
                local[v[k]] = (int16_t)cur_verts.size();
                cur_verts.push_back(v[k]);
            }
            cur_tris.push_back((uint8_t)local[v[k]]);
        }
    }
    flush();

    mesh.wide_indices = num_vertices > 65536;
    if (mesh.wide_indices) {
        mesh.idx32 = vertex_indices;
    } else {
        mesh.idx16.resize(vertex_indices.size());
        for (size_t i = 0; i < vertex_indices.size(); i++) mesh.idx16[i] = (uint16_t)vertex_indices[i];
    }
}

}  // namespace

struct LhScene {
    std::vector<std::unique_ptr<Mesh>> meshes;
    std::vector<InstanceRecord> instances;
};

extern "C" {

int lh_scene_create(LhScene** out) {
    if (!out) return -1;
    *out = new LhScene();
    return 0;
}

void lh_scene_destroy(LhScene* scene) { delete scene; }

int lh_scene_add_mesh(LhScene* scene, const LhMeshDesc* desc) {
    if (!scene || !desc) return -1;
    FloatMesh fm;
    make_float_mesh(*desc, fm);
    size_t nv = fm.pos.size() / 3;
    auto mesh = std::make_unique<Mesh>();

    // quantise: q = round((p - min) / (max - min) * 65535); dequantisation is
    // folded into the instance transform
    float lo[3], hi[3];
    for (int a = 0; a < 3; a++) {
        lo[a] = hi[a] = fm.pos[a];
    }
    for (size_t i = 0; i < nv; i++)
        for (int a = 0; a < 3; a++) {
            lo[a] = std::min(lo[a], fm.pos[3 * i + a]);
            hi[a] = std::max(hi[a], fm.pos[3 * i + a]);
        }
    // One scale for all three axes: the reference's normal_transform is
    // mat3(inverse(transform)) WITHOUT the transpose (src/shared_cpu_gpu.h:92-97),
    // which is exact only when the linear part is a symmetric rotation times a
    // uniform scale.  Scenes stay inside that domain (see place_instances).
    float max_ext = 0.0f;
    for (int a = 0; a < 3; a++) max_ext = std::max(max_ext, hi[a] - lo[a]);
    if (!(max_ext > 1e-20f)) max_ext = 1.0f;
    for (int a = 0; a < 3; a++) {
        mesh->dq_scale[a] = max_ext / 65535.0f;
        mesh->dq_offset[a] = lo[a];
    }
    mesh->positions.resize(nv * 3);
    mesh->normals.resize(nv * 3);
    mesh->uvs.resize(nv * 2);
    std::vector<float> raw(nv * 3);
    for (size_t i = 0; i < nv; i++) {
        for (int a = 0; a < 3; a++) {
            float q = std::nearbyint((fm.pos[3 * i + a] - lo[a]) / (mesh->dq_scale[a] * 65535.0f) * 65535.0f);
            q = std::min(65535.0f, std::max(0.0f, q));
            mesh->positions[3 * i + a] = (uint16_t)q;
            raw[3 * i + a] = q;  // bounds are computed on the raw integer values (mesh_loading.cpp:584-597)
            float nq = std::nearbyint(fm.nrm[3 * i + a] * 127.0f);
            mesh->normals[3 * i + a] = (int8_t)std::min(127.0f, std::max(-127.0f, nq));
        }
        for (int a = 0; a < 2; a++) {
            float q = std::nearbyint(fm.uv[2 * i + a] * 65535.0f);
            mesh->uvs[2 * i + a] = (uint16_t)std::min(65535.0f, std::max(0.0f, q));
        }
    }
    mesh->num_triangles = fm.tris.size() / 3;
    build_meshlets(*mesh, fm.tris, raw, nv);
    mesh->triangles = fm.tris;
    if (mesh->meshlets.size() > 65535) return -2;  // 16-bit meshlet index cap (mesh_loading.cpp:710-713)

    LhMeshInfo& info = mesh->info;
    memset(&info, 0, sizeof(info));
    info.num_meshlets = (uint16_t)mesh->meshlets.size();
    info.flags = (uint8_t)((mesh->wide_indices ? LH_MESH_INFO_FLAGS_32_BIT_INDICES : 0u) |
                           (desc->flags & LH_MESH_INFO_FLAGS_ALPHA_CLIP));
    bounding_sphere(info.bounding_sphere, raw.data(), nv);
    float rep = desc->uv_repeat > 0 ? desc->uv_repeat : 1.0f;
    info.texture_scale[0] = info.texture_scale[1] = rep / 65535.0f;  // mesh_loading.cpp:555-557
    info.texture_offset[0] = info.texture_offset[1] = 0.0f;
    info.base_color_texture_index = desc->base_color_texture_index;
    info.metallic_roughness_texture_index = desc->metallic_roughness_texture_index;
    info.normal_texture_index = desc->normal_texture_index;
    for (int a = 0; a < 3; a++) info.base_color_factor[a] = desc->base_color_factor[a];
    scene->meshes.push_back(std::move(mesh));
    return (int)scene->meshes.size() - 1;
}

int lh_scene_add_mesh_indexed(LhScene* scene, const LhIndexedMesh* d) {
    if (!scene || !d || !d->positions || !d->normals || !d->uvs || !d->indices) return -1;
    if (d->position_stride < 6 || d->normal_stride < 3 || d->uv_stride < 4 || (d->index_bytes != 2 && d->index_bytes != 4)) return -1;
    if (d->num_vertices == 0 || d->num_indices % 3 != 0) return -1;
    const size_t nv = (size_t)d->num_vertices;
    auto mesh = std::make_unique<Mesh>();
    mesh->positions.resize(nv * 3);
    mesh->normals.resize(nv * 3);
    mesh->uvs.resize(nv * 2);
    std::vector<float> raw(nv * 3);
    for (size_t i = 0; i < nv; i++) {
        uint16_t p[3], uv[2];
        int8_t n[3];
        memcpy(p, (const uint8_t*)d->positions + i * d->position_stride, 6);
        memcpy(n, (const uint8_t*)d->normals + i * d->normal_stride, 3);
        memcpy(uv, (const uint8_t*)d->uvs + i * d->uv_stride, 4);
        for (int a = 0; a < 3; a++) {
            mesh->positions[3 * i + a] = p[a];
            raw[3 * i + a] = (float)p[a];   // meshlets and bounds are built on the raw integers (mesh_loading.cpp:584-597)
            mesh->normals[3 * i + a] = n[a];
        }
        mesh->uvs[2 * i] = uv[0];
        mesh->uvs[2 * i + 1] = uv[1];
    }
    std::vector<uint32_t> tris((size_t)d->num_indices);
    for (size_t i = 0; i < tris.size(); i++) {
        tris[i] = d->index_bytes == 4 ? ((const uint32_t*)d->indices)[i] : (uint32_t)((const uint16_t*)d->indices)[i];
        if (tris[i] >= nv) return -3;   // an index past the vertex streams
    }
    mesh->num_triangles = tris.size() / 3;
    build_meshlets(*mesh, tris, raw, nv);
    mesh->triangles = tris;
    if (d->index_bytes == 4 && !mesh->wide_indices) {   // the index width follows the accessor (mesh_loading.cpp:455-462)
        mesh->idx32.assign(mesh->idx16.begin(), mesh->idx16.end());
        mesh->idx16.clear();
        mesh->wide_indices = true;
    }
    if (mesh->meshlets.size() > 65535) return -2;  // 16-bit meshlet index cap (mesh_loading.cpp:710-713)
    LhMeshInfo& info = mesh->info;
    memset(&info, 0, sizeof(info));
    info.num_meshlets = (uint16_t)mesh->meshlets.size();
    info.flags = (uint8_t)((mesh->wide_indices ? LH_MESH_INFO_FLAGS_32_BIT_INDICES : 0u) | (d->flags & LH_MESH_INFO_FLAGS_ALPHA_CLIP));
    bounding_sphere(info.bounding_sphere, raw.data(), nv);
    for (int a = 0; a < 2; a++) {
        info.texture_scale[a] = d->texture_scale[a];
        info.texture_offset[a] = d->texture_offset[a];
    }
    info.base_color_texture_index = d->base_color_texture_index;
    info.metallic_roughness_texture_index = d->metallic_roughness_texture_index;
    info.normal_texture_index = d->normal_texture_index;
    for (int a = 0; a < 3; a++) info.base_color_factor[a] = d->base_color_factor[a];
    scene->meshes.push_back(std::move(mesh));
    return (int)scene->meshes.size() - 1;
}

// A material variant: a second MeshInfo (own texture indices / colour factor) over the SAME vertex, index and meshlet
// buffers -- in the reference every glTF primitive has its own MeshInfo buffer (mesh_loading.cpp:699-708), and
// primitives that differ only in material are exactly this.
int lh_scene_add_mesh_variant(LhScene* scene, int mesh, uint16_t base_color, uint16_t metallic_roughness, uint16_t normal,
                              const float base_color_factor[3]) {
    if (!scene || mesh < 0 || mesh >= (int)scene->meshes.size()) return -1;
    const Mesh& src = *scene->meshes[mesh];
    auto v = std::make_unique<Mesh>(src);
    v->shares_geometry_of = src.shares_geometry_of >= 0 ? src.shares_geometry_of : mesh;
    v->info.base_color_texture_index = base_color;
    v->info.metallic_roughness_texture_index = metallic_roughness;
    v->info.normal_texture_index = normal;
    if (base_color_factor)
        for (int a = 0; a < 3; a++) v->info.base_color_factor[a] = base_color_factor[a];
    scene->meshes.push_back(std::move(v));
    return (int)scene->meshes.size() - 1;
}

int lh_scene_add_instance_matrix(LhScene* scene, int mesh, const float transform[16]) {
    if (!scene || !transform || mesh < 0 || mesh >= (int)scene->meshes.size()) return -1;
    LhMat4 m;
    memcpy(&m, transform, sizeof(m));
    scene->instances.push_back({mesh, m});
    return 0;
}

int lh_scene_add_mesh_raw(LhScene* scene, const LhMeshView* v) {
    if (!scene || !v) return -1;
    auto mesh = std::make_unique<Mesh>();
    mesh->positions.assign(v->positions, v->positions + v->num_vertices * 3);
    mesh->normals.assign(v->normals, v->normals + v->num_vertices * 3);
    mesh->uvs.assign(v->uvs, v->uvs + v->num_vertices * 2);
    mesh->wide_indices = v->index_bytes == 4;
    if (mesh->wide_indices)
        mesh->idx32.assign((const uint32_t*)v->indices, (const uint32_t*)v->indices + v->num_indices);
    else
        mesh->idx16.assign((const uint16_t*)v->indices, (const uint16_t*)v->indices + v->num_indices);
    mesh->micro.assign(v->micro_indices, v->micro_indices + v->micro_indices_bytes);
    mesh->meshlets.assign(v->meshlets, v->meshlets + v->num_meshlets);
    mesh->num_triangles = v->num_triangles;
    mesh->info = v->info;
    mesh->info.num_meshlets = (uint16_t)v->num_meshlets;
    for (int a = 0; a < 3; a++) {
        mesh->dq_scale[a] = v->dequant_scale[a];
        mesh->dq_offset[a] = v->dequant_offset[a];
    }
    scene->meshes.push_back(std::move(mesh));
    return (int)scene->meshes.size() - 1;
}

int lh_scene_num_meshes(const LhScene* scene) { return scene ? (int)scene->meshes.size() : -1; }

int lh_scene_mesh_view(const LhScene* scene, int id, LhMeshView* out) {
    if (!scene || !out || id < 0 || id >= (int)scene->meshes.size()) return -1;
    const Mesh& m = *scene->meshes[id];
    memset(out, 0, sizeof(*out));
    out->positions = m.positions.data();
    out->num_vertices = m.positions.size() / 3;
    out->normals = m.normals.data();
    out->uvs = m.uvs.data();
    out->indices = m.wide_indices ? (const void*)m.idx32.data() : (const void*)m.idx16.data();
    out->num_indices = m.wide_indices ? m.idx32.size() : m.idx16.size();
    out->index_bytes = m.wide_indices ? 4 : 2;
    out->micro_indices = m.micro.data();
    out->micro_indices_bytes = m.micro.size();
    out->meshlets = m.meshlets.data();
    out->num_meshlets = m.meshlets.size();
    out->num_triangles = m.num_triangles;
    out->info = m.info;
    for (int a = 0; a < 3; a++) {
        out->dequant_scale[a] = m.dq_scale[a];
        out->dequant_offset[a] = m.dq_offset[a];
    }
    return 0;
}

int lh_scene_add_instance(LhScene* scene, int mesh, const float t[3], const float q[4], float scale) {
    if (!scene || mesh < 0 || mesh >= (int)scene->meshes.size()) return -1;
    const Mesh& m = *scene->meshes[mesh];
    LhMat4 dequant = mul(translation({m.dq_offset[0], m.dq_offset[1], m.dq_offset[2]}),
                         scaling({m.dq_scale[0], m.dq_scale[1], m.dq_scale[2]}));
    LhMat4 trs = mul(translation({t[0], t[1], t[2]}), mul(rotation(q), scaling({scale, scale, scale})));
    scene->instances.push_back({mesh, mul(trs, dequant)});
    return 0;
}

int lh_scene_place_instances(LhScene* scene, uint32_t count, uint64_t seed, const float extent[3], float smin,
                             float smax, const int32_t* mesh_ids, uint32_t num_mesh_ids) {
    if (!scene || !mesh_ids || num_mesh_ids == 0) return -1;
    uint32_t first = (uint32_t)scene->instances.size();
    for (uint32_t k = 0; k < count; k++) {
        uint64_t c = (uint64_t)(first + k) * 8;
        float t[3];
        for (int a = 0; a < 3; a++) t[a] = (u01(seed, c + a) * 2.0f - 1.0f) * extent[a];
        float s = smin + (smax - smin) * u01(seed, c + 3);
        // Half-turn about a uniformly random axis n: R = 2 n n^T - I is symmetric,
        // so inverse(T) needs no transpose (the reference's own scenes have no
        // rotation at all, mesh_loading.cpp:8-49), yet orientations still vary.
        float u1 = u01(seed, c + 4), u2 = u01(seed, c + 5);
        const float two_pi = 6.28318530717958647692f;
        float nz = 2.0f * u1 - 1.0f, nr = std::sqrt(std::max(0.0f, 1.0f - nz * nz));
        float q[4] = {nr * std::cos(two_pi * u2), nr * std::sin(two_pi * u2), nz, 0.0f};
        int rc = lh_scene_add_instance(scene, mesh_ids[(first + k) % num_mesh_ids], t, q, s);
        if (rc) return rc;
    }
    return 0;
}

// Uniformly random rotations (unit quaternions) and a different scale per axis: here mat3(inverse(T)) differs from the
// inverse transpose, which exercises the reference's untransposed normal transform (shared_cpu_gpu.h:92-97,
// culling.glsl:141) where it is NOT the correct normal matrix (SURVEY 8d "random rotation").  The reference's own scenes
// never rotate (mesh_loading.cpp:8-49), so this placement is for parity tests; the bench scenes keep half-turns.
int lh_scene_place_instances_general(LhScene* scene, uint32_t count, uint64_t seed, const float extent[3], float smin,
                                     float smax, const int32_t* mesh_ids, uint32_t num_mesh_ids) {
    if (!scene || !mesh_ids || num_mesh_ids == 0) return -1;
    uint32_t first = (uint32_t)scene->instances.size();
    for (uint32_t k = 0; k < count; k++) {
        uint64_t c = (uint64_t)(first + k) * 16;
        int mesh = mesh_ids[(first + k) % num_mesh_ids];
        if (mesh < 0 || mesh >= (int)scene->meshes.size()) return -1;
        const Mesh& m = *scene->meshes[mesh];
        float t[3], s[3];
        for (int a = 0; a < 3; a++) t[a] = (u01(seed, c + a) * 2.0f - 1.0f) * extent[a];
        for (int a = 0; a < 3; a++) s[a] = smin + (smax - smin) * u01(seed, c + 3 + a);
        // Shoemake: a uniformly distributed unit quaternion from three uniforms
        const float two_pi = 6.28318530717958647692f;
        float u1 = u01(seed, c + 6), u2 = u01(seed, c + 7), u3 = u01(seed, c + 8);
        float a = std::sqrt(1.0f - u1), b = std::sqrt(u1);
        float q[4] = {a * std::sin(two_pi * u2), a * std::cos(two_pi * u2), b * std::sin(two_pi * u3), b * std::cos(two_pi * u3)};
        LhMat4 dequant = mul(translation({m.dq_offset[0], m.dq_offset[1], m.dq_offset[2]}),
                             scaling({m.dq_scale[0], m.dq_scale[1], m.dq_scale[2]}));
        LhMat4 trs = mul(translation({t[0], t[1], t[2]}), mul(rotation(q), scaling({s[0], s[1], s[2]})));
        scene->instances.push_back({mesh, mul(trs, dequant)});
    }
    return 0;
}

uint32_t lh_scene_num_instances(const LhScene* scene) { return scene ? (uint32_t)scene->instances.size() : 0; }

uint64_t lh_scene_total_triangles(const LhScene* scene) {
    uint64_t n = 0;
    for (auto& i : scene->instances) n += scene->meshes[i.mesh]->num_triangles;
    return n;
}

uint64_t lh_scene_total_meshlets(const LhScene* scene) {
    uint64_t n = 0;
    for (auto& i : scene->instances) n += scene->meshes[i.mesh]->meshlets.size();
    return n;
}

int lh_scene_build_instances(const LhScene* scene, const uint64_t* mesh_info_addresses, LhInstance* out) {
    if (!scene || !mesh_info_addresses || !out) return -1;
    for (size_t i = 0; i < scene->instances.size(); i++) {
        const InstanceRecord& r = scene->instances[i];
        LhInstance inst;
        memset(&inst, 0, sizeof(inst));
        inst.transform = r.transform;
        inst.mesh_info_address = mesh_info_addresses[r.mesh];
        LhMat4 inv = inverse(r.transform);  // mat3(inverse(T)), no transpose
        for (int c = 0; c < 3; c++)
            for (int row = 0; row < 3; row++) inst.normal_transform.m[c][row] = inv.m[c][row];
        out[i] = inst;
    }
    return 0;
}

void lh_camera_default(LhCamera* c) {
    if (!c) return;
    memset(c, 0, sizeof(*c));
    c->fov_degrees = 45.0f;       // src/main.cpp:606-613
    c->sun_latitude = -5.6f;
    c->sun_longitude = 1.37f;
    c->shadow_cam_distance = 1024.0f;  // src/main.cpp:653-654
    c->cascade_split_pow = 3.0f;
}

int lh_scene_uniforms(const LhScene* scene, const LhCamera* cam, uint32_t width, uint32_t height, LhUniforms* u) {
    if (!cam || !u || width == 0 || height == 0) return -1;
    memset(u, 0, sizeof(*u));
    Vec3 pos{cam->position[0], cam->position[1], cam->position[2]};
    Vec3 facing{std::cos(cam->yaw) * std::cos(cam->pitch), std::sin(cam->pitch),
                std::sin(cam->yaw) * std::cos(cam->pitch)};
    LhMat4 view = look_at(pos, pos + facing, {0, 1, 0});
    const float deg2rad = 0.01745329251994329577f;
    LhMat4 persp = infinite_reverse_z_perspective(cam->fov_degrees * deg2rad, (float)width, (float)height, LH_NEAR_PLANE);
    LhMat4 pv = mul(persp, view);
    u->combined_perspective_view = pv;
    u->inv_perspective_view = inverse(pv);
    u->view = view;
    u->initial_view = view;  // bench/tests always track the live camera (copy_view on, main.cpp:756-759)
    u->perspective = persp;
    u->perspective_inverse = inverse(persp);
    u->view_inverse = inverse(view);
    u->camera_pos[0] = pos.x; u->camera_pos[1] = pos.y; u->camera_pos[2] = pos.z;
    u->sun_dir[0] = std::cos(cam->sun_latitude) * std::cos(cam->sun_longitude);  // src/input.h:43-49
    u->sun_dir[1] = std::sin(cam->sun_longitude);
    u->sun_dir[2] = std::sin(cam->sun_latitude) * std::cos(cam->sun_longitude);
    u->sun_intensity[0] = u->sun_intensity[1] = u->sun_intensity[2] = 1.0f;
    u->num_instances = scene ? (uint32_t)scene->instances.size() : 0;
    u->window_size[0] = width;
    u->window_size[1] = height;
    u->shadow_cam_distance = cam->shadow_cam_distance;
    u->cascade_split_pow = cam->cascade_split_pow;
    u->debug = LH_UNIFORMS_DEBUG_OFF;
    u->debug_shadowmaps = 0;
    return 0;
}

uint64_t lh_scene_make_texture(int kind, uint32_t size, uint32_t variant, int alpha_cutout, uint8_t* dst,
                               uint32_t* mip_levels) {
    uint32_t levels = 0;
    uint64_t total = 0;
    for (uint32_t s = size; s >= 1; s >>= 1) {
        total += (uint64_t)s * s * 4;
        levels++;
        if (s == 1) break;
    }
    if (mip_levels) *mip_levels = levels;
    if (!dst) return total;
    const float two_pi = 6.28318530717958647692f;
    // level 0: smooth, low-frequency content so that filtering differences stay small
    float hue[3] = {0.55f + 0.4f * u01(0xC0FFEEull + variant, 1), 0.55f + 0.4f * u01(0xC0FFEEull + variant, 2),
                    0.55f + 0.4f * u01(0xC0FFEEull + variant, 3)};
    float ph = u01(0xC0FFEEull + variant, 4);
    for (uint32_t y = 0; y < size; y++)
        for (uint32_t x = 0; x < size; x++) {
            float u = (x + 0.5f) / size, v = (y + 0.5f) / size;
            float w1 = 0.5f + 0.5f * std::sin(two_pi * (u + ph)) * std::sin(two_pi * (v * 2 + ph));
            float w2 = 0.5f + 0.5f * std::sin(two_pi * (u * 3 - v + ph));
            float px[4];
            if (kind == 0) {
                for (int c = 0; c < 3; c++) px[c] = hue[c] * (0.55f + 0.45f * w1);
                px[3] = alpha_cutout ? 0.5f + 0.6f * std::sin(two_pi * 3 * u) * std::sin(two_pi * 3 * v) : 1.0f;
            } else if (kind == 1) {
                px[0] = 1.0f;
                // roughness (g).  The reference samples this map through an sRGB format like every DDS texture
                // (image_loading.cpp:20-34), so 0.6 .. 0.95 stored is a perceptual roughness of 0.32 .. 0.89 in the shader:
                // the range of everyday materials.  (A mirror-like 0.1 makes the GGX lobe so narrow that the 1/512
                // difference two texture units may have in a bilinear weight of the normal map changes a highlight pixel's
                // radiance by tens of percent -- content no implementation pair agrees on.)
                px[1] = 0.6f + 0.35f * w2;
                px[2] = 0.1f + 0.5f * w1;   // metallic (b)
                px[3] = 1.0f;
            } else {
                float nx = 0.25f * std::cos(two_pi * (u * 2 + ph)), ny = 0.25f * std::cos(two_pi * (v * 2 - ph));
                float nz = std::sqrt(std::max(0.0f, 1 - nx * nx - ny * ny));
                px[0] = (nx * 127.0f + 128.0f) / 255.0f;
                px[1] = (ny * 127.0f + 128.0f) / 255.0f;
                px[2] = (nz * 127.0f + 128.0f) / 255.0f;
                px[3] = 1.0f;
            }
            for (int c = 0; c < 4; c++)
                dst[((uint64_t)y * size + x) * 4 + c] = (uint8_t)std::nearbyint(std::min(1.0f, std::max(0.0f, px[c])) * 255.0f);
        }
    // box-filtered mips
    uint8_t* prev = dst;
    uint32_t ps = size;
    uint8_t* cur = dst + (uint64_t)size * size * 4;
    while (ps > 1) {
        uint32_t cs = ps / 2;
        for (uint32_t y = 0; y < cs; y++)
            for (uint32_t x = 0; x < cs; x++)
                for (int c = 0; c < 4; c++) {
                    uint32_t s = prev[((uint64_t)(2 * y) * ps + 2 * x) * 4 + c] + prev[((uint64_t)(2 * y) * ps + 2 * x + 1) * 4 + c] +
                                 prev[((uint64_t)(2 * y + 1) * ps + 2 * x) * 4 + c] +
                                 prev[((uint64_t)(2 * y + 1) * ps + 2 * x + 1) * 4 + c];
                    cur[((uint64_t)y * cs + x) * 4 + c] = (uint8_t)((s + 2) / 4);
                }
        prev = cur;
        cur += (uint64_t)cs * cs * 4;
        ps = cs;
    }
    return total;
}

uint64_t lh_scene_make_skybox(uint32_t size, float* dst) {
    uint64_t total = (uint64_t)6 * size * size * 4 * sizeof(float);
    if (!dst) return total;
    for (int f = 0; f < 6; f++)
        for (uint32_t y = 0; y < size; y++)
            for (uint32_t x = 0; x < size; x++) {
                float sc = 2.0f * (x + 0.5f) / size - 1.0f, tc = 2.0f * (y + 0.5f) / size - 1.0f;
                Vec3 d;
                switch (f) {  // standard cube-map face orientation
                case 0: d = {1, -tc, -sc}; break;
                case 1: d = {-1, -tc, sc}; break;
                case 2: d = {sc, 1, tc}; break;
                case 3: d = {sc, -1, -tc}; break;
                case 4: d = {sc, -tc, 1}; break;
                default: d = {-sc, -tc, -1}; break;
                }
                d = normalize(d);
                float t = 0.5f + 0.5f * d.y;
                float* px = dst + (((uint64_t)f * size + y) * size + x) * 4;
                px[0] = 0.08f + 0.10f * t;
                px[1] = 0.10f + 0.18f * t;
                px[2] = 0.14f + 0.35f * t;
                px[3] = 1.0f;
            }
    return total;
}

// ---- presets ------------------------------------------------------------------

int lh_scene_build_preset(LhScene* scene, int preset, uint32_t instance_count, uint64_t seed, LhPresetInfo* out) {
    if (!scene || !out || !scene->meshes.empty() || !scene->instances.empty()) return -1;
    memset(out, 0, sizeof(*out));
    lh_camera_default(&out->camera);
    out->alpha_cutout_set = 0xFFFFFFFFu;
    if (seed == 0) seed = 0x5EED0001ull;
    struct M { int kind; uint32_t su, sv; bool alpha; };
    std::vector<M> meshes;
    uint32_t count = 0;
    float extent = 1.0f, smin = 0.5f, smax = 2.0f;
    switch (preset) {
    case LH_PRESET_TINY:
        meshes = {{LH_MESH_TORUS, 32, 16, false}};
        count = 64; extent = 6.0f; out->width = 256; out->height = 144;
        break;
    case LH_PRESET_C1:
        meshes = {{LH_MESH_TORUS, 125, 80, false}};
        count = 1000; extent = 18.0f; out->width = 1920; out->height = 1080;
        break;
    case LH_PRESET_C2:
        meshes = {{LH_MESH_TORUS, 250, 200, false}};
        count = 10000; extent = 40.0f; out->width = 3840; out->height = 2160;
        break;
    case LH_PRESET_C5:
        meshes = {{LH_MESH_TORUS, 250, 200, false}};
        count = 10000; extent = 40.0f; out->width = 7680; out->height = 4320;
        break;
    case LH_PRESET_C4:
        // mixed meshes 5 k - 200 k tri; every 20th instance is alpha-clip foliage
        meshes = {{LH_MESH_TORUS, 100, 50, false},  {LH_MESH_SPHERE, 128, 64, false}, {LH_MESH_PLANE, 100, 100, false},
                  {LH_MESH_BOX, 40, 40, false},     {LH_MESH_TORUS, 250, 200, false}, {LH_MESH_SPHERE, 200, 100, false},
                  {LH_MESH_TORUS, 160, 100, false}, {LH_MESH_BOX, 64, 64, false},     {LH_MESH_PLANE, 50, 50, true}};
        count = 3000; extent = 30.0f; out->width = 3840; out->height = 2160;
        break;
    default:
        return -1;
    }
    if (instance_count) {
        // keep the instance density of the preset when the count is overridden
        extent *= std::cbrt((float)instance_count / (float)count);
        count = instance_count;
    }
    // 16 base-colour + 16 metallic-roughness + 16 normal maps.  A preset with fewer than 16 meshes gets
    // material variants (same geometry buffers, own MeshInfo and texture set) up to 16; the unit-test preset keeps one.
    const uint32_t n_variants = (preset == LH_PRESET_TINY || meshes.size() >= 16) ? 1u : (uint32_t)(16 / meshes.size());
    uint32_t n_sets = (uint32_t)std::min<size_t>(16, meshes.size() * n_variants);
    out->num_texture_sets = n_sets;
    std::vector<int32_t> opaque_ids, alpha_ids;
    for (size_t k = 0; k < meshes.size(); k++) {
        LhMeshDesc d;
        memset(&d, 0, sizeof(d));
        d.kind = meshes[k].kind;
        d.segments_u = meshes[k].su;
        d.segments_v = meshes[k].sv;
        d.flags = meshes[k].alpha ? LH_MESH_INFO_FLAGS_ALPHA_CLIP : 0u;
        uint32_t set = (uint32_t)(k % n_sets);
        if (meshes[k].alpha) out->alpha_cutout_set = set;
        d.base_color_texture_index = (uint16_t)set;
        d.metallic_roughness_texture_index = (uint16_t)(n_sets + set);
        d.normal_texture_index = (uint16_t)(2 * n_sets + set);
        d.base_color_factor[0] = d.base_color_factor[1] = d.base_color_factor[2] = 1.0f;
        d.uv_repeat = 4.0f;
        int id = lh_scene_add_mesh(scene, &d);
        if (id < 0) return id;
        (meshes[k].alpha ? alpha_ids : opaque_ids).push_back(id);
        for (uint32_t v = 1; v < n_variants; v++) {
            const uint32_t vset = (uint32_t)((k + v * meshes.size()) % n_sets);
            int vid = lh_scene_add_mesh_variant(scene, id, (uint16_t)vset, (uint16_t)(n_sets + vset), (uint16_t)(2 * n_sets + vset), nullptr);
            if (vid < 0) return vid;
            (meshes[k].alpha ? alpha_ids : opaque_ids).push_back(vid);
        }
    }
    // mesh id per instance: every 20th is foliage when the preset has any
    std::vector<int32_t> ids(alpha_ids.empty() ? opaque_ids.size() : 20 * opaque_ids.size());
    for (size_t i = 0; i < ids.size(); i++) {
        if (!alpha_ids.empty() && i % 20 == 19) ids[i] = alpha_ids[(i / 20) % alpha_ids.size()];
        else ids[i] = opaque_ids[i % opaque_ids.size()];
    }
    float ext[3] = {extent * 2.0f, extent * 0.5f, extent};
    int rc = lh_scene_place_instances(scene, count, seed, ext, smin, smax, ids.data(), (uint32_t)ids.size());
    if (rc) return rc;
    // camera outside the -z face of the (wide, flat) box looking at +z: the frustum holds about 40 % of it
    out->camera.position[0] = 0.0f;
    out->camera.position[1] = 0.0f;
    out->camera.position[2] = -extent * 1.15f;
    out->camera.yaw = 1.57079632679f;
    out->camera.pitch = 0.0f;
    return 0;
}

// ---- binding ----------------------------------------------------------------

int lh_host_alloc(void*, uint64_t bytes, uint64_t* addr) {
    if (!addr) return -1;
    void* p = nullptr;
    uint64_t rounded = (bytes + 63) & ~63ull;
    if (rounded == 0) rounded = 64;
    if (posix_memalign(&p, 64, rounded) != 0) return -3;
    memset(p, 0, rounded);
    *addr = (uint64_t)(uintptr_t)p;
    return 0;
}
int lh_host_upload(void*, uint64_t addr, const void* src, uint64_t bytes) {
    if (bytes) memcpy((void*)(uintptr_t)addr, src, bytes);
    return 0;
}
int lh_host_free(void*, uint64_t addr) {
    free((void*)(uintptr_t)addr);
    return 0;
}

struct LhBinding {
    LhFreeFn free_fn = nullptr;
    void* user = nullptr;
    std::vector<uint64_t> allocations;
    std::vector<uint64_t> mesh_info_addresses;
    std::vector<LhInstance> instances;
};

int lh_scene_bind(const LhScene* scene, const LhCapacities* caps, LhAllocFn alloc, LhUploadFn upload, LhFreeFn free_fn,
                  void* user, LhUniforms* addresses, LhBinding** out) {
    if (!scene || !caps || !alloc || !upload || !addresses || !out) return -1;
    auto b = std::make_unique<LhBinding>();
    b->free_fn = free_fn;
    b->user = user;
    int rc = 0;
    auto put = [&](const void* src, uint64_t bytes, uint64_t* addr) {
        // 16 bytes of slack: the device reads streams with 128-bit loads
        if ((rc = alloc(user, bytes + 16, addr))) return false;
        b->allocations.push_back(*addr);
        if (src && bytes && (rc = upload(user, *addr, src, bytes))) return false;
        return true;
    };
    auto fail = [&]() {
        LhBinding* raw = b.release();
        lh_binding_destroy(raw);
        return rc ? rc : -1;
    };
    std::vector<LhMeshInfo> bound_infos;
    for (auto& mp : scene->meshes) {
        const Mesh& m = *mp;
        LhMeshInfo info = m.info;
        if (m.shares_geometry_of >= 0) {   // a material variant: the original's stream addresses, its own MeshInfo buffer
            const LhMeshInfo& g = bound_infos[(size_t)m.shares_geometry_of];
            info.positions = g.positions; info.indices = g.indices; info.normals = g.normals; info.uvs = g.uvs;
            info.micro_indices = g.micro_indices; info.meshlets = g.meshlets;
            uint64_t info_addr = 0;
            if (!put(&info, sizeof(info), &info_addr)) return fail();
            b->mesh_info_addresses.push_back(info_addr);
            bound_infos.push_back(info);
            continue;
        }
        if (!put(m.positions.data(), m.positions.size() * 2, &info.positions)) return fail();
        if (m.wide_indices) {
            if (!put(m.idx32.data(), m.idx32.size() * 4, &info.indices)) return fail();
        } else {
            if (!put(m.idx16.data(), m.idx16.size() * 2, &info.indices)) return fail();
        }
        if (!put(m.normals.data(), m.normals.size(), &info.normals)) return fail();
        if (!put(m.uvs.data(), m.uvs.size() * 2, &info.uvs)) return fail();
        if (!put(m.micro.data(), m.micro.size(), &info.micro_indices)) return fail();
        if (!put(m.meshlets.data(), m.meshlets.size() * sizeof(LhMeshlet), &info.meshlets)) return fail();
        uint64_t info_addr = 0;
        if (!put(&info, sizeof(info), &info_addr)) return fail();  // one buffer per mesh (mesh_loading.cpp:699-708)
        b->mesh_info_addresses.push_back(info_addr);
        bound_infos.push_back(info);
    }
    b->instances.resize(scene->instances.size());
    if (!b->instances.empty()) lh_scene_build_instances(scene, b->mesh_info_addresses.data(), b->instances.data());
    uint64_t draws = (uint64_t)caps->max_opaque_draws + caps->max_alpha_clip_draws;
    if (!put(b->instances.data(), b->instances.size() * sizeof(LhInstance), &addresses->instances)) return fail();
    if (!put(nullptr, 8 * draws * 2, &addresses->meshlet_references)) return fail();               // main.cpp:453-467
    if (!put(nullptr, ((uint64_t)caps->max_instances * 8 + 8) * 4, &addresses->num_meshlets_prefix_sum)) return fail();  // :468-481
    if (!put(nullptr, 8 + 16 * draws, &addresses->draw_calls)) return fail();                      // :506-520
    if (!put(nullptr, 12 * 3, &addresses->dispatches)) return fail();                              // :521-532
    if (!put(nullptr, sizeof(LhMiscStorage), &addresses->misc_storage)) return fail();             // :494-505
    *out = b.release();
    return 0;
}

const LhInstance* lh_binding_instances(const LhBinding* b) { return b ? b->instances.data() : nullptr; }
uint64_t lh_binding_mesh_info_address(const LhBinding* b, int mesh) {
    return (b && mesh >= 0 && mesh < (int)b->mesh_info_addresses.size()) ? b->mesh_info_addresses[mesh] : 0;
}
void lh_binding_destroy(LhBinding* b) {
    if (!b) return;
    if (b->free_fn)
        for (uint64_t a : b->allocations) b->free_fn(b->user, a);
    delete b;
}

uint64_t lh_scene_mesh_triangles(const LhScene* scene, int id, uint32_t* out) {
    if (!scene || id < 0 || id >= (int)scene->meshes.size()) return 0;
    const Mesh& m = *scene->meshes[id];
    if (out) memcpy(out, m.triangles.data(), m.triangles.size() * sizeof(uint32_t));
    return m.triangles.size();
}

int lh_host_has_meshoptimizer(void) { return use_meshoptimizer() ? 1 : 0; }

void lh_compute_bounding_sphere(float result[4], const float* points, uint64_t count) {
    bounding_sphere(result, points, (size_t)count);
}

void lh_compute_meshlet_bounds(const uint32_t* verts, const uint8_t* tris, uint32_t tri_count, const float* positions,
                               LhMeshlet* out) {
    meshlet_bounds(verts, tris, tri_count, positions, out);
}

}  // extern "C"
