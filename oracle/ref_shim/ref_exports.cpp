// C exports around the reference's own scene-prep code, compiled IN PLACE from
// /root/reference (never copied): src/resources/meshlets.cpp (build_meshlets),
// src/resources/bounding_sphere.h (computeBoundingSphere) and the vendored
// meshoptimizer v0.20.  Output: oracle/_ref/libref_sceneprep.so.  Used by
// tests (and tests/golden/make_golden.py) to pin this repo's scene
// preparation against the reference.  TEST INFRASTRUCTURE.
#include "ref_pch.h"

#include "resources/meshlets.cpp"            // -I/root/reference/src
#include "resources/bounding_sphere.h"

extern "C" {

// Runs the reference's build_meshlets on float positions (xyz) + u32 indices.
// Call once with all outputs NULL to get the sizes, then again with buffers.
__attribute__((visibility("default"))) int ref_build_meshlets(
    const uint32_t* indices, uint64_t index_count, const float* positions, uint64_t vertex_count,
    uint64_t* num_meshlets, uint64_t* micro_bytes, uint64_t* num_vertex_indices,
    void* meshlets_out /* 56 B each */, uint8_t* micro_out, uint16_t* vertex_indices_out) {
    static Meshlets cache;
    static const uint32_t* cache_key = nullptr;
    if (cache_key != indices || cache.meshlets.empty()) {
        std::vector<uint32_t> idx(indices, indices + index_count);
        std::vector<float> pos(positions, positions + vertex_count * 3);
        bool wide = vertex_count > 65536;
        if (wide) return -1;  // fixtures use 16-bit meshes
        std::vector<uint16_t> idx16(idx.begin(), idx.end());
        cache = build_meshlets(reinterpret_cast<uint8_t*>(idx16.data()), index_count, pos.data(), vertex_count, false);
        cache_key = indices;
    }
    static_assert(sizeof(Meshlet) == 56, "Meshlet layout");
    if (num_meshlets) *num_meshlets = cache.meshlets.size();
    if (micro_bytes) *micro_bytes = cache.micro_indices.size();
    if (num_vertex_indices) *num_vertex_indices = cache.indices_16bit.size();
    if (meshlets_out) memcpy(meshlets_out, cache.meshlets.data(), cache.meshlets.size() * sizeof(Meshlet));
    if (micro_out) memcpy(micro_out, cache.micro_indices.data(), cache.micro_indices.size());
    if (vertex_indices_out) memcpy(vertex_indices_out, cache.indices_16bit.data(), cache.indices_16bit.size() * 2);
    return 0;
}

__attribute__((visibility("default"))) void ref_compute_bounding_sphere(float result[4], const float* points, uint64_t count) {
    computeBoundingSphere(result, points, (size_t)count);
}

__attribute__((visibility("default"))) void ref_compute_meshlet_bounds(const uint32_t* meshlet_vertices, const uint8_t* meshlet_triangles,
                                                                    uint64_t triangle_count, const float* positions,
                                                                    uint64_t vertex_count, float out[11]) {
    meshopt_Bounds b = meshopt_computeMeshletBounds(meshlet_vertices, meshlet_triangles, (size_t)triangle_count, positions,
                                                    (size_t)vertex_count, sizeof(float) * 3);
    memcpy(out + 0, b.cone_apex, 12);
    memcpy(out + 3, b.cone_axis, 12);
    out[6] = b.cone_cutoff;
    memcpy(out + 7, b.center, 12);
    out[10] = b.radius;
}

// struct sizes of the reference header as compiled here (layout pin)
__attribute__((visibility("default"))) void ref_struct_sizes(uint32_t out[9]) {
    out[0] = sizeof(Instance); out[1] = sizeof(MeshInfo); out[2] = sizeof(Meshlet); out[3] = sizeof(MeshletReference);
    out[4] = sizeof(PrefixSumValue); out[5] = sizeof(DrawIndirectCommand); out[6] = sizeof(DispatchIndirectCommand);
    out[7] = sizeof(MiscStorage); out[8] = sizeof(Uniforms);
}

}  // extern "C"
