/*
 * lh_shared_cpu_gpu.h -- the CPU<->GPU data contract of the geometry pipeline.
 *
 * Restates the layouts and capacity constants of the reference's
 * src/shared_cpu_gpu.h (which is compiled both as C++ against glm and as GLSL
 * with `scalar` block layout).  Nothing here depends on glm: vectors and
 * matrices are plain float arrays, matrices are column-major (m[col][row]),
 * so the structs are byte-identical to the reference's on both sides.
 *
 * Every "uint64_t address" field is a raw device address.  In the reference
 * they are Vulkan buffer device addresses (src/main.cpp:655-670,
 * src/resources/mesh_loading.cpp:660-677); here they are CUDA device pointers
 * returned by lh_buffer_alloc() (include/lighthugger_b200.h).
 *
 * Usable from C, C++ (host, oracle) and CUDA.
 */
#ifndef LH_SHARED_CPU_GPU_H
#define LH_SHARED_CPU_GPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
#define LH_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define LH_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

/* reference: src/shared_cpu_gpu.h:8-9 */
#define LH_MESH_INFO_FLAGS_32_BIT_INDICES 1u
#define LH_MESH_INFO_FLAGS_ALPHA_CLIP 2u

/* reference: src/shared_cpu_gpu.h:119-127 (defaults; the run-time capacities
 * live in LhConfig because BASELINE configs 2-5 exceed them) */
#define LH_MAX_OPAQUE_DRAWS 200000u
#define LH_MAX_ALPHA_CLIP_DRAWS 200000u
#define LH_ALPHA_CLIP_DRAWS_OFFSET LH_MAX_OPAQUE_DRAWS
#define LH_MESHLET_INDICES_BUFFER_SECTION_OFFSET (LH_ALPHA_CLIP_DRAWS_OFFSET + LH_MAX_ALPHA_CLIP_DRAWS)
#define LH_NEAR_PLANE 0.01f

/* reference: src/shared_cpu_gpu.h:148-152 */
#define LH_MAX_MESHLET_UNIQUE_VERTICES 64u
#define LH_MAX_MESHLET_TRIANGLES 124u
#define LH_MAX_MESHLET_VERTICES (LH_MAX_MESHLET_TRIANGLES * 3u)
#define LH_UNUSED_TEXTURE_INDEX 0xFFFFu

/* reference: src/shared_cpu_gpu.h:159-166 */
#define LH_MAX_INSTANCES 1024u
#define LH_PREFIX_SUM_BUFFER_SECTOR_SIZE (LH_MAX_INSTANCES * 8u + 8u)
#define LH_PER_INSTANCE_DISPATCH 0u
#define LH_PER_SHADOW_INSTANCE_DISPATCH 1u
#define LH_PER_MESHLET_DISPATCH 2u

#define LH_NUM_CASCADES 4u
#define LH_SHADOWMAP_SIZE 1024u          /* src/main.cpp:385-409 */
#define LH_MAX_TEXTURES 512u             /* src/descriptor_set.cpp:31-37 */

/* reference: src/shared_cpu_gpu.h:46-51 */
#define LH_UNIFORMS_DEBUG_OFF 0
#define LH_UNIFORMS_DEBUG_CASCADES 1
#define LH_UNIFORMS_DEBUG_TRIANGLE_INDEX 2
#define LH_UNIFORMS_DEBUG_INSTANCE_INDEX 3
#define LH_UNIFORMS_DEBUG_SHADER_CLOCK 4
#define LH_UNIFORMS_DEBUG_NORMALS 5

typedef struct LhMat4 { float m[4][4]; } LhMat4;   /* m[column][row] */
typedef struct LhMat3 { float m[3][3]; } LhMat3;   /* m[column][row] */

/* reference: src/shared_cpu_gpu.h:12-16 (== VkDispatchIndirectCommand) */
typedef struct LhDispatchIndirectCommand {
    uint32_t x, y, z;
} LhDispatchIndirectCommand;

/* reference: src/shared_cpu_gpu.h:18-34 */
typedef struct LhMeshInfo {
    uint64_t positions;      /* u16 x3 per vertex, raw integers */
    uint64_t indices;        /* u16 or u32 (flag bit 0) vertex indices */
    uint64_t normals;        /* i8 x3 per vertex */
    uint64_t uvs;            /* u16 x2 per vertex */
    uint64_t micro_indices;  /* u8 x3 per triangle, meshlet-local */
    uint64_t meshlets;       /* LhMeshlet[] */
    uint16_t num_meshlets;
    uint8_t flags;
    uint8_t _pad0;
    float bounding_sphere[4];
    float texture_scale[2];
    float texture_offset[2];
    uint16_t base_color_texture_index;
    uint16_t metallic_roughness_texture_index;
    uint16_t normal_texture_index;
    uint16_t _pad1;
    float base_color_factor[3];
} LhMeshInfo;

/* reference: src/shared_cpu_gpu.h:36-43 */
typedef struct LhMiscStorage {
    LhMat4 shadow_matrices[4];
    LhMat4 uv_space_shadow_matrices[4];
    LhMat4 shadow_view_matrices[4];
    float shadow_sphere_radii[4];
    uint32_t min_depth;
    uint32_t max_depth;
} LhMiscStorage;

/* reference: src/shared_cpu_gpu.h:53-76 */
typedef struct LhUniforms {
    LhMat4 combined_perspective_view;
    LhMat4 inv_perspective_view;
    LhMat4 view;
    LhMat4 initial_view;
    LhMat4 perspective;
    LhMat4 perspective_inverse;
    LhMat4 view_inverse;
    uint64_t meshlet_references;
    uint64_t instances;
    uint64_t draw_calls;
    uint64_t misc_storage;
    uint64_t num_meshlets_prefix_sum;
    uint64_t dispatches;
    float camera_pos[3];
    float sun_dir[3];
    float sun_intensity[3];
    uint32_t num_instances;
    uint32_t window_size[2];
    float shadow_cam_distance;
    float cascade_split_pow;
    int32_t debug;
    uint32_t debug_shadowmaps;   /* GLSL bool (4 bytes) at offset 556 */
} LhUniforms;

/* reference: src/shared_cpu_gpu.h:79-84 (== VkDrawIndirectCommand) */
typedef struct LhDrawIndirectCommand {
    uint32_t vertexCount;
    uint32_t instanceCount;
    uint32_t firstVertex;
    uint32_t firstInstance;
} LhDrawIndirectCommand;

/* reference: src/shared_cpu_gpu.h:86-99.  normal_transform is
 * mat3(inverse(transform)) WITHOUT a transpose (:92-97). */
typedef struct LhInstance {
    LhMat4 transform;
    uint64_t mesh_info_address;
    LhMat3 normal_transform;
    uint32_t _pad0;
} LhInstance;

/* reference: src/shared_cpu_gpu.h:129-139 */
typedef struct LhMeshlet {
    float cone_apex[3];
    float cone_axis[3];
    float cone_cutoff;
    float bounding_sphere[4];
    uint32_t triangle_offset;  /* byte offset into micro_indices */
    uint32_t index_offset;     /* element offset into indices */
    uint8_t triangle_count;
    uint8_t index_count;
    uint16_t _pad0;
} LhMeshlet;

/* reference: src/shared_cpu_gpu.h:142-145 */
typedef struct LhMeshletReference {
    uint32_t instance_index;
    uint16_t meshlet_index;
    uint16_t _pad0;
} LhMeshletReference;

/* reference: src/shared_cpu_gpu.h:154-157.  A prefix-sum sector is one
 * uint64 counter (hi32 = number of values, lo32 = running sum) followed by
 * the values (src/shaders/common/prefix_sum.glsl:6-9). */
typedef struct LhPrefixSumValue {
    uint32_t index;
    uint32_t sum;
} LhPrefixSumValue;

/* Byte offsets inside the draw-call buffer
 * (src/shaders/common/buffer_references.glsl:51-55): two u32 counters, then
 * the commands; the alpha-clip commands start at index max_opaque_draws. */
#define LH_DRAW_CALLS_NUM_OPAQUE_OFFSET 0u
#define LH_DRAW_CALLS_NUM_ALPHA_CLIP_OFFSET 4u
#define LH_DRAW_CALLS_COMMANDS_OFFSET 8u

/* ---- layout pins (numbers verified against the reference header compiled
 * with glm) ---- */
LH_STATIC_ASSERT(sizeof(LhDispatchIndirectCommand) == 12, "DispatchIndirectCommand");
LH_STATIC_ASSERT(sizeof(LhMeshInfo) == 104, "MeshInfo size");
LH_STATIC_ASSERT(offsetof(LhMeshInfo, meshlets) == 40, "MeshInfo.meshlets");
LH_STATIC_ASSERT(offsetof(LhMeshInfo, num_meshlets) == 48, "MeshInfo.num_meshlets");
LH_STATIC_ASSERT(offsetof(LhMeshInfo, flags) == 50, "MeshInfo.flags");
LH_STATIC_ASSERT(offsetof(LhMeshInfo, bounding_sphere) == 52, "MeshInfo.bounding_sphere");
LH_STATIC_ASSERT(offsetof(LhMeshInfo, texture_scale) == 68, "MeshInfo.texture_scale");
LH_STATIC_ASSERT(offsetof(LhMeshInfo, texture_offset) == 76, "MeshInfo.texture_offset");
LH_STATIC_ASSERT(offsetof(LhMeshInfo, base_color_texture_index) == 84, "MeshInfo.base_color_texture_index");
LH_STATIC_ASSERT(offsetof(LhMeshInfo, metallic_roughness_texture_index) == 86, "MeshInfo.mr_texture_index");
LH_STATIC_ASSERT(offsetof(LhMeshInfo, normal_texture_index) == 88, "MeshInfo.normal_texture_index");
LH_STATIC_ASSERT(offsetof(LhMeshInfo, base_color_factor) == 92, "MeshInfo.base_color_factor");
LH_STATIC_ASSERT(sizeof(LhMiscStorage) == 792, "MiscStorage size");
LH_STATIC_ASSERT(offsetof(LhMiscStorage, uv_space_shadow_matrices) == 256, "MiscStorage.uv");
LH_STATIC_ASSERT(offsetof(LhMiscStorage, shadow_view_matrices) == 512, "MiscStorage.view");
LH_STATIC_ASSERT(offsetof(LhMiscStorage, shadow_sphere_radii) == 768, "MiscStorage.radii");
LH_STATIC_ASSERT(offsetof(LhMiscStorage, min_depth) == 784, "MiscStorage.min_depth");
LH_STATIC_ASSERT(offsetof(LhMiscStorage, max_depth) == 788, "MiscStorage.max_depth");
LH_STATIC_ASSERT(sizeof(LhUniforms) == 560, "Uniforms size");
LH_STATIC_ASSERT(offsetof(LhUniforms, meshlet_references) == 448, "Uniforms.meshlet_references");
LH_STATIC_ASSERT(offsetof(LhUniforms, dispatches) == 488, "Uniforms.dispatches");
LH_STATIC_ASSERT(offsetof(LhUniforms, camera_pos) == 496, "Uniforms.camera_pos");
LH_STATIC_ASSERT(offsetof(LhUniforms, sun_dir) == 508, "Uniforms.sun_dir");
LH_STATIC_ASSERT(offsetof(LhUniforms, sun_intensity) == 520, "Uniforms.sun_intensity");
LH_STATIC_ASSERT(offsetof(LhUniforms, num_instances) == 532, "Uniforms.num_instances");
LH_STATIC_ASSERT(offsetof(LhUniforms, window_size) == 536, "Uniforms.window_size");
LH_STATIC_ASSERT(offsetof(LhUniforms, shadow_cam_distance) == 544, "Uniforms.shadow_cam_distance");
LH_STATIC_ASSERT(offsetof(LhUniforms, cascade_split_pow) == 548, "Uniforms.cascade_split_pow");
LH_STATIC_ASSERT(offsetof(LhUniforms, debug) == 552, "Uniforms.debug");
LH_STATIC_ASSERT(offsetof(LhUniforms, debug_shadowmaps) == 556, "Uniforms.debug_shadowmaps");
LH_STATIC_ASSERT(sizeof(LhDrawIndirectCommand) == 16, "DrawIndirectCommand");
LH_STATIC_ASSERT(sizeof(LhInstance) == 112, "Instance size");
LH_STATIC_ASSERT(offsetof(LhInstance, mesh_info_address) == 64, "Instance.mesh_info_address");
LH_STATIC_ASSERT(offsetof(LhInstance, normal_transform) == 72, "Instance.normal_transform");
LH_STATIC_ASSERT(sizeof(LhMeshlet) == 56, "Meshlet size");
LH_STATIC_ASSERT(offsetof(LhMeshlet, cone_axis) == 12, "Meshlet.cone_axis");
LH_STATIC_ASSERT(offsetof(LhMeshlet, cone_cutoff) == 24, "Meshlet.cone_cutoff");
LH_STATIC_ASSERT(offsetof(LhMeshlet, bounding_sphere) == 28, "Meshlet.bounding_sphere");
LH_STATIC_ASSERT(offsetof(LhMeshlet, triangle_offset) == 44, "Meshlet.triangle_offset");
LH_STATIC_ASSERT(offsetof(LhMeshlet, index_offset) == 48, "Meshlet.index_offset");
LH_STATIC_ASSERT(offsetof(LhMeshlet, triangle_count) == 52, "Meshlet.triangle_count");
LH_STATIC_ASSERT(offsetof(LhMeshlet, index_count) == 53, "Meshlet.index_count");
LH_STATIC_ASSERT(sizeof(LhMeshletReference) == 8, "MeshletReference size");
LH_STATIC_ASSERT(offsetof(LhMeshletReference, meshlet_index) == 4, "MeshletReference.meshlet_index");
LH_STATIC_ASSERT(sizeof(LhPrefixSumValue) == 8, "PrefixSumValue size");
LH_STATIC_ASSERT(LH_PREFIX_SUM_BUFFER_SECTOR_SIZE == 8200, "sector size");
LH_STATIC_ASSERT(LH_MESHLET_INDICES_BUFFER_SECTION_OFFSET == 400000, "section offset");

/* ---- packed visibility word (replaces the reference's
 * R32_UINT `tri << 24 | slot` + separate D32F, src/shaders/common/util.glsl:18-20)
 *
 *   key = (uint64)(~floatBits(depth)) << 32 | slot << 7 | tri
 *
 * depth is reverse-Z (1 at the near plane, ->0 at infinity, GREATER wins in the
 * reference), so ~bits makes "nearest" the numerically smallest key and a
 * 64-bit atomicMin resolves visibility.  Clear value = all ones
 * (depth bits 0 == the reference's depth clear 0.0 == "sky").
 */
#define LH_VIS_CLEAR 0xFFFFFFFFFFFFFFFFull
#define LH_VIS_TRI_BITS 7u
#define LH_VIS_TRI_MASK 0x7Fu
#define LH_VIS_MAX_SLOTS (1u << 25)

#if defined(__CUDACC__)
#define LH_HD __host__ __device__
#else
#define LH_HD
#endif
#ifdef __cplusplus
LH_HD static inline uint64_t lh_vis_pack(uint32_t depth_bits, uint32_t slot, uint32_t tri) {
    return ((uint64_t)(~depth_bits) << 32) | (uint64_t)((slot << LH_VIS_TRI_BITS) | tri);
}
LH_HD static inline uint32_t lh_vis_depth_bits(uint64_t key) { return ~(uint32_t)(key >> 32); }
LH_HD static inline uint32_t lh_vis_slot(uint64_t key) { return (uint32_t)(key & 0xFFFFFFFFull) >> LH_VIS_TRI_BITS; }
LH_HD static inline uint32_t lh_vis_tri(uint64_t key) { return (uint32_t)key & LH_VIS_TRI_MASK; }
#endif

#endif /* LH_SHARED_CPU_GPU_H */
