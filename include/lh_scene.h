/*
 * lh_scene.h -- C interface of the host-side scene preparation library
 * (liblh_host.so): procedural meshes, quantisation, meshlet build + bounds,
 * instance placement and the per-frame Uniforms derivation.
 *
 * This is the INPUT side of the hot path (SURVEY.md 3.4 / 8d "Synthetic
 * inputs"): it produces buffers in exactly the layouts of
 * lh_shared_cpu_gpu.h so that either the CUDA layer (device addresses) or the
 * CPU oracle (host addresses) can consume them.  It replaces, for synthetic
 * scenes, what the reference does at load time in
 * src/resources/mesh_loading.cpp:584-697 (quantised streams, MeshInfo fill),
 * src/resources/meshlets.cpp:3-140 (meshlet build + pack),
 * src/resources/bounding_sphere.h:3-74, src/main.cpp:433-440 (Instance
 * construction) and src/main.cpp:734-760 (Uniforms derivation).
 */
#ifndef LH_SCENE_H
#define LH_SCENE_H

#include "lh_shared_cpu_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define LH_HOST_API __attribute__((visibility("default")))

typedef struct LhScene LhScene;

enum { LH_MESH_TORUS = 0, LH_MESH_SPHERE = 1, LH_MESH_PLANE = 2, LH_MESH_BOX = 3 };

typedef struct LhMeshDesc {
    int32_t kind;          /* LH_MESH_* */
    uint32_t segments_u;   /* torus: around the main ring; sphere: longitude; plane/box: x */
    uint32_t segments_v;   /* torus: around the tube; sphere: latitude; plane/box: y */
    uint32_t flags;        /* LH_MESH_INFO_FLAGS_ALPHA_CLIP to mark foliage-like meshes */
    uint16_t base_color_texture_index;
    uint16_t metallic_roughness_texture_index;
    uint16_t normal_texture_index;
    uint16_t _pad;
    float base_color_factor[3];
    float uv_repeat;       /* texture_scale = uv_repeat / 65535 */
} LhMeshDesc;

/* Read-only view of one prepared mesh.  info holds everything but the six
 * stream addresses (left 0): the binder fills those. */
typedef struct LhMeshView {
    const uint16_t* positions;  uint64_t num_vertices;       /* u16 x3 */
    const int8_t* normals;                                   /* i8 x3 */
    const uint16_t* uvs;                                     /* u16 x2 */
    const void* indices;        uint64_t num_indices;  uint32_t index_bytes; /* 2 or 4 */
    const uint8_t* micro_indices; uint64_t micro_indices_bytes;
    const LhMeshlet* meshlets;  uint64_t num_meshlets;
    uint64_t num_triangles;
    LhMeshInfo info;
    float dequant_scale[3];     /* object = q * scale + offset */
    float dequant_offset[3];
} LhMeshView;

typedef struct LhCamera {
    float position[3];
    float yaw, pitch;           /* facing = (cos yaw cos pitch, sin pitch, sin yaw cos pitch), src/input.h:31-37 */
    float fov_degrees;          /* 45, src/main.cpp:606-613 */
    float sun_latitude, sun_longitude; /* -5.6, 1.37 */
    float shadow_cam_distance;  /* 1024, src/main.cpp:653 */
    float cascade_split_pow;    /* 3, src/main.cpp:654 */
} LhCamera;

LH_HOST_API int lh_scene_create(LhScene** out);
LH_HOST_API void lh_scene_destroy(LhScene* scene);

/* meshes */
LH_HOST_API int lh_scene_add_mesh(LhScene* scene, const LhMeshDesc* desc);          /* returns mesh id or <0 */
/* pre-built streams (e.g. the golden reference-meshletised torus in tests/golden) */
LH_HOST_API int lh_scene_add_mesh_raw(LhScene* scene, const LhMeshView* view);
LH_HOST_API int lh_scene_num_meshes(const LhScene* scene);
LH_HOST_API int lh_scene_mesh_view(const LhScene* scene, int mesh, LhMeshView* out);

/* instances: counter-based splitmix64 placement (seed 0x5EED0001 by default):
 * position uniform in [-extent, extent]^3 (per axis), uniform scale in
 * [scale_min, scale_max], a half-turn about a uniformly random axis; mesh id cycles through
 * mesh_ids[i % num_mesh_ids].  transform = T * R * S * dequant(mesh). */
LH_HOST_API int lh_scene_place_instances(LhScene* scene, uint32_t count, uint64_t seed,
                                         const float extent[3], float scale_min, float scale_max,
                                         const int32_t* mesh_ids, uint32_t num_mesh_ids);
/* the same with uniformly random rotations and a different scale per axis (parity tests: general transforms) */
LH_HOST_API int lh_scene_place_instances_general(LhScene* scene, uint32_t count, uint64_t seed,
                                                 const float extent[3], float scale_min, float scale_max,
                                                 const int32_t* mesh_ids, uint32_t num_mesh_ids);
LH_HOST_API int lh_scene_add_instance(LhScene* scene, int mesh, const float translation[3],
                                      const float rotation_quat_xyzw[4], float scale);
LH_HOST_API uint32_t lh_scene_num_instances(const LhScene* scene);
LH_HOST_API uint64_t lh_scene_total_triangles(const LhScene* scene);   /* sum over instances of mesh triangle counts */
LH_HOST_API uint64_t lh_scene_total_meshlets(const LhScene* scene);
/* Build the Instance array exactly as src/shared_cpu_gpu.h:92-97 does
 * (normal_transform = mat3(inverse(transform)), no transpose);
 * mesh_info_addresses[mesh] is where each mesh's LhMeshInfo lives. */
LH_HOST_API int lh_scene_build_instances(const LhScene* scene, const uint64_t* mesh_info_addresses,
                                         LhInstance* out);

/* per-frame uniforms (src/main.cpp:734-760, src/projection.h:3-20); buffer
 * addresses are left 0 for the binder. */
LH_HOST_API int lh_scene_uniforms(const LhScene* scene, const LhCamera* camera,
                                  uint32_t width, uint32_t height, LhUniforms* out);
LH_HOST_API void lh_camera_default(LhCamera* camera);

/* ---- BASELINE.json configs as presets (SURVEY.md 8d "Synthetic inputs").
 * Adds the meshes and seeded instances to an empty scene and fills camera /
 * resolution.  instance_count 0 = the preset's own count. ---- */
enum {
    LH_PRESET_TINY = 0,      /* 64 x torus 32x16 (1 024 tri), 256x144: unit tests */
    LH_PRESET_C1 = 1,        /* 1 000 x torus 125x80 (20 000 tri), 1920x1080 */
    LH_PRESET_C2 = 2,        /* 10 000 x torus 250x200 (100 000 tri), 3840x2160 */
    LH_PRESET_C4 = 4,        /* ~100 M tri mixed meshes, 5 % alpha-clip foliage, 3840x2160 */
    LH_PRESET_C5 = 5         /* 1 G tri stress: 10 000 x torus 250x200, 7680x4320 */
};
typedef struct LhPresetInfo {
    uint32_t width, height;
    uint32_t num_texture_sets;  /* textures: base v, metallic-roughness n+v, normal 2n+v (v < n) */
    uint32_t alpha_cutout_set;  /* texture set whose base colour carries alpha cut-outs (or 0xFFFFFFFF) */
    LhCamera camera;
} LhPresetInfo;
/* 1 when meshlets are built by the reference's vendored meshoptimizer (meshopt_buildMeshlets, 64 / 124 / 0.25, as
 * src/resources/meshlets.cpp:10-57), 0 when the built-in greedy fallback is in use (library built without
 * /root/reference, or LH_MESHLET_BUILDER=greedy in the environment) */
LH_HOST_API int lh_host_has_meshoptimizer(void);
/* the triangle list (3 vertex indices per triangle, in the order handed to the meshlet builder) of a mesh built by
 * lh_scene_add_mesh / lh_scene_add_mesh_indexed; returns the index count (call with out NULL to query) */
LH_HOST_API uint64_t lh_scene_mesh_triangles(const LhScene* scene, int mesh, uint32_t* out);
LH_HOST_API int lh_scene_build_preset(LhScene* scene, int preset, uint32_t instance_count, uint64_t seed, LhPresetInfo* out);

/* procedural textures: RGBA8 with a full box-filtered mip chain, tightly
 * packed largest first.  kind: 0 base colour (with alpha cut-outs if
 * alpha_cutout), 1 metallic-roughness, 2 normal map.  Returns bytes written
 * (call with dst NULL to query). */
LH_HOST_API uint64_t lh_scene_make_texture(int kind, uint32_t size, uint32_t variant, int alpha_cutout,
                                           uint8_t* dst, uint32_t* mip_levels);
/* BC7 (mode 6) encoding of one RGBA8 level; returns bytes (call with dst NULL to query) */
LH_HOST_API uint64_t lh_bc7_encode(const uint8_t* rgba8, uint32_t width, uint32_t height, uint8_t* dst);
/* procedural sky: 6 faces of size^2 RGBA32F */
LH_HOST_API uint64_t lh_scene_make_skybox(uint32_t size, float* dst);

/* ---- binding: place the scene in an address space.  alloc/upload/free have
 * the signatures of lh_buffer_alloc / lh_upload / lh_buffer_free
 * (include/lighthugger_b200.h) with the context as `user`, so the CUDA layer's
 * exports can be passed directly; lh_host_alloc / lh_host_upload / lh_host_free
 * place it in host memory instead (what the CPU oracle consumes).  Replaces the
 * buffer creation + address wiring of src/main.cpp:442-532,646-670 and
 * src/resources/mesh_loading.cpp:660-708. ---- */
typedef int (*LhAllocFn)(void* user, uint64_t bytes, uint64_t* addr);
typedef int (*LhUploadFn)(void* user, uint64_t addr, const void* src, uint64_t bytes);
typedef int (*LhFreeFn)(void* user, uint64_t addr);

typedef struct LhCapacities {
    uint32_t max_instances;         /* values per prefix-sum sector */
    uint32_t max_opaque_draws;
    uint32_t max_alpha_clip_draws;
} LhCapacities;

typedef struct LhBinding LhBinding;
/* Uploads every mesh stream, one LhMeshInfo per mesh, the instance array and
 * allocates (zeroed) the five pipeline buffers; writes the six buffer
 * addresses into `addresses` (other fields untouched). */
LH_HOST_API int lh_scene_bind(const LhScene* scene, const LhCapacities* caps, LhAllocFn alloc, LhUploadFn upload,
                              LhFreeFn free_fn, void* user, LhUniforms* addresses, LhBinding** out);
LH_HOST_API const LhInstance* lh_binding_instances(const LhBinding* b);   /* host copy, as uploaded */
LH_HOST_API uint64_t lh_binding_mesh_info_address(const LhBinding* b, int mesh);
LH_HOST_API void lh_binding_destroy(LhBinding* b);                        /* frees through free_fn */
LH_HOST_API int lh_host_alloc(void* user, uint64_t bytes, uint64_t* addr);
LH_HOST_API int lh_host_upload(void* user, uint64_t addr, const void* src, uint64_t bytes);
LH_HOST_API int lh_host_free(void* user, uint64_t addr);

/* ---- glTF ingestion: one KHR_mesh_quantization primitive as the reference reads it
 * (src/resources/mesh_loading.cpp:400-597): POSITION u16 x3 padded to 8 bytes, NORMAL i8 x3 padded to 4, TEXCOORD_0
 * u16 x2, a u16 / u32 triangle list; the material's texture slots, KHR_texture_transform scale / offset (already
 * divided by 65535 for normalised uvs, :555-557), MASK -> LH_MESH_INFO_FLAGS_ALPHA_CLIP and the base colour factor.
 * Compacts the streams to 3 components (copy_quantized_positions / _normals), builds meshlets + bounds and the mesh
 * bounding sphere from the raw integer positions, and takes 32-bit indices from index_bytes as the reference does. ---- */
typedef struct LhIndexedMesh {
    const void* positions;  uint32_t position_stride;   /* >= 6 */
    const void* normals;    uint32_t normal_stride;     /* >= 3 */
    const void* uvs;        uint32_t uv_stride;         /* >= 4 */
    uint64_t num_vertices;
    const void* indices;    uint64_t num_indices;  uint32_t index_bytes;   /* 2 or 4; num_indices % 3 == 0 */
    uint32_t flags;                                      /* LH_MESH_INFO_FLAGS_ALPHA_CLIP or 0 */
    float texture_scale[2], texture_offset[2];
    uint16_t base_color_texture_index, metallic_roughness_texture_index, normal_texture_index, _pad;
    float base_color_factor[3];
} LhIndexedMesh;
LH_HOST_API int lh_scene_add_mesh_indexed(LhScene* scene, const LhIndexedMesh* mesh);        /* returns mesh id or <0 */
/* an instance with the node's matrix as it is (column-major; the quantisation is already part of a glTF node's
 * scale / translation): Instance(transform, mesh_info) of src/main.cpp:433-440 */
LH_HOST_API int lh_scene_add_instance_matrix(LhScene* scene, int mesh, const float transform[16]);
/* a material variant of `mesh`: its own MeshInfo (texture indices, colour factor; NULL keeps the factor) over the same
 * vertex / index / meshlet buffers; returns the new mesh id */
LH_HOST_API int lh_scene_add_mesh_variant(LhScene* scene, int mesh, uint16_t base_color_texture_index,
                                          uint16_t metallic_roughness_texture_index, uint16_t normal_texture_index,
                                          const float base_color_factor[3]);

/* ---- the reference's meshlet cache (src/resources/fs_cache.h:3-25, src/resources/mesh_loading.cpp:602-649) ----
 * Three raw blobs per primitive under <dir>/<std::hash of "<name> meshlets" | "<name> indices" | "<name> micro indices">:
 * the Meshlet records, the per-meshlet vertex indices (u16 / u32) and the micro-indices.  which: 0, 1, 2 in that order. */
typedef struct LhMeshletBlobs LhMeshletBlobs;
LH_HOST_API int lh_meshlet_cache_path(const char* dir, const char* primitive_name, int which, char* out, uint64_t out_bytes);
LH_HOST_API int lh_meshlet_cache_store(const char* dir, const char* primitive_name, const LhMeshView* mesh);
/* 0: hit (*out owns the blobs); 1: miss (build and store); < 0: unreadable or inconsistent blobs */
LH_HOST_API int lh_meshlet_cache_load(const char* dir, const char* primitive_name, uint32_t index_bytes, LhMeshletBlobs** out);
/* fills the meshlet / index / micro-index fields (and num_triangles) of a view; the vertex streams are the caller's */
LH_HOST_API int lh_meshlet_blobs_view(const LhMeshletBlobs* blobs, uint32_t index_bytes, LhMeshView* view);
LH_HOST_API void lh_meshlet_blobs_free(LhMeshletBlobs* blobs);

/* stand-alone pieces, exported for tests against oracle/_ref */
LH_HOST_API void lh_compute_bounding_sphere(float result[4], const float* points, uint64_t count);
LH_HOST_API void lh_compute_meshlet_bounds(const uint32_t* meshlet_vertices, const uint8_t* meshlet_triangles,
                                           uint32_t triangle_count, const float* positions /* xyz */,
                                           LhMeshlet* out /* cone + sphere fields only */);

#ifdef __cplusplus
}
#endif
#endif /* LH_SCENE_H */
