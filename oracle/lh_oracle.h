/*
 * lh_oracle.h -- CPU restatement of the reference's geometry pipeline.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under lighthugger_b200/ may include,
 * link or call this: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker / CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" by the reference -- the reference ships no
 * tests, golden vectors or fixtures for this path and
 * its GLSL cannot be executed in this image (no Vulkan loader / lavapipe /
 * glslc).  What IS pinned against the reference: struct layouts and constants
 * (include/lh_shared_cpu_gpu.h static_asserts), and the scene-prep inputs
 * (meshlet bounds, bounding sphere) against the reference's own
 * src/resources/meshlets.cpp + vendored meshoptimizer compiled into
 * oracle/_ref (see oracle/Makefile, tests/golden/).
 *
 * The functions operate on the same byte layouts as the device path; every
 * "address" field of LhUniforms / LhInstance / LhMeshInfo is a HOST pointer
 * here.  Each function cites the reference file:line it follows.
 *
 * Pinned arithmetic model (the reference leaves these implementation-defined,
 * SURVEY.md 8c): IEEE-754 binary32, round-to-nearest-even, no FMA contraction
 * (build with -ffp-contract=off), mat4*vec4 = (c0*x + c1*y) + (c2*z + c3*w),
 * mat3*vec3 = (c0*x + c1*y) + c2*z, dot = (ax*bx + ay*by) + az*bz,
 * length = sqrt(dot(v,v)), normalize(v) = v * (1 / length(v)) (glm's, detail/func_geometric.inl:82-91),
 * appends in ascending invocation order.
 */
#ifndef LH_ORACLE_H
#define LH_ORACLE_H

#include "lh_shared_cpu_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct LhoTexture {
    const uint8_t* rgba8;   /* all mips, tightly packed, largest first */
    uint32_t width, height, mip_levels;
    uint32_t srgb;          /* decode sRGB -> linear on fetch (reference registers every BC texture as sRGB) */
} LhoTexture;

typedef struct LhoConfig {
    uint32_t max_instances;
    uint32_t max_opaque_draws;
    uint32_t max_alpha_clip_draws;
    uint32_t shadowmap_size;
    uint32_t threads;            /* 0 = hardware_concurrency */
    uint32_t instance_first;     /* sort-last partition: instances [first, first+count) */
    uint32_t instance_count;     /* 0xFFFFFFFF = all */
    const LhoTexture* textures;  /* bindless table, may be NULL */
    uint32_t num_textures;
    const float* skybox;         /* 6 faces RGBA32F or NULL */
    uint32_t skybox_size;
} LhoConfig;

void lho_config_default(LhoConfig* cfg);

/* reset_buffers_a (compute/reset_buffers.glsl:9-29) + reset_counters (common/util.glsl:61-70) */
int lho_reset(const LhUniforms* u, const LhoConfig* cfg);
/* cull_instances (cull_instances.comp:7-27) + prefix_sum_inclusive_append (common/prefix_sum.glsl:17-30),
 * appends in ascending instance order */
int lho_cull_instances(const LhUniforms* u, const LhoConfig* cfg);
/* reset_buffers_b + write_draw_calls (write_draw_calls.comp:7-57) incl. get_meshlet_reference
 * (common/util.glsl:42-59) and prefix_sum_binary_search (common/prefix_sum.glsl:33-46) */
int lho_cull_meshlets(const LhUniforms* u, const LhoConfig* cfg);
/* visbuffer_opaque_* / visbuffer_alpha_clip_* (rasterization.glsl:38-97) + fixed-function state
 * (pipelines.cpp:10-25,67-71; rendering.cpp:195-266).  vis holds width*height packed keys and is cleared first. */
int lho_raster_visbuffer(const LhUniforms* u, const LhoConfig* cfg, uint64_t* vis);
/* read_depth (compute/read_depth.glsl:22-83) */
int lho_reduce_depth(const LhUniforms* u, const LhoConfig* cfg, const uint64_t* vis);
/* generate_shadow_matrices (compute/generate_shadow_matrices.glsl:15-110) */
int lho_generate_shadow_matrices(const LhUniforms* u, const LhoConfig* cfg);
/* cull_instances_shadows (cull_instances.comp:31-56) */
int lho_cull_instances_shadows(const LhUniforms* u, const LhoConfig* cfg);
/* reset_buffers_c + write_draw_calls_shadows (write_draw_calls.comp:61-117) + shadowmap_* raster
 * (rasterization.glsl:101-145; pipelines.cpp:15-19,73-77; rendering.cpp:379-436).
 * shadow holds 4 layers of shadowmap_size^2 floats; layer `cascade` is cleared to 1.0 and rendered. */
int lho_shadow_cascade(const LhUniforms* u, const LhoConfig* cfg, uint32_t cascade, float* shadow);
/* render_geometry main (render_geometry.comp:70-266); out = width*height RGBA32F (the device stores RGBA16F) */
int lho_resolve(const LhUniforms* u, const LhoConfig* cfg, const uint64_t* vis, const float* shadow, float* out_rgba32f);
/* one input of the lighting resolve instead of the colour (100 shadow term, 101 base colour, 102 normal, 103 roughness /
 * metallic / cascade, 104 world position, 105 uv): attribution of differences, not part of the reference */
int lho_resolve_debug(const LhUniforms* u, const LhoConfig* cfg, const uint64_t* vis, const float* shadow, int debug_term, float* out_rgba32f);
/* the texture unit's bilinear filter (level 0 of an RGBA8 texture, repeat addressing) at `count` positions: replayed
 * against dumps of the real texture unit (tests/golden/texture_unit_*.bin) */
void lho_test_sample_bilinear(const uint8_t* rgba8, uint32_t width, uint32_t height, uint32_t srgb, uint64_t count, const float* uv,
                              float* out_rgba);
/* textureLod(sampler3D, uvw, 0) of a size^3 RGB32F table with clamp addressing at `count` positions */
void lho_test_sample_lut(const float* lut_rgb32f, uint32_t size, uint64_t count, const float* uvw, float* out_rgb);
/* textureLod(samplerCube, direction, 0) of a size x size RGBA32F cube map at `count` directions (out: rgb) */
void lho_test_sample_cube(const float* faces_rgba32f, uint32_t size, uint64_t count, const float* xyz, float* out_rgb);
/* textureLod(level 0) with repeat addressing at arbitrary coordinates (wraps into [0, 1) first, then filters) */
void lho_test_sample_lod0(const uint8_t* rgba8, uint32_t width, uint32_t height, uint32_t srgb, uint64_t count, const float* uv,
                          float* out_rgba);
/* the lit colour of `count` listed pixels (y * width + x) with `bias` added to the rgb of every material texture sample:
 * the conditioning of the reference's own result at those pixels (not part of the reference) */
int lho_resolve_pixels(const LhUniforms* u, const LhoConfig* cfg, const uint64_t* vis, const float* shadow, uint64_t count,
                       const uint32_t* pixels, const float bias[3], float* out_rgba32f);
/* display_transform (display_transform.comp:5-81, debug overlay off): scene RGBA32F + LUT (size^3 RGB32F) -> the
 * sRGB-encoded RGBA32F value stored to the UNORM swapchain image */
int lho_display_transform(const LhUniforms* u, const LhoConfig* cfg, const float* scene_rgba32f, const float* lut_rgb32f, uint32_t lut_size,
                          float* out_rgba32f);
/* steps 1-11 of rendering.cpp (reset ... resolve); out_rgba32f may be NULL to skip the resolve */
int lho_render_frame(const LhUniforms* u, const LhoConfig* cfg, uint64_t* vis, float* shadow, float* out_rgba32f);

/* sort-last composite: dst[i] = min(dst[i], src[i]) (new) */
void lho_composite_min(uint64_t* dst, const uint64_t* src, uint64_t count);

/* single culling decisions (1 = culled) on caller-supplied inputs, for the random-input comparison with the reference's GLSL */
int lho_test_cull_sphere(const LhUniforms* u, const LhInstance* inst, const float sphere[4]);
int lho_test_cull_sphere_shadows(const LhUniforms* u, const LhInstance* inst, const float sphere[4], uint32_t cascade);
int lho_test_cull_cone_perspective(const LhUniforms* u, const LhInstance* inst, const LhMeshlet* m);
int lho_test_cull_cone_orthographic(const LhUniforms* u, const LhInstance* inst, const LhMeshlet* m);

uint32_t lho_test_clip_positions(const LhUniforms* u, int cascade, uint32_t slot, float* out_xyzw);

/* pinned pow used by generate_shadow_matrices (double arithmetic, identical op sequence on host and device) */
float lho_pinned_powf(float x, float p);

#ifdef __cplusplus
}
#endif
#endif
