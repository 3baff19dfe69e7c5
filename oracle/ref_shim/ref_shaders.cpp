// The reference's own GLSL, compiled as C++ from where it lies under /root/reference (through
// oracle/ref_shim/glsl_filter.py -> oracle/_ref/gen/, see the Makefile) and run on the host, one invocation at a
// time in ascending gl_GlobalInvocationID order, over the same host-bound buffers the oracle uses.
// Output: oracle/_ref/libref_shaders.so.  This is what pins oracle/lh_oracle*.cpp: tests/test_ref_shaders.py runs
// both on the same scenes and compares the buffers.  TEST INFRASTRUCTURE; never linked into the product.
//
// Each shader "translation unit" gets its own namespace, because every .comp/.glsl of the reference includes the
// common files again.
#include "glsl_shim.h"

#define REF_EXPORT extern "C" __attribute__((visibility("default")))

namespace refsh {
namespace culling_tu {   // cull_instances.comp + write_draw_calls.comp (both include culling.glsl and util.glsl)
#include "gen/common/prefix_sum.glsl"
#include "gen/common/util.glsl"
#include "gen/common/culling.glsl"
#include "gen/cull_instances.comp"
#include "gen/write_draw_calls.comp"
}  // namespace culling_tu

namespace reset_tu {
#include "gen/common/prefix_sum.glsl"
#include "gen/common/util.glsl"
#include "gen/compute/reset_buffers.glsl"
}  // namespace reset_tu

namespace shadow_matrices_tu {
#include "gen/common/prefix_sum.glsl"
#include "gen/common/matrices.glsl"
#include "gen/common/util.glsl"
#include "gen/compute/generate_shadow_matrices.glsl"
}  // namespace shadow_matrices_tu

namespace raster_tu {   // rasterization.glsl: the vertex shaders (the fragment stages only pass `packed` through / discard)
// stage interface variables (layout(location = ...) in/out)
thread_local uint32_t packed, out_packed, base_texture_index;
thread_local vec2 uv;
inline float texture_alpha(sampler2D, vec2) { return 1.0f; }
struct AlphaOnly { float a; };
inline AlphaOnly texture(sampler2D, vec2) { return AlphaOnly{1.0f}; }   // the alpha test itself is exercised through the oracle's raster
#include "gen/common/prefix_sum.glsl"
#include "gen/common/util.glsl"
#include "gen/rasterization.glsl"
}  // namespace raster_tu

namespace depth_tu {
#include "gen/compute/read_depth.glsl"
}  // namespace depth_tu

namespace resolve_tu {
// GLSL passes `const` arrays to unqualified array parameters (evaluate_spherical_harmonic(normal, AMBIENT_SH)); C++ does not
#define const
#include "gen/common/debug.glsl"
#include "gen/common/pbr.glsl"
#include "gen/common/prefix_sum.glsl"
#include "gen/common/util.glsl"
#include "gen/common/vbuffer.glsl"
#define main render_geometry_main
#include "gen/render_geometry.comp"
#undef main
#undef const
}  // namespace resolve_tu

namespace display_tu {
#define const
#include "gen/display_transform.comp"
#undef const
}  // namespace display_tu

// vkCmdDispatch(groups) of a shader with the given local size: every invocation, ascending (x fastest)
template <class F>
void dispatch(uint32_t gx, uint32_t gy, uint32_t lx, uint32_t ly, F&& entry) {
    for (uint32_t y = 0; y < gy * ly; y++)
        for (uint32_t x = 0; x < gx * lx; x++) {
            gl_GlobalInvocationID.set(x, y, 0);
            entry();
        }
}

uint32_t groups(uint32_t width, uint32_t wg) { return ((width == 0 ? 1u : width - 1u) / wg) + 1u; }   // src/util.cpp:56-58 guarded

}  // namespace refsh

using namespace refsh;

// ---- whole passes, recorded like src/rendering.cpp does (indirect dispatch sizes are read from the buffer the
// ---- reset shaders wrote, as vkCmdDispatchIndirect would) ----
static const DispatchIndirectCommand& indirect(const Uniforms* u, uint32_t which) {
    return reinterpret_cast<const DispatchIndirectCommand*>(u->dispatches)[which];
}

REF_EXPORT void ref_sh_reset_buffers_a(const Uniforms* u) {   // rendering.cpp:79
    uniform_buffer.address = (uint64_t)u;
    dispatch(1, 1, 1, 1, [] { reset_tu::reset_buffers_a(); });
}

REF_EXPORT void ref_sh_cull_instances(const Uniforms* u) {   // rendering.cpp:140-168, cull_instances.comp:5
    uniform_buffer.address = (uint64_t)u;
    dispatch(indirect(u, PER_INSTANCE_DISPATCH).x, 1, 64, 1, [] { culling_tu::cull_instances(); });
}

REF_EXPORT void ref_sh_write_draw_calls(const Uniforms* u) {   // rendering.cpp:172-193
    uniform_buffer.address = (uint64_t)u;
    dispatch(1, 1, 1, 1, [] { reset_tu::reset_buffers_b(); });
    dispatch(indirect(u, PER_MESHLET_DISPATCH).x, 1, 64, 1, [] { culling_tu::write_draw_calls(); });
}

REF_EXPORT void ref_sh_generate_shadow_matrices(const Uniforms* u) {   // rendering.cpp:305-322, local_size_x = 4
    uniform_buffer.address = (uint64_t)u;
    dispatch(1, 1, 4, 1, [] { shadow_matrices_tu::generate_shadow_matrices(); });
}

REF_EXPORT void ref_sh_cull_instances_shadows(const Uniforms* u) {   // rendering.cpp:326-352, local 16 x 4
    uniform_buffer.address = (uint64_t)u;
    dispatch(indirect(u, PER_SHADOW_INSTANCE_DISPATCH).x, 1, 16, 4, [] { culling_tu::cull_instances_shadows(); });
}

REF_EXPORT void ref_sh_write_draw_calls_shadows(const Uniforms* u, uint32_t cascade) {   // rendering.cpp:355-377
    uniform_buffer.address = (uint64_t)u;
    shadow_constant.cascade_index = cascade;
    dispatch(1, 1, 1, 1, [] { reset_tu::reset_buffers_c(); });
    dispatch(indirect(u, PER_MESHLET_DISPATCH).x, 1, 64, 1, [] { culling_tu::write_draw_calls_shadows(); });
}

// ---- single functions on caller-supplied inputs (random-input tests) ----
REF_EXPORT int ref_sh_cull_bounding_sphere(const Uniforms* u, const Instance* instance, const float sphere[4]) {
    uniform_buffer.address = (uint64_t)u;
    return culling_tu::cull_bounding_sphere(*instance, vec4(sphere[0], sphere[1], sphere[2], sphere[3])) ? 1 : 0;
}
REF_EXPORT int ref_sh_cull_bounding_sphere_shadows(const Uniforms* u, const Instance* instance, const float sphere[4], uint32_t cascade) {
    uniform_buffer.address = (uint64_t)u;
    return culling_tu::cull_bounding_sphere_shadows(*instance, vec4(sphere[0], sphere[1], sphere[2], sphere[3]), cascade) ? 1 : 0;
}
REF_EXPORT int ref_sh_cull_cone_perspective(const Uniforms* u, const Instance* instance, const Meshlet* meshlet) {
    uniform_buffer.address = (uint64_t)u;
    return culling_tu::cull_cone_perspective(*instance, *meshlet) ? 1 : 0;
}
REF_EXPORT int ref_sh_cull_cone_orthographic(const Uniforms* u, const Instance* instance, const Meshlet* meshlet) {
    uniform_buffer.address = (uint64_t)u;
    return culling_tu::cull_cone_orthographic(*instance, *meshlet) ? 1 : 0;
}
REF_EXPORT void ref_sh_look_at(const float eye[3], const float at[3], const float up[3], float out[16]) {
    mat4 m = shadow_matrices_tu::lookAt(vec3(eye[0], eye[1], eye[2]), vec3(at[0], at[1], at[2]), vec3(up[0], up[1], up[2]));
    memcpy(out, &m, 64);
}
REF_EXPORT void ref_sh_orthographic(float l, float r, float b, float t, float n, float f, float out[16]) {
    mat4 m = shadow_matrices_tu::OrthographicProjection(l, r, b, t, n, f);
    memcpy(out, &m, 64);
}
REF_EXPORT void ref_sh_struct_sizes(uint32_t out[9]) {
    out[0] = sizeof(Instance); out[1] = sizeof(MeshInfo); out[2] = sizeof(Meshlet); out[3] = sizeof(MeshletReference);
    out[4] = sizeof(PrefixSumValue); out[5] = sizeof(DrawIndirectCommand); out[6] = sizeof(DispatchIndirectCommand);
    out[7] = sizeof(MiscStorage); out[8] = sizeof(Uniforms);
}

// ---- stage 2: vertex shaders, depth reduction, lighting resolve, display transform ----

// gl_Position of vertex `vertex_index` (0 .. 3 * triangle_count - 1) of the draw with firstInstance == slot, and the
// flat `packed` output: visbuffer_opaque_vertex (cascade < 0) / shadowmap_opaque_vertex (rasterization.glsl:38-50,101-114)
REF_EXPORT uint32_t ref_sh_vertex(const Uniforms* u, int cascade, uint32_t slot, uint32_t vertex_index, float position[4]) {
    uniform_buffer.address = (uint64_t)u;
    gl_VertexIndex = vertex_index;
    gl_InstanceIndex = slot;
    raster_tu::packed = 0;
    if (cascade < 0) {
        raster_tu::visbuffer_opaque_vertex();
    } else {
        shadow_constant.cascade_index = (uint32_t)cascade;
        raster_tu::shadowmap_opaque_vertex();
    }
    memcpy(position, &gl_Position, 16);
    return raster_tu::packed;
}
// ... and the alpha-clip variants, which also output the uv (rasterization.glsl:66-88,118-135)
REF_EXPORT uint32_t ref_sh_vertex_alpha_clip(const Uniforms* u, int cascade, uint32_t slot, uint32_t vertex_index, float position[4], float uv_out[2]) {
    uniform_buffer.address = (uint64_t)u;
    gl_VertexIndex = vertex_index;
    gl_InstanceIndex = slot;
    raster_tu::packed = 0;
    if (cascade < 0) {
        raster_tu::visbuffer_alpha_clip_vertex();
    } else {
        shadow_constant.cascade_index = (uint32_t)cascade;
        raster_tu::shadowmap_alpha_clip_vertex();
    }
    memcpy(position, &gl_Position, 16);
    uv_out[0] = raster_tu::uv.x;
    uv_out[1] = raster_tu::uv.y;
    return raster_tu::packed;
}

static void set_frame(const LhoConfig* cfg, const uint64_t* vis, const float* shadow, uint32_t width, uint32_t height) {
    shim_frame = ShimFrame{};
    shim_frame.cfg = cfg;
    shim_frame.vis = vis;
    shim_frame.shadow = shadow;
    shim_frame.width = width;
    shim_frame.height = height;
    shim_frame.shadow_size = cfg ? cfg->shadowmap_size : 0;
}

// read_depth over the whole target (rendering.cpp:290-303: dispatch_size(w, 8 * 4) x dispatch_size(h, 8 * 4) groups of 8 x 8)
REF_EXPORT void ref_sh_read_depth(const Uniforms* u, const LhoConfig* cfg, const uint64_t* vis) {
    uniform_buffer.address = (uint64_t)u;
    set_frame(cfg, vis, nullptr, u->window_size.x, u->window_size.y);
    dispatch(groups(u->window_size.x, 32), groups(u->window_size.y, 32), 8, 8, [] { depth_tu::read_depth(); });
}

// render_geometry over the rows [y0, y1) (rendering.cpp:463-490: groups of 8 x 8 over the window); out = RGBA32F
REF_EXPORT void ref_sh_render_geometry(const Uniforms* u, const LhoConfig* cfg, const uint64_t* vis, const float* shadow, float* out,
                                       uint32_t y0, uint32_t y1) {
    const uint32_t W = u->window_size.x;
    lho::parallel_for(y1 - y0, lho::thread_count(cfg), [&](uint64_t a, uint64_t b) {
        uniform_buffer.address = (uint64_t)u;
        set_frame(cfg, vis, shadow, W, u->window_size.y);
        shim_frame.scene_out = out;
        for (uint64_t y = y0 + a; y < y0 + b; y++)
            for (uint32_t x = 0; x < W; x++) {
                gl_GlobalInvocationID.set(x, (uint32_t)y, 0);
                resolve_tu::render_geometry_main();
            }
    });
}

// display_transform over the whole target: scene = RGBA32F in, lut = size^3 RGB32F, out = RGBA32F (the sRGB-encoded
// value the reference stores to its UNORM swapchain image)
REF_EXPORT void ref_sh_display_transform(const Uniforms* u, const LhoConfig* cfg, const float* scene, const float* shadow, const float* lut,
                                         uint32_t lut_size, float* out) {
    const uint32_t W = u->window_size.x, H = u->window_size.y;
    lho::parallel_for(H, lho::thread_count(cfg), [&](uint64_t a, uint64_t b) {
        uniform_buffer.address = (uint64_t)u;
        set_frame(cfg, nullptr, shadow, W, H);
        shim_frame.scene_in = scene;
        shim_frame.lut = lut;
        shim_frame.lut_size = lut_size;
        shim_frame.swapchain_out = out;
        for (uint64_t y = a; y < b; y++)
            for (uint32_t x = 0; x < W; x++) {
                gl_GlobalInvocationID.set(x, (uint32_t)y, 0);
                display_tu::display_transform();
            }
    });
}
