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
