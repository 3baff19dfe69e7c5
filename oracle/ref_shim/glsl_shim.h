// A C++ environment in which the reference's GLSL compiles as it stands: its own vendored glm supplies the vector
// and matrix types and their arithmetic (GLSL's built-in types are glm's model), its own src/shared_cpu_gpu.h (which
// the reference already dual-compiles as C++ and GLSL) supplies the structs, and this header supplies what a Vulkan
// device would: buffer references over plain 64-bit addresses (host pointers here), built-in variables, atomics,
// subgroup operations (subgroup size 1) and the HLSL-style aliases of src/shaders/common/hlsl4glsl.glsl.
//
// TEST INFRASTRUCTURE (oracle/_ref/libref_shaders.so); written for this repo.  Arithmetic: glm's scalar code paths
// (no SIMD), compiled -ffp-contract=off; mat*vec, dot, length, normalize, inverse are therefore glm's.
#pragma once
// operator-style swizzles (v.xyz as a member) need glm's "MS extensions" language flag, which glm 0.9.9.7 only sets for
// gcc together with SIMD code paths; anonymous structs are a gcc extension too, so the flag is claimed directly.
#define _MSC_EXTENSIONS 1
#define GLM_FORCE_SWIZZLE 1
#define GLM_FORCE_PURE 1   // scalar code paths only
#include <cmath>
#include <cstdint>
#include <cstring>

#include <glm/glm.hpp>
#include <glm/gtc/type_precision.hpp>
#undef _MSC_EXTENSIONS

// The reference's own src/shared_cpu_gpu.h, GLSL side (through the filter, which only drops its #include of
// hlsl4glsl.glsl): the structs are then plain aggregates, as GLSL sees them (the C++ side gives Instance a constructor
// and no default one, which GLSL code such as `VertexData vertex_data;` cannot live with).  Layouts are identical
// (tests/test_ref_shaders.py compares the sizes with the C++ side's).
using namespace glm;
#define GLSL 1
#include "gen/shared_cpu_gpu.h"
#undef GLSL

#include <functional>
#include "lh_oracle.h"
#include "lh_oracle_internal.hpp"   // the samplers that stand in for the texture unit

namespace glm {
// ---- swizzle proxies in generic functions: forward to the vector they stand for ----
#define SHIM_SWZ template <int N, typename T, qualifier Q, int E0, int E1, int E2, int E3>
#define SHIM_SWZ_T detail::_swizzle<N, T, Q, E0, E1, E2, E3>
#define SHIM_UNARY(fn) \
    SHIM_SWZ inline auto fn(SHIM_SWZ_T const& s) { return fn(s()); }
SHIM_UNARY(length)
SHIM_UNARY(normalize)
SHIM_UNARY(abs)
#undef SHIM_UNARY
#undef SHIM_SWZ
#undef SHIM_SWZ_T

// GLSL converts integer vectors implicitly where C++ templates cannot deduce
inline uvec3 operator+(uint32_t a, u8vec3 const& b) { return uvec3(a + b.x, a + b.y, a + b.z); }
inline uvec2 operator*(uvec2 const& a, int b) { return a * (uint32_t)b; }
inline vec2 operator*(uvec2 const& a, vec2 const& b) { return vec2(a) * b; }
inline vec2 operator/(float a, uvec2 const& b) { return a / vec2(b); }
inline vec3 operator*(vec3 const& a, int b) { return a * (float)b; }
inline vec2 operator*(vec2 const& a, int b) { return a * (float)b; }
inline vec2 operator/(vec2 const& a, uvec2 const& b) { return a / vec2(b); }
inline float pow(int a, float b) { return std::pow((float)a, b); }
// pow(x, p) as pinned by the oracle and the CUDA pass: applied per component (the display transform calls it on a vec3)
inline float shim_pinned_pow(float x, float p) { return lho_pinned_powf(x, p); }
template <typename V>
inline V shim_pinned_pow(const V& x, float p) {
    V r = x;
    for (int i = 0; i < (int)x.length(); i++) r[i] = lho_pinned_powf(x[i], p);
    return r;
}
// pow(x, 5.0) as pinned by the oracle and the CUDA pass (GLSL leaves pow's precision to the implementation)
inline float shim_pow5(float x) {
    float x2 = x * x;
    return x2 * x2 * x;
}
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
}  // namespace glm

namespace refsh {

// ---- src/shaders/common/hlsl4glsl.glsl, C++ side ----
using float2 = vec2;
using float3 = vec3;
using float4 = vec4;
using float2x2 = mat2;
using float3x3 = mat3;
using float4x4 = mat4;
using uint32_t2 = uvec2;
using uint32_t3 = uvec3;
using uint32_t4 = uvec4;
using uint2 = uvec2;
using uint3 = uvec3;
using uint4 = uvec4;
using uint8_t3 = u8vec3;
using uint8_t4 = u8vec4;
using uint16_t2 = u16vec2;
using uint16_t3 = u16vec3;
using uint16_t4 = u16vec4;
using int8_t4 = i8vec4;
using int8_t3 = i8vec3;
inline uint32_t asuint(float v) { return floatBitsToUint(v); }
inline uvec4 asuint(vec4 const& v) { return floatBitsToUint(v); }
inline float asfloat(uint32_t v) { return uintBitsToFloat(v); }
template <class A, class B, class C>
inline auto lerp(A a, B b, C c) { return mix(a, b, c); }
inline float3 rcp(float3 value) { return 1.0f / value; }
inline float rcp(float value) { return 1.0f / value; }
inline float select(bool b, float t, float f) { return mix(f, t, b); }
inline uint32_t select(bool b, uint32_t t, uint32_t f) { return mix(f, t, b); }
inline uint32_t select(bool b, int t, uint32_t f) { return b ? (uint32_t)t : f; }   // select(width == 0, 1, width - 1)
inline float3 select(bool b, float3 t, float3 f) { return mix(f, t, bvec3(b)); }
inline float3 saturate(float3 value) { return clamp(value, float3(0), float3(1)); }

// ---- built-in variables and push constants (set by the dispatch loops in ref_shaders.cpp) ----
struct InvocationId {
    uint32_t x = 0, y = 0, z = 0;
    uvec2 xy{0, 0};
    operator uvec3() const { return uvec3(x, y, z); }
    void set(uint32_t x_, uint32_t y_, uint32_t z_) { x = x_; y = y_; z = z_; xy = uvec2(x_, y_); }
};
inline thread_local InvocationId gl_GlobalInvocationID;
inline thread_local uint32_t gl_VertexIndex = 0, gl_InstanceIndex = 0;
inline thread_local vec4 gl_Position;
inline thread_local UniformBufferAddressConstant uniform_buffer;
inline thread_local ShadowPassConstant shadow_constant;

// ---- buffer references (src/shaders/common/buffer_references.glsl, prefix_sum.glsl): a handle over an address ----
#define SHIM_BUFFER_SINGLE(Name, Type, member)                          \
    struct Name {                                                       \
        Type& member;                                                   \
        explicit Name(uint64_t a) : member(*reinterpret_cast<Type*>(a)) {} \
    };
#define SHIM_BUFFER_ARRAY(Name, Type, member)                           \
    struct Name {                                                       \
        Type* member;                                                   \
        explicit Name(uint64_t a) : member(reinterpret_cast<Type*>(a)) {} \
    };
SHIM_BUFFER_SINGLE(MeshInfoBuffer, MeshInfo, mesh_info)
SHIM_BUFFER_ARRAY(MeshletReferenceBuffer, MeshletReference, meshlet_reference)
SHIM_BUFFER_ARRAY(MeshletBuffer, Meshlet, meshlets)
SHIM_BUFFER_ARRAY(MicroIndexBuffer, u8vec3, indices)
SHIM_BUFFER_ARRAY(MicroIndexBufferSingle, uint8_t, indices)
SHIM_BUFFER_ARRAY(Index32Buffer, uint32_t, indices)
SHIM_BUFFER_ARRAY(Index16Buffer, uint16_t, indices)
SHIM_BUFFER_ARRAY(QuantizedPositionBuffer, u16vec3, positions)
SHIM_BUFFER_ARRAY(QuanitizedUvs, u16vec2, uvs)
SHIM_BUFFER_ARRAY(QuanitizedNormals, i8vec3, normals)
SHIM_BUFFER_ARRAY(InstanceBuffer, Instance, instances)
SHIM_BUFFER_SINGLE(MiscStorageBuffer, MiscStorage, misc_storage)
SHIM_BUFFER_SINGLE(UniformsBuffer, Uniforms, uniforms)
SHIM_BUFFER_ARRAY(DispatchCommandsBuffer, DispatchIndirectCommand, commands)
struct DrawCallBuffer {
    uint32_t& num_opaque;
    uint32_t& num_alpha_clip;
    DrawIndirectCommand* draw_calls;
    explicit DrawCallBuffer(uint64_t a)
        : num_opaque(*reinterpret_cast<uint32_t*>(a)), num_alpha_clip(*reinterpret_cast<uint32_t*>(a + 4)),
          draw_calls(reinterpret_cast<DrawIndirectCommand*>(a + 8)) {}
};
struct PrefixSumBuffer {
    uint64_t& counter;
    PrefixSumValue* values;
    explicit PrefixSumBuffer(uint64_t a) : counter(*reinterpret_cast<uint64_t*>(a)), values(reinterpret_cast<PrefixSumValue*>(a + 8)) {}
};
static_assert(sizeof(u8vec3) == 3 && sizeof(u16vec3) == 6 && sizeof(i8vec3) == 3 && sizeof(u16vec2) == 4, "scalar block layout");

// common/bindings.glsl:27-29
inline Uniforms get_uniforms() { return UniformsBuffer(uniform_buffer.address).uniforms; }

// ---- atomics and subgroup operations (one invocation at a time, ascending order: a legal schedule) ----
template <class T, class V>
inline T atomicAdd(T& mem, V v) { T old = mem; mem = old + (T)v; return old; }
template <class T>
inline T atomicMin(T& mem, T v) { T old = mem; mem = v < old ? v : old; return old; }
template <class T>
inline T atomicMax(T& mem, T v) { T old = mem; mem = v > old ? v : old; return old; }
inline uint64_t clockARB() { return 0; }
template <class T>
inline T nonuniformEXT(T v) { return v; }
inline void shim_discard() {}
inline uint32_t subgroupMin(uint32_t v) { return v; }
inline uint32_t subgroupMax(uint32_t v) { return v; }
inline bool subgroupElect() { return true; }


// ---- descriptors (src/shaders/common/bindings.glsl:7-25): what the texture unit of a device would do is done by the
// ---- oracle's samplers (oracle/lh_oracle_internal.hpp); the reference's GLSL decides WHAT is sampled and how it is used
struct ShimFrame {
    const LhoConfig* cfg = nullptr;       // textures + skybox
    const uint64_t* vis = nullptr;        // packed keys: depth = ~hi32, lo = slot << 7 | triangle
    const float* shadow = nullptr;        // 4 layers
    uint32_t width = 0, height = 0, shadow_size = 0;
    float* scene_out = nullptr;           // rw_scene_referred_framebuffer (RGBA32F)
    const float* scene_in = nullptr;      // scene_referred_framebuffer
    const float* lut = nullptr;           // display_transform_lut: size^3 RGB32F
    uint32_t lut_size = 0;
    float* swapchain_out = nullptr;       // swapchain_image (RGBA32F, before the UNORM conversion)
};
inline thread_local ShimFrame shim_frame;

struct sampler { int id; };
inline const sampler clamp_sampler{0}, repeat_sampler{1}, shadowmap_comparison_sampler{2};
struct texture2D { int id; uint32_t index; };          // id 0: textures[index], 1: depth_buffer, 2: scene_referred_framebuffer
struct utexture2D { int id; };
struct texture2DArray { int id; };
struct texture3D { int id; };
struct textureCube { int id; };
struct image2D { int id; };                            // 0: rw_scene_referred_framebuffer, 1: swapchain_image
struct TextureTable { texture2D operator[](uint32_t i) const { return texture2D{0, i}; } };
inline const TextureTable textures;
inline const texture2D scene_referred_framebuffer{2, 0}, depth_buffer{1, 0};
inline const utexture2D visibility_buffer{0};
inline const texture2DArray shadowmap{0};
inline const texture3D display_transform_lut{0};
inline const textureCube skybox{0};
inline const image2D rw_scene_referred_framebuffer{0}, swapchain_image{1};
struct sampler2D { texture2D t; sampler s; sampler2D(texture2D t_, sampler s_) : t(t_), s(s_) {} };
struct sampler2DArray { sampler2DArray(texture2DArray, sampler) {} };
struct sampler2DArrayShadow { sampler2DArrayShadow(texture2DArray, sampler) {} };
struct sampler3D { sampler3D(texture3D, sampler) {} };
struct samplerCube { samplerCube(textureCube, sampler) {} };

inline float shim_depth_at(int x, int y) {   // D32F: the float whose bits are ~hi32 of the key (0 where nothing was drawn)
    const uint64_t key = shim_frame.vis[(size_t)y * shim_frame.width + x];
    return uintBitsToFloat(~(uint32_t)(key >> 32));
}
inline vec4 texelFetch(texture2D t, ivec2 c, int) {
    if (t.id == 1) return vec4(shim_depth_at(c.x, c.y), 0, 0, 0);
    const float* p = shim_frame.scene_in + ((size_t)c.y * shim_frame.width + c.x) * 4;
    return vec4(p[0], p[1], p[2], p[3]);
}
inline uvec4 texelFetch(utexture2D, ivec2 c, int) {   // R32_UINT: triangle << 24 | slot (util.glsl:18-20)
    const uint32_t lo = (uint32_t)shim_frame.vis[(size_t)c.y * shim_frame.width + c.x];
    return uvec4(((lo & 0x7Fu) << 24) | (lo >> 7), 0, 0, 0);
}
// textureGather(depth_buffer + clamp_sampler): texels (i0,j1), (i1,j1), (i1,j0), (i0,j0) around the sample position
inline vec4 textureGather(sampler2D, vec2 uv) {
    const int W = (int)shim_frame.width, H = (int)shim_frame.height;
    const float x = uv.x * (float)W - 0.5f, y = uv.y * (float)H - 0.5f;
    auto cx = [&](int v) { return v < 0 ? 0 : (v > W - 1 ? W - 1 : v); };
    auto cy = [&](int v) { return v < 0 ? 0 : (v > H - 1 ? H - 1 : v); };
    const int i0 = (int)std::floor(x), j0 = (int)std::floor(y);
    return vec4(shim_depth_at(cx(i0), cy(j0 + 1)), shim_depth_at(cx(i0 + 1), cy(j0 + 1)), shim_depth_at(cx(i0 + 1), cy(j0)), shim_depth_at(cx(i0), cy(j0)));
}
inline vec4 textureGrad(sampler2D s, vec2 uv, vec2 dx, vec2 dy) {
    const LhoConfig* cfg = shim_frame.cfg;
    if (!cfg || !cfg->textures || s.t.index >= cfg->num_textures) return vec4(1, 1, 1, 1);   // an unbound slot (the oracle's convention)
    float out[4];
    lho::sample_texture_grad(cfg->textures[s.t.index], &uv.x, &dx.x, &dy.x, out);
    return vec4(out[0], out[1], out[2], out[3]);
}
inline vec4 textureLod(samplerCube, vec3 dir, float) {
    float out[3];
    lho::sample_skybox(shim_frame.cfg, lho::V3{dir.x, dir.y, dir.z}, out);
    return vec4(out[0], out[1], out[2], 1.0f);
}
inline float texture(sampler2DArrayShadow, vec4 c) {   // (u, v, layer, reference depth); the layer index is clamped
    const uint32_t n = shim_frame.shadow_size;
    int layer = (int)std::floor(c.z + 0.5f);
    layer = layer < 0 ? 0 : (layer > 3 ? 3 : layer);
    return lho::sample_shadow(shim_frame.shadow + (size_t)layer * n * n, n, c.x, c.y, c.w);
}
inline vec4 textureLod(sampler2DArray, vec3 c, float) {   // display_transform's debug overlay: nearest layer, bilinear, clamp
    const uint32_t n = shim_frame.shadow_size;
    int layer = (int)std::floor(c.z + 0.5f);
    layer = layer < 0 ? 0 : (layer > 3 ? 3 : layer);
    const float* L = shim_frame.shadow + (size_t)layer * n * n;
    const float x = c.x * (float)n - 0.5f, y = c.y * (float)n - 0.5f, fx = std::floor(x), fy = std::floor(y), wx = x - fx, wy = y - fy;
    auto at = [&](int xi, int yi) { xi = xi < 0 ? 0 : (xi > (int)n - 1 ? (int)n - 1 : xi); yi = yi < 0 ? 0 : (yi > (int)n - 1 ? (int)n - 1 : yi); return L[(size_t)yi * n + xi]; };
    const float top = at((int)fx, (int)fy) + wx * (at((int)fx + 1, (int)fy) - at((int)fx, (int)fy));
    const float bot = at((int)fx, (int)fy + 1) + wx * (at((int)fx + 1, (int)fy + 1) - at((int)fx, (int)fy + 1));
    return vec4(top + wy * (bot - top));
}
// textureLod(display_transform_lut + clamp_sampler): trilinear, clamp to edge
inline vec4 textureLod(sampler3D, vec3 uvw, float) {
    float out[3];
    lho::sample_lut(shim_frame.lut, shim_frame.lut_size, &uvw.x, out);
    return vec4(out[0], out[1], out[2], 1.0f);
}
inline void imageStore(image2D img, ivec2 c, vec4 v) {
    float* dst = (img.id == 0 ? shim_frame.scene_out : shim_frame.swapchain_out) + ((size_t)c.y * shim_frame.width + c.x) * 4;
    dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
}

}  // namespace refsh
