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

#include <shared_cpu_gpu.h>   // the reference's own header, C++ side (-I$(REF)/src); brings `using namespace glm`

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
SHIM_SWZ inline vec<N, T, Q> operator-(SHIM_SWZ_T const& s, T b) { return s() - b; }
SHIM_SWZ inline vec<N, T, Q> operator+(SHIM_SWZ_T const& s, T b) { return s() + b; }
SHIM_SWZ inline vec<N, T, Q> operator*(SHIM_SWZ_T const& s, T b) { return s() * b; }
SHIM_SWZ inline vec<N, T, Q> operator/(SHIM_SWZ_T const& s, T b) { return s() / b; }
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
inline uint32_t subgroupMin(uint32_t v) { return v; }
inline uint32_t subgroupMax(uint32_t v) { return v; }
inline bool subgroupElect() { return true; }

}  // namespace refsh
