// Full-screen lighting resolve from the visibility buffer, and the display
// transform that follows it.
//
// Replaces (reference file:line)
//   render_geometry main            src/shaders/render_geometry.comp:70-266
//   CalcFullBary / interpolate      src/shaders/common/vbuffer.glsl:11-119
//   BRDF and helpers                src/shaders/common/pbr.glsl:3-96
//   display_transform               src/shaders/display_transform.comp:37-81
//   samplers                        src/main.cpp:550-581
//
// B200 design.  One thread per pixel; each warp owns an 8x4 pixel block so the
// three-vertex attribute refetch and the texture footprints of a warp stay
// inside one or two meshlets.  Depth and ids arrive in ONE 8-byte word per
// pixel (the reference reads a D32F image and an R32U image); material
// textures are CUDA block-compressed (BC7 sRGB) mipmapped arrays: the texture
// unit decodes the blocks, converts sRGB and filters each bilinear tap, while
// the anisotropic footprint (16x, from the analytic barycentric derivatives)
// is laid out in the kernel per the Vulkan model so that the lit colour is
// reproducible; the shadow compare-and-filter, which CUDA
// has no sampler for, is done by hand from the u32 depth layers.  Output is
// written as one 8-byte RGBA16F store per pixel.
//
// This pass is floating point, and REPRODUCIBLE: the file is compiled with -fmad=false (Makefile), divisions and
// square roots are IEEE, and every expression keeps the operation order of the reference's shader as restated by the
// checker (which models the texture unit's filter bit for bit), so the RGBA16F colour of a covered pixel is meant to
// be bit-identical to the checker's -- the pass has several discontinuous choices (shadow compares, anisotropic probe
// count, the BRDF where N.V is clamped) on which a last-bit difference would otherwise become a visible one.
#include <cuda_fp16.h>

#include "lh_internal.cuh"

namespace lh {

namespace {

struct F3 {
    float x, y, z;
};
__device__ __forceinline__ F3 operator+(F3 a, F3 b) { return F3{a.x + b.x, a.y + b.y, a.z + b.z}; }
__device__ __forceinline__ F3 operator-(F3 a, F3 b) { return F3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ F3 operator*(F3 a, float s) { return F3{a.x * s, a.y * s, a.z * s}; }
__device__ __forceinline__ float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ F3 cross(F3 a, F3 b) { return F3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
__device__ __forceinline__ F3 normalize(F3 a) {
    float inv = 1.0f / sqrtf(dot(a, a));
    return a * inv;
}
struct Bary {
    F3 lambda, ddx, ddy;
};

// CalcFullBary, vbuffer.glsl:11-63.  Everything up to the barycentrics themselves is PINNED arithmetic (the reference's
// operation order with IEEE operations, lh_internal.cuh): the barycentrics give the world position, the world position
// gives the shadow coordinate, and the shadow lookup COMPARES it with stored depths -- without a depth bias for opaque
// meshes -- so a last-bit difference would flip comparisons on every surface that grazes its own shadow-map texels.
// With this part bit-identical to the oracle (which is bit-identical to the reference's GLSL compiled on the host), the
// shadow term only differs by the rounding of the filter weights.
__device__ Bary calc_full_bary(float4 p0, float4 p1, float4 p2, float ndc_x, float ndc_y, float win_w, float win_h) {
    Bary r;
    F3 invw{pdiv(1.0f, p0.w), pdiv(1.0f, p1.w), pdiv(1.0f, p2.w)};
    float n0x = pmul(p0.x, invw.x), n0y = pmul(p0.y, invw.x);
    float n1x = pmul(p1.x, invw.y), n1y = pmul(p1.y, invw.y);
    float n2x = pmul(p2.x, invw.z), n2y = pmul(p2.y, invw.z);
    float inv_det = pdiv(1.0f, psub(pmul(psub(n2x, n1x), psub(n0y, n1y)), pmul(psub(n0x, n1x), psub(n2y, n1y))));
    r.ddx = F3{pmul(pmul(psub(n1y, n2y), inv_det), invw.x), pmul(pmul(psub(n2y, n0y), inv_det), invw.y), pmul(pmul(psub(n0y, n1y), inv_det), invw.z)};
    r.ddy = F3{pmul(pmul(psub(n2x, n1x), inv_det), invw.x), pmul(pmul(psub(n0x, n2x), inv_det), invw.y), pmul(pmul(psub(n1x, n0x), inv_det), invw.z)};
    float ddx_sum = padd(padd(r.ddx.x, r.ddx.y), r.ddx.z);
    float ddy_sum = padd(padd(r.ddy.x, r.ddy.y), r.ddy.z);
    float dvx = psub(ndc_x, n0x), dvy = psub(ndc_y, n0y);
    float interp_inv_w = padd(padd(invw.x, pmul(dvx, ddx_sum)), pmul(dvy, ddy_sum));
    float interp_w = pdiv(1.0f, interp_inv_w);
    r.lambda.x = pmul(interp_w, padd(padd(invw.x, pmul(dvx, r.ddx.x)), pmul(dvy, r.ddy.x)));
    r.lambda.y = pmul(interp_w, padd(padd(0.0f, pmul(dvx, r.ddx.y)), pmul(dvy, r.ddy.y)));
    r.lambda.z = pmul(interp_w, padd(padd(0.0f, pmul(dvx, r.ddx.z)), pmul(dvy, r.ddy.z)));
    // The derivatives decide the anisotropic footprint (probe count, major axis: discontinuous choices), so they are
    // pinned as well.
    float sx = pdiv(2.0f, win_w), sy = -pdiv(2.0f, win_h);  // the y derivative is flipped (vbuffer.glsl:50-51)
    r.ddx = F3{pmul(r.ddx.x, sx), pmul(r.ddx.y, sx), pmul(r.ddx.z, sx)};
    r.ddy = F3{pmul(r.ddy.x, sy), pmul(r.ddy.y, sy), pmul(r.ddy.z, sy)};
    ddx_sum = pmul(ddx_sum, sx);
    ddy_sum = pmul(ddy_sum, sy);
    const float iw_ddx = pdiv(1.0f, padd(interp_inv_w, ddx_sum)), iw_ddy = pdiv(1.0f, padd(interp_inv_w, ddy_sum));
    auto perspective = [&](float lambda, float d, float iw) { return psub(pmul(iw, padd(pmul(lambda, interp_inv_w), d)), lambda); };
    r.ddx = F3{perspective(r.lambda.x, r.ddx.x, iw_ddx), perspective(r.lambda.y, r.ddx.y, iw_ddx), perspective(r.lambda.z, r.ddx.z, iw_ddx)};
    r.ddy = F3{perspective(r.lambda.x, r.ddy.x, iw_ddy), perspective(r.lambda.y, r.ddy.y, iw_ddy), perspective(r.lambda.z, r.ddy.z, iw_ddy)};
    return r;
}

struct Interp {
    float value, dx, dy;
};
__device__ __forceinline__ Interp interp1(const Bary& b, float v0, float v1, float v2) {
    F3 m{v0, v1, v2};
    return Interp{dot(m, b.lambda), dot(m, b.ddx), dot(m, b.ddy)};
}
// ... with the value in pinned arithmetic (the world position)
__device__ __forceinline__ float dot_pinned(float v0, float v1, float v2, F3 w) { return padd(padd(pmul(v0, w.x), pmul(v1, w.y)), pmul(v2, w.z)); }
__device__ __forceinline__ Interp interp1_pinned(const Bary& b, float v0, float v1, float v2) {
    F3 m{v0, v1, v2};
    return Interp{dot_pinned(v0, v1, v2, b.lambda), dot(m, b.ddx), dot(m, b.ddy)};
}
// ... with value and derivatives pinned (the texture coordinate: its derivatives lay out the footprint)
__device__ __forceinline__ Interp interp1_all_pinned(const Bary& b, float v0, float v1, float v2) {
    return Interp{dot_pinned(v0, v1, v2, b.lambda), dot_pinned(v0, v1, v2, b.ddx), dot_pinned(v0, v1, v2, b.ddy)};
}

// pbr.glsl:3-96, in the checker's operation order (no contraction in this file)
constexpr float kPi = 3.14159265358979323846f;
__device__ __forceinline__ float pow5(float x) {   // pow(x, 5.0): implementation-defined in GLSL, pinned as ((x*x)*(x*x))*x
    float x2 = x * x;
    return x2 * x2 * x;
}
__device__ __forceinline__ float f_schlick1(float u, float f0, float f90) { return f0 + (f90 - f0) * pow5(1.0f - u); }
__device__ __forceinline__ float fd_burley(float nov, float nol, float loh, float roughness) {
    float f90 = 0.5f + 2.0f * roughness * loh * loh;
    return f_schlick1(nol, 1.0f, f90) * f_schlick1(nov, 1.0f, f90) * (1.0f / kPi);
}
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

__device__ F3 brdf(F3 v, F3 l, F3 n, float perceptual_roughness, float metallic, F3 base, bool thin_walled) {
    F3 h = normalize(F3{v.x + l.x, v.y + l.y, v.z + l.z});
    const float refl = ((1.5f - 1.0f) / (1.5f + 1.0f)) * ((1.5f - 1.0f) / (1.5f + 1.0f));   // reflectance_for_ior(1.5)
    const float* bp = &base.x;
    float diffuse[3], f0[3];
#pragma unroll
    for (int c = 0; c < 3; c++) {
        diffuse[c] = (1.0f - metallic) * bp[c];
        f0[c] = refl * (1.0f - metallic) + bp[c] * metallic;
    }
    float nov = clampf(dot(n, v), 1e-5f, 1.0f), nol = clampf(dot(n, l), 0.0f, 1.0f);
    float noh = clampf(dot(n, h), 0.0f, 1.0f), loh = clampf(dot(l, h), 0.0f, 1.0f);
    float roughness = perceptual_roughness * perceptual_roughness;
    float a2 = roughness * roughness;
    float f = (noh * a2 - noh) * noh + 1.0f;
    float D = a2 / (kPi * f * f);
    float ggxl = nov * sqrtf((-nol * a2 + nol) * nol + a2);
    float ggxv = nol * sqrtf((-nov * a2 + nov) * nov + a2);
    float V = 0.5f / (ggxv + ggxl);
    float p5 = pow5(1.0f - loh);
    float burley = fd_burley(nov, nol, loh, roughness);
    F3 nn{-n.x, -n.y, -n.z};
    float rev_nov = clampf(dot(nn, v), 1e-5f, 1.0f), rev_nol = clampf(dot(nn, l), 0.0f, 1.0f);
    float rev_burley = thin_walled ? fd_burley(rev_nov, rev_nol, loh, roughness) : 0.0f;
    F3 out;
    float* o = &out.x;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        float F = f0[c] + (1.0f - f0[c]) * p5;
        float fr = (D * V) * F;
        float fd = diffuse[c] * burley;
        float irr = fmaxf((fd + fr) * nol, 0.0f);
        if (thin_walled) irr += diffuse[c] * rev_burley * rev_nol;
        o[c] = isnan(irr) ? 0.0f : irr;
    }
    return out;
}

// log2 of a positive normal float by a fixed sequence of IEEE single operations (the checker runs the same sequence):
// the level of detail decides the blend of two mip levels, and libm / the SFU would each round it their own way
__device__ __forceinline__ float pinned_log2f(float x) {
    uint32_t bits = __float_as_uint(x);
    int e = (int)(bits >> 23) - 127;
    float m = __uint_as_float((bits & 0x007FFFFFu) | 0x3F800000u);   // [1, 2)
    if (m > 1.41421354f) {
        m = m * 0.5f;
        e += 1;
    }
    const float s = (m - 1.0f) / (m + 1.0f), s2 = s * s;
    float q = 0.142857149f;            // 2 * atanh(s) = 2 s (1 + s^2/3 + s^4/5 + s^6/7), |s| < 0.1716
    q = q * s2 + 0.2f;
    q = q * s2 + 0.333333343f;
    q = q * s2 + 1.0f;
    return (float)e + (2.0f * s * q) * 1.44269502f;
}

// sampler2DArrayShadow + compare LESS + linear (src/main.cpp:569-581), by hand
__device__ __forceinline__ float sample_shadow(const uint32_t* layer, int n, float u, float v, float ref) {
    float x = u * (float)n - 0.5f, y = v * (float)n - 0.5f;
    if (!(isfinite(x) && isfinite(y))) return 0.0f;
    float fx = floorf(x), fy = floorf(y), wx = x - fx, wy = y - fy;
    fx = fminf(fmaxf(fx, -2.0f), (float)n + 1.0f);
    fy = fminf(fmaxf(fy, -2.0f), (float)n + 1.0f);
    int x0 = min(max((int)fx, 0), n - 1), x1 = min(max((int)fx + 1, 0), n - 1);
    int y0 = min(max((int)fy, 0), n - 1), y1 = min(max((int)fy + 1, 0), n - 1);
    float c00 = ref < __uint_as_float(__ldg(layer + (size_t)y0 * n + x0)) ? 1.0f : 0.0f;
    float c10 = ref < __uint_as_float(__ldg(layer + (size_t)y0 * n + x1)) ? 1.0f : 0.0f;
    float c01 = ref < __uint_as_float(__ldg(layer + (size_t)y1 * n + x0)) ? 1.0f : 0.0f;
    float c11 = ref < __uint_as_float(__ldg(layer + (size_t)y1 * n + x1)) ? 1.0f : 0.0f;
    float top = c00 + wx * (c10 - c00), bot = c01 + wx * (c11 - c01);
    return top + wy * (bot - top);
}

// The 3x3 PCF of render_geometry.comp:200-226 when the tap spacing (1/1024) is one texel: the nine bilinear compare taps
// touch a 4x4 texel neighbourhood, which is loaded and compared ONCE (16 loads instead of 36, one bit per texel).  Every
// tap keeps its own texel position and weights, computed exactly as sample_shadow computes them (u + k/1024 rounds, so
// the three fractions along an axis differ in their last bits), and the sum runs in sample_shadow's order, so the
// result is bit-identical to nine sample_shadow calls.  Common case: tap k along an axis starts at texel floor + k, so
// row r of tap column k lerps the mask bits (r, k + 1) and (r, k + 2) -- twelve x-lerps shared by the nine taps.  A tap
// whose floor lands elsewhere (possible only through that rounding) sends the pixel down the literal path.
__device__ __forceinline__ float pcf3x3_texel_spaced(const uint32_t* layer, int n, float u, float v, float ref) {
    const float x = u * (float)n - 0.5f, y = v * (float)n - 0.5f;
    if (!(isfinite(x) && isfinite(y))) return 0.0f;   // then every tap is 0
    const float fx = fminf(fmaxf(floorf(x), -3.0f), (float)n + 2.0f), fy = fminf(fmaxf(floorf(y), -3.0f), (float)n + 2.0f);
    const int bx = (int)fx - 1, by = (int)fy - 1;
    float wx[3], wy[3];
    bool regular = true;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float tx = (u + (float)(k - 1) / 1024.0f) * (float)n - 0.5f, ty = (v + (float)(k - 1) / 1024.0f) * (float)n - 0.5f;
        const float tfx = floorf(tx), tfy = floorf(ty);
        wx[k] = tx - tfx;
        wy[k] = ty - tfy;
        // tfx == fx + (k - 1) also rules out a non-finite tap and the clamps of sample_shadow (fx is inside them by 1)
        regular = regular && tfx == fx + (float)(k - 1) && tfy == fy + (float)(k - 1);
    }
    regular = regular && fx >= -1.0f && fx <= (float)n && fy >= -1.0f && fy <= (float)n;   // so that fx +- 1 stays inside sample_shadow's clamp [-2, n + 1]
    if (!regular) {
        float sum = 0.0f;
        for (int ox = -1; ox <= 1; ox++)
            for (int oy = -1; oy <= 1; oy++) sum += sample_shadow(layer, n, u + (float)ox / 1024.0f, v + (float)oy / 1024.0f, ref);
        return sum;
    }
    uint32_t mask = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        const uint32_t* row = layer + (size_t)min(max(by + r, 0), n - 1) * n;
#pragma unroll
        for (int c = 0; c < 4; c++)
            if (ref < __uint_as_float(__ldg(row + min(max(bx + c, 0), n - 1)))) mask |= 1u << (4 * r + c);
    }
    // c + w * (d - c) with c, d in {0, 1} is one of 0, 1, w, 1 - w -- each exactly what the expression rounds to
    float lerp_x[3][4];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float w = wx[k], omw = 1.0f - w;
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const bool c = (mask >> (4 * r + k)) & 1u, d = (mask >> (4 * r + k + 1)) & 1u;
            lerp_x[k][r] = c ? (d ? 1.0f : omw) : (d ? w : 0.0f);
        }
    }
    float sum = 0.0f;
#pragma unroll
    for (int kx = 0; kx < 3; kx++)
#pragma unroll
        for (int ky = 0; ky < 3; ky++) {
            const float top = lerp_x[kx][ky], bot = lerp_x[kx][ky + 1];
            sum += top + wy[ky] * (bot - top);
        }
    return sum;
}

// textureGrad through repeat_sampler (render_geometry.comp:59-66, src/main.cpp:550-560).  Footprint per the Vulkan
// specification's anisotropic model; every tap is one hardware bilinear fetch (BC decode + sRGB in the texture unit)
// at an integer mip level.
__device__ __forceinline__ float4 sample_trilinear(cudaTextureObject_t t, float u, float v, float lod, float max_level) {
    lod = fminf(fmaxf(lod, 0.0f), max_level);
    const float l0 = floorf(lod), f = lod - l0;
    float4 a = tex2DLod<float4>(t, u, v, l0);
    if (f > 0.0f && l0 + 1.0f <= max_level) {
        const float4 b = tex2DLod<float4>(t, u, v, l0 + 1.0f);
        // pinned (no contraction): the oracle models the texture unit bit for bit, so the blend of two fetches should not differ either
        a.x = padd(a.x, pmul(f, psub(b.x, a.x))); a.y = padd(a.y, pmul(f, psub(b.y, a.y)));
        a.z = padd(a.z, pmul(f, psub(b.z, a.z))); a.w = padd(a.w, pmul(f, psub(b.w, a.w)));
    }
    return a;
}
// The footprint depends on the texture's size and the uv derivatives only, so the material's three textures (same size
// in practice) share one: probes == 0 stands for the degenerate case (a single trilinear fetch at lod 0).
struct Footprint {
    float lod, max_level;
    float2 major;
    int probes;
};
__device__ __forceinline__ Footprint texture_footprint(uint4 dims, float2 ddx, float2 ddy) {
    Footprint f;
    const float w = (float)dims.x, h = (float)dims.y;
    f.max_level = (float)(dims.z - 1u);
    const float ax0 = pmul(ddx.x, w), ax1 = pmul(ddx.y, h), ay0 = pmul(ddy.x, w), ay1 = pmul(ddy.y, h);
    const float px = __fsqrt_rn(padd(pmul(ax0, ax0), pmul(ax1, ax1))), py = __fsqrt_rn(padd(pmul(ay0, ay0), pmul(ay1, ay1)));
    const float pmax = fmaxf(px, py), pmin = fminf(px, py);
    f.probes = 0;
    f.lod = 0.0f;
    f.major = make_float2(0.0f, 0.0f);
    if (!(pmax > 0.0f) || !isfinite(pmax)) return f;
    float n = pmin > 0.0f ? fminf(ceilf(pdiv(pmax, pmin)), 16.0f) : 16.0f;
    if (pmax <= 1.0f) n = 1.0f;  // magnification: a single bilinear probe
    f.lod = pinned_log2f(pdiv(pmax, n));
    f.probes = (int)n;
    f.major = px >= py ? ddx : ddy;
    return f;
}
__device__ float4 sample_texture_footprint(cudaTextureObject_t t, const Footprint& f, float u, float v) {
    if (f.probes == 0) return sample_trilinear(t, u, v, 0.0f, f.max_level);
    float4 acc = make_float4(0, 0, 0, 0);
    for (int i = 0; i < f.probes; i++) {
        const float o = psub(pdiv((float)i + 0.5f, (float)f.probes), 0.5f);
        const float4 s = sample_trilinear(t, padd(u, pmul(o, f.major.x)), padd(v, pmul(o, f.major.y)), f.lod, f.max_level);
        acc.x = padd(acc.x, s.x); acc.y = padd(acc.y, s.y); acc.z = padd(acc.z, s.z); acc.w = padd(acc.w, s.w);
    }
    const float inv = pdiv(1.0f, (float)f.probes);
    return make_float4(pmul(acc.x, inv), pmul(acc.y, inv), pmul(acc.z, inv), pmul(acc.w, inv));
}

// scene_remote: rank 0's target in peer memory (sort-last ranks != 0 deliver their rows there; a warp row is 64 B)
__device__ __forceinline__ void store_rgba16f(void* scene, void* scene_remote, size_t pixel, float r, float g, float b, float a) {
    __half2 lo = __floats2half2_rn(r, g), hi = __floats2half2_rn(b, a);
    uint2 packed;
    packed.x = *reinterpret_cast<uint32_t*>(&lo);
    packed.y = *reinterpret_cast<uint32_t*>(&hi);
    reinterpret_cast<uint2*>(scene)[pixel] = packed;
    if (scene_remote) reinterpret_cast<uint2*>(scene_remote)[pixel] = packed;
}

// a MeshletReference, possibly in a peer's memory (which the read-only cache path must not serve)
__device__ __forceinline__ unsigned long long owner_load(const LhMeshletReference* p) {
    return __ldcg((const unsigned long long*)p);
}

// PHASE 0: the whole pass.  The pass can also run in two launches, so that everything that does not need the shadow
// maps overlaps the shadow passes (lh_render_frame):
// PHASE 1 (material): visibility decode, attribute refetch, barycentrics, textures, BRDF -> per pixel the un-shadowed
//          sun term, the ambient term, the world position and the alpha-clip flag (40 B in P.resolve_tmp); sky pixels
//          are finished here;
// PHASE 2 (shadow): cascade selection, 3x3 PCF, combination, RGBA16F store.
// The combination is written with explicit roundings so that all phases produce the same bits.
#ifndef LH_RESOLVE_MIN_CTAS
#define LH_RESOLVE_MIN_CTAS 4   // 64 registers, 48 bytes of spills, 32 warps per SM: the pass waits on dependent loads (measured: 0.90 -> 0.74 ms against 3 CTAs at 79 registers)
#endif
template <int PHASE>
__global__ void __launch_bounds__(256, LH_RESOLVE_MIN_CTAS) resolve_kernel(const __grid_constant__ FrameParams P) {
    const uint32_t W = P.u.window_size[0], H = P.u.window_size[1];
    // the four cascade matrices are produced on the device (MiscStorage): staged once per CTA instead of 16 loads per
    // thread and cascade
    __shared__ float cascade_matrix[4][16];
    if (PHASE != 1) {
        const LhMiscStorage* misc = (const LhMiscStorage*)P.u.misc_storage;
        if (threadIdx.x < 64) cascade_matrix[threadIdx.x >> 4][threadIdx.x & 15] = (&misc->uv_space_shadow_matrices[threadIdx.x >> 4].m[0][0])[threadIdx.x & 15];
        __syncthreads();
    }
    // 32x8 pixels per CTA, one 8x4 block per warp
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t x = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    const uint32_t y = P.resolve_y0 + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    if (x >= W || y >= min(H, P.resolve_y1)) return;
    const size_t pixel = (size_t)y * W + x;
    const float ndc_x = psub(pmul(pdiv((float)x, (float)W), 2.0f), 1.0f);  // pixel corner, render_geometry.comp:82-84
    const float ndc_y = psub(pmul(pdiv((float)y, (float)H), 2.0f), 1.0f);
    float4* tmp0 = (float4*)P.resolve_tmp;
    float4* tmp1 = tmp0 + (size_t)W * H;
    float2* tmp2 = (float2*)(tmp1 + (size_t)W * H);
    F3 world, sun, ambient;
    F3 dbg_base{0, 0, 0}, dbg_normal{0, 0, 0}, dbg_view{0, 0, 0}, dbg_fp{0, 0, 0};
    float2 dbg_rm = make_float2(0, 0), dbg_uv = make_float2(0, 0);
    bool alpha;
    if (PHASE != 2) {
    const unsigned long long key = P.vis[pixel];
    if (lh_vis_depth_bits(key) == 0u) {  // depth == 0: sky (:86-101)
        const V4 un = mat4_mul(&P.u.perspective_inverse.m[0][0], ndc_x, ndc_y, 0.0f, 1.0f);
        F3 local = normalize(F3{un.x, un.y, un.z});
        const V4 ray = mat4_mul(&P.u.view_inverse.m[0][0], local.x, local.y, local.z, 0.0f);
        float4 s = P.sky ? texCubemapLod<float4>(P.sky, ray.x, ray.y, ray.z, 0.0f) : make_float4(0, 0, 0, 0);
        store_rgba16f(P.scene, P.scene_remote, pixel, s.x * 8.0f, s.y * 8.0f, s.z * 8.0f, 1.0f);
        if (PHASE == 1) tmp2[pixel] = make_float2(0.0f, -1.0f);   // nothing left to do for this pixel
        return;
    }
    const uint32_t slot = lh_vis_slot(key), tri = lh_vis_tri(key);
    const LhInstance* inst;
    const LhMeshInfo* mi;
    const uint8_t* micro;
    const uint8_t* first_index;   // byte address of the meshlet's first vertex index
    uint32_t flags;
    {
        // sort-last: the slot is global (rank * slots_per_rank + local) and the reference lives in the owning rank's table
        const LhMeshletReference* ref_at = P.resolve_refs + slot;
        if (P.peer_refs) {
            const uint32_t owner = slot / P.slots_per_rank;
            ref_at = P.peer_refs[owner] + (slot - owner * P.slots_per_rank);
        }
        const unsigned long long ref = owner_load(ref_at);
        inst = (const LhInstance*)P.u.instances + (uint32_t)ref;
        const uint32_t meshlet_index = (uint32_t)(ref >> 32) & 0xFFFFu;
        mi = (const LhMeshInfo*)__ldg((const unsigned long long*)&inst->mesh_info_address);
        const LhMeshlet* m = (const LhMeshlet*)__ldg((const unsigned long long*)&mi->meshlets) + meshlet_index;
        flags = __ldg(&mi->flags);
        micro = (const uint8_t*)__ldg((const unsigned long long*)&mi->micro_indices) + __ldg(&m->triangle_offset);
        first_index = (const uint8_t*)__ldg((const unsigned long long*)&mi->indices) +
                      (size_t)__ldg(&m->index_offset) * ((flags & LH_MESH_INFO_FLAGS_32_BIT_INDICES) ? 4u : 2u);
    }
    micro += 3 * tri;
    float T[16];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float4 c = __ldg((const float4*)&inst->transform + k);
        T[4 * k] = c.x; T[4 * k + 1] = c.y; T[4 * k + 2] = c.z; T[4 * k + 3] = c.w;
    }
    const uint16_t* positions = (const uint16_t*)__ldg((const unsigned long long*)&mi->positions);
    const int8_t* normals = (const int8_t*)__ldg((const unsigned long long*)&mi->normals);
    const uint16_t* uvs = (const uint16_t*)__ldg((const unsigned long long*)&mi->uvs);
    const bool wide = (flags & LH_MESH_INFO_FLAGS_32_BIT_INDICES) != 0;
    alpha = (flags & LH_MESH_INFO_FLAGS_ALPHA_CLIP) != 0;

    F3 pos[3], nrm[3];
    float4 clip[3];
    float tu[3], tv[3];
    const float ts0 = __ldg(&mi->texture_scale[0]), ts1 = __ldg(&mi->texture_scale[1]);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const uint32_t vi = __ldg(micro + k);
        const uint32_t idx = wide ? __ldg((const uint32_t*)first_index + vi) : (uint32_t)__ldg((const uint16_t*)first_index + vi);
        const uint16_t* p = positions + 3 * (size_t)idx;
        const V4 w = mat4_mul(T, (float)__ldg(p), (float)__ldg(p + 1), (float)__ldg(p + 2), 1.0f);   // pinned, like the raster pass
        pos[k] = F3{w.x, w.y, w.z};
        const V4 c = mat4_mul(&P.u.combined_perspective_view.m[0][0], w.x, w.y, w.z, 1.0f);
        clip[k] = make_float4(c.x, c.y, c.z, c.w);
        const int8_t* n = normals + 3 * (size_t)idx;
        nrm[k] = normalize(F3{(float)__ldg(n), (float)__ldg(n + 1), (float)__ldg(n + 2)});
        tu[k] = (float)__ldg(uvs + 2 * (size_t)idx) * ts0;
        tv[k] = (float)__ldg(uvs + 2 * (size_t)idx + 1) * ts1;
    }
    const Bary bary = calc_full_bary(clip[0], clip[1], clip[2], ndc_x, ndc_y, (float)W, (float)H);
    const Interp wx = interp1_pinned(bary, pos[0].x, pos[1].x, pos[2].x);
    const Interp wy = interp1_pinned(bary, pos[0].y, pos[1].y, pos[2].y);
    const Interp wz = interp1_pinned(bary, pos[0].z, pos[1].z, pos[2].z);
    world = F3{wx.value, wy.value, wz.value};
    const Interp iu = interp1_all_pinned(bary, tu[0], tu[1], tu[2]), iv = interp1_all_pinned(bary, tv[0], tv[1], tv[2]);
    const float u = iu.value + __ldg(&mi->texture_offset[0]), v = iv.value + __ldg(&mi->texture_offset[1]);
    const float2 duv_dx = make_float2(iu.dx, iv.dx), duv_dy = make_float2(iu.dy, iv.dy);

    const F3 view_vec = normalize(F3{P.u.camera_pos[0], P.u.camera_pos[1], P.u.camera_pos[2]} - world);
    F3 normal{interp1(bary, nrm[0].x, nrm[1].x, nrm[2].x).value, interp1(bary, nrm[0].y, nrm[1].y, nrm[2].y).value,
              interp1(bary, nrm[0].z, nrm[1].z, nrm[2].z).value};
    {
        const float* N = &inst->normal_transform.m[0][0];
        float n9[9];
#pragma unroll
        for (int k = 0; k < 9; k++) n9[k] = __ldg(N + k);
        normal = normalize(F3{n9[0] * normal.x + n9[3] * normal.y + n9[6] * normal.z, n9[1] * normal.x + n9[4] * normal.y + n9[7] * normal.z,
                              n9[2] * normal.x + n9[5] * normal.y + n9[8] * normal.z});
    }
    if (dot(cross(pos[1] - pos[0], pos[2] - pos[0]), view_vec) < 0.0f) normal = normal * -1.0f;

    // material (:228-246)
    F3 base{1, 1, 1};
    float roughness = 1.0f, metallic = 1.0f;
    const uint32_t t_base = __ldg(&mi->base_color_texture_index), t_mr = __ldg(&mi->metallic_roughness_texture_index);
    const uint32_t t_nrm = __ldg(&mi->normal_texture_index);
    uint4 fp_dims = make_uint4(0u, 0u, 0u, 0u);   // the size the current footprint was laid out for
    Footprint fp{};
    auto footprint_for = [&](uint32_t slot) {
        const uint4 dims = __ldg(&P.tex_dims[slot]);
        if (dims.x != fp_dims.x || dims.y != fp_dims.y || dims.z != fp_dims.z) {
            fp = texture_footprint(dims, duv_dx, duv_dy);
            fp_dims = dims;
        }
    };
    if (t_base < LH_MAX_TEXTURES && P.textures[t_base]) {
        footprint_for(t_base);
        float4 s = sample_texture_footprint(P.textures[t_base], fp, u, v);
        base = F3{s.x, s.y, s.z};
    }
    base = F3{base.x * __ldg(&mi->base_color_factor[0]), base.y * __ldg(&mi->base_color_factor[1]), base.z * __ldg(&mi->base_color_factor[2])};
    if (t_mr < LH_MAX_TEXTURES && P.textures[t_mr]) {
        footprint_for(t_mr);
        float4 s = sample_texture_footprint(P.textures[t_mr], fp, u, v);
        roughness = s.y;
        metallic = s.z;
    }
    if (t_nrm != LH_UNUSED_TEXTURE_INDEX && t_nrm < LH_MAX_TEXTURES && P.textures[t_nrm]) {
        footprint_for(t_nrm);
        float4 s = sample_texture_footprint(P.textures[t_nrm], fp, u, v);
        F3 mn{s.x * 255.0f / 127.0f - 128.0f / 127.0f, s.y * 255.0f / 127.0f - 128.0f / 127.0f, s.z * 255.0f / 127.0f - 128.0f / 127.0f};
        // compute_cotangent_frame (:25-38)
        F3 pdx{wx.dx, wy.dx, wz.dx}, pdy{wx.dy, wy.dy, wz.dy};
        F3 dpy_perp = cross(pdy, normal), dpx_perp = cross(normal, pdx);
        F3 t3 = dpy_perp * duv_dx.x + dpx_perp * duv_dy.x;
        F3 b3 = dpy_perp * duv_dx.y + dpx_perp * duv_dy.y;
        float invmax = 1.0f / sqrtf(fmaxf(dot(t3, t3), dot(b3, b3)));
        normal = normalize(F3{(t3.x * invmax) * mn.x + (b3.x * invmax) * mn.y + normal.x * mn.z,
                              (t3.y * invmax) * mn.x + (b3.y * invmax) * mn.y + normal.y * mn.z,
                              (t3.z * invmax) * mn.x + (b3.z * invmax) * mn.y + normal.z * mn.z});
    }
    if (PHASE == 0) {
        dbg_base = base; dbg_normal = normal; dbg_rm = make_float2(roughness, metallic); dbg_uv = make_float2(u, v); dbg_view = view_vec;
        dbg_fp = F3{(float)fp.probes, fp.lod, 0.0f};
    }
    const float amb = 0.025f + 0.025f * normal.y;  // SH with sh[0] = sh[2] = 0.025 (:40-57)
    sun = brdf(view_vec, F3{P.u.sun_dir[0], P.u.sun_dir[1], P.u.sun_dir[2]}, normal, roughness, metallic, base, alpha);
    ambient = F3{amb * ((1.0f - metallic) * base.x), amb * ((1.0f - metallic) * base.y), amb * ((1.0f - metallic) * base.z)};
    if (PHASE == 1) {
        tmp0[pixel] = make_float4(sun.x, sun.y, sun.z, world.x);
        tmp1[pixel] = make_float4(ambient.x, ambient.y, ambient.z, world.y);
        tmp2[pixel] = make_float2(world.z, alpha ? 1.0f : 0.0f);
        return;
    }
    } else {   // PHASE 2: pick up what the material phase left
        const float2 c = tmp2[pixel];
        if (c.y < 0.0f) return;
        const float4 a = tmp0[pixel], b = tmp1[pixel];
        sun = F3{a.x, a.y, a.z};
        ambient = F3{b.x, b.y, b.z};
        world = F3{a.w, b.w, c.x};
        alpha = c.y > 0.5f;
    }

    // cascade selection (:183-196) and 3x3 PCF (:200-226)
    uint32_t cascade;
    float4 sc = make_float4(0, 0, 0, 0);
    for (cascade = 0; cascade < 4; cascade++) {
        const V4 c = mat4_mul(cascade_matrix[cascade], world.x, world.y, world.z, 1.0f);   // pinned
        sc = make_float4(c.x, c.y, c.z, c.w);
        if (fabsf(psub(sc.x, 0.5f)) < 0.5f && fabsf(psub(sc.y, 0.5f)) < 0.5f) break;
    }
    const float bias = alpha ? 0.000002f : 0.0f;
    if (sc.w != 1.0f) {   // x / 1.0f == x exactly: an orthographic cascade skips the three IEEE divisions
        sc.x = pdiv(sc.x, sc.w); sc.y = pdiv(sc.y, sc.w); sc.z = pdiv(sc.z, sc.w);
    }
    const int sn = (int)P.shadowmap_size;
    const uint32_t* layer = P.shadow + (size_t)min(cascade, 3u) * sn * sn;
    float shadow_sum = 0.0f;
    if (sn == 1024) {
        shadow_sum = pcf3x3_texel_spaced(layer, sn, sc.x, sc.y, psub(sc.z, bias));
    } else {
#pragma unroll
        for (int ox = -1; ox <= 1; ox++)
#pragma unroll
            for (int oy = -1; oy <= 1; oy++)
                shadow_sum += sample_shadow(layer, sn, padd(sc.x, (float)ox / 1024.0f), padd(sc.y, (float)oy / 1024.0f), psub(sc.z, bias));
    }
    shadow_sum /= 9.0f;
    if (sc.z > 1.0f) shadow_sum = 1.0f;

    if (PHASE == 0 && P.u.debug >= 100) {
        // attribution aid (not part of the reference's debug views, which are not reproduced): one input of the pass
        // instead of the colour (tools/resolve_diag.py compares it with the checker's).  Only the single-launch pass (lh_resolve) supports it.
        F3 t = world;
        if (P.u.debug == 100) t = F3{shadow_sum, shadow_sum, shadow_sum};
        else if (P.u.debug == 101) t = dbg_base;
        else if (P.u.debug == 102) t = dbg_normal;
        else if (P.u.debug == 103) t = F3{dbg_rm.x, dbg_rm.y, (float)cascade * 0.25f};
        else if (P.u.debug == 105) t = F3{dbg_uv.x, dbg_uv.y, 0.0f};
        else if (P.u.debug == 106) t = dbg_view;
        else if (P.u.debug == 107) t = sun;
        else if (P.u.debug == 108) t = dbg_fp;
        store_rgba16f(P.scene, P.scene_remote, pixel, t.x, t.y, t.z, 1.0f);
        return;
    }
    const float sun_intensity = kPi * shadow_sum;
    store_rgba16f(P.scene, P.scene_remote, pixel, sun.x * sun_intensity + ambient.x, sun.y * sun_intensity + ambient.y,
                  sun.z * sun_intensity + ambient.z, 1.0f);
}

// display_transform.comp:37-81: Tony McMapface LUT (48^3) + sRGB OETF -> BGRA8
// pow is implementation-defined in GLSL: pinned as the fixed IEEE sequence the checker uses (lh_internal.cuh)
__device__ __forceinline__ float srgb_oetf(float v) { return v <= 0.003130f ? 12.92f * v : 1.055f * pinned_powf(v, 1.0f / 2.4f) - 0.055f; }

__global__ void __launch_bounds__(256) display_transform_kernel(const __grid_constant__ FrameParams P) {
    const uint32_t W = P.u.window_size[0], H = P.u.window_size[1];
    const uint64_t pixels = (uint64_t)W * H;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (uint64_t)gridDim.x * blockDim.x) {
        uint2 packed = reinterpret_cast<const uint2*>(P.scene)[i];
        __half2 lo = *reinterpret_cast<__half2*>(&packed.x), hi = *reinterpret_cast<__half2*>(&packed.y);
        float3 c = make_float3(fmaxf(__low2float(lo), 0.0f), fmaxf(__high2float(lo), 0.0f), fmaxf(__low2float(hi), 0.0f));
        // encoded = c / (c + 1); uv = encoded * 47/48 + 0.5/48 (display_transform.comp:5-20)
        float3 e = make_float3(c.x / (c.x + 1.0f), c.y / (c.y + 1.0f), c.z / (c.z + 1.0f));
        float3 mapped = e;
        if (P.lut) {
            const float s = 47.0f / 48.0f, o = 0.5f / 48.0f;
            float4 t = tex3D<float4>(P.lut, e.x * s + o, e.y * s + o, e.z * s + o);
            mapped = make_float3(t.x, t.y, t.z);
        }
        uint32_t r = (uint32_t)__float2int_rn(fminf(fmaxf(srgb_oetf(mapped.x), 0.0f), 1.0f) * 255.0f);
        uint32_t g = (uint32_t)__float2int_rn(fminf(fmaxf(srgb_oetf(mapped.y), 0.0f), 1.0f) * 255.0f);
        uint32_t b = (uint32_t)__float2int_rn(fminf(fmaxf(srgb_oetf(mapped.z), 0.0f), 1.0f) * 255.0f);
        reinterpret_cast<uint32_t*>(P.display)[i] = b | (g << 8) | (r << 16) | 0xFF000000u;  // B8G8R8A8
    }
}

}  // namespace

void launch_resolve(LhContext* ctx, const FrameParams& p, int phase, cudaStream_t stream) {
    uint32_t W = p.u.window_size[0], rows = min(p.u.window_size[1], p.resolve_y1) - p.resolve_y0;
    if (!W || !rows) return;
    if (!stream) stream = ctx->stream;
    dim3 grid((W + 31) / 32, (rows + 7) / 8);
    if (phase == 0) resolve_kernel<0><<<grid, 256, 0, stream>>>(p);
    else if (phase == 1) resolve_kernel<1><<<grid, 256, 0, stream>>>(p);
    else resolve_kernel<2><<<grid, 256, 0, stream>>>(p);
    ctx->launches++;
}

void launch_display_transform(LhContext* ctx, const FrameParams& p) {
    display_transform_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(p); ctx->launches++;
}

}  // namespace lh
