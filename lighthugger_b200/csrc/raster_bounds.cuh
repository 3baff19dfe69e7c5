// Conservative screen bound of a meshlet, shared by the meshlet cull (which computes it for every survivor while the
// instance transform and the meshlet's sphere are in registers) and the rasteriser (which orders, tests and skips
// meshlets by it).  See raster.cu for how the bounds are used.
#pragma once
#include "lh_internal.cuh"

namespace lh {

struct MeshletBound {
    uint32_t rect_x, rect_y;   // px0 | px1 << 16, py0 | py1 << 16; rect_x == kUntestable: rasterise without a test
    uint32_t edge;             // main view: ~floatBits(nearest depth); cascade: floatBits(nearest depth)
    uint32_t key;              // depth key, smaller = nearer; kKeyEmpty: the rectangle holds no pixel centre
};
constexpr uint32_t kUntestable = 0xFFFFFFFFu, kKeyEmpty = 0xFFFFFFFFu;
constexpr int kMaxBands = 8;   // bands per view: FrameParams::num_bands (LhContext::Tuning), 2..8
constexpr uint32_t kHistShift = 19, kHistBins = 1u << (32 - kHistShift);   // 8192 bins over the top bits of the depth key
constexpr int kHizShift0 = 2, kHizShift1 = 4;   // hierarchical-Z levels: 4x4-pixel and 16x16-pixel tiles
#ifndef LH_MAX_HIZ_TILES
#define LH_MAX_HIZ_TILES 24
#endif
constexpr int kMaxHizTiles = LH_MAX_HIZ_TILES;   // selection reads at most this many tiles of one level per meshlet


// the view a bound is taken in: the matrix, the lengths of its rows' xyz parts, the target size
struct BoundView {
    float M[16];
    float nrm_x, nrm_y, nrm_z, nrm_w;
    int W, H;
};

template <bool SHADOW>
__device__ __forceinline__ void make_bound_view(const FrameParams& P, uint32_t cascade, BoundView& bv) {
    const float* src = SHADOW ? &((const LhMiscStorage*)P.u.misc_storage)->shadow_matrices[cascade].m[0][0] : &P.u.combined_perspective_view.m[0][0];
    for (int k = 0; k < 16; k++) bv.M[k] = src[k];
    const float* M = bv.M;
    bv.nrm_x = sqrtf(M[0] * M[0] + M[4] * M[4] + M[8] * M[8]);
    bv.nrm_y = sqrtf(M[1] * M[1] + M[5] * M[5] + M[9] * M[9]);
    bv.nrm_z = sqrtf(M[2] * M[2] + M[6] * M[6] + M[10] * M[10]);
    bv.nrm_w = sqrtf(M[3] * M[3] + M[7] * M[7] + M[11] * M[11]);
    bv.W = (int)(SHADOW ? P.shadowmap_size : P.u.window_size[0]);
    bv.H = (int)(SHADOW ? P.shadowmap_size : P.u.window_size[1]);
}

// sigma_max(T) * r bounds |T (v - c)| for v in the object-space sphere (c, r).  Gershgorin on the Gram matrix and the
// Frobenius norm are both valid for any matrix; for rotation * scale the first equals the reference's max column length.
// T here is the 3x3 part + translation of the instance transform, columns packed: T[0..2], T[3..5], T[6..8], T[9..11].
__device__ __forceinline__ float sigma_max_bound(const float* T) {
    const float c00 = T[0] * T[0] + T[1] * T[1] + T[2] * T[2], c11 = T[3] * T[3] + T[4] * T[4] + T[5] * T[5],
                c22 = T[6] * T[6] + T[7] * T[7] + T[8] * T[8];
    const float c01 = fabsf(T[0] * T[3] + T[1] * T[4] + T[2] * T[5]), c02 = fabsf(T[0] * T[6] + T[1] * T[7] + T[2] * T[8]),
                c12 = fabsf(T[3] * T[6] + T[4] * T[7] + T[5] * T[8]);
    const float gersh = fmaxf(fmaxf(c00 + c01 + c02, c11 + c01 + c12), c22 + c02 + c12);
    return sqrtf(fminf(gersh, c00 + c11 + c22)) * 1.00001f;
}


// T16: the instance transform (column-major 4x4); sphere: the meshlet's object-space bounding sphere.
// Interval arithmetic through T and the view's matrix: a conservative pixel rectangle, the nearest depth the meshlet
// can produce ("edge") and a depth key.
__device__ __forceinline__ float sigma_max_bound16(const float* T16) {
    const float T[9] = {T16[0], T16[1], T16[2], T16[4], T16[5], T16[6], T16[8], T16[9], T16[10]};
    return sigma_max_bound(T);
}

// sigma: sigma_max_bound of T16's 3x3 part (a per-instance value the caller may have at hand), or a negative number
template <bool SHADOW>
__device__ __forceinline__ MeshletBound meshlet_bound(const BoundView& bv, const float* T16, const float* sphere, float sigma = -1.0f) {
    const float* M = bv.M;
    const int W = bv.W, H = bv.H;
    const float half_w = (float)W * 0.5f, half_h = (float)H * 0.5f;
    const float T[12] = {T16[0], T16[1], T16[2], T16[4], T16[5], T16[6], T16[8], T16[9], T16[10], T16[12], T16[13], T16[14]};
    MeshletBound mb{kUntestable, 0u, 0u, 0u};   // no usable bound: nearest band, rasterised without a test
    const float R0 = sphere[3] * (sigma >= 0.0f ? sigma : sigma_max_bound(T));
    const float wx = T[0] * sphere[0] + T[3] * sphere[1] + T[6] * sphere[2] + T[9];
    const float wy = T[1] * sphere[0] + T[4] * sphere[1] + T[7] * sphere[2] + T[10];
    const float wz = T[2] * sphere[0] + T[5] * sphere[1] + T[8] * sphere[2] + T[11];
    // radius padded for the rounding of the vertex transform itself (a few ulp of the terms' magnitude)
    const float R = R0 * 1.0001f + 2e-6f * (fabsf(wx) + fabsf(wy) + fabsf(wz) + fabsf(T[9]) + fabsf(T[10]) + fabsf(T[11]));
    const float ccx = M[0] * wx + M[4] * wy + M[8] * wz + M[12], ccy = M[1] * wx + M[5] * wy + M[9] * wz + M[13];
    const float ccz = M[2] * wx + M[6] * wy + M[10] * wz + M[14], ccw = M[3] * wx + M[7] * wy + M[11] * wz + M[15];
    const float ex = R * bv.nrm_x, ey = R * bv.nrm_y, ez = R * bv.nrm_z, ew = R * bv.nrm_w;
    const float w_lo = ccw - ew, w_hi = ccw + ew;
    if (w_lo > 1e-6f) {   // otherwise the sphere reaches the w = 0 plane: no bound
        const float x_lo = ccx - ex, x_hi = ccx + ex, y_lo = ccy - ey, y_hi = ccy + ey;
        const float nx_lo = x_lo / (x_lo >= 0.0f ? w_hi : w_lo), nx_hi = x_hi / (x_hi >= 0.0f ? w_lo : w_hi);
        const float ny_lo = y_lo / (y_lo >= 0.0f ? w_hi : w_lo), ny_hi = y_hi / (y_hi >= 0.0f ? w_lo : w_hi);
        float fx_lo = nx_lo * half_w + half_w, fx_hi = nx_hi * half_w + half_w;
        float fy_lo = ny_lo * half_h + half_h, fy_hi = ny_hi * half_h + half_h;
        // nearest depth the meshlet can produce: largest z/w for the reverse-Z view, smallest for a cascade
        const float z_edge = SHADOW ? ccz - ez : ccz + ez;
        float d_edge = SHADOW ? z_edge / (z_edge >= 0.0f ? w_hi : w_lo) : z_edge / (z_edge >= 0.0f ? w_lo : w_hi);
        if (fabsf(fx_lo) + fabsf(fx_hi) + fabsf(fy_lo) + fabsf(fy_hi) + fabsf(d_edge) < 3.0e38f) {
            fx_lo -= 0.125f + 1e-5f * fabsf(fx_lo); fx_hi += 0.125f + 1e-5f * fabsf(fx_hi);
            fy_lo -= 0.125f + 1e-5f * fabsf(fy_lo); fy_hi += 0.125f + 1e-5f * fabsf(fy_hi);
            const int px0 = max(0, (int)ceilf(fx_lo - 0.5f)), px1 = min(W - 1, (int)floorf(fx_hi - 0.5f));
            const int py0 = max(0, (int)ceilf(fy_lo - 0.5f)), py1 = min(H - 1, (int)floorf(fy_hi - 0.5f));
            // depth key, smaller = nearer.  Main view: float bits of the view distance of the nearest point (bins are
            // logarithmic in distance); cascade: the nearest light-space depth in [0,1) as 31-bit fixed point (linear bins)
            mb.key = SHADOW ? (uint32_t)(fminf(fmaxf(d_edge, 0.0f), 0.99999994f) * 2147483648.0f) : min(__float_as_uint(w_lo), 0x7F800000u);
            if (px0 > px1 || py0 > py1) {
                mb.key = kKeyEmpty;   // no pixel centre inside the bound
            } else {
                d_edge = SHADOW ? d_edge - (4e-5f * fabsf(d_edge) + 1e-7f) : d_edge + (4e-5f * fabsf(d_edge) + 1e-30f);
                const int ntx = (px1 >> kHizShift1) - (px0 >> kHizShift1) + 1, nty = (py1 >> kHizShift1) - (py0 >> kHizShift1) + 1;
                if (ntx * nty <= kMaxHizTiles && d_edge > 0.0f) {
                    mb.rect_x = (uint32_t)px0 | ((uint32_t)px1 << 16);
                    mb.rect_y = (uint32_t)py0 | ((uint32_t)py1 << 16);
                    // main view: a fragment of depth d loses when floatBits(d) < ~hi(key) for the stored key, i.e.
                    // hi(key) < ~floatBits(d); cascade: it loses when d >= stored depth (LESS test)
                    mb.edge = SHADOW ? __float_as_uint(d_edge) : ~__float_as_uint(d_edge);
                }
            }
        }
    }
    return mb;
}

}  // namespace lh
