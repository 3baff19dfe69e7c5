// Software rasteriser: visibility buffer (packed 64-bit depth|slot|triangle
// words resolved by atomicMin) and the four shadow-map layers (u32 atomicMin
// on float bits).
//
// Replaces the reference's hardware raster passes (file:line)
//   load_vertex_data / visbuffer_* / shadowmap_*   src/shaders/rasterization.glsl:12-145
//   calculate_world_pos, load_index, pack          src/shaders/common/util.glsl:3-20
//   fixed-function state                           src/pipelines.cpp:10-25,67-77,255-315
//   clears, drawIndirectCount                      src/rendering.cpp:195-266,379-436
//
// Rasterisation rules (Vulkan; pinned in DESIGN.md): clip against w - z >= 0,
// ndc = clip / w, viewport transform, snap to 8 sub-pixel bits (round half
// even), integer edge functions, samples at pixel centres, top-left rule,
// front face = counter-clockwise (A2 < 0), screen-linear depth.
//
// B200 design.  The reference issues one indirect draw per surviving meshlet
// and runs the vertex shader three times per triangle through a 6-deep
// pointer chase.  Here one WARP owns a meshlet: its <= 64 vertices are
// transformed once into shared memory (2 per lane), then every lane sets up
// triangles (stride 32).  Sub-pixel and small triangles -- the overwhelming
// majority at 1 G triangles per frame -- are finished by their lane with
// 32-bit edge functions; triangles with a large footprint or crossing the
// near plane are deferred, one queue entry per 128x32-pixel screen tile, to a
// second kernel where a whole warp scans a tile in 8x4-pixel blocks with
// 64-bit edge functions.  Both paths evaluate identical integer coverage and
// the identical depth expression, so the result is independent of the split.
#include "lh_internal.cuh"
#include "raster_bounds.cuh"

namespace lh {

namespace {

#ifndef LH_WARPS_PER_CTA
#define LH_WARPS_PER_CTA 8
#endif
constexpr int kWarpsPerCta = LH_WARPS_PER_CTA;
constexpr uint32_t kTileW = 128, kTileH = 32;   // deferred-triangle screen tile
constexpr uint32_t kWholeTriangle = 0xFFFFFFFFu;
constexpr int kSmallMaxSpan = 32;               // small path: pixel bbox < 32 on each axis ...
// ... and at most FrameParams::small_max_area pixel tests (LhContext::Tuning, default 128; swept on B200: 16..1024)
constexpr int kSmallMaxExtent = 1 << 14;        // ... and sub-pixel extent < 2^14 (32-bit edge functions stay exact)

enum { CULL_NONE = 0, CULL_BACK = 1, CULL_FRONT = 2 };
enum { FLAG_INSIDE = 1, FLAG_SNAPPED = 2 };

// What the band selection hands to the raster kernel about one queued meshlet (64 bytes): the end of the pointer
// chain reference -> instance -> mesh info -> meshlet, walked once by the selection's lane (one lane per meshlet, so
// the chain's dependent round trips hide behind the other lanes' work), and the meshlet's bound.  The raster kernel
// then never chases a pointer: it pulls records, and through them the meshlet's micro-indices, vertex indices and
// instance transform, into shared memory with bulk-asynchronous copies (cp.async.bulk + mbarrier) ahead of use.
struct alignas(16) RasterRecord {
    const uint16_t* positions;
    const uint8_t* indices;        // byte address of the meshlet's first vertex index (u16 or u32)
    const uint8_t* micro;          // the meshlet's first micro-index
    const LhMeshInfo* mi;
    const LhInstance* instance;    // 16-byte aligned; the transform is its first 64 bytes
    uint32_t counts_flags;         // triangle_count | index_count << 8 | MeshInfo.flags << 16
    uint32_t local;                // list index of the meshlet (its slot minus the view's slot base)
    uint32_t rect_x, rect_y, edge; // MeshletBound
    uint32_t pad;
};
static_assert(sizeof(RasterRecord) == 64, "records are moved in 16-byte units");

#ifndef LH_CHUNK
#define LH_CHUNK 2
#endif
constexpr int kChunk = LH_CHUNK;   // queue entries a warp takes per ticket
// LH_STAGE_STREAMS: the chunk's micro-indices, vertex indices and instance transforms are copied to shared memory
// (cp.async, 16 bytes per lane) as soon as its records have landed, instead of being fetched where they are used
#ifndef LH_STAGE_STREAMS
#define LH_STAGE_STREAMS 1
#endif
// LH_RECORD_BULK: the chunk's records arrive by ONE bulk-asynchronous copy (cp.async.bulk + mbarrier, issued by an
// elected lane) instead of four 16-byte loads + stores by each of the chunk's lanes
#ifndef LH_RECORD_BULK
#define LH_RECORD_BULK 0
#endif
constexpr bool kStageStreams = LH_STAGE_STREAMS != 0, kRecordBulk = LH_RECORD_BULK != 0;
// a meshlet's micro-indices (<= 124 triangles x 3 bytes) and vertex indices (<= 64 x 4 bytes) start at any byte: the
// copies are the enclosing 16-byte-aligned ranges
constexpr int kMicroBuf = (15 + 3 * (int)LH_MAX_MESHLET_TRIANGLES + 15) & ~15;           // 400
constexpr int kIndexBuf = (15 + 4 * (int)LH_MAX_MESHLET_UNIQUE_VERTICES + 15) & ~15;     // 272

// Per-warp scratch of the meshlet kernel.
// sx holds two sentinels instead of a flag byte: kVertOutside (behind the w - z >= 0 plane) and kVertNonFinite.
struct alignas(16) WarpScratch {
    RasterRecord rec[kChunk];
    float T[kStageStreams ? kChunk : 1][16];
    uint8_t micro[kStageStreams ? kChunk : 1][kMicroBuf];
    uint8_t index[kStageStreams ? kChunk : 1][kIndexBuf];
    int sx[LH_MAX_MESHLET_UNIQUE_VERTICES];
    int sy[LH_MAX_MESHLET_UNIQUE_VERTICES];
    float z[LH_MAX_MESHLET_UNIQUE_VERTICES];
    float w[LH_MAX_MESHLET_UNIQUE_VERTICES];   // clip w, written for alpha-clip meshlets only (perspective-correct uv)
    // per vertex, packed int16 pairs (x | y << 16): .x = first pixel whose centre is >= the vertex, .y = last pixel whose
    // centre is <= it.  min/max over a triangle's vertices (SIMD halfword ops) give its pixel rectangle directly.
    uint2 cover[LH_MAX_MESHLET_UNIQUE_VERTICES];
    uint8_t surv[128];
    unsigned long long rec_bar;   // mbarrier: the chunk's records have landed (LH_RECORD_BULK)
};
constexpr int kVertOutside = INT_MIN, kVertNonFinite = INT_MIN + 1;   // |snapped| <= 2^29, so neither can be a coordinate
#ifndef LH_SELECT_EXACT_PIXELS
#define LH_SELECT_EXACT_PIXELS 16
#endif
constexpr int kSelectExactPixels = LH_SELECT_EXACT_PIXELS;   // selection tests rectangles of at most this many pixels against the stored values themselves (0: hierarchical Z only)
#ifndef LH_SELECT_CTA_CLAIM
#define LH_SELECT_CTA_CLAIM 0
#endif
constexpr bool kSelectCtaClaim = LH_SELECT_CTA_CLAIM != 0;   // the selection claims queue space once per CTA and iteration instead of once per warp (measured: -1.5 % at C2, the two barriers cost more than the atomics)
#ifndef LH_STATIC_FIRST
#define LH_STATIC_FIRST 1
#endif
constexpr bool kStaticFirst = LH_STATIC_FIRST != 0;          // first-wave CTAs of the raster kernel take their first chunk without a ticket
#ifndef LH_LANE_PINNED
#define LH_LANE_PINNED 1
#endif
#ifndef LH_FEW_SAMPLES_PATH
#define LH_FEW_SAMPLES_PATH 0
#endif
constexpr bool kFewSamplesPath = LH_FEW_SAMPLES_PATH != 0;   // meshlets touching at most 2x2 samples skip the per-triangle rectangle / compaction machinery
constexpr int kMaxOcclusionBlocks = 16;    // early-out reads at most 16 blocks of 8x4 pixels (512 keys)

// ---- bulk-asynchronous copies global -> shared, completion counted in bytes on an mbarrier ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(void* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LH_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LH_MBAR_DONE;\n"
        "bra LH_MBAR_WAIT;\n"
        "LH_MBAR_DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 16 bytes global -> shared, asynchronously (cp.async: the data bypasses the register file and L1)
__device__ __forceinline__ void async_copy16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void async_copy_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// dst, src and bytes are multiples of 16
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct SnapVert {
    int x, y;
    float z, w, u, v;
};

// (sat16(hi) << 16) | (sat16(lo) & 0xFFFF)
__device__ __forceinline__ uint32_t pack_sat_s16(int hi, int lo) {
    uint32_t d;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(d) : "r"(hi), "r"(lo));
    return d;
}

// project_snap of the pinned model: false when the vertex is not finite
__device__ __forceinline__ bool project_snap(float cx, float cy, float cz, float cw, float half_w, float half_h, int* sx, int* sy, float* z) {
    // x / 1.0f == x exactly: an orthographic (cascade) matrix yields w == 1 and skips the three IEEE divisions
    float ndc_x = cx, ndc_y = cy;
    *z = cz;
    if (cw != 1.0f) {
        ndc_x = pdiv(cx, cw);
        ndc_y = pdiv(cy, cw);
        *z = pdiv(cz, cw);
    }
    float fx = pmul(padd(pmul(ndc_x, half_w), half_w), 256.0f);
    float fy = pmul(padd(pmul(ndc_y, half_h), half_h), 256.0f);
    if (!(isfinite(fx) && isfinite(fy) && isfinite(*z))) return false;
    const float lim = 536870912.0f;  // 2^29
    fx = fminf(fmaxf(fx, -lim), lim);
    fy = fminf(fmaxf(fy, -lim), lim);
    *sx = __float2int_rn(fx);
    *sy = __float2int_rn(fy);
    return true;
}

template <bool SHADOW>
struct Target {
    unsigned long long* vis;
    uint32_t* layer;
    uint32_t width, height;
    bool count;                 // LH_CONFIG_COUNT_SAMPLES
    mutable uint32_t samples;   // reductions this thread has issued (counted only then)
    // depth test + write of one covered sample
    __device__ __forceinline__ void emit(int x, int y, float z, uint32_t lo) const {
        if (count) samples++;
        if (SHADOW) {
            z = padd(z, 0.0f);                      // canonicalise -0
            if (!(z >= 0.0f && z < 1.0f)) return;   // 0 <= z <= w clip; LESS against the 1.0 clear
            uint32_t bits = __float_as_uint(z);
            // fire-and-forget reduction: no read-back, so a lane never waits on the L2 round trip of its own sample
            atomicMin(layer + (size_t)y * width + x, bits);
        } else {
            z = fminf(fmaxf(z, 0.0f), 1.0f);
            if (!(z > 0.0f)) return;                // GREATER against the 0.0 clear
            unsigned long long key = ((unsigned long long)(~__float_as_uint(z)) << 32) | lo;
            atomicMin(vis + (size_t)y * width + x, key);
        }
    }
};

struct AlphaTest {
    bool enabled;
    cudaTextureObject_t tex;
};

__device__ __forceinline__ bool top_left(long long dx, long long dy) { return dy < 0 || (dy == 0 && dx > 0); }

// visbuffer_alpha_clip_pixel (rasterization.glsl:90-97): perspective-correct uv, alpha < 0.5 discards
__device__ __forceinline__ bool alpha_discard(const AlphaTest& at, float w0, float w1, float w2, float fA, const SnapVert& a,
                                              const SnapVert& b, const SnapVert& c) {
    float b0 = pdiv(w0, fA), b1 = pdiv(w1, fA), b2 = pdiv(w2, fA);
    float q0 = pdiv(b0, a.w), q1 = pdiv(b1, b.w), q2 = pdiv(b2, c.w);
    float den = padd(padd(q0, q1), q2);
    float tu = pdiv(padd(padd(pmul(q0, a.u), pmul(q1, b.u)), pmul(q2, c.u)), den);
    float tv = pdiv(padd(padd(pmul(q0, a.v), pmul(q1, b.v)), pmul(q2, c.v)), den);
    return tex2DLod<float4>(at.tex, tu, tv, 0.0f).w < 0.5f;
}

// ---- warp-cooperative rasterisation of one snapped triangle inside a pixel rectangle ----
// I: I in general; int when the triangle's sub-pixel extent is below 2^14 on both axes (every product and every
// stepped edge value then stays below 2^31, so both instantiations compute the same integers)
template <bool SHADOW, typename I>
__device__ void raster_snapped_coop(SnapVert a, SnapVert b, SnapVert c, int cull, const Target<SHADOW>& tg, int rx0, int ry0, int rx1,
                                    int ry1, const AlphaTest& at, uint32_t lo) {
    const uint32_t lane = threadIdx.x & 31;
    I A2 = (I)(b.x - a.x) * (c.y - a.y) - (I)(b.y - a.y) * (c.x - a.x);
    if (A2 == 0) return;
    bool front = A2 < 0;
    if (cull == CULL_BACK && !front) return;
    if (cull == CULL_FRONT && front) return;
    if (A2 < 0) {
        SnapVert t = b;
        b = c;
        c = t;
        A2 = -A2;
    }
    int minx = min(a.x, min(b.x, c.x)), maxx = max(a.x, max(b.x, c.x));
    int miny = min(a.y, min(b.y, c.y)), maxy = max(a.y, max(b.y, c.y));
    int px0 = max(rx0, (minx - 128 + 255) >> 8), px1 = min(rx1, (maxx - 128) >> 8);
    int py0 = max(ry0, (miny - 128 + 255) >> 8), py1 = min(ry1, (maxy - 128) >> 8);
    if (px0 > px1 || py0 > py1) return;
    const I e0x = (I)c.x - b.x, e0y = (I)c.y - b.y;
    const I e1x = (I)a.x - c.x, e1y = (I)a.y - c.y;
    const I e2x = (I)b.x - a.x, e2y = (I)b.y - a.y;
    const I bias0 = top_left(e0x, e0y) ? 0 : -1, bias1 = top_left(e1x, e1y) ? 0 : -1, bias2 = top_left(e2x, e2y) ? 0 : -1;
    const float fA = (float)A2;
    const float inv_area = pdiv(1.0f, fA);
    const float dz1 = psub(b.z, a.z), dz2 = psub(c.z, a.z);
    const int nbx = (px1 - px0 + 8) >> 3, nby = (py1 - py0 + 4) >> 2;
    // The edge functions are affine in the pixel position, so everything below is additions: E_k at the centre of the
    // block's first pixel is carried from block to block (step_x per 8 pixels, step_y per 4 rows); the lane's own pixel
    // and the block corner where E_k is largest (for the trivial reject) are fixed offsets from it.
    const I cx0 = (I)px0 * 256 + 128, cy0 = (I)py0 * 256 + 128;
    const I dx0 = -e0y * 256, dx1 = -e1y * 256, dx2 = -e2y * 256;   // E_k(x + 1 pixel) - E_k(x)
    const I dy0 = e0x * 256, dy1 = e1x * 256, dy2 = e2x * 256;      // E_k(y + 1 pixel) - E_k(y)
    I row0 = e0x * (cy0 - b.y) - e0y * (cx0 - b.x) + bias0;          // top-left bias folded in: inside <=> all >= 0
    I row1 = e1x * (cy0 - c.y) - e1y * (cx0 - c.x) + bias1;
    I row2 = e2x * (cy0 - a.y) - e2y * (cx0 - a.x) + bias2;
    const I corner0 = (dx0 > 0 ? 7 * dx0 : 0) + (dy0 > 0 ? 3 * dy0 : 0);
    const I corner1 = (dx1 > 0 ? 7 * dx1 : 0) + (dy1 > 0 ? 3 * dy1 : 0);
    const I corner2 = (dx2 > 0 ? 7 * dx2 : 0) + (dy2 > 0 ? 3 * dy2 : 0);
    const int lx = (int)(lane & 7), ly = (int)(lane >> 3);
    const I mine0 = lx * dx0 + ly * dy0, mine1 = lx * dx1 + ly * dy1, mine2 = lx * dx2 + ly * dy2;
    for (int by = 0; by < nby; by++) {
        I blk0 = row0, blk1 = row1, blk2 = row2;
        const int py = py0 + by * 4 + ly;
        for (int bx = 0; bx < nbx; bx++) {
            // trivial reject: the block corner where each (biased) edge function is largest is still outside
            const bool reject = ((blk0 + corner0) | (blk1 + corner1) | (blk2 + corner2)) < 0;
            const I w0b = blk0 + mine0, w1b = blk1 + mine1, w2b = blk2 + mine2;
            blk0 += 8 * dx0;
            blk1 += 8 * dx1;
            blk2 += 8 * dx2;
            if (reject) continue;
            const int px = px0 + bx * 8 + lx;
            if (px > px1 || py > py1) continue;
            if ((w0b | w1b | w2b) < 0) continue;
            const I w0 = w0b - bias0, w1 = w1b - bias1, w2 = w2b - bias2;
            if (at.enabled && alpha_discard(at, (float)w0, (float)w1, (float)w2, fA, a, b, c)) continue;
            float z = padd(a.z, pmul(padd(pmul((float)w1, dz1), pmul((float)w2, dz2)), inv_area));
            tg.emit(px, py, z, lo);
        }
        row0 += 4 * dy0;
        row1 += 4 * dy1;
        row2 += 4 * dy2;
    }
}

struct ClipVert {
    float x, y, z, w, u, v;
};

// The generic path of the pinned model: near-plane clip (w - z >= 0; a cascade has no such plane) and snap.
// Returns the number of vertices of the clipped polygon in s[] (0: nothing to draw, else 3 or 4).
template <bool SHADOW>
__device__ int clip_and_snap(const ClipVert v[3], const Target<SHADOW>& tg, SnapVert s[4]) {
    const float half_w = pmul((float)tg.width, 0.5f), half_h = pmul((float)tg.height, 0.5f);
    bool in[3] = {true, true, true};
    float d[3] = {0, 0, 0};
    int n_in = 3;
    if (!SHADOW) {
        n_in = 0;
        for (int k = 0; k < 3; k++) {
            d[k] = psub(v[k].w, v[k].z);
            in[k] = d[k] >= 0.0f;
            n_in += in[k] ? 1 : 0;
        }
    }
    if (n_in == 0) return 0;
    ClipVert poly[4];
    int n = 0;
    if (n_in == 3) {
        poly[0] = v[0];
        poly[1] = v[1];
        poly[2] = v[2];
        n = 3;
    } else {
        for (int e = 0; e < 3; e++) {
            int ia = e, ib = (e + 1) % 3;
            if (in[ia]) poly[n++] = v[ia];
            if (in[ia] != in[ib]) {
                int ii = in[ia] ? ia : ib, io = in[ia] ? ib : ia;
                float t = pdiv(d[ii], psub(d[ii], d[io]));
                const ClipVert& I = v[ii];
                const ClipVert& O = v[io];
                poly[n++] = ClipVert{padd(I.x, pmul(t, psub(O.x, I.x))), padd(I.y, pmul(t, psub(O.y, I.y))),
                                     padd(I.z, pmul(t, psub(O.z, I.z))), padd(I.w, pmul(t, psub(O.w, I.w))),
                                     padd(I.u, pmul(t, psub(O.u, I.u))), padd(I.v, pmul(t, psub(O.v, I.v)))};
            }
        }
    }
    for (int k = 0; k < n; k++) {
        if (!project_snap(poly[k].x, poly[k].y, poly[k].z, poly[k].w, half_w, half_h, &s[k].x, &s[k].y, &s[k].z)) return 0;
        s[k].w = poly[k].w;
        s[k].u = poly[k].u;
        s[k].v = poly[k].v;
    }
    return n;
}

// Everything needed to locate one meshlet's data; loaded by every lane from the same addresses (broadcast).
struct MeshletHeader {
    float T[16];
    const uint16_t* positions;
    const void* indices;
    const uint16_t* uvs;
    const uint8_t* micro;      // already offset by triangle_offset
    uint32_t index_offset;
    uint32_t tri_count, vert_count;
    bool wide, alpha;
    float tex_scale[2], tex_offset[2];
    uint32_t base_color_texture;
};

__device__ __forceinline__ void load_header(const FrameParams& P, uint32_t slot, MeshletHeader& h) {
    const LhMeshletReference* refs = (const LhMeshletReference*)P.pass_refs;
    const unsigned long long ref = __ldg((const unsigned long long*)(refs + slot));
    const uint32_t instance_index = (uint32_t)ref, meshlet_index = (uint32_t)(ref >> 32) & 0xFFFFu;
    const LhInstance* inst = (const LhInstance*)P.u.instances + instance_index;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        float4 c = __ldg((const float4*)&inst->transform + k);
        h.T[4 * k] = c.x; h.T[4 * k + 1] = c.y; h.T[4 * k + 2] = c.z; h.T[4 * k + 3] = c.w;
    }
    const LhMeshInfo* mi = (const LhMeshInfo*)__ldg((const unsigned long long*)&inst->mesh_info_address);
    h.positions = (const uint16_t*)__ldg((const unsigned long long*)&mi->positions);
    h.indices = (const void*)__ldg((const unsigned long long*)&mi->indices);
    h.uvs = (const uint16_t*)__ldg((const unsigned long long*)&mi->uvs);
    const uint8_t* micro = (const uint8_t*)__ldg((const unsigned long long*)&mi->micro_indices);
    const LhMeshlet* m = (const LhMeshlet*)__ldg((const unsigned long long*)&mi->meshlets) + meshlet_index;
    const uint32_t flags = __ldg(&mi->flags);
    h.wide = (flags & LH_MESH_INFO_FLAGS_32_BIT_INDICES) != 0;
    h.alpha = (flags & LH_MESH_INFO_FLAGS_ALPHA_CLIP) != 0;
    const unsigned long long tail = __ldg((const unsigned long long*)m + 6);  // index_offset | counts
    h.micro = micro + __ldg(&m->triangle_offset);
    h.index_offset = (uint32_t)tail;
    h.tri_count = (uint32_t)(tail >> 32) & 0xFFu;
    h.vert_count = (uint32_t)(tail >> 40) & 0xFFu;
    h.tex_scale[0] = __ldg(&mi->texture_scale[0]);
    h.tex_scale[1] = __ldg(&mi->texture_scale[1]);
    h.tex_offset[0] = __ldg(&mi->texture_offset[0]);
    h.tex_offset[1] = __ldg(&mi->texture_offset[1]);
    h.base_color_texture = __ldg(&mi->base_color_texture_index);
}

// The pointer chase of load_header by ONE lane of the selection, left as a RasterRecord for the raster kernel
__device__ __forceinline__ void make_record(const FrameParams& P, uint32_t k, uint4 mb, RasterRecord* out) {
    const LhMeshletReference* refs = (const LhMeshletReference*)P.pass_refs;
    const unsigned long long ref = __ldg((const unsigned long long*)(refs + k));
    const uint32_t instance_index = (uint32_t)ref, meshlet_index = (uint32_t)(ref >> 32) & 0xFFFFu;
    const LhInstance* inst = (const LhInstance*)P.u.instances + instance_index;
    const LhMeshInfo* mi = (const LhMeshInfo*)__ldg((const unsigned long long*)&inst->mesh_info_address);
    const unsigned long long positions = __ldg((const unsigned long long*)&mi->positions), indices = __ldg((const unsigned long long*)&mi->indices);
    const unsigned long long micro = __ldg((const unsigned long long*)&mi->micro_indices), meshlets = __ldg((const unsigned long long*)&mi->meshlets);
    const uint32_t flags = __ldg(&mi->flags);
    const LhMeshlet* m = (const LhMeshlet*)meshlets + meshlet_index;
    const unsigned long long tail = __ldg((const unsigned long long*)m + 6);  // index_offset | counts
    const uint32_t triangle_offset = __ldg(&m->triangle_offset);
    const unsigned long long first_index = indices + (unsigned long long)(uint32_t)tail * ((flags & LH_MESH_INFO_FLAGS_32_BIT_INDICES) ? 4u : 2u);
    uint4* dst = (uint4*)out;
    dst[0] = make_uint4((uint32_t)positions, (uint32_t)(positions >> 32), (uint32_t)first_index, (uint32_t)(first_index >> 32));
    const unsigned long long micro_first = micro + triangle_offset, mi_bits = (unsigned long long)(uintptr_t)mi;
    dst[1] = make_uint4((uint32_t)micro_first, (uint32_t)(micro_first >> 32), (uint32_t)mi_bits, (uint32_t)(mi_bits >> 32));
    const unsigned long long inst_bits = (unsigned long long)(uintptr_t)inst;
    dst[2] = make_uint4((uint32_t)inst_bits, (uint32_t)(inst_bits >> 32), ((uint32_t)(tail >> 32) & 0xFFFFu) | (flags << 16), k);
    dst[3] = make_uint4(mb.x, mb.y, mb.z, 0u);
}

__device__ __forceinline__ uint32_t load_vertex_index(const MeshletHeader& h, uint32_t v) {
    return h.wide ? __ldg((const uint32_t*)h.indices + h.index_offset + v) : (uint32_t)__ldg((const uint16_t*)h.indices + h.index_offset + v);
}

__device__ __forceinline__ AlphaTest make_alpha_test(const FrameParams& P, bool alpha, uint32_t base_color_texture) {
    AlphaTest at{false, 0};
    if (alpha && base_color_texture < LH_MAX_TEXTURES) {
        at.tex = P.textures[base_color_texture];
        at.enabled = at.tex != 0;
    }
    return at;
}

__device__ __forceinline__ void push_deferred(const FrameParams& P, uint32_t lo, uint32_t tile) {
    uint32_t at = atomicAdd(&P.pass->large_count, 1u);
    if (at < P.max_large) {
        P.large_queue[2 * (size_t)at] = lo;
        P.large_queue[2 * (size_t)at + 1] = tile;
    } else {
        atomicOr(&P.state->overflow, (uint32_t)LH_OVERFLOW_LARGE_TRIANGLES);
    }
}

// ---- the meshlet rasteriser: bounds -> depth bands -> (select, rasterise) per band ----
//
// The reference's hardware rasteriser rejects hidden geometry early (hierarchical Z); a software rasteriser that
// pushes 10^8 sub-pixel triangles through 64-bit atomics needs the same, and gets it at meshlet granularity:
//
//   (cull_meshlets_kernel) for every survivor, while its transform and sphere are in registers: interval arithmetic
//                          through the view's matrix on the bounding sphere gives a conservative pixel rectangle, a
//                          conservative nearest depth ("edge") and a depth key (raster_bounds.cuh); stored as a
//                          16-byte MeshletBound next to the survivor's list entry, keys histogrammed.
//   raster_bands_kernel    one CTA: histogram -> kBands key thresholds at fixed count fractions (nearest band small).
//   raster_select_kernel   per band, nearest first: meshlets of the band whose rectangle holds a pixel centre and in
//                          which some stored key is NOT already nearer than the edge depth are appended to a queue.
//   raster_meshlets_kernel per band: one warp per queued meshlet (ticketed), re-tests against the live buffer, then
//                          vertex phase (lane per vertex) -> triangle cull phase (lane per triangle, ballot compaction)
//                          -> raster phase over the compacted survivors.
//
// Keys only ever decrease and a dropped meshlet could not have won any pixel, so the final buffer is bit-identical to
// rasterising everything (tests compare it with the oracle, which does exactly that).

// One CTA of 1024 threads: scans the key histogram and writes the kBands upper key bounds (exclusive) at fixed
// fractions of the meshlet count -- a small nearest band, so that the next bands already find most of the screen
// covered -- and zeroes the histogram for the next pass.
__global__ void __launch_bounds__(1024) raster_bands_kernel(const __grid_constant__ FrameParams P) {
    __shared__ uint32_t partial[1024];
    __shared__ uint32_t grand_total;
    const uint32_t tid = threadIdx.x;
    constexpr uint32_t kPer = kHistBins / 1024;   // 8 consecutive bins per thread: two coalesced 16-byte loads
    uint32_t local[kPer], sum = 0;
    {
        uint4* src = (uint4*)(P.depth_hist + tid * kPer);
        const uint4 a = src[0], b = src[1];
        src[0] = make_uint4(0, 0, 0, 0);
        src[1] = make_uint4(0, 0, 0, 0);
        local[0] = a.x; local[1] = a.y; local[2] = a.z; local[3] = a.w;
        local[4] = b.x; local[5] = b.y; local[6] = b.z; local[7] = b.w;
#pragma unroll
        for (uint32_t j = 0; j < kPer; j++) sum += local[j];
    }
    // inclusive scan of the 1024 partials: shuffles inside each warp, then the 32 warp totals by warp 0
    const uint32_t lane = tid & 31, warp = tid >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (uint32_t o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) partial[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = partial[lane];
#pragma unroll
        for (uint32_t o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, w, o);
            if (lane >= o) w += v;
        }
        partial[32 + lane] = w;   // inclusive totals of the warps
        if (lane == 31) grand_total = w;
    }
    __syncthreads();
    const uint32_t total = grand_total, before = incl - sum + (warp ? partial[32 + warp - 1] : 0u);
    // cumulative count fractions at which the bands end, for each band count
    const float tables[kMaxBands + 1][kMaxBands - 1] = {
        {}, {}, {0.10f}, {0.04f, 0.25f}, {0.02f, 0.08f, 0.30f}, {0.025f, 0.09f, 0.25f, 0.55f}, {0.02f, 0.07f, 0.17f, 0.35f, 0.60f},
        {0.015f, 0.05f, 0.11f, 0.22f, 0.40f, 0.65f}, {0.01f, 0.03f, 0.07f, 0.14f, 0.26f, 0.42f, 0.65f}};
    const int kBands = (int)min(max(P.num_bands, 2u), (uint32_t)kMaxBands);
    const float* fractions = P.band_fractions[0] > 0.0f ? P.band_fractions : tables[kBands];
    PassState* st = P.pass;
    if (tid == 0) {
        st->band_key[kBands - 1] = kKeyEmpty;   // the last band takes the rest (empty meshlets carry kKeyEmpty and are never selected)
        st->queue_count = 0;
        st->band_mark[0] = 0;
        // a raster pass begins: empty deferred queue
        st->large_count = st->split_count = 0;
        st->drain_end[0] = st->split_end[0] = 0;
        st->cta_done = st->cta_done_deferred = 0;
    }
    if (tid < 8) st->band_ticket[tid] = 0;
    for (int b = 0; b < kBands - 1; b++) {
        const uint32_t target = (uint32_t)((float)total * fractions[b]);
        // the thread whose bins hold the target-th meshlet names the bin after it as the band's upper bound
        if (sum && target >= before && target < before + sum) {
            uint32_t run = before, j = 0;
            for (; j < kPer; j++) {
                run += local[j];
                if (run > target) break;
            }
            st->band_key[b] = (tid * kPer + min(j, kPer - 1) + 1u) << kHistShift;
        }
        if (total == 0 && tid == 0) st->band_key[b] = 0u;
    }
}

// Hierarchical Z, two levels: per 4x4-pixel and per 16x16-pixel tile the FARTHEST value stored in it (main view:
// largest key high word, all-ones while any pixel is uncovered; cascade: largest depth).  Built between bands; stored
// values only get nearer afterwards, so a tile value stays a valid bound.  One thread per 4x4 tile (rows read as two
// 16-byte loads, a warp row covering 256 contiguous bytes); a warp owns 8x4 fine tiles = two coarse tiles, reduced by shuffles.
__device__ __forceinline__ int hiz_dim(int pixels, int shift) { return (pixels + (1 << shift) - 1) >> shift; }

template <bool SHADOW>
__global__ void __launch_bounds__(256) raster_hiz_kernel(const __grid_constant__ FrameParams P, uint32_t cascade) {
    const int W = (int)(SHADOW ? P.shadowmap_size : P.u.window_size[0]), H = (int)(SHADOW ? P.shadowmap_size : P.u.window_size[1]);
    const int TW0 = hiz_dim(W, kHizShift0), TH0 = hiz_dim(H, kHizShift0), TW1 = hiz_dim(W, kHizShift1), TH1 = hiz_dim(H, kHizShift1);
    uint32_t* level0 = P.hiz;
    uint32_t* level1 = P.hiz + (size_t)TW0 * TH0;
    const uint32_t* layer = SHADOW ? P.shadow + (size_t)cascade * P.shadowmap_size * P.shadowmap_size : nullptr;
    const uint32_t lane = threadIdx.x & 31;
    const int warps = (int)(gridDim.x * (blockDim.x >> 5));
    const int groups_x = (TW0 + 7) >> 3, groups_y = (TH0 + 3) >> 2;   // warp groups of 8x4 fine tiles
    for (int g = (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)); g < groups_x * groups_y; g += warps) {
        const int tx = (g % groups_x) * 8 + (int)(lane & 7), ty = (g / groups_x) * 4 + (int)(lane >> 3);
        uint32_t worst = 0;
        if (tx < TW0 && ty < TH0) {
            const int x0 = tx << kHizShift0, y0 = ty << kHizShift0;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int y = y0 + r;
                if (y >= H) break;
                if (x0 + 3 < W && (W & 3) == 0) {
                    if (SHADOW) {
                        const uint4 q = __ldcg((const uint4*)(layer + (size_t)y * W + x0));
                        worst = max(worst, max(max(q.x, q.y), max(q.z, q.w)));
                    } else {
                        const uint4 a = __ldcg((const uint4*)(P.vis + (size_t)y * W + x0)), b = __ldcg((const uint4*)(P.vis + (size_t)y * W + x0 + 2));
                        worst = max(worst, max(max(a.y, a.w), max(b.y, b.w)));
                    }
                } else {
                    for (int x = x0; x < min(W, x0 + 4); x++)
                        worst = max(worst, SHADOW ? __ldcg(layer + (size_t)y * W + x) : (uint32_t)(__ldcg(P.vis + (size_t)y * W + x) >> 32));
                }
            }
            level0[(size_t)ty * TW0 + tx] = worst;
        }
        // coarse tile = 4x4 fine tiles: lanes with equal bit 2 (tx >> 2) over all four rows
        uint32_t coarse = worst;
        coarse = max(coarse, __shfl_xor_sync(0xFFFFFFFFu, coarse, 1));
        coarse = max(coarse, __shfl_xor_sync(0xFFFFFFFFu, coarse, 2));
        coarse = max(coarse, __shfl_xor_sync(0xFFFFFFFFu, coarse, 8));
        coarse = max(coarse, __shfl_xor_sync(0xFFFFFFFFu, coarse, 16));
        if ((lane & 27u) == 0 && tx < TW0 && ty < TH0) level1[(size_t)(ty >> 2) * TW1 + (tx >> 2)] = coarse;
    }
}

// Band selection, one lane per meshlet: appends the band's meshlets that might still win a pixel to the raster queue.
// A meshlet is hidden when its nearest possible depth is farther than the farthest value of every tile its rectangle
// touches (fine tiles when the rectangle is small, coarse ones otherwise).
template <bool SHADOW>
__global__ void __launch_bounds__(256) raster_select_kernel(const __grid_constant__ FrameParams P, uint32_t cascade, uint32_t band) {
    const uint32_t* draw = (const uint32_t*)P.pass_draw_calls;
    const uint32_t n_opaque = min(draw[0], P.max_opaque), total = n_opaque + min(draw[1], P.max_alpha);
    const uint32_t lane = threadIdx.x & 31;
    const int W = (int)(SHADOW ? P.shadowmap_size : P.u.window_size[0]), H = (int)(SHADOW ? P.shadowmap_size : P.u.window_size[1]);
    const int TW0 = hiz_dim(W, kHizShift0), TH0 = hiz_dim(H, kHizShift0), TW1 = hiz_dim(W, kHizShift1);
    const uint32_t* level0 = P.hiz;
    const uint32_t* level1 = P.hiz + (size_t)TW0 * TH0;
    const uint32_t key_lo = band ? P.pass->band_key[band - 1] : 0u, key_hi = P.pass->band_key[band];
    const MeshletBound* bounds = (const MeshletBound*)P.meshlet_bounds;
    uint32_t occluded = 0;
    const uint32_t stride = gridDim.x * blockDim.x;
    // queue space is claimed once per CTA and iteration (eight warps' survivors in one atomic): the queue counter is a
    // single address, and one atomic per warp made it the bottleneck of the long cascade lists
    __shared__ uint32_t warp_keep[2][8], cta_at[2];
    const uint32_t warp = threadIdx.x >> 5;
    uint32_t iteration = 0;
    for (uint32_t cta_base = blockIdx.x * blockDim.x; cta_base < total; cta_base += stride, iteration ^= 1u) {
        const uint32_t base = cta_base + (threadIdx.x & ~31u);
        // bounds and keys sit at the meshlets' list indices: the opaque draws from 0, the alpha-clip draws from max_opaque
        const uint32_t k = base + lane < n_opaque ? base + lane : P.max_opaque + (base + lane - n_opaque);
        // the scan reads the 4-byte keys only; the 16-byte bound is fetched for the band's own meshlets
        const uint32_t key = base + lane < total ? __ldg(P.depth_keys + k) : kKeyEmpty;
        bool keep = key >= key_lo && key < key_hi;
        uint4 mb = make_uint4(kUntestable, 0u, 0u, kKeyEmpty);
        if (keep) mb = __ldg((const uint4*)(bounds + k));
        if (keep && band != 0 && mb.x != kUntestable) {
            const int px0 = (int)(mb.x & 0xFFFFu), px1 = (int)(mb.x >> 16), py0 = (int)(mb.y & 0xFFFFu), py1 = (int)(mb.y >> 16);
            uint32_t worst = 0;   // main view: largest stored key high word; cascade: largest stored depth
            if ((px1 - px0 + 1) * (py1 - py0 + 1) <= kSelectExactPixels) {
                // a rectangle of a few pixels (most meshlets of a far cascade): the stored values themselves, which is
                // what the raster kernel's own re-test would read -- with one LANE per meshlet instead of one warp
                const uint32_t* layer = SHADOW ? P.shadow + (size_t)cascade * P.shadowmap_size * P.shadowmap_size : nullptr;
                for (int y = py0; y <= py1; y++)
                    for (int x = px0; x <= px1; x++) {
                        const size_t at = (size_t)y * W + x;
                        worst = max(worst, SHADOW ? __ldcg(layer + at) : __ldcg((const uint32_t*)P.vis + 2 * at + 1));
                    }
            } else {
                const bool fine = ((px1 >> kHizShift0) - (px0 >> kHizShift0) + 1) * ((py1 >> kHizShift0) - (py0 >> kHizShift0) + 1) <= kMaxHizTiles;
                const int shift = fine ? kHizShift0 : kHizShift1, TW = fine ? TW0 : TW1;
                const uint32_t* level = fine ? level0 : level1;
                for (int ty = py0 >> shift; ty <= (py1 >> shift); ty++)
                    for (int tx = px0 >> shift; tx <= (px1 >> shift); tx++) worst = max(worst, __ldcg(level + (size_t)ty * TW + tx));
            }
            if (SHADOW ? mb.z >= worst : worst < mb.z) {
                keep = false;
                occluded++;
            }
        }
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, keep);
        if (kSelectCtaClaim) {
            if (lane == 0) warp_keep[iteration][warp] = (uint32_t)__popc(ballot);
            __syncthreads();
            if (threadIdx.x == 0) {
                uint32_t sum = 0;
#pragma unroll
                for (int w = 0; w < 8; w++) sum += warp_keep[iteration][w];
                cta_at[iteration] = sum ? atomicAdd(&P.pass->queue_count, sum) : 0u;
            }
            __syncthreads();
            if (keep) {
                uint32_t at = cta_at[iteration] + (uint32_t)__popc(ballot & ((1u << lane) - 1u));
                for (uint32_t w = 0; w < warp; w++) at += warp_keep[iteration][w];
                make_record(P, k, mb, (RasterRecord*)P.raster_records + at);
            }
        } else if (ballot) {
            uint32_t at = 0;
            if (lane == 0) at = atomicAdd(&P.pass->queue_count, (uint32_t)__popc(ballot));
            at = __shfl_sync(0xFFFFFFFFu, at, 0);
            if (keep) make_record(P, k, mb, (RasterRecord*)P.raster_records + at + __popc(ballot & ((1u << lane) - 1u)));
        }
    }
    if (P.flags & LH_CONFIG_PASS_TIMINGS) {
        occluded = __reduce_add_sync(0xFFFFFFFFu, occluded);
        if (lane == 0 && occluded) atomicAdd(&P.state->skipped_occluded[SHADOW ? 1u + cascade : 0u], occluded);
    }
}

// What the last CTA of a launch to finish does: it publishes how far the queue the launch appended to has grown
// (raster launch of band b: drain_end[b + 1]; first round of drain d: split_end[d + 1]).  Each slot is written once and
// read only by launches ordered after this one, whatever else runs beside them.
__device__ __forceinline__ void publish_queue_end(uint32_t* done_counter, const uint32_t* count, uint32_t* slot, uint32_t cap) {
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(done_counter, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        *done_counter = 0;
        *slot = min(*(volatile const uint32_t*)count, cap);
    }
}

#ifndef LH_RASTER_MIN_CTAS
#define LH_RASTER_MIN_CTAS 4
#endif

struct RasterCounters {
    uint32_t tris_done = 0, skipped_occluded = 0, skipped_empty = 0;
};

// One queued meshlet by one warp.  rec: its record (shared memory); T, micro_s, index_s: its instance transform,
// micro-indices and vertex indices staged in shared memory (LH_STAGE_STREAMS; micro_s / index_s point at the meshlet's
// first byte inside the staged 16-byte-aligned ranges), otherwise fetched through the record's pointers;
// first_keys: the lane's stored key under the first 8x4 block of the bound, prefetched with the chunk;
// M: the view's matrix (main view: the kernel parameter, i.e. constant-bank operands; cascade: shared memory).
template <bool SHADOW, bool ALPHA>
__device__ __forceinline__ void raster_one_meshlet(const FrameParams& P, const Target<SHADOW>& tg, WarpScratch& sv, const RasterRecord& rec,
                                                   const float* T, const uint8_t* micro_s, const uint8_t* index_s, uint32_t first_keys, const float* M,
                                                   bool affine, RasterCounters& cnt, uint32_t lane) {
    const int W = (int)tg.width, H = (int)tg.height;
    // ---- (0) re-test against the keys stored since the selection (the band's own meshlets occlude one another) ----
    {
        const uint4 mb = *(const uint4*)&rec.rect_x;   // rect_x, rect_y, edge, pad
        const int px0 = (int)(mb.x & 0xFFFFu), px1 = (int)(mb.x >> 16), py0 = (int)(mb.y & 0xFFFFu), py1 = (int)(mb.y >> 16);
        if (mb.x != kUntestable && (((px1 - px0) >> 3) + 1) * (((py1 - py0) >> 2) + 1) <= kMaxOcclusionBlocks) {
            uint32_t worst = first_keys;
            if (px1 - px0 >= 8 || py1 - py0 >= 4) {   // more than the prefetched block
                for (int y = py0 + (int)(lane >> 3); y <= py1; y += 4) {
                    const uint32_t* row = SHADOW ? tg.layer + (size_t)y * W : (const uint32_t*)tg.vis + 2 * (size_t)y * W + 1;
                    for (int x = px0 + (int)(lane & 7); x <= px1; x += 8) worst = max(worst, __ldcg(row + (SHADOW ? x : 2 * x)));
                }
            }
            worst = __reduce_max_sync(0xFFFFFFFFu, worst);
            if (SHADOW ? mb.z >= worst : worst < mb.z) {
                cnt.skipped_occluded++;
                return;
            }
        }
    }
    const uint32_t cf = rec.counts_flags;
    // more triangles / vertices than a meshlet may have (MAX_MESHLET_TRIANGLES, MAX_MESHLET_UNIQUE_VERTICES) are cut off
    const uint32_t tri_count = min(cf & 0xFFu, (uint32_t)LH_MAX_MESHLET_TRIANGLES);
    const uint32_t nv = min((cf >> 8) & 0xFFu, (uint32_t)LH_MAX_MESHLET_UNIQUE_VERTICES);
    const bool wide = (cf >> 16) & LH_MESH_INFO_FLAGS_32_BIT_INDICES;
    constexpr bool alpha = ALPHA;   // MESH_INFO_FLAGS_ALPHA_CLIP: the caller branches once per meshlet, so the opaque path carries none of its state
    const float half_w = pmul((float)tg.width, 0.5f), half_h = pmul((float)tg.height, 0.5f);

    // ---- (1) vertex phase: load_vertex_data + calculate_world_pos + matrix * (world, 1) ----
    int vminx = INT_MAX, vmaxx = INT_MIN, vminy = INT_MAX, vmaxy = INT_MIN;
    uint32_t vnear = SHADOW ? 0xFFFFFFFFu : 0u;   // bits of the nearest vertex depth (depths are >= 0, so bits order like values)
    bool special = false;                          // some vertex is behind the near plane, non-finite or of negative depth
    {
        const uint16_t* positions = rec.positions;
        // the instance transform: live for this phase only
        float Tl[16];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const float4 c = kStageStreams ? *(const float4*)(T + 4 * k) : __ldg((const float4*)rec.instance + k);
            Tl[4 * k] = c.x; Tl[4 * k + 1] = c.y; Tl[4 * k + 2] = c.z; Tl[4 * k + 3] = c.w;
        }
        const uint8_t* index_g = rec.indices;
        // A meshlet has at most 64 vertices: two per lane.  Both vertex indices, then all six position loads, are issued
        // before anything is used, so the phase pays the index and the position round trip once, not once per vertex.
        auto load_index = [&](uint32_t v) -> uint32_t {
            if (kStageStreams) return wide ? *(const uint32_t*)(index_s + 4 * v) : (uint32_t)*(const uint16_t*)(index_s + 2 * v);
            return wide ? __ldg((const uint32_t*)index_g + v) : (uint32_t)__ldg((const uint16_t*)index_g + v);
        };
        const uint32_t v0 = lane, v1 = lane + 32u;
        const bool has0 = v0 < nv, has1 = v1 < nv;
        const uint32_t idx0 = has0 ? load_index(v0) : 0u, idx1 = has1 ? load_index(v1) : 0u;
        const uint16_t* p0 = positions + 3 * (size_t)idx0;
        const uint16_t* p1 = positions + 3 * (size_t)idx1;
        uint16_t q[6] = {0, 0, 0, 0, 0, 0};
        if (has0) { q[0] = __ldg(p0); q[1] = __ldg(p0 + 1); q[2] = __ldg(p0 + 2); }
        if (has1) { q[3] = __ldg(p1); q[4] = __ldg(p1 + 1); q[5] = __ldg(p1 + 2); }
#pragma unroll
        for (int k = 0; k < 2; k++) {
            if (!(k ? has1 : has0)) continue;
            const uint32_t v = k ? v1 : v0;
            const float px = (float)q[3 * k], py = (float)q[3 * k + 1], pz = (float)q[3 * k + 2];
            const V4 world = mat4_mul(Tl, px, py, pz, 1.0f);
            V4 clip;
            if (SHADOW && affine) {
                // a cascade matrix ends in the row (0, 0, 0, 1): w is exactly 1 for every finite position, and a
                // non-finite one makes x, y and z non-finite too, so the vertex is classified the same without w
                clip.x = padd(padd(pmul(M[0], world.x), pmul(M[4], world.y)), padd(pmul(M[8], world.z), M[12]));
                clip.y = padd(padd(pmul(M[1], world.x), pmul(M[5], world.y)), padd(pmul(M[9], world.z), M[13]));
                clip.z = padd(padd(pmul(M[2], world.x), pmul(M[6], world.y)), padd(pmul(M[10], world.z), M[14]));
                clip.w = 1.0f;
            } else {
                clip = mat4_mul(M, world.x, world.y, world.z, 1.0f);
            }
            int sx = kVertOutside, sy = 0;
            float z = 0.0f;
            if (SHADOW || psub(clip.w, clip.z) >= 0.0f) {
                if (!project_snap(clip.x, clip.y, clip.z, clip.w, half_w, half_h, &sx, &sy, &z)) sx = kVertNonFinite;
            }
            sv.sx[v] = sx;
            sv.sy[v] = sy;
            sv.z[v] = z;
            if (alpha) sv.w[v] = clip.w;
            // first / last pixel per axis, saturated to int16 and packed (x low, y high) in one instruction each
            sv.cover[v] = make_uint2(pack_sat_s16((sy + 127) >> 8, (sx + 127) >> 8), pack_sat_s16((sy - 128) >> 8, (sx - 128) >> 8));
            if (sx == kVertOutside || sx == kVertNonFinite || !(z >= 0.0f)) {
                special = true;
            } else {
                vminx = min(vminx, sx); vmaxx = max(vmaxx, sx);
                vminy = min(vminy, sy); vmaxy = max(vmaxy, sy);
                vnear = SHADOW ? min(vnear, __float_as_uint(z)) : max(vnear, __float_as_uint(z));
            }
        }
    }
    __syncwarp();

    // ---- (1b) the same two early-outs with the exact vertex bounds: every triangle lies inside the vertices'
    // rectangle, and (screen-linear depth) no fragment is nearer than the nearest vertex ----
    const bool any_special = __any_sync(0xFFFFFFFFu, special);
    int bx0 = 0, bx1 = INT_MAX, by0 = 0, by1 = INT_MAX;   // pixels whose centres lie inside the vertices' rectangle
    if (!any_special) {
        vminx = __reduce_min_sync(0xFFFFFFFFu, vminx); vmaxx = __reduce_max_sync(0xFFFFFFFFu, vmaxx);
        vminy = __reduce_min_sync(0xFFFFFFFFu, vminy); vmaxy = __reduce_max_sync(0xFFFFFFFFu, vmaxy);
        bx0 = max(0, (vminx - 128 + 255) >> 8); bx1 = min(W - 1, (vmaxx - 128) >> 8);
        by0 = max(0, (vminy - 128 + 255) >> 8); by1 = min(H - 1, (vmaxy - 128) >> 8);
        if (bx0 > bx1 || by0 > by1) {   // no pixel centre inside
            cnt.skipped_empty++;
            return;
        }
        if (bx1 - bx0 < 8 && by1 - by0 < 4) {   // one 8x4 block: test the stored keys against the nearest vertex depth
            vnear = SHADOW ? __reduce_min_sync(0xFFFFFFFFu, vnear) : __reduce_max_sync(0xFFFFFFFFu, vnear);
            const float zn = __uint_as_float(vnear);
            // the interpolated depth can leave the vertices' range by rounding only: pad by a few ulp
            const float d = SHADOW ? zn - (4e-6f * zn + 1e-30f) : zn + (4e-6f * zn + 1e-30f);
            const int x = bx0 + (int)(lane & 7), y = by0 + (int)(lane >> 3);
            uint32_t worst = 0;
            if (x <= bx1 && y <= by1) {
                const size_t at = (size_t)y * W + x;
                worst = SHADOW ? __ldcg(tg.layer + at) : __ldcg((const uint32_t*)tg.vis + 2 * at + 1);
            }
            worst = __reduce_max_sync(0xFFFFFFFFu, worst);
            if (d > 0.0f && (SHADOW ? __float_as_uint(d) >= worst : worst < ~__float_as_uint(d))) {
                cnt.skipped_occluded++;
                return;
            }
        }
    }

    // cull modes: pipelines.cpp:10-25,255-315
    const int cull = alpha ? CULL_NONE : (SHADOW ? CULL_FRONT : CULL_BACK);
    const uint32_t lo_base = (P.slot_offset + rec.local) << LH_VIS_TRI_BITS;
    const uint8_t* micro_g = rec.micro;
    auto micro = [&](uint32_t k) -> uint32_t { return kStageStreams ? (uint32_t)micro_s[k] : (uint32_t)__ldg(micro_g + k); };

    // ---- (2') a meshlet that can touch at most 2x2 samples -- most meshlets of a far cascade or of distant instances:
    // every triangle tests those samples directly (same integer edge functions, same top-left rule, same depth
    // expression as below), with no per-triangle rectangle, no compaction and no second pass over the survivors ----
    if (kFewSamplesPath && !ALPHA && !any_special && bx1 - bx0 <= 1 && by1 - by0 <= 1) {
        const int ccx = bx0 * 256 + 128, ccy = by0 * 256 + 128;
        const int nx = bx1 - bx0, ny = by1 - by0;   // 0 or 1 more sample per axis
        for (uint32_t t = lane; t < tri_count; t += 32) {
            const uint32_t ia = micro(3 * t);
            uint32_t ib = micro(3 * t + 1), ic = micro(3 * t + 2);
            if (max(ia, max(ib, ic)) >= nv) continue;
            const int ax = sv.sx[ia], ay = sv.sy[ia];
            int bx = sv.sx[ib], by = sv.sy[ib], cx = sv.sx[ic], cy = sv.sy[ic];
            // the vertices span fewer than three pixels: every product below is far inside 32 bits
            int A2 = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
            const bool front = A2 < 0;
            if (A2 == 0 || (cull == CULL_BACK && !front) || (cull == CULL_FRONT && front)) continue;
            if (front) {
                int s0 = bx; bx = cx; cx = s0;
                s0 = by; by = cy; cy = s0;
                const uint32_t si = ib; ib = ic; ic = si;
                A2 = -A2;
            }
            const int e0x = cx - bx, e0y = cy - by, e1x = ax - cx, e1y = ay - cy, e2x = bx - ax, e2y = by - ay;
            const int bias0 = top_left(e0x, e0y) ? 0 : -1, bias1 = top_left(e1x, e1y) ? 0 : -1, bias2 = top_left(e2x, e2y) ? 0 : -1;
            int w0_row = e0x * (ccy - by) - e0y * (ccx - bx);
            int w1_row = e1x * (ccy - cy) - e1y * (ccx - cx);
            int w2_row = e2x * (ccy - ay) - e2y * (ccx - ax);
            bool any = false;
            float az = 0.0f, dz1 = 0.0f, dz2 = 0.0f, inv_area = 0.0f;
            for (int dy = 0; dy <= ny; dy++) {
                int w0 = w0_row, w1 = w1_row, w2 = w2_row;
                for (int dx = 0; dx <= nx; dx++) {
                    if (((w0 + bias0) | (w1 + bias1) | (w2 + bias2)) >= 0) {
                        if (!any) {
                            any = true;
                            az = sv.z[ia];
                            dz1 = psub(sv.z[ib], az);
                            dz2 = psub(sv.z[ic], az);
                            inv_area = pdiv(1.0f, (float)A2);
                            cnt.tris_done++;
                        }
                        const float z = padd(az, pmul(padd(pmul((float)w1, dz1), pmul((float)w2, dz2)), inv_area));
                        tg.emit(bx0 + dx, by0 + dy, z, lo_base | t);
                    }
                    w0 -= e0y * 256;
                    w1 -= e1y * 256;
                    w2 -= e2y * 256;
                }
                w0_row += e0x * 256;
                w1_row += e1x * 256;
                w2_row += e2x * 256;
            }
        }
        return;
    }

    // ---- (2) triangle cull phase ----
    uint32_t n_surv = 0;
    const uint32_t screen_max = (uint32_t)(W - 1) | ((uint32_t)(H - 1) << 16);
    for (uint32_t t0 = 0; t0 < tri_count; t0 += 32) {
        const uint32_t t = t0 + lane;
        bool keep = false;
        if (t < tri_count) {
            const uint32_t i0 = micro(3 * t), i1 = micro(3 * t + 1), i2 = micro(3 * t + 2);
            if (max(i0, max(i1, i2)) < nv) {
                const uint32_t lo = lo_base | t;
                // facing test and routing of a triangle whose pixel rectangle [px0,px1] x [py0,py1] is not empty
                auto covered = [&](int px0, int px1, int py0, int py1) {
                    const int ax = sv.sx[i0], bx = sv.sx[i1], cx = sv.sx[i2], ay = sv.sy[i0], by = sv.sy[i1], cy = sv.sy[i2];
                    const int minx = min(ax, min(bx, cx)), maxx = max(ax, max(bx, cx));
                    const int miny = min(ay, min(by, cy)), maxy = max(ay, max(by, cy));
                    cnt.tris_done++;
                    // exact in 32 bits when the sub-pixel extent is < 2^14 (every product < 2^28)
                    const bool small_extent = (maxx - minx) < kSmallMaxExtent && (maxy - miny) < kSmallMaxExtent;
                    const long long A2 = small_extent
                                             ? (long long)((bx - ax) * (cy - ay) - (by - ay) * (cx - ax))
                                             : (long long)(bx - ax) * (cy - ay) - (long long)(by - ay) * (cx - ax);
                    const bool front = A2 < 0;
                    if (A2 == 0 || (cull == CULL_BACK && !front) || (cull == CULL_FRONT && front)) return;
                    const int spanx = px1 - px0 + 1, spany = py1 - py0 + 1;
                    if (small_extent && spanx <= kSmallMaxSpan && spany <= kSmallMaxSpan && spanx * spany <= (int)P.small_max_area) {
                        keep = true;
                    } else {
                        // large footprint: one queue entry per 128x32 screen tile the bounding box touches
                        const uint32_t tx0 = (uint32_t)px0 / kTileW, tx1 = (uint32_t)px1 / kTileW;
                        const uint32_t ty0 = (uint32_t)py0 / kTileH, ty1 = (uint32_t)py1 / kTileH;
                        if ((tx1 - tx0 + 1) * (ty1 - ty0 + 1) > 2) {
                            push_deferred(P, lo, kWholeTriangle);   // split by the deferred kernel, a warp per triangle
                        } else {
                            for (uint32_t ty = ty0; ty <= ty1; ty++)
                                for (uint32_t tx = tx0; tx <= tx1; tx++) push_deferred(P, lo, tx | (ty << 16));
                        }
                    }
                };
                if (!any_special) {
                    // the common case: every vertex is in front of the near plane and finite.  Rectangle by
                    // halfword min/max of the per-vertex pixel bounds; most triangles end here (no sample inside).
                    const uint2 c0 = sv.cover[i0], c1 = sv.cover[i1], c2 = sv.cover[i2];
                    const uint32_t first = __vmaxs2(__vmins2(__vmins2(c0.x, c1.x), c2.x), 0u);
                    const uint32_t last = __vmins2(__vmaxs2(__vmaxs2(c0.y, c1.y), c2.y), screen_max);
                    if (__vcmpgts2(first, last) == 0u)
                        covered((int)(first & 0xFFFFu), (int)(last & 0xFFFFu), (int)(first >> 16), (int)(last >> 16));
                } else {
                    const int ax = sv.sx[i0], bx = sv.sx[i1], cx = sv.sx[i2];
                    const int n_out = (ax == kVertOutside) + (bx == kVertOutside) + (cx == kVertOutside);
                    if (n_out == 0) {
                        if (ax != kVertNonFinite && bx != kVertNonFinite && cx != kVertNonFinite) {
                            const int ay = sv.sy[i0], by = sv.sy[i1], cy = sv.sy[i2];
                            const int minx = min(ax, min(bx, cx)), maxx = max(ax, max(bx, cx));
                            const int miny = min(ay, min(by, cy)), maxy = max(ay, max(by, cy));
                            const int px0 = max(0, (minx - 128 + 255) >> 8), px1 = min(W - 1, (maxx - 128) >> 8);
                            const int py0 = max(0, (miny - 128 + 255) >> 8), py1 = min(H - 1, (maxy - 128) >> 8);
                            if (px0 <= px1 && py0 <= py1) covered(px0, px1, py0, py1);
                        }
                    } else if (n_out < 3) {
                        push_deferred(P, lo, kWholeTriangle);   // crosses the near plane: generic path
                    }
                }
            }
        }
        const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, keep);
        if (keep) sv.surv[n_surv + __popc(ballot & ((1u << lane) - 1u))] = (uint8_t)t;
        n_surv += __popc(ballot);
    }
    __syncwarp();
    if (n_surv == 0) return;

    // ---- (3) raster phase over the compacted survivors ----
    AlphaTest at{false, 0};
    const uint16_t* uvs = nullptr;
    float tex_scale[2] = {0.0f, 0.0f}, tex_offset[2] = {0.0f, 0.0f};
    if (ALPHA) {
        const LhMeshInfo* mi = rec.mi;
        at = make_alpha_test(P, true, __ldg(&mi->base_color_texture_index));
        uvs = (const uint16_t*)__ldg((const unsigned long long*)&mi->uvs);
        tex_scale[0] = __ldg(&mi->texture_scale[0]);
        tex_scale[1] = __ldg(&mi->texture_scale[1]);
        tex_offset[0] = __ldg(&mi->texture_offset[0]);
        tex_offset[1] = __ldg(&mi->texture_offset[1]);
    }
    for (uint32_t s = lane; s < n_surv; s += 32) {
        const uint32_t t = sv.surv[s];
        const uint32_t lo = lo_base | t;
        uint32_t ia = micro(3 * t), ib = micro(3 * t + 1), ic = micro(3 * t + 2);
        int ax = sv.sx[ia], ay = sv.sy[ia], bx = sv.sx[ib], by = sv.sy[ib], cx = sv.sx[ic], cy = sv.sy[ic];
        const int minx = min(ax, min(bx, cx)), maxx = max(ax, max(bx, cx));
        const int miny = min(ay, min(by, cy)), maxy = max(ay, max(by, cy));
        const int px0 = max(0, (minx - 128 + 255) >> 8), px1 = min(W - 1, (maxx - 128) >> 8);
        const int py0 = max(0, (miny - 128 + 255) >> 8), py1 = min(H - 1, (maxy - 128) >> 8);
        int A2 = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
        float az = sv.z[ia], bz = sv.z[ib], cz = sv.z[ic];
        if (A2 < 0) {
            int t0 = bx; bx = cx; cx = t0;
            t0 = by; by = cy; cy = t0;
            float tz = bz; bz = cz; cz = tz;
            const uint32_t ti = ib; ib = ic; ic = ti;
            A2 = -A2;
        }
        const int e0x = cx - bx, e0y = cy - by, e1x = ax - cx, e1y = ay - cy, e2x = bx - ax, e2y = by - ay;
        const int bias0 = top_left(e0x, e0y) ? 0 : -1, bias1 = top_left(e1x, e1y) ? 0 : -1, bias2 = top_left(e2x, e2y) ? 0 : -1;
        const float fA = (float)A2;
        const float inv_area = pdiv(1.0f, fA);
        const float dz1 = psub(bz, az), dz2 = psub(cz, az);
        const int ccx = px0 * 256 + 128, ccy = py0 * 256 + 128;
        int w0_row = e0x * (ccy - by) - e0y * (ccx - bx);
        int w1_row = e1x * (ccy - cy) - e1y * (ccx - cx);
        int w2_row = e2x * (ccy - ay) - e2y * (ccx - ax);
        SnapVert A, B, C;
        if (ALPHA && at.enabled) {
            // perspective-correct uv needs the clip w and the vertex uvs
            const uint32_t ids[3] = {ia, ib, ic};
            SnapVert* out[3] = {&A, &B, &C};
            for (int q = 0; q < 3; q++) {
                const uint8_t* index_g = rec.indices;
                const uint32_t idx = wide ? __ldg((const uint32_t*)index_g + ids[q]) : (uint32_t)__ldg((const uint16_t*)index_g + ids[q]);
                out[q]->w = sv.w[ids[q]];
                out[q]->u = padd(pmul((float)__ldg(uvs + 2 * (size_t)idx), tex_scale[0]), tex_offset[0]);
                out[q]->v = padd(pmul((float)__ldg(uvs + 2 * (size_t)idx + 1), tex_scale[1]), tex_offset[1]);
            }
        }
        for (int py = py0; py <= py1; py++) {
            int w0 = w0_row, w1 = w1_row, w2 = w2_row;
            for (int px = px0; px <= px1; px++) {
                if (((w0 + bias0) | (w1 + bias1) | (w2 + bias2)) >= 0) {
                    if (!(ALPHA && at.enabled && alpha_discard(at, (float)w0, (float)w1, (float)w2, fA, A, B, C))) {
                        float z = padd(az, pmul(padd(pmul((float)w1, dz1), pmul((float)w2, dz2)), inv_area));
                        tg.emit(px, py, z, lo);
                    }
                }
                w0 -= e0y * 256;
                w1 -= e1y * 256;
                w2 -= e2y * 256;
            }
            w0_row += e0x * 256;
            w1_row += e1x * 256;
            w2_row += e2x * 256;
        }
    }
}

// One warp per queued meshlet.  A warp takes up to kChunk queue entries per ticket (one when the band is small, so that
// every warp gets work and a heavy near meshlet does not serialise its neighbours).  The front end costs three round
// trips per CHUNK, none of them a pointer chase: the ticket (taken one chunk ahead, so it is already there), the
// chunk's records (what the selection's lanes left: pointers, counts, bound), then -- issued together -- the stored
// keys under each bound and the asynchronous copies of the chunk's micro-indices, vertex indices and transforms.
template <bool SHADOW>
__global__ void __launch_bounds__(kWarpsPerCta * 32, LH_RASTER_MIN_CTAS) raster_meshlets_kernel(const __grid_constant__ FrameParams P, uint32_t cascade, uint32_t band) {
    __shared__ WarpScratch scratch_all[kWarpsPerCta];
    __shared__ float cascade_matrix[16];
#if LH_LANE_PINNED
    // read once through an opaque instruction: under the 64-register cap the compiler otherwise re-reads the special
    // register (S2R) and rebuilds the warp's shared-memory base all over the chunk loop
    uint32_t lane, warp;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    asm volatile("shr.u32 %0, %1, 5;" : "=r"(warp) : "r"(threadIdx.x));
#else
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#endif
    WarpScratch& sv = scratch_all[warp];
    // band b owns the queue entries appended by its selection: [band_mark[b], queue_count); nothing appends while
    // this kernel runs, and the mark it leaves for band b + 1 is read by that band's launch only
    const uint32_t band_first = P.pass->band_mark[band], band_end = P.pass->queue_count;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        P.pass->band_mark[band + 1] = band_end;
        if (P.flags & LH_CONFIG_PASS_TIMINGS) atomicAdd(&P.state->queued[SHADOW ? 1u + cascade : 0u], band_end - band_first);
    }
    if (kRecordBulk && lane == 0) {
        mbar_init(&sv.rec_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // The main view multiplies by the kernel parameter itself (constant-bank operands, no registers); a cascade's
    // matrix was generated on the device: it is read from shared memory where it is used.
    if (SHADOW && threadIdx.x < 16) cascade_matrix[threadIdx.x] = (&((const LhMiscStorage*)P.u.misc_storage)->shadow_matrices[cascade].m[0][0])[threadIdx.x];
    __syncthreads();

    Target<SHADOW> tg;
    tg.vis = P.vis;
    tg.layer = SHADOW ? P.shadow + (size_t)cascade * P.shadowmap_size * P.shadowmap_size : nullptr;
    tg.width = SHADOW ? P.shadowmap_size : P.u.window_size[0];
    tg.height = SHADOW ? P.shadowmap_size : P.u.window_size[1];
    tg.count = (P.flags & LH_CONFIG_COUNT_SAMPLES) != 0;
    tg.samples = 0;
    const int W = (int)tg.width;
    const float* M = SHADOW ? cascade_matrix : &P.u.combined_perspective_view.m[0][0];
    const bool affine = SHADOW && cascade_matrix[3] == 0.0f && cascade_matrix[7] == 0.0f && cascade_matrix[11] == 0.0f && cascade_matrix[15] == 1.0f;
    const RasterRecord* records = (const RasterRecord*)P.raster_records;
    RasterCounters cnt;

    const uint32_t band_count = band_end - band_first;
    const uint32_t chunk = max(1u, min((uint32_t)kChunk, band_count / (gridDim.x * kWarpsPerCta * 8u)));
    // The CTAs of the first wave -- resident together when the launch has the GPU to itself -- take their FIRST chunk
    // statically (their warp index); the rest of the band is handed out by tickets, and a CTA of a later wave looks at
    // the counter once (one load per CTA) before any of its warps asks it for work.  A short band -- a rank's share on
    // eight GPUs, the near cascades -- then costs no atomic at all instead of one per warp of the grid on a single
    // address, and a long one saves a round trip per first-wave warp.
    // lane 0 holds the NEXT ticket (relative to band_first) from the moment it is requested; nothing reads it before the
    // following iteration, so the atomic's round trip overlaps the current chunk.
    const uint32_t first_wave = kStaticFirst ? min(gridDim.x, P.first_wave_ctas) : 0u;
    const uint32_t static_end = min(band_count, first_wave * kWarpsPerCta * chunk);   // entries [0, static_end) are assigned statically
    uint32_t next_ticket;
    if (blockIdx.x < first_wave) {
        next_ticket = (blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5)) * chunk;
        if (next_ticket >= static_end) next_ticket = band_count;   // nothing static left for this warp: the band is short
    } else {
        __shared__ uint32_t taken;
        if (threadIdx.x == 0) taken = __ldcg(&P.pass->band_ticket[band]);
        __syncthreads();
        next_ticket = band_count;
        if (static_end + taken < band_count && lane == 0) next_ticket = static_end + atomicAdd(&P.pass->band_ticket[band], chunk);
    }
    for (uint32_t iteration = 0;; iteration++) {
        const uint32_t ticket = band_first + __shfl_sync(0xFFFFFFFFu, next_ticket, 0);
        if (ticket >= band_end) break;
        if (lane == 0) next_ticket = static_end < band_count ? static_end + atomicAdd(&P.pass->band_ticket[band], chunk) : band_count;
        const uint32_t n = min(chunk, band_end - ticket);
        // ---- the chunk's records ----
        uint32_t my_rx = kUntestable, my_ry = 0;
        if (kRecordBulk) {
            if (lane == 0) {
                mbar_arrive_expect_tx(&sv.rec_bar, n * (uint32_t)sizeof(RasterRecord));
                bulk_copy_g2s(&sv.rec[0], records + ticket, n * (uint32_t)sizeof(RasterRecord), &sv.rec_bar);
            }
            mbar_wait(&sv.rec_bar, iteration & 1u);
            if (lane < n) {
                my_rx = sv.rec[lane].rect_x;
                my_ry = sv.rec[lane].rect_y;
            }
        } else {
            if (lane < n) {
                const uint4* src = (const uint4*)(records + ticket + lane);
                const uint4 a = __ldg(src), b = __ldg(src + 1), c = __ldg(src + 2), d = __ldg(src + 3);
                uint4* dst = (uint4*)&sv.rec[lane];
                dst[0] = a; dst[1] = b; dst[2] = c; dst[3] = d;
                my_rx = d.x;
                my_ry = d.y;
            }
            __syncwarp();
        }
        // ---- the stored keys under the first 8x4 block of every bound, and the chunk's streams, all in flight together ----
        uint32_t first_keys[kChunk];
#pragma unroll
        for (int j = 0; j < kChunk; j++) {
            first_keys[j] = 0;
            const uint32_t rx = __shfl_sync(0xFFFFFFFFu, my_rx, j), ry = __shfl_sync(0xFFFFFFFFu, my_ry, j);
            const int x = (int)(rx & 0xFFFFu) + (int)(lane & 7), y = (int)(ry & 0xFFFFu) + (int)(lane >> 3);
            if (rx != kUntestable && x <= (int)(rx >> 16) && y <= (int)(ry >> 16)) {
                const size_t at = (size_t)y * W + x;
                first_keys[j] = SHADOW ? __ldcg(tg.layer + at) : __ldcg((const uint32_t*)tg.vis + 2 * at + 1);
            }
        }
        if (kStageStreams) {
#pragma unroll
            for (int j = 0; j < kChunk; j++) {
                if ((uint32_t)j >= n) break;
                const RasterRecord& r = sv.rec[j];
                const uint32_t cf = r.counts_flags;
                const uint32_t tris = min(cf & 0xFFu, (uint32_t)LH_MAX_MESHLET_TRIANGLES), verts = min((cf >> 8) & 0xFFu, (uint32_t)LH_MAX_MESHLET_UNIQUE_VERTICES);
                const uintptr_t m = (uintptr_t)r.micro, x = (uintptr_t)r.indices;
                const uint32_t m_bytes = (uint32_t)(m & 15u) + 3u * tris;                                                    // <= 387: pieces 0..24
                const uint32_t x_bytes = (uint32_t)(x & 15u) + (((cf >> 16) & LH_MESH_INFO_FLAGS_32_BIT_INDICES) ? 4u : 2u) * verts;   // <= 271: pieces 0..16
                // lanes 0..24: the micro-indices; lanes 28..31: the instance transform; then lanes 0..16: the vertex indices
                if (lane * 16u < m_bytes) async_copy16(&sv.micro[kStageStreams ? j : 0][lane * 16u], (const uint8_t*)(m & ~(uintptr_t)15) + lane * 16u);
                else if (lane >= 28u) async_copy16(&sv.T[kStageStreams ? j : 0][(lane - 28u) * 4u], (const uint8_t*)r.instance + (lane - 28u) * 16u);
                if (lane * 16u < x_bytes) async_copy16(&sv.index[kStageStreams ? j : 0][lane * 16u], (const uint8_t*)(x & ~(uintptr_t)15) + lane * 16u);
            }
            async_copy_wait_all();
            __syncwarp();
        }
#pragma unroll
        for (int j = 0; j < kChunk; j++) {
            if ((uint32_t)j >= n) break;
            const RasterRecord& rec = sv.rec[j];
            const int js = kStageStreams ? j : 0;
            const uint8_t* micro = &sv.micro[js][0] + ((uintptr_t)rec.micro & 15u);
            const uint8_t* index = &sv.index[js][0] + ((uintptr_t)rec.indices & 15u);
            if ((rec.counts_flags >> 16) & LH_MESH_INFO_FLAGS_ALPHA_CLIP)
                raster_one_meshlet<SHADOW, true>(P, tg, sv, rec, &sv.T[js][0], micro, index, first_keys[j], M, affine, cnt, lane);
            else
                raster_one_meshlet<SHADOW, false>(P, tg, sv, rec, &sv.T[js][0], micro, index, first_keys[j], M, affine, cnt, lane);
            __syncwarp();
        }
    }
    if (P.flags & LH_CONFIG_PASS_TIMINGS) {
        const uint32_t tris_done = __reduce_add_sync(0xFFFFFFFFu, cnt.tris_done);
        if (lane == 0) {
            const uint32_t pass = SHADOW ? 1u + cascade : 0u;
            if (tris_done && !SHADOW) atomicAdd(&P.state->rasterized_triangles, (unsigned long long)tris_done);
            if (cnt.skipped_occluded) atomicAdd(&P.state->skipped_occluded[pass], cnt.skipped_occluded);
            if (cnt.skipped_empty) atomicAdd(&P.state->skipped_empty[pass], cnt.skipped_empty);
        }
    }
    if (tg.count) {
        const uint32_t n = __reduce_add_sync(0xFFFFFFFFu, tg.samples);
        if (lane == 0 && n) atomicAdd(&P.state->samples[SHADOW ? 1u + cascade : 0u], (unsigned long long)n);
    }
    publish_queue_end(&P.pass->cta_done, &P.pass->large_count, &P.pass->drain_end[band + 1], P.max_large);
}

// ---- the deferred queue: large and near-plane-crossing triangles, one warp per entry ----
// An entry is (key low word, screen tile) or (key low word, kWholeTriangle).  Whole-triangle entries are clipped and,
// when their rectangle spans more than two 128x32 tiles, split into per-tile entries appended to the queue -- a
// screen-filling triangle would otherwise be scanned by a single warp -- which the next launch rasterises.
// A warp takes 32 entries at a time: (a) one lane per entry walks the pointer chain, fetches and transforms the three
// vertices, clips and snaps, and parks the result in shared memory -- the chain's seven dependent round trips are paid
// once per 32 entries; (b) the warp then goes through the parked triangles one by one, all lanes scanning 8x4 blocks.
struct DeferredTriangle {
    SnapVert s[4];
    cudaTextureObject_t tex;     // 0: no alpha test
    uint32_t lo;
    uint32_t rect_x, rect_y;     // x0 | x1 << 16, y0 | y1 << 16 (pixels); split: tile range instead
    uint16_t n;                  // vertices after clipping; 0: nothing to do
    uint8_t cull, split;         // split: append one entry per tile of the rectangle instead of rasterising; bit 1: 32-bit edge functions suffice
};
constexpr int kDeferredWarps = 4;

template <bool SHADOW>
__global__ void __launch_bounds__(kDeferredWarps * 32) raster_deferred_kernel(const __grid_constant__ FrameParams P, uint32_t cascade, uint32_t round, uint32_t from, uint32_t to, uint32_t drain) {
    __shared__ DeferredTriangle parked_all[kDeferredWarps][32];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    DeferredTriangle* parked = parked_all[warp];
    // round 0: the deferred queue's entries appended by the bands from .. to - 1 (drain_end[from] .. drain_end[to]);
    // round 1: the split queue's entries appended by round 0 of the same drain
    const uint32_t range_begin = round == 0 ? P.pass->drain_end[from] : P.pass->split_end[drain];
    const uint32_t range_end = round == 0 ? P.pass->drain_end[to] : P.pass->split_end[drain + 1];
    const uint32_t* queue = round == 0 ? P.large_queue : P.large_queue + 2 * (size_t)P.max_large;
    uint32_t* split_queue = P.large_queue + 2 * (size_t)P.max_large;
    Target<SHADOW> tg;
    tg.vis = P.vis;
    tg.layer = SHADOW ? P.shadow + (size_t)cascade * P.shadowmap_size * P.shadowmap_size : nullptr;
    tg.width = SHADOW ? P.shadowmap_size : P.u.window_size[0];
    tg.height = SHADOW ? P.shadowmap_size : P.u.window_size[1];
    tg.count = (P.flags & LH_CONFIG_COUNT_SAMPLES) != 0;
    tg.samples = 0;
    const float* M = SHADOW ? &((const LhMiscStorage*)P.u.misc_storage)->shadow_matrices[cascade].m[0][0] : &P.u.combined_perspective_view.m[0][0];

    for (uint32_t base = range_begin + (blockIdx.x * kDeferredWarps + warp) * 32; base < range_end; base += gridDim.x * kDeferredWarps * 32) {
        // ---- (a) one lane per entry ----
        {
            DeferredTriangle& d = parked[lane];
            d.n = 0;
            const uint32_t e = base + lane;
            if (e < range_end) {
                const uint32_t lo = queue[2 * (size_t)e], tile = queue[2 * (size_t)e + 1];
                const uint32_t local = (lo >> LH_VIS_TRI_BITS) - P.slot_offset, t = lo & LH_VIS_TRI_MASK;
                MeshletHeader h;
                load_header(P, local, h);
                const AlphaTest at = make_alpha_test(P, h.alpha, h.base_color_texture);
                ClipVert v[3];
                for (int q = 0; q < 3; q++) {
                    const uint32_t mi = __ldg(h.micro + 3 * t + q);
                    const uint32_t idx = load_vertex_index(h, mi);
                    const uint16_t* p = h.positions + 3 * (size_t)idx;
                    const V4 world = mat4_mul(h.T, (float)__ldg(p), (float)__ldg(p + 1), (float)__ldg(p + 2), 1.0f);
                    const V4 clip = mat4_mul(M, world.x, world.y, world.z, 1.0f);
                    v[q] = ClipVert{clip.x, clip.y, clip.z, clip.w, 0.0f, 0.0f};
                    if (at.enabled) {
                        v[q].u = padd(pmul((float)__ldg(h.uvs + 2 * (size_t)idx), h.tex_scale[0]), h.tex_offset[0]);
                        v[q].v = padd(pmul((float)__ldg(h.uvs + 2 * (size_t)idx + 1), h.tex_scale[1]), h.tex_offset[1]);
                    }
                }
                int n = clip_and_snap<SHADOW>(v, tg, d.s);
                int rx0 = 0, ry0 = 0, rx1 = (int)tg.width - 1, ry1 = (int)tg.height - 1;
                bool split = false;
                if (n != 0) {
                    if (tile != kWholeTriangle) {
                        rx0 = (int)((tile & 0xFFFFu) * kTileW);
                        ry0 = (int)((tile >> 16) * kTileH);
                        rx1 = min(rx1, rx0 + (int)kTileW - 1);
                        ry1 = min(ry1, ry0 + (int)kTileH - 1);
                    } else {
                        int minx = d.s[0].x, maxx = d.s[0].x, miny = d.s[0].y, maxy = d.s[0].y;
                        for (int k = 1; k < n; k++) {
                            minx = min(minx, d.s[k].x); maxx = max(maxx, d.s[k].x);
                            miny = min(miny, d.s[k].y); maxy = max(maxy, d.s[k].y);
                        }
                        const int px0 = max(0, (minx - 128 + 255) >> 8), px1 = min(rx1, (maxx - 128) >> 8);
                        const int py0 = max(0, (miny - 128 + 255) >> 8), py1 = min(ry1, (maxy - 128) >> 8);
                        if (px0 > px1 || py0 > py1) {
                            n = 0;
                        } else {
                            const int tx0 = px0 / (int)kTileW, tx1 = px1 / (int)kTileW, ty0 = py0 / (int)kTileH, ty1 = py1 / (int)kTileH;
                            if ((tx1 - tx0 + 1) * (ty1 - ty0 + 1) > 2) {
                                split = true;
                                rx0 = tx0; rx1 = tx1; ry0 = ty0; ry1 = ty1;
                            }
                        }
                    }
                }
                d.n = (uint16_t)n;
                d.lo = lo;
                d.rect_x = (uint32_t)rx0 | ((uint32_t)rx1 << 16);
                d.rect_y = (uint32_t)ry0 | ((uint32_t)ry1 << 16);
                d.cull = (uint8_t)(h.alpha ? CULL_NONE : (SHADOW ? CULL_FRONT : CULL_BACK));
                bool narrow = n != 0;   // sub-pixel extent of the (clipped) polygon below 2^14 on both axes
                if (n != 0) {
                    int minx = d.s[0].x, maxx = d.s[0].x, miny = d.s[0].y, maxy = d.s[0].y;
                    for (int k = 1; k < n; k++) {
                        minx = min(minx, d.s[k].x); maxx = max(maxx, d.s[k].x);
                        miny = min(miny, d.s[k].y); maxy = max(maxy, d.s[k].y);
                    }
                    narrow = (long long)maxx - minx < kSmallMaxExtent && (long long)maxy - miny < kSmallMaxExtent;
                }
                d.split = (split ? 1 : 0) | (narrow ? 2 : 0);
                d.tex = at.enabled ? at.tex : 0;
            }
        }
        __syncwarp();
        // ---- (b) the warp, one parked triangle at a time ----
        for (int i = 0; i < 32; i++) {
            const DeferredTriangle& d = parked[i];
            const int n = d.n;
            if (n == 0) continue;
            const int x0 = (int)(d.rect_x & 0xFFFFu), x1 = (int)(d.rect_x >> 16), y0 = (int)(d.rect_y & 0xFFFFu), y1 = (int)(d.rect_y >> 16);
            const uint32_t lo = d.lo;
            if (d.split & 1) {
                const uint32_t ntx = (uint32_t)(x1 - x0 + 1), ntiles = ntx * (uint32_t)(y1 - y0 + 1);
                uint32_t at = 0;
                if (lane == 0) at = atomicAdd(&P.pass->split_count, ntiles);
                at = __shfl_sync(0xFFFFFFFFu, at, 0);
                for (uint32_t k = lane; k < ntiles; k += 32) {
                    if (at + k < P.max_large) {
                        split_queue[2 * (size_t)(at + k)] = lo;
                        split_queue[2 * (size_t)(at + k) + 1] = ((uint32_t)x0 + k % ntx) | (((uint32_t)y0 + k / ntx) << 16);
                    } else {
                        atomicOr(&P.state->overflow, (uint32_t)LH_OVERFLOW_LARGE_TRIANGLES);
                    }
                }
                continue;
            }
            const AlphaTest at{d.tex != 0, d.tex};
            if (d.split & 2) {
                raster_snapped_coop<SHADOW, int>(d.s[0], d.s[1], d.s[2], (int)d.cull, tg, x0, y0, x1, y1, at, lo);
                if (n == 4) raster_snapped_coop<SHADOW, int>(d.s[0], d.s[2], d.s[3], (int)d.cull, tg, x0, y0, x1, y1, at, lo);
            } else {
                raster_snapped_coop<SHADOW, long long>(d.s[0], d.s[1], d.s[2], (int)d.cull, tg, x0, y0, x1, y1, at, lo);
                if (n == 4) raster_snapped_coop<SHADOW, long long>(d.s[0], d.s[2], d.s[3], (int)d.cull, tg, x0, y0, x1, y1, at, lo);
            }
        }
        __syncwarp();
    }
    if (tg.count) {
        const uint32_t n = __reduce_add_sync(0xFFFFFFFFu, tg.samples);
        if (lane == 0 && n) atomicAdd(&P.state->samples[SHADOW ? 1u + cascade : 0u], (unsigned long long)n);
    }
    if (round == 0) publish_queue_end(&P.pass->cta_done_deferred, &P.pass->split_count, &P.pass->split_end[drain + 1], P.max_large);
}

__global__ void end_main_pass_kernel(const __grid_constant__ FrameParams P) {
    if (threadIdx.x == 0 && blockIdx.x == 0) P.state->main_large_count = P.pass->large_count + P.pass->split_count;
}

}  // namespace

void launch_raster(LhContext* ctx, const FrameParams& params, int cascade, cudaStream_t stream) {
    if (!stream) stream = ctx->stream;
    // The main view runs alone: many small CTAs even out its heavy near meshlets.  The cascades run side by side on
    // four streams: grids of about one wave each let their kernels share the SMs instead of queueing behind one another.
    // Every band costs a hierarchical-Z rebuild, a selection and a raster launch in a dependent chain.  A rank's share of
    // the lists shrinks with the world size while those launches do not, so fewer bands pay off on many GPUs (measured
    // at C2 on 8 x B200: 4/6 bands 568, 3/4 594, 2/3 599, 2/2 576 frames/s; on 2 x B200 4/6 wins).  LH_TUNE overrides.
    uint32_t bands = cascade < 0 ? ctx->tune.main_bands : ctx->tune.shadow_bands;
    if (!ctx->tune.from_environment && ctx->nranks >= 4)
        bands = ctx->nranks >= 8 ? (cascade < 0 ? 2u : 3u) : (cascade < 0 ? 3u : 4u);
    const uint32_t grid = (uint32_t)ctx->num_sms * (cascade < 0 ? ctx->tune.main_grid : ctx->tune.shadow_grid);
    const uint32_t select_grid = (uint32_t)ctx->num_sms * ctx->tune.select_grid;
    const uint32_t deferred_grid = (uint32_t)ctx->num_sms * (cascade < 0 ? 8u : ctx->tune.shadow_grid);
    uint32_t drained_to = 0, drain = 0;
    FrameParams p = params;
    p.num_bands = bands;
    for (int i = 0; i < 7; i++) p.band_fractions[i] = cascade < 0 ? ctx->tune.main_fractions[i] : 0.0f;
    const bool timed = cascade < 0 && (ctx->cfg.flags & LH_CONFIG_PASS_TIMINGS) != 0;
    raster_bands_kernel<<<1, 1024, 0, stream>>>(p);   // also opens the pass: empty raster and deferred queues
    ctx->launches++;
    for (uint32_t band = 0; band < bands; band++) {
        if (band != 0) {
            if (cascade < 0) raster_hiz_kernel<false><<<grid, 256, 0, stream>>>(p, 0);
            else raster_hiz_kernel<true><<<grid, 256, 0, stream>>>(p, (uint32_t)cascade);
            ctx->launches++;
        }
        // selection is a chain of dependent L2 round trips per entry (key -> bound -> tiles): many light CTAs
        if (cascade < 0) raster_select_kernel<false><<<select_grid, 256, 0, stream>>>(p, 0, band);
        else raster_select_kernel<true><<<select_grid, 256, 0, stream>>>(p, (uint32_t)cascade, band);
        if (timed) cudaEventRecord(ctx->raster_band_a[band], stream);
        if (cascade < 0) raster_meshlets_kernel<false><<<grid, kWarpsPerCta * 32, 0, stream>>>(p, 0, band);
        else raster_meshlets_kernel<true><<<grid, kWarpsPerCta * 32, 0, stream>>>(p, (uint32_t)cascade, band);
        if (timed) {
            cudaEventRecord(ctx->raster_band_b[band], stream);
            ctx->raster_bands_recorded = band + 1;
        }
        ctx->launches += 2;
        // The deferred queue is drained after the near bands (whose large triangles are the best occluders for the
        // bands that follow) and at the end; each drain is two launches: the first splits and rasterises, the second
        // rasterises the pieces.  The main view's drains run on their own stream BESIDE the next band's passes (the
        // hierarchical Z stays valid when keys arrive late: stored values only get nearer); a cascade defers next to
        // nothing and drains once, at the end, on its own stream.
        if ((cascade < 0 && band < 3) || band + 1 == bands) {
            cudaStream_t ds = stream;
            if (cascade < 0) {
                ds = ctx->deferred_stream;
                cudaEventRecord(ctx->band_done[band], stream);
                cudaStreamWaitEvent(ds, ctx->band_done[band], 0);
            }
            for (uint32_t round = 0; round < 2; round++) {
                if (cascade < 0) raster_deferred_kernel<false><<<deferred_grid, kDeferredWarps * 32, 0, ds>>>(p, 0, round, drained_to, band + 1, drain);
                else raster_deferred_kernel<true><<<deferred_grid, kDeferredWarps * 32, 0, ds>>>(p, (uint32_t)cascade, round, drained_to, band + 1, drain);
                ctx->launches++;
            }
            drained_to = band + 1;
            drain++;
        }
    }
    if (cascade < 0) {   // the drains join the main view before anything reads the finished buffer
        cudaEventRecord(ctx->drains_done, ctx->deferred_stream);
        cudaStreamWaitEvent(stream, ctx->drains_done, 0);
    }
    if (cascade < 0) {
        end_main_pass_kernel<<<1, 32, 0, stream>>>(p); ctx->launches++;
    }
}

}  // namespace lh
