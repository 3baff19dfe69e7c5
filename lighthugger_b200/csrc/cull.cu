// Culling passes: reset, instance cull + single-pass prefix sum, meshlet cull
// + stable compaction ("mesh-shader emulation").
//
// Replaces (reference file:line)
//   reset_buffers_a/b/c            src/shaders/compute/reset_buffers.glsl:9-57
//   cull_instances[_shadows]       src/shaders/cull_instances.comp:7-56
//   prefix_sum_inclusive_append    src/shaders/common/prefix_sum.glsl:17-30
//   write_draw_calls[_shadows]     src/shaders/write_draw_calls.comp:7-117
//   cull_bounding_sphere*, cone    src/shaders/common/culling.glsl:1-151
//
// B200 design.  The reference appends with one contended 64-bit atomicAdd per
// surviving instance / two u32 atomics per surviving meshlet, which makes the
// list order schedule-dependent.  Here both lists are produced by a
// decoupled-lookback scan over tiles (one 64-bit status|aggregate word per
// tile, no global atomics on the data path), so the output is the canonical
// ascending-invocation order, deterministically.  Instances are read with
// coalesced 128-bit loads into shared memory; the meshlet pass is a persistent
// grid that takes tiles by ticket, finds the tile's instance window with two
// binary searches and lets every thread finish its search inside that window.
#include "lh_internal.cuh"
#include "raster_bounds.cuh"

namespace lh {

namespace {

constexpr unsigned long long kStatusShift = 62;
constexpr unsigned long long kStatusAggregate = 1ull << kStatusShift;
constexpr unsigned long long kStatusInclusive = 2ull << kStatusShift;
constexpr unsigned long long kPayloadMask = (1ull << kStatusShift) - 1;

__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Decoupled look-back (Merrill & Garland): called by all 32 lanes of one warp.
// `aggregate` is this tile's packed total; returns the packed exclusive prefix
// (sum of all previous tiles, plus `base` for tile 0's predecessors).
__device__ unsigned long long lookback(unsigned long long* states, uint32_t tile, unsigned long long aggregate,
                                       unsigned long long base) {
    const uint32_t lane = threadIdx.x & 31;
    if (tile == 0) {
        if (lane == 0) st_relaxed(&states[0], kStatusInclusive | ((base + aggregate) & kPayloadMask));
        return base;
    }
    if (lane == 0) st_relaxed(&states[tile], kStatusAggregate | (aggregate & kPayloadMask));
    unsigned long long exclusive = 0;
    int64_t j = (int64_t)tile - 1;
    for (;;) {
        int64_t mine = j - lane;
        unsigned long long v = kStatusInclusive;  // lanes before tile 0 contribute nothing
        if (mine >= 0) {
            do {
                v = ld_relaxed(&states[mine]);
            } while ((v >> kStatusShift) == 0);
        }
        uint32_t incl = __ballot_sync(0xFFFFFFFFu, (v >> kStatusShift) == 2);
        // lanes nearer than the first inclusive one (lower lane index) add their aggregates
        uint32_t first = incl ? (uint32_t)(__ffs(incl) - 1) : 32u;
        unsigned long long part = (lane <= first && mine >= 0) ? (v & kPayloadMask) : 0ull;
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
        exclusive += part;
        if (incl) break;
        j -= 32;
    }
    exclusive &= kPayloadMask;
    if (lane == 0) st_relaxed(&states[tile], kStatusInclusive | ((exclusive + aggregate) & kPayloadMask));
    return exclusive;
}

__device__ __forceinline__ uint32_t dispatch_size(uint32_t width, uint32_t wg) {  // common/util.glsl:38-40
    return ((width == 0 ? 1u : width - 1u) / wg) + 1u;
}

__device__ __forceinline__ unsigned long long* sector_counter(const FrameParams& P, uint32_t cascade) {
    return (unsigned long long*)(P.u.num_meshlets_prefix_sum + (uint64_t)cascade * ((uint64_t)P.max_instances * 8 + 8));
}

// ---- sphere / cone tests (common/culling.glsl) ----

struct FrustumPlanes {
    V3 fx, fy;
};
__device__ __forceinline__ FrustumPlanes frustum_planes(const LhUniforms& u) {
    const float(*Pm)[4] = u.perspective.m;
    FrustumPlanes f;
    f.fx = normalize3(V3{padd(Pm[0][3], Pm[0][0]), padd(Pm[1][3], Pm[1][0]), padd(Pm[2][3], Pm[2][0])});
    f.fy = normalize3(V3{padd(Pm[0][3], Pm[0][1]), padd(Pm[1][3], Pm[1][1]), padd(Pm[2][3], Pm[2][1])});
    return f;
}

// world-space centre + scaled radius of a sphere under an instance transform
__device__ __forceinline__ V4 world_sphere(const float* T, const float* sphere) {
    V4 w = mat4_mul(T, sphere[0], sphere[1], sphere[2], 1.0f);
    float sx = length3(V3{T[0], T[1], T[2]});
    float sy = length3(V3{T[4], T[5], T[6]});
    float sz = length3(V3{T[8], T[9], T[10]});
    w.w = pmul(sphere[3], fmaxf(fmaxf(sx, sy), sz));
    return w;
}

// the same with the instance's scale max(|T[0].xyz|, |T[1].xyz|, |T[2].xyz|) already at hand (identical arithmetic)
__device__ __forceinline__ float instance_scale(const float* T) {
    return fmaxf(fmaxf(length3(V3{T[0], T[1], T[2]}), length3(V3{T[4], T[5], T[6]})), length3(V3{T[8], T[9], T[10]}));
}
__device__ __forceinline__ V4 world_sphere_scaled(const float* T, const float* sphere, float scale) {
    V4 w = mat4_mul(T, sphere[0], sphere[1], sphere[2], 1.0f);
    w.w = pmul(sphere[3], scale);
    return w;
}

// cull_bounding_sphere, culling.glsl:1-46: true when culled
__device__ __forceinline__ bool cull_sphere_main(const LhUniforms& u, const FrustumPlanes& f, V4 ws) {
    V4 v = mat4_mul(&u.initial_view.m[0][0], ws.x, ws.y, ws.z, 1.0f);
    float vz = -v.z;
    bool visible = padd(vz, ws.w) > LH_NEAR_PLANE;
    visible = visible && (padd(pmul(vz, f.fx.z), pmul(fabsf(v.x), f.fx.x)) < ws.w);
    visible = visible && (psub(pmul(vz, f.fy.z), pmul(fabsf(v.y), f.fy.y)) < ws.w);
    return !visible;
}

// cull_bounding_sphere_shadows, culling.glsl:82-135
__device__ __forceinline__ bool cull_sphere_shadows(const LhMiscStorage* misc, V4 ws, uint32_t cascade) {
    V4 v = mat4_mul(&misc->shadow_view_matrices[cascade].m[0][0], ws.x, ws.y, ws.z, 1.0f);
    float R = misc->shadow_sphere_radii[cascade];
    bool partially = (psub(fabsf(v.x), ws.w) < R) && (psub(fabsf(v.y), ws.w) < R);
    if (!partially) return true;
    for (uint32_t k = 0; k < cascade; k++) {
        V4 vk = mat4_mul(&misc->shadow_view_matrices[k].m[0][0], ws.x, ws.y, ws.z, 1.0f);
        float Rk = misc->shadow_sphere_radii[k];
        if ((padd(fabsf(vk.x), ws.w) < Rk) && (padd(fabsf(vk.y), ws.w) < Rk)) return true;
    }
    return false;
}

// Sort-last ownership by screen stripes (new; lh_set_instance_stripes).  A pure function of (instance, camera), evaluated
// identically by every rank, so the ranks' lists partition the single-GPU list.  Main view: stripe of the centre's
// ndc x under initial_view (the culling camera); cascade: stripe of its light-space x over [-R, R].
// the stripe whose [bounds[s], bounds[s + 1]) holds t01; the outer stripes are open-ended, NaN goes to stripe 0
__device__ __forceinline__ uint32_t stripe_of(float t01, uint32_t stripes, const float* bounds) {
    uint32_t s = 0;
    while (s + 1 < stripes && t01 >= bounds[s + 1]) s++;
    return s;
}
__device__ __forceinline__ bool owned_main(const FrameParams& P, V4 ws) {
    if (P.stripes[0] == 0) return true;
    const V4 v = mat4_mul(&P.u.initial_view.m[0][0], ws.x, ws.y, ws.z, 1.0f);
    const float ndc = pdiv(pmul(P.u.perspective.m[0][0], v.x), fmaxf(-v.z, LH_NEAR_PLANE));
    return stripe_of(padd(pmul(ndc, 0.5f), 0.5f), P.stripes[0], P.stripe_bounds[0]) % P.stripe_count[0] == P.stripe_index[0];
}
__device__ __forceinline__ bool owned_shadow(const FrameParams& P, const LhMiscStorage* misc, V4 ws, uint32_t cascade) {
    if (P.stripes[1] == 0) return true;
    const V4 v = mat4_mul(&misc->shadow_view_matrices[cascade].m[0][0], ws.x, ws.y, ws.z, 1.0f);
    const float t = padd(pmul(pdiv(v.x, misc->shadow_sphere_radii[cascade]), 0.5f), 0.5f);
    return stripe_of(t, P.stripes[1], P.stripe_bounds[1]) % P.stripe_count[1] == P.stripe_index[1];
}

// ---- K1: reset_buffers_a + reset_counters ----
__global__ void reset_kernel(const __grid_constant__ FrameParams P) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    LhMiscStorage* misc = (LhMiscStorage*)P.u.misc_storage;
    misc->min_depth = 0xFFFFFFFFu;
    misc->max_depth = 0u;
    LhDispatchIndirectCommand* disp = (LhDispatchIndirectCommand*)P.u.dispatches;
    disp[LH_PER_INSTANCE_DISPATCH] = LhDispatchIndirectCommand{dispatch_size(P.u.num_instances, 64), 1, 1};
    disp[LH_PER_SHADOW_INSTANCE_DISPATCH] = LhDispatchIndirectCommand{dispatch_size(P.u.num_instances, 16), 1, 1};
    uint32_t* draw = (uint32_t*)P.u.draw_calls;
    draw[0] = 0;
    draw[1] = 0;
    for (uint32_t c = 0; c < 4; c++) *sector_counter(P, c) = 0ull;
    FrameState* s = P.state;
    s->overflow = 0;
    s->main_large_count = 0;
    s->main_visible = s->main_candidates = s->main_opaque = s->main_alpha_clip = 0;
    s->rasterized_triangles = 0ull;
    for (int c = 0; c < 5; c++) {
        s->skipped_empty[c] = s->skipped_occluded[c] = s->queued[c] = 0;
        s->samples[c] = 0ull;
    }
    for (int c = 0; c < 4; c++) {
        s->shadow_opaque[c] = 0;
        s->shadow_alpha_clip[c] = 0;
        s->shadow_visible[c] = 0;
        s->shadow_candidates[c] = 0;
    }
}

// reset_counters() as run by generate_shadow_matrices (common/util.glsl:61-70)
__global__ void reset_counters_kernel(const __grid_constant__ FrameParams P) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    uint32_t* draw = (uint32_t*)P.u.draw_calls;
    draw[0] = 0;
    draw[1] = 0;
    for (uint32_t c = 0; c < 4; c++) *sector_counter(P, c) = 0ull;
}

// ---- K2 / K9: instance cull + scan ----
// One CTA per 256-instance tile (blockIdx.y = cascade for the shadow variant).
template <bool SHADOWS>
__global__ void __launch_bounds__(kInstanceTile) cull_instances_kernel(const __grid_constant__ FrameParams P, uint32_t tiles_per_sector) {
    __shared__ int4 stage[kInstanceTile * 7];  // 256 Instances, 112 B each
    __shared__ uint32_t warp_count[8], warp_sum[8];
    __shared__ unsigned long long tile_exclusive;

    const uint32_t tile = blockIdx.x, cascade = SHADOWS ? blockIdx.y : 0u;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t first = (SHADOWS ? P.shadow_instance_first : P.instance_first) + tile * kInstanceTile;
    const uint32_t n_here = min(kInstanceTile, (SHADOWS ? P.shadow_instance_end : P.instance_end) - first);

    // coalesced 128-bit loads of the tile's Instance records
    const int4* src = (const int4*)((const LhInstance*)P.u.instances + first);
    for (uint32_t k = tid; k < n_here * 7; k += kInstanceTile) stage[k] = __ldg(src + k);
    __syncthreads();

    bool keep = false;
    uint32_t value = 0;
    if (tid < n_here) {
        const float* T = (const float*)&stage[tid * 7];
        const LhMeshInfo* mi = (const LhMeshInfo*)(*(const unsigned long long*)(T + 16));
        float sphere[4] = {__ldg(&mi->bounding_sphere[0]), __ldg(&mi->bounding_sphere[1]), __ldg(&mi->bounding_sphere[2]),
                           __ldg(&mi->bounding_sphere[3])};
        value = __ldg(&mi->num_meshlets);
        V4 ws = world_sphere(T, sphere);
        if (SHADOWS) {
            keep = !cull_sphere_shadows((const LhMiscStorage*)P.u.misc_storage, ws, cascade) &&
                   owned_shadow(P, (const LhMiscStorage*)P.u.misc_storage, ws, cascade);
        } else {
            FrustumPlanes f = frustum_planes(P.u);
            keep = !cull_sphere_main(P.u, f, ws) && owned_main(P, ws);
        }
    }
    if (!keep) value = 0;

    // block scan of (keep, value)
    uint32_t ballot = __ballot_sync(0xFFFFFFFFu, keep);
    uint32_t incl = value;
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += n;
    }
    if (lane == 31) {
        warp_count[warp] = __popc(ballot);
        warp_sum[warp] = incl;
    }
    __syncthreads();
    uint32_t count_before = 0, sum_before = 0, tile_count = 0, tile_sum = 0;
    for (uint32_t w = 0; w < 8; w++) {
        if (w < warp) {
            count_before += warp_count[w];
            sum_before += warp_sum[w];
        }
        tile_count += warp_count[w];
        tile_sum += warp_sum[w];
    }
    unsigned long long* counter = sector_counter(P, cascade);
    unsigned long long* states = P.instance_tiles + (uint64_t)cascade * tiles_per_sector;
    if (warp == 0) {
        // only tile 0 reads the counter (appends continue from its current value, prefix_sum.glsl:22-23)
        unsigned long long base = tile == 0 ? *counter : 0ull;
        unsigned long long aggregate = ((unsigned long long)tile_count << 32) + tile_sum;
        unsigned long long ex = lookback(states, tile, aggregate, base);
        if (lane == 0) tile_exclusive = ex;
    }
    __syncthreads();
    const unsigned long long ex = tile_exclusive;
    if (keep) {
        uint32_t slot = (uint32_t)(ex >> 32) + count_before + __popc(ballot & ((1u << lane) - 1u));
        uint32_t sum = (uint32_t)ex + sum_before + incl;
        if (slot < P.max_instances) {
            LhPrefixSumValue* values = (LhPrefixSumValue*)(counter + 1);
            values[slot] = LhPrefixSumValue{first + tid, sum};
        } else {
            atomicOr(&P.state->overflow, (uint32_t)LH_OVERFLOW_INSTANCES);
        }
    }
    // the last tile owns the final counter; the last CTA out cleans the tile states
    if (tid == 0) {
        if (tile == gridDim.x - 1) {
            unsigned long long total = (ex + (((unsigned long long)tile_count << 32) + tile_sum)) & kPayloadMask;
            *counter = total;
            if (SHADOWS) {
                P.state->shadow_visible[cascade] = (uint32_t)(total >> 32);
                P.state->shadow_candidates[cascade] = (uint32_t)total;
            } else {
                P.state->main_visible = (uint32_t)(total >> 32);
                P.state->main_candidates = (uint32_t)total;
            }
        }
        __threadfence();
        uint32_t finished = atomicAdd(&P.state->done[cascade], 1u);
        if (finished == gridDim.x - 1) {
            for (uint32_t t = 0; t < gridDim.x; t++) st_relaxed(&states[t], 0ull);
            P.state->done[cascade] = 0;
        }
    }
}

// ---- K4 / K11: meshlet cull + stable compaction ----

// upper_bound on .sum over values[lo, hi): prefix_sum_binary_search, prefix_sum.glsl:33-46
__device__ __forceinline__ uint32_t upper_bound_sum(const LhPrefixSumValue* values, uint32_t lo, uint32_t hi, uint32_t target) {
    uint32_t first = lo, count = hi - lo;
    while (count > 0) {
        uint32_t step = count / 2, cur = first + step;
        bool greater = target >= __ldg(&values[cur].sum);
        first = greater ? cur + 1 : first;
        count = greater ? count - (step + 1) : step;
    }
    return first;
}

// What the meshlet cull needs of one instance, staged once per tile: a tile of 1024 candidates spans only a few
// instances (one or two at 1000 meshlets per instance), so the two dependent loads behind `values[]`, the transform,
// its scale (three square roots) and the bound's sigma are fetched / computed per instance instead of per candidate.
constexpr uint32_t kWindowMax = 8;
struct alignas(16) WindowInstance {
    float T[16];
    float N[9];
    float scale, sigma;
    uint32_t index, sum, num_meshlets, flags;
    unsigned long long meshlets;
};

struct Survivor {
    uint32_t instance_index;
    uint32_t meshlet_and_tris;  // meshlet_index | triangle_count << 16
};

template <bool SHADOWS>
__global__ void __launch_bounds__(256) cull_meshlets_kernel(const __grid_constant__ FrameParams P, uint32_t cascade) {
    // a tile has at most kMeshletTile survivors over both classes: the opaque ones are staged from the front (ascending
    // candidate order), the alpha-clip ones from the back; each with the screen bound the rasteriser will want
    __shared__ Survivor staged[kMeshletTile];
    __shared__ uint4 staged_bound[kMeshletTile];
    __shared__ BoundView bview;
    __shared__ WindowInstance win[kWindowMax];
    __shared__ uint32_t warp_counts[2][8];   // per class, the eight warps' survivor counts of the current tile
    __shared__ uint32_t window[2];
    __shared__ uint32_t tile_ticket;
    __shared__ unsigned long long tile_exclusive;
    __shared__ FrustumPlanes planes;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t sector = SHADOWS ? cascade : 0u;
    unsigned long long* counter = sector_counter(P, sector);
    const LhPrefixSumValue* values = (const LhPrefixSumValue*)(counter + 1);
    const unsigned long long c64 = *counter;
    const uint32_t n_values = min((uint32_t)(c64 >> 32), P.max_instances);
    // when the value list overflowed, only the candidates the stored values cover are valid
    const uint32_t total = (uint32_t)(c64 >> 32) <= P.max_instances ? (uint32_t)c64 : (n_values ? __ldg(&values[n_values - 1].sum) : 0u);
    const uint32_t num_tiles = (total + kMeshletTile - 1) / kMeshletTile;
    const LhInstance* instances = (const LhInstance*)P.u.instances;
    const LhMiscStorage* misc = (const LhMiscStorage*)P.u.misc_storage;
    uint32_t* draw = (uint32_t*)P.pass_draw_calls;
    LhDrawIndirectCommand* cmds = (LhDrawIndirectCommand*)(draw + 2);
    LhMeshletReference* refs = (LhMeshletReference*)P.pass_refs;   // the view's own list: slot = pass_slot_base + index
    FrameState* st = P.state;
    PassState* ps = P.pass;

    if (tid == 0) planes = frustum_planes(P.u);
    if (tid == 32) make_bound_view<SHADOWS>(P, cascade, bview);
    uint32_t empty_bounds = 0;
    if (blockIdx.x == 0 && tid == 0) {
        // reset_buffers_b / _c: indirect size for the reference's 64-wide dispatch
        // (a cascade that runs concurrently on a side stream with internal lists leaves the host-visible slot alone, so
        // what is left there is the last cascade's, as in the reference)
        if (P.pass_draw_calls == P.u.draw_calls)
            ((LhDispatchIndirectCommand*)P.u.dispatches)[LH_PER_MESHLET_DISPATCH] = LhDispatchIndirectCommand{dispatch_size((uint32_t)c64, 64), 1, 1};
        if (num_tiles == 0) {
            draw[0] = 0;
            draw[1] = 0;
            if (SHADOWS) {
                st->shadow_opaque[cascade] = 0;
                st->shadow_alpha_clip[cascade] = 0;
            } else {
                st->main_opaque = 0;
                st->main_alpha_clip = 0;
            }
        }
    }

    for (;;) {
        __syncthreads();
        if (tid == 0) tile_ticket = atomicAdd(&ps->cull_ticket, 1u);
        __syncthreads();
        const uint32_t tile = tile_ticket;
        if (tile >= num_tiles) break;
        const uint32_t g0 = tile * kMeshletTile;
        const uint32_t g_end = min(total, g0 + kMeshletTile);
        // instance window of the tile: two searches, then every thread searches inside it
        if (tid < 2) window[tid] = upper_bound_sum(values, 0, n_values, tid == 0 ? g0 : g_end - 1);
        __syncthreads();
        const uint32_t win_lo = window[0], win_hi = min(window[1] + 1, n_values);
        const uint32_t win_n = win_hi - win_lo;
        const bool windowed = win_n <= kWindowMax;
        if (windowed && tid < win_n) {
            WindowInstance& wi = win[tid];
            const LhPrefixSumValue r = values[win_lo + tid];
            const LhInstance* inst = instances + r.index;
#pragma unroll
            for (int k = 0; k < 4; k++) *(float4*)&wi.T[4 * k] = __ldg((const float4*)&inst->transform + k);
#pragma unroll
            for (int k = 0; k < 9; k++) wi.N[k] = __ldg(&inst->normal_transform.m[0][0] + k);
            const LhMeshInfo* mi = (const LhMeshInfo*)__ldg((const unsigned long long*)&inst->mesh_info_address);
            wi.index = r.index;
            wi.sum = r.sum;
            wi.num_meshlets = __ldg(&mi->num_meshlets);
            wi.flags = __ldg(&mi->flags);
            wi.meshlets = __ldg((const unsigned long long*)&mi->meshlets);
            wi.scale = instance_scale(wi.T);
            wi.sigma = sigma_max_bound16(wi.T);
        }
        __syncthreads();

        // Each warp owns a contiguous eighth of the tile (128 candidates, four rounds of 32) and stages its survivors in
        // its own eighth of the staging arrays -- opaque from the front, alpha-clip from the back -- so the rounds need no
        // CTA barrier; warps are in candidate order, so the lists stay in ascending candidate order.
        constexpr uint32_t kPerWarp = kMeshletTile / 8;
        uint32_t base_count[2] = {0, 0};   // this warp's survivors so far, per class
        for (uint32_t round = 0; round < kPerWarp / 32; round++) {
            const uint32_t g = g0 + warp * kPerWarp + round * 32 + lane;
            uint32_t cls = 2;  // 0 opaque survivor, 1 alpha-clip survivor, 2 culled
            Survivor sv{0, 0};
            uint4 bound = make_uint4(0u, 0u, 0u, 0u);
            if (g < g_end) {
                // get_meshlet_reference, common/util.glsl:42-59
                LhPrefixSumValue r;
                float Tl[16], scale, sigma;
                uint32_t nm, flags;
                const LhMeshlet* meshlets;
                const float* N;
                if (windowed) {
                    uint32_t w = 0;
                    while (w + 1 < win_n && g >= win[w].sum) w++;   // the first instance whose inclusive sum exceeds g
                    const WindowInstance& wi = win[w];
                    r.index = wi.index;
                    r.sum = wi.sum;
#pragma unroll
                    for (int k = 0; k < 16; k++) Tl[k] = wi.T[k];
                    scale = wi.scale;
                    sigma = wi.sigma;
                    nm = wi.num_meshlets;
                    flags = wi.flags;
                    meshlets = (const LhMeshlet*)wi.meshlets;
                    N = wi.N;
                } else {
                    const uint32_t vi = upper_bound_sum(values, win_lo, win_hi, g);
                    r = values[vi];
                    const LhInstance* inst = instances + r.index;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        float4 c = __ldg((const float4*)&inst->transform + k);
                        Tl[4 * k] = c.x; Tl[4 * k + 1] = c.y; Tl[4 * k + 2] = c.z; Tl[4 * k + 3] = c.w;
                    }
                    const LhMeshInfo* mi = (const LhMeshInfo*)__ldg((const unsigned long long*)&inst->mesh_info_address);
                    nm = __ldg(&mi->num_meshlets);
                    flags = __ldg(&mi->flags);
                    meshlets = (const LhMeshlet*)__ldg((const unsigned long long*)&mi->meshlets);
                    scale = instance_scale(Tl);
                    sigma = -1.0f;
                    N = &inst->normal_transform.m[0][0];
                }
                const uint32_t local = (g - (r.sum - nm)) & 0xFFFFu;
                const bool alpha = (flags & LH_MESH_INFO_FLAGS_ALPHA_CLIP) != 0;
                const LhMeshlet* m = meshlets + local;
                // 56-byte meshlet as seven 8-byte loads
                unsigned long long raw[7];
#pragma unroll
                for (int k = 0; k < 7; k++) raw[k] = __ldg((const unsigned long long*)m + k);
                const float* mf = (const float*)raw;  // cone_apex[3] cone_axis[3] cutoff sphere[4]
                const uint32_t tri_count = ((const uint8_t*)raw)[52];
                V4 ws = world_sphere_scaled(Tl, mf + 7, scale);
                bool culled;
                if (SHADOWS) culled = cull_sphere_shadows(misc, ws, cascade);
                else culled = cull_sphere_main(P.u, planes, ws);
                if (!culled && !alpha) {
                    float Nl[9];
#pragma unroll
                    for (int k = 0; k < 9; k++) Nl[k] = N[k];
                    V3 axis = normalize3(mat3_mul(Nl, mf[3], mf[4], mf[5]));
                    if (SHADOWS) {
                        // cull_cone_orthographic, culling.glsl:147-151
                        culled = dot3(V3{P.u.sun_dir[0], P.u.sun_dir[1], P.u.sun_dir[2]}, axis) >= mf[6];
                    } else {
                        // cull_cone_perspective, culling.glsl:137-145 (axis normalised twice)
                        V4 apex = mat4_mul(Tl, mf[0], mf[1], mf[2], 1.0f);
                        V3 to = normalize3(V3{psub(apex.x, P.u.camera_pos[0]), psub(apex.y, P.u.camera_pos[1]), psub(apex.z, P.u.camera_pos[2])});
                        culled = dot3(to, normalize3(axis)) >= mf[6];
                    }
                }
                if (!culled) {
                    cls = alpha ? 1u : 0u;
                    sv.instance_index = r.index;
                    sv.meshlet_and_tris = local | (tri_count << 16);
                    // the rasteriser's bound of this meshlet, and its depth key into the band histogram (aggregated over
                    // the lanes that fall into the same bin: neighbouring meshlets mostly do)
                    const MeshletBound mb = meshlet_bound<SHADOWS>(bview, Tl, mf + 7, sigma);
                    bound = make_uint4(mb.rect_x, mb.rect_y, mb.edge, mb.key);
                    const uint32_t bin = mb.key == kKeyEmpty ? kHistBins : mb.key >> kHistShift;
                    if (mb.key == kKeyEmpty) empty_bounds++;
                    const uint32_t peers = __match_any_sync(__activemask(), bin);
                    if (bin < kHistBins && lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(P.depth_hist + bin, (uint32_t)__popc(peers));
                }
            }
            // stable ranks inside the round, per class
            const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, cls == 0), b1 = __ballot_sync(0xFFFFFFFFu, cls == 1);
            if (cls < 2) {
                const uint32_t rank = (cls ? base_count[1] : base_count[0]) + __popc((cls ? b1 : b0) & ((1u << lane) - 1u));
                const uint32_t at = warp * kPerWarp + (cls ? kPerWarp - 1u - rank : rank);
                staged[at] = sv;
                staged_bound[at] = bound;
            }
            base_count[0] += __popc(b0);
            base_count[1] += __popc(b1);
        }
        if (lane == 0) {
            warp_counts[0][warp] = base_count[0];
            warp_counts[1][warp] = base_count[1];
        }
        __syncthreads();
        uint32_t before[2] = {0, 0}, tile_count[2] = {0, 0};   // survivors of the earlier warps / of the whole tile
#pragma unroll
        for (uint32_t w = 0; w < 8; w++) {
            const uint32_t c0 = warp_counts[0][w], c1 = warp_counts[1][w];
            if (w < warp) {
                before[0] += c0;
                before[1] += c1;
            }
            tile_count[0] += c0;
            tile_count[1] += c1;
        }
        // tile aggregate -> exclusive prefix over all earlier tiles
        if (warp == 0) {
            unsigned long long aggregate = ((unsigned long long)tile_count[0] << 31) | tile_count[1];
            unsigned long long ex = lookback(P.meshlet_tiles, tile, aggregate, 0ull);
            if (lane == 0) tile_exclusive = ex;
        }
        __syncthreads();
        const unsigned long long ex = tile_exclusive;
        const uint32_t ex_count[2] = {(uint32_t)(ex >> 31), (uint32_t)(ex & 0x7FFFFFFFu)};
        // write-out of the compacted lists: every warp moves its own staged survivors, coalesced
#pragma unroll
        for (uint32_t cls = 0; cls < 2; cls++) {
            const uint32_t cap = cls ? P.max_alpha : P.max_opaque;
            const uint32_t first_draw = cls ? P.max_opaque : 0u;  // ALPHA_CLIP_DRAWS_OFFSET
            for (uint32_t k = lane; k < base_count[cls]; k += 32) {
                uint32_t n = ex_count[cls] + before[cls] + k;
                if (n >= cap) {
                    atomicOr(&st->overflow, cls ? (uint32_t)LH_OVERFLOW_ALPHA_CLIP : (uint32_t)LH_OVERFLOW_OPAQUE);
                    continue;
                }
                const uint32_t at = warp * kPerWarp + (cls ? kPerWarp - 1u - k : k);
                Survivor s = staged[at];
                uint32_t draw_index = first_draw + n, slot = draw_index + P.pass_slot_base;
                const uint4 bound = staged_bound[at];
                ((uint4*)P.meshlet_bounds)[draw_index] = bound;
                P.depth_keys[draw_index] = bound.w;
                LhMeshletReference ref;
                ref.instance_index = s.instance_index;
                ref.meshlet_index = (uint16_t)(s.meshlet_and_tris & 0xFFFFu);
                ref._pad0 = 0;
                refs[draw_index] = ref;
                if (P.flags & LH_CONFIG_WRITE_DRAW_COMMANDS)
                    cmds[draw_index] = LhDrawIndirectCommand{(s.meshlet_and_tris >> 16) * 3u, 1u, 0u, slot};
            }
        }
        if (tile == num_tiles - 1 && tid == 0) {
            uint32_t n0 = ex_count[0] + tile_count[0], n1 = ex_count[1] + tile_count[1];
            draw[0] = n0;
            draw[1] = n1;
            if (SHADOWS) {
                st->shadow_opaque[cascade] = n0;
                st->shadow_alpha_clip[cascade] = n1;
            } else {
                st->main_opaque = n0;
                st->main_alpha_clip = n1;
            }
        }
    }
    if (P.flags & LH_CONFIG_PASS_TIMINGS) {
        empty_bounds = __reduce_add_sync(0xFFFFFFFFu, empty_bounds);
        if (lane == 0 && empty_bounds) atomicAdd(&st->skipped_empty[SHADOWS ? 1u + cascade : 0u], empty_bounds);
    }
    // self-cleaning: the last CTA out zeroes the tile states and the ticket
    __shared__ uint32_t last_out;
    if (tid == 0) {
        __threadfence();
        last_out = atomicAdd(&ps->cull_done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last_out) {
        for (uint32_t t = tid; t < num_tiles; t += 256) st_relaxed(&P.meshlet_tiles[t], 0ull);
        if (tid == 0) {
            ps->cull_ticket = 0;
            ps->cull_done = 0;
        }
    }
}

__global__ void fill_u64_kernel(unsigned long long* dst, uint64_t count, unsigned long long value) {
    // 128-bit stores; count is even-padded by the caller's allocation
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    ulonglong2 v{value, value};
    for (uint64_t k = i; k < count / 2; k += stride) ((ulonglong2*)dst)[k] = v;
    if (i == 0 && (count & 1)) dst[count - 1] = value;
}
__global__ void fill_u32_kernel(uint32_t* dst, uint64_t count, uint32_t value) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint4 v{value, value, value, value};
    for (uint64_t k = i; k < count / 4; k += stride) ((uint4*)dst)[k] = v;
    if (i < (count & 3)) dst[(count & ~3ull) + i] = value;
}

}  // namespace

void launch_reset(LhContext* ctx, const FrameParams& p) { reset_kernel<<<1, 32, 0, ctx->stream>>>(p); ctx->launches++; }

void launch_reset_counters(LhContext* ctx, const FrameParams& p) { reset_counters_kernel<<<1, 32, 0, ctx->stream>>>(p); ctx->launches++; }

void launch_cull_instances(LhContext* ctx, const FrameParams& p, bool shadows) {
    uint32_t n = shadows ? p.shadow_instance_end - p.shadow_instance_first : p.instance_end - p.instance_first;
    if (n == 0) return;
    uint32_t tiles = (n + kInstanceTile - 1) / kInstanceTile;
    if (shadows) cull_instances_kernel<true><<<dim3(tiles, 4), kInstanceTile, 0, ctx->stream>>>(p, ctx->instance_tiles_per_sector);
    else cull_instances_kernel<false><<<dim3(tiles, 1), kInstanceTile, 0, ctx->stream>>>(p, ctx->instance_tiles_per_sector);
    ctx->launches++;
}

void launch_cull_meshlets(LhContext* ctx, const FrameParams& p, int cascade, cudaStream_t stream) {
    if (!stream) stream = ctx->stream;
    // persistent grid: tiles are taken by ticket, so any grid size is correct
    uint32_t grid = (uint32_t)ctx->num_sms * 4;
    if (cascade < 0) cull_meshlets_kernel<false><<<grid, 256, 0, stream>>>(p, 0);
    else cull_meshlets_kernel<true><<<grid, 256, 0, stream>>>(p, (uint32_t)cascade);
    ctx->launches++;
}

void launch_fill_u64(LhContext* ctx, unsigned long long* dst, uint64_t count, unsigned long long value) {
    if (!count) return;
    fill_u64_kernel<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(dst, count, value); ctx->launches++;
}
void launch_fill_u32(LhContext* ctx, uint32_t* dst, uint64_t count, uint32_t value, cudaStream_t stream) {
    if (!count) return;
    fill_u32_kernel<<<ctx->num_sms * 8, 256, 0, stream ? stream : ctx->stream>>>(dst, count, value); ctx->launches++;
}

}  // namespace lh
