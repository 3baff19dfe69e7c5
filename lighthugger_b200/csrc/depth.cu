// Depth min/max reduction, cascade matrices, and small readback helpers.
//
// Replaces (reference file:line)
//   read_depth                  src/shaders/compute/read_depth.glsl:22-83
//   generate_shadow_matrices    src/shaders/compute/generate_shadow_matrices.glsl:15-110
//   lookAt / OrthographicProjection   src/shaders/common/matrices.glsl:9-40
//
// B200 design.  The reference gathers a separate D32F image 4x4 texels per
// thread and issues one atomic per subgroup.  Depth here lives in the high
// word of the packed visibility key, so the pass is a pure streaming read of
// the 8 B/pixel buffer with 128-bit loads (two pixels per load), a register
// min/max, one shuffle tree per warp and one atomic pair per CTA: HBM/L2
// bandwidth bound by construction.
#include "lh_internal.cuh"

namespace lh {

namespace {

__global__ void __launch_bounds__(256) reduce_depth_kernel(const __grid_constant__ FrameParams P, uint64_t pixels) {
    const ulonglong2* src = (const ulonglong2*)P.vis;
    const uint64_t pairs = pixels / 2;
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
    auto take = [&](unsigned long long key) {
        uint32_t d = ~(uint32_t)(key >> 32);
        if (d != 0u) mn = min(mn, d);  // min_if_not_zero, read_depth.glsl:3-7
        mx = max(mx, d);
    };
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // 4 independent 128-bit loads in flight per thread
    for (; i + 3 * stride < pairs; i += 4 * stride) {
        ulonglong2 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
        take(a.x); take(a.y); take(b.x); take(b.y); take(c.x); take(c.y); take(d.x); take(d.y);
    }
    for (; i < pairs; i += stride) {
        ulonglong2 a = src[i];
        take(a.x); take(a.y);
    }
    if ((pixels & 1) && blockIdx.x == 0 && threadIdx.x == 0) take(P.vis[pixels - 1]);
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
    }
    __shared__ uint32_t smn[8], smx[8];
    if ((threadIdx.x & 31) == 0) {
        smn[threadIdx.x >> 5] = mn;
        smx[threadIdx.x >> 5] = mx;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) {
            mn = min(mn, smn[w]);
            mx = max(mx, smx[w]);
        }
        LhMiscStorage* misc = (LhMiscStorage*)P.u.misc_storage;
        if (mn != 0xFFFFFFFFu) atomicMin(&misc->min_depth, mn);
        if (mx != 0u) atomicMax(&misc->max_depth, mx);
    }
}

__device__ void mat4_mat4(const float* a, const float* b, float* out) {
    for (int c = 0; c < 4; c++) {
        V4 col = mat4_mul(a, b[4 * c], b[4 * c + 1], b[4 * c + 2], b[4 * c + 3]);
        out[4 * c] = col.x; out[4 * c + 1] = col.y; out[4 * c + 2] = col.z; out[4 * c + 3] = col.w;
    }
}

// K8: four threads, one cascade each (generate_shadow_matrices.glsl:15-110)
__global__ void shadow_matrices_kernel(const __grid_constant__ FrameParams P) {
    const uint32_t cascade = threadIdx.x;
    LhMiscStorage* misc = (LhMiscStorage*)P.u.misc_storage;
    if (cascade < 4) {
        const float min_depth = __uint_as_float(misc->min_depth), max_depth = __uint_as_float(misc->max_depth);
        const float min_distance = pdiv(LH_NEAR_PLANE, max_depth);  // convert_infinite_reverze_z_depth, util.glsl:23-25
        const float max_distance = pdiv(LH_NEAR_PLANE, min_depth);
        const float t1 = pinned_powf(0.5f, P.u.cascade_split_pow), t2 = pinned_powf(0.75f, P.u.cascade_split_pow);
        float splits[4];
        splits[0] = max_depth;
        splits[1] = pdiv(LH_NEAR_PLANE, padd(pmul(min_distance, psub(1.0f, t1)), pmul(max_distance, t1)));
        splits[2] = pdiv(LH_NEAR_PLANE, padd(pmul(min_distance, psub(1.0f, t2)), pmul(max_distance, t2)));
        splits[3] = min_depth;
        const float last_split = cascade > 0 ? splits[cascade - 1] : max_depth;
        const float split = splits[cascade];
        V3 corners[8] = {{-1, 1, last_split}, {1, 1, last_split}, {1, -1, last_split}, {-1, -1, last_split},
                         {-1, 1, split},      {1, 1, split},      {1, -1, split},      {-1, -1, split}};
        const float* inv = &P.u.inv_perspective_view.m[0][0];
        for (int i = 0; i < 8; i++) {
            V4 c = mat4_mul(inv, corners[i].x, corners[i].y, corners[i].z, 1.0f);
            corners[i] = V3{pdiv(c.x, c.w), pdiv(c.y, c.w), pdiv(c.z, c.w)};
        }
        V3 center{0, 0, 0};
        for (int i = 0; i < 8; i++) center = V3{padd(center.x, corners[i].x), padd(center.y, corners[i].y), padd(center.z, corners[i].z)};
        center = V3{pdiv(center.x, 8.0f), pdiv(center.y, 8.0f), pdiv(center.z, 8.0f)};
        float radius = 0.0f;
        for (int i = 0; i < 8; i++)
            radius = fmaxf(radius, length3(V3{psub(corners[i].x, center.x), psub(corners[i].y, center.y), psub(corners[i].z, center.z)}));
        radius = pdiv(ceilf(pmul(radius, 16.0f)), 16.0f);

        const V3 sun{P.u.sun_dir[0], P.u.sun_dir[1], P.u.sun_dir[2]};
        const float dist = P.u.shadow_cam_distance;
        const V3 eye{padd(pmul(dist, sun.x), center.x), padd(pmul(dist, sun.y), center.y), padd(pmul(dist, sun.z), center.z)};
        // lookAt, matrices.glsl:9-19
        const V3 f = normalize3(V3{psub(center.x, eye.x), psub(center.y, eye.y), psub(center.z, eye.z)});
        const V3 r = normalize3(cross3(f, V3{0, 1, 0}));
        const V3 up = cross3(r, f);
        float view[16] = {r.x, up.x, -f.x, 0, r.y, up.y, -f.y, 0, r.z, up.z, -f.z, 0, -dot3(r, eye), -dot3(up, eye), dot3(f, eye), 1};
        // OrthographicProjection(-radius, radius, -radius, radius, 0, 2 dist), matrices.glsl:21-40
        const float left = -radius, right = radius, bottom = -radius, top = radius, nearp = 0.0f, farp = pmul(dist, 2.0f);
        const float rml = psub(right, left), rpl = padd(right, left), tmb = psub(top, bottom), tpb = padd(top, bottom), fmn = psub(farp, nearp);
        float proj[16] = {};
        proj[0] = pdiv(2.0f, rml);
        proj[5] = pdiv(-2.0f, tmb);
        proj[10] = pdiv(-1.0f, fmn);
        proj[12] = -pdiv(rpl, rml);
        proj[13] = -pdiv(tpb, tmb);
        proj[14] = -pdiv(nearp, fmn);
        proj[15] = 1.0f;
        float matrix[16], uv_matrix[16];
        mat4_mat4(proj, view, matrix);
        const float bias[16] = {0.5f, 0, 0, 0, 0, 0.5f, 0, 0, 0, 0, 1.0f, 0, 0.5f, 0.5f, 0, 1.0f};
        mat4_mat4(bias, matrix, uv_matrix);
        for (int k = 0; k < 16; k++) {
            (&misc->shadow_matrices[cascade].m[0][0])[k] = matrix[k];
            (&misc->uv_space_shadow_matrices[cascade].m[0][0])[k] = uv_matrix[k];
            (&misc->shadow_view_matrices[cascade].m[0][0])[k] = view[k];
        }
        misc->shadow_sphere_radii[cascade] = radius;
    }
    // reset_counters(), common/util.glsl:61-70 (every invocation writes the same zeros)
    if (threadIdx.x == 0) {
        uint32_t* draw = (uint32_t*)P.u.draw_calls;
        draw[0] = 0;
        draw[1] = 0;
        for (uint32_t c = 0; c < 4; c++)
            *(unsigned long long*)(P.u.num_meshlets_prefix_sum + (uint64_t)c * ((uint64_t)P.max_instances * 8 + 8)) = 0ull;
    }
}

__global__ void extract_depth_kernel(const unsigned long long* vis, float* depth, uint64_t pixels) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i < pixels; i += (uint64_t)gridDim.x * blockDim.x) depth[i] = __uint_as_float(~(uint32_t)(vis[i] >> 32));
}

// collects the per-frame counters into one struct (one D2H copy for the host)
__global__ void gather_stats_kernel(const __grid_constant__ FrameParams P, LhFrameStats* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    LhFrameStats s;
    s._pad1 = 0;
    const FrameState* st = P.state;
    const LhMiscStorage* misc = (const LhMiscStorage*)P.u.misc_storage;
    s.num_visible_instances = st->main_visible;
    s.num_candidate_meshlets = st->main_candidates;
    s.num_opaque = st->main_opaque;
    s.num_alpha_clip = st->main_alpha_clip;
    s.min_depth = misc->min_depth;
    s.max_depth = misc->max_depth;
    for (int c = 0; c < 4; c++) {
        s.shadow_visible_instances[c] = st->shadow_visible[c];
        s.shadow_candidate_meshlets[c] = st->shadow_candidates[c];
        s.shadow_opaque[c] = st->shadow_opaque[c];
        s.shadow_alpha_clip[c] = st->shadow_alpha_clip[c];
    }
    s.num_large_triangles = st->main_large_count;
    s.overflow = st->overflow;
    s.rasterized_triangles = st->rasterized_triangles;
    for (int c = 0; c < 5; c++) {
        s.meshlets_skipped_no_sample[c] = st->skipped_empty[c];
        s.meshlets_skipped_occluded[c] = st->skipped_occluded[c];
        s.meshlets_queued[c] = st->queued[c];
        s.emitted_samples[c] = st->samples[c];
    }
    *out = s;
}

}  // namespace

void launch_reduce_depth(LhContext* ctx, const FrameParams& p) {
    uint64_t pixels = (uint64_t)p.u.window_size[0] * p.u.window_size[1];
    if (!pixels) return;
    // 4K: 4.1 M 128-bit loads; 148 x 8 CTAs x 256 threads x 4 loads in flight covers it in ~4 passes
    uint32_t grid = (uint32_t)ctx->num_sms * 8;
    reduce_depth_kernel<<<grid, 256, 0, ctx->stream>>>(p, pixels); ctx->launches++;
}

void launch_shadow_matrices(LhContext* ctx, const FrameParams& p) { shadow_matrices_kernel<<<1, 32, 0, ctx->stream>>>(p); ctx->launches++; }

void launch_extract_depth(LhContext* ctx, const FrameParams& p) {
    uint64_t pixels = (uint64_t)ctx->cfg.width * ctx->cfg.height;
    extract_depth_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->vis, ctx->depth_plane, pixels); ctx->launches++;
}

void launch_gather_stats(LhContext* ctx, const FrameParams& p, LhFrameStats* dst) { gather_stats_kernel<<<1, 32, 0, ctx->stream>>>(p, dst); ctx->launches++; }

}  // namespace lh
