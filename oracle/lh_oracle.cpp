// CPU restatement of the reference's geometry passes (everything up to and
// including the shadow maps).  TEST INFRASTRUCTURE -- see lh_oracle.h for the
// rules, the parity status ("parity unpinned" by the reference: it ships no
// golden vectors) and the pinned arithmetic model.  Build with
// -ffp-contract=off (oracle/Makefile).
#include "lh_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "lh_oracle_internal.hpp"

namespace lho {

unsigned thread_count(const LhoConfig* cfg) {
    unsigned n = cfg && cfg->threads ? cfg->threads : std::thread::hardware_concurrency();
    return n ? n : 1;
}

void parallel_for(uint64_t count, unsigned threads, const std::function<void(uint64_t, uint64_t)>& body) {
    if (count == 0) return;
    if (threads <= 1 || count < 2) {
        body(0, count);
        return;
    }
    // dynamic chunks: work per item is very uneven (raster)
    uint64_t chunk = std::max<uint64_t>(1, count / (threads * 16));
    std::atomic<uint64_t> next{0};
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < threads; t++)
        pool.emplace_back([&]() {
            for (;;) {
                uint64_t b = next.fetch_add(chunk);
                if (b >= count) break;
                body(b, std::min(count, b + chunk));
            }
        });
    for (auto& th : pool) th.join();
}

// ---- buffer views ---------------------------------------------------------

uint64_t* sector_counter(const LhUniforms* u, const LhoConfig* cfg, uint32_t cascade) {
    // prefix_sum_buffer_for_cascade, common/util.glsl:27-32
    return (uint64_t*)(uintptr_t)(u->num_meshlets_prefix_sum + (uint64_t)cascade * ((uint64_t)cfg->max_instances * 8 + 8));
}
LhPrefixSumValue* sector_values(const LhUniforms* u, const LhoConfig* cfg, uint32_t cascade) {
    return (LhPrefixSumValue*)(sector_counter(u, cfg, cascade) + 1);
}

// ---- culling (common/culling.glsl) ---------------------------------------------

// cull_bounding_sphere, common/culling.glsl:1-46.  Returns true when culled.
bool cull_bounding_sphere(const LhUniforms* u, const LhInstance* inst, const float sphere[4]) {
    V4 world = mat4_mul(inst->transform, V4{sphere[0], sphere[1], sphere[2], 1.0f});
    float radius = sphere[3];
    float sx = length3(V3{inst->transform.m[0][0], inst->transform.m[0][1], inst->transform.m[0][2]});
    float sy = length3(V3{inst->transform.m[1][0], inst->transform.m[1][1], inst->transform.m[1][2]});
    float sz = length3(V3{inst->transform.m[2][0], inst->transform.m[2][1], inst->transform.m[2][2]});
    float scale = fmaxf(fmaxf(sx, sy), sz);
    radius = radius * scale;

    V4 view = mat4_mul(u->initial_view, V4{world.x, world.y, world.z, 1.0f});
    float vz = -view.z;
    bool visible = vz + radius > LH_NEAR_PLANE;

    const LhMat4& P = u->perspective;
    V3 fx = normalize3(V3{P.m[0][3] + P.m[0][0], P.m[1][3] + P.m[1][0], P.m[2][3] + P.m[2][0]});
    V3 fy = normalize3(V3{P.m[0][3] + P.m[0][1], P.m[1][3] + P.m[1][1], P.m[2][3] + P.m[2][1]});
    visible = visible && (vz * fx.z + fabsf(view.x) * fx.x < radius);
    visible = visible && (vz * fy.z - fabsf(view.y) * fy.y < radius);
    return !visible;
}

// fits_entirely_inside_cascade, common/culling.glsl:64-80
static bool fits_entirely_inside_cascade(const LhMiscStorage* misc, V3 world, float radius, uint32_t k) {
    V4 v = mat4_mul(misc->shadow_view_matrices[k], V4{world.x, world.y, world.z, 1.0f});
    float R = misc->shadow_sphere_radii[k];
    return (fabsf(v.x) + radius < R) && (fabsf(v.y) + radius < R);
}

// cull_bounding_sphere_shadows, common/culling.glsl:82-135
bool cull_bounding_sphere_shadows(const LhUniforms* u, const LhInstance* inst, const float sphere[4], uint32_t cascade) {
    const LhMiscStorage* misc = (const LhMiscStorage*)(uintptr_t)u->misc_storage;
    V4 world = mat4_mul(inst->transform, V4{sphere[0], sphere[1], sphere[2], 1.0f});
    float radius = sphere[3];
    float sx = length3(V3{inst->transform.m[0][0], inst->transform.m[0][1], inst->transform.m[0][2]});
    float sy = length3(V3{inst->transform.m[1][0], inst->transform.m[1][1], inst->transform.m[1][2]});
    float sz = length3(V3{inst->transform.m[2][0], inst->transform.m[2][1], inst->transform.m[2][2]});
    radius = radius * fmaxf(fmaxf(sx, sy), sz);

    V4 v = mat4_mul(misc->shadow_view_matrices[cascade], V4{world.x, world.y, world.z, 1.0f});
    float R = misc->shadow_sphere_radii[cascade];
    bool partially = (fabsf(v.x) - radius < R) && (fabsf(v.y) - radius < R);
    if (!partially) return true;
    V3 w3{world.x, world.y, world.z};
    if (cascade > 0 && fits_entirely_inside_cascade(misc, w3, radius, 0)) return true;
    if (cascade > 1 && fits_entirely_inside_cascade(misc, w3, radius, 1)) return true;
    if (cascade > 2 && fits_entirely_inside_cascade(misc, w3, radius, 2)) return true;
    return false;
}

// cull_cone_perspective, common/culling.glsl:137-145 (axis is normalised twice)
bool cull_cone_perspective(const LhUniforms* u, const LhInstance* inst, const LhMeshlet* m) {
    V4 apex = mat4_mul(inst->transform, V4{m->cone_apex[0], m->cone_apex[1], m->cone_apex[2], 1.0f});
    V3 axis = normalize3(mat3_mul(inst->normal_transform, V3{m->cone_axis[0], m->cone_axis[1], m->cone_axis[2]}));
    V3 to = normalize3(V3{apex.x - u->camera_pos[0], apex.y - u->camera_pos[1], apex.z - u->camera_pos[2]});
    return dot3(to, normalize3(axis)) >= m->cone_cutoff;
}

// cull_cone_orthographic, common/culling.glsl:147-151
bool cull_cone_orthographic(const LhUniforms* u, const LhInstance* inst, const LhMeshlet* m) {
    V3 axis = normalize3(mat3_mul(inst->normal_transform, V3{m->cone_axis[0], m->cone_axis[1], m->cone_axis[2]}));
    return dot3(V3{u->sun_dir[0], u->sun_dir[1], u->sun_dir[2]}, axis) >= m->cone_cutoff;
}

static uint32_t dispatch_size(uint32_t width, uint32_t wg) {  // common/util.glsl:38-40
    return ((width == 0 ? 1 : width - 1) / wg) + 1;
}

// prefix_sum_binary_search, common/prefix_sum.glsl:33-46 (upper_bound on .sum)
static LhPrefixSumValue prefix_sum_binary_search(const uint64_t* counter, const LhPrefixSumValue* values, uint32_t target) {
    uint32_t count = (uint32_t)(*counter >> 32);
    uint32_t first = 0;
    while (count > 0) {
        uint32_t step = count / 2;
        uint32_t current = first + step;
        bool greater = target >= values[current].sum;
        first = greater ? current + 1 : first;
        count = greater ? count - (step + 1) : step;
    }
    return values[first];
}

static void reset_counters(const LhUniforms* u, const LhoConfig* cfg) {  // common/util.glsl:61-70
    uint32_t* draw = (uint32_t*)(uintptr_t)u->draw_calls;
    draw[0] = 0;
    draw[1] = 0;
    for (uint32_t c = 0; c < 4; c++) *sector_counter(u, cfg, c) = 0;
}

static void instance_range(const LhUniforms* u, const LhoConfig* cfg, uint32_t* first, uint32_t* end) {
    uint32_t f = cfg->instance_first, n = cfg->instance_count;
    f = std::min(f, u->num_instances);
    uint32_t e = (n == 0xFFFFFFFFu || (uint64_t)f + n > u->num_instances) ? u->num_instances : f + n;
    *first = f;
    *end = e;
}

// Meshlet cull shared by the main view (write_draw_calls) and the cascades
// (write_draw_calls_shadows).  Candidates are visited in ascending global
// meshlet id, the canonical order.
static int cull_meshlets_common(const LhUniforms* u, const LhoConfig* cfg, int cascade /* -1 = main view */) {
    uint32_t sector = cascade < 0 ? 0u : (uint32_t)cascade;
    const uint64_t* counter = sector_counter(u, cfg, sector);
    const LhPrefixSumValue* values = sector_values(u, cfg, sector);
    uint32_t total = (uint32_t)(*counter);  // prefix_sum_total
    const LhInstance* instances = (const LhInstance*)(uintptr_t)u->instances;
    uint32_t* draw = (uint32_t*)(uintptr_t)u->draw_calls;
    LhDrawIndirectCommand* cmds = (LhDrawIndirectCommand*)(draw + 2);
    LhMeshletReference* refs = (LhMeshletReference*)(uintptr_t)u->meshlet_references;
    LhDispatchIndirectCommand* disp = (LhDispatchIndirectCommand*)(uintptr_t)u->dispatches;
    disp[LH_PER_MESHLET_DISPATCH] = {dispatch_size(total, 64), 1, 1};  // reset_buffers_b / _c
    if (cascade >= 0) {
        draw[0] = 0;  // reset_buffers_c
        draw[1] = 0;
    }
    uint32_t section = cascade < 0 ? 0u : cfg->max_opaque_draws + cfg->max_alpha_clip_draws;

    std::vector<uint8_t> state(total);
    std::vector<LhMeshletReference> cand(total);
    parallel_for(total, thread_count(cfg), [&](uint64_t b, uint64_t e) {
        for (uint64_t g = b; g < e; g++) {
            // get_meshlet_reference, common/util.glsl:42-59
            LhPrefixSumValue r = prefix_sum_binary_search(counter, values, (uint32_t)g);
            const LhInstance* inst = &instances[r.index];
            const LhMeshInfo* mi = (const LhMeshInfo*)(uintptr_t)inst->mesh_info_address;
            uint32_t local = (uint32_t)g - (r.sum - mi->num_meshlets);
            cand[g] = LhMeshletReference{r.index, (uint16_t)local, 0};
            const LhMeshlet* m = (const LhMeshlet*)(uintptr_t)mi->meshlets + (uint16_t)local;
            bool alpha = (mi->flags & LH_MESH_INFO_FLAGS_ALPHA_CLIP) != 0;
            bool culled;
            if (cascade < 0) {
                culled = cull_bounding_sphere(u, inst, m->bounding_sphere);
                if (!culled && !alpha) culled = cull_cone_perspective(u, inst, m);
            } else {
                culled = cull_bounding_sphere_shadows(u, inst, m->bounding_sphere, (uint32_t)cascade);
                if (!culled && !alpha) culled = cull_cone_orthographic(u, inst, m);
            }
            state[g] = culled ? 0 : (alpha ? 2 : 1);
        }
    });
    int rc = 0;
    uint32_t n_opaque = 0, n_alpha = 0;
    for (uint32_t g = 0; g < total; g++) {
        if (!state[g]) continue;
        const LhInstance* inst = &instances[cand[g].instance_index];
        const LhMeshInfo* mi = (const LhMeshInfo*)(uintptr_t)inst->mesh_info_address;
        const LhMeshlet* m = (const LhMeshlet*)(uintptr_t)mi->meshlets + cand[g].meshlet_index;
        uint32_t draw_index;
        if (state[g] == 2) {
            if (n_alpha >= cfg->max_alpha_clip_draws) { n_alpha++; rc = -4; continue; }
            draw_index = cfg->max_opaque_draws + n_alpha++;
        } else {
            if (n_opaque >= cfg->max_opaque_draws) { n_opaque++; rc = -4; continue; }
            draw_index = n_opaque++;
        }
        uint32_t slot = draw_index + section;
        cmds[draw_index] = LhDrawIndirectCommand{(uint32_t)m->triangle_count * 3u, 1u, 0u, slot};
        refs[slot] = cand[g];
    }
    draw[0] = n_opaque;
    draw[1] = n_alpha;
    return rc;
}

// ---- rasterisation ------------------------------------------------------------

// Transforms the vertices of one meshlet: load_vertex_data + calculate_world_pos
// (rasterization.glsl:12-34, common/util.glsl:3-16) then `matrix * (world, 1)`.
void transform_meshlet_vertices(const LhInstance* inst, const LhMeshInfo* mi, const LhMeshlet* m, const LhMat4& matrix,
                                V4* clip_out, V3* world_out) {
    const uint16_t* positions = (const uint16_t*)(uintptr_t)mi->positions;
    bool wide = (mi->flags & LH_MESH_INFO_FLAGS_32_BIT_INDICES) != 0;
    for (uint32_t v = 0; v < m->index_count; v++) {
        uint32_t idx = wide ? ((const uint32_t*)(uintptr_t)mi->indices)[m->index_offset + v]
                            : ((const uint16_t*)(uintptr_t)mi->indices)[m->index_offset + v];
        V4 p{(float)positions[3 * idx + 0], (float)positions[3 * idx + 1], (float)positions[3 * idx + 2], 1.0f};
        V4 world = mat4_mul(inst->transform, p);
        if (world_out) world_out[v] = V3{world.x, world.y, world.z};
        clip_out[v] = mat4_mul(matrix, V4{world.x, world.y, world.z, 1.0f});
    }
}

struct RasterTarget {
    uint32_t width, height;
    uint64_t* vis;      // main view
    uint32_t* shadow;   // shadow layer (float bits)
};

static inline void atomic_min_u64(uint64_t* p, uint64_t v) {
    uint64_t cur = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
}
static inline void atomic_min_u32(uint32_t* p, uint32_t v) {
    uint32_t cur = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
    }
}

// Rasterises the meshlet in `slot` of the meshlet-reference table.
static void raster_meshlet(const LhUniforms* u, const LhoConfig* cfg, uint32_t slot, uint32_t local_slot, bool alpha, int cascade,
                           const RasterTarget& rt) {
    const LhMeshletReference* refs = (const LhMeshletReference*)(uintptr_t)u->meshlet_references;
    const LhInstance* instances = (const LhInstance*)(uintptr_t)u->instances;
    LhMeshletReference ref = refs[slot];
    const LhInstance* inst = &instances[ref.instance_index];
    const LhMeshInfo* mi = (const LhMeshInfo*)(uintptr_t)inst->mesh_info_address;
    const LhMeshlet* m = (const LhMeshlet*)(uintptr_t)mi->meshlets + ref.meshlet_index;
    const LhMiscStorage* misc = (const LhMiscStorage*)(uintptr_t)u->misc_storage;
    const LhMat4& matrix = cascade < 0 ? u->combined_perspective_view : misc->shadow_matrices[cascade];

    V4 clip[LH_MAX_MESHLET_UNIQUE_VERTICES * 4];  // index_count is a u8: up to 255
    transform_meshlet_vertices(inst, mi, m, matrix, clip, nullptr);

    // cull modes: pipelines.cpp:10-25,255-315 (vis opaque BACK, shadow opaque FRONT, alpha-clip NONE)
    int cull = alpha ? CULL_NONE : (cascade < 0 ? CULL_BACK : CULL_FRONT);
    const uint8_t* micro = (const uint8_t*)(uintptr_t)mi->micro_indices + m->triangle_offset;
    AlphaTest at = make_alpha_test(cfg, mi, alpha);
    for (uint32_t t = 0; t < m->triangle_count; t++) {
        uint32_t i0 = micro[3 * t + 0], i1 = micro[3 * t + 1], i2 = micro[3 * t + 2];
        V4 tri[3] = {clip[i0], clip[i1], clip[i2]};
        float uv[3][2] = {{0, 0}, {0, 0}, {0, 0}};
        if (at.enabled) load_triangle_uvs(mi, m, i0, i1, i2, uv);
        if (cascade < 0) {
            uint32_t lo = (local_slot << LH_VIS_TRI_BITS) | t;
            raster_clip_triangle(tri, uv, cull, rt.width, rt.height, /*clip_near=*/true, at,
                                 [&](int x, int y, float z) {
                                     z = fminf(fmaxf(z, 0.0f), 1.0f);
                                     if (!(z > 0.0f)) return;  // depth test GREATER against the 0.0 clear
                                     uint32_t bits;
                                     memcpy(&bits, &z, 4);
                                     uint64_t key = ((uint64_t)(~bits) << 32) | lo;
                                     atomic_min_u64(&rt.vis[(uint64_t)y * rt.width + x], key);
                                 });
        } else {
            raster_clip_triangle(tri, uv, cull, rt.width, rt.height, /*clip_near=*/false, at,
                                 [&](int x, int y, float z) {
                                     z = z + 0.0f;  // canonicalise -0
                                     if (!(z >= 0.0f && z < 1.0f)) return;  // 0 <= z <= w clip, LESS against the 1.0 clear
                                     uint32_t bits;
                                     memcpy(&bits, &z, 4);
                                     atomic_min_u32(&rt.shadow[(uint64_t)y * rt.width + x], bits);
                                 });
        }
    }
}

static void raster_lists(const LhUniforms* u, const LhoConfig* cfg, int cascade, const RasterTarget& rt) {
    const uint32_t* draw = (const uint32_t*)(uintptr_t)u->draw_calls;
    uint32_t n_opaque = std::min(draw[0], cfg->max_opaque_draws);   // drawIndirectCount clamps to maxDrawCount
    uint32_t n_alpha = std::min(draw[1], cfg->max_alpha_clip_draws);
    uint32_t section = cascade < 0 ? 0u : cfg->max_opaque_draws + cfg->max_alpha_clip_draws;
    parallel_for((uint64_t)n_opaque + n_alpha, thread_count(cfg), [&](uint64_t b, uint64_t e) {
        for (uint64_t k = b; k < e; k++) {
            bool alpha = k >= n_opaque;
            uint32_t local = alpha ? cfg->max_opaque_draws + (uint32_t)(k - n_opaque) : (uint32_t)k;
            raster_meshlet(u, cfg, section + local, local, alpha, cascade, rt);
        }
    });
}

}  // namespace lho

using namespace lho;

extern "C" {

void lho_config_default(LhoConfig* cfg) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->max_instances = LH_MAX_INSTANCES;
    cfg->max_opaque_draws = LH_MAX_OPAQUE_DRAWS;
    cfg->max_alpha_clip_draws = LH_MAX_ALPHA_CLIP_DRAWS;
    cfg->shadowmap_size = LH_SHADOWMAP_SIZE;
    cfg->instance_count = 0xFFFFFFFFu;
}

int lho_reset(const LhUniforms* u, const LhoConfig* cfg) {
    LhMiscStorage* misc = (LhMiscStorage*)(uintptr_t)u->misc_storage;
    misc->min_depth = 0xFFFFFFFFu;
    misc->max_depth = 0;
    LhDispatchIndirectCommand* disp = (LhDispatchIndirectCommand*)(uintptr_t)u->dispatches;
    disp[LH_PER_INSTANCE_DISPATCH] = {dispatch_size(u->num_instances, 64), 1, 1};
    disp[LH_PER_SHADOW_INSTANCE_DISPATCH] = {dispatch_size(u->num_instances, 16), 1, 1};
    reset_counters(u, cfg);
    return 0;
}

int lho_cull_instances(const LhUniforms* u, const LhoConfig* cfg) {
    const LhInstance* instances = (const LhInstance*)(uintptr_t)u->instances;
    uint64_t* counter = sector_counter(u, cfg, 0);
    LhPrefixSumValue* values = sector_values(u, cfg, 0);
    uint32_t first, end;
    instance_range(u, cfg, &first, &end);
    std::vector<uint8_t> vis(end - first);
    parallel_for(end - first, thread_count(cfg), [&](uint64_t b, uint64_t e) {
        for (uint64_t k = b; k < e; k++) {
            const LhInstance* inst = &instances[first + k];
            const LhMeshInfo* mi = (const LhMeshInfo*)(uintptr_t)inst->mesh_info_address;
            vis[k] = !cull_bounding_sphere(u, inst, mi->bounding_sphere);
        }
    });
    int rc = 0;
    uint64_t c = *counter;
    for (uint32_t i = first; i < end; i++) {
        if (!vis[i - first]) continue;
        const LhMeshInfo* mi = (const LhMeshInfo*)(uintptr_t)instances[i].mesh_info_address;
        // prefix_sum_inclusive_append, common/prefix_sum.glsl:17-30
        uint64_t old = c;
        c += (uint64_t)mi->num_meshlets | (1ull << 32);
        uint32_t slot = (uint32_t)(old >> 32);
        if (slot >= cfg->max_instances) { rc = -4; continue; }
        values[slot].index = i;
        values[slot].sum = (uint32_t)old + mi->num_meshlets;
    }
    *counter = c;
    return rc;
}

int lho_cull_meshlets(const LhUniforms* u, const LhoConfig* cfg) { return cull_meshlets_common(u, cfg, -1); }

int lho_raster_visbuffer(const LhUniforms* u, const LhoConfig* cfg, uint64_t* vis) {
    uint64_t n = (uint64_t)u->window_size[0] * u->window_size[1];
    for (uint64_t i = 0; i < n; i++) vis[i] = LH_VIS_CLEAR;
    RasterTarget rt{u->window_size[0], u->window_size[1], vis, nullptr};
    raster_lists(u, cfg, -1, rt);
    return 0;
}

int lho_reduce_depth(const LhUniforms* u, const LhoConfig* cfg, const uint64_t* vis) {
    (void)cfg;
    LhMiscStorage* misc = (LhMiscStorage*)(uintptr_t)u->misc_storage;
    uint64_t n = (uint64_t)u->window_size[0] * u->window_size[1];
    uint32_t mn = misc->min_depth, mx = misc->max_depth;
    for (uint64_t i = 0; i < n; i++) {
        uint32_t d = ~(uint32_t)(vis[i] >> 32);
        if (d != 0 && d < mn) mn = d;  // min_if_not_zero, read_depth.glsl:3-7
        if (d > mx) mx = d;
    }
    misc->min_depth = mn;
    misc->max_depth = mx;
    return 0;
}

int lho_generate_shadow_matrices(const LhUniforms* u, const LhoConfig* cfg) {
    LhMiscStorage* misc = (LhMiscStorage*)(uintptr_t)u->misc_storage;
    float min_depth, max_depth;
    memcpy(&min_depth, &misc->min_depth, 4);
    memcpy(&max_depth, &misc->max_depth, 4);
    for (uint32_t cascade = 0; cascade < 4; cascade++) {
        float min_distance = LH_NEAR_PLANE / max_depth;  // convert_infinite_reverze_z_depth, util.glsl:23-25
        float max_distance = LH_NEAR_PLANE / min_depth;
        float t1 = lho_pinned_powf(0.5f, u->cascade_split_pow);
        float t2 = lho_pinned_powf(0.75f, u->cascade_split_pow);
        float splits[4] = {max_depth, LH_NEAR_PLANE / (min_distance * (1.0f - t1) + max_distance * t1),
                           LH_NEAR_PLANE / (min_distance * (1.0f - t2) + max_distance * t2), min_depth};
        float last_split = cascade > 0 ? splits[cascade - 1] : max_depth;
        float split = splits[cascade];
        V3 corners[8] = {{-1, 1, last_split}, {1, 1, last_split}, {1, -1, last_split}, {-1, -1, last_split},
                         {-1, 1, split},      {1, 1, split},      {1, -1, split},      {-1, -1, split}};
        for (int i = 0; i < 8; i++) {
            V4 c = mat4_mul(u->inv_perspective_view, V4{corners[i].x, corners[i].y, corners[i].z, 1.0f});
            corners[i] = V3{c.x / c.w, c.y / c.w, c.z / c.w};
        }
        V3 center{0, 0, 0};
        for (int i = 0; i < 8; i++) center = V3{center.x + corners[i].x, center.y + corners[i].y, center.z + corners[i].z};
        center = V3{center.x / 8.0f, center.y / 8.0f, center.z / 8.0f};
        float radius = 0.0f;
        for (int i = 0; i < 8; i++) {
            float d = length3(V3{corners[i].x - center.x, corners[i].y - center.y, corners[i].z - center.z});
            radius = fmaxf(radius, d);
        }
        radius = ceilf(radius * 16.0f) / 16.0f;

        V3 sun{u->sun_dir[0], u->sun_dir[1], u->sun_dir[2]};
        V3 eye{u->shadow_cam_distance * sun.x + center.x, u->shadow_cam_distance * sun.y + center.y,
               u->shadow_cam_distance * sun.z + center.z};
        // lookAt, common/matrices.glsl:9-19
        V3 f = normalize3(V3{center.x - eye.x, center.y - eye.y, center.z - eye.z});
        V3 r = normalize3(cross3(f, V3{0, 1, 0}));
        V3 up = cross3(r, f);
        LhMat4 view{};
        view.m[0][0] = r.x; view.m[0][1] = up.x; view.m[0][2] = -f.x; view.m[0][3] = 0;
        view.m[1][0] = r.y; view.m[1][1] = up.y; view.m[1][2] = -f.y; view.m[1][3] = 0;
        view.m[2][0] = r.z; view.m[2][1] = up.z; view.m[2][2] = -f.z; view.m[2][3] = 0;
        view.m[3][0] = -dot3(r, eye); view.m[3][1] = -dot3(up, eye); view.m[3][2] = dot3(f, eye); view.m[3][3] = 1;
        // OrthographicProjection(-r, r, -r, r, 0, 2*dist), common/matrices.glsl:21-40
        float left = -radius, right = radius, bottom = -radius, top = radius, nearp = 0.0f;
        float farp = u->shadow_cam_distance * 2.0f;
        float rml = right - left, rpl = right + left, tmb = top - bottom, tpb = top + bottom, fmn = farp - nearp;
        LhMat4 proj{};
        proj.m[0][0] = 2.0f / rml;
        proj.m[1][1] = -2.0f / tmb;
        proj.m[2][2] = -1.0f / fmn;
        proj.m[3][0] = -(rpl / rml); proj.m[3][1] = -(tpb / tmb); proj.m[3][2] = -(nearp / fmn); proj.m[3][3] = 1.0f;
        LhMat4 matrix = mat4_mat4(proj, view);
        LhMat4 bias{};  // generate_shadow_matrices.glsl:6-11
        bias.m[0][0] = 0.5f; bias.m[1][1] = 0.5f; bias.m[2][2] = 1.0f;
        bias.m[3][0] = 0.5f; bias.m[3][1] = 0.5f; bias.m[3][3] = 1.0f;
        misc->shadow_matrices[cascade] = matrix;
        misc->uv_space_shadow_matrices[cascade] = mat4_mat4(bias, matrix);
        misc->shadow_view_matrices[cascade] = view;
        misc->shadow_sphere_radii[cascade] = radius;
    }
    reset_counters(u, cfg);
    return 0;
}

int lho_cull_instances_shadows(const LhUniforms* u, const LhoConfig* cfg) {
    const LhInstance* instances = (const LhInstance*)(uintptr_t)u->instances;
    uint32_t first, end;
    instance_range(u, cfg, &first, &end);
    int rc = 0;
    for (uint32_t cascade = 0; cascade < 4; cascade++) {
        uint64_t* counter = sector_counter(u, cfg, cascade);
        LhPrefixSumValue* values = sector_values(u, cfg, cascade);
        uint64_t c = *counter;
        for (uint32_t i = first; i < end; i++) {
            const LhInstance* inst = &instances[i];
            const LhMeshInfo* mi = (const LhMeshInfo*)(uintptr_t)inst->mesh_info_address;
            if (cull_bounding_sphere_shadows(u, inst, mi->bounding_sphere, cascade)) continue;
            uint64_t old = c;
            c += (uint64_t)mi->num_meshlets | (1ull << 32);
            uint32_t slot = (uint32_t)(old >> 32);
            if (slot >= cfg->max_instances) { rc = -4; continue; }
            values[slot].index = i;
            values[slot].sum = (uint32_t)old + mi->num_meshlets;
        }
        *counter = c;
    }
    return rc;
}

int lho_shadow_cascade(const LhUniforms* u, const LhoConfig* cfg, uint32_t cascade, float* shadow) {
    if (cascade >= 4) return -1;
    int rc = cull_meshlets_common(u, cfg, (int)cascade);
    uint64_t n = (uint64_t)cfg->shadowmap_size * cfg->shadowmap_size;
    float* layer = shadow + n * cascade;
    for (uint64_t i = 0; i < n; i++) layer[i] = 1.0f;  // clear 1.0, rendering.cpp:384
    RasterTarget rt{cfg->shadowmap_size, cfg->shadowmap_size, nullptr, (uint32_t*)layer};
    raster_lists(u, cfg, (int)cascade, rt);
    return rc;
}

int lho_render_frame(const LhUniforms* u, const LhoConfig* cfg, uint64_t* vis, float* shadow, float* out) {
    int rc = 0, r;
    if ((r = lho_reset(u, cfg))) rc = r;
    if ((r = lho_cull_instances(u, cfg))) rc = r;
    if ((r = lho_cull_meshlets(u, cfg))) rc = r;
    if ((r = lho_raster_visbuffer(u, cfg, vis))) rc = r;
    if ((r = lho_reduce_depth(u, cfg, vis))) rc = r;
    if ((r = lho_generate_shadow_matrices(u, cfg))) rc = r;
    if ((r = lho_cull_instances_shadows(u, cfg))) rc = r;
    for (uint32_t c = 0; c < 4; c++)
        if ((r = lho_shadow_cascade(u, cfg, c, shadow))) rc = r;
    if (out && (r = lho_resolve(u, cfg, vis, shadow, out))) rc = r;
    return rc;
}

void lho_composite_min(uint64_t* dst, const uint64_t* src, uint64_t count) {
    for (uint64_t i = 0; i < count; i++) dst[i] = src[i] < dst[i] ? src[i] : dst[i];
}

// pow(x, p) for x > 0 evaluated with a fixed sequence of IEEE double
// operations (no libm), so the device reproduces it bit for bit.
// single decisions on caller-supplied inputs: compared with the reference's GLSL on random inputs (tests/test_ref_shaders.py)
int lho_test_cull_sphere(const LhUniforms* u, const LhInstance* inst, const float sphere[4]) { return lho::cull_bounding_sphere(u, inst, sphere) ? 1 : 0; }
int lho_test_cull_sphere_shadows(const LhUniforms* u, const LhInstance* inst, const float sphere[4], uint32_t cascade) {
    return lho::cull_bounding_sphere_shadows(u, inst, sphere, cascade) ? 1 : 0;
}
int lho_test_cull_cone_perspective(const LhUniforms* u, const LhInstance* inst, const LhMeshlet* m) { return lho::cull_cone_perspective(u, inst, m) ? 1 : 0; }
int lho_test_cull_cone_orthographic(const LhUniforms* u, const LhInstance* inst, const LhMeshlet* m) { return lho::cull_cone_orthographic(u, inst, m) ? 1 : 0; }

// clip-space positions the oracle's rasteriser starts from for the draw in `slot`: 3 * triangle_count vertices in draw
// order (what the reference's vertex shader sees as gl_VertexIndex).  cascade < 0: the main view.  Returns the count.
uint32_t lho_test_clip_positions(const LhUniforms* u, int cascade, uint32_t slot, float* out) {
    const LhMeshletReference* refs = (const LhMeshletReference*)(uintptr_t)u->meshlet_references;
    const LhInstance* inst = (const LhInstance*)(uintptr_t)u->instances + refs[slot].instance_index;
    const LhMeshInfo* mi = (const LhMeshInfo*)(uintptr_t)inst->mesh_info_address;
    const LhMeshlet* m = (const LhMeshlet*)(uintptr_t)mi->meshlets + refs[slot].meshlet_index;
    const LhMiscStorage* misc = (const LhMiscStorage*)(uintptr_t)u->misc_storage;
    lho::V4 clip[256];
    lho::transform_meshlet_vertices(inst, mi, m, cascade < 0 ? u->combined_perspective_view : misc->shadow_matrices[cascade], clip, nullptr);
    const uint8_t* micro = (const uint8_t*)(uintptr_t)mi->micro_indices + m->triangle_offset;
    for (uint32_t k = 0; k < 3u * m->triangle_count; k++) memcpy(out + 4 * k, &clip[micro[k]], 16);
    return 3u * m->triangle_count;
}

float lho_pinned_powf(float xf, float pf) {
    if (!(xf > 0.0f)) return xf == 0.0f ? 0.0f : NAN;
    double x = (double)xf;
    // x = m * 2^e with m in [sqrt(1/2), sqrt(2))
    uint64_t bits;
    memcpy(&bits, &x, 8);
    int e = (int)((bits >> 52) & 0x7FF) - 1023;
    bits = (bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
    double m;
    memcpy(&m, &bits, 8);
    if (m > 1.4142135623730951) {
        m = m * 0.5;
        e += 1;
    }
    double s = (m - 1.0) / (m + 1.0);
    double s2 = s * s;
    double term = s, acc = 0.0;
    for (int k = 1; k <= 25; k += 2) {
        acc = acc + term / (double)k;
        term = term * s2;
    }
    const double ln2 = 0.6931471805599453;
    double lnx = (double)e * ln2 + 2.0 * acc;
    double y = (double)pf * lnx;
    double kf = floor(y / ln2 + 0.5);
    double r = y - kf * ln2;
    double et = 1.0, ex = 1.0;
    for (int k = 1; k <= 18; k++) {
        et = et * r / (double)k;
        ex = ex + et;
    }
    int ki = (int)kf;
    if (ki > 1000) return INFINITY;
    if (ki < -1000) return 0.0f;
    uint64_t pb = (uint64_t)(ki + 1023) << 52;
    double p2;
    memcpy(&p2, &pb, 8);
    return (float)(ex * p2);
}

}  // extern "C"
