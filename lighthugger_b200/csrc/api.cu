// C ABI of the CUDA layer (include/lighthugger_b200.h): context, buffers,
// textures, targets, the per-pass entry points in src/rendering.cpp order, and
// readback.  There is no CPU fallback anywhere in this library.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <nvtx3/nvToolsExt.h>

#include "lh_internal.cuh"

using namespace lh;

namespace {
thread_local std::string g_create_error;
}

namespace lh {

int fail(LhContext* ctx, int code, const char* what, cudaError_t e) {
    std::string msg = what ? what : "";
    if (e != cudaSuccess) {
        msg += ": ";
        msg += cudaGetErrorString(e);
        if (e == cudaErrorMemoryAllocation) code = LH_ERR_OUT_OF_MEMORY;
        cudaGetLastError();
    }
    if (ctx) ctx->error = msg;
    else g_create_error = msg;
    return code;
}

FrameParams make_params(LhContext* ctx, int cascade, int set, bool internal_lists) {
    FrameParams p;
    memset(&p, 0, sizeof(p));
    p.u = ctx->uniforms;
    p.max_instances = ctx->cfg.max_instances;
    p.max_opaque = ctx->cfg.max_opaque_draws;
    p.max_alpha = ctx->cfg.max_alpha_clip_draws;
    uint32_t first = ctx->instance_first < p.u.num_instances ? ctx->instance_first : p.u.num_instances;
    uint64_t end = (uint64_t)first + ctx->instance_count;
    p.instance_first = first;
    p.instance_end = (ctx->instance_count == 0xFFFFFFFFu || end > p.u.num_instances) ? p.u.num_instances : (uint32_t)end;
    p.shadow_instance_first = p.instance_first;
    p.shadow_instance_end = p.instance_end;
    if (ctx->shadow_range_set) {
        const uint32_t sfirst = ctx->shadow_instance_first < p.u.num_instances ? ctx->shadow_instance_first : p.u.num_instances;
        const uint64_t send = (uint64_t)sfirst + ctx->shadow_instance_count;
        p.shadow_instance_first = sfirst;
        p.shadow_instance_end = send > p.u.num_instances ? p.u.num_instances : (uint32_t)send;
    }
    for (int v = 0; v < 2; v++) {
        p.stripes[v] = ctx->stripes[v];
        p.stripe_index[v] = ctx->stripe_index[v];
        p.stripe_count[v] = ctx->stripe_count[v];
        for (uint32_t s = 0; s <= kMaxStripes; s++) p.stripe_bounds[v][s] = ctx->stripe_bounds[v][s];
    }
    p.small_max_area = ctx->tune.small_max_area;
    p.shadowmap_size = ctx->cfg.shadowmap_size;
    p.flags = ctx->cfg.flags;
    p.max_large = ctx->cfg.max_large_triangles;
    p.state = ctx->state;
    p.instance_tiles = ctx->instance_tiles;
    const LhContext::PassScratch& sc = ctx->scratch[set];
    p.pass = sc.pass;
    p.meshlet_tiles = sc.meshlet_tiles;
    p.large_queue = sc.large_queue;
    p.meshlet_bounds = sc.meshlet_bounds;
    p.raster_records = sc.raster_records;
    p.first_wave_ctas = (uint32_t)ctx->num_sms * 4u;   // LH_RASTER_MIN_CTAS
    p.depth_keys = sc.depth_keys;
    p.depth_hist = sc.depth_hist;
    p.hiz = sc.hiz;
    // a cascade's lists live in section 1 of the host-visible buffers (MESHLET_INDICES_BUFFER_SECTION_OFFSET), unless it
    // runs concurrently with another cascade, in which case they live in the scratch set
    const uint64_t section = (uint64_t)p.max_opaque + p.max_alpha;
    p.pass_slot_base = cascade < 0 ? 0u : (uint32_t)section;
    if (internal_lists) {
        p.pass_draw_calls = (uint64_t)(uintptr_t)sc.draw_calls;
        p.pass_refs = (uint64_t)(uintptr_t)sc.refs;
    } else {
        p.pass_draw_calls = p.u.draw_calls;
        p.pass_refs = p.u.meshlet_references + (cascade < 0 ? 0ull : section * sizeof(LhMeshletReference));
    }
    p.vis = ctx->vis;
    p.shadow = ctx->shadow;
    p.textures = ctx->tex_dev;
    p.tex_dims = ctx->tex_dims_dev;
    p.slot_offset = ctx->slot_offset;
    p.resolve_refs = ctx->gathered_refs ? (const LhMeshletReference*)ctx->gathered_refs : (const LhMeshletReference*)p.u.meshlet_references;
    p.scene = ctx->scene;
    p.resolve_tmp = ctx->resolve_tmp;
    p.display = ctx->display;
    p.sky = ctx->sky_tex;
    p.lut = ctx->lut_tex;
    p.resolve_y0 = 0;
    p.resolve_y1 = p.u.window_size[1];
    p.rank = (uint32_t)ctx->rank;
    p.nranks = (uint32_t)ctx->nranks;
    p.slots_per_rank = ctx->slots_per_rank;
    p.depth_pair = ctx->comm_scratch;
    if (peer_exchange(ctx)) {
        p.peer_vis = (const unsigned long long* const*)ctx->peer_vis_dev;
        p.peer_refs = (const LhMeshletReference* const*)ctx->peer_refs_dev;
        p.scene_remote = ctx->scene_remote;
        p.resolve_y0 = ctx->row0;
        p.resolve_y1 = ctx->row1;
    }
    return p;
}

}  // namespace lh

namespace {

// NVTX ranges named like the reference's Tracy zones (src/rendering.cpp:42,140,172,198,290,305,326,463,492), so a
// timeline of this layer reads like one of the reference; passes the reference has no zone for get descriptive names
const char* const kPassZone[LH_PASS_COUNT] = {
    "reset buffers", "cull instances", "cull meshlets and write draw calls", "visbuffer rendering", "composite (sort-last)",
    "depth reduction", "generate shadow matrices", "cull instances for shadows", "shadowmap rasterization", "render geometry",
    "display transform", "shadow exchange (sort-last)", "frame barrier (sort-last)"};

struct PassScope {
    LhContext* ctx;
    int pass;
    PassScope(LhContext* c, int p) : ctx(c), pass(p) {
        nvtxRangePushA(kPassZone[pass]);
        if (ctx->cfg.flags & LH_CONFIG_PASS_TIMINGS) cudaEventRecord(ctx->pass_start[pass], ctx->stream);
    }
    ~PassScope() {
        if (ctx->cfg.flags & LH_CONFIG_PASS_TIMINGS) {
            cudaEventRecord(ctx->pass_stop[pass], ctx->stream);
            ctx->pass_recorded[pass] = true;
            if (pass != LH_PASS_DISPLAY_TRANSFORM) {
                if (ctx->first_pass < 0) ctx->first_pass = pass;
                ctx->last_pass = pass;
            }
        }
        nvtxRangePop();
    }
};
struct ZoneScope {   // a range without a timed pass behind it
    explicit ZoneScope(const char* name) { nvtxRangePushA(name); }
    ~ZoneScope() { nvtxRangePop(); }
};

int check_launch(LhContext* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, LH_ERR_CUDA, what, e);
    return LH_OK;
}

int require_frame(LhContext* ctx) {
    if (!ctx) return LH_ERR_INVALID_ARGUMENT;
    if (!ctx->have_uniforms) return fail(ctx, LH_ERR_NOT_READY, "lh_frame_begin has not been called");
    return LH_OK;
}

int alloc_targets(LhContext* ctx, uint32_t w, uint32_t h) {
    cudaFree(ctx->vis);
    cudaFree(ctx->scene);
    cudaFree(ctx->display);
    cudaFree(ctx->depth_plane);
    cudaFree(ctx->resolve_tmp);
    ctx->resolve_tmp = nullptr;
    ctx->vis = nullptr;
    ctx->scene = ctx->display = nullptr;
    ctx->depth_plane = nullptr;
    uint64_t px = (uint64_t)w * h;
    if (px == 0) return fail(ctx, LH_ERR_INVALID_ARGUMENT, "zero-sized render target");
    // + one byte per 64x16-pixel tile behind the keys: "this rank drew something here" (sort-last composite, comm.cu)
    const size_t flag_bytes = (size_t)((w + 63) / 64) * ((h + 15) / 16);
    LH_CUDA(ctx, cudaMalloc(&ctx->vis, px * 8 + 16 + flag_bytes));
    LH_CUDA(ctx, cudaMalloc(&ctx->scene, px * 8));
    LH_CUDA(ctx, cudaMalloc(&ctx->display, px * 4));
    LH_CUDA(ctx, cudaMalloc(&ctx->depth_plane, px * 4));
    if (ctx->scratch[0].pass) {   // the main view's hierarchical-Z follows the target size
        auto tiles = [](size_t tw, size_t th) { return ((tw + 3) / 4) * ((th + 3) / 4) + ((tw + 15) / 16) * ((th + 15) / 16); };
        const size_t main_tiles = tiles(w, h), shadow_tiles = tiles(ctx->cfg.shadowmap_size, ctx->cfg.shadowmap_size);
        cudaFree(ctx->scratch[0].hiz);
        ctx->scratch[0].hiz = nullptr;
        LH_CUDA(ctx, cudaMalloc(&ctx->scratch[0].hiz, (main_tiles > shadow_tiles ? main_tiles : shadow_tiles) * 4));
    }
    LH_CUDA(ctx, cudaMemset(ctx->vis, 0xFF, px * 8 + 16 + flag_bytes));
    LH_CUDA(ctx, cudaMemset(ctx->scene, 0, px * 8));
    LH_CUDA(ctx, cudaMemset(ctx->display, 0, px * 4));
    ctx->cfg.width = w;
    ctx->cfg.height = h;
    return LH_OK;
}

// scratch of one view; sets 1..3 (concurrent cascades) also get internal draw-call counters and a reference list
int alloc_pass_scratch(LhContext* ctx, int set) {
    LhContext::PassScratch& sc = ctx->scratch[set];
    if (sc.pass) return LH_OK;
    const size_t draws = (size_t)ctx->cfg.max_opaque_draws + ctx->cfg.max_alpha_clip_draws;
    // candidate meshlets <= max_instances * 65535 (u16 meshlets per mesh) and < 2^32 (lo32 of the prefix counter)
    const unsigned long long cand = (unsigned long long)ctx->cfg.max_instances * 65535ull;
    ctx->meshlet_tile_count = (size_t)((cand < (1ull << 32) ? cand : (1ull << 32)) / kMeshletTile + 2);
    LH_CUDA(ctx, cudaMalloc(&sc.pass, sizeof(PassState)));
    LH_CUDA(ctx, cudaMemset(sc.pass, 0, sizeof(PassState)));
    LH_CUDA(ctx, cudaMalloc(&sc.meshlet_tiles, ctx->meshlet_tile_count * 8));
    LH_CUDA(ctx, cudaMemset(sc.meshlet_tiles, 0, ctx->meshlet_tile_count * 8));
    LH_CUDA(ctx, cudaMalloc(&sc.large_queue, (size_t)(ctx->cfg.max_large_triangles ? ctx->cfg.max_large_triangles : 1) * 8 * 2));   // deferred + split queue
    LH_CUDA(ctx, cudaMalloc(&sc.meshlet_bounds, (draws ? draws : 1) * 16));
    LH_CUDA(ctx, cudaMalloc(&sc.raster_records, (draws ? draws : 1) * 64));
    LH_CUDA(ctx, cudaMalloc(&sc.depth_keys, (draws ? draws : 1) * 4));
    LH_CUDA(ctx, cudaMalloc(&sc.depth_hist, 8192 * 4));
    LH_CUDA(ctx, cudaMemset(sc.depth_hist, 0, 8192 * 4));
    // two hierarchical-Z levels (4x4 and 16x16 pixel tiles): the main view only ever uses set 0
    auto tiles = [](size_t tw, size_t th) { return ((tw + 3) / 4) * ((th + 3) / 4) + ((tw + 15) / 16) * ((th + 15) / 16); };
    size_t hiz_tiles = tiles(ctx->cfg.shadowmap_size, ctx->cfg.shadowmap_size);
    if (set == 0 && tiles(ctx->cfg.width, ctx->cfg.height) > hiz_tiles) hiz_tiles = tiles(ctx->cfg.width, ctx->cfg.height);
    LH_CUDA(ctx, cudaMalloc(&sc.hiz, hiz_tiles * 4));
    if (set != 0) {
        LH_CUDA(ctx, cudaMalloc(&sc.draw_calls, 16));   // counters only: commands are not materialised for internal lists
        LH_CUDA(ctx, cudaMemset(sc.draw_calls, 0, 16));
        LH_CUDA(ctx, cudaMalloc(&sc.refs, (draws ? draws : 1) * sizeof(LhMeshletReference)));
        if (!ctx->side_streams[set - 1]) {
            int prio_least = 0, prio_greatest = 0;
            cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
            LH_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->side_streams[set - 1], cudaStreamNonBlocking, prio_greatest));
            LH_CUDA(ctx, cudaEventCreateWithFlags(&ctx->join_events[set - 1], cudaEventDisableTiming));
        }
        if (!ctx->fork_event) LH_CUDA(ctx, cudaEventCreateWithFlags(&ctx->fork_event, cudaEventDisableTiming));
    }
    return LH_OK;
}

void free_pass_scratch(LhContext* ctx) {
    for (auto& sc : ctx->scratch) {
        cudaFree(sc.pass); cudaFree(sc.meshlet_tiles); cudaFree(sc.large_queue); cudaFree(sc.meshlet_bounds);
        cudaFree(sc.raster_records); cudaFree(sc.depth_keys); cudaFree(sc.depth_hist); cudaFree(sc.hiz); cudaFree(sc.draw_calls); cudaFree(sc.refs);
        sc = LhContext::PassScratch{};
    }
}

int ensure_instance_tiles(LhContext* ctx, uint32_t num_instances) {
    uint32_t tiles = (num_instances + kInstanceTile - 1) / kInstanceTile + 1;
    if (tiles <= ctx->instance_tiles_per_sector) return LH_OK;
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->instance_tiles);
    ctx->instance_tiles = nullptr;
    LH_CUDA(ctx, cudaMalloc(&ctx->instance_tiles, (size_t)tiles * 4 * 8));
    LH_CUDA(ctx, cudaMemset(ctx->instance_tiles, 0, (size_t)tiles * 4 * 8));
    ctx->instance_tiles_per_sector = tiles;
    return LH_OK;
}

}  // namespace

extern "C" {

int lh_config_default(LhConfig* cfg) {
    if (!cfg) return LH_ERR_INVALID_ARGUMENT;
    memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = sizeof(LhConfig);
    cfg->device = 0;
    cfg->max_instances = LH_MAX_INSTANCES;
    cfg->max_opaque_draws = LH_MAX_OPAQUE_DRAWS;
    cfg->max_alpha_clip_draws = LH_MAX_ALPHA_CLIP_DRAWS;
    cfg->width = 640;   // src/main.cpp:103-106
    cfg->height = 480;
    cfg->shadowmap_size = LH_SHADOWMAP_SIZE;
    cfg->flags = LH_CONFIG_WRITE_DRAW_COMMANDS;
    cfg->max_large_triangles = 1u << 22;
    return LH_OK;
}

int lh_buffer_sizes(const LhConfig* cfg, LhBufferSizes* out) {
    if (!cfg || !out) return LH_ERR_INVALID_ARGUMENT;
    uint64_t draws = (uint64_t)cfg->max_opaque_draws + cfg->max_alpha_clip_draws;
    out->meshlet_references = 8 * draws * 2;
    out->num_meshlets_prefix_sum = ((uint64_t)cfg->max_instances * 8 + 8) * 4;
    out->draw_calls = 8 + 16 * draws;
    out->dispatches = 12 * 3;
    out->misc_storage = sizeof(LhMiscStorage);
    out->uniforms = sizeof(LhUniforms);
    return LH_OK;
}

const char* lh_version(void) { return "lighthugger_b200 0.1 (sm_100a)"; }

const char* lh_last_error(const LhContext* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int lh_context_create(const LhConfig* cfg, LhContext** out) {
    if (!cfg || !out || cfg->struct_size != sizeof(LhConfig)) return fail(nullptr, LH_ERR_INVALID_ARGUMENT, "bad LhConfig");
    if ((uint64_t)cfg->max_opaque_draws + cfg->max_alpha_clip_draws > LH_VIS_MAX_SLOTS)
        return fail(nullptr, LH_ERR_INVALID_ARGUMENT, "max_opaque_draws + max_alpha_clip_draws exceeds the 25-bit slot field");
    if (cfg->max_instances == 0 || cfg->shadowmap_size == 0) return fail(nullptr, LH_ERR_INVALID_ARGUMENT, "zero capacity");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return fail(nullptr, LH_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)", e);
    if (cfg->device < 0 || cfg->device >= count) return fail(nullptr, LH_ERR_NO_DEVICE, "device ordinal out of range");
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return fail(nullptr, LH_ERR_CUDA, "cudaSetDevice", e);
    LhContext* ctx = new (std::nothrow) LhContext();
    if (!ctx) return fail(nullptr, LH_ERR_OUT_OF_MEMORY, "host allocation");
    ctx->device = cfg->device;
    ctx->cfg = *cfg;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, cfg->device);
    ctx->num_sms = prop.multiProcessorCount;
    if (const char* t = getenv("LH_TUNE_FRACTIONS")) {
        float* f = ctx->tune.main_fractions;
        sscanf(t, "%f,%f,%f,%f,%f,%f,%f", &f[0], &f[1], &f[2], &f[3], &f[4], &f[5], &f[6]);
    }
    if (const char* t = getenv("LH_TUNE")) {
        unsigned v[8] = {ctx->tune.main_bands, ctx->tune.shadow_bands, ctx->tune.main_grid, ctx->tune.shadow_grid, ctx->tune.resolve_low_priority, ctx->tune.select_grid, ctx->tune.resolve_split, ctx->tune.small_max_area};
        sscanf(t, "%u,%u,%u,%u,%u,%u,%u,%u", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]);
        ctx->tune.small_max_area = v[7] < 1 ? 1 : (v[7] > 1024 ? 1024 : v[7]);
        ctx->tune.resolve_low_priority = v[4] != 0;
        ctx->tune.resolve_split = v[6] != 0;
        auto clamp = [](unsigned x, unsigned lo, unsigned hi) { return x < lo ? lo : (x > hi ? hi : x); };
        ctx->tune.from_environment = true;
        ctx->tune.main_bands = clamp(v[0], 2, 8);
        ctx->tune.shadow_bands = clamp(v[1], 2, 8);
        ctx->tune.main_grid = clamp(v[2], 1, 64);
        ctx->tune.shadow_grid = clamp(v[3], 1, 64);
        ctx->tune.select_grid = clamp(v[5], 1, 64);
    }
    auto bail = [&](int rc) {
        g_create_error = ctx->error;
        lh_context_destroy(ctx);
        return rc;
    };
#define TRY(expr)                                                              \
    do {                                                                       \
        cudaError_t _e = (expr);                                               \
        if (_e != cudaSuccess) return bail(fail(ctx, LH_ERR_CUDA, #expr, _e)); \
    } while (0)
    // the context stream and the cascades' side streams carry the frame's critical path; the material phase of the
    // resolve, which only fills the gaps beside the shadow passes, runs on a stream of the lowest priority
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    TRY(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_greatest));
    TRY(cudaStreamCreateWithPriority(&ctx->deferred_stream, cudaStreamNonBlocking, prio_greatest));
    for (int i = 0; i < 8; i++) TRY(cudaEventCreateWithFlags(&ctx->band_done[i], cudaEventDisableTiming));
    TRY(cudaEventCreateWithFlags(&ctx->drains_done, cudaEventDisableTiming));
    TRY(cudaStreamCreateWithPriority(&ctx->resolve_stream, cudaStreamNonBlocking, ctx->tune.resolve_low_priority ? prio_least : prio_greatest));
    TRY(cudaEventCreateWithFlags(&ctx->vis_ready, cudaEventDisableTiming));
    TRY(cudaEventCreateWithFlags(&ctx->material_done, cudaEventDisableTiming));
    int rc = alloc_targets(ctx, cfg->width, cfg->height);
    if (rc) return bail(rc);
    size_t shadow_bytes = (size_t)4 * cfg->shadowmap_size * cfg->shadowmap_size * 4;
    TRY(cudaMalloc(&ctx->shadow, shadow_bytes));
    TRY(cudaMemset(ctx->shadow, 0, shadow_bytes));
    TRY(cudaMalloc(&ctx->state, sizeof(FrameState)));
    TRY(cudaMemset(ctx->state, 0, sizeof(FrameState)));
    rc = alloc_pass_scratch(ctx, 0);
    if (rc) return bail(rc);
    TRY(cudaMalloc(&ctx->tex_dev, sizeof(cudaTextureObject_t) * LH_MAX_TEXTURES));
    TRY(cudaMemset(ctx->tex_dev, 0, sizeof(cudaTextureObject_t) * LH_MAX_TEXTURES));
    TRY(cudaMalloc(&ctx->tex_dims_dev, sizeof(uint4) * LH_MAX_TEXTURES));
    TRY(cudaMemset(ctx->tex_dims_dev, 0, sizeof(uint4) * LH_MAX_TEXTURES));
    TRY(cudaMallocHost(&ctx->pinned_stats, sizeof(LhFrameStats)));
    for (int i = 0; i < LH_PASS_COUNT; i++) {
        TRY(cudaEventCreate(&ctx->pass_start[i]));
        TRY(cudaEventCreate(&ctx->pass_stop[i]));
    }
    TRY(cudaEventCreate(&ctx->timer_a));
    TRY(cudaEventCreate(&ctx->timer_b));
    for (int i = 0; i < 8; i++) {
        TRY(cudaEventCreate(&ctx->raster_band_a[i]));
        TRY(cudaEventCreate(&ctx->raster_band_b[i]));
    }
    rc = ensure_instance_tiles(ctx, cfg->max_instances);
    if (rc) return bail(rc);
#undef TRY
    *out = ctx;
    return LH_OK;
}

int lh_context_destroy(LhContext* ctx) {
    if (!ctx) return LH_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    lh_comm_destroy(ctx);
    for (uint32_t i = 0; i < LH_MAX_TEXTURES; i++) {
        if (ctx->tex_host[i]) cudaDestroyTextureObject(ctx->tex_host[i]);
        if (ctx->tex_arrays[i]) cudaFreeMipmappedArray(ctx->tex_arrays[i]);
    }
    if (ctx->sky_tex) cudaDestroyTextureObject(ctx->sky_tex);
    if (ctx->sky_array) cudaFreeArray(ctx->sky_array);
    if (ctx->lut_tex) cudaDestroyTextureObject(ctx->lut_tex);
    if (ctx->lut_array) cudaFreeArray(ctx->lut_array);
    for (void* b : ctx->buffers) cudaFree(b);
    cudaFree(ctx->vis);
    cudaFree(ctx->scene);
    cudaFree(ctx->display);
    cudaFree(ctx->depth_plane);
    cudaFree(ctx->resolve_tmp);
    if (ctx->resolve_stream) cudaStreamDestroy(ctx->resolve_stream);
    if (ctx->deferred_stream) cudaStreamDestroy(ctx->deferred_stream);
    for (int i = 0; i < 8; i++)
        if (ctx->band_done[i]) cudaEventDestroy(ctx->band_done[i]);
    if (ctx->drains_done) cudaEventDestroy(ctx->drains_done);
    if (ctx->vis_ready) cudaEventDestroy(ctx->vis_ready);
    if (ctx->material_done) cudaEventDestroy(ctx->material_done);
    cudaFree(ctx->shadow);
    cudaFree(ctx->state);
    cudaFree(ctx->instance_tiles);
    free_pass_scratch(ctx);
    for (int i = 0; i < 3; i++) {
        if (ctx->side_streams[i]) cudaStreamDestroy(ctx->side_streams[i]);
        if (ctx->join_events[i]) cudaEventDestroy(ctx->join_events[i]);
    }
    if (ctx->fork_event) cudaEventDestroy(ctx->fork_event);
    cudaFree(ctx->tex_dev);
    cudaFree(ctx->tex_dims_dev);
    cudaFree(ctx->l2_flush);
    cudaFree(ctx->gathered_refs);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->pinned_stats) cudaFreeHost(ctx->pinned_stats);
    for (int i = 0; i < LH_PASS_COUNT; i++) {
        if (ctx->pass_start[i]) cudaEventDestroy(ctx->pass_start[i]);
        if (ctx->pass_stop[i]) cudaEventDestroy(ctx->pass_stop[i]);
    }
    if (ctx->timer_a) cudaEventDestroy(ctx->timer_a);
    if (ctx->timer_b) cudaEventDestroy(ctx->timer_b);
    for (int i = 0; i < 8; i++) {
        if (ctx->raster_band_a[i]) cudaEventDestroy(ctx->raster_band_a[i]);
        if (ctx->raster_band_b[i]) cudaEventDestroy(ctx->raster_band_b[i]);
    }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return LH_OK;
}

// ---- buffers ----
int lh_buffer_alloc(LhContext* ctx, uint64_t bytes, uint64_t* dev_addr) {
    if (!ctx || !dev_addr) return LH_ERR_INVALID_ARGUMENT;
    void* p = nullptr;
    LH_CUDA(ctx, cudaMalloc(&p, bytes ? bytes : 16));
    LH_CUDA(ctx, cudaMemsetAsync(p, 0, bytes ? bytes : 16, ctx->stream));
    ctx->buffers.push_back(p);
    *dev_addr = (uint64_t)(uintptr_t)p;
    return LH_OK;
}

int lh_buffer_free(LhContext* ctx, uint64_t dev_addr) {
    if (!ctx) return LH_ERR_INVALID_ARGUMENT;
    void* p = (void*)(uintptr_t)dev_addr;
    for (size_t i = 0; i < ctx->buffers.size(); i++)
        if (ctx->buffers[i] == p) {
            cudaStreamSynchronize(ctx->stream);
            cudaFree(p);
            ctx->buffers[i] = ctx->buffers.back();
            ctx->buffers.pop_back();
            return LH_OK;
        }
    return fail(ctx, LH_ERR_INVALID_ARGUMENT, "lh_buffer_free: unknown address");
}

int lh_upload(LhContext* ctx, uint64_t dev_addr, const void* src, uint64_t bytes) {
    if (!ctx || (!src && bytes)) return LH_ERR_INVALID_ARGUMENT;
    if (!bytes) return LH_OK;
    LH_CUDA(ctx, cudaMemcpyAsync((void*)(uintptr_t)dev_addr, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    LH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // no host pointer is retained past the call
    return LH_OK;
}

int lh_download(LhContext* ctx, uint64_t dev_addr, void* dst, uint64_t bytes) {
    if (!ctx || (!dst && bytes)) return LH_ERR_INVALID_ARGUMENT;
    if (!bytes) return LH_OK;
    LH_CUDA(ctx, cudaMemcpyAsync(dst, (const void*)(uintptr_t)dev_addr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    LH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LH_OK;
}

// ---- textures ----
int lh_texture_create(LhContext* ctx, int format, uint32_t width, uint32_t height, uint32_t mip_levels, const void* data,
                      uint64_t bytes, uint16_t* slot_out) {
    if (!ctx || !data || !slot_out || !width || !height || !mip_levels) return LH_ERR_INVALID_ARGUMENT;
    if (width > 65536u || height > 65536u || mip_levels > 17u) return fail(ctx, LH_ERR_INVALID_ARGUMENT, "texture dimensions out of range");
    uint32_t slot = 0;
    while (slot < LH_MAX_TEXTURES && ctx->tex_host[slot]) slot++;
    if (slot == LH_MAX_TEXTURES) return fail(ctx, LH_ERR_CAPACITY, "all 512 bindless texture slots are in use");
    cudaChannelFormatDesc desc;
    bool block = false, srgb = false;
    switch (format) {
    case LH_TEXTURE_RGBA8_SRGB: desc = cudaCreateChannelDesc<uchar4>(); srgb = true; break;
    case LH_TEXTURE_RGBA8_UNORM: desc = cudaCreateChannelDesc<uchar4>(); break;
    case LH_TEXTURE_BC7_SRGB: desc = cudaCreateChannelDesc<cudaChannelFormatKindUnsignedBlockCompressed7SRGB>(); block = true; srgb = true; break;
    case LH_TEXTURE_BC1_SRGB: desc = cudaCreateChannelDesc<cudaChannelFormatKindUnsignedBlockCompressed1SRGB>(); block = true; srgb = true; break;
    default: return fail(ctx, LH_ERR_UNSUPPORTED, "unknown texture format");
    }
    const uint32_t block_bytes = format == LH_TEXTURE_BC1_SRGB ? 8 : 16;
    uint64_t need = 0;
    for (uint32_t l = 0; l < mip_levels; l++) {
        uint32_t w = width >> l ? width >> l : 1, h = height >> l ? height >> l : 1;
        need += block ? (uint64_t)((w + 3) / 4) * ((h + 3) / 4) * block_bytes : (uint64_t)w * h * 4;
    }
    if (bytes < need) return fail(ctx, LH_ERR_INVALID_ARGUMENT, "texture data is smaller than its mip chain");
    cudaMipmappedArray_t arr = nullptr;
    LH_CUDA(ctx, cudaMallocMipmappedArray(&arr, &desc, make_cudaExtent(width, height, 0), mip_levels));
    const uint8_t* src = (const uint8_t*)data;
    for (uint32_t l = 0; l < mip_levels; l++) {
        uint32_t w = width >> l ? width >> l : 1, h = height >> l ? height >> l : 1;
        cudaArray_t level;
        cudaError_t e = cudaGetMipmappedArrayLevel(&level, arr, l);
        size_t row_bytes = block ? (size_t)((w + 3) / 4) * block_bytes : (size_t)w * 4;
        size_t rows = block ? (h + 3) / 4 : h;
        if (e == cudaSuccess) e = cudaMemcpy2DToArray(level, 0, 0, src, row_bytes, row_bytes, rows, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFreeMipmappedArray(arr);
            return fail(ctx, LH_ERR_CUDA, "texture level upload", e);
        }
        src += row_bytes * rows;
    }
    // repeat_sampler (src/main.cpp:550-560): linear mag/min/mip, repeat, anisotropy 16.  The texture unit does the
    // block decode, the sRGB conversion and the bilinear taps; mip selection, the trilinear blend and the anisotropic
    // footprint (Vulkan's model: N = min(ceil(Pmax/Pmin), 16) probes along the major axis at lod log2(Pmax/N)) are
    // computed in the resolve kernel, because the hardware's own footprint is implementation-defined and the lit
    // colour is parity-checked against that model.  Hence a point mip filter here: every fetch names an integer level.
    cudaResourceDesc res{};
    res.resType = cudaResourceTypeMipmappedArray;
    res.res.mipmap.mipmap = arr;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeWrap;
    td.filterMode = cudaFilterModeLinear;
    td.mipmapFilterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    td.maxAnisotropy = 1;
    td.minMipmapLevelClamp = 0.0f;
    td.maxMipmapLevelClamp = (float)(mip_levels - 1);
    td.sRGB = srgb ? 1 : 0;  // the driver rejects a BC*_SRGB array unless the sRGB read flag is set too (probed on B200)
    cudaTextureObject_t tex = 0;
    cudaError_t e = cudaCreateTextureObject(&tex, &res, &td, nullptr);
    if (e != cudaSuccess) {
        cudaFreeMipmappedArray(arr);
        return fail(ctx, LH_ERR_CUDA, "cudaCreateTextureObject", e);
    }
    ctx->tex_host[slot] = tex;
    ctx->tex_arrays[slot] = arr;
    LH_CUDA(ctx, cudaMemcpyAsync(ctx->tex_dev + slot, &ctx->tex_host[slot], sizeof(tex), cudaMemcpyHostToDevice, ctx->stream));
    const uint4 dims = make_uint4(width, height, mip_levels, 0u);
    LH_CUDA(ctx, cudaMemcpyAsync(ctx->tex_dims_dev + slot, &dims, sizeof(dims), cudaMemcpyHostToDevice, ctx->stream));
    LH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *slot_out = (uint16_t)slot;
    return LH_OK;
}

int lh_texture_destroy(LhContext* ctx, uint16_t slot) {
    if (!ctx || slot >= LH_MAX_TEXTURES || !ctx->tex_host[slot]) return LH_ERR_INVALID_ARGUMENT;
    cudaStreamSynchronize(ctx->stream);
    cudaDestroyTextureObject(ctx->tex_host[slot]);
    cudaFreeMipmappedArray(ctx->tex_arrays[slot]);
    ctx->tex_host[slot] = 0;
    ctx->tex_arrays[slot] = nullptr;
    LH_CUDA(ctx, cudaMemsetAsync(ctx->tex_dev + slot, 0, sizeof(cudaTextureObject_t), ctx->stream));
    return LH_OK;
}

int lh_skybox_create(LhContext* ctx, uint32_t size, const float* faces) {
    if (!ctx || !size || !faces) return LH_ERR_INVALID_ARGUMENT;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->sky_tex) cudaDestroyTextureObject(ctx->sky_tex);
    if (ctx->sky_array) cudaFreeArray(ctx->sky_array);
    ctx->sky_tex = 0;
    ctx->sky_array = nullptr;
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<float4>();
    LH_CUDA(ctx, cudaMalloc3DArray(&ctx->sky_array, &desc, make_cudaExtent(size, size, 6), cudaArrayCubemap));
    cudaMemcpy3DParms cp{};
    cp.srcPtr = make_cudaPitchedPtr((void*)faces, (size_t)size * 16, size, size);
    cp.dstArray = ctx->sky_array;
    cp.extent = make_cudaExtent(size, size, 6);
    cp.kind = cudaMemcpyHostToDevice;
    LH_CUDA(ctx, cudaMemcpy3D(&cp));
    cudaResourceDesc res{};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = ctx->sky_array;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 1;
    LH_CUDA(ctx, cudaCreateTextureObject(&ctx->sky_tex, &res, &td, nullptr));
    return LH_OK;
}

int lh_set_display_lut(LhContext* ctx, uint32_t size, const float* rgb) {
    if (!ctx || !size || !rgb) return LH_ERR_INVALID_ARGUMENT;
    cudaStreamSynchronize(ctx->stream);
    if (ctx->lut_tex) cudaDestroyTextureObject(ctx->lut_tex);
    if (ctx->lut_array) cudaFreeArray(ctx->lut_array);
    ctx->lut_tex = 0;
    ctx->lut_array = nullptr;
    std::vector<float> rgba((size_t)size * size * size * 4);
    for (size_t i = 0; i < (size_t)size * size * size; i++) {
        rgba[4 * i] = rgb[3 * i];
        rgba[4 * i + 1] = rgb[3 * i + 1];
        rgba[4 * i + 2] = rgb[3 * i + 2];
        rgba[4 * i + 3] = 1.0f;
    }
    cudaChannelFormatDesc desc = cudaCreateChannelDesc<float4>();
    LH_CUDA(ctx, cudaMalloc3DArray(&ctx->lut_array, &desc, make_cudaExtent(size, size, size), 0));
    cudaMemcpy3DParms cp{};
    cp.srcPtr = make_cudaPitchedPtr(rgba.data(), (size_t)size * 16, size, size);
    cp.dstArray = ctx->lut_array;
    cp.extent = make_cudaExtent(size, size, size);
    cp.kind = cudaMemcpyHostToDevice;
    LH_CUDA(ctx, cudaMemcpy3D(&cp));
    cudaResourceDesc res{};
    res.resType = cudaResourceTypeArray;
    res.res.array.array = ctx->lut_array;
    cudaTextureDesc td{};  // clamp_sampler (src/main.cpp:561-568)
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 1;
    LH_CUDA(ctx, cudaCreateTextureObject(&ctx->lut_tex, &res, &td, nullptr));
    return LH_OK;
}

int lh_targets_resize(LhContext* ctx, uint32_t width, uint32_t height) {
    if (!ctx) return LH_ERR_INVALID_ARGUMENT;
    cudaStreamSynchronize(ctx->stream);
    // sort-last ranks resize together: unmap -> barrier -> reallocate, and the row slices follow the new height
    if (int rc = peer_release_collective(ctx)) return rc;
    if (int rc = alloc_targets(ctx, width, height)) return rc;
    comm_update_rows(ctx);
    return LH_OK;
}

int lh_set_instance_range(LhContext* ctx, uint32_t first, uint32_t count) {
    if (!ctx) return LH_ERR_INVALID_ARGUMENT;
    ctx->instance_first = first;
    ctx->instance_count = count;
    return LH_OK;
}

int lh_set_shadow_instance_range(LhContext* ctx, uint32_t first, uint32_t count) {
    if (!ctx) return LH_ERR_INVALID_ARGUMENT;
    ctx->shadow_range_set = true;
    ctx->shadow_instance_first = first;
    ctx->shadow_instance_count = count;
    return LH_OK;
}

int lh_set_instance_stripes(LhContext* ctx, int view, uint32_t stripes, uint32_t index, uint32_t count) {
    if (!ctx || view < 0 || view > 1 || stripes > kMaxStripes || (stripes && (count == 0 || index >= count))) return LH_ERR_INVALID_ARGUMENT;
    ctx->stripes[view] = stripes;
    ctx->stripe_index[view] = index;
    ctx->stripe_count[view] = count ? count : 1;
    for (uint32_t s = 0; s <= kMaxStripes; s++) ctx->stripe_bounds[view][s] = stripes ? (s < stripes ? (float)s / (float)stripes : 1.0f) : 0.0f;
    return LH_OK;
}

int lh_set_instance_stripe_bounds(LhContext* ctx, int view, const float* bounds, uint32_t stripes, uint32_t index, uint32_t count) {
    if (!ctx || !bounds || view < 0 || view > 1 || stripes == 0 || stripes > kMaxStripes || count == 0 || index >= count) return LH_ERR_INVALID_ARGUMENT;
    for (uint32_t s = 0; s < stripes; s++)
        if (!(bounds[s] <= bounds[s + 1])) return fail(ctx, LH_ERR_INVALID_ARGUMENT, "stripe bounds must ascend");
    ctx->stripes[view] = stripes;
    ctx->stripe_index[view] = index;
    ctx->stripe_count[view] = count;
    // the outer bounds are open: every centre belongs to some stripe
    for (uint32_t s = 0; s <= kMaxStripes; s++) ctx->stripe_bounds[view][s] = s == 0 ? 0.0f : (s < stripes ? bounds[s] : 1.0f);
    return LH_OK;
}

int lh_set_shadow_cascade_mask(LhContext* ctx, uint32_t mask) {
    if (!ctx) return LH_ERR_INVALID_ARGUMENT;
    ctx->cascade_mask = mask & 0xFu;
    return LH_OK;
}

// ---- per frame ----
int lh_frame_begin(LhContext* ctx, const LhUniforms* u) {
    if (!ctx || !u) return LH_ERR_INVALID_ARGUMENT;
    if (u->window_size[0] != ctx->cfg.width || u->window_size[1] != ctx->cfg.height)
        return fail(ctx, LH_ERR_INVALID_ARGUMENT, "uniforms.window_size differs from the render targets (call lh_targets_resize)");
    if (!u->instances || !u->meshlet_references || !u->draw_calls || !u->misc_storage || !u->num_meshlets_prefix_sum || !u->dispatches)
        return fail(ctx, LH_ERR_INVALID_ARGUMENT, "uniforms carry a null buffer address");
    ctx->uniforms = *u;
    ctx->have_uniforms = true;
    for (int i = 0; i < LH_PASS_COUNT; i++) ctx->pass_recorded[i] = false;
    ctx->first_pass = ctx->last_pass = -1;
    ctx->raster_bands_recorded = 0;
    if (int rc = ensure_peer_maps(ctx)) return rc;
    return ensure_instance_tiles(ctx, u->num_instances);
}

int lh_reset(LhContext* ctx) {
    if (int rc = require_frame(ctx)) return rc;
    PassScope scope(ctx, LH_PASS_RESET);
    launch_reset(ctx, make_params(ctx));
    return check_launch(ctx, "reset");
}

int lh_cull_instances(LhContext* ctx) {
    if (int rc = require_frame(ctx)) return rc;
    PassScope scope(ctx, LH_PASS_CULL_INSTANCES);
    launch_cull_instances(ctx, make_params(ctx), false);
    return check_launch(ctx, "cull_instances");
}

int lh_cull_meshlets(LhContext* ctx) {
    if (int rc = require_frame(ctx)) return rc;
    PassScope scope(ctx, LH_PASS_CULL_MESHLETS);
    launch_cull_meshlets(ctx, make_params(ctx), -1);
    return check_launch(ctx, "cull_meshlets");
}

int lh_raster_visbuffer(LhContext* ctx) {
    if (int rc = require_frame(ctx)) return rc;
    PassScope scope(ctx, LH_PASS_VISBUFFER);
    // render-pass begin: depth clear 0.0 (rendering.cpp:206-211) == all-ones key
    launch_fill_u64(ctx, ctx->vis, (uint64_t)ctx->cfg.width * ctx->cfg.height, LH_VIS_CLEAR);
    launch_raster(ctx, make_params(ctx), -1);
    return check_launch(ctx, "raster_visbuffer");
}

int lh_composite(LhContext* ctx) {
    if (int rc = require_frame(ctx)) return rc;
    if (!ctx->nccl_comm || ctx->nranks <= 1) return LH_OK;
    PassScope scope(ctx, LH_PASS_COMPOSITE);
    if (peer_exchange(ctx)) {
        if (int rc = composite_pull(ctx)) return rc;
        return check_launch(ctx, "composite");
    }
    return composite_allreduce_min(ctx);
}

int lh_reduce_depth(LhContext* ctx) {
    if (int rc = require_frame(ctx)) return rc;
    PassScope scope(ctx, LH_PASS_DEPTH_REDUCTION);
    if (peer_exchange(ctx)) {
        // the composite kernel already reduced this rank's rows; what is left is the 8-byte exchange
        if (int rc = reduce_depth_exchange(ctx)) return rc;
    } else {
        launch_reduce_depth(ctx, make_params(ctx));
    }
    return check_launch(ctx, "reduce_depth");
}

int lh_generate_shadow_matrices(LhContext* ctx) {
    if (int rc = require_frame(ctx)) return rc;
    PassScope scope(ctx, LH_PASS_SHADOW_MATRICES);
    launch_shadow_matrices(ctx, make_params(ctx));
    return check_launch(ctx, "generate_shadow_matrices");
}

int lh_cull_instances_shadows(LhContext* ctx) {
    if (int rc = require_frame(ctx)) return rc;
    PassScope scope(ctx, LH_PASS_CULL_INSTANCES_SHADOWS);
    launch_cull_instances(ctx, make_params(ctx), true);
    return check_launch(ctx, "cull_instances_shadows");
}

// reset_buffers_c + write_draw_calls_shadows + the layer clear + the depth-only raster of one cascade
static void enqueue_cascade(LhContext* ctx, uint32_t cascade, int set, bool internal_lists, cudaStream_t stream) {
    ZoneScope zone("shadowmap inner");
    FrameParams p = make_params(ctx, (int)cascade, set, internal_lists);
    if (internal_lists) p.flags &= ~(uint32_t)LH_CONFIG_WRITE_DRAW_COMMANDS;
    launch_cull_meshlets(ctx, p, (int)cascade, stream);
    size_t texels = (size_t)ctx->cfg.shadowmap_size * ctx->cfg.shadowmap_size;
    launch_fill_u32(ctx, ctx->shadow + texels * cascade, texels, 0x3F800000u, stream);  // clear 1.0, rendering.cpp:384
    launch_raster(ctx, p, (int)cascade, stream);
}

int lh_shadow_cascade(LhContext* ctx, uint32_t cascade) {
    if (int rc = require_frame(ctx)) return rc;
    if (cascade >= 4) return fail(ctx, LH_ERR_INVALID_ARGUMENT, "cascade index out of range");
    enqueue_cascade(ctx, cascade, 0, false, nullptr);
    return check_launch(ctx, "shadow_cascade");
}

int lh_resolve(LhContext* ctx) {
    if (int rc = require_frame(ctx)) return rc;
    PassScope scope(ctx, LH_PASS_RENDER_GEOMETRY);
    launch_resolve(ctx, make_params(ctx));
    return check_launch(ctx, "resolve");
}

int lh_display_transform(LhContext* ctx) {
    if (int rc = require_frame(ctx)) return rc;
    PassScope scope(ctx, LH_PASS_DISPLAY_TRANSFORM);
    launch_display_transform(ctx, make_params(ctx));
    return check_launch(ctx, "display_transform");
}

int lh_render_frame(LhContext* ctx) {
    ZoneScope zone("render");
    int rc;
    if ((rc = lh_reset(ctx))) return rc;
    if ((rc = lh_cull_instances(ctx))) return rc;
    if ((rc = lh_cull_meshlets(ctx))) return rc;
    if ((rc = lh_raster_visbuffer(ctx))) return rc;
    if ((rc = lh_composite(ctx))) return rc;
    // The resolve in two launches: its material phase needs the visibility buffer only and runs beside the depth and
    // shadow passes; its shadow phase follows them.
    const bool split = ctx->tune.resolve_split;
    if (split) {
        if (!ctx->resolve_tmp) LH_CUDA(ctx, cudaMalloc(&ctx->resolve_tmp, (size_t)ctx->cfg.width * ctx->cfg.height * 40));
        LH_CUDA(ctx, cudaEventRecord(ctx->vis_ready, ctx->stream));
        LH_CUDA(ctx, cudaStreamWaitEvent(ctx->resolve_stream, ctx->vis_ready, 0));
        launch_resolve(ctx, make_params(ctx), 1, ctx->resolve_stream);
        LH_CUDA(ctx, cudaEventRecord(ctx->material_done, ctx->resolve_stream));
    }
    if ((rc = lh_reduce_depth(ctx))) return rc;
    if ((rc = lh_generate_shadow_matrices(ctx))) return rc;
    if ((rc = lh_cull_instances_shadows(ctx))) return rc;
    {
        // The cascades are independent of one another (own layer, own prefix-sum sector).  The last one of the mask runs
        // on the context stream with the host-visible lists, which is the state the reference leaves behind
        // (rendering.cpp:325-438 reuses one scratch section); the others run concurrently on side streams with internal
        // lists, so that a small cascade's latency-bound launches overlap a large cascade's.
        PassScope scope(ctx, LH_PASS_SHADOWMAPS);
        int last = -1, others[4], n_others = 0;
        for (int c = 3; c >= 0; c--)
            if ((ctx->cascade_mask >> c) & 1u) {
                if (last < 0) last = c;
                else others[n_others++] = c;
            }
        for (int i = 0; i < n_others; i++)
            if ((rc = alloc_pass_scratch(ctx, 1 + i))) return rc;
        if (n_others) LH_CUDA(ctx, cudaEventRecord(ctx->fork_event, ctx->stream));
        // the last cascade is the largest and the critical path: it is enqueued first, so that it does not wait for
        // the host to enqueue the other three
        if (last >= 0) enqueue_cascade(ctx, (uint32_t)last, 0, false, nullptr);
        for (int i = 0; i < n_others; i++) {
            cudaStream_t s = ctx->side_streams[i];
            LH_CUDA(ctx, cudaStreamWaitEvent(s, ctx->fork_event, 0));
            enqueue_cascade(ctx, (uint32_t)others[i], 1 + i, true, s);
            LH_CUDA(ctx, cudaEventRecord(ctx->join_events[i], s));
        }
        for (int i = 0; i < n_others; i++) LH_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->join_events[i], 0));
        if ((rc = check_launch(ctx, "shadow cascades"))) return rc;
    }
    if (ctx->nccl_comm && ctx->nranks > 1) {
        PassScope scope(ctx, LH_PASS_SHADOW_EXCHANGE);
        if ((rc = lh_shadow_exchange(ctx))) return rc;
    }
    {
        PassScope scope(ctx, LH_PASS_RENDER_GEOMETRY);   // what is left of the resolve on the critical path
        if (split) LH_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->material_done, 0));
        launch_resolve(ctx, make_params(ctx), split ? 2 : 0);
        if ((rc = check_launch(ctx, "resolve"))) return rc;
    }
    // rank 0's frame is complete, and every table a peer reads may be rewritten, once all ranks have resolved their rows
    if (peer_exchange(ctx)) {
        PassScope scope(ctx, LH_PASS_FRAME_BARRIER);
        return frame_barrier(ctx);
    }
    return LH_OK;
}

int lh_frame_sync(LhContext* ctx) {
    if (!ctx) return LH_ERR_INVALID_ARGUMENT;
    // the frame's overflow word rides along (4 bytes into pinned memory), so a list that outgrew its LhConfig capacity is
    // an error here too, not only in lh_get_frame_stats
    ctx->pinned_stats->overflow = 0;
    LH_CUDA(ctx, cudaMemcpyAsync(&ctx->pinned_stats->overflow, &ctx->state->overflow, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    LH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (int rc = check_launch(ctx, "frame")) return rc;
    if (int rc = comm_poll_errors(ctx)) return rc;
    if (ctx->have_uniforms && ctx->pinned_stats->overflow)
        return fail(ctx, LH_ERR_CAPACITY, "a list overflowed its LhConfig capacity (see LhFrameStats.overflow)");
    return LH_OK;
}

int lh_get_frame_stats(LhContext* ctx, LhFrameStats* out) {
    if (int rc = require_frame(ctx)) return rc;
    if (!out) return LH_ERR_INVALID_ARGUMENT;
    launch_gather_stats(ctx, make_params(ctx), ctx->pinned_stats);
    LH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (int rc = check_launch(ctx, "gather_stats")) return rc;
    *out = *ctx->pinned_stats;
    if (out->overflow) return fail(ctx, LH_ERR_CAPACITY, "a list overflowed its LhConfig capacity (see LhFrameStats.overflow)");
    return LH_OK;
}

int lh_render_frame_host(LhContext* ctx, const LhUniforms* u, const LhInstance* instances, uint32_t num_instances, LhFrameStats* stats) {
    if (!ctx || !u) return LH_ERR_INVALID_ARGUMENT;
    int rc;
    if ((rc = lh_frame_begin(ctx, u))) return rc;
    if (instances && num_instances) {
        // stage through pinned memory so the copy is a true async DMA
        size_t bytes = (size_t)num_instances * sizeof(LhInstance);
        if (bytes > ctx->pinned_bytes) {
            if (ctx->pinned) cudaFreeHost(ctx->pinned);
            ctx->pinned = nullptr;
            ctx->pinned_bytes = 0;
            LH_CUDA(ctx, cudaMallocHost(&ctx->pinned, bytes));
            ctx->pinned_bytes = bytes;
        }
        memcpy(ctx->pinned, instances, bytes);
        LH_CUDA(ctx, cudaMemcpyAsync((void*)(uintptr_t)u->instances, ctx->pinned, bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    if ((rc = lh_render_frame(ctx))) return rc;
    if (stats) return lh_get_frame_stats(ctx, stats);
    return lh_frame_sync(ctx);
}

int lh_get_target(LhContext* ctx, int target, void* dst, uint64_t bytes) {
    if (!ctx || !dst) return LH_ERR_INVALID_ARGUMENT;
    uint64_t px = (uint64_t)ctx->cfg.width * ctx->cfg.height;
    const void* src = nullptr;
    uint64_t need = 0;
    switch (target) {
    case LH_TARGET_VIS64: src = ctx->vis; need = px * 8; break;
    case LH_TARGET_DEPTH:
        launch_extract_depth(ctx, make_params(ctx));
        src = ctx->depth_plane; need = px * 4; break;
    case LH_TARGET_SCENE_RGBA16F: src = ctx->scene; need = px * 8; break;
    case LH_TARGET_SHADOW: src = ctx->shadow; need = (uint64_t)4 * ctx->cfg.shadowmap_size * ctx->cfg.shadowmap_size * 4; break;
    case LH_TARGET_DISPLAY_BGRA8: src = ctx->display; need = px * 4; break;
    default: return fail(ctx, LH_ERR_INVALID_ARGUMENT, "unknown target");
    }
    if (bytes < need) return fail(ctx, LH_ERR_INVALID_ARGUMENT, "lh_get_target: destination too small");
    LH_CUDA(ctx, cudaMemcpyAsync(dst, src, need, cudaMemcpyDeviceToHost, ctx->stream));
    LH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LH_OK;
}

int lh_get_target_address(LhContext* ctx, int target, uint64_t* dev_addr) {
    if (!ctx || !dev_addr) return LH_ERR_INVALID_ARGUMENT;
    switch (target) {
    case LH_TARGET_VIS64: *dev_addr = (uint64_t)(uintptr_t)ctx->vis; break;
    case LH_TARGET_DEPTH: *dev_addr = (uint64_t)(uintptr_t)ctx->depth_plane; break;
    case LH_TARGET_SCENE_RGBA16F: *dev_addr = (uint64_t)(uintptr_t)ctx->scene; break;
    case LH_TARGET_SHADOW: *dev_addr = (uint64_t)(uintptr_t)ctx->shadow; break;
    case LH_TARGET_DISPLAY_BGRA8: *dev_addr = (uint64_t)(uintptr_t)ctx->display; break;
    default: return fail(ctx, LH_ERR_INVALID_ARGUMENT, "unknown target");
    }
    return LH_OK;
}

int lh_get_timings(LhContext* ctx, LhPassTimes* out) {
    if (!ctx || !out) return LH_ERR_INVALID_ARGUMENT;
    if (!(ctx->cfg.flags & LH_CONFIG_PASS_TIMINGS)) return fail(ctx, LH_ERR_NOT_READY, "context was created without LH_CONFIG_PASS_TIMINGS");
    LH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memset(out, 0, sizeof(*out));
    for (int i = 0; i < LH_PASS_COUNT; i++)
        if (ctx->pass_recorded[i]) cudaEventElapsedTime(&out->ms[i], ctx->pass_start[i], ctx->pass_stop[i]);
    // the frame: first pass recorded to last pass recorded, in time order (the display transform is timed separately:
    // SURVEY 8d, the frame is steps 1-11)
    if (ctx->first_pass >= 0) cudaEventElapsedTime(&out->frame_ms, ctx->pass_start[ctx->first_pass], ctx->pass_stop[ctx->last_pass]);
    out->raster_meshlets_kernel_launches = ctx->raster_bands_recorded;
    for (uint32_t i = 0; i < ctx->raster_bands_recorded; i++) {   // sum over the main view's launches of the kernel
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, ctx->raster_band_a[i], ctx->raster_band_b[i]);
        out->raster_meshlets_kernel_ms += ms;
    }
    return LH_OK;
}

int lh_set_config_flags(LhContext* ctx, uint32_t flags) {
    if (!ctx) return LH_ERR_INVALID_ARGUMENT;
    ctx->cfg.flags = flags;
    return LH_OK;
}

int lh_get_launch_count(LhContext* ctx, uint64_t* n) {
    if (!ctx || !n) return LH_ERR_INVALID_ARGUMENT;
    *n = ctx->launches;
    return LH_OK;
}

int lh_timer_start(LhContext* ctx) {
    if (!ctx) return LH_ERR_INVALID_ARGUMENT;
    LH_CUDA(ctx, cudaEventRecord(ctx->timer_a, ctx->stream));
    return LH_OK;
}

int lh_timer_stop(LhContext* ctx, float* ms) {
    if (!ctx || !ms) return LH_ERR_INVALID_ARGUMENT;
    LH_CUDA(ctx, cudaEventRecord(ctx->timer_b, ctx->stream));
    LH_CUDA(ctx, cudaEventSynchronize(ctx->timer_b));
    LH_CUDA(ctx, cudaEventElapsedTime(ms, ctx->timer_a, ctx->timer_b));
    return LH_OK;
}

int lh_flush_l2(LhContext* ctx) {
    if (!ctx) return LH_ERR_INVALID_ARGUMENT;
    const size_t bytes = (size_t)256 << 20;  // > 126 MB of L2
    if (!ctx->l2_flush) {
        LH_CUDA(ctx, cudaMalloc(&ctx->l2_flush, bytes));
        ctx->l2_flush_bytes = bytes;
    }
    launch_fill_u32(ctx, (uint32_t*)ctx->l2_flush, bytes / 4, 0u);
    return check_launch(ctx, "flush_l2");
}

}  // extern "C"
