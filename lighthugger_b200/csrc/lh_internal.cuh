// Internal declarations shared by the CUDA translation units: the context,
// the per-launch parameter block and the pinned device arithmetic.
//
// Pinned arithmetic model: everything that decides a cull, a
// coverage bit or a depth value is written with __fmul_rn / __fadd_rn /
// __fdiv_rn / __fsqrt_rn so that no FMA contraction or approximate
// division can occur whatever the compiler flags are.  The operation order
// restates what the reference's GLSL leaves implementation-defined
// (src/shaders/common/culling.glsl, rasterization.glsl) in glm's order
// (mat4*vec4 = (c0*x + c1*y) + (c2*z + c3*w)).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "lighthugger_b200.h"

namespace lh {

// ---- pinned scalar / vector arithmetic ------------------------------------
__device__ __forceinline__ float pmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float padd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float psub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float pdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float psqrt(float a) { return __fsqrt_rn(a); }

struct V3 {
    float x, y, z;
};
struct V4 {
    float x, y, z, w;
};

// (c0*x + c1*y) + (c2*z + c3*w); m points at 16 floats, column-major
__device__ __forceinline__ V4 mat4_mul(const float* m, float x, float y, float z, float w) {
    V4 r;
    r.x = padd(padd(pmul(m[0], x), pmul(m[4], y)), padd(pmul(m[8], z), pmul(m[12], w)));
    r.y = padd(padd(pmul(m[1], x), pmul(m[5], y)), padd(pmul(m[9], z), pmul(m[13], w)));
    r.z = padd(padd(pmul(m[2], x), pmul(m[6], y)), padd(pmul(m[10], z), pmul(m[14], w)));
    r.w = padd(padd(pmul(m[3], x), pmul(m[7], y)), padd(pmul(m[11], z), pmul(m[15], w)));
    return r;
}
// (c0*x + c1*y) + c2*z; m points at 9 floats, column-major
__device__ __forceinline__ V3 mat3_mul(const float* m, float x, float y, float z) {
    V3 r;
    r.x = padd(padd(pmul(m[0], x), pmul(m[3], y)), pmul(m[6], z));
    r.y = padd(padd(pmul(m[1], x), pmul(m[4], y)), pmul(m[7], z));
    r.z = padd(padd(pmul(m[2], x), pmul(m[5], y)), pmul(m[8], z));
    return r;
}
__device__ __forceinline__ float dot3(V3 a, V3 b) { return padd(padd(pmul(a.x, b.x), pmul(a.y, b.y)), pmul(a.z, b.z)); }
__device__ __forceinline__ float length3(V3 a) { return psqrt(dot3(a, a)); }
// glm's normalize (what the reference's GLSL compiles to against its own glm, and one IEEE division instead of three):
// v * inversesqrt(dot(v, v)), inversesqrt(x) = 1 / sqrt(x)
__device__ __forceinline__ V3 normalize3(V3 a) {
    const float inv = pdiv(1.0f, length3(a));
    return V3{pmul(a.x, inv), pmul(a.y, inv), pmul(a.z, inv)};
}
__device__ __forceinline__ V3 cross3(V3 a, V3 b) {
    return V3{psub(pmul(a.y, b.z), pmul(b.y, a.z)), psub(pmul(a.z, b.x), pmul(b.z, a.x)), psub(pmul(a.x, b.y), pmul(b.x, a.y))};
}

// pow(x, p) for x > 0 by a fixed sequence of IEEE double operations (no libm),
// the same sequence the CPU checker uses (the reference's GLSL pow is
// implementation-defined).
__device__ inline float pinned_powf(float xf, float pf) {
    if (!(xf > 0.0f)) return xf == 0.0f ? 0.0f : __int_as_float(0x7FC00000);
    double x = (double)xf;
    unsigned long long bits = (unsigned long long)__double_as_longlong(x);
    int e = (int)((bits >> 52) & 0x7FF) - 1023;
    bits = (bits & 0x000FFFFFFFFFFFFFull) | 0x3FF0000000000000ull;
    double m = __longlong_as_double((long long)bits);
    if (m > 1.4142135623730951) {
        m = __dmul_rn(m, 0.5);
        e += 1;
    }
    double s = __ddiv_rn(__dsub_rn(m, 1.0), __dadd_rn(m, 1.0));
    double s2 = __dmul_rn(s, s);
    double term = s, acc = 0.0;
    for (int k = 1; k <= 25; k += 2) {
        acc = __dadd_rn(acc, __ddiv_rn(term, (double)k));
        term = __dmul_rn(term, s2);
    }
    const double ln2 = 0.6931471805599453;
    double lnx = __dadd_rn(__dmul_rn((double)e, ln2), __dmul_rn(2.0, acc));
    double y = __dmul_rn((double)pf, lnx);
    double kf = floor(__dadd_rn(__ddiv_rn(y, ln2), 0.5));
    double r = __dsub_rn(y, __dmul_rn(kf, ln2));
    double et = 1.0, ex = 1.0;
    for (int k = 1; k <= 18; k++) {
        et = __ddiv_rn(__dmul_rn(et, r), (double)k);
        ex = __dadd_rn(ex, et);
    }
    int ki = (int)kf;
    if (ki > 1000) return __int_as_float(0x7F800000);
    if (ki < -1000) return 0.0f;
    double p2 = __longlong_as_double((long long)((unsigned long long)(ki + 1023) << 52));
    return (float)__dmul_rn(ex, p2);
}


// ---- device-side scratch that lives across passes of one frame -------------
struct FrameState {
    uint32_t overflow;             // LH_OVERFLOW_* bits
    uint32_t main_large_count;     // ... as left by the main-view pass (stats)
    uint32_t main_visible, main_candidates, main_opaque, main_alpha_clip;  // main-view counts (the shared counters are reused by the cascades)
    uint32_t shadow_opaque[4];     // draw counts of each cascade, saved for the stats
    uint32_t shadow_alpha_clip[4];
    uint32_t shadow_visible[4];
    uint32_t shadow_candidates[4];
    uint32_t done[4];              // instance cull: CTAs finished per sector (self-cleaning of the tile states)
    unsigned long long rasterized_triangles;
    uint32_t skipped_empty[5], skipped_occluded[5], queued[5];
    unsigned long long samples[5];      // reductions issued by the rasteriser (LH_CONFIG_COUNT_SAMPLES): [0] main view, [1..4] cascades   // meshlets dropped by the rasteriser's early-outs: [0] main view, [1..4] cascades (stats)
};

// Work counters of one view (the main view or one cascade) through meshlet cull and rasterisation.  Each view that
// can be in flight at the same time has its own, with its own scratch buffers, so cascades can run on separate streams.
struct PassState {
    uint32_t cull_ticket, cull_done;   // meshlet cull: next tile, CTAs finished (self-cleaning)
    // Deferred triangles.  Raster launches append to the deferred queue, the first round of a drain appends the
    // per-tile pieces of the triangles it splits to the split queue (the second half of the same allocation), the
    // second round rasterises those.  Ranges are handed over by index, written once by the last CTA of the producing
    // launch, so a drain may run on its own stream beside the next band.
    uint32_t large_count, split_count;
    uint32_t drain_end[9];             // [b + 1]: large_count when band b's rasterisation finished; [0] = 0
    uint32_t split_end[9];             // [d + 1]: split_count when drain d's first round finished; [0] = 0
    uint32_t cta_done_deferred;        // CTAs of the running deferred launch that have finished (self-cleaning)
    uint32_t queue_count;              // entries appended to the raster queue so far
    uint32_t band_mark[9];             // queue_count when band b's rasterisation began: band b owns [band_mark[b], band_mark[b+1])
    uint32_t band_ticket[8];           // next unclaimed entry of band b's range, relative to band_mark[b]
    uint32_t band_key[8];              // exclusive upper depth key of each band
    uint32_t cta_done;                 // CTAs of the running raster / deferred launch that have finished (self-cleaning)
};

constexpr uint32_t kMaxStripes = 32;          // sort-last ownership stripes per view (lh_set_instance_stripes / _stripe_bounds)
constexpr uint32_t kMeshletTile = 1024;        // candidate meshlets per scan tile
constexpr uint32_t kInstanceTile = 256;        // instances per scan tile
constexpr uint64_t kMaxMeshletTiles = (1ull << 32) / kMeshletTile;  // lo32 of the prefix counter bounds the candidates

// Everything a kernel needs, passed by value (__grid_constant__).
struct FrameParams {
    LhUniforms u;
    uint32_t max_instances, max_opaque, max_alpha;
    uint32_t instance_first, instance_end;   // sort-last partition
    uint32_t shadow_instance_first, shadow_instance_end;   // ... of the cascades (lh_set_shadow_instance_range)
    // sort-last ownership by screen stripes (lh_set_instance_stripes): stripes == 0 means off; otherwise an instance
    // belongs to the context whose index equals (stripe of its projected centre) % count.  [0] main view, [1] cascades
    uint32_t stripes[2], stripe_index[2], stripe_count[2];
    float stripe_bounds[2][kMaxStripes + 1]; // stripe s holds the centres with bounds[s] <= t < bounds[s + 1], t in [0, 1] across the view
    uint32_t shadowmap_size;
    uint32_t flags;
    uint32_t max_large;
    uint32_t num_bands;                      // depth bands of the raster pass being launched (raster.cu)
    uint32_t small_max_area;                 // a lane finishes triangles of at most this many pixel tests itself (raster.cu)
    float band_fractions[7];                 // cumulative meshlet-count fractions at which its bands end; [0] == 0: built-in table
    FrameState* state;
    unsigned long long* instance_tiles;      // [4][tiles] decoupled-lookback states
    // ---- the view being culled / rasterised (see PassState) ----
    PassState* pass;
    uint64_t pass_draw_calls;                // its DrawCallBuffer: u.draw_calls, or an internal one for a concurrent cascade
    uint64_t pass_refs;                      // first MeshletReference of its list (u.meshlet_references + section, or internal)
    uint32_t pass_slot_base;                 // DrawIndirectCommand.firstInstance of its first draw (MESHLET_INDICES_BUFFER_SECTION_OFFSET for cascades)
    unsigned long long* meshlet_tiles;       // decoupled-lookback states of the meshlet cull
    uint32_t* large_queue;
    void* meshlet_bounds;                    // 16 B per surviving meshlet of the current raster pass (raster.cu MeshletBound)
    void* raster_records;                    // 64 B per meshlet selected for rasterisation, band after band (raster.cu RasterRecord)
    uint32_t first_wave_ctas;                // CTAs of the raster kernel that are resident together when it has the GPU to itself
    uint32_t* depth_keys;                    // the bounds' depth keys on their own (4 B per meshlet): what band selection scans
    uint32_t* depth_hist;                    // 8192 bins of the depth-key histogram (self-cleaning)
    uint32_t* hiz;                           // two hierarchical-Z levels of the current raster target: farthest stored value per 4x4- and per 16x16-pixel tile
    unsigned long long* vis;
    uint32_t* shadow;                        // 4 layers of float bits
    const cudaTextureObject_t* textures;
    const uint4* tex_dims;                   // per texture slot: width, height, mip levels, 0
    uint32_t slot_offset;                    // sort-last: global slot of this rank's local slot 0
    const LhMeshletReference* resolve_refs;  // table indexed by GLOBAL slot (== u.meshlet_references on one GPU)
    // peer exchange (multi-GPU, LH_EXCHANGE_PEER): device arrays of one pointer per rank into the peers' memory
    const LhMeshletReference* const* peer_refs;   // rank q's table, indexed by q's LOCAL slot; nullptr: use resolve_refs
    const unsigned long long* const* peer_vis;    // rank q's key buffer
    uint32_t slots_per_rank, rank, nranks;
    void* scene_remote;                      // rank 0's RGBA16F target mapped here (ranks != 0), else nullptr
    uint32_t* depth_pair;                    // {min depth bits, ~max depth bits} of this rank's rows, then of the frame
    void* scene;                             // RGBA16F target
    void* resolve_tmp;                       // 40 B per pixel handed from the resolve's material phase to its shadow phase
    void* display;                           // BGRA8 target
    cudaTextureObject_t sky, lut;
    uint32_t resolve_y0, resolve_y1;         // screen rows this rank resolves
};

}  // namespace lh

struct LhContext {
    int device = 0;
    LhConfig cfg{};
    cudaStream_t stream = nullptr;
    std::string error;
    LhUniforms uniforms{};
    bool have_uniforms = false;
    uint32_t instance_first = 0, instance_count = 0xFFFFFFFFu;
    bool shadow_range_set = false;
    uint32_t stripes[2] = {0, 0}, stripe_index[2] = {0, 0}, stripe_count[2] = {1, 1};
    float stripe_bounds[2][33] = {};   // kMaxStripes + 1
    uint32_t shadow_instance_first = 0, shadow_instance_count = 0xFFFFFFFFu;
    uint32_t cascade_mask = 0xF;

    // targets
    unsigned long long* vis = nullptr;
    void* scene = nullptr;       // RGBA16F
    uint32_t* shadow = nullptr;
    void* display = nullptr;     // BGRA8
    float* depth_plane = nullptr;
    void* resolve_tmp = nullptr;           // allocated by the first lh_render_frame
    cudaStream_t resolve_stream = nullptr; // the material phase of the resolve runs here, beside the shadow passes
    cudaEvent_t vis_ready = nullptr, material_done = nullptr;

    // scratch
    lh::FrameState* state = nullptr;
    unsigned long long* instance_tiles = nullptr;
    uint32_t instance_tiles_per_sector = 0;
    // Per-view scratch.  Set 0 serves the main view and whichever cascade runs on the context stream with the
    // host-visible lists (u.draw_calls, u.meshlet_references); sets 1..3 serve cascades that lh_render_frame runs
    // concurrently on side streams, with internal lists (allocated on first use).
    struct PassScratch {
        lh::PassState* pass = nullptr;
        unsigned long long* meshlet_tiles = nullptr;
        uint32_t* large_queue = nullptr;
        void* meshlet_bounds = nullptr;
        void* raster_records = nullptr;
        uint32_t* depth_keys = nullptr;
        uint32_t* depth_hist = nullptr;
        uint32_t* hiz = nullptr;
        void* draw_calls = nullptr;   // internal lists (sets 1..3)
        void* refs = nullptr;
    } scratch[4];
    size_t meshlet_tile_count = 0;
    cudaStream_t deferred_stream = nullptr;   // the main view's deferred-triangle drains run here, beside the next band
    cudaEvent_t band_done[8] = {}, drains_done = nullptr;
    cudaStream_t side_streams[3] = {};
    cudaEvent_t fork_event = nullptr, join_events[3] = {};
    void* l2_flush = nullptr;
    size_t l2_flush_bytes = 0;
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    LhFrameStats* pinned_stats = nullptr;

    // textures
    cudaTextureObject_t tex_host[LH_MAX_TEXTURES] = {};
    cudaMipmappedArray_t tex_arrays[LH_MAX_TEXTURES] = {};
    cudaTextureObject_t* tex_dev = nullptr;
    uint4* tex_dims_dev = nullptr;
    cudaArray_t sky_array = nullptr;
    cudaTextureObject_t sky_tex = 0;
    cudaArray_t lut_array = nullptr;
    cudaTextureObject_t lut_tex = 0;

    // timing
    cudaEvent_t pass_start[LH_PASS_COUNT] = {}, pass_stop[LH_PASS_COUNT] = {};
    bool pass_recorded[LH_PASS_COUNT] = {};
    int first_pass = -1, last_pass = -1;   // of the current frame, in the order they were recorded
    cudaEvent_t timer_a = nullptr, timer_b = nullptr;
    cudaEvent_t raster_band_a[8] = {}, raster_band_b[8] = {};   // around each main-view raster_meshlets_kernel launch
    uint32_t raster_bands_recorded = 0;
    uint64_t launches = 0;

    // multi-GPU
    void* nccl_comm = nullptr;
    int rank = 0, nranks = 1;
    uint32_t slot_offset = 0;          // rank * slots_per_rank
    uint32_t slots_per_rank = 0;
    void* gathered_refs = nullptr;     // all ranks' meshlet-reference tables, rank-major (LH_EXCHANGE_NCCL only)
    // LH_EXCHANGE_PEER: buffers of the peers mapped through CUDA IPC (comm.cu)
    int exchange = LH_EXCHANGE_PEER;
    struct PeerMap {
        void* local = nullptr;             // this rank's buffer
        std::vector<void*> ptrs;           // per rank: the buffer as seen from this device (own entry == local)
        void** table_dev = nullptr;        // the same, on the device
    };
    std::vector<PeerMap> peer_maps;
    uint32_t* comm_scratch = nullptr;      // [0..1] depth pair, [2] barrier word, [16..] handle staging
    void** peer_vis_dev = nullptr;         // device tables of the current frame (owned by peer_maps)
    void** peer_refs_dev = nullptr;
    void* scene_remote = nullptr;
    uint64_t mapped_refs = 0;              // the MeshletReference buffer the peers currently map (0: none)
    uint32_t row0 = 0, row1 = 0;           // screen rows this rank composites and resolves
    int num_sms = 148;
    // launch shapes of the rasteriser; defaults measured on B200 at BASELINE config 2.  Developer overrides:
    // LH_TUNE="main_bands,shadow_bands,main_grid,shadow_grid,resolve_low_priority,select_grid,resolve_split,small_max_area" in the environment at context creation.
    struct Tuning {
        float main_fractions[7] = {0, 0, 0, 0, 0, 0, 0};   // LH_TUNE_FRACTIONS="f0,f1,..." (main view)
        uint32_t main_bands = 4, shadow_bands = 6, main_grid = 16, shadow_grid = 3, resolve_low_priority = 1, select_grid = 16, resolve_split = 1, small_max_area = 128;
        bool from_environment = false;     // LH_TUNE was given: no automatic choices
    } tune;

    std::vector<void*> buffers;
};

namespace lh {

int fail(LhContext* ctx, int code, const char* what, cudaError_t e = cudaSuccess);
// cascade < 0: the main view.  set: which PassScratch; internal_lists: the view's lists live in the set's own buffers
FrameParams make_params(LhContext* ctx, int cascade = -1, int set = 0, bool internal_lists = false);

// pass launchers (one per translation unit); stream == nullptr means the context stream
void launch_reset(LhContext* ctx, const FrameParams& p);
void launch_cull_instances(LhContext* ctx, const FrameParams& p, bool shadows);
void launch_cull_meshlets(LhContext* ctx, const FrameParams& p, int cascade, cudaStream_t stream = nullptr);
void launch_raster(LhContext* ctx, const FrameParams& p, int cascade, cudaStream_t stream = nullptr);
void launch_reduce_depth(LhContext* ctx, const FrameParams& p);
void launch_shadow_matrices(LhContext* ctx, const FrameParams& p);
// phase 0: the whole pass; 1: material (no shadow maps needed); 2: shadows + combination (resolve.cu)
void launch_resolve(LhContext* ctx, const FrameParams& p, int phase = 0, cudaStream_t stream = nullptr);
void launch_display_transform(LhContext* ctx, const FrameParams& p);
void launch_extract_depth(LhContext* ctx, const FrameParams& p);
void launch_gather_stats(LhContext* ctx, const FrameParams& p, LhFrameStats* dst_dev_or_pinned);
void launch_fill_u64(LhContext* ctx, unsigned long long* dst, uint64_t count, unsigned long long value);
void launch_fill_u32(LhContext* ctx, uint32_t* dst, uint64_t count, uint32_t value, cudaStream_t stream = nullptr);
int composite_allreduce_min(LhContext* ctx);
int composite_pull(LhContext* ctx);          // LH_EXCHANGE_PEER: barrier + pull-min kernel with the depth bounds folded in
int reduce_depth_exchange(LhContext* ctx);   // ... the frame's depth bounds from the ranks' row bounds
int frame_barrier(LhContext* ctx);           // every rank has finished its part of the frame
int comm_poll_errors(LhContext* ctx);        // ncclCommGetAsyncError: LH_ERR_NCCL when the communicator has failed
bool peer_exchange(const LhContext* ctx);    // multi-GPU and LH_EXCHANGE_PEER
int ensure_peer_maps(LhContext* ctx);        // maps what the frame needs of the peers (collective the first time)
int peer_release_collective(LhContext* ctx); // every rank unmaps everything, then a barrier: buffers may now be freed / moved
void comm_update_rows(LhContext* ctx);       // row slices follow the target height
extern "C" int lh_shadow_exchange(LhContext* ctx);

}  // namespace lh

#define LH_CUDA(ctx, expr)                                              \
    do {                                                                \
        cudaError_t _e = (expr);                                        \
        if (_e != cudaSuccess) return lh::fail(ctx, LH_ERR_CUDA, #expr, _e); \
    } while (0)
