/*
 * lighthugger_b200.h -- C ABI of the B200-native geometry pipeline.
 *
 * The reference (expenses/lighthugger) has no plugin/FFI boundary: its
 * de-facto boundary is the argument list of render() (src/rendering.h:4-16)
 * plus the memory contract of src/shared_cpu_gpu.h, everything on the GPU
 * being reached from one 64-bit uniform-buffer address (src/rendering.cpp:56-61,
 * src/shaders/common/bindings.glsl:23-30).  This header is what a C++20 host
 * calls where the reference recorded Vulkan compute and raster passes.  Every
 * export cites the reference interface it replaces.
 *
 * Conventions
 *  - every function returns 0 (LH_OK) or a negative LH_ERR_* code; no
 *    exception or abort crosses the boundary (the reference aborts on any
 *    negative VkResult, src/util.cpp:4-11); lh_last_error() gives the text;
 *  - a context is externally synchronised (like a VkQueue): one host thread
 *    at a time, one CUDA stream, one GPU.  Pass functions enqueue and return;
 *    lh_frame_sync() is the fence wait (src/main.cpp:766-773);
 *  - the library owns device memory, the host owns host memory; no host
 *    pointer is retained past the call;
 *  - there is no CPU fallback: if no CUDA device is usable,
 *    lh_context_create() fails.
 */
#ifndef LIGHTHUGGER_B200_H
#define LIGHTHUGGER_B200_H

#include "lh_shared_cpu_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define LH_API __attribute__((visibility("default")))

enum {
    LH_OK = 0,
    LH_ERR_INVALID_ARGUMENT = -1,
    LH_ERR_CUDA = -2,
    LH_ERR_OUT_OF_MEMORY = -3,
    LH_ERR_CAPACITY = -4,        /* a list overflowed its LhConfig capacity */
    LH_ERR_NO_DEVICE = -5,
    LH_ERR_NCCL = -6,
    LH_ERR_NOT_READY = -7,
    LH_ERR_UNSUPPORTED = -8
};

enum {
    LH_CONFIG_WRITE_DRAW_COMMANDS = 1u << 0,  /* keep materialising DrawIndirectCommand[] (default on) */
    LH_CONFIG_PASS_TIMINGS = 1u << 1,         /* record a CUDA event pair around every pass */
    LH_CONFIG_COUNT_SAMPLES = 1u << 2         /* count the depth-tested samples the rasteriser writes (LhFrameStats.emitted_samples) */
};

typedef struct LhContext LhContext;

/* Capacities default to the reference's compile-time constants
 * (src/shared_cpu_gpu.h:119-125,159); BASELINE configs 2-5 need larger ones. */
typedef struct LhConfig {
    uint32_t struct_size;          /* = sizeof(LhConfig) */
    int32_t device;                /* CUDA device ordinal */
    uint32_t max_instances;        /* values per prefix-sum sector (1024) */
    uint32_t max_opaque_draws;     /* 200000 */
    uint32_t max_alpha_clip_draws; /* 200000 */
    uint32_t width, height;        /* render-target size */
    uint32_t shadowmap_size;       /* 1024 */
    uint32_t flags;                /* LH_CONFIG_* */
    uint32_t max_large_triangles;  /* capacity of the deferred large-triangle queue */
} LhConfig;

/* Sizes the host must give the six buffers named by LhUniforms
 * (src/main.cpp:442-532 with the capacities above). */
typedef struct LhBufferSizes {
    uint64_t meshlet_references;       /* 8 * (max_opaque+max_alpha) * 2 */
    uint64_t num_meshlets_prefix_sum;  /* (8 * max_instances + 8) * 4 */
    uint64_t draw_calls;               /* 8 + 16 * (max_opaque+max_alpha) */
    uint64_t dispatches;               /* 12 * 3 */
    uint64_t misc_storage;             /* 792 */
    uint64_t uniforms;                 /* 560 */
} LhBufferSizes;

/* Per-frame counters read back by lh_get_frame_stats (new: the reference has
 * no readback, only debug views). */
typedef struct LhFrameStats {
    uint32_t num_visible_instances;     /* counter hi32, sector 0 */
    uint32_t num_candidate_meshlets;    /* counter lo32, sector 0 */
    uint32_t num_opaque;                /* draw_calls.num_opaque after the main-view cull */
    uint32_t num_alpha_clip;
    uint32_t min_depth;                 /* MiscStorage.min_depth (float bits) */
    uint32_t max_depth;
    uint32_t shadow_visible_instances[4];
    uint32_t shadow_candidate_meshlets[4];
    uint32_t shadow_opaque[4];
    uint32_t shadow_alpha_clip[4];
    uint32_t num_large_triangles;       /* triangles deferred to the binned path (main view) */
    uint32_t overflow;                  /* bit mask of LH_OVERFLOW_* */
    uint64_t rasterized_triangles;      /* main view, after cull / clip (0 unless LH_CONFIG_PASS_TIMINGS) */
    /* rasteriser early-outs, [0] main view, [1..4] cascades (counted only with LH_CONFIG_PASS_TIMINGS) */
    uint32_t meshlets_skipped_no_sample[5]; /* the meshlet's bound holds no pixel centre */
    uint32_t meshlets_skipped_occluded[5];  /* every key already stored in the bound is nearer */
    uint32_t meshlets_queued[5];            /* meshlets the band selection handed to the raster kernel */
    uint32_t _pad1;
    uint64_t emitted_samples[5];            /* 64-/32-bit min reductions issued, [0] main view, [1..4] cascades (LH_CONFIG_COUNT_SAMPLES) */
} LhFrameStats;

enum {
    LH_OVERFLOW_INSTANCES = 1u << 0,
    LH_OVERFLOW_OPAQUE = 1u << 1,
    LH_OVERFLOW_ALPHA_CLIP = 1u << 2,
    LH_OVERFLOW_LARGE_TRIANGLES = 1u << 3
};

/* Pass names follow the reference's Tracy zones (src/rendering.cpp:140-492). */
enum {
    LH_PASS_RESET = 0,
    LH_PASS_CULL_INSTANCES,
    LH_PASS_CULL_MESHLETS,
    LH_PASS_VISBUFFER,
    LH_PASS_COMPOSITE,
    LH_PASS_DEPTH_REDUCTION,
    LH_PASS_SHADOW_MATRICES,
    LH_PASS_CULL_INSTANCES_SHADOWS,
    LH_PASS_SHADOWMAPS,
    LH_PASS_RENDER_GEOMETRY,
    LH_PASS_DISPLAY_TRANSFORM,
    LH_PASS_SHADOW_EXCHANGE,   /* sort-last: merging the ranks' shadow layers (new) */
    LH_PASS_FRAME_BARRIER,     /* sort-last: every rank has delivered its rows of the frame (new) */
    LH_PASS_COUNT
};

typedef struct LhPassTimes {
    float ms[LH_PASS_COUNT];
    float frame_ms;                /* first pass start -> last pass end */
    float raster_meshlets_kernel_ms; /* the main view's launches of the one-warp-per-meshlet raster kernel, summed */
    uint32_t raster_meshlets_kernel_launches; /* how many launches that was (one per depth band) */
} LhPassTimes;

enum {
    LH_TARGET_VIS64 = 0,        /* uint64 per pixel: packed key (see lh_shared_cpu_gpu.h) */
    LH_TARGET_DEPTH = 1,        /* float per pixel, derived from VIS64 (reference D32F) */
    LH_TARGET_SCENE_RGBA16F = 2,/* 4 x half per pixel (reference scene-referred framebuffer) */
    LH_TARGET_SHADOW = 3,       /* float per texel, 4 layers of shadowmap_size^2 */
    LH_TARGET_DISPLAY_BGRA8 = 4 /* 4 x u8 per pixel (reference swapchain image) */
};

enum {
    LH_TEXTURE_RGBA8_SRGB = 0,
    LH_TEXTURE_RGBA8_UNORM = 1,
    LH_TEXTURE_BC7_SRGB = 2,
    LH_TEXTURE_BC1_SRGB = 3
};

/* ---- lifetime (replaces VMA/RAII wrappers, src/allocations/base.cpp:5-61) ---- */
LH_API int lh_config_default(LhConfig* cfg);
LH_API int lh_buffer_sizes(const LhConfig* cfg, LhBufferSizes* out);
LH_API int lh_context_create(const LhConfig* cfg, LhContext** out);
LH_API int lh_context_destroy(LhContext* ctx);
LH_API const char* lh_last_error(const LhContext* ctx);   /* ctx may be NULL: last create error */
LH_API const char* lh_version(void);

/* ---- buffers (replaces AllocatedBuffer + upload_via_staging_buffer,
 * src/allocations/staging.cpp:3-47).  dev_addr is a plain device address the
 * host stores in LhUniforms / LhInstance / LhMeshInfo fields. ---- */
LH_API int lh_buffer_alloc(LhContext* ctx, uint64_t bytes, uint64_t* dev_addr);
LH_API int lh_buffer_free(LhContext* ctx, uint64_t dev_addr);
LH_API int lh_upload(LhContext* ctx, uint64_t dev_addr, const void* src, uint64_t bytes);
LH_API int lh_download(LhContext* ctx, uint64_t dev_addr, void* dst, uint64_t bytes);

/* ---- bindless textures (replaces DescriptorSet::write_image,
 * src/descriptor_set.cpp:140-165; 512 slots, 16-bit indices).  data holds all
 * mip levels tightly packed, largest first.  Sampled like the reference's
 * repeat_sampler (src/main.cpp:550-560): linear, linear-mip, anisotropy 16. ---- */
LH_API int lh_texture_create(LhContext* ctx, int format, uint32_t width, uint32_t height,
                             uint32_t mip_levels, const void* data, uint64_t bytes, uint16_t* slot);
LH_API int lh_texture_destroy(LhContext* ctx, uint16_t slot);
/* skybox cubemap, 6 faces (+X,-X,+Y,-Y,+Z,-Z) of size^2 RGBA32F texels
 * (the reference loads an RGB9E5 DDS cubemap, src/main.cpp:542-549). */
LH_API int lh_skybox_create(LhContext* ctx, uint32_t size, const float* rgba32f_faces);
/* display-transform LUT: size^3 RGB32F texels, x fastest (the reference loads the
 * 48^3 RGB9E5 Tony McMapface DDS, src/main.cpp:534-541; decode on the host). */
LH_API int lh_set_display_lut(LhContext* ctx, uint32_t size, const float* rgb32f);

/* ---- render targets (replaces ResizingResources, src/frame_resources.h:7-73) ---- */
LH_API int lh_targets_resize(LhContext* ctx, uint32_t width, uint32_t height);

/* ---- sort-last partition (new; no reference counterpart).
 * The context culls and rasterises instances [first, first+count) only. ---- */
LH_API int lh_set_instance_range(LhContext* ctx, uint32_t first, uint32_t count);
/* Sort-last ownership by SCREEN STRIPES instead of (or on top of) an index range: the view is cut into `stripes`
 * vertical stripes and an instance belongs to the context whose `index` equals (stripe of its bounding-sphere centre)
 * % `count`.  view 0: the main view (stripes of the camera image); view 1: the cascades (stripes of each cascade's
 * light-space square).  Neighbouring instances then land on the same GPU, so each GPU's own geometry keeps occluding
 * itself (an index range hands every GPU a sparse copy of the whole scene whose hidden layers all get rasterised);
 * interleaving several stripes per GPU evens the load.  stripes == 0 turns it off (default). */
LH_API int lh_set_instance_stripes(LhContext* ctx, int view, uint32_t stripes, uint32_t index, uint32_t count);
/* The same with stripes of unequal width: bounds[0 .. stripes] ascend from 0 to 1 across the view (the outer two are
 * taken as open-ended); every context of the group must be given the same bounds.  A host that measures each GPU's
 * pass times can move the bounds from frame to frame to even the load (lighthugger_b200/sortlast.py does). */
LH_API int lh_set_instance_stripe_bounds(LhContext* ctx, int view, const float* bounds, uint32_t stripes, uint32_t index, uint32_t count);
/* cascades this context renders in lh_render_frame (bit c = cascade c; default 0xF).
 * "One cascade per GPU" = mask 1 << rank; the layers are merged by lh_shadow_exchange. */
LH_API int lh_set_shadow_cascade_mask(LhContext* ctx, uint32_t mask);
/* instances this context's cascades draw, [first, first+count); default: the range of lh_set_instance_range.
 * Over all ranks the (cascade, instance) pairs must be covered exactly once: either every rank draws all four
 * cascades of its own instances, or one cascade per GPU over all instances (BASELINE config 3). */
LH_API int lh_set_shadow_instance_range(LhContext* ctx, uint32_t first, uint32_t count);

/* ---- per frame (replaces uniform_buffer.flush + render() + submit + fence,
 * src/main.cpp:784-817).  One export per pass, in src/rendering.cpp order. ---- */
LH_API int lh_frame_begin(LhContext* ctx, const LhUniforms* uniforms);          /* main.cpp:784 */
LH_API int lh_reset(LhContext* ctx);                    /* reset_buffers_a, rendering.cpp:79 */
LH_API int lh_cull_instances(LhContext* ctx);           /* cull_instances.comp:7-27, rendering.cpp:139-147 */
LH_API int lh_cull_meshlets(LhContext* ctx);            /* reset_buffers_b + write_draw_calls.comp:7-57, rendering.cpp:156-183 */
LH_API int lh_raster_visbuffer(LhContext* ctx);         /* rasterization.glsl:38-97 + fixed function incl. the depth clear 0.0, rendering.cpp:195-266 */
LH_API int lh_composite(LhContext* ctx);                /* new: sort-last u64-min composite (peer-memory kernel or NCCL, lh_comm_set_exchange); no-op without lh_comm_init */
LH_API int lh_reduce_depth(LhContext* ctx);             /* compute/read_depth.glsl:22-83, rendering.cpp:289-300 */
LH_API int lh_generate_shadow_matrices(LhContext* ctx); /* compute/generate_shadow_matrices.glsl:15-110, rendering.cpp:302 */
LH_API int lh_cull_instances_shadows(LhContext* ctx);   /* cull_instances.comp:31-56, rendering.cpp:304-312 */
LH_API int lh_shadow_cascade(LhContext* ctx, uint32_t cascade); /* reset_buffers_c + write_draw_calls_shadows + shadowmap_*; rendering.cpp:325-438 */
LH_API int lh_shadow_exchange(LhContext* ctx);          /* new: uint32-min merge of shadow layers over NCCL; no-op without lh_comm_init */
LH_API int lh_resolve(LhContext* ctx);                  /* render_geometry.comp:70-266, rendering.cpp:462-474 */
LH_API int lh_display_transform(LhContext* ctx);        /* display_transform.comp:37-81, rendering.cpp:491-502 */
LH_API int lh_render_frame(LhContext* ctx);             /* all of the above in rendering.cpp order (display transform excluded) */
LH_API int lh_frame_sync(LhContext* ctx);               /* fence wait; reports LH_ERR_CAPACITY on list overflow */

/* Whole frame from HOST buffers: uploads the uniforms and (if non-NULL) the
 * instance array to uniforms->instances, renders, reads the stats back and
 * waits.  This is the end-to-end call timed by bench.py's "e2e". */
LH_API int lh_render_frame_host(LhContext* ctx, const LhUniforms* uniforms,
                                const LhInstance* instances, uint32_t num_instances,
                                LhFrameStats* stats_out);

/* ---- readback (new; needed for parity and bench) ---- */
LH_API int lh_get_frame_stats(LhContext* ctx, LhFrameStats* out);   /* synchronises */
/* replaces LhConfig.flags for the frames that follow (e.g. to count samples on one untimed frame) */
LH_API int lh_set_config_flags(LhContext* ctx, uint32_t flags);
LH_API int lh_get_target(LhContext* ctx, int target, void* dst, uint64_t bytes);
LH_API int lh_get_timings(LhContext* ctx, LhPassTimes* out);
LH_API int lh_get_launch_count(LhContext* ctx, uint64_t* kernels_launched);   /* kernels this context has launched so far */
LH_API int lh_get_target_address(LhContext* ctx, int target, uint64_t* dev_addr);
/* time one span on the context stream (CUDA events): bench helpers */
LH_API int lh_timer_start(LhContext* ctx);
LH_API int lh_timer_stop(LhContext* ctx, float* ms);    /* synchronises */
LH_API int lh_flush_l2(LhContext* ctx);                 /* writes a buffer larger than L2 */

/* ---- multi-GPU (new).  One context per process/GPU.  The id is
 * ncclUniqueId bytes (128) created on rank 0 by lh_comm_unique_id and
 * distributed by the host (e.g. a torch.distributed broadcast). ---- */
LH_API int lh_comm_unique_id(void* id128);
LH_API int lh_comm_init(LhContext* ctx, int rank, int nranks, const void* id128);
LH_API int lh_comm_destroy(LhContext* ctx);
LH_API int lh_comm_info(LhContext* ctx, int* rank, int* nranks);
/* How the exchange steps move data.  LH_EXCHANGE_PEER (default when every rank can map its peers' buffers through
 * CUDA IPC over NVLink): each rank owns the screen rows [y0, y1) of lh_comm_rows; every rank flags the 64x16-pixel tiles
 * of its key buffer it drew into, and lh_composite is ONE kernel that pulls the rank's rows of the peers that drew
 * there through peer memory, takes the uint64 min and folds the depth min/max in (the frame width must be even: keys
 * move as 16-byte pairs; LH_ERR_UNSUPPORTED otherwise); the resolve reads the winners' MeshletReference from the owning rank's table through peer memory and stores
 * its rows of lit colour into rank 0's frame.  LH_EXCHANGE_NCCL: ncclAllReduce(uint64, min) of whole buffers and an
 * all-gather of the reference tables; every rank then holds and resolves the whole frame.
 * LH_EXCHANGE_PEER maps whole allocations: LhUniforms.meshlet_references must be an address lh_buffer_alloc returned
 * (not an offset into a larger buffer), and ranks must move that buffer, resize their targets or destroy their
 * communicator together. */
enum { LH_EXCHANGE_PEER = 0, LH_EXCHANGE_NCCL = 1 };
LH_API int lh_comm_set_exchange(LhContext* ctx, int mode);
/* Collective (LH_EXCHANGE_PEER): every rank drops its mappings of its peers' buffers and waits for the others to have done
 * so; call it on all ranks before moving LhUniforms.meshlet_references to another buffer.  lh_targets_resize does the
 * same internally (ranks resize together); the next lh_frame_begin maps the current buffers again. */
LH_API int lh_comm_remap(LhContext* ctx);
LH_API int lh_comm_rows(LhContext* ctx, uint32_t* y0, uint32_t* y1);   /* rows this rank composites and resolves */
/* Replaces the rank's even share of the screen rows (lh_comm_init / lh_targets_resize restore it): rows of sky resolve
 * in a fraction of the time of rows of geometry, so a host may move the boundaries towards equal measured pass times
 * (sortlast.RowBalancer).  The ranks' slices must tile [0, height) between them; the library cannot check that. */
LH_API int lh_comm_set_rows(LhContext* ctx, uint32_t y0, uint32_t y1);

#ifdef __cplusplus
}
#endif
#endif /* LIGHTHUGGER_B200_H */
