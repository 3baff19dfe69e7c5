// Sort-last multi-GPU exchange (new; the reference is single-GPU).
//
// One process / context per GPU.  Instances are partitioned across ranks; each
// rank rasterises its partition into a full-resolution packed visibility
// buffer whose keys carry GLOBAL slots (rank * slots_per_rank + local), so the
// per-pixel uint64 MIN over ranks is exactly the single-GPU buffer.
//
// LH_EXCHANGE_PEER (default).  Every rank maps its peers' key buffer,
// MeshletReference table and rank 0's colour target through CUDA IPC (NVLink
// peer memory; the handles travel through one ncclAllGather) and owns the
// screen rows [row0, row1).  Per frame, on the context stream:
//   composite        barrier, then ONE kernel: for its rows, 128-bit loads of
//                    every rank's keys straight from peer memory, uint64 min,
//                    result stored in place, depth min/max of the rows folded in
//                    (in place is safe: a rank writes only its own rows of its
//                    own buffer, and its peers read only THEIR rows of it)
//   depth bounds     ncclAllReduce(min) of {min, ~max}: 8 bytes, doubles as the
//                    barrier after which key buffers may be reused
//   shadow layers    ncclAllReduce(uint32, min) over the 4 depth layers
//   resolve          rows [row0,row1) only; the winner's MeshletReference is read
//                    from the owning rank's table through peer memory; the lit
//                    colour is stored into rank 0's frame through peer memory
//   frame barrier    4-byte all-reduce: rank 0's frame is complete
// LH_EXCHANGE_NCCL.  Whole-buffer collectives; every rank resolves everything:
//   composite        ncclAllReduce(uint64, min) over width*height keys
//   reference table  ncclAllGather of each rank's MeshletReference section
//   shadow layers    ncclAllReduce(uint32, min) over the 4 depth layers
// NCCL is resolved at run time (dlsym in the process image first -- the
// torch-bundled NCCL when the host is Python -- then dlopen), so the library
// has no link-time NCCL dependency and single-GPU use never touches it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "lh_internal.cuh"

using namespace lh;

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
    bool tried = false;
};
NcclApi g_nccl;

bool load_nccl() {
    if (g_nccl.tried) return g_nccl.ok;
    g_nccl.tried = true;
    void* handles[3] = {RTLD_DEFAULT, nullptr, nullptr};
    auto find = [&](const char* name) -> void* {
        for (void* h : handles) {
            if (h == nullptr) continue;
            if (void* s = dlsym(h, name)) return s;
        }
        return nullptr;
    };
    if (!dlsym(RTLD_DEFAULT, "ncclAllReduce")) {
        handles[0] = nullptr;
        handles[1] = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!handles[1]) handles[1] = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!handles[1]) return false;
    }
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))find("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))find("ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))find("ncclCommDestroy");
    g_nccl.CommGetAsyncError = (decltype(g_nccl.CommGetAsyncError))find("ncclCommGetAsyncError");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))find("ncclAllReduce");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))find("ncclAllGather");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))find("ncclGetErrorString");
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.CommDestroy && g_nccl.AllReduce && g_nccl.AllGather;
    return g_nccl.ok;
}

int nccl_fail(LhContext* ctx, const char* what, ncclResult_t r) {
    std::string msg = what;
    msg += ": ";
    msg += g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "NCCL error";
    return fail(ctx, LH_ERR_NCCL, msg.c_str());
}

void close_peer_handles(LhContext* ctx, LhContext::PeerMap& m) {
    for (size_t q = 0; q < m.ptrs.size(); q++)
        if ((int)q != ctx->rank && m.ptrs[q]) cudaIpcCloseMemHandle(m.ptrs[q]);
    m.ptrs.clear();
    cudaFree(m.table_dev);
    m.table_dev = nullptr;
}

// Maps `local` (a cudaMalloc allocation of this rank) on every peer and the peers' counterparts here.  Collective:
// every rank must call it with its own counterpart at the same point of the stream.
int peer_map(LhContext* ctx, void* local, LhContext::PeerMap** out) {
    for (auto& m : ctx->peer_maps)
        if (m.local == local) {
            *out = &m;
            return LH_OK;
        }
    const int n = ctx->nranks;
    cudaIpcMemHandle_t mine;
    LH_CUDA(ctx, cudaIpcGetMemHandle(&mine, local));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle staging assumes 64-byte IPC handles");
    uint8_t* stage = (uint8_t*)(ctx->comm_scratch + 16);
    LH_CUDA(ctx, cudaMemcpyAsync(stage + 64 * (size_t)ctx->rank, &mine, 64, cudaMemcpyHostToDevice, ctx->stream));
    ncclResult_t r = g_nccl.AllGather(stage + 64 * (size_t)ctx->rank, stage, 64, ncclUint8, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclAllGather(ipc handles)", r);
    std::vector<cudaIpcMemHandle_t> all((size_t)n);
    LH_CUDA(ctx, cudaMemcpyAsync(all.data(), stage, 64 * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    LH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    LhContext::PeerMap m;
    m.local = local;
    m.ptrs.assign((size_t)n, nullptr);
    for (int q = 0; q < n; q++) {
        if (q == ctx->rank) {
            m.ptrs[q] = local;
            continue;
        }
        cudaError_t e = cudaIpcOpenMemHandle(&m.ptrs[q], all[q], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            close_peer_handles(ctx, m);
            return fail(ctx, LH_ERR_CUDA, "cudaIpcOpenMemHandle (peer buffer; LH_EXCHANGE_NCCL does not need peer memory)", e);
        }
    }
    cudaError_t e = cudaMalloc(&m.table_dev, sizeof(void*) * (size_t)n);
    if (e == cudaSuccess) e = cudaMemcpy(m.table_dev, m.ptrs.data(), sizeof(void*) * (size_t)n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        close_peer_handles(ctx, m);
        return fail(ctx, LH_ERR_CUDA, "peer pointer table", e);
    }
    ctx->peer_maps.push_back(std::move(m));
    *out = &ctx->peer_maps.back();
    return LH_OK;
}

void peer_unmap_all(LhContext* ctx) {
    for (auto& m : ctx->peer_maps) close_peer_handles(ctx, m);
    ctx->peer_maps.clear();
    ctx->peer_vis_dev = ctx->peer_refs_dev = nullptr;
    ctx->scene_remote = nullptr;
    ctx->mapped_refs = 0;
}

// screen rows of this rank: even row counts so that a slice is a whole number of 16-byte key pairs
void update_rows(LhContext* ctx) {
    const uint32_t H = ctx->cfg.height, n = (uint32_t)ctx->nranks, per = ((H + n - 1) / n + 1) & ~1u;
    ctx->row0 = per * (uint32_t)ctx->rank < H ? per * (uint32_t)ctx->rank : H;
    ctx->row1 = ctx->row0 + per < H ? ctx->row0 + per : H;
}

// A collective on 4 bytes: nothing after it on any rank's stream starts before everything ahead of it on every rank's
// stream has finished.
int stream_barrier(LhContext* ctx) {
    ncclResult_t r = g_nccl.AllReduce(ctx->comm_scratch + 2, ctx->comm_scratch + 2, 1, ncclUint32, ncclMin, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclAllReduce(barrier)", r);
    return LH_OK;
}

__global__ void depth_pair_init_kernel(uint32_t* pair) {
    if (threadIdx.x < 2) pair[threadIdx.x] = 0xFFFFFFFFu;   // min identity for both {min, ~max}
}

__global__ void depth_pair_store_kernel(const __grid_constant__ FrameParams P) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        LhMiscStorage* misc = (LhMiscStorage*)P.u.misc_storage;
        misc->min_depth = P.depth_pair[0];     // 0xFFFFFFFF / 0 when nothing was drawn, like read_depth.glsl
        misc->max_depth = ~P.depth_pair[1];
    }
}

// The sort-last composite of this rank's rows: min over every rank's keys, read through peer memory.
//
// A rank draws the instances of its screen stripe, so most of its key buffer is still clear when the frame is
// composited, and pulling every rank's copy of every row would move mostly all-ones words over NVLink (58 of 66 MB per
// rank on 8 GPUs).  Each rank therefore first marks which 64x16-pixel tiles of its own buffer hold anything
// (tile_flags_kernel: one streaming pass over its L2-resident keys, one byte per tile behind the keys, so the flags are
// peer-mapped with them); the composite then works tile by tile, asks every rank's flag, and reads only the ranks that
// drew there.  A tile nobody drew in is skipped outright; a tile only this rank drew in needs no store.  Two pixels per
// 128-bit load; the loads of all flagged ranks are issued before the first min so the NVLink round trips overlap.
constexpr int kMaxRanks = 16;
constexpr uint32_t kFlagTileW = 64, kFlagTileH = 16;

__device__ __forceinline__ const uint8_t* tile_flags_of(const unsigned long long* vis, uint64_t pixels) { return (const uint8_t*)(vis + pixels) + 16; }

__global__ void __launch_bounds__(256) tile_flags_kernel(const __grid_constant__ FrameParams P) {
    const uint32_t W = P.u.window_size[0], H = P.u.window_size[1];
    const uint32_t tiles_x = (W + kFlagTileW - 1) / kFlagTileW, tiles = tiles_x * ((H + kFlagTileH - 1) / kFlagTileH);
    const uint32_t tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (tile >= tiles) return;
    const uint32_t x = (tile % tiles_x) * kFlagTileW + 2 * lane, y0 = (tile / tiles_x) * kFlagTileH;
    bool any = false;
    if (x < W) {
#pragma unroll 4
        for (uint32_t r = 0; r < kFlagTileH; r++) {
            if (y0 + r >= H) break;
            const unsigned long long* at = P.vis + (size_t)(y0 + r) * W + x;
            if (x + 1 < W) {
                const ulonglong2 v = __ldcg((const ulonglong2*)at);
                any = any || (v.x & v.y) != ~0ull;
            } else {
                any = any || __ldcg(at) != ~0ull;
            }
        }
    }
    any = __any_sync(0xFFFFFFFFu, any);
    if (lane == 0) ((uint8_t*)tile_flags_of(P.vis, (uint64_t)W * H))[tile] = any ? 1 : 0;
}

// one CTA per tile that touches the rank's rows [row0, row1)
__global__ void __launch_bounds__(256) composite_pull_kernel(const __grid_constant__ FrameParams P, uint32_t row0, uint32_t row1) {
    const uint32_t n = P.nranks, self = P.rank;
    const uint32_t W = P.u.window_size[0], H = P.u.window_size[1];
    const uint64_t pixels = (uint64_t)W * H;
    const uint32_t tiles_x = (W + kFlagTileW - 1) / kFlagTileW;
    const uint32_t tx = blockIdx.x % tiles_x, ty = row0 / kFlagTileH + blockIdx.x / tiles_x, tile = ty * tiles_x + tx;
    __shared__ uint32_t drew;
    if (threadIdx.x < 32) {
        const bool flag = threadIdx.x < n && __ldcg(tile_flags_of(P.peer_vis[threadIdx.x], pixels) + tile) != 0;
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, flag);
        if (threadIdx.x == 0) drew = mask;
    }
    __syncthreads();
    const uint32_t mask = drew;
    if (mask == 0) return;   // clear on every rank: sky, which moves neither depth bound
    uint32_t mn = 0xFFFFFFFFu, inv_mx = 0xFFFFFFFFu;
#pragma unroll
    for (uint32_t j = 0; j < kFlagTileW * kFlagTileH / 2 / 256; j++) {
        const uint32_t item = threadIdx.x + 256 * j;
        const uint32_t x = tx * kFlagTileW + 2 * (item % (kFlagTileW / 2)), y = ty * kFlagTileH + item / (kFlagTileW / 2);
        if (x >= W || y < row0 || y >= row1) continue;   // W is even (checked by the caller): a pair never straddles a row
        const uint64_t pair = ((uint64_t)y * W + x) / 2;
        ulonglong2 v[kMaxRanks];
#pragma unroll
        for (uint32_t q = 0; q < (uint32_t)kMaxRanks; q++)
            if ((mask >> q) & 1u) v[q] = __ldcg((const ulonglong2*)P.peer_vis[q] + pair);
        ulonglong2 best = make_ulonglong2(~0ull, ~0ull);
#pragma unroll
        for (uint32_t q = 0; q < (uint32_t)kMaxRanks; q++)
            if ((mask >> q) & 1u) {
                best.x = min(best.x, v[q].x);
                best.y = min(best.y, v[q].y);
            }
        if (mask != (1u << self)) ((ulonglong2*)P.peer_vis[self])[pair] = best;
        const uint32_t d0 = ~(uint32_t)(best.x >> 32), d1 = ~(uint32_t)(best.y >> 32);
        if (d0) mn = min(mn, d0);                 // min_if_not_zero, read_depth.glsl:3-7
        if (d1) mn = min(mn, d1);
        inv_mx = min(inv_mx, min(~d0, ~d1));
    }
    mn = __reduce_min_sync(0xFFFFFFFFu, mn);
    inv_mx = __reduce_min_sync(0xFFFFFFFFu, inv_mx);
    if ((threadIdx.x & 31) == 0) {
        if (mn != 0xFFFFFFFFu) atomicMin(P.depth_pair, mn);
        if (inv_mx != 0xFFFFFFFFu) atomicMin(P.depth_pair + 1, inv_mx);
    }
}

}  // namespace

namespace lh {

bool peer_exchange(const LhContext* ctx) { return ctx->nccl_comm && ctx->nranks > 1 && ctx->exchange == LH_EXCHANGE_PEER; }

// Called by lh_frame_begin.  The first frame (and any frame after the host moved its MeshletReference buffer or
// resized the targets -- which all ranks must do together) exchanges IPC handles; afterwards this is three lookups.
int ensure_peer_maps(LhContext* ctx) {
    ctx->peer_vis_dev = ctx->peer_refs_dev = nullptr;
    ctx->scene_remote = nullptr;
    if (!peer_exchange(ctx)) return LH_OK;
    // The handle exchange is a collective, so WHETHER it runs must not depend on rank-local state such as the address
    // cudaMalloc happened to return: buffers are mapped once after lh_comm_init / lh_targets_resize / lh_comm_remap (all
    // of which every rank calls together and which drop every mapping), and a host that moves its MeshletReference
    // buffer in between is told so instead of deadlocking its peers.
    if (ctx->mapped_refs && ctx->mapped_refs != ctx->uniforms.meshlet_references)
        return fail(ctx, LH_ERR_NOT_READY, "LhUniforms.meshlet_references moved while peers map the old buffer: call lh_comm_remap on every rank");
    LhContext::PeerMap* m = nullptr;
    if (int rc = peer_map(ctx, ctx->vis, &m)) return rc;
    ctx->peer_vis_dev = m->table_dev;
    if (int rc = peer_map(ctx, (void*)(uintptr_t)ctx->uniforms.meshlet_references, &m)) return rc;
    ctx->peer_refs_dev = m->table_dev;
    ctx->mapped_refs = ctx->uniforms.meshlet_references;
    if (int rc = peer_map(ctx, ctx->scene, &m)) return rc;
    ctx->scene_remote = ctx->rank == 0 ? nullptr : m->ptrs[0];
    return LH_OK;
}

int composite_pull(LhContext* ctx) {
    if (ctx->nranks > kMaxRanks) return fail(ctx, LH_ERR_UNSUPPORTED, "LH_EXCHANGE_PEER supports at most 16 ranks");
    FrameParams p = make_params(ctx);
    depth_pair_init_kernel<<<1, 32, 0, ctx->stream>>>(ctx->comm_scratch); ctx->launches++;
    const uint32_t W = ctx->cfg.width, H = ctx->cfg.height;
    // the keys are read as 16-byte pairs that must not straddle rows, and row slices hold an even number of rows
    if (W & 1u) return fail(ctx, LH_ERR_UNSUPPORTED, "LH_EXCHANGE_PEER needs an even frame width");
    const uint32_t tiles_x = (W + kFlagTileW - 1) / kFlagTileW, tiles = tiles_x * ((H + kFlagTileH - 1) / kFlagTileH);
    tile_flags_kernel<<<(tiles + 7) / 8, 256, 0, ctx->stream>>>(p); ctx->launches++;
    if (int rc = stream_barrier(ctx)) return rc;   // every rank has rasterised its instances and flagged its tiles
    if (ctx->row1 > ctx->row0) {
        const uint32_t tile_rows = (ctx->row1 - 1) / kFlagTileH - ctx->row0 / kFlagTileH + 1;
        composite_pull_kernel<<<tiles_x * tile_rows, 256, 0, ctx->stream>>>(p, ctx->row0, ctx->row1); ctx->launches++;
    }
    return LH_OK;
}

int reduce_depth_exchange(LhContext* ctx) {
    ncclResult_t r = g_nccl.AllReduce(ctx->comm_scratch, ctx->comm_scratch, 2, ncclUint32, ncclMin, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclAllReduce(depth bounds)", r);
    depth_pair_store_kernel<<<1, 32, 0, ctx->stream>>>(make_params(ctx)); ctx->launches++;
    return LH_OK;
}

int frame_barrier(LhContext* ctx) { return stream_barrier(ctx); }

// an asynchronous NCCL failure (a peer died, a link went down) surfaces here instead of as a hang at the next collective
int comm_poll_errors(LhContext* ctx) {
    if (!ctx->nccl_comm || !g_nccl.CommGetAsyncError) return LH_OK;
    ncclResult_t async = ncclSuccess;
    ncclResult_t r = g_nccl.CommGetAsyncError((ncclComm_t)ctx->nccl_comm, &async);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclCommGetAsyncError", r);
    if (async != ncclSuccess && async != ncclInProgress) return nccl_fail(ctx, "asynchronous NCCL error", async);
    return LH_OK;
}

// Collective: every rank drops its mappings of everybody's buffers; when it returns no rank maps any peer buffer any
// more, so targets may be freed / reallocated and the MeshletReference buffer moved.
int peer_release_collective(LhContext* ctx) {
    if (!ctx->nccl_comm || ctx->nranks <= 1) return LH_OK;
    LH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    peer_unmap_all(ctx);
    if (int rc = stream_barrier(ctx)) return rc;
    LH_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return LH_OK;
}

void comm_update_rows(LhContext* ctx) {
    if (ctx->nccl_comm) update_rows(ctx);
}

int composite_allreduce_min(LhContext* ctx) {
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const size_t pixels = (size_t)ctx->cfg.width * ctx->cfg.height;
    ncclResult_t r = g_nccl.AllReduce(ctx->vis, ctx->vis, pixels, ncclUint64, ncclMin, comm, ctx->stream);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclAllReduce(vis, uint64, min)", r);
    // every rank needs the (instance, meshlet) of any slot that can win a pixel it resolves
    const void* local_refs = (const void*)(uintptr_t)ctx->uniforms.meshlet_references;
    if (!ctx->gathered_refs) {   // lh_comm_set_exchange before lh_comm_init, or a re-initialised communicator
        const size_t bytes = (size_t)ctx->slots_per_rank * ctx->nranks * sizeof(LhMeshletReference);
        LH_CUDA(ctx, cudaMalloc(&ctx->gathered_refs, bytes));
        LH_CUDA(ctx, cudaMemsetAsync(ctx->gathered_refs, 0, bytes, ctx->stream));
    }
    r = g_nccl.AllGather(local_refs, ctx->gathered_refs, (size_t)ctx->slots_per_rank * sizeof(LhMeshletReference), ncclUint8, comm, ctx->stream);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclAllGather(meshlet references)", r);
    return LH_OK;
}

}  // namespace lh

extern "C" {

int lh_comm_unique_id(void* id128) {
    if (!id128) return LH_ERR_INVALID_ARGUMENT;
    if (!load_nccl()) return fail(nullptr, LH_ERR_NCCL, "NCCL library not found");
    ncclUniqueId id;
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return nccl_fail(nullptr, "ncclGetUniqueId", r);
    memcpy(id128, &id, sizeof(id));
    return LH_OK;
}

int lh_comm_init(LhContext* ctx, int rank, int nranks, const void* id128) {
    if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks) return LH_ERR_INVALID_ARGUMENT;
    if (!load_nccl()) return fail(ctx, LH_ERR_NCCL, "NCCL library not found");
    const uint64_t per_rank = (uint64_t)ctx->cfg.max_opaque_draws + ctx->cfg.max_alpha_clip_draws;
    if (per_rank * nranks > LH_VIS_MAX_SLOTS)
        return fail(ctx, LH_ERR_INVALID_ARGUMENT, "nranks * (max_opaque_draws + max_alpha_clip_draws) exceeds the 25-bit slot field");
    cudaSetDevice(ctx->device);
    if (ctx->nccl_comm) lh_comm_destroy(ctx);   // a second init replaces the communicator (and everything sized by the old one)
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclComm_t comm = nullptr;
    ncclResult_t r = g_nccl.CommInitRank(&comm, nranks, id, rank);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclCommInitRank", r);
    ctx->nccl_comm = comm;
    ctx->rank = rank;
    ctx->nranks = nranks;
    ctx->slots_per_rank = (uint32_t)per_rank;
    ctx->slot_offset = (uint32_t)(per_rank * rank);
    cudaFree(ctx->comm_scratch);
    ctx->comm_scratch = nullptr;
    LH_CUDA(ctx, cudaMalloc(&ctx->comm_scratch, 64 + 64 * (size_t)nranks));
    LH_CUDA(ctx, cudaMemset(ctx->comm_scratch, 0, 64 + 64 * (size_t)nranks));
    update_rows(ctx);
    return LH_OK;
}

int lh_comm_set_exchange(LhContext* ctx, int mode) {
    if (!ctx || (mode != LH_EXCHANGE_PEER && mode != LH_EXCHANGE_NCCL)) return LH_ERR_INVALID_ARGUMENT;
    ctx->exchange = mode;
    if (mode == LH_EXCHANGE_NCCL && ctx->nranks > 1 && !ctx->gathered_refs) {
        const size_t bytes = (size_t)ctx->slots_per_rank * ctx->nranks * sizeof(LhMeshletReference);
        LH_CUDA(ctx, cudaMalloc(&ctx->gathered_refs, bytes));
        LH_CUDA(ctx, cudaMemset(ctx->gathered_refs, 0, bytes));
    }
    return LH_OK;
}

int lh_comm_rows(LhContext* ctx, uint32_t* y0, uint32_t* y1) {
    if (!ctx) return LH_ERR_INVALID_ARGUMENT;
    const bool split = peer_exchange(ctx);
    if (y0) *y0 = split ? ctx->row0 : 0u;
    if (y1) *y1 = split ? ctx->row1 : ctx->cfg.height;
    return LH_OK;
}

int lh_comm_set_rows(LhContext* ctx, uint32_t y0, uint32_t y1) {
    if (!ctx) return LH_ERR_INVALID_ARGUMENT;
    if (!ctx->nccl_comm) return fail(ctx, LH_ERR_NOT_READY, "lh_comm_set_rows before lh_comm_init");
    if (y0 > y1 || y1 > ctx->cfg.height) return fail(ctx, LH_ERR_INVALID_ARGUMENT, "lh_comm_set_rows: need y0 <= y1 <= height");
    ctx->row0 = y0;
    ctx->row1 = y1;
    return LH_OK;
}

int lh_comm_destroy(LhContext* ctx) {
    if (!ctx) return LH_ERR_INVALID_ARGUMENT;
    if (ctx->nccl_comm && g_nccl.CommDestroy) {
        // unmap -> barrier -> destroy: no rank frees a buffer a peer still maps (ranks destroy their communicators together)
        if (peer_exchange(ctx) && !ctx->peer_maps.empty()) peer_release_collective(ctx);
        else cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    }
    cudaFree(ctx->comm_scratch);
    ctx->comm_scratch = nullptr;
    cudaFree(ctx->gathered_refs);
    ctx->gathered_refs = nullptr;
    ctx->nccl_comm = nullptr;
    ctx->rank = 0;
    ctx->nranks = 1;
    ctx->slot_offset = 0;
    return LH_OK;
}

int lh_comm_remap(LhContext* ctx) {
    if (!ctx) return LH_ERR_INVALID_ARGUMENT;
    return peer_release_collective(ctx);   // the next lh_frame_begin maps the current buffers again (collectively)
}

int lh_comm_info(LhContext* ctx, int* rank, int* nranks) {
    if (!ctx) return LH_ERR_INVALID_ARGUMENT;
    if (rank) *rank = ctx->rank;
    if (nranks) *nranks = ctx->nranks;
    return LH_OK;
}

// Shadow layers rendered by different ranks (one cascade per GPU, or instance
// partitions of one cascade) are merged with a uint32 min: unrendered layers
// hold the 1.0 clear, and LESS-tested depth is order-independent.
int lh_shadow_exchange(LhContext* ctx) {
    if (!ctx) return LH_ERR_INVALID_ARGUMENT;
    if (!ctx->nccl_comm || ctx->nranks <= 1) return LH_OK;
    // layers this rank did not render must hold the clear value before the reduction
    const size_t texels = (size_t)ctx->cfg.shadowmap_size * ctx->cfg.shadowmap_size;
    for (uint32_t c = 0; c < 4; c++)
        if (!((ctx->cascade_mask >> c) & 1u)) launch_fill_u32(ctx, ctx->shadow + texels * c, texels, 0x3F800000u);
    ncclResult_t r = g_nccl.AllReduce(ctx->shadow, ctx->shadow, texels * 4, ncclUint32, ncclMin, (ncclComm_t)ctx->nccl_comm, ctx->stream);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclAllReduce(shadow, uint32, min)", r);
    return LH_OK;
}

}  // extern "C"
