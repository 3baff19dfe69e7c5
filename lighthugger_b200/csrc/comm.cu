// Sort-last multi-GPU exchange (new; the reference is single-GPU).
//
// One process / context per GPU.  Instances are partitioned across ranks; each
// rank rasterises its partition into a full-resolution packed visibility
// buffer whose keys carry GLOBAL slots (rank * slots_per_rank + local), so the
// per-pixel uint64 MIN over ranks is exactly the single-GPU buffer.  Exchange
// steps, all on the context stream:
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

}  // namespace

namespace lh {

int composite_allreduce_min(LhContext* ctx) {
    ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
    const size_t pixels = (size_t)ctx->cfg.width * ctx->cfg.height;
    ncclResult_t r = g_nccl.AllReduce(ctx->vis, ctx->vis, pixels, ncclUint64, ncclMin, comm, ctx->stream);
    if (r != ncclSuccess) return nccl_fail(ctx, "ncclAllReduce(vis, uint64, min)", r);
    // every rank needs the (instance, meshlet) of any slot that can win a pixel it resolves
    const void* local_refs = (const void*)(uintptr_t)ctx->uniforms.meshlet_references;
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
    if (nranks > 1) {
        cudaFree(ctx->gathered_refs);
        ctx->gathered_refs = nullptr;
        LH_CUDA(ctx, cudaMalloc(&ctx->gathered_refs, per_rank * nranks * sizeof(LhMeshletReference)));
        LH_CUDA(ctx, cudaMemset(ctx->gathered_refs, 0, per_rank * nranks * sizeof(LhMeshletReference)));
    }
    return LH_OK;
}

int lh_comm_destroy(LhContext* ctx) {
    if (!ctx) return LH_ERR_INVALID_ARGUMENT;
    if (ctx->nccl_comm && g_nccl.CommDestroy) {
        cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy((ncclComm_t)ctx->nccl_comm);
    }
    ctx->nccl_comm = nullptr;
    ctx->rank = 0;
    ctx->nranks = 1;
    ctx->slot_offset = 0;
    return LH_OK;
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
