// The reference's on-disk meshlet cache: FsCache (src/resources/fs_cache.h:3-25) keeps, per glTF
// primitive, three raw blobs under cache/<decimal std::hash of the key>: "<name> meshlets" (Meshlet[], 56 B each),
// "<name> indices" (the per-meshlet vertex indices, u16 or u32) and "<name> micro indices" (u8 triples, each meshlet's
// run padded to 4 bytes) -- written after build_meshlets and read back instead of rebuilding
// (src/resources/mesh_loading.cpp:602-649).  Same keys, same file names (std::hash<std::string> of the same standard
// library), same bytes; unlike the reference a short or oversized blob is an error, not undefined behaviour.
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <string>
#include <vector>

#include "lh_scene.h"

namespace {

std::filesystem::path blob_path(const char* dir, const std::string& key) {
    return std::filesystem::path(dir) / std::to_string(std::hash<std::string>{}(key));
}

bool write_blob(const std::filesystem::path& p, const void* data, uint64_t bytes) {
    std::ofstream s(p, std::ios::binary | std::ios::trunc);
    if (!s) return false;
    s.write((const char*)data, (std::streamsize)bytes);
    return (bool)s;
}

bool read_blob(const std::filesystem::path& p, std::vector<uint8_t>& out) {
    std::error_code ec;
    const auto size = std::filesystem::file_size(p, ec);
    if (ec) return false;
    out.resize((size_t)size);
    std::ifstream s(p, std::ios::binary);
    if (!s) return false;
    s.read((char*)out.data(), (std::streamsize)size);
    return (bool)s || size == 0;
}

}  // namespace

struct LhMeshletBlobs {
    std::vector<uint8_t> meshlets, indices, micro;
};

extern "C" {

int lh_meshlet_cache_path(const char* dir, const char* primitive_name, int which, char* out, uint64_t out_bytes) {
    static const char* suffix[3] = {" meshlets", " indices", " micro indices"};
    if (!dir || !primitive_name || !out || which < 0 || which > 2) return -1;
    const std::string p = blob_path(dir, std::string(primitive_name) + suffix[which]).string();
    if (p.size() + 1 > out_bytes) return -1;
    memcpy(out, p.c_str(), p.size() + 1);
    return 0;
}

int lh_meshlet_cache_store(const char* dir, const char* primitive_name, const LhMeshView* mesh) {
    if (!dir || !primitive_name || !mesh) return -1;
    std::error_code ec;
    std::filesystem::create_directories(dir, ec);
    const std::string name(primitive_name);
    bool ok = write_blob(blob_path(dir, name + " meshlets"), mesh->meshlets, mesh->num_meshlets * sizeof(LhMeshlet));
    ok = ok && write_blob(blob_path(dir, name + " micro indices"), mesh->micro_indices, mesh->micro_indices_bytes);
    ok = ok && write_blob(blob_path(dir, name + " indices"), mesh->indices, mesh->num_indices * mesh->index_bytes);
    return ok ? 0 : -2;
}

int lh_meshlet_cache_load(const char* dir, const char* primitive_name, uint32_t index_bytes, LhMeshletBlobs** out) {
    if (!dir || !primitive_name || !out || (index_bytes != 2 && index_bytes != 4)) return -1;
    const std::string name(primitive_name);
    auto blobs = new LhMeshletBlobs();
    // all three or nothing, as mesh_loading.cpp:612 ("if (opt_meshlets && opt_indices && opt_micro_indices)")
    if (!read_blob(blob_path(dir, name + " meshlets"), blobs->meshlets) || !read_blob(blob_path(dir, name + " indices"), blobs->indices) ||
        !read_blob(blob_path(dir, name + " micro indices"), blobs->micro)) {
        delete blobs;
        return 1;   // a miss: the caller builds the meshlets and stores them
    }
    // consistency the reference never checks: whole records, and every meshlet's ranges inside the two index blobs
    bool sane = blobs->meshlets.size() % sizeof(LhMeshlet) == 0 && blobs->indices.size() % index_bytes == 0;
    const uint64_t n = blobs->meshlets.size() / sizeof(LhMeshlet), n_idx = blobs->indices.size() / index_bytes;
    for (uint64_t i = 0; sane && i < n; i++) {
        LhMeshlet m;
        memcpy(&m, blobs->meshlets.data() + i * sizeof(LhMeshlet), sizeof(m));
        sane = m.triangle_count <= LH_MAX_MESHLET_TRIANGLES && m.index_count <= LH_MAX_MESHLET_UNIQUE_VERTICES &&
               (uint64_t)m.index_offset + m.index_count <= n_idx && (uint64_t)m.triangle_offset + 3ull * m.triangle_count <= blobs->micro.size();
        // every micro index names one of the meshlet's own vertices (the kernels index per-meshlet arrays with them)
        for (uint32_t k = 0; sane && k < 3u * m.triangle_count; k++) sane = blobs->micro[m.triangle_offset + k] < m.index_count;
    }
    if (!sane) {
        delete blobs;
        return -3;
    }
    *out = blobs;
    return 0;
}

int lh_meshlet_blobs_view(const LhMeshletBlobs* blobs, uint32_t index_bytes, LhMeshView* view) {
    if (!blobs || !view || (index_bytes != 2 && index_bytes != 4)) return -1;
    view->meshlets = (const LhMeshlet*)blobs->meshlets.data();
    view->num_meshlets = blobs->meshlets.size() / sizeof(LhMeshlet);
    view->indices = blobs->indices.data();
    view->num_indices = blobs->indices.size() / index_bytes;
    view->index_bytes = index_bytes;
    view->micro_indices = blobs->micro.data();
    view->micro_indices_bytes = blobs->micro.size();
    uint64_t tris = 0;
    for (uint64_t i = 0; i < view->num_meshlets; i++) tris += view->meshlets[i].triangle_count;
    view->num_triangles = tris;
    return 0;
}

void lh_meshlet_blobs_free(LhMeshletBlobs* blobs) { delete blobs; }

}  // extern "C"
