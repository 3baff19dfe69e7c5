// DDS (DX10) and KTX2 containers -> level layout + payload, for the textures the lighting resolve samples.
// Follows what the reference accepts (src/resources/image_loading.cpp:14-39 formats, :67-217 DDS, :242-330 KTX2;
// src/resources/dds.h:6-175, src/resources/ktx2.h:1-36) but is written against a memory image of the file and
// reports malformed input instead of aborting.  See include/lh_image_files.h.
#include "lh_image_files.h"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>

namespace {

// Byte-exact field readers: the file is not assumed to be aligned.
struct Reader {
    const uint8_t* base;
    uint64_t size;
    bool ok = true;
    template <typename T>
    T at(uint64_t offset) {
        T v{};
        if (offset > size || sizeof(T) > size - offset) {
            ok = false;
            return v;
        }
        memcpy(&v, base + offset, sizeof(T));
        return v;
    }
};

// ---- DDS: "DDS " + DDS_HEADER (124 bytes) + DDS_HEADER_DXT10 (20 bytes) ----
constexpr uint64_t kDdsMagicBytes = 4, kDdsHeaderBytes = 124, kDdsHeader10Bytes = 20;
constexpr uint64_t kDdsHeight = 4 + 8, kDdsWidth = 4 + 12, kDdsDepth = 4 + 20, kDdsMipCount = 4 + 24, kDdsCaps2 = 4 + 108;
constexpr uint64_t kDds10Format = 4 + 124 + 0, kDds10Dimension = 4 + 124 + 4;
constexpr uint32_t kDdsCaps2Cubemap = 0x200;
constexpr uint32_t kDxgiR9G9B9E5 = 67, kDxgiBc1Unorm = 71, kDxgiBc1UnormSrgb = 72, kDxgiBc7Unorm = 98, kDxgiBc7UnormSrgb = 99;
constexpr uint32_t kDxgiRgba8Unorm = 28, kDxgiRgba8UnormSrgb = 29;
constexpr uint32_t kD3d10Texture2d = 3, kD3d10Texture3d = 4;
constexpr uint32_t kMaxDimension = 65536;   // per axis; larger headers are rejected before any size arithmetic

bool dxgi_format(uint32_t dxgi, LhImageInfo* info) {
    switch (dxgi) {
    case kDxgiR9G9B9E5: info->format = LH_PIXEL_RGB9E5; info->bits_per_pixel = 32; info->block_compressed = 0; return true;
    // every block-compressed DDS is created as an sRGB image, UNORM or not (image_loading.cpp:20-34)
    case kDxgiBc1Unorm:
    case kDxgiBc1UnormSrgb: info->format = LH_PIXEL_BC1_SRGB; info->bits_per_pixel = 4; info->block_compressed = 1; return true;
    case kDxgiBc7Unorm:
    case kDxgiBc7UnormSrgb: info->format = LH_PIXEL_BC7_SRGB; info->bits_per_pixel = 8; info->block_compressed = 1; return true;
    // not produced by the reference's asset pipeline, but what an uncompressed export of the same textures holds
    case kDxgiRgba8Unorm: info->format = LH_PIXEL_RGBA8_UNORM; info->bits_per_pixel = 32; info->block_compressed = 0; return true;
    case kDxgiRgba8UnormSrgb: info->format = LH_PIXEL_RGBA8_SRGB; info->bits_per_pixel = 32; info->block_compressed = 0; return true;
    default: return false;
    }
}

uint32_t round_up4(uint32_t v) { return (v + 3u) & ~3u; }

int parse_dds(Reader& r, LhImageInfo* info) {
    const uint64_t data_offset = kDdsMagicBytes + kDdsHeaderBytes + kDdsHeader10Bytes;
    if (r.size < data_offset) return LH_IMAGE_ERR_TRUNCATED;
    info->container = LH_IMAGE_CONTAINER_DDS;
    info->height = r.at<uint32_t>(kDdsHeight);
    info->width = r.at<uint32_t>(kDdsWidth);
    info->depth = std::max(r.at<uint32_t>(kDdsDepth), 1u);
    info->levels = std::max(r.at<uint32_t>(kDdsMipCount), 1u);   // writers leave 0 for "no mip chain"
    info->faces = (r.at<uint32_t>(kDdsCaps2) & kDdsCaps2Cubemap) ? 6u : 1u;
    info->source_format = (int)r.at<uint32_t>(kDds10Format);
    const uint32_t dimension = r.at<uint32_t>(kDds10Dimension);
    if (!dxgi_format((uint32_t)info->source_format, info)) return LH_IMAGE_ERR_FORMAT;
    if (dimension == kD3d10Texture2d) info->dimension = 2;
    else if (dimension == kD3d10Texture3d) info->dimension = 3;
    else return LH_IMAGE_ERR_FORMAT;
    if (info->width == 0 || info->height == 0 || info->levels > LH_IMAGE_MAX_LEVELS) return LH_IMAGE_ERR_FORMAT;
    // dimensions a GPU can bind; also keeps the 64-bit size products below from wrapping (four 32-bit header fields)
    if (info->width > kMaxDimension || info->height > kMaxDimension || info->depth > kMaxDimension || info->faces > 6) return LH_IMAGE_ERR_FORMAT;
    uint64_t offset = data_offset;
    for (uint32_t l = 0; l < info->levels; l++) {
        LhImageLevel& level = info->levels_info[l];
        level.width = std::max(info->width >> l, 1u);
        level.height = std::max(info->height >> l, 1u);
        // a block-compressed level holds at least one 4x4 block per axis (image_loading.cpp:196-211); like the
        // reference, the depth is NOT reduced per level
        const uint64_t w = info->block_compressed ? round_up4(level.width) : level.width;
        const uint64_t h = info->block_compressed ? round_up4(level.height) : level.height;
        level.bytes = w * h * info->depth * info->faces * info->bits_per_pixel / 8;
        level.stored_bytes = level.bytes;
        level.offset = offset;
        offset += level.bytes;
        info->decoded_bytes += level.bytes;
    }
    // "buffer_offset == bytes_remaining" (image_loading.cpp:214-217)
    if (offset > r.size) return LH_IMAGE_ERR_TRUNCATED;
    if (offset != r.size) return LH_IMAGE_ERR_SIZE;
    return LH_IMAGE_OK;
}

// ---- KTX2: 12-byte identifier, header (9 x u32), index (4 x u32 + 2 x u64), level index (3 x u64 per level) ----
const uint8_t kKtx2Identifier[12] = {0xAB, 0x4B, 0x54, 0x58, 0x20, 0x32, 0x30, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A};
constexpr uint64_t kKtx2Header = 12, kKtx2Index = 12 + 36, kKtx2Levels = 12 + 36 + 32;
// VkFormat values (vulkan_core.h) of the formats this path samples
constexpr uint32_t kVkRgba8Unorm = 37, kVkRgba8Srgb = 43, kVkE5B9G9R9 = 123, kVkBc1RgbSrgb = 132, kVkBc1RgbaSrgb = 134, kVkBc7Srgb = 146;
constexpr uint32_t kVkBc1RgbUnorm = 131, kVkBc1RgbaUnorm = 133, kVkBc7Unorm = 145;

bool vk_format(uint32_t vk, LhImageInfo* info) {
    switch (vk) {
    case kVkRgba8Unorm: info->format = LH_PIXEL_RGBA8_UNORM; info->bits_per_pixel = 32; info->block_compressed = 0; return true;
    case kVkRgba8Srgb: info->format = LH_PIXEL_RGBA8_SRGB; info->bits_per_pixel = 32; info->block_compressed = 0; return true;
    case kVkE5B9G9R9: info->format = LH_PIXEL_RGB9E5; info->bits_per_pixel = 32; info->block_compressed = 0; return true;
    case kVkBc1RgbSrgb:
    case kVkBc1RgbaSrgb: info->format = LH_PIXEL_BC1_SRGB; info->bits_per_pixel = 4; info->block_compressed = 1; return true;
    case kVkBc7Srgb: info->format = LH_PIXEL_BC7_SRGB; info->bits_per_pixel = 8; info->block_compressed = 1; return true;
    // the KTX2 path passes header.format straight to the image (image_loading.cpp:293): a UNORM block format would be
    // sampled without the sRGB decode, which the CUDA layer's texture formats do not offer -- reported, not guessed
    case kVkBc1RgbUnorm:
    case kVkBc1RgbaUnorm:
    case kVkBc7Unorm:
    default: return false;
    }
}

int parse_ktx2(Reader& r, LhImageInfo* info) {
    if (r.size < kKtx2Levels) return LH_IMAGE_ERR_TRUNCATED;
    info->container = LH_IMAGE_CONTAINER_KTX2;
    info->source_format = (int)r.at<uint32_t>(kKtx2Header + 0);
    info->width = r.at<uint32_t>(kKtx2Header + 8);
    info->height = std::max(r.at<uint32_t>(kKtx2Header + 12), 1u);
    info->depth = std::max(r.at<uint32_t>(kKtx2Header + 16), 1u);
    info->faces = std::max(r.at<uint32_t>(kKtx2Header + 24), 1u);
    info->levels = std::max(r.at<uint32_t>(kKtx2Header + 28), 1u);
    const uint32_t scheme = r.at<uint32_t>(kKtx2Header + 32);
    if (!vk_format((uint32_t)info->source_format, info)) return LH_IMAGE_ERR_FORMAT;
    if (scheme != 0 && scheme != 2) return LH_IMAGE_ERR_FORMAT;   // None / Zstandard (image_loading.cpp:339-365)
    info->zstd = scheme == 2;
    info->dimension = info->depth > 1 ? 3 : 2;
    if (info->width == 0 || info->levels > LH_IMAGE_MAX_LEVELS || (info->faces != 1 && info->faces != 6)) return LH_IMAGE_ERR_FORMAT;
    if (r.size < kKtx2Levels + 24ull * info->levels) return LH_IMAGE_ERR_TRUNCATED;
    if (info->width > kMaxDimension || info->height > kMaxDimension || info->depth > kMaxDimension || info->faces > 6) return LH_IMAGE_ERR_FORMAT;
    for (uint32_t l = 0; l < info->levels; l++) {
        LhImageLevel& level = info->levels_info[l];
        level.offset = r.at<uint64_t>(kKtx2Levels + 24ull * l);
        level.stored_bytes = r.at<uint64_t>(kKtx2Levels + 24ull * l + 8);
        level.bytes = r.at<uint64_t>(kKtx2Levels + 24ull * l + 16);
        level.width = std::max(info->width >> l, 1u);
        level.height = std::max(info->height >> l, 1u);
        if (level.offset > r.size || level.stored_bytes > r.size - level.offset) return LH_IMAGE_ERR_TRUNCATED;
        if (!info->zstd && level.stored_bytes != level.bytes) return LH_IMAGE_ERR_SIZE;
        // the level index must agree with the dimensions: the CUDA upload sizes its copies from the latter
        const uint64_t w = info->block_compressed ? round_up4(level.width) : level.width;
        const uint64_t h = info->block_compressed ? round_up4(level.height) : level.height;
        const uint64_t d = std::max(info->depth >> l, 1u);
        if (level.bytes != w * h * d * info->faces * info->bits_per_pixel / 8) return LH_IMAGE_ERR_SIZE;
        info->decoded_bytes += level.bytes;
    }
    return LH_IMAGE_OK;
}

// ZSTD_decompress / ZSTD_isError from libzstd.so.1, resolved on first use (the image has the runtime, not the header)
using ZstdDecompress = size_t (*)(void*, size_t, const void*, size_t);
using ZstdIsError = unsigned (*)(size_t);
ZstdDecompress g_zstd_decompress = nullptr;
ZstdIsError g_zstd_is_error = nullptr;
std::once_flag g_zstd_once;

bool load_zstd() {
    std::call_once(g_zstd_once, [] {
        void* h = dlopen("libzstd.so.1", RTLD_NOW);
        if (!h) h = dlopen("libzstd.so", RTLD_NOW);
        if (!h) return;
        g_zstd_decompress = (ZstdDecompress)dlsym(h, "ZSTD_decompress");
        g_zstd_is_error = (ZstdIsError)dlsym(h, "ZSTD_isError");
    });
    return g_zstd_decompress && g_zstd_is_error;
}

}  // namespace

extern "C" {

int lh_image_parse(const void* file, uint64_t file_bytes, LhImageInfo* out) {
    if (!file || !out) return LH_IMAGE_ERR_ARGUMENT;
    memset(out, 0, sizeof(*out));
    Reader r{(const uint8_t*)file, file_bytes};
    int rc = LH_IMAGE_ERR_MAGIC;
    if (file_bytes >= 4 && memcmp(file, "DDS ", 4) == 0) rc = parse_dds(r, out);
    else if (file_bytes >= 12 && memcmp(file, kKtx2Identifier, 12) == 0) rc = parse_ktx2(r, out);
    if (rc == LH_IMAGE_OK && !r.ok) rc = LH_IMAGE_ERR_TRUNCATED;
    return rc;
}

int lh_image_decode(const void* file, uint64_t file_bytes, const LhImageInfo* info, void* dst, uint64_t dst_bytes) {
    if (!file || !info || !dst || info->levels == 0 || info->levels > LH_IMAGE_MAX_LEVELS) return LH_IMAGE_ERR_ARGUMENT;
    if (dst_bytes < info->decoded_bytes) return LH_IMAGE_ERR_ARGUMENT;
    uint8_t* out = (uint8_t*)dst;
    for (uint32_t l = 0; l < info->levels; l++) {
        const LhImageLevel& level = info->levels_info[l];
        if (level.offset > file_bytes || level.stored_bytes > file_bytes - level.offset) return LH_IMAGE_ERR_TRUNCATED;
        const uint8_t* src = (const uint8_t*)file + level.offset;
        if (info->zstd) {
            if (!load_zstd()) return LH_IMAGE_ERR_ZSTD;
            const size_t n = g_zstd_decompress(out, (size_t)level.bytes, src, (size_t)level.stored_bytes);
            if (g_zstd_is_error(n) || n != level.bytes) return LH_IMAGE_ERR_ZSTD;   // image_loading.cpp:353
        } else {
            memcpy(out, src, (size_t)level.bytes);
        }
        out += level.bytes;
    }
    return LH_IMAGE_OK;
}

void lh_rgb9e5_decode(const uint32_t* texels, uint64_t count, float* rgb) {
    // value = mantissa * 2^(exponent - 15 - 9): 9-bit mantissas without a hidden bit, 5-bit shared exponent, bias 15
    for (uint64_t i = 0; i < count; i++) {
        const uint32_t t = texels[i];
        const float scale = std::ldexp(1.0f, (int)(t >> 27) - 15 - 9);
        rgb[3 * i + 0] = (float)(t & 0x1FFu) * scale;
        rgb[3 * i + 1] = (float)((t >> 9) & 0x1FFu) * scale;
        rgb[3 * i + 2] = (float)((t >> 18) & 0x1FFu) * scale;
    }
}

const char* lh_image_error_string(int code) {
    switch (code) {
    case LH_IMAGE_OK: return "ok";
    case LH_IMAGE_ERR_ARGUMENT: return "invalid argument";
    case LH_IMAGE_ERR_MAGIC: return "not a DDS or KTX2 file";
    case LH_IMAGE_ERR_TRUNCATED: return "file is truncated";
    case LH_IMAGE_ERR_FORMAT: return "unsupported pixel format, dimension or supercompression scheme";
    case LH_IMAGE_ERR_SIZE: return "payload size does not match the mip chain";
    case LH_IMAGE_ERR_ZSTD: return "Zstandard decompression failed (or libzstd.so.1 is missing)";
    default: return "unknown error";
    }
}

}  // extern "C"
