/* On-disk texture formats feeding the resolve: the DDS (DX10 header) and KTX2 containers the
 * reference loads with load_dds / load_ktx2_image (src/resources/image_loading.cpp:67-240, 242-411; headers
 * src/resources/dds.h, src/resources/ktx2.h).  Host-side C++ (liblh_host.so), no CUDA: a parsed file is handed to
 * lh_texture_create / lh_skybox_create / lh_set_display_lut of the CUDA layer.
 *
 * Differences from the reference, by design: parsing works on a memory image of the file (the caller reads or maps
 * it), nothing aborts -- malformed input is an error code -- and Zstandard is resolved at run time (libzstd.so.1). */
#ifndef LH_IMAGE_FILES_H
#define LH_IMAGE_FILES_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#ifndef LH_HOST_API
#define LH_HOST_API __attribute__((visibility("default")))
#endif

enum {
    LH_IMAGE_OK = 0,
    LH_IMAGE_ERR_ARGUMENT = -1,
    LH_IMAGE_ERR_MAGIC = -2,        /* neither "DDS " nor the KTX2 identifier */
    LH_IMAGE_ERR_TRUNCATED = -3,    /* headers or level data run past the end of the file */
    LH_IMAGE_ERR_FORMAT = -4,       /* a pixel format / dimension / supercompression this path does not handle */
    LH_IMAGE_ERR_SIZE = -5,         /* payload size does not match the mip chain (image_loading.cpp:214-217) */
    LH_IMAGE_ERR_ZSTD = -6          /* libzstd missing, or a level failed to decompress to its stated size */
};

enum { LH_IMAGE_CONTAINER_DDS = 1, LH_IMAGE_CONTAINER_KTX2 = 2 };

/* Pixel formats the path uses.  The first four are the values lh_texture_create takes (LH_TEXTURE_*). */
enum {
    LH_PIXEL_RGBA8_SRGB = 0,
    LH_PIXEL_RGBA8_UNORM = 1,
    LH_PIXEL_BC7_SRGB = 2,
    LH_PIXEL_BC1_SRGB = 3,
    LH_PIXEL_RGB9E5 = 16            /* shared-exponent HDR (skybox cube, Tony McMapface LUT): decode with lh_rgb9e5_decode */
};

#define LH_IMAGE_MAX_LEVELS 16

typedef struct LhImageLevel {
    uint64_t offset;               /* of the level's data from the start of the file */
    uint64_t stored_bytes;         /* bytes in the file (Zstandard-compressed size for a supercompressed KTX2) */
    uint64_t bytes;                /* bytes of the level once decoded: all faces / depth slices of the level */
    uint32_t width, height;
} LhImageLevel;

typedef struct LhImageInfo {
    int container;                 /* LH_IMAGE_CONTAINER_* */
    int format;                    /* LH_PIXEL_* */
    int source_format;             /* DXGI_FORMAT (DDS) or VkFormat (KTX2) as stored */
    uint32_t width, height, depth; /* depth >= 1 */
    uint32_t levels;               /* >= 1 */
    uint32_t faces;                /* 6 for a cubemap, else 1 */
    uint32_t dimension;            /* 2 or 3 */
    uint32_t bits_per_pixel;
    uint32_t block_compressed;     /* 4x4 blocks: a level occupies at least one block per axis */
    uint32_t zstd;                 /* KTX2 supercompression scheme 2 */
    uint64_t decoded_bytes;        /* sum of levels[].bytes */
    LhImageLevel levels_info[LH_IMAGE_MAX_LEVELS];
} LhImageInfo;

/* Parses the headers and lays out the levels.  DDS: image_loading.cpp:67-217 incl. its quirks -- every block-compressed
 * DXGI format is taken as sRGB (BC1_UNORM -> BC1 sRGB, BC7_UNORM -> BC7 sRGB, :14-39), a level's size is computed from
 * its dimensions rounded up to 4 (:196-211) and multiplied by 6 for a cubemap, and the sizes must add up to the payload.
 * KTX2: image_loading.cpp:242-330 -- header, index, level index; sizes come from the level index. */
LH_HOST_API int lh_image_parse(const void* file, uint64_t file_bytes, LhImageInfo* out);

/* Copies (DDS, plain KTX2) or Zstandard-decompresses (KTX2 scheme 2) the levels, mip 0 first, tightly packed, into
 * dst[0 .. info->decoded_bytes): the layout lh_texture_create expects. */
LH_HOST_API int lh_image_decode(const void* file, uint64_t file_bytes, const LhImageInfo* info, void* dst, uint64_t dst_bytes);

/* R9G9B9E5 shared-exponent texels -> 3 floats each (the reference samples them through the hardware format; CUDA has
 * no such array format, so the host decodes at upload). */
LH_HOST_API void lh_rgb9e5_decode(const uint32_t* texels, uint64_t count, float* rgb);

LH_HOST_API const char* lh_image_error_string(int code);

#ifdef __cplusplus
}
#endif
#endif /* LH_IMAGE_FILES_H */
