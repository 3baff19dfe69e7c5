"""ctypes binding of the on-disk texture formats (include/lh_image_files.h, liblh_host.so): DDS / KTX2 -> level layout
+ payload, ready for Context.create_texture / create_skybox / set_display_lut.  Mirrors what the reference's
load_dds / load_ktx2_image accept (src/resources/image_loading.cpp:67-411)."""
import ctypes as C

import numpy as np

from .scene import host_lib

LH_PIXEL_RGBA8_SRGB, LH_PIXEL_RGBA8_UNORM, LH_PIXEL_BC7_SRGB, LH_PIXEL_BC1_SRGB, LH_PIXEL_RGB9E5 = 0, 1, 2, 3, 16
LH_IMAGE_CONTAINER_DDS, LH_IMAGE_CONTAINER_KTX2 = 1, 2
LH_IMAGE_MAX_LEVELS = 16


class LhImageLevel(C.Structure):
    _fields_ = [("offset", C.c_uint64), ("stored_bytes", C.c_uint64), ("bytes", C.c_uint64), ("width", C.c_uint32),
                ("height", C.c_uint32)]


class LhImageInfo(C.Structure):
    _fields_ = [("container", C.c_int), ("format", C.c_int), ("source_format", C.c_int), ("width", C.c_uint32),
                ("height", C.c_uint32), ("depth", C.c_uint32), ("levels", C.c_uint32), ("faces", C.c_uint32),
                ("dimension", C.c_uint32), ("bits_per_pixel", C.c_uint32), ("block_compressed", C.c_uint32),
                ("zstd", C.c_uint32), ("decoded_bytes", C.c_uint64), ("levels_info", LhImageLevel * LH_IMAGE_MAX_LEVELS)]


class ImageError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"image file error {code}: {message}")
        self.code = code


_bound = False


def _lib():
    global _bound
    lib = host_lib()
    if not _bound:
        lib.lh_image_parse.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(LhImageInfo)]
        lib.lh_image_decode.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(LhImageInfo), C.c_void_p, C.c_uint64]
        lib.lh_rgb9e5_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        lib.lh_rgb9e5_decode.restype = None
        lib.lh_image_error_string.argtypes = [C.c_int]
        lib.lh_image_error_string.restype = C.c_char_p
        _bound = True
    return lib


def _check(lib, rc):
    if rc != 0:
        raise ImageError(rc, lib.lh_image_error_string(rc).decode())


def parse(data):
    """Header + level layout of a DDS / KTX2 file given as bytes."""
    lib = _lib()
    buf = np.frombuffer(bytes(data), np.uint8)
    info = LhImageInfo()
    _check(lib, lib.lh_image_parse(buf.ctypes.data, buf.size, C.byref(info)))
    return info


def decode(data, info=None):
    """(info, payload): the levels, mip 0 first, tightly packed (Zstandard levels decompressed) as a uint8 array."""
    lib = _lib()
    buf = np.frombuffer(bytes(data), np.uint8)
    if info is None:
        info = parse(data)
    out = np.empty(info.decoded_bytes, np.uint8)
    _check(lib, lib.lh_image_decode(buf.ctypes.data, buf.size, C.byref(info), out.ctypes.data, out.size))
    return info, out


def rgb9e5_to_float(texels):
    lib = _lib()
    texels = np.ascontiguousarray(texels, np.uint32).reshape(-1)
    out = np.empty((texels.size, 3), np.float32)
    lib.lh_rgb9e5_decode(texels.ctypes.data, texels.size, out.ctypes.data)
    return out


def upload(ctx, data):
    """Creates what the file holds on a renderer.Context and says what it was: ("texture", slot) for a 2-D texture with
    its mip chain (DescriptorSet::write_image in the reference), ("skybox", size) for an RGB9E5 cubemap, ("lut", size)
    for an RGB9E5 3-D table (main.cpp:534-549)."""
    info, payload = decode(data)
    if info.format == LH_PIXEL_RGB9E5:
        texels = payload[:info.levels_info[0].bytes].view(np.uint32)
        rgb = rgb9e5_to_float(texels)
        if info.faces == 6 and info.width == info.height:
            rgba = np.ones((texels.size, 4), np.float32)
            rgba[:, :3] = rgb
            ctx.create_skybox(info.width, rgba.reshape(-1))
            return "skybox", info.width
        if info.dimension == 3 and info.width == info.height == info.depth:
            ctx.set_display_lut(info.width, rgb.reshape(-1))
            return "lut", info.width
        raise ImageError(-4, "an RGB9E5 image must be a cubemap or a cubic 3-D table")
    if info.faces != 1 or info.depth != 1:
        raise ImageError(-4, "material textures are plain 2-D images")
    return "texture", ctx.create_texture(info.format, info.width, info.height, info.levels, payload)
