"""On-disk texture formats: DDS (DX10 header) and KTX2 parsing / decoding of liblh_host.so against
files built here field by field from the container specifications, with the quirks of the reference's loaders
(src/resources/image_loading.cpp:14-39,196-217,339-365), and against the reference's own LUT asset."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from lighthugger_b200 import image_files as IF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DXGI = dict(RGB9E5=67, BC1_UNORM=71, BC1_SRGB=72, BC7_UNORM=98, BC7_SRGB=99, RGBA8=28, RGBA8_SRGB=29, BC3_UNORM=77)
VK = dict(RGBA8_UNORM=37, RGBA8_SRGB=43, E5B9G9R9=123, BC1_RGB_SRGB=132, BC7_SRGB=146, BC7_UNORM=145)
KTX2_ID = bytes([0xAB, 0x4B, 0x54, 0x58, 0x20, 0x32, 0x30, 0xBB, 0x0D, 0x0A, 0x1A, 0x0A])


def dds_file(dxgi, width, height, mips, payload, depth=0, cubemap=False, dimension=3):
    """"DDS " + DDS_HEADER (124 B) + DDS_HEADER_DXT10 (20 B) + payload, per the DDS programming guide."""
    pixel_format = struct.pack("<II4sIIIII", 32, 0x4, b"DX10", 0, 0, 0, 0, 0)
    header = struct.pack("<IIIIIII", 124, 0x1007 | 0x20000, height, width, 0, depth, mips) + bytes(44) + pixel_format
    header += struct.pack("<IIIII", 0x1000, 0x200 | 0xFC00 if cubemap else 0, 0, 0, 0)
    assert len(header) == 124
    header10 = struct.pack("<IIIII", dxgi, dimension, 0x4 if cubemap else 0, 1, 0)
    return b"DDS " + header + header10 + bytes(payload)


def level_bytes(width, height, mips, bits, block, faces=1, depth=1):
    out = []
    for l in range(mips):
        w, h = max(width >> l, 1), max(height >> l, 1)
        if block:
            w, h = (w + 3) & ~3, (h + 3) & ~3
        out.append(w * h * depth * faces * bits // 8)
    return out


def ktx2_file(vk_format, width, height, levels, level_data, faces=1, depth=0, scheme=0, stored=None):
    stored = stored or level_data
    head = KTX2_ID + struct.pack("<9I", vk_format, 1, width, height, depth, 0, faces, levels, scheme) + struct.pack("<4I2Q", 0, 0, 0, 0, 0, 0)
    offset = len(head) + 24 * len(level_data)
    index, blob = b"", b""
    for raw, kept in zip(level_data, stored):
        index += struct.pack("<3Q", offset + len(blob), len(kept), len(raw))
        blob += bytes(kept)
    return head + index + blob


def test_dds_bc7_mip_chain_layout_and_srgb_quirk():
    rng = np.random.default_rng(1)
    sizes = level_bytes(64, 32, 7, 8, True)
    assert sizes[-3:] == [16, 16, 16]           # 4x2, 2x1, 1x1 still occupy one 4x4 block each (image_loading.cpp:196-211)
    payload = rng.integers(0, 256, sum(sizes), dtype=np.uint8)
    for dxgi in (DXGI["BC7_UNORM"], DXGI["BC7_SRGB"]):     # both become BC7 sRGB (image_loading.cpp:26-34)
        info, data = IF.decode(dds_file(dxgi, 64, 32, 7, payload))
        assert (info.container, info.format, info.source_format) == (IF.LH_IMAGE_CONTAINER_DDS, IF.LH_PIXEL_BC7_SRGB, dxgi)
        assert (info.width, info.height, info.depth, info.levels, info.faces, info.dimension) == (64, 32, 1, 7, 1, 2)
        assert [info.levels_info[l].bytes for l in range(7)] == sizes
        assert [(info.levels_info[l].width, info.levels_info[l].height) for l in range(7)] == [(max(64 >> l, 1), max(32 >> l, 1)) for l in range(7)]
        assert info.levels_info[0].offset == 148 and info.decoded_bytes == sum(sizes)
        assert np.array_equal(data, payload)


def test_dds_bc1_is_srgb_and_half_byte_per_pixel():
    sizes = level_bytes(16, 16, 5, 4, True)
    assert sizes == [128, 32, 8, 8, 8]
    info = IF.parse(dds_file(DXGI["BC1_UNORM"], 16, 16, 5, bytes(sum(sizes))))
    assert info.format == IF.LH_PIXEL_BC1_SRGB and info.bits_per_pixel == 4 and info.block_compressed == 1


def test_dds_rgb9e5_cubemap_and_volume():
    n = 8
    cube = ((np.arange(6 * n * n, dtype=np.uint64) * 2654435761) & 0xFFFFFFFF).astype(np.uint32)
    info, data = IF.decode(dds_file(DXGI["RGB9E5"], n, n, 1, cube.tobytes(), cubemap=True))
    assert (info.format, info.faces, info.dimension, info.levels) == (IF.LH_PIXEL_RGB9E5, 6, 2, 1)
    assert info.levels_info[0].bytes == 6 * n * n * 4 and np.array_equal(data.view(np.uint32), cube)
    vol = np.arange(n ** 3, dtype=np.uint32)
    info = IF.parse(dds_file(DXGI["RGB9E5"], n, n, 1, vol.tobytes(), depth=n, dimension=4))
    assert (info.dimension, info.depth, info.faces, info.decoded_bytes) == (3, n, 1, n ** 3 * 4)


def test_dds_malformed_inputs_are_reported_not_aborted():
    good = dds_file(DXGI["BC7_SRGB"], 8, 8, 1, bytes(64))
    cases = [(b"DDX " + good[4:], -2), (good[:100], -3), (good[:-1], -3), (good + b"\0", -5),
             (dds_file(DXGI["BC3_UNORM"], 8, 8, 1, bytes(64)), -4), (dds_file(DXGI["BC7_SRGB"], 8, 8, 1, bytes(64), dimension=2), -4),
             (dds_file(DXGI["BC7_SRGB"], 0, 8, 1, b""), -4), (b"", -2)]
    for data, code in cases:
        with pytest.raises(IF.ImageError) as e:
            IF.parse(data)
        assert e.value.code == code, (len(data), code, e.value.code)


def _zstd_compress(raw):
    lib = C.CDLL("libzstd.so.1")
    lib.ZSTD_compressBound.restype = C.c_size_t
    lib.ZSTD_compressBound.argtypes = [C.c_size_t]
    lib.ZSTD_compress.restype = C.c_size_t
    lib.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    dst = C.create_string_buffer(lib.ZSTD_compressBound(len(raw)))
    n = lib.ZSTD_compress(dst, len(dst), bytes(raw), len(raw), 3)
    return dst.raw[:n]


def test_ktx2_plain_and_zstandard_levels():
    rng = np.random.default_rng(2)
    sizes = level_bytes(32, 32, 6, 8, True)
    levels = [np.repeat(rng.integers(0, 256, max(1, s // 16), dtype=np.uint8), 16)[:s].tobytes() for s in sizes]   # compressible
    plain = ktx2_file(VK["BC7_SRGB"], 32, 32, 6, levels)
    info, data = IF.decode(plain)
    assert (info.container, info.format, info.levels, info.zstd, info.faces) == (IF.LH_IMAGE_CONTAINER_KTX2, IF.LH_PIXEL_BC7_SRGB, 6, 0, 1)
    assert bytes(data) == b"".join(levels)
    packed = ktx2_file(VK["BC7_SRGB"], 32, 32, 6, levels, scheme=2, stored=[_zstd_compress(l) for l in levels])
    assert len(packed) < len(plain)
    info, data = IF.decode(packed)
    assert info.zstd == 1 and [info.levels_info[l].bytes for l in range(6)] == sizes
    assert info.levels_info[0].stored_bytes < info.levels_info[0].bytes
    assert bytes(data) == b"".join(levels)


def test_ktx2_malformed_inputs():
    levels = [bytes(64)]
    good = ktx2_file(VK["BC7_SRGB"], 8, 8, 1, levels)
    corrupt_zstd = ktx2_file(VK["BC7_SRGB"], 8, 8, 1, levels, scheme=2, stored=[b"not a zstd frame"])
    cases = [(good[:60], -3), (good[:-8], -3), (ktx2_file(VK["BC7_UNORM"], 8, 8, 1, levels), -4),
             (ktx2_file(VK["BC7_SRGB"], 8, 8, 1, levels, scheme=1), -4), (ktx2_file(VK["BC7_SRGB"], 8, 8, 1, [bytes(32)]), -5)]
    for data, code in cases:
        with pytest.raises(IF.ImageError) as e:
            IF.parse(data)
        assert e.value.code == code
    with pytest.raises(IF.ImageError) as e:
        IF.decode(corrupt_zstd)
    assert e.value.code == -6


def test_rgb9e5_decode_known_answers_and_reference_lut_head():
    # known answers: 1.0 = mantissa 256 at exponent 16; the largest value; zero
    one = 256 | (16 << 27)
    texels = np.array([one, (256 << 9) | (16 << 27), 0x1FF | (31 << 27), 0], np.uint32)
    rgb = IF.rgb9e5_to_float(texels)
    assert rgb[0].tolist() == [1.0, 0.0, 0.0] and rgb[1].tolist() == [0.0, 1.0, 0.0]
    assert rgb[2, 0] == 511 * 2.0 ** (31 - 24) and rgb[3].tolist() == [0.0, 0.0, 0.0]
    g = np.load(os.path.join(ROOT, "tests", "golden", "tony_mc_mapface_head.npz"))
    assert (int(g["file_bytes"]), int(g["width"]), int(g["depth"]), int(g["dxgi"]), int(g["dimension"])) == (442516, 48, 48, 67, 4)
    assert np.array_equal(IF.rgb9e5_to_float(g["texels"]), g["rgb"])


def test_reference_lut_file_when_present():
    path = "/root/reference/external/tony-mc-mapface/shader/tony_mc_mapface.dds"
    if not os.path.exists(path):
        pytest.skip("the reference tree is not mounted")
    g = np.load(os.path.join(ROOT, "tests", "golden", "tony_mc_mapface_head.npz"))
    info, data = IF.decode(open(path, "rb").read())
    assert (info.format, info.dimension, info.width, info.height, info.depth, info.levels) == (IF.LH_PIXEL_RGB9E5, 3, 48, 48, 48, 1)
    texels = data.view(np.uint32)
    assert texels.size == int(g["texel_count"]) and np.bitwise_xor.reduce(texels) == g["xor_all"]
    assert int(texels.astype(np.uint64).sum()) == int(g["sum_all"])


def test_parser_survives_mutated_files_under_address_sanitizer(tmp_path):
    # the host library's parser + decoder compiled with -fsanitize=address,undefined and driven by a mutation fuzzer
    # (tests/native/fuzz_image_files.cpp): malformed input must come back as an error code, never as an out-of-bounds access
    import shutil
    import subprocess
    cxx = shutil.which("g++")
    if cxx is None:
        pytest.skip("no C++ compiler")
    exe = str(tmp_path / "fuzz_image_files")
    build = subprocess.run([cxx, "-std=c++20", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-I", os.path.join(ROOT, "include"),
                            os.path.join(ROOT, "tests", "native", "fuzz_image_files.cpp"), os.path.join(ROOT, "lighthugger_b200", "host", "image_files.cpp"),
                            "-o", exe, "-ldl"], capture_output=True, text=True)
    if build.returncode != 0 and "sanitize" in build.stderr:
        pytest.skip("the sanitizer runtimes are not installed")
    assert build.returncode == 0, build.stderr[-2000:]
    rng = np.random.default_rng(3)
    sizes = level_bytes(32, 16, 6, 8, True)
    seeds = {
        "bc7.dds": dds_file(DXGI["BC7_SRGB"], 32, 16, 6, rng.integers(0, 256, sum(sizes), dtype=np.uint8)),
        "cube.dds": dds_file(DXGI["RGB9E5"], 4, 4, 1, bytes(6 * 16 * 4), cubemap=True),
        "lut.dds": dds_file(DXGI["RGB9E5"], 4, 4, 1, bytes(64 * 4), depth=4, dimension=4),
        "plain.ktx2": ktx2_file(VK["BC7_SRGB"], 32, 16, 6, [bytes(s) for s in sizes]),
        "zstd.ktx2": ktx2_file(VK["BC7_SRGB"], 32, 16, 6, [bytes(s) for s in sizes], scheme=2, stored=[_zstd_compress(bytes(s)) for s in sizes]),
    }
    paths = []
    for name, data in seeds.items():
        p = str(tmp_path / name)
        open(p, "wb").write(data)
        paths.append(p)
    run = subprocess.run([exe, "3000"] + paths, capture_output=True, text=True, timeout=300,
                         env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    assert run.returncode == 0 and "fuzz ok" in run.stdout, (run.stdout + run.stderr)[-3000:]
