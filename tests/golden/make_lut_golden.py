"""Golden vector for the RGB9E5 path from the reference's own asset: the header facts of
external/tony-mc-mapface/shader/tony_mc_mapface.dds (the display-transform LUT the reference loads at main.cpp:534-541)
and its first 256 texels, decoded here by an independent numpy implementation of the shared-exponent format.
Run in the build container (where /root/reference exists): python tests/golden/make_lut_golden.py"""
import struct
import sys

import numpy as np

path = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/external/tony-mc-mapface/shader/tony_mc_mapface.dds"
raw = open(path, "rb").read()
assert raw[:4] == b"DDS "
height, width = struct.unpack_from("<II", raw, 4 + 8)
depth, mips = struct.unpack_from("<II", raw, 4 + 20)
dxgi, dimension = struct.unpack_from("<II", raw, 4 + 124)
texels = np.frombuffer(raw, np.uint32, count=256, offset=4 + 124 + 20).copy()
# independent decode: mantissa * 2^(e - 15 - 9)
e = (texels >> 27).astype(np.int32)
scale = np.exp2((e - 24).astype(np.float64))
rgb = np.stack([(texels & 0x1FF) * scale, ((texels >> 9) & 0x1FF) * scale, ((texels >> 18) & 0x1FF) * scale], 1).astype(np.float32)
whole = np.frombuffer(raw, np.uint32, offset=4 + 124 + 20)
np.savez_compressed("tests/golden/tony_mc_mapface_head.npz", file_bytes=len(raw), width=width, height=height, depth=depth, mips=mips,
                    dxgi=dxgi, dimension=dimension, texels=texels, rgb=rgb, texel_count=whole.size,
                    xor_all=np.bitwise_xor.reduce(whole), sum_all=np.uint64(whole.astype(np.uint64).sum()))
print(len(raw), width, height, depth, mips, dxgi, dimension, whole.size)
