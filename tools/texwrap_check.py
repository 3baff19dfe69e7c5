"""Compares tools/texwrap_probe.cu's dumps (48 x 80 texture, u, v in [-4, 4), repeat addressing) with the oracle's filter
model, fed (a) the raw coordinate and (b) the coordinate wrapped into [0, 1) first, as the oracle's textureLod does.
usage: python tools/texwrap_check.py [dir]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
W, H = 48, 80
lib = C.CDLL(os.path.join(ROOT, "oracle", "liblh_oracle.so"))
lib.lho_test_sample_bilinear.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]
lib.lho_test_sample_bilinear.restype = None
for name in ("unorm", "srgb"):
    raw = open(f"{d}/texwrap_{name}.bin", "rb").read()
    tex = np.frombuffer(raw[:W * H * 4], np.uint8).copy()
    rec = np.frombuffer(raw[W * H * 4:], np.float32).reshape(-1, 6)
    uv = np.ascontiguousarray(rec[:, :2])
    want = np.ascontiguousarray(rec[:, 2:])
    for label, coords in (("raw coordinate", uv), ("wrapped first", (uv - np.floor(uv)).astype(np.float32))):
        coords = np.ascontiguousarray(coords, np.float32)
        out = np.zeros((len(uv), 4), np.float32)
        lib.lho_test_sample_bilinear(tex.ctypes.data, W, H, int(name == "srgb"), len(uv), coords.ctypes.data, out.ctypes.data)
        same = (out.view(np.uint32) == want.view(np.uint32)).all(axis=1)
        print(f"{name:6s} {label:16s}: {same.sum()} of {len(same)} fetches bit-identical, max |diff| {np.abs(out - want).max():.3e}")
