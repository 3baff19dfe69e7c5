"""The reference's on-disk meshlet cache (src/resources/fs_cache.h:3-25, mesh_loading.cpp:602-649):
three raw blobs per primitive, file name = decimal std::hash of the key.  Exercised with the meshlets the REFERENCE's
own build_meshlets produced (tests/golden/torus_*_ref.npz) and with a mesh prepared by the host library."""
import ctypes as C
import os

import numpy as np

from lighthugger_b200 import _structs as S
from lighthugger_b200.scene import Scene, host_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    lib = host_lib()
    lib.lh_meshlet_cache_path.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_uint64]
    lib.lh_meshlet_cache_store.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(S.LhMeshView)]
    lib.lh_meshlet_cache_load.argtypes = [C.c_char_p, C.c_char_p, C.c_uint32, C.POINTER(C.c_void_p)]
    lib.lh_meshlet_blobs_view.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(S.LhMeshView)]
    lib.lh_meshlet_blobs_free.argtypes = [C.c_void_p]
    return lib


def _paths(lib, d, name):
    out = []
    for which in range(3):
        buf = C.create_string_buffer(1024)
        assert lib.lh_meshlet_cache_path(d, name, which, buf, len(buf)) == 0
        out.append(buf.value.decode())
    return out


def test_reference_built_meshlets_round_trip(tmp_path):
    lib = _lib()
    g = np.load(os.path.join(ROOT, "tests", "golden", "torus_125x80_ref.npz"))
    meshlets = np.ascontiguousarray(g["meshlets"])                 # [214, 56] bytes, as the reference packed them
    indices = np.ascontiguousarray(g["indices"].astype(np.uint16))
    micro = np.ascontiguousarray(g["micro_indices"].astype(np.uint8))
    view = S.LhMeshView()
    view.meshlets, view.num_meshlets = meshlets.ctypes.data, len(meshlets)
    view.indices, view.num_indices, view.index_bytes = indices.ctypes.data, indices.size, 2
    view.micro_indices, view.micro_indices_bytes = micro.ctypes.data, micro.size
    d, name = str(tmp_path / "cache").encode(), b"torus.gltf mesh 0 primitive 0"
    assert lib.lh_meshlet_cache_load(d, name, 2, C.byref(C.c_void_p())) == 1          # miss before the store
    assert lib.lh_meshlet_cache_store(d, name, C.byref(view)) == 0
    paths = _paths(lib, d, name)
    assert len(set(paths)) == 3 and all(os.path.basename(p).isdigit() for p in paths)   # std::to_string(std::hash(key))
    assert [os.path.getsize(p) for p in paths] == [214 * 56, indices.size * 2, micro.size]
    assert open(paths[0], "rb").read() == meshlets.tobytes() and open(paths[2], "rb").read() == micro.tobytes()
    blobs = C.c_void_p()
    assert lib.lh_meshlet_cache_load(d, name, 2, C.byref(blobs)) == 0
    back = S.LhMeshView()
    assert lib.lh_meshlet_blobs_view(blobs, 2, C.byref(back)) == 0
    assert (back.num_meshlets, back.num_indices, back.micro_indices_bytes) == (214, indices.size, micro.size)
    assert back.num_triangles == int(meshlets[:, 52].sum()) == 20000
    assert C.string_at(back.meshlets, 214 * 56) == meshlets.tobytes()
    assert C.string_at(back.indices, indices.size * 2) == indices.tobytes()
    lib.lh_meshlet_blobs_free(blobs)
    # a damaged blob is reported (the reference would read past its buffers)
    with open(paths[0], "r+b") as f:
        f.truncate(214 * 56 - 1)
    assert lib.lh_meshlet_cache_load(d, name, 2, C.byref(C.c_void_p())) == -3
    with open(paths[0], "wb") as f:
        bad = meshlets.copy()
        bad[7, 48:52] = np.frombuffer(np.uint32(10 ** 7).tobytes(), np.uint8)          # index_offset far outside the index blob
        f.write(bad.tobytes())
    assert lib.lh_meshlet_cache_load(d, name, 2, C.byref(C.c_void_p())) == -3


def test_cached_meshlets_rebuild_the_same_scene_mesh(tmp_path):
    lib = _lib()
    sc = Scene()
    mesh = sc.add_mesh(S.LH_MESH_TORUS, 40, 24)
    view = sc.mesh_view(mesh)
    d, name = str(tmp_path).encode(), b"procedural torus 40x24"
    assert lib.lh_meshlet_cache_store(d, name, C.byref(view)) == 0
    blobs = C.c_void_p()
    assert lib.lh_meshlet_cache_load(d, name, view.index_bytes, C.byref(blobs)) == 0
    cached = sc.mesh_view(mesh)            # vertex streams and MeshInfo from the asset, meshlet buffers from the cache
    assert lib.lh_meshlet_blobs_view(blobs, view.index_bytes, C.byref(cached)) == 0
    other = Scene()
    again = other.add_mesh_raw(cached)
    a, b = sc.mesh_arrays(mesh), other.mesh_arrays(again)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    lib.lh_meshlet_blobs_free(blobs)
