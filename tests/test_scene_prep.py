"""Host scene preparation against the reference's own scene-prep code
(golden fixtures made by tests/golden/make_golden.py from oracle/_ref)."""
import ctypes as C
import os

import numpy as np
import pytest

from lighthugger_b200 import _structs as S
from lighthugger_b200.scene import Scene, host_lib
from helpers import bc7_decode_mode6

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("name,n_meshlets", [("torus_32x16_ref.npz", 11), ("torus_125x80_ref.npz", 214)])
def test_bounds_match_reference_bit_for_bit(name, n_meshlets):
    g = np.load(os.path.join(GOLDEN, name))
    lib = host_lib()
    raw = np.ascontiguousarray(g["positions"].astype(np.float32))
    meshlets = g["meshlets"]
    assert len(meshlets) == n_meshlets
    # mesh bounding sphere (src/resources/bounding_sphere.h:3-74)
    sphere = (C.c_float * 4)()
    lib.lh_compute_bounding_sphere(sphere, raw.ctypes.data, len(raw))
    assert np.array_equal(np.array(sphere[:], np.float32).view(np.uint32), g["bounding_sphere"].view(np.uint32))
    # per-meshlet cone + sphere (meshopt_computeMeshletBounds, called at src/resources/meshlets.cpp:69-86)
    tri_off = meshlets[:, 44:48].copy().view(np.uint32).reshape(-1)
    idx_off = meshlets[:, 48:52].copy().view(np.uint32).reshape(-1)
    for m in range(len(meshlets)):
        tc, ic = int(meshlets[m, 52]), int(meshlets[m, 53])
        mv = np.ascontiguousarray(g["indices"][idx_off[m]:idx_off[m] + ic].astype(np.uint32))
        mt = np.ascontiguousarray(g["micro_indices"][tri_off[m]:tri_off[m] + 3 * tc])
        out = S.LhMeshlet()
        lib.lh_compute_meshlet_bounds(mv.ctypes.data, mt.ctypes.data, tc, raw.ctypes.data, C.byref(out))
        ours = np.frombuffer(bytes(out), np.uint8)[:44]
        assert np.array_equal(ours, meshlets[m, :44]), f"meshlet {m} bounds differ from the reference's"
        assert np.array_equal(ours.view(np.float32), np.concatenate([g["bounds"][m][:7], g["bounds"][m][7:]]))


def test_reference_meshlet_packing_facts():
    g = np.load(os.path.join(GOLDEN, "torus_125x80_ref.npz"))
    m = g["meshlets"]
    tc, ic = m[:, 52].astype(int), m[:, 53].astype(int)
    assert tc.max() <= 124 and ic.max() <= 64 and tc.sum() == 20000
    tri_off = m[:, 44:48].copy().view(np.uint32).reshape(-1)
    assert np.all(tri_off % 4 == 0)  # each run padded to 4 bytes (clusterizer.cpp:253)
    assert len(g["micro_indices"]) == tri_off[-1] + ((tc[-1] * 3 + 3) & ~3)


@pytest.mark.parametrize("kind,su,sv", [(S.LH_MESH_TORUS, 32, 16), (S.LH_MESH_TORUS, 125, 80), (S.LH_MESH_SPHERE, 24, 12),
                                        (S.LH_MESH_PLANE, 9, 5), (S.LH_MESH_BOX, 6, 6)])
def test_own_meshlet_builder_invariants(kind, su, sv):
    sc = Scene()
    mid = sc.add_mesh(kind, su, sv)
    a = sc.mesh_arrays(mid)
    m = a["meshlets"]
    tc, ic = m[:, 52].astype(int), m[:, 53].astype(int)
    tri_off = m[:, 44:48].copy().view(np.uint32).reshape(-1)
    idx_off = m[:, 48:52].copy().view(np.uint32).reshape(-1)
    assert tc.max() <= 124 and ic.max() <= 64 and tc.min() >= 1
    assert tc.sum() == a["num_triangles"]
    assert np.all(tri_off % 4 == 0)
    assert a["info"].num_meshlets == len(m)
    # every meshlet-local index is in range and the reassembled triangle set is the mesh's triangle set
    tris = []
    for k in range(len(m)):
        mt = a["micro_indices"][tri_off[k]:tri_off[k] + 3 * tc[k]].astype(int)
        assert mt.max() < ic[k]
        tris.append(a["indices"][idx_off[k] + mt].reshape(-1, 3))
    tris = np.concatenate(tris)
    assert len(np.unique(np.sort(tris, axis=1), axis=0)) == len(tris) or kind in (S.LH_MESH_SPHERE,)  # sphere poles repeat
    assert tris.max() < len(a["positions"])
    # bounds contain every vertex of the meshlet (in quantised space)
    pos = a["positions"].astype(np.float64)
    spheres = m[:, 28:44].copy().view(np.float32).reshape(-1, 4).astype(np.float64)
    for k in range(len(m)):
        vid = a["indices"][idx_off[k]:idx_off[k] + ic[k]]
        d = np.linalg.norm(pos[vid] - spheres[k, :3], axis=1)
        assert d.max() <= spheres[k, 3] * (1 + 1e-5) + 1e-3


def test_instance_normal_transform_is_untransposed_inverse():
    sc = Scene.preset("tiny")
    from lighthugger_b200.scene import HostBinding
    hb = HostBinding(sc)
    ptr = sc.lib.lh_binding_instances(hb.handle)
    inst = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), (sc.num_instances, 28))
    for i in (0, 5, 63):
        T = inst[i, :16].reshape(4, 4).T.astype(np.float64)  # column-major -> math layout
        N = inst[i, 18:27].reshape(3, 3).T.astype(np.float64)
        assert np.allclose(N, np.linalg.inv(T)[:3, :3], rtol=1e-4, atol=1e-7)  # shared_cpu_gpu.h:92-97, no transpose
        # the scene generator stays where that is exact: symmetric rotation x uniform scale
        L = T[:3, :3]
        assert np.allclose(L, L.T, rtol=1e-5, atol=1e-9)


def test_bc7_mode6_round_trip():
    sc = Scene()
    for kind, cutout, mean_tol in ((0, False, 2.0), (1, False, 2.0), (2, False, 2.0), (0, True, 4.0)):
        rgba, levels = sc.make_texture(kind, 64, 3, alpha_cutout=cutout)
        level0 = np.ascontiguousarray(rgba[:64 * 64 * 4])
        blocks = sc.bc7_encode(level0, 64, 64)
        assert blocks.nbytes == 16 * 16 * 16
        dec = bc7_decode_mode6(blocks, 64, 64).astype(int)
        err = np.abs(dec - level0.reshape(64, 64, 4).astype(int))
        assert err.mean() < mean_tol, (kind, cutout, err.mean(), err.max())


def test_presets_build():
    for name, n, tris in (("tiny", 64, 1024), ("c1", 1000, 20000)):
        sc = Scene.preset(name)
        assert sc.num_instances == n and sc.total_triangles == n * tris


# ---- meshlet building: the reference's meshopt_buildMeshlets path (src/resources/meshlets.cpp:3-140) ----

def _add_indexed(sc, positions, normals, uvs, triangles):
    from lighthugger_b200.gltf import LhIndexedMesh
    lib = host_lib()
    lib.lh_scene_add_mesh_indexed.argtypes = [C.c_void_p, C.POINTER(LhIndexedMesh)]
    d = LhIndexedMesh()
    keep = [np.ascontiguousarray(positions, np.uint16), np.ascontiguousarray(normals, np.int8), np.ascontiguousarray(uvs, np.uint16),
            np.ascontiguousarray(triangles, np.uint32)]
    d.positions, d.position_stride = keep[0].ctypes.data, 6
    d.normals, d.normal_stride = keep[1].ctypes.data, 3
    d.uvs, d.uv_stride = keep[2].ctypes.data, 4
    d.num_vertices = len(keep[0])
    d.indices, d.num_indices, d.index_bytes = keep[3].ctypes.data, len(keep[3]), 4
    d.texture_scale[:] = (1 / 65535.0, 1 / 65535.0)
    d.base_color_texture_index = d.metallic_roughness_texture_index = d.normal_texture_index = 0xFFFF
    d.base_color_factor[:] = (1, 1, 1)
    mid = lib.lh_scene_add_mesh_indexed(sc.handle, C.byref(d))
    assert mid >= 0
    return mid


needs_meshopt = pytest.mark.skipif(not host_lib().lh_host_has_meshoptimizer(),
                                   reason="liblh_host.so was built without the reference's vendored meshoptimizer (greedy fallback)")


@needs_meshopt
@pytest.mark.parametrize("name,n_meshlets", [("torus_32x16_ref.npz", 11), ("torus_125x80_ref.npz", 214)])
def test_meshlets_equal_the_reference_fixture_byte_for_byte(name, n_meshlets):
    """The golden fixtures hold what the REFERENCE's build_meshlets made of these very streams."""
    g = np.load(os.path.join(GOLDEN, name))
    sc = Scene()
    a = sc.mesh_arrays(_add_indexed(sc, g["positions"], g["normals"], g["uvs"], g["triangles"]))
    assert len(a["meshlets"]) == n_meshlets
    assert np.array_equal(a["meshlets"].reshape(-1, 56)[:, :54], g["meshlets"][:, :54])   # 2 trailing pad bytes are unspecified
    assert np.array_equal(a["micro_indices"], g["micro_indices"])
    assert np.array_equal(a["indices"], g["indices"])


@needs_meshopt
@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_sceneprep.so")), reason="oracle/_ref not built")
@pytest.mark.parametrize("kind,su,sv", [(S.LH_MESH_TORUS, 125, 80), (S.LH_MESH_TORUS, 250, 200), (S.LH_MESH_SPHERE, 128, 64), (S.LH_MESH_BOX, 40, 40)])
def test_procedural_meshes_are_meshletised_like_the_reference_would(kind, su, sv):
    """The bench's own meshes through the reference's build_meshlets (compiled in place) and through the host library."""
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_sceneprep.so"))
    vp = C.c_void_p
    ref.ref_build_meshlets.argtypes = [vp, C.c_uint64, vp, C.c_uint64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), vp, vp, vp]
    sc = Scene()
    a = sc.mesh_arrays(sc.add_mesh(kind, su, sv))
    tris = np.ascontiguousarray(sc.mesh_triangles(0), np.uint32)
    raw = np.ascontiguousarray(a["positions"].reshape(-1, 3).astype(np.float32))
    n_m, n_micro, n_idx = C.c_uint64(), C.c_uint64(), C.c_uint64()
    assert ref.ref_build_meshlets(tris.ctypes.data, len(tris), raw.ctypes.data, len(raw), C.byref(n_m), C.byref(n_micro), C.byref(n_idx), None, None, None) == 0
    meshlets = np.zeros((n_m.value, 56), np.uint8)
    micro = np.zeros(n_micro.value, np.uint8)
    idx = np.zeros(n_idx.value, np.uint16)
    ref.ref_build_meshlets(tris.ctypes.data, len(tris), raw.ctypes.data, len(raw), C.byref(n_m), C.byref(n_micro), C.byref(n_idx),
                           meshlets.ctypes.data, micro.ctypes.data, idx.ctypes.data)
    assert len(a["meshlets"]) == n_m.value
    assert np.array_equal(a["meshlets"].reshape(-1, 56)[:, :54], meshlets[:, :54])
    assert np.array_equal(a["micro_indices"], micro) and np.array_equal(a["indices"], idx)


def test_bench_scene_meshlet_counts():
    """Candidate meshlets of the BASELINE configurations (meshopt_buildMeshlets, 64 / 124 / 0.25)."""
    if not host_lib().lh_host_has_meshoptimizer():
        pytest.skip("greedy fallback in use")
    sc = Scene()
    assert len(sc.mesh_arrays(sc.add_mesh(S.LH_MESH_TORUS, 125, 80))["meshlets"]) == 214     # C1 mesh
    n = len(sc.mesh_arrays(sc.add_mesh(S.LH_MESH_TORUS, 250, 200))["meshlets"])
    assert 1040 <= n <= 1090, n                                                              # SURVEY 8d measured 1 063 on its own tessellation
