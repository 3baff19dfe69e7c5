"""Layout and constant pins against the reference header."""
import ctypes as C
import os

import numpy as np
import pytest

from lighthugger_b200 import _structs as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libref_sceneprep.so")


def test_struct_sizes_and_offsets():
    assert C.sizeof(S.LhInstance) == 112 and S.LhInstance.mesh_info_address.offset == 64 and S.LhInstance.normal_transform.offset == 72
    assert C.sizeof(S.LhMeshInfo) == 104
    for name, off in (("meshlets", 40), ("num_meshlets", 48), ("flags", 50), ("bounding_sphere", 52), ("texture_scale", 68),
                      ("texture_offset", 76), ("base_color_texture_index", 84), ("metallic_roughness_texture_index", 86),
                      ("normal_texture_index", 88), ("base_color_factor", 92)):
        assert getattr(S.LhMeshInfo, name).offset == off, name
    assert C.sizeof(S.LhMeshlet) == 56
    for name, off in (("cone_axis", 12), ("cone_cutoff", 24), ("bounding_sphere", 28), ("triangle_offset", 44),
                      ("index_offset", 48), ("triangle_count", 52), ("index_count", 53)):
        assert getattr(S.LhMeshlet, name).offset == off, name
    assert C.sizeof(S.LhMiscStorage) == 792
    for name, off in (("uv_space_shadow_matrices", 256), ("shadow_view_matrices", 512), ("shadow_sphere_radii", 768),
                      ("min_depth", 784), ("max_depth", 788)):
        assert getattr(S.LhMiscStorage, name).offset == off, name
    assert C.sizeof(S.LhUniforms) == 560
    for name, off in (("meshlet_references", 448), ("instances", 456), ("draw_calls", 464), ("misc_storage", 472),
                      ("num_meshlets_prefix_sum", 480), ("dispatches", 488), ("camera_pos", 496), ("sun_dir", 508),
                      ("sun_intensity", 520), ("num_instances", 532), ("window_size", 536), ("shadow_cam_distance", 544),
                      ("cascade_split_pow", 548), ("debug", 552), ("debug_shadowmaps", 556)):
        assert getattr(S.LhUniforms, name).offset == off, name
    assert C.sizeof(S.LhMeshletReference) == 8 and S.LhMeshletReference.meshlet_index.offset == 4
    assert C.sizeof(S.LhPrefixSumValue) == 8 and C.sizeof(S.LhDrawIndirectCommand) == 16
    assert C.sizeof(S.LhDispatchIndirectCommand) == 12


def test_constants():
    # PREFIX_SUM_BUFFER_SECTOR_SIZE, MESHLET_INDICES_BUFFER_SECTION_OFFSET (shared_cpu_gpu.h:119-125,159-162)
    assert 1024 * 8 + 8 == 8200
    assert 200000 + 200000 == 400000


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref not built (needs /root/reference)")
def test_sizes_match_reference_header_compiled_here():
    lib = C.CDLL(REF)
    out = (C.c_uint32 * 9)()
    lib.ref_struct_sizes(out)
    assert list(out) == [112, 104, 56, 8, 8, 16, 12, 792, 560]


def test_golden_struct_sizes_fixture():
    # committed copy of the numbers above as produced from the reference header (tests/golden/make_golden.py)
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_layout.json")) as f:
        g = json.load(f)
    assert g["sizes"] == {"Instance": 112, "MeshInfo": 104, "Meshlet": 56, "MeshletReference": 8, "PrefixSumValue": 8,
                          "DrawIndirectCommand": 16, "DispatchIndirectCommand": 12, "MiscStorage": 792, "Uniforms": 560}
    for name, n in g["sizes"].items():
        assert C.sizeof(getattr(S, "Lh" + name)) == n
