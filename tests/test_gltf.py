"""glTF ingestion: a KHR_mesh_quantization asset written by tests/gltf_fixture.py must come out of
lighthugger_b200.gltf.load as the scene the host library builds from the same arrays -- streams, meshlets, bounds,
MeshInfo and Instance records byte for byte -- with the reference loader's conventions (src/resources/mesh_loading.cpp)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from lighthugger_b200 import _structs as S
from lighthugger_b200 import gltf as G
from lighthugger_b200.scene import HostBinding
import gltf_fixture


def _instances(scene):
    hb = HostBinding(scene)
    ptr = scene.lib.lh_binding_instances(hb.handle)
    raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), (scene.num_instances * 112,)).reshape(-1, 112).copy()
    addresses = [scene.lib.lh_binding_mesh_info_address(hb.handle, m) for m in range(scene.num_meshes)]
    which = [addresses.index(int(np.frombuffer(r[64:72].tobytes(), np.uint64)[0])) for r in raw]
    hb.close()
    return raw, which


def test_gltf_scene_equals_the_directly_built_scene(tmp_path):
    path, expected = gltf_fixture.build(str(tmp_path))
    scene, images = G.load(path)
    assert images == [] and scene.num_meshes == expected.num_meshes == 2 and scene.num_instances == expected.num_instances == 3
    for m in range(2):
        a, b = scene.mesh_arrays(m), expected.mesh_arrays(m)
        for key in ("positions", "normals", "uvs", "indices", "micro_indices", "meshlets"):
            assert np.array_equal(a[key], b[key]), (m, key)
        assert a["num_triangles"] == b["num_triangles"] and bytes(a["info"]) == bytes(b["info"])
    assert scene.mesh_arrays(0)["indices"].dtype == np.uint16 and scene.mesh_arrays(1)["indices"].dtype == np.uint32   # accessor width
    assert scene.mesh_arrays(1)["info"].flags & S.LH_MESH_INFO_FLAGS_32_BIT_INDICES
    ra, wa = _instances(scene)
    rb, wb = _instances(expected)
    assert wa == wb == [0, 1, 0]
    assert np.array_equal(ra[:, :64], rb[:, :64]) and np.array_equal(ra[:, 72:108], rb[:, 72:108])   # transform, normal_transform
    # node transform = translate(scale(I, s), t): a point p goes to s * (p + t)  (mesh_loading.cpp:8-49)
    t0 = ra[0, :64].view(np.float32).reshape(4, 4)
    s, t = np.float32(0.0001), np.array([-72767.0, -32767.0, -182767.0], np.float32)
    assert np.array_equal(t0[3, :3], s * t) and t0[0, 0] == s and t0[3, 3] == 1.0


def test_gltf_material_conventions(tmp_path):
    from test_image_files import dds_file, DXGI
    files = [(f"t{i}.dds", dds_file(DXGI["BC7_SRGB"], 4, 4, 1, bytes(16))) for i in range(3)]
    path, expected = gltf_fixture.build(str(tmp_path), files)
    scene, images = G.load(path)
    assert [os.path.basename(p) for p in images] == ["t0.dds", "t1.dds", "t2.dds"]
    i0, i1 = scene.mesh_arrays(0)["info"], scene.mesh_arrays(1)["info"]
    assert (i0.base_color_texture_index, i0.metallic_roughness_texture_index, i0.normal_texture_index) == (0, 1, 2)
    assert (i1.base_color_texture_index, i1.metallic_roughness_texture_index, i1.normal_texture_index) == (0, 0xFFFF, 0xFFFF)
    # KHR_texture_transform scale / 65535 for normalised uvs; MASK -> alpha clip
    assert np.allclose(list(i0.texture_scale), np.array([2.0, 3.0], np.float32) / np.float32(65535), rtol=0, atol=0)
    assert list(i0.texture_offset) == [0.25, 0.5] and i1.flags & S.LH_MESH_INFO_FLAGS_ALPHA_CLIP and not i0.flags & S.LH_MESH_INFO_FLAGS_ALPHA_CLIP
    assert np.allclose(list(i0.base_color_factor), [0.9, 0.7, 0.5])
    for m in range(2):
        assert bytes(scene.mesh_arrays(m)["info"]) == bytes(expected.mesh_arrays(m)["info"])


def test_gltf_rejections(tmp_path):
    path, _ = gltf_fixture.build(str(tmp_path))
    doc = json.load(open(path))

    def variant(change):
        d = json.loads(json.dumps(doc))
        change(d)
        p = os.path.join(str(tmp_path), "variant.gltf")
        json.dump(d, open(p, "w"))
        return p

    with pytest.raises(G.GltfError):      # float positions: the reference asserts UnsignedShort
        G.load(variant(lambda d: d["accessors"][0].update(componentType=5126)))
    with pytest.raises(G.GltfError):      # alphaCutoff != 0.5 aborts in the reference
        G.load(variant(lambda d: d["materials"][0].update(alphaCutoff=0.3)))
    with pytest.raises(G.GltfError):      # embedded buffers are not read
        G.load(variant(lambda d: d["buffers"][0].update(uri="data:application/octet-stream;base64,AAAA")))
    with pytest.raises(G.GltfError):      # only .dds / .ktx2 images
        G.load(variant(lambda d: d.update(images=[dict(uri="a.png")])))
    with pytest.raises(G.GltfError):      # an accessor past the end of its buffer
        G.load(variant(lambda d: d["accessors"][3].update(count=10 ** 7)))
    # a primitive without TEXCOORD_0 is skipped, not an error
    scene, _ = G.load(variant(lambda d: d["meshes"][1]["primitives"][0]["attributes"].pop("TEXCOORD_0")))
    assert scene.num_meshes == 1 and scene.num_instances == 2   # two torus instances; the sphere primitive is gone
