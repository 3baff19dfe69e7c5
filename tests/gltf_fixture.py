"""Writes a small glTF asset the way the reference's pipeline expects them (KHR_mesh_quantization, padded vertex
streams, external .bin, .dds / .ktx2 images, KHR_texture_transform) together with the SAME scene built through the host
library directly, so that tests can compare the two (tests/test_gltf.py, tests/test_resolve_gpu.py)."""
import ctypes as C
import json
import os

import numpy as np

from lighthugger_b200 import _structs as S
from lighthugger_b200 import gltf as G
from lighthugger_b200.scene import Scene


def triangle_list(arrays):
    """The mesh's triangles in meshlet order (global vertex indices): feeding them back to the greedy meshlet builder
    reproduces the same meshlets."""
    m = arrays["meshlets"]
    tri_off = m[:, 44:48].copy().view(np.uint32).reshape(-1)
    idx_off = m[:, 48:52].copy().view(np.uint32).reshape(-1)
    out = []
    for k in range(len(m)):
        tc = int(m[k, 52])
        micro = arrays["micro_indices"][tri_off[k]:tri_off[k] + 3 * tc].astype(np.int64)
        out.append(arrays["indices"][idx_off[k] + micro])
    return np.concatenate(out)


# node transform = translate(scale(I, s), t): a quantised point q (0..65535) lands at s * (q + t); the three instances
# sit in front of a camera at the origin looking down -z
NODES = [dict(mesh=0, scale=[0.0001, 0.0001, 0.0001], translation=[-72767.0, -32767.0, -182767.0]),
         dict(mesh=1, scale=[0.0001, 0.00015, 0.0001], translation=[7233.0, -26100.0, -212767.0]),
         dict(mesh=0, scale=[0.00005, 0.00005, 0.00005], translation=[-32767.0, -92767.0, -272767.0]),
         dict(name="an empty node", translation=[1.0, 2.0, 3.0])]


def build(directory, image_files=()):
    """Returns (path of the .gltf, the equivalent Scene built directly).  image_files: (name, bytes) per bindless slot."""
    src = Scene()
    torus = src.add_mesh(S.LH_MESH_TORUS, 48, 24)
    sphere = src.add_mesh(S.LH_MESH_SPHERE, 32, 16)
    specs = [dict(mesh=torus, alpha=False, slots=(0, 1, 2) if image_files else (G.UNUSED,) * 3, color=(0.9, 0.7, 0.5), scale=(2.0, 3.0), offset=(0.25, 0.5)),
             dict(mesh=sphere, alpha=bool(image_files), slots=(0, G.UNUSED, G.UNUSED) if image_files else (G.UNUSED,) * 3, color=(0.4, 0.8, 1.0), scale=(1.0, 1.0), offset=(0.0, 0.0))]
    blob, views, accessors, meshes, materials = bytearray(), [], [], [], []

    def add_view(data, stride=None):
        while len(blob) % 4:
            blob.append(0)
        views.append(dict(buffer=0, byteOffset=len(blob), byteLength=len(data), **({"byteStride": stride} if stride else {})))
        blob.extend(data)
        return len(views) - 1

    expected = Scene()
    expected.lib.lh_scene_add_mesh_indexed.argtypes = [C.c_void_p, C.POINTER(G.LhIndexedMesh)]
    expected.lib.lh_scene_add_instance_matrix.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
    keep = []
    for k, spec in enumerate(specs):
        a = src.mesh_arrays(spec["mesh"])
        nv = len(a["positions"])
        pos4 = np.zeros((nv, 4), np.uint16); pos4[:, :3] = a["positions"]
        nrm4 = np.zeros((nv, 4), np.int8); nrm4[:, :3] = a["normals"]
        tris = triangle_list(a).astype(np.uint32 if k == 1 else np.uint16)      # the second primitive uses 32-bit indices
        accessors += [dict(bufferView=add_view(pos4.tobytes(), 8), componentType=5123, count=nv, type="VEC3"),
                      dict(bufferView=add_view(nrm4.tobytes(), 4), componentType=5120, count=nv, type="VEC3", normalized=True),
                      dict(bufferView=add_view(a["uvs"].tobytes(), 4), componentType=5123, count=nv, type="VEC2", normalized=True),
                      dict(bufferView=add_view(tris.tobytes()), componentType=5125 if k == 1 else 5123, count=tris.size, type="SCALAR")]
        base = 4 * k
        meshes.append(dict(primitives=[dict(attributes=dict(POSITION=base, NORMAL=base + 1, TEXCOORD_0=base + 2), indices=base + 3, material=k)]))
        tt = dict(KHR_texture_transform=dict(scale=list(spec["scale"]), offset=list(spec["offset"])))
        mat = dict(pbrMetallicRoughness=dict(baseColorFactor=list(spec["color"]) + [1.0]), alphaMode="MASK" if spec["alpha"] else "OPAQUE",
                   doubleSided=spec["alpha"])
        if spec["slots"][0] != G.UNUSED:
            mat["pbrMetallicRoughness"]["baseColorTexture"] = dict(index=spec["slots"][0], extensions=tt)
        if spec["slots"][1] != G.UNUSED:
            mat["pbrMetallicRoughness"]["metallicRoughnessTexture"] = dict(index=spec["slots"][1], extensions=tt)
        if spec["slots"][2] != G.UNUSED:
            mat["normalTexture"] = dict(index=spec["slots"][2], extensions=tt)
        materials.append(mat)
        # the same primitive through the host library directly
        d = G.LhIndexedMesh()
        p3, n3, uv = np.ascontiguousarray(a["positions"]), np.ascontiguousarray(a["normals"]), np.ascontiguousarray(a["uvs"])
        keep += [p3, n3, uv, tris]
        d.positions, d.position_stride, d.normals, d.normal_stride, d.uvs, d.uv_stride = p3.ctypes.data, 6, n3.ctypes.data, 3, uv.ctypes.data, 4
        d.num_vertices, d.indices, d.num_indices, d.index_bytes = nv, tris.ctypes.data, tris.size, tris.dtype.itemsize
        d.flags = S.LH_MESH_INFO_FLAGS_ALPHA_CLIP if spec["alpha"] else 0
        has_tex = spec["slots"][0] != G.UNUSED
        sc = np.array(spec["scale"] if has_tex else (1.0, 1.0), np.float32) / np.float32(65535)
        d.texture_scale[:] = sc.tolist()
        d.texture_offset[:] = list(spec["offset"] if has_tex else (0.0, 0.0))
        d.base_color_texture_index, d.metallic_roughness_texture_index, d.normal_texture_index = spec["slots"]
        d.base_color_factor[:] = list(spec["color"])
        assert expected.lib.lh_scene_add_mesh_indexed(expected.handle, C.byref(d)) == k
    for node in NODES:
        if "mesh" in node:
            m = np.ascontiguousarray(G.node_matrix(node).reshape(-1))
            assert expected.lib.lh_scene_add_instance_matrix(expected.handle, node["mesh"], m.ctypes.data_as(C.POINTER(C.c_float))) == 0
    doc = dict(asset=dict(version="2.0"), extensionsUsed=["KHR_mesh_quantization", "KHR_texture_transform"],
               extensionsRequired=["KHR_mesh_quantization"], buffers=[dict(uri="scene.bin", byteLength=len(blob))], bufferViews=views,
               accessors=accessors, meshes=meshes, materials=materials, nodes=NODES, scenes=[dict(nodes=list(range(len(NODES))))], scene=0)
    if image_files:
        doc["images"] = [dict(uri=name) for name, _ in image_files]
        doc["textures"] = [dict(source=i) for i in range(len(image_files))]
        for name, data in image_files:
            with open(os.path.join(directory, name), "wb") as f:
                f.write(data)
    with open(os.path.join(directory, "scene.bin"), "wb") as f:
        f.write(bytes(blob))
    path = os.path.join(directory, "scene.gltf")
    with open(path, "w") as f:
        json.dump(doc, f)
    look_at_instances(expected)
    return path, expected


def look_at_instances(scene):
    scene.camera.position[:] = [0.0, 0.0, 0.0]
    scene.camera.yaw, scene.camera.pitch = -1.5707964, 0.0     # facing (0, 0, -1), src/input.h:31-37
    scene.width, scene.height = 320, 180
