"""glTF scene ingestion, following what the reference's load_gltf accepts and how it interprets it
(src/resources/mesh_loading.cpp:8-49, 241-738; src/main.cpp:433-440):

* extensions KHR_mesh_quantization and KHR_texture_transform; buffers and images are external files (URIs);
* per primitive: POSITION u16 vec3 (padded to 8 bytes), NORMAL i8 vec3 (padded to 4), TEXCOORD_0 u16 vec2, u16 / u32
  indices; a primitive without TEXCOORD_0 is skipped (:412-416);
* material: base colour / metallic-roughness / normal textures -> bindless slots in image order, the LAST texture
  transform seen wins for scale and offset (:500-552), normalised uvs divide the scale by 65535 (:555-557), alphaMode MASK
  -> alpha-clip flag, alphaCutoff must be 0.5 (:567-570), baseColorFactor rgb;
* node transform = translate(scale(I, s), t) = S * T -- no rotation, no parent chain (:8-49, "these are probably not
  correct" in the reference; kept, because the instances it produces are the path's input);
* one Instance per primitive with its node's matrix (main.cpp:433-440).

The JSON and accessor plumbing lives here; compaction, meshlets, bounds and MeshInfo are built by the C++ host library
(lh_scene_add_mesh_indexed).  Images are returned as file paths in slot order for image_files.upload."""
import ctypes as C
import json
import os

import numpy as np

from . import _structs as S
from .scene import Scene

_COMPONENT = {5120: np.int8, 5121: np.uint8, 5122: np.int16, 5123: np.uint16, 5125: np.uint32, 5126: np.float32}
_WIDTH = {"SCALAR": 1, "VEC2": 2, "VEC3": 3, "VEC4": 4}
UNUSED = 0xFFFF


class GltfError(ValueError):
    pass


class LhIndexedMesh(C.Structure):
    _fields_ = [("positions", C.c_void_p), ("position_stride", C.c_uint32), ("normals", C.c_void_p), ("normal_stride", C.c_uint32),
                ("uvs", C.c_void_p), ("uv_stride", C.c_uint32), ("num_vertices", C.c_uint64), ("indices", C.c_void_p),
                ("num_indices", C.c_uint64), ("index_bytes", C.c_uint32), ("flags", C.c_uint32), ("texture_scale", C.c_float * 2),
                ("texture_offset", C.c_float * 2), ("base_color_texture_index", C.c_uint16),
                ("metallic_roughness_texture_index", C.c_uint16), ("normal_texture_index", C.c_uint16), ("_pad", C.c_uint16),
                ("base_color_factor", C.c_float * 3)]


def node_matrix(node):
    """translate(scale(I, s), t), column-major: a point p goes to s * (p + t)."""
    if "matrix" in node:
        raise GltfError("node matrices are not read by the reference (TRS only)")
    s = np.array(node.get("scale", [1, 1, 1]), np.float32)
    t = np.array(node.get("translation", [0, 0, 0]), np.float32)
    m = np.zeros((4, 4), np.float32)          # m[column][row]
    m[0, 0], m[1, 1], m[2, 2], m[3, 3] = s[0], s[1], s[2], 1.0
    m[3, :3] = s * t
    return m


class _Asset:
    def __init__(self, path):
        self.dir = os.path.dirname(os.path.abspath(path))
        with open(path, "r") as f:
            self.j = json.load(f)
        self.buffers = []
        for b in self.j.get("buffers", []):
            if "uri" not in b or b["uri"].startswith("data:"):
                raise GltfError("buffers must be external files (mesh_loading.cpp:286-316)")
            raw = np.fromfile(os.path.join(self.dir, b["uri"]), np.uint8)
            if raw.size < b["byteLength"]:
                raise GltfError(f"buffer {b['uri']} is shorter than its byteLength")
            self.buffers.append(raw)

    def accessor(self, index, component, kind):
        """(base address array, stride, count, normalized) of an accessor whose type is checked like the reference's asserts."""
        a = self.j["accessors"][index]
        if _COMPONENT[a["componentType"]] != component or a["type"] != kind:
            raise GltfError(f"accessor {index}: expected {np.dtype(component).name} {kind}, found {a['componentType']} {a['type']}")
        view = self.j["bufferViews"][a["bufferView"]]
        offset = view.get("byteOffset", 0) + a.get("byteOffset", 0)
        element = np.dtype(component).itemsize * _WIDTH[kind]
        stride = view.get("byteStride", 0) or ((element + 3) & ~3 if kind == "VEC3" else element)   # quantised vec3s are padded to 4-byte multiples
        need = offset + (a["count"] - 1) * stride + element if a["count"] else offset
        buf = self.buffers[view["buffer"]]
        if need > buf.size:
            raise GltfError(f"accessor {index} runs past its buffer")
        return buf[offset:], stride, a["count"], bool(a.get("normalized", False))


def load(path):
    """Reads a .gltf file.  Returns (Scene, image_paths): image i of the asset goes to bindless slot i."""
    asset = _Asset(path)
    j = asset.j
    images = []
    for img in j.get("images", []):
        p = os.path.join(asset.dir, img.get("uri", ""))
        if os.path.splitext(p)[1].lower() not in (".dds", ".ktx2"):
            raise GltfError(f"image {img.get('uri')}: only .dds and .ktx2 are read (mesh_loading.cpp:327-356)")
        images.append(p)

    def slot(texture_info):
        return j["textures"][texture_info["index"]]["source"]

    scene = Scene()
    lib = scene.lib
    lib.lh_scene_add_mesh_indexed.argtypes = [C.c_void_p, C.POINTER(LhIndexedMesh)]
    lib.lh_scene_add_instance_matrix.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
    built = {}   # (mesh, primitive) -> mesh id: nodes that share a glTF mesh share its buffers and MeshInfo here (the
                 # reference rebuilds them per node, mesh_loading.cpp:361-368; the Instance records are the same)
    for node in j.get("nodes", []):
        if "mesh" not in node:
            continue
        matrix = np.ascontiguousarray(node_matrix(node).reshape(-1))
        for pi, prim in enumerate(j["meshes"][node["mesh"]]["primitives"]):
            attrs = prim["attributes"]
            if "TEXCOORD_0" not in attrs:
                continue                     # "missing uvs. Skipping."
            if (node["mesh"], pi) in built:
                if lib.lh_scene_add_instance_matrix(scene.handle, built[(node["mesh"], pi)], matrix.ctypes.data_as(C.POINTER(C.c_float))):
                    raise GltfError("lh_scene_add_instance_matrix failed")
                continue
            pos, pos_stride, count, _ = asset.accessor(attrs["POSITION"], np.uint16, "VEC3")
            nrm, nrm_stride, n_count, _ = asset.accessor(attrs["NORMAL"], np.int8, "VEC3")
            uvs, uv_stride, uv_count, uv_normalized = asset.accessor(attrs["TEXCOORD_0"], np.uint16, "VEC2")
            if not (count == n_count == uv_count):
                raise GltfError("attribute counts differ")
            ia = j["accessors"][prim["indices"]]
            wide = ia["componentType"] == 5125
            idx, idx_stride, n_idx, _ = asset.accessor(prim["indices"], np.uint32 if wide else np.uint16, "SCALAR")
            mat = j["materials"][prim["material"]]
            pbr = mat.get("pbrMetallicRoughness", {})
            d = LhIndexedMesh()
            scale, offset = np.ones(2, np.float32), np.zeros(2, np.float32)
            slots = [UNUSED, UNUSED, UNUSED]
            for k, info in enumerate((pbr.get("baseColorTexture"), pbr.get("metallicRoughnessTexture"), mat.get("normalTexture"))):
                if info is None:
                    continue
                slots[k] = slot(info)
                tt = info.get("extensions", {}).get("KHR_texture_transform")
                if tt is not None:           # the last transform seen wins (mesh_loading.cpp:500-552)
                    scale = np.array(tt.get("scale", [1, 1]), np.float32)
                    offset = np.array(tt.get("offset", [0, 0]), np.float32)
            if uv_normalized:
                scale = scale / np.float32(65535)
            if mat.get("alphaCutoff", 0.5) != 0.5:
                raise GltfError("alphaCutoff must be 0.5 (mesh_loading.cpp:567-570)")
            d.positions, d.position_stride = pos.ctypes.data, pos_stride
            d.normals, d.normal_stride = nrm.ctypes.data, nrm_stride
            d.uvs, d.uv_stride = uvs.ctypes.data, uv_stride
            d.num_vertices = count
            d.indices, d.num_indices, d.index_bytes = idx.ctypes.data, n_idx, 4 if wide else 2
            d.flags = S.LH_MESH_INFO_FLAGS_ALPHA_CLIP if mat.get("alphaMode") == "MASK" else 0
            d.texture_scale[:] = scale.tolist()
            d.texture_offset[:] = offset.tolist()
            d.base_color_texture_index, d.metallic_roughness_texture_index, d.normal_texture_index = slots
            d.base_color_factor[:] = [float(x) for x in pbr.get("baseColorFactor", [1, 1, 1, 1])[:3]]
            mesh = lib.lh_scene_add_mesh_indexed(scene.handle, C.byref(d))
            if mesh < 0:
                raise GltfError(f"lh_scene_add_mesh_indexed failed: {mesh}")
            built[(node["mesh"], pi)] = mesh
            rc = lib.lh_scene_add_instance_matrix(scene.handle, mesh, matrix.ctypes.data_as(C.POINTER(C.c_float)))
            if rc:
                raise GltfError("lh_scene_add_instance_matrix failed")
    return scene, images
