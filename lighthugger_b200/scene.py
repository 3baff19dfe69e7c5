"""ctypes binding of liblh_host.so (include/lh_scene.h): procedural scenes in
the reference's buffer layouts."""
import ctypes as C
import os

import numpy as np

from . import _structs as S

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def host_lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "lib", "liblh_host.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `make host` (or __graft_entry__.build())")
        lib = C.CDLL(path)
        vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
        lib.lh_scene_create.argtypes = [C.POINTER(vp)]
        lib.lh_scene_destroy.argtypes = [vp]
        lib.lh_scene_destroy.restype = None
        lib.lh_scene_add_mesh.argtypes = [vp, C.POINTER(S.LhMeshDesc)]
        lib.lh_scene_add_mesh_raw.argtypes = [vp, C.POINTER(S.LhMeshView)]
        lib.lh_scene_num_meshes.argtypes = [vp]
        lib.lh_scene_mesh_view.argtypes = [vp, C.c_int, C.POINTER(S.LhMeshView)]
        lib.lh_scene_place_instances.argtypes = [vp, u32, u64, C.POINTER(C.c_float), C.c_float, C.c_float,
                                                 C.POINTER(C.c_int32), u32]
        lib.lh_scene_place_instances_general.argtypes = lib.lh_scene_place_instances.argtypes
        lib.lh_scene_add_instance.argtypes = [vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_float]
        lib.lh_scene_num_instances.argtypes = [vp]
        lib.lh_scene_num_instances.restype = u32
        lib.lh_scene_total_triangles.argtypes = [vp]
        lib.lh_scene_total_triangles.restype = u64
        lib.lh_scene_total_meshlets.argtypes = [vp]
        lib.lh_scene_total_meshlets.restype = u64
        lib.lh_scene_uniforms.argtypes = [vp, C.POINTER(S.LhCamera), u32, u32, C.POINTER(S.LhUniforms)]
        lib.lh_camera_default.argtypes = [C.POINTER(S.LhCamera)]
        lib.lh_camera_default.restype = None
        lib.lh_scene_build_preset.argtypes = [vp, C.c_int, u32, u64, C.POINTER(S.LhPresetInfo)]
        lib.lh_scene_make_texture.argtypes = [C.c_int, u32, u32, C.c_int, vp, C.POINTER(u32)]
        lib.lh_scene_make_texture.restype = u64
        lib.lh_scene_make_skybox.argtypes = [u32, vp]
        lib.lh_scene_make_skybox.restype = u64
        lib.lh_scene_bind.argtypes = [vp, C.POINTER(S.LhCapacities), vp, vp, vp, vp, C.POINTER(S.LhUniforms), C.POINTER(vp)]
        lib.lh_binding_instances.argtypes = [vp]
        lib.lh_binding_instances.restype = vp
        lib.lh_binding_mesh_info_address.argtypes = [vp, C.c_int]
        lib.lh_binding_mesh_info_address.restype = u64
        lib.lh_binding_destroy.argtypes = [vp]
        lib.lh_binding_destroy.restype = None
        lib.lh_compute_bounding_sphere.argtypes = [C.POINTER(C.c_float), vp, u64]
        lib.lh_compute_bounding_sphere.restype = None
        lib.lh_compute_meshlet_bounds.argtypes = [vp, vp, u32, vp, C.POINTER(S.LhMeshlet)]
        lib.lh_compute_meshlet_bounds.restype = None
        lib.lh_bc7_encode.argtypes = [vp, u32, u32, vp]
        lib.lh_bc7_encode.restype = u64
        _lib = lib
    return _lib


PRESETS = {"tiny": 0, "c1": 1, "c2": 2, "c4": 4, "c5": 5}


class Scene:
    """A procedural scene (meshes + instances + camera)."""

    def __init__(self):
        self.lib = host_lib()
        self.handle = C.c_void_p()
        if self.lib.lh_scene_create(C.byref(self.handle)) != 0:
            raise RuntimeError("lh_scene_create failed")
        self.camera = S.LhCamera()
        self.lib.lh_camera_default(C.byref(self.camera))
        self.width, self.height = 256, 144
        self.num_texture_sets = 0
        self.alpha_cutout_set = 0xFFFFFFFF
        self._keep = []

    @classmethod
    def preset(cls, name, instance_count=0, seed=0, width=None, height=None):
        sc = cls()
        info = S.LhPresetInfo()
        rc = sc.lib.lh_scene_build_preset(sc.handle, PRESETS[name], instance_count, seed, C.byref(info))
        if rc != 0:
            raise RuntimeError(f"lh_scene_build_preset({name}) failed: {rc}")
        sc.camera = info.camera
        sc.width, sc.height = width or info.width, height or info.height
        sc.num_texture_sets = info.num_texture_sets
        sc.alpha_cutout_set = info.alpha_cutout_set
        return sc

    def close(self):
        if self.handle:
            self.lib.lh_scene_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_mesh(self, kind, su, sv, flags=0, textures=(0xFFFF, 0xFFFF, 0xFFFF), base_color=(1, 1, 1), uv_repeat=1.0):
        d = S.LhMeshDesc(kind=kind, segments_u=su, segments_v=sv, flags=flags, base_color_texture_index=textures[0],
                         metallic_roughness_texture_index=textures[1], normal_texture_index=textures[2], uv_repeat=uv_repeat)
        d.base_color_factor[:] = base_color
        mid = self.lib.lh_scene_add_mesh(self.handle, C.byref(d))
        if mid < 0:
            raise RuntimeError(f"lh_scene_add_mesh failed: {mid}")
        return mid

    def add_mesh_raw(self, view, keep=()):
        self._keep.extend(keep)
        mid = self.lib.lh_scene_add_mesh_raw(self.handle, C.byref(view))
        if mid < 0:
            raise RuntimeError("lh_scene_add_mesh_raw failed")
        return mid

    def mesh_view(self, mesh):
        v = S.LhMeshView()
        if self.lib.lh_scene_mesh_view(self.handle, mesh, C.byref(v)) != 0:
            raise IndexError(mesh)
        return v

    def mesh_triangles(self, mesh):
        """The triangle list (uint32 vertex indices, 3 per triangle) the mesh's meshlets were built from."""
        self.lib.lh_scene_mesh_triangles.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        self.lib.lh_scene_mesh_triangles.restype = C.c_uint64
        n = self.lib.lh_scene_mesh_triangles(self.handle, mesh, None)
        out = np.zeros(n, np.uint32)
        self.lib.lh_scene_mesh_triangles(self.handle, mesh, out.ctypes.data)
        return out

    def mesh_arrays(self, mesh):
        """numpy copies of one mesh's streams."""
        v = self.mesh_view(mesh)

        def arr(ptr, dtype, n):
            if n == 0:
                return np.zeros(0, dtype)
            return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(np.ctypeslib.as_ctypes_type(dtype))), (n,)).copy()

        nv = v.num_vertices
        return {
            "positions": arr(v.positions, np.uint16, nv * 3).reshape(-1, 3),
            "normals": arr(v.normals, np.int8, nv * 3).reshape(-1, 3),
            "uvs": arr(v.uvs, np.uint16, nv * 2).reshape(-1, 2),
            "indices": arr(v.indices, np.uint32 if v.index_bytes == 4 else np.uint16, v.num_indices),
            "micro_indices": arr(v.micro_indices, np.uint8, v.micro_indices_bytes),
            "meshlets": arr(v.meshlets, np.uint8, v.num_meshlets * 56).reshape(-1, 56),
            "num_triangles": int(v.num_triangles),
            "info": v.info,
        }

    def place_instances(self, count, extent, scale=(0.5, 2.0), mesh_ids=(0,), seed=0x5EED0001, general=False):
        """general: uniformly random rotations and per-axis scales instead of half-turns and uniform scales."""
        ext = (C.c_float * 3)(*extent)
        ids = (C.c_int32 * len(mesh_ids))(*mesh_ids)
        fn = self.lib.lh_scene_place_instances_general if general else self.lib.lh_scene_place_instances
        rc = fn(self.handle, count, seed, ext, scale[0], scale[1], ids, len(mesh_ids))
        if rc != 0:
            raise RuntimeError(f"lh_scene_place_instances failed: {rc}")

    def add_instance(self, mesh, translation=(0, 0, 0), quat=(0, 0, 0, 1), scale=1.0):
        t = (C.c_float * 3)(*translation)
        q = (C.c_float * 4)(*quat)
        if self.lib.lh_scene_add_instance(self.handle, mesh, t, q, scale) != 0:
            raise RuntimeError("lh_scene_add_instance failed")

    @property
    def num_instances(self):
        return int(self.lib.lh_scene_num_instances(self.handle))

    @property
    def num_meshes(self):
        return int(self.lib.lh_scene_num_meshes(self.handle))

    @property
    def total_triangles(self):
        return int(self.lib.lh_scene_total_triangles(self.handle))

    @property
    def total_meshlets(self):
        return int(self.lib.lh_scene_total_meshlets(self.handle))

    def uniforms(self, width=None, height=None, camera=None):
        u = S.LhUniforms()
        cam = camera or self.camera
        rc = self.lib.lh_scene_uniforms(self.handle, C.byref(cam), width or self.width, height or self.height, C.byref(u))
        if rc != 0:
            raise RuntimeError("lh_scene_uniforms failed")
        return u

    def default_capacities(self, slack=1.0):
        """Capacities that hold the whole scene (no overflow possible)."""
        n = max(1, self.num_instances)
        m = max(1, self.total_meshlets)
        return S.LhCapacities(max_instances=n, max_opaque_draws=int(m * slack) + 1, max_alpha_clip_draws=int(m * slack) + 1)

    # ---- textures ----
    def make_texture(self, kind, size, variant, alpha_cutout=False):
        """RGBA8 mip chain (largest first) as a flat uint8 array + level count."""
        levels = C.c_uint32()
        n = self.lib.lh_scene_make_texture(kind, size, variant, int(alpha_cutout), None, C.byref(levels))
        buf = np.zeros(n, np.uint8)
        self.lib.lh_scene_make_texture(kind, size, variant, int(alpha_cutout), buf.ctypes.data, C.byref(levels))
        return buf, int(levels.value)

    def texture_table(self, size=256):
        """The preset's texture table in slot order: (rgba8 mips, levels, srgb) per slot."""
        n = self.num_texture_sets
        out = []
        for kind in range(3):
            for v in range(n):
                buf, levels = self.make_texture(kind, size, v, alpha_cutout=(kind == 0 and v == self.alpha_cutout_set))
                out.append((buf, levels, size))
        return out

    def make_skybox(self, size=64):
        buf = np.zeros(6 * size * size * 4, np.float32)
        self.lib.lh_scene_make_skybox(size, buf.ctypes.data)
        return buf

    def bc7_encode(self, rgba8, width, height):
        """BC7 (mode 6) blocks of one RGBA8 level."""
        n = self.lib.lh_bc7_encode(rgba8.ctypes.data, width, height, None)
        out = np.zeros(n, np.uint8)
        self.lib.lh_bc7_encode(rgba8.ctypes.data, width, height, out.ctypes.data)
        return out


class HostBinding:
    """The scene placed in HOST memory (every address field is a host pointer):
    what the CPU oracle consumes."""

    def __init__(self, scene, caps=None):
        lib = scene.lib
        self.scene = scene
        self.caps = caps or scene.default_capacities()
        self.addresses = S.LhUniforms()
        self.handle = C.c_void_p()
        fn = lambda f: C.cast(f, C.c_void_p)
        rc = lib.lh_scene_bind(scene.handle, C.byref(self.caps), fn(lib.lh_host_alloc), fn(lib.lh_host_upload),
                               fn(lib.lh_host_free), None, C.byref(self.addresses), C.byref(self.handle))
        if rc != 0:
            raise RuntimeError(f"lh_scene_bind(host) failed: {rc}")

    def uniforms(self, **kw):
        u = self.scene.uniforms(**kw)
        for f in ("meshlet_references", "instances", "draw_calls", "misc_storage", "num_meshlets_prefix_sum", "dispatches"):
            setattr(u, f, getattr(self.addresses, f))
        return u

    def close(self):
        if self.handle:
            self.scene.lib.lh_binding_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
