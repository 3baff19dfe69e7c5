"""ctypes binding of oracle/liblh_oracle.so -- the CPU restatement of the
reference (TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs import this)."""
import ctypes as C
import os

import numpy as np

from lighthugger_b200 import _structs as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_lib = None


class LhoTexture(C.Structure):
    _fields_ = [("rgba8", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("mip_levels", C.c_uint32),
                ("srgb", C.c_uint32)]


class LhoConfig(C.Structure):
    _fields_ = [("max_instances", C.c_uint32), ("max_opaque_draws", C.c_uint32), ("max_alpha_clip_draws", C.c_uint32),
                ("shadowmap_size", C.c_uint32), ("threads", C.c_uint32), ("instance_first", C.c_uint32),
                ("instance_count", C.c_uint32), ("textures", C.POINTER(LhoTexture)), ("num_textures", C.c_uint32),
                ("skybox", C.c_void_p), ("skybox_size", C.c_uint32)]


def oracle_lib():
    global _lib
    if _lib is None:
        # LH_ORACLE_LIB: a differently built oracle (the sanitizer test builds one with -fsanitize=address,undefined)
        path = os.environ.get("LH_ORACLE_LIB") or os.path.join(ROOT, "oracle", "liblh_oracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `make oracle`")
        lib = C.CDLL(path)
        U, K = C.POINTER(S.LhUniforms), C.POINTER(LhoConfig)
        vp = C.c_void_p
        lib.lho_config_default.argtypes = [K]
        lib.lho_config_default.restype = None
        for name in ("lho_reset", "lho_cull_instances", "lho_cull_meshlets", "lho_generate_shadow_matrices",
                     "lho_cull_instances_shadows"):
            getattr(lib, name).argtypes = [U, K]
        lib.lho_raster_visbuffer.argtypes = [U, K, vp]
        lib.lho_reduce_depth.argtypes = [U, K, vp]
        lib.lho_shadow_cascade.argtypes = [U, K, C.c_uint32, vp]
        lib.lho_resolve.argtypes = [U, K, vp, vp, vp]
        lib.lho_render_frame.argtypes = [U, K, vp, vp, vp]
        lib.lho_display_transform.argtypes = [U, K, vp, vp, C.c_uint32, vp]
        lib.lho_resolve_debug.argtypes = [U, K, vp, vp, C.c_int, vp]
        lib.lho_resolve_pixels.argtypes = [U, K, vp, vp, C.c_uint64, vp, vp, vp]
        lib.lho_composite_min.argtypes = [vp, vp, C.c_uint64]
        lib.lho_composite_min.restype = None
        lib.lho_pinned_powf.argtypes = [C.c_float, C.c_float]
        lib.lho_pinned_powf.restype = C.c_float
        _lib = lib
    return _lib


def _view(addr, ctype, count):
    return np.ctypeslib.as_array(C.cast(addr, C.POINTER(ctype)), (count,))


class Oracle:
    """Runs the oracle passes over a HostBinding and exposes the results as numpy arrays."""

    def __init__(self, binding, width=None, height=None, threads=0, textures=None, skybox=None, shadowmap_size=1024,
                 instance_range=None, camera=None):
        self.lib = oracle_lib()
        self.binding = binding
        self.u = binding.uniforms(width=width, height=height, camera=camera)
        self.cfg = LhoConfig()
        self.lib.lho_config_default(C.byref(self.cfg))
        caps = binding.caps
        self.cfg.max_instances = caps.max_instances
        self.cfg.max_opaque_draws = caps.max_opaque_draws
        self.cfg.max_alpha_clip_draws = caps.max_alpha_clip_draws
        self.cfg.shadowmap_size = shadowmap_size
        self.cfg.threads = threads
        if instance_range is not None:
            self.cfg.instance_first, self.cfg.instance_count = instance_range
        self._tex_keep = []
        if textures:
            arr = (LhoTexture * len(textures))()
            for i, (rgba8, levels, size, srgb) in enumerate(textures):
                self._tex_keep.append(rgba8)
                arr[i] = LhoTexture(rgba8.ctypes.data, size, size, levels, int(srgb))
            self._tex_arr = arr
            self.cfg.textures = arr
            self.cfg.num_textures = len(textures)
        if skybox is not None:
            self._sky = skybox
            self.cfg.skybox = skybox.ctypes.data
            self.cfg.skybox_size = int(round((skybox.size // 24) ** 0.5))
        w, h = self.u.window_size[0], self.u.window_size[1]
        self.width, self.height = w, h
        self.vis = np.full(w * h, S.LH_VIS_CLEAR, np.uint64)
        self.shadow = np.ones(4 * shadowmap_size * shadowmap_size, np.float32)
        self.color = None

    def _call(self, name, *args):
        return getattr(self.lib, name)(C.byref(self.u), C.byref(self.cfg), *args)

    def reset(self): return self._call("lho_reset")
    def cull_instances(self): return self._call("lho_cull_instances")
    def cull_meshlets(self): return self._call("lho_cull_meshlets")
    def raster_visbuffer(self): return self._call("lho_raster_visbuffer", self.vis.ctypes.data)
    def reduce_depth(self): return self._call("lho_reduce_depth", self.vis.ctypes.data)
    def generate_shadow_matrices(self): return self._call("lho_generate_shadow_matrices")
    def cull_instances_shadows(self): return self._call("lho_cull_instances_shadows")
    def shadow_cascade(self, c): return self._call("lho_shadow_cascade", c, self.shadow.ctypes.data)

    def resolve(self):
        self.color = np.zeros(self.width * self.height * 4, np.float32)
        return self._call("lho_resolve", self.vis.ctypes.data, self.shadow.ctypes.data, self.color.ctypes.data)

    def resolve_debug(self, term):
        out = np.zeros(self.width * self.height * 4, np.float32)
        self._call("lho_resolve_debug", self.vis.ctypes.data, self.shadow.ctypes.data, term, out.ctypes.data)
        return out.reshape(-1, 4)

    def resolve_pixels(self, pixels, bias=(0.0, 0.0, 0.0)):
        """The lit colour of the listed pixels with `bias` added to every material texture sample (conditioning probe)."""
        px = np.ascontiguousarray(pixels, np.uint32)
        b = np.asarray(bias, np.float32)
        out = np.zeros((len(px), 4), np.float32)
        self._call("lho_resolve_pixels", self.vis.ctypes.data, self.shadow.ctypes.data, len(px), px.ctypes.data, b.ctypes.data, out.ctypes.data)
        return out

    def display_transform(self, scene_rgba32f, lut_rgb32f, lut_size=48):
        """display_transform.comp on a scene-referred RGBA32F image: the sRGB-encoded RGBA32F the reference stores."""
        scene = np.ascontiguousarray(scene_rgba32f, np.float32).reshape(-1)
        lut = np.ascontiguousarray(lut_rgb32f, np.float32).reshape(-1)
        out = np.zeros(self.width * self.height * 4, np.float32)
        assert scene.size == out.size and lut.size == 3 * lut_size ** 3
        self._call("lho_display_transform", scene.ctypes.data, lut.ctypes.data, lut_size, out.ctypes.data)
        return out.reshape(-1, 4)

    def render_frame(self, resolve=True):
        """Steps 1-11 of rendering.cpp, pass by pass (so the main view's draw counts can be noted before the cascades
        reuse the counters)."""
        self.reset(); self.cull_instances(); self.cull_meshlets()
        self.main_draw_counts = self.draw_counts()
        self.raster_visbuffer(); self.reduce_depth(); self.generate_shadow_matrices(); self.cull_instances_shadows()
        for c in range(4):
            self.shadow_cascade(c)
        if resolve:
            self.resolve()
        return 0

    # ---- views into the host-bound pipeline buffers ----
    def sector(self, cascade=0):
        """(count, total, values[count] as (index, sum) uint32 pairs) of prefix-sum sector `cascade`."""
        base = self.u.num_meshlets_prefix_sum + cascade * (self.cfg.max_instances * 8 + 8)
        counter = int(_view(base, C.c_uint64, 1)[0])
        count, total = counter >> 32, counter & 0xFFFFFFFF
        vals = _view(base + 8, C.c_uint32, 2 * min(count, self.cfg.max_instances)).reshape(-1, 2).copy()
        return count, total, vals

    def draw_counts(self):
        d = _view(self.u.draw_calls, C.c_uint32, 2)
        return int(d[0]), int(d[1])

    def draw_commands(self, first, count):
        return _view(self.u.draw_calls + 8 + 16 * first, C.c_uint32, 4 * count).reshape(-1, 4).copy()

    def meshlet_references(self, first, count):
        """(instance_index u32, meshlet_index u16) pairs as a uint32 [count, 2] array."""
        raw = _view(self.u.meshlet_references + 8 * first, C.c_uint32, 2 * count).reshape(-1, 2).copy()
        raw[:, 1] &= 0xFFFF
        return raw

    def misc(self):
        return S.LhMiscStorage.from_buffer_copy(_view(self.u.misc_storage, C.c_uint8, 792).tobytes())

    def dispatches(self):
        return _view(self.u.dispatches, C.c_uint32, 9).reshape(3, 3).copy()
