"""ctypes binding of liblighthugger_b200.so (include/lighthugger_b200.h): the
C-ABI CUDA layer a C++20 host calls where the reference recorded Vulkan passes
(/root/reference/src/rendering.cpp:28-548).  No compute happens here and there
is no CPU fallback: everything raises if the library or a GPU is missing."""
import ctypes as C
import os

import numpy as np

from . import _structs as S

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

EXPORTS = [
    "lh_config_default", "lh_buffer_sizes", "lh_context_create", "lh_context_destroy", "lh_last_error", "lh_version",
    "lh_buffer_alloc", "lh_buffer_free", "lh_upload", "lh_download", "lh_texture_create", "lh_texture_destroy",
    "lh_skybox_create", "lh_targets_resize", "lh_set_instance_range", "lh_frame_begin", "lh_reset", "lh_cull_instances",
    "lh_cull_meshlets", "lh_raster_visbuffer", "lh_composite", "lh_reduce_depth", "lh_generate_shadow_matrices",
    "lh_cull_instances_shadows", "lh_shadow_cascade", "lh_resolve", "lh_display_transform", "lh_render_frame",
    "lh_frame_sync", "lh_render_frame_host", "lh_get_frame_stats", "lh_set_config_flags", "lh_get_target", "lh_get_timings",
    "lh_get_target_address", "lh_get_launch_count", "lh_timer_start", "lh_timer_stop", "lh_flush_l2", "lh_comm_unique_id", "lh_comm_init",
    "lh_comm_destroy", "lh_comm_info", "lh_comm_set_exchange", "lh_comm_remap", "lh_comm_rows", "lh_comm_set_rows", "lh_set_shadow_instance_range", "lh_set_instance_stripes", "lh_set_instance_stripe_bounds", "lh_set_display_lut", "lh_set_shadow_cascade_mask", "lh_shadow_exchange",
]


class LhError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"lighthugger_b200 error {code}: {message}")
        self.code = code


def cuda_lib_path():
    # LH_CUDA_LIB: developer hook for A/B-testing a differently built library (still the CUDA library, never a fallback)
    return os.environ.get("LH_CUDA_LIB") or os.path.join(_HERE, "lib", "liblighthugger_b200.so")


def cuda_lib():
    global _lib
    if _lib is None:
        path = cuda_lib_path()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `make cuda` (or __graft_entry__.build()); there is no CPU fallback")
        lib = C.CDLL(path)
        vp, u64, u32, u16 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint16
        lib.lh_config_default.argtypes = [C.POINTER(S.LhConfig)]
        lib.lh_buffer_sizes.argtypes = [C.POINTER(S.LhConfig), C.POINTER(S.LhBufferSizes)]
        lib.lh_context_create.argtypes = [C.POINTER(S.LhConfig), C.POINTER(vp)]
        lib.lh_context_destroy.argtypes = [vp]
        lib.lh_last_error.argtypes = [vp]
        lib.lh_last_error.restype = C.c_char_p
        lib.lh_version.restype = C.c_char_p
        lib.lh_buffer_alloc.argtypes = [vp, u64, C.POINTER(u64)]
        lib.lh_buffer_free.argtypes = [vp, u64]
        lib.lh_upload.argtypes = [vp, u64, vp, u64]
        lib.lh_download.argtypes = [vp, u64, vp, u64]
        lib.lh_texture_create.argtypes = [vp, C.c_int, u32, u32, u32, vp, u64, C.POINTER(u16)]
        lib.lh_texture_destroy.argtypes = [vp, u16]
        lib.lh_skybox_create.argtypes = [vp, u32, vp]
        lib.lh_set_display_lut.argtypes = [vp, u32, vp]
        lib.lh_targets_resize.argtypes = [vp, u32, u32]
        lib.lh_set_instance_range.argtypes = [vp, u32, u32]
        lib.lh_set_shadow_cascade_mask.argtypes = [vp, u32]
        lib.lh_set_shadow_instance_range.argtypes = [vp, u32, u32]
        lib.lh_comm_set_exchange.argtypes = [vp, C.c_int]
        lib.lh_comm_remap.argtypes = [vp]
        if hasattr(lib, "lh_set_instance_stripes") or not os.environ.get("LH_CUDA_LIB"):   # an older A/B build may lack it
            lib.lh_set_instance_stripes.argtypes = [vp, C.c_int, u32, u32, u32]
        if hasattr(lib, "lh_set_instance_stripe_bounds"):
            lib.lh_set_instance_stripe_bounds.argtypes = [vp, C.c_int, C.POINTER(C.c_float), u32, u32, u32]
        lib.lh_comm_rows.argtypes = [vp, C.POINTER(u32), C.POINTER(u32)]
        lib.lh_comm_set_rows.argtypes = [vp, u32, u32]
        lib.lh_frame_begin.argtypes = [vp, C.POINTER(S.LhUniforms)]
        for name in ("lh_reset", "lh_cull_instances", "lh_cull_meshlets", "lh_raster_visbuffer", "lh_composite",
                     "lh_reduce_depth", "lh_generate_shadow_matrices", "lh_cull_instances_shadows", "lh_resolve",
                     "lh_display_transform", "lh_render_frame", "lh_frame_sync", "lh_timer_start", "lh_flush_l2",
                     "lh_comm_destroy", "lh_shadow_exchange"):
            getattr(lib, name).argtypes = [vp]
        lib.lh_shadow_cascade.argtypes = [vp, u32]
        lib.lh_render_frame_host.argtypes = [vp, C.POINTER(S.LhUniforms), vp, u32, C.POINTER(S.LhFrameStats)]
        lib.lh_get_frame_stats.argtypes = [vp, C.POINTER(S.LhFrameStats)]
        lib.lh_set_config_flags.argtypes = [vp, C.c_uint32]
        lib.lh_get_target.argtypes = [vp, C.c_int, vp, u64]
        lib.lh_get_timings.argtypes = [vp, C.POINTER(S.LhPassTimes)]
        lib.lh_get_target_address.argtypes = [vp, C.c_int, C.POINTER(u64)]
        lib.lh_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
        lib.lh_get_launch_count.argtypes = [vp, C.POINTER(u64)]
        lib.lh_comm_unique_id.argtypes = [vp]
        lib.lh_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
        lib.lh_comm_info.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib = lib
    return _lib


class Context:
    """One GPU context (one CUDA stream, externally synchronised), mirroring the
    per-frame sequence of the reference's main loop (src/main.cpp:784-817)."""

    def __init__(self, width, height, caps=None, device=0, flags=None, shadowmap_size=1024, max_large_triangles=0):
        self.lib = cuda_lib()
        cfg = S.LhConfig()
        self.lib.lh_config_default(C.byref(cfg))
        cfg.device = device
        cfg.width, cfg.height = width, height
        cfg.shadowmap_size = shadowmap_size
        if caps is not None:
            cfg.max_instances = caps.max_instances
            cfg.max_opaque_draws = caps.max_opaque_draws
            cfg.max_alpha_clip_draws = caps.max_alpha_clip_draws
        if flags is not None:
            cfg.flags = flags
        if max_large_triangles:
            cfg.max_large_triangles = max_large_triangles
        self.cfg = cfg
        self.caps = S.LhCapacities(cfg.max_instances, cfg.max_opaque_draws, cfg.max_alpha_clip_draws)
        self.handle = C.c_void_p()
        rc = self.lib.lh_context_create(C.byref(cfg), C.byref(self.handle))
        if rc != 0:
            raise LhError(rc, (self.lib.lh_last_error(None) or b"").decode())
        self.width, self.height = width, height
        self.binding = None
        self.scene = None
        self.addresses = None

    def _check(self, rc):
        if rc != 0:
            raise LhError(rc, (self.lib.lh_last_error(self.handle) or b"").decode())

    def close(self):
        if getattr(self, "handle", None):
            if self.binding:
                self.scene.lib.lh_binding_destroy(self.binding)
                self.binding = None
            self.lib.lh_context_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- scene upload: lh_buffer_alloc / lh_upload are passed straight to the host binder ----
    def bind_scene(self, scene):
        hl = scene.lib
        self.scene = scene
        self.addresses = S.LhUniforms()
        self.binding = C.c_void_p()
        fn = lambda f: C.cast(f, C.c_void_p)
        rc = hl.lh_scene_bind(scene.handle, C.byref(self.caps), fn(self.lib.lh_buffer_alloc), fn(self.lib.lh_upload),
                              fn(self.lib.lh_buffer_free), self.handle, C.byref(self.addresses), C.byref(self.binding))
        if rc != 0:
            raise LhError(rc, (self.lib.lh_last_error(self.handle) or b"").decode())

    def uniforms(self, **kw):
        u = self.scene.uniforms(**({"width": self.width, "height": self.height} | kw))
        for f in ("meshlet_references", "instances", "draw_calls", "misc_storage", "num_meshlets_prefix_sum", "dispatches"):
            setattr(u, f, getattr(self.addresses, f))
        return u

    def host_instances(self):
        """(pointer, count) of the host copy of the instance array (for lh_render_frame_host)."""
        return self.scene.lib.lh_binding_instances(self.binding), self.scene.num_instances

    def create_texture(self, fmt, width, height, levels, data):
        slot = C.c_uint16()
        data = np.ascontiguousarray(data)
        self._check(self.lib.lh_texture_create(self.handle, fmt, width, height, levels, data.ctypes.data, data.nbytes,
                                               C.byref(slot)))
        return int(slot.value)

    def create_skybox(self, size, rgba32f):
        rgba32f = np.ascontiguousarray(rgba32f, np.float32)
        self._check(self.lib.lh_skybox_create(self.handle, size, rgba32f.ctypes.data))

    def set_display_lut(self, size, rgb32f):
        rgb32f = np.ascontiguousarray(rgb32f, np.float32)
        self._check(self.lib.lh_set_display_lut(self.handle, size, rgb32f.ctypes.data))

    def set_instance_range(self, first, count):
        self._check(self.lib.lh_set_instance_range(self.handle, first, count))

    def set_shadow_instance_range(self, first, count):
        self._check(self.lib.lh_set_shadow_instance_range(self.handle, first, count))

    def set_instance_stripe_bounds(self, view, bounds, index, count):
        arr = (C.c_float * len(bounds))(*[float(b) for b in bounds])
        self._check(self.lib.lh_set_instance_stripe_bounds(self.handle, view, arr, len(bounds) - 1, index, count))

    def set_instance_stripes(self, view, stripes, index=0, count=1):
        self._check(self.lib.lh_set_instance_stripes(self.handle, view, stripes, index, count))

    def comm_set_exchange(self, mode):
        self._check(self.lib.lh_comm_set_exchange(self.handle, mode))

    def resize(self, width, height):
        """lh_targets_resize (sort-last ranks call it together: it is collective under LH_EXCHANGE_PEER)."""
        self._check(self.lib.lh_targets_resize(self.handle, width, height))
        self.width, self.height = width, height

    def comm_remap(self):
        """Collective: drop every peer mapping (call on all ranks before moving the MeshletReference buffer)."""
        self._check(self.lib.lh_comm_remap(self.handle))

    def comm_set_rows(self, y0, y1):
        self._check(self.lib.lh_comm_set_rows(self.handle, int(y0), int(y1)))

    def comm_rows(self):
        y0, y1 = C.c_uint32(0), C.c_uint32(0)
        self._check(self.lib.lh_comm_rows(self.handle, C.byref(y0), C.byref(y1)))
        return y0.value, y1.value

    def set_shadow_cascade_mask(self, mask):
        self._check(self.lib.lh_set_shadow_cascade_mask(self.handle, mask))

    # ---- per frame ----
    def frame_begin(self, u):
        self._check(self.lib.lh_frame_begin(self.handle, C.byref(u)))

    def _pass(self, name, *args):
        self._check(getattr(self.lib, name)(self.handle, *args))

    def reset(self): self._pass("lh_reset")
    def cull_instances(self): self._pass("lh_cull_instances")
    def cull_meshlets(self): self._pass("lh_cull_meshlets")
    def raster_visbuffer(self): self._pass("lh_raster_visbuffer")
    def composite(self): self._pass("lh_composite")
    def reduce_depth(self): self._pass("lh_reduce_depth")
    def generate_shadow_matrices(self): self._pass("lh_generate_shadow_matrices")
    def cull_instances_shadows(self): self._pass("lh_cull_instances_shadows")
    def shadow_cascade(self, c): self._pass("lh_shadow_cascade", c)
    def resolve(self): self._pass("lh_resolve")
    def display_transform(self): self._pass("lh_display_transform")
    def render_frame(self): self._pass("lh_render_frame")
    def sync(self): self._pass("lh_frame_sync")
    def flush_l2(self): self._pass("lh_flush_l2")
    def timer_start(self): self._pass("lh_timer_start")

    def timer_stop(self):
        ms = C.c_float()
        self._check(self.lib.lh_timer_stop(self.handle, C.byref(ms)))
        return float(ms.value)

    def render_frame_host(self, u, upload_instances=True):
        stats = S.LhFrameStats()
        ptr, n = self.host_instances() if upload_instances else (None, 0)
        self._check(self.lib.lh_render_frame_host(self.handle, C.byref(u), ptr, n, C.byref(stats)))
        return stats

    def set_config_flags(self, flags):
        self._check(self.lib.lh_set_config_flags(self.handle, flags))
        self.cfg.flags = flags

    # ---- readback ----
    def stats(self):
        st = S.LhFrameStats()
        self._check(self.lib.lh_get_frame_stats(self.handle, C.byref(st)))
        return st

    def timings(self):
        t = S.LhPassTimes()
        self._check(self.lib.lh_get_timings(self.handle, C.byref(t)))
        return {n: float(t.ms[i]) for i, n in enumerate(S.LH_PASS_NAMES)} | {
            "frame": float(t.frame_ms), "raster_meshlets_kernel": float(t.raster_meshlets_kernel_ms),
                "raster_meshlets_kernel_launches": int(t.raster_meshlets_kernel_launches)}

    def launch_count(self):
        n = C.c_uint64()
        self._check(self.lib.lh_get_launch_count(self.handle, C.byref(n)))
        return int(n.value)

    def download(self, addr, nbytes, dtype=np.uint8):
        out = np.zeros(nbytes // np.dtype(dtype).itemsize, dtype)
        self._check(self.lib.lh_download(self.handle, addr, out.ctypes.data, out.nbytes))
        return out

    def upload(self, addr, array):
        array = np.ascontiguousarray(array)
        self._check(self.lib.lh_upload(self.handle, addr, array.ctypes.data, array.nbytes))

    def target(self, which):
        n = self.width * self.height
        sm = self.cfg.shadowmap_size
        shape_dtype = {
            S.LH_TARGET_VIS64: (n, np.uint64), S.LH_TARGET_DEPTH: (n, np.float32),
            S.LH_TARGET_SCENE_RGBA16F: (n * 4, np.float16), S.LH_TARGET_SHADOW: (4 * sm * sm, np.float32),
            S.LH_TARGET_DISPLAY_BGRA8: (n * 4, np.uint8),
        }[which]
        out = np.zeros(shape_dtype[0], shape_dtype[1])
        self._check(self.lib.lh_get_target(self.handle, which, out.ctypes.data, out.nbytes))
        return out

    # views of the pipeline buffers (the test-side checker exposes the same shapes)
    def sector(self, cascade=0):
        base = self.addresses.num_meshlets_prefix_sum + cascade * (self.cfg.max_instances * 8 + 8)
        counter = int(self.download(base, 8, np.uint64)[0])
        count, total = counter >> 32, counter & 0xFFFFFFFF
        vals = self.download(base + 8, 8 * min(count, self.cfg.max_instances), np.uint32).reshape(-1, 2)
        return count, total, vals

    def draw_counts(self):
        d = self.download(self.addresses.draw_calls, 8, np.uint32)
        return int(d[0]), int(d[1])

    def draw_commands(self, first, count):
        return self.download(self.addresses.draw_calls + 8 + 16 * first, 16 * count, np.uint32).reshape(-1, 4)

    def meshlet_references(self, first, count):
        raw = self.download(self.addresses.meshlet_references + 8 * first, 8 * count, np.uint32).reshape(-1, 2).copy()
        raw[:, 1] &= 0xFFFF
        return raw

    def misc(self):
        return S.LhMiscStorage.from_buffer_copy(self.download(self.addresses.misc_storage, 792).tobytes())

    def dispatches(self):
        return self.download(self.addresses.dispatches, 36, np.uint32).reshape(3, 3)

    # ---- multi-GPU ----
    def comm_unique_id(self):
        buf = (C.c_uint8 * 128)()
        self._check(self.lib.lh_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, rank, nranks, unique_id):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        self._check(self.lib.lh_comm_init(self.handle, rank, nranks, buf))
