"""Pins the oracle against the REFERENCE'S OWN GLSL: oracle/_ref/libref_shaders.so is the reference's shader source
compiled in place as C++ (oracle/ref_shim/, Makefile target `ref`) and run on the host over the same host-bound
buffers the oracle uses.  Whole compute passes (reset, instance cull + prefix-sum append, meshlet cull / draw-call
writing, cascade matrices, shadow instance cull, per-cascade draw-call writing) must leave byte-identical buffers;
single functions are compared on random inputs.  The capacities are the reference's compile-time ones
(MAX_INSTANCES 1024, 200 000 + 200 000 draws), because its shaders hard-code the sector and section strides.

The library is built where /root/reference exists and travels to the GPU box prebuilt; these tests need no GPU."""
import ctypes as C
import os

import numpy as np
import pytest

from lighthugger_b200 import _structs as S
from lighthugger_b200.scene import Scene
from helpers import golden_torus_scene, make_oracle, rotated_scene

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libref_shaders.so")
pytestmark = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_shaders.so not built (needs /root/reference: make ref)")

U = C.POINTER(S.LhUniforms)
_ref = None


def ref_lib():
    global _ref
    if _ref is None:
        lib = C.CDLL(LIB)
        for n in ("ref_sh_reset_buffers_a", "ref_sh_cull_instances", "ref_sh_write_draw_calls", "ref_sh_generate_shadow_matrices",
                  "ref_sh_cull_instances_shadows"):
            getattr(lib, n).argtypes = [U]
            getattr(lib, n).restype = None
        lib.ref_sh_write_draw_calls_shadows.argtypes = [U, C.c_uint32]
        lib.ref_sh_write_draw_calls_shadows.restype = None
        fp = C.POINTER(C.c_float)
        lib.ref_sh_cull_bounding_sphere.argtypes = [U, C.c_void_p, fp]
        lib.ref_sh_cull_bounding_sphere_shadows.argtypes = [U, C.c_void_p, fp, C.c_uint32]
        lib.ref_sh_cull_cone_perspective.argtypes = [U, C.c_void_p, C.c_void_p]
        lib.ref_sh_cull_cone_orthographic.argtypes = [U, C.c_void_p, C.c_void_p]
        lib.ref_sh_struct_sizes.argtypes = [C.POINTER(C.c_uint32)]
        _ref = lib
    return _ref


def reference_caps():
    return S.LhCapacities(max_instances=S.LH_MAX_INSTANCES, max_opaque_draws=S.LH_MAX_OPAQUE_DRAWS,
                          max_alpha_clip_draws=S.LH_MAX_ALPHA_CLIP_DRAWS)


def test_reference_structs_have_the_layout_of_ours():
    sizes = (C.c_uint32 * 9)()
    ref_lib().ref_sh_struct_sizes(sizes)
    ours = [C.sizeof(t) for t in (S.LhInstance, S.LhMeshInfo, S.LhMeshlet, S.LhMeshletReference, S.LhPrefixSumValue,
                                  S.LhDrawIndirectCommand, S.LhDispatchIndirectCommand, S.LhMiscStorage, S.LhUniforms)]
    assert list(sizes) == ours


def _lists(o, section=0):
    n_op, n_al = o.draw_counts()
    first = section * (o.cfg.max_opaque_draws + o.cfg.max_alpha_clip_draws)
    return (n_op, n_al, o.meshlet_references(first, n_op).tobytes(), o.meshlet_references(first + o.cfg.max_opaque_draws, n_al).tobytes(),
            o.draw_commands(0, n_op).tobytes(), o.draw_commands(o.cfg.max_opaque_draws, n_al).tobytes())


def _scene(name):
    if name == "golden":
        return golden_torus_scene(), {}
    if name == "rotated":
        return rotated_scene(), {}
    if name == "c1":
        return Scene.preset("c1"), dict(width=960, height=540)
    if name == "c4small":   # mixed meshes incl. alpha-clip foliage and 32-bit indices
        return Scene.preset("c4", instance_count=120), dict(width=480, height=270)
    return Scene.preset(name), {}


@pytest.mark.parametrize("preset", ["tiny", "golden", "rotated", "c1", "c4small"])
def test_compute_passes_match_the_reference_shaders(preset):
    """The frame's compute passes, oracle on one binding and the reference's GLSL on another, compared pass by pass
    (the raster passes in between are the oracle's on both: the reference rasterises with fixed-function hardware)."""
    ref = ref_lib()
    sc, kw = _scene(preset)
    a = make_oracle(sc, caps=reference_caps(), **kw)
    b = make_oracle(sc, caps=reference_caps(), **kw)
    ub = C.byref(b.u)

    a.reset()
    ref.ref_sh_reset_buffers_a(ub)                              # compute/reset_buffers.glsl:9-29
    assert np.array_equal(a.dispatches()[:2], b.dispatches()[:2])
    assert bytes(a.misc())[784:] == bytes(b.misc())[784:]

    a.cull_instances()
    ref.ref_sh_cull_instances(ub)                               # cull_instances.comp:7-27, culling.glsl:1-46, prefix_sum.glsl:17-30
    ca, ta, va = a.sector(0)
    cb, tb, vb = b.sector(0)
    assert (ca, ta) == (cb, tb) and np.array_equal(va, vb)
    assert 0 < ca <= S.LH_MAX_INSTANCES, "the scene must stay inside the reference's MAX_INSTANCES for this comparison"

    a.cull_meshlets()
    ref.ref_sh_write_draw_calls(ub)                             # write_draw_calls.comp:7-57, util.glsl:42-59, culling.glsl:137-145
    assert np.array_equal(a.dispatches(), b.dispatches())
    la, lb = _lists(a), _lists(b)
    assert la[0] > 0 and la[0] <= S.LH_MAX_OPAQUE_DRAWS
    assert la == lb

    for o in (a, b):                                            # fixed-function raster + read_depth: the oracle's on both
        o.raster_visbuffer()
        o.reduce_depth()
    a.generate_shadow_matrices()
    ref.ref_sh_generate_shadow_matrices(ub)                     # compute/generate_shadow_matrices.glsl:15-110, matrices.glsl:9-40
    assert bytes(a.misc()) == bytes(b.misc())

    a.cull_instances_shadows()
    ref.ref_sh_cull_instances_shadows(ub)                       # cull_instances.comp:31-56, culling.glsl:48-135
    for c in range(4):
        sa, sb = a.sector(c), b.sector(c)
        assert sa[:2] == sb[:2] and np.array_equal(sa[2], sb[2]), f"cascade {c}"
    assert sum(a.sector(c)[0] for c in range(4)) > 0

    for c in range(4):
        a.shadow_cascade(c)
        ref.ref_sh_write_draw_calls_shadows(ub, c)              # reset_buffers.glsl:43-57, write_draw_calls.comp:61-117
        assert _lists(a, 1) == _lists(b, 1), f"cascade {c}"
        assert np.array_equal(a.dispatches(), b.dispatches())


def _random_uniforms(rng, sc, o):
    """The scene's uniforms with a random camera pose (so the frustum planes and view matrices are exercised)."""
    cam = S.LhCamera.from_buffer_copy(bytes(sc.camera))
    cam.position[0] += float(rng.uniform(-20, 20))
    cam.position[1] += float(rng.uniform(-5, 5))
    cam.position[2] += float(rng.uniform(-20, 20))
    cam.yaw += float(rng.uniform(-3, 3))
    cam.pitch = float(rng.uniform(-1.2, 1.2))
    return o.binding.uniforms(camera=cam)


def test_cull_functions_match_the_reference_on_random_inputs():
    """cull_bounding_sphere, cull_bounding_sphere_shadows, cull_cone_perspective, cull_cone_orthographic on 10^5 random
    (instance, sphere / meshlet) pairs with general rotations and non-uniform scales, against the oracle's exports."""
    from oracle_api import oracle_lib
    ref, orc = ref_lib(), oracle_lib()
    for n in ("lho_test_cull_sphere", "lho_test_cull_sphere_shadows", "lho_test_cull_cone_perspective", "lho_test_cull_cone_orthographic"):
        assert hasattr(orc, n)
    orc.lho_test_cull_sphere.argtypes = [U, C.c_void_p, C.POINTER(C.c_float)]
    orc.lho_test_cull_sphere_shadows.argtypes = [U, C.c_void_p, C.POINTER(C.c_float), C.c_uint32]
    orc.lho_test_cull_cone_perspective.argtypes = [U, C.c_void_p, C.c_void_p]
    orc.lho_test_cull_cone_orthographic.argtypes = [U, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(0x5EED0002)
    sc = Scene.preset("tiny")
    o = make_oracle(sc, caps=reference_caps())
    o.render_frame(resolve=False)           # fills MiscStorage with real cascade matrices
    N = 100_000
    flips = np.zeros(4, np.int64)
    kept = np.zeros(4, np.int64)
    inst = S.LhInstance()
    meshlet = S.LhMeshlet()
    sphere = (C.c_float * 4)()
    u = o.u
    for i in range(N):
        if i % 500 == 0:
            u = _random_uniforms(rng, sc, o)
        # general affine transform: random rotation (QR), non-uniform scale, translation
        q, _ = np.linalg.qr(rng.normal(size=(3, 3)))
        m3 = (q * rng.uniform(0.2, 3.0, 3)).astype(np.float32)
        t = rng.uniform(-30, 30, 3).astype(np.float32)
        T = np.eye(4, dtype=np.float32)
        T[:3, :3] = m3
        T[:3, 3] = t
        Ninv = np.linalg.inv(T.astype(np.float64))[:3, :3].astype(np.float32)   # mat3(inverse(T)), untransposed (shared_cpu_gpu.h:92-97)
        for c in range(4):
            for r in range(4):
                inst.transform.m[c][r] = float(T[r, c])
        for c in range(3):
            for r in range(3):
                inst.normal_transform.m[c][r] = float(Ninv[r, c])
        s = rng.uniform(-2, 2, 3)
        sphere[:] = [float(s[0]), float(s[1]), float(s[2]), float(rng.uniform(0.0, 1.5))]
        axis = rng.normal(size=3)
        axis /= np.linalg.norm(axis)
        meshlet.cone_apex[:] = [float(x) for x in rng.uniform(-2, 2, 3)]
        meshlet.cone_axis[:] = [float(x) for x in axis]
        meshlet.cone_cutoff = float(rng.uniform(-1, 1))
        pi, pm = C.addressof(inst), C.addressof(meshlet)
        cascade = int(rng.integers(0, 4))
        got = (orc.lho_test_cull_sphere(C.byref(u), pi, sphere), orc.lho_test_cull_sphere_shadows(C.byref(u), pi, sphere, cascade),
               orc.lho_test_cull_cone_perspective(C.byref(u), pi, pm), orc.lho_test_cull_cone_orthographic(C.byref(u), pi, pm))
        want = (ref.ref_sh_cull_bounding_sphere(C.byref(u), pi, sphere), ref.ref_sh_cull_bounding_sphere_shadows(C.byref(u), pi, sphere, cascade),
                ref.ref_sh_cull_cone_perspective(C.byref(u), pi, pm), ref.ref_sh_cull_cone_orthographic(C.byref(u), pi, pm))
        for k in range(4):
            flips[k] += got[k] != want[k]
            kept[k] += want[k] == 0
    assert flips.sum() == 0, f"decisions differing from the reference's GLSL: {flips.tolist()} of {N}"
    assert np.all(kept > N // 100) and np.all(kept < N - N // 100), f"both outcomes must be exercised: kept {kept.tolist()}"


# ---- stage 2: vertex shaders, depth reduction, lighting resolve ----

def _stage2_lib():
    from oracle_api import LhoConfig
    lib = ref_lib()
    K = C.POINTER(LhoConfig)
    fp = C.POINTER(C.c_float)
    lib.ref_sh_vertex.argtypes = [U, C.c_int, C.c_uint32, C.c_uint32, fp]
    lib.ref_sh_vertex.restype = C.c_uint32
    lib.ref_sh_vertex_alpha_clip.argtypes = [U, C.c_int, C.c_uint32, C.c_uint32, fp, fp]
    lib.ref_sh_vertex_alpha_clip.restype = C.c_uint32
    lib.ref_sh_read_depth.argtypes = [U, K, C.c_void_p]
    lib.ref_sh_read_depth.restype = None
    lib.ref_sh_render_geometry.argtypes = [U, K, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    lib.ref_sh_render_geometry.restype = None
    return lib


@pytest.mark.parametrize("preset", ["tiny", "rotated", "c4small"])
def test_vertex_shaders_match_the_oracle_vertex_transform(preset):
    """gl_Position of the reference's vertex shaders (rasterization.glsl:12-50,101-135) for every vertex of a sample of the
    surviving meshlets, main view and one cascade, against the clip positions the oracle's rasteriser starts from."""
    from oracle_api import oracle_lib
    lib, orc = _stage2_lib(), oracle_lib()
    orc.lho_test_clip_positions.argtypes = [U, C.c_int, C.c_uint32, C.POINTER(C.c_float)]
    orc.lho_test_clip_positions.restype = C.c_uint32
    sc, kw = _scene(preset)
    o = make_oracle(sc, caps=reference_caps(), **kw)
    o.reset(); o.cull_instances(); o.cull_meshlets()
    o.raster_visbuffer(); o.reduce_depth(); o.generate_shadow_matrices(); o.cull_instances_shadows()
    n_main = o.draw_counts()
    checked = 0
    pos = (C.c_float * 4)()
    uv = (C.c_float * 2)()
    buf = (C.c_float * (4 * 3 * 255))()
    rng = np.random.default_rng(7)

    def check(cascade, first_slot, n_op, n_al):
        nonlocal checked
        slots = [first_slot + int(s) for s in rng.choice(n_op, min(n_op, 60), replace=False)] if n_op else []
        slots += [first_slot + o.cfg.max_opaque_draws + int(s) for s in rng.choice(n_al, min(n_al, 30), replace=False)] if n_al else []
        for slot in slots:
            nverts = orc.lho_test_clip_positions(C.byref(o.u), cascade, slot, buf)   # 3 * triangle_count positions, draw order
            want = np.frombuffer(buf, np.float32, 4 * nverts).reshape(-1, 4)
            alpha = slot - first_slot >= o.cfg.max_opaque_draws
            for v in range(nverts):
                if alpha:
                    packed = lib.ref_sh_vertex_alpha_clip(C.byref(o.u), cascade, slot, v, pos, uv)
                else:
                    packed = lib.ref_sh_vertex(C.byref(o.u), cascade, slot, v, pos)
                assert np.array_equal(np.frombuffer(pos, np.uint32), want[v].view(np.uint32)), (cascade, slot, v)
                if cascade < 0:
                    assert packed == ((v // 3) << 24 | slot)          # pack(), util.glsl:18-20
                checked += 1

    check(-1, 0, *n_main)
    section = o.cfg.max_opaque_draws + o.cfg.max_alpha_clip_draws
    o.shadow_cascade(2)
    check(2, section, *o.draw_counts())
    assert checked > 3000


@pytest.mark.parametrize("preset", ["tiny", "rotated"])
def test_depth_reduction_matches_the_reference_shader(preset):
    lib = _stage2_lib()
    sc, kw = _scene(preset)
    o = make_oracle(sc, caps=reference_caps(), **kw)
    o.reset(); o.cull_instances(); o.cull_meshlets(); o.raster_visbuffer()
    o.reduce_depth()
    want = (o.misc().min_depth, o.misc().max_depth)
    o.reset()                                                      # min = 0xFFFFFFFF, max = 0 again
    lib.ref_sh_read_depth(C.byref(o.u), C.byref(o.cfg), o.vis.ctypes.data)   # compute/read_depth.glsl:22-83
    assert (o.misc().min_depth, o.misc().max_depth) == want and want[1] != 0


@pytest.mark.parametrize("preset", ["tiny", "rotated", "c4small"])
def test_lighting_resolve_matches_the_reference_shader(preset):
    """render_geometry.comp main() -- the reference's own text, incl. vbuffer.glsl and pbr.glsl -- against the oracle's
    restatement on every pixel.  Both sample through the same texture functions, so what is compared is everything the
    shader computes around them: attribute refetch, barycentrics and derivatives, normals, cascade selection, PCF,
    BRDF.  The restatement keeps the reference's operation order, so the lit colour is BIT-IDENTICAL."""
    from helpers import Textures
    lib = _stage2_lib()
    sc, kw = _scene(preset)
    tex = Textures(sc, 32) if sc.num_texture_sets else None
    sky = sc.make_skybox(8)
    o = make_oracle(sc, textures=tex, skybox=sky, caps=reference_caps(), **kw)
    o.render_frame(resolve=True)
    ref = np.zeros_like(o.color)
    lib.ref_sh_render_geometry(C.byref(o.u), C.byref(o.cfg), o.vis.ctypes.data, o.shadow.ctypes.data, ref.ctypes.data, 0, o.height)
    a, b = o.color.reshape(-1, 4), ref.reshape(-1, 4)
    assert np.isfinite(b).all() and (b[:, 3] == 1).all()
    geometry = (o.vis != S.LH_VIS_CLEAR)
    assert geometry.any() and (~geometry).any()
    differing = np.count_nonzero((a.view(np.uint32) != b.view(np.uint32)).any(axis=1))
    assert differing == 0, f"{differing} of {len(a)} pixels differ from the reference shader's output"


def test_display_transform_matches_the_reference_shader_with_the_real_lut():
    """display_transform.comp (tony_mc_mapface through the reference's own 48^3 table + sRGB transfer) compiled in place,
    against the oracle's lho_display_transform: bit-identical on a rendered frame and on a synthetic HDR ramp."""
    from helpers import Textures, real_display_lut
    from oracle_api import LhoConfig
    lut = real_display_lut()
    if lut is None:
        pytest.skip("oracle/_ref/tony_mc_mapface.dds not built (make ref)")
    lib = ref_lib()
    lib.ref_sh_display_transform.argtypes = [U, C.POINTER(LhoConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    lib.ref_sh_display_transform.restype = None
    sc = Scene.preset("tiny")
    o = make_oracle(sc, textures=Textures(sc, 32), skybox=sc.make_skybox(8), caps=reference_caps())
    o.render_frame(resolve=True)
    rng = np.random.default_rng(5)
    ramp = np.ones((o.width * o.height, 4), np.float32)
    ramp[:, :3] = np.exp(rng.uniform(np.log(1e-4), np.log(64.0), (len(ramp), 3))).astype(np.float32)   # 20 stops of scene-referred values
    for scene in (o.color, ramp.reshape(-1)):
        want = np.zeros(o.width * o.height * 4, np.float32)
        lib.ref_sh_display_transform(C.byref(o.u), C.byref(o.cfg), scene.ctypes.data, o.shadow.ctypes.data, lut.ctypes.data, 48, want.ctypes.data)
        got = o.display_transform(scene, lut)
        assert np.array_equal(got.view(np.uint32), want.reshape(-1, 4).view(np.uint32))
        assert 0.0 < got[:, :3].min() and got[:, :3].max() <= 1.0 + 1e-6
