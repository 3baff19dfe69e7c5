"""Parity at the BENCHMARK configurations themselves (BASELINE.json configs 2, 4, 5 at full size, with bench.py's
capacities and flags): everything that makes the bench fast -- depth bands, hierarchical Z, the deferred-triangle
queue, concurrent cascades at full occupancy -- is validated here on the frame the bench times, not on toy sizes.

Integer / index / depth-bit results are bit-exact against the oracle; the lit colour is held to SURVEY 8c's tolerance
(helpers.compare_color).  The oracle renders a 4K frame of C2 in a few seconds on the GPU box's host cores."""
import numpy as np
import pytest

from lighthugger_b200 import _structs as S
from lighthugger_b200 import sortlast
from lighthugger_b200.scene import Scene
from helpers import Textures, classify_vis_mismatches, compare_color, make_context, make_oracle, rotated_scene

pytestmark = pytest.mark.gpu


def frame_pair(sc, tex_size=64):
    tex = Textures(sc, tex_size) if sc.num_texture_sets else None
    sky = sc.make_skybox(16)
    caps = sortlast.capacities(sc, 1)     # what bench.py uses on one GPU
    o = make_oracle(sc, textures=tex, skybox=sky, caps=caps)
    ctx = make_context(sc, textures=tex, skybox=sky, caps=caps, flags=S.LH_CONFIG_WRITE_DRAW_COMMANDS | S.LH_CONFIG_PASS_TIMINGS)
    return o, ctx


def assert_frame_bit_exact(o, ctx, st, alpha_clip=False):
    mo = ctx.cfg.max_opaque_draws
    # main-view lists: prefix-sum sector 0, surviving meshlets, draw commands
    co, to, vo = o.sector(0)
    cg, tg, vg = ctx.sector(0)
    assert (cg, tg) == (co, to) and np.array_equal(vg, vo)
    # (section 0 of the reference buffer survives the cascades, which use section 1; the counters do not, so the
    # oracle's main-view counts come from the visible instances' meshlets that are listed)
    n_opaque, n_alpha = st.num_opaque, st.num_alpha_clip
    assert n_opaque > 0
    assert np.array_equal(ctx.meshlet_references(0, n_opaque), o.meshlet_references(0, n_opaque))
    assert np.array_equal(ctx.meshlet_references(mo, n_alpha), o.meshlet_references(mo, n_alpha))
    assert (n_opaque, n_alpha) == o.main_draw_counts
    # depth bounds + cascade matrices, visibility keys, shadow layers
    assert bytes(ctx.misc()) == bytes(o.misc())
    got = ctx.target(S.LH_TARGET_VIS64)
    covered = np.count_nonzero(o.vis != S.LH_VIS_CLEAR)
    mismatch = np.count_nonzero(got != o.vis)
    sh_g, sh_o = ctx.target(S.LH_TARGET_SHADOW).view(np.uint32), o.shadow.view(np.uint32)
    import helpers
    if helpers.COLOR_LOG:
        with open(helpers.COLOR_LOG, "a") as f:
            f.write('{"test": "fullsize_vis", "alpha_clip": %d, "covered": %d, "mismatch": %d, "shadow_mismatch": %d}\n' % (
                int(alpha_clip), covered, mismatch, np.count_nonzero(sh_g != sh_o)))
    if alpha_clip:   # the alpha test samples a texture inside the raster: 0.1 % of covered pixels
        assert mismatch <= 0.001 * covered, f"{mismatch} / {covered} pixels differ"
        assert classify_vis_mismatches(got, o.vis, o.width, o.height) == 0   # all of them edge or tie pixels
        assert np.count_nonzero(sh_g != sh_o) <= 0.001 * np.count_nonzero(sh_o != 0x3F800000)
    else:
        assert mismatch == 0, f"{mismatch} of {covered} covered pixels differ"
        assert np.array_equal(sh_g, sh_o), f"{np.count_nonzero(sh_g != sh_o)} shadow texels differ"
    for c in range(4):
        assert ctx.sector(c)[:2] == o.sector(c)[:2] and np.array_equal(ctx.sector(c)[2], o.sector(c)[2])
    assert st.overflow == 0
    return covered


def test_c2_at_3840x2160_matches_the_oracle():
    """BASELINE config 2 as benched: 10 000 x 100 k-triangle torus, 3840x2160."""
    sc = Scene.preset("c2")
    o, ctx = frame_pair(sc)
    o.render_frame(resolve=True)
    st = ctx.render_frame_host(ctx.uniforms())
    covered = assert_frame_bit_exact(o, ctx, st)
    assert covered > 0.5 * sc.width * sc.height and st.num_opaque > 1_000_000
    compare_color(ctx, o)
    # a second frame from the same context (warm scratch, self-cleaned counters) is the same frame
    ctx.render_frame_host(ctx.uniforms())
    assert np.array_equal(ctx.target(S.LH_TARGET_VIS64), o.vis)
    ctx.close()


def test_c4_mixed_meshes_full_size_matches_the_oracle():
    """BASELINE config 4: ~100 M triangles of mixed meshes, 5 % alpha-clip foliage, BC7 textures, 3840x2160."""
    sc = Scene.preset("c4")
    o, ctx = frame_pair(sc)
    o.render_frame(resolve=True)
    st = ctx.render_frame_host(ctx.uniforms())
    assert st.num_alpha_clip > 0
    assert_frame_bit_exact(o, ctx, st, alpha_clip=True)
    compare_color(ctx, o, ignore=np.flatnonzero(ctx.target(S.LH_TARGET_VIS64) != o.vis))
    ctx.close()


def test_c5_at_7680x4320_matches_the_oracle():
    """BASELINE config 5: the 1 G-triangle scene at 7680x4320 (the key buffer no longer fits L2)."""
    sc = Scene.preset("c5")
    o, ctx = frame_pair(sc)
    o.render_frame(resolve=False)
    st = ctx.render_frame_host(ctx.uniforms())
    assert_frame_bit_exact(o, ctx, st)
    ctx.close()


def test_general_rotations_and_non_uniform_scales():
    """Instances with uniformly random rotations and per-axis scales: the reference's untransposed normal transform
    (shared_cpu_gpu.h:92-97, culling.glsl:141) is then not the inverse transpose, and max-column-length scaling of the
    bounding spheres matters."""
    sc = rotated_scene()
    o, ctx = frame_pair(sc)
    o.render_frame(resolve=True)
    st = ctx.render_frame_host(ctx.uniforms())
    assert_frame_bit_exact(o, ctx, st)
    compare_color(ctx, o)
    ctx.close()
