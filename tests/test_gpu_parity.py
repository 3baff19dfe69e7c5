"""CUDA path vs the oracle, pass by pass, through the C ABI (run on the B200
box: pytest -m gpu).  Integer / index / depth-bit results must be bit-exact;
the floating-point resolve is held to the tolerance stated in test_resolve."""
import numpy as np
import pytest

from lighthugger_b200 import _structs as S
from lighthugger_b200.scene import Scene
from helpers import Textures, golden_torus_scene, make_context, make_oracle, rotated_scene, vis_fields

pytestmark = pytest.mark.gpu


def build(name, tex_size=32, **kw):
    sc = Scene.preset(name, **kw) if isinstance(name, str) else name
    tex = Textures(sc, tex_size) if sc.num_texture_sets else None
    sky = sc.make_skybox(8)
    return sc, tex, sky


def run_cull(o, ctx, u):
    o.reset(); o.cull_instances(); o.cull_meshlets()
    ctx.frame_begin(u); ctx.reset(); ctx.cull_instances(); ctx.cull_meshlets(); ctx.sync()


def assert_lists_equal(o, ctx, section=0):
    n_opaque, n_alpha = o.draw_counts()
    assert ctx.draw_counts() == (n_opaque, n_alpha)
    mo = ctx.cfg.max_opaque_draws
    assert np.array_equal(ctx.meshlet_references(section, n_opaque), o.meshlet_references(section, n_opaque))
    assert np.array_equal(ctx.meshlet_references(section + mo, n_alpha), o.meshlet_references(section + mo, n_alpha))
    assert np.array_equal(ctx.draw_commands(0, n_opaque), o.draw_commands(0, n_opaque))
    assert np.array_equal(ctx.draw_commands(mo, n_alpha), o.draw_commands(mo, n_alpha))
    return n_opaque, n_alpha


SCENES = ["tiny", "golden", "c1"]


@pytest.fixture(scope="module", params=SCENES)
def pair(request):
    if request.param == "golden":
        sc = golden_torus_scene()
    elif request.param == "c1":
        sc = Scene.preset("c1", width=960, height=540)
    else:
        sc = Scene.preset(request.param)
    sc, tex, sky = build(sc)
    o = make_oracle(sc, textures=tex, skybox=sky)
    ctx = make_context(sc, textures=tex, skybox=sky)
    yield sc, o, ctx
    ctx.close()


def test_reset_matches(pair):
    sc, o, ctx = pair
    u = ctx.uniforms()
    o.reset()
    ctx.frame_begin(u); ctx.reset(); ctx.sync()
    assert np.array_equal(ctx.dispatches()[:2], o.dispatches()[:2])
    m = ctx.misc()
    assert (m.min_depth, m.max_depth) == (0xFFFFFFFF, 0)
    assert ctx.draw_counts() == (0, 0) and ctx.sector(0)[:2] == (0, 0)


def test_instance_cull_and_prefix_sum_bit_exact(pair):
    sc, o, ctx = pair
    o.reset(); o.cull_instances()
    ctx.frame_begin(ctx.uniforms()); ctx.reset(); ctx.cull_instances(); ctx.sync()
    co, to, vo = o.sector(0)
    cg, tg, vg = ctx.sector(0)
    assert (cg, tg) == (co, to) and co > 0
    assert np.array_equal(vg, vo)


def test_meshlet_cull_lists_bit_exact(pair):
    sc, o, ctx = pair
    run_cull(o, ctx, ctx.uniforms())
    n_opaque, n_alpha = assert_lists_equal(o, ctx)
    assert n_opaque > 0
    assert np.array_equal(ctx.dispatches(), o.dispatches())


def test_visbuffer_bit_exact(pair):
    sc, o, ctx = pair
    run_cull(o, ctx, ctx.uniforms())
    o.raster_visbuffer()
    ctx.raster_visbuffer(); ctx.sync()
    got = ctx.target(S.LH_TARGET_VIS64)
    covered = (o.vis != S.LH_VIS_CLEAR)
    assert covered.any()
    mismatch = np.count_nonzero(got != o.vis)
    # integer coverage, pinned depth expression and min-resolve: no tie-break slack is needed for opaque geometry
    assert mismatch == 0, f"{mismatch} of {covered.sum()} covered pixels differ"
    depth = ctx.target(S.LH_TARGET_DEPTH)
    assert np.array_equal(depth.view(np.uint32), vis_fields(o.vis)[0])


def test_depth_reduction_bit_exact(pair):
    sc, o, ctx = pair
    run_cull(o, ctx, ctx.uniforms())
    o.raster_visbuffer(); o.reduce_depth()
    ctx.raster_visbuffer(); ctx.reduce_depth(); ctx.sync()
    mo, mg = o.misc(), ctx.misc()
    assert (mg.min_depth, mg.max_depth) == (mo.min_depth, mo.max_depth)
    assert mo.max_depth != 0


def test_whole_frame_matches_pass_by_pass_and_oracle(pair):
    sc, o, ctx = pair
    o.render_frame(resolve=False)
    st = ctx.render_frame_host(ctx.uniforms())
    # cascades: matrices, radii and depth bounds are bit-exact (pinned pow / sqrt / div)
    assert bytes(ctx.misc()) == bytes(o.misc())
    assert np.array_equal(ctx.target(S.LH_TARGET_VIS64), o.vis)
    sh_g = ctx.target(S.LH_TARGET_SHADOW).view(np.uint32)
    sh_o = o.shadow.view(np.uint32)
    assert np.array_equal(sh_g, sh_o), f"{np.count_nonzero(sh_g != sh_o)} shadow texels differ"
    # last cascade's lists are what is left in the shared buffers
    section = ctx.cfg.max_opaque_draws + ctx.cfg.max_alpha_clip_draws
    assert_lists_equal(o, ctx, section)
    mo = o.misc()
    assert (st.min_depth, st.max_depth) == (mo.min_depth, mo.max_depth)
    assert st.overflow == 0
    for c in range(4):
        assert ctx.sector(c)[:2] == o.sector(c)[:2]
        assert np.array_equal(ctx.sector(c)[2], o.sector(c)[2])


def test_shadow_cascades_one_by_one(pair):
    sc, o, ctx = pair
    u = ctx.uniforms()
    run_cull(o, ctx, u)
    o.raster_visbuffer(); o.reduce_depth(); o.generate_shadow_matrices(); o.cull_instances_shadows()
    ctx.raster_visbuffer(); ctx.reduce_depth(); ctx.generate_shadow_matrices(); ctx.cull_instances_shadows(); ctx.sync()
    section = ctx.cfg.max_opaque_draws + ctx.cfg.max_alpha_clip_draws
    for c in range(4):
        assert ctx.sector(c)[:2] == o.sector(c)[:2]
        o.shadow_cascade(c)
        ctx.shadow_cascade(c); ctx.sync()
        assert_lists_equal(o, ctx, section)
    assert np.array_equal(ctx.target(S.LH_TARGET_SHADOW).view(np.uint32), o.shadow.view(np.uint32))


def test_frame_stats(pair):
    sc, o, ctx = pair
    o.reset(); o.cull_instances(); o.cull_meshlets()
    co, to, _ = o.sector(0)
    n_opaque, n_alpha = o.draw_counts()
    st = ctx.render_frame_host(ctx.uniforms())
    assert (st.num_visible_instances, st.num_candidate_meshlets, st.num_opaque, st.num_alpha_clip) == (co, to, n_opaque, n_alpha)


# ---- edge cases ----

def test_empty_scene_and_empty_view():
    sc = Scene.preset("tiny")
    sc.camera.yaw = -np.pi / 2  # nothing in view
    o = make_oracle(sc)
    ctx = make_context(sc)
    o.render_frame(resolve=False)
    st = ctx.render_frame_host(ctx.uniforms())
    assert st.num_visible_instances == 0 and st.num_opaque == 0
    assert (st.min_depth, st.max_depth) == (0xFFFFFFFF, 0)
    assert (ctx.target(S.LH_TARGET_VIS64) == S.LH_VIS_CLEAR).all()
    ctx.close()


def test_ragged_resolution_and_near_plane_clipping():
    # odd target size (not a multiple of any tile) and a camera inside the instance cloud (triangles cross w = near)
    sc = Scene.preset("tiny", width=251, height=141)
    sc.camera.position[:] = (0.4, 0.1, 0.3)
    o = make_oracle(sc)
    ctx = make_context(sc)
    o.render_frame(resolve=False)
    st = ctx.render_frame_host(ctx.uniforms())
    assert st.num_large_triangles > 0  # the deferred path ran
    got = ctx.target(S.LH_TARGET_VIS64)
    assert np.count_nonzero(got != o.vis) == 0
    assert np.array_equal(ctx.target(S.LH_TARGET_SHADOW).view(np.uint32), o.shadow.view(np.uint32))
    ctx.close()


def test_capacity_overflow_is_reported_not_silent():
    from lighthugger_b200.renderer import LhError
    sc = Scene.preset("tiny")
    caps = S.LhCapacities(max_instances=8, max_opaque_draws=64, max_alpha_clip_draws=16)
    ctx = make_context(sc, caps=caps)
    with pytest.raises(LhError) as e:
        ctx.render_frame_host(ctx.uniforms())
    assert e.value.code == S.LH_ERR_CAPACITY
    ctx.close()
    caps = S.LhCapacities(max_instances=64, max_opaque_draws=50, max_alpha_clip_draws=16)
    ctx = make_context(sc, caps=caps)
    with pytest.raises(LhError) as e:
        ctx.render_frame_host(ctx.uniforms())
    assert e.value.code == S.LH_ERR_CAPACITY
    ctx.close()


def test_reference_capacities_default_config():
    # the reference's compile-time capacities (1024 / 200 000 / 200 000) with a scene that fits them
    sc = Scene.preset("tiny")
    caps = S.LhCapacities(1024, 200000, 200000)
    o = make_oracle(sc, caps=caps)
    ctx = make_context(sc, caps=caps)
    o.render_frame(resolve=False)
    ctx.render_frame_host(ctx.uniforms())
    assert np.array_equal(ctx.target(S.LH_TARGET_VIS64), o.vis)
    ctx.close()


def test_sort_last_partition_on_one_gpu():
    # two contexts with disjoint instance ranges: min-composite == whole frame (SURVEY.md 8e "determinism")
    sc = Scene.preset("tiny")
    whole = make_context(sc)
    whole.render_frame_host(whole.uniforms())
    vw = whole.target(S.LH_TARGET_VIS64)
    rw = whole.meshlet_references(0, whole.stats().num_opaque)
    n = sc.num_instances
    acc = np.full_like(vw, S.LH_VIS_CLEAR)
    parts = []
    for first, count in ((0, 20), (20, n - 20)):
        c = make_context(sc)
        c.set_instance_range(first, count)
        c.render_frame_host(c.uniforms())
        v = c.target(S.LH_TARGET_VIS64)
        acc = np.minimum(acc, v)
        parts.append((v, c.meshlet_references(0, c.stats().num_opaque)))
        c.close()
    dw, sw, tw = vis_fields(vw)
    da, _, _ = vis_fields(acc)
    assert np.array_equal(dw, da)
    cov = dw != 0
    got = np.zeros((acc.size, 3), np.int64) - 1
    for v, refs in parts:
        d, s, t = vis_fields(v)
        win = (v == acc) & (d != 0)
        got[win, 0], got[win, 1], got[win, 2] = refs[s[win], 0], refs[s[win], 1], t[win]
    want = np.stack([rw[sw[cov], 0], rw[sw[cov], 1], tw[cov]], 1)
    assert np.array_equal(got[cov], want)
    whole.close()


def test_sort_last_stripes_on_one_gpu():
    # screen-stripe ownership (lh_set_instance_stripes): three contexts own interleaved stripes of the main view and of
    # the cascades; the min of their key buffers and of their shadow layers is the whole frame's
    sc = Scene.preset("tiny")
    whole = make_context(sc)
    whole.render_frame_host(whole.uniforms())
    vw = whole.target(S.LH_TARGET_VIS64)
    sw = whole.target(S.LH_TARGET_SHADOW).view(np.uint32)
    rw = whole.meshlet_references(0, whole.stats().num_opaque)
    dw, slot_w, tri_w = vis_fields(vw)
    parts, acc, visible = [], np.full_like(vw, S.LH_VIS_CLEAR), 0
    members = 3
    for index in range(members):
        c = make_context(sc)
        c.set_instance_stripes(0, 4 * members, index, members)
        c.set_instance_stripes(1, 4 * members, index, members)
        c.render_frame_host(c.uniforms())
        # the cascade matrices come from this context's own depth bounds, so only the main view is compared here; the
        # cascades' partition is checked below with the whole frame's depth bounds
        v = c.target(S.LH_TARGET_VIS64)
        acc = np.minimum(acc, v)
        st = c.stats()
        visible += st.num_visible_instances
        parts.append((v, c.meshlet_references(0, st.num_opaque)))
        c.close()
    assert visible == whole.stats().num_visible_instances, "the stripes do not partition the visible instances"
    da, _, _ = vis_fields(acc)
    assert np.array_equal(dw, da)
    cov = dw != 0
    got = np.zeros((acc.size, 3), np.int64) - 1
    for v, refs in parts:
        d, s, t = vis_fields(v)
        win = (v == acc) & (d != 0)
        got[win, 0], got[win, 1], got[win, 2] = refs[s[win], 0], refs[s[win], 1], t[win]
    want = np.stack([rw[slot_w[cov], 0], rw[slot_w[cov], 1], tri_w[cov]], 1)
    assert np.array_equal(got[cov], want)
    # cascades: pass by pass with the whole frame's visibility buffer, so that every context derives the same matrices
    layers = np.full_like(sw, 0x3F800000)
    shadow_visible = np.zeros(4, np.int64)
    for index in range(members):
        c = make_context(sc)
        c.set_instance_stripes(1, 4 * members, index, members)   # main view whole, cascades striped
        c.frame_begin(c.uniforms())
        c.reset(); c.cull_instances(); c.cull_meshlets(); c.raster_visbuffer(); c.reduce_depth()
        c.generate_shadow_matrices(); c.cull_instances_shadows()
        for k in range(4):
            c.shadow_cascade(k)
        c.sync()
        st = c.stats()
        shadow_visible += np.array(list(st.shadow_visible_instances), np.int64)
        layers = np.minimum(layers, c.target(S.LH_TARGET_SHADOW).view(np.uint32))
        c.close()
    assert np.array_equal(shadow_visible, np.array(list(whole.stats().shadow_visible_instances), np.int64))
    assert np.array_equal(layers, sw)
    whole.close()
