"""Lighting resolve and alpha-clip raster against the oracle.  Both used to be tolerance-checked; with the
oracle modelling the texture unit's filter bit for bit (tests/test_texture_unit.py) they are exact:
  lit colour: covered pixels bit-identical in RGBA16F; every pixel within 0.125/255 after the display transfer on
              >= 99.9 % of the pixels and 2/255 on the rest (helpers.compare_color states the bound and the exceptions);
  vis ids and shadow layers with alpha-clip geometry: bit-exact (the alpha test samples the texture unit too)."""
import numpy as np
import pytest

from lighthugger_b200 import _structs as S
from lighthugger_b200.scene import Scene
from helpers import Textures, classify_vis_mismatches, compare_color, display, make_context, make_oracle, vis_fields

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,bc7", [("tiny", True), ("tiny", False)])
def test_resolve_matches_oracle(name, bc7):
    sc = Scene.preset(name)
    tex = Textures(sc, 64, bc7=bc7)
    sky = sc.make_skybox(16)
    o = make_oracle(sc, textures=tex, skybox=sky)
    ctx = make_context(sc, textures=tex, skybox=sky)
    o.render_frame(resolve=True)
    ctx.render_frame_host(ctx.uniforms())
    assert np.array_equal(ctx.target(S.LH_TARGET_VIS64), o.vis)  # same visibility going in
    compare_color(ctx, o)
    ctx.close()


def test_resolve_c1_quarter_resolution():
    sc = Scene.preset("c1", width=480, height=270)
    tex = Textures(sc, 64)
    sky = sc.make_skybox(16)
    o = make_oracle(sc, textures=tex, skybox=sky)
    ctx = make_context(sc, textures=tex, skybox=sky)
    o.render_frame(resolve=True)
    ctx.render_frame_host(ctx.uniforms())
    assert np.array_equal(ctx.target(S.LH_TARGET_VIS64), o.vis)
    compare_color(ctx, o)
    ctx.close()


def test_alpha_clip_geometry_mixed_scene():
    sc = Scene.preset("c4", instance_count=120, width=480, height=270)
    tex = Textures(sc, 64)
    sky = sc.make_skybox(16)
    o = make_oracle(sc, textures=tex, skybox=sky)
    ctx = make_context(sc, textures=tex, skybox=sky)
    o.render_frame(resolve=True)
    st = ctx.render_frame_host(ctx.uniforms())
    assert st.num_alpha_clip > 0 and o.draw_counts is not None
    got = ctx.target(S.LH_TARGET_VIS64)
    covered = np.count_nonzero(o.vis != S.LH_VIS_CLEAR)
    mismatch = np.count_nonzero(got != o.vis)
    import helpers
    if helpers.COLOR_LOG:
        with open(helpers.COLOR_LOG, "a") as f:
            f.write('{"test": "alpha_clip_vis", "covered": %d, "mismatch": %d, "shadow_mismatch": %d}\n' % (
                covered, mismatch, np.count_nonzero(ctx.target(S.LH_TARGET_SHADOW).view(np.uint32) != o.shadow.view(np.uint32))))
    # the alpha test reads the texture unit, which the oracle models bit for bit: no tolerance (if a different GPU's
    # texture unit ever filtered differently, classify_vis_mismatches tells edge / tie pixels from real errors)
    assert mismatch == 0, f"{mismatch} / {covered} pixels differ, of which {classify_vis_mismatches(got, o.vis, sc.width, sc.height)} are neither edge nor tie pixels"
    sh_g, sh_o = ctx.target(S.LH_TARGET_SHADOW).view(np.uint32), o.shadow.view(np.uint32)
    assert np.array_equal(sh_g, sh_o)
    ctx.close()


def test_display_transform_runs():
    sc = Scene.preset("tiny")
    ctx = make_context(sc, skybox=sc.make_skybox(8))
    ctx.render_frame_host(ctx.uniforms())
    ctx.display_transform(); ctx.sync()
    bgra = ctx.target(S.LH_TARGET_DISPLAY_BGRA8).reshape(-1, 4)
    assert (bgra[:, 3] == 255).all() and bgra[:, :3].max() > 0
    # without a LUT the transform is c/(1+c) followed by the sRGB OETF (display_transform.comp:20-27)
    scene = ctx.target(S.LH_TARGET_SCENE_RGBA16F).astype(np.float32).reshape(-1, 4)
    want = np.rint(np.clip(display(scene[:, :3]), 0, 1) * 255)
    assert np.abs(want[:, ::-1] - bgra[:, :3]).max() <= 1
    ctx.close()


def test_textures_from_dds_and_ktx2_files_render_like_direct_uploads():
    # the preset's BC7 mip chains wrapped into DDS (DX10) and KTX2 + Zstandard containers, parsed by
    # lh_image_parse / lh_image_decode and handed to lh_texture_create, must give the frame the direct upload gives
    import test_image_files as F
    from lighthugger_b200 import image_files as IF
    sc = Scene.preset("tiny")
    tex = Textures(sc, 32)
    direct = make_context(sc, textures=tex)
    direct.render_frame_host(direct.uniforms())
    want = direct.target(S.LH_TARGET_SCENE_RGBA16F).view(np.uint16).copy()
    direct.close()
    files = make_context(sc)
    for i, (blocks, levels, _) in enumerate(tex.entries):
        if i % 2 == 0:
            data = F.dds_file(F.DXGI["BC7_UNORM"], tex.size, tex.size, levels, blocks)   # UNORM on disk, sRGB when sampled
        else:
            sizes = F.level_bytes(tex.size, tex.size, levels, 8, True)
            cuts = np.cumsum([0] + sizes)
            chain = [blocks[cuts[l]:cuts[l + 1]].tobytes() for l in range(levels)]
            data = F.ktx2_file(F.VK["BC7_SRGB"], tex.size, tex.size, levels, chain, scheme=2, stored=[F._zstd_compress(c) for c in chain])
        assert IF.upload(files, data) == ("texture", i)
    files.render_frame_host(files.uniforms())
    got = files.target(S.LH_TARGET_SCENE_RGBA16F).view(np.uint16)
    assert np.array_equal(got, want)
    files.close()


def test_display_lut_and_skybox_from_rgb9e5_dds():
    import test_image_files as F
    from lighthugger_b200 import image_files as IF
    sc = Scene.preset("tiny")
    n = 48   # the size display_transform.comp's texel-centre mapping (encoded * 47/48 + 0.5/48) is written for
    # an identity table and a constant sky, written as R9G9B9E5 texels (exponent 15: mantissa / 512)
    g = np.arange(n, dtype=np.uint32) * 511 // (n - 1)
    b, gg, r = np.meshgrid(g, g, g, indexing="ij")
    lut = (r | (gg << 9) | (b << 18) | (15 << 27)).astype(np.uint32).reshape(-1)
    sky = np.full(6 * 4 * 4, 128 | (128 << 9) | (256 << 18) | (15 << 27), np.uint32)
    ctx = make_context(sc)
    assert IF.upload(ctx, F.dds_file(F.DXGI["RGB9E5"], n, n, 1, lut.tobytes(), depth=n, dimension=4)) == ("lut", n)
    assert IF.upload(ctx, F.dds_file(F.DXGI["RGB9E5"], 4, 4, 1, sky.tobytes(), cubemap=True)) == ("skybox", 4)
    ctx.render_frame_host(ctx.uniforms())
    ctx.display_transform(); ctx.sync()
    scene = ctx.target(S.LH_TARGET_SCENE_RGBA16F).astype(np.float32).reshape(-1, 4)
    depth, _, _ = vis_fields(ctx.target(S.LH_TARGET_VIS64))
    skyp = depth == 0
    assert skyp.any() and np.allclose(scene[skyp, :3], np.array([0.25, 0.25, 0.5]) * 8, rtol=1e-3)   # skybox * 2^3 (render_geometry.comp:86-101)
    # the table maps encoded c/(1+c) to itself (trilinear between exact grid points), so the result is the OETF of it
    bgra = ctx.target(S.LH_TARGET_DISPLAY_BGRA8).reshape(-1, 4)
    enc = np.clip(scene[:, :3], 0, None) / (1 + np.clip(scene[:, :3], 0, None))
    want = np.rint(np.clip(np.where(enc <= 0.0031308, 12.92 * enc, 1.055 * np.power(enc, 1 / 2.4) - 0.055), 0, 1) * 255)
    assert np.abs(want[:, ::-1] - bgra[:, :3]).max() <= 3
    ctx.close()


def test_gltf_scene_renders_like_the_directly_built_scene(tmp_path):
    # .gltf + .bin + .dds / .ktx2 files -> lighthugger_b200.gltf.load + image_files.upload -> the same frame
    # as the scene assembled from the same arrays and textures through the host library
    import gltf_fixture
    import test_image_files as F
    from lighthugger_b200 import gltf as G, image_files as IF
    donor = Scene.preset("tiny")
    tex = Textures(donor, 32)
    files = []
    for i in range(3):
        blocks, levels, _ = tex.entries[i]
        files.append((f"tex{i}.dds", F.dds_file(F.DXGI["BC7_SRGB"], tex.size, tex.size, levels, blocks)))
    path, expected = gltf_fixture.build(str(tmp_path), files)
    loaded, images = G.load(path)
    gltf_fixture.look_at_instances(loaded)
    frames = []
    for scene in (loaded, expected):
        ctx = make_context(scene)
        for i, image in enumerate(images):
            assert IF.upload(ctx, open(image, "rb").read()) == ("texture", i)
        st = ctx.render_frame_host(ctx.uniforms())
        assert st.num_visible_instances == 3 and st.num_opaque > 0 and st.num_alpha_clip > 0
        frames.append((ctx.target(S.LH_TARGET_VIS64).copy(), ctx.target(S.LH_TARGET_SCENE_RGBA16F).view(np.uint16).copy()))
        ctx.close()
    assert np.array_equal(frames[0][0], frames[1][0]) and np.array_equal(frames[0][1], frames[1][1])
    depth, _, _ = vis_fields(frames[0][0])
    assert np.count_nonzero(depth) > 0.02 * depth.size


def test_display_transform_with_the_reference_lut_matches_the_oracle():
    """the REAL Tony McMapface table (the reference's own DDS, read through the DDS parser and
    lh_set_display_lut) on a lit frame: B8G8R8A8 output within 1/255 of the oracle's display_transform.comp restatement on
    every byte and identical on 99.9 % of them; the restatement is itself bit-identical to the reference's shader
    (tests/test_ref_shaders.py)."""
    from helpers import REAL_LUT, real_display_lut
    from lighthugger_b200 import image_files as IF
    lut = real_display_lut()
    if lut is None:
        pytest.skip("oracle/_ref/tony_mc_mapface.dds not present (built by `make ref` where /root/reference exists)")
    sc = Scene.preset("c1", width=960, height=540)
    tex = Textures(sc, 64)
    sky = sc.make_skybox(16)
    ctx = make_context(sc, textures=tex, skybox=sky)
    assert IF.upload(ctx, open(REAL_LUT, "rb").read()) == ("lut", 48)
    ctx.render_frame_host(ctx.uniforms())
    ctx.display_transform(); ctx.sync()
    scene = ctx.target(S.LH_TARGET_SCENE_RGBA16F).astype(np.float32).reshape(-1, 4)
    bgra = ctx.target(S.LH_TARGET_DISPLAY_BGRA8).reshape(-1, 4).astype(np.int32)
    o = make_oracle(sc, textures=tex, skybox=sky)
    want = np.rint(np.clip(o.display_transform(scene, lut)[:, :3], 0, 1) * 255).astype(np.int32)   # the UNORM store
    diff = np.abs(want[:, ::-1] - bgra[:, :3])
    import helpers
    if helpers.COLOR_LOG:
        with open(helpers.COLOR_LOG, "a") as f:
            f.write('{"test": "display_transform_real_lut", "pixels": %d, "bytes_differing": %d, "max": %d}\n' % (len(diff), np.count_nonzero(diff), diff.max()))
    assert (bgra[:, 3] == 255).all()
    assert diff.max() <= 1, f"max difference {diff.max()}/255 on {np.count_nonzero(diff > 1)} pixels"
    # the oracle models the unit's 3D float filter (99.86 % of fetches bit for bit) and both sides pin pow: measured on
    # B200, 339 of 1,555,200 bytes differ, by 1
    assert np.count_nonzero(diff) <= 0.001 * diff.size
    assert len(np.unique(bgra[:, :3])) > 100            # a real image, not a constant
    ctx.close()
