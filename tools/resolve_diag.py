"""Diagnostic (GPU box): where does the lit colour differ from the oracle?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from lighthugger_b200 import _structs as S
from lighthugger_b200.scene import Scene
from helpers import Textures, make_context, make_oracle, vis_fields
from test_resolve_gpu import display

for name, kw, bc7, ts in [("tiny", {}, True, 64), ("tiny", {}, False, 64), ("tiny", {}, None, 64), ("c1", dict(width=480, height=270), True, 64)]:
    sc = Scene.preset(name, **kw)
    tex = Textures(sc, ts, bc7=bc7) if bc7 is not None else None
    sky = sc.make_skybox(16)
    o = make_oracle(sc, textures=tex, skybox=sky)
    ctx = make_context(sc, textures=tex, skybox=sky)
    o.render_frame(resolve=True)
    ctx.render_frame_host(ctx.uniforms())
    got = ctx.target(S.LH_TARGET_SCENE_RGBA16F).astype(np.float32).reshape(-1, 4)
    want = o.color.reshape(-1, 4)
    d = np.abs(display(got[:, :3]) - display(want[:, :3])).max(axis=1)
    depth, slot, tri = vis_fields(o.vis.reshape(-1))
    cov = depth != 0
    print(name, "bc7", bc7, "pixels", len(d), "covered", cov.sum(), "bad>2/255", (d > 2/255).sum(), "bad sky", ((d > 2/255) & ~cov).sum(),
          "max", d.max(), "q99", np.quantile(d, .99), "q999", np.quantile(d, .999))
    bad = np.nonzero(d > 2/255)[0][:8]
    for i in bad:
        print("   px", i % sc.width, i // sc.width, "got", got[i, :3], "want", want[i, :3], "slot", slot[i], "tri", tri[i])
    ctx.close()
