"""Diagnostic (GPU box): where does the lit colour differ from the oracle's, and through which input of the pass?
Renders a scene on both sides, then re-runs the single-launch resolve with uniforms.debug = 100..105 (shadow term, base
colour, normal, roughness/metallic/cascade, uv) against lho_resolve_debug, and prints, for the pixels whose colour is
off by more than 2/255, how each term differs.  usage: python tools/resolve_diag.py [scene ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np
from lighthugger_b200 import _structs as S
from lighthugger_b200.scene import Scene
from helpers import Textures, display, make_context, make_oracle, rotated_scene, vis_fields

TERMS = {100: "shadow", 101: "base colour", 102: "normal", 103: "rough/metal/cascade", 105: "uv", 106: "view vector", 107: "sun term", 108: "footprint (probes, lod)"}


def scene_of(name):
    if name == "rotated":
        return rotated_scene(), 64
    if name == "c1q":
        return Scene.preset("c1", width=480, height=270), 64
    if name == "c4s":
        return Scene.preset("c4", instance_count=120, width=480, height=270), 64
    if name == "c2":
        return Scene.preset("c2"), 64
    return Scene.preset(name), 64


for name in (sys.argv[1:] or ["tiny", "c1q", "rotated"]):
    sc, ts = scene_of(name)
    tex = Textures(sc, ts) if sc.num_texture_sets else None
    sky = sc.make_skybox(16)
    o = make_oracle(sc, textures=tex, skybox=sky)
    ctx = make_context(sc, textures=tex, skybox=sky)
    o.render_frame(resolve=True)
    ctx.render_frame_host(ctx.uniforms())
    assert np.array_equal(ctx.target(S.LH_TARGET_VIS64), o.vis)
    got = ctx.target(S.LH_TARGET_SCENE_RGBA16F).astype(np.float32).reshape(-1, 4)
    want = o.color.reshape(-1, 4)
    d = np.abs(display(got[:, :3]) - display(want[:, :3])).max(axis=1)
    depth, slot, tri = vis_fields(o.vis.reshape(-1))
    cov = depth != 0
    bad = d > 2 / 255
    print(f"== {name}: {len(d)} px, covered {cov.sum()}, >2/255: {bad.sum()} ({bad.mean():.4%}), of which sky {np.count_nonzero(bad & ~cov)}, "
          f"max {d.max() * 255:.2f}/255, q99.9 {np.quantile(d, .999) * 255:.2f}/255")
    diffs = {}
    for term, label in TERMS.items():
        u = ctx.uniforms()
        u.debug = term
        ctx.frame_begin(u)
        ctx.resolve()
        ctx.sync()
        g = ctx.target(S.LH_TARGET_SCENE_RGBA16F).astype(np.float32).reshape(-1, 4)[:, :3]
        w = o.resolve_debug(term)[:, :3]
        e = np.abs(g - w).max(axis=1)
        e[~cov] = 0
        diffs[term] = (e, g, w)
        q = np.quantile(e[cov], [0.5, 0.9, 0.99, 0.999]) if cov.any() else [0] * 4
        print(f"   {label:22s} q50/q90/q99/q99.9 {q[0]:.5f} {q[1]:.5f} {q[2]:.5f} {q[3]:.5f}  max |diff| all {e.max():.5f}  on bad pixels: median {np.median(e[bad]) if bad.any() else 0:.5f} max {e[bad].max() if bad.any() else 0:.5f};"
              f"  pixels with diff > 0.01: {np.count_nonzero(e > 0.01)} (bad among them {np.count_nonzero((e > 0.01) & bad)})")
    order = np.argsort(-d)[:6]
    for i in order:
        terms = "  ".join(f"{TERMS[t][:6]} {diffs[t][1][i].round(4).tolist()} vs {diffs[t][2][i].round(4).tolist()}" for t in (100, 101, 102, 103, 106, 107))
        print(f"   px ({i % sc.width},{i // sc.width}) err {d[i] * 255:.2f}/255 got {got[i, :3].round(4).tolist()} want {want[i, :3].round(4).tolist()} | {terms}")
    ctx.close()
