"""The C++20 host (examples/frame_loop.cpp -> lighthugger_b200/bin/lh_frame_loop): the reference's main() frame loop
re-hosted on the two C ABIs.  Without a GPU it must fail loudly (no CPU fallback); on a GPU its presented image must be
the one the Python binding produces for the same camera."""
import copy
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "lighthugger_b200", "bin", "lh_frame_loop")


def _built():
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", ROOT, "examples"])
    return EXE


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_cpp_host_fails_loudly_without_a_gpu():
    if _have_gpu():
        pytest.skip("a GPU is present")
    r = subprocess.run([_built(), "tiny", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cpp_host_presents_the_same_image_as_the_python_binding():
    from lighthugger_b200 import _structs as S
    from lighthugger_b200.scene import Scene
    from helpers import Textures, make_context
    frames = 3
    r = subprocess.run([_built(), "tiny", str(frames)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    last = r.stdout.strip().splitlines()[-1]
    assert f"{frames} frames of tiny" in last
    theirs = int(last.split("checksum")[1].strip(), 16)
    sc = Scene.preset("tiny")
    ctx = make_context(sc, textures=Textures(sc, 64), skybox=sc.make_skybox(16))
    cam = copy.copy(sc.camera)
    yaw = np.float32(cam.yaw)
    for _ in range(frames):
        yaw = np.float32(yaw + np.float32(0.002))
    cam.yaw = float(yaw)
    ctx.frame_begin(ctx.uniforms(camera=cam))
    ctx.render_frame(); ctx.display_transform(); ctx.sync()
    image = ctx.target(S.LH_TARGET_DISPLAY_BGRA8).view(np.uint32).reshape(-1)
    h = 1469598103934665603
    for p in image.tolist():                       # FNV-1a, as examples/frame_loop.cpp
        h = ((h ^ p) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    assert h == theirs
    ctx.close()
