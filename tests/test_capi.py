"""The C-ABI library loads and exports every symbol include/lighthugger_b200.h
declares (no compute calls: runs without a GPU)."""
import ctypes as C
import os
import re

import pytest

from lighthugger_b200 import renderer
from lighthugger_b200 import _structs as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header, prefix):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z0-9_]+)\s*\(", text)))


@pytest.mark.skipif(not os.path.exists(renderer.cuda_lib_path()), reason="CUDA library not built (make cuda)")
def test_every_declared_export_is_present():
    lib = C.CDLL(renderer.cuda_lib_path())
    names = [n for n in declared("lighthugger_b200.h", "lh_") if not n.startswith("lh_vis_")]
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/lighthugger_b200.h but not exported"
    assert sorted(set(renderer.EXPORTS)) == names


def test_host_library_exports():
    lib = C.CDLL(os.path.join(ROOT, "lighthugger_b200", "lib", "liblh_host.so"))
    for n in declared("lh_scene.h", "lh_"):
        assert hasattr(lib, n), n


@pytest.mark.skipif(not os.path.exists(renderer.cuda_lib_path()), reason="CUDA library not built (make cuda)")
def test_config_defaults_are_the_reference_capacities():
    lib = renderer.cuda_lib()
    cfg = S.LhConfig()
    assert lib.lh_config_default(C.byref(cfg)) == 0
    # shared_cpu_gpu.h:119-125,159; main.cpp:103-106,385-409
    assert (cfg.max_instances, cfg.max_opaque_draws, cfg.max_alpha_clip_draws) == (1024, 200000, 200000)
    assert (cfg.width, cfg.height, cfg.shadowmap_size) == (640, 480, 1024)
    sizes = S.LhBufferSizes()
    assert lib.lh_buffer_sizes(C.byref(cfg), C.byref(sizes)) == 0
    # main.cpp:442-532
    assert sizes.meshlet_references == 8 * 400000 * 2 and sizes.num_meshlets_prefix_sum == (8 * 1024 + 8) * 4
    assert sizes.draw_calls == 8 + 16 * 400000 and sizes.dispatches == 36 and sizes.misc_storage == 792 and sizes.uniforms == 560


@pytest.mark.skipif(not os.path.exists(renderer.cuda_lib_path()), reason="CUDA library not built (make cuda)")
def test_no_gpu_means_a_loud_failure_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(renderer.LhError) as e:
        renderer.Context(64, 64)
    assert e.value.code in (-5, -2)


def test_product_never_touches_the_oracle():
    # nothing under lighthugger_b200/ may include, link or import oracle/
    for base, _, files in os.walk(os.path.join(ROOT, "lighthugger_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(base, f)).read()
                assert "lh_oracle" not in text and "oracle_api" not in text and "lho_" not in text, os.path.join(base, f)
