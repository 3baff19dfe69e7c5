"""lighthugger_b200 -- B200-native geometry pipeline of expenses/lighthugger.

The product is the C-ABI CUDA library (include/lighthugger_b200.h,
lighthugger_b200/csrc) plus the C++20 host scene preparation
(include/lh_scene.h, lighthugger_b200/host).  This Python package is a thin
ctypes binding over both, used by tests/ and bench.py; it contains no compute
and has no CPU fallback: creating a Context fails loudly when the CUDA library
or a GPU is missing.
"""
from . import _structs as structs  # noqa: F401
from .scene import Scene, HostBinding, host_lib  # noqa: F401
from .renderer import Context, LhError, cuda_lib, cuda_lib_path  # noqa: F401
