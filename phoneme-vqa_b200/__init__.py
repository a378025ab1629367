"""phoneme-vqa_b200 — B200-native (sm_100a) hot path of PhonoVQA (hieunghia-pat/phoneme-VQA).

Layout
  csrc/      hand-written CUDA kernels + the C-ABI (include/pvqa.h) -> libpvqa_sm100.so
  _lib.py    ctypes loader / in-tree nvcc build
  ops.py     autograd wrappers around the C-ABI entry points
  modules.py / models.py   host-side mirror of the reference model classes
             (same class names, ctor/forward/generate signatures, state_dict keys)
"""
from . import _lib  # noqa: F401
from ._lib import build, load, launch_count  # noqa: F401

__version__ = "0.1.0"
