"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/pvqa.h declares, reports argument errors without touching the device, and the Python host side
refuses CPU tensors (there is no CPU fallback in the product path)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "pvqa.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pvqa_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import phoneme_vqa_b200 as pv
    lib = pv.load()
    names = _declared_symbols()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pvqa.h but not exported by libpvqa_sm100.so"
    assert lib.pvqa_abi_version() == int(re.search(r"#define PVQA_ABI_VERSION (\d+)", open(
        os.path.join(ROOT, "include", "pvqa.h")).read()).group(1))


def test_every_binding_has_a_signature():
    from phoneme_vqa_b200 import _lib
    missing = [n for n in _declared_symbols() if n not in _lib._SIGNATURES]
    assert not missing, missing


def test_argument_errors_are_reported_without_a_device():
    import phoneme_vqa_b200 as pv
    lib = pv.load()
    # d not a multiple of 8 -> PVQA_ERR_SHAPE, message available, nothing launched
    before = lib.pvqa_launch_count()
    rc = lib.pvqa_embed_mm_fwd(None, None, None, None, None, None, None, None, None, None,
                               2, 3, 4, 5, 7, 10, 1024, 0, 0, None, None)
    assert rc == 1
    assert b"multiple of 8" in lib.pvqa_last_error()
    rc = lib.pvqa_attn_fwd(None, None, None, None, None, None, None, 1, 1, 8, 8, 32,
                           0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1.0, 0, 0.0, 0, 0, None, None, 0, 0, None)
    assert rc == 1 and b"head dim" in lib.pvqa_last_error()
    rc = lib.pvqa_embed_tgt_fwd(None, None, None, None, None, None, 1, 1, 10, 4, 4, 1, 1, 1, 0, 0, 0.0, 0, 0, None, None)
    assert rc == 1 and b"on_dim" in lib.pvqa_last_error()
    rc = lib.pvqa_phoneme_head_ce_fwd(None, None, 3, None, None, None, None, None, None, None, None, None, None, None,
                                      None, 4, 12, 4, 4, 5, 5, 5, 0, 7, 0, None)
    assert rc != 0
    assert lib.pvqa_launch_count() == before
    # empty problems are OK and launch nothing
    assert lib.pvqa_embed_mm_fwd(None, None, None, None, None, None, None, None, None, None,
                                 0, 3, 4, 5, 8, 10, 1024, 0, 0, None, None) == 0
    assert lib.pvqa_launch_count() == before


def test_product_ops_refuse_cpu_tensors():
    from phoneme_vqa_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.embed_multimodal(torch.zeros(1, 2, 8), None, None, torch.zeros(1, 3, dtype=torch.long), None,
                             torch.ones(1, 3), torch.zeros(10, 8), ())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.attention_self(torch.zeros(1, 4, 3, 1, 64), 1.0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.rms_norm(torch.zeros(2, 128), torch.ones(128), 1e-6, torch.float32)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from phoneme_vqa_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "phoneme-vqa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f
