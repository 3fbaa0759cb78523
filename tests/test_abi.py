"""The C-ABI library loads without a GPU and exports every symbol include/wf_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    src = open(os.path.join(ROOT, "include", "wf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wf_[a-z0-9_]+)\s*\(", src)))


def test_build_and_load():
    import __graft_entry__ as g
    g.build()
    from worldforge_b200 import lib
    h = lib.load()
    assert h.wf_abi_version() == 1
    assert h.wf_last_error() is not None


def test_header_and_binding_agree_and_every_symbol_is_exported():
    from worldforge_b200 import lib
    names = declared()
    assert len(names) >= 30
    assert names == lib.exported_symbols(), set(names) ^ set(lib.exported_symbols())
    dll = ctypes.CDLL(lib.LIB_PATH)
    for n in names:
        assert hasattr(dll, n), n


def test_no_cpu_fallback():
    """Host tensors are refused; nothing in the product imports the oracle."""
    import torch
    from worldforge_b200 import lib
    with pytest.raises(lib.WfError):
        lib.cfg_combine(torch.zeros(8), torch.zeros(8), 1.0)
    pkg = os.path.join(ROOT, "worldforge_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            text = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in text and "from oracle" not in text, fn


def test_argument_validation_without_gpu():
    """WF_EINVAL paths return before any CUDA call."""
    from worldforge_b200 import lib
    h = lib.load()
    assert h.wf_gemm_bf16(None, 8, None, 8, None, None, 8, None, 0, 1, 32, 8, 0, None) == -1
    assert b"null" in h.wf_last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert h.wf_gemm_bf16(p, 8, p, 8, None, p, 8, None, 0, 4, 33, 8, 0, None) == -1      # N not a multiple of 32
    assert h.wf_cfg_combine(p, p, p, 1, 1.0, 6, None) == -1                             # not a multiple of 4
    assert h.wf_rms_norm_rope(p, 8, p, p, 1, 72, 1e-6, None) == -1                      # RoPE needs head_dim 128
