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


def test_binding_argument_types_are_the_header_prototypes():
    """Every ctypes prototype in worldforge_b200/lib.py has the argument list of its declaration in include/wf_b200.h
    (pointers -> c_void_p, int / float / long long / unsigned by value).  The .cu files include the same header, so the
    compiler holds the definitions to it; this holds the Python side to it."""
    import ctypes as C
    from worldforge_b200 import lib
    src = open(os.path.join(ROOT, "include", "wf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = re.findall(r"\b([A-Za-z_][A-Za-z0-9_ \*]*?)\b(wf_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src)
    assert sorted(n for _, n, _ in protos) == declared()
    scalar = {"int": C.c_int, "float": C.c_float, "long long": C.c_longlong, "unsigned": C.c_uint, "unsigned int": C.c_uint}

    def ctype(param):
        param = param.strip()
        if param in ("void", ""):
            return None
        if "*" in param:
            return C.c_void_p
        return scalar[re.sub(r"\bconst\b", "", re.sub(r"\b[A-Za-z_][A-Za-z0-9_]*$", "", param)).strip()]

    for ret, name, args in protos:
        got = [t for t in (ctype(a) for a in args.split(",")) if t is not None]
        want = lib._SIGNATURES[name] if name in lib._SIGNATURES else lib._PLAIN[name][1]
        assert got == want, (name, [t.__name__ for t in got], [t.__name__ for t in want])
        if name in lib._SIGNATURES:
            assert ret.strip() == "int", name                    # error code
