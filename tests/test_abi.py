"""CPU-side checks of the C-ABI boundary: the library loads and exports every symbol include/shgvqa.h declares
(no compute calls without a GPU), and the Python binding covers the same set."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "shgvqa.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(shg_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as ge
    from shg_vqa_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        ge.build()
    return _lib


def test_header_symbols_are_exported(built):
    names = _declared()
    assert len(names) >= 10
    so = ctypes.CDLL(built.LIB_PATH)
    for n in names:
        assert hasattr(so, n), "libshgvqa.so does not export %s" % n


def test_binding_matches_header(built):
    names = set(_declared()) - {"shg_last_error"}
    assert names == set(built.SIGNATURES), names ^ set(built.SIGNATURES)


def test_version_and_error_text(built):
    lib = built.lib()
    assert lib.shg_version() >= 100
    assert isinstance(lib.shg_last_error(), bytes)


def test_product_ops_refuse_cpu_tensors(built):
    import torch
    from shg_vqa_b200 import ops
    with pytest.raises(built.ShgError):
        ops.layernorm_fwd(torch.zeros(4, 768, dtype=torch.bfloat16), torch.ones(768), torch.zeros(768), 1e-12)
