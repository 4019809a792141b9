"""CPU-side check of the drop-in boundary: the shared library builds (nvcc cross
compiles without a GPU), loads, and exports every function include/*.h declares.
No compute call is made here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as g
    if not os.path.isfile(g.LIB):
        g.build()
    return g.LIB


def declared_functions():
    with open(os.path.join(ROOT, "include", "breakmer_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bk_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = declared_functions()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), n


def test_binding_lists_the_same_symbols():
    from breakmer_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared_functions()


def test_no_cpu_fallback_without_a_device(lib_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from breakmer_b200 import _lib
    with pytest.raises(_lib.BreakmerError):
        _lib.Handle(0)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "breakmer_b200")
    for dp, _dn, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dp, fn)) as f:
                    src = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn
                assert "liboracle" not in src, fn


def test_concat_str_bytes_and_non_ascii():
    """The binding's packer: one join for all-str input, per-record encoding otherwise; offsets are byte offsets."""
    import numpy as np
    from breakmer_b200._lib import concat
    for case in ([], ["ACGT", "", "NNA"], [b"AC", b"GT"], ["AC", b"GT"], ["é", "AC"], [""], (s for s in ["A", "CG"])):
        case = list(case)
        data, off = concat(iter(case))
        bs = [s.encode() if isinstance(s, str) else bytes(s) for s in case]
        assert data.tobytes() == b"".join(bs)
        assert off.dtype == np.int64 and list(off) == [0] + list(np.cumsum([len(b) for b in bs]))
