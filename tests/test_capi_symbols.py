"""CPU: the C-ABI library loads and exports every symbol include/tf2b200.h declares; without a GPU
the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from tests.conftest import ROOT
from tf2_b200 import capi, nets


def _declared():
    with open(os.path.join(ROOT, "include", "tf2b200.h")) as f:
        txt = f.read()
    return sorted(set(re.findall(r"\b(tf2b_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    lib = capi.load()
    decl = _declared()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in tf2b200.h but not exported"
    assert sorted(capi.SYMBOLS) == decl, "capi.SYMBOLS out of sync with include/tf2b200.h"
    assert b"sm_100a" in lib.tf2b_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tf2_b200.network import NetWork, Tf2bError
    with pytest.raises(Tf2bError) as ei:
        NetWork(nets.chain((16, 8, 8), [dict(N=16, k=1)]), device=0)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_create_rejects_bad_tables():
    lib = capi.load()
    h = C.c_void_p()
    assert lib.tf2b_create(None, 0, None, 0, 0, C.byref(h)) == -1
    assert b"null" in lib.tf2b_last_error(None)
