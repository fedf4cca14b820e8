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


def _prototypes():
    """name -> (return type, [argument types]) parsed from include/tf2b200.h"""
    with open(os.path.join(ROOT, "include", "tf2b200.h")) as f:
        txt = re.sub(r"/\*.*?\*/", " ", f.read(), flags=re.S)
    out = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(tf2b_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", txt):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                types.append("ptr" if "*" in a else ("i64" if "int64_t" in a else "int"))
        out[name] = (ret, types)
    return out


def test_ctypes_prototypes_match_the_header():
    """Every argtypes / restype capi.py sets is the header's declaration: argument count, pointer vs integer,
    64-bit returns — a mismatch would corrupt the call silently."""
    lib = capi.load()
    protos = _prototypes()
    assert sorted(protos) == sorted(capi.SYMBOLS)
    for name, (ret, types) in protos.items():
        fn = getattr(lib, name)
        at = fn.argtypes
        if types:
            assert at is not None and len(at) == len(types), f"{name}: {len(at or [])} argtypes for {len(types)} parameters"
            for i, (want, got) in enumerate(zip(types, at)):
                is_ptr = got in (C.c_void_p, C.c_char_p) or hasattr(got, "contents") or issubclass(got, C._Pointer)
                if want == "ptr":
                    assert is_ptr, f"{name} arg {i}: header pointer, ctypes {got}"
                else:
                    assert not is_ptr and C.sizeof(got) == (8 if want == "i64" else 4), f"{name} arg {i}: header {want}, ctypes {got}"
        if "char" in ret and "*" in ret:
            assert fn.restype is C.c_char_p, name
        elif ret == "void":
            assert fn.restype is None, name
        elif "int64_t" in ret:
            assert C.sizeof(fn.restype) == 8, name
        else:
            assert fn.restype is C.c_int, name


def test_ctypes_structs_match_the_header():
    with open(os.path.join(ROOT, "include", "tf2b200.h")) as f:
        txt = re.sub(r"/\*.*?\*/", " ", f.read(), flags=re.S)
    m = re.search(r"typedef struct \{([^}]*)\}\s*tf2b_layer_desc;", txt)
    fields = [n.strip() for decl in m.group(1).split(";") if decl.strip() for n in decl.replace("int32_t", "").split(",")]
    assert fields == [n for n, _ in capi.LayerDescC._fields_]
    assert C.sizeof(capi.LayerDescC) == 4 * len(fields)
    m = re.search(r"typedef struct \{([^}]*)\}\s*tf2b_tensor_desc;", txt)
    fields = [n.strip() for decl in m.group(1).split(";") if decl.strip() for n in decl.replace("int32_t", "").split(",")]
    assert fields == [n for n, _ in capi.TensorDescC._fields_] == ["C", "H", "W"]
