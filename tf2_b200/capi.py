"""ctypes binding of libtf2b200.so (the C ABI declared in include/tf2b200.h).

Fails loudly when the CUDA extension is missing or no GPU is present — there is no CPU fallback
on the product path."""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

# TF2B_LIB: development override (A/B builds of the same sources under tools/); the product is lib/libtf2b200.so
_LIB_PATH = os.environ.get("TF2B_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libtf2b200.so")

TF2B_OK = 0
VARIANT_AUTO, VARIANT_SHIFT, VARIANT_MMA = 0, 1, 2
LAYOUT_CHW, LAYOUT_HWC = 0, 1
WEIGHTS_PLANES, WEIGHTS_PACKED4 = 0, 1


class BiasBn(C.Structure):
    _fields_ = [("bias", C.c_int32), ("alpha", C.c_int32), ("beta", C.c_int32)]


class TensorDescC(C.Structure):
    _fields_ = [("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32)]


class LayerDescC(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "in_tensor", "out_tensor", "out_ch0", "add_tensor", "C", "N", "k", "pad", "stride", "OH", "OW",
        "relu", "pool", "pool_stride", "pool_pad", "PH", "PW", "add_relu", "gap", "ipool",
        "in_may_be_m128")]


# every symbol include/tf2b200.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "tf2b_create", "tf2b_load_layer", "tf2b_load_layer_packed4", "tf2b_finalize", "tf2b_set_variant", "tf2b_set_graph", "tf2b_set_weight_staging", "tf2b_set_stem_chunk",
    "tf2b_weight_blob_bytes", "tf2b_export_weight_blob", "tf2b_import_weight_blob", "tf2b_run",
    "tf2b_run_raw224", "tf2b_run_raw224_host", "tf2b_submit_raw224_host", "tf2b_submit_host", "tf2b_wait", "tf2b_run_host", "tf2b_set_result", "tf2b_read_tensor",
    "tf2b_dump_acc", "tf2b_set_profile", "tf2b_get_profile", "tf2b_last_launches", "tf2b_layer_kernel", "tf2b_layer_mode", "tf2b_last_error", "tf2b_version",
    "tf2b_destroy",
]

_lib: Optional[C.CDLL] = None


def lib_path() -> str:
    return _LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"{_LIB_PATH} not found: build the CUDA extension first "
            f"(python -c 'import __graft_entry__ as g; g.build()' or tf2_b200/csrc/build.sh). "
            f"tf2_b200 has no CPU fallback.")
    lib = C.CDLL(_LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    lib.tf2b_create.argtypes = [C.POINTER(TensorDescC), i32, C.POINTER(LayerDescC), i32, i32, C.POINTER(vp)]
    lib.tf2b_load_layer.argtypes = [vp, i32, vp, vp]
    lib.tf2b_load_layer_packed4.argtypes = [vp, i32, vp, i32, vp, vp, vp]
    lib.tf2b_finalize.argtypes = [vp, i32]
    lib.tf2b_set_variant.argtypes = [vp, i32]
    lib.tf2b_set_graph.argtypes = [vp, i32]
    lib.tf2b_set_weight_staging.argtypes = [vp, i32]
    lib.tf2b_set_stem_chunk.argtypes = [vp, i32]
    lib.tf2b_weight_blob_bytes.argtypes = [vp]
    lib.tf2b_weight_blob_bytes.restype = i64
    lib.tf2b_export_weight_blob.argtypes = [vp, vp, vp]
    lib.tf2b_import_weight_blob.argtypes = [vp, vp, i64, vp]
    lib.tf2b_run.argtypes = [vp, vp, i32, i32, vp, i32, vp]
    lib.tf2b_run_raw224.argtypes = [vp, vp, i32, vp, i32, vp]
    lib.tf2b_run_raw224_host.argtypes = [vp, vp, i32, vp, i32]
    lib.tf2b_run_host.argtypes = [vp, vp, i32, i32, vp, i32]
    lib.tf2b_submit_raw224_host.argtypes = [vp, vp, i32, vp, i32, i32]
    lib.tf2b_submit_host.argtypes = [vp, vp, i32, i32, vp, i32, i32]
    lib.tf2b_wait.argtypes = [vp, i32]
    lib.tf2b_set_result.argtypes = [vp, i32]
    lib.tf2b_read_tensor.argtypes = [vp, i32, i32, vp, i32, vp]
    lib.tf2b_dump_acc.argtypes = [vp, i32, i32, vp, vp]
    lib.tf2b_set_profile.argtypes = [vp, i32]
    lib.tf2b_get_profile.argtypes = [vp, vp, vp, i32]
    lib.tf2b_last_launches.argtypes = [vp]
    lib.tf2b_layer_kernel.argtypes = [vp, i32]
    lib.tf2b_layer_kernel.restype = C.c_char_p
    lib.tf2b_layer_mode.argtypes = [vp, i32, i32]
    lib.tf2b_layer_mode.restype = C.c_char_p
    lib.tf2b_last_error.argtypes = [vp]
    lib.tf2b_last_error.restype = C.c_char_p
    lib.tf2b_version.restype = C.c_char_p
    lib.tf2b_destroy.argtypes = [vp]
    lib.tf2b_destroy.restype = None
    for name in ("tf2b_create", "tf2b_load_layer", "tf2b_load_layer_packed4", "tf2b_finalize",
                 "tf2b_set_variant", "tf2b_set_graph", "tf2b_set_weight_staging", "tf2b_set_stem_chunk", "tf2b_export_weight_blob", "tf2b_import_weight_blob", "tf2b_run",
                 "tf2b_run_raw224", "tf2b_run_raw224_host", "tf2b_submit_raw224_host", "tf2b_submit_host", "tf2b_wait", "tf2b_run_host", "tf2b_set_result",
                 "tf2b_read_tensor", "tf2b_dump_acc", "tf2b_last_launches", "tf2b_set_profile", "tf2b_get_profile"):
        getattr(lib, name).restype = i32
    _lib = lib
    return lib


class Tf2bError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"tf2b error {code}: {msg}")
        self.code = code
