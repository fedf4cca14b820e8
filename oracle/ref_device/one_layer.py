"""Builds the REFERENCE's whole device pipeline (cnn.cl compiled as C) for ONE arbitrary layer run as
"layer 0" of a one-layer network — the compiled-reference oracle for the convolution geometry of
layers the shipped tables never place first (1x1, padded 3x3, stride 2, 5x5).

The one-layer table header is GENERATED from the reference's own googlenet.h where it lies (its
macro block is kept as is, every per-layer table is cut to one entry with this layer's values, the
static cycle tables are switched off so that the reference's cycle.cl derives the schedule:
`#ifndef STATIC_CYCLE`, googlenet.h / cycle.cl:27-252).  The header, the host loaders compiled
against it (InputConvert / FilterConvert address the device buffers through the same tables) and
the device pipeline go to oracle/_ref/one_<hash>/ — generated, git-ignored, never committed.
Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import re
import subprocess
from typing import Dict

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TF2_REFERENCE", "/root/reference")


def ceil(a, b):
    return -(-a // b)


def layer_tables(cfg: Dict[str, int]) -> Dict[str, int]:
    """Per-layer table entries of a one-layer network; the derived ones follow the formulas the
    shipped headers obey on every layer (checked in tests/test_single_layer_ref.py)."""
    C_, N, k, pad, s = cfg["C"], cfg["N"], cfg["k"], cfg["pad"], cfg["stride"]
    IH, IW = cfg["IH"], cfg["IW"]
    oh1, ow1 = IH + 2 * pad - k + 1, IW + 2 * pad - k + 1          # stride-1 output (kOutputHeight/Width)
    ph, pw = (oh1 - 1) // s + 1, (ow1 - 1) // s + 1                 # what reaches the feature writer
    t = dict(
        kCacheReadBase=0, kCacheWriteBase="C2", kDDRReadBase=0, kDDRWriteBase=0, kCacheWriteEnable=1, kDDRWriteEnable=0,
        kEndPoolEnable=0, kAdditionEnable=0, kAdditionReluEnable=0, kReluEnable=cfg.get("relu", 1), kFilterSize=k,
        kPadWidth=pad, kPadHeight=pad, kInputWidth=IW, kInputHeight=IH, kOutputWidth=ow1, kOutputHeight=oh1,
        kInputChannels=C_, kOutputChannels=N, kWvecEnd=ceil(IW, 7), kConvStride=s, kIpoolEnable=0, kPoolEnable=0,
        kBiasEnable=1, kPoolWindow=3, kPoolType=0, kPoolStride2=0, kPoolOutputWidth=pw, kPoolOutputHeight=ph,
        kPoolOutputWvecEnd=ceil(pw, 7), kOhEndWithOffset=ceil(oh1, s) + 2, kOwEndWithOffset=ow1 + 2,
        kFWvecEnd=ceil(k, 3), kCvecEnd=ceil(C_, 16), kFilterCvecEnd=ceil(C_, 48) if k == 1 else ceil(C_, 16),
        kNvecEnd=ceil(N, 16), kNEndWithOffset=N, kNStart=0, kNEnd=N, kPoolPad=0, kBnEnable=1, kInputLayer=0,
        kBranchTail=0, kConcatLayer=0, kSequencerIdleCycle=0)
    t["kFilterLoadSize"] = t["kFilterCvecEnd"] * k * t["kFWvecEnd"]
    return t


def generate_header(cfg: Dict[str, int]) -> str:
    src = open(os.path.join(REF, "Runtime_Engine", "cnn", "host", "inc", "googlenet.h")).read()
    t = layer_tables(cfg)
    src = src.replace("#define STATIC_CYCLE", "//#define STATIC_CYCLE")
    src = re.sub(r"#define NUM_LAYER\s+\d+", "#define NUM_LAYER 1", src)
    src = re.sub(r"#define NUM_CONVOLUTIONS\s+\d+", "#define NUM_CONVOLUTIONS 1", src)
    src = re.sub(r"#define INPUT_IMAGE_C\s+\d+", f"#define INPUT_IMAGE_C {cfg['C']}", src)
    src = re.sub(r"#define INPUT_IMAGE_H\s+\d+", f"#define INPUT_IMAGE_H {cfg['IH']}", src)
    src = re.sub(r"#define INPUT_IMAGE_W\s+\d+", f"#define INPUT_IMAGE_W {cfg['IW']}", src)
    seen = set()

    def repl(m):
        typ, name = m.group(1), m.group(2)
        if name not in t:
            return m.group(0)            # static cycle tables (inside #ifdef STATIC_CYCLE) stay untouched
        seen.add(name)
        v = t[name]
        if name == "kFilterLoadSize":    # cycle.cl:58-63 also reads the NEXT layer's entry
            return f"CONSTANT {typ} {name}[NUM_CONVOLUTIONS + 1] = {{ {v}, {v} }};"
        return f"CONSTANT {typ} {name}[NUM_CONVOLUTIONS] = {{ {v} }};"

    src = re.sub(r"CONSTANT\s+(\w+)\s+(\w+)\s*\[\s*NUM_CONVOLUTIONS\s*\]\s*=\s*\{[^}]*\}\s*;", repl, src)
    missing = sorted(set(t) - seen)
    if missing:
        raise RuntimeError(f"tables not found in googlenet.h: {missing}")
    return src


def build(cfg: Dict[str, int]) -> str:
    """Returns the directory holding libhost.so / libdev.so for this layer geometry (built on demand)."""
    if not os.path.isdir(os.path.join(REF, "Runtime_Engine")):
        raise FileNotFoundError(REF)
    h = hashlib.sha1(repr(sorted(cfg.items())).encode())
    for src in ("full_harness.c", "fifo_shim.h", "one_layer.py", os.path.join("..", "ref_host_shim.cpp")):
        with open(os.path.join(_HERE, src), "rb") as f:      # a changed harness must not meet a cached library
            h.update(f.read())
    key = h.hexdigest()[:12]
    out = os.path.join(os.path.dirname(_HERE), "_ref", f"one_{key}")
    if os.path.exists(os.path.join(out, "libdev.so")) and os.path.exists(os.path.join(out, "libhost.so")):
        return out
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "one_layer.h"), "w") as f:
        f.write(generate_header(cfg))
    cnn = os.path.join(REF, "Runtime_Engine", "cnn")
    host, dev, common = os.path.join(cnn, "host"), os.path.join(cnn, "device", "src"), os.path.join(REF, "Runtime_Engine", "common", "inc")
    # the reference's cnn.h is skipped (__CNN_H__), its three generic headers + the generated tables take its place
    with open(os.path.join(out, "prelude.h"), "w") as f:
        f.write('#define __CNN_H__\n#include "archs.h"\n#include "defines.h"\n#include "types.h"\n#include "one_layer.h"\n')
    subprocess.check_call(["g++", "-std=c++11", "-O1", "-fPIC", "-shared", "-w", "-fopenmp", "-DPRINT_LEVEL_QUIET",
                           "-I" + os.path.join(os.path.dirname(_HERE), "stub"), "-I" + common, "-I" + os.path.join(host, "inc"),
                           "-I" + out, "-include", os.path.join(out, "prelude.h"),
                           os.path.join(host, "src", "model_loader.cpp"), os.path.join(host, "src", "quantization.cpp"),
                           os.path.join(host, "src", "input_loader.cpp"), os.path.join(host, "src", "debug.cpp"),
                           os.path.join(host, "src", "network_helper.cpp"),
                           os.path.join(os.path.dirname(_HERE), "ref_host_shim.cpp"), "-o", os.path.join(out, "libhost.so")])
    subprocess.check_call(["/usr/bin/gcc", "-x", "c", "-std=gnu11", "-O1", "-fPIC", "-shared", "-w", "-DTF2_ONE_LAYER",
                           "-I" + dev, "-I" + os.path.join(host, "inc"), "-I" + _HERE, "-I" + out,
                           os.path.join(_HERE, "full_harness.c"), "-o", os.path.join(out, "libdev.so")])
    return out


def run(cfg: Dict[str, int], x: np.ndarray, codes: np.ndarray, params: np.ndarray):
    """x int8 [C][IH][IW]; codes uint8 [N][C][k][k]; params int32 [N][3] -> (int8 [N][PH][PW], counts, consts)"""
    d = build(cfg)
    Lh, Lf = C.CDLL(os.path.join(d, "libhost.so")), C.CDLL(os.path.join(d, "libdev.so"))
    Lf.full_const.restype = C.c_longlong
    Lh.ref_input_device_size.restype = C.c_longlong
    Lh.ref_filter_device_size.restype = C.c_longlong
    consts = [Lf.full_const(i) for i in range(5)]
    isz, fsz, mb = Lh.ref_input_device_size(), Lh.ref_filter_device_size(), Lh.ref_max_bias_size()
    inp_f = np.zeros(isz, np.float32)
    xr = np.ascontiguousarray(x, dtype=np.float32)
    Lh.ref_input_convert.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    Lh.ref_input_convert(xr.ctypes.data, inp_f.ctypes.data, 1)
    inp = inp_f.astype(np.int8)
    fraw = np.full(fsz, 64, np.uint8)
    fraw[:codes.size] = np.ascontiguousarray(codes, dtype=np.uint8).reshape(-1)
    freal = np.full(fsz, 64, np.uint8)
    scratch = np.zeros(fsz, np.uint8)
    Lh.ref_filter_convert.argtypes = [C.c_void_p] * 3
    Lh.ref_filter_convert(scratch.ctypes.data, fraw.ctypes.data, freal.ctypes.data)
    N = codes.shape[0]
    bb = np.zeros((mb + 16, 3), np.int32)
    bb[:N] = params
    ddr = np.zeros(8 << 20, np.int8)
    ib, io = Lf.full_item_bytes(), Lf.full_item_data_offset()
    cap = 40000
    cache = np.zeros(cap * ib, np.uint8)
    nc = C.c_longlong(0)
    counts = np.zeros(8, np.int64)
    Lf.full_run_layer0.argtypes = [C.c_void_p] * 3 + [C.c_longlong] + [C.c_void_p] * 2 + [C.c_longlong, C.c_void_p, C.c_void_p]
    rc = Lf.full_run_layer0(inp.ctypes.data, freal.ctypes.data, bb.ctypes.data, consts[0], ddr.ctypes.data,
                            cache.ctypes.data, cap, C.byref(nc), counts.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"full_run_layer0 failed: {rc}")
    n = nc.value
    t = layer_tables(cfg)
    ph, pw = t["kPoolOutputHeight"], t["kPoolOutputWidth"]
    nvec, pwv = ceil(N, 16), ceil(pw, 7)
    if n != nvec * ph * pwv:
        raise RuntimeError(f"reference pipeline emitted {n} output tiles, expected {nvec * ph * pwv} (counts {counts}, consts {consts})")
    data = cache[: n * ib].reshape(n, ib)[:, io:io + 128].view(np.int8).reshape(nvec, ph, pwv, 8, 16)[:, :, :, :7, :]
    out = np.ascontiguousarray(data.transpose(0, 4, 1, 2, 3).reshape(nvec * 16, ph, pwv * 7)[:N, :, :pw])
    return out, counts, consts
