"""Builds the REFERENCE's device program (cnn.cl compiled as C, kernels as coroutines: net_harness.c) for
an ARBITRARY chain / concat network given as a NetDesc — the compiled-reference oracle for the BASELINE
networks the reference ships no tables for (VGG16, SqueezeNet).

Like one_layer.py, the table header is GENERATED from the reference's own googlenet.h where it lies: every
per-layer table is rewritten with this network's values (the derived entries follow the formulas the
shipped headers obey on every layer, tests/test_single_layer_ref.py), the size macros are raised to fit,
the static cycle tables are switched off so that the reference's cycle.cl derives the schedule.  What the
tables cannot say for themselves is planned here the way the shipped headers do it: every tensor gets a
page of the on-chip feature cache for as long as a later layer reads it (kCacheReadBase / kCacheWriteBase),
concat branches write into their buffer at kNStart.  Residual adds and ipool layers are not generated (the
shipped ResNet50 / GoogLeNet tables cover them).  Outputs go to oracle/_ref/net_<hash>/ — generated,
git-ignored, never committed.  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import hashlib
import json
import os
import re
import subprocess
from typing import Dict, List

from .one_layer import REF, ceil, layer_tables

_HERE = os.path.dirname(os.path.abspath(__file__))


def _tensor_entries(t) -> int:
    return ceil(t.C, 16) * t.H * ceil(t.W, 7)


def plan(net):
    """-> (per-layer table dict name -> list, macro dict).  Cache pages by liveness."""
    L = len(net.layers)
    last_use: Dict[int, int] = {}
    for l, ld in enumerate(net.layers):
        if ld.ipool or ld.add_tensor >= 0 or ld.gap:
            raise NotImplementedError("multi_layer: ipool / residual / end-pool layers are covered by the shipped tables")
        last_use[ld.in_tensor] = l
    page_of: Dict[int, int] = {0: 0}
    free: List[int] = []
    n_pages = 1
    rows: List[Dict[str, int]] = []
    concat_id: Dict[int, int] = {}
    for l, ld in enumerate(net.layers):
        tin = net.tensors[ld.in_tensor]
        t = layer_tables(dict(C=ld.C, N=ld.N, k=ld.k, pad=ld.pad, stride=ld.stride, IH=tin.H, IW=tin.W, relu=ld.relu))
        if ld.out_tensor not in page_of:
            if free:
                page_of[ld.out_tensor] = free.pop(0)
            else:
                page_of[ld.out_tensor] = n_pages
                n_pages += 1
        tail = 1 if net.branch_tail and net.branch_tail[l] else 0
        if tail:
            concat_id.setdefault(ld.out_tensor, net.concat_layer[l])
        t.update(
            kCacheReadBase=f"C{page_of[ld.in_tensor] + 1}", kCacheWriteBase=f"C{page_of[ld.out_tensor] + 1}",
            kCacheWriteEnable=1 if l < L - 1 else 0, kDDRWriteEnable=0, kDDRReadBase=0, kDDRWriteBase=0,
            kPoolEnable=ld.pool, kPoolStride2=1 if (ld.pool and ld.pool_stride == 2) else 0, kPoolPad=ld.pool_pad,
            kPoolOutputWidth=ld.PW, kPoolOutputHeight=ld.PH, kPoolOutputWvecEnd=ceil(ld.PW, 7),
            kBiasEnable=ld.bias_en, kBnEnable=1, kInputLayer=ld.q_in_row, kBranchTail=tail,
            kConcatLayer=net.concat_layer[l] if tail else 0, kNStart=ld.out_ch0, kNEnd=ld.out_ch0 + ld.N)
        rows.append(t)
        # pages whose tensor nobody reads any more are free for the NEXT layer's output
        for tid in list(page_of):
            if last_use.get(tid, -1) <= l and tid != ld.out_tensor and tid in page_of and last_use.get(tid, -1) >= 0:
                if last_use[tid] == l:
                    free.append(page_of[tid])
                    last_use[tid] = -2
    tables = {k: [r[k] for r in rows] for k in rows[0]}
    conv = [ld for ld in net.layers]
    kmax = max(ld.k for ld in conv)
    t0 = net.tensors[0]
    page = max(_tensor_entries(t) for t in net.tensors)
    macros = dict(
        NUM_LAYER=L, NUM_CONVOLUTIONS=L, INPUT_IMAGE_C=t0.C, INPUT_IMAGE_H=t0.H, INPUT_IMAGE_W=t0.W,
        FIRST_FILTER_SIZE=net.layers[0].k, MAX_OUT_CHANNEL=max(1024, max(ld.N for ld in conv)),
        MAX_POOL_OUTPUT_WVEC=max(ceil(ld.PW, 7) for ld in conv),
        DDR_PAGE_SIZE0=page, DDR_PAGE_SIZE1=page, CACHE_PAGE_SIZE=page, CACHE_SIZE=f"(CACHE_PAGE_SIZE * {n_pages})",
        FILTER_CACHE_PAGE_SIZE1=max([16 * ceil(ld.C, 16) * ld.k * ceil(ld.k, 3) for ld in conv if ld.k > 1] or [48]),
        FILTER_CACHE_PAGE_SIZE2=max([48 * ceil(ld.C, 48) for ld in conv if ld.k == 1] or [48]),
        MAX_BIAS_SIZE=16 * ceil(max(ld.N for ld in conv), 16), POOL_WINDOW_MAX=3)
    per_layer_filter = max((ceil(ld.C, 48) if ld.k == 1 else ceil(ld.C, 16)) * ld.k * ceil(ld.k, 3) * 16 * ceil(ld.N, 16) * 64 for ld in conv)
    p2 = 1
    while p2 < per_layer_filter:
        p2 *= 2
    macros["MAX_FILTER_SIZE1"] = macros["MAX_FILTER_SIZE2"] = p2
    scal = dict(
        kFilterSizeMax=kmax, kInputWidthMax=max(tables["kInputWidth"]), kInputHeightMax=max(tables["kInputHeight"]),
        kOutputWidthMax=max(tables["kOutputWidth"]), kOutputHeightMax=max(tables["kOutputHeight"]),
        kOutputChannelsMax=max(tables["kOutputChannels"]), kWvecEndMax=max(tables["kWvecEnd"]),
        kPoolOutputWidthMax=max(tables["kPoolOutputWidth"]), kPoolOutputHeightMax=max(tables["kPoolOutputHeight"]),
        kOhEndWithOffsetMax=max(tables["kOhEndWithOffset"]), kOwEndWithOffsetMax=max(tables["kOwEndWithOffset"]),
        kFWvecEndMax=max(tables["kFWvecEnd"]), kCvecEndMax=max(tables["kCvecEnd"]),
        kFilterCvecEndMax=max(tables["kFilterCvecEnd"]), END_WW_MAX_INPUT_READER=ceil(max(tables["kInputWidth"]), 3),
        kNvecEndMax=max(tables["kNvecEnd"]), kNEndWithOffsetMax=max(tables["kNEndWithOffset"]))
    return tables, macros, scal, n_pages


def generate_header(net) -> str:
    src = open(os.path.join(REF, "Runtime_Engine", "cnn", "host", "inc", "googlenet.h")).read()
    tables, macros, scal, n_pages = plan(net)
    L = len(net.layers)
    src = src.replace("#define STATIC_CYCLE", "//#define STATIC_CYCLE")
    for name, val in macros.items():
        src, n = re.subn(r"(#define\s+" + name + r")[ \t]+[^\n]*", lambda m: f"{m.group(1)} {val}", src, count=1)
        if n != 1:
            raise RuntimeError(f"macro {name} not found in googlenet.h")
    src, n = re.subn(r"#define NUM_Q_LAYERS[^\n]*", f"#define NUM_Q_LAYERS {net.num_q_rows}", src, count=1)
    pages = "\n".join(f"#define C{i + 1} ({i} * CACHE_PAGE_SIZE)" for i in range(max(n_pages, 3)))
    src = re.sub(r"#define C1 0\s*\n#define C2 CACHE_PAGE_SIZE\s*\n#define C3 \(2 \* CACHE_PAGE_SIZE\)", pages, src, count=1)
    if "#define C1 (0 * CACHE_PAGE_SIZE)" not in src:
        raise RuntimeError("cache page macros not found in googlenet.h")
    seen = set()

    def repl(m):
        typ, name = m.group(1), m.group(2)
        if name not in tables:
            return m.group(0)            # static cycle tables (inside #ifdef STATIC_CYCLE) stay untouched
        seen.add(name)
        v = list(tables[name])
        if name == "kFilterLoadSize":    # cycle.cl:58-63 also reads the NEXT layer's entry
            return f"CONSTANT {typ} {name}[NUM_CONVOLUTIONS + 1] = {{ {', '.join(map(str, v + [v[-1]]))} }};"
        return f"CONSTANT {typ} {name}[NUM_CONVOLUTIONS] = {{ {', '.join(map(str, v))} }};"

    src = re.sub(r"CONSTANT\s+(\w+)\s+(\w+)\s*\[\s*NUM_CONVOLUTIONS\s*\]\s*=\s*\{[^}]*\}\s*;", repl, src)
    missing = sorted(set(tables) - seen)
    if missing:
        raise RuntimeError(f"tables not found in googlenet.h: {missing}")
    for name, val in scal.items():
        src, n = re.subn(r"(CONSTANT\s+int\s+" + name + r"\s*=)[^;]*;", lambda m: f"{m.group(1)} {val};", src, count=1)
        if n != 1:
            raise RuntimeError(f"scalar {name} not found in googlenet.h")
    return src


def build(net) -> str:
    """Returns the directory holding libhost.so / libnet.so for this network (built on demand)."""
    if not os.path.isdir(os.path.join(REF, "Runtime_Engine")):
        raise FileNotFoundError(REF)
    h = hashlib.sha1(json.dumps(net.to_json(), sort_keys=True).encode())
    for src in ("net_harness.c", "fifo_shim.h", "multi_layer.py", "one_layer.py", os.path.join("..", "ref_host_shim.cpp")):
        with open(os.path.join(_HERE, src), "rb") as f:      # a changed harness must not meet a cached library
            h.update(f.read())
    key = h.hexdigest()[:12]
    out = os.path.join(os.path.dirname(_HERE), "_ref", f"net_{key}")
    if os.path.exists(os.path.join(out, "libnet.so")) and os.path.exists(os.path.join(out, "libhost.so")):
        return out
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "one_layer.h"), "w") as f:
        f.write(generate_header(net))
    cnn = os.path.join(REF, "Runtime_Engine", "cnn")
    host, dev, common = os.path.join(cnn, "host"), os.path.join(cnn, "device", "src"), os.path.join(REF, "Runtime_Engine", "common", "inc")
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
                           os.path.join(_HERE, "net_harness.c"), "-o", os.path.join(out, "libnet.so")])
    return out


def libs(net):
    d = build(net)
    return C.CDLL(os.path.join(d, "libhost.so")), C.CDLL(os.path.join(d, "libnet.so"))
