#!/usr/bin/env python
"""Generates tests/golden/loader_golden.json from the REFERENCE's own LoadModel (compiled unmodified
by oracle/build_ref.sh into oracle/_ref/): sha256 of the shift codes and BiasBnParam of every layer
for a seeded synthetic float blob + the shipped Q files.  Run in the build container."""
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tf2_b200 import formats, nets, synth  # noqa: E402

out = {}
for name, seed in (("resnet50", 3), ("googlenet", 3)):
    L = O.ref_host_lib(name)
    assert L is not None, "run oracle/build_ref.sh first"
    net = nets.load(name)
    qfile = os.path.join(ROOT, "tests", "golden", f"{name}_Q")
    nq, mo = L.ref_num_q_layers(), L.ref_max_out_channel()
    qref = np.zeros((nq + 2) * mo, dtype=np.int8)
    L.ref_quantization(qref.ctypes.data, qfile.encode())
    q = formats.parse_q_file(net, qfile)
    blob = synth.synth_float_blob(net, seed=seed, q=q)
    with tempfile.NamedTemporaryFile(suffix=".bin") as tf:
        tf.write(blob); tf.flush()
        stride, nl = L.ref_filter_layer_stride(), net.num_layers
        fr = np.zeros(nl * stride + 1024, np.uint8)
        bb = np.zeros((nl * L.ref_max_bias_size() + 16, 3), np.int32)
        L.ref_load_model(tf.name.encode(), fr.ctypes.data, bb.ctypes.data, qref.ctypes.data)
    rec = {"seed": seed, "blob_sha256": hashlib.sha256(blob).hexdigest(), "codes": {}, "params": {}}
    for l, ld in enumerate(net.layers):
        if ld.ipool:
            continue
        n = ld.N * ld.C * ld.k * ld.k
        rec["codes"][str(l)] = hashlib.sha256(fr[l * stride:l * stride + n].tobytes()).hexdigest()
        rec["params"][str(l)] = hashlib.sha256(
            np.ascontiguousarray(bb[l * L.ref_max_bias_size():l * L.ref_max_bias_size() + ld.N]).tobytes()).hexdigest()
    out[name] = rec
with open(os.path.join(ROOT, "tests", "golden", "loader_golden.json"), "w") as f:
    json.dump(out, f, indent=0)
print("wrote loader_golden.json")
