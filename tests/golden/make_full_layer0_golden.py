#!/usr/bin/env python
"""Generates tests/golden/full_layer0_golden.json: SHA-256 of layer 0's output as produced by the
REFERENCE's whole device pipeline (cnn.cl compiled as C, oracle/ref_device/full_harness.c) for the
seeded inputs of tests/test_full_layer0.py.  Run in the build container (needs /root/reference)."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.test_full_layer0 import NETS, seeded_case, tiles_to_map  # noqa: E402

out = {}
for name in NETS:
    net, ld, tin, x, codes, params = seeded_case(name)
    data, counts, consts = O.ref_full_layer0(name, x, codes, params)
    ref = tiles_to_map(data, ld.N, ld.PH, ld.PW)
    out[name] = {"sha256": hashlib.sha256(ref.tobytes()).hexdigest(), "counts": [int(c) for c in counts],
                 "consts": [int(c) for c in consts], "shape": list(ref.shape)}
with open(os.path.join(ROOT, "tests", "golden", "full_layer0_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
print(out)
