#!/usr/bin/env python
"""Generates tests/golden/single_layer_golden.json: SHA-256 of the output of the REFERENCE's whole device
pipeline run on one-layer networks (oracle/ref_device/one_layer.py) for the seeded cases of
tests/test_single_layer_ref.py.  Run in the build container (needs /root/reference)."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "ref_device"))
import one_layer as OL  # noqa: E402
from tests.test_single_layer_ref import CASES, seeded  # noqa: E402

out = {}
for case in CASES:
    cfg, ld, tin, x, codes, params = seeded(case)
    ref, counts, consts = OL.run(cfg, x, codes, params)
    key = "c%d_n%d_k%d_p%d_s%d_%dx%d" % case[:7]
    out[key] = {"sha256": hashlib.sha256(ref.tobytes()).hexdigest(), "counts": [int(c) for c in counts[:6]],
                "consts": [int(c) for c in consts], "shape": list(ref.shape)}
    print(key, out[key]["shape"], out[key]["counts"])
with open(os.path.join(ROOT, "tests", "golden", "single_layer_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
