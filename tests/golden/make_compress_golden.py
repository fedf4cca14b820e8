#!/usr/bin/env python
"""Generates tests/golden/compress_golden.json from the REFERENCE's own ComputeQuantumRange /
ShapeIntoTwoPower (TransForm_Kit/Compression/compress_net/core/compress_core.py), executed from the source
where it lies on the seeded inputs of tests/test_compress.py.  Build container only; the JSON travels."""
import contextlib
import hashlib
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("TF2_REFERENCE", "/root/reference")
ns = {}
exec(compile(open(os.path.join(REF, "TransForm_Kit", "Compression", "compress_net", "core", "compress_core.py")).read(),
             "compress_core.py", "exec"), ns)
sys.path.insert(0, ROOT)
from tests.test_compress import golden_inputs  # noqa: E402

out = []
for w, portions in golden_inputs():
    w = w.astype(np.float64).copy()
    mask = np.ones(w.shape, np.int8)
    steps = []
    prev = 0.0
    for cur in portions:
        with contextlib.redirect_stdout(io.StringIO()):
            mx, mn = ns["ComputeQuantumRange"](w, mask, 7)
            w, mask = ns["ShapeIntoTwoPower"](w, mask, prev, cur, mx, mn)
        steps.append({"max_exp": int(mx), "min_exp": int(mn), "n_quantized": int((mask == 0).sum()),
                      "sha256": hashlib.sha256(np.ascontiguousarray(w, dtype=np.float64).tobytes()).hexdigest(),
                      "mask_sha256": hashlib.sha256(np.ascontiguousarray(mask, dtype=np.int8).tobytes()).hexdigest()})
        prev = cur
    out.append(steps)
with open(os.path.join(ROOT, "tests", "golden", "compress_golden.json"), "w") as f:
    json.dump(out, f, indent=1)
print(len(out), "cases")
