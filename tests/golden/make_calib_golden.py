#!/usr/bin/env python
"""Generates tests/golden/calib_golden.json from the REFERENCE's own QuantizeForShift / QuantizeChannel
(TransForm_Kit/Quantization/quantization.py:33-69): the two function definitions are compiled out of the
reference source where it lies (the module itself cannot be imported: it drags in Caffe-era loaders) and
run on seeded inputs.  Run in the build container; the JSON travels."""
import ast
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("TF2_REFERENCE", "/root/reference")
src = open(os.path.join(REF, "TransForm_Kit", "Quantization", "quantization.py")).read()
tree = ast.parse(src)
keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("QuantizeForShift", "QuantizeChannel")]
ns = {"np": np, "math": math}
exec(compile(ast.Module(body=keep, type_ignores=[]), "quantization.py", "exec"), ns)

sys.path.insert(0, ROOT)
from tests.test_calibrate import golden_inputs  # noqa: E402

out = {"for_shift": [], "channel": []}
for x in golden_inputs()["for_shift"]:
    out["for_shift"].append(int(ns["QuantizeForShift"](x)))
for x in golden_inputs()["channel"]:
    out["channel"].append([int(v) for v in ns["QuantizeChannel"]("shift", x)])
with open(os.path.join(ROOT, "tests", "golden", "calib_golden.json"), "w") as f:
    json.dump(out, f)
print({k: len(v) for k, v in out.items()})
