#!/usr/bin/env python
"""Generates tests/golden/post_golden.json: SHA-256 of every layer's output as produced by the
REFERENCE's post-PE kernels (relu.cl, pool.cl, pool_tail.cl, feature_writer.cl, full_size_pool.cl
compiled as C by oracle/build_ref.sh) for seeded random PE-output maps, for the three shipped
networks.  Run in the build container (needs /root/reference); the JSON is what travels."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.test_post_golden import seeded_maps  # noqa: E402

out = {}
for net in ("resnet50", "googlenet", "resnet50_pruned"):
    R = O.RefPost(net)
    maps = seeded_maps(R.n_layers, [R.pe_shape(l) for l in range(R.n_layers)])
    outs, counts = R.run(maps)
    out[net] = {"shapes": [list(R.pe_shape(l)) for l in range(R.n_layers)],
                "counts": [int(c) for c in counts[:3]], "consts": [int(c) for c in R.consts],
                "sha256": [hashlib.sha256(np.ascontiguousarray(o).tobytes()).hexdigest() for o in outs]}
with open(os.path.join(ROOT, "tests", "golden", "post_golden.json"), "w") as f:
    json.dump(out, f, indent=0)
print({k: len(v["sha256"]) for k, v in out.items()})
