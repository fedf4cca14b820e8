"""Generates tests/golden/whole_net_golden.json: SHA-256 of what the REFERENCE's own device program (cnn.cl
compiled as C, all kernels running as coroutines: oracle/ref_device/net_harness.c) sends to the on-chip cache
for every layer, and of the final map it leaves in feature_ddr, for the seeded whole-network cases of
tests/test_whole_net_ref.py.  Build container only (needs /root/reference + oracle/build_ref.sh):
    python tests/golden/make_whole_net_golden.py"""
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import oracle as O  # noqa: E402
from tests.test_whole_net_ref import CASES, build_case  # noqa: E402

out = {}
for case in CASES:
    name, net, model, x = build_case(case)
    t = time.time()
    per, final, st = O.ref_run_network(name, x, model)
    assert st["parked"] == 0 and st["fifo_bytes_left"] == 0 and st["tap_dropped"] == 0, st
    out[case] = {
        "layers": [None if p is None else hashlib.sha256(np.ascontiguousarray(p).tobytes()).hexdigest() for p in per],
        "final": hashlib.sha256(np.ascontiguousarray(final).tobytes()).hexdigest(),
        "final_std": float(final.std()),
        "kernels_finished": st["done"], "tiles": st["tap_counts"], "seconds": round(time.time() - t, 1),
    }
    print(case, out[case]["seconds"], "s", flush=True)
from tests.test_whole_net_ref import _frame_hashes, frames_case  # noqa: E402
net, model, t0 = frames_case()
t = time.time()
finals, st = O.ref_run_frames("googlenet", t0, model)
assert st["parked"] == 0 and st["fifo_bytes_left"] == 0, st
out["googlenet_frames"] = {"finals": _frame_hashes(finals), "kernels_finished": st["done"], "tiles": st["tap_counts"],
                           "seconds": round(time.time() - t, 1)}
print("googlenet_frames", out["googlenet_frames"]["seconds"], "s", flush=True)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "whole_net_golden.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print("wrote", path)
