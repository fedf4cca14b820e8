#!/usr/bin/env python
"""Regenerates tf2_b200/nets/*.json (network descriptions) from the reference's generated headers
and copies the shipped Q files used as fixtures into tests/golden/.  Run in the build container
(needs /root/reference); the outputs are committed so nothing reads the reference at run time."""
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tf2_b200.netdesc import NetDesc  # noqa: E402

REF = os.environ.get("TF2_REFERENCE", "/root/reference")
INC = os.path.join(REF, "Runtime_Engine/cnn/host/inc")
MODEL = os.path.join(REF, "Runtime_Engine/cnn/host/model")

for name in ("resnet50", "googlenet", "resnet50_pruned"):
    net = NetDesc.from_header(os.path.join(INC, name + ".h"), name)
    with open(os.path.join(ROOT, "tf2_b200", "nets", name + ".json"), "w") as f:
        json.dump(net.to_json(), f, separators=(",", ":"))
    print(name, net.num_layers, "layers", len(net.tensors), "tensors", net.macs_per_image(), "MAC/img")
for q in ("resnet50_Q", "googlenet_Q", "resnet50_pruned_Q"):
    shutil.copy(os.path.join(MODEL, q), os.path.join(ROOT, "tests", "golden", q))
