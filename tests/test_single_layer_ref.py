"""The convolution geometry of sequencer.cl / retriever.cl for layers the shipped tables never place
first — 1x1, padded 3x3, stride 2, 5x5, ragged channel counts, the map sizes of the BASELINE configs —
pinned against the reference's whole device pipeline: cnn.cl compiled as C runs each geometry as
"layer 0" of a one-layer network whose table header is generated from the reference's googlenet.h
(oracle/ref_device/one_layer.py; schedule from the reference's own cycle.cl).  Live in the build
container; SHA-256 of the reference outputs committed so the pin travels."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as H
from tf2_b200 import nets

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "ref_device"))
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "single_layer_golden.json")

# C, N, k, pad, stride, IH, IW, relu
CASES = [
    (64, 32, 1, 0, 1, 14, 14, 1), (40, 24, 1, 0, 1, 14, 14, 1), (64, 64, 3, 1, 1, 14, 14, 1),
    (32, 32, 1, 0, 2, 14, 14, 0), (32, 48, 1, 0, 2, 28, 28, 1), (16, 16, 1, 0, 2, 56, 56, 1),
    (64, 64, 3, 1, 2, 28, 28, 1), (32, 32, 3, 1, 2, 14, 14, 1), (16, 32, 3, 1, 2, 56, 56, 1), (32, 32, 3, 1, 2, 17, 17, 1),
    (16, 32, 5, 2, 1, 14, 14, 1), (100, 48, 3, 1, 1, 9, 9, 0), (16, 16, 3, 1, 1, 56, 56, 1), (32, 32, 3, 1, 1, 27, 27, 1),
    (32, 16, 3, 1, 1, 13, 13, 1), (48, 32, 3, 1, 1, 7, 7, 1), (64, 16, 1, 0, 1, 55, 55, 1), (64, 32, 1, 0, 1, 7, 7, 0),
    (3, 16, 3, 1, 1, 56, 56, 1), (27, 64, 3, 0, 1, 30, 30, 1),
]
IDS = ["c%d_n%d_k%d_p%d_s%d_%dx%d" % c[:7] for c in CASES]


def seeded(case):
    C, N, k, pad, s, IH, IW, relu = case
    net = nets.chain((C, IH, IW), [dict(N=N, k=k, pad=pad, stride=s, relu=relu)])
    ld, tin = net.layers[0], net.tensors[0]
    rng = np.random.default_rng(1000 * C + 10 * N + k + pad + s + IH)
    x = H.random_input(rng, C, IH, IW, nonneg=False)
    codes = H.random_codes(rng, N, C, k)
    params = H.fit_params(rng, ld, tin, x, codes)
    cfg = dict(C=C, N=N, k=k, pad=pad, stride=s, IH=IH, IW=IW, relu=relu)
    return cfg, ld, tin, x, codes, params


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_vs_golden_hash(case):
    with open(GOLDEN) as f:
        g = json.load(f)
    cfg, ld, tin, x, codes, params = seeded(case)
    out = O.layer_forward(ld, tin, x, codes, params)
    key = "c%d_n%d_k%d_p%d_s%d_%dx%d" % case[:7]
    assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == g[key]["sha256"]
    assert g[key]["counts"][1] == g[key]["consts"][0] and g[key]["counts"][0] == g[key]["consts"][2]


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_vs_compiled_reference_pipeline(case, reference_dir):
    import one_layer as OL
    cfg, ld, tin, x, codes, params = seeded(case)
    ref, counts, consts = OL.run(cfg, x, codes, params)
    # the item counts of the run equal what the reference's cycle.cl derives for this layer
    assert counts[0] == consts[2] and counts[1] == consts[0] and counts[5] == consts[3]
    out = O.layer_forward(ld, tin, x, codes, params)
    assert np.array_equal(out, ref), f"{(out != ref).sum()} of {ref.size} outputs differ from the reference pipeline"


def test_derived_table_formulas_hold_on_every_shipped_layer(reference_dir):
    """the formulas one_layer.layer_tables() uses for the derived tables (kWvecEnd, kFilterLoadSize,
    kOhEndWithOffset, ...) reproduce every entry of the three shipped headers"""
    import one_layer as OL
    from tf2_b200.header_tables import parse_header_file
    for name in ("resnet50", "googlenet", "resnet50_pruned"):
        t = parse_header_file(os.path.join(reference_dir, "Runtime_Engine", "cnn", "host", "inc", name + ".h"))
        L = int(t["NUM_LAYER"])
        for l in range(L):
            if t["kIpoolEnable"][l]:
                continue
            cfg = dict(C=t["kInputChannels"][l], N=t["kOutputChannels"][l], k=t["kFilterSize"][l], pad=t["kPadWidth"][l],
                       stride=t["kConvStride"][l], IH=t["kInputHeight"][l], IW=t["kInputWidth"][l])
            d = OL.layer_tables(cfg)
            for key in ("kWvecEnd", "kFWvecEnd", "kCvecEnd", "kFilterCvecEnd", "kFilterLoadSize", "kNvecEnd", "kOhEndWithOffset",
                        "kOwEndWithOffset", "kOutputWidth", "kOutputHeight", "kNEndWithOffset"):
                assert d[key] == t[key][l], f"{name} layer {l} {key}: formula {d[key]} header {t[key][l]}"
            if not t["kPoolEnable"][l]:
                assert d["kPoolOutputWidth"] == t["kPoolOutputWidth"][l] and d["kPoolOutputWvecEnd"] == t["kPoolOutputWvecEnd"][l]
