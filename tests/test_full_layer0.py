"""Step 1 of SURVEY.md Appendix A (the convolution geometry of sequencer.cl / retriever.cl chained to
the real PE arithmetic and post-PE kernels) pinned against the reference's whole device pipeline:
cnn.cl compiled as C (oracle/ref_device/full_harness.c) runs layer 0 of each shipped network — 3x3 over
the 27-channel 114x114 transformed image, 64 outputs, ReLU, 3x3/stride-2 max pool — from device
buffers laid out by the reference's own InputConvert / FilterConvert.  Live in the build container;
the SHA-256 of the reference outputs for the seeded inputs is committed so the pin travels."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as H
from tf2_b200 import nets

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "full_layer0_golden.json")
NETS = ["resnet50", "googlenet", "resnet50_pruned"]


def seeded_case(name):
    net = nets.load(name)
    ld, tin = net.layers[0], net.tensors[0]
    rng = np.random.default_rng(20190510)
    x = H.random_input(rng, tin.C, tin.H, tin.W, nonneg=False)       # includes -128 (pe.cl:32-34 quirk)
    codes = H.random_codes(rng, ld.N, ld.C, ld.k)
    codes[rng.random(codes.shape) < 0.05] = 0x00                     # the code-0 taps of the transformed conv1
    params = H.fit_params(rng, ld, tin, x, codes)
    return net, ld, tin, x, codes, params


def tiles_to_map(data, N, PH, PW):
    nvec, pwv = -(-N // 16), -(-PW // 7)
    t = data.reshape(nvec, PH, pwv, 8, 16)[:, :, :, :7, :]
    return np.ascontiguousarray(t.transpose(0, 4, 1, 2, 3).reshape(nvec * 16, PH, pwv * 7)[:N, :, :PW])


@pytest.mark.parametrize("name", NETS)
def test_oracle_layer0_vs_golden_hash(name):
    with open(GOLDEN) as f:
        g = json.load(f)[name]
    net, ld, tin, x, codes, params = seeded_case(name)
    out = O.layer_forward(ld, tin, x, codes, params)
    assert hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() == g["sha256"]
    assert g["counts"][1] == g["consts"][0]          # sequencer items == CONV_CYCLE(0) of the reference header


@pytest.mark.parametrize("name", NETS)
def test_oracle_layer0_vs_compiled_reference_pipeline(name):
    net, ld, tin, x, codes, params = seeded_case(name)
    try:
        data, counts, consts = O.ref_full_layer0(name, x, codes, params)
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    assert counts[0] == consts[2] and counts[1] == consts[0] and counts[4] == consts[4] and counts[5] == consts[3]
    ref = tiles_to_map(data, ld.N, ld.PH, ld.PW)
    out = O.layer_forward(ld, tin, x, codes, params)
    assert np.array_equal(out, ref), f"{name}: {(out != ref).sum()} of {ref.size} outputs differ from the reference pipeline"
    assert ref.std() > 5                              # a live map, not a saturated one
