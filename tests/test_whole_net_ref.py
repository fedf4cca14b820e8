"""Whole networks through the REFERENCE's own device program.

cnn.cl compiled as C with all of its kernels running concurrently as coroutines
(oracle/ref_device/net_harness.c: input_reader, filter_reader, sequencer, retriever, 16 PEs, relu, pool,
full_size_pool, pool_tail, feature_writer) executes every layer of ResNet50, GoogLeNet and pruned ResNet50
for one image — including what the one-layer pins cannot reach: the feedback of each layer's output into
the on-chip feature cache through the retriever's non-blocking reads (retriever.cl:328-329), the ipool
pseudo layers fed from that cache (retriever.cl:285-302), residual operands through the DDR ping-pong and
concat offsets across layers.  The oracle must equal it on every layer.

The cases are the very models and image the GPU parity tests run (tests/helpers.py: synth_case), so
"CUDA == oracle" (pytest -m gpu) and "oracle == compiled reference" (here) meet on the same tensors.
Live where oracle/_ref is built (~80 s); the SHA-256 of the reference's outputs is committed
(tests/golden/whole_net_golden.json, tests/golden/make_whole_net_golden.py) so the pin travels."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as H
from tf2_b200 import nets

CASES = ["resnet50", "googlenet", "resnet50_pruned", "googlenet_dense"]


def build_case(case):
    """-> (network name, net, model, tensor 0 of the image)"""
    if case in H.SYNTH_CASES:
        net, q, model, x = H.synth_case(case)
        return case, net, model, x
    # dense random codes, signed input with -128s, every layer re-fitted to stay alive (1x1 maps included)
    name = case.split("_")[0]
    net = nets.load(name)
    t0 = net.tensors[0]
    rng = np.random.default_rng(11)
    x = H.random_input(rng, t0.C, t0.H, t0.W, nonneg=False)
    model = H.random_model(net, rng, x[None])
    codes, params = model[-1]                       # fit_params sees one sample per channel on a 1x1 map:
    _, accs = H.oracle_tensors(net, model, x)       # scale the classifier by the spread ACROSS channels instead
    params = params.copy()
    params[:, 1] = int(40.0 * 2.0 ** 35 / (accs[len(model) - 1].astype(np.float64).std() + 1.0))
    model[-1] = (codes, params)
    return name, net, model, x


def _have_ref(name):
    return O.ref_host_lib(name) is not None and os.path.exists(os.path.join(os.path.dirname(O.__file__), "_ref", f"libtf2ref_net_{name}.so"))


@pytest.mark.parametrize("case", CASES)
def test_oracle_equals_reference_hashes(case):
    name, net, model, x = build_case(case)
    tens, _ = H.oracle_tensors(net, model, x)
    g = H.assert_reference_hashes(case, net, tens)
    assert g["final_std"] > 3                                   # the final map, read back from feature_ddr; alive
    assert g["kernels_finished"] == 25                         # every reference kernel ran to its last cycle


@pytest.mark.parametrize("case", CASES)
def test_oracle_equals_reference_live(case):
    name, net, model, x = build_case(case)
    if not _have_ref(name):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    per, final, st = O.ref_run_network(name, x, model)
    assert st["done"] == 25 and st["parked"] == 0 and st["fifo_bytes_left"] == 0 and st["tap_dropped"] == 0
    assert st["tap_used"] == st["tap_counts"]                  # every tile the feature writer emitted is accounted for
    tens, _ = H.oracle_tensors(net, model, x)
    for l, ld in enumerate(net.layers):
        if per[l] is None:
            continue
        want = tens[ld.out_tensor][ld.out_ch0:ld.out_ch0 + ld.N]
        assert per[l].shape == want.shape, (l, per[l].shape, want.shape)
        assert np.array_equal(per[l], want), f"{case}: layer {l}: {(per[l] != want).sum()} of {want.size} differ"
    ld = net.layers[-1]
    want = tens[ld.out_tensor][ld.out_ch0:ld.out_ch0 + ld.N]
    assert np.array_equal(final, want.reshape(final.shape))
    assert sum(p is None for p in per) <= 5


# ---- images back to back: the premise of sharding a batch over GPUs -------------------------------------
def frames_case():
    """GoogLeNet (synthetic model of the GPU tests) on three different images."""
    from tf2_b200 import formats, synth
    net, q, model, _ = H.synth_case("googlenet")
    _, t0 = formats.prepare_input(net, synth.synth_images(3, seed=23), q)
    return net, model, t0


def _frame_hashes(maps):
    import hashlib
    return [hashlib.sha256(np.ascontiguousarray(m, dtype=np.int8).tobytes()).hexdigest() for m in maps]


def test_frames_back_to_back_hashes():
    """The reference runs num_images frames through ONE invocation of its kernels (runner.cpp:61-176: frame_num),
    re-using the on-chip cache, the DDR pages and the filter double buffer from frame to frame.  Every frame's
    result equals the oracle's result for that image alone: images are independent, which is what lets a batch
    be sharded over GPUs with no data-path collective."""
    import json
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "whole_net_golden.json")) as f:
        g = json.load(f)["googlenet_frames"]
    net, model, t0 = frames_case()
    exp = O.run_network(net, model, t0)
    assert _frame_hashes(exp) == g["finals"]
    assert len(set(g["finals"])) == 3 and g["kernels_finished"] == 25


def test_frames_back_to_back_live():
    if not _have_ref("googlenet"):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    net, model, t0 = frames_case()
    finals, st = O.ref_run_frames("googlenet", t0, model)
    assert st["done"] == 25 and st["parked"] == 0 and st["fifo_bytes_left"] == 0
    assert st["tap_counts"] == [3 * c for c in st["tap_per_frame"]]
    exp = O.run_network(net, model, t0)
    for f in range(3):
        assert np.array_equal(finals[f].reshape(-1), exp[f].reshape(-1)), f"frame {f}"
    assert exp.std() > 3 and not np.array_equal(exp[0], exp[1])
