"""Steps 3-6 of SURVEY.md Appendix A (ReLU, 3x3 max pool + alignment, conv stride in W, residual add,
concat offset, global average) pinned against the reference's own device kernels.

The reference's relu.cl / pool.cl / pool_tail.cl / feature_writer.cl / full_size_pool.cl are compiled
as C (oracle/ref_device/post_harness.c) and run over the layer tables of the three shipped networks
on seeded random PE-output maps.  `test_oracle_vs_compiled_reference` compares the CPU oracle with
that live (build container only); `test_oracle_vs_golden_hashes` compares it with the committed
SHA-256 of the reference outputs (tests/golden/post_golden.json, made by make_post_golden.py), so the
pin travels to boxes without /root/reference."""
import dataclasses
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tf2_b200 import nets
from tf2_b200.netdesc import TensorDesc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "post_golden.json")
NETS = ["resnet50", "googlenet", "resnet50_pruned"]


def seeded_maps(n_layers, shapes):
    rng = np.random.default_rng(20190510)
    return [rng.integers(-128, 128, shapes[l], dtype=np.int8) for l in range(n_layers)]


def pe_shapes_from_netdesc(net):
    """[channels][rows after the conv stride][columns at stride 1], from OUR parsed layer tables"""
    shp = []
    for ld in net.layers:
        ti = net.tensors[ld.in_tensor]
        if ld.ipool:
            shp.append((ld.N, ti.H, ti.W))
        else:
            w1 = ti.W + 2 * ld.pad - ld.k + 1
            shp.append((ld.N, ld.OH, w1))
    return shp


def oracle_post(net, maps):
    """Runs the oracle's steps 3-6 on given PE outputs: every conv layer becomes an identity 1x1
    convolution (code = shift 15 on the diagonal, alpha = 2^20, beta = bias = 0 reproduces its input
    exactly), strided like the original, followed by the layer's real relu/pool/add/gap flags."""
    tens = {}
    outs = []
    for l, ld in enumerate(net.layers):
        y1 = maps[l]
        nch = ld.N
        if ld.ipool:
            tin = TensorDesc(nch, y1.shape[1], y1.shape[2])
            out = O.layer_forward(ld, tin, y1, None, None)
        else:
            s = ld.stride
            ih, iw = (ld.OH - 1) * s + 1, (ld.OW - 1) * s + 1
            x = np.zeros((nch, ih, iw), np.int8)
            x[:, ::s, :] = y1[:, :, :iw]          # stride-2 layers keep stride-1 column 2*ow (pool_tail.cl:128-139)
            codes = np.full((nch, nch, 1, 1), 0x40, np.uint8)
            codes[np.arange(nch), np.arange(nch), 0, 0] = 15
            params = np.zeros((nch, 3), np.int32)
            params[:, 1] = 1 << 20
            ld1 = dataclasses.replace(ld, C=nch, k=1, pad=0)
            tin = TensorDesc(nch, ih, iw)
            res = tens[ld.add_tensor][:nch] if ld.add_tensor >= 0 else None
            out = O.layer_forward(ld1, tin, x, codes, params, R=res)
        outs.append(out)
        t = net.tensors[ld.out_tensor]
        if ld.out_tensor not in tens:
            tens[ld.out_tensor] = np.zeros((t.C, t.H, t.W), np.int8)
        tens[ld.out_tensor][ld.out_ch0:ld.out_ch0 + nch] = out.reshape(nch, t.H, t.W)
    return outs


@pytest.mark.parametrize("name", NETS)
def test_oracle_vs_golden_hashes(name):
    with open(GOLDEN) as f:
        g = json.load(f)[name]
    net = nets.load(name)
    shapes = pe_shapes_from_netdesc(net)
    assert [list(s) for s in shapes] == g["shapes"], "layer tables disagree with the reference header"
    assert g["counts"] == g["consts"][:3]      # item counts == the reference's own cycle constants
    outs = oracle_post(net, seeded_maps(net.num_layers, shapes))
    for l, o in enumerate(outs):
        assert hashlib.sha256(np.ascontiguousarray(o).tobytes()).hexdigest() == g["sha256"][l], f"{name} layer {l}"


@pytest.mark.parametrize("name", NETS)
def test_oracle_vs_compiled_reference(name):
    try:
        R = O.RefPost(name)
    except FileNotFoundError:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    net = nets.load(name)
    shapes = pe_shapes_from_netdesc(net)
    assert shapes == [R.pe_shape(l) for l in range(R.n_layers)]
    rng = np.random.default_rng(7)
    maps = [rng.integers(-128, 128, s, dtype=np.int8) for s in shapes]
    ref, counts = R.run(maps)
    assert list(counts[:3]) == R.consts[:3]
    got = oracle_post(net, maps)
    for l in range(net.num_layers):
        assert np.array_equal(got[l].reshape(ref[l].shape), ref[l]), f"{name} layer {l} differs from the reference kernels"
