"""fpganetwork.bin -> layer tables (tf2_b200/fpganet.py), the replacement of the reference's
TF2_auto_config front half.  The reference's own example file
(Runtime_Engine/TF2_auto_config/examples/resnet50/fpganetwork.bin, kept as a fixture under
tests/golden/) must compile to exactly the tables of the shipped resnet50.h."""
import os
from dataclasses import asdict

import numpy as np
import pytest

from tests import helpers as H
from tests.conftest import GOLDEN
from tf2_b200 import fpganet as F
from tf2_b200 import nets

FIXTURE = os.path.join(GOLDEN, "resnet50_fpganetwork.bin")


def test_parse_reference_example_byte_exact():
    data = open(FIXTURE, "rb").read()
    fn = F.parse_fpganetwork(data)
    assert len(data) == 37705 and len(fn.layers) == 73 and fn.version == F.VERSION
    kinds = [tuple(F.OP_NAMES[o.type] for o in L.ops) for L in fn.layers]
    assert kinds.count(("conv", "bn", "scale", "relu")) + kinds.count(("conv", "bn", "scale")) == 53
    assert kinds.count(("eltwise", "relu")) == 16 and kinds[-2:] == [("fc",), ("softmax",)]
    # writer . parser is the identity on everything but the (garbage) pointer fields
    again = F.parse_fpganetwork(F.write_fpganetwork(fn))
    assert again.layers == fn.layers and len(F.write_fpganetwork(fn)) == len(data)


def test_resnet50_example_compiles_to_the_shipped_header_tables():
    got = F.load_fpganetwork(FIXTURE, "resnet50", max_pool_pad=1)   # the reference forces kPoolPad=1 by net name
    want = nets.resnet50()                                           # parsed from resnet50.h
    assert len(got.layers) == len(want.layers) == 54 and len(got.tensors) == len(want.tensors)
    for a, b in zip(got.layers, want.layers):
        assert asdict(a) == asdict(b), a.name
    for a, b in zip(got.tensors, want.tensors):
        assert asdict(a) == asdict(b)
    assert got.num_q_rows == want.num_q_rows and got.max_out_channel == want.max_out_channel
    assert got.macs_per_image() == want.macs_per_image() == 4089184256
    # what the Caffe file itself says (ceil-mode pool, pad 0) differs only in that one table entry
    raw = F.load_fpganetwork(FIXTURE, "resnet50")
    assert raw.layers[0].pool_pad == 0 and all(asdict(x) == asdict(y) for x, y in zip(raw.layers[1:], want.layers[1:]))


def test_reference_example_against_live_reference_tree(reference_dir):
    p = os.path.join(reference_dir, "Runtime_Engine", "TF2_auto_config", "examples", "resnet50", "fpganetwork.bin")
    assert open(p, "rb").read() == open(FIXTURE, "rb").read()


def _blob(t, p=None):
    return F.FpgaOp(t, p or {})


def _conv(oc, k=1, pad=0, s=1, bias=1):
    return _blob(F.OP_CONV, dict(out_c=oc, bias=bias, pad_l=pad, pad_t=pad, pad_r=pad, pad_b=pad, kh=k, kw=k, sh=s, sw=s, dilation=1))


def inception_fpganet():
    """stem conv + max pool, then one GoogLeNet-style inception block (1x1 | 1x1->3x3 | 1x1->5x5 |
    pool->1x1, concat), a 7x7 average and a classifier — as caffe2fpga would dump it."""
    S = lambda c, h: (1, c, h, h)
    relu = _blob(F.OP_RELU)
    mp = lambda s, pad: _blob(F.OP_POOL, dict(method=0, pad_l=pad, pad_t=pad, pad_r=pad, pad_b=pad, kh=3, kw=3, sh=s, sw=s, global_pool=0))
    Ls = [
        F.FpgaLayer(0, [_conv(32, 3, 1), relu], [-1], [0], [S(8, 14)], [S(32, 14)]),
        F.FpgaLayer(1, [mp(2, 0)], [0], [1], [S(32, 14)], [S(32, 7)]),
        F.FpgaLayer(2, [_conv(16), relu], [1], [2], [S(32, 7)], [S(16, 7)]),
        F.FpgaLayer(3, [_conv(16), relu], [1], [3], [S(32, 7)], [S(16, 7)]),
        F.FpgaLayer(4, [_conv(32, 3, 1), relu], [3], [4], [S(16, 7)], [S(32, 7)]),
        F.FpgaLayer(5, [_conv(16), relu], [1], [5], [S(32, 7)], [S(16, 7)]),
        F.FpgaLayer(6, [_conv(16, 5, 2), relu], [5], [6], [S(16, 7)], [S(16, 7)]),
        F.FpgaLayer(7, [mp(1, 1)], [1], [7], [S(32, 7)], [S(32, 7)]),
        F.FpgaLayer(8, [_conv(16), relu], [7], [8], [S(32, 7)], [S(16, 7)]),
        F.FpgaLayer(9, [_blob(F.OP_CONCAT)], [2, 4, 6, 8], [9], [S(16, 7), S(32, 7), S(16, 7), S(16, 7)], [S(80, 7)]),
        F.FpgaLayer(10, [_conv(48), relu], [9], [10], [S(80, 7)], [S(48, 7)]),
        F.FpgaLayer(11, [_blob(F.OP_POOL, dict(method=1, pad_l=0, pad_t=0, pad_r=0, pad_b=0, kh=7, kw=7, sh=1, sw=1, global_pool=0))],
                    [10], [11], [S(48, 7)], [S(48, 1)]),
        F.FpgaLayer(12, [_blob(F.OP_FC, dict(out_c=10, bias=1))], [11], [12], [S(48, 1)], [S(10, 1)]),
        F.FpgaLayer(13, [_blob(F.OP_SOFTMAX)], [12], [13], [S(10, 1)], [S(10, 1)]),
    ]
    return F.FpgaNet(F.VERSION, "mini_inception", Ls)


def test_inception_block_compiles_to_concat_offsets_and_ipool():
    net = F.to_netdesc(F.parse_fpganetwork(F.write_fpganetwork(inception_fpganet())))
    L = net.layers
    assert [l.k for l in L] == [3, 1, 1, 3, 1, 5, 3, 1, 1, 1]
    assert (L[0].pool, L[0].pool_stride, L[0].pool_pad, L[0].PH) == (1, 2, 0, 7)        # max pool fused into the stem
    cat = L[1].out_tensor
    assert net.tensors[cat].C == 80 and [(l.out_tensor, l.out_ch0) for l in (L[1], L[3], L[5], L[7])] == \
        [(cat, 0), (cat, 16), (cat, 48), (cat, 64)]
    assert L[6].ipool == 1 and L[6].in_tensor == L[0].out_tensor and L[7].in_tensor == L[6].out_tensor
    assert L[8].in_tensor == cat and L[8].gap == 1 and L[9].C == 48 and L[9].N == 10 and L[9].bias_en == 1
    assert net.branch_tail == [0, 1, 0, 1, 0, 1, 0, 1, 0, 0] and net.num_q_rows == 10 + 1 + 1
    # the compiled description is executable: the CPU oracle runs it end to end
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    x = H.random_input(rng, 8, 14, 14, nonneg=False, B=2)
    model = H.random_model(net, rng, x)
    y = O.run_network(net, model, x)
    assert y.shape == (2, 10, 1, 1) and y.std() > 0


@pytest.mark.parametrize("breaker,msg", [
    (lambda n: n.layers[1].ops[0].p.update(kh=2, kw=2), "3x3 max"),
    (lambda n: n.layers[11].ops[0].p.update(kh=5, kw=5), "7x7 global average"),
    (lambda n: n.layers[0].ops[0].p.update(pad_l=2), "symmetric"),
])
def test_unsupported_networks_are_rejected_not_mis_compiled(breaker, msg):
    fn = inception_fpganet()
    breaker(fn)
    with pytest.raises(ValueError, match=msg):
        F.to_netdesc(fn)


@pytest.mark.gpu
def test_compiled_inception_runs_bit_exact_on_gpu():
    import torch
    from oracle import oracle as O
    from tf2_b200 import capi
    from tf2_b200.network import NetWork, Runner
    net = F.to_netdesc(inception_fpganet())
    rng = np.random.default_rng(9)
    x = H.random_input(rng, 8, 14, 14, nonneg=False, B=3)
    model = H.random_model(net, rng, x)
    exp = O.run_network(net, model, x)
    for variant in (capi.VARIANT_SHIFT, capi.VARIANT_AUTO):
        nw = NetWork(net, 0)
        nw.InitFromCodes(model, None, max_images=3, variant=variant)
        got = Runner(nw).run_device(torch.from_numpy(x).cuda()).cpu().numpy()
        assert np.array_equal(got, exp)
        nw.CleanUp()
