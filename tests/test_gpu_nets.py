"""GPU parity for the other shipped networks — GoogLeNet (BASELINE configs[4]: 1x1 inception branches,
5x5 convs, concat offsets, ipool pseudo layers, ceil-mode pools) and channel-pruned ResNet50 (ragged
channel counts) — whole network vs the CPU oracle, every tensor of image 0."""
import os

import numpy as np
import pytest

from tests import helpers as H
from tests.conftest import GOLDEN
from tf2_b200 import capi, formats, nets, synth

pytestmark = pytest.mark.gpu


def _model(name):
    net = nets.load(name)
    q = formats.parse_q_file(net, os.path.join(GOLDEN, f"{name}_Q"))
    blob = synth.synth_float_blob(net, seed=5, q=q)
    return net, q, formats.load_float_blob(net, blob, q)


@pytest.mark.parametrize("name,variant", [("googlenet", capi.VARIANT_AUTO), ("googlenet", capi.VARIANT_SHIFT),
                                          ("resnet50_pruned", capi.VARIANT_AUTO)],
                         ids=["googlenet-auto", "googlenet-shift", "resnet50_pruned-auto"])
def test_network_matches_oracle(name, variant):
    import torch
    from oracle import oracle as O
    from tf2_b200.network import NetWork, Runner
    net, q, model = _model(name)
    B = 3
    imgs = synth.synth_images(B, seed=17)
    raw, t0 = formats.prepare_input(net, imgs, q)
    nw = NetWork(net, 0)
    nw.InitFromCodes(model, q, max_images=B, variant=variant)
    r = Runner(nw)
    got = r.run_device(torch.from_numpy(raw).cuda(), raw224=True).cpu().numpy()
    exp = O.run_network(net, model, t0)
    assert got.shape == exp.shape
    assert np.array_equal(got, exp), f"{name}: result differs in {(got != exp).sum()} of {exp.size}"
    tens, accs = H.oracle_tensors(net, model, t0[0])
    dev = {}
    for t in range(1, len(net.tensors)):
        g = dev[t] = r.read_tensor(t, B).cpu().numpy()[0]
        assert np.array_equal(g, tens[t]), f"{name}: tensor {t} ({net.tensors[t].name}) differs in {(g != tens[t]).sum()}"
    # ... and directly against what the reference's own device program produced for this model and image
    H.assert_reference_hashes(name, net, dev)
    kinds = set(nw.layer_kernels())
    assert ("mma" in kinds) == (variant != capi.VARIANT_SHIFT)
    nw.CleanUp()


def test_raw224_input_at_an_unaligned_device_address():
    """The space-to-depth kernel stages the raw image with 16-byte loads; a caller's buffer at an odd address takes
    the byte-load path.  Same results, and tensor 0 (27 transformed channels, input_loader.cpp:27-73) equals
    feature_trans of the quantised image either way."""
    import torch
    from tf2_b200.network import NetWork, Runner
    net, q, model = _model("googlenet")
    B = 2
    imgs = synth.synth_images(B, seed=23)
    raw, t0 = formats.prepare_input(net, imgs, q)
    nw = NetWork(net, 0)
    nw.InitFromCodes(model, q, max_images=B, variant=capi.VARIANT_AUTO)
    r = Runner(nw)
    x = torch.from_numpy(raw).cuda()
    a = r.run_device(x, raw224=True).cpu().numpy()
    ta = r.read_tensor(0, B).cpu().numpy()
    assert np.array_equal(ta[:, :t0.shape[1]], t0), "tensor 0 differs from feature_trans"
    flat = torch.empty(x.numel() + 16, dtype=torch.int8, device="cuda")
    for off in (1, 7):
        y = flat[off:off + x.numel()].view(x.shape)
        y.copy_(x)
        assert y.data_ptr() % 16 != 0
        b = r.run_device(y, raw224=True).cpu().numpy()
        assert np.array_equal(a, b), off
        assert np.array_equal(r.read_tensor(0, B).cpu().numpy(), ta), off
    nw.CleanUp()
