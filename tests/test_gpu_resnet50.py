"""GPU parity for the whole ResNet50 INT4/INT8 network (BASELINE configs[1]/[2]) vs the CPU oracle,
plus size-independent properties at the full batch size."""
import os

import numpy as np
import pytest

from tests import helpers as H
from tests.conftest import GOLDEN
from tf2_b200 import capi, formats, nets, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def resnet_model():
    net = nets.resnet50()
    q = formats.parse_q_file(net, os.path.join(GOLDEN, "resnet50_Q"))
    blob = synth.synth_float_blob(net, seed=3, q=q)
    model = formats.load_float_blob(net, blob, q)
    return net, q, model


@pytest.mark.parametrize("variant", [capi.VARIANT_SHIFT, capi.VARIANT_AUTO], ids=["shift", "auto"])
def test_resnet50_matches_oracle(resnet_model, variant):
    import torch
    from oracle import oracle as O
    from tf2_b200.network import NetWork, Runner
    net, q, model = resnet_model
    B = 2
    imgs = synth.synth_images(B, seed=11)
    raw, t0 = formats.prepare_input(net, imgs, q)
    nw = NetWork(net, 0)
    nw.InitFromCodes(model, q, max_images=B, variant=variant)
    r = Runner(nw)
    got = r.run_device(torch.from_numpy(raw).cuda(), raw224=True).cpu().numpy()
    exp = O.run_network(net, model, t0)
    assert got.shape == exp.shape == (B, 1000, 1, 1)
    assert np.array_equal(got, exp), f"fc1000 int8 differs in {(got != exp).sum()} of {exp.size}"
    # the device-side space-to-depth transform equals feature_trans
    t0_dev = r.read_tensor(0, B).cpu().numpy()
    assert np.array_equal(t0_dev, t0)
    # every intermediate feature map and every INT32 accumulator of image 0
    tens, accs = H.oracle_tensors(net, model, t0[0])
    dev = {}
    for t in range(1, len(net.tensors)):
        g = dev[t] = r.read_tensor(t, B).cpu().numpy()[0]
        assert np.array_equal(g, tens[t]), f"tensor {t} ({net.tensors[t].name}) differs in {(g != tens[t]).sum()}"
    # ... and directly against what the reference's own device program produced for this model and image
    H.assert_reference_hashes("resnet50", net, dev)
    for l in (0, 1, 3, 4, 13, 26, 45, 52, 53):
        g = r.dump_acc(l, B).cpu().numpy()[0]
        assert np.array_equal(g, accs[l]), f"layer {l} accumulators differ"
    # the reference-facing host call gives the same answer (Runner::Run, float images in)
    got_host = r.Run(imgs)
    assert np.array_equal(got_host, exp)
    assert nw.last_launches() >= net.num_layers
    nw.CleanUp()


def test_resnet50_batch256_properties(resnet_model):
    """Full BASELINE batch: results are independent of batch composition and order."""
    import torch
    from tf2_b200.network import NetWork, Runner
    net, q, model = resnet_model
    B = 256
    rng = np.random.default_rng(0)
    raw = rng.integers(-128, 128, size=(B, 3, 224, 224), dtype=np.int8)
    raw[1] = raw[0]                      # duplicate images must give duplicate outputs
    nw = NetWork(net, 0)
    nw.InitFromCodes(model, q, max_images=B)
    r = Runner(nw)
    x = torch.from_numpy(raw).cuda()
    a = r.run_device(x, raw224=True).cpu().numpy()
    assert np.array_equal(a[0], a[1])
    perm = rng.permutation(B)
    b = r.run_device(x[torch.from_numpy(perm).cuda()].contiguous(), raw224=True).cpu().numpy()
    assert np.array_equal(b, a[perm])    # permutation equivariance over images
    c = r.run_device(x[:5].contiguous(), raw224=True).cpu().numpy()
    assert np.array_equal(c, a[:5])      # a sub-batch reproduces the same images
    a2 = r.run_device(x, raw224=True).cpu().numpy()
    assert np.array_equal(a, a2)         # deterministic
    assert a.std() > 1.0                 # the synthetic model keeps the logits alive
    nw.CleanUp()


def test_resnet50_batch256_plan_matches_oracle(resnet_model):
    """The BENCHMARKED configuration itself (B = 256, variant AUTO: 30 layers run as CTA pairs, which a B = 2 run
    never selects on the 7x7 / 14x14 maps) against the oracle and the executed reference: image 0 is the image of
    tests/golden/whole_net_golden.json, every one of its 55 tensors is compared with the oracle and with the
    hashes of what the reference's own device program produced; the first, a middle and the last image of the
    batch are compared at the logits and at the INT32 accumulators of CTA-pair / halo / flat layers."""
    import torch
    from oracle import oracle as O
    from tf2_b200.network import NetWork, Runner
    net, q, model = resnet_model
    B = 256
    rng = np.random.default_rng(5)
    raw = rng.integers(-128, 128, size=(B, 3, 224, 224), dtype=np.int8)
    raw0, _ = formats.prepare_input(net, synth.synth_images(1, seed=11), q)      # the golden case's image
    raw[0] = raw0[0]
    nw = NetWork(net, 0)
    nw.InitFromCodes(model, q, max_images=B, variant=capi.VARIANT_AUTO)
    modes = nw.layer_modes(B)
    assert sum("ctapair" in m for m in modes) >= 20, modes          # the plan under test IS the CTA-pair plan
    assert any("halo" in m for m in modes) and all(k == "mma" for k in nw.layer_kernels())
    r = Runner(nw)
    got = r.run_device(torch.from_numpy(raw).cuda(), raw224=True).cpu().numpy()
    pick = [0, 131, 255]
    t0 = formats.feature_trans(raw[pick]).reshape(len(pick), -1, 114, 114)
    exp = O.run_network(net, model, t0)
    for i, b in enumerate(pick):
        assert np.array_equal(got[b], exp[i]), f"image {b}: logits differ in {(got[b] != exp[i]).sum()} of {exp[i].size}"
    # every tensor of image 0: oracle and reference-device hashes
    tens, accs = H.oracle_tensors(net, model, t0[0])
    dev = {}
    for t in range(1, len(net.tensors)):
        g = dev[t] = r.read_tensor(t, 1).cpu().numpy()[0]
        assert np.array_equal(g, tens[t]), f"tensor {t} ({net.tensors[t].name}) differs in {(g != tens[t]).sum()}"
    H.assert_reference_hashes("resnet50", net, dev)
    # INT32 accumulators out of the tensor-core kernel in this launch plan: conv1 (halo, low plane), a halo 3x3,
    # flat + residual, CTA-pair 1x1 / 3x3 / stride 2, the last bottleneck and the fc layer
    _, accs_last = H.oracle_tensors(net, model, t0[2])
    for l in (0, 3, 4, 13, 26, 29, 45, 47, 52, 53):
        g = r.dump_acc(l, B).cpu().numpy()
        assert np.array_equal(g[0], accs[l]), f"layer {l} [{modes[l]}]: image 0 accumulators differ in {(g[0] != accs[l]).sum()}"
        assert np.array_equal(g[255], accs_last[l]), f"layer {l} [{modes[l]}]: image 255 accumulators differ"
    # the taps left the run's feature maps alone
    assert np.array_equal(r.read_tensor(net.result_tensor(), B).cpu().numpy(), got)
    nw.CleanUp()
