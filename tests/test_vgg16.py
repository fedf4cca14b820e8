"""VGG16 in the runtime's vocabulary (BASELINE configs[3]; tf2_b200.nets.vgg16): 224-wide maps (two
tiles per row), a pool after every stage, fc6 as a 7x7 convolution, a plain 3-channel first layer
that can see -128.  CPU: the description is executable by the oracle and goes through the whole
model path (synthetic param.bin -> calibrated Q -> LoadModel); GPU: bit-exact against the oracle on
both kernel families."""
import numpy as np
import pytest

from tests import helpers as H
from tf2_b200 import calibrate as K
from tf2_b200 import capi, formats, nets, synth


def test_vgg16_description():
    net = nets.vgg16()
    assert net.num_layers == 16 and net.macs_per_image() == 15470264320      # 15.470 GMAC
    assert [net.tensors[l.out_tensor].H for l in net.layers if l.pool] == [112, 56, 28, 14, 7]
    assert (net.layers[13].k, net.layers[13].C, net.layers[13].N, net.layers[13].OH) == (7, 512, 4096, 1)
    assert net.layers[0].in_may_be_m128 == 1 and all(l.in_may_be_m128 == 0 for l in net.layers[1:])


def test_vgg16_lite_model_path_and_oracle():
    from oracle import oracle as O
    net = nets.vgg16(width_div=16)
    blob = synth.synth_float_blob(net, seed=4)
    assert len(blob) == formats.float_blob_size(net)
    imgs = synth.synth_images(2, seed=8)
    qtext, _ = K.calibrate(net, blob, imgs)
    q = formats.parse_q_text(net, qtext)
    model = formats.load_float_blob(net, blob, q)
    x = formats.quantize_input(imgs, int(q[0, 0]))            # runner.cpp:158-164; no 7x7 transform for this stem
    y = O.run_network(net, model, x)
    ref = K.float_forward(net, blob, imgs)[0][net.result_tensor()].reshape(2, -1)
    deq = y.reshape(2, -1).astype(np.float64) * np.exp2(q[net.num_layers, :1000].astype(np.float64))
    for b in range(2):
        assert np.corrcoef(ref[b], deq[b])[0, 1] > 0.9


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [capi.VARIANT_AUTO, capi.VARIANT_SHIFT], ids=["auto", "shift"])
def test_vgg16_lite_matches_oracle_on_gpu(variant):
    import torch
    from oracle import oracle as O
    from tf2_b200.network import NetWork, Runner
    net = nets.vgg16(width_div=8)
    rng = np.random.default_rng(12)
    B = 2
    x = H.random_input(rng, 3, 224, 224, nonneg=False, B=B)
    model = H.random_model(net, rng, x)
    exp = O.run_network(net, model, x)
    nw = NetWork(net, 0)
    nw.InitFromCodes(model, None, max_images=B, variant=variant)
    r = Runner(nw)
    got = r.run_device(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(got, exp), f"logits differ in {(got != exp).sum()} of {exp.size}"
    tens, _ = H.oracle_tensors(net, model, x[0])
    dev = {}
    for t in range(1, len(net.tensors)):
        g = dev[t] = r.read_tensor(t, B).cpu().numpy()[0]
        assert np.array_equal(g, tens[t]), f"tensor {t} differs in {(g != tens[t]).sum()}"
    # ... and directly against what the reference's own device program produced for this model and image
    # (tests/test_generated_nets_ref.py: case vgg16_div8)
    H.assert_reference_hashes("vgg16_div8", net, dev, golden="generated_nets_golden.json")
    if variant == capi.VARIANT_AUTO:
        assert "mma" in nw.layer_kernels()
    nw.CleanUp()


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_vgg16_full_size_matches_oracle_on_gpu():
    """BASELINE configs[3]'s network at FULL width (15.47 GMAC per image, 138 M weights: the arena holds them as
    int8 planes for the tensor cores and int16 planes for the shift kernel), B = 3, against the oracle: every
    tensor of image 0, the logits of all images, INT32 accumulators of the stem (which sees -128), a 224-wide 3x3,
    a CTA-pair 3x3 and the 7x7 fc6 convolution."""
    import torch
    from oracle import oracle as O
    from tf2_b200.network import NetWork, Runner
    net = nets.vgg16()
    rng = np.random.default_rng(12)
    B = 3
    x = H.random_input(rng, 3, 224, 224, nonneg=False, B=B)
    model = H.random_model(net, rng, x)
    exp = O.run_network(net, model, x)
    nw = NetWork(net, 0)
    nw.InitFromCodes(model, None, max_images=B, variant=capi.VARIANT_AUTO)
    kern = nw.layer_kernels()
    assert kern.count("mma") >= 15, kern
    r = Runner(nw)
    got = r.run_device(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(got, exp), f"logits differ in {(got != exp).sum()} of {exp.size}"
    tens, accs = H.oracle_tensors(net, model, x[0])
    for t in range(1, len(net.tensors)):
        g = r.read_tensor(t, 1).cpu().numpy()[0]
        assert np.array_equal(g, tens[t]), f"tensor {t} differs in {(g != tens[t]).sum()}"
    for l in (0, 1, 7, 13, 15):
        g = r.dump_acc(l, B).cpu().numpy()[0]
        assert np.array_equal(g, accs[l]), f"layer {l} ({kern[l]}): INT32 accumulators differ in {(g != accs[l]).sum()}"
    assert np.array_equal(r.run_host(x), exp)
    nw.CleanUp()
