"""SqueezeNet 1.1 in the runtime's vocabulary (BASELINE configs[0]: one 224x224 image through the CPU
host path, plumbing): the description is executable end to end on the CPU side — synthetic param.bin
-> calibrated per-channel Q (concat rows included) -> LoadModel -> oracle — and the INT8 result tracks
the float network."""
import numpy as np

from tf2_b200 import calibrate as K
from tf2_b200 import formats, nets, synth


def test_squeezenet_description():
    net = nets.squeezenet()
    assert net.num_layers == 26 and abs(net.macs_per_image() / 1e9 - 0.349) < 0.001
    squeezes = [l for l in net.layers[1:-1] if not net.tensors[l.out_tensor].name.startswith("concat")]
    assert [l.N for l in squeezes] == [16, 16, 32, 32, 48, 48, 64, 64] and all(l.k == 1 for l in squeezes)
    cat = {l.out_tensor for l in net.layers if net.tensors[l.out_tensor].name.startswith("concat")}
    assert len(cat) == 8 and all(net.tensors[t].C in (128, 256, 384, 512) for t in cat)
    assert [net.tensors[l.out_tensor].H for l in net.layers if l.pool] == [55, 27, 27, 13, 13]
    r = net.tensors[net.result_tensor()]
    assert (r.C, r.H, r.W) == (1000, 13, 13)


def test_squeezenet_single_image_through_the_cpu_path():
    from oracle import oracle as O
    net = nets.squeezenet()
    blob = synth.synth_float_blob(net, seed=6)
    assert len(blob) == formats.float_blob_size(net)
    img = synth.synth_images(1, seed=3)
    qtext, _ = K.calibrate(net, blob, img)
    assert len(qtext.split()) == formats.q_file_value_count(net)
    q = formats.parse_q_text(net, qtext)
    model = formats.load_float_blob(net, blob, q)
    x = formats.quantize_input(img, int(q[0, 0]))
    y = O.run_network(net, model, x)[0].astype(np.float64)                 # [1000][13][13]
    ref = K.float_forward(net, blob, img)[0][net.result_tensor()][0]
    deq = y * np.exp2(q[net.num_layers, :1000].astype(np.float64))[:, None, None]
    assert np.corrcoef(ref.reshape(-1), deq.reshape(-1))[0, 1] > 0.9
    assert (np.abs(y) >= 127).mean() < 0.05 and y.std() > 2
    # the class scores after the (float) 13x13 average agree on the winner's neighbourhood
    top_ref = np.argsort(ref.mean(axis=(1, 2)))[-5:]
    top_int = np.argsort(deq.mean(axis=(1, 2)))[-20:]
    assert len(set(top_ref) & set(top_int)) >= 3


import pytest  # noqa: E402


@pytest.mark.gpu
def test_squeezenet_matches_oracle_on_gpu():
    """every tensor of image 0 and the 13x13 class maps of both images, both kernel families"""
    import torch
    from oracle import oracle as O
    from tests import helpers as H
    from tf2_b200 import capi
    from tf2_b200.network import NetWork, Runner
    net = nets.squeezenet()
    rng = np.random.default_rng(21)
    B = 2
    x = H.random_input(rng, 3, 224, 224, nonneg=False, B=B)
    model = H.random_model(net, rng, x)
    exp = O.run_network(net, model, x)
    tens, _ = H.oracle_tensors(net, model, x[0])
    for variant in (capi.VARIANT_AUTO, capi.VARIANT_SHIFT):
        nw = NetWork(net, 0)
        nw.InitFromCodes(model, None, max_images=B, variant=variant)
        r = Runner(nw)
        got = r.run_device(torch.from_numpy(x).cuda()).cpu().numpy()
        assert np.array_equal(got, exp)
        for t in range(1, len(net.tensors)):
            assert np.array_equal(r.read_tensor(t, B).cpu().numpy()[0], tens[t]), f"tensor {t}"
        nw.CleanUp()
