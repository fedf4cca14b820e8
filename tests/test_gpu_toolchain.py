"""GPU tests of SURVEY.md 8f-3 / 8f-4: the offline tool chain feeding the ENGINE (not only the oracle), and the engine's
output going through the reference's result tools.

  torch model -> param.bin (from_torch) -> INQ grid (compress.quantize_blob) -> 4-bit model file and back ->
  calibrated Q (calibrate) -> NetWork.Init4bit (LoadModel) -> CUDA engine == oracle, bit for bit;
  engine logits -> feature_ddr tile layout (network_helper.cpp:95-118) -> Verify / Evaluation
  (network_helper.cpp:120-207) == the same tools over the oracle's logits.
"""
import os

import numpy as np
import pytest

from tf2_b200 import calibrate as K
from tf2_b200 import capi, formats, nets, synth
from tf2_b200 import compress as Z

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def _float_model():
    """A torchvision ResNet50 when torchvision is here (random initialisation, trained-like BatchNorm statistics),
    else the synthetic float blob: either way a FLOAT model the INQ projection has to quantise."""
    net = nets.resnet50()
    try:
        import torch
        import torchvision
        from tf2_b200 import from_torch
        torch.manual_seed(7)
        m = torchvision.models.resnet50(weights=None)
        g = torch.Generator().manual_seed(8)
        for mod in m.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
                mod.running_var.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
                mod.weight.data.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
                mod.bias.data.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
        return net, from_torch.blob_from_modules(net, from_torch.torchvision_resnet50(m.eval())), 1.0 / 50.0, "torchvision"
    except ImportError:
        return net, synth.synth_float_blob(net, seed=9), 1.0, "synthetic"


def test_calibrated_model_through_the_engine(tmp_path):
    import torch
    from oracle import oracle as O
    from tf2_b200.network import Evaluation, NetWork, Runner, Verify
    net, blob, img_scale, origin = _float_model()
    qblob, _ = Z.quantize_blob(net, blob)                       # INQ projection onto the power-of-two grid
    imgs = (synth.synth_images(3, seed=2) * img_scale).astype(np.float32)
    qtext, _ = K.calibrate(net, qblob, imgs[:2])                # per-channel Q from two calibration images
    # the files a TransForm_Kit user hands to the runtime: 4-bit model + Q text
    m4 = formats.float_blob_to_4bit(net, qblob)
    (tmp_path / "model4.bin").write_bytes(m4)
    (tmp_path / "net_Q").write_text(qtext)
    nw = NetWork(net, 0)
    nw.Init4bit(str(tmp_path / "model4.bin"), str(tmp_path / "net_Q"), max_images=3)
    assert "mma" in nw.layer_kernels()
    r = Runner(nw)
    got = r.Run(imgs)                                           # float images in, like Runner::Run
    # the oracle over the same files
    q = formats.parse_q_text(net, qtext)
    model = formats.load_float_blob(net, qblob, q)
    _, t0 = formats.prepare_input(net, imgs, q)
    exp = O.run_network(net, model, t0)
    assert np.array_equal(got, exp), f"{origin}: engine differs from the oracle in {(got != exp).sum()} of {exp.size}"
    assert exp.std() > 2.0
    # the calibrated INT8 network still computes the float network (image 2 was not a calibration image)
    ref = K.float_forward(net, qblob, imgs)[0][net.result_tensor()].reshape(3, -1)
    qrow = q[net.num_layers]
    for b in range(3):
        deq = got[b].reshape(-1).astype(np.float64) * np.exp2(qrow[:1000].astype(np.float64))
        assert np.corrcoef(ref[b], deq)[0, 1] > 0.9
    # engine output -> the reference's result tools
    for b in range(3):
        tiles = formats.to_device_layout(got[b])                # what the FPGA leaves in feature_ddr
        back = formats.from_device_layout(tiles, 1000, 1, 1)
        assert np.array_equal(back, got[b])
        ev_g, ev_o = Evaluation(got[b], qrow), Evaluation(exp[b], qrow)
        assert ev_g == ev_o and len(ev_g) == 5
        assert all(0 <= l < 1000 and 0.0 <= p <= 1.0 for l, p in ev_g)
        v = Verify(got[b], ref[b].reshape(1000, 1, 1).astype(np.float32), qrow)
        assert v == Verify(exp[b], ref[b].reshape(1000, 1, 1).astype(np.float32), qrow) and np.isfinite(v)
    nw.CleanUp()


def test_two_engines_in_one_process():
    """Two handles alive in one process (on two devices when the box has them, else both on device 0): the
    per-device kernel attributes and SM counts belong to the handle, results do not depend on which engine ran
    first or on interleaving."""
    import torch
    from tests import helpers as H
    from tf2_b200.network import NetWork, Runner
    rng = np.random.default_rng(31)
    net = nets.chain((64, 14, 14), [dict(N=256, k=1, relu=0), dict(N=64, k=1, src=-1), dict(N=64, k=3, pad=1),
                                    dict(N=256, k=1, relu=0, add=0, add_relu=1), dict(N=512, k=1)], "twoeng")
    B = 5
    x = H.random_input(rng, 64, 14, 14, nonneg=False, B=B)
    model = H.random_model(net, rng, x)
    exp = np.stack([H.oracle_tensors(net, model, x[b])[0][net.result_tensor()] for b in range(B)])
    devs = [0, 1] if torch.cuda.device_count() >= 2 else [0, 0]
    # create the second engine FIRST on the other device: a process-wide "attributes already set" flag would
    # leave one device without the > 48 KB shared-memory opt-in
    engines = []
    for d in reversed(devs):
        nw = NetWork(net, d)
        nw.InitFromCodes(model, None, max_images=B, variant=capi.VARIANT_AUTO)
        assert "mma" in nw.layer_kernels()
        engines.append((d, nw, Runner(nw)))
    for rep in range(2):
        for d, nw, r in engines:
            with torch.cuda.device(d):
                got = r.run_device(torch.from_numpy(x).cuda(d)).cpu().numpy()
            assert np.array_equal(got, exp), f"device {d} rep {rep}"
            assert np.array_equal(r.run_host(x), exp)
    for _, nw, _ in engines:
        nw.CleanUp()
