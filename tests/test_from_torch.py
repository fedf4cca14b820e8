"""torch model -> param.bin -> (INQ projection -> calibrated Q -> LoadModel -> oracle): SURVEY.md 8f-3 on a
torchvision ResNet50 (random initialisation: there is no network for pretrained weights).  The blob order, the
BatchNorm folding and the table topology are checked by comparing the float forward pass over the blob with the
torch model itself."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
torchvision = pytest.importorskip("torchvision")

from tf2_b200 import calibrate as K  # noqa: E402
from tf2_b200 import compress as Z  # noqa: E402
from tf2_b200 import formats, from_torch, nets, synth  # noqa: E402


@pytest.fixture(scope="module")
def torch_resnet50():
    torch.manual_seed(7)
    m = torchvision.models.resnet50(weights=None)
    g = torch.Generator().manual_seed(8)
    for mod in m.modules():                                 # BatchNorm statistics as after training, not 0 / 1
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
            mod.weight.data.copy_(torch.rand(mod.num_features, generator=g) * 0.5 + 0.75)
            mod.bias.data.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
    return m.eval()


def test_blob_of_a_torchvision_resnet50_computes_the_same_function(torch_resnet50):
    net = nets.resnet50()
    blob = from_torch.blob_from_modules(net, from_torch.torchvision_resnet50(torch_resnet50))
    assert len(blob) == formats.float_blob_size(net)
    imgs = synth.synth_images(2, seed=1) / 50.0
    with torch.no_grad():
        want = torch_resnet50(torch.from_numpy(imgs)).numpy()
    got = K.float_forward(net, blob, imgs)[0][net.result_tensor()].reshape(2, -1)
    assert np.allclose(got, want, rtol=1e-3, atol=1e-3 * np.abs(want).max())
    with pytest.raises(ValueError):
        from_torch.blob_from_modules(net, from_torch.torchvision_resnet50(torch_resnet50)[:-1])


def test_torch_model_to_served_int8_model(torch_resnet50):
    from oracle import oracle as O
    net = nets.resnet50()
    blob = from_torch.blob_from_modules(net, from_torch.torchvision_resnet50(torch_resnet50))
    qblob, min_exps = Z.quantize_blob(net, blob)            # INQ projection, no retraining
    imgs = synth.synth_images(2, seed=2) / 50.0
    qtext, _ = K.calibrate(net, qblob, imgs)
    q = formats.parse_q_text(net, qtext)
    model = formats.load_float_blob(net, qblob, q)
    raw, t0 = formats.prepare_input(net, imgs, q)
    y = O.run_network(net, model, t0).reshape(2, -1).astype(np.float64)
    deq = y * np.exp2(q[net.num_layers, :1000].astype(np.float64))
    ref = K.float_forward(net, qblob, imgs)[0][net.result_tensor()].reshape(2, -1)
    for b in range(2):
        assert np.corrcoef(ref[b], deq[b])[0, 1] > 0.9
    # the 4-bit model file of this network: an eighth of the float blob, identity round trip
    m4 = formats.float_blob_to_4bit(net, qblob)
    assert len(m4) < len(qblob) / 6 and formats.float_blob_from_4bit(net, m4) == qblob


def test_squeezenet_tables_compute_torchvision_squeezenet1_1():
    """nets.squeezenet() (BASELINE configs[0]; the reference ships no SqueezeNet tables) against torchvision's own
    SqueezeNet 1.1: stride-2 stem with its pool, fire-module concats, the stride-2 pools fused into both expand
    layers, the 1x1 classifier with ReLU over the 13x13 map."""
    torch.manual_seed(3)
    m = torchvision.models.squeezenet1_1(weights=None).eval()
    net = nets.squeezenet()
    blob = from_torch.blob_from_modules(net, from_torch.torchvision_squeezenet1_1(m))
    assert len(blob) == formats.float_blob_size(net)
    imgs = synth.synth_images(2, seed=5) / 50.0
    with torch.no_grad():
        x = m.features(torch.from_numpy(imgs))
        want = m.classifier[2](m.classifier[1](x)).numpy()             # conv + ReLU, before the 13x13 average
    got = K.float_forward(net, blob, imgs)[0][net.result_tensor()]
    assert got.shape == want.shape == (2, 1000, 13, 13)
    assert np.allclose(got, want, rtol=1e-3, atol=1e-3 * np.abs(want).max())


def test_vgg16_tables_compute_a_torch_vgg16():
    """nets.vgg16() (BASELINE configs[3]) against a torchvision VGG16 at a quarter of the width whose 2x2 pools are
    replaced by the runtime's 3x3 / stride-2 window (pool.cl knows no other; 224 -> 112 needs ceil mode): the 13
    convolutions, the five pools, fc6 as a 7x7 convolution in torch's flatten order, fc7, fc8."""
    from torchvision.models.vgg import make_layers
    torch.manual_seed(11)
    cfg = [16, 16, "M", 32, 32, "M", 64, 64, 64, "M", 128, 128, 128, "M", 128, 128, 128, "M"]
    features = make_layers(cfg)
    for i, mod in enumerate(features):
        if isinstance(mod, torch.nn.MaxPool2d):
            features[i] = torch.nn.MaxPool2d(3, 2, ceil_mode=True)
    classifier = torch.nn.Sequential(torch.nn.Linear(128 * 7 * 7, 1024), torch.nn.ReLU(), torch.nn.Linear(1024, 1024),
                                     torch.nn.ReLU(), torch.nn.Linear(1024, 1000))
    net = nets.vgg16(width_div=4)
    blob = from_torch.blob_from_modules(net, from_torch.vgg_modules(features, classifier))
    assert len(blob) == formats.float_blob_size(net)
    imgs = synth.synth_images(2, seed=6) / 50.0
    with torch.no_grad():
        want = classifier(torch.flatten(features.eval()(torch.from_numpy(imgs)), 1)).numpy()
    got = K.float_forward(net, blob, imgs)[0][net.result_tensor()].reshape(2, -1)
    assert np.allclose(got, want, rtol=1e-3, atol=1e-3 * np.abs(want).max())
