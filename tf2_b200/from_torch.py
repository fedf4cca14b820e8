"""PyTorch model -> the reference's float `param.bin` (calibration and INQ projection start
from a torch model; the reference's own route is Caffe -> caffe2fpga, `ModelConvert/caffe2fpga/src/
caffe2fpga.cpp:48-123`, which dumps every blob as float32 in layer order).

  * `blob_from_modules(net, modules)` — `modules[i]` = (conv or linear, batch-norm or None) of the i-th
    non-ipool layer of `net`; writes weight, bias, and BatchNorm (mean, var, scale factor 1, gamma, beta) in the
    order `model_loader.cpp:154-231` reads them.
  * `torchvision_resnet50(model)` — the module list of a torchvision ResNet50 in the order of the shipped
    `resnet50.h` tables (projection shortcut first).  The shipped tables ARE the torchvision topology: stride 2 on
    the 3x3 convolution of a stage's first block (kConvStride of layers 13 / 26 / 45; the reference ships a
    `pytorch_resnet50_q` next to them), so a torchvision model maps one to one.

Host-side tooling; nothing here is on the inference hot path."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from .netdesc import NetDesc


def _f32(t) -> bytes:
    return np.ascontiguousarray(t.detach().cpu().numpy(), dtype="<f4").tobytes()


def blob_from_modules(net: NetDesc, modules: List[Tuple[object, Optional[object]]]) -> bytes:
    import torch
    conv_layers = [ld for ld in net.layers if not ld.ipool]
    if len(modules) != len(conv_layers):
        raise ValueError(f"{len(modules)} modules for {len(conv_layers)} weight layers")
    out = []
    for ld, (conv, bn) in zip(conv_layers, modules):
        w = conv.weight
        if w.dim() == 2:                                   # Linear: a 1x1 convolution over the pooled map
            w = w[:, :, None, None]
        want = (ld.N, net.input_c, 7, 7) if ld.first_layer_7x7 else (ld.N, ld.C, ld.k, ld.k)
        if tuple(w.shape) != want:
            raise ValueError(f"{ld.name}: module weight {tuple(w.shape)} does not match the tables {want}")
        out.append(_f32(w))
        if ld.bias_en:
            b = conv.bias if getattr(conv, "bias", None) is not None else torch.zeros(ld.N)
            out.append(_f32(b))
        elif getattr(conv, "bias", None) is not None and bool(torch.any(conv.bias != 0)):
            raise ValueError(f"{ld.name}: the module has a bias but the tables have kBiasEnable = 0")
        if ld.bn_en:
            if bn is None:
                raise ValueError(f"{ld.name}: the tables expect BatchNorm / Scale parameters")
            if abs(bn.eps - 1e-5) > 1e-12:
                raise ValueError(f"{ld.name}: LoadModel folds BatchNorm with eps 1e-5 (model_loader.cpp:205), module has {bn.eps}")
            gamma = bn.weight if bn.weight is not None else torch.ones(ld.N)
            beta = bn.bias if bn.bias is not None else torch.zeros(ld.N)
            out += [_f32(bn.running_mean), _f32(bn.running_var), np.array([1.0], "<f4").tobytes(), _f32(gamma), _f32(beta)]
        elif bn is not None:
            raise ValueError(f"{ld.name}: a BatchNorm module for a layer with kBnEnable = 0")
    return b"".join(out)


def torchvision_resnet50(model) -> List[Tuple[object, Optional[object]]]:
    mods = [(model.conv1, model.bn1)]
    for stage in (model.layer1, model.layer2, model.layer3, model.layer4):
        for blk in stage:
            if blk.downsample is not None:
                mods.append((blk.downsample[0], blk.downsample[1]))     # res*_branch1 comes first in the tables
            mods += [(blk.conv1, blk.bn1), (blk.conv2, blk.bn2), (blk.conv3, blk.bn3)]
    mods.append((model.fc, None))
    return mods


def torchvision_squeezenet1_1(model) -> List[Tuple[object, Optional[object]]]:
    """Module list of a torchvision SqueezeNet 1.1 in the layer order of `nets.squeezenet()`: conv1, then squeeze /
    expand1x1 / expand3x3 of every fire module, then the 1x1 classifier convolution (its ReLU is the layer's; the
    13x13 average after it is outside the runtime: full_size_pool.cl is a 7x7 average)."""
    import torch
    mods = [(model.features[0], None)]
    for f in model.features:
        if hasattr(f, "squeeze"):
            mods += [(f.squeeze, None), (f.expand1x1, None), (f.expand3x3, None)]
    conv = [c for c in model.classifier if isinstance(c, torch.nn.Conv2d)]
    mods.append((conv[0], None))
    return mods


def vgg_modules(features, classifier) -> List[Tuple[object, Optional[object]]]:
    """Module list of a torchvision-style VGG (`features`: Conv2d / ReLU / pool sequence, `classifier`: Linear /
    ReLU / Dropout sequence) in the layer order of `nets.vgg16()`: the convolutions, then fc6 as a 7x7 convolution
    over the last map (torch flattens [C][H][W], which is the convolution's own weight order), fc7, fc8."""
    import torch

    class _AsConv:
        def __init__(self, lin, c, k):
            self.weight = lin.weight.reshape(lin.out_features, c, k, k)
            self.bias = lin.bias

    convs = [m for m in features if isinstance(m, torch.nn.Conv2d)]
    lins = [m for m in classifier if isinstance(m, torch.nn.Linear)]
    c_last = convs[-1].out_channels
    k = int(round((lins[0].in_features // c_last) ** 0.5))
    return [(c, None) for c in convs] + [(_AsConv(lins[0], c_last, k), None)] + [(l, None) for l in lins[1:]]
