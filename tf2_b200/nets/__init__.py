"""Built-in network descriptions.

`resnet50`, `googlenet`, `resnet50_pruned` are the reference's own layer tables
(`Runtime_Engine/cnn/host/inc/{resnet50,googlenet,resnet50_pruned}.h`) resolved by
`NetDesc.from_header` and stored as JSON by tools/gen_nets.py.  Any other TF2 header can be loaded
at run time with `NetDesc.from_header(path)`.
"""
from __future__ import annotations

import json
import os
from typing import List, Optional

from ..netdesc import LayerDesc, NetDesc, TensorDesc

_HERE = os.path.dirname(os.path.abspath(__file__))


def load(name: str) -> NetDesc:
    if name == "vgg16":
        return vgg16()
    if name == "squeezenet":
        return squeezenet()
    p = os.path.join(_HERE, name + ".json")
    if not os.path.exists(p):
        raise KeyError(f"unknown built-in network {name!r}")
    with open(p) as f:
        return NetDesc.from_json(json.load(f))


def resnet50() -> NetDesc:
    return load("resnet50")


def googlenet() -> NetDesc:
    return load("googlenet")


def resnet50_pruned() -> NetDesc:
    return load("resnet50_pruned")


def chain(input_chw, specs: List[dict], name: str = "chain") -> NetDesc:
    """Small hand-made networks for tests: a list of layer specs applied in sequence.

    spec keys: N, k=1, pad=0, stride=1, relu=1, pool=0, pool_stride=1, pool_pad=0, add=None (index
    of an earlier layer whose output is the residual operand, -1 = the input), add_relu=0, gap=0,
    ipool=0, src=None (index of the producing layer, default previous; -1 = input), bias_en=0,
    bn_en=1, concat=None ((group id, channel offset, total channels))."""
    C0, H0, W0 = input_chw
    tensors = [TensorDesc(C0, H0, W0, 0, "input")]
    layers: List[LayerDesc] = []
    out_of = {-1: 0}
    concat_t = {}
    for i, s in enumerate(specs):
        src = s.get("src", i - 1)
        tin = out_of[src]
        ti = tensors[tin]
        k, pad, stride = s.get("k", 1), s.get("pad", 0), s.get("stride", 1)
        ipool = s.get("ipool", 0)
        if ipool:
            N, C, k, pad, stride = ti.C, ti.C, 3, 1, 1
            OH, OW, PH, PW = ti.H, ti.W, ti.H, ti.W
            pool, ps, pp = 1, 1, 1
        else:
            N, C = s["N"], s.get("C", ti.C)
            OH = (ti.H + 2 * pad - k) // stride + 1
            OW = (ti.W + 2 * pad - k) // stride + 1
            pool, ps, pp = s.get("pool", 0), s.get("pool_stride", 1), s.get("pool_pad", 0)
            if pool:
                PH = s.get("PH", (OH + 2 * pp - 3) // ps + 1)
                PW = s.get("PW", (OW + 2 * pp - 3) // ps + 1)
            else:
                PH, PW = OH, OW
        gap = s.get("gap", 0)
        oh, ow = (1, 1) if gap else (PH, PW)
        cc = s.get("concat")
        if cc is not None:
            gid, ch0, ctot = cc
            if gid not in concat_t:
                tensors.append(TensorDesc(ctot, oh, ow, -1, f"concat{gid}"))
                concat_t[gid] = len(tensors) - 1
            tout = concat_t[gid]
        else:
            ch0 = 0
            tensors.append(TensorDesc(N, oh, ow, i + 1, f"out{i}"))
            tout = len(tensors) - 1
        out_of[i] = tout
        add = s.get("add")
        layers.append(LayerDesc(
            name=f"layer{i}", in_tensor=tin, out_tensor=tout, out_ch0=ch0,
            add_tensor=(out_of[add] if add is not None else -1), C=C, N=N, k=k, pad=pad, stride=stride,
            OH=OH, OW=OW, relu=s.get("relu", 1), pool=pool, pool_stride=ps, pool_pad=pp, PH=PH, PW=PW,
            add_relu=s.get("add_relu", 0), gap=gap, ipool=ipool, bias_en=s.get("bias_en", 0),
            bn_en=s.get("bn_en", 1), q_in_row=0, q_out_row=i + 1))
    may = [False] * len(tensors)
    may[0] = True
    for ld in layers:
        if ld.ipool:
            nonneg = not may[ld.in_tensor]
        else:
            nonneg = bool(ld.add_relu if ld.add_tensor >= 0 else ld.relu)
        if not nonneg:
            may[ld.out_tensor] = True
    for ld in layers:
        ld.in_may_be_m128 = 1 if may[ld.in_tensor] else 0
    mo = max(max(t.C for t in tensors), 1)
    return NetDesc(name=name, tensors=tensors, layers=layers, max_out_channel=mo,
                   num_q_rows=len(layers) + 1, input_c=C0, input_h=H0, input_w=W0)


def vgg16(width_div: int = 1, classes: int = 1000, name: Optional[str] = None) -> NetDesc:
    """VGG16 in the runtime's vocabulary (BASELINE configs[3]; the reference ships no table for it):
    thirteen 3x3/pad-1 convolutions with bias + ReLU on 224x224x3, a pool after each of the five
    stages, fc6 as a 7x7 convolution on the 7x7 map, fc7/fc8 as 1x1 convolutions.  VGG's 2x2/stride-2
    max pools do not exist in the reference (`pool.cl:194-199` pools 3x3 only), so the stages end in
    the runtime's 3x3/stride-2 pool without padding (GoogLeNet's form: 224 -> 112 -> ... -> 7, zero
    fill past the far edge).  The first layer is a plain 3-channel convolution: the 7x7 -> 3x3 input
    transform of `input_loader.cpp:27-73` is specific to ResNet/GoogLeNet stems.  `width_div` scales
    all channel counts down (tests)."""
    w = lambda c: max(16, c // width_div)
    specs: List[dict] = []
    size = 224
    for n, reps in ((64, 2), (128, 2), (256, 3), (512, 3), (512, 3)):
        for r in range(reps):
            s = dict(N=w(n), k=3, pad=1, bias_en=1, bn_en=0)
            if r == reps - 1:
                s.update(pool=1, pool_stride=2, pool_pad=0, PH=size // 2, PW=size // 2)
                size //= 2
            specs.append(s)
    specs.append(dict(N=w(4096), k=7, pad=0, bias_en=1, bn_en=0))
    specs.append(dict(N=w(4096), k=1, bias_en=1, bn_en=0))
    specs.append(dict(N=classes, k=1, relu=0, bias_en=1, bn_en=0))
    net = chain((3, 224, 224), specs, name or ("vgg16" if width_div == 1 else f"vgg16_div{width_div}"))
    for ld in net.layers:                      # Q-table rows as quantization.cpp lays them out
        ld.q_in_row = net.tensors[ld.in_tensor].q_row
    return net


def squeezenet(classes: int = 1000, name: str = "squeezenet") -> NetDesc:
    """SqueezeNet 1.1 in the runtime's vocabulary (BASELINE configs[0]; the reference only carries a
    160x160 FaceNet variant as a PyTorch model, `TransForm_Kit/Quantization/models/SqueezeNet/
    SqueezeNet.py:45-87`, and no table): conv1 3x3/stride 2 on 224x224x3, 3x3/stride-2 pools
    (zero fill past the far edge), eight fire modules (1x1 squeeze, then 1x1 and 3x3 expand writing
    one concat buffer), conv10 1x1 -> `classes` maps of 13x13.  The network ends there: the 13x13
    global average that follows in the original is outside `full_size_pool.cl` (7x7 only,
    constant 669)."""
    specs: List[dict] = []
    cat = [0]

    def fire(src, squeeze, expand, pool_after=None):
        """src: index of the producing spec, or a concat group id given as ('cat', gid)"""
        specs.append(dict(N=squeeze, k=1, bias_en=1, bn_en=0, src=src))
        sq = len(specs) - 1
        gid = cat[0]
        cat[0] += 1
        e1 = dict(N=expand, k=1, bias_en=1, bn_en=0, src=sq, concat=(gid, 0, 2 * expand))
        e3 = dict(N=expand, k=3, pad=1, bias_en=1, bn_en=0, src=sq, concat=(gid, expand, 2 * expand))
        specs.extend([e1, e3])
        return len(specs) - 1            # any member of the group names the concat tensor

    specs.append(dict(N=64, k=3, stride=2, pad=0, bias_en=1, bn_en=0, pool=1, pool_stride=2, pool_pad=0, PH=55, PW=55))
    last = 0
    last = fire(last, 16, 64)
    last = fire(last, 16, 64)
    # the pools between fire modules act on a concat buffer: stand-alone 3x3/stride-2 pools do not exist
    # in the reference (ipool is stride 1), so the stride-2 pool is fused into BOTH expand layers of the
    # preceding module (max pooling commutes with channel concatenation)
    for s in specs[-2:]:
        s.update(pool=1, pool_stride=2, pool_pad=0, PH=27, PW=27)
    last = fire(last, 32, 128)
    last = fire(last, 32, 128)
    for s in specs[-2:]:
        s.update(pool=1, pool_stride=2, pool_pad=0, PH=13, PW=13)
    last = fire(last, 48, 192)
    last = fire(last, 48, 192)
    last = fire(last, 64, 256)
    last = fire(last, 64, 256)
    specs.append(dict(N=classes, k=1, bias_en=1, bn_en=0, src=last))
    net = chain((3, 224, 224), specs, name)
    # Q rows: l + 1 per layer, one more per concat buffer (quantization.cpp:36-50)
    nl = net.num_layers
    tails, cats = [0] * nl, [0] * nl
    for l, ld in enumerate(net.layers):
        t = net.tensors[ld.out_tensor]
        if t.name.startswith("concat"):
            gid = int(t.name[len("concat"):])
            tails[l], cats[l] = 1, gid
            t.q_row = nl + 1 + gid
    for ld in net.layers:
        ld.q_in_row = net.tensors[ld.in_tensor].q_row
    net.branch_tail, net.concat_layer = tails, cats
    net.num_q_rows = nl + 1 + cat[0]
    return net
