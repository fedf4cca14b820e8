"""Per-channel Q calibration for power-of-two ("shift") quantisation — SURVEY.md 8f-3.

Restates what the reference's TransForm_Kit does offline to produce the `<net>_Q` file the runtime
reads (`TransForm_Kit/Quantization/quantization.py:33-69`, driven over the dumped float feature maps
of `feature_write.py:72-84` with the layer list of `config.py: resnet50_layer_name_q`):

  * `quantize_for_shift(x)`   — Q with |x| * 2^Q <= 127 (quantization.py:33-46)
  * `quantize_channel(x)`     — one Q per channel, outliers pulled to the mean (quantization.py:48-69)
  * `float_forward(...)`      — float32 forward of a network given as NetDesc + param.bin blob
                                (the feature maps feature_write.py dumps from Caffe/PyTorch)
  * `calibrate(...)`          — Q text in the runtime's file order (README.md:51-55 of Runtime_Engine)

Which feature map defines a layer's Q follows `resnet50_layer_name_q`: a plain layer is measured at
its BatchNorm/Scale output BEFORE ReLU and pooling ('conv1-scale', 'res2a_branch2a-scale'); every
tensor that takes part in a chain of residual adds shares ONE row measured on the LAST tensor of
that chain after add + ReLU ('res2c' for all of stage 2) — the runtime adds int8 values directly
(feature_writer.cl:124-127), so both operands must carry the same Q.

Host-side tooling: float math in PyTorch on the CPU; nothing here is on the inference hot path.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np

from .netdesc import NetDesc


def quantize_for_shift(x: np.ndarray) -> int:
    """quantization.py:33-46."""
    x = np.asarray(x, dtype=np.float64)
    mx = float(np.max(np.abs(x))) if x.size else 0.0
    q = 0.0
    if mx > 0:
        q = math.log2(127.0 / mx)
        q = math.floor(q) if q < 0 else float(np.round(q))
        out = x * (2.0 ** q)
        while np.max(out) > 127 or np.min(out) < -128:
            q -= 1
            out = x * (2.0 ** q)
    return int(q)


def quantize_channel(x: np.ndarray) -> np.ndarray:
    """quantization.py:48-69, style 'shift'.  x is [batch][C][...] (Q per channel, axis 1) or [C]."""
    x = np.asarray(x)
    if x.ndim == 1:
        power = np.array([quantize_for_shift(x[i]) for i in range(x.shape[0])], dtype=np.float64)
    else:
        power = np.array([quantize_for_shift(x[:, i]) for i in range(x.shape[1])], dtype=np.float64)
    mean = power.mean()
    for i in range(power.shape[0]):
        if power[i] > 0 and power[i] > mean:
            power[i] = math.floor(mean)
        if power[i] < 0 and power[i] < mean:
            power[i] = math.ceil(mean)
    return power.astype(np.int64)


# --------------------------------------------------------------------------------------------
def _read(buf: memoryview, pos: int, n: int) -> Tuple[np.ndarray, int]:
    end = pos + 4 * n
    if end > len(buf):
        raise ValueError("model blob too short")
    return np.frombuffer(buf[pos:end], dtype="<f4").copy(), end


def float_forward(net: NetDesc, blob: bytes, images: np.ndarray):
    """Float32 forward pass of `net` with the weights of a param.bin blob (blob order and BatchNorm
    folding as model_loader.cpp:154-231: eps 1e-5, mean/var divided by the scale factor).

    images: float32 [B][3][H][W] (the mean-subtracted input the runtime quantises).
    Returns (tensors, pre, post): tensors[t] = float output of tensor t as the NEXT layer sees it
    (after ReLU / pool / add / global average), pre[l] = layer l's BatchNorm/Scale output before ReLU
    and pooling (None for ipool layers), post[l] = layer l's output after add + ReLU, before the
    global average."""
    import torch
    import torch.nn.functional as F
    buf = memoryview(blob)
    pos = 0
    x0 = torch.from_numpy(np.ascontiguousarray(images, dtype=np.float32))
    tens: Dict[int, "torch.Tensor"] = {}
    if not any(ld.first_layer_7x7 for ld in net.layers):
        tens[0] = x0     # plain stem: tensor 0 is the image itself (no 7x7 -> 3x3 input transform)
    pre: List[Optional[np.ndarray]] = []
    post: List[np.ndarray] = []
    for l, ld in enumerate(net.layers):
        if ld.ipool:
            y = F.max_pool2d(F.pad(tens[ld.in_tensor], (1, 1, 1, 1), value=0.0), 3, 1)
            pre.append(None)
        else:
            N = ld.N
            if ld.first_layer_7x7:
                C, H, W = net.input_c, 7, 7
            else:
                C, H, W = ld.C, ld.k, ld.k
            w, pos = _read(buf, pos, N * C * H * W)
            w = torch.from_numpy(w.reshape(N, C, H, W))
            b = None
            if ld.bias_en:
                bb, pos = _read(buf, pos, N)
                b = torch.from_numpy(bb)
            if ld.first_layer_7x7:
                y = F.conv2d(x0, w, b, stride=2, padding=3)     # what the 27-channel 3x3 form computes
            else:
                xin = tens[ld.in_tensor]
                if xin.shape[1] > ld.C:
                    xin = xin[:, :ld.C]
                y = F.conv2d(xin, w, b, stride=ld.stride, padding=ld.pad)
            if ld.bn_en:
                mean, pos = _read(buf, pos, N)
                var, pos = _read(buf, pos, N)
                sf, pos = _read(buf, pos, 1)
                gamma, pos = _read(buf, pos, N)
                beta, pos = _read(buf, pos, N)
                a = mean / sf[0]
                bsd = np.sqrt(var / sf[0] + np.float32(1e-5))
                alpha = (gamma / bsd).astype(np.float32)
                shift = (-(alpha * a) + beta).astype(np.float32)
                y = y * torch.from_numpy(alpha).view(1, -1, 1, 1) + torch.from_numpy(shift).view(1, -1, 1, 1)
            pre.append(y.numpy().copy())
            if ld.relu:
                y = torch.relu(y)
            if ld.pool:
                p = ld.pool_pad
                need_h = (ld.PH - 1) * ld.pool_stride + 3 - (y.shape[2] + p)
                need_w = (ld.PW - 1) * ld.pool_stride + 3 - (y.shape[3] + p)
                y = F.max_pool2d(F.pad(y, (p, max(need_w, 0), p, max(need_h, 0)), value=0.0), 3, ld.pool_stride)
                y = y[:, :, :ld.PH, :ld.PW]
        if ld.add_tensor >= 0:
            y = y + tens[ld.add_tensor][:, :ld.N]
            if ld.add_relu:
                y = torch.relu(y)
        post.append(y.numpy().copy())
        if ld.gap:
            y = y.mean(dim=(2, 3), keepdim=True)
        t = net.tensors[ld.out_tensor]
        if ld.out_tensor not in tens:
            tens[ld.out_tensor] = torch.zeros((x0.shape[0], t.C, t.H, t.W), dtype=torch.float32)
        tens[ld.out_tensor][:, ld.out_ch0:ld.out_ch0 + ld.N] = y
    if pos != len(buf):
        raise ValueError(f"model blob has {len(buf) - pos} trailing bytes")
    return {k: v.numpy() for k, v in tens.items()}, pre, post


def calibrate(net: NetDesc, blob: bytes, images: np.ndarray) -> Tuple[str, List[np.ndarray]]:
    """Q table for `net`: returns (text in `<net>_Q` file order, per-layer Q arrays).

    File order (Runtime_Engine README.md:51-55, quantization.cpp:36-53): the Q of the 3 image
    channels, then for every non-ipool layer its kOutputChannels values, in runtime layer order."""
    tens, pre, post = float_forward(net, blob, images)
    # residual chains: tensors joined by adds share the Q of the chain's last tensor
    parent: Dict[int, int] = {}

    def find(a: int) -> int:
        while parent.get(a, a) != a:
            a = parent[a]
        return a

    last_in_chain: Dict[int, int] = {}   # chain root -> index of the last layer that adds into the chain
    for ld in net.layers:
        if ld.add_tensor >= 0:
            ra, rb = find(ld.out_tensor), find(ld.add_tensor)
            if ra != rb:
                parent[ra] = rb
    for ld in net.layers:
        if ld.add_tensor >= 0:
            last_in_chain[find(ld.out_tensor)] = net.layers.index(ld)     # layers are in execution order
    chain_q: Dict[int, np.ndarray] = {root: quantize_channel(post[l]) for root, l in last_in_chain.items()}

    img_q = quantize_channel(np.ascontiguousarray(images, dtype=np.float32))
    vals: List[int] = [int(v) for v in img_q]
    per_layer: List[np.ndarray] = []
    for l, ld in enumerate(net.layers):
        if ld.ipool:
            per_layer.append(np.zeros(0, np.int64))
            continue
        root = find(ld.out_tensor)
        if root in chain_q:
            q = chain_q[root][:ld.N]
        elif ld.gap:
            q = quantize_channel(tens[ld.out_tensor][:, ld.out_ch0:ld.out_ch0 + ld.N])
        else:
            q = quantize_channel(pre[l])
        per_layer.append(np.asarray(q, dtype=np.int64))
        vals.extend(int(v) for v in q)
    return "\n".join(str(v) for v in vals) + "\n", per_layer
