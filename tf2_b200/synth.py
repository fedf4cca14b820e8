"""Synthetic TF2 models: float weight blobs in the reference's `param.bin` format and Q tables.

The reference's real weights are not in its repository (`Runtime_Engine/cnn/host/model/README:1-7`)
and there is no network access, so benchmarks and parity tests use random-initialised models of the
right architecture: INQ-style power-of-two weights with 7 magnitude levels per layer
(`TransForm_Kit/Compression/compress_net/compress_core.py` ComputeQuantumRange), BatchNorm
statistics and a per-channel Q table (the shipped `resnet50_Q` / `googlenet_Q` when given, else a
synthetic one).  The blob is written exactly as caffe2fpga would (`caffe2fpga.cpp:72-107`) so it
goes through the same loader as a real model.
"""
from __future__ import annotations

import io
from typing import Optional

import numpy as np

from .netdesc import NetDesc


def synth_q_text(net: NetDesc, seed: int = 0, spread: int = 2) -> str:
    """A plausible Q file: image Q = 0, feature maps Q in [2, 2+spread] per channel; operands of a
    residual add share their Q rows (the reference requires it)."""
    rng = np.random.default_rng(seed)
    rows = {}
    vals = [0] * net.input_c
    # tensors joined by a residual add must carry identical Q
    same = {}
    for l, ld in enumerate(net.layers):
        if ld.add_tensor >= 0:
            same[ld.out_tensor] = same.get(ld.add_tensor, ld.add_tensor)
    tq = {}
    for l, ld in enumerate(net.layers):
        if ld.ipool:
            continue
        key = same.get(ld.out_tensor, ld.out_tensor)
        if key in tq and net.tensors[key].C == ld.N and not (net.branch_tail and net.branch_tail[l]):
            v = tq[key]
        else:
            v = rng.integers(2, 3 + spread, size=ld.N)
            if not (net.branch_tail and net.branch_tail[l]):
                tq[ld.out_tensor] = v
                tq.setdefault(key, v)
        vals.extend(int(x) for x in v)
    return "\n".join(str(v) for v in vals) + "\n"


def synth_float_blob(net: NetDesc, seed: int = 0, zero_frac: float = 0.2,
                     q: Optional[np.ndarray] = None, target_rms: float = 20.0) -> bytes:
    """Random model in param.bin order (model_loader.cpp:154-213).

    When the Q table `q` (formats.parse_q_*) is given, the BatchNorm scale of every channel is
    calibrated analytically so that the INT8 feature maps keep an rms of about `target_rms` at
    every depth (otherwise a random net decays to a few LSBs and parity tests exercise few bits)."""
    rng = np.random.default_rng(seed)
    out = io.BytesIO()
    lvl_p = np.array([1, 2, 4, 6, 6, 5, 4], dtype=np.float64) / 28.0
    for ld in net.layers:
        if ld.ipool:
            continue
        if ld.first_layer_7x7:
            C, H, W = net.input_c, 7, 7
        else:
            C, H, W = ld.C, ld.k, ld.k
        N = ld.N
        fan_in = C * H * W
        # He-style scale -> largest magnitude level 2^max_exp, 7 levels below it
        std = np.sqrt(2.0 / fan_in)
        max_exp = int(np.clip(np.round(np.log2(std * 2.5)), -8, 0))
        if q is not None and not ld.bn_en:
            # no BatchNorm to rescale (GoogLeNet, fc layers): pick the layer's largest weight level so
            # that the conv output itself has about target_rms LSBs of 2^-Q_out
            Cq = net.input_c if ld.first_layer_7x7 else C
            Qin = -q[ld.q_in_row, :Cq].astype(np.float64)
            Qout = -q[ld.q_out_row, :N].astype(np.float64)
            x_rms = (50.0 if ld.q_in_row == 0 else target_rms / np.sqrt(2.0)) * np.mean(np.exp2(-Qin))
            rel2 = float(np.sum(lvl_p * np.exp2(-2.0 * np.arange(7))))
            want = target_rms * np.mean(np.exp2(-Qout)) / (np.sqrt(fan_in * (1.0 - zero_frac) * rel2) * x_rms)
            max_exp = int(np.clip(np.round(np.log2(want)), -8, 0))
        lv = rng.choice(7, size=(N, C, H, W), p=lvl_p)
        mag = np.exp2((max_exp - lv).astype(np.float32))
        sign = rng.choice(np.array([-1.0, 1.0], dtype=np.float32), size=(N, C, H, W))
        w = (mag * sign).astype(np.float32)
        w[rng.random((N, C, H, W)) < zero_frac] = 0.0
        out.write(w.astype("<f4").tobytes())
        if ld.bias_en:
            out.write(rng.normal(0, 0.01, N).astype("<f4").tobytes())
        if ld.bn_en:
            out.write(rng.normal(0, 0.05, N).astype("<f4").tobytes())          # mean
            out.write(rng.uniform(0.5, 1.5, N).astype("<f4").tobytes())        # variance
            out.write(np.array([1.0], dtype="<f4").tobytes())                  # scale_factor
            var = rng.uniform(0.5, 1.5, N)
            if q is not None:
                # real-domain rms of the input (post-ReLU ~ target/sqrt(2) LSBs of 2^-Q_in) and of
                # the conv output; gamma/sqrt(var) maps it to target_rms LSBs of 2^-Q_out
                Cq = net.input_c if ld.first_layer_7x7 else C
                Qin = -q[ld.q_in_row, :Cq].astype(np.float64)
                Qout = -q[ld.q_out_row, :N].astype(np.float64)
                x_rms = (50.0 if ld.q_in_row == 0 else target_rms / np.sqrt(2.0)) * np.mean(np.exp2(-Qin))
                ew2 = float(np.sum(lvl_p * np.exp2(2.0 * (max_exp - np.arange(7)))))
                conv_rms = np.sqrt(fan_in * (1.0 - zero_frac) * ew2) * x_rms
                gamma = (target_rms * np.exp2(-Qout) / conv_rms) * np.sqrt(var) * rng.uniform(0.8, 1.25, N)
            else:
                gamma = rng.uniform(0.4, 1.2, N)
            out.seek(out.tell() - 4 * N - 4)   # rewrite variance with the values gamma was fitted to
            out.write(var.astype("<f4").tobytes())
            out.write(np.array([1.0], dtype="<f4").tobytes())
            out.write(gamma.astype("<f4").tobytes())                           # gamma
            out.write(rng.normal(0, 0.2, N).astype("<f4").tobytes())           # beta
    return out.getvalue()


def synth_images(n: int, seed: int = 0, c: int = 3, h: int = 224, w: int = 224) -> np.ndarray:
    """Float images ~ N(0, 50) clipped to +-150, the range of the reference's mean-subtracted
    fixtures (`host/test_images/*.bin`: -126..154)."""
    rng = np.random.default_rng(seed)
    return np.clip(rng.normal(0, 50, size=(n, c, h, w)), -150, 150).astype(np.float32)
