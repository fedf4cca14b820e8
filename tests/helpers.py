"""Shared helpers for the parity tests: random layers/models and oracle-side execution."""
from __future__ import annotations

import numpy as np

from oracle import oracle as O
from tf2_b200 import nets
from tf2_b200.netdesc import NetDesc


def random_codes(rng, N, C, k, shift_lo=3, shift_hi=12, zero_frac=0.15, per_n_offset=3, per_c_offset=2):
    """Shift codes as LoadModel would emit them: shift = base + q-like per-n and per-c offsets +
    a 7-level weight exponent; bit7 sign; 0x40 zeros."""
    lvl = rng.integers(0, 7, size=(N, C, k, k))
    sh = shift_lo + lvl + rng.integers(0, per_n_offset + 1, size=(N, 1, 1, 1)) + rng.integers(0, per_c_offset + 1, size=(1, C, 1, 1))
    sh = np.clip(sh, 0, max(shift_hi, shift_lo + 6 + per_n_offset + per_c_offset)).astype(np.uint8)
    sign = (rng.random((N, C, k, k)) < 0.5).astype(np.uint8) << 7
    codes = (sh & 0x1f) | sign
    codes[rng.random((N, C, k, k)) < zero_frac] = 0x40
    return codes.astype(np.uint8)


def random_input(rng, C, H, W, nonneg, B=None):
    lo = 0 if nonneg else -128
    shape = (C, H, W) if B is None else (B, C, H, W)
    x = rng.integers(lo, 128, size=shape).astype(np.int8)
    if not nonneg:
        # make sure the int8 negate quirk (pe.cl:32-34) is exercised
        m = rng.random(shape) < 0.02
        x[m] = -128
    return x


def fit_params(rng, ld, tin, X, codes, bias_en=True):
    """Random BiasBnParam whose alpha/beta put the requantised map inside int8 without saturating
    everywhere (derived from the oracle's accumulators of one sample)."""
    N = ld.N
    params = np.zeros((N, 3), dtype=np.int32)
    if bias_en:
        params[:, 0] = rng.integers(-(1 << 18), 1 << 18, size=N)
    tmp = np.zeros((N, 3), dtype=np.int32)
    tmp[:, 0] = params[:, 0]
    tmp[:, 1] = 1 << 20
    _, acc = O.layer_forward(_conv_only(ld), tin, X, codes, tmp, want_acc=True)
    std = acc.reshape(N, -1).astype(np.float64).std(axis=1) + 1.0
    # y ~ acc*alpha/2^35 ; aim at |y| ~ 40
    alpha = (rng.uniform(0.5, 1.5, N) * 40.0 * (2.0 ** 35) / std)
    params[:, 1] = np.clip(alpha, 1, 2 ** 31 - 1).astype(np.int64).astype(np.int32)
    params[:, 2] = rng.integers(-(1 << 19), 1 << 19, size=N)
    return params


def _conv_only(ld):
    import copy
    c = copy.copy(ld)
    c.pool, c.gap, c.add_tensor, c.PH, c.PW = 0, 0, -1, ld.OH, ld.OW
    return c


def random_model(net: NetDesc, rng, t0):
    """Random codes/params for every conv layer, fitted layer by layer on image 0 of `t0`
    ([B][C][H][W]) so activations stay alive through the net.  Returns model list."""
    model = []
    tens = {0: t0[0]}
    for l, ld in enumerate(net.layers):
        tin = net.tensors[ld.in_tensor]
        X = tens[ld.in_tensor]
        if ld.ipool:
            model.append((None, None))
            codes = params = None
        else:
            codes = random_codes(rng, ld.N, ld.C, ld.k)
            params = fit_params(rng, ld, tin, X, codes, bias_en=True)
            model.append((codes, params))
        R = tens[ld.add_tensor] if ld.add_tensor >= 0 else None
        y = O.layer_forward(ld, tin, X, codes, params, R=R)
        to = net.tensors[ld.out_tensor]
        if ld.out_tensor not in tens:
            tens[ld.out_tensor] = np.zeros((to.C, to.H, to.W), np.int8)
        tens[ld.out_tensor][ld.out_ch0:ld.out_ch0 + ld.N] = y.reshape(ld.N, to.H, to.W)
    return model


def oracle_tensors(net: NetDesc, model, x0):
    """All tensors of one image from the oracle (dict tensor id -> [C][H][W]) + per-layer accs."""
    tens = {0: x0}
    accs = {}
    for l, ld in enumerate(net.layers):
        tin = net.tensors[ld.in_tensor]
        codes, params = model[l]
        R = tens[ld.add_tensor] if ld.add_tensor >= 0 else None
        if ld.ipool:
            y = O.layer_forward(ld, tin, tens[ld.in_tensor], None, None)
        else:
            y, acc = O.layer_forward(ld, tin, tens[ld.in_tensor], codes, params, R=R, want_acc=True)
            accs[l] = acc
        to = net.tensors[ld.out_tensor]
        if ld.out_tensor not in tens:
            tens[ld.out_tensor] = np.zeros((to.C, to.H, to.W), np.int8)
        tens[ld.out_tensor][ld.out_ch0:ld.out_ch0 + ld.N] = y.reshape(ld.N, to.H, to.W)
    return tens, accs


# ---- the synthetic whole-network cases shared by the GPU parity tests and the compiled-reference pin ----
SYNTH_CASES = {"resnet50": (3, 11), "googlenet": (5, 17), "resnet50_pruned": (5, 17)}      # (blob seed, image seed)


def synth_case(name):
    """(net, q, model, t0 of image 0) exactly as tests/test_gpu_resnet50.py / test_gpu_nets.py build them."""
    import os
    from tf2_b200 import formats, synth
    golden = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    blob_seed, img_seed = SYNTH_CASES[name]
    net = nets.load(name)
    q = formats.parse_q_file(net, os.path.join(golden, f"{name}_Q"))
    model = formats.load_float_blob(net, synth.synth_float_blob(net, seed=blob_seed, q=q), q)
    _, t0 = formats.prepare_input(net, synth.synth_images(1, seed=img_seed), q)
    return net, q, model, t0[0]


def layer_output_hashes(net: NetDesc, tens):
    """SHA-256 of every layer's slice of its output tensor (int8 [N][H][W]) — the form in which
    tests/golden/whole_net_golden.json records what the reference's device program produced."""
    import hashlib
    out = []
    for ld in net.layers:
        a = np.ascontiguousarray(np.asarray(tens[ld.out_tensor])[ld.out_ch0:ld.out_ch0 + ld.N], dtype=np.int8)
        out.append(hashlib.sha256(a.tobytes()).hexdigest())
    return out


def assert_reference_hashes(case, net: NetDesc, tens, golden="whole_net_golden.json"):
    """`tens` (tensor id -> int8 [C][H][W] of image 0, from the oracle or read back from the GPU) against
    what the reference's own device program produced for this case (tests/golden/whole_net_golden.json,
    generated_nets_golden.json)."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", golden)) as f:
        g = json.load(f)[case]
    got = layer_output_hashes(net, tens)
    assert len(got) == len(g["layers"])
    checked = 0
    for l, (a, b) in enumerate(zip(got, g["layers"])):
        if b is not None:                            # layers that only write feature_ddr are checked by their consumers
            assert a == b, f"{case}: layer {l} differs from the reference device program"
            checked += 1
    assert checked >= len(got) - 5
    assert got[-1] == g["final"], f"{case}: final map differs from the reference device program"
    return g
