"""GPU parity of the tcgen05 INT8 tensor-core path (kernel B) against the CPU oracle, on the tile
geometries the networks use (flat 1x1 tiles; (tw,th,tn) boxes for 56/28/14/7-wide maps), channel
counts that exercise TMA zero fill, multi-plane weights and odd batch sizes."""
import zlib

import numpy as np
import pytest

from tests import helpers as H
from tf2_b200 import capi, nets

pytestmark = pytest.mark.gpu

# name, input CHW, specs, batch, random_codes kwargs
CASES = [
    ("flat_1x1_c256_n64_14", (256, 14, 14), [dict(N=64, k=1)], 3, {}),
    ("flat_1x1_c64_n256_56", (64, 56, 56), [dict(N=256, k=1, relu=0)], 2, {}),
    ("flat_1x1_c1024_n256", (1024, 14, 14), [dict(N=256, k=1)], 2, {}),
    ("flat_1x1_c192_n96_28", (192, 28, 28), [dict(N=96, k=1)], 3, {}),
    ("flat_1x1_c48_n24", (48, 7, 7), [dict(N=24, k=1)], 5, {}),
    ("box_3x3_c64_56", (64, 56, 56), [dict(N=64, k=3, pad=1)], 2, {}),
    ("box_3x3_c128_28", (128, 28, 28), [dict(N=128, k=3, pad=1)], 3, {}),
    ("box_3x3_c256_14", (256, 14, 14), [dict(N=256, k=3, pad=1)], 3, {}),
    ("box_3x3_c512_7", (512, 7, 7), [dict(N=512, k=3, pad=1)], 5, {}),
    ("box_5x5_c16_28", (16, 28, 28), [dict(N=32, k=5, pad=2)], 2, {}),
    ("box_3x3_c96_n208_14", (96, 14, 14), [dict(N=208, k=3, pad=1)], 2, {}),
    ("box_3x3_nopad_c32_20", (32, 20, 20), [dict(N=64, k=3, pad=0)], 2, {}),
    ("planes2", (128, 14, 14), [dict(N=128, k=1)], 2, dict(per_c_offset=4, per_n_offset=5)),
    ("planes3_3x3", (64, 14, 14), [dict(N=64, k=3, pad=1)], 2, dict(per_c_offset=9, per_n_offset=3)),
    ("fc_like_c2048_n1000", (2048, 1, 1), [dict(N=1000, k=1, relu=0, bias_en=1, bn_en=0)], 7, {}),
    ("bottleneck_residual", (256, 14, 14), [dict(N=64, k=1), dict(N=64, k=3, pad=1),
                                            dict(N=256, k=1, relu=0, add=-1, add_relu=1)], 3, {}),
    ("wide_map_112", (32, 112, 112), [dict(N=64, k=3, pad=1)], 1, {}),
    ("s2_3x3_c128_56to28", (128, 56, 56), [dict(N=128, k=3, pad=1, stride=2)], 2, {}),
    ("s2_3x3_c256_28to14", (256, 28, 28), [dict(N=256, k=3, pad=1, stride=2)], 3, {}),
    ("s2_3x3_c512_14to7", (512, 14, 14), [dict(N=512, k=3, pad=1, stride=2)], 5, {}),
    ("s2_1x1_c256_56to28", (256, 56, 56), [dict(N=512, k=1, stride=2, relu=0)], 2, {}),
    ("s2_1x1_c1024_14to7", (1024, 14, 14), [dict(N=2048, k=1, stride=2, relu=0)], 3, {}),
    ("s2_3x3_odd_15to8", (64, 15, 15), [dict(N=64, k=3, pad=1, stride=2)], 3, {}),
    # single-plane layers take the 256-wide N tile (two epilogue passes per warp)
    ("bn256_flat_c256_n512", (256, 14, 14), [dict(N=512, k=1)], 3, dict(per_c_offset=0)),
    ("bn256_box_c128_n256", (128, 14, 14), [dict(N=256, k=3, pad=1)], 3, dict(per_c_offset=0)),
    ("bn256_n384_ragged", (64, 10, 10), [dict(N=384, k=1, relu=0)], 2, dict(per_c_offset=0)),
    ("bn256_residual", (512, 7, 7), [dict(N=128, k=1), dict(N=512, k=1, relu=0, add=-1, add_relu=1)], 3, dict(per_c_offset=0)),
    ("bn256_fc_n1000", (512, 1, 1), [dict(N=1000, k=1, relu=0, bias_en=1, bn_en=0)], 5, dict(per_c_offset=0)),
]


def test_mma_first_layer_quirk():
    """A layer fed by the network input (may hold -128) runs on the tensor cores through the
    negated copy of tensor 0 and still reproduces pe.cl:32-34's int8 negate."""
    import torch
    from tf2_b200.network import NetWork, Runner
    rng = np.random.default_rng(77)
    for chw, spec in (((27, 30, 30), dict(N=64, k=3, pad=0)), ((3, 20, 20), dict(N=32, k=3, pad=1, relu=0)),
                      ((40, 9, 9), dict(N=48, k=1))):
        net = nets.chain(chw, [spec])
        x = H.random_input(rng, *chw, nonneg=False, B=3)
        x[0, :, :2, :] = -128
        model = H.random_model(net, rng, x)
        # sprinkle low absolute shifts (like the code-0 taps of the transformed first layer) among
        # the high ones: exercises the unscaled "low" plane of the tensor-core path
        codes, params = model[0]
        m = (rng.random(codes.shape) < 0.1) & ((codes & 0x40) == 0)
        codes[m] = (codes[m] & 0x80) | rng.integers(0, 4, size=int(m.sum())).astype(np.uint8)
        nw = NetWork(net, 0)
        nw.InitFromCodes(model, None, max_images=3, variant=capi.VARIANT_MMA)
        assert nw.layer_kernels() == ["mma"], nw.layer_kernels()
        r = Runner(nw)
        out = r.run_device(torch.from_numpy(x).cuda()).cpu().numpy()
        xh = np.ascontiguousarray(x.transpose(0, 2, 3, 1))
        out2 = r.run_device(torch.from_numpy(xh).cuda(), in_layout=capi.LAYOUT_HWC).cpu().numpy()
        for b in range(3):
            tens, _ = H.oracle_tensors(net, model, x[b])
            assert np.array_equal(out[b], tens[1]), f"{chw}: image {b} differs in {(out[b] != tens[1]).sum()}"
            assert np.array_equal(out2[b], tens[1])
        nw.CleanUp()


@pytest.mark.parametrize("name,chw,specs,B,ckw", CASES, ids=[c[0] for c in CASES])
def test_mma_parity(name, chw, specs, B, ckw):
    import torch
    from tf2_b200.network import NetWork, Runner
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    net = nets.chain(chw, specs, name)
    x = H.random_input(rng, *chw, nonneg=True, B=B)
    # the network input is flagged "may hold -128" (exact path); feed the conv under test from a
    # ReLU'd identity-like 1x1 layer instead so that it is eligible for the tensor-core path
    pre = [dict(N=chw[0], k=1, relu=1)]
    net = nets.chain(chw, pre + [dict(s, **({"src": s["src"] + 1} if "src" in s and s["src"] >= 0 else {}),
                                      **({"add": s["add"] + 1} if "add" in s and s["add"] is not None else {}))
                                 for s in specs], name)
    old = H.random_codes
    try:
        if ckw:
            H.random_codes = lambda rng_, N, C, k, **kw: old(rng_, N, C, k, **ckw)
        model = H.random_model(net, rng, x)
    finally:
        H.random_codes = old
    nw = NetWork(net, 0)
    nw.InitFromCodes(model, None, max_images=B, variant=capi.VARIANT_MMA)
    kern = nw.layer_kernels()
    assert "mma" in kern[1:], f"{name}: tensor-core path was not selected ({kern})"
    r = Runner(nw)
    out = r.run_device(torch.from_numpy(x).cuda()).cpu().numpy()
    for b in range(B):
        tens, accs = H.oracle_tensors(net, model, x[b])
        for t in range(1, len(net.tensors)):
            got = r.read_tensor(t, B).cpu().numpy()[b]
            bad = (got != tens[t])
            assert not bad.any(), f"{name}: image {b} tensor {t} differs in {bad.sum()} of {bad.size} (first {np.argwhere(bad)[:4].tolist()})"
        assert np.array_equal(out[b], tens[net.result_tensor()])
        if b == 0:
            # the INT32 accumulators of the tensor-core kernel itself (pe.cl:196-199's tap): the same launch plan
            # with the exact epilogue writing bias + sum of shifted features before requantisation
            for l in range(1, net.num_layers):
                if kern[l] != "mma":
                    continue
                g = r.dump_acc(l, B).cpu().numpy()
                for bb in range(B):
                    want = accs[l] if bb == 0 else H.oracle_tensors(net, model, x[bb])[1][l]
                    assert np.array_equal(g[bb], want), f"{name}: layer {l} image {bb} INT32 accumulators differ in {(g[bb] != want).sum()}"
            # the tap must leave the feature maps of the run untouched
            assert np.array_equal(r.read_tensor(net.result_tensor(), B).cpu().numpy(), out)
    # a second run with fewer images than max_images (tiles past the batch end)
    if B > 1:
        out1 = r.run_device(torch.from_numpy(x[:1].copy()).cuda()).cpu().numpy()
        assert np.array_equal(out1[0], out[0])
    nw.CleanUp()


# ---- staging / MMA modes: make sure each one is actually selected and is bit-exact -----------------
# name, input CHW, spec, batch, random_codes kwargs, substring the launch plan must contain
MODE_CASES = [
    # CTA pairs (tcgen05.mma.cta_group::2): streamed weights, 256-wide MMA; odd and even m-tile counts
    ("pair_flat_p1_c512_n256", (512, 14, 14), dict(N=256, k=1), 5, dict(per_c_offset=0), "ctapair"),
    ("pair_flat_p2_c512_n128", (512, 28, 28), dict(N=128, k=1), 1, dict(per_c_offset=4, per_n_offset=5), "ctapair"),
    ("pair_flat_odd_tiles", (1024, 9, 9), dict(N=256, k=1), 7, dict(per_c_offset=0), "ctapair"),   # 567 pixels: 5 m-tiles
    ("pair_box_s2_c256", (256, 28, 28), dict(N=256, k=3, pad=1, stride=2), 5, dict(per_c_offset=0), "box"),
    ("pair_box_c512_7", (512, 7, 7), dict(N=256, k=3, pad=1), 16, dict(per_c_offset=0), "box"),     # 49 of 128 rows: stays on boxes
    # CTA pairs on halo tiles with streamed weights (stride-1 k x k, C >= 128): one activation box per tile and chunk
    ("pair_hs_p1_c256_14", (256, 14, 14), dict(N=256, k=3, pad=1), 9, dict(per_c_offset=0), "halo wstream"),
    ("pair_hs_p2_c128_28", (128, 28, 28), dict(N=128, k=3, pad=1), 3, dict(shift_lo=1, per_c_offset=4, per_n_offset=0), "halo wstream"),
    ("pair_hs_residual", (256, 14, 14), dict(N=256, k=3, pad=1, relu=0, add=0, add_relu=1), 6, dict(per_c_offset=0), "halo wstream"),
    ("pair_hs_p2_c256_n256_two_ntiles", (256, 14, 14), dict(N=256, k=3, pad=1), 5, dict(shift_lo=1, per_c_offset=4, per_n_offset=0), "halo wstream"),
    ("pair_hs_ragged_rows_c128_13", (128, 13, 13), dict(N=128, k=3, pad=1), 3, dict(per_c_offset=4, per_n_offset=0), "halo wstream"),
    ("pair_hs_5x5_c128_20", (128, 20, 20), dict(N=256, k=5, pad=2), 2, dict(per_c_offset=0), "halo wstream"),
    ("pair_hs_nopad_c128_30", (128, 30, 30), dict(N=128, k=3, pad=0), 2, dict(per_c_offset=4, per_n_offset=0), "halo wstream"),
    ("pair_hs_wide_c128_150", (128, 20, 150), dict(N=128, k=3, pad=1), 1, dict(per_c_offset=4, per_n_offset=0), "halo wstream"),   # W tiled: 126 + 24 columns
    ("pair_hs_c384_n200_28", (384, 28, 28), dict(N=200, k=3, pad=1), 1, dict(per_c_offset=4, per_n_offset=0), "halo wstream"),    # 3 chunks, ragged N
    # halo tiles: one TMA box per tile, taps as row-shifted descriptor views
    ("halo_3x3_c64_56", (64, 56, 56), dict(N=64, k=3, pad=1), 2, {}, "halo"),
    ("halo_3x3_c64_56_p1", (64, 56, 56), dict(N=64, k=3, pad=1), 2, dict(per_c_offset=0), "halo"),
    ("halo_5x5_c16_28", (16, 28, 28), dict(N=32, k=5, pad=2), 3, dict(per_c_offset=0), "halo"),
    ("halo_3x3_nopad_c32_30", (32, 30, 30), dict(N=64, k=3, pad=0), 2, {}, "halo"),
    ("halo_3x3_c192_9_three_chunks", (192, 9, 9), dict(N=64, k=3, pad=1), 4, dict(per_c_offset=0), "halo"),
    ("halo_ragged_rows_c32_13", (32, 13, 13), dict(N=32, k=3, pad=1), 3, {}, "halo"),
]


@pytest.mark.parametrize("name,chw,spec,B,ckw,want", MODE_CASES, ids=[c[0] for c in MODE_CASES])
def test_mma_modes(name, chw, spec, B, ckw, want):
    import torch
    from tf2_b200.network import NetWork, Runner
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    net = nets.chain(chw, [dict(N=chw[0], k=1, relu=1), dict(spec)], name)
    x = H.random_input(rng, *chw, nonneg=True, B=B)
    old = H.random_codes
    try:
        if ckw:
            H.random_codes = lambda rng_, N, C, k, **kw: old(rng_, N, C, k, **ckw)
        model = H.random_model(net, rng, x)
    finally:
        H.random_codes = old
    nw = NetWork(net, 0)
    nw.InitFromCodes(model, None, max_images=B, variant=capi.VARIANT_MMA)
    plan = nw.layer_modes(B)[1]
    assert want in plan, f"{name}: expected '{want}' in the launch plan, got '{plan}'"
    if name.startswith("pair"):
        assert "ctapair" in plan and "fold" in plan, plan
    r = Runner(nw)
    out = r.run_device(torch.from_numpy(x).cuda()).cpu().numpy()
    for b in range(B):
        tens, _ = H.oracle_tensors(net, model, x[b])
        bad = out[b] != tens[net.result_tensor()]
        assert not bad.any(), f"{name} [{plan}]: image {b} differs in {bad.sum()} of {bad.size} (first {np.argwhere(bad)[:4].tolist()})"
    # INT32 accumulators of this very launch plan (CTA pairs / halo tiles included)
    g = r.dump_acc(1, B).cpu().numpy()
    for b in range(B):
        want = H.oracle_tensors(net, model, x[b])[1][1]
        assert np.array_equal(g[b], want), f"{name} [{plan}]: image {b} INT32 accumulators differ in {(g[b] != want).sum()}"
    # fewer images than max_images: other tile counts (pairs with a missing partner, ragged last tile)
    for nb in {1, max(1, B - 1)}:
        o = r.run_device(torch.from_numpy(x[:nb].copy()).cuda()).cpu().numpy()
        assert np.array_equal(o, out[:nb]), f"{name}: sub-batch of {nb} differs"
    nw.CleanUp()


# ---- weights from packed 4-bit tiles, expanded on the fly (tf2b_set_weight_staging) ----------------------------------
PACKED4_CASES = [
    ("p4_halo_3x3_c64_56", (64, 56, 56), dict(N=64, k=3, pad=1), 2, {}),
    ("p4_halo_5x5_c16_28", (16, 28, 28), dict(N=32, k=5, pad=2), 3, dict(per_c_offset=0)),
    ("p4_flat_wres_c64_n256", (64, 28, 28), dict(N=256, k=1, relu=0), 3, {}),
    ("p4_flat_wres_c256_n64_bk128", (256, 14, 14), dict(N=64, k=1), 3, {}),
    ("p4_box_wres_s2_c64", (64, 28, 28), dict(N=64, k=3, pad=1, stride=2), 2, dict(per_c_offset=0)),
    ("p4_residual_bn256", (512, 7, 7), dict(N=128, k=1), 3, dict(per_c_offset=0)),
]


@pytest.mark.parametrize("name,chw,spec,B,ckw", PACKED4_CASES, ids=[c[0] for c in PACKED4_CASES])
def test_mma_packed4_weight_tiles(name, chw, spec, B, ckw):
    """TF2B_WEIGHTS_PACKED4: the resident weight slab of the tensor-core kernel arrives as 4-bit codes through TMA and
    every CTA expands it on the fly into the swizzled K-major int8 layout (SWIZZLE_64B and SWIZZLE_128B tiles, one
    and two planes).  Same results as the int8 planes and as the oracle, INT32 accumulators included, and the launch
    plan says `packed4`."""
    import torch
    from tf2_b200.network import NetWork, Runner
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    net = nets.chain(chw, [dict(N=chw[0], k=1, relu=1), dict(spec)], name)
    x = H.random_input(rng, *chw, nonneg=True, B=B)
    old = H.random_codes
    try:
        if ckw:
            H.random_codes = lambda rng_, N, C, k, **kw: old(rng_, N, C, k, **ckw)
        model = H.random_model(net, rng, x)
    finally:
        H.random_codes = old
    outs = {}
    for staging in (capi.WEIGHTS_PLANES, capi.WEIGHTS_PACKED4):
        nw = NetWork(net, 0)
        nw.set_weight_staging(staging)
        nw.InitFromCodes(model, None, max_images=B, variant=capi.VARIANT_MMA)
        plan = nw.layer_modes(B)[1]
        assert ("packed4" in plan) == (staging == capi.WEIGHTS_PACKED4), plan
        assert "wres" in plan or "halo" in plan, plan
        r = Runner(nw)
        outs[staging] = r.run_device(torch.from_numpy(x).cuda()).cpu().numpy()
        if staging == capi.WEIGHTS_PACKED4:
            g = r.dump_acc(1, B).cpu().numpy()
            for b in range(B):
                tens, accs = H.oracle_tensors(net, model, x[b])
                assert np.array_equal(outs[staging][b], tens[net.result_tensor()]), f"{name} [{plan}]: image {b} differs from the oracle"
                assert np.array_equal(g[b], accs[1]), f"{name} [{plan}]: image {b} INT32 accumulators differ"
        nw.CleanUp()
    assert np.array_equal(outs[capi.WEIGHTS_PLANES], outs[capi.WEIGHTS_PACKED4])
