"""GPU parity: single fused layers and small chains through the C ABI vs the CPU oracle.
Bit-exact at the INT32 accumulator and at the INT8 requantised feature map."""
import numpy as np
import pytest

from tests import helpers as H
from tf2_b200 import capi, nets

pytestmark = pytest.mark.gpu

# (name, input CHW, layer specs, nonneg input)
CASES = [
    ("3x3_c64_p1", (64, 14, 14), [dict(N=64, k=3, pad=1)], True),
    ("1x1_c256_n64", (256, 14, 14), [dict(N=64, k=1)], True),
    ("1x1_c64_n256_norelu", (64, 28, 28), [dict(N=256, k=1, relu=0)], True),
    ("conv1_like_quirk", (27, 30, 30), [dict(N=64, k=3, pad=0)], False),
    ("3x3_s2", (128, 28, 28), [dict(N=128, k=3, pad=1, stride=2)], True),
    ("1x1_s2", (256, 28, 28), [dict(N=512, k=1, stride=2, relu=0)], True),
    ("5x5_p2", (16, 14, 14), [dict(N=32, k=5, pad=2)], True),
    ("odd_c40_n24", (40, 13, 13), [dict(N=24, k=1)], True),
    ("odd_c100_n48_3x3", (100, 9, 9), [dict(N=48, k=3, pad=1)], True),
    ("fc_2048_1000", (2048, 1, 1), [dict(N=1000, k=1, relu=0, bias_en=1, bn_en=0)], True),
    ("7x7_map_c512", (512, 7, 7), [dict(N=512, k=3, pad=1)], True),
    ("negative_input_3x3", (32, 12, 12), [dict(N=32, k=3, pad=1, relu=0)], False),
    ("pool_s2_p1", (27, 24, 24), [dict(N=64, k=3, pad=0, pool=1, pool_stride=2, pool_pad=1, PH=11, PW=11)], False),
    ("pool_s2_p0", (64, 28, 28), [dict(N=192, k=3, pad=1, pool=1, pool_stride=2, pool_pad=0, PH=14, PW=14)], True),
    ("pool_s1_p1", (64, 14, 14), [dict(N=32, k=1, pool=1, pool_stride=1, pool_pad=1, PH=14, PW=14)], True),
    ("residual", (64, 14, 14), [dict(N=256, k=1, relu=0), dict(N=64, k=1, src=-1), dict(N=64, k=3, pad=1),
                                dict(N=256, k=1, relu=0, add=0, add_relu=1)], True),
    ("residual_norelu", (32, 10, 10), [dict(N=48, k=1, relu=0), dict(N=48, k=3, pad=1, src=-1, relu=0, add=0, add_relu=0)], True),
    ("gap_with_add", (128, 7, 7), [dict(N=256, k=1, relu=0), dict(N=256, k=1, src=-1, relu=0, add=0, add_relu=1, gap=1)], True),
    ("ipool_concat", (48, 14, 14), [dict(N=32, k=1, concat=(0, 0, 80)), dict(ipool=1, src=-1),
                                    dict(N=48, k=1, concat=(0, 32, 80)), dict(N=16, k=1, src=2)], True),
]


@pytest.mark.parametrize("name,chw,specs,nonneg", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("variant", [capi.VARIANT_SHIFT, capi.VARIANT_AUTO], ids=["shift", "auto"])
def test_layer_parity(name, chw, specs, nonneg, variant):
    import torch
    from tf2_b200.network import NetWork, Runner
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    net = nets.chain(chw, specs, name)
    B = 3
    x = H.random_input(rng, *chw, nonneg=nonneg, B=B)
    model = H.random_model(net, rng, x)
    nw = NetWork(net, device=0)
    nw.InitFromCodes(model, None, max_images=B, variant=variant)
    r = Runner(nw)
    out = r.run_device(torch.from_numpy(x).cuda()).cpu().numpy()
    for b in range(B):
        tens, accs = H.oracle_tensors(net, model, x[b])
        exp = tens[net.result_tensor()]
        assert np.array_equal(out[b], exp), f"{name}: image {b} final int8 differs ({(out[b] != exp).sum()} of {exp.size})"
        if b == 0:
            for t in range(1, len(net.tensors)):
                got = r.read_tensor(t, B).cpu().numpy()[0]
                assert np.array_equal(got, tens[t]), f"{name}: tensor {t} differs"
            for l, acc in accs.items():
                got = r.dump_acc(l, B).cpu().numpy()[0]
                assert np.array_equal(got, acc), f"{name}: layer {l} INT32 accumulators differ ({(got != acc).sum()})"
    nw.CleanUp()


def test_layouts_roundtrip():
    import torch
    from tf2_b200.network import NetWork, Runner
    rng = np.random.default_rng(5)
    net = nets.chain((40, 9, 9), [dict(N=24, k=1)])
    x = H.random_input(rng, 40, 9, 9, nonneg=True, B=2)
    model = H.random_model(net, rng, x)
    nw = NetWork(net, 0)
    nw.InitFromCodes(model, None, max_images=2)
    r = Runner(nw)
    a = r.run_device(torch.from_numpy(x).cuda()).cpu().numpy()
    xh = np.ascontiguousarray(x.transpose(0, 2, 3, 1))
    b = r.run_device(torch.from_numpy(xh).cuda(), in_layout=capi.LAYOUT_HWC, out_layout=capi.LAYOUT_HWC).cpu().numpy()
    assert np.array_equal(a, b.transpose(0, 3, 1, 2))
    c = r.run_host(x)
    assert np.array_equal(a, c)
    nw.CleanUp()


def test_errors_do_not_exit():
    from tf2_b200.network import NetWork, Runner, Tf2bError
    net = nets.chain((16, 8, 8), [dict(N=16, k=1)])
    nw = NetWork(net, 0)
    with pytest.raises(Tf2bError):
        nw._check(nw._lib.tf2b_finalize(nw.handle, 4))  # no weights loaded
    rng = np.random.default_rng(0)
    x = H.random_input(rng, 16, 8, 8, True, B=1)
    nw.InitFromCodes(H.random_model(net, rng, x), None, max_images=1)
    with pytest.raises(Tf2bError):
        Runner(nw).run_host(np.zeros((2, 16, 8, 8), np.int8))  # more images than max_images
    nw.CleanUp()
    # pool.cl / pool_tail.cl know stride 1 and 2 only: anything else is refused when the tables are handed over
    import dataclasses
    bad = nets.chain((16, 12, 12), [dict(N=16, k=1, pool=1, pool_stride=2, pool_pad=0, PH=5, PW=5)])
    bad.layers[0] = dataclasses.replace(bad.layers[0], pool_stride=3)
    with pytest.raises(Tf2bError) as ei:
        NetWork(bad, 0)
    assert "pool stride" in str(ei.value)


def test_weight_blob_roundtrip():
    """The init-time broadcast payload: export from a loaded engine, import into a fresh one built
    from the same tables (no model file), identical results on both kernel families."""
    import torch
    from tf2_b200.network import NetWork, Runner
    rng = np.random.default_rng(21)
    net = nets.chain((64, 14, 14), [dict(N=256, k=1, relu=0), dict(N=64, k=1, src=-1), dict(N=64, k=3, pad=1),
                                    dict(N=256, k=1, relu=0, add=0, add_relu=1)])
    x = H.random_input(rng, 64, 14, 14, nonneg=False, B=2)
    model = H.random_model(net, rng, x)
    a = NetWork(net, 0)
    a.InitFromCodes(model, None, max_images=2)
    blob = torch.empty(a.weight_blob_bytes(), dtype=torch.uint8, device="cuda")
    a.export_weight_blob(blob.data_ptr())
    ya = Runner(a).run_device(torch.from_numpy(x).cuda()).cpu().numpy()
    for variant in (capi.VARIANT_AUTO, capi.VARIANT_SHIFT):
        b = NetWork(net, 0)
        b.InitFromBlob(blob.data_ptr(), int(blob.numel()), max_images=2, variant=variant)
        yb = Runner(b).run_device(torch.from_numpy(x).cuda()).cpu().numpy()
        assert np.array_equal(ya, yb)
        if variant == capi.VARIANT_AUTO:
            assert b.layer_kernels() == a.layer_kernels()
        b.CleanUp()
    a.CleanUp()


def test_pipelined_host_api():
    """tf2b_submit_raw224_host / tf2b_wait: same results as the synchronous call, slots reusable."""
    import torch
    from tf2_b200.network import NetWork, Runner
    rng = np.random.default_rng(31)
    net = nets.chain((27, 114, 114), [dict(N=64, k=3, pad=0, pool=1, pool_stride=2, pool_pad=1, PH=56, PW=56),
                                      dict(N=32, k=1, gap=0)])
    B = 3
    raws = [rng.integers(-128, 128, size=(B, 3, 224, 224), dtype=np.int8) for _ in range(5)]
    from tf2_b200 import formats
    t0 = formats.feature_trans(raws[0]).reshape(B, 27, 114, 114)
    model = H.random_model(net, rng, t0)
    nw = NetWork(net, 0)
    nw.InitFromCodes(model, None, max_images=B)
    r = Runner(nw)
    ref = [r.run_host(x, raw224=True).copy() for x in raws]
    pins = [torch.from_numpy(x).pin_memory() for x in raws]
    outs = [torch.empty(ref[0].shape, dtype=torch.int8).pin_memory() for _ in raws]
    for i in range(len(raws)):
        r.submit_host(pins[i], outs[i], i % 2)
        if i >= 1:
            r.wait((i - 1) % 2)
    r.wait((len(raws) - 1) % 2)
    for i in range(len(raws)):
        assert np.array_equal(outs[i].numpy(), ref[i]), f"batch {i}"
    nw.CleanUp()
