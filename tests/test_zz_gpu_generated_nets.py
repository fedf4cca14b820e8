"""GPU engine against what the REFERENCE's device program produced for the generated networks of
tests/test_generated_nets_ref.py (SqueezeNet fire modules on even maps, the pool / stride probes): the golden
hashes come from the executed reference, so this is CUDA vs compiled reference without the oracle in between.

The VGG16 case of the same golden file is checked inside tests/test_vgg16.py."""
import numpy as np
import pytest

from tests import helpers as H
from tests.test_generated_nets_ref import CASES, build_case

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]


@pytest.mark.parametrize("case", [c for c in CASES if c != "vgg16_div8"])
def test_engine_equals_reference_device_program(case):
    import torch
    from tf2_b200 import capi
    from tf2_b200.network import NetWork, Runner
    net, model, x = build_case(case)
    for variant in (capi.VARIANT_AUTO, capi.VARIANT_SHIFT):
        nw = NetWork(net, 0)
        nw.InitFromCodes(model, None, max_images=1, variant=variant)
        r = Runner(nw)
        r.run_device(torch.from_numpy(np.ascontiguousarray(x[None])).cuda())
        dev = {t: r.read_tensor(t, 1).cpu().numpy()[0] for t in range(1, len(net.tensors))}
        H.assert_reference_hashes(case, net, dev, golden="generated_nets_golden.json")
        nw.CleanUp()


def test_packed4_entry_point_equals_codes_entry_point():
    """tf2b_load_layer_packed4 (4-bit nibbles + Q rows, expanded on the host side of the C ABI like Get_real)
    gives the same engine as tf2b_load_layer fed with formats.codes_from_nibbles — and both equal the oracle."""
    import torch
    from oracle import oracle as O
    from tf2_b200 import capi, formats, nets
    from tf2_b200.network import NetWork, Runner
    net = nets.chain((16, 14, 14), [dict(N=32, k=3, pad=1), dict(N=24, k=1), dict(N=16, k=5, pad=2)], "packed4")
    rng = np.random.default_rng(44)
    B = 2
    x = H.random_input(rng, 16, 14, 14, nonneg=False, B=B)
    tens = {0: x[0]}
    model, packed = [], []
    for l, ld in enumerate(net.layers):
        tin = net.tensors[ld.in_tensor]
        nib = rng.integers(0, 15, (ld.N, ld.C, ld.k, ld.k)).astype(np.uint8)          # 7 = zero, 15 never
        min_exp = int(rng.integers(-12, -5))
        q_in = rng.integers(-4, 1, ld.C).astype(np.int8)
        q_out = rng.integers(-4, 1, ld.N).astype(np.int8)
        codes = formats.codes_from_nibbles(nib, min_exp, q_in, q_out)
        params = H.fit_params(rng, ld, tin, tens[ld.in_tensor], codes)
        tens[ld.out_tensor] = O.layer_forward(ld, tin, tens[ld.in_tensor], codes, params)
        model.append((codes, params))
        packed.append((formats.nibbles_dense(nib), min_exp, q_in, q_out, np.ascontiguousarray(params, dtype=np.int32)))
    exp = O.run_network(net, model, x)
    a = NetWork(net, 0)
    a.InitFromCodes(model, None, max_images=B)
    got_a = Runner(a).run_device(torch.from_numpy(x).cuda()).cpu().numpy()
    b = NetWork(net, 0)
    for l, (dense, min_exp, q_in, q_out, params) in enumerate(packed):
        b._check(b._lib.tf2b_load_layer_packed4(b.handle, l, dense.ctypes.data, min_exp, q_in.ctypes.data, q_out.ctypes.data,
                                                params.ctypes.data))
    b._check(b._lib.tf2b_set_variant(b.handle, capi.VARIANT_AUTO))
    b._check(b._lib.tf2b_finalize(b.handle, B))
    b.max_images = B
    got_b = Runner(b).run_device(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(got_a, exp) and np.array_equal(got_b, exp)
    assert exp.std() > 3
    a.CleanUp()
    b.CleanUp()


def test_set_result_returns_an_inner_tensor():
    """tf2b_set_result: the run calls hand back any tensor of the graph (device and host entry points)."""
    import torch
    from tf2_b200 import nets
    from tf2_b200.network import NetWork, Runner
    net = nets.chain((16, 14, 14), [dict(N=32, k=3, pad=1), dict(N=24, k=1), dict(N=16, k=3, pad=1, stride=2)], "setres")
    rng = np.random.default_rng(8)
    B = 3
    x = H.random_input(rng, 16, 14, 14, nonneg=False, B=B)
    model = H.random_model(net, rng, x)
    nw = NetWork(net, 0)
    nw.InitFromCodes(model, None, max_images=B)
    r = Runner(nw)
    last = r.run_device(torch.from_numpy(x).cuda()).cpu().numpy()
    assert last.shape == (B, 16, 7, 7)
    for t in (1, 2):
        nw.set_result(t)
        td = net.tensors[t]
        got = r.run_device(torch.from_numpy(x).cuda()).cpu().numpy()
        assert got.shape == (B, td.C, td.H, td.W)
        assert np.array_equal(got, r.read_tensor(t, B).cpu().numpy())
        assert np.array_equal(r.run_host(x), got)
        for b in range(B):
            assert np.array_equal(got[b], H.oracle_tensors(net, model, x[b])[0][t])
    nw.set_result(net.result_tensor())
    assert np.array_equal(r.run_device(torch.from_numpy(x).cuda()).cpu().numpy(), last)
    with pytest.raises(Exception):
        nw.set_result(99)
    nw.CleanUp()
