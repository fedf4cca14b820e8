"""Golden vectors produced by the reference's own PE kernel (pe.cl compiled as C, see
tests/golden/make_pe_golden.py): the CPU oracle must reproduce them (CPU test) and so must the CUDA
engine through the C ABI (GPU test) — INT8 outputs of PeFunction = requantised INT32 accumulators."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.conftest import GOLDEN
from tf2_b200 import capi, nets


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "pe_golden.npz"))


def _oracle_reduce(x, codes, p, mode):
    o = O.lib()
    steps = x.shape[0]
    res = []
    for w in range(7 if mode == 1 else 5):
        tot = int(p[0])
        for s in range(steps):
            if mode == 1:
                for c in range(16):
                    tot += o.tf2o_mul(int(x[s, w, c]), int(codes[s, c]))
            else:
                for fw in range(3):
                    for c in range(16):
                        tot += o.tf2o_mul(int(x[s, w + fw, c]), int(codes[s, fw, c]))
        tot = (tot + 2 ** 31) % 2 ** 32 - 2 ** 31
        res.append(o.tf2o_requant(tot, int(p[1]), int(p[2])))
    return res


@pytest.mark.parametrize("mode,key", [(1, "t1"), (3, "t3")])
def test_oracle_reproduces_reference_pe(gold, mode, key):
    steps = gold[key + "_steps"]
    xs, cs = gold[key + "_x"], gold[key + "_codes"]
    xo = co = 0
    for t, s in enumerate(steps):
        s = int(s)
        x = xs[xo:xo + s * 112].reshape(s, 7, 16); xo += s * 112
        n = s * 16 * (1 if mode == 1 else 3)
        c = cs[co:co + n].reshape((s, 16) if mode == 1 else (s, 3, 16)); co += n
        exp = [int(v) for v in gold[key + "_y"][t][:7 if mode == 1 else 5]]
        assert _oracle_reduce(x, c, gold[key + "_params"][t], mode) == exp, f"trial {t}"


def _layer_net(gold):
    C = gold["layer_x"].shape[0]
    return nets.chain((C, 1, 7), [dict(N=gold["layer_codes"].shape[0], k=1, relu=0)], "pe_golden")


def test_oracle_layer_matches_reference_pe(gold):
    net = _layer_net(gold)
    y = O.layer_forward(net.layers[0], net.tensors[0], gold["layer_x"], gold["layer_codes"], gold["layer_params"])
    assert np.array_equal(y, gold["layer_y"])


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [capi.VARIANT_SHIFT, capi.VARIANT_MMA], ids=["shift", "mma"])
def test_cuda_layer_matches_reference_pe(gold, variant):
    import torch
    from tf2_b200.network import NetWork, Runner
    net = _layer_net(gold)
    nw = NetWork(net, 0)
    nw.InitFromCodes([(gold["layer_codes"], gold["layer_params"])], None, max_images=2, variant=variant)
    assert nw.layer_kernels() == ["shift" if variant == capi.VARIANT_SHIFT else "mma"]
    x = np.stack([gold["layer_x"], gold["layer_x"][:, :, ::-1]]).copy()
    got = Runner(nw).run_device(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.array_equal(got[0], gold["layer_y"]), f"differs in {(got[0] != gold['layer_y']).sum()} of {got[0].size}"
    assert np.array_equal(got[1], gold["layer_y"][:, :, ::-1])   # mirrored input -> mirrored output (1x1 conv)
    nw.CleanUp()
