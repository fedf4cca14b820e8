"""CPU: the C oracle against known answers taken from the reference's own device code compiled as
C and against a brute-force NumPy statement of SURVEY.md Appendix A."""
import numpy as np
import pytest

from oracle import oracle as O
from tests import helpers as H
from tf2_b200 import nets


def test_mul_known_answers():
    L = O.lib()
    # reference MUL (pe.cl:27-40) compiled as C: SURVEY.md 8c
    assert L.tf2o_mul(5, 0x03) == 40 and L.tf2o_mul(5, 0x83) == -40 and L.tf2o_mul(5, 0x40) == 0
    assert L.tf2o_mul(-128, 0x14) == -134217728
    assert L.tf2o_mul(127, 0x1f) == -2147483648 and L.tf2o_mul(-1, 0x9f) == -2147483648
    # int8 negate quirk: -(-128) stays -128 (pe.cl:32-34)
    assert L.tf2o_mul(-128, 0x80) == -128 and L.tf2o_mul(-128, 0x83) == -1024


def test_requant_known_answers():
    L = O.lib()
    assert L.tf2o_requant(1000000, 1 << 20, 0) == 31
    assert L.tf2o_requant(-1000000, 1 << 20, 0) == -31         # (-1000000>>14 = -62) -> (-62+1)>>1 = -31
    assert L.tf2o_requant(1 << 30, 1 << 20, 0) == 127 and L.tf2o_requant(-(1 << 30), 1 << 20, 0) == -128
    assert L.tf2o_requant(16384, 1 << 20, 0) == 1 and L.tf2o_requant(16383, 1 << 20, 0) == 0
    assert L.tf2o_requant(0, 12345, 1 << 14) == 1
    assert L.tf2o_gap_finish(49 * 100) == 100 and L.tf2o_gap_finish(-49 * 100) == -100 and L.tf2o_gap_finish(49) == 1


def _ref_layer(ld, tin, X, codes, params, R):
    """Brute-force NumPy statement of SURVEY.md Appendix A (int64 arithmetic, explicit wrap)."""
    N, k, s, pad = ld.N, ld.k, ld.stride, ld.pad
    Xp = np.zeros((tin.C, tin.H + 2 * pad, tin.W + 2 * pad), np.int64)
    Xp[:, pad:pad + tin.H, pad:pad + tin.W] = X
    acc = np.zeros((N, ld.OH, ld.OW), np.int64) + params[:, 0].astype(np.int64)[:, None, None]
    for n in range(N):
        for c in range(ld.C):
            for fh in range(k):
                for fw in range(k):
                    cd = int(codes[n, c, fh, fw])
                    if cd & 0x40:
                        continue
                    f = Xp[c, fh:fh + s * ld.OH:s, fw:fw + s * ld.OW:s]
                    if cd & 0x80:
                        f = np.where(f == -128, -128, -f)
                    acc[n] += f << (cd & 0x1f)
    acc = ((acc + 2 ** 31) % 2 ** 32 - 2 ** 31)
    a = (acc * params[:, 1].astype(np.int64)[:, None, None]) >> 20
    a = ((a + 2 ** 31) % 2 ** 32 - 2 ** 31)
    sres = ((a + params[:, 2].astype(np.int64)[:, None, None] + 2 ** 31) % 2 ** 32 - 2 ** 31)
    y = np.clip(((sres >> 14) + 1) >> 1, -128, 127)
    if ld.relu:
        y = np.maximum(y, 0)
    if ld.pool:
        yp = np.zeros((N, ld.PH, ld.PW), np.int64)
        for ph in range(ld.PH):
            for pw in range(ld.PW):
                m = np.full(N, -128, np.int64)
                for dh in range(3):
                    for dw in range(3):
                        h, w = ph * ld.pool_stride - ld.pool_pad + dh, pw * ld.pool_stride - ld.pool_pad + dw
                        v = y[:, h, w] if (0 <= h < ld.OH and 0 <= w < ld.OW) else np.zeros(N, np.int64)
                        m = np.maximum(m, v)
                yp[:, ph, pw] = m
        y = yp
    if ld.add_tensor >= 0:
        y = np.clip(y + R.astype(np.int64), -128, 127)
        if ld.add_relu:
            y = np.maximum(y, 0)
    if ld.gap:
        S = y.reshape(N, -1).sum(axis=1)
        S = ((S + 2 ** 15) % 2 ** 16 - 2 ** 15)
        y = np.clip((((S * 669) >> 14) + 1) >> 1, -128, 127)
    return y.astype(np.int8), acc.astype(np.int32)


CASES = [
    ((5, 9, 9), dict(N=6, k=3, pad=1), False),
    ((7, 8, 8), dict(N=4, k=1, relu=0), False),
    ((4, 11, 11), dict(N=5, k=3, pad=1, stride=2), True),
    ((3, 10, 10), dict(N=4, k=5, pad=2), False),
    ((6, 12, 12), dict(N=8, k=3, pad=0, pool=1, pool_stride=2, pool_pad=1, PH=5, PW=5), False),
    ((6, 9, 9), dict(N=8, k=1, pool=1, pool_stride=1, pool_pad=1, PH=9, PW=9), True),
]


@pytest.mark.parametrize("chw,spec,nonneg", CASES)
def test_oracle_matches_bruteforce(chw, spec, nonneg):
    rng = np.random.default_rng(11)
    net = nets.chain(chw, [spec])
    ld, tin = net.layers[0], net.tensors[0]
    X = H.random_input(rng, *chw, nonneg=nonneg)
    codes = H.random_codes(rng, ld.N, ld.C, ld.k, shift_lo=0, per_n_offset=6, per_c_offset=6)
    codes[0, 0, 0, 0] = 0x9f  # maximal shift, negative
    params = H.fit_params(rng, ld, tin, X, codes)
    y, acc = O.layer_forward(ld, tin, X, codes, params, want_acc=True)
    y_ref, acc_ref = _ref_layer(ld, tin, X, codes, params, None)
    assert np.array_equal(acc, acc_ref)
    assert np.array_equal(y, y_ref)


def test_oracle_residual_gap_and_network():
    rng = np.random.default_rng(12)
    net = nets.chain((8, 7, 7), [dict(N=12, k=1, relu=0), dict(N=12, k=3, pad=1, src=-1, relu=0, add=0, add_relu=1, gap=1)])
    x = H.random_input(rng, 8, 7, 7, nonneg=False, B=3)
    model = H.random_model(net, rng, x)
    full = O.run_network(net, model, x)
    for b in range(3):
        t1, _ = _ref_layer(net.layers[0], net.tensors[0], x[b], model[0][0], model[0][1], None)
        t2, _ = _ref_layer(net.layers[1], net.tensors[0], x[b], model[1][0], model[1][1], t1)
        assert np.array_equal(full[b].reshape(-1), t2.reshape(-1))


def test_oracle_ipool_concat():
    rng = np.random.default_rng(13)
    net = nets.chain((6, 8, 8), [dict(N=4, k=1, concat=(0, 0, 20)), dict(ipool=1, src=-1), dict(N=16, k=1, concat=(0, 4, 20))])
    # concat offsets in the oracle are channel-granular (the CUDA engine needs multiples of 16)
    x = H.random_input(rng, 6, 8, 8, nonneg=True, B=2)
    model = H.random_model(net, rng, x)
    out = O.run_network(net, model, x)
    tens, _ = H.oracle_tensors(net, model, x[0])
    assert out.shape == (2, 20, 8, 8) and np.array_equal(out[0], tens[net.result_tensor()])
    # ipool = 3x3 / s1 / p1 max with zero outside
    xp = np.pad(x[0].astype(np.int64), ((0, 0), (1, 1), (1, 1)))
    mp = np.max(np.stack([xp[:, i:i + 8, j:j + 8] for i in range(3) for j in range(3)]), axis=0)
    assert np.array_equal(tens[2], mp.astype(np.int8))
