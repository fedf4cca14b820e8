"""Verify() / Evaluation() and the feature_ddr tile layout they read, against the reference's own
network_helper.cpp compiled unmodified (oracle/_ref/libtf2ref_host_<net>.so) — live when the build
container has it, and against tests/golden/eval_golden.json (made by tests/golden/make_eval_golden.py)."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

from oracle import oracle
from tf2_b200 import formats, nets
from tf2_b200.network import Evaluation, Verify

_GOLD = os.path.join(os.path.dirname(__file__), "golden", "eval_golden.json")
_MAX_OUT = {"resnet50": 2048, "googlenet": 1024, "resnet50_pruned": 2048}     # <net>.h: MAX_OUT_CHANNEL


def _last(net_name):
    net = getattr(nets, net_name)()
    ld = net.layers[-1]
    return net, ld


def _ref_buffers(net_name, fmap, q, image):
    """feature_ddr image + the q array exactly as runner.cpp hands them to Verify / Evaluation."""
    L = oracle.ref_host_lib(net_name)
    L.ref_output_offset.restype = C.c_longlong
    L.ref_last_ddr_write_base.restype = C.c_longlong
    off, base, nl = L.ref_output_offset(), L.ref_last_ddr_write_base(), L.ref_num_layer()
    tiles = formats.to_device_layout(fmap)
    ddr = np.full(off * (1 + image) + base + tiles.size + 4096, 99, np.int8)       # 99: reads outside the map show up
    ddr[off * (1 + image) + base: off * (1 + image) + base + tiles.size] = tiles
    qa = np.full((nl + 1) * _MAX_OUT[net_name], 77, np.int8)
    qa[nl * _MAX_OUT[net_name]: nl * _MAX_OUT[net_name] + q.size] = q
    return L, ddr, qa


def ref_evaluation(net_name, x, q, image=0):
    L, ddr, qa = _ref_buffers(net_name, np.asarray(x, np.int8).reshape(-1, 1, 1), q, image)
    top = np.zeros(5, np.int32)
    L.ref_evaluation.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_evaluation(image, qa.ctypes.data, ddr.ctypes.data, top.ctypes.data)
    return [int(v) for v in top]


_LINE = re.compile(r"error=(\S+) expect1=(\S+) q=(-?\d+) expect_trans=(\S+) output=(\S+) addr=(\d+) n=(\d+) h=(\d+) w=(\d+)")


def ref_verify_lines(net_name, fmap, expect, q, image, tmp):
    L, ddr, qa = _ref_buffers(net_name, fmap, q, image)
    fn = os.path.join(str(tmp), "expect.bin")
    np.asarray(expect, "<f4").tofile(fn)
    L.ref_verify.argtypes = [C.c_int, C.c_char_p, C.c_void_p, C.c_void_p]
    cwd = os.getcwd()
    os.chdir(str(tmp))
    try:
        L.ref_verify(image, fn.encode(), qa.ctypes.data, ddr.ctypes.data)
    finally:
        os.chdir(cwd)
    rows = []
    with open(os.path.join(str(tmp), f"Lastconv{image}.dat")) as f:
        for line in f:
            m = _LINE.match(line)
            rows.append({"error": float(m[1]), "expect": float(m[2]), "q": int(m[3]), "expect_trans": float(m[4]),
                         "output": float(m[5]), "addr": int(m[6]), "n": int(m[7]), "h": int(m[8]), "w": int(m[9])})
    return rows


def eval_cases():
    """(name, net, int8 logits, Q row).  Q <= 0 only: `1 << (-q)` with q > 0 is undefined in C."""
    rng = np.random.default_rng(20240917)
    for net_name in ("resnet50", "googlenet"):
        n = 1000
        yield f"{net_name}_random", net_name, rng.integers(-128, 128, n).astype(np.int8), rng.integers(-3, 1, n).astype(np.int8)
        x = rng.integers(-20, 20, n).astype(np.int8)
        x[[5, 17, 400, 999]] = 21                                              # four-way tie for rank 0
        x[[3, 998]] = 20
        yield f"{net_name}_ties", net_name, x, np.zeros(n, np.int8)
        yield f"{net_name}_all_equal", net_name, np.full(n, -7, np.int8), np.full(n, -1, np.int8)
        x = rng.integers(-128, 128, n).astype(np.int8)
        q = rng.integers(-2, 1, n).astype(np.int8)                            # equal features from different (x, q)
        x[10], q[10], x[20], q[20], x[30], q[30] = 127, -2, 127, -2, 127, -2
        yield f"{net_name}_mixed_q", net_name, x, q


def verify_case(net_name):
    """The shipped networks all end in a 1x1 map; the tile arithmetic for wide maps is covered by
    test_device_layout_* below and, through feature_writer.cl itself, by tests/test_post_golden.py."""
    net, ld = _last(net_name)
    rng = np.random.default_rng(7 + len(net_name))
    fmap = rng.integers(-128, 128, (ld.N, ld.PH if not ld.gap else 1, ld.PW if not ld.gap else 1)).astype(np.int8)
    q = rng.integers(-3, 1, ld.N).astype(np.int8)
    expect = (fmap.astype(np.float32) * np.exp2(q.astype(np.float32)).reshape(-1, 1, 1)
              + rng.normal(0, 0.4, fmap.shape).astype(np.float32) * np.exp2(q.astype(np.float32)).reshape(-1, 1, 1))
    return fmap, expect.astype(np.float32), q


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(64, 112, 112), (1000, 1, 1), (24, 7, 7), (100, 13, 13), (16, 14, 14), (3, 8, 15)])
def test_device_layout_roundtrip_and_addresses(shape):
    rng = np.random.default_rng(sum(shape))
    fmap = rng.integers(-128, 128, shape).astype(np.int8)
    buf = formats.to_device_layout(fmap)
    C_, H, W = shape
    wv = -(-W // 7)
    assert buf.size == -(-C_ // 16) * H * wv * 128
    np.testing.assert_array_equal(formats.from_device_layout(buf, *shape), fmap)
    for _ in range(200):                                                       # network_helper.cpp:107-118
        n, h, w = rng.integers(0, C_), rng.integers(0, H), rng.integers(0, W)
        addr = (n // 16) * H * wv * 128 + h * wv * 128 + (w // 7) * 128 + (w % 7) * 16 + n % 16
        assert buf[addr] == fmap[n, h, w]
    mask = np.ones(buf.size, bool)                                             # everything else is zero
    nn, hh, ww = np.meshgrid(np.arange(C_), np.arange(H), np.arange(W), indexing="ij")
    mask[(nn // 16) * H * wv * 128 + hh * wv * 128 + (ww // 7) * 128 + (ww % 7) * 16 + nn % 16] = False
    assert not buf[mask].any()


def test_evaluation_golden():
    gold = json.load(open(_GOLD))["evaluation"]
    seen = 0
    for name, net, x, q in eval_cases():
        got = Evaluation(x, q)
        assert [l for l, _ in got] == gold[name], name
        p = np.array([p for _, p in got])          # float sum_exp overflows to inf for features > 88: p = 0, as there
        assert np.all(np.diff(p) <= 0) and np.all(p >= 0) and p.sum() <= 1 + 1e-6
        seen += 1
    assert seen == len(gold)


def test_evaluation_probabilities():
    rng = np.random.default_rng(3)
    x = rng.integers(-128, 128, 1000).astype(np.int8)
    q = rng.integers(-3, 0, 1000).astype(np.int8)                              # features <= 63.5: exp fits a float
    feat = x.astype(np.float64) * np.exp2(q.astype(np.float64))
    sm = np.exp(feat - feat.max())
    sm /= sm.sum()
    for label, p in Evaluation(x, q):
        assert p == pytest.approx(sm[label], rel=1e-4)


@pytest.mark.skipif(oracle.ref_host_lib("resnet50") is None or not hasattr(oracle.ref_host_lib("resnet50"), "ref_evaluation"),
                    reason="compiled reference not built (oracle/build_ref.sh needs /root/reference)")
def test_evaluation_against_reference(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)                                                # Evaluation() writes Lastconv.dat in the cwd
    for name, net, x, q in eval_cases():
        for image in (0, 3):
            assert [l for l, _ in Evaluation(x, q)] == ref_evaluation(net, x, q, image), (name, image)


def test_verify_golden():
    gold = json.load(open(_GOLD))["verify"]
    for net_name, g in gold.items():
        fmap, expect, q = verify_case(net_name)
        assert fmap.size == g["n"]
        assert Verify(fmap, expect, q) == pytest.approx(g["sum_error"] / g["sum_expect"], rel=1e-4)
    fmap, expect, q = verify_case("resnet50")
    assert Verify(fmap, fmap.astype(np.float32) * np.exp2(q.astype(np.float32)).reshape(-1, 1, 1), q) == 0.0


@pytest.mark.skipif(oracle.ref_host_lib("resnet50") is None or not hasattr(oracle.ref_host_lib("resnet50"), "ref_verify"),
                    reason="compiled reference not built (oracle/build_ref.sh needs /root/reference)")
@pytest.mark.parametrize("net_name", ["resnet50", "googlenet", "resnet50_pruned"])
def test_verify_against_reference(net_name, tmp_path):
    fmap, expect, q = verify_case(net_name)
    image = 2
    rows = ref_verify_lines(net_name, fmap, expect, q, image, tmp_path)
    assert len(rows) == fmap.size
    L = oracle.ref_host_lib(net_name)
    L.ref_output_offset.restype = C.c_longlong
    off = L.ref_output_offset() * (1 + image)
    tiles = formats.to_device_layout(fmap)
    err = tot = 0.0
    for r in rows:                                                             # every element read where we put it
        assert tiles[r["addr"] - off] == fmap[r["n"], r["h"], r["w"]] == int(r["output"])
        assert r["q"] == q[r["n"]]
        err += r["error"]
        tot += abs(r["expect_trans"])
    assert Verify(fmap, expect, q) == pytest.approx(err / tot, rel=1e-4)
