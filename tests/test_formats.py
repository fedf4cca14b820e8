"""CPU: file formats and host numeric preparation against (a) known answers measured on the
reference's own compiled sources, (b) the compiled reference itself when oracle/_ref is present,
(c) the independent C oracle."""
import ctypes as C
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.conftest import GOLDEN, REFERENCE
from tf2_b200 import formats, nets, synth

# Get_real known answers from the reference compiled as-is
GET_REAL_KAT = [(0.5, 15, 0x0e), (-0.25, 12, 0x8a), (0.0, 3, 0x40), (1.0, 3, 0x03), (2.0 ** -14, 10, 0x00),
                (0.3, 15, 0x0f), (-(2.0 ** -6), 20, 0x8e)]


def test_get_real_known_answers():
    for w, e, code in GET_REAL_KAT:
        assert int(formats.get_real(np.float32(w), np.int8(e))) == code
        assert O.lib().tf2o_get_real(float(w), int(e)) == code


def test_get_real_matches_oracle_random():
    rng = np.random.default_rng(1)
    w = np.concatenate([rng.normal(0, 0.2, 3000), np.exp2(-rng.integers(0, 18, 3000)) * rng.choice([-1, 1], 3000),
                        [0, 1e-5, 9.9e-6, 0.99, 1.01, 0.505]]).astype(np.float32)
    e = rng.integers(-5, 30, w.size).astype(np.int8)
    mine = formats.get_real(w, e)
    orc = np.array([O.lib().tf2o_get_real(float(a), int(b)) for a, b in zip(w, e)], dtype=np.uint8)
    assert np.array_equal(mine, orc)


def test_transforms_match_oracle():
    rng = np.random.default_rng(2)
    f = rng.integers(0, 256, 49).astype(np.uint8)
    out_or = np.zeros(81, np.uint8)
    O.lib().tf2o_filter_trans(f.ctypes.data, out_or.ctypes.data)
    assert np.array_equal(formats.filter_trans(f.reshape(7, 7)).reshape(-1), out_or)
    # the 18 taps LoadModel never writes keep the fill value 0 = code of +1
    ft = formats.filter_trans(np.full((7, 7), 0x40, np.uint8))
    assert int((ft == 0).sum()) == 18 and int((ft == 0x40).sum()) == 63
    img = rng.normal(0, 50, (224, 224)).astype(np.float32)
    o_or = np.zeros(9 * 115 * 115, np.float32)
    O.lib().tf2o_feature_trans(img.ctypes.data, o_or.ctypes.data)
    assert np.array_equal(formats.feature_trans(img), o_or.reshape(9, 115, 115)[:, :114, :114])


def test_conv1_transform_is_exact_7x7():
    """3x3 over the 9 derived planes == 7x7/s2/p3 over the raw plane (ignoring the code-0 quirk)."""
    rng = np.random.default_rng(3)
    x = rng.integers(-8, 8, (20 * 2 + 0, 20 * 2)).astype(np.int64)
    x224 = np.zeros((224, 224), np.int64); x224[:40, :40] = x
    w = rng.integers(-3, 4, (7, 7)).astype(np.int64)
    P = np.pad(x224, 3)
    ref = np.array([[(P[2 * i:2 * i + 7, 2 * j:2 * j + 7] * w).sum() for j in range(12)] for i in range(12)])
    d = formats.feature_trans(x224)
    # weights through filter_trans with an out-of-band fill so unwritten taps are identifiable
    wt = formats.filter_trans((w + 10).astype(np.uint8), fill=0xFF).astype(np.int64)
    wt = np.where(wt == 0xFF, 0, np.where(wt == 0x40, 0, wt - 10))
    got = np.array([[sum((d[p, i:i + 3, j:j + 3] * wt[p]).sum() for p in range(9)) for j in range(12)] for i in range(12)])
    assert np.array_equal(ref, got)


def test_quantize_input_matches_oracle():
    rng = np.random.default_rng(4)
    x = np.concatenate([rng.normal(0, 60, 4000), [0.5, -0.5, 1.5, -1.5, 127.5, -128.5, 300, -300]]).astype(np.float32)
    for q0 in (0, -1, -3, 2):
        mine = formats.quantize_input(x, q0)
        orc = np.array([O.lib().tf2o_quantize_input(float(v), q0) for v in x], dtype=np.int8)
        assert np.array_equal(mine, orc)


@pytest.mark.parametrize("name,count,rows", [("resnet50", 27563, 55), ("googlenet", 8283, 77)])
def test_q_file_golden(name, count, rows):
    net = nets.load(name)
    path = os.path.join(GOLDEN, f"{name}_Q")
    assert formats.q_file_value_count(net) == count == len(open(path).read().split())
    q = formats.parse_q_file(net, path)
    assert q.shape == (rows, net.max_out_channel)
    if name == "resnet50":
        assert list(q[0, :3]) == [0, 0, 0] and set(np.unique(-q[1, :64])) == {3, 4, 5}
        assert (q[54, :1000] == -2).all()
    else:
        # ipool rows copy their input row, branch tails fill the concat row (quantization.cpp:42-49)
        l8 = net.layers[8]
        assert l8.ipool and np.array_equal(q[9, :l8.N], q[l8.q_in_row, :l8.N])
        tail = [l for l, t in enumerate(net.branch_tail) if t][0]
        ld = net.layers[tail]
        assert np.array_equal(q[net.num_layers + 1 + net.concat_layer[tail], ld.out_ch0:ld.out_ch0 + ld.N], q[tail + 1, :ld.N])


def test_loader_golden_hashes():
    """codes / BiasBnParam of a seeded synthetic blob: hashes recorded from the reference's own
    LoadModel compiled by oracle/build_ref.sh (tests/golden/make_golden.py)."""
    with open(os.path.join(GOLDEN, "loader_golden.json")) as f:
        gold = json.load(f)
    for name in ("resnet50", "googlenet"):
        net = nets.load(name)
        q = formats.parse_q_file(net, os.path.join(GOLDEN, f"{name}_Q"))
        blob = synth.synth_float_blob(net, seed=gold[name]["seed"], q=q)
        assert hashlib.sha256(blob).hexdigest() == gold[name]["blob_sha256"], "synthetic blob generator changed"
        model = formats.load_float_blob(net, blob, q)
        for l, (codes, params) in enumerate(model):
            if codes is None:
                continue
            assert hashlib.sha256(codes.tobytes()).hexdigest() == gold[name]["codes"][str(l)], f"{name} layer {l} codes"
            assert hashlib.sha256(params.tobytes()).hexdigest() == gold[name]["params"][str(l)], f"{name} layer {l} params"


def test_loader_against_compiled_reference():
    L = O.ref_host_lib("googlenet")
    if L is None:
        pytest.skip("oracle/_ref not built on this box")
    net = nets.load("googlenet")
    qfile = os.path.join(GOLDEN, "googlenet_Q")
    nq, mo = L.ref_num_q_layers(), L.ref_max_out_channel()
    qref = np.zeros((nq + 2) * mo, dtype=np.int8)
    L.ref_quantization(qref.ctypes.data, qfile.encode())
    q = formats.parse_q_file(net, qfile)
    assert np.array_equal(q, qref[:nq * mo].reshape(nq, mo))
    blob = synth.synth_float_blob(net, seed=9, q=q)
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".bin") as tf:
        tf.write(blob); tf.flush()
        stride, nl = L.ref_filter_layer_stride(), net.num_layers
        fr = np.zeros(nl * stride + 1024, np.uint8)
        bb = np.zeros((nl * L.ref_max_bias_size() + 16, 3), np.int32)
        L.ref_load_model(tf.name.encode(), fr.ctypes.data, bb.ctypes.data, qref.ctypes.data)
    model = formats.load_float_blob(net, blob, q)
    for l, (codes, params) in enumerate(model):
        if codes is None:
            continue
        assert np.array_equal(fr[l * stride:l * stride + codes.size], codes.reshape(-1)), f"layer {l} codes"
        assert np.array_equal(bb[l * L.ref_max_bias_size():l * L.ref_max_bias_size() + codes.shape[0]], params), f"layer {l} params"


def test_4bit_blob_roundtrip():
    rng = np.random.default_rng(5)
    for shape in ((24, 40, 1, 1), (16, 8, 3, 3), (8, 4, 5, 5), (4, 3, 7, 7), (6, 5, 2, 2)):
        min_exp = -9
        lv = rng.integers(0, 7, shape)
        w = (np.exp2((min_exp + lv).astype(np.float32)) * rng.choice([-1.0, 1.0], shape)).astype(np.float32)
        w[rng.random(shape) < 0.2] = 0
        nib = formats.weights_to_nibbles(w, min_exp)
        assert np.array_equal(formats.nibbles_to_weights(nib, min_exp), w)
        rec = formats.pack4_layer(nib, min_exp)
        nib2, me2, pos = formats.unpack4_layer(memoryview(rec), 0)
        assert me2 == min_exp and pos == len(rec) and np.array_equal(nib, nib2)
        q_in = rng.integers(-5, 0, shape[1]).astype(np.int8)
        q_out = rng.integers(-5, 0, shape[0]).astype(np.int8)
        codes = formats.codes_from_nibbles(nib, min_exp, q_in, q_out)
        expand = (15 + q_in.astype(np.int32)[None, :] - q_out.astype(np.int32)[:, None]).astype(np.int8)
        assert np.array_equal(codes, formats.get_real(w, expand[:, :, None, None]))


def test_load_input_image_matches_reference(tmp_path):
    """LoadInputImage (input_loader.cpp:76-118: read float [3][224][224], feature_trans each plane, crop
    115 -> 114) compiled unmodified vs load_image_bin + feature_trans — on a seeded synthetic image file and,
    where the reference tree is present, on its two shipped test images."""
    L = O.ref_host_lib("resnet50")
    if L is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    import ctypes as C
    from tf2_b200 import synth
    files = []
    img = synth.synth_images(1, seed=99)[0]
    p = tmp_path / "img.bin"
    img.astype("<f4").tofile(p)
    files.append(str(p))
    for name in ("resnet50_data_label_100.bin", "googlenet_data_label_391.bin"):
        f = os.path.join(REFERENCE, "Runtime_Engine", "cnn", "host", "test_images", name)
        if os.path.exists(f):
            files.append(f)
    L.ref_load_input_image.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p]
    for f in files:
        input_raw = np.full(27 * 114 * 114 + 64, np.float32(7.5), np.float32)
        raw_images = np.zeros(3 * 224 * 224, np.float32)
        L.ref_load_input_image(f.encode(), input_raw.ctypes.data, raw_images.ctypes.data)
        mine = formats.load_image_bin(f)
        assert np.array_equal(mine.reshape(-1), raw_images)
        assert np.array_equal(formats.feature_trans(mine).reshape(-1), input_raw[:27 * 114 * 114])
        assert np.all(input_raw[27 * 114 * 114:] == np.float32(7.5))          # nothing written past the 27x114x114 image
    with pytest.raises(ValueError):
        (tmp_path / "short.bin").write_bytes(b"\0" * 100)
        formats.load_image_bin(str(tmp_path / "short.bin"))


def test_4bit_model_file_roundtrip():
    """Whole-model 4-bit file (4bit_data_format.txt: short-coded weights, float records for everything else):
    param.bin -> 4-bit -> param.bin is the identity on an INQ-grid model, the file is ~7x smaller, and the
    shift codes LoadModel derives from either are the same."""
    for name, seed in (("resnet50", 3), ("googlenet", 5)):
        net = nets.load(name)
        q = formats.parse_q_file(net, os.path.join(GOLDEN, f"{name}_Q"))
        blob = synth.synth_float_blob(net, seed=seed, q=q)
        m4 = formats.float_blob_to_4bit(net, blob)
        assert len(m4) < len(blob) / 6
        back = formats.float_blob_from_4bit(net, m4)
        assert back == blob
        with pytest.raises(ValueError):
            formats.float_blob_from_4bit(net, m4[:-2])
        with pytest.raises(ValueError):
            formats.float_blob_from_4bit(net, m4 + b"\\0\\0")
    # weights off the grid are refused (the packer is not a quantiser: tf2_b200.compress is)
    net = nets.vgg16(width_div=16)
    blob = bytearray(synth.synth_float_blob(net, seed=1))
    blob[0:4] = np.float32(0.3).tobytes()
    with pytest.raises(ValueError):
        formats.float_blob_to_4bit(net, bytes(blob))


def test_every_shipped_q_file_parses_like_the_reference():
    """The four Q files the reference ships (host/model/: resnet50_Q, pytorch_resnet50_q, resnet50_pruned_Q,
    googlenet_Q) through its own Quantization() (quantization.cpp:25-55, compiled unmodified) vs parse_q_file."""
    model_dir = os.path.join(REFERENCE, "Runtime_Engine", "cnn", "host", "model")
    if not os.path.isdir(model_dir) or O.ref_host_lib("resnet50") is None:
        pytest.skip("reference tree / oracle/_ref absent")
    seen = 0
    for fname, name in (("resnet50_Q", "resnet50"), ("pytorch_resnet50_q", "resnet50"), ("resnet50_pruned_Q", "resnet50_pruned"),
                        ("googlenet_Q", "googlenet")):
        path = os.path.join(model_dir, fname)
        L = O.ref_host_lib(name)
        net = nets.load(name)
        nq, mo = L.ref_num_q_layers(), L.ref_max_out_channel()
        qref = np.zeros((nq + 2) * mo, dtype=np.int8)
        L.ref_quantization(qref.ctypes.data, path.encode())
        q = formats.parse_q_file(net, path)
        assert q.shape == (nq, mo) and np.array_equal(q, qref[:nq * mo].reshape(nq, mo)), fname
        with open(path) as f:
            assert len(f.read().split()) == formats.q_file_value_count(net), fname
        seen += 1
    assert seen == 4
