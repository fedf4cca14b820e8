"""INQ power-of-two grid (tf2_b200/compress.py) against the reference's own ComputeQuantumRange /
ShapeIntoTwoPower executed from source (tests/golden/make_compress_golden.py), and through the rest of the
weight path: grid -> 4-bit nibbles -> packed blob -> shift codes."""
import hashlib
import json
import os

import numpy as np
import pytest

from tf2_b200 import compress as Z
from tf2_b200 import formats

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "compress_golden.json")


def golden_inputs():
    rng = np.random.default_rng(77)
    cases = []
    for shape, scale in (((16, 8, 3, 3), 0.05), ((32, 64, 1, 1), 0.2), ((10, 3, 3, 3), 1.3), ((8, 4, 5, 5), 0.004)):
        cases.append(((rng.standard_normal(shape) * scale).astype(np.float32), [1.0]))                 # one shot
    w = (rng.standard_normal((24, 16, 3, 3)) * 0.07).astype(np.float32)
    cases.append((w, [0.5, 0.75, 0.875, 1.0]))                                                          # the INQ schedule
    cases.append(((rng.laplace(0, 0.03, (12, 12, 1, 1))).astype(np.float32), [0.3, 0.6, 1.0]))
    return cases


def test_grid_equals_the_reference():
    with open(GOLDEN) as f:
        gold = json.load(f)
    for (w, portions), steps in zip(golden_inputs(), gold):
        w = w.astype(np.float64).copy()
        mask = np.ones(w.shape, np.int8)
        prev = 0.0
        for cur, g in zip(portions, steps):
            mx, mn = Z.quantum_range(w, mask)
            assert (mx, mn) == (g["max_exp"], g["min_exp"])
            w, mask = Z.shape_into_two_power(w, mask, prev, cur, mx, mn)
            assert int((mask == 0).sum()) == g["n_quantized"]
            assert hashlib.sha256(np.ascontiguousarray(mask, dtype=np.int8).tobytes()).hexdigest() == g["mask_sha256"]
            assert hashlib.sha256(np.ascontiguousarray(w, dtype=np.float64).tobytes()).hexdigest() == g["sha256"]
            prev = cur
        assert not mask.any()


def test_one_shot_layer_goes_through_the_4bit_path():
    rng = np.random.default_rng(5)
    w = (rng.standard_normal((16, 24, 3, 3)) * 0.06).astype(np.float32)
    q, min_exp = Z.quantize_layer(w)
    nz = q != 0
    e = np.log2(np.abs(q[nz]))
    assert np.array_equal(e, np.round(e)) and e.min() >= min_exp and e.max() <= min_exp + 6
    assert np.all(np.sign(q[nz]) == np.sign(w[nz]))
    # floor(log2(4|w|/3)): every weight lands on the nearest grid point in the INQ sense (ratio in [3/4, 3/2))
    r = np.abs(w[nz]) / np.abs(q[nz])
    assert r.min() >= 0.75 - 1e-6 and r.max() < 1.5 + 1e-6
    assert np.all(np.abs(w[~nz]) < 0.75 * 2.0 ** min_exp + 1e-12)           # what fell below the smallest level
    nib = formats.weights_to_nibbles(q, min_exp)
    assert np.array_equal(formats.nibbles_to_weights(nib, min_exp), q)
    rec = formats.pack4_layer(nib, min_exp)
    nib2, me2, pos = formats.unpack4_layer(memoryview(rec), 0)
    assert me2 == min_exp and pos == len(rec) and np.array_equal(nib, nib2)
    assert Z.quantize_layer(np.zeros((4, 4, 1, 1), np.float32))[0].sum() == 0


def test_quantum_range_rejects_a_late_outlier():
    w = np.array([0.5, 0.25, 3.0], np.float64)
    with pytest.raises(ValueError):
        Z.quantum_range(w, np.array([0, 0, 1]))


def test_float_model_to_served_int8_model():
    """An arbitrary float param.bin -> INQ grid (quantize_blob) -> calibrated Q -> LoadModel -> oracle: the
    whole offline path of TransForm_Kit in one go; the INT8 network tracks the float network it came from."""
    from oracle import oracle as O
    from tf2_b200 import calibrate as K
    from tf2_b200 import nets, synth
    net = nets.vgg16(width_div=16)
    base = synth.synth_float_blob(net, seed=9)
    # knock the synthetic power-of-two weights off the grid: a float model as a training framework would save it
    rng = np.random.default_rng(4)
    arr = np.frombuffer(base, dtype="<f4").copy()
    pos = 0
    for ld in net.layers:
        cnt = ld.N * ld.C * ld.k * ld.k
        arr[pos:pos + cnt] *= rng.uniform(0.8, 1.3, cnt).astype(np.float32)
        pos += cnt + (ld.N if ld.bias_en else 0) + ((4 * ld.N + 1) if ld.bn_en else 0)
    assert pos == arr.size
    float_blob = arr.astype("<f4").tobytes()
    qblob, min_exps = Z.quantize_blob(net, float_blob)
    assert len(qblob) == len(float_blob) and len(min_exps) == net.num_layers
    qa = np.frombuffer(qblob, dtype="<f4")
    pos = 0
    for ld, me in zip(net.layers, min_exps):
        cnt = ld.N * ld.C * ld.k * ld.k
        w = qa[pos:pos + cnt]
        nib = formats.weights_to_nibbles(w.reshape(ld.N, ld.C, ld.k, ld.k), me)        # raises unless on the 7-level grid
        assert np.array_equal(formats.nibbles_to_weights(nib, me).reshape(-1), w)
        pos += cnt + (ld.N if ld.bias_en else 0) + ((4 * ld.N + 1) if ld.bn_en else 0)
    imgs = synth.synth_images(2, seed=8)
    qtext, _ = K.calibrate(net, qblob, imgs)
    q = formats.parse_q_text(net, qtext)
    model = formats.load_float_blob(net, qblob, q)
    y = O.run_network(net, model, formats.quantize_input(imgs, int(q[0, 0])))
    deq = y.reshape(2, -1).astype(np.float64) * np.exp2(q[net.num_layers, :1000].astype(np.float64))
    ref_q = K.float_forward(net, qblob, imgs)[0][net.result_tensor()].reshape(2, -1)
    ref_f = K.float_forward(net, float_blob, imgs)[0][net.result_tensor()].reshape(2, -1)
    for b in range(2):
        assert np.corrcoef(ref_q[b], deq[b])[0, 1] > 0.99         # INT8 engine arithmetic vs the float net on the grid (0.9997)
        assert np.corrcoef(ref_f[b], deq[b])[0, 1] > 0.8          # ... and vs the original float model, no retraining (0.94)
    with pytest.raises(ValueError):
        Z.quantize_blob(net, float_blob[:-4])
