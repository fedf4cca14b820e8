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
