"""Power-of-two weight grid of the reference's INQ compression (TransForm_Kit/Compression/compress_net/
core/compress_core.py) — the producer side of the 4-bit weight format.

The reference quantises weights INCREMENTALLY during fine-tuning (portions 0 -> 1 with retraining in
between, compress_train_eval.py:26-162; training is out of scope here).  What the runtime needs from it
is the grid itself, restated:

  * `quantum_range(w, mask)`      — ComputeQuantumRange (compress_core.py:19-51): the layer's largest
                                    exponent and the smallest of its 7 levels;
  * `shape_into_two_power(...)`   — ShapeIntoTwoPower (compress_core.py:53-103): move the largest
                                    `current_portion` of the not-yet-quantised weights onto the grid;
  * `quantize_layer(w)`           — both in one shot (portion 0 -> 1): every weight becomes 0 or
                                    +-2^e, min_exp <= e <= max_exp, ready for formats.weights_to_nibbles.

Host-side tooling in numpy; nothing here is on the inference hot path."""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np

NUM_QUANTUM_VALUES = 7          # 4-bit code: sign + 7 magnitude levels (4bit_data_format.txt)
_ALL_QUANTIZED = -100           # compress_core.py:47


def quantum_range(w: np.ndarray, mask: np.ndarray, num_quantum_values: int = NUM_QUANTUM_VALUES) -> Tuple[int, int]:
    """compress_core.py:19-51.  mask 1 = still float, 0 = already on the grid."""
    w = np.abs(np.asarray(w, dtype=np.float64).reshape(-1))
    mask = np.asarray(mask).reshape(-1)
    if not np.all((mask == 0) | (mask == 1)):
        raise ValueError("mask value is not 0, nor 1")
    todo, done = w[mask == 1], w[mask == 0]
    # the reference counts how often the running maximum of the quantised part moved (`updated`); only
    # "0", "all" and "in between" matter, and a quantised part of all zeros behaves like "none yet"
    updated = 0 if done.size == 0 else int(np.count_nonzero(np.maximum.accumulate(done) > np.concatenate(([-np.inf], np.maximum.accumulate(done)[:-1]))))
    if updated == 0:
        max_exp = math.floor(math.log(4.0 * todo.max() / 3.0) / math.log(2.0))
    elif updated < w.size:
        max_exp = round(math.log(done.max()) / math.log(2.0))
        if todo.size and max_exp < math.floor(math.log(4.0 * todo.max() / 3.0) / math.log(2.0)):
            raise ValueError("a weight still to be quantised is larger than the grid of the quantised ones")
    else:
        max_exp = _ALL_QUANTIZED
    return int(max_exp), int(max_exp - num_quantum_values + 1)


def shape_into_two_power(w: np.ndarray, mask: np.ndarray, previous_portion: float, current_portion: float,
                         max_exp: int, min_exp: int) -> Tuple[np.ndarray, np.ndarray]:
    """compress_core.py:53-103.  Returns (weights, mask) with the largest weights moved onto the grid."""
    if current_portion == 0 or max_exp == _ALL_QUANTIZED:
        return w, mask
    shape = np.shape(w)
    param = np.array(w, dtype=np.float64).reshape(-1)
    m = np.array(mask).reshape(-1)
    todo = m == 1
    n_todo = int(todo.sum())
    n_init = round(float(n_todo) / (1.0 - previous_portion))
    n_keep = round(float(n_init) * (1.0 - current_portion))
    if n_todo - n_keep > 0:
        thr = np.sort(np.abs(param[todo]))[n_keep]
        sel = todo & (np.abs(param) >= thr)
        if thr == 0:                                   # log(0): the reference fails here; zeros stay zeros
            sel &= param != 0
        e = np.floor(np.log(4.0 * np.abs(param[sel]) / 3.0) / math.log(2.0))
        param[sel] = np.where(e >= min_exp, np.sign(param[sel]) * np.exp2(e), 0.0)
        m[sel] = 0
    return param.reshape(shape).astype(np.asarray(w).dtype, copy=False), m.reshape(shape)


def quantize_layer(w: np.ndarray, num_quantum_values: int = NUM_QUANTUM_VALUES) -> Tuple[np.ndarray, int]:
    """One-shot INQ (portion 0 -> 1, no retraining): (weights on the grid, min_exp)."""
    w = np.asarray(w, dtype=np.float32)
    if not np.any(w):
        return w.copy(), 0
    mask = np.ones(w.shape, np.int8)
    max_exp, min_exp = quantum_range(w, mask, num_quantum_values)
    q, _ = shape_into_two_power(w, mask, 0.0, 1.0, max_exp, min_exp)
    return q.astype(np.float32), min_exp


def quantize_blob(net, blob: bytes, num_quantum_values: int = NUM_QUANTUM_VALUES):
    """A float `param.bin` (blob order of model_loader.cpp:154-231) with arbitrary float convolution weights ->
    (the same blob with every layer's weights on its own INQ grid, list of per-layer min_exp; None for ipool
    layers).  Biases and BatchNorm / Scale parameters pass through untouched.  The result loads through
    LoadModel (`Get_real` needs exact powers of two) and packs into the 4-bit format."""
    buf = memoryview(blob)
    out = bytearray(blob)
    pos = 0
    min_exps = []
    for ld in net.layers:
        if ld.ipool:
            min_exps.append(None)
            continue
        if ld.first_layer_7x7:
            cnt = ld.N * net.input_c * 49
        else:
            cnt = ld.N * ld.C * ld.k * ld.k
        end = pos + 4 * cnt
        if end > len(buf):
            raise ValueError("model blob too short")
        w = np.frombuffer(buf[pos:end], dtype="<f4")
        q, me = quantize_layer(w, num_quantum_values)
        out[pos:end] = q.astype("<f4").tobytes()
        min_exps.append(me)
        pos = end
        if ld.bias_en:
            pos += 4 * ld.N
        if ld.bn_en:
            pos += 4 * (4 * ld.N + 1)
    if pos != len(buf):
        raise ValueError(f"model blob has {len(buf) - pos} trailing bytes")
    return bytes(out), min_exps
