#!/usr/bin/env python
"""Golden vectors from the REFERENCE's PE kernel (Runtime_Engine/cnn/device/src/pe.cl compiled as C
by oracle/build_ref.sh -> oracle/_ref/libtf2ref_pe.so, driven by oracle/ref_device/pe_harness.c).

pe_golden.npz holds
  * layer_*: one 1x1 convolution layer (C=2048 -> N=64 on a 1x7 map, image may hold -128): input,
    LoadModel-style codes, BiasBnParam and the int8 outputs PeFunction produced (its requantised
    accumulators), used by the CPU oracle test and directly by the GPU test;
  * t1_* / t3_*: random 1x1-mode and 3x3-mode reductions of random length.
Run in the build container (needs /root/reference)."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtf2ref_pe.so"))
for f in (L.pe_run_1x1, L.pe_run_3x3):
    f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
rng = np.random.default_rng(2024)
D = L.pe_filter_cache_page_depth()


def fit_alpha(x, codes, mode):
    # scale so outputs are spread over int8 (reference-free estimate of the accumulator spread)
    sh = (codes & 0x1f).astype(np.float64)
    mag = np.where(codes & 0x40, 0.0, np.exp2(sh))
    est = np.sqrt((mag ** 2).sum()) * 74.0 + 1.0
    return int(min(2 ** 31 - 1, max(1, 40.0 * 2 ** 35 / est * rng.uniform(0.5, 1.5))))


out = {}
# ---- one full layer: C = 128 steps x 16 channels, 7 pixels, 64 output channels
steps = D
x = rng.integers(-128, 128, (steps, 7, 16)).astype(np.int8)
x[rng.random(x.shape) < 0.02] = -128
N = 64
codes = np.zeros((N, steps, 16), np.uint8)
params = np.zeros((N, 3), np.int32)
y = np.zeros((N, 7), np.int8)
for n in range(N):
    base = rng.integers(2, 9)
    c = (base + rng.integers(0, 7, (steps, 16)) + rng.integers(0, 3, (1, 16))).astype(np.uint8) & 0x1f
    c |= (rng.random((steps, 16)) < 0.5).astype(np.uint8) << 7
    c[rng.random((steps, 16)) < 0.15] = 0x40
    codes[n] = c
    params[n] = (int(rng.integers(-2 ** 18, 2 ** 18)), fit_alpha(x, c, 1), int(rng.integers(-2 ** 19, 2 ** 19)))
    o = np.zeros(7, np.int8)
    rc = L.pe_run_1x1(steps, x.ctypes.data, np.ascontiguousarray(c).ctypes.data, int(params[n, 0]), int(params[n, 1]),
                      int(params[n, 2]), o.ctypes.data)
    assert rc == 0
    y[n] = o
# tensor layout for the engines: X[C][H=1][W=7] with c = step*16 + lane ; codes[N][C][1][1]
out["layer_x"] = np.ascontiguousarray(x.transpose(0, 2, 1).reshape(steps * 16, 1, 7))
out["layer_codes"] = codes.reshape(N, steps * 16, 1, 1)
out["layer_params"] = params
out["layer_y"] = y.reshape(N, 1, 7)
# ---- random reductions, both PE modes
for mode, key in ((1, "t1"), (3, "t3")):
    xs, cs, ps, ys, ss = [], [], [], [], []
    for t in range(40):
        s = int(rng.integers(1, D + 1))
        xx = rng.integers(-128, 128, (s, 7, 16)).astype(np.int8)
        cshape = (s, 16) if mode == 1 else (s, 3, 16)
        cc = rng.integers(0, 256, cshape).astype(np.uint8)
        if t % 4:
            cc = ((cc & 0xc0) | rng.integers(0, 14, cshape)).astype(np.uint8)
        p = (int(rng.integers(-2 ** 20, 2 ** 20)), fit_alpha(xx, cc, mode), int(rng.integers(-2 ** 19, 2 ** 19)))
        if t % 7 == 0:
            p = (p[0], int(rng.integers(2 ** 28, 2 ** 31 - 1)), p[2])   # large alpha: int64 product / wrap paths
        o = np.zeros(7, np.int8)
        rc = (L.pe_run_1x1 if mode == 1 else L.pe_run_3x3)(s, xx.ctypes.data, cc.ctypes.data, p[0], p[1], p[2], o.ctypes.data)
        assert rc == 0
        xs.append(xx.reshape(-1)); cs.append(cc.reshape(-1)); ps.append(p); ys.append(o.copy()); ss.append(s)
    out[key + "_steps"] = np.array(ss, np.int32)
    out[key + "_x"] = np.concatenate(xs)
    out[key + "_codes"] = np.concatenate(cs)
    out[key + "_params"] = np.array(ps, np.int32)
    out[key + "_y"] = np.stack(ys)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pe_golden.npz"), **out)
print("wrote pe_golden.npz", {k: v.shape for k, v in out.items()})
