"""Host logic of the product library on the CPU.  tests/csrc/host_logic_test.cu #includes the shipped
tf2_b200/csrc/api.cu and calls its weight preparation directly (no GPU, no CUDA call): every weight is rebuilt
from the prepared planes of BOTH kernel families and compared with the LoadModel code it came from; the same
layers loaded through tf2b_load_layer_packed4 (4-bit nibbles + Q rows) must leave the identical prepared state;
the range analysis must pick the epilogue form this file predicts.  Needs nvcc and the library's object files
(python -c "import __graft_entry__ as g; g.build()"); skipped otherwise."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from tests import helpers as H
from tests.conftest import GOLDEN, ROOT
from tf2_b200 import formats, nets, synth
from tf2_b200.network import _layer_descs

SRC = os.path.join(ROOT, "tests", "csrc", "host_logic_test.cu")
EXE = os.path.join(ROOT, "tests", "_build", "host_logic_test")
OBJS = [os.path.join(ROOT, "tf2_b200", "lib", f"{n}.o") for n in ("conv_mma", "conv_sa", "aux_kernels")]
DEPS = [SRC, os.path.join(ROOT, "tf2_b200", "csrc", "api.cu"), os.path.join(ROOT, "tf2_b200", "csrc", "common.cuh"),
        os.path.join(ROOT, "include", "tf2b200.h")] + OBJS


def _driver():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc) or not all(os.path.exists(o) for o in OBJS):
        pytest.skip("nvcc or the library's object files are not here")
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in DEPS):
        os.makedirs(os.path.dirname(EXE), exist_ok=True)
        subprocess.check_call([nvcc, "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr",
                               "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "tf2_b200", "csrc"), SRC] + OBJS
                              + ["-o", EXE])
    return EXE


def _case_file(path, net, model, packed=None):
    """packed: per layer None or (nib [N][C][k][k], min_exp, q_in, q_out)"""
    tarr, larr = _layer_descs(net)
    with open(path, "wb") as f:
        f.write(struct.pack("<ii", len(net.tensors), net.num_layers))
        f.write(bytes(tarr))
        for l, ld in enumerate(net.layers):
            f.write(bytes(larr[l]))
            if ld.ipool:
                continue
            codes, params = model[l]
            assert codes.shape == (ld.N, ld.C, ld.k, ld.k)
            f.write(np.ascontiguousarray(codes, dtype=np.uint8).tobytes())
            f.write(np.ascontiguousarray(params, dtype="<i4").tobytes())
            p4 = packed[l] if packed else None
            f.write(struct.pack("<i", 1 if p4 else 0))
            if p4:
                nib, min_exp, q_in, q_out = p4
                f.write(struct.pack("<i", min_exp))
                f.write(formats.nibbles_dense(nib).tobytes())
                f.write(np.ascontiguousarray(q_in, dtype=np.int8).tobytes())
                f.write(np.ascontiguousarray(q_out, dtype=np.int8).tobytes())


def _run(exe, path, batch=256):
    r = subprocess.run([exe, path, str(batch)], capture_output=True, text=True, timeout=600)
    rows = []
    for line in r.stdout.splitlines():
        if not line.startswith("layer "):
            continue
        kv = dict(tok.split("=", 1) for tok in line.split()[2:] if "=" in tok)
        rows.append({k: (v if k == "mode" else int(v)) for k, v in kv.items()})
    return r.returncode, rows, r.stdout


def _packed_from_blob(net, blob, q):
    """Per layer the 4-bit form of a power-of-two param.bin + the Q rows LoadModel would use."""
    buf = memoryview(blob)
    pos = 0
    out = []
    for l, kind, shape in formats._blob_fields(net):
        a, pos = formats._read_f32(buf, pos, int(np.prod(shape)))
        if kind != "weight":
            continue
        ld = net.layers[l]
        while len(out) < l:
            out.append(None)                                  # ipool layers
        if ld.first_layer_7x7:
            out.append(None)                                  # stored as 7x7x3, expanded by filter_trans: not a packed4 case
            continue
        nz = a[a != 0]
        min_exp = int(np.round(np.log2(np.abs(nz).min())))
        nib = formats.weights_to_nibbles(a.reshape(shape), min_exp)
        out.append((nib, min_exp, q[ld.q_in_row][:ld.C], q[ld.q_out_row][:ld.N]))
    while len(out) < net.num_layers:
        out.append(None)
    return out


@pytest.mark.parametrize("name,seed", [("resnet50", 3), ("googlenet", 5)])
def test_shipped_networks_weight_preparation(name, seed, tmp_path):
    exe = _driver()
    net = nets.load(name)
    q = formats.parse_q_file(net, os.path.join(GOLDEN, f"{name}_Q"))
    blob = synth.synth_float_blob(net, seed=seed, q=q)
    model = formats.load_float_blob(net, blob, q)
    packed = _packed_from_blob(net, blob, q)
    for l, p4 in enumerate(packed):                           # the Python mirror of the 4-bit expansion agrees with LoadModel
        if p4:
            assert np.array_equal(formats.codes_from_nibbles(*p4), model[l][0]), l
    path = str(tmp_path / "case.bin")
    _case_file(path, net, model, packed)
    rc, rows, out = _run(exe, path)
    assert rc == 0, out
    conv = [ld for ld in net.layers if not ld.ipool]
    assert len(rows) == len(conv)
    assert all(r["rc"] == 0 and r["bad"] == 0 for r in rows)
    assert sum(r["packed4_same"] == 1 for r in rows) == sum(1 for p in packed if p) >= len(conv) - 1
    assert all(r["mma_ok"] == 1 for r in rows)                # every layer of the shipped nets runs on the tensor cores
    if name == "resnet50":
        assert all(r["fast_requant"] >= 2 for r in rows)              # range analysis: folded epilogue everywhere (bench.py
                                                                      # reports "fold": 53 — conv1's low plane keeps the literal form)
        assert sum(r["fast_requant"] == 3 for r in rows) >= 40        # most layers: every base shift >= 3 -> hi32
        # the launch plan at the BASELINE batch (what tf2b_layer_mode reports on the GPU) is the one the committed
        # bench line was measured with (profiles/r02_bench_resnet50_v4.json: roofline.staging_modes)
        import collections
        import json
        plan = collections.Counter(tok for r in rows for tok in r["mode"].split("_"))
        with open(os.path.join(ROOT, "profiles", "r02_bench_resnet50_v4.json")) as f:
            measured = json.loads(f.read().strip().splitlines()[-1])["roofline"]["staging_modes"]
        assert {k: plan[k] for k in measured} == measured, (dict(plan), measured)
        # stride-1 3x3 layers with C >= 128 on 28 x 28 / 14 x 14 maps: CTA pairs on halo tiles with streamed weights;
        # the 7 x 7 ones (49 of 128 accumulator rows) and the strided ones stay on one box per tap
        assert all("halo_wstream" in rows[l]["mode"] and "ctapair" in rows[l]["mode"] for l in (16, 19, 22, 29, 32, 35, 38, 41))
        assert all("box" in rows[l]["mode"] for l in (13, 26, 45, 48, 51))
        assert rows[0]["mode"] == "mma_BN64_BK64_planes2_halo_wres_stages4"                  # conv1: halo tile, literal epilogue
        assert rows[1]["mode"] == "mma_BN128_BK64_planes2_flat_wres_fold_hi32_tmastore_stages8"
        assert plan["tmastore"] >= 30 and plan["sparse2"] == 0 and plan["packed4"] == 0      # both measured and left off
        assert plan["mma"] == 54 and plan["hi32"] == 53
        assert rows[0]["low"] >= 0                                    # conv1's code-0 taps sit in the unscaled low plane


def test_random_layers_weight_preparation(tmp_path):
    """Quirk layers (input may hold -128) on tensor 0 and elsewhere, wide shift ranges (several planes), low
    codes, 5x5, ragged channel counts, pixel-pair geometry, and parameters that defeat the range analysis."""
    exe = _driver()
    rng = np.random.default_rng(31)
    specs = [dict(N=24, k=3, pad=1, relu=0), dict(N=40, k=1, relu=1), dict(N=64, k=5, pad=2), dict(N=64, k=3, pad=1),
             dict(N=100, k=1), dict(N=16, k=3, pad=1, stride=2)]
    net = nets.chain((3, 20, 20), specs, "hostlogic")
    x = H.random_input(rng, 3, 20, 20, nonneg=False)
    model = H.random_model(net, rng, x[None])
    # widen layer 1's shift range to three 7-level planes and give layer 2 low absolute shifts next to high ones
    c1, p1 = model[1]
    c1 = c1.copy()
    c1[0, 0, 0, 0], c1[0, 1, 0, 0] = 0x01, 0x13
    model[1] = (c1, p1)
    c4, p4 = model[4]
    p4 = p4.copy()
    p4[:, 1] = 2 ** 31 - 1                                     # alpha so large that a + beta can wrap: literal epilogue
    model[4] = (c4, p4)
    path = str(tmp_path / "case.bin")
    _case_file(path, net, model)
    rc, rows, out = _run(exe, path)
    assert rc == 0, out
    assert all(r["rc"] == 0 and r["bad"] == 0 for r in rows), out
    assert net.layers[0].in_may_be_m128 == 1 and rows[0]["segs_s"] >= 1            # tensor 0: negated copy = extra channels
    assert net.layers[1].in_may_be_m128 == 1 and rows[1]["mma_ok"] == 0            # -128 beyond tensor 0: exact kernel only
    assert rows[1]["segs_s"] >= 4                                                  # three exponent levels x (plain, negating)
    assert rows[4]["fast_requant"] == 0
    assert all(r["mma_ok"] == 1 for i, r in enumerate(rows) if i != 1)


def test_every_layer_of_the_baseline_networks_is_planned_on_the_tensor_cores(tmp_path):
    """BASELINE configs[3] (VGG16, batch 1024 over 8 GPUs = 128 per GPU) and configs[0] (SqueezeNet, one image):
    full-size tables, random codes — the weight preparation is exact for every layer (25 088-deep fc6 included)
    and no layer falls back to the shift-accumulate kernel.  (ResNet50 and GoogLeNet: see above.)"""
    exe = _driver()
    rng = np.random.default_rng(1)
    for net, batch in ((nets.vgg16(), 128), (nets.squeezenet(), 1)):
        model = []
        for ld in net.layers:
            params = np.zeros((ld.N, 3), np.int32)
            params[:, 1] = 1 << 10
            model.append((H.random_codes(rng, ld.N, ld.C, ld.k), params))
        path = str(tmp_path / f"{net.name}.bin")
        _case_file(path, net, model)
        rc, rows, out = _run(exe, path, batch)
        assert rc == 0, out[-2000:]
        assert len(rows) == net.num_layers and all(r["bad"] == 0 and r["mma_ok"] == 1 for r in rows)
        assert all(r["mode"].startswith("mma_") for r in rows)
        assert net.layers[0].in_may_be_m128 == 1           # the 3-channel stems see -128: negated copy of tensor 0
