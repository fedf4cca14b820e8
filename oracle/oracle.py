"""ctypes front-end of the CPU oracle (oracle/tf2_oracle.c) and of the compiled-reference host
loaders (oracle/_ref/).  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs — never by tf2_b200/."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libtf2oracle.so")


class BiasBn(C.Structure):
    _fields_ = [("bias", C.c_int32), ("alpha", C.c_int32), ("beta", C.c_int32)]


class OLayer(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "C", "IH", "IW", "N", "k", "pad", "stride", "OH", "OW", "relu", "pool", "pool_stride", "pool_pad",
        "PH", "PW", "add", "add_relu", "gap", "ipool")]


_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "tf2_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libtf2oracle.so"])
    return _LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        L.tf2o_mul.argtypes = [C.c_int8, C.c_uint8]
        L.tf2o_mul.restype = C.c_int32
        L.tf2o_get_real.argtypes = [C.c_float, C.c_int8]
        L.tf2o_get_real.restype = C.c_uint8
        L.tf2o_quantize_input.argtypes = [C.c_float, C.c_int]
        L.tf2o_quantize_input.restype = C.c_int8
        L.tf2o_requant.argtypes = [C.c_int32, C.c_int32, C.c_int32]
        L.tf2o_requant.restype = C.c_int8
        L.tf2o_gap_finish.argtypes = [C.c_int32]
        L.tf2o_gap_finish.restype = C.c_int8
        L.tf2o_filter_trans.argtypes = [C.c_void_p, C.c_void_p]
        L.tf2o_feature_trans.argtypes = [C.c_void_p, C.c_void_p]
        L.tf2o_conv_acc.argtypes = [C.POINTER(OLayer), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.tf2o_layer_forward.argtypes = [C.POINTER(OLayer), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
        L.tf2o_run_network.argtypes = [C.c_int, C.POINTER(OLayer)] + [C.c_void_p] * 4 + [C.c_int] + \
            [C.c_void_p] * 8 + [C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.tf2o_run_network.restype = C.c_int
        _lib = L
    return _lib


def olayer(ld, tin) -> OLayer:
    """LayerDesc (+ its input TensorDesc) -> oracle layer struct."""
    o = OLayer()
    o.C, o.IH, o.IW = (tin.C if ld.ipool else ld.C), tin.H, tin.W
    o.N, o.k, o.pad, o.stride, o.OH, o.OW = ld.N, ld.k, ld.pad, ld.stride, ld.OH, ld.OW
    o.relu, o.pool, o.pool_stride, o.pool_pad, o.PH, o.PW = ld.relu, ld.pool, ld.pool_stride, ld.pool_pad, ld.PH, ld.PW
    o.add, o.add_relu, o.gap, o.ipool = (1 if ld.add_tensor >= 0 else 0), ld.add_relu, ld.gap, ld.ipool
    return o


def layer_forward(ld, tin, X, codes, params, R=None, want_acc=False):
    """X int8 [C][IH][IW]; codes uint8 [N][C][k][k]; params int32 [N][3]; R int8 [N][PH][PW]."""
    L = lib()
    o = olayer(ld, tin)
    X = np.ascontiguousarray(X, dtype=np.int8)
    # the oracle reads only ld.C channels of a wider tensor when C < tin.C (never in shipped nets)
    osz = (ld.N,) if ld.gap else (ld.N, ld.PH, ld.PW)
    out = np.zeros(osz, dtype=np.int8)
    acc = np.zeros((ld.N, ld.OH, ld.OW), dtype=np.int32) if want_acc else None
    cptr = np.ascontiguousarray(codes, dtype=np.uint8).ctypes.data if codes is not None else None
    pptr = np.ascontiguousarray(params, dtype=np.int32).ctypes.data if params is not None else None
    if codes is not None:
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        params = np.ascontiguousarray(params, dtype=np.int32)
        cptr, pptr = codes.ctypes.data, params.ctypes.data
    Rc = np.ascontiguousarray(R, dtype=np.int8) if R is not None else None
    L.tf2o_layer_forward(C.byref(o), X.ctypes.data, cptr, pptr, Rc.ctypes.data if Rc is not None else None,
                         out.ctypes.data, acc.ctypes.data if acc is not None else None)
    return (out, acc) if want_acc else out


def run_network(net, model, t0_int8: np.ndarray, result_tensor: Optional[int] = None, n_threads: int = 0):
    """t0_int8: [B][C0][H0][W0] int8 (tensor 0).  Returns int8 [B][C][H][W] of the result tensor."""
    L = lib()
    nl = net.num_layers
    layers = (OLayer * nl)()
    in_idx = np.zeros(nl, np.int32); out_idx = np.zeros(nl, np.int32)
    out_ch0 = np.zeros(nl, np.int32); add_idx = np.zeros(nl, np.int32)
    code_off = np.zeros(nl, np.int64); param_off = np.zeros(nl, np.int64)
    code_chunks: List[np.ndarray] = []; param_chunks: List[np.ndarray] = []
    co = po = 0
    for l, ld in enumerate(net.layers):
        layers[l] = olayer(ld, net.tensors[ld.in_tensor])
        in_idx[l], out_idx[l], out_ch0[l], add_idx[l] = ld.in_tensor, ld.out_tensor, ld.out_ch0, ld.add_tensor
        code_off[l], param_off[l] = co, po
        codes, params = model[l]
        if codes is not None:
            c = np.ascontiguousarray(codes, dtype=np.uint8).reshape(-1)
            p = np.ascontiguousarray(params, dtype=np.int32).reshape(-1, 3)
            code_chunks.append(c); param_chunks.append(p)
            co += c.size; po += p.shape[0]
    codes_all = np.concatenate(code_chunks) if code_chunks else np.zeros(1, np.uint8)
    params_all = np.concatenate(param_chunks) if param_chunks else np.zeros((1, 3), np.int32)
    tC = np.array([t.C for t in net.tensors], np.int32)
    tH = np.array([t.H for t in net.tensors], np.int32)
    tW = np.array([t.W for t in net.tensors], np.int32)
    rt = net.result_tensor() if result_tensor is None else result_tensor
    x = np.ascontiguousarray(t0_int8, dtype=np.int8)
    B = x.shape[0]
    out = np.zeros((B, int(tC[rt]), int(tH[rt]), int(tW[rt])), np.int8)
    rc = L.tf2o_run_network(nl, layers, in_idx.ctypes.data, out_idx.ctypes.data, out_ch0.ctypes.data,
                            add_idx.ctypes.data, len(net.tensors), tC.ctypes.data, tH.ctypes.data, tW.ctypes.data,
                            codes_all.ctypes.data, code_off.ctypes.data, params_all.ctypes.data,
                            param_off.ctypes.data, x.ctypes.data, B, rt, out.ctypes.data, n_threads)
    if rc != 0:
        raise MemoryError("oracle run_network failed")
    return out


# ---- compiled reference host loaders (oracle/_ref/, built by oracle/build_ref.sh) --------------
def ref_host_lib(net_name: str):
    """Returns the ctypes handle of oracle/_ref/libtf2ref_host_<net>.so or None if not built."""
    p = os.path.join(_HERE, "_ref", f"libtf2ref_host_{net_name}.so")
    if not os.path.exists(p):
        return None
    L = C.CDLL(p)
    L.ref_get_real.argtypes = [C.c_float, C.c_char]
    L.ref_get_real.restype = C.c_char
    L.ref_filter_layer_stride.restype = C.c_longlong
    vp = C.c_void_p
    L.ref_filter_trans.argtypes = [vp, vp]
    L.ref_feature_trans.argtypes = [vp, vp]
    L.ref_quantization.argtypes = [vp, C.c_char_p]
    L.ref_load_model.argtypes = [C.c_char_p, vp, vp, vp]
    L.ref_load_input_image.argtypes = [C.c_char_p, vp, vp]
    return L


# ---- compiled reference post-PE kernels (oracle/_ref/libtf2ref_post_<net>.so) --------------------
POST_TABLES = ["N", "OH1", "OW1", "stride", "k", "relu", "pool", "pool_s2", "pool_pad", "PH", "PW", "add",
               "add_relu", "gap", "ipool", "n_start", "n_end", "ddr_wen", "ddr_wbase", "ddr_rbase", "cache_wen",
               "cache_wbase", "nvec", "pwvec"]


class RefPost:
    """relu.cl -> pool.cl -> pool_tail.cl -> feature_writer.cl -> full_size_pool.cl of the reference,
    compiled as C with the network's own tables (oracle/ref_device/post_harness.c)."""

    def __init__(self, net_name: str):
        p = os.path.join(_HERE, "_ref", f"libtf2ref_post_{net_name}.so")
        if not os.path.exists(p):
            raise FileNotFoundError(p)
        L = C.CDLL(p)
        L.post_output_offset.restype = C.c_longlong
        L.post_ddr_bytes.restype = C.c_longlong
        L.post_const.restype = C.c_longlong
        L.post_run.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong,
                                                              C.c_void_p, C.c_void_p]
        self.L = L
        self.n_layers = L.post_num_layers()
        self.tab = [{k: L.post_table(i, l) for i, k in enumerate(POST_TABLES)} for l in range(self.n_layers)]
        self.consts = [L.post_const(i) for i in range(4)]
        self.item_bytes = L.post_item_bytes()
        self.item_off = L.post_item_data_offset()

    def pe_shape(self, l):
        """shape of the PE-output map of layer l: channels, rows (stride applied), columns (stride 1)"""
        t = self.tab[l]
        return (t["n_end"] - t["n_start"], -(-t["OH1"] // t["stride"]), t["OW1"])

    def _run(self, maps, n_feed):
        offs = np.zeros(self.n_layers, np.int64)
        chunks, o = [], 0
        for l in range(self.n_layers):
            m = np.ascontiguousarray(maps[l], dtype=np.int8).reshape(-1)
            offs[l] = o
            chunks.append(m)
            o += m.size
        y = np.concatenate(chunks)
        ddr = np.zeros(self.L.post_ddr_bytes(), np.int8)
        cap = 200000
        cache = np.zeros(cap * self.item_bytes, np.uint8)
        gap = np.zeros(4096 * self.item_bytes, np.uint8)
        nc, ng = C.c_longlong(0), C.c_longlong(0)
        counts = np.zeros(4, np.int64)
        rc = self.L.post_run(n_feed, y.ctypes.data, offs.ctypes.data, ddr.ctypes.data, cache.ctypes.data, cap,
                             C.byref(nc), gap.ctypes.data, 4096, C.byref(ng), counts.ctypes.data)
        if rc != 0:
            raise RuntimeError(f"post_run failed: {rc}")
        ci = cache[: nc.value * self.item_bytes].reshape(nc.value, self.item_bytes)
        gi = gap[: ng.value * self.item_bytes].reshape(ng.value, self.item_bytes)
        return ddr, ci, gi, counts

    def _tiles_to_map(self, tiles, nch, PH, PW):
        """[nvec][PH][pwvec][8][16] tiles (w_inc, n_inc) -> [nch][PH][PW]"""
        nvec, _, pwvec = tiles.shape[:3]
        t = tiles[:, :, :, :7, :]                                  # w_inc 0..6 used (W_VECTOR = 7)
        t = t.transpose(0, 4, 1, 2, 3).reshape(nvec * 16, PH, pwvec * 7)
        return np.ascontiguousarray(t[:nch, :, :PW])

    def run(self, maps):
        """maps[l]: int8 [nch][H][W1] PE outputs of layer l.  Returns (outs, counts): outs[l] is the
        int8 [nch][PH][PW] map the reference produced for layer l ([nch] for global-average layers)."""
        ddr, ci, gi, counts = self._run(maps, self.n_layers)
        data = ci[:, self.item_off:self.item_off + 128].view(np.int8).reshape(-1, 8, 16)
        gdata = gi[:, self.item_off:self.item_off + 128].view(np.int8).reshape(-1, 8, 16)
        outs = [None] * self.n_layers
        pos = gpos = 0
        for l, t in enumerate(self.tab):
            nch = t["n_end"] - t["n_start"]
            PH, PW, nvec, pwvec = t["PH"], t["PW"], t["nvec"], t["pwvec"]
            cnt = nvec * PH * pwvec
            if t["gap"]:
                g = gdata[gpos:gpos + nvec]
                gpos += nvec
                outs[l] = np.ascontiguousarray(g[:, 0, :].reshape(-1)[:nch])
            elif t["cache_wen"]:
                tiles = data[pos:pos + cnt].reshape(nvec, PH, pwvec, 8, 16)
                pos += cnt
                outs[l] = self._tiles_to_map(tiles, nch, PH, PW)
        assert pos == data.shape[0] and gpos == gdata.shape[0], "unconsumed reference items"
        # layers that only go to feature_ddr: read the DDR image right after that layer ran
        oo = self.L.post_output_offset()
        for l, t in enumerate(self.tab):
            if outs[l] is not None:
                continue
            last = l == self.n_layers - 1
            d = ddr if last else self._run(maps, l + 1)[0]
            nch, PH, PW, nvec = t["n_end"] - t["n_start"], t["PH"], t["PW"], t["nvec"]
            pwv = -(-PW // 7)
            base = (t["ddr_wbase"] + (t["n_start"] // 16) * PH * pwv) * 128 + (oo if last else 0)
            tiles = d[base: base + nvec * PH * pwv * 128].reshape(nvec, PH, pwv, 8, 16)
            outs[l] = self._tiles_to_map(tiles, nch, PH, PW)
        return outs, counts


# ---- compiled reference device pipeline for layer 0 (oracle/_ref/libtf2ref_full_<net>.so) -------
def ref_full_layer0(net_name: str, x: np.ndarray, codes: np.ndarray, params: np.ndarray):
    """Runs the reference's own input_reader -> filter_reader -> sequencer -> retriever -> 16 PEs ->
    relu -> pool -> pool_tail -> feature_writer (cnn.cl compiled as C, oracle/ref_device/
    full_harness.c) for layer 0, from device buffers laid out by the reference's InputConvert /
    FilterConvert.  x int8 [C0][H0][W0], codes uint8 [N][C0][k][k], params int32 [N][3].
    Returns (int8 [N][PH][PW], item counts, the reference's cycle constants)."""
    Lh = ref_host_lib(net_name)
    p = os.path.join(_HERE, "_ref", f"libtf2ref_full_{net_name}.so")
    if Lh is None or not os.path.exists(p):
        raise FileNotFoundError(p)
    Lf = C.CDLL(p)
    Lf.full_const.restype = C.c_longlong
    Lh.ref_input_device_size.restype = C.c_longlong
    Lh.ref_filter_device_size.restype = C.c_longlong
    consts = [Lf.full_const(i) for i in range(5)]
    isz, fsz, mb = Lh.ref_input_device_size(), Lh.ref_filter_device_size(), Lh.ref_max_bias_size()
    nl = Lh.ref_num_layer()
    inp_f = np.zeros(isz, np.float32)
    xr = np.ascontiguousarray(x, dtype=np.float32)
    Lh.ref_input_convert.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    Lh.ref_input_convert(xr.ctypes.data, inp_f.ctypes.data, 1)
    inp = inp_f.astype(np.int8)
    fraw = np.full(fsz, 64, np.uint8)
    fraw[:codes.size] = np.ascontiguousarray(codes, dtype=np.uint8).reshape(-1)
    freal = np.full(fsz, 64, np.uint8)
    scratch = np.zeros(fsz, np.uint8)
    Lh.ref_filter_convert.argtypes = [C.c_void_p] * 3
    Lh.ref_filter_convert(scratch.ctypes.data, fraw.ctypes.data, freal.ctypes.data)
    N = codes.shape[0]
    bb = np.zeros((nl * mb, 3), np.int32)
    bb[:N] = params
    ddr = np.zeros(8 << 20, np.int8)
    ib, io = Lf.full_item_bytes(), Lf.full_item_data_offset()
    cap = 8000
    cache = np.zeros(cap * ib, np.uint8)
    nc = C.c_longlong(0)
    counts = np.zeros(8, np.int64)
    Lf.full_run_layer0.argtypes = [C.c_void_p] * 3 + [C.c_longlong] + [C.c_void_p] * 2 + [C.c_longlong, C.c_void_p, C.c_void_p]
    rc = Lf.full_run_layer0(inp.ctypes.data, freal.ctypes.data, bb.ctypes.data, consts[0], ddr.ctypes.data,
                            cache.ctypes.data, cap, C.byref(nc), counts.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"full_run_layer0 failed: {rc}")
    n = nc.value
    data = cache[: n * ib].reshape(n, ib)[:, io:io + 128].view(np.int8)
    return data, counts, consts


# ---- compiled reference device program over ALL layers (oracle/_ref/libtf2ref_net_<net>.so) -------
def ref_run_frames(net_name, xs: np.ndarray, model):
    """`xs` int8 [F][C0][H0][W0]: F images through the reference's device program BACK TO BACK in one run
    (frame_num = F, as Runner::Run does for num_images > 1).  Returns (finals [F] of int8 [N][PH][PW], stats)."""
    return ref_run_network(net_name, xs, model, _frames=True)


def ref_run_network(net_name: str, x: np.ndarray, model, _frames: bool = False):
    """Runs the reference's whole device program (cnn.cl compiled as C, kernels as coroutines:
    oracle/ref_device/net_harness.c) for one image through every layer of a shipped network.
    x int8 [C0][H0][W0] (tensor 0: the transformed, quantised image); model: per layer (codes uint8
    [N][C][k][k], params int32 [N][3]) or (None, None) for ipool layers.
    Returns (per_layer, final, stats): per_layer[l] = int8 [N][PH][PW] the feature writer sent to the
    on-chip cache for layer l (after add / ReLU; [N][1][1] from full_size_pool for end-pool layers; None
    when the layer writes no cache), final = the last layer's map read back from feature_ddr."""
    if isinstance(net_name, tuple):          # (libhost, libnet) of a generated network (ref_device/multi_layer.py)
        Lh, Ln = net_name
    else:
        Lh = ref_host_lib(net_name)
        p = os.path.join(_HERE, "_ref", f"libtf2ref_net_{net_name}.so")
        if Lh is None or not os.path.exists(p):
            raise FileNotFoundError(p)
        Ln = C.CDLL(p)
    Lh.ref_filter_layer_stride.restype = C.c_longlong
    for fn in ("net_ddr_bytes", "net_output_offset", "net_layer_info"):
        getattr(Ln, fn).restype = C.c_longlong
    Lh.ref_input_device_size.restype = C.c_longlong
    Lh.ref_filter_device_size.restype = C.c_longlong
    nl = Ln.net_num_layer()
    assert nl == len(model)
    isz, fsz, mb, stride = Lh.ref_input_device_size(), Lh.ref_filter_device_size(), Lh.ref_max_bias_size(), Lh.ref_filter_layer_stride()
    frames = x.shape[0] if _frames else 1
    Lh.ref_input_convert.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    parts = []
    for f in range(frames):                      # InputConvert lays out ONE image (input_loader.cpp:123-155 ignores n on the source side)
        inp_f = np.zeros(isz, np.float32)
        xr = np.ascontiguousarray(x[f] if _frames else x, dtype=np.float32)
        Lh.ref_input_convert(xr.ctypes.data, inp_f.ctypes.data, 1)
        parts.append(inp_f.astype(np.int8))
    inp = np.concatenate(parts)
    fraw = np.full(fsz, 64, np.uint8)
    bb = np.zeros((nl * mb + 16, 3), np.int32)
    for l, (codes, params) in enumerate(model):
        if codes is None:
            continue
        fraw[l * stride: l * stride + codes.size] = np.ascontiguousarray(codes, dtype=np.uint8).reshape(-1)
        bb[l * mb: l * mb + params.shape[0]] = params
    freal = np.full(fsz, 64, np.uint8)
    scratch = np.zeros(fsz, np.uint8)
    Lh.ref_filter_convert.argtypes = [C.c_void_p] * 3
    Lh.ref_filter_convert(scratch.ctypes.data, fraw.ctypes.data, freal.ctypes.data)
    ddr = np.zeros(Ln.net_ddr_bytes() + frames * Ln.net_output_offset() + (1 << 20), np.int8)
    info = lambda l, w: Ln.net_layer_info(l, w)
    total = frames * sum(info(l, 2) for l in range(nl)) + 4096
    rb, do = Ln.net_tap_record_bytes(), Ln.net_tap_data_offset()
    tap = np.zeros(total * rb, np.uint8)
    ntap = C.c_longlong(0)
    stats = np.zeros(64, np.int64)
    Ln.net_run_frames.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_longlong, C.c_void_p, C.c_void_p]
    rc = Ln.net_run_frames(frames, inp.ctypes.data, freal.ctypes.data, bb.ctypes.data, ddr.ctypes.data, tap.ctypes.data, total,
                           C.byref(ntap), stats.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"net_run failed: {rc}")
    rec = tap[: ntap.value * rb].reshape(ntap.value, rb)
    which = rec[:, :4].copy().view(np.int32).reshape(-1)
    tiles = [rec[which == w][:, do:do + 128].view(np.int8) for w in (0, 1)]
    pos = [0, 0]
    per_layer = []

    def untile(t, nvec, ph, pwv, N, pw):
        t = t.reshape(nvec, ph, pwv, 8, 16)[:, :, :, :7, :]
        return np.ascontiguousarray(t.transpose(0, 4, 1, 2, 3).reshape(nvec * 16, ph, pwv * 7)[:N, :, :pw])

    for l in range(nl):
        nvec, ph, pwv, N, pw = info(l, 9), info(l, 10), info(l, 11), info(l, 12), info(l, 13)
        if info(l, 1):                       # end pool: one tile per 16 channels from full_size_pool
            t = tiles[1][pos[1]: pos[1] + nvec]
            pos[1] += nvec
            per_layer.append(untile(t, nvec, 1, 1, N, 1) if t.shape[0] == nvec else None)
        elif info(l, 0):
            n = nvec * ph * pwv
            t = tiles[0][pos[0]: pos[0] + n]
            pos[0] += n
            per_layer.append(untile(t, nvec, ph, pwv, N, pw) if t.shape[0] == n else None)
        else:
            per_layer.append(None)
    l = nl - 1
    nvec, ph, pwv, N, pw = info(l, 9), info(l, 10), info(l, 11), info(l, 12), info(l, 13)
    if info(l, 1):
        ph = pw = pwv = 1
    finals = []
    for f in range(frames):
        base = Ln.net_output_offset() * (1 + f) + info(l, 4)
        finals.append(untile(ddr[base: base + nvec * ph * pwv * 128].copy(), nvec, ph, pwv, N, pw))
    final = finals[0]
    st = {"done": int(stats[0]), "parked": int(stats[1]), "parked_ids": [int(v) for v in stats[8:8 + int(stats[1])]],
          "switches": int(stats[2]), "tap_dropped": int(stats[3]), "fifo_bytes_left": int(stats[4]),
          "tap_used": pos, "tap_counts": [int(t.shape[0]) for t in tiles], "ddr": ddr}
    if _frames:
        st["tap_per_frame"] = [c // frames for c in st["tap_counts"]]
        return finals, st
    return per_layer, final, st
