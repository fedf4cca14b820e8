"""File formats at the TF2 boundary and the host-side numeric preparation (NumPy, vectorised).

Mirrors, bit for bit, what the reference host does before anything reaches the accelerator:
  * Q text file            -> q table        (`Runtime_Engine/cnn/host/src/quantization.cpp:25-55`)
  * float weight blob      -> shift codes + BiasBnParam (`model_loader.cpp:98-258`)
  * 4-bit weight blob      (`TransForm_Kit/Compression/compress_net/4bit_data_format.txt:1-44`)
  * float image .bin       -> int8 image     (`input_loader.cpp:76-118`, `runner.cpp:158-164`)
The tests compare every function here with the reference's own sources compiled by
`oracle/build_ref.sh` and with the independent C restatement in `oracle/tf2_oracle.c`.
"""
from __future__ import annotations

import struct
from typing import List, Tuple

import numpy as np

from .netdesc import NetDesc

INFLAT = 15        # types.h:33
ALPHA_INFLAT = 20  # types.h:32


# --------------------------------------------------------------------------------------------
# Q file
# --------------------------------------------------------------------------------------------
def parse_q_text(net: NetDesc, text: str) -> np.ndarray:
    """quantization.cpp:25-55.  Returns q[num_q_rows][max_out_channel] int8 holding -Q.

    Row 0 = image (3 values), row l+1 = output of layer l (kOutputChannels[l] values, read in layer
    order); an ipool layer's row copies its input row instead of consuming values (:42-43); a branch
    tail also fills its slice of the concat row (:47-49).
    """
    vals = [int(v) for v in text.split()]
    q = np.zeros((net.num_q_rows, net.max_out_channel), dtype=np.int8)
    pos = 0

    def take(n):
        nonlocal pos
        if pos + n > len(vals):
            raise ValueError(f"Q file too short: need {pos + n} values, have {len(vals)}")
        out = vals[pos:pos + n]
        pos += n
        return out

    q[0, :net.input_c] = -np.array(take(net.input_c), dtype=np.int64)
    for l, ld in enumerate(net.layers):
        n = ld.N
        if ld.ipool:
            q[l + 1, :n] = q[ld.q_in_row, :n]
            continue
        v = -np.array(take(n), dtype=np.int64)
        q[l + 1, :n] = v
        if net.branch_tail and net.branch_tail[l]:
            row = net.num_layers + 1 + net.concat_layer[l]
            q[row, ld.out_ch0:ld.out_ch0 + n] = v
    return q


def parse_q_file(net: NetDesc, path: str) -> np.ndarray:
    with open(path, "r") as f:
        return parse_q_text(net, f.read())


def q_file_value_count(net: NetDesc) -> int:
    return net.input_c + sum(ld.N for ld in net.layers if not ld.ipool)


# --------------------------------------------------------------------------------------------
# float -> shift code  (Get_real)
# --------------------------------------------------------------------------------------------
def get_real(w: np.ndarray, expand: np.ndarray) -> np.ndarray:
    """model_loader.cpp:98-126, vectorised.  `w` float32, `expand` int8 (broadcastable)."""
    w = np.asarray(w, dtype=np.float32)
    expand = np.asarray(expand).astype(np.int8)
    absw = np.abs(w).astype(np.float64)
    zero = absw < 1.0e-05
    vals = np.zeros(w.shape, dtype=np.int32)
    found = np.zeros(w.shape, dtype=bool)
    for i in range(15):
        temps = float(np.float32(1.0) / np.float32(1 << i))
        hit = (~found) & (absw > 0.99 * temps) & (absw < 1.01 * temps)
        vals[hit] = i
        found |= hit
    oups = (expand.astype(np.int32) - vals).astype(np.int8)  # `char oups = expand - vals`
    oups = np.where(oups < 0, 0, oups).astype(np.uint8)
    oups = np.where(w < 0, oups | 0x80, oups).astype(np.uint8)
    return np.where(zero, np.uint8(0x40), oups).astype(np.uint8)


_W_COL = ((0, 2, 4), (1, 3, 5), (None, None, 6))  # filter column carried by (wsel, j)


def filter_trans(codes_7x7: np.ndarray, fill: int = 0) -> np.ndarray:
    """model_loader.cpp:25-96 + :244-257.  codes [..., 7, 7] -> [..., 9, 3, 3].

    Taps the reference never writes keep `fill`; LoadModel memsets the destination to 0, which is
    the code of weight +1 — that is the default here.
    """
    c = np.asarray(codes_7x7, dtype=np.uint8)
    out = np.full(c.shape[:-2] + (9, 3, 3), fill, dtype=np.uint8)
    for w in range(3):
        for hp in range(2):
            d = 2 * w + hp
            for r in range(3):
                for j in range(3):
                    col = _W_COL[w][j]
                    out[..., d, r, j] = 0x40 if col is None else c[..., 2 * r + hp, col]
        for j in range(3):
            col = _W_COL[w][j]
            out[..., 6 + w, 2, j] = 0x40 if col is None else c[..., 6, col]
    return out


def feature_trans(img: np.ndarray) -> np.ndarray:
    """input_loader.cpp:27-73 (+ the 114x114 crop of :99-115).  [..., 224, 224] -> [..., 9, 114, 114].

    Works for float images and for already-quantised int8 images alike (pure permutation + zero
    padding)."""
    img = np.asarray(img)
    P = np.zeros(img.shape[:-2] + (232, 232), dtype=img.dtype)
    P[..., 3:227, 3:227] = img
    out = np.zeros(img.shape[:-2] + (9, 114, 114), dtype=img.dtype)
    for d in range(9):
        if d < 6:
            w, r0 = d // 2, d % 2
        else:
            w, r0 = d - 6, 2
        out[..., d, :, :] = P[..., r0:r0 + 228:2, w:w + 228:2]
    return out


def quantize_input(x: np.ndarray, q0: int) -> np.ndarray:
    """runner.cpp:158-164: x * 2^Q0 (q0 = -Q0 from the Q table), round half away, clamp int8."""
    x = np.asarray(x, dtype=np.float32)
    trans = np.float32(1.0 / (1 << q0)) if q0 > 0 else np.float32(1 << (-q0))
    tmp = (x * trans).astype(np.float32)
    t = np.where(tmp > 0, tmp.astype(np.float64) + 0.5, tmp.astype(np.float64) - 0.5)
    t = np.trunc(t)
    return np.clip(t, -128, 127).astype(np.int8)


def load_image_bin(path: str, c: int = 3, h: int = 224, w: int = 224) -> np.ndarray:
    """input_loader.cpp:76-97: raw float32 [C][H][W], already mean-subtracted."""
    a = np.fromfile(path, dtype="<f4", count=c * h * w)
    if a.size != c * h * w:
        raise ValueError(f"{path}: expected {c * h * w} floats, got {a.size}")
    return a.reshape(c, h, w)


def prepare_input(net: NetDesc, images_f32: np.ndarray, q: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """float images [B][3][224][224] -> (raw int8 [B][3][224][224], tensor-0 int8 [B][27][114][114])."""
    raw = quantize_input(images_f32, int(q[0, 0]))
    t0 = feature_trans(raw).reshape(raw.shape[0], -1, 114, 114)
    return raw, t0


# --------------------------------------------------------------------------------------------
# float weight blob (param.bin / fpgamodel.bin)
# --------------------------------------------------------------------------------------------
def _read_f32(buf: memoryview, pos: int, n: int) -> Tuple[np.ndarray, int]:
    end = pos + 4 * n
    if end > len(buf):
        raise ValueError("model blob too short")
    return np.frombuffer(buf[pos:end], dtype="<f4"), end


def load_float_blob(net: NetDesc, data: bytes, q: np.ndarray) -> List[Tuple[np.ndarray, np.ndarray]]:
    """model_loader.cpp:129-258.  Returns per layer (codes uint8 [N][C][k][k], params int32 [N][3]);
    ipool layers yield (None, None).  Blob order per layer: W[N][C][H][W], bias[N] if kBiasEnable,
    then if kBnEnable mean[N], var[N], scale_factor[1], gamma[N], beta[N]."""
    buf = memoryview(data)
    pos = 0
    out: List[Tuple[np.ndarray, np.ndarray]] = []
    for l, ld in enumerate(net.layers):
        if ld.ipool:
            # LoadModel still walks bias/BN flags for the pseudo layer; the shipped tables have none
            if ld.bias_en or ld.bn_en:
                raise ValueError("ipool layer with bias/BN is outside the reference's use")
            out.append((None, None))
            continue
        N = ld.N
        if ld.first_layer_7x7:
            C, H, W = net.input_c, 7, 7
        else:
            C, H, W = ld.C, ld.k, ld.k
        q_in = q[ld.q_in_row, :C].astype(np.int32)
        q_out = q[ld.q_out_row, :N].astype(np.int32)
        w, pos = _read_f32(buf, pos, N * C * H * W)
        w = w.reshape(N, C, H, W)
        expand = (INFLAT + q_in[None, :] - q_out[:, None]).astype(np.int8)  # `char expand`
        codes = get_real(w, expand[:, :, None, None])
        params = np.zeros((N, 3), dtype=np.int32)
        coe = (np.int64(1) << (INFLAT - q_out).astype(np.int64)).astype(np.float32)  # bias_trans_coe
        if ld.bias_en:
            b, pos = _read_f32(buf, pos, N)
            params[:, 0] = (b.astype(np.float32) * coe).astype(np.float32).astype(np.int32)
        if ld.bn_en:
            mean, pos = _read_f32(buf, pos, N)
            var, pos = _read_f32(buf, pos, N)
            sf, pos = _read_f32(buf, pos, 1)
            gamma, pos = _read_f32(buf, pos, N)
            beta, pos = _read_f32(buf, pos, N)
            sf = np.float32(sf[0])
            with np.errstate(divide="ignore", invalid="ignore"):
                a = (mean / sf).astype(np.float32)
                bsd = np.sqrt((var / sf).astype(np.float32) + np.float32(0.00001)).astype(np.float32)
                alpha_data = (gamma / bsd).astype(np.float32)
                beta_data = (-(alpha_data * a).astype(np.float32) + beta).astype(np.float32)
        else:
            alpha_data = np.ones(N, dtype=np.float32)
            beta_data = np.zeros(N, dtype=np.float32)
        params[:, 1] = np.trunc(alpha_data.astype(np.float64) * float(2 ** ALPHA_INFLAT)).astype(np.int32)
        cb = (coe * beta_data).astype(np.float32).astype(np.float64)
        params[:, 2] = np.trunc(np.where(beta_data > 0, cb + 0.5, cb - 0.5)).astype(np.int32)
        if ld.first_layer_7x7:
            codes = filter_trans(codes, fill=0).reshape(N, C * 9, 3, 3)
        out.append((np.ascontiguousarray(codes), params))
    return out


def load_float_blob_file(net: NetDesc, path: str, q: np.ndarray):
    with open(path, "rb") as f:
        return load_float_blob(net, f.read(), q)


def float_blob_size(net: NetDesc) -> int:
    n = 0
    for ld in net.layers:
        if ld.ipool:
            continue
        if ld.first_layer_7x7:
            n += ld.N * net.input_c * 49
        else:
            n += ld.N * ld.C * ld.k * ld.k
        if ld.bias_en:
            n += ld.N
        if ld.bn_en:
            n += 4 * ld.N + 1
    return 4 * n


# --------------------------------------------------------------------------------------------
# 4-bit weight blob (4bit_data_format.txt).  The reference ships the spec only ("code will be
# available soon", TransForm_Kit/Compression/README.md:36); nibble order inside a short is not
# specified there — this implementation fixes it as LOW NIBBLE FIRST and documents it here.
# Layer record: int8 min_exp, int8 dtype(0), int16 N, C, H, W, then shorts; a filter row of width
# w is stored as ceil(w/3) shorts of up to 3 nibbles (1x1: 4 consecutive weights per short along
# the flattened [N][C] order, 2x2: 2 per short).
# --------------------------------------------------------------------------------------------
def weights_to_nibbles(w: np.ndarray, min_exp: int) -> np.ndarray:
    """float weights (+-2^(min_exp+e), e in 0..6, or 0) -> 4-bit codes ((positive?8:0)|e, 7 = zero)."""
    w = np.asarray(w, dtype=np.float32)
    nib = np.full(w.shape, 7, dtype=np.uint8)
    nz = w != 0
    e = np.zeros(w.shape, dtype=np.int32)
    e[nz] = np.round(np.log2(np.abs(w[nz]))).astype(np.int32) - min_exp
    if nz.any() and (e[nz].min() < 0 or e[nz].max() > 6):
        raise ValueError("weights do not fit 7 levels above min_exp")
    if nz.any() and not np.array_equal(np.abs(w[nz]), np.exp2((e[nz] + min_exp).astype(np.float32))):
        raise ValueError("weights are not powers of two")
    nib[nz] = (e[nz] | np.where(w[nz] > 0, 8, 0)).astype(np.uint8)
    return nib


def nibbles_to_weights(nib: np.ndarray, min_exp: int) -> np.ndarray:
    nib = np.asarray(nib, dtype=np.uint8)
    e = (nib & 7).astype(np.int32)
    mag = np.exp2((e + min_exp).astype(np.float32))
    w = np.where(nib & 8, mag, -mag).astype(np.float32)
    return np.where(e == 7, np.float32(0), w)


def _group_sizes(kw: int) -> List[int]:
    if kw == 1:
        return [1]
    if kw == 2:
        return [2]
    a, b = divmod(kw, 3)
    return [3] * a + ([b] if b else [])


def pack4_layer(nib: np.ndarray, min_exp: int) -> bytes:
    """nib uint8 [N][C][H][W] -> one layer record of the 4-bit blob."""
    N, C, H, W = nib.shape
    hdr = struct.pack("<bbhhhh", min_exp, 0, N, C, H, W)
    if H == 1 and W == 1:
        flat = nib.reshape(-1)
        pad = (-flat.size) % 4
        flat = np.concatenate([flat, np.full(pad, 7, dtype=np.uint8)]).reshape(-1, 4).astype(np.uint16)
        shorts = flat[:, 0] | (flat[:, 1] << 4) | (flat[:, 2] << 8) | (flat[:, 3] << 12)
    else:
        rows = nib.reshape(-1, W).astype(np.uint16)
        cols = []
        start = 0
        for g in _group_sizes(W):
            s = np.zeros(rows.shape[0], dtype=np.uint16)
            for i in range(g):
                s |= rows[:, start + i] << (4 * i)
            cols.append(s)
            start += g
        shorts = np.stack(cols, axis=1).reshape(-1)
    return hdr + shorts.astype("<u2").tobytes()


def unpack4_layer(buf: memoryview, pos: int) -> Tuple[np.ndarray, int, int]:
    """-> (nib uint8 [N][C][H][W], min_exp, new_pos)"""
    min_exp, dtype, N, C, H, W = struct.unpack_from("<bbhhhh", buf, pos)
    pos += 10
    if dtype != 0:
        raise ValueError("only short-coded (dtype 0) records hold weights")
    if H == 1 and W == 1:
        cnt = N * C
        ns = (cnt + 3) // 4
        s = np.frombuffer(buf[pos:pos + 2 * ns], dtype="<u2").astype(np.uint16)
        pos += 2 * ns
        nib = np.stack([(s >> (4 * i)) & 0xF for i in range(4)], axis=1).reshape(-1)[:cnt]
        return nib.astype(np.uint8).reshape(N, C, 1, 1), min_exp, pos
    groups = _group_sizes(W)
    nrows = N * C * H
    ns = nrows * len(groups)
    s = np.frombuffer(buf[pos:pos + 2 * ns], dtype="<u2").astype(np.uint16).reshape(nrows, len(groups))
    pos += 2 * ns
    cols = []
    for gi, g in enumerate(groups):
        for i in range(g):
            cols.append((s[:, gi] >> (4 * i)) & 0xF)
    nib = np.stack(cols, axis=1).astype(np.uint8).reshape(N, C, H, W)
    return nib, min_exp, pos


def _float_record(a: np.ndarray) -> bytes:
    """dtype 1 record (non-convolution parameters stay float): header + float32 values."""
    a = np.asarray(a, dtype="<f4")
    dims = list(a.shape) + [1] * (4 - a.ndim)
    return struct.pack("<bbhhhh", 0, 1, *dims) + a.tobytes()


def _blob_fields(net: NetDesc):
    """The fields of a float param.bin in file order (model_loader.cpp:154-231): (layer, kind, shape)."""
    for l, ld in enumerate(net.layers):
        if ld.ipool:
            continue
        if ld.first_layer_7x7:
            yield l, "weight", (ld.N, net.input_c, 7, 7)
        else:
            yield l, "weight", (ld.N, ld.C, ld.k, ld.k)
        if ld.bias_en:
            yield l, "bias", (ld.N,)
        if ld.bn_en:
            for kind, shape in (("mean", (ld.N,)), ("var", (ld.N,)), ("scale_factor", (1,)), ("gamma", (ld.N,)), ("beta", (ld.N,))):
                yield l, kind, shape


def float_blob_to_4bit(net: NetDesc, blob: bytes) -> bytes:
    """A float param.bin whose convolution weights sit on a 7-level power-of-two grid per layer (INQ output,
    tf2_b200.compress) -> the 4-bit model file of 4bit_data_format.txt: one record per blob, in param.bin order —
    weights short-coded (dtype 0), biases and BatchNorm / Scale parameters as float records (dtype 1)."""
    buf = memoryview(blob)
    pos = 0
    out = []
    for l, kind, shape in _blob_fields(net):
        a, pos = _read_f32(buf, pos, int(np.prod(shape)))
        if kind == "weight":
            nz = a[a != 0]
            min_exp = int(np.round(np.log2(np.abs(nz).min()))) if nz.size else 0
            out.append(pack4_layer(weights_to_nibbles(a.reshape(shape), min_exp), min_exp))
        else:
            out.append(_float_record(a.reshape(shape)))
    if pos != len(buf):
        raise ValueError(f"model blob has {len(buf) - pos} trailing bytes")
    return b"".join(out)


def float_blob_from_4bit(net: NetDesc, data: bytes) -> bytes:
    """Inverse of float_blob_to_4bit: the float param.bin a 4-bit model file stands for (bit-exact: every
    weight is a power of two), ready for load_float_blob / LoadModel."""
    buf = memoryview(data)
    pos = 0
    out = []
    for l, kind, shape in _blob_fields(net):
        if pos + 10 > len(buf):
            raise ValueError("4-bit model file too short")
        dtype = struct.unpack_from("<b", buf, pos + 1)[0]
        if kind == "weight":
            if dtype != 0:
                raise ValueError(f"layer {l}: expected a short-coded weight record")
            nib, min_exp, pos = unpack4_layer(buf, pos)
            if nib.shape != tuple(shape):
                raise ValueError(f"layer {l}: weight record {nib.shape} does not match the tables {tuple(shape)}")
            out.append(nibbles_to_weights(nib, min_exp).astype("<f4").tobytes())
        else:
            if dtype != 1:
                raise ValueError(f"layer {l}: expected a float record for {kind}")
            dims = struct.unpack_from("<hhhh", buf, pos + 2)
            cnt = int(np.prod(shape))
            if int(np.prod(dims)) != cnt:
                raise ValueError(f"layer {l}: {kind} record {dims} does not match the tables {tuple(shape)}")
            pos += 10
            a, pos = _read_f32(buf, pos, cnt)
            out.append(a.astype("<f4").tobytes())
    if pos != len(buf):
        raise ValueError(f"4-bit model file has {len(buf) - pos} trailing bytes")
    return b"".join(out)


def nibbles_dense(nib: np.ndarray) -> np.ndarray:
    """[N][C][H][W] 4-bit codes -> dense bytes, two per byte, low nibble first (C-ABI layout)."""
    flat = np.asarray(nib, dtype=np.uint8).reshape(-1)
    if flat.size % 2:
        flat = np.concatenate([flat, np.zeros(1, dtype=np.uint8)])
    return (flat[0::2] | (flat[1::2] << 4)).astype(np.uint8)


def codes_from_nibbles(nib: np.ndarray, min_exp: int, q_in: np.ndarray, q_out: np.ndarray) -> np.ndarray:
    """What Get_real would return for the floats a 4-bit record encodes (host mirror of
    tf2b_load_layer_packed4)."""
    w = nibbles_to_weights(nib, min_exp)
    expand = (INFLAT + q_in.astype(np.int32)[None, :] - q_out.astype(np.int32)[:, None]).astype(np.int8)
    return get_real(w, expand[:, :, None, None])


# --------------------------------------------------------------------------------------------
# Device (feature_ddr) layout of a feature map
# --------------------------------------------------------------------------------------------
W_VECTOR = 7          # archs.h:41 (FW_VECTOR + OW_VECTOR - 1)
N_VECTOR = 16         # archs.h:26 / NARROW_N_VECTOR
TILE = 128            # NEXT_POWER_OF_2(W_VECTOR * NARROW_N_VECTOR)


def to_device_layout(fmap: np.ndarray) -> np.ndarray:
    """[C][H][W] int8 -> the tile order feature_writer.cl:116-137 writes to feature_ddr and
    network_helper.cpp:95-118 / 160-170 reads back: [C/16][H][ceil(W/7)][128 slots: (w % 7) * 16 + c % 16]
    (slots 112..127 and the padding of ragged C / W are zero)."""
    fmap = np.asarray(fmap, dtype=np.int8)
    C_, H, W = fmap.shape
    nv, wv = -(-C_ // N_VECTOR), -(-W // W_VECTOR)
    pad = np.zeros((nv * N_VECTOR, H, wv * W_VECTOR), np.int8)
    pad[:C_, :, :W] = fmap
    t = pad.reshape(nv, N_VECTOR, H, wv, W_VECTOR).transpose(0, 2, 3, 4, 1)          # [nv][H][wv][7][16]
    out = np.zeros((nv, H, wv, TILE), np.int8)
    out[..., :W_VECTOR * N_VECTOR] = t.reshape(nv, H, wv, W_VECTOR * N_VECTOR)
    return out.reshape(-1)


def from_device_layout(buf: np.ndarray, C_: int, H: int, W: int) -> np.ndarray:
    """Inverse of to_device_layout: the un-tiling Verify() does element by element."""
    nv, wv = -(-C_ // N_VECTOR), -(-W // W_VECTOR)
    t = np.asarray(buf, dtype=np.int8)[: nv * H * wv * TILE].reshape(nv, H, wv, TILE)[..., :W_VECTOR * N_VECTOR]
    t = t.reshape(nv, H, wv, W_VECTOR, N_VECTOR).transpose(0, 4, 1, 2, 3).reshape(nv * N_VECTOR, H, wv * W_VECTOR)
    return np.ascontiguousarray(t[:C_, :, :W])
