"""`fpganetwork.bin` -> layer tables: the front half of the reference's TF2_auto_config.

caffe2fpga dumps the network as raw C structs (`TransForm_Kit/ModelConvert/caffe2fpga/src/tools.cpp:
437-568`; structs in `Runtime_Engine/TF2_auto_config/inc/fpganetworkinterface.h:36-257`, x86-64
layout, embedded pointers are garbage); `tf2_auto_config` turns it into the `k*` tables of `<net>.h`
(`tf2_auto_param.cpp:33-1058`).  This module reads the same file and produces the same *network*
tables as a `NetDesc` (fused conv / bias / BN / ReLU / pool / eltwise / global-average layers with
explicit tensor ids), so a TransForm_Kit-emitted model drops in without a hand-written header.  The
FPGA-only outputs of tf2_auto_config (cache pages, DDR bases, cycle counts) are not produced: they
describe the FPGA's memories, not the network.

File layout (little endian): StFpgaNetInfo 144 B {int version; char name[128]; int nLayers; ptr};
per layer StFpgaLayerInfo 424 B {int id, nOps, nIn, nOut; int inId[10]; int outId[10];
StBlobShape in[10], out[10] (N, C, H, W); ptr}; per op StFpgaOpInfo 16 B {int type; ptr} + payload:
Conv 48 B, Fc 8 B, Bn 4 B, Scale 1 B (unaligned), Pool 40 B, Eltwise 4 B, others none.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

from .netdesc import LayerDesc, NetDesc, TensorDesc

# EN_fpgaop (fpganetworkinterface.h:36-55)
OP_CONV, OP_FC, OP_BN, OP_SCALE, OP_LRN, OP_RELU, OP_RELU6, OP_PRELU, OP_ELU, OP_SIGMOID, OP_POOL, OP_ELTWISE, \
    OP_CONCAT, OP_FLATTEN, OP_SOFTMAX, OP_SLICE, OP_UNSUPPORTED = range(1, 18)
OP_NAMES = {OP_CONV: "conv", OP_FC: "fc", OP_BN: "bn", OP_SCALE: "scale", OP_LRN: "lrn", OP_RELU: "relu",
            OP_RELU6: "relu6", OP_PRELU: "prelu", OP_ELU: "elu", OP_SIGMOID: "sigmoid", OP_POOL: "pool",
            OP_ELTWISE: "eltwise", OP_CONCAT: "concat", OP_FLATTEN: "flatten", OP_SOFTMAX: "softmax",
            OP_SLICE: "slice", OP_UNSUPPORTED: "unsupported"}
VERSION = 20190510


@dataclass
class FpgaOp:
    type: int
    p: Dict[str, float] = field(default_factory=dict)


@dataclass
class FpgaLayer:
    id: int
    ops: List[FpgaOp]
    in_ids: List[int]
    out_ids: List[int]
    in_shapes: List[Tuple[int, int, int, int]]
    out_shapes: List[Tuple[int, int, int, int]]

    def op(self, t: int) -> Optional[FpgaOp]:
        for o in self.ops:
            if o.type == t:
                return o
        return None


@dataclass
class FpgaNet:
    version: int
    name: str
    layers: List[FpgaLayer]


def parse_fpganetwork(data: bytes) -> FpgaNet:
    """Reads the struct dump (reader of the reference: fpganetworkinterface.cpp:27-183)."""
    if len(data) < 144:
        raise ValueError("fpganetwork.bin: file shorter than the 144-byte network header")
    version, = struct.unpack_from("<i", data, 0)
    name = data[4:132].split(b"\0", 1)[0].decode("latin-1")
    nlayers, = struct.unpack_from("<i", data, 132)
    if version != VERSION:
        raise ValueError(f"fpganetwork.bin: version {version}, expected {VERSION}")
    if not 0 < nlayers < 100000:
        raise ValueError(f"fpganetwork.bin: implausible layer count {nlayers}")
    off = 144
    layers: List[FpgaLayer] = []
    for _ in range(nlayers):
        if off + 424 > len(data):
            raise ValueError("fpganetwork.bin: truncated layer record")
        lid, nops, nin, nout = struct.unpack_from("<4i", data, off)
        in_ids = list(struct.unpack_from("<10i", data, off + 16))
        out_ids = list(struct.unpack_from("<10i", data, off + 56))
        ins = [struct.unpack_from("<4i", data, off + 96 + 16 * i) for i in range(10)]
        outs = [struct.unpack_from("<4i", data, off + 256 + 16 * i) for i in range(10)]
        off += 424
        if not (0 <= nin <= 10 and 0 <= nout <= 10 and 0 <= nops <= 64):
            raise ValueError(f"fpganetwork.bin: layer {lid}: bad counts ops={nops} in={nin} out={nout}")
        ops: List[FpgaOp] = []
        for _ in range(nops):
            t, = struct.unpack_from("<i", data, off)
            off += 16
            op = FpgaOp(t)
            if t == OP_CONV:
                outc, bias, pl, pt, pr, pb, padv, kh, kw, sh, sw, dil = struct.unpack_from("<i?3x4if5i", data, off)
                op.p = dict(out_c=outc, bias=int(bias), pad_l=pl, pad_t=pt, pad_r=pr, pad_b=pb, kh=kh, kw=kw, sh=sh, sw=sw,
                            dilation=dil)
                off += 48
            elif t == OP_FC:
                outc, bias = struct.unpack_from("<i?3x", data, off)
                op.p = dict(out_c=outc, bias=int(bias))
                off += 8
            elif t == OP_BN:
                op.p = dict(eps=struct.unpack_from("<f", data, off)[0])
                off += 4
            elif t == OP_SCALE:
                op.p = dict(bias=int(data[off] != 0))
                off += 1
            elif t == OP_POOL:
                m, pl, pt, pr, pb, kh, kw, sh, sw, glob = struct.unpack_from("<9i?3x", data, off)
                op.p = dict(method=m, pad_l=pl, pad_t=pt, pad_r=pr, pad_b=pb, kh=kh, kw=kw, sh=sh, sw=sw, global_pool=int(glob))
                off += 40
            elif t == OP_ELTWISE:
                op.p = dict(method=struct.unpack_from("<i", data, off)[0])
                off += 4
            elif t in (OP_RELU, OP_CONCAT, OP_SOFTMAX, OP_FLATTEN):
                pass
            else:
                raise ValueError(f"fpganetwork.bin: layer {lid}: op {OP_NAMES.get(t, t)} is not supported by the TF2 runtime")
            ops.append(op)
        layers.append(FpgaLayer(lid, ops, in_ids[:nin], out_ids[:nout], [tuple(s) for s in ins[:nin]],
                                [tuple(s) for s in outs[:nout]]))
    if off != len(data):
        raise ValueError(f"fpganetwork.bin: {len(data) - off} trailing bytes")
    return FpgaNet(version, name, layers)


def write_fpganetwork(net: FpgaNet) -> bytes:
    """Inverse of parse_fpganetwork (writer of the reference: caffe2fpga tools.cpp:437-568); pointer
    fields are written as zero."""
    out = bytearray()
    out += struct.pack("<i", net.version) + net.name.encode("latin-1")[:127].ljust(128, b"\0") + struct.pack("<i", len(net.layers))
    out += b"\0" * 8
    for L in net.layers:
        rec = bytearray(424)
        struct.pack_into("<4i", rec, 0, L.id, len(L.ops), len(L.in_ids), len(L.out_ids))
        struct.pack_into("<10i", rec, 16, *(L.in_ids + [0] * (10 - len(L.in_ids))))
        struct.pack_into("<10i", rec, 56, *(L.out_ids + [0] * (10 - len(L.out_ids))))
        for i, s in enumerate(L.in_shapes):
            struct.pack_into("<4i", rec, 96 + 16 * i, *s)
        for i, s in enumerate(L.out_shapes):
            struct.pack_into("<4i", rec, 256 + 16 * i, *s)
        out += rec
        for o in L.ops:
            out += struct.pack("<i", o.type) + b"\0" * 12
            p = o.p
            if o.type == OP_CONV:
                out += struct.pack("<i?3x4if5i", p["out_c"], bool(p["bias"]), p["pad_l"], p["pad_t"], p["pad_r"], p["pad_b"], 0.0,
                                   p["kh"], p["kw"], p["sh"], p["sw"], p.get("dilation", 1))
            elif o.type == OP_FC:
                out += struct.pack("<i?3x", p["out_c"], bool(p["bias"]))
            elif o.type == OP_BN:
                out += struct.pack("<f", p.get("eps", 1e-5))
            elif o.type == OP_SCALE:
                out += struct.pack("<?", bool(p.get("bias", 1)))
            elif o.type == OP_POOL:
                out += struct.pack("<9i?3x", p["method"], p["pad_l"], p["pad_t"], p["pad_r"], p["pad_b"], p["kh"], p["kw"], p["sh"],
                                   p["sw"], bool(p.get("global_pool", 0)))
            elif o.type == OP_ELTWISE:
                out += struct.pack("<i", p["method"])
    return bytes(out)


def to_netdesc(fn: FpgaNet, name: Optional[str] = None, max_pool_pad: Optional[int] = None) -> NetDesc:
    """Fuses the Caffe-level layers into the runtime's fused layers and resolves the tensor graph —
    what `FrameParseStore` + `ParamGeneration` do for the network tables (tf2_auto_param.cpp:33-1847):

      conv|fc [+ bn + scale] [+ relu]   -> one layer (kBnEnable / kBiasEnable / kReluEnable)
      3x3 max pool after a conv          -> kPoolEnable of that conv (kPoolStride2, kPoolPad); a pool whose
                                            producer cannot absorb it becomes an ipool pseudo layer
      eltwise sum [+ relu]               -> kAdditionEnable / kAdditionReluEnable of the LATER producer
      7x7 average pool                   -> kEndPoolEnable of its producer
      concat                             -> the producers write one shared tensor at channel offsets
      softmax                            -> dropped (the runtime stops at the logits)
      first 7x7 / stride 2 / pad 3 conv  -> the 27-channel 114x114 3x3 form (tf2_auto_param.cpp:214-254)

    `max_pool_pad`: the reference forces kPoolPad = 1 for networks it knows by NAME as ResNet50
    (tf2_auto_param.cpp:1593-1594; the Caffe file says 0 = ceil mode).  Pass 1 to get the shipped
    resnet50.h tables; None keeps what the file says."""
    name = name or (fn.name if fn.name and fn.name != "undefine" else "net")
    consumers: Dict[int, List[int]] = {}
    for i, L in enumerate(fn.layers):
        for b in L.in_ids:
            consumers.setdefault(b, []).append(i)
    first = fn.layers[0]
    if not first.op(OP_CONV) or first.in_ids != [-1]:
        raise ValueError("fpganetwork: the first layer must be a convolution fed by the image (blob -1)")
    _, ic, ih, iw = first.in_shapes[0]
    c0 = first.op(OP_CONV).p
    tensors: List[TensorDesc] = []
    layers: List[LayerDesc] = []
    blob_t: Dict[int, Tuple[int, int]] = {}     # blob id -> (tensor id, channel offset)
    blob_layer: Dict[int, int] = {}             # blob id -> index of the fused layer that produces it
    seven = (c0["kh"], c0["sh"], c0["pad_l"], ic, ih, iw) == (7, 2, 3, 3, 224, 224)
    tensors.append(TensorDesc(27, 114, 114, 0, "input") if seven else TensorDesc(ic, ih, iw, 0, "input"))
    blob_t[-1] = (0, 0)

    # concat outputs own a tensor; their inputs are written into it at channel offsets
    concat_of: Dict[int, Tuple[int, int]] = {}  # producer blob -> (concat index, channel offset)
    concat_shape: Dict[int, Tuple[int, int, int]] = {}
    concat_out_blob: Dict[int, int] = {}
    ncat = 0
    for L in fn.layers:
        if L.op(OP_CONCAT):
            off = 0
            for b, s in zip(L.in_ids, L.in_shapes):
                concat_of[b] = (ncat, off)
                off += s[1]
            concat_shape[ncat] = (off, L.out_shapes[0][2], L.out_shapes[0][3])
            concat_out_blob[ncat] = L.out_ids[0]
            ncat += 1
    concat_tensor: Dict[int, int] = {}
    branch_tail: List[int] = []
    concat_layer: List[int] = []

    def new_layer(**kw) -> LayerDesc:
        l = len(layers)
        d = dict(name=f"layer{l}", in_tensor=-1, out_tensor=-1, out_ch0=0, add_tensor=-1, C=0, N=0, k=1, pad=0, stride=1,
                 OH=0, OW=0, relu=0, pool=0, pool_stride=1, pool_pad=0, PH=0, PW=0, add_relu=0, gap=0, ipool=0,
                 bias_en=0, bn_en=0, in_may_be_m128=0, q_in_row=0, q_out_row=l + 1, first_layer_7x7=0)
        d.update(kw)
        ld = LayerDesc(**d)
        layers.append(ld)
        branch_tail.append(0)
        concat_layer.append(0)
        return ld

    def bind_output(ld: LayerDesc, blob: int, C: int, H: int, W: int):
        """gives the fused layer its output tensor: private, or a slice of a concat buffer"""
        l = layers.index(ld)
        if blob in concat_of:
            cid, off = concat_of[blob]
            if cid not in concat_tensor:
                cc, ch, cw = concat_shape[cid]
                tensors.append(TensorDesc(cc, ch, cw, -1, f"concat{cid}"))
                concat_tensor[cid] = len(tensors) - 1
                blob_t[concat_out_blob[cid]] = (concat_tensor[cid], 0)
            ld.out_tensor, ld.out_ch0 = concat_tensor[cid], off
            branch_tail[l], concat_layer[l] = 1, cid
        else:
            tensors.append(TensorDesc(C, H, W, l + 1, f"out{l}"))
            ld.out_tensor, ld.out_ch0 = len(tensors) - 1, 0
        blob_t[blob] = (ld.out_tensor, ld.out_ch0)
        blob_layer[blob] = l

    def rebind(ld: LayerDesc, old_blob: int, new_blob: int, C: int, H: int, W: int):
        """the layer absorbed a following op: its result is now `new_blob` with a new shape"""
        l = layers.index(ld)
        if old_blob in concat_of or new_blob in concat_of:
            if old_blob in concat_of:
                raise ValueError(f"layer {l}: an op after a concat input cannot be fused")
            # result goes into a concat buffer: drop the private tensor made for old_blob (it is the last one)
            assert ld.out_tensor == len(tensors) - 1
            tensors.pop()
            del blob_t[old_blob]
            bind_output(ld, new_blob, C, H, W)
            del blob_layer[old_blob]
            return
        t = tensors[ld.out_tensor]
        t.C, t.H, t.W = C, H, W
        blob_t[new_blob] = blob_t.pop(old_blob)
        blob_layer[new_blob] = blob_layer.pop(old_blob)

    for i, L in enumerate(fn.layers):
        conv, fc, pool, elt = L.op(OP_CONV), L.op(OP_FC), L.op(OP_POOL), L.op(OP_ELTWISE)
        if L.op(OP_SOFTMAX) or L.op(OP_CONCAT) or L.op(OP_FLATTEN):
            if L.op(OP_FLATTEN) or L.op(OP_SOFTMAX):
                if L.in_ids[0] in blob_t and not L.op(OP_SOFTMAX):
                    blob_t[L.out_ids[0]] = blob_t[L.in_ids[0]]
                    if L.in_ids[0] in blob_layer:
                        blob_layer[L.out_ids[0]] = blob_layer[L.in_ids[0]]
            continue
        if conv or fc:
            if len(L.in_ids) != 1 or L.in_ids[0] not in blob_t:
                raise ValueError(f"fpga layer {L.id}: convolution input blob {L.in_ids} is not available")
            tin, ch0 = blob_t[L.in_ids[0]]
            if ch0 != 0:
                raise ValueError(f"fpga layer {L.id}: reading a slice of a concat buffer is not supported")
            ti = tensors[tin]
            _, oc, oh, ow = L.out_shapes[0]
            if conv:
                p = conv.p
                if not (p["kh"] == p["kw"] and p["sh"] == p["sw"] and p["pad_l"] == p["pad_r"] == p["pad_t"] == p["pad_b"]):
                    raise ValueError(f"fpga layer {L.id}: only square kernels / strides / symmetric padding exist in the runtime")
                k, pad, stride, C = p["kh"], p["pad_l"], p["sh"], L.in_shapes[0][1]
                f7 = 0
                if i == 0 and seven:
                    k, pad, stride, C, f7 = 3, 0, 1, 27, 1
                ld = new_layer(in_tensor=tin, C=C, N=p["out_c"], k=k, pad=pad, stride=stride, OH=oh, OW=ow, PH=oh, PW=ow,
                               bias_en=int(p["bias"]), first_layer_7x7=f7)
            else:
                if ti.H != 1 or ti.W != 1:
                    raise ValueError(f"fpga layer {L.id}: fully connected layers run as 1x1 convolutions on a 1x1 map")
                ld = new_layer(in_tensor=tin, C=ti.C, N=fc.p["out_c"], k=1, OH=1, OW=1, PH=1, PW=1, bias_en=int(fc.p["bias"]))
            if (ti.H + 2 * ld.pad - ld.k) // ld.stride + 1 != ld.OH:
                raise ValueError(f"fpga layer {L.id}: output height {ld.OH} does not follow from the geometry")
            ld.bn_en = 1 if L.op(OP_BN) else 0
            if ld.bn_en and not L.op(OP_SCALE):
                raise ValueError(f"fpga layer {L.id}: BatchNorm without Scale (the blob stream carries gamma/beta)")
            ld.relu = 1 if L.op(OP_RELU) else 0
            bind_output(ld, L.out_ids[0], ld.N, oh, ow)
            continue
        if pool:
            p = pool.p
            src = L.in_ids[0]
            _, oc, oh, ow = L.out_shapes[0]
            prod = blob_layer.get(src)
            sole = len(consumers.get(src, [])) == 1
            is_gap = p["method"] == 1
            if is_gap:
                ish = L.in_shapes[0]
                if not (p["global_pool"] or (p["kh"], p["kw"]) == (ish[2], ish[3])) or (ish[2], ish[3]) != (7, 7):
                    raise ValueError(f"fpga layer {L.id}: only the 7x7 global average exists (full_size_pool.cl:115-118)")
                if prod is None or not sole or layers[prod].gap or layers[prod].ipool:
                    raise ValueError(f"fpga layer {L.id}: the global average must follow a convolution it can fuse with")
                ld = layers[prod]
                ld.gap = 1
                rebind(ld, src, L.out_ids[0], ld.N, 1, 1)
                continue
            if (p["kh"], p["kw"]) != (3, 3) or p["sh"] not in (1, 2) or p["method"] != 0:
                raise ValueError(f"fpga layer {L.id}: the runtime pools 3x3 max only (pool.cl:194-199)")
            ppad = p["pad_l"] if max_pool_pad is None else (max_pool_pad if p["sh"] == 2 else p["pad_l"])
            if prod is not None and sole and not layers[prod].pool and not layers[prod].gap and not layers[prod].ipool \
                    and layers[prod].add_tensor < 0 and layers[prod].stride == 1:
                ld = layers[prod]
                ld.pool, ld.pool_stride, ld.pool_pad, ld.PH, ld.PW = 1, p["sh"], ppad, oh, ow
                rebind(ld, src, L.out_ids[0], ld.N, oh, ow)
            else:
                if (p["sh"], p["pad_l"]) != (1, 1):
                    raise ValueError(f"fpga layer {L.id}: a stand-alone pool must be 3x3 / stride 1 / pad 1 (retriever.cl:285-302)")
                tin, ch0 = blob_t[src]
                if ch0 != 0:
                    raise ValueError(f"fpga layer {L.id}: pooling a slice of a concat buffer is not supported")
                ti = tensors[tin]
                ld = new_layer(in_tensor=tin, C=ti.C, N=ti.C, k=3, pad=1, stride=1, OH=ti.H, OW=ti.W, pool=1, pool_stride=1,
                               pool_pad=1, PH=oh, PW=ow, ipool=1)
                bind_output(ld, L.out_ids[0], ti.C, oh, ow)
            continue
        if elt:
            if elt.p["method"] != 1 or len(L.in_ids) != 2:
                raise ValueError(f"fpga layer {L.id}: only the two-operand eltwise SUM exists (feature_writer.cl:124)")
            a, b = L.in_ids
            la, lb = blob_layer.get(a, -1), blob_layer.get(b, -1)
            main, other = (a, b) if la > lb else (b, a)
            lm = blob_layer.get(main, -1)
            if lm < 0 or len(consumers.get(main, [])) != 1 or layers[lm].pool or layers[lm].gap or layers[lm].ipool \
                    or layers[lm].add_tensor >= 0:
                raise ValueError(f"fpga layer {L.id}: the later operand of the sum must be a plain convolution used only here")
            ld = layers[lm]
            ot, och0 = blob_t[other]
            if och0 != 0:
                raise ValueError(f"fpga layer {L.id}: residual operand inside a concat buffer is not supported")
            ld.add_tensor = ot
            ld.add_relu = 1 if L.op(OP_RELU) else 0
            rebind(ld, main, L.out_ids[0], ld.N, ld.PH, ld.PW)
            continue
        if L.op(OP_RELU) and len(L.ops) == 1:
            raise ValueError(f"fpga layer {L.id}: a stand-alone ReLU should have been merged by caffe2fpga (hebing_op)")
        raise ValueError(f"fpga layer {L.id}: unsupported op group {[OP_NAMES.get(o.type, o.type) for o in L.ops]}")

    # Q-table rows: 0 = image, l + 1 = output of layer l, then one row per concat buffer (quantization.cpp:36-50)
    nconv = len(layers)
    for cid, t in concat_tensor.items():
        tensors[t].q_row = nconv + 1 + cid
    for ld in layers:
        ld.q_in_row = tensors[ld.in_tensor].q_row
    may = [False] * len(tensors)
    may[0] = True
    for ld in layers:
        if ld.ipool:
            nonneg = not may[ld.in_tensor]
        else:
            nonneg = bool(ld.add_relu if ld.add_tensor >= 0 else ld.relu)
        if not nonneg:
            may[ld.out_tensor] = True
    for ld in layers:
        ld.in_may_be_m128 = 1 if may[ld.in_tensor] else 0
    return NetDesc(name=name, tensors=tensors, layers=layers, max_out_channel=max(t.C for t in tensors[1:]),
                   num_q_rows=nconv + 1 + ncat, input_c=ic, input_h=ih, input_w=iw, branch_tail=branch_tail,
                   concat_layer=concat_layer)


def load_fpganetwork(path: str, name: Optional[str] = None, max_pool_pad: Optional[int] = None) -> NetDesc:
    with open(path, "rb") as f:
        return to_netdesc(parse_fpganetwork(f.read()), name, max_pool_pad)
