"""Network description: the TF2 per-layer tables resolved into a tensor graph.

The reference describes a network by compile-time tables `k*[NUM_CONVOLUTIONS]`
(`Runtime_Engine/cnn/host/inc/resnet50.h:119-1368`, `googlenet.h:47-1304`) whose tensor plumbing is
implicit in FPGA addresses: `kInputLayer` is a row index of the Q table (row 0 = image, row l+1 =
output of layer l, rows > NUM_CONVOLUTIONS = concat buffers, `quantization.cpp:36-50`), the
residual operand is whatever an earlier layer left at `kDDRReadBase` (`feature_writer.cl:91-107`),
concat is a channel offset `kNStart` into a shared buffer (`feature_writer.cl:109-137`).
`NetDesc.from_tables` turns that into explicit tensor ids so the same tables drive the CUDA engine
and the CPU oracle.
"""
from __future__ import annotations

from dataclasses import dataclass, field, asdict
from typing import Dict, List, Optional

from .header_tables import HeaderTables, parse_header_file


@dataclass
class TensorDesc:
    C: int
    H: int
    W: int
    q_row: int = -1          # row of the Q table that holds this tensor's per-channel Q
    name: str = ""


@dataclass
class LayerDesc:
    name: str
    in_tensor: int
    out_tensor: int
    out_ch0: int
    add_tensor: int
    C: int
    N: int
    k: int
    pad: int
    stride: int
    OH: int
    OW: int
    relu: int
    pool: int
    pool_stride: int
    pool_pad: int
    PH: int
    PW: int
    add_relu: int
    gap: int
    ipool: int
    bias_en: int
    bn_en: int
    in_may_be_m128: int = 0
    q_in_row: int = 0        # kInputLayer (Q row of the input)
    q_out_row: int = 0       # layer index + 1
    first_layer_7x7: int = 0  # layer 0 stored as 7x7x3 in the model file (model_loader.cpp:148-151)


@dataclass
class NetDesc:
    name: str
    tensors: List[TensorDesc]
    layers: List[LayerDesc]
    max_out_channel: int
    num_q_rows: int
    input_c: int = 3
    input_h: int = 224
    input_w: int = 224
    # Q rows that alias other rows: (layer idx of a branch tail) handled in formats.parse_q_file
    branch_tail: List[int] = field(default_factory=list)
    concat_layer: List[int] = field(default_factory=list)

    @property
    def num_layers(self) -> int:
        return len(self.layers)

    def result_tensor(self) -> int:
        return self.layers[-1].out_tensor

    # ------------------------------------------------------------------
    @staticmethod
    def from_header(path: str, name: Optional[str] = None) -> "NetDesc":
        t = parse_header_file(path)
        if name is None:
            import os
            name = os.path.splitext(os.path.basename(path))[0]
        return NetDesc.from_tables(t, name)

    @staticmethod
    def from_tables(t: HeaderTables, name: str) -> "NetDesc":
        L = int(t["NUM_LAYER"])
        nconv = int(t["NUM_CONVOLUTIONS"])
        max_oc = int(t["MAX_OUT_CHANNEL"])

        def tab(key, default=0):
            v = list(t.get(key, []))
            # the shipped googlenet.h leaves a few tables short; C zero-initialises the rest
            return v + [default] * (L - len(v))

        kFilter, kPadW, kPadH = tab("kFilterSize"), tab("kPadWidth"), tab("kPadHeight")
        kInW, kInH = tab("kInputWidth"), tab("kInputHeight")
        kOutW, kOutH = tab("kOutputWidth"), tab("kOutputHeight")
        kInC, kOutC = tab("kInputChannels"), tab("kOutputChannels")
        kStride = tab("kConvStride", 1)
        kPool, kPoolS2, kPoolPad = tab("kPoolEnable"), tab("kPoolStride2"), tab("kPoolPad")
        kPoolOW, kPoolOH = tab("kPoolOutputWidth"), tab("kPoolOutputHeight")
        kBias, kBn, kRelu = tab("kBiasEnable"), tab("kBnEnable"), tab("kReluEnable")
        kAdd, kAddRelu, kGap = tab("kAdditionEnable"), tab("kAdditionReluEnable"), tab("kEndPoolEnable")
        kDDRRead, kDDRWrite, kDDRWen = tab("kDDRReadBase"), tab("kDDRWriteBase"), tab("kDDRWriteEnable")
        kNStart = tab("kNStart")
        kIpool, kTail, kConcat = tab("kIpoolEnable"), tab("kBranchTail"), tab("kConcatLayer")
        kInputLayer = tab("kInputLayer")

        n_q_rows = int(t.get("NUM_Q_LAYERS", nconv + 1))
        # tensor 0 = network input as layer 0 sees it; tensor l+1 = private output of layer l;
        # concat buffers follow.
        tensors: List[TensorDesc] = [TensorDesc(kInC[0], kInH[0], kInW[0], 0, "input")]
        row_to_tensor: Dict[int, int] = {0: 0}
        layers: List[LayerDesc] = []
        concat_tensor: Dict[int, int] = {}

        def out_hw(l):
            if kGap[l]:
                return 1, 1
            return kPoolOH[l], kPoolOW[l]

        for l in range(L):
            oh, ow = out_hw(l)
            if kTail[l]:
                cid = kConcat[l]
                if cid not in concat_tensor:
                    # channel count = max kNEnd over the tails of this concat
                    cmax = max(kNStart[j] + kOutC[j] for j in range(L) if kTail[j] and kConcat[j] == cid)
                    tensors.append(TensorDesc(cmax, oh, ow, nconv + 1 + cid, f"concat{cid}"))
                    concat_tensor[cid] = len(tensors) - 1
                    row_to_tensor[nconv + 1 + cid] = concat_tensor[cid]
                out_t = concat_tensor[cid]
                # a tail's own Q row (l+1) still exists (quantization.cpp:45-49 fills both)
                row_to_tensor[l + 1] = -1
            else:
                tensors.append(TensorDesc(kOutC[l], oh, ow, l + 1, f"out{l}"))
                out_t = len(tensors) - 1
                row_to_tensor[l + 1] = out_t
            layers.append(LayerDesc(
                name=f"layer{l}", in_tensor=-1, out_tensor=out_t, out_ch0=kNStart[l] if kTail[l] else 0,
                add_tensor=-1, C=kInC[l], N=kOutC[l], k=kFilter[l], pad=kPadW[l], stride=kStride[l],
                OH=0, OW=0, relu=kRelu[l], pool=kPool[l], pool_stride=2 if kPoolS2[l] else 1,
                pool_pad=kPoolPad[l], PH=kPoolOH[l], PW=kPoolOW[l], add_relu=kAddRelu[l], gap=kGap[l],
                ipool=kIpool[l], bias_en=kBias[l], bn_en=kBn[l], q_in_row=kInputLayer[l], q_out_row=l + 1,
                first_layer_7x7=1 if (l == 0 and int(t.get("FIRST_FILTER_SIZE", 0)) == 7 and kFilter[0] == 3) else 0))

        # resolve inputs, conv output sizes, residual operands
        ddr_owner: Dict[int, int] = {}  # DDR base -> tensor id of the last layer that wrote there
        for l, ld in enumerate(layers):
            row = kInputLayer[l]
            tin = row_to_tensor.get(row, None)
            if tin is None or tin < 0:
                raise ValueError(f"layer {l}: input Q row {row} is not a tensor")
            ld.in_tensor = tin
            ti = tensors[tin]
            if kPadH[l] != kPadW[l]:
                raise ValueError(f"layer {l}: asymmetric padding is outside the reference's use")
            if ld.ipool:
                ld.C = ld.N = ti.C
                ld.OH, ld.OW = ti.H, ti.W
                ld.k, ld.pad, ld.stride = 3, 1, 1
            else:
                ld.OH = (ti.H + 2 * ld.pad - ld.k) // ld.stride + 1
                ld.OW = (ti.W + 2 * ld.pad - ld.k) // ld.stride + 1
                if not ld.pool:
                    # stride-2 layers carry the stride-1 width in kOutputWidth and the real size
                    # in kPoolOutputWidth (pool_tail.cl:128-154 drops the odd columns)
                    if (ld.PH, ld.PW) != (ld.OH, ld.OW):
                        raise ValueError(f"layer {l}: tables give {ld.PH}x{ld.PW}, geometry {ld.OH}x{ld.OW}")
                else:
                    if ld.stride != 1:
                        raise ValueError(f"layer {l}: conv stride with pooling is outside the reference's use")
                    if (kOutH[l], kOutW[l]) != (ld.OH, ld.OW):
                        raise ValueError(f"layer {l}: kOutput size mismatch")
            if kAdd[l]:
                base = kDDRRead[l]
                if base not in ddr_owner:
                    raise ValueError(f"layer {l}: residual reads DDR base {base} that nothing wrote")
                ld.add_tensor = ddr_owner[base]
            if kDDRWen[l]:
                ddr_owner[kDDRWrite[l]] = ld.out_tensor

        # which tensors may hold -128: the image, and outputs of layers without a trailing ReLU
        may_m128 = [False] * len(tensors)
        may_m128[0] = True
        for ld in layers:
            if ld.ipool:
                nonneg = not may_m128[ld.in_tensor]
            else:
                last_relu = ld.add_relu if ld.add_tensor >= 0 else ld.relu
                nonneg = bool(last_relu)
            if not nonneg:
                may_m128[ld.out_tensor] = True
        for ld in layers:
            ld.in_may_be_m128 = 1 if may_m128[ld.in_tensor] else 0

        return NetDesc(name=name, tensors=tensors, layers=layers, max_out_channel=max_oc,
                       num_q_rows=max(n_q_rows, nconv + 1 + (max(concat_tensor) + 1 if concat_tensor else 0)),
                       input_c=int(t.get("INPUT_IMAGE_C", 3)), input_h=int(t.get("INPUT_IMAGE_H", 224)),
                       input_w=int(t.get("INPUT_IMAGE_W", 224)), branch_tail=kTail, concat_layer=kConcat)

    # ------------------------------------------------------------------
    def to_json(self) -> dict:
        return {"name": self.name, "max_out_channel": self.max_out_channel, "num_q_rows": self.num_q_rows,
                "input": [self.input_c, self.input_h, self.input_w],
                "branch_tail": self.branch_tail, "concat_layer": self.concat_layer,
                "tensors": [asdict(x) for x in self.tensors], "layers": [asdict(x) for x in self.layers]}

    @staticmethod
    def from_json(d: dict) -> "NetDesc":
        return NetDesc(name=d["name"], tensors=[TensorDesc(**x) for x in d["tensors"]],
                       layers=[LayerDesc(**x) for x in d["layers"]], max_out_channel=d["max_out_channel"],
                       num_q_rows=d["num_q_rows"], input_c=d["input"][0], input_h=d["input"][1],
                       input_w=d["input"][2], branch_tail=d.get("branch_tail", []),
                       concat_layer=d.get("concat_layer", []))

    def macs_per_image(self) -> int:
        """True-convolution MACs (layer 0 counted in its 7x7x3 form when it was transformed)."""
        total = 0
        for ld in self.layers:
            if ld.ipool:
                continue
            if ld.first_layer_7x7:
                total += ld.OH * ld.OW * ld.N * 49 * self.input_c
            else:
                total += ld.OH * ld.OW * ld.N * ld.C * ld.k * ld.k
        return total
