"""Host-side mirror of the reference runtime classes for the quantised-convolution hot path.

Reference interface being mirrored (names, argument meaning, order of operations):
  * `NetWork::Init(platform, model_file, q_file, image_file, num_images)`
    (`Runtime_Engine/cnn/host/src/network.cpp:22-38`): Quantization() -> LoadModel() ->
    FilterConvert() -> device buffers.
  * `Runner::Run()` (`runner.cpp:54-196`): load + transform + quantise the image(s), write the
    input buffer, run every layer on the device, read the last feature map back.
  * `Verify` / `Evaluation` (`network_helper.cpp:18-207`).
Device memory, streams and (for multi-GPU) `torch.distributed` come from PyTorch; all arithmetic on
the path runs in libtf2b200.so (hand-written sm_100a CUDA) through the C ABI of include/tf2b200.h.
Errors are raised as `Tf2bError` instead of the reference's print-and-exit (opencl.cpp:226-250).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

from . import capi, formats
from .capi import Tf2bError
from .netdesc import NetDesc


def _layer_descs(net: NetDesc):
    arr = (capi.LayerDescC * net.num_layers)()
    for i, ld in enumerate(net.layers):
        a = arr[i]
        for name, _ in capi.LayerDescC._fields_:
            setattr(a, name, int(getattr(ld, name)))
    tarr = (capi.TensorDescC * len(net.tensors))()
    for i, t in enumerate(net.tensors):
        tarr[i].C, tarr[i].H, tarr[i].W = t.C, t.H, t.W
    return tarr, arr


class NetWork:
    """Owns the engine handle and the model (reference: class NetWork, network.h / network.cpp)."""

    def __init__(self, net: NetDesc, device: int = 0):
        self.net = net
        self.device = device
        self.q: Optional[np.ndarray] = None
        self.model = None  # per layer (codes, params) as LoadModel produced them
        self._lib = capi.load()
        self._h = C.c_void_p()
        tarr, larr = _layer_descs(net)
        rc = self._lib.tf2b_create(tarr, len(net.tensors), larr, net.num_layers, device, C.byref(self._h))
        if rc != capi.TF2B_OK:
            raise Tf2bError(rc, self._lib.tf2b_last_error(None).decode())
        self.max_images = 0
        self._result_tensor: Optional[int] = None

    # -- error plumbing -------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != capi.TF2B_OK:
            raise Tf2bError(rc, self._lib.tf2b_last_error(self._h).decode())

    @property
    def handle(self):
        return self._h

    # -- NetWork::InitNetwork (network.cpp:40-98) ----------------------------------------
    def Init(self, model_file: str, q_file: str, max_images: int = 1, variant: int = capi.VARIANT_AUTO):
        """Quantization(q_file) -> LoadModel(model_file) -> upload.  `model_file` is the float32
        blob (param.bin / fpgamodel.bin)."""
        self.q = formats.parse_q_file(self.net, q_file)
        self.model = formats.load_float_blob_file(self.net, model_file, self.q)
        self._upload(max_images, variant)
        return True

    def InitFromMemory(self, model_blob: bytes, q_text: str, max_images: int = 1,
                       variant: int = capi.VARIANT_AUTO):
        self.q = formats.parse_q_text(self.net, q_text)
        self.model = formats.load_float_blob(self.net, model_blob, self.q)
        self._upload(max_images, variant)
        return True

    def Init4bit(self, model4_file: str, q_file: str, max_images: int = 1, variant: int = capi.VARIANT_AUTO):
        """Like Init, from the 4-bit model file of TransForm_Kit/Compression/compress_net/4bit_data_format.txt
        (formats.float_blob_to_4bit writes one): short-coded power-of-two weights, float biases / BatchNorm."""
        with open(model4_file, "rb") as f:
            blob = formats.float_blob_from_4bit(self.net, f.read())
        with open(q_file, "r") as f:
            return self.InitFromMemory(blob, f.read(), max_images, variant)

    def set_stem_chunk(self, on: bool):
        """Chunked L2-resident stem on / off (default off: measured slower); before Init*."""
        self._check(self._lib.tf2b_set_stem_chunk(self._h, 1 if on else 0))

    def set_weight_staging(self, mode: int):
        """capi.WEIGHTS_PLANES (default) / capi.WEIGHTS_PACKED4 for the tensor-core kernel's resident-weight layers;
        before Init*."""
        self._check(self._lib.tf2b_set_weight_staging(self._h, mode))

    def InitFromCodes(self, model, q: Optional[np.ndarray], max_images: int = 1,
                      variant: int = capi.VARIANT_AUTO):
        """`model`: per layer (codes uint8 [N][C][k][k], params int32 [N][3]) or (None, None)."""
        self.q = q
        self.model = model
        self._upload(max_images, variant)
        return True

    def _upload(self, max_images: int, variant: int):
        for l, (codes, params) in enumerate(self.model):
            if codes is None:
                continue
            ld = self.net.layers[l]
            codes = np.ascontiguousarray(codes, dtype=np.uint8)
            params = np.ascontiguousarray(params, dtype=np.int32)
            if codes.shape != (ld.N, ld.C, ld.k, ld.k) or params.shape != (ld.N, 3):
                raise ValueError(f"layer {l}: codes {codes.shape} / params {params.shape} do not match the tables")
            self._check(self._lib.tf2b_load_layer(self._h, l, codes.ctypes.data, params.ctypes.data))
        self._check(self._lib.tf2b_set_variant(self._h, variant))
        self._check(self._lib.tf2b_finalize(self._h, max_images))
        self.max_images = max_images

    def InitFromBlob(self, blob_dev_ptr: int, blob_bytes: int, max_images: int, variant: int = capi.VARIANT_AUTO,
                     stream: int = 0, q: Optional[np.ndarray] = None):
        """Receiving side of the init-time weight broadcast (no model file on this rank).  `q` = the Q table
        (Runner.Run needs its first entry to quantise float images)."""
        self.q = q
        self._check(self._lib.tf2b_import_weight_blob(self._h, blob_dev_ptr, blob_bytes, stream))
        self._check(self._lib.tf2b_set_variant(self._h, variant))
        self._check(self._lib.tf2b_finalize(self._h, max_images))
        self.max_images = max_images

    def set_variant(self, variant: int):
        self._check(self._lib.tf2b_set_variant(self._h, variant))

    def set_graph(self, on: bool):
        """CUDA-graph executor on (default) / off (kernel-by-kernel launches)."""
        self._check(self._lib.tf2b_set_graph(self._h, 1 if on else 0))

    def set_result(self, tensor: int):
        """Which tensor the run calls return (default: the last layer's output) — the reference's way of
        verifying an inner layer is to rebuild with a shorter table (CONCAT_LAYER_DEBUG, network_helper.cpp:19-23)."""
        self._check(self._lib.tf2b_set_result(self._h, tensor))
        self._result_tensor = tensor

    def layer_kernels(self) -> List[str]:
        return [self._lib.tf2b_layer_kernel(self._h, l).decode() for l in range(self.net.num_layers)]

    def layer_modes(self, n_images: int = 0) -> List[str]:
        """Launch plan of every layer (tile sizes, staging mode, CTA-pair MMA ...) for `n_images`."""
        return [self._lib.tf2b_layer_mode(self._h, l, n_images).decode() for l in range(self.net.num_layers)]

    def weight_blob_bytes(self) -> int:
        n = self._lib.tf2b_weight_blob_bytes(self._h)
        if n < 0:
            self._check(int(n))
        return int(n)

    def export_weight_blob(self, dst_dev_ptr: int, stream: int = 0):
        self._check(self._lib.tf2b_export_weight_blob(self._h, dst_dev_ptr, stream))

    def set_profile(self, on: bool):
        self._check(self._lib.tf2b_set_profile(self._h, 1 if on else 0))

    def get_profile(self):
        """(conv_ms, layer_ms) float32 arrays of the last profiled run."""
        n = self.net.num_layers
        conv = np.zeros(n, np.float32)
        layer = np.zeros(n, np.float32)
        self._check(self._lib.tf2b_get_profile(self._h, conv.ctypes.data, layer.ctypes.data, n))
        return conv, layer

    def last_launches(self) -> int:
        return int(self._lib.tf2b_last_launches(self._h))

    def CleanUp(self):
        if self._h:
            self._lib.tf2b_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.CleanUp()
        except Exception:
            pass


class Runner:
    """Runs batches through a NetWork (reference: class Runner, runner.h / runner.cpp)."""

    def __init__(self, network: NetWork):
        self.network = network
        self._lib = network._lib

    @property
    def net(self) -> NetDesc:
        return self.network.net

    def result_shape(self):
        tid = getattr(self.network, "_result_tensor", None)
        t = self.net.tensors[self.net.result_tensor() if tid is None else tid]
        return (t.C, t.H, t.W)

    def _check_input_shape(self, shape, raw224: bool, in_layout: int):
        """The C ABI takes plain pointers: a wrong-shaped buffer would be read out of bounds."""
        t0 = self.net.tensors[0]
        if raw224:
            want = (3, 224, 224)
        else:
            want = (t0.C, t0.H, t0.W) if in_layout == capi.LAYOUT_CHW else (t0.H, t0.W, t0.C)
        if tuple(shape) != want:
            raise ValueError(f"input images have shape {tuple(shape)}, the network expects {want}")

    def _check_out_size(self, out, B: int):
        C_, H_, W_ = self.result_shape()
        n = out.numel() if hasattr(out, "numel") else np.asarray(out).size
        if n != B * C_ * H_ * W_:
            raise ValueError(f"output buffer holds {n} values, the result of {B} images has {B * C_ * H_ * W_}")

    # -- device-resident path ---------------------------------------------------------------
    def run_device(self, x, out=None, raw224: bool = False, in_layout: int = capi.LAYOUT_CHW,
                   out_layout: int = capi.LAYOUT_CHW, stream=None):
        """x: torch int8 CUDA tensor, either raw quantised images [B,3,224,224] (raw224=True; the
        device applies feature_trans) or tensor-0 images [B,C,H,W] / [B,H,W,C]."""
        import torch
        if not (x.is_cuda and x.dtype == torch.int8 and x.is_contiguous()):
            raise ValueError("run_device: x must be a contiguous int8 CUDA tensor")
        if x.device.index != self.network.device:
            raise ValueError(f"run_device: x lives on {x.device}, the engine on cuda:{self.network.device}")
        B = x.shape[0]
        self._check_input_shape(tuple(x.shape[1:]), raw224, in_layout)
        C_, H_, W_ = self.result_shape()
        if out is None:
            shape = (B, C_, H_, W_) if out_layout == capi.LAYOUT_CHW else (B, H_, W_, C_)
            out = torch.empty(shape, dtype=torch.int8, device=x.device)
        elif not (out.is_cuda and out.device == x.device and out.dtype == torch.int8 and out.is_contiguous()
                  and out.numel() == B * C_ * H_ * W_):
            raise ValueError(f"run_device: out must be a contiguous int8 tensor of {B * C_ * H_ * W_} elements on {x.device}")
        st = stream if stream is not None else torch.cuda.current_stream(x.device)
        sp = C.c_void_p(st.cuda_stream)
        if raw224:
            rc = self._lib.tf2b_run_raw224(self.network.handle, x.data_ptr(), B, out.data_ptr(), out_layout, sp)
        else:
            rc = self._lib.tf2b_run(self.network.handle, x.data_ptr(), in_layout, B, out.data_ptr(), out_layout, sp)
        self.network._check(rc)
        return out

    # -- host-buffer path (the reference-facing call: H2D + run + D2H inside) ----------------
    def run_host(self, x_host, out_host=None, raw224: bool = False, in_layout: int = capi.LAYOUT_CHW,
                 out_layout: int = capi.LAYOUT_CHW):
        """x_host / out_host: int8 numpy arrays or CPU torch tensors (pinned for async copies)."""
        xp, B = _host_ptr(x_host)
        self._check_input_shape(tuple(x_host.shape[1:]), raw224, in_layout)
        C_, H_, W_ = self.result_shape()
        if out_host is None:
            out_host = np.empty((B, C_, H_, W_) if out_layout == capi.LAYOUT_CHW else (B, H_, W_, C_), dtype=np.int8)
        op, _ = _host_ptr(out_host)
        self._check_out_size(out_host, B)
        if raw224:
            rc = self._lib.tf2b_run_raw224_host(self.network.handle, xp, B, op, out_layout)
        else:
            rc = self._lib.tf2b_run_host(self.network.handle, xp, in_layout, B, op, out_layout)
        self.network._check(rc)
        return out_host

    # -- pipelined host path: EnqueueKernels now, WaitForAllKernels later (runner.cpp:32,183) ------
    def submit_host(self, x_host, out_host, slot: int, out_layout: int = capi.LAYOUT_CHW, raw224: bool = True,
                    in_layout: int = capi.LAYOUT_CHW):
        """Enqueue H2D + run + D2H of one int8 batch on `slot` (0/1); pinned buffers.  raw224: raw quantised
        3x224x224 images (the device applies feature_trans), else tensor-0 images in `in_layout`."""
        xp, B = _host_ptr(x_host)
        self._check_input_shape(tuple(x_host.shape[1:]), raw224, in_layout)
        op, _ = _host_ptr(out_host)
        self._check_out_size(out_host, B)
        if raw224:
            rc = self._lib.tf2b_submit_raw224_host(self.network.handle, xp, B, op, out_layout, slot)
        else:
            rc = self._lib.tf2b_submit_host(self.network.handle, xp, in_layout, B, op, out_layout, slot)
        self.network._check(rc)

    def wait(self, slot: int):
        self.network._check(self._lib.tf2b_wait(self.network.handle, slot))

    # -- Runner::Run (runner.cpp:54-196) ------------------------------------------------------
    def Run(self, images_f32: np.ndarray) -> np.ndarray:
        """Float images [B][3][224][224] (mean-subtracted, as the reference's .bin files) ->
        int8 output of the last layer [B][C][H][W].  Quantisation follows runner.cpp:158-164; the
        7x7 -> 3x3 space-to-depth transform of input_loader.cpp:27-73 runs on the device."""
        q = self.network.q
        if q is None:
            raise RuntimeError("NetWork has no Q table (use Init / InitFromMemory)")
        raw = formats.quantize_input(images_f32, int(q[0, 0]))
        t0 = self.net.tensors[0]
        if (t0.C, t0.H, t0.W) == (27, 114, 114):
            return self.run_host(np.ascontiguousarray(raw), raw224=True)
        return self.run_host(np.ascontiguousarray(raw))

    # -- debug taps -----------------------------------------------------------------------------
    def read_tensor(self, tensor: int, n_images: int, layout: int = capi.LAYOUT_CHW):
        import torch
        t = self.net.tensors[tensor]
        shape = (n_images, t.C, t.H, t.W) if layout == capi.LAYOUT_CHW else (n_images, t.H, t.W, t.C)
        out = torch.empty(shape, dtype=torch.int8, device=f"cuda:{self.network.device}")
        st = torch.cuda.current_stream(out.device)
        self.network._check(self._lib.tf2b_read_tensor(self.network.handle, tensor, n_images, out.data_ptr(),
                                                       layout, C.c_void_p(st.cuda_stream)))
        return out

    def dump_acc(self, layer: int, n_images: int):
        import torch
        ld = self.net.layers[layer]
        out = torch.empty((n_images, ld.N, ld.OH, ld.OW), dtype=torch.int32, device=f"cuda:{self.network.device}")
        st = torch.cuda.current_stream(out.device)
        self.network._check(self._lib.tf2b_dump_acc(self.network.handle, layer, n_images, out.data_ptr(),
                                                    C.c_void_p(st.cuda_stream)))
        return out


def _host_ptr(a):
    try:
        import torch
        if isinstance(a, torch.Tensor):
            assert not a.is_cuda and a.dtype == torch.int8 and a.is_contiguous()
            return a.data_ptr(), a.shape[0]
    except ImportError:
        pass
    a = np.asarray(a)
    assert a.dtype == np.int8 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data, a.shape[0]


# -- network_helper.cpp:18-207 ------------------------------------------------------------------
def Verify(output_int8: np.ndarray, expect_f32: np.ndarray, q_row: np.ndarray) -> float:
    """network_helper.cpp:120-139: relative L1 error of the dequantised last layer vs a float dump.
    output [C][H][W] int8; q_row holds -Q per channel (so value = int8 * 2^q)."""
    C_ = output_int8.shape[0]
    scale = np.exp2(-q_row[:C_].astype(np.float64)).reshape(C_, *([1] * (output_int8.ndim - 1)))
    exp = expect_f32.reshape(output_int8.shape).astype(np.float64) * scale
    err = np.abs(output_int8.astype(np.float64) - exp).sum()
    tot = np.abs(exp).sum()
    return float(err / tot) if tot else float("inf")


def Evaluation(output_int8: np.ndarray, q_row: np.ndarray, top: int = 5):
    """network_helper.cpp:143-207: dequantise (float32: feature = int8 / (1 << Q)), softmax, top-5 as
    (label, probability).  Ties are ordered like the reference's five bubble passes with a strict
    `>` (:187-193): among equal features the HIGHER label ranks first."""
    x = output_int8.reshape(-1).astype(np.float32)
    trans = np.exp2(-q_row[:x.size].astype(np.float64)).astype(np.float32)        # 1 << (-current_q)
    feat = (x / trans).astype(np.float32)
    sum_exp = np.float32(0)
    with np.errstate(over="ignore"):                                               # features > 88: inf, like the float there
        for v in np.exp(feat.astype(np.float64)):                                  # float sum_exp += exp(double)
            sum_exp = np.float32(np.float64(sum_exp) + v)
    order = np.lexsort((np.arange(x.size), feat))[::-1][:top]                      # feature desc, label desc
    return [(int(i), float(np.exp(np.float64(feat[i])) / np.float64(sum_exp))) for i in order]
