/*
 * tf2b200.h — C ABI of the B200-native TF2 quantised-convolution engine (libtf2b200.so).
 *
 * This is the drop-in boundary for the one hot path of TF2's Runtime_Engine/cnn.  In the reference
 * the host talks to the accelerator through an OpenCL program (`cnn.aocx`) — buffers created in
 * NetWork::InitBuffer (Runtime_Engine/cnn/host/src/network.cpp:100-150), kernel arguments set and
 * tasks enqueued in Runner::Run (Runtime_Engine/cnn/host/src/runner.cpp:54-196).  Each entry point
 * below names the reference interface it replaces.  Plain pointers and sizes only; no torch types.
 *
 * Error behaviour: the reference prints and exit()s (common/src/AOCLUtils/opencl.cpp:226-250).
 * Here every call returns TF2B_OK or a negative status and never exits; the text of the last
 * failure is available from tf2b_last_error().  There is no CPU fallback: on a machine without a
 * CUDA device tf2b_create fails with TF2B_ERR_CUDA.
 */
#ifndef TF2B200_H
#define TF2B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TF2B_OK 0
#define TF2B_ERR_ARG (-1)
#define TF2B_ERR_CUDA (-2)
#define TF2B_ERR_STATE (-3)
#define TF2B_ERR_NOMEM (-4)

/* kernel selection for tf2b_set_variant: which convolution kernels the executor may use */
#define TF2B_VARIANT_AUTO 0  /* tcgen05 INT8 MMA where the layer is a dense contraction, else shift */
#define TF2B_VARIANT_SHIFT 1 /* CUDA-core shift-accumulate kernel for every layer (BASELINE configs[1]) */
#define TF2B_VARIANT_MMA 2   /* force the tensor-core path on every layer it supports (configs[2]) */

/* how the tensor-core kernel receives the weights of layers whose weight slab stays resident in shared memory
 * (tf2b_set_weight_staging; the CUDA-core kernel ALWAYS reads packed 4-bit tiles and expands them on the fly) */
#define TF2B_WEIGHTS_PLANES 0  /* int8 +-2^e planes expanded at load time (default: fastest) */
#define TF2B_WEIGHTS_PACKED4 1 /* packed 4-bit tiles (4bit_data_format.txt's information content) fetched by TMA and */
                               /* expanded on the fly, once per CTA, into the swizzled K-major int8 slab             */

/* layouts accepted/produced at the boundary */
#define TF2B_LAYOUT_CHW 0 /* [image][C][H][W] — the reference host order (input_loader.cpp:76-97) */
#define TF2B_LAYOUT_HWC 1 /* [image][H][W][C] — the engine's native order */

/* Runtime_Engine/cnn/host/inc/types.h:39-43 (BiasBnParam): bias has 15-q_out fractional bits,
 * alpha 20, beta 15-q_out. */
typedef struct {
  int32_t bias;
  int32_t alpha;
  int32_t beta;
} tf2b_bias_bn;

/* One feature-map tensor of the network graph.  Tensor 0 is the network input as the first
 * convolution sees it (for ResNet50/GoogLeNet the 27-channel 114x114 space-to-depth form produced
 * by feature_trans, input_loader.cpp:27-73).  Concat outputs (googlenet.h kConcatLayer) are one
 * tensor written at channel offsets. */
typedef struct {
  int32_t C, H, W;
} tf2b_tensor_desc;

/* One fused layer = one row of the reference's k* tables (resnet50.h:119-1368), with the tensor
 * plumbing (kInputLayer / kDDRReadBase / kNStart / kConcatLayer) resolved to tensor ids. */
typedef struct {
  int32_t in_tensor;    /* kInputLayer (row of the Q table) resolved to a tensor id            */
  int32_t out_tensor;   /* tensor written (the layer's own tensor or a concat buffer)          */
  int32_t out_ch0;      /* kNStart: channel offset inside out_tensor                           */
  int32_t add_tensor;   /* residual operand (kAdditionEnable + kDDRReadBase resolved), or -1   */
  int32_t C;            /* kInputChannels                                                      */
  int32_t N;            /* kOutputChannels                                                     */
  int32_t k;            /* kFilterSize                                                         */
  int32_t pad;          /* kPadWidth == kPadHeight                                             */
  int32_t stride;       /* kConvStride                                                         */
  int32_t OH, OW;       /* convolution output size after the stride                            */
  int32_t relu;         /* kReluEnable                                                         */
  int32_t pool;         /* kPoolEnable: 3x3 max pool, taps outside the map read 0              */
  int32_t pool_stride;  /* kPoolStride2 ? 2 : 1                                                */
  int32_t pool_pad;     /* kPoolPad                                                            */
  int32_t PH, PW;       /* kPoolOutputHeight/Width (== OH,OW without pool)                     */
  int32_t add_relu;     /* kAdditionReluEnable                                                 */
  int32_t gap;          /* kEndPoolEnable: 7x7 global average                                  */
  int32_t ipool;        /* kIpoolEnable: pooling pseudo layer (no convolution)                 */
  int32_t in_may_be_m128; /* 1 if the input tensor can hold -128 (not produced behind a ReLU): */
                          /* the int8 negate quirk of pe.cl:32-34 then needs the exact path    */
} tf2b_layer_desc;

typedef struct tf2b_net tf2b_net;

/* Replaces OpenCLFPGA::Init + NetWork::InitNetwork's static tables (opencl_fpga.cpp:22-98,
 * network.cpp:40-98): builds an engine for the given layer graph on CUDA device `device`. */
int tf2b_create(const tf2b_tensor_desc* tensors, int n_tensors, const tf2b_layer_desc* layers,
                int n_layers, int device, tf2b_net** out);

/* Replaces the filter_buffer / bias_bn_buffer uploads of NetWork::InitBuffer
 * (network.cpp:118-147).  `codes` is LoadModel's output for this layer (model_loader.cpp:129-258):
 * [N][C][k][k] bytes, 0x40 = zero, bit7 = negative, low 5 bits = shift.  `params` is [N]. */
int tf2b_load_layer(tf2b_net* net, int layer, const uint8_t* codes, const tf2b_bias_bn* params);

/* 4-bit weight blob path (TransForm_Kit/Compression/compress_net/4bit_data_format.txt:1-44):
 * `nibbles` holds N*C*k*k 4-bit codes, two per byte, low nibble first, in [N][C][k][k] order
 * ((positive?8:0)|e, e in 0..6, value +-2^(min_exp+e); nibble 7 = 0.0).  q_in[C] / q_out[N] are
 * the (negated) Q rows used by LoadModel (model_loader.cpp:159-162).  Expansion to shift codes
 * follows Get_real (model_loader.cpp:98-126). */
int tf2b_load_layer_packed4(tf2b_net* net, int layer, const uint8_t* nibbles, int min_exp,
                            const int8_t* q_in, const int8_t* q_out, const tf2b_bias_bn* params);

/* Allocates activation storage for up to `max_images` images per call and freezes the plan. */
int tf2b_finalize(tf2b_net* net, int max_images);

int tf2b_set_variant(tf2b_net* net, int variant);

/* Chunked stem (default OFF — measured 4 % slower on ResNet50 B=256, kept as an option): with the raw 3x224x224 entry points and a first layer that is convolution + pool
 * (input_loader.cpp:27-73's transformed stem), batches of more than 32 images go through space-to-depth -> conv1 ->
 * max pool in chunks of 32 whose intermediates stay in the L2 cache.  Tensor 0 is then not kept for the whole batch;
 * tf2b_read_tensor(0) / tf2b_dump_acc(layer 0) re-create it from the last raw input (which must still be valid).
 * Call before tf2b_finalize. */
int tf2b_set_stem_chunk(tf2b_net* net, int on);

/* Weight staging of the tensor-core kernel (TF2B_WEIGHTS_*); call before tf2b_finalize. */
int tf2b_set_weight_staging(tf2b_net* net, int mode);

/* Executor: by default (on = 1) the layer sequence of a batch size is captured once into a CUDA graph (the
 * programmatic-dependent-launch edges between consecutive layers included) and replayed by every later run of
 * that size — the counterpart of the reference enqueueing all its kernels once per frame batch
 * (Runner::EnqueueKernels, runner.cpp:32) — and independent branches of the layer graph (inception branches,
 * the shortcut convolution of a residual block) run concurrently on the engine's side streams.
 * on = 0: kernel by kernel on the caller's stream; 2: graph, one stream; 3: several streams, no graph
 * (A/B measurements, debugging). */
int tf2b_set_graph(tf2b_net* net, int on);

/* Copies packed device weights/params from/to a flat device buffer so one rank can load the model
 * and the others receive it with a single NCCL broadcast. */
int64_t tf2b_weight_blob_bytes(tf2b_net* net);
int tf2b_export_weight_blob(tf2b_net* net, void* dev_dst, void* stream);
/* `blob_bytes` = length of the buffer at dev_src; the header carries a hash of the layer/tensor tables and
 * the arena size, every offset is checked against them before anything is read. */
int tf2b_import_weight_blob(tf2b_net* net, const void* dev_src, int64_t blob_bytes, void* stream);

/* Replaces Runner::Run's enqueue of the finite kernels (runner.cpp:166-183) with inputs already on
 * the device.  `in_dev`: int8 tensor-0 images in `in_layout`; `out_dev`: int8 result tensor
 * (`result_tensor` of tf2b_set_result, default = the last layer's output) in `out_layout`.
 * Asynchronous on `stream` (a cudaStream_t, may be NULL).
 * A handle owns ONE set of feature-map buffers: run calls on one handle must be issued from one thread and
 * onto one stream at a time (or ordered by the caller with events); use one handle per concurrent stream. */
int tf2b_run(tf2b_net* net, const int8_t* in_dev, int in_layout, int n_images, int8_t* out_dev,
             int out_layout, void* stream);

/* Same for the raw 3x224x224 int8 image (quantised by runner.cpp:158-164 but NOT yet through
 * feature_trans): the device performs the space-to-depth transform of input_loader.cpp:27-73. */
int tf2b_run_raw224(tf2b_net* net, const int8_t* raw_dev, int n_images, int8_t* out_dev,
                    int out_layout, void* stream);

/* Host-buffer form of the same call: the H2D write of runner.cpp:166 and the D2H read of
 * runner.cpp:195 happen inside (synchronous; pinned buffers give async copies). */
int tf2b_run_raw224_host(tf2b_net* net, const int8_t* raw_host, int n_images, int8_t* out_host,
                         int out_layout);
/* Pipelined form of tf2b_run_raw224_host — the analogue of Runner::EnqueueKernels (runner.cpp:32)
 * followed later by WaitForAllKernels (runner.cpp:183): returns as soon as the H2D copy, the run and
 * the D2H copy of this batch are enqueued on `slot` (0 or 1); tf2b_wait(slot) blocks until out_host
 * holds the result.  With two slots the copies of neighbouring batches overlap the computation.
 * Host buffers should be pinned; they must stay valid until tf2b_wait returns. */
int tf2b_submit_raw224_host(tf2b_net* net, const int8_t* raw_host, int n_images, int8_t* out_host,
                            int out_layout, int slot);
/* Same for tensor-0 images in `in_layout` (networks without the 7x7 -> 3x3 stem transform). */
int tf2b_submit_host(tf2b_net* net, const int8_t* in_host, int in_layout, int n_images, int8_t* out_host,
                     int out_layout, int slot);
int tf2b_wait(tf2b_net* net, int slot);

int tf2b_run_host(tf2b_net* net, const int8_t* in_host, int in_layout, int n_images,
                  int8_t* out_host, int out_layout);

int tf2b_set_result(tf2b_net* net, int tensor);

/* Debug taps (the reference exposes these only through -DPRINT_PE_OUTPUT printf, pe.cl:196-199):
 * copy tensor `tensor` of the last run ([n_images][C][H][W] or HWC) to `dst_dev`. */
int tf2b_read_tensor(tf2b_net* net, int tensor, int n_images, int8_t* dst_dev, int layout,
                     void* stream);
/* Re-runs layer `layer` of the last run and writes its int32 accumulators
 * (bias + sum of shifted features, before requantisation) to acc_dev as [image][N][OH][OW]. */
int tf2b_dump_acc(tf2b_net* net, int layer, int n_images, int32_t* acc_dev, void* stream);

/* Per-layer device timing (the reference gets its latency from OpenCL event profiling of the
 * sequencer kernel, runner.cpp:187-189).  With profiling on, every run records CUDA events on the
 * launching stream around each layer and around its convolution kernel; tf2b_get_profile waits
 * for the last run and returns milliseconds per layer (arrays of n_layers floats). */
int tf2b_set_profile(tf2b_net* net, int on);
int tf2b_get_profile(tf2b_net* net, float* conv_ms, float* layer_ms, int n_layers);

/* Number of kernels the last tf2b_run* call launched (bench.py's gpu_launches). */
int tf2b_last_launches(tf2b_net* net);
/* Name of the convolution kernel the plan uses for a layer ("shift", "mma", "none"). */
const char* tf2b_layer_kernel(tf2b_net* net, int layer);
/* Launch plan of `layer` for a batch of n_images (0 = max_images): "shift", "pool", or for the
 * tensor-core kernel e.g. "mma BN128 BK128 planes2 box fold ctapair stages5" (tile sizes, operand
 * staging mode — flat / box / halo / pixelpair —, resident weights, folded requantisation, CTA-pair
 * MMA).  Introspection for tests and profiles; the string lives until the next call for that layer. */
const char* tf2b_layer_mode(tf2b_net* net, int layer, int n_images);

const char* tf2b_last_error(tf2b_net* net);
const char* tf2b_version(void);
void tf2b_destroy(tf2b_net* net);

#ifdef __cplusplus
}
#endif
#endif
