/*
 * tf2_oracle.h — CPU oracle for the TF2 Runtime_Engine/cnn quantised-convolution hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or executed by the product
 * (tf2_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may call it, and only as the checker / the timed CPU baseline.
 *
 * It is a plain-C restatement of the reference algorithm; every function cites the reference
 * file:line it follows (paths relative to the reference repo root).
 *
 * Parity pinning.  The reference ships no integer golden vectors for this path (its Verify() is a
 * float tolerance check and the weights are not in the repository), so the restatement is pinned
 * against the reference's OWN sources compiled here by oracle/build_ref.sh into oracle/_ref/:
 *   - host loaders (model_loader.cpp, quantization.cpp, input_loader.cpp) compiled unmodified:
 *     Get_real, filter_trans, feature_trans, LoadModel, Quantization — equal bit for bit on
 *     ResNet50 / GoogLeNet / pruned ResNet50 (tests/test_formats.py, tests/golden/loader_golden.json);
 *   - the PE kernel device/src/pe.cl compiled as C behind a FIFO shim (oracle/ref_device/
 *     pe_harness.c): MUL exhaustively (65 536 cases) and PeFunction's MAC, bias seed, int32
 *     wrap-around, requantisation and clamp for 1x1 and 3x3 mode reductions of up to 128 steps —
 *     committed as tests/golden/pe_golden.npz (tests/test_pe_golden.py).
 *   - the post-PE kernels device/src/{relu,pool,pool_tail,feature_writer,full_size_pool}.cl compiled
 *     as C with each network's own tables (oracle/ref_device/post_harness.c) and run over EVERY
 *     layer of ResNet50, GoogLeNet and pruned ResNet50 on seeded random PE-output maps: ReLU, the
 *     3x3 max pool with zero border and its alignment, the conv stride in W, the ipool pseudo layer,
 *     the residual add through the DDR ping-pong, concat offsets and the 7x7 global average — equal
 *     bit for bit (tests/test_post_golden.py live in the build container; SHA-256 of the reference
 *     outputs committed as tests/golden/post_golden.json); the item counts the harness feeds equal
 *     the reference's own cycle constants (CONV_TOTAL_WRITE_CACHE, POOL_TOTAL_CYCLE, ...).
 *   - the WHOLE device pipeline device/src/cnn.cl compiled as C (oracle/ref_device/full_harness.c:
 *     input_reader -> filter_reader -> sequencer -> retriever -> 16 PEs -> relu -> pool -> pool_tail ->
 *     feature_writer, fed from InputConvert / FilterConvert device buffers): layer 0 of the three
 *     shipped networks (tests/test_full_layer0.py) and 20 one-layer networks generated from the
 *     reference's googlenet.h (oracle/ref_device/one_layer.py: 1x1, padded 3x3, stride 2, 5x5, ragged
 *     channel counts, 7..56-wide maps; tests/test_single_layer_ref.py) — the convolution geometry of
 *     sequencer.cl / retriever.cl; 0 mismatches, item counts equal the reference's cycle model.
 *   - host result readers network_helper.cpp (Verify, Evaluation) compiled unmodified: the feature_ddr
 *     tile addressing, the top-5 order including ties, and the relative-error figure
 *     (tests/test_verify_eval.py, tests/golden/eval_golden.json).
 *   - WHOLE NETWORKS through the complete device program: cnn.cl as C with all 25 kernels running
 *     concurrently as coroutines (oracle/ref_device/net_harness.c) — every layer of ResNet50, GoogLeNet
 *     and pruned ResNet50 chained through the on-chip feature cache (the retriever's non-blocking reads,
 *     retriever.cl:328-329), the ipool feed (retriever.cl:285-302), the DDR residual ping-pong and concat
 *     offsets; 0 mismatches on every layer (tests/test_whole_net_ref.py, tests/golden/whole_net_golden.json).
 *
 *   - networks the reference ships no tables for (VGG16, SqueezeNet fire modules, pool probes) through the
 *     same device program built against a generated table header (oracle/ref_device/multi_layer.py;
 *     tests/test_generated_nets_ref.py, tests/golden/generated_nets_golden.json).
 *
 * Layouts follow the reference host side: features [C][H][W] int8, codes [N][C][FH][FW] uint8.
 */
#ifndef TF2_ORACLE_H
#define TF2_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Runtime_Engine/cnn/host/inc/types.h:39-43 */
typedef struct {
  int32_t bias;
  int32_t alpha;
  int32_t beta;
} tf2o_bias_bn;

/* One fused layer, SURVEY.md Appendix A.  All fields are plain ints so ctypes can fill it. */
typedef struct {
  int32_t C, IH, IW;        /* input feature map (already transformed for layer 0) */
  int32_t N;                /* output channels of this layer */
  int32_t k;                /* kFilterSize */
  int32_t pad;              /* kPadHeight == kPadWidth */
  int32_t stride;           /* kConvStride */
  int32_t OH, OW;           /* conv output size AFTER the stride */
  int32_t relu;             /* kReluEnable */
  int32_t pool;             /* kPoolEnable (3x3 max, zero outside) */
  int32_t pool_stride;      /* 2 if kPoolStride2 else 1 */
  int32_t pool_pad;         /* kPoolPad */
  int32_t PH, PW;           /* size after pool (== OH,OW when !pool) */
  int32_t add;              /* kAdditionEnable */
  int32_t add_relu;         /* kAdditionReluEnable */
  int32_t gap;              /* kEndPoolEnable: 7x7 global average */
  int32_t ipool;            /* kIpoolEnable: no conv, 3x3/s1/p1 max pool of the input */
} tf2o_layer;

/* pe.cl:27-40 */
int32_t tf2o_mul(int8_t feature, uint8_t code);
/* model_loader.cpp:98-126 */
uint8_t tf2o_get_real(float w, int8_t expand);
/* model_loader.cpp:25-96: one 7x7 code plane -> 9 derived 3x3 planes (81 bytes, caller pre-fills) */
void tf2o_filter_trans(const uint8_t* in49, uint8_t* out81);
/* input_loader.cpp:27-73: one 224x224 float plane -> 9 planes of 115x115 floats */
void tf2o_feature_trans(const float* in, float* out9x115x115);
/* runner.cpp:158-164 */
int8_t tf2o_quantize_input(float x, int q0);
/* pe.cl:185-203 */
int8_t tf2o_requant(int32_t acc, int32_t alpha, int32_t beta);
/* full_size_pool.cl:95-119 */
int8_t tf2o_gap_finish(int32_t sum);

/* Steps 1 of Appendix A: acc[N][OH][OW] (bias seeded).  pe.cl:144-180, sequencer.cl:268-311 */
void tf2o_conv_acc(const tf2o_layer* L, const int8_t* X, const uint8_t* code,
                   const tf2o_bias_bn* P, int32_t* acc);

/* Steps 1-7 for one image.  `out` has layout [N][PH][PW] (or [N] when gap).  R is the residual
 * operand [N][PH][PW] or NULL.  acc_out (nullable) receives the int32 accumulators [N][OH][OW]. */
void tf2o_layer_forward(const tf2o_layer* L, const int8_t* X, const uint8_t* code,
                        const tf2o_bias_bn* P, const int8_t* R, int8_t* out, int32_t* acc_out);

/* Whole-network executor over `n_images` images (OpenMP across images and channels).
 * in_idx[l]  : tensor id read by layer l (0 = network input, t>0 = tensor t)
 * out_idx[l] : tensor id written by layer l, out_ch0[l] the channel offset inside it (concat)
 * add_idx[l] : tensor id of the residual operand (or -1)
 * tensor_C/H/W[t]: geometry of tensor t.  Tensor 0 is the (transformed, quantised) input.
 * code_off / param_off: offsets of each layer's codes / params in the packed arrays.
 * result: tensor `result_tensor` of every image is copied to `out` ([n_images][C*H*W]). */
int tf2o_run_network(int n_layers, const tf2o_layer* layers, const int32_t* in_idx,
                     const int32_t* out_idx, const int32_t* out_ch0, const int32_t* add_idx,
                     int n_tensors, const int32_t* tensor_C, const int32_t* tensor_H,
                     const int32_t* tensor_W, const uint8_t* codes, const int64_t* code_off,
                     const tf2o_bias_bn* params, const int64_t* param_off, const int8_t* input,
                     int n_images, int result_tensor, int8_t* out, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
