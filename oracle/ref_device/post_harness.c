/*
 * post_harness.c — runs the REFERENCE's own post-PE kernels (Runtime_Engine/cnn/device/src/
 * relu.cl, pool.cl, pool_tail.cl, feature_writer.cl, full_size_pool.cl — compiled as plain C where
 * they lie; nothing is copied) over a whole network's layer tables for one frame.
 *
 * Test infrastructure only.  It pins steps 3-6 of SURVEY.md Appendix A against executed reference
 * code: ReLU (relu.cl:54), the separable 3x3 max pool with zero border (pool.cl:178-260), the pool
 * alignment / stride-2 column and row selection (pool_tail.cl:91-216, which also realises the conv
 * stride in W), the residual add + clamp + ReLU through the DDR ping-pong (feature_writer.cl:88-147),
 * the concat channel offset (feature_writer.cl:109) and the 7x7 global average (full_size_pool.cl:
 * 95-125).  The caller supplies, per layer, the int8 map the PEs would emit (pe.cl output, pinned
 * separately by pe_harness.c): [N][H][W1] with H = ceil(kOutputHeight / kConvStride) rows (the
 * sequencer applies the stride to rows, sequencer.cl:116) and W1 = kOutputWidth columns at stride 1.
 * For ipool pseudo layers the same array feeds `ipool_channel` (what retriever.cl:285-302 sends).
 *
 * Built by oracle/build_ref.sh with -DRESNET50 / -DGOOGLENET into oracle/_ref/libtf2ref_post_<net>.so.
 * OpenCL-isms are mapped by fifo_shim.h (channels -> unbounded FIFOs).
 */
#include "fifo_shim.h"

/* the reference sources, unmodified (each includes ../../host/inc/cnn.h = tables of the chosen net) */
#include "cycle.cl"
#include "relu.cl"
#include "pool.cl"
#include "pool_tail.cl"
#include "feature_writer.cl"
#include "full_size_pool.cl"

int post_num_layers(void) { return NUM_CONVOLUTIONS; }
int post_w_vector(void) { return W_VECTOR; }
int post_n_vector(void) { return NARROW_N_VECTOR; }
long long post_output_offset(void) { return (long long)OUTPUT_OFFSET; }
long long post_ddr_bytes(void) { return (long long)OUTPUT_OFFSET * 3 + (1 << 20); }
int post_item_bytes(void) { return (int)sizeof(PoolTailOutput); }
int post_item_data_offset(void) { return (int)((char*)&((PoolTailOutput*)0)->write_data - (char*)0); }
int post_item_row_stride(void) { return NEXT_POWER_OF_2(C_VECTOR); }

/* table getter: the values the kernels themselves read */
int post_table(int which, int l) {
  switch (which) {
    case 0: return kOutputChannels[l];
    case 1: return kOutputHeight[l];
    case 2: return kOutputWidth[l];
    case 3: return kConvStride[l];
    case 4: return kFilterSize[l];
    case 5: return kReluEnable[l];
    case 6: return kPoolEnable[l];
    case 7: return kPoolStride2[l];
    case 8: return kPoolPad[l];
    case 9: return kPoolOutputHeight[l];
    case 10: return kPoolOutputWidth[l];
    case 11: return kAdditionEnable[l];
    case 12: return kAdditionReluEnable[l];
    case 13: return kEndPoolEnable[l];
    case 14: return kIpoolEnable[l];
    case 15: return kNStart[l];
    case 16: return kNEnd[l];
    case 17: return kDDRWriteEnable[l];
    case 18: return kDDRWriteBase[l];
    case 19: return kDDRReadBase[l];
    case 20: return kCacheWriteEnable[l];
    case 21: return kCacheWriteBase[l];
    case 22: return kNvecEnd[l];
    case 23: return kPoolOutputWvecEnd[l];
    case 24: return kNEndWithOffset[l];
    case 25: return kOhEndWithOffset[l];
    case 26: return kOwEndWithOffset[l];
    case 27: return kPoolWindow[l];
    default: return -1;
  }
}

/* Feeds the PE-output / ipool streams in the order pool.cl consumes them (pool.cl:85-119: n by
 * N_VECTOR, oh, ow by OW_VECTOR (3x3 mode) or W_VECTOR (1x1 mode); an item exists where
 * kNStart+n < kNEnd, oh < H, ow < W), then runs relu -> pool -> pool_tail -> feature_writer ->
 * full_size_pool to completion for ONE frame.
 *   n_feed_layers : only layers [0, n_feed_layers) are fed; the kernels stop when their input runs
 *                   dry, so feature_ddr then shows the state right after that layer (the DDR pages
 *                   are ping-pong buffers that later layers overwrite)
 *   y, y_off : concatenated per-layer maps [N][H][W1] and their byte offsets
 *   ddr      : feature_ddr image (caller-zeroed, post_ddr_bytes() long)
 *   cache_items / n_cache : raw PoolTailOutput items feature_writer sent to the retriever
 *   gap_items / n_gap     : raw PoolTailOutput items full_size_pool emitted
 * returns 0, or a negative code when an item count disagrees with the reference's cycle constants */
int post_run(int n_feed_layers, const signed char* y, const long long* y_off, signed char* ddr, unsigned char* cache_items,
             long long cache_cap, long long* n_cache, unsigned char* gap_items, long long gap_cap,
             long long* n_gap, long long* counts /* [4]: relu in, pool out, fw in, fw out */) {
  fifo_reset_all();
  long long n_relu_in = 0;
  for (int l = 0; l < NUM_CONVOLUTIONS && l < n_feed_layers; l++) {
    const int N = kNEndWithOffset[l], OH = kOhEndWithOffset[l], OW = kOwEndWithOffset[l];
    const int H = CEIL(kOutputHeight[l], kConvStride[l]), W = kOutputWidth[l];
    const int FH = kFilterSize[l];
    const int WOW = FH != 1 ? OW_VECTOR : W_VECTOR;
    const int nch = kNEnd[l] - kNStart[l];
    const signed char* yl = y + y_off[l];
    for (int n = 0; n < N; n += N_VECTOR)
      for (int oh = 0; oh < OH; oh++)
        for (int ow = 0; ow < OW; ow += WOW) {
          if (!((kNStart[l] + n) < kNEnd[l] && oh < H && ow < W)) continue;
          if (kIpoolEnable[l]) {
            ReluOutput r;
            memset(&r, 0, sizeof r);
            for (int ni = 0; ni < NARROW_N_VECTOR; ni++)
              for (int wi = 0; wi < W_VECTOR; wi++)
                if (n + ni < nch && ow + wi < W) r.data[ni].v[wi] = yl[((long long)(n + ni) * H + oh) * W + ow + wi];
            write_channel_altera(ipool_channel, r);
          } else {
            for (int ni = 0; ni < NARROW_N_VECTOR; ni++) {
              PeOutput o;
              memset(&o, 0, sizeof o);
              o.is_QVECTOR = FH != 1;
              o.pe_output_relu = kReluEnable[l];
              for (int wi = 0; wi < W_VECTOR; wi++)
                if (n + ni < nch && ow + wi < W) o.data.v[wi] = yl[((long long)(n + ni) * H + oh) * W + ow + wi];
              write_channel_altera(pe_output_channel[ni], o);
            }
            n_relu_in++;
          }
        }
  }
  counts[0] = n_relu_in;
  if (!setjmp(g_exit)) relu(1);
  if (!setjmp(g_exit)) pool(1);
  counts[1] = (long long)fifo_count(&pool_output_channel, sizeof(PoolOutput));
  if (!setjmp(g_exit)) pool_tail(1, (real*)ddr);
  counts[2] = (long long)fifo_count(&feature_writer_input_channel, sizeof(PoolTailOutput));
  if (!setjmp(g_exit)) feature_writer(1, (real*)ddr);
  counts[3] = (long long)fifo_count(&feature_writer_input_channel, sizeof(PoolTailOutput));
  if (!setjmp(g_exit)) full_size_pool(1);
  long long nc = (long long)fifo_count(&retriever_input_channel, sizeof(PoolTailOutput));
  long long ng = (long long)fifo_count(&end_pool_output_channel, sizeof(PoolTailOutput));
  if (nc > cache_cap || ng > gap_cap) return -3;
  for (long long i = 0; i < nc; i++) {
    PoolTailOutput o = read_channel_altera(retriever_input_channel);
    memcpy(cache_items + i * sizeof o, &o, sizeof o);
  }
  for (long long i = 0; i < ng; i++) {
    PoolTailOutput o = read_channel_altera(end_pool_output_channel);
    memcpy(gap_items + i * sizeof o, &o, sizeof o);
  }
  *n_cache = nc;
  *n_gap = ng;
  return 0;
}

/* the reference's own totals (resnet50.h:1613-1616 or the cycle.cl functions) */
long long post_const(int which) {
  switch (which) {
    case 0: return CONV_TOTAL_WRITE_CACHE;
    case 1: return POOL_TOTAL_CYCLE;
    case 2: return FEATURE_WRITER_TOTAL_CYCLE;
    case 3: return END_POOL_TOTAL_CYCLE;
    default: return -1;
  }
}
