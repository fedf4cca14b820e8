/*
 * full_harness.c — runs the REFERENCE's whole device pipeline (Runtime_Engine/cnn/device/src/cnn.cl:
 * input_reader, filter_reader, sequencer, retriever, the 16 PE kernels, relu, pool, pool_tail,
 * feature_writer — compiled as plain C where they lie; nothing is copied) for LAYER 0 of the
 * network the tables describe, from the same device buffers the host would upload.
 *
 * Test infrastructure only.  It pins step 1 of SURVEY.md Appendix A — the convolution geometry of
 * sequencer.cl:268-311 / retriever.cl:134-213 (which feature feeds which tap, zero fill beyond the
 * map and beyond the channel count, filter cache addressing) chained to the real PE arithmetic and
 * the real post-PE kernels — for the first layer of each shipped network (3x3 on the 27-channel
 * 114x114 input, 64 outputs, ReLU, 3x3/s2 max pool).  Layers > 0 read the on-chip cache that
 * feature_writer fills through a non-blocking channel (retriever.cl:328-329), a feedback loop that
 * needs cycle-accurate co-scheduling, so they are not run here.
 *
 * Built by oracle/build_ref.sh with -DRESNET50 / -DGOOGLENET into oracle/_ref/libtf2ref_full_<net>.so.
 */
#include "fifo_shim.h"

#include "cnn.cl"

long long full_const(int which) {
  switch (which) {
    case 0: return CONV_CYCLE(0);
    case 1: return FILTER_PRELOAD_CYCLE;
    case 2: return INPUT_READER_CYCLE;
    case 3: return FEATURE_WRITER_CYCLE(0);
    case 4: return POOL_CYCLE(0);
    default: return -1;
  }
}
int full_item_bytes(void) { return (int)sizeof(PoolTailOutput); }
int full_item_data_offset(void) { return (int)((char*)&((PoolTailOutput*)0)->write_data - (char*)0); }

/* input_buffer : int8 image in the InputConvert layout; gl_filter : FilterConvert output (whole net
 * sized, only layer 0 matters); bias_bn : BiasBnParam[NUM_CONVOLUTIONS * MAX_BIAS_SIZE];
 * seq_items : how many sequencer items to let through (the first layer's schedule).
 * cache_items : raw PoolTailOutput items feature_writer sent to the retriever (layer 0's output).
 * counts[8] : items seen on the way (input reader, sequencer, PE outputs of PE 0, relu, pool, pool_tail, writer) */
int full_run_layer0(const signed char* input_buffer, signed char* gl_filter, BiasBnParam* bias_bn, long long seq_items,
                    signed char* ddr, unsigned char* cache_items, long long cache_cap, long long* n_cache, long long* counts) {
  static int idle[NUM_CONVOLUTIONS];
  fifo_reset_all();
  g_limit_key = NULL;
  if (!setjmp(g_exit)) input_reader(1, (const real*)input_buffer);
  counts[0] = (long long)fifo_count(&input_reader_output_channel, sizeof(InputReaderOutput));
  /* the filter stream of layer 0 (+ the preload of layer 1 the reader interleaves): cap it generously */
  g_limit_key = &filter_reader_output_channel;
  g_limit_bytes = (size_t)400000 * sizeof(FilterReaderOutput);
  if (!setjmp(g_exit)) filter_reader(1, (real*)gl_filter, bias_bn);
  g_limit_key = &sequencer_output_channel;
  g_limit_bytes = (size_t)seq_items * sizeof(SequencerOutput);
  if (!setjmp(g_exit)) sequencer(1);
  g_limit_key = NULL;
  counts[1] = (long long)fifo_count(&sequencer_output_channel, sizeof(SequencerOutput));
  if (!setjmp(g_exit)) retriever(1, idle);
  for (int n = 0; n < N_VECTOR; n++)
    if (!setjmp(g_exit)) PeFunction(n);
  counts[2] = (long long)fifo_count(&pe_output_channel[0], sizeof(PeOutput));
  if (!setjmp(g_exit)) relu(1);
  counts[3] = (long long)fifo_count(&relu_output_channel, sizeof(ReluOutput));
  if (!setjmp(g_exit)) pool(1);
  counts[4] = (long long)fifo_count(&pool_output_channel, sizeof(PoolOutput));
  if (!setjmp(g_exit)) pool_tail(1, (real*)ddr);
  counts[5] = (long long)fifo_count(&feature_writer_input_channel, sizeof(PoolTailOutput));
  if (!setjmp(g_exit)) feature_writer(1, (real*)ddr);
  long long nc = (long long)fifo_count(&retriever_input_channel, sizeof(PoolTailOutput));
  if (nc > cache_cap) return -3;
  for (long long i = 0; i < nc; i++) {
    PoolTailOutput o = read_channel_altera(retriever_input_channel);
    memcpy(cache_items + i * sizeof o, &o, sizeof o);
  }
  *n_cache = nc;
  return 0;
}
