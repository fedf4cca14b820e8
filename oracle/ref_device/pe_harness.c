/*
 * pe_harness.c — runs the REFERENCE's own PE kernel source (Runtime_Engine/cnn/device/src/pe.cl,
 * compiled as plain C where it lies; nothing is copied) on caller-supplied control / filter / data
 * items.  Test infrastructure only: it pins steps 1-2 of SURVEY.md Appendix A (MUL, DotProduct, the
 * bias seed, int32 wrap-around accumulation, requantisation and clamp of PeFunction) for arbitrary
 * reduction lengths.  Built by oracle/build_ref.sh into oracle/_ref/libtf2ref_pe.so.
 *
 * OpenCL-isms are mapped by fifo_shim.h (channels -> unbounded FIFOs).
 */
#include "fifo_shim.h"

/* the reference source, unmodified */
#include "pe.cl"

/* ---- C entry points for ctypes ---- */
int pe_w_vector(void) { return W_VECTOR; }
int pe_c_vector(void) { return C_VECTOR; }
int pe_ow_vector(void) { return OW_VECTOR; }
int pe_fw_vector(void) { return FW_VECTOR; }

/* reference MUL / DotProduct as compiled from pe.cl */
int pe_ref_mul(signed char feature, signed char code) { return MUL(feature, code); }

/* 1x1 mode (pe.cl:160-171): for each of the W_VECTOR=7 columns an independent reduction over
 * `steps` x C_VECTOR=16 channels.  x[steps][7][16] int8, codes[steps][16], one BiasBnParam.
 * Returns the 7 requantised int8 outputs of PE 0 in out[7].  Filters are first written into the
 * PE's cache through the same channel protocol the retriever uses (pe.cl:116-127). */
int pe_run_1x1(int steps, const signed char* x, const unsigned char* codes, int bias, int alpha, int beta,
               signed char* out) {
  if (steps < 1 || steps > FILTER_CACHE_PAGE_DEPTH) return -1;
  fifo_reset_all();
  PeControlSignal cont;
  PeInputFilter filt;
  PeInputData din;
  /* phase 1: load `steps` filter vectors into cache page 0 (write addr = step), no valid data */
  for (int s = 0; s < steps; s++) {
    memset(&cont, 0, sizeof cont);
    memset(&filt, 0, sizeof filt);
    memset(&din, 0, sizeof din);
    cont.filter_write_addr = s;
    cont.filter_bias_read_page = 1;  /* write page = !read page = 0 */
    filt.data_valid = true;
    filt.n_inc = 0;
    for (int c = 0; c < C_VECTOR; c++) filt.filter_data.v[0].v[c] = (real)codes[s * C_VECTOR + c];
    filt.bias_bn_data.bias = bias;
    filt.bias_bn_data.alpha = alpha;
    filt.bias_bn_data.beta = beta;
    write_channel_altera(pe_control_channel_first, cont);
    write_channel_altera(pe_input_filter_channel_first, filt);
    write_channel_altera(pe_input_data_channel_first, din);
  }
  /* phase 2: the reduction; conv_start on the first step, conv_done on the last */
  for (int s = 0; s < steps; s++) {
    memset(&cont, 0, sizeof cont);
    memset(&filt, 0, sizeof filt);
    memset(&din, 0, sizeof din);
    cont.is_QVECTOR = false;
    cont.conv_start = (s == 0);
    cont.conv_done[0] = (s == steps - 1);
    cont.filter_read_addr = s;
    cont.filter_read_fw_vec = 0;
    cont.filter_bias_read_page = 0;
    din.input_data_valid = true;
    for (int w = 0; w < W_VECTOR; w++)
      for (int c = 0; c < C_VECTOR; c++) din.input_data.v[w].v[c] = x[(s * W_VECTOR + w) * C_VECTOR + c];
    write_channel_altera(pe_control_channel_first, cont);
    write_channel_altera(pe_input_filter_channel_first, filt);
    write_channel_altera(pe_input_data_channel_first, din);
  }
  if (!setjmp(g_exit)) PeFunction(0);
  if (fifo_count(&pe_output_channel[0], sizeof(PeOutput)) != 1) return -2;
  PeOutput o = read_channel_altera(pe_output_channel[0]);
  for (int w = 0; w < W_VECTOR; w++) out[w] = o.data.v[w];
  return 0;
}

/* 3x3 mode (pe.cl:146-159): OW_VECTOR=5 outputs, out[ow] += sum_fw Dot(in[ow+fw], w[fw]) per step;
 * x[steps][7][16], codes[steps][3][16]. */
int pe_run_3x3(int steps, const signed char* x, const unsigned char* codes, int bias, int alpha, int beta,
               signed char* out) {
  if (steps < 1 || steps > FILTER_CACHE_PAGE_DEPTH) return -1;
  fifo_reset_all();
  PeControlSignal cont;
  PeInputFilter filt;
  PeInputData din;
  for (int s = 0; s < steps; s++) {
    memset(&cont, 0, sizeof cont);
    memset(&filt, 0, sizeof filt);
    memset(&din, 0, sizeof din);
    cont.filter_write_addr = s;
    cont.filter_bias_read_page = 1;
    filt.data_valid = true;
    filt.n_inc = 0;
    for (int fw = 0; fw < FW_VECTOR; fw++)
      for (int c = 0; c < C_VECTOR; c++) filt.filter_data.v[fw].v[c] = (real)codes[(s * FW_VECTOR + fw) * C_VECTOR + c];
    filt.bias_bn_data.bias = bias;
    filt.bias_bn_data.alpha = alpha;
    filt.bias_bn_data.beta = beta;
    write_channel_altera(pe_control_channel_first, cont);
    write_channel_altera(pe_input_filter_channel_first, filt);
    write_channel_altera(pe_input_data_channel_first, din);
  }
  for (int s = 0; s < steps; s++) {
    memset(&cont, 0, sizeof cont);
    memset(&filt, 0, sizeof filt);
    memset(&din, 0, sizeof din);
    cont.is_QVECTOR = true;
    cont.conv_start = (s == 0);
    cont.conv_done[0] = (s == steps - 1);
    cont.filter_read_addr = s;
    cont.filter_bias_read_page = 0;
    din.input_data_valid = true;
    for (int w = 0; w < W_VECTOR; w++)
      for (int c = 0; c < C_VECTOR; c++) din.input_data.v[w].v[c] = x[(s * W_VECTOR + w) * C_VECTOR + c];
    write_channel_altera(pe_control_channel_first, cont);
    write_channel_altera(pe_input_filter_channel_first, filt);
    write_channel_altera(pe_input_data_channel_first, din);
  }
  if (!setjmp(g_exit)) PeFunction(0);
  if (fifo_count(&pe_output_channel[0], sizeof(PeOutput)) != 1) return -2;
  PeOutput o = read_channel_altera(pe_output_channel[0]);
  for (int w = 0; w < W_VECTOR; w++) out[w] = o.data.v[w];
  return 0;
}
int pe_filter_cache_page_depth(void) { return FILTER_CACHE_PAGE_DEPTH; }
