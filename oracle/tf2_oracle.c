/*
 * tf2_oracle.c — see tf2_oracle.h.  TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C restatement of the TF2 Runtime_Engine/cnn device arithmetic and
 * of the host-side numeric preparation, each function citing the reference lines it follows.
 * Integer conventions: two's complement, arithmetic >> on signed values, all int32 sums wrap
 * (computed in uint32_t to stay free of C undefined behaviour).
 */
#include "tf2_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* pe.cl:27-40.  bit6 => 0; bit7 => negate the feature in the 8-bit `real` type (so -(-128) stays
 * -128, pe.cl:32-34); shift by the low 5 bits, result wraps in 32 bits. */
int32_t tf2o_mul(int8_t feature, uint8_t code) {
  if (code & 0x40) return 0;
  int8_t f = feature;
  if (code & 0x80) f = (int8_t)(uint8_t)(0u - (uint8_t)feature);
  return (int32_t)((uint32_t)(int32_t)f << (code & 0x1f));
}

/* model_loader.cpp:98-126 */
uint8_t tf2o_get_real(float data, int8_t expand) {
  int sign = 0;
  int vals = 0;
  if (fabs(data) < 1.0e-05) return 0x40;
  if (data < 0) {
    sign = 1;
    data = -data;
  }
  for (int i = 0; i < 15; i++) {
    float temps = 1.0f / (float)(1 << i);
    /* the reference compares float against double products (0.99 * temps is double) */
    if ((double)data > 0.99 * (double)temps && (double)data < 1.01 * (double)temps) {
      vals = i;
      break;
    }
  }
  int8_t oups = (int8_t)(expand - vals);
  if (oups < 0) oups = 0;
  uint8_t r = (uint8_t)oups;
  if (sign) r |= 0x80;
  return r;
}

/* model_loader.cpp:25-96.  A 7x7/stride-2 filter plane becomes nine 3x3 planes acting on the
 * nine derived input planes of feature_trans.  Derived plane index d:
 *   d = 2*wsel + hpar (0..5): wsel 0/1 = even/odd input columns, wsel 2 = even columns shifted by
 *       one (carries filter column 6); hpar = row parity;   taps [r][j] = f[2r+hpar][col(wsel,j)]
 *   d = 6 + wsel: filter row 6 (the 4th even row), placed in tap row 2.
 * Positions the reference never writes keep the caller's fill (LoadModel memsets 0, which is the
 * code for +1: model_loader.cpp:247). */
void tf2o_filter_trans(const uint8_t* f, uint8_t* out) {
  uint8_t wsplit[3][7][3]; /* [wsel][row][j] */
  for (int w = 0; w < 3; w++)
    for (int r = 0; r < 7; r++)
      for (int j = 0; j < 3; j++) wsplit[w][r][j] = 0x40;
  for (int r = 0; r < 7; r++) {
    for (int j = 0; j < 3; j++) {
      wsplit[0][r][j] = f[r * 7 + 2 * j];     /* even columns 0,2,4 */
      wsplit[1][r][j] = f[r * 7 + 2 * j + 1]; /* odd columns 1,3,5  */
    }
    wsplit[2][r][2] = f[r * 7 + 6];           /* column 6 rides on the shifted even plane */
  }
  for (int w = 0; w < 3; w++) {
    for (int hp = 0; hp < 2; hp++) {
      int d = 2 * w + hp;
      for (int r = 0; r < 3; r++)
        for (int j = 0; j < 3; j++) out[d * 9 + r * 3 + j] = wsplit[w][2 * r + hp][j];
    }
    /* row 6 = fourth even row -> derived planes 6..8, tap row 2 (rows 0-1 never written) */
    for (int j = 0; j < 3; j++) out[(6 + w) * 9 + 2 * 3 + j] = wsplit[w][6][j];
  }
}

/* input_loader.cpp:27-73.  224x224 plane, zero pad 3 -> 230x230, split by column parity
 * (+ shifted even plane) and row parity into 6 planes of 115x115, then 3 more planes that are the
 * even-row planes shifted down by one row. */
void tf2o_feature_trans(const float* in, float* out) {
  enum { D = 224, P = 3, ND = 230, HD = 115 };
  float* pad = (float*)calloc((size_t)ND * ND, sizeof(float));
  float* med = (float*)calloc((size_t)3 * ND * ND, sizeof(float));
  float* ht = (float*)calloc((size_t)6 * HD * HD, sizeof(float));
  for (int i = 0; i < D; i++)
    for (int j = 0; j < D; j++) pad[(i + P) * ND + j + P] = in[i * D + j];
  for (int i = 0; i < ND; i++)
    for (int j = 0; j < ND; j++) med[((j % 2) * ND + i) * ND + j / 2] = pad[i * ND + j];
  for (int i = 0; i < ND; i++)
    for (int j = 0; j < ND - 1; j++) med[(2 * ND + i) * ND + j] = med[(0 * ND + i) * ND + j + 1];
  for (int w = 0; w < 3; w++)
    for (int j = 0; j < ND; j++)
      for (int k = 0; k < HD; k++)
        ht[((w * 2 + j % 2) * HD + j / 2) * HD + k] = med[(w * ND + j) * ND + k];
  for (int d = 0; d < 6; d++)
    for (int j = 0; j < HD; j++)
      for (int k = 0; k < HD; k++) out[(d * HD + j) * HD + k] = ht[(d * HD + j) * HD + k];
  for (int w = 0; w < 3; w++)
    for (int j = 0; j < HD; j++)
      for (int k = 0; k < HD; k++)
        out[((w + 6) * HD + j) * HD + k] = (j + 1 < HD) ? ht[((w * 2) * HD + j + 1) * HD + k] : 0.0f;
  free(pad);
  free(med);
  free(ht);
}

/* runner.cpp:158-164: x * 2^Q0 (q0 = -Q0), round half away from zero, clamp to int8 */
int8_t tf2o_quantize_input(float x, int q0) {
  float trans = q0 > 0 ? (1.0f / (float)(1 << q0)) : (float)(1 << (-q0));
  float tmp = x * trans;
  int t = (int)(tmp > 0 ? tmp + 0.5 : tmp - 0.5);
  return (int8_t)(t > 127 ? 127 : t < -128 ? -128 : t);
}

/* pe.cl:185-203: int64 product, >>20 truncated to int32, +beta wraps, >>14, +1, >>1, clamp */
int8_t tf2o_requant(int32_t acc, int32_t alpha, int32_t beta) {
  int64_t t = (int64_t)acc * (int64_t)alpha;
  int32_t a = (int32_t)(t >> 20);
  int32_t s = (int32_t)((uint32_t)a + (uint32_t)beta);
  int32_t y = ((s >> 14) + 1) >> 1;
  return (int8_t)(y > 127 ? 127 : y < -128 ? -128 : y);
}

/* full_size_pool.cl:95-119: ((sum*669 >> 14) + 1) >> 1, clamp (669 ~ 2^15/49) */
int8_t tf2o_gap_finish(int32_t sum) {
  int32_t y = (((sum * 669) >> 14) + 1) >> 1;
  return (int8_t)(y > 127 ? 127 : y < -128 ? -128 : y);
}

/* pe.cl:144-180 with the geometry of sequencer.cl:268-311 / retriever.cl:134-213 (zero padding,
 * h = oh*stride - pad + fh; the FPGA's stride-1-in-W + column drop is the same strided conv). */
static void conv_acc_channel(const tf2o_layer* L, const int8_t* X, const uint8_t* code_n,
                             int32_t bias, int32_t* acc /* [OH][OW] */) {
  const int C = L->C, IH = L->IH, IW = L->IW, k = L->k, pad = L->pad, s = L->stride;
  const int OH = L->OH, OW = L->OW;
  uint32_t* a = (uint32_t*)acc;
  for (int i = 0; i < OH * OW; i++) a[i] = (uint32_t)bias;
  if (k == 1 && s == 1 && pad == 0) {
    /* 1x1: the map is one contiguous run of OH*OW pixels per channel (same arithmetic, longer vector loops) */
    const int HW = OH * OW;
    for (int c = 0; c < C; c++) {
      const uint8_t cd = code_n[c];
      if (cd & 0x40) continue;
      const int sh = cd & 0x1f;
      const int8_t* xr = X + (size_t)c * HW;
      if (cd & 0x80) {
        for (int i = 0; i < HW; i++) {
          int8_t f = (int8_t)(uint8_t)(0u - (uint8_t)xr[i]); /* pe.cl:32-34 */
          a[i] += (uint32_t)(int32_t)f << sh;
        }
      } else {
        for (int i = 0; i < HW; i++) a[i] += (uint32_t)(int32_t)xr[i] << sh;
      }
    }
    return;
  }
  for (int c = 0; c < C; c++) {
    const int8_t* Xc = X + (size_t)c * IH * IW;
    for (int fh = 0; fh < k; fh++) {
      for (int fw = 0; fw < k; fw++) {
        uint8_t cd = code_n[(c * k + fh) * k + fw];
        if (cd & 0x40) continue;
        const int sh = cd & 0x1f;
        const int neg = cd & 0x80;
        /* valid output range so that 0 <= oh*s - pad + fh < IH (same for w) */
        int oh0 = 0, ow0 = 0;
        while (oh0 < OH && oh0 * s - pad + fh < 0) oh0++;
        while (ow0 < OW && ow0 * s - pad + fw < 0) ow0++;
        int oh1 = OH, ow1 = OW;
        while (oh1 > oh0 && (oh1 - 1) * s - pad + fh >= IH) oh1--;
        while (ow1 > ow0 && (ow1 - 1) * s - pad + fw >= IW) ow1--;
        for (int oh = oh0; oh < oh1; oh++) {
          const int8_t* xr = Xc + (size_t)(oh * s - pad + fh) * IW + (-pad + fw);
          uint32_t* ar = a + (size_t)oh * OW;
          if (neg) {
            for (int ow = ow0; ow < ow1; ow++) {
              int8_t f = (int8_t)(uint8_t)(0u - (uint8_t)xr[ow * s]); /* pe.cl:32-34 */
              ar[ow] += (uint32_t)(int32_t)f << sh;
            }
          } else {
            for (int ow = ow0; ow < ow1; ow++) ar[ow] += (uint32_t)(int32_t)xr[ow * s] << sh;
          }
        }
      }
    }
  }
}

void tf2o_conv_acc(const tf2o_layer* L, const int8_t* X, const uint8_t* code,
                   const tf2o_bias_bn* P, int32_t* acc) {
  const size_t kk = (size_t)L->C * L->k * L->k;
  for (int n = 0; n < L->N; n++)
    conv_acc_channel(L, X, code + n * kk, P[n].bias, acc + (size_t)n * L->OH * L->OW);
}

/* pool.cl:178-260 + pool_tail.cl:91-216: separable 3x3 max; window of output j starts at
 * j*ps - ppad; taps outside the map contribute 0 (not -inf). */
static void pool3x3(const int8_t* in, int H, int W, int ps, int ppad, int PH, int PW, int8_t* out) {
  for (int ph = 0; ph < PH; ph++) {
    for (int pw = 0; pw < PW; pw++) {
      int m = -128;
      for (int dh = 0; dh < 3; dh++) {
        for (int dw = 0; dw < 3; dw++) {
          int h = ph * ps - ppad + dh, w = pw * ps - ppad + dw;
          int v = (h >= 0 && h < H && w >= 0 && w < W) ? in[h * W + w] : 0;
          if (v > m) m = v;
        }
      }
      out[ph * PW + pw] = (int8_t)m;
    }
  }
}

/* one output channel of one layer: steps 1-6 of Appendix A; scratch sized by caller */
static void layer_channel(const tf2o_layer* L, const int8_t* X, const uint8_t* code_n,
                          const tf2o_bias_bn* Pn, const int8_t* Rn, int8_t* out_n,
                          int32_t* acc_n /* [OH][OW] scratch or user buffer */,
                          int8_t* y /* [OH][OW] scratch */, int8_t* yp /* [PH][PW] scratch */) {
  const int OH = L->OH, OW = L->OW, PH = L->PH, PW = L->PW;
  const int8_t* cur;
  if (L->ipool) {
    /* retriever.cl:285-302: the pseudo layer pools channel n of its input, 3x3 / s1 / p1 */
    pool3x3(X, L->IH, L->IW, 1, 1, PH, PW, yp);
    cur = yp;
  } else {
    conv_acc_channel(L, X, code_n, Pn->bias, acc_n);
    for (int i = 0; i < OH * OW; i++) {
      int8_t v = tf2o_requant(acc_n[i], Pn->alpha, Pn->beta);
      if (L->relu && v < 0) v = 0; /* relu.cl:54 */
      y[i] = v;
    }
    cur = y;
    if (L->pool) {
      pool3x3(y, OH, OW, L->pool_stride, L->pool_pad, PH, PW, yp);
      cur = yp;
    }
  }
  if (L->gap) {
    /* feature_writer.cl:124-127 then full_size_pool.cl:95-119 */
    int32_t sum = 0;
    for (int i = 0; i < PH * PW; i++) {
      int v = cur[i];
      if (L->add) {
        int t = v + (int)Rn[i];
        t = t > 127 ? 127 : t < -128 ? -128 : t;
        if (L->add_relu && t < 0) t = 0;
        v = t;
      }
      sum = (int16_t)(sum + v);
    }
    out_n[0] = tf2o_gap_finish(sum);
    return;
  }
  for (int i = 0; i < PH * PW; i++) {
    int v = cur[i];
    if (L->add) { /* feature_writer.cl:124-127 */
      int t = v + (int)Rn[i];
      t = t > 127 ? 127 : t < -128 ? -128 : t;
      if (L->add_relu && t < 0) t = 0;
      v = t;
    }
    out_n[i] = (int8_t)v;
  }
}

void tf2o_layer_forward(const tf2o_layer* L, const int8_t* X, const uint8_t* code,
                        const tf2o_bias_bn* P, const int8_t* R, int8_t* out, int32_t* acc_out) {
  const size_t kk = (size_t)L->C * L->k * L->k;
  const size_t ohw = (size_t)L->OH * L->OW, phw = (size_t)L->PH * L->PW;
  const size_t osz = L->gap ? 1 : phw;
#pragma omp parallel
  {
    int32_t* acc = (int32_t*)malloc(sizeof(int32_t) * (ohw ? ohw : 1));
    int8_t* y = (int8_t*)malloc(ohw ? ohw : 1);
    int8_t* yp = (int8_t*)malloc(phw ? phw : 1);
#pragma omp for schedule(dynamic, 1)
    for (int n = 0; n < L->N; n++) {
      const int8_t* Xn = L->ipool ? X + (size_t)n * L->IH * L->IW : X;
      int32_t* an = acc_out ? acc_out + (size_t)n * ohw : acc;
      layer_channel(L, Xn, L->ipool ? NULL : code + n * kk, L->ipool ? NULL : &P[n],
                    R ? R + n * phw : NULL, out + n * osz, an, y, yp);
    }
    free(acc);
    free(y);
    free(yp);
  }
}

int tf2o_run_network(int n_layers, const tf2o_layer* layers, const int32_t* in_idx,
                     const int32_t* out_idx, const int32_t* out_ch0, const int32_t* add_idx,
                     int n_tensors, const int32_t* tC, const int32_t* tH, const int32_t* tW,
                     const uint8_t* codes, const int64_t* code_off, const tf2o_bias_bn* params,
                     const int64_t* param_off, const int8_t* input, int n_images,
                     int result_tensor, int8_t* out, int n_threads) {
#ifdef _OPENMP
  if (n_threads > 0) omp_set_num_threads(n_threads);
#else
  (void)n_threads;
#endif
  size_t* toff = (size_t*)malloc(sizeof(size_t) * (n_tensors + 1));
  toff[0] = 0;
  for (int t = 0; t < n_tensors; t++) toff[t + 1] = toff[t] + (size_t)tC[t] * tH[t] * tW[t];
  const size_t in_sz = (size_t)tC[0] * tH[0] * tW[0];
  const size_t res_sz = (size_t)tC[result_tensor] * tH[result_tensor] * tW[result_tensor];
  size_t max_ohw = 1, max_phw = 1;
  for (int l = 0; l < n_layers; l++) {
    size_t a = (size_t)layers[l].OH * layers[l].OW, b = (size_t)layers[l].PH * layers[l].PW;
    if (a > max_ohw) max_ohw = a;
    if (b > max_phw) max_phw = b;
  }
  int rc = 0;
  /* images are independent (sequencer.cl:58-62 iterates frames sequentially): parallel over
   * (image) when there are many, else over output channels inside each layer */
  const int par_images = n_images >= 2;
#pragma omp parallel if (par_images)
  {
    int8_t* buf = (int8_t*)malloc(toff[n_tensors]);
    int32_t* acc = (int32_t*)malloc(sizeof(int32_t) * max_ohw);
    int8_t* y = (int8_t*)malloc(max_ohw);
    int8_t* yp = (int8_t*)malloc(max_phw);
    if (!buf || !acc || !y || !yp) {
      rc = -1;
    } else {
#pragma omp for schedule(dynamic, 1)
      for (int img = 0; img < n_images; img++) {
        memcpy(buf, input + (size_t)img * in_sz, in_sz);
        for (int l = 0; l < n_layers; l++) {
          const tf2o_layer* L = &layers[l];
          const int8_t* X = buf + toff[in_idx[l]];
          const int ot = out_idx[l];
          const size_t osz = L->gap ? 1 : (size_t)L->PH * L->PW;
          int8_t* O = buf + toff[ot] + (size_t)out_ch0[l] * osz;
          const int8_t* R = add_idx[l] >= 0 ? buf + toff[add_idx[l]] : NULL;
          const size_t kk = (size_t)L->C * L->k * L->k;
          if (par_images) {
            for (int n = 0; n < L->N; n++) {
              const int8_t* Xn = L->ipool ? X + (size_t)n * L->IH * L->IW : X;
              layer_channel(L, Xn, L->ipool ? NULL : codes + code_off[l] + n * kk,
                            L->ipool ? NULL : &params[param_off[l] + n],
                            R ? R + (size_t)n * L->PH * L->PW : NULL, O + n * osz, acc, y, yp);
            }
          } else {
            tf2o_layer_forward(L, X, codes + code_off[l], params + param_off[l], R, O, NULL);
          }
        }
        memcpy(out + (size_t)img * res_sz, buf + toff[result_tensor], res_sz);
      }
    }
    free(buf);
    free(acc);
    free(y);
    free(yp);
  }
  free(toff);
  return rc;
}
