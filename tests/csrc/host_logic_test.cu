// CPU test driver for the HOST logic of the product library (no GPU, no CUDA call): it #includes
// tf2_b200/csrc/api.cu, so the code under test is the shipped translation unit itself — weight preparation
// (LoadModel codes -> per-channel base shift + power-of-two planes for both kernel families, the negated-copy
// handling of the int8 -128 quirk, the unscaled low plane, the pixel-pair K layout), the range analysis that
// selects the folded / hi32 epilogue, and the 4-bit expansion of tf2b_load_layer_packed4.
//
// Input: a case file written by tests/test_host_logic.py (layer descriptors as the C ABI structs, codes as
// formats.codes_from_nibbles / the synthetic models produce them, nibbles + Q rows, BiasBnParam).  For every
// layer the driver loads the codes, rebuilds every weight from the prepared planes and compares it with the
// code it came from; loads the same layer through the packed4 entry point and compares the whole prepared
// state; prints one line per layer.  Exit code 0 = all equal.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../tf2_b200/csrc/api.cu"

static std::vector<unsigned char> read_all(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { perror(path); exit(2); }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  std::vector<unsigned char> b((size_t)n);
  if (fread(b.data(), 1, (size_t)n, f) != (size_t)n) exit(2);
  fclose(f);
  return b;
}

struct Reader {
  const unsigned char* p;
  int32_t i32() { int32_t v; memcpy(&v, p, 4); p += 4; return v; }
  const unsigned char* bytes(size_t n) { const unsigned char* q = p; p += n; return q; }
};

// value the code stands for, as (sign, shift); zero codes -> shift = -1
static void decode(uint8_t cd, int& sign, int& shift) {
  if (cd & 0x40) { sign = 0; shift = -1; return; }
  sign = (cd & 0x80) ? -1 : 1;
  shift = cd & 0x1f;
}

static long long check_planes(const tf2b_net* net, const LayerState& S, const uint8_t* codes, char* why) {
  const tf2b_layer_desc& d = S.d;
  const int C = d.C, N = d.N, k = d.k;
  const bool quirk = d.in_may_be_m128 != 0;
  long long bad = 0;
  // stored position of a logical input / output channel (the engine may reorder the channels of inner tensors)
  const std::vector<int>* vin = net->tpos[d.in_tensor].empty() ? nullptr : &net->tpos[d.in_tensor];
  const std::vector<int>* vout = net->tpos[d.out_tensor].empty() ? nullptr : &net->tpos[d.out_tensor];
  auto pin = [&](int c) { return vin ? (*vin)[c] : c; };
  auto pout = [&](int n) { return vout ? (*vout)[n] : n; };
  if (vin) {   // a permutation, and sorted by the consumer-side offset classes it was derived for
    std::vector<char> seen(vin->size(), 0);
    for (int v : *vin) { if (v < 0 || v >= (int)vin->size() || seen[v]) { snprintf(why, 200, "input order is not a permutation"); return 1; } seen[v] = 1; }
  }
  // ---- shift-accumulate kernel: packed 4-bit codes in segments.  weight = +-2^e << seg_shift << base[n]; a segment
  //      flagged "neg" multiplies the int8-negated activation (stands for negative weights of a layer whose input may
  //      hold -128), and for tensor 0 the negative weights sit as magnitudes on the channels of its negated copy
  {
    const bool dual = quirk && d.in_tensor == 0;
    const int KC = tf2b::sa_kc(S.Cp_s);
    const int cchunks = (S.Cp_s + KC - 1) / KC;
    const size_t row_bytes = (size_t)S.Kp_s / 2;
    if ((size_t)S.nseg_s * S.Npad_s * row_bytes != S.h_w4.size()) { snprintf(why, 200, "w4 size"); return 1; }
    size_t nz = 0, nz_codes = 0;
    for (size_t i = 0; i < S.h_w4.size(); i++) nz += ((S.h_w4[i] & 7) != 7) + (((S.h_w4[i] >> 4) & 7) != 7);
    for (int g = 1; g < S.nseg_s; g++)
      if (S.seg_shift_s[g] > S.seg_shift_s[g - 1]) { if (!bad) snprintf(why, 200, "segments not in descending shift order"); bad++; }
    for (int n = 0; n < N; n++)
      for (int c = 0; c < C; c++)
        for (int t = 0; t < k * k; t++) {
          int sign, shift;
          decode(codes[((size_t)n * C + c) * k * k + t], sign, shift);
          long long pos[2] = {0, 0}, neg[2] = {0, 0};   // [half]: plain / negated-copy channel; pos / neg: plain / negating segment
          for (int half = 0; half < (dual ? 2 : 1); half++) {
            const size_t kidx = (size_t)t * cchunks * KC + (half ? S.Cp + c : pin(c));
            for (int g = 0; g < S.nseg_s; g++) {
              const uint8_t byte = S.h_w4[((size_t)g * S.Npad_s + pout(n)) * row_bytes + kidx / 2];
              const unsigned nib = (kidx & 1) ? (byte >> 4) : (byte & 15);
              if ((nib & 7) == 7) { if (nib & 8) { if (!bad) snprintf(why, 200, "negative zero code"); bad++; } continue; }
              long long v = (1ll << (nib & 7)) * (1ll << S.seg_shift_s[g]) * (1ll << S.h_nshift[pout(n)]);
              if (nib & 8) v = -v;
              if (S.seg_neg_s[g]) neg[half] += v; else pos[half] += v;
            }
          }
          long long want_pos[2] = {0, 0}, want_neg[2] = {0, 0};
          if (shift >= 0) {
            nz_codes++;
            if (!quirk) want_pos[0] = sign * (1ll << shift);
            else if (sign > 0) want_pos[0] = 1ll << shift;
            else if (dual) want_pos[1] = 1ll << shift;
            else want_neg[0] = 1ll << shift;
          }
          if (pos[0] != want_pos[0] || pos[1] != want_pos[1] || neg[0] != want_neg[0] || neg[1] != want_neg[1]) {
            if (!bad) snprintf(why, 200, "w4 n=%d c=%d t=%d", n, c, t);
            bad++;
          }
        }
    if (nz != nz_codes) { if (!bad) snprintf(why, 200, "w4 holds %zu non-zeros for %zu codes", nz, nz_codes); bad++; }
  }
  // ---- tensor-core planes (int8)
  if (S.mma_ok) {
    const bool dual = quirk;
    const int in_pitch = net->tpitch[d.in_tensor];
    const bool pair = tf2b::mma_pair_mode(k, d.stride, d.pad, S.Cp_m, in_pitch, d.OW, d.OH, N, S.planes_m);
    const int Cpm = pair ? ((k + 1) / 2) * 128 : round_up(S.Cp_m, tf2b::mma_pick_bk(S.Cp_m));
    size_t nz = 0, nz_codes = 0;
    for (size_t i = 0; i < S.h_w8.size(); i++) nz += S.h_w8[i] != 0;
    for (int n = 0; n < N; n++)
      for (int c = 0; c < C; c++)
        for (int t = 0; t < k * k; t++) {
          int sign, shift;
          decode(codes[((size_t)n * C + c) * k * k + t], sign, shift);
          long long got[2] = {0, 0};   // [0]: multiplies channel c, [1]: multiplies the negated copy Cp + c
          for (int half = 0; half < (dual ? 2 : 1); half++) {
            const int cc = half ? S.Cp + c : pin(c);
            size_t kidx = (size_t)t * Cpm + cc;
            if (pair) {
              const int fh = t / k, fw = t - fh * k;
              kidx = (size_t)fh * Cpm + (size_t)(fw / 2) * 128 + (size_t)(fw & 1) * 64 + cc;
            }
            for (int p = 0; p < S.planes_m; p++) {
              long long v = S.h_w8[((size_t)p * S.Npad_m + pout(n)) * S.Kp_m + kidx];
              if (v && (v & (v - 1)) && ((-v) & (-v - 1))) { if (!bad) snprintf(why, 200, "w8 not a power of two"); bad++; }
              if (v > 64 || v < -64) { if (!bad) snprintf(why, 200, "w8 magnitude above 2^6"); bad++; }
              v *= 1ll << S.plane_shift_m[p];
              if (p != S.low_plane_m) v *= 1ll << S.h_nshift_m[pout(n)];
              got[half] += v;
            }
          }
          long long want[2] = {0, 0};
          if (shift >= 0) {
            nz_codes++;
            if (dual && sign < 0) want[1] = 1ll << shift; else want[0] = sign * (1ll << shift);
          }
          if (got[0] != want[0] || got[1] != want[1]) { if (!bad) snprintf(why, 200, "w8 n=%d c=%d t=%d", n, c, t); bad++; }
        }
    if (nz != nz_codes) { if (!bad) snprintf(why, 200, "w8 holds %zu non-zeros for %zu codes", nz, nz_codes); bad++; }
  }
  return bad;
}

static bool same_state(const LayerState& a, const LayerState& b) {
  return a.h_w4 == b.h_w4 && a.h_w8 == b.h_w8 && a.h_nshift == b.h_nshift && a.h_nshift_m == b.h_nshift_m &&
         a.h_bias == b.h_bias && a.h_alpha == b.h_alpha && a.h_beta == b.h_beta && a.nseg_s == b.nseg_s &&
         a.planes_m == b.planes_m && a.mma_ok == b.mma_ok && a.fast_requant == b.fast_requant &&
         a.low_plane_m == b.low_plane_m && a.Kp_m == b.Kp_m && a.Kp_s == b.Kp_s;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::vector<unsigned char> file = read_all(argv[1]);
  Reader r{file.data()};
  const int n_tensors = r.i32(), n_layers = r.i32();
  tf2b_net net;                                  // built by hand: tf2b_create needs a CUDA device
  net.tensors.resize(n_tensors);
  memcpy(net.tensors.data(), r.bytes(sizeof(tf2b_tensor_desc) * n_tensors), sizeof(tf2b_tensor_desc) * n_tensors);
  net.tpitch.resize(n_tensors);
  for (int t = 0; t < n_tensors; t++) net.tpitch[t] = round_up(net.tensors[t].C, 16);
  net.t0_neg_off = round_up(net.tensors[0].C, 16);   // as tf2b_create
  net.tpitch[0] = 2 * net.t0_neg_off;
  net.layers.resize(n_layers);
  const int B = argc > 2 ? atoi(argv[2]) : 256;
  net.tbuf.resize(n_tensors);
  for (int t = 0; t < n_tensors; t++) net.tbuf[t] = reinterpret_cast<int8_t*>((uintptr_t)0x100000000ull + (uintptr_t)t * 0x10000000ull);
  net.scratch0 = reinterpret_cast<int8_t*>((uintptr_t)0x4000000000ull);
  tf2b_net net4 = net;
  struct Case { const uint8_t* codes; const tf2b_bias_bn* params; int has4; int rc; int rc4; };
  std::vector<Case> cases(n_layers, Case{nullptr, nullptr, 0, 0, 0});
  for (int l = 0; l < n_layers; l++) {
    tf2b_layer_desc d;
    memcpy(&d, r.bytes(sizeof d), sizeof d);
    net.layers[l].d = d;
    net4.layers[l].d = d;
    if (d.ipool) continue;                       // pseudo layer: no weights follow in the case file
    const size_t cnt = (size_t)d.N * d.C * d.k * d.k;
    cases[l].codes = r.bytes(cnt);
    cases[l].params = (const tf2b_bias_bn*)r.bytes(sizeof(tf2b_bias_bn) * d.N);
    cases[l].has4 = r.i32();
    cases[l].rc = tf2b_load_layer(&net, l, cases[l].codes, cases[l].params);
    if (cases[l].has4) {
      const int min_exp = r.i32();
      const uint8_t* nib = r.bytes((cnt + 1) / 2);
      const int8_t* q_in = (const int8_t*)r.bytes(d.C);
      const int8_t* q_out = (const int8_t*)r.bytes(d.N);
      cases[l].rc4 = tf2b_load_layer_packed4(&net4, l, nib, min_exp, q_in, q_out, cases[l].params);
    } else {
      cases[l].rc4 = tf2b_load_layer(&net4, l, cases[l].codes, cases[l].params);
    }
  }
  // what tf2b_finalize does before anything touches the device: channel orders, then every layer's weight forms
  const int rcp = prepare_network(&net, true), rcp4 = prepare_network(&net4, true);
  int rc_all = (rcp != TF2B_OK || rcp4 != TF2B_OK) ? 1 : 0;
  int n_ordered = 0;
  for (auto& v : net.tpos) n_ordered += !v.empty();
  printf("prepare rc=%d rc4=%d ordered_tensors=%d %s\n", rcp, rcp4, n_ordered, net.err.c_str());
  for (int l = 0; l < n_layers; l++) {
    const tf2b_layer_desc d = net.layers[l].d;
    if (d.ipool) continue;
    const int rc = cases[l].rc != TF2B_OK ? cases[l].rc : rcp;
    char why[200] = "";
    long long bad = rc == TF2B_OK ? check_planes(&net, net.layers[l], cases[l].codes, why) : -1;
    int same4 = -1;
    if (cases[l].has4) same4 = cases[l].rc4 == TF2B_OK && same_state(net.layers[l], net4.layers[l]) && net.tpos == net4.tpos;
    const LayerState& S = net.layers[l];
    // launch plan of the tensor-core path for a batch of B images: what tf2b_layer_mode reports after finalize
    // (geometry only: the buffer addresses are placeholders, nothing is dereferenced)
    std::string mode = "shift";
    if (rc == TF2B_OK && S.mma_ok) {
      const bool to_scratch = d.pool || d.gap;
      int8_t* dst = to_scratch ? net.scratch0 : net.tbuf[d.out_tensor] + d.out_ch0;
      const int dstC = to_scratch ? round_up(d.N, 16) : net.tpitch[d.out_tensor];
      const int8_t* res = (d.add_tensor >= 0 && !d.pool) ? net.tbuf[d.add_tensor] : nullptr;
      const int resC = (d.add_tensor >= 0 && !d.pool) ? net.tpitch[d.add_tensor] : 0;
      ConvParams p = conv_params(&net, S, B, dst, dstC, res, resC, true);
      mode = tf2b::mma_describe(p, S.planes_m);
      for (auto& ch : mode) if (ch == ' ') ch = '_';
    }
    printf("layer %d rc=%d bad=%lld segs_s=%d planes_m=%d low=%d mma_ok=%d fast_requant=%d packed4_same=%d mode=%s %s\n", l, rc,
           bad, S.nseg_s, S.planes_m, S.low_plane_m, (int)S.mma_ok, S.fast_requant, same4, mode.c_str(), why);
    if (rc != TF2B_OK || bad != 0 || same4 == 0) rc_all = 1;
  }
  return rc_all;
}
