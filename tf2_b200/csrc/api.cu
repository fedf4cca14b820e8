// api.cu — engine object, weight preparation, execution plan and the C ABI of libtf2b200.so.
//
// Host-side mirror of the reference's NetWork (Runtime_Engine/cnn/host/src/network.cpp:22-168:
// owns filter / bias_bn / feature buffers) and Runner (runner.cpp:54-196: runs the layer
// sequence), but for a CUDA device: tensors live in HBM as NHWC int8, layers run as kernel launches
// on one stream.  See include/tf2b200.h for the contract of every entry point.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/tf2b200.h"
#include "common.cuh"

namespace tf2b {
// CUDA-core shift-accumulate path (conv_sa.cu)
cudaError_t launch_conv_sa(const ConvParams& p, int nseg, const int* seg_shift, const int* seg_neg, const int* seg_cbeg,
                           const int* seg_cend, const unsigned char* kmask_dev, const void* tmaps, int ksplit, int num_sms,
                           cudaStream_t stream);
cudaError_t sa_prepare_device();
int sa_kc(int Cp);
int sa_npad(int N);
int sa_max_segments();
size_t sa_tmap_bytes();
int sa_ksplit(const ConvParams& p, int num_sms);
std::string sa_describe(const ConvParams& p, int nseg, int ksplit);
int sa_build_tmaps(void* host_tmaps, const ConvParams& p, const uint8_t* wgt4, int nseg, int ksplit, std::string* err);
cudaError_t launch_chw_to_hwc(const int8_t*, int8_t*, int, int, int, int, int, int, cudaStream_t);
cudaError_t launch_hwc_to_chw(const int8_t*, int8_t*, int, int, int, int, int, const int*, cudaStream_t);
cudaError_t launch_hwc_repitch(const int8_t*, int8_t*, size_t, int, int, int, int, const int*, cudaStream_t);
cudaError_t launch_raw224_to_s2d(const int8_t*, int8_t*, int, int, cudaStream_t);
cudaError_t launch_maxpool3x3(const int8_t*, int8_t*, const int8_t*, int, int, int, int, int, int,
                              int, int, int, int, int, int, cudaStream_t);
cudaError_t launch_gap(const int8_t*, int8_t*, int, int, int, int, int, cudaStream_t);
// tensor-core path (conv_mma.cu)
bool mma_layer_supported(const tf2b_layer_desc& L, int in_pitch, int planes8);
struct MmaHostParams {
  const int32_t* bias;
  const int32_t* alpha;
  const int32_t* beta;
  const uint8_t* nshift;
};
cudaError_t launch_conv_mma(const ConvParams& p, const MmaHostParams& hp, int planes8,
                            const int* plane8_shift, void* tmaps, int num_sms, cudaStream_t stream);
cudaError_t mma_prepare_device(int* num_sms);
size_t mma_tmap_bytes();
std::string mma_describe(const ConvParams& p, int planes8);
int mma_build_tmaps(void* host_tmaps, const ConvParams& p, const int8_t* wgt8, const uint8_t* wgt4, int planes8,
                    std::string* err);
long long mma_slab_bytes(int k, int Cp, int N, int planes8);
int mma_bn();
int mma_pick_bk(int Cp);
bool mma_pair_mode(int k, int stride, int pad, int Cp, int xC, int OW, int OH, int N, int planes8);
bool mma_sparse2_ok(const tf2b_layer_desc& L, int in_pitch, int N);
}  // namespace tf2b

using tf2b::ConvParams;

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct LayerState {
  tf2b_layer_desc d;
  bool loaded = false;     // weights were handed over (tf2b_load_layer*) or imported with a blob
  bool prepared = false;   // device-side weight forms have been built (prepare_network / blob import)
  std::vector<uint8_t> h_codes;          // LoadModel's codes [N][C][k][k], kept until tf2b_finalize
  std::vector<tf2b_bias_bn> h_params;    // [N]
  int Cp = 0;  // reduction channels padded to 16
  int Cp_m = 0;  // reduction channels seen by the tensor-core path (2*Cp when it reads the negated copy)
  // --- shift-accumulate kernel: packed 4-bit codes, one row block per segment (exponent level x sign-quirk) ---
  int Npad_s = 0, Kp_s = 0, Cp_s = 0, nseg_s = 0, ksplit_s = 0;
  int seg_shift_s[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // descending: the kernel combines the segment sums by Horner's rule
  int seg_neg_s[8] = {0, 0, 0, 0, 0, 0, 0, 0};     // 1: the segment multiplies the int8-negated activations
  int seg_cbeg_s[8] = {0, 0, 0, 0, 0, 0, 0, 0};    // channel chunks [cbeg, cend) of every tap the segment spans
  int seg_cend_s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::vector<uint8_t> h_w4;
  std::vector<uint8_t> h_kmask;   // [nseg][taps * chunks]: 16-channel steps of a (segment, tap, chunk) that hold weights
  std::vector<unsigned char> h_tmaps_s;
  // --- mma kernel (int8 planes) ---
  int Npad_m = 0, Kp_m = 0, planes_m = 0;
  int plane_shift_m[tf2b::kMaxPlanes] = {0, 0, 0, 0};
  int low_plane_m = -1;  // index of the plane that holds absolute shifts 0..6 (not scaled by 2^nshift), or -1
  int sparse2_m = 0;     // two planes whose second one is sparse in K: per (tap, 32-channel block) a 2-bit mask says
                         // which planes hold weights; the MMA warp issues per plane and skips empty blocks
  std::vector<uint8_t> h_blkmask;   // [taps][K chunks]: 2 bits per 32-channel block of the chunk
  int fast_requant = 0;  // range analysis: no int32 intermediate of pe.cl:191-194 can wrap for this layer
  std::vector<int8_t> h_w8;
  std::vector<uint8_t> h_w8p;   // the same planes as 4-bit codes (layers whose weight slab is resident in shared memory)
  size_t off_w8p = 0;
  std::vector<uint8_t> h_nshift_m;  // per-channel base shift of the tensor-core planes
  bool mma_ok = false;
  // --- per-channel params (padded to max(Npad_s, Npad_m)) ---
  int Npar = 0;
  std::vector<int32_t> h_bias, h_alpha, h_beta;
  std::vector<uint8_t> h_nshift;
  // device views inside the arena
  size_t off_kmask = 0;
  size_t off_w4 = 0, off_w8 = 0, off_bias = 0, off_alpha = 0, off_beta = 0, off_nshift = 0, off_nshift_m = 0;
  int kernel = 0;  // 0 none (ipool), 1 shift, 2 mma
  std::vector<unsigned char> h_tmaps;  // CUtensorMap blobs for the mma path (host copy)
  std::string mode_desc;               // tf2b_layer_mode() text
  size_t off_tmaps = 0;
};

struct BlobLayerMeta {  // fixed-size, trivially copyable: travels inside the weight blob
  int32_t loaded, Cp, Cp_m, Npad_s, Kp_s, Cp_s, nseg_s, seg_shift_s[8], seg_neg_s[8], seg_cbeg_s[8], seg_cend_s[8];
  int32_t Npad_m, Kp_m, planes_m, plane_shift_m[4], mma_ok, Npar, low_plane_m, nshift_m_len, fast_requant;
  int32_t sparse2_m, blkmask_len;
  uint8_t blkmask[320];
  int64_t off_kmask, kmask_len, off_w8p, w8p_len;
  int64_t off_w4, off_w8, off_bias, off_alpha, off_beta, off_nshift, off_nshift_m;
};

struct tf2b_net {
  int device = 0;
  int num_sms = 0;          // of `device` (grids of the persistent kernels)
  std::vector<tf2b_tensor_desc> tensors;
  std::vector<int> tpitch;  // channel pitch of each tensor (C rounded up to 16)
  int t0_neg_off = 0;       // channel offset of the negated copy inside tensor 0
  // Channel order of a tensor in memory (empty = natural): tpos[t][c] = position of logical channel c.  Chosen at
  // tf2b_finalize so that channels whose weights share a power-of-two offset sit next to each other in K (the
  // producers' weight rows and the consumers' weight columns are laid out accordingly: free at run time); reads
  // at the boundary (tf2b_read_tensor, the result) and the accumulator tap undo it.
  std::vector<std::vector<int>> tpos;
  std::vector<size_t> off_tpos, off_tinv;   // device copies inside the arena (int32 [C]): logical->position, position->logical
  std::vector<LayerState> layers;
  int variant = TF2B_VARIANT_AUTO;
  int max_images = 0;
  bool finalized = false;
  int result_tensor = -1;
  // device memory
  unsigned char* arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<int8_t*> tbuf;
  int8_t* scratch0 = nullptr;
  int8_t* scratch1 = nullptr;
  size_t scratch_bytes = 0;
  int8_t* io_in = nullptr;   // staging for host-buffer entry points
  int8_t* io_out = nullptr;
  size_t io_in_bytes = 0, io_out_bytes = 0;
  cudaStream_t own_stream = nullptr;
  // pipelined host entry point (tf2b_submit_raw224_host / tf2b_wait): two slots, three streams
  int8_t* slot_in[2] = {nullptr, nullptr};
  int8_t* slot_out[2] = {nullptr, nullptr};
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
  bool slot_used[2] = {false, false};
  int last_launches = 0;
  int last_images = 0;
  // CUDA-graph executor: the layer sequence of a batch size is captured once (programmatic-dependent-launch
  // edges included) and replayed; key = number of images
  // layer schedule over streams (finalize): independent branches (inception branches, the shortcut convolution of a
  // residual block) run concurrently; lane 0 is the caller's stream, lanes 1.. are the engine's side streams
  static constexpr int kLanes = 4;
  std::vector<int> lane;                       // per layer
  std::vector<std::vector<int>> wait_on;       // per layer: layers on other lanes it must wait for
  std::vector<char> need_event;                // per layer: another lane waits for it
  std::vector<cudaEvent_t> ev_layer;
  cudaStream_t side[kLanes] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[kLanes] = {nullptr, nullptr, nullptr, nullptr};
  bool lane_used[kLanes] = {true, false, false, false};
  bool multi_stream = true;
  bool permute = true;      // choose_permutations at finalize (off: natural channel order everywhere)
  // Chunked stem (raw 3x224x224 entry point, first layer = convolution + pool): the batch goes through
  // space-to-depth -> conv1 -> max pool in chunks of kStemChunk images whose intermediates live in two chunk-sized
  // buffers that are REUSED by every chunk, so they stay in the 126 MB L2 (no DRAM round trip of the 64-byte/pixel
  // tensor 0 and of conv1's 112x112x64 map).  Tensor 0 is then not materialised for the whole batch; the debug taps
  // re-create it from the last raw input on demand.
  static constexpr int kStemChunk = 32;
  bool stem_chunked = false;
  bool stem_chunk_on = false;   // measured (profiles/r02_ab_options.md): 21 more launches of smaller kernels cost more than the
                                // saved DRAM round trips — 100.6 k images/s with the chunked stem, 104.6 k without
  int8_t* stem_t0 = nullptr;          // [kStemChunk][114][114][64]
  int8_t* stem_conv = nullptr;        // [kStemChunk][OH][OW][Np16]
  std::vector<unsigned char> stem_tmaps;
  const int8_t* last_raw = nullptr;   // raw input of the last chunked run (debug taps)
  int last_raw_images = 0;
  bool t0_stale = false;
  int weight_staging = TF2B_WEIGHTS_PLANES;   // tensor-core path, resident-weight layers: int8 planes or packed 4-bit tiles
  bool use_graph = true;
  struct GraphEntry { int B; const void* raw; int launches; cudaGraphExec_t exec; };
  std::vector<GraphEntry> graphs;
  bool profile = false;
  std::vector<cudaEvent_t> ev;  // 3 per layer: layer start, conv end, layer end
  std::string err;
};

static int fail(tf2b_net* n, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (n) n->err = buf;
  return code;
}

#define CUDA_TRY(n, expr)                                                                  \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return fail(n, TF2B_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                  __FILE__, __LINE__);                                                     \
  } while (0)

// tf2b_create has no handle to hang its error text on: one slot per calling thread
static thread_local std::string g_create_err;
extern "C" {
static int ensure_io(tf2b_net* net);
static int build_tmaps(tf2b_net* net);
}

// ------------------------------------------------------------------------------------------------
// weight preparation: LoadModel codes -> per-channel base shift + power-of-two weight planes
// ------------------------------------------------------------------------------------------------
static int prepare_layer(tf2b_net* net, LayerState& S, const uint8_t* codes,
                         const tf2b_bias_bn* params) {
  const tf2b_layer_desc& d = S.d;
  const int C = d.C, N = d.N, k = d.k;
  // positions of the logical input / output channels in their tensors (natural order when the tensor has none)
  const std::vector<int>* vin = (d.in_tensor < (int)net->tpos.size() && !net->tpos[d.in_tensor].empty()) ? &net->tpos[d.in_tensor] : nullptr;
  const std::vector<int>* vout = (d.out_tensor < (int)net->tpos.size() && !net->tpos[d.out_tensor].empty()) ? &net->tpos[d.out_tensor] : nullptr;
  auto pin = [&](int c) { return vin ? (*vin)[c] : c; };
  auto pout = [&](int n) { return vout ? (*vout)[n] : n; };
  S.Cp = round_up(C, 16);
  const int Ktot = k * k * S.Cp;
  // per-channel base shift = smallest shift among the channel's non-zero codes
  std::vector<uint8_t> base(N, 0);
  int max_rel = 0;
  for (int n = 0; n < N; n++) {
    int mn = 99, mx = -1;
    const uint8_t* cn = codes + (size_t)n * C * k * k;
    for (int i = 0; i < C * k * k; i++) {
      if (cn[i] & 0x40) continue;
      int s = cn[i] & 0x1f;
      mn = std::min(mn, s);
      mx = std::max(mx, s);
    }
    if (mx < 0) { mn = 0; mx = 0; }
    base[n] = (uint8_t)mn;
    max_rel = std::max(max_rel, mx - mn);
  }
  const bool quirk = d.in_may_be_m128 != 0;
  // ---- shift-accumulate kernel: 4-bit codes (bit 3 = negative, bits 0..2 = exponent e in 0..6, 7 = zero) in
  //      segments: shift = base[n] + D + e with one D per segment.  Two ways to pick D, the cheaper one wins:
  //        plain levels    D = 7 * floor(rel / 7): every segment walks the whole K range;
  //        offset classes  D = off[c] + 7 * floor((rel - off[c]) / 7) with off[c] the input channel's own offset
  //                        (q_in[c] - min q_in): when the tensor's channels are ordered by that offset a segment
  //                        only spans the chunks of its class and K is walked ONCE in total.
  //      The int8 negate quirk (pe.cl:32-34) is carried either by the negated copy of tensor 0 (extra channels,
  //      positive magnitudes) or, for any other tensor that may hold -128, by segments that multiply the
  //      byte-negated activations.
  {
    const int lv = 7;
    const bool dual_s = quirk && d.in_tensor == 0 && S.Cp == net->t0_neg_off;
    const bool negseg = quirk && !dual_s;
    S.Cp_s = dual_s ? 2 * S.Cp : S.Cp;
    const int KC = tf2b::sa_kc(S.Cp_s);
    const int cchunks = (S.Cp_s + KC - 1) / KC;
    S.Npad_s = tf2b::sa_npad(N);
    S.Kp_s = k * k * cchunks * KC;
    const int kk = k * k;
    // per input channel: smallest relative shift of any of its weights
    std::vector<int> off(C, 99);
    for (int n = 0; n < N; n++)
      for (int c = 0; c < C; c++)
        for (int t = 0; t < kk; t++) {
          const uint8_t cd = codes[((size_t)n * C + c) * kk + t];
          if (!(cd & 0x40)) off[c] = std::min(off[c], (cd & 0x1f) - base[n]);
        }
    for (int c = 0; c < C; c++) if (off[c] == 99) off[c] = 0;
    struct Seg { int D, ng, cb, ce; };
    auto plan = [&](bool classes, std::vector<Seg>& segs) -> long long {
      segs.clear();
      for (int n = 0; n < N; n++)
        for (int c = 0; c < C; c++)
          for (int t = 0; t < kk; t++) {
            const uint8_t cd = codes[((size_t)n * C + c) * kk + t];
            if (cd & 0x40) continue;
            const int rel = (cd & 0x1f) - base[n];
            const int o = classes ? off[c] : 0;
            const int D = o + lv * ((rel - o) / lv);
            const bool negw = (cd & 0x80) != 0;
            const int ng = (negseg && negw) ? 1 : 0;
            const int cc = (negw && dual_s) ? S.Cp + c : pin(c);
            const int chunk = cc / KC;
            Seg* f = nullptr;
            for (auto& sg : segs) if (sg.D == D && sg.ng == ng) { f = &sg; break; }
            if (!f) { segs.push_back({D, ng, chunk, chunk + 1}); continue; }
            f->cb = std::min(f->cb, chunk);
            f->ce = std::max(f->ce, chunk + 1);
          }
      if (segs.empty()) segs.push_back({0, 0, 0, cchunks});
      std::sort(segs.begin(), segs.end(), [](const Seg& a, const Seg& b) { return a.D != b.D ? a.D > b.D : a.ng < b.ng; });
      long long cost = 0;
      for (auto& sg : segs) cost += sg.ce - sg.cb;
      return (int)segs.size() <= tf2b::sa_max_segments() ? cost : -1;
    };
    std::vector<Seg> seg_plain, seg_class;
    const long long cost_plain = plan(false, seg_plain), cost_class = plan(true, seg_class);
    if (cost_plain < 0 && cost_class < 0)
      return fail(net, TF2B_ERR_ARG, "layer needs more than %d weight segments", tf2b::sa_max_segments());
    const bool classes = cost_class >= 0 && (cost_plain < 0 || cost_class < cost_plain);
    const std::vector<Seg>& segs = classes ? seg_class : seg_plain;
    S.nseg_s = (int)segs.size();
    for (int g = 0; g < 8; g++) {
      S.seg_shift_s[g] = g < S.nseg_s ? segs[g].D : 0;
      S.seg_neg_s[g] = g < S.nseg_s ? segs[g].ng : 0;
      S.seg_cbeg_s[g] = g < S.nseg_s ? segs[g].cb : 0;
      S.seg_cend_s[g] = g < S.nseg_s ? segs[g].ce : 0;
    }
    S.h_w4.assign((size_t)S.nseg_s * S.Npad_s * (S.Kp_s / 2), 0x77);
    S.h_kmask.assign((size_t)S.nseg_s * kk * cchunks, 0);
    for (int n = 0; n < N; n++)
      for (int c = 0; c < C; c++)
        for (int t = 0; t < kk; t++) {
          const uint8_t cd = codes[((size_t)n * C + c) * kk + t];
          if (cd & 0x40) continue;
          const int rel = (cd & 0x1f) - base[n];
          const int o = classes ? off[c] : 0;
          const int D = o + lv * ((rel - o) / lv);
          const bool negw = (cd & 0x80) != 0;
          const int ng = (negseg && negw) ? 1 : 0;
          int g = 0;
          while (g < S.nseg_s && !(segs[g].D == D && segs[g].ng == ng)) g++;
          int cc = pin(c);
          unsigned nib = (unsigned)(rel - D);
          if (negw) {
            if (dual_s) cc = S.Cp + c;          // magnitude times the negated copy (tensor 0 keeps its order)
            else if (!negseg) nib |= 8u;        // plain negative weight
          }
          const size_t kidx = ((size_t)t * cchunks * KC) + cc;
          uint8_t& byte = S.h_w4[((size_t)g * S.Npad_s + pout(n)) * (S.Kp_s / 2) + kidx / 2];
          byte = (kidx & 1) ? (uint8_t)((byte & 0x0f) | (nib << 4)) : (uint8_t)((byte & 0xf0) | nib);
          S.h_kmask[(size_t)g * kk * cchunks + (size_t)t * cchunks + cc / KC] |= (uint8_t)(1u << ((cc % KC) / 16));
        }
  }
  // ---- mma kernel planes: int8 +-2^e, e in 0..6.  A layer whose input may hold -128 can use the
  //      tensor cores only when that input is tensor 0 (which carries the negated copy): positive
  //      weights multiply channel c, magnitudes of negative weights multiply channel Cp + c.
  S.mma_ok = false;
  S.planes_m = 0;
  S.h_w8.clear();
  S.Cp_m = S.Cp;
  if (!quirk || d.in_tensor == 0) {
    const int lv = 7;
    // Two decompositions: (a) every code relative to the channel's smallest shift; (b) codes with an
    // absolute shift <= 6 (e.g. the code-0 taps LoadModel leaves in the transformed first layer,
    // model_loader.cpp:247) go to a "low" plane that is added unscaled, the rest relative to the
    // smallest shift above 6.  Take whichever needs fewer planes.
    std::vector<uint8_t> base_hi(N, 0);
    int max_rel_hi = 0;
    bool any_low = false;
    for (int n = 0; n < N; n++) {
      int mn = 99, mx = -1;
      const uint8_t* cn = codes + (size_t)n * C * k * k;
      for (int i = 0; i < C * k * k; i++) {
        if (cn[i] & 0x40) continue;
        int sft = cn[i] & 0x1f;
        if (sft < lv) { any_low = true; continue; }
        mn = std::min(mn, sft);
        mx = std::max(mx, sft);
      }
      if (mx < 0) { mn = 0; mx = 0; }
      base_hi[n] = (uint8_t)mn;
      max_rel_hi = std::max(max_rel_hi, mx - mn);
    }
    const int np_std = max_rel / lv + 1;
    const int np_alt = (max_rel_hi / lv + 1) + (any_low ? 1 : 0);
    const bool use_low = any_low && np_alt < np_std;
    const int np = use_low ? np_alt : np_std;
    const std::vector<uint8_t>& bm = use_low ? base_hi : base;
    const int in_pitch = net->tpitch[d.in_tensor];
    const bool dual = quirk;
    if (dual) S.Cp_m = 2 * S.Cp;
    if ((!dual || S.Cp == net->t0_neg_off) && np <= tf2b::kMaxPlanes && tf2b::mma_layer_supported(d, in_pitch, np)) {
      S.planes_m = np;
      S.low_plane_m = use_low ? np - 1 : -1;
      S.Npad_m = round_up(N, tf2b::mma_bn());
      // pixel-pair rows (conv_mma.cu pair_mode): K = filter row x ceil(k/2) chunks of [tap 2j | tap 2j+1]
      const bool pair = tf2b::mma_pair_mode(k, d.stride, d.pad, S.Cp_m, in_pitch, d.OW, d.OH, N, np);
      const int pair_chunks = (k + 1) / 2;
      const int Cpm = pair ? pair_chunks * 128 : round_up(S.Cp_m, tf2b::mma_pick_bk(S.Cp_m));
      S.Kp_m = pair ? k * Cpm : k * k * Cpm;
      S.h_w8.assign((size_t)np * S.Npad_m * S.Kp_m, 0);
      for (int p = 0; p < np; p++) S.plane_shift_m[p] = lv * p;
      if (use_low) S.plane_shift_m[np - 1] = 0;
      // Sparse second plane (two planes, no low plane, no negated copy): the second plane's shift D1 is free, and a
      // channel whose weights all fit [D1, D1 + 6] can live in plane 1 alone, one whose weights all fit [0, 6] in
      // plane 0 alone; only the others straddle.  With the tensor's channels ordered by their offset, whole
      // 32-channel blocks of K then hold weights in ONE plane and the MMA warp skips the other (api: blkmask).
      // D1 is chosen to minimise the number of (block, plane) pairs that have to be issued.
      S.sparse2_m = 0;
      std::vector<char> chan_plane;   // per logical channel: 0 / 1 = that plane only, 2 = per weight (rel <= 6 -> plane 0)
      if (np == 2 && !use_low && !dual && !pair) {
        const int kk2 = k * k;
        std::vector<int> lo(C, 99), hi(C, -1);
        for (int n = 0; n < N; n++)
          for (int c = 0; c < C; c++)
            for (int t = 0; t < kk2; t++) {
              const uint8_t cd = codes[((size_t)n * C + c) * kk2 + t];
              if (cd & 0x40) continue;
              const int rel = (cd & 0x1f) - bm[n];
              lo[c] = std::min(lo[c], rel);
              hi[c] = std::max(hi[c], rel);
            }
        const int nblk = Cpm / 32;
        long long best_cost = -1;
        int best_d1 = lv;
        std::vector<char> best_plane;
        for (int d1 = 1; d1 <= lv; d1++) {
          bool ok = true;
          std::vector<char> cp(C, 0);
          std::vector<char> use0(nblk, 0), use1(nblk, 0);
          for (int c = 0; c < C && ok; c++) {
            if (hi[c] < 0) continue;                                  // no weights at all
            const int b = pin(c) / 32;
            if (hi[c] <= 6) { cp[c] = 0; use0[b] = 1; }
            else if (lo[c] >= d1 && hi[c] <= d1 + 6) { cp[c] = 1; use1[b] = 1; }
            else {
              cp[c] = 2;
              use0[b] = use1[b] = 1;
              if (hi[c] > d1 + 6) ok = false;                         // a weight above plane 1's range
              // weights in (6, d1) cannot exist: d1 <= 7
            }
          }
          if (!ok) continue;
          long long cost = 0;
          for (int b = 0; b < nblk; b++) cost += use0[b] + use1[b];
          if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_d1 = d1; best_plane = cp; }
        }
        if (best_cost >= 0 && best_cost * 10 <= (long long)nblk * 2 * 9 && tf2b::mma_sparse2_ok(d, in_pitch, N) &&
            (size_t)(k * k) * (size_t)(Cpm / tf2b::mma_pick_bk(S.Cp_m)) <= sizeof(((ConvParams*)nullptr)->blkmask)) {
          S.sparse2_m = 1;
          S.plane_shift_m[1] = best_d1;
          chan_plane = best_plane;
        }
      }
      const int d1s = S.plane_shift_m[1];
      for (int n = 0; n < N; n++)
        for (int c = 0; c < C; c++)
          for (int t = 0; t < k * k; t++) {
            uint8_t cd = codes[((size_t)n * C + c) * k * k + t];
            if (cd & 0x40) continue;
            const int sft = cd & 0x1f;
            int p, e;
            if (use_low && sft < lv) {
              p = np - 1;
              e = sft;
            } else if (S.sparse2_m) {
              const int rel = sft - bm[n];
              p = chan_plane[c] == 2 ? (rel <= 6 ? 0 : 1) : chan_plane[c];
              e = p ? rel - d1s : rel;
            } else {
              int rel = sft - bm[n];
              p = rel / lv;
              e = rel - p * lv;
            }
            int v = 1 << e;
            int cc = pin(c);
            if (cd & 0x80) {
              if (dual) cc = S.Cp + c; else v = -v;
            }
            size_t kidx = (size_t)t * Cpm + cc;
            if (pair) {
              const int fh = t / k, fw = t - fh * k;
              kidx = (size_t)fh * Cpm + (size_t)(fw / 2) * 128 + (size_t)(fw & 1) * 64 + cc;
            }
            S.h_w8[((size_t)p * S.Npad_m + pout(n)) * S.Kp_m + kidx] = (int8_t)v;
          }
      // which planes hold weights in each (tap, 32-channel block): 2 bits per block, one byte per (tap, K chunk)
      {
        const int BKm = tf2b::mma_pick_bk(S.Cp_m);
        const int kch = pair ? pair_chunks : Cpm / BKm;
        const int ntap = pair ? k : k * k;
        S.h_blkmask.assign((size_t)ntap * kch, 0);
        if (S.sparse2_m) {
          for (int p = 0; p < np; p++)
            for (int n = 0; n < N; n++) {
              const int8_t* row = S.h_w8.data() + ((size_t)p * S.Npad_m + pout(n)) * S.Kp_m;
              for (int kx = 0; kx < S.Kp_m; kx++)
                if (row[kx]) {
                  const int tap = kx / Cpm, cc2 = kx - tap * Cpm;
                  S.h_blkmask[(size_t)tap * kch + cc2 / BKm] |= (uint8_t)(1u << (2 * ((cc2 % BKm) / 32) + p));
                }
            }
          // every plane needs at least one issued block per tile (its first MMA clears the accumulator)
          for (int p = 0; p < 2; p++) {
            bool any = false;
            for (auto m : S.h_blkmask) any = any || (m & (0x55u << p));
            if (!any) S.h_blkmask[0] |= (uint8_t)(1u << p);
          }
        }
      }
      // packed 4-bit copy of the planes for the layers that keep their weights resident in shared memory: every
      // entry of a plane is 0 or +-2^e, e <= 6 — exactly one 4-bit code (bit 3 = negative, bits 0..2 = e, 7 = zero)
      S.h_w8p.clear();
      if (net->weight_staging == TF2B_WEIGHTS_PACKED4 && !pair && tf2b::mma_slab_bytes(k, S.Cp_m, N, np) > 0 &&
          S.Kp_m % 32 == 0) {
        S.h_w8p.assign(S.h_w8.size() / 2, 0x77);
        for (size_t i = 0; i < S.h_w8.size(); i++) {
          const int v = S.h_w8[i];
          if (!v) continue;
          unsigned e = 0;
          for (int m = v < 0 ? -v : v; m > 1; m >>= 1) e++;
          const unsigned nib = e | (v < 0 ? 8u : 0u);
          uint8_t& b8 = S.h_w8p[i >> 1];
          b8 = (i & 1) ? (uint8_t)((b8 & 0x0f) | (nib << 4)) : (uint8_t)((b8 & 0xf0) | nib);
        }
      }
      S.h_nshift_m.assign(round_up(std::max(tf2b::sa_npad(N), S.Npad_m), 16), 0);
      for (int n = 0; n < N; n++) S.h_nshift_m[pout(n)] = bm[n];
      S.mma_ok = true;
    }
  }
  // Range analysis for the fused requantisation of the tensor-core epilogue.  |acc| <= |bias| +
  // 128 * sum_k 2^shift_k (capped at 2^31: a wrapped accumulator is still an int32), a = acc*alpha >> 20,
  // s = a + beta, then (s >> 14 + 1) >> 1.  If |a| + |beta| + 2^14 < 2^31 for every channel nothing wraps and
  // ((a + beta) >> 14 + 1) >> 1 == (acc*alpha + ((beta + 2^14) << 20)) >> 35 exactly.
  S.fast_requant = 1;
  for (int n = 0; n < N && S.fast_requant; n++) {
    long double sum = fabsl((long double)params[n].bias);
    const uint8_t* cn = codes + (size_t)n * C * k * k;
    for (int i = 0; i < C * k * k; i++)
      if (!(cn[i] & 0x40)) sum += 128.0L * (long double)(1ull << (cn[i] & 0x1f));
    if (sum > 2147483648.0L) sum = 2147483648.0L;
    const long double amax = sum * fabsl((long double)params[n].alpha) / 1048576.0L + 1.0L;
    if (amax + fabsl((long double)params[n].beta) + 16384.0L + 2.0L >= 2147483647.0L) S.fast_requant = 0;
  }
  // Folded form of the tensor-core epilogue: acc = tot * 2^nshift + bias must hold in the integers (no
  // wrap-around of the accumulator) and alpha << nshift must be an int32; then
  // acc*alpha + ((beta + 2^14) << 20) = tot * (alpha << nshift) + [bias*alpha + ((beta + 2^14) << 20)].
  if (S.fast_requant && S.mma_ok) {
    bool fold = true;
    for (int n = 0; n < N && fold; n++) {
      long double sum = fabsl((long double)params[n].bias);
      const uint8_t* cn = codes + (size_t)n * C * k * k;
      for (int i = 0; i < C * k * k; i++)
        if (!(cn[i] & 0x40)) sum += 128.0L * (long double)(1ull << (cn[i] & 0x1f));
      if (sum >= 2147483647.0L) fold = false;
      const long double a = fabsl((long double)params[n].alpha) * (long double)(1ull << S.h_nshift_m[pout(n)]);
      if (a >= 2147483647.0L) fold = false;
    }
    if (fold) {
      S.fast_requant = 2;
      // every base shift >= 3: the epilogue can drop the final shift (conv_mma.cu HI32)
      bool hi = true;
      for (int n = 0; n < N; n++) hi = hi && S.h_nshift_m[pout(n)] >= 3;
      if (hi) S.fast_requant = 3;
    }
  }
  S.Npar = std::max(S.Npad_s, S.Npad_m);
  S.h_bias.assign(S.Npar, 0);
  S.h_alpha.assign(S.Npar, 0);
  S.h_beta.assign(S.Npar, 0);
  S.h_nshift.assign(round_up(S.Npar, 16), 0);
  for (int n = 0; n < N; n++) {
    S.h_bias[pout(n)] = params[n].bias;
    S.h_alpha[pout(n)] = params[n].alpha;
    S.h_beta[pout(n)] = params[n].beta;
    S.h_nshift[pout(n)] = base[n];
  }
  S.prepared = true;
  return TF2B_OK;
}

// ------------------------------------------------------------------------------------------------
// channel order of the inner tensors
// ------------------------------------------------------------------------------------------------
// Per input channel of a layer: the smallest shift, relative to each output channel's base shift, any of its
// weights carries.  With TF2's per-channel Q tables this is (Q_in[c] - min Q_in): shift = 15 + q_in[c] - q_out[n]
// - exponent (model_loader.cpp:159-168).
static std::vector<int> channel_offsets(const LayerState& S) {
  const tf2b_layer_desc& d = S.d;
  const int C = d.C, N = d.N, kk = d.k * d.k;
  std::vector<int> base(N, 0), off(C, 99);
  for (int n = 0; n < N; n++) {
    int mn = 99;
    const uint8_t* cn = S.h_codes.data() + (size_t)n * C * kk;
    for (int i = 0; i < C * kk; i++)
      if (!(cn[i] & 0x40)) mn = std::min(mn, cn[i] & 0x1f);
    base[n] = mn == 99 ? 0 : mn;
  }
  for (int n = 0; n < N; n++) {
    const uint8_t* cn = S.h_codes.data() + (size_t)n * C * kk;
    for (int c = 0; c < C; c++)
      for (int t = 0; t < kk; t++) {
        const uint8_t cd = cn[(size_t)c * kk + t];
        if (!(cd & 0x40)) off[c] = std::min(off[c], (cd & 0x1f) - base[n]);
      }
  }
  for (int c = 0; c < C; c++)
    if (off[c] == 99) off[c] = 0;
  return off;
}

// Tensors joined by a residual add share one order (feature_writer.cl:124-127 adds element by element); concat
// buffers, tensors touched by an ipool pseudo layer and tensor 0 keep the natural order.  Within a class the
// channels are sorted (stably) by the offset its heaviest consumer sees, so that channels of equal offset are
// contiguous in that consumer's K dimension.
static void choose_permutations(tf2b_net* net) {
  const int T = (int)net->tensors.size();
  net->tpos.assign(T, {});
  std::vector<int> parent(T);
  for (int t = 0; t < T; t++) parent[t] = t;
  auto find = [&](int x) { while (parent[x] != x) x = parent[x] = parent[parent[x]]; return x; };
  std::vector<char> ok(T, 1);
  std::vector<int> writers(T, 0);
  ok[0] = 0;
  for (auto& S : net->layers) {
    const tf2b_layer_desc& d = S.d;
    writers[d.out_tensor]++;
    if (d.ipool) { ok[d.in_tensor] = ok[d.out_tensor] = 0; continue; }
    if (d.out_ch0 != 0 || d.N != net->tensors[d.out_tensor].C) ok[d.out_tensor] = 0;
    if (d.C != net->tensors[d.in_tensor].C) ok[d.in_tensor] = 0;
    if (!S.loaded || S.h_codes.empty()) ok[d.in_tensor] = ok[d.out_tensor] = 0;
    if (d.add_tensor >= 0) parent[find(d.out_tensor)] = find(d.add_tensor);
  }
  for (int t = 0; t < T; t++)
    if (writers[t] != 1) ok[t] = 0;
  std::vector<char> class_ok(T, 1);
  for (int t = 0; t < T; t++)
    if (!ok[t]) class_ok[find(t)] = 0;
  for (int root = 0; root < T; root++) {
    if (find(root) != root || !class_ok[root]) continue;
    const LayerState* best = nullptr;
    double best_work = -1.0;
    for (auto& S : net->layers) {
      if (S.d.ipool || find(S.d.in_tensor) != root) continue;
      const double w = (double)S.d.k * S.d.k * S.d.C * S.d.N * S.d.OH * S.d.OW;
      if (w > best_work) { best_work = w; best = &S; }
    }
    if (!best) continue;
    const std::vector<int> off = channel_offsets(*best);
    const int C = (int)off.size();
    if (*std::min_element(off.begin(), off.end()) == *std::max_element(off.begin(), off.end())) continue;
    std::vector<int> order(C);
    for (int c = 0; c < C; c++) order[c] = c;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return off[a] < off[b]; });
    std::vector<int> pos(C);
    for (int p = 0; p < C; p++) pos[order[p]] = p;
    for (int t = 0; t < T; t++)
      if (find(t) == root) net->tpos[t] = pos;
  }
}

// Builds the device-side weight forms of every loaded layer (tf2b_finalize; the CPU test driver calls it directly).
static int prepare_network(tf2b_net* net, bool permute) {
  if (permute) choose_permutations(net);
  else net->tpos.assign(net->tensors.size(), {});
  for (auto& S : net->layers) {
    if (S.d.ipool || S.prepared) continue;
    if (S.h_codes.empty()) return fail(net, TF2B_ERR_STATE, "a layer has no weights loaded");
    int rc = prepare_layer(net, S, S.h_codes.data(), S.h_params.data());
    if (rc != TF2B_OK) return rc;
  }
  return TF2B_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

const char* tf2b_version(void) { return "tf2b200 0.1 (sm_100a)"; }

const char* tf2b_last_error(tf2b_net* net) { return net ? net->err.c_str() : g_create_err.c_str(); }

int tf2b_create(const tf2b_tensor_desc* tensors, int n_tensors, const tf2b_layer_desc* layers,
                int n_layers, int device, tf2b_net** out) {
  if (!tensors || !layers || !out || n_tensors <= 0 || n_layers <= 0) {
    g_create_err = "tf2b_create: null/empty argument";
    return TF2B_ERR_ARG;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    g_create_err = std::string("tf2b_create: no usable CUDA device (") +
                   (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range") +
                   "); this engine has no CPU fallback";
    return TF2B_ERR_CUDA;
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    g_create_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
    return TF2B_ERR_CUDA;
  }
  tf2b_net* n = new tf2b_net();
  n->device = device;
  n->tensors.assign(tensors, tensors + n_tensors);
  n->tpitch.resize(n_tensors);
  for (int t = 0; t < n_tensors; t++) {
    if (tensors[t].C <= 0 || tensors[t].H <= 0 || tensors[t].W <= 0) {
      g_create_err = "tf2b_create: tensor with non-positive dimension";
      delete n;
      return TF2B_ERR_ARG;
    }
    n->tpitch[t] = round_up(tensors[t].C, 16);
  }
  // Tensor 0 (the quantised image) can hold -128, where the reference negates inside int8
  // (pe.cl:32-34).  It is stored with a second, int8-negated copy of its channels so that the
  // tensor-core path can multiply the copy with the magnitudes of the negative weights.
  n->t0_neg_off = round_up(tensors[0].C, 16);
  n->tpitch[0] = 2 * n->t0_neg_off;
  n->layers.resize(n_layers);
  for (int l = 0; l < n_layers; l++) {
    const tf2b_layer_desc& d = layers[l];
    auto bad = [&](const char* why) {
      char b[256];
      snprintf(b, sizeof b, "tf2b_create: layer %d: %s", l, why);
      g_create_err = b;
      delete n;
      return TF2B_ERR_ARG;
    };
    if (d.in_tensor < 0 || d.in_tensor >= n_tensors || d.out_tensor < 0 || d.out_tensor >= n_tensors)
      return bad("tensor id out of range");
    if (d.add_tensor >= n_tensors) return bad("add_tensor out of range");
    if (d.out_ch0 % 16 != 0) return bad("concat channel offset must be a multiple of 16");
    if (d.out_ch0 + d.N > tensors[d.out_tensor].C) return bad("output channels exceed out_tensor");
    const tf2b_tensor_desc& ti = tensors[d.in_tensor];
    const tf2b_tensor_desc& to = tensors[d.out_tensor];
    if (d.ipool) {
      if (d.N != ti.C) return bad("ipool must keep the channel count");
      if (d.PH != to.H || d.PW != to.W) return bad("ipool output size mismatch");
    } else {
      if (d.C > ti.C || d.C <= 0 || d.N <= 0 || d.k <= 0 || d.stride <= 0) return bad("bad conv geometry");
      if ((ti.H + 2 * d.pad - d.k) / d.stride + 1 != d.OH || (ti.W + 2 * d.pad - d.k) / d.stride + 1 != d.OW)
        return bad("OH/OW do not match (IH + 2*pad - k)/stride + 1");
      int eh = d.gap ? 1 : d.PH, ew = d.gap ? 1 : d.PW;
      if (eh != to.H || ew != to.W) return bad("output tensor size mismatch");
      if (!d.pool && (d.PH != d.OH || d.PW != d.OW)) return bad("PH/PW must equal OH/OW without pool");
    }
    if ((d.pool || d.ipool) && d.pool_stride != 1 && d.pool_stride != 2)
      return bad("pool stride must be 1 or 2 (pool.cl / pool_tail.cl)");
    if (d.add_tensor >= 0) {
      const tf2b_tensor_desc& tr = tensors[d.add_tensor];
      if (tr.H != d.PH || tr.W != d.PW || tr.C < d.N) return bad("residual tensor shape mismatch");
    }
    n->layers[l].d = d;
  }
  n->result_tensor = layers[n_layers - 1].out_tensor;
  e = cudaStreamCreateWithFlags(&n->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    g_create_err = std::string("cudaStreamCreate: ") + cudaGetErrorString(e);
    delete n;
    return TF2B_ERR_CUDA;
  }
  *out = n;
  return TF2B_OK;
}

int tf2b_load_layer(tf2b_net* net, int layer, const uint8_t* codes, const tf2b_bias_bn* params) {
  if (!net) return TF2B_ERR_ARG;
  if (layer < 0 || layer >= (int)net->layers.size()) return fail(net, TF2B_ERR_ARG, "layer %d out of range", layer);
  if (net->finalized) return fail(net, TF2B_ERR_STATE, "tf2b_load_layer after tf2b_finalize");
  LayerState& S = net->layers[layer];
  if (S.d.ipool) return fail(net, TF2B_ERR_ARG, "layer %d is an ipool pseudo layer (no weights)", layer);
  if (!codes || !params) return fail(net, TF2B_ERR_ARG, "null codes/params");
  // The device-side weight forms depend on the channel order chosen for the tensors, which needs every layer's
  // codes: keep a copy, build everything in tf2b_finalize (prepare_network)
  const size_t cnt = (size_t)S.d.N * S.d.C * S.d.k * S.d.k;
  S.h_codes.assign(codes, codes + cnt);
  S.h_params.assign(params, params + S.d.N);
  S.loaded = true;
  S.prepared = false;
  return TF2B_OK;
}

int tf2b_load_layer_packed4(tf2b_net* net, int layer, const uint8_t* nibbles, int min_exp,
                            const int8_t* q_in, const int8_t* q_out, const tf2b_bias_bn* params) {
  if (!net) return TF2B_ERR_ARG;
  if (layer < 0 || layer >= (int)net->layers.size()) return fail(net, TF2B_ERR_ARG, "layer %d out of range", layer);
  if (!nibbles || !q_in || !q_out || !params) return fail(net, TF2B_ERR_ARG, "null argument");
  const tf2b_layer_desc& d = net->layers[layer].d;
  if (d.ipool) return fail(net, TF2B_ERR_ARG, "layer %d is an ipool pseudo layer (no weights)", layer);
  const size_t kk = (size_t)d.k * d.k;
  std::vector<uint8_t> codes((size_t)d.N * d.C * kk);
  for (int n = 0; n < d.N; n++)
    for (int c = 0; c < d.C; c++) {
      // model_loader.cpp:159-162: expand = INFLAT + q_in[c] - q_out[n] (kept in a char)
      int8_t expand = (int8_t)(15 + q_in[c] - q_out[n]);
      for (size_t t = 0; t < kk; t++) {
        size_t i = ((size_t)n * d.C + c) * kk + t;
        uint8_t nib = (nibbles[i >> 1] >> ((i & 1) * 4)) & 0xF;
        uint8_t code;
        int e = nib & 7;
        if (e == 7) {
          // 4bit_data_format.txt: e=7 with sign bit 0 is 0.0; 15 is invalid -> treated as zero
          code = 0x40;
        } else {
          // value = +-2^(min_exp+e); Get_real finds i = -(min_exp+e) in 0..14, else i = 0
          int i2 = -(min_exp + e);
          if (i2 >= 17) {
            code = 0x40;  // |w| = 2^-17 < 1e-5: Get_real's zero test (model_loader.cpp:101-102)
          } else {
            if (i2 < 0 || i2 > 14) i2 = 0;
            int8_t sh = (int8_t)(expand - i2);
            if (sh < 0) sh = 0;
            code = (uint8_t)sh;
            if (!(nib & 8)) code |= 0x80;  // bit3 set = positive
          }
        }
        codes[i] = code;
      }
    }
  return tf2b_load_layer(net, layer, codes.data(), params);
}

static void drop_graphs(tf2b_net* net) {
  for (auto& g : net->graphs) cudaGraphExecDestroy(g.exec);
  net->graphs.clear();
}

int tf2b_set_graph(tf2b_net* net, int on) {
  if (!net) return TF2B_ERR_ARG;
  net->use_graph = on == 1 || on == 2;
  net->multi_stream = on == 1 || on == 3;
  drop_graphs(net);
  return TF2B_OK;
}

int tf2b_set_variant(tf2b_net* net, int variant) {
  if (!net) return TF2B_ERR_ARG;
  if (variant < 0 || variant > 2) return fail(net, TF2B_ERR_ARG, "unknown variant %d", variant);
  net->variant = variant;
  drop_graphs(net);   // captured graphs hold the kernels of the old plan
  for (auto& S : net->layers) {
    if (S.d.ipool) { S.kernel = 0; continue; }
    S.kernel = (variant != TF2B_VARIANT_SHIFT && S.mma_ok) ? 2 : 1;
  }
  if (net->finalized && net->stem_chunked) return build_tmaps(net);   // the chunked stem's maps belong to one kernel family
  return TF2B_OK;
}

int tf2b_set_weight_staging(tf2b_net* net, int mode) {
  if (!net) return TF2B_ERR_ARG;
  if (net->finalized) return fail(net, TF2B_ERR_STATE, "tf2b_set_weight_staging after tf2b_finalize");
  if (mode != TF2B_WEIGHTS_PLANES && mode != TF2B_WEIGHTS_PACKED4) return fail(net, TF2B_ERR_ARG, "unknown weight staging %d", mode);
  net->weight_staging = mode;
  return TF2B_OK;
}

int tf2b_set_stem_chunk(tf2b_net* net, int on) {
  if (!net) return TF2B_ERR_ARG;
  if (net->finalized) return fail(net, TF2B_ERR_STATE, "tf2b_set_stem_chunk after tf2b_finalize");
  net->stem_chunk_on = on != 0;
  return TF2B_OK;
}

int tf2b_set_result(tf2b_net* net, int tensor) {
  if (!net) return TF2B_ERR_ARG;
  if (tensor < 0 || tensor >= (int)net->tensors.size()) return fail(net, TF2B_ERR_ARG, "tensor out of range");
  net->result_tensor = tensor;
  if (net->finalized) {
    CUDA_TRY(net, cudaSetDevice(net->device));
    return ensure_io(net);
  }
  return TF2B_OK;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static size_t layout_arena(tf2b_net* net) {
  size_t off = 0;
  for (auto& S : net->layers) {
    if (S.d.ipool || !S.loaded) continue;
    S.off_w4 = off; off = align256(off + S.h_w4.size());
    S.off_kmask = off; off = align256(off + S.h_kmask.size());
    S.off_w8 = off; off = align256(off + S.h_w8.size());
    S.off_w8p = off; off = align256(off + S.h_w8p.size());
    S.off_bias = off; off = align256(off + (size_t)S.Npar * 4);
    S.off_alpha = off; off = align256(off + (size_t)S.Npar * 4);
    S.off_beta = off; off = align256(off + (size_t)S.Npar * 4);
    S.off_nshift = off; off = align256(off + S.h_nshift.size());
    S.off_nshift_m = off; off = align256(off + S.h_nshift_m.size());
  }
  net->off_tpos.assign(net->tensors.size(), 0);
  net->off_tinv.assign(net->tensors.size(), 0);
  for (size_t t = 0; t < net->tensors.size(); t++) {
    if (t >= net->tpos.size() || net->tpos[t].empty()) continue;
    net->off_tpos[t] = off; off = align256(off + net->tpos[t].size() * 4);
    net->off_tinv[t] = off; off = align256(off + net->tpos[t].size() * 4);
  }
  return off;
}

constexpr size_t kBlobFixed = 32;   // magic, layer count, hash of the layer / tensor tables, arena bytes
struct BlobTensorMeta {             // channel order of a tensor (position of every logical channel), if it has one
  int64_t off_pos, off_inv;
  int32_t channels, has_order;
};
static size_t blob_header_bytes(const tf2b_net* net) {
  return align256(kBlobFixed + net->layers.size() * sizeof(BlobLayerMeta) + net->tensors.size() * sizeof(BlobTensorMeta));
}
// FNV-1a over the tensor and layer tables: a blob only fits an engine created from the same tables
static uint64_t tables_hash(const tf2b_net* net) {
  uint64_t h = 1469598103934665603ull;
  auto mix = [&](const void* p, size_t n) {
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
  };
  for (auto& t : net->tensors) mix(&t, sizeof t);
  for (auto& S : net->layers) mix(&S.d, sizeof S.d);
  return h;
}

int64_t tf2b_weight_blob_bytes(tf2b_net* net) {
  if (!net) return TF2B_ERR_ARG;
  if (!net->finalized) return fail(net, TF2B_ERR_STATE, "finalize first");
  return (int64_t)(blob_header_bytes(net) + net->arena_bytes);
}

static int alloc_runtime(tf2b_net* net);
static int build_schedule(tf2b_net* net);

int tf2b_finalize(tf2b_net* net, int max_images) {
  if (!net) return TF2B_ERR_ARG;
  if (net->finalized) return fail(net, TF2B_ERR_STATE, "already finalized");
  if (max_images <= 0) return fail(net, TF2B_ERR_ARG, "max_images must be positive");
  for (size_t l = 0; l < net->layers.size(); l++)
    if (!net->layers[l].d.ipool && !net->layers[l].loaded)
      return fail(net, TF2B_ERR_STATE, "layer %zu has no weights loaded", l);
  CUDA_TRY(net, cudaSetDevice(net->device));
  CUDA_TRY(net, tf2b::mma_prepare_device(&net->num_sms));
  CUDA_TRY(net, tf2b::sa_prepare_device());
  net->max_images = max_images;
  {
    bool any_raw = false;
    for (auto& S : net->layers) any_raw = any_raw || (!S.d.ipool && !S.prepared);
    if (any_raw) {   // (an imported blob brings prepared layers and their channel orders)
      int rcp = prepare_network(net, net->permute);
      if (rcp != TF2B_OK) return rcp;
    }
    if (net->tpos.size() != net->tensors.size()) net->tpos.assign(net->tensors.size(), {});
    for (auto& S : net->layers) { S.h_codes.clear(); S.h_codes.shrink_to_fit(); }
  }
  net->arena_bytes = layout_arena(net);
  CUDA_TRY(net, cudaMalloc(&net->arena, std::max<size_t>(net->arena_bytes, 256)));
  for (auto& S : net->layers) {
    if (S.d.ipool) continue;
    auto up = [&](size_t off, const void* src, size_t bytes) -> cudaError_t {
      if (!bytes) return cudaSuccess;
      return cudaMemcpy(net->arena + off, src, bytes, cudaMemcpyHostToDevice);
    };
    CUDA_TRY(net, up(S.off_w4, S.h_w4.data(), S.h_w4.size()));
    CUDA_TRY(net, up(S.off_kmask, S.h_kmask.data(), S.h_kmask.size()));
    CUDA_TRY(net, up(S.off_w8, S.h_w8.data(), S.h_w8.size()));
    CUDA_TRY(net, up(S.off_w8p, S.h_w8p.data(), S.h_w8p.size()));
    CUDA_TRY(net, up(S.off_bias, S.h_bias.data(), (size_t)S.Npar * 4));
    CUDA_TRY(net, up(S.off_alpha, S.h_alpha.data(), (size_t)S.Npar * 4));
    CUDA_TRY(net, up(S.off_beta, S.h_beta.data(), (size_t)S.Npar * 4));
    CUDA_TRY(net, up(S.off_nshift, S.h_nshift.data(), S.h_nshift.size()));
    CUDA_TRY(net, up(S.off_nshift_m, S.h_nshift_m.data(), S.h_nshift_m.size()));
  }
  for (size_t t = 0; t < net->tensors.size(); t++) {
    if (net->tpos[t].empty()) continue;
    std::vector<int> inv(net->tpos[t].size());
    for (size_t c = 0; c < inv.size(); c++) inv[net->tpos[t][c]] = (int)c;
    CUDA_TRY(net, cudaMemcpy(net->arena + net->off_tpos[t], net->tpos[t].data(), inv.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(net, cudaMemcpy(net->arena + net->off_tinv[t], inv.data(), inv.size() * 4, cudaMemcpyHostToDevice));
  }
  int rc = alloc_runtime(net);
  if (rc != TF2B_OK) return rc;
  rc = build_schedule(net);
  if (rc != TF2B_OK) return rc;
  net->finalized = true;
  return tf2b_set_variant(net, net->variant);
}

// Builds the ConvParams of a layer for `B` images (pointers into the arena and tensor buffers).
static ConvParams conv_params(tf2b_net* net, const LayerState& S, int B, int8_t* dst, int dstC,
                              const int8_t* res, int resC, bool mma) {
  const tf2b_layer_desc& d = S.d;
  ConvParams p;
  memset(&p, 0, sizeof p);
  const tf2b_tensor_desc& ti = net->tensors[d.in_tensor];
  p.x = net->tbuf[d.in_tensor];
  p.y = dst;
  p.r = res;
  p.bias = reinterpret_cast<const int32_t*>(net->arena + S.off_bias);
  p.alpha = reinterpret_cast<const int32_t*>(net->arena + S.off_alpha);
  p.beta = reinterpret_cast<const int32_t*>(net->arena + S.off_beta);
  p.nshift = net->arena + (mma ? S.off_nshift_m : S.off_nshift);
  p.low_plane = mma ? S.low_plane_m : -1;
  p.fast_requant = mma ? S.fast_requant : 0;
  p.acc_dump = nullptr;
  p.B = B; p.IH = ti.H; p.IW = ti.W; p.Cp = mma ? S.Cp_m : S.Cp; p.xC = net->tpitch[d.in_tensor];
  p.OH = d.OH; p.OW = d.OW; p.N = d.N; p.yC = dstC; p.rC = resC;
  p.k = d.k; p.pad = d.pad; p.stride = d.stride;
  p.relu = d.relu; p.add_relu = d.add_relu;
  p.sparse2 = 0;
  p.w4_avail = 0;
  if (mma) {
    p.sparse2 = S.sparse2_m;
    p.w4_avail = (net->weight_staging == TF2B_WEIGHTS_PACKED4 && !S.h_w8p.empty()) ? 1 : 0;
    if (S.sparse2_m) memcpy(p.blkmask, S.h_blkmask.data(), std::min(S.h_blkmask.size(), sizeof p.blkmask));
    p.Npad = S.Npad_m; p.Kp = S.Kp_m; p.Ktot = S.Kp_m; p.planes = S.planes_m;
    for (int i = 0; i < tf2b::kMaxPlanes; i++) { p.plane_shift[i] = S.plane_shift_m[i]; p.plane_neg[i] = 0; }
  } else {
    p.Npad = S.Npad_s; p.Kp = S.Kp_s; p.Ktot = S.Kp_s; p.planes = S.nseg_s; p.Cp = S.Cp_s;
  }
  return p;
}

// Staging of the host-buffer entry points, allocated when the plan is frozen (never inside a run call):
// the synchronous pair io_in / io_out and the two slots of the pipelined call with their copy streams and
// events.  Inputs are sized for max_images raw or tensor-0 images, outputs for the RESULT tensor; a later
// tf2b_set_result to a larger tensor grows the output side.
static int ensure_io(tf2b_net* net) {
  const size_t B = (size_t)net->max_images;
  const size_t in_b = std::max((size_t)3 * 224 * 224, (size_t)net->tensors[0].C * net->tensors[0].H * net->tensors[0].W) * B;
  const tf2b_tensor_desc& tr = net->tensors[net->result_tensor];
  const size_t out_b = std::max((size_t)tr.C * tr.H * tr.W * B, (size_t)256);
  if (!net->s_h2d) {
    CUDA_TRY(net, cudaStreamCreateWithFlags(&net->s_h2d, cudaStreamNonBlocking));
    CUDA_TRY(net, cudaStreamCreateWithFlags(&net->s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      CUDA_TRY(net, cudaEventCreateWithFlags(&net->ev_h2d[i], cudaEventDisableTiming));
      CUDA_TRY(net, cudaEventCreateWithFlags(&net->ev_comp[i], cudaEventDisableTiming));
      CUDA_TRY(net, cudaEventCreateWithFlags(&net->ev_d2h[i], cudaEventDisableTiming));
    }
  }
  if (in_b > net->io_in_bytes) {
    CUDA_TRY(net, cudaDeviceSynchronize());
    int8_t** bufs[3] = {&net->io_in, &net->slot_in[0], &net->slot_in[1]};
    for (auto b : bufs) {
      if (*b) CUDA_TRY(net, cudaFree(*b));
      *b = nullptr;
      CUDA_TRY(net, cudaMalloc(b, in_b));
    }
    net->io_in_bytes = in_b;
  }
  if (out_b > net->io_out_bytes) {
    CUDA_TRY(net, cudaDeviceSynchronize());
    int8_t** bufs[3] = {&net->io_out, &net->slot_out[0], &net->slot_out[1]};
    for (auto b : bufs) {
      if (*b) CUDA_TRY(net, cudaFree(*b));
      *b = nullptr;
      CUDA_TRY(net, cudaMalloc(b, out_b));
    }
    net->io_out_bytes = out_b;
  }
  return TF2B_OK;
}

static int alloc_runtime(tf2b_net* net) {
  const int B = net->max_images;
  net->tbuf.assign(net->tensors.size(), nullptr);
  for (size_t t = 0; t < net->tensors.size(); t++) {
    // + slack: the pixel-pair rows of the tensor-core path read one pixel past the last one
    size_t bytes = (size_t)B * net->tensors[t].H * net->tensors[t].W * net->tpitch[t] + 256;
    CUDA_TRY(net, cudaMalloc(&net->tbuf[t], bytes));
    CUDA_TRY(net, cudaMemset(net->tbuf[t], 0, bytes));
  }
  size_t sc = 256;
  for (auto& S : net->layers) {
    const tf2b_layer_desc& d = S.d;
    if (d.ipool) continue;
    if (d.pool || d.gap) sc = std::max(sc, (size_t)B * d.OH * d.OW * round_up(d.N, 16));
  }
  net->scratch_bytes = sc;
  CUDA_TRY(net, cudaMalloc(&net->scratch0, sc));
  CUDA_TRY(net, cudaMalloc(&net->scratch1, sc));
  {
    int rc = ensure_io(net);
    if (rc != TF2B_OK) return rc;
  }
  // chunk buffers of the stem (only when a run can be larger than one chunk)
  {
    const tf2b_layer_desc& d0 = net->layers[0].d;
    const tf2b_tensor_desc& t0 = net->tensors[0];
    net->stem_chunked = net->stem_chunk_on && B > tf2b_net::kStemChunk && !d0.ipool && d0.in_tensor == 0 && d0.pool && !d0.gap &&
                        t0.C == 27 && t0.H == 114 && t0.W == 114;
    if (net->stem_chunked) {
      const size_t n = tf2b_net::kStemChunk;
      CUDA_TRY(net, cudaMalloc(&net->stem_t0, n * t0.H * t0.W * net->tpitch[0] + 256));
      CUDA_TRY(net, cudaMalloc(&net->stem_conv, n * d0.OH * d0.OW * round_up(d0.N, 16) + 256));
      CUDA_TRY(net, cudaMemset(net->stem_t0, 0, n * t0.H * t0.W * net->tpitch[0] + 256));
    }
  }
  return build_tmaps(net);
}

// Tensor maps depend on buffer addresses (tensors, scratch, arena): built when the plan is frozen and again
// whenever a buffer moves (tf2b_dump_acc growing the scratch).
static int build_tmaps(tf2b_net* net) {
  const int B = net->max_images;
  for (auto& S : net->layers) {
    if (!S.mma_ok) continue;
    const tf2b_layer_desc& d = S.d;
    bool to_scratch = d.pool || d.gap;
    int8_t* dst = to_scratch ? net->scratch0 : net->tbuf[d.out_tensor] + d.out_ch0;
    int dstC = to_scratch ? round_up(d.N, 16) : net->tpitch[d.out_tensor];
    const int8_t* res = (d.add_tensor >= 0 && !d.pool) ? net->tbuf[d.add_tensor] : nullptr;
    const int resC = (d.add_tensor >= 0 && !d.pool) ? net->tpitch[d.add_tensor] : 0;
    ConvParams p = conv_params(net, S, B, dst, dstC, res, resC, true);
    S.h_tmaps.assign(tf2b::mma_tmap_bytes(), 0);
    std::string err;
    int rc = tf2b::mma_build_tmaps(S.h_tmaps.data(), p, reinterpret_cast<const int8_t*>(net->arena + S.off_w8),
                                   S.h_w8p.empty() ? nullptr : net->arena + S.off_w8p, S.planes_m, &err);
    if (rc != 0) {
      // not fatal: fall back to the shift kernel for this layer, remember why
      S.mma_ok = false;
      net->err = "mma tensor map: " + err;
    }
  }
  // chunked stem: conv1's maps over the chunk buffers
  if (net->stem_chunked) {
    LayerState& S0 = net->layers[0];
    const tf2b_layer_desc& d0 = S0.d;
    const bool mma0 = S0.kernel == 2 ? S0.mma_ok : (net->variant != TF2B_VARIANT_SHIFT && S0.mma_ok);
    ConvParams p = conv_params(net, S0, tf2b_net::kStemChunk, net->stem_conv, round_up(d0.N, 16), nullptr, 0, mma0);
    p.x = net->stem_t0;
    std::string err;
    int rc;
    if (mma0) {
      net->stem_tmaps.assign(tf2b::mma_tmap_bytes(), 0);
      rc = tf2b::mma_build_tmaps(net->stem_tmaps.data(), p, reinterpret_cast<const int8_t*>(net->arena + S0.off_w8),
                                 S0.h_w8p.empty() ? nullptr : net->arena + S0.off_w8p, S0.planes_m, &err);
    } else {
      net->stem_tmaps.assign(tf2b::sa_tmap_bytes(), 0);
      rc = tf2b::sa_build_tmaps(net->stem_tmaps.data(), p, net->arena + S0.off_w4, S0.nseg_s, S0.ksplit_s, &err);
    }
    if (rc != 0) net->stem_chunked = false;   // fall back to the whole-batch stem
  }
  // ... and those of the shift-accumulate path (every convolution has one: it is the exact fallback)
  for (auto& S : net->layers) {
    if (S.d.ipool) continue;
    const tf2b_layer_desc& d = S.d;
    ConvParams p = conv_params(net, S, B, net->scratch0, round_up(d.N, 16), nullptr, 0, false);
    S.ksplit_s = tf2b::sa_ksplit(p, net->num_sms);
    S.h_tmaps_s.assign(tf2b::sa_tmap_bytes(), 0);
    std::string err;
    int rc = tf2b::sa_build_tmaps(S.h_tmaps_s.data(), p, net->arena + S.off_w4, S.nseg_s, S.ksplit_s, &err);
    if (rc != 0) return fail(net, TF2B_ERR_CUDA, "shift-accumulate tensor map: %s", err.c_str());
  }
  return TF2B_OK;
}

// Dependency schedule of the layer graph over kLanes streams.  A layer depends on every earlier layer that writes
// its input tensor (all branch tails of a concat buffer) or its residual operand, and on the previous user of the
// shared pool / gap scratch.  It continues the lane of a producer that is still the tail of its lane (so that
// programmatic dependent launch keeps working along chains), else it opens the next lane round robin.
static int build_schedule(tf2b_net* net) {
  const int L = (int)net->layers.size();
  net->lane.assign(L, 0);
  net->wait_on.assign(L, {});
  net->need_event.assign(L, 0);
  for (int i = 0; i < tf2b_net::kLanes; i++) net->lane_used[i] = i == 0;
  std::vector<int> tail(tf2b_net::kLanes, -1);
  int last_scratch = -1, rr = 0;
  for (int l = 0; l < L; l++) {
    const tf2b_layer_desc& d = net->layers[l].d;
    std::vector<int> deps;
    for (int j = 0; j < l; j++) {
      const tf2b_layer_desc& e = net->layers[j].d;
      if (e.out_tensor == d.in_tensor || (d.add_tensor >= 0 && e.out_tensor == d.add_tensor)) deps.push_back(j);
      // a slice of a concat buffer is written once, but keep writers of the same tensor ordered with later readers
    }
    const bool scratch = !d.ipool && (d.pool || d.gap);
    if (scratch && last_scratch >= 0) deps.push_back(last_scratch);
    int ln = -1;
    for (int j : deps)
      if (tail[net->lane[j]] == j && (ln < 0 || j == l - 1)) ln = net->lane[j];
    if (ln < 0) {
      if (deps.empty()) ln = 0;
      else { rr = rr % (tf2b_net::kLanes - 1) + 1; ln = rr; }
    }
    net->lane[l] = ln;
    net->lane_used[ln] = true;
    for (int j : deps)
      if (net->lane[j] != ln) { net->wait_on[l].push_back(j); net->need_event[j] = 1; }
    tail[ln] = l;
    if (scratch) last_scratch = l;
  }
  CUDA_TRY(net, cudaEventCreateWithFlags(&net->ev_fork, cudaEventDisableTiming));
  for (int i = 1; i < tf2b_net::kLanes; i++) {
    CUDA_TRY(net, cudaStreamCreateWithFlags(&net->side[i], cudaStreamNonBlocking));
    CUDA_TRY(net, cudaEventCreateWithFlags(&net->ev_join[i], cudaEventDisableTiming));
  }
  net->ev_layer.assign(L, nullptr);
  for (int l = 0; l < L; l++)
    if (net->need_event[l]) CUDA_TRY(net, cudaEventCreateWithFlags(&net->ev_layer[l], cudaEventDisableTiming));
  return TF2B_OK;
}

static int run_layers(tf2b_net* net, int B, cudaStream_t st0, int only_layer, int32_t* acc_dump,
                      const int8_t* raw_chunked = nullptr) {
  int launches = 0;
  const bool prof = net->profile && only_layer < 0 && !net->ev.empty();
  // several lanes only for whole-network runs outside the per-layer profiler
  const bool lanes = net->multi_stream && !prof && only_layer < 0 && !net->lane.empty();
  if (lanes) {
    CUDA_TRY(net, cudaEventRecord(net->ev_fork, st0));
    for (int i = 1; i < tf2b_net::kLanes; i++)
      if (net->lane_used[i]) CUDA_TRY(net, cudaStreamWaitEvent(net->side[i], net->ev_fork, 0));
  }
  for (size_t l = 0; l < net->layers.size(); l++) {
    if (only_layer >= 0 && (int)l != only_layer) continue;
    LayerState& S = net->layers[l];
    cudaStream_t st = st0;
    if (lanes) {
      if (net->lane[l] > 0) st = net->side[net->lane[l]];
      for (int j : net->wait_on[l]) CUDA_TRY(net, cudaStreamWaitEvent(st, net->ev_layer[j], 0));
    }
    if (prof) CUDA_TRY(net, cudaEventRecord(net->ev[3 * l], st));
    const tf2b_layer_desc& d = S.d;
    const tf2b_tensor_desc& ti = net->tensors[d.in_tensor];
    int8_t* out = net->tbuf[d.out_tensor] + d.out_ch0;
    const int outC = net->tpitch[d.out_tensor];
    const int8_t* res = d.add_tensor >= 0 ? net->tbuf[d.add_tensor] : nullptr;
    const int resC = d.add_tensor >= 0 ? net->tpitch[d.add_tensor] : 0;
    if (d.ipool) {
      if (only_layer >= 0) return fail(net, TF2B_ERR_ARG, "ipool layer has no accumulators");
      CUDA_TRY(net, tf2b::launch_maxpool3x3(net->tbuf[d.in_tensor], out, nullptr, B, ti.H, ti.W,
                                            net->tpitch[d.in_tensor], d.PH, d.PW, outC, 0, d.N, 1, 1, 0, st));
      launches++;
      if (prof) {
        CUDA_TRY(net, cudaEventRecord(net->ev[3 * l + 1], st));
        CUDA_TRY(net, cudaEventRecord(net->ev[3 * l + 2], st));
      }
      if (lanes && net->need_event[l]) CUDA_TRY(net, cudaEventRecord(net->ev_layer[l], st));
      continue;
    }
    if (l == 0 && raw_chunked != nullptr) {
      // chunked stem: space-to-depth -> conv1 -> pool per chunk through the L2-resident chunk buffers
      const int Np16c = round_up(d.N, 16);
      const bool mma0 = S.kernel == 2 && S.mma_ok;
      for (int b0 = 0; b0 < B; b0 += tf2b_net::kStemChunk) {
        const int nb = std::min(tf2b_net::kStemChunk, B - b0);
        CUDA_TRY(net, tf2b::launch_raw224_to_s2d(raw_chunked + (size_t)b0 * 3 * 224 * 224, net->stem_t0, nb, 1, st));
        ConvParams pc = conv_params(net, S, nb, net->stem_conv, Np16c, nullptr, 0, mma0);
        pc.x = net->stem_t0;
        if (mma0) {
          const tf2b::MmaHostParams hp = {S.h_bias.data(), S.h_alpha.data(), S.h_beta.data(), S.h_nshift_m.data()};
          CUDA_TRY(net, tf2b::launch_conv_mma(pc, hp, S.planes_m, S.plane_shift_m, net->stem_tmaps.data(), net->num_sms, st));
        } else {
          CUDA_TRY(net, tf2b::launch_conv_sa(pc, S.nseg_s, S.seg_shift_s, S.seg_neg_s, S.seg_cbeg_s, S.seg_cend_s,
                                             net->arena + S.off_kmask, net->stem_tmaps.data(), S.ksplit_s, net->num_sms, st));
        }
        const size_t poff = (size_t)b0 * d.PH * d.PW;
        CUDA_TRY(net, tf2b::launch_maxpool3x3(net->stem_conv, out + poff * outC, res ? res + poff * resC : nullptr, nb, d.OH, d.OW,
                                              Np16c, d.PH, d.PW, outC, resC, d.N, d.pool_stride, d.pool_pad, d.add_relu, st));
        launches += 3;
      }
      if (lanes && net->need_event[l]) CUDA_TRY(net, cudaEventRecord(net->ev_layer[l], st));
      continue;
    }
    // the accumulator tap re-runs one convolution with its feature map diverted to scratch; everything
    // else of the layer's launch plan (kernel family, staging mode, CTA pairs, residual) stays as it is
    const bool to_scratch = d.pool || d.gap || acc_dump != nullptr;
    const int Np16 = round_up(d.N, 16);
    int8_t* cdst = to_scratch ? net->scratch0 : out;
    const int cdstC = to_scratch ? Np16 : outC;
    // the residual add sits after the pool (pool_tail -> feature_writer): fuse it into the conv
    // epilogue only when there is no pool
    const int8_t* cres = d.pool ? nullptr : res;
    bool use_mma = S.kernel == 2 && S.mma_ok;
    ConvParams p = conv_params(net, S, B, cdst, cdstC, cres, resC, use_mma);
    p.acc_dump = acc_dump;
    p.acc_perm = (acc_dump != nullptr && !net->tpos[d.out_tensor].empty())
                     ? reinterpret_cast<const int*>(net->arena + net->off_tinv[d.out_tensor]) : nullptr;
    if (use_mma) {
      const tf2b::MmaHostParams hp = {S.h_bias.data(), S.h_alpha.data(), S.h_beta.data(), S.h_nshift_m.data()};
      CUDA_TRY(net, tf2b::launch_conv_mma(p, hp, S.planes_m, S.plane_shift_m, S.h_tmaps.data(), net->num_sms, st));
    } else {
      CUDA_TRY(net, tf2b::launch_conv_sa(p, S.nseg_s, S.seg_shift_s, S.seg_neg_s, S.seg_cbeg_s, S.seg_cend_s,
                                         net->arena + S.off_kmask, S.h_tmaps_s.data(), S.ksplit_s, net->num_sms, st));
    }
    launches++;
    if (prof) CUDA_TRY(net, cudaEventRecord(net->ev[3 * l + 1], st));
    if (only_layer >= 0) break;
    const int8_t* cur = cdst;
    int curC = cdstC, curH = d.OH, curW = d.OW;
    if (d.pool) {
      int8_t* pdst = d.gap ? net->scratch1 : out;
      int pdstC = d.gap ? Np16 : outC;
      CUDA_TRY(net, tf2b::launch_maxpool3x3(cur, pdst, res, B, curH, curW, curC, d.PH, d.PW, pdstC, resC,
                                            d.N, d.pool_stride, d.pool_pad, d.add_relu, st));
      launches++;
      cur = pdst; curC = pdstC; curH = d.PH; curW = d.PW;
    }
    if (d.gap) {
      CUDA_TRY(net, tf2b::launch_gap(cur, out, B, curH * curW, curC, outC, d.N, st));
      launches++;
    }
    if (prof) CUDA_TRY(net, cudaEventRecord(net->ev[3 * l + 2], st));
    if (lanes && net->need_event[l]) CUDA_TRY(net, cudaEventRecord(net->ev_layer[l], st));
  }
  if (lanes) {
    for (int i = 1; i < tf2b_net::kLanes; i++)
      if (net->lane_used[i]) {
        CUDA_TRY(net, cudaEventRecord(net->ev_join[i], net->side[i]));
        CUDA_TRY(net, cudaStreamWaitEvent(st0, net->ev_join[i], 0));
      }
  }
  net->last_launches += launches;
  return TF2B_OK;
}

// Runs the layer sequence for B images on `st`: replays the captured CUDA graph of this batch size (captured on
// first use), or launches kernel by kernel when graphs are off, the stream cannot capture (legacy default
// stream), or per-layer profiling is on.
static int run_layers_exec(tf2b_net* net, int B, cudaStream_t st, const int8_t* raw_chunked = nullptr) {
  if (!net->use_graph || net->profile || st == nullptr || st == cudaStreamLegacy)
    return run_layers(net, B, st, -1, nullptr, raw_chunked);
  for (auto& g : net->graphs)
    if (g.B == B && g.raw == (const void*)raw_chunked) {
      CUDA_TRY(net, cudaGraphLaunch(g.exec, st));
      net->last_launches += g.launches;
      return TF2B_OK;
    }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone)
    return run_layers(net, B, st, -1, nullptr, raw_chunked);   // the caller is capturing already: become part of its graph
  if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    return run_layers(net, B, st, -1, nullptr, raw_chunked);
  }
  const int before = net->last_launches;
  int rc = run_layers(net, B, st, -1, nullptr, raw_chunked);
  cudaGraph_t graph = nullptr;
  cudaError_t ce = cudaStreamEndCapture(st, &graph);
  const int launches = net->last_launches - before;
  if (rc != TF2B_OK || ce != cudaSuccess || !graph) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    net->use_graph = false;                      // capture is not possible here: stay with plain launches
    net->last_launches = before;
    return rc != TF2B_OK ? rc : run_layers(net, B, st, -1, nullptr, raw_chunked);
  }
  cudaGraphExec_t exec = nullptr;
  ce = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ce != cudaSuccess || !exec) {
    cudaGetLastError();
    net->use_graph = false;
    net->last_launches = before;
    return run_layers(net, B, st, -1, nullptr, raw_chunked);
  }
  if (net->graphs.size() >= 12) {                // a handful of batch sizes / input buffers at most
    cudaGraphExecDestroy(net->graphs.front().exec);
    net->graphs.erase(net->graphs.begin());
  }
  net->graphs.push_back({B, (const void*)raw_chunked, launches, exec});
  CUDA_TRY(net, cudaGraphLaunch(exec, st));
  return TF2B_OK;
}

// debug taps of a chunked run: tensor 0 was never materialised for the whole batch — rebuild it from the last raw input
static int refresh_t0(tf2b_net* net, cudaStream_t st) {
  if (!net->t0_stale || !net->last_raw) return TF2B_OK;
  CUDA_TRY(net, tf2b::launch_raw224_to_s2d(net->last_raw, net->tbuf[0], net->last_raw_images, 1, st));
  net->t0_stale = false;
  return TF2B_OK;
}

static int check_run(tf2b_net* net, int n_images) {
  if (!net) return TF2B_ERR_ARG;
  if (!net->finalized) return fail(net, TF2B_ERR_STATE, "tf2b_finalize has not been called");
  if (n_images <= 0 || n_images > net->max_images)
    return fail(net, TF2B_ERR_ARG, "n_images %d outside 1..%d", n_images, net->max_images);
  return TF2B_OK;
}

static int write_result(tf2b_net* net, int tensor, int B, int8_t* out_dev, int layout, cudaStream_t st) {
  const tf2b_tensor_desc& t = net->tensors[tensor];
  const int Cp = net->tpitch[tensor];
  // logical channel c lives at position pos[c] of the stored tensor
  const int* pos = net->tpos[tensor].empty() ? nullptr : reinterpret_cast<const int*>(net->arena + net->off_tpos[tensor]);
  if (layout == TF2B_LAYOUT_CHW) {
    CUDA_TRY(net, tf2b::launch_hwc_to_chw(net->tbuf[tensor], out_dev, B, t.C, t.H, t.W, Cp, pos, st));
  } else if (layout == TF2B_LAYOUT_HWC) {
    CUDA_TRY(net, tf2b::launch_hwc_repitch(net->tbuf[tensor], out_dev, (size_t)B * t.H * t.W, t.C, Cp, t.C, 0, pos, st));
  } else {
    return fail(net, TF2B_ERR_ARG, "unknown layout %d", layout);
  }
  net->last_launches++;
  return TF2B_OK;
}

int tf2b_run(tf2b_net* net, const int8_t* in_dev, int in_layout, int n_images, int8_t* out_dev,
             int out_layout, void* stream) {
  int rc = check_run(net, n_images);
  if (rc) return rc;
  if (!in_dev || !out_dev) return fail(net, TF2B_ERR_ARG, "null device pointer");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(net, cudaSetDevice(net->device));
  net->last_launches = 0;
  net->last_images = n_images;
  const tf2b_tensor_desc& t0 = net->tensors[0];
  if (in_layout == TF2B_LAYOUT_CHW) {
    CUDA_TRY(net, tf2b::launch_chw_to_hwc(in_dev, net->tbuf[0], n_images, t0.C, t0.H, t0.W, net->tpitch[0],
                                          net->t0_neg_off, st));
  } else if (in_layout == TF2B_LAYOUT_HWC) {
    CUDA_TRY(net, tf2b::launch_hwc_repitch(in_dev, net->tbuf[0], (size_t)n_images * t0.H * t0.W, t0.C, t0.C,
                                           net->tpitch[0], net->t0_neg_off, nullptr, st));
  } else {
    return fail(net, TF2B_ERR_ARG, "unknown layout %d", in_layout);
  }
  net->last_launches++;
  net->t0_stale = false;
  rc = run_layers_exec(net, n_images, st);
  if (rc) return rc;
  return write_result(net, net->result_tensor, n_images, out_dev, out_layout, st);
}

int tf2b_run_raw224(tf2b_net* net, const int8_t* raw_dev, int n_images, int8_t* out_dev, int out_layout,
                    void* stream) {
  int rc = check_run(net, n_images);
  if (rc) return rc;
  if (!raw_dev || !out_dev) return fail(net, TF2B_ERR_ARG, "null device pointer");
  const tf2b_tensor_desc& t0 = net->tensors[0];
  if (t0.C != 27 || t0.H != 114 || t0.W != 114)
    return fail(net, TF2B_ERR_ARG, "tensor 0 is %dx%dx%d, not the 27x114x114 space-to-depth input", t0.C, t0.H, t0.W);
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(net, cudaSetDevice(net->device));
  net->last_launches = 0;
  net->last_images = n_images;
  if (net->stem_chunked && n_images > tf2b_net::kStemChunk && !net->profile) {
    net->last_raw = raw_dev;
    net->last_raw_images = n_images;
    net->t0_stale = true;
    rc = run_layers_exec(net, n_images, st, raw_dev);
  } else {
    CUDA_TRY(net, tf2b::launch_raw224_to_s2d(raw_dev, net->tbuf[0], n_images, 1, st));
    net->last_launches++;
    net->t0_stale = false;
    rc = run_layers_exec(net, n_images, st);
  }
  if (rc) return rc;
  return write_result(net, net->result_tensor, n_images, out_dev, out_layout, st);
}

int tf2b_run_raw224_host(tf2b_net* net, const int8_t* raw_host, int n_images, int8_t* out_host, int out_layout) {
  int rc = check_run(net, n_images);
  if (rc) return rc;
  if (!raw_host || !out_host) return fail(net, TF2B_ERR_ARG, "null host pointer");
  cudaStream_t st = net->own_stream;
  CUDA_TRY(net, cudaSetDevice(net->device));
  const tf2b_tensor_desc& tr = net->tensors[net->result_tensor];
  CUDA_TRY(net, cudaMemcpyAsync(net->io_in, raw_host, (size_t)n_images * 3 * 224 * 224, cudaMemcpyHostToDevice, st));
  rc = tf2b_run_raw224(net, net->io_in, n_images, net->io_out, out_layout, st);
  if (rc) return rc;
  CUDA_TRY(net, cudaMemcpyAsync(out_host, net->io_out, (size_t)n_images * tr.C * tr.H * tr.W, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(net, cudaStreamSynchronize(st));
  return TF2B_OK;
}

// Pipelined form of the host-buffer calls: the reference's host also only *enqueues* its finite
// kernels (Runner::EnqueueKernels, runner.cpp:32) and waits later (WaitForAllKernels).  Two slots:
// while batch i computes, the H2D copy of batch i+1 and the D2H copy of batch i-1 run on their own
// copy engines.  tf2b_wait(slot) returns when that slot's result is in out_host.  All staging buffers,
// streams and events exist since tf2b_finalize (ensure_io): nothing is allocated here.
static int submit_common(tf2b_net* net, const int8_t* in_host, bool raw224, int in_layout, int n_images,
                         int8_t* out_host, int out_layout, int slot) {
  int rc = check_run(net, n_images);
  if (rc) return rc;
  if (!in_host || !out_host || slot < 0 || slot > 1) return fail(net, TF2B_ERR_ARG, "bad pointer or slot");
  CUDA_TRY(net, cudaSetDevice(net->device));
  const tf2b_tensor_desc& t0 = net->tensors[0];
  const tf2b_tensor_desc& tr = net->tensors[net->result_tensor];
  const size_t in_bytes = (size_t)n_images * (raw224 ? (size_t)3 * 224 * 224 : (size_t)t0.C * t0.H * t0.W);
  cudaStream_t sc = net->own_stream;
  // the slot's input staging buffer is free once the previous batch that used it has been consumed
  if (net->slot_used[slot]) CUDA_TRY(net, cudaStreamWaitEvent(net->s_h2d, net->ev_comp[slot], 0));
  CUDA_TRY(net, cudaMemcpyAsync(net->slot_in[slot], in_host, in_bytes, cudaMemcpyHostToDevice, net->s_h2d));
  CUDA_TRY(net, cudaEventRecord(net->ev_h2d[slot], net->s_h2d));
  CUDA_TRY(net, cudaStreamWaitEvent(sc, net->ev_h2d[slot], 0));
  if (net->slot_used[slot]) CUDA_TRY(net, cudaStreamWaitEvent(sc, net->ev_d2h[slot], 0));  // slot_out still being read back
  rc = raw224 ? tf2b_run_raw224(net, net->slot_in[slot], n_images, net->slot_out[slot], out_layout, sc)
              : tf2b_run(net, net->slot_in[slot], in_layout, n_images, net->slot_out[slot], out_layout, sc);
  if (rc) return rc;
  CUDA_TRY(net, cudaEventRecord(net->ev_comp[slot], sc));
  CUDA_TRY(net, cudaStreamWaitEvent(net->s_d2h, net->ev_comp[slot], 0));
  CUDA_TRY(net, cudaMemcpyAsync(out_host, net->slot_out[slot], (size_t)n_images * tr.C * tr.H * tr.W,
                                cudaMemcpyDeviceToHost, net->s_d2h));
  CUDA_TRY(net, cudaEventRecord(net->ev_d2h[slot], net->s_d2h));
  net->slot_used[slot] = true;
  return TF2B_OK;
}

int tf2b_submit_raw224_host(tf2b_net* net, const int8_t* raw_host, int n_images, int8_t* out_host, int out_layout,
                            int slot) {
  return submit_common(net, raw_host, true, TF2B_LAYOUT_CHW, n_images, out_host, out_layout, slot);
}

int tf2b_submit_host(tf2b_net* net, const int8_t* in_host, int in_layout, int n_images, int8_t* out_host,
                     int out_layout, int slot) {
  return submit_common(net, in_host, false, in_layout, n_images, out_host, out_layout, slot);
}

int tf2b_wait(tf2b_net* net, int slot) {
  if (!net || slot < 0 || slot > 1) return TF2B_ERR_ARG;
  if (!net->slot_used[slot]) return TF2B_OK;
  CUDA_TRY(net, cudaSetDevice(net->device));
  CUDA_TRY(net, cudaEventSynchronize(net->ev_d2h[slot]));
  return TF2B_OK;
}

int tf2b_run_host(tf2b_net* net, const int8_t* in_host, int in_layout, int n_images, int8_t* out_host,
                  int out_layout) {
  int rc = check_run(net, n_images);
  if (rc) return rc;
  if (!in_host || !out_host) return fail(net, TF2B_ERR_ARG, "null host pointer");
  cudaStream_t st = net->own_stream;
  CUDA_TRY(net, cudaSetDevice(net->device));
  const tf2b_tensor_desc& t0 = net->tensors[0];
  const tf2b_tensor_desc& tr = net->tensors[net->result_tensor];
  CUDA_TRY(net, cudaMemcpyAsync(net->io_in, in_host, (size_t)n_images * t0.C * t0.H * t0.W, cudaMemcpyHostToDevice, st));
  rc = tf2b_run(net, net->io_in, in_layout, n_images, net->io_out, out_layout, st);
  if (rc) return rc;
  CUDA_TRY(net, cudaMemcpyAsync(out_host, net->io_out, (size_t)n_images * tr.C * tr.H * tr.W, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(net, cudaStreamSynchronize(st));
  return TF2B_OK;
}

int tf2b_read_tensor(tf2b_net* net, int tensor, int n_images, int8_t* dst_dev, int layout, void* stream) {
  int rc = check_run(net, n_images);
  if (rc) return rc;
  if (tensor < 0 || tensor >= (int)net->tensors.size() || !dst_dev) return fail(net, TF2B_ERR_ARG, "bad tensor/pointer");
  CUDA_TRY(net, cudaSetDevice(net->device));
  if (tensor == 0) {
    int rc0 = refresh_t0(net, (cudaStream_t)stream);
    if (rc0) return rc0;
  }
  return write_result(net, tensor, n_images, dst_dev, layout, (cudaStream_t)stream);
}

int tf2b_dump_acc(tf2b_net* net, int layer, int n_images, int32_t* acc_dev, void* stream) {
  int rc = check_run(net, n_images);
  if (rc) return rc;
  if (layer < 0 || layer >= (int)net->layers.size() || !acc_dev) return fail(net, TF2B_ERR_ARG, "bad layer/pointer");
  CUDA_TRY(net, cudaSetDevice(net->device));
  // the conv of this layer is recomputed from its (still resident) input tensor into scratch, so the
  // feature maps of the last run are left untouched
  LayerState& S = net->layers[layer];
  if (S.d.ipool) return fail(net, TF2B_ERR_ARG, "ipool layer has no accumulators");
  if (S.d.in_tensor == 0) {
    int rc0 = refresh_t0(net, (cudaStream_t)stream);
    if (rc0) return rc0;
  }
  size_t need = (size_t)n_images * S.d.OH * S.d.OW * round_up(S.d.N, 16);
  if (need > net->scratch_bytes) {
    // layers that normally write straight to their tensor may exceed the scratch: grow it
    CUDA_TRY(net, cudaDeviceSynchronize());
    CUDA_TRY(net, cudaFree(net->scratch0));
    CUDA_TRY(net, cudaFree(net->scratch1));
    net->scratch0 = net->scratch1 = nullptr;
    drop_graphs(net);   // captured graphs hold the old scratch addresses
    net->scratch_bytes = need;
    CUDA_TRY(net, cudaMalloc(&net->scratch0, need));
    CUDA_TRY(net, cudaMalloc(&net->scratch1, need));
    int rcb = build_tmaps(net);   // layers that write to the scratch carry its address in their tensor maps
    if (rcb != TF2B_OK) return rcb;
  }
  return run_layers(net, n_images, (cudaStream_t)stream, layer, acc_dev);
}

int tf2b_export_weight_blob(tf2b_net* net, void* dev_dst, void* stream) {
  if (!net || !dev_dst) return TF2B_ERR_ARG;
  if (!net->finalized) return fail(net, TF2B_ERR_STATE, "finalize first");
  CUDA_TRY(net, cudaSetDevice(net->device));
  const size_t hb = blob_header_bytes(net);
  std::vector<unsigned char> hdr(hb, 0);
  uint64_t magic = 0x54463242424c4f42ull, nl = net->layers.size(), th = tables_hash(net), ab = net->arena_bytes;
  memcpy(hdr.data(), &magic, 8);
  memcpy(hdr.data() + 8, &nl, 8);
  memcpy(hdr.data() + 16, &th, 8);
  memcpy(hdr.data() + 24, &ab, 8);
  for (size_t l = 0; l < net->layers.size(); l++) {
    const LayerState& S = net->layers[l];
    BlobLayerMeta m;
    memset(&m, 0, sizeof m);
    m.loaded = S.loaded; m.Cp = S.Cp; m.Cp_m = S.Cp_m; m.Npad_s = S.Npad_s; m.Kp_s = S.Kp_s; m.Cp_s = S.Cp_s; m.nseg_s = S.nseg_s;
    m.Npad_m = S.Npad_m; m.Kp_m = S.Kp_m; m.planes_m = S.planes_m; m.mma_ok = S.planes_m > 0; m.Npar = S.Npar;
    for (int i = 0; i < 4; i++) m.plane_shift_m[i] = S.plane_shift_m[i];
    for (int i = 0; i < 8; i++) {
      m.seg_shift_s[i] = S.seg_shift_s[i]; m.seg_neg_s[i] = S.seg_neg_s[i];
      m.seg_cbeg_s[i] = S.seg_cbeg_s[i]; m.seg_cend_s[i] = S.seg_cend_s[i];
    }
    m.off_kmask = S.off_kmask; m.kmask_len = (int64_t)S.h_kmask.size();
    m.off_w8p = S.off_w8p; m.w8p_len = (int64_t)S.h_w8p.size();
    m.off_w4 = S.off_w4; m.off_w8 = S.off_w8; m.off_bias = S.off_bias; m.off_alpha = S.off_alpha;
    m.off_beta = S.off_beta; m.off_nshift = S.off_nshift; m.off_nshift_m = S.off_nshift_m;
    m.low_plane_m = S.low_plane_m; m.nshift_m_len = (int32_t)S.h_nshift_m.size(); m.fast_requant = S.fast_requant;
    m.sparse2_m = S.sparse2_m; m.blkmask_len = (int32_t)std::min(S.h_blkmask.size(), sizeof m.blkmask);
    memcpy(m.blkmask, S.h_blkmask.data(), (size_t)m.blkmask_len);
    memcpy(hdr.data() + kBlobFixed + l * sizeof m, &m, sizeof m);
  }
  for (size_t t = 0; t < net->tensors.size(); t++) {
    BlobTensorMeta tm;
    memset(&tm, 0, sizeof tm);
    tm.has_order = net->tpos[t].empty() ? 0 : 1;
    tm.channels = (int32_t)net->tpos[t].size();
    tm.off_pos = (int64_t)net->off_tpos[t];
    tm.off_inv = (int64_t)net->off_tinv[t];
    memcpy(hdr.data() + kBlobFixed + net->layers.size() * sizeof(BlobLayerMeta) + t * sizeof tm, &tm, sizeof tm);
  }
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(net, cudaMemcpyAsync(dev_dst, hdr.data(), hb, cudaMemcpyHostToDevice, st));
  CUDA_TRY(net, cudaMemcpyAsync((unsigned char*)dev_dst + hb, net->arena, net->arena_bytes, cudaMemcpyDeviceToDevice, st));
  CUDA_TRY(net, cudaStreamSynchronize(st));
  return TF2B_OK;
}

// Import = the receiving side of the init-time NCCL broadcast: engine created from the same layer
// tables, no model file read; call instead of tf2b_load_layer, then tf2b_finalize.
int tf2b_import_weight_blob(tf2b_net* net, const void* dev_src, int64_t blob_bytes, void* stream) {
  if (!net || !dev_src) return TF2B_ERR_ARG;
  if (net->finalized) return fail(net, TF2B_ERR_STATE, "import must precede tf2b_finalize");
  CUDA_TRY(net, cudaSetDevice(net->device));
  const size_t hb = blob_header_bytes(net);
  if (blob_bytes < (int64_t)hb) return fail(net, TF2B_ERR_ARG, "weight blob of %lld bytes is shorter than its header", (long long)blob_bytes);
  std::vector<unsigned char> hdr(hb);
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(net, cudaMemcpyAsync(hdr.data(), dev_src, hb, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(net, cudaStreamSynchronize(st));
  uint64_t magic, nl, th, ab;
  memcpy(&magic, hdr.data(), 8);
  memcpy(&nl, hdr.data() + 8, 8);
  memcpy(&th, hdr.data() + 16, 8);
  memcpy(&ab, hdr.data() + 24, 8);
  if (magic != 0x54463242424c4f42ull || nl != net->layers.size() || th != tables_hash(net))
    return fail(net, TF2B_ERR_ARG, "weight blob does not match this network (magic / layer count / table hash)");
  if ((uint64_t)blob_bytes != hb + ab)
    return fail(net, TF2B_ERR_ARG, "weight blob is %lld bytes, its header says %llu", (long long)blob_bytes,
                (unsigned long long)(hb + ab));
  for (size_t l = 0; l < net->layers.size(); l++) {
    LayerState& S = net->layers[l];
    BlobLayerMeta m;
    memcpy(&m, hdr.data() + kBlobFixed + l * sizeof m, sizeof m);
    if (S.d.ipool) continue;
    if (!m.loaded) return fail(net, TF2B_ERR_ARG, "blob layer %zu has no weights", l);
    S.loaded = true; S.Cp = m.Cp; S.Cp_m = m.Cp_m; S.Npad_s = m.Npad_s; S.Kp_s = m.Kp_s; S.Cp_s = m.Cp_s; S.nseg_s = m.nseg_s;
    S.Npad_m = m.Npad_m; S.Kp_m = m.Kp_m; S.planes_m = m.planes_m; S.mma_ok = m.mma_ok != 0; S.Npar = m.Npar;
    if (S.nseg_s < 1 || S.nseg_s > 8 || S.planes_m < 0 || S.planes_m > tf2b::kMaxPlanes)
      return fail(net, TF2B_ERR_ARG, "blob layer %zu: segment / plane counts out of range", l);
    for (int i = 0; i < 4; i++) S.plane_shift_m[i] = m.plane_shift_m[i];
    for (int i = 0; i < 8; i++) {
      S.seg_shift_s[i] = m.seg_shift_s[i]; S.seg_neg_s[i] = m.seg_neg_s[i];
      S.seg_cbeg_s[i] = m.seg_cbeg_s[i]; S.seg_cend_s[i] = m.seg_cend_s[i];
      if (i < S.nseg_s && (S.seg_cbeg_s[i] < 0 || S.seg_cend_s[i] <= S.seg_cbeg_s[i] || S.seg_cend_s[i] > 4096))
        return fail(net, TF2B_ERR_ARG, "blob layer %zu: segment range out of bounds", l);
    }
    // pull the arrays back to the host so finalize() can lay out and upload them uniformly
    auto pull = [&](auto& vec, size_t count, int64_t off) -> cudaError_t {
      // every array must lie inside the arena part of the blob the caller handed over
      if (off < 0 || count > (size_t)1 << 40 || (uint64_t)off + count * sizeof(vec[0]) > ab) return cudaErrorInvalidValue;
      vec.resize(count);
      if (!count) return cudaSuccess;
      return cudaMemcpy(vec.data(), (const unsigned char*)dev_src + hb + off, count * sizeof(vec[0]), cudaMemcpyDeviceToHost);
    };
    CUDA_TRY(net, pull(S.h_w4, (size_t)S.nseg_s * S.Npad_s * (S.Kp_s / 2), m.off_w4));
    if (m.kmask_len < 0 || m.kmask_len > (1 << 20)) return fail(net, TF2B_ERR_ARG, "blob layer %zu: bad step mask", l);
    CUDA_TRY(net, pull(S.h_kmask, (size_t)m.kmask_len, m.off_kmask));
    CUDA_TRY(net, pull(S.h_w8, (size_t)S.planes_m * S.Npad_m * S.Kp_m, m.off_w8));
    if (m.w8p_len != 0 && m.w8p_len != (int64_t)S.h_w8.size() / 2) return fail(net, TF2B_ERR_ARG, "blob layer %zu: bad packed planes", l);
    CUDA_TRY(net, pull(S.h_w8p, (size_t)m.w8p_len, m.off_w8p));
    CUDA_TRY(net, pull(S.h_bias, (size_t)S.Npar, m.off_bias));
    CUDA_TRY(net, pull(S.h_alpha, (size_t)S.Npar, m.off_alpha));
    CUDA_TRY(net, pull(S.h_beta, (size_t)S.Npar, m.off_beta));
    CUDA_TRY(net, pull(S.h_nshift, (size_t)round_up(S.Npar, 16), m.off_nshift));
    CUDA_TRY(net, pull(S.h_nshift_m, (size_t)m.nshift_m_len, m.off_nshift_m));
    S.low_plane_m = m.low_plane_m;
    S.fast_requant = m.fast_requant;
    S.sparse2_m = m.sparse2_m;
    if (m.blkmask_len < 0 || m.blkmask_len > (int32_t)sizeof m.blkmask) return fail(net, TF2B_ERR_ARG, "blob layer %zu: bad block mask", l);
    S.h_blkmask.assign(m.blkmask, m.blkmask + m.blkmask_len);
    S.prepared = true;
  }
  net->tpos.assign(net->tensors.size(), {});
  for (size_t t = 0; t < net->tensors.size(); t++) {
    BlobTensorMeta tm;
    memcpy(&tm, hdr.data() + kBlobFixed + net->layers.size() * sizeof(BlobLayerMeta) + t * sizeof tm, sizeof tm);
    if (!tm.has_order) continue;
    if (tm.channels != net->tensors[t].C || tm.off_pos < 0 || (uint64_t)tm.off_pos + (uint64_t)tm.channels * 4 > ab)
      return fail(net, TF2B_ERR_ARG, "blob tensor %zu: channel order out of range", t);
    net->tpos[t].resize(tm.channels);
    CUDA_TRY(net, cudaMemcpy(net->tpos[t].data(), (const unsigned char*)dev_src + hb + tm.off_pos, (size_t)tm.channels * 4,
                             cudaMemcpyDeviceToHost));
    std::vector<char> seen(tm.channels, 0);
    for (int c = 0; c < tm.channels; c++) {
      const int ppos = net->tpos[t][c];
      if (ppos < 0 || ppos >= tm.channels || seen[ppos]) return fail(net, TF2B_ERR_ARG, "blob tensor %zu: not a permutation", t);
      seen[ppos] = 1;
    }
  }
  return TF2B_OK;
}

int tf2b_set_profile(tf2b_net* net, int on) {
  if (!net) return TF2B_ERR_ARG;
  CUDA_TRY(net, cudaSetDevice(net->device));
  if (on && net->ev.empty()) {
    net->ev.resize(3 * net->layers.size());
    for (auto& e : net->ev) CUDA_TRY(net, cudaEventCreate(&e));
  }
  net->profile = on != 0;
  return TF2B_OK;
}

int tf2b_get_profile(tf2b_net* net, float* conv_ms, float* layer_ms, int n_layers) {
  if (!net || !conv_ms || !layer_ms) return TF2B_ERR_ARG;
  if (n_layers != (int)net->layers.size()) return fail(net, TF2B_ERR_ARG, "n_layers mismatch");
  if (net->ev.empty()) return fail(net, TF2B_ERR_STATE, "profiling was never enabled");
  CUDA_TRY(net, cudaSetDevice(net->device));
  CUDA_TRY(net, cudaEventSynchronize(net->ev.back()));
  for (int l = 0; l < n_layers; l++) {
    CUDA_TRY(net, cudaEventElapsedTime(&conv_ms[l], net->ev[3 * l], net->ev[3 * l + 1]));
    CUDA_TRY(net, cudaEventElapsedTime(&layer_ms[l], net->ev[3 * l], net->ev[3 * l + 2]));
  }
  return TF2B_OK;
}

int tf2b_last_launches(tf2b_net* net) { return net ? net->last_launches : 0; }

const char* tf2b_layer_kernel(tf2b_net* net, int layer) {
  if (!net || layer < 0 || layer >= (int)net->layers.size()) return "none";
  const LayerState& S = net->layers[layer];
  if (S.d.ipool) return "none";
  return (S.kernel == 2 && S.mma_ok) ? "mma" : "shift";
}

const char* tf2b_layer_mode(tf2b_net* net, int layer, int n_images) {
  if (!net || !net->finalized || layer < 0 || layer >= (int)net->layers.size()) return "none";
  LayerState& S = net->layers[layer];
  if (S.d.ipool) return "pool";
  const tf2b_layer_desc& d = S.d;
  if (!(S.kernel == 2 && S.mma_ok)) {
    ConvParams ps = conv_params(net, S, n_images > 0 ? n_images : net->max_images, net->scratch0, round_up(d.N, 16), nullptr, 0, false);
    S.mode_desc = tf2b::sa_describe(ps, S.nseg_s, S.ksplit_s);
    return S.mode_desc.c_str();
  }
  const bool to_scratch = d.pool || d.gap;
  int8_t* dst = to_scratch ? net->scratch0 : net->tbuf[d.out_tensor] + d.out_ch0;
  const int dstC = to_scratch ? round_up(d.N, 16) : net->tpitch[d.out_tensor];
  const int8_t* res = (d.add_tensor >= 0 && !d.pool) ? net->tbuf[d.add_tensor] : nullptr;
  const int resC = (d.add_tensor >= 0 && !d.pool) ? net->tpitch[d.add_tensor] : 0;
  ConvParams p = conv_params(net, S, n_images > 0 ? n_images : net->max_images, dst, dstC, res, resC, true);
  S.mode_desc = tf2b::mma_describe(p, S.planes_m);
  return S.mode_desc.c_str();
}

void tf2b_destroy(tf2b_net* net) {
  if (!net) return;
  cudaSetDevice(net->device);
  drop_graphs(net);
  for (auto p : net->tbuf) if (p) cudaFree(p);
  if (net->arena) cudaFree(net->arena);
  if (net->scratch0) cudaFree(net->scratch0);
  if (net->scratch1) cudaFree(net->scratch1);
  if (net->stem_t0) cudaFree(net->stem_t0);
  if (net->stem_conv) cudaFree(net->stem_conv);
  if (net->io_in) cudaFree(net->io_in);
  if (net->io_out) cudaFree(net->io_out);
  if (net->own_stream) cudaStreamDestroy(net->own_stream);
  for (int i = 0; i < 2; i++) {
    if (net->slot_in[i]) cudaFree(net->slot_in[i]);
    if (net->slot_out[i]) cudaFree(net->slot_out[i]);
    if (net->ev_h2d[i]) cudaEventDestroy(net->ev_h2d[i]);
    if (net->ev_comp[i]) cudaEventDestroy(net->ev_comp[i]);
    if (net->ev_d2h[i]) cudaEventDestroy(net->ev_d2h[i]);
  }
  if (net->s_h2d) cudaStreamDestroy(net->s_h2d);
  if (net->s_d2h) cudaStreamDestroy(net->s_d2h);
  for (auto e : net->ev) cudaEventDestroy(e);
  for (auto e : net->ev_layer) if (e) cudaEventDestroy(e);
  if (net->ev_fork) cudaEventDestroy(net->ev_fork);
  for (int i = 1; i < tf2b_net::kLanes; i++) {
    if (net->ev_join[i]) cudaEventDestroy(net->ev_join[i]);
    if (net->side[i]) cudaStreamDestroy(net->side[i]);
  }
  delete net;
}

}  // extern "C"
