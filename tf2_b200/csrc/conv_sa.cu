// conv_sa.cu — kernel A: CUDA-core shift-accumulate convolution, exact for every layer shape.
//
// Reference semantics: PeFunction's MAC  result += sum_c MUL(feature, code)  with  MUL = (+-feature) << shift
// (Runtime_Engine/cnn/device/src/pe.cl:27-49,144-180), zero padding / stride geometry of sequencer.cl:268-311
// and retriever.cl:134-213, then the requantisation of pe.cl:185-203, ReLU (relu.cl:54) and the residual add
// of feature_writer.cl:124-127 in the epilogue.  INT32 accumulator tap = the value pe.cl:196-199 prints.
//
// B200 mapping (implicit GEMM, M = pixels, N = output channels, K = taps x channels):
//   * weights stay PACKED in HBM / L2 as 4-bit codes (sign + 3-bit exponent, 7 = zero — the information content
//     of TransForm_Kit's 4-bit format, 4bit_data_format.txt:1-44): a shift by s is a multiplication by 2^s
//     modulo 2^32, s = base[n] + 7*segment + e, so a weight is +-2^e inside its segment;
//   * per K chunk one thread stages the int8 activation tile (TMA box of the NHWC tensor: the hardware zero fill
//     is the padding, the traversal stride is the convolution stride) and the packed weight tile (TMA box of the
//     nibble matrix) into an mbarrier ring that runs ahead across segments and tiles (persistent CTAs);
//   * the eight warps expand the nibbles of the chunk ON THE FLY into an int8 tile in shared memory
//     (PRMT look-ups, ~14 instructions per 8 weights, once per CTA and chunk = 5 % of the MAC instructions) and
//     run the MACs as 4-way int8 dot products (IDP.4A) on 8 x 8 register tiles fed by 16-byte shared loads;
//   * a segment's int32 sum wraps exactly like the FPGA accumulator; sums are combined as
//     total += sum << 7*segment, then bias + (total << base[n]) — all modulo 2^32, i.e. bit-exact;
//   * the int8 negate quirk (pe.cl:32-34: -(-128) = -128) is a segment that multiplies the byte-negated
//     activations, or — for tensor 0, which carries a negated copy of its channels — plain extra channels;
//   * tiles whose m-range is small (7x7 maps, the fc layer) split K over the two half-warps and combine the
//     partial sums with warp shuffles.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>

#include "common.cuh"

namespace tf2b {

int sa_kc(int Cp);

namespace {

constexpr int SA_BM = 128;            // pixels per CTA tile (64 with the K split)
constexpr int SA_CONSUMERS = 256;     // 8 warps, all of them compute
constexpr int SA_THREADS = SA_CONSUMERS;
constexpr int SA_STAGES = 4;
constexpr int SA_MAXSEG = 8;
constexpr int SA_BSTRIDE = 80;        // bytes per row of the expanded weight tile (64 data + 16 pad)

struct SaParams {
  ConvParams c;
  int BN;                   // 64 or 128 output channels per tile
  int KC;                   // channels per K chunk: 32 or 64
  int mode;                 // 0 = flat (1x1, stride 1, pad 0), 1 = box
  int tw, th, tn, tiles_w, tiles_h, tiles_b, m_tiles, n_tiles;
  int taps, cchunks;        // k*k, channel chunks per tap
  int nseg;
  int seg_shift[SA_MAXSEG];
  int seg_neg[SA_MAXSEG];
  int seg_cbeg[SA_MAXSEG];  // channel chunks [cbeg, cend) of every tap the segment covers: with the channels of a tensor
  int seg_cend[SA_MAXSEG];  // ordered by their power-of-two offset, a segment only spans the chunks of its offset class
  int Npad;                 // rows per segment in the packed weight matrix
  int a_bytes, b_bytes;     // bytes one stage receives
  int ksplit;               // 1: the two half-warps take alternate 16-channel steps, partial sums via shuffles
  const unsigned char* kmask;   // [nseg][taps * cchunks]: which 16-channel steps of a (segment, tap, chunk) hold weights;
                                // 0 = the whole chunk is skipped by the producer and the MACs alike
};

struct SaMaps {
  CUtensorMap a;   // activations
  CUtensorMap b;   // packed 4-bit weights [nseg * Npad][Kp / 2]
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(unsigned smem, const CUtensorMap* map, unsigned bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(unsigned smem, const CUtensorMap* map, unsigned bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void bar_consumers() { __syncthreads(); }

__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel) {
  unsigned d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// Eight 4-bit codes (code i in bits 4i..4i+3: bit 3 = negative, bits 0..2 = exponent e, 7 = zero, never
// "negative zero") -> eight int8 weights +-2^e: lo = weights 0..3, hi = weights 4..7.
//   magnitude  PRMT look-up in the byte table {1,2,4,8 | 16,32,64,0} by the exponent
//   sign mask  PRMT in sign-replicate mode over (w, w << 4): byte msb = bit 3 of each code
//   negate     (m ^ mask) + (mask & 0x01..): -m = ~m + 1 per byte; m >= 1 wherever the mask is set, no carry
__device__ __forceinline__ void expand8(unsigned w, unsigned& lo, unsigned& hi) {
  const unsigned T0 = 0x08040201u, T1 = 0x00402010u;
  const unsigned w4 = w << 4;
  const unsigned mlo = prmt(T0, T1, w & 0x7777u);
  const unsigned mhi = prmt(T0, T1, (w >> 16) & 0x7777u);
  const unsigned slo = prmt(w, w4, 0x9D8Cu);
  const unsigned shi = prmt(w, w4, 0xBFAEu);
  lo = (mlo ^ slo) + (slo & 0x01010101u);
  hi = (mhi ^ shi) + (shi & 0x01010101u);
}

struct SaTile {
  int n0, m0, b0, oh0, ow0;
};
__device__ __forceinline__ SaTile sa_decode(const SaParams& P, int tile, int BM) {
  SaTile t;
  const int mt = tile / P.n_tiles;
  t.n0 = (tile - mt * P.n_tiles) * P.BN;
  t.m0 = mt * BM;
  t.b0 = t.oh0 = t.ow0 = 0;
  if (P.mode == 1) {
    const int r = mt / P.tiles_w;
    t.ow0 = (mt - r * P.tiles_w) * P.tw;
    const int bt = r / P.tiles_h;
    t.oh0 = (r - bt * P.tiles_h) * P.th;
    t.b0 = bt * P.tn;
  }
  return t;
}

// TN = output channels per thread (BN = 16 * TN).  KSPLIT (layers with few tiles): the tile is 64 pixels, the two
// half-warps (lane bit 4) take alternate 16-channel steps of every chunk and their partial sums meet in a
// warp-shuffle reduction; otherwise the tile is 128 pixels and lane bit 4 is a second pixel group.
template <int TN, bool KSPLIT, int KC>
__global__ void __launch_bounds__(SA_THREADS, 2)
conv_sa_kernel(const __grid_constant__ SaParams P, const __grid_constant__ SaMaps maps) {
  constexpr int BN = 16 * TN;
  constexpr int BM = KSPLIT ? 64 : SA_BM;
  constexpr int TM = 8;                                    // pixels per thread
  constexpr int TYN = BM / TM;                             // row distance between a thread's pixels
  extern __shared__ __align__(128) unsigned char smem[];
  // carve: [stages][A tile | packed B tile]  [2][expanded B tile, reused as the int8 output staging tile]
  constexpr int a_stage = BM * KC;
  constexpr int b_stage = BN * (KC / 2);
  constexpr int stage_bytes = a_stage + b_stage;
  unsigned char* const bx_base = smem + SA_STAGES * stage_bytes;
  constexpr int BX_BYTES = BN * SA_BSTRIDE;
  __shared__ __align__(8) unsigned long long bars[SA_STAGES];
  __shared__ unsigned row_lut[SA_BM];
  const unsigned full_bar = smem_u32(&bars[0]);
  const int t = threadIdx.x;
  const int lane = t & 31;
  const ConvParams& c = P.c;
  if (t == 0) {
    for (int s = 0; s < SA_STAGES; s++) mbar_init(full_bar + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (P.mode == 1 && t < BM) {
    const int wl = t % P.tw;
    const int r = t / P.tw;
    const int hl = r % P.th;
    const int nl = r / P.th;
    row_lut[t] = (unsigned)wl | ((unsigned)hl << 8) | ((unsigned)nl << 16) | ((nl < P.tn ? 1u : 0u) << 24);
  }
  __syncthreads();
  const int num_tiles = P.m_tiles * P.n_tiles;

  // ---- producer cursor (thread 0): the TMA loads run SA_STAGES - 1 chunks ahead of the MACs, across segment and
  //      tile boundaries (persistent CTAs: the ring never drains).  No "empty" barriers: chunk i + STAGES - 1 goes
  //      into the stage of chunk i - 1, which every warp has left once it is past the CTA barrier of chunk i.
  int p_tile = blockIdx.x, p_g = 0, p_tap = 0, p_cc = P.seg_cbeg[0], p_stage = 0;
  SaTile p_tc = sa_decode(P, p_tile < num_tiles ? p_tile : 0, BM);
  const int mask_row = P.taps * P.cchunks;
  auto p_advance = [&]() {   // next (segment, tap, chunk) of the cursor, wrapping into the next tile
    if (++p_cc == P.seg_cend[p_g]) {
      if (++p_tap == P.taps) {
        p_tap = 0;
        if (++p_g == P.nseg) {
          p_g = 0;
          p_tile += gridDim.x;
          if (p_tile < num_tiles) p_tc = sa_decode(P, p_tile, BM);
        }
      }
      p_cc = P.seg_cbeg[p_g];
    }
  };
  auto produce = [&]() {
    // chunks without weights are never staged (the consumers skip them by the same table)
    while (p_tile < num_tiles && P.kmask[p_g * mask_row + p_tap * P.cchunks + p_cc] == 0) p_advance();
    if (p_tile >= num_tiles) return;
    const int fh = p_tap / c.k, fw = p_tap - fh * c.k;
    const unsigned fb = full_bar + 8 * p_stage;
    const unsigned sa = smem_u32(smem) + p_stage * stage_bytes;
    mbar_expect_tx(fb, (unsigned)(P.a_bytes + P.b_bytes));
    if (P.mode == 0) tma_load_2d(sa, &maps.a, fb, p_cc * KC, p_tc.m0);
    else tma_load_4d(sa, &maps.a, fb, p_cc * KC, p_tc.ow0 * c.stride - c.pad + fw, p_tc.oh0 * c.stride - c.pad + fh, p_tc.b0);
    tma_load_2d(sa + a_stage, &maps.b, fb, (p_tap * P.cchunks + p_cc) * (KC / 2), p_g * P.Npad + p_tc.n0);
    if (++p_stage == SA_STAGES) p_stage = 0;
    p_advance();
  };
  if (t == 0)
    for (int i = 0; i < SA_STAGES - 1; i++) produce();

  // ---- thread -> (ty pixel group, tx channel group)
  const int tx = t & 15;
  const int kh = KSPLIT ? ((t >> 4) & 1) : 0;
  const int ty = KSPLIT ? (t >> 5) : (t >> 4);
  const int M = c.B * c.OH * c.OW;
  int stage = 0;
  unsigned phase = 0;
  int bxsel = 0;
  constexpr int ksteps = KC / 16;
  for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const SaTile tc = sa_decode(P, tile, BM);
    // ONE accumulator set: segments come in descending shift order and are combined by Horner's rule,
    // acc = (acc << (shift[g-1] - shift[g])) + sum_g  — exact modulo 2^32 like the FPGA's wrapping accumulator
    int acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; i++)
#pragma unroll
      for (int j = 0; j < TN; j++) acc[i][j] = 0;
    for (int g = 0; g < P.nseg; g++) {
      const bool neg = P.seg_neg[g] != 0;
      const unsigned char* mrow = P.kmask + g * mask_row;
      for (int tap = 0; tap < P.taps; tap++)
      for (int cc = P.seg_cbeg[g]; cc < P.seg_cend[g]; cc++) {
        const unsigned kmask = mrow[tap * P.cchunks + cc];
        if (kmask == 0) continue;   // (warp-uniform, CTA-uniform: the producer skipped it too)
        mbar_wait(full_bar + 8 * stage, phase);
        const unsigned char* As = smem + stage * stage_bytes;
        const unsigned char* Bp = As + a_stage;
        unsigned char* Bx = bx_base + bxsel * BX_BYTES;
        // ---- expand the packed weight tile of this chunk: 16 packed bytes -> 32 int8 weights per item
        {
          constexpr int per_row = KC / 32;
          constexpr int items = BN * per_row;
#pragma unroll
          for (int e0 = 0; e0 < items; e0 += SA_CONSUMERS) {
            const int e = e0 + t;
            if (items % SA_CONSUMERS == 0 || e < items) {
              const int row = e / per_row, part = e % per_row;
              const uint4 pk = *reinterpret_cast<const uint4*>(Bp + row * (KC / 2) + part * 16);
              uint4 o0, o1;
              expand8(pk.x, o0.x, o0.y);
              expand8(pk.y, o0.z, o0.w);
              expand8(pk.z, o1.x, o1.y);
              expand8(pk.w, o1.z, o1.w);
              uint4* dst = reinterpret_cast<uint4*>(Bx + row * SA_BSTRIDE + part * 32);
              dst[0] = o0;
              dst[1] = o1;
            }
          }
        }
        bar_consumers();
        if (t == 0) produce();   // refills the stage of the previous chunk
        // ---- MACs: 16 channels per step as 4 x IDP.4A per (pixel, channel) pair
#pragma unroll
        for (int ks = kh; ks < ksteps; ks += (KSPLIT ? 2 : 1)) {
          if (!((kmask >> ks) & 1u)) continue;   // no weight of this segment in these 16 channels
          uint4 a[TM];
#pragma unroll
          for (int i = 0; i < TM; i++) {
            a[i] = *reinterpret_cast<const uint4*>(As + (ty + TYN * i) * KC + ks * 16);
            if (neg) {   // pe.cl:32-34: negate inside int8 (wraps: -(-128) = -128)
              a[i].x = __vneg4(a[i].x); a[i].y = __vneg4(a[i].y); a[i].z = __vneg4(a[i].z); a[i].w = __vneg4(a[i].w);
            }
          }
#pragma unroll
          for (int j = 0; j < TN; j++) {
            const uint4 b = *reinterpret_cast<const uint4*>(Bx + (tx + 16 * j) * SA_BSTRIDE + ks * 16);
#pragma unroll
            for (int i = 0; i < TM; i++) {
              int s = acc[i][j];
              s = __dp4a((int)a[i].x, (int)b.x, s);
              s = __dp4a((int)a[i].y, (int)b.y, s);
              s = __dp4a((int)a[i].z, (int)b.z, s);
              s = __dp4a((int)a[i].w, (int)b.w, s);
              acc[i][j] = s;
            }
          }
        }
        // (the expanded tile is double buffered: the barrier of the NEXT chunk orders its reuse)
        if (++stage == SA_STAGES) { stage = 0; phase ^= 1; }
        bxsel ^= 1;
      }
      const int sh = P.seg_shift[g] - (g + 1 < P.nseg ? P.seg_shift[g + 1] : 0);
      if (sh != 0) {
#pragma unroll
        for (int i = 0; i < TM; i++)
#pragma unroll
          for (int j = 0; j < TN; j++) acc[i][j] = (int)((unsigned)acc[i][j] << sh);
      }
    }
    if (KSPLIT) {
      // warp-shuffle reduction of the two K halves (lanes l and l ^ 16 hold partial sums of the same outputs)
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], 16);
    }
    // ---- epilogue: bias seed + base shift, requantisation, ReLU -> int8 staging tile (reuses the expanded-weight
    //      buffers: every warp is past its last MAC of this tile after the barrier)
    bar_consumers();
    unsigned char* Cs = bx_base;
    constexpr int CS_STRIDE = BN + 16;
    static_assert(BM * CS_STRIDE <= 2 * BX_BYTES, "staging tile must fit the expanded-weight buffers");
    // output pixel (flat index b*OH*OW + oh*OW + ow, < 2^31) of a tile row, or -1
    auto pixel_of = [&](int row) -> int {
      if (P.mode == 0) return (tc.m0 + row < M) ? tc.m0 + row : -1;
      const unsigned lu = row_lut[row];
      const int ow = tc.ow0 + (int)(lu & 0xff), oh = tc.oh0 + (int)((lu >> 8) & 0xff), b = tc.b0 + (int)((lu >> 16) & 0xff);
      if (!(lu >> 24) || ow >= c.OW || oh >= c.OH || b >= c.B) return -1;
      return (b * c.OH + oh) * c.OW + ow;
    };
    if (!KSPLIT || kh == 0) {
      // INT32 accumulator tap (pe.cl:196-199 prints this value): [image][N][OH][OW]
      long long dump_off[TM];
      if (c.acc_dump != nullptr) {
        const int hw = c.OH * c.OW;
#pragma unroll
        for (int i = 0; i < TM; i++) {
          const int pix = pixel_of(ty + TYN * i);
          const int b = pix >= 0 ? pix / hw : 0;
          dump_off[i] = pix >= 0 ? (long long)b * c.N * hw + (pix - b * hw) : -1ll;   // + n * hw below
        }
      }
#pragma unroll
      for (int j = 0; j < TN; j++) {
        const int n = tc.n0 + tx + 16 * j;   // < Npad always
        const int bias = c.bias[n], alpha = c.alpha[n], beta = c.beta[n];
        const int nsh = c.nshift[n];
#pragma unroll
        for (int i = 0; i < TM; i++) {
          const int row = ty + TYN * i;
          const int a32 = (int)((unsigned)bias + ((unsigned)acc[i][j] << nsh));
          if (c.acc_dump != nullptr && n < c.N && dump_off[i] >= 0)
            c.acc_dump[(size_t)dump_off[i] + (size_t)(c.acc_perm ? c.acc_perm[n] : n) * (size_t)(c.OH * c.OW)] = a32;
          int y = requant(a32, alpha, beta);
          if (c.relu) y = max(y, 0);
          Cs[row * CS_STRIDE + tx + 16 * j] = (unsigned char)(signed char)y;
        }
      }
    }
    bar_consumers();
    // ---- coalesced store (+ residual add): 16-byte chunks, consecutive threads along the channels
    for (int chk = t; chk < BM * (BN / 16); chk += SA_CONSUMERS) {
      const int row = chk / (BN / 16), seg = chk - row * (BN / 16);
      const int nb = tc.n0 + seg * 16;
      if (nb >= c.N) continue;
      const long long pix = pixel_of(row);
      if (pix < 0) continue;
      uint4 v = *reinterpret_cast<const uint4*>(&Cs[row * CS_STRIDE + seg * 16]);
      int8_t* dst = c.y + pix * c.yC + nb;
      const int nvalid = min(16, c.N - nb);
      if (nvalid == 16) {
        if (c.r != nullptr) {
          const uint4 rv = *reinterpret_cast<const uint4*>(c.r + pix * c.rC + nb);
          v.x = add_res4(v.x, rv.x, c.add_relu);
          v.y = add_res4(v.y, rv.y, c.add_relu);
          v.z = add_res4(v.z, rv.z, c.add_relu);
          v.w = add_res4(v.w, rv.w, c.add_relu);
        }
        *reinterpret_cast<uint4*>(dst) = v;
      } else {
        const unsigned char* vb = &Cs[row * CS_STRIDE + seg * 16];
        for (int e = 0; e < nvalid; e++) {
          int yv = (signed char)vb[e];
          if (c.r != nullptr) {
            int s2 = yv + (int)c.r[pix * c.rC + nb + e];
            s2 = max(-128, min(127, s2));
            if (c.add_relu) s2 = max(s2, 0);
            yv = s2;
          }
          dst[e] = (int8_t)yv;
        }
      }
    }
    bar_consumers();   // the staging tile is the next tile's expanded-weight buffer
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn sa_encode_fn(std::string* err) {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
    if (err) *err = "cuTensorMapEncodeTiled entry point not available";
    return nullptr;
  }
  fn = (EncodeTiledFn)p;
  return fn;
}

void sa_geometry(SaParams& P, const ConvParams& c, int nseg, const int* seg_shift, const int* seg_neg, int ksplit,
                 const int* seg_cbeg = nullptr, const int* seg_cend = nullptr, const unsigned char* kmask = nullptr) {
  memset(&P, 0, sizeof P);
  P.c = c;
  P.kmask = kmask;
  P.ksplit = ksplit;
  const int BM = ksplit ? 64 : SA_BM;
  P.BN = c.N > 64 ? 128 : 64;
  P.KC = sa_kc(c.Cp);
  P.mode = (c.k == 1 && c.stride == 1 && c.pad == 0) ? 0 : 1;
  P.taps = c.k * c.k;
  P.cchunks = (c.Cp + P.KC - 1) / P.KC;
  P.nseg = nseg;
  for (int i = 0; i < SA_MAXSEG; i++) {
    P.seg_shift[i] = (i < nseg && seg_shift) ? seg_shift[i] : 0;
    P.seg_neg[i] = (i < nseg && seg_neg) ? seg_neg[i] : 0;
    P.seg_cbeg[i] = (i < nseg && seg_cbeg) ? seg_cbeg[i] : 0;
    P.seg_cend[i] = (i < nseg && seg_cend) ? seg_cend[i] : P.cchunks;
  }
  P.Npad = c.Npad;
  P.n_tiles = (c.N + P.BN - 1) / P.BN;
  if (P.mode == 0) {
    const long long M = (long long)c.B * c.OH * c.OW;
    P.m_tiles = (int)((M + BM - 1) / BM);
    P.a_bytes = BM * P.KC;
  } else {
    // (tw, th, tn) box of output pixels: whole rows, then whole images when the map is small; must not depend
    // on the batch size of a run (the tensor map is built once for max_images)
    P.tw = c.OW < BM ? c.OW : BM;
    P.th = BM / P.tw;
    if (P.th > c.OH) P.th = c.OH;
    const int th_tiles = (c.OH + P.th - 1) / P.th;
    P.th = (c.OH + th_tiles - 1) / th_tiles;
    P.tn = (P.th == c.OH) ? BM / (P.tw * P.th) : 1;
    if (P.tn < 1) P.tn = 1;
    P.tiles_w = (c.OW + P.tw - 1) / P.tw;
    P.tiles_h = (c.OH + P.th - 1) / P.th;
    P.tiles_b = (c.B + P.tn - 1) / P.tn;
    P.m_tiles = P.tiles_w * P.tiles_h * P.tiles_b;
    P.a_bytes = P.tw * P.th * P.tn * P.KC;
  }
  P.b_bytes = P.BN * (P.KC / 2);
}

}  // namespace

int sa_kc(int Cp) { return Cp > 32 ? 64 : 32; }
int sa_npad(int N) { return (N + 127) / 128 * 128; }
int sa_max_segments() { return SA_MAXSEG; }
size_t sa_tmap_bytes() { return sizeof(SaMaps); }

// K split (decided once, for max_images: the tensor maps depend on it): the layer has fewer 128-pixel tiles than
// CTA slots and its chunks have at least two 16-channel steps per half-warp
int sa_ksplit(const ConvParams& c, int num_sms) {
  SaParams P;
  sa_geometry(P, c, 1, nullptr, nullptr, 0);
  return (P.m_tiles * P.n_tiles < num_sms && P.KC == 64 && c.OW <= 64) ? 1 : 0;
}

std::string sa_describe(const ConvParams& c, int nseg, int ksplit) {
  SaParams P;
  sa_geometry(P, c, nseg, nullptr, nullptr, ksplit);
  char b[160];
  snprintf(b, sizeof b, "shift BM%d BN%d KC%d segs%d %s packed4 dp4a%s", ksplit ? 64 : SA_BM, P.BN, P.KC, nseg,
           P.mode == 0 ? "flat" : "box", ksplit ? " ksplit-shfl" : "");
  return std::string(b);
}

int sa_build_tmaps(void* host_tmaps, const ConvParams& c, const uint8_t* wgt4, int nseg, int ksplit, std::string* err) {
  EncodeTiledFn enc = sa_encode_fn(err);
  if (!enc) return -1;
  SaParams P;
  sa_geometry(P, c, nseg, nullptr, nullptr, ksplit);
  const int BM = ksplit ? 64 : SA_BM;
  SaMaps* tp = reinterpret_cast<SaMaps*>(host_tmaps);
  CUresult r;
  if (P.mode == 0) {
    cuuint64_t dims[2] = {(cuuint64_t)c.xC, (cuuint64_t)c.B * c.IH * c.IW};
    cuuint64_t strides[1] = {(cuuint64_t)c.xC};
    cuuint32_t box[2] = {(cuuint32_t)P.KC, (cuuint32_t)BM};
    cuuint32_t es[2] = {1, 1};
    r = enc(&tp->a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)c.x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)c.xC, (cuuint64_t)c.IW, (cuuint64_t)c.IH, (cuuint64_t)c.B};
    cuuint64_t strides[3] = {(cuuint64_t)c.xC, (cuuint64_t)c.xC * c.IW, (cuuint64_t)c.xC * c.IW * c.IH};
    const cuuint32_t st = (cuuint32_t)c.stride;
    cuuint32_t box[4] = {(cuuint32_t)P.KC, (cuuint32_t)((P.tw - 1) * st + 1), (cuuint32_t)((P.th - 1) * st + 1), (cuuint32_t)P.tn};
    cuuint32_t es[4] = {1, st, st, 1};
    r = enc(&tp->a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void*)c.x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled(A, shift kernel) failed with CUresult " + std::to_string((int)r);
    return -1;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)(c.Kp / 2), (cuuint64_t)nseg * c.Npad};
    cuuint64_t strides[1] = {(cuuint64_t)(c.Kp / 2)};
    cuuint32_t box[2] = {(cuuint32_t)(P.KC / 2), (cuuint32_t)P.BN};
    cuuint32_t es[2] = {1, 1};
    r = enc(&tp->b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)wgt4, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled(packed weights) failed with CUresult " + std::to_string((int)r);
    return -1;
  }
  return 0;
}

namespace {
template <int TN, bool KS, int KC>
cudaError_t sa_set_attr() {
  return cudaFuncSetAttribute(conv_sa_kernel<TN, KS, KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
}
template <int TN, bool KS, int KC>
void sa_launch(int grid, size_t smem, cudaStream_t stream, const SaParams& P, const SaMaps& m) {
  conv_sa_kernel<TN, KS, KC><<<grid, SA_THREADS, smem, stream>>>(P, m);
}
}  // namespace

// per-device: opt in to > 48 KB dynamic shared memory (called from tf2b_finalize after cudaSetDevice)
cudaError_t sa_prepare_device() {
  cudaError_t e;
  if ((e = sa_set_attr<4, false, 64>()) != cudaSuccess) return e;
  if ((e = sa_set_attr<8, false, 64>()) != cudaSuccess) return e;
  if ((e = sa_set_attr<4, true, 64>()) != cudaSuccess) return e;
  if ((e = sa_set_attr<8, true, 64>()) != cudaSuccess) return e;
  if ((e = sa_set_attr<4, false, 32>()) != cudaSuccess) return e;
  if ((e = sa_set_attr<8, false, 32>()) != cudaSuccess) return e;
  return cudaSuccess;
}

cudaError_t launch_conv_sa(const ConvParams& c, int nseg, const int* seg_shift, const int* seg_neg, const int* seg_cbeg,
                           const int* seg_cend, const unsigned char* kmask_dev, const void* tmaps, int ksplit, int num_sms,
                           cudaStream_t stream) {
  SaParams P;
  sa_geometry(P, c, nseg, seg_shift, seg_neg, ksplit, seg_cbeg, seg_cend, kmask_dev);
  const int num_tiles = P.m_tiles * P.n_tiles;
  const int BM = ksplit ? 64 : SA_BM;
  const int stage_bytes = BM * P.KC + P.BN * (P.KC / 2);
  const size_t smem = (size_t)SA_STAGES * stage_bytes + 2 * (size_t)P.BN * SA_BSTRIDE;
  int grid = std::min(num_tiles, 2 * num_sms);
  if (grid < 1) grid = 1;
  const SaMaps* tp = reinterpret_cast<const SaMaps*>(tmaps);
  if (P.KC == 64) {
    if (P.BN == 64) {
      if (ksplit) sa_launch<4, true, 64>(grid, smem, stream, P, *tp);
      else sa_launch<4, false, 64>(grid, smem, stream, P, *tp);
    } else {
      if (ksplit) sa_launch<8, true, 64>(grid, smem, stream, P, *tp);
      else sa_launch<8, false, 64>(grid, smem, stream, P, *tp);
    }
  } else {
    if (P.BN == 64) sa_launch<4, false, 32>(grid, smem, stream, P, *tp);
    else sa_launch<8, false, 32>(grid, smem, stream, P, *tp);
  }
  return cudaGetLastError();
}

}  // namespace tf2b
