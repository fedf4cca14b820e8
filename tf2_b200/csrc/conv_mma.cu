// conv_mma.cu — kernel B: INT8 x INT4->INT8 implicit-GEMM convolution on the 5th-gen tensor cores.
//
// Reference semantics are those of kernel A (pe.cl:27-49,144-203; relu.cl:54; feature_writer.cl:
// 124-127).  The power-of-two weights make the shift-accumulate a true integer GEMM: a code with
// shift s contributes feature * (+-2^s) mod 2^32.  Per output channel the smallest shift is factored
// out (applied to the finished sum, exact mod 2^32); the remaining exponent e = s - base is split
// into planes of 7 levels so every weight is an int8 +-2^(e mod 7).  tcgen05.mma kind::i8 multiplies
// the int8 activation tile with each plane into its own int32 TMEM accumulator (|sum| <= 127*64*K
// < 2^31, so nothing saturates or wraps inside the tensor core) and the epilogue recombines
// sum_p acc_p << 7p, adds the bias seed and requantises exactly like pe.cl:185-203.
//
// Structure (one persistent CTA per SM, 18 warps, warp specialised):
//   warp 0       TMA producer: the activation tile into an mbarrier ring, in one of the staging modes
//                  flat      1x1/stride 1: the NHWC tensor as a [pixels x BK] matrix
//                  box       one (BK, tw, th, tn) box per filter tap, hardware zero fill = the padding,
//                            traversal stride = the convolution stride
//                  halo      stride-1 kxk with resident weights: ONE box per tile (the input window in
//                            raster order); the taps are row-shifted UMMA descriptor views of it
//                  pixelpair 64-byte pixels, two horizontal taps per 128-byte row
//                + one weight tile per plane (or the whole slab of the n-tile once: resident weights)
//   warp 1       MMA issuer: tcgen05.mma kind::i8, M=128 (cta_group::1) or M=256 over a CTA pair
//                (cta_group::2: each CTA stages its activation rows and half of the weight tile),
//                N = planes*BN <= 256, K=32 per instruction, accumulators double-buffered in TMEM
//   warps 2..17  epilogue: tcgen05.ld -> plane recombination -> requantisation (folded: one IMAD.HI per
//                output) -> packed int8 ReLU / residual (s16x2 lanes) -> 32-byte row stores
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>

#include "../../include/tf2b200.h"
#include "common.cuh"

namespace tf2b {

// host copies of a layer's per-channel parameters (the launcher builds the constant-bank table from them)
struct MmaHostParams {
  const int32_t* bias;
  const int32_t* alpha;
  const int32_t* beta;
  const uint8_t* nshift;
};

namespace {

// Experiment switches (TF2B_MMA_DEBUG role counters, TF2B_MMA_NOEPI, TF2B_MMA_POLL0, TF2B_MMA_TOP,
// TF2B_MMA_L2PF) cost a few uniform branches per tile in the hot loops; they are compiled in only
// with -DTF2B_EXPERIMENTS (TF2B_EXPERIMENTS=1 tf2_b200/csrc/build.sh).
#ifdef TF2B_EXPERIMENTS
constexpr bool kExp = true;
#else
constexpr bool kExp = false;
#endif

// Mode switches (TF2B_MMA_HALO, _FOLD, _HI32, _CG2, _PAIR, _BRES, _RESTMA, _RBUFS, _BN256, _DIRECT, _PDL,
// _STAGES) are read from the environment only in experiment builds; the shipped library has no getenv.
inline int env_int(const char* name, int dflt) {
#ifdef TF2B_EXPERIMENTS
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
#else
  (void)name;
  return dflt;
#endif
}

constexpr int MMA_M = 128;
constexpr int NUM_EPI_WARPS = 16;
constexpr int NUM_THREADS = 32 * (2 + NUM_EPI_WARPS);
constexpr int MAX_STAGES = 8;
constexpr int MAX_RBUFS = 6;   // residual tile ring (flat layers with a TMA-fed residual operand)
constexpr int TMEM_COLS = 512;
constexpr int EPI_ROW = 48;   // bytes per staging row: 32 data + 16 pad (conflict-free 16-byte accesses)
constexpr int PSTR = 64;      // per-channel parameter arrays of one epilogue warp: [6][PSTR] int32
constexpr int EPI_WARP_BYTES = 32 * EPI_ROW + 6 * PSTR * 4;   // staging tile + params of one epilogue warp
constexpr int EPI_BYTES = NUM_EPI_WARPS * EPI_WARP_BYTES;

// x / d for 0 <= x, d < 2^20:  (x * ceil(2^40/d)) >> 40   (exact in that range)
struct FastDiv {
  unsigned long long mul;
  int d;
};
__host__ __device__ __forceinline__ FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = d < 1 ? 1 : d;
  f.mul = ((1ull << 40) + (unsigned long long)f.d - 1) / (unsigned long long)f.d;
  return f;
}
__device__ __forceinline__ int fdiv(int x, const FastDiv& f) { return (int)(((unsigned long long)(unsigned)x * f.mul) >> 40); }

struct MmaParams {
  ConvParams c;
  int mode;                 // 0 = flat (1x1, stride 1, pad 0), 1 = box (one TMA box per filter tap)
  int tw, th, tn;           // box tile: output columns, rows, images (tw*th*tn <= 128)
  int tiles_w, tiles_h, tiles_b;
  int m_tiles, n_tiles;
  int BK, BN, planes, stages;
  int kchunks;              // channel chunks per tap = Cpm / BK
  int Cpm;                  // input channels rounded up to BK (weight K layout: tap*Cpm + c)
  int taps;                 // k*k
  int a_bytes, b_bytes;     // bytes one TMA load delivers (A box, one B plane tile)
  int plane8_shift[kMaxPlanes];
  int Npad;                 // rows per weight plane
  unsigned idesc;           // tcgen05 instruction descriptor
  unsigned sbo16;           // stride byte offset >> 4 of the smem descriptors
  unsigned layout_type;     // UMMA smem layout type (2 = SWIZZLE_128B, 4 = SWIZZLE_64B)
  long long* dbg;           // optional per-CTA cycle counters [grid][8] (TF2B_MMA_DEBUG), else nullptr
  FastDiv d_ntiles, d_tiles_w, d_tiles_h;
  int direct256;            // output/residual rows are 32-byte aligned: row-per-lane 32-byte accesses
  int pair;                 // "pixel pair" rows: 64-byte pixels, pad 0: one 128-byte TMA row = 2 adjacent
                            // pixels = 2 horizontal taps (taps = filter rows, kchunks = ceil(k/2))
  int noepi;                // experiment switch: epilogue warps only hand the accumulators back (no math, no stores)
  int poll_lane0;           // experiment switch: one lane polls mbarriers (else all lanes)
  int roles_top;            // experiment switch: producer/MMA warps at the highest warp ids
  int l2_prefetch;          // tiles ahead whose activation boxes are prefetched into L2 (0 = off)
  int res_tma;              // residual tiles arrive through TMA into a smem ring (flat layers, BN >= 128)
  int res_bufs;             // ring depth (2..MAX_RBUFS)
  int b_resident;           // the CTA's weight slab (all taps/chunks/planes of its n-tile) stays in smem
  int res_bytes;            // bytes of that slab
  int halo;                 // halo mode (box tiles, stride 1, resident weights): ONE TMA box per tile and
                            // channel chunk holds the (th + k - 1) x Wp input pixels of the tile in raster
                            // order; every filter tap is a row-shifted view of it (UMMA descriptor start
                            // address + (fh * Wp + fw) * BK: the swizzle is a function of the address bits)
  int Wp;                   // halo mode: row width of the position space = tw + k - 1
  int hstream;              // halo tiles with STREAMED weights (CTA pairs, K-heavy k x k / stride-1 layers): the halo
                            // tile of a channel chunk sits in its own ring (a_bufs x a_ring_bytes) and is fetched once
                            // per tile and chunk instead of once per tap; the pipeline stages carry weights only
  int a_bufs, a_ring_bytes; // halo-tile ring of that mode
  int lean_roles;           // producer / MMA roles run their lean loops (always in product builds; experiment builds
                            // keep the generic loops for the role counters and switches)
  int a_stage_bytes;        // bytes reserved per pipeline stage for the activation tile
  int idx32;                // output / residual tensors are smaller than 2 GiB: 32-bit byte offsets
  int cg2;                  // CTA-pair mode: clusters of two CTAs, tcgen05.mma.cta_group::2 (M = 256 over two
                            // SMs, each CTA stages its own 128 activation rows and HALF of the weight tile)
  int b_stage_bytes;        // bytes of weights one CTA stages per k-iteration
  int egroups;              // epilogue warp groups (1: all 16 warps share every tile; 2: 8 warps per tile,
                            // the groups take alternate tiles = alternate TMEM buffers)
  int b_packed4;            // resident weights arrive as PACKED 4-bit codes (TMA) and are expanded on the fly, once per CTA,
                            // into the swizzled K-major int8 slab the MMAs read (4bit_data_format.txt's information content)
  int sparse2;              // two planes issued separately (N = BN each), empty (tap, block, plane) combinations skipped
  unsigned p1mul;           // 2^shift of the second plane (128 with the plain 7-level planes)
  unsigned idesc1;          // instruction descriptor of a one-plane MMA (N = BN)
  int n_tile0;              // first n-tile of this launch (channel window of the parameter table); n_tiles counts
                            // the n-tiles of the launch
  int tstore;               // flat layers, folded epilogue: finished int8 rows leave through per-warp TMA stores
};

struct TmapPair {
  CUtensorMap a;
  CUtensorMap b;
  CUtensorMap r;   // residual operand (flat layers): [pixels][channels], 128 x 128-byte boxes
  CUtensorMap h;   // CTA pairs: the half weight tile one CTA of the pair stages
  CUtensorMap y;   // output of flat layers with the folded epilogue: [pixels][channels], per-warp 32-row boxes (TMA store)
  CUtensorMap p;   // packed 4-bit weight planes [planes * Npad][Kp / 2] (resident-weight layers)
};


// ---------------------------------------------------------------- PTX wrappers
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
// bounded wait: a descriptor/protocol bug must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
// converged-warp wait whose loop branch is a vote: provably uniform, so the compiler keeps the surrounding loop state
// (ring addresses, descriptors) in uniform registers
__device__ __forceinline__ void mbar_wait_u(unsigned bar, unsigned parity) {
  if (__any_sync(0xffffffffu, mbar_try_wait(bar, parity))) return;
  long long t0 = clock64();
  while (!__any_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
// same, accumulating the cycles spent waiting (experiment builds, TF2B_MMA_DEBUG)
__device__ __forceinline__ void mbar_wait_ut(unsigned bar, unsigned parity, long long& acc, bool on) {
  if (!on) { mbar_wait_u(bar, parity); return; }
  long long t0 = clock64();
  mbar_wait_u(bar, parity);
  acc += clock64() - t0;
}
// warp-collective wait: lane 0 polls, __syncwarp releases the others
__device__ __forceinline__ void mbar_wait_warp(unsigned bar, unsigned parity, int lane0_only) {
  if (lane0_only) {
    if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
    __syncwarp();
  } else {
    mbar_wait(bar, parity);
  }
}
// same, accumulating the cycles spent waiting (role breakdown for TF2B_MMA_DEBUG)
__device__ __forceinline__ void mbar_wait_timed(unsigned bar, unsigned parity, long long& acc, bool on, int lane0_only) {
  if (!on) { mbar_wait_warp(bar, parity, lane0_only); return; }
  long long t0 = clock64();
  mbar_wait_warp(bar, parity, lane0_only);
  acc += clock64() - t0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

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
// L2 prefetch of a tile (no smem, no barrier): hides the DRAM latency of streamed activations
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2),
               "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

// TMA store of a finished shared-memory tile (bulk async-group of the issuing thread)
__device__ __forceinline__ void tma_store_2d(unsigned smem, const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void tmem_alloc(unsigned smem_dst, unsigned ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(unsigned taddr, unsigned ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], int8 x int8 -> int32
__device__ __forceinline__ void umma_i8(unsigned tmem_d, unsigned long long desc_a, unsigned long long desc_b,
                                        unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, unsigned (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(unsigned taddr, unsigned (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- CTA-pair (cta_group::2) forms.  Shared-memory window addresses carry the CTA rank of the pair in
// bit 24 (cute/arch/copy_sm100_tma.hpp Sm100MmaPeerBitMask): clearing it names the leader's copy.
constexpr unsigned kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the leader CTA's copy of a barrier (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(unsigned bar) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar)
      : "memory");
}
// TMA loads issued by either CTA of the pair into its OWN shared memory, completing on the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_cg2(unsigned smem, const CUtensorMap* map, unsigned bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_cg2(unsigned smem, const CUtensorMap* map, unsigned bar, int c0, int c1,
                                                int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem), "l"(map), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_cg2(unsigned smem_dst, unsigned ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_cg2(unsigned taddr, unsigned ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// completion of all prior MMAs of the pair -> the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_cg2(unsigned bar) {
  asm volatile(
      "{\n\t.reg .b16 m;\n\tmov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
      ::"r"(bar)
      : "memory");
}
// D[tmem of both CTAs] (+)= A[smem of both] * B[smem halves of both], M = 256
__device__ __forceinline__ void umma_i8_cg2(unsigned tmem_d, unsigned long long desc_a, unsigned long long desc_b,
                                            unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// One elected lane of a converged warp (the form CUTLASS uses: keeps the surrounding code
// warp-uniform so descriptors live in uniform registers and UTCIMMA/UTMALDG issue back to back).
__device__ __forceinline__ bool elect_one() {
  unsigned pred = 0;
  unsigned laneid = 0;
  asm volatile(
      "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\t"
      "elect.sync %%rx|%%px, %2;\n\t"
      "@%%px mov.s32 %1, 1;\n\t"
      "mov.s32 %0, %%rx;\n\t}"
      : "+r"(laneid), "+r"(pred)
      : "r"(0xFFFFFFFF));
  return pred != 0;
}

// K-major smem matrix descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start address and
// byte offsets in 16-byte units, version 1 (Blackwell), swizzle mode in bits 61..63.
__device__ __forceinline__ unsigned long long make_smem_desc(unsigned saddr, unsigned sbo16, unsigned layout_type) {
  unsigned long long d = 0;
  d |= (unsigned long long)((saddr >> 4) & 0x3FFF);
  d |= (unsigned long long)1 << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= (unsigned long long)(sbo16 & 0x3FFF) << 32;        // stride between 8-row groups
  d |= (unsigned long long)1 << 46;                       // descriptor version
  d |= (unsigned long long)(layout_type & 7) << 61;
  return d;
}

struct TileCoord {
  int n0;            // first output channel
  int m0;            // flat: first pixel
  int b0, oh0, ow0;  // box: first image / row / column
};

__device__ __forceinline__ TileCoord decode_tile(const MmaParams& P, int tile) {
  TileCoord t;
  const int mt = fdiv(tile, P.d_ntiles);
  const int nt = tile - mt * P.n_tiles;
  t.n0 = (P.n_tile0 + nt) * P.BN;
  t.m0 = mt * MMA_M;
  t.b0 = t.oh0 = t.ow0 = 0;
  if (P.mode == 1) {
    const int r = fdiv(mt, P.d_tiles_w);
    const int wt = mt - r * P.tiles_w;
    const int bt = fdiv(r, P.d_tiles_h);
    const int ht = r - bt * P.tiles_h;
    t.ow0 = wt * P.tw;
    t.oh0 = ht * P.th;
    t.b0 = bt * P.tn;
  }
  return t;
}

// 32-byte global accesses (LDG/STG.E.ENL2.256): one whole sector per lane and instruction
__device__ __forceinline__ void ldg256(const void* p, uint4& lo, uint4& hi) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint4& lo, const uint4& hi) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w),
               "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
               : "memory");
}

// Eight 4-bit codes (code i in bits 4i..4i+3: bit 3 = negative, bits 0..2 = exponent e, 7 = zero) -> eight int8
// weights +-2^e (lo = weights 0..3, hi = 4..7): PRMT look-up of the magnitude, PRMT sign-replicate for the mask,
// -m = ~m + 1 per byte (m >= 1 wherever the mask is set: no carry between bytes)
__device__ __forceinline__ void expand8(unsigned w, unsigned& lo, unsigned& hi) {
  const unsigned T0 = 0x08040201u, T1 = 0x00402010u;
  const unsigned w4 = w << 4;
  unsigned mlo, mhi, slo, shi;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(mlo) : "r"(T0), "r"(T1), "r"(w & 0x7777u));
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(mhi) : "r"(T0), "r"(T1), "r"((w >> 16) & 0x7777u));
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(slo) : "r"(w), "r"(w4), "r"(0x9D8Cu));
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(shi) : "r"(w), "r"(w4), "r"(0xBFAEu));
  lo = (mlo ^ slo) + (slo & 0x01010101u);
  hi = (mhi ^ shi) + (shi & 0x01010101u);
}

// saturating pack of four int32 into int8x4 (y0 in byte 0): cvt.pack.sat clamps to [-128,127]
__device__ __forceinline__ unsigned pack_sat4(int y0, int y1, int y2, int y3) {
  unsigned hi, r;
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(y3), "r"(y2), "r"(0));
  asm("cvt.pack.sat.s8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(y1), "r"(y0), "r"(hi));
  return r;
}

// pe.cl:185-203 without the final clamp (done by the saturating pack)
__device__ __forceinline__ int requant_raw(int32_t acc, int32_t alpha, int32_t beta) {
  long long t = (long long)acc * (long long)alpha;
  int a = (int)(t >> 20);
  int s = (int)((unsigned)a + (unsigned)beta);
  return ((s >> 14) + 1) >> 1;
}

// relu.cl:54 on four packed int8: PRMT replicates each byte's sign bit into a byte mask
__device__ __forceinline__ unsigned relu_s8x4(unsigned v) {
  unsigned m;
  asm("prmt.b32 %0, %1, %1, 0xba98;" : "=r"(m) : "r"(v));
  return v & ~m;
}

// feature_writer.cl:124-127 on four packed int8: y + r in int16 lanes (VIADD.16x2), clamp to int8
// (with ReLU: VIMNMX.S16x2.RELU clamps to [0,127] in one instruction), repack
template <bool RELU>
__device__ __forceinline__ unsigned add_res_s8x4(unsigned y4, unsigned r4) {
  unsigned ylo, yhi, rlo, rhi, slo, shi, out;
  asm("prmt.b32 %0, %1, %1, 0x9180;" : "=r"(ylo) : "r"(y4));   // bytes 0,1 sign-extended to s16x2
  asm("prmt.b32 %0, %1, %1, 0xb3a2;" : "=r"(yhi) : "r"(y4));   // bytes 2,3
  asm("prmt.b32 %0, %1, %1, 0x9180;" : "=r"(rlo) : "r"(r4));
  asm("prmt.b32 %0, %1, %1, 0xb3a2;" : "=r"(rhi) : "r"(r4));
  asm("add.s16x2 %0, %1, %2;" : "=r"(slo) : "r"(ylo), "r"(rlo));
  asm("add.s16x2 %0, %1, %2;" : "=r"(shi) : "r"(yhi), "r"(rhi));
  if (RELU) {
    asm("min.relu.s16x2 %0, %1, %2;" : "=r"(slo) : "r"(slo), "r"(0x007f007fu));
    asm("min.relu.s16x2 %0, %1, %2;" : "=r"(shi) : "r"(shi), "r"(0x007f007fu));
  } else {
    asm("min.s16x2 %0, %1, %2;" : "=r"(slo) : "r"(slo), "r"(0x007f007fu));
    asm("min.s16x2 %0, %1, %2;" : "=r"(shi) : "r"(shi), "r"(0x007f007fu));
    asm("max.s16x2 %0, %1, %2;" : "=r"(slo) : "r"(slo), "r"(0xff80ff80u));
    asm("max.s16x2 %0, %1, %2;" : "=r"(shi) : "r"(shi), "r"(0xff80ff80u));
  }
  asm("prmt.b32 %0, %1, %2, 0x6420;" : "=r"(out) : "r"(slo), "r"(shi));
  return out;
}

// EPI < 0: exact requantisation, every option decided at run time.  EPI >= 0: fused 64-bit
// requantisation (range-analysed layers) specialised on bit0 = second scaled plane, bit1 = low plane,
// bit2 = residual operand, bit3 = folded form: y = (tot * (alpha << nshift) + (bias*alpha +
// ((beta + 2^14) << 20))) >> 35 with tot = plane0 + (plane1 << 7) — one IMAD.HI per output; bit4 (with
// bit3) = every nshift >= 3, so alpha << (nshift-3) and the addend >> 3 make the high word the result.
template <int BN, int MODE, int EPI, bool CG2 = false, int GRP = 1>
// (18 warps = 5 on one scheduler: 16 384 registers / (5 warps x 32 lanes) caps the kernel at 96 registers per thread)
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_mma_kernel(const __grid_constant__ MmaParams P, const __grid_constant__ TmapPair maps) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // experiment builds: phase time stamps of CTA 0 (entry, prologue done, previous layer done, first operands landed,
  // last MMA issued, last tile stored, teardown) -> dbg[kTlBase ..]
  long long tl_entry = 0;
  if (kExp && P.dbg != nullptr && blockIdx.x == 0) tl_entry = clock64();
#define TF2B_TL(i) do { if (kExp && P.dbg != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0) P.dbg[8 * 148 + 3 * 64 + 8 + (i)] = clock64() - tl_entry; } while (0)
  // carve: [resident weight slab] [stages][A | B planes] (1024-aligned) [epilogue scratch]
  const unsigned smem_res = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const unsigned smem_base = smem_res + (unsigned)P.res_bytes;
  const int a_stage = P.a_stage_bytes;
  const int b_plane = BN * P.BK;
  const int stage_bytes = a_stage + (P.b_resident ? 0 : P.b_stage_bytes);
  // CG2 is a template parameter: a kernel that contains cta_group::2 instructions can only be launched
  // as a cluster of CTA pairs
  constexpr bool cg2 = CG2;
  unsigned cta_rank = 0u;
  if constexpr (CG2) cta_rank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);   // (shuffle: provably warp-uniform)
  const int res_tile = BN * 128;   // one residual tile: 128 rows x BN bytes as BN/128 SWIZZLE_128B sub-tiles
  const unsigned smem_rres = smem_base + (unsigned)(P.stages * stage_bytes + EPI_BYTES);   // residual ring
  const unsigned smem_aring = smem_rres;   // streamed-weight halo mode (box layers have no residual ring): halo-tile ring

  __shared__ __align__(8) unsigned long long bars[2 * MAX_STAGES + 5 + 2 * MAX_RBUFS];
  __shared__ unsigned tmem_base_slot;
  __shared__ unsigned row_lut[MMA_M];   // box mode: row -> (wl | hl<<8 | nl<<16 | inbox<<24)
  const unsigned full_bar = smem_u32(&bars[0]);                  // [stages]
  const unsigned empty_bar = smem_u32(&bars[MAX_STAGES]);        // [stages]
  const unsigned tfull_bar = smem_u32(&bars[2 * MAX_STAGES]);    // [2]
  const unsigned tempty_bar = smem_u32(&bars[2 * MAX_STAGES + 2]);  // [2]
  const unsigned bres_bar = smem_u32(&bars[2 * MAX_STAGES + 4]);    // resident weight slab landed
  const unsigned rfull_bar = smem_u32(&bars[2 * MAX_STAGES + 5]);                // [MAX_RBUFS] residual tile landed
  const unsigned rempty_bar = smem_u32(&bars[2 * MAX_STAGES + 5 + MAX_RBUFS]);   // [MAX_RBUFS] residual tile consumed

  // Warp roles.  The SM's issue arbiter favours higher warp ids, so the two single-issuer warps that
  // feed the tensor pipe sit at the top and are never starved by the ALU-heavy epilogue warps:
  //   hardware warps 0..15 -> epilogue (role ids 2..17), warp 16 -> TMA producer (role 0), warp 17 -> MMA (role 1)
  const int hw_warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // canonical warp index: warp-uniform for the compiler
  const int warp = (kExp && P.roles_top) ? (hw_warp < NUM_EPI_WARPS ? hw_warp + 2 : hw_warp - NUM_EPI_WARPS) : hw_warp;
  const int lane = threadIdx.x & 31;
  const int num_tiles = P.m_tiles * P.n_tiles;
  const int kiters = (P.halo && !P.hstream) ? P.kchunks : P.taps * P.kchunks;   // pipeline stages consumed per tile

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&maps.a);
    tma_prefetch_desc(&maps.b);
    if (cg2) tma_prefetch_desc(&maps.h);
    if (MODE == 0 && P.res_tma) tma_prefetch_desc(&maps.r);
    if (MODE == 0 && P.tstore) tma_prefetch_desc(&maps.y);
    for (int s = 0; s < P.stages; s++) {
      mbar_init(full_bar + 8 * s, cg2 ? 2 : 1);   // pair mode: the producers of both CTAs arrive on the leader's
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(tfull_bar + 8 * b, 1);
      mbar_init(tempty_bar + 8 * b, (cg2 ? 2 : 1) * NUM_EPI_WARPS / P.egroups);   // pair: both CTAs' epilogues
    }
    mbar_init(bres_bar, 1);
    for (int b = 0; b < MAX_RBUFS; b++) {
      // residual ring (flat layers) or halo-tile ring (streamed-weight halo mode: filled by both producers of the
      // pair, released by the MMA warp's commit)
      mbar_init(rfull_bar + 8 * b, P.hstream ? (cg2 ? 2 : 1) : 1);
      mbar_init(rempty_bar + 8 * b, P.hstream ? 1 : NUM_EPI_WARPS / P.egroups);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (cg2) tmem_alloc_cg2(smem_u32(&tmem_base_slot), TMEM_COLS);
    else tmem_alloc(smem_u32(&tmem_base_slot), TMEM_COLS);
  }
  if (MODE == 1 && threadIdx.x >= 64 && threadIdx.x < 64 + MMA_M) {
    const int row = threadIdx.x - 64;
    const int roww = P.halo ? P.Wp : P.tw;   // halo mode: accumulator rows walk the padded raster
    const int wl = row % roww;
    const int r = row / roww;
    const int hl = r % P.th;
    const int nl = r / P.th;
    row_lut[row] = (unsigned)wl | ((unsigned)hl << 8) | ((unsigned)nl << 16) | (((nl < P.tn && wl < P.tw) ? 1u : 0u) << 24);
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (cg2) cluster_sync_all();   // the peer's barriers must be initialised before anything signals them
  tc_fence_after();
  const unsigned tmem_base = tmem_base_slot;
  const int acc_cols = P.planes * BN;  // TMEM columns of one accumulator buffer
  // Work items of this CTA: tiles blockIdx.x, +gridDim.x, ...  Pair mode: work item q = (pair of adjacent
  // m-tiles, n-tile); the pair blockIdx.x/2 walks q, rank r of the pair owns m-tile 2*mp + r.
  const int q_first = cg2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int q_step = cg2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int q_count = cg2 ? ((P.m_tiles + 1) >> 1) * P.n_tiles : num_tiles;
  auto tile_of = [&](int q) -> int {
    if (!cg2) return q;
    const int mp = fdiv(q, P.d_ntiles);
    return (2 * mp + (int)cta_rank) * P.n_tiles + (q - mp * P.n_tiles);
  };
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor
  // prefetch, row LUT) overlaps the tail of the previous layer's kernel; from here on this grid
  // reads what that kernel wrote, so wait for it to complete and flush.  The next layer's grid may
  // be scheduled as soon as SMs free up (it blocks at its own wait).
  // Resident weights from PACKED 4-bit tiles (weights do not depend on the previous layer, so this whole block
  // overlaps that layer's tail under programmatic dependent launch): one thread TMA-loads the CTA's slab as
  // nibbles (bit 3 = negative, bits 0..2 = exponent, 7 = zero) into the still idle pipeline ring, every thread
  // expands 16 codes at a time into the swizzled K-major int8 layout the UMMA descriptors expect, and a proxy
  // fence hands the slab to the tensor core.
  if (P.b_packed4 && (int)blockIdx.x < num_tiles) {
    const int n0p = decode_tile(P, blockIdx.x).n0;
    const int ntile = P.taps * P.kchunks * P.planes;
    const int ptile = b_plane >> 1;   // bytes of one packed tile
    if (warp == 0) {
      if (elect_one()) {
        mbar_expect_tx(bres_bar, (unsigned)(ntile * ptile));
        for (int tap = 0; tap < P.taps; tap++)
          for (int kc = 0; kc < P.kchunks; kc++)
            for (int pl = 0; pl < P.planes; pl++)
              tma_load_2d(smem_base + ((tap * P.kchunks + kc) * P.planes + pl) * ptile, &maps.p, bres_bar,
                          (tap * P.Cpm + kc * P.BK) >> 1, pl * P.Npad + n0p);
      }
      __syncwarp();
    }
    mbar_wait(bres_bar, 0);
    const int cpr = P.BK >> 4;                  // 16-byte output chunks per row
    const int cpt = BN * cpr;                   // ... per tile
    const unsigned char* pk_base = smem_raw + (smem_base - smem_u32(smem_raw));
    unsigned char* out_base = smem_raw + (smem_res - smem_u32(smem_raw));
    for (int idx = threadIdx.x; idx < ntile * cpt; idx += NUM_THREADS) {
      const int ti = idx / cpt, rem = idx - ti * cpt;
      const int row = rem / cpr, j = rem - row * cpr;
      const uint2 pk = *reinterpret_cast<const uint2*>(pk_base + ti * ptile + row * (P.BK >> 1) + j * 8);
      uint4 o;
      expand8(pk.x, o.x, o.y);
      expand8(pk.y, o.z, o.w);
      const int swz = (P.BK == 128) ? (row & 7) : ((row >> 1) & 3);   // SWIZZLE_128B / SWIZZLE_64B of the address bits
      *reinterpret_cast<uint4*>(out_base + ti * b_plane + row * P.BK + ((j ^ swz) << 4)) = o;
    }
    fence_proxy_async();
    __syncthreads();   // slab complete; the ring is free for the activation tiles from here on
  }
  // Weight-stationary layers: every tile of this CTA has the same n-tile (the grid is a multiple of n_tiles), so its
  // whole weight slab is fetched once — and BEFORE the wait on the previous layer: weights do not depend on it, so the
  // fetch overlaps that layer's tail on every SM that CTA has already left.
  if (warp == 0 && P.b_resident && !P.b_packed4 && (int)blockIdx.x < num_tiles) {
    const int n0 = decode_tile(P, blockIdx.x).n0;
    if (elect_one()) {
      mbar_expect_tx(bres_bar, (unsigned)P.res_bytes);
      for (int tap = 0; tap < P.taps; tap++)
        for (int kc = 0; kc < P.kchunks; kc++)
          for (int pl = 0; pl < P.planes; pl++)
            tma_load_2d(smem_res + ((tap * P.kchunks + kc) * P.planes + pl) * b_plane, &maps.b, bres_bar,
                        tap * P.Cpm + kc * P.BK, pl * P.Npad + n0);
    }
    __syncwarp();
  }
  if (threadIdx.x == 32) TF2B_TL(0);
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 32) TF2B_TL(1);

  if (warp == 0) {
    // ===================================================== TMA producer (whole warp, one elected lane issues)
    int stage = 0;
    unsigned phase = 0;
    const bool dbg = kExp && P.dbg != nullptr;
    long long w_empty = 0, t_start = clock64();
    int rb = 0;
    unsigned rphase = 0;
    // Lean form of this role: tile-constant coordinates hoisted, ring addresses advanced by additions, no per-stage
    // division / descriptor rebuild / warp barrier, and waits whose loop branch is a warp vote — with provably uniform
    // control flow the compiler keeps the loop state in uniform registers and issues UTMALDG / UTCIMMA back to back.
    // Measured with the role counters (profiles/r02_mma_roles.md): the generic loop spent 500-800 clk per pipeline
    // stage in its own instructions, more than the 512 clk the stage's four M256 x N256 x K32 MMAs take.
    if (!kExp || P.lean_roles) {
      // (all 32 lanes walk the loop so that every address stays in uniform registers; one elected lane issues)
      const int nst = P.stages, BK = P.BK, kch = P.kchunks, kk = P.c.k, pad = P.c.pad;
      const unsigned a_bytes = (unsigned)P.a_bytes;
      unsigned sa = smem_base, fb = full_bar, eb = empty_bar;
      auto advance = [&]() {
        if (++stage == nst) { stage = 0; phase ^= 1; sa = smem_base; fb = full_bar; eb = empty_bar; }
        else { sa += (unsigned)stage_bytes; fb += 8; eb += 8; }
      };
      for (int q = q_first; q < q_count; q += q_step) {
        const TileCoord t = decode_tile(P, tile_of(q));
        if (MODE == 0 && P.res_tma) {
          // residual operand of this tile: full 128-byte lines, no L1, latency hidden by the run-ahead
          mbar_wait_u(rempty_bar + 8 * rb, rphase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(rfull_bar + 8 * rb, (unsigned)res_tile);
            for (int j = 0; j < BN / 128; j++)
              tma_load_2d(smem_rres + rb * res_tile + j * (128 * 128), &maps.r, rfull_bar + 8 * rb, t.n0 + j * 128, t.m0);
          }
          if (++rb == P.res_bufs) { rb = 0; rphase ^= 1; }
        }
        if constexpr (cg2) {
          const int brow = P.planes == 2 ? (int)cta_rank * P.Npad + t.n0 : t.n0 + (int)cta_rank * (BN / 2);
          if (MODE == 0) {
            const unsigned tx = 2u * (a_bytes + (unsigned)P.b_stage_bytes);
            for (int kc = 0, c0 = 0; kc < kch; kc++, c0 += BK) {
              mbar_wait_ut(eb, phase ^ 1, w_empty, kExp && dbg);
              if (elect_one()) {
                if (cta_rank == 0) mbar_expect_tx(fb, tx);
                else mbar_arrive_leader(fb);
                tma_load_2d_cg2(sa, &maps.a, fb, c0, t.m0);
                tma_load_2d_cg2(sa + a_stage, &maps.h, fb, c0, brow);
              }
              advance();
            }
          } else if (P.hstream) {
            const unsigned txa = 2u * a_bytes, txb = 2u * (unsigned)P.b_stage_bytes;
            const int x0 = t.ow0 - pad, y0 = t.oh0 - pad, ntap = P.taps, Cpm = P.Cpm;
            for (int kc = 0, c0 = 0; kc < kch; kc++, c0 += BK) {
              mbar_wait_u(rempty_bar + 8 * rb, rphase ^ 1);
              if (elect_one()) {
                const unsigned ab = rfull_bar + 8 * rb;
                if (cta_rank == 0) mbar_expect_tx(ab, txa);
                else mbar_arrive_leader(ab);
                tma_load_4d_cg2(smem_aring + rb * P.a_ring_bytes, &maps.a, ab, c0, x0, y0, t.b0);
              }
              if (++rb == P.a_bufs) { rb = 0; rphase ^= 1; }
              for (int tap = 0, kcol = c0; tap < ntap; tap++, kcol += Cpm) {
                mbar_wait_ut(eb, phase ^ 1, w_empty, kExp && dbg);
                if (elect_one()) {
                  if (cta_rank == 0) mbar_expect_tx(fb, txb);
                  else mbar_arrive_leader(fb);
                  tma_load_2d_cg2(sa, &maps.h, fb, kcol, brow);
                }
                advance();
              }
            }
          } else {
            const unsigned tx = 2u * (a_bytes + (unsigned)P.b_stage_bytes);
            const int x0 = t.ow0 * P.c.stride - pad, y0 = t.oh0 * P.c.stride - pad, Cpm = P.Cpm;
            int kcol = 0;
            for (int fh = 0; fh < kk; fh++)
              for (int fw = 0; fw < kk; fw++, kcol += Cpm)
                for (int kc = 0, c0 = 0; kc < kch; kc++, c0 += BK) {
                  mbar_wait_ut(eb, phase ^ 1, w_empty, kExp && dbg);
                  if (elect_one()) {
                    if (cta_rank == 0) mbar_expect_tx(fb, tx);
                    else mbar_arrive_leader(fb);
                    tma_load_4d_cg2(sa, &maps.a, fb, c0, x0 + fw, y0 + fh, t.b0);
                    tma_load_2d_cg2(sa + a_stage, &maps.h, fb, kcol + c0, brow);
                  }
                  advance();
                }
          }
        } else if (P.halo) {
          // halo tiles, resident weights: one box per tile and channel chunk
          const int x0 = t.ow0 - pad, y0 = t.oh0 - pad;
          for (int kc = 0, c0 = 0; kc < kch; kc++, c0 += BK) {
            mbar_wait_ut(eb, phase ^ 1, w_empty, kExp && dbg);
            if (elect_one()) {
              mbar_expect_tx(fb, a_bytes);
              tma_load_4d(sa, &maps.a, fb, c0, x0, y0, t.b0);
            }
            advance();
          }
        } else {
          // single CTA, flat rows or one box per tap; weights resident or streamed with the activations
          const bool bres = P.b_resident != 0;
          const int npl = P.planes, Npad = P.Npad;
          const unsigned tx = a_bytes + (bres ? 0u : (unsigned)(npl * P.b_bytes));
          if (MODE == 0) {
            for (int kc = 0, c0 = 0; kc < kch; kc++, c0 += BK) {
              mbar_wait_ut(eb, phase ^ 1, w_empty, kExp && dbg);
              if (elect_one()) {
                mbar_expect_tx(fb, tx);
                tma_load_2d(sa, &maps.a, fb, c0, t.m0);
                if (!bres)
                  for (int pl = 0; pl < npl; pl++) tma_load_2d(sa + a_stage + pl * b_plane, &maps.b, fb, c0, pl * Npad + t.n0);
              }
              advance();
            }
          } else if (P.pair) {
            // pixel pairs: tap = filter row, chunk kc = horizontal taps 2kc, 2kc+1 = pixels ow+2kc, ow+2kc+1 of one 128-byte row
            const int Cpm = P.Cpm, ntap = P.taps;
            int row = (t.b0 * P.c.IH + t.oh0) * P.c.IW + t.ow0;
            for (int tap = 0, kcol = 0; tap < ntap; tap++, kcol += Cpm, row += P.c.IW)
              for (int kc = 0, c0 = 0; kc < kch; kc++, c0 += BK) {
                mbar_wait_ut(eb, phase ^ 1, w_empty, kExp && dbg);
                if (elect_one()) {
                  mbar_expect_tx(fb, tx);
                  tma_load_2d(sa, &maps.a, fb, 0, row + 2 * kc);
                  if (!bres)
                    for (int pl = 0; pl < npl; pl++)
                      tma_load_2d(sa + a_stage + pl * b_plane, &maps.b, fb, kcol + c0, pl * Npad + t.n0);
                }
                advance();
              }
          } else {
            const int x0 = t.ow0 * P.c.stride - pad, y0 = t.oh0 * P.c.stride - pad, Cpm = P.Cpm;
            int kcol = 0;
            for (int fh = 0; fh < kk; fh++)
              for (int fw = 0; fw < kk; fw++, kcol += Cpm)
                for (int kc = 0, c0 = 0; kc < kch; kc++, c0 += BK) {
                  mbar_wait_ut(eb, phase ^ 1, w_empty, kExp && dbg);
                  if (elect_one()) {
                    mbar_expect_tx(fb, tx);
                    tma_load_4d(sa, &maps.a, fb, c0, x0 + fw, y0 + fh, t.b0);
                    if (!bres)
                      for (int pl = 0; pl < npl; pl++)
                        tma_load_2d(sa + a_stage + pl * b_plane, &maps.b, fb, kcol + c0, pl * Npad + t.n0);
                  }
                  advance();
                }
          }
        }
      }
      __syncwarp();
    } else if constexpr (kExp)   // the generic loop (role counters, experiment switches) exists in experiment builds only
    for (int q = q_first; q < q_count; q += q_step) {
      const int tile = tile_of(q);
      const TileCoord t = decode_tile(P, tile);
      if (kExp && P.l2_prefetch > 0 && !cg2) {
        // activations are streamed from HBM once; with only a few stages in flight the ~2 us DRAM
        // latency is not covered, so the boxes of a later tile of this CTA are pulled into L2 now
        const int ptile = tile + P.l2_prefetch * (int)gridDim.x;
        if (ptile < num_tiles && elect_one()) {
          const TileCoord pt = decode_tile(P, ptile);
          if (MODE == 0) {
            for (int kc = 0; kc < P.kchunks; kc++) tma_prefetch_2d(&maps.a, kc * P.BK, pt.m0);
            if (P.res_tma)
              for (int j = 0; j < BN / 128; j++) tma_prefetch_2d(&maps.r, pt.n0 + j * 128, pt.m0);
          } else if (P.pair) {
            for (int fh = 0; fh < P.taps; fh++)
              tma_prefetch_2d(&maps.a, 0, ((pt.b0 * P.c.IH + pt.oh0 + fh) * P.c.IW + pt.ow0));
          } else {
            // the centre column of taps covers every input row/column the tile reads (side taps only
            // add one pixel that belongs to the neighbouring tile or to the zero padding)
            const int fwc = P.c.k / 2;
            for (int fh = 0; fh < P.c.k; fh++)
              for (int kc = 0; kc < P.kchunks; kc++)
                tma_prefetch_4d(&maps.a, kc * P.BK, pt.ow0 * P.c.stride - P.c.pad + fwc,
                                pt.oh0 * P.c.stride - P.c.pad + fh, pt.b0);
          }
        }
        __syncwarp();
      }
      if (MODE == 0 && P.res_tma) {
        // residual operand of this tile: full 128-byte lines, no L1, latency hidden by the run-ahead
        mbar_wait_warp(rempty_bar + 8 * rb, rphase ^ 1, (kExp ? P.poll_lane0 : 0));
        if (elect_one()) {
          mbar_expect_tx(rfull_bar + 8 * rb, (unsigned)res_tile);
          for (int j = 0; j < BN / 128; j++)
            tma_load_2d(smem_rres + rb * res_tile + j * (128 * 128), &maps.r, rfull_bar + 8 * rb, t.n0 + j * 128, t.m0);
        }
        __syncwarp();
        if (++rb == P.res_bufs) { rb = 0; rphase ^= 1; }
      }
      if constexpr (cg2) {
        if (P.hstream) {
          // halo tiles, streamed weights: per channel chunk ONE box with the tile's whole input window into the
          // halo ring, then the chunk's k*k half weight tiles through the pipeline stages
          for (int kc = 0; kc < P.kchunks; kc++) {
            mbar_wait_timed(rempty_bar + 8 * rb, rphase ^ 1, w_empty, dbg, (kExp ? P.poll_lane0 : 0));
            if (elect_one()) {
              const unsigned ab = rfull_bar + 8 * rb;
              if (cta_rank == 0) mbar_expect_tx(ab, 2u * (unsigned)P.a_bytes);
              else mbar_arrive_leader(ab);
              tma_load_4d_cg2(smem_aring + rb * P.a_ring_bytes, &maps.a, ab, kc * P.BK, t.ow0 - P.c.pad, t.oh0 - P.c.pad, t.b0);
            }
            __syncwarp();
            if (++rb == P.a_bufs) { rb = 0; rphase ^= 1; }
            const int brow = P.planes == 2 ? (int)cta_rank * P.Npad + t.n0 : t.n0 + (int)cta_rank * (BN / 2);
            for (int tap = 0; tap < P.taps; tap++) {
              mbar_wait_timed(empty_bar + 8 * stage, phase ^ 1, w_empty, dbg, (kExp ? P.poll_lane0 : 0));
              if (elect_one()) {
                const unsigned fb = full_bar + 8 * stage;
                if (cta_rank == 0) mbar_expect_tx(fb, 2u * (unsigned)P.b_stage_bytes);
                else mbar_arrive_leader(fb);
                tma_load_2d_cg2(smem_base + stage * stage_bytes, &maps.h, fb, tap * P.Cpm + kc * P.BK, brow);
              }
              __syncwarp();
              if (++stage == P.stages) { stage = 0; phase ^= 1; }
            }
          }
          continue;
        }
      }
      if (P.halo) {
        for (int kc = 0; kc < P.kchunks; kc++) {
          mbar_wait_timed(empty_bar + 8 * stage, phase ^ 1, w_empty, dbg, (kExp ? P.poll_lane0 : 0));
          const unsigned fb = full_bar + 8 * stage;
          if (elect_one()) {
            mbar_expect_tx(fb, (unsigned)P.a_bytes);
            tma_load_4d(smem_base + stage * stage_bytes, &maps.a, fb, kc * P.BK, t.ow0 - P.c.pad, t.oh0 - P.c.pad, t.b0);
          }
          __syncwarp();
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
        continue;
      }
      for (int tap = 0; tap < P.taps; tap++) {
        const int fh = P.pair ? tap : tap / P.c.k, fw = P.pair ? 0 : tap - fh * P.c.k;
        for (int kc = 0; kc < P.kchunks; kc++) {
          mbar_wait_timed(empty_bar + 8 * stage, phase ^ 1, w_empty, dbg, (kExp ? P.poll_lane0 : 0));
          const unsigned fb = full_bar + 8 * stage;
          const unsigned sa = smem_base + stage * stage_bytes;
          // sparse second plane: which planes hold weights anywhere in this (tap, K chunk)
          const unsigned cmask = (kExp && P.sparse2) ? (unsigned)P.c.blkmask[tap * P.kchunks + kc] : 0xffu;
          const bool ld0 = (cmask & 0x55u) != 0, ld1 = (cmask & 0xaau) != 0;
          if constexpr (cg2) {
            // Pair mode: both producers signal the LEADER's full barrier (its MMA warp drives both SMs).
            // This CTA stages its own 128 activation rows and its half of the weight tile: plane
            // `rank` of a two-plane layer, or rows rank*128.. of a 256-wide single plane — or, with the sparse
            // second plane, rows rank*64.. of EACH plane that holds weights in this chunk (one-plane MMAs, N = BN,
            // take half of their rows from either CTA).
            if (kExp && P.sparse2) {
              if (elect_one()) {
                const unsigned half = (unsigned)(64 * P.BK);
                const unsigned bbytes = (ld0 ? half : 0u) + (ld1 ? half : 0u);
                if (cta_rank == 0) mbar_expect_tx(fb, 2u * ((unsigned)P.a_bytes + bbytes));
                else mbar_arrive_leader(fb);
                if (MODE == 0) tma_load_2d_cg2(sa, &maps.a, fb, kc * P.BK, t.m0);
                else tma_load_4d_cg2(sa, &maps.a, fb, kc * P.BK, t.ow0 * P.c.stride - P.c.pad + fw, t.oh0 * P.c.stride - P.c.pad + fh, t.b0);
                if (ld0) tma_load_2d_cg2(sa + a_stage, &maps.h, fb, tap * P.Cpm + kc * P.BK, t.n0 + (int)cta_rank * 64);
                if (ld1) tma_load_2d_cg2(sa + a_stage + half, &maps.h, fb, tap * P.Cpm + kc * P.BK, P.Npad + t.n0 + (int)cta_rank * 64);
              }
            } else if (elect_one()) {
              if (cta_rank == 0) mbar_expect_tx(fb, 2u * (unsigned)(P.a_bytes + P.b_stage_bytes));
              else mbar_arrive_leader(fb);
              if (MODE == 0) {
                tma_load_2d_cg2(sa, &maps.a, fb, kc * P.BK, t.m0);
              } else {
                tma_load_4d_cg2(sa, &maps.a, fb, kc * P.BK, t.ow0 * P.c.stride - P.c.pad + fw,
                                t.oh0 * P.c.stride - P.c.pad + fh, t.b0);
              }
              const int brow = P.planes == 2 ? (int)cta_rank * P.Npad + t.n0 : t.n0 + (int)cta_rank * (BN / 2);
              tma_load_2d_cg2(sa + a_stage, &maps.h, fb, tap * P.Cpm + kc * P.BK, brow);
            }
          } else if (elect_one()) {
            const int nld = (kExp && P.sparse2) ? (int)ld0 + (int)ld1 : P.planes;
            mbar_expect_tx(fb, (unsigned)(P.a_bytes + (P.b_resident ? 0 : nld * P.b_bytes)));
            if (MODE == 0) {
              tma_load_2d(sa, &maps.a, fb, kc * P.BK, t.m0);
            } else if (P.pair) {
              // tap = filter row; chunk kc covers horizontal taps 2kc, 2kc+1 = pixels ow+2kc, ow+2kc+1
              tma_load_2d(sa, &maps.a, fb, 0, ((t.b0 * P.c.IH + t.oh0 + tap) * P.c.IW + t.ow0 + 2 * kc));
            } else {
              tma_load_4d(sa, &maps.a, fb, kc * P.BK, t.ow0 * P.c.stride - P.c.pad + fw,
                          t.oh0 * P.c.stride - P.c.pad + fh, t.b0);
            }
            if (!P.b_resident)
              for (int pl = 0; pl < P.planes; pl++)
                if (!(kExp && P.sparse2) || (pl == 0 ? ld0 : ld1))
                  tma_load_2d(sa + a_stage + pl * b_plane, &maps.b, fb, tap * P.Cpm + kc * P.BK, pl * P.Npad + t.n0);
          }
          __syncwarp();
          if (++stage == P.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
    if (dbg && lane == 0) {
      P.dbg[blockIdx.x * 8 + 0] = w_empty;
      P.dbg[blockIdx.x * 8 + 1] = clock64() - t_start;
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer (whole warp, one elected lane issues)
    int stage = 0;
    unsigned phase = 0;
    const bool dbg = kExp && P.dbg != nullptr;
    long long w_full = 0, w_tempty = 0, t_issue = 0, t_start = clock64();
    if (P.b_resident && (int)blockIdx.x < num_tiles) mbar_wait_warp(bres_bar, 0, (kExp ? P.poll_lane0 : 0));
    int li = 0;   // CTA-local tile index: TMEM buffer li & 1, its phase (li >> 1) & 1
    int ab = 0, hs_tap = 0, hs_fh = 0, hs_fw = 0;   // streamed-weight halo mode: ring slot, filter tap of the next stage
    unsigned aphase = 0;
    // experiment builds: where a steady-state stage of CTA 0's MMA thread spends its cycles
    long long fa[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = 0;
    const bool ftr = dbg && blockIdx.x == 0;
#define TF2B_TICK(i) do { if (kExp && ftr) { const long long t_ = clock64(); fa[i] += t_ - tprev; tprev = t_; } } while (0)
    // Lean form (see the producer): descriptors advanced by additions, nothing between the stage's barrier and its MMAs.  No tcgen05 fence after the full barrier: the operands arrive through TMA (async proxy,
    // completion on the mbarrier), not through tcgen05 operations of other threads.
    if (!kExp || P.lean_roles) {
      if (!cg2 || cta_rank == 0) {
        const int nst = P.stages, kch = P.kchunks, kk = P.c.k;
        const unsigned idesc = P.idesc;
        const unsigned long long d0 = make_smem_desc(smem_base, P.sbo16, P.layout_type);
        const unsigned long long dstep = (unsigned long long)((unsigned)stage_bytes >> 4);
        const unsigned long long px16 = (unsigned long long)(P.BK >> 4);              // one pixel of a halo tile
        const unsigned long long row16 = (unsigned long long)((P.Wp * P.BK) >> 4);   // one raster row of it
        unsigned long long ds = d0;
        unsigned fb = full_bar, eb = empty_bar;
        auto advance = [&]() {
          if (++stage == nst) { stage = 0; phase ^= 1; ds = d0; fb = full_bar; eb = empty_bar; }
          else { ds += dstep; fb += 8; eb += 8; }
        };
        bool tl_first = true;
        auto tl_mark = [&]() { if (kExp && tl_first) { tl_first = false; TF2B_TL(2); } };
        for (int q = q_first; q < q_count; q += q_step, li++) {
          const int buf = li & 1;
          mbar_wait_ut(tempty_bar + 8 * buf, ((unsigned)(li >> 1) & 1u) ^ 1u, w_tempty, kExp && dbg);   // epilogue has drained this accumulator
          tc_fence_after();
          const unsigned d_tmem = tmem_base + buf * acc_cols;
          unsigned acc = 0u;
          if constexpr (cg2) {
            if (P.hstream) {
              const unsigned long long dring0 = make_smem_desc(smem_aring, P.sbo16, P.layout_type);
              const unsigned long long ring16 = (unsigned long long)((unsigned)P.a_ring_bytes >> 4);
              for (int kc = 0; kc < kch; kc++) {
                mbar_wait_ut(rfull_bar + 8 * ab, aphase, w_full, kExp && dbg);
                unsigned long long da_row = dring0 + (unsigned long long)ab * ring16;
                for (int fh = 0; fh < kk; fh++) {
                  unsigned long long dat = da_row;
                  for (int fw = 0; fw < kk; fw++) {
                    mbar_wait_ut(fb, phase, w_full, kExp && dbg); tl_mark();
                    if (elect_one()) {
                      umma_i8_cg2(d_tmem, dat, ds, idesc, acc);
                      umma_i8_cg2(d_tmem, dat + 2ull, ds + 2ull, idesc, 1u);
                      umma_i8_cg2(d_tmem, dat + 4ull, ds + 4ull, idesc, 1u);
                      umma_i8_cg2(d_tmem, dat + 6ull, ds + 6ull, idesc, 1u);
                      umma_commit_cg2(eb);
                    }
                    acc = 1u;
                    dat += px16;
                    advance();
                  }
                  da_row += row16;
                }
                if (elect_one()) umma_commit_cg2(rempty_bar + 8 * ab);   // halo tile of this chunk consumed
                if (++ab == P.a_bufs) { ab = 0; aphase ^= 1; }
              }
            } else {
              const unsigned long long boff = (unsigned long long)((unsigned)a_stage >> 4);
              for (int it = 0; it < kiters; it++) {
                mbar_wait_ut(fb, phase, w_full, kExp && dbg); tl_mark();
                if (elect_one()) {
                  const unsigned long long db = ds + boff;
                  umma_i8_cg2(d_tmem, ds, db, idesc, acc);
                  umma_i8_cg2(d_tmem, ds + 2ull, db + 2ull, idesc, 1u);
                  umma_i8_cg2(d_tmem, ds + 4ull, db + 4ull, idesc, 1u);
                  umma_i8_cg2(d_tmem, ds + 6ull, db + 6ull, idesc, 1u);
                  umma_commit_cg2(eb);
                }
                acc = 1u;
                advance();
              }
            }
            if (elect_one()) umma_commit_cg2(tfull_bar + 8 * buf);
          } else if (!P.halo) {
            // single CTA, flat rows or one box per tap: the planes of a chunk lie back to back, ONE N = planes * BN
            // instruction per 32 channels covers them
            const bool bres = P.b_resident != 0, k128 = P.BK == 128;
            const unsigned long long boff = (unsigned long long)((unsigned)a_stage >> 4);
            const unsigned long long b_it16 = (unsigned long long)((unsigned)(P.planes * b_plane) >> 4);
            unsigned long long db_res = make_smem_desc(smem_res, P.sbo16, P.layout_type);
            for (int it = 0; it < kiters; it++, db_res += b_it16) {
              mbar_wait_ut(fb, phase, w_full, kExp && dbg); tl_mark();
              if (elect_one()) {
                const unsigned long long db = bres ? db_res : ds + boff;
                umma_i8(d_tmem, ds, db, idesc, acc);
                umma_i8(d_tmem, ds + 2ull, db + 2ull, idesc, 1u);
                if (k128) {
                  umma_i8(d_tmem, ds + 4ull, db + 4ull, idesc, 1u);
                  umma_i8(d_tmem, ds + 6ull, db + 6ull, idesc, 1u);
                }
                umma_commit(eb);
              }
              acc = 1u;
              advance();
            }
            if (elect_one()) umma_commit(tfull_bar + 8 * buf);
          } else {
            // halo tiles, resident weights: every tap of a channel chunk is a row-shifted view of the chunk's tile
            const unsigned long long dres = make_smem_desc(smem_res, P.sbo16, P.layout_type);
            const unsigned long long b_kc16 = (unsigned long long)((unsigned)(P.planes * b_plane) >> 4);
            const unsigned long long b_tap16 = (unsigned long long)kch * b_kc16;
            const bool k128 = P.BK == 128;
            unsigned long long db_kc = dres;
            for (int kc = 0; kc < kch; kc++, db_kc += b_kc16) {
              mbar_wait_ut(fb, phase, w_full, kExp && dbg); tl_mark();
              if (elect_one()) {
                unsigned long long da_row = ds, db = db_kc;
                for (int fh = 0; fh < kk; fh++) {
                  unsigned long long dat = da_row;
                  for (int fw = 0; fw < kk; fw++) {
                    umma_i8(d_tmem, dat, db, idesc, acc);
                    umma_i8(d_tmem, dat + 2ull, db + 2ull, idesc, 1u);
                    if (k128) {
                      umma_i8(d_tmem, dat + 4ull, db + 4ull, idesc, 1u);
                      umma_i8(d_tmem, dat + 6ull, db + 6ull, idesc, 1u);
                    }
                    acc = 1u;
                    dat += px16;
                    db += b_tap16;
                  }
                  da_row += row16;
                }
                umma_commit(eb);
              }
              acc = 1u;
              advance();
            }
            if (elect_one()) umma_commit(tfull_bar + 8 * buf);
          }
        }
      }
      __syncwarp();
    } else if constexpr (kExp)
    for (int q = q_first; q < q_count && (!cg2 || cta_rank == 0); q += q_step, li++) {   // pair mode: the leader issues
      const int buf = li & 1;
      mbar_wait_timed(tempty_bar + 8 * buf, ((unsigned)(li >> 1) & 1u) ^ 1u, w_tempty, dbg, (kExp ? P.poll_lane0 : 0));   // epilogue has drained this accumulator
      tc_fence_after();
      const unsigned d_tmem = tmem_base + buf * acc_cols;
      unsigned touched0 = 0u, touched1 = 0u;   // sparse second plane: has the plane's accumulator been written in this tile
      for (int it = 0; it < kiters; it++) {
        long long tr0 = 0;
        if (dbg) tr0 = clock64();
        if (kExp && ftr) tprev = clock64();
        if (kExp && ftr) mbar_wait_warp(full_bar + 8 * stage, phase, 0);
        else mbar_wait_timed(full_bar + 8 * stage, phase, w_full, dbg, (kExp ? P.poll_lane0 : 0));
        TF2B_TICK(0);
#ifndef TF2B_NO_STAGE_FENCE
        tc_fence_after();
#endif
        TF2B_TICK(1);
        long long tr1 = 0;
        if (dbg) tr1 = clock64();
        const unsigned sa = smem_base + stage * stage_bytes;
        const unsigned long long da = make_smem_desc(sa, P.sbo16, P.layout_type);
        long long ti0 = 0;
        if (dbg) ti0 = clock64();
        if (cg2 && P.hstream) {
          if constexpr (cg2) {
            // this stage = the half weight tiles of one (channel chunk, tap); the activations are the tap's
            // row-shifted view of the chunk's halo tile
            if (hs_tap == 0) {
              mbar_wait_timed(rfull_bar + 8 * ab, aphase, w_full, dbg, (kExp ? P.poll_lane0 : 0));
              tc_fence_after();
            }
            if (elect_one()) {
              const unsigned long long dah = make_smem_desc(smem_aring + ab * P.a_ring_bytes, P.sbo16, P.layout_type) +
                                             (unsigned long long)(((hs_fh * P.Wp + hs_fw) * P.BK) >> 4);
              const unsigned long long db = make_smem_desc(sa, P.sbo16, P.layout_type);
              umma_i8_cg2(d_tmem, dah, db, P.idesc, it > 0 ? 1u : 0u);
              umma_i8_cg2(d_tmem, dah + 2ull, db + 2ull, P.idesc, 1u);
              umma_i8_cg2(d_tmem, dah + 4ull, db + 4ull, P.idesc, 1u);
              umma_i8_cg2(d_tmem, dah + 6ull, db + 6ull, P.idesc, 1u);
              umma_commit_cg2(empty_bar + 8 * stage);
              if (hs_tap == P.taps - 1) umma_commit_cg2(rempty_bar + 8 * ab);   // halo tile of this chunk consumed
              if (it == kiters - 1) umma_commit_cg2(tfull_bar + 8 * buf);
            }
            if (++hs_fw == P.c.k) { hs_fw = 0; hs_fh++; }
            if (++hs_tap == P.taps) {
              hs_tap = hs_fh = hs_fw = 0;
              if (++ab == P.a_bufs) { ab = 0; aphase ^= 1; }
            }
          }
        } else if (P.halo) {
          if (elect_one()) {
            // all taps of this channel chunk read the same halo tile through row-shifted descriptors.
            // Only the 14-bit start-address field differs between taps, so the descriptors advance by
            // plain additions (no per-tap division / descriptor rebuild on this single-thread path).
            const unsigned long long a_px16 = (unsigned long long)(P.BK >> 4);
            const unsigned long long a_row16 = (unsigned long long)((P.Wp * P.BK) >> 4);
            const unsigned long long b_tap16 = (unsigned long long)((P.kchunks * P.planes * b_plane) >> 4);
            unsigned long long da_row = da;
            unsigned long long db = make_smem_desc(smem_res + it * P.planes * b_plane, P.sbo16, P.layout_type);
            unsigned acc0 = it > 0 ? 1u : 0u;
            const int kk = P.c.k;
            for (int fh = 0; fh < kk; fh++) {
              unsigned long long dat = da_row;
              for (int fw = 0; fw < kk; fw++) {
                umma_i8(d_tmem, dat, db, P.idesc, acc0);
                umma_i8(d_tmem, dat + 2ull, db + 2ull, P.idesc, 1u);
                if (P.BK == 128) {
                  umma_i8(d_tmem, dat + 4ull, db + 4ull, P.idesc, 1u);
                  umma_i8(d_tmem, dat + 6ull, db + 6ull, P.idesc, 1u);
                }
                acc0 = 1u;
                dat += a_px16;
                db += b_tap16;
              }
              da_row += a_row16;
            }
            umma_commit(empty_bar + 8 * stage);
            if (it == kiters - 1) umma_commit(tfull_bar + 8 * buf);
          }
        } else if (elect_one()) {
          // The planes of one k-chunk lie back to back in shared memory and in TMEM, so ONE instruction
          // of N = planes*BN (<= 256) covers them all: an N <= 128 instruction occupies the tensor pipe
          // as long as an N = 128 one, and fewer, wider instructions leave no issue bubble.
          const unsigned bsrc = P.b_resident ? smem_res + it * P.planes * b_plane : sa + a_stage;
          const unsigned long long db = make_smem_desc(bsrc, P.sbo16, P.layout_type);
          const unsigned acc0 = it > 0 ? 1u : 0u;
          if (kExp && P.sparse2) {
            // Sparse second plane: one-plane instructions (N = BN) into each plane's own TMEM columns, issued only
            // for the (32-channel block, plane) combinations that hold weights.  The first instruction a plane
            // sees in a tile overwrites its accumulator.
            const unsigned cmask = (unsigned)P.c.blkmask[it];
            const unsigned pstride = cg2 ? (unsigned)(64 * P.BK) : (unsigned)b_plane;   // pair: half the rows per CTA
            const unsigned long long db1 = make_smem_desc(bsrc + pstride, P.sbo16, P.layout_type);
            const int nblk = P.BK >> 5;
            for (int kb = 0; kb < nblk; kb++) {
              const unsigned m2 = (cmask >> (2 * kb)) & 3u;
              const unsigned long long koff = (unsigned long long)(2 * kb);
              if (m2 & 1u) {
                if constexpr (cg2) umma_i8_cg2(d_tmem, da + koff, db + koff, P.idesc1, touched0);
                else umma_i8(d_tmem, da + koff, db + koff, P.idesc1, touched0);
                touched0 = 1u;
              }
              if (m2 & 2u) {
                if constexpr (cg2) umma_i8_cg2(d_tmem + BN, da + koff, db1 + koff, P.idesc1, touched1);
                else umma_i8(d_tmem + BN, da + koff, db1 + koff, P.idesc1, touched1);
                touched1 = 1u;
              }
            }
            if constexpr (cg2) {
              umma_commit_cg2(empty_bar + 8 * stage);
              if (it == kiters - 1) umma_commit_cg2(tfull_bar + 8 * buf);
            } else {
              umma_commit(empty_bar + 8 * stage);
              if (it == kiters - 1) umma_commit(tfull_bar + 8 * buf);
            }
          } else
          // advance both descriptors by 32 bytes of K inside the swizzled row
          if constexpr (cg2) {
            // one M = 256 instruction spans both SMs: A = the two CTAs' activation tiles, B = their halves
            TF2B_TICK(2);
            umma_i8_cg2(d_tmem, da, db, P.idesc, acc0);
            TF2B_TICK(3);
            umma_i8_cg2(d_tmem, da + 2ull, db + 2ull, P.idesc, 1u);
            if (P.BK == 128) {
              umma_i8_cg2(d_tmem, da + 4ull, db + 4ull, P.idesc, 1u);
              umma_i8_cg2(d_tmem, da + 6ull, db + 6ull, P.idesc, 1u);
            }
            TF2B_TICK(4);
            umma_commit_cg2(empty_bar + 8 * stage);
            if (it == kiters - 1) umma_commit_cg2(tfull_bar + 8 * buf);
            TF2B_TICK(5);
          } else if (P.BK == 128) {
            umma_i8(d_tmem, da, db, P.idesc, acc0);
            umma_i8(d_tmem, da + 2ull, db + 2ull, P.idesc, 1u);
            umma_i8(d_tmem, da + 4ull, db + 4ull, P.idesc, 1u);
            umma_i8(d_tmem, da + 6ull, db + 6ull, P.idesc, 1u);
          } else {
            umma_i8(d_tmem, da, db, P.idesc, acc0);
            umma_i8(d_tmem, da + 2ull, db + 2ull, P.idesc, 1u);
          }
          if (!cg2 && !(kExp && P.sparse2)) {
            umma_commit(empty_bar + 8 * stage);             // frees the smem stage when the MMAs retire
            if (it == kiters - 1) umma_commit(tfull_bar + 8 * buf);   // accumulators complete -> epilogue
          }
        }
        __syncwarp();
        TF2B_TICK(6);
        if (kExp && ftr) fa[7] += 1;
        if (dbg) t_issue += clock64() - ti0;
        if (dbg && blockIdx.x == 0 && lane == 0 && it < 64) {
          // stage-position profile of CTA 0: cycles waiting for the stage, and from wait start to issue end
          const long long tr2 = clock64();
          P.dbg[8 * 148 + 3 * it + 0] += tr1 - tr0;
          P.dbg[8 * 148 + 3 * it + 1] += tr2 - tr0;
          P.dbg[8 * 148 + 3 * it + 2] += 1;
        }
        if (++stage == P.stages) { stage = 0; phase ^= 1; }
      }
    }
    if (kExp && ftr && lane == 0)
      for (int i = 0; i < 8; i++) P.dbg[8 * 148 + 3 * 64 + i] = fa[i];
#undef TF2B_TICK
    if (dbg && lane == 0) {
      P.dbg[blockIdx.x * 8 + 2] = w_full;
      P.dbg[blockIdx.x * 8 + 3] = w_tempty;
      P.dbg[blockIdx.x * 8 + 4] = clock64() - t_start;
      TF2B_TL(3);
      P.dbg[blockIdx.x * 8 + 7] = t_issue;
    }
  } else {
    // ===================================================== epilogue warps
    // Warp (quarter q, slice s) owns rows 32q..32q+31 (its TMEM lane quarter) x W = BN/4 columns.
    // Per tile: (1) before the accumulators are ready, fetch its per-channel params into its private
    // smem slice and prefetch its residual operand with the coalesced mapping; (2) tcgen05.ld ->
    // recombine planes -> requantise -> 16-byte st.shared into a private [32][W] staging tile;
    // (3) re-read the staging tile with lanes along the channel dimension so every global access
    // covers whole 32-byte sectors: residual add, 16-byte stores.  No cross-warp synchronisation.
    constexpr bool FAST = EPI >= 0;
    constexpr bool CT_TWO = FAST && (EPI & 1);
    constexpr bool CT_LOW = FAST && (EPI & 2);
    constexpr bool CT_RES = FAST && (EPI & 4);
    constexpr bool FOLD = FAST && (EPI & 8);
    constexpr bool HI32 = FOLD && (EPI & 16);
    constexpr bool TSTORE_GLOBAL = FOLD && MODE == 0;   // every channel has nshift >= 3: y = hi32(tot*(alpha<<(nshift-3)) + (B>>3))
    // epilogue groups; group g owns the tiles with local index % G == g.  Two groups (flat folded layers with a short
    // K loop, where the epilogue is the whole kernel): the TMEM reads and barrier waits of one tile overlap the
    // arithmetic of the other — measured -10 % on the 64 -> 256 layers, +13 % (worse) on 256 -> 1024.
    constexpr int G = (FOLD && MODE == 0 && !CG2 && BN == 128) ? GRP : 1;
    constexpr int SLICES = 4 / G;         // column slices of a tile (one warp per lane quarter and slice)
    constexpr int WT = BN / SLICES;       // columns per warp: 16..128
    constexpr int W = WT > 32 ? 32 : WT;  // columns per pass (staging tile width)
    constexpr int PASSES = WT / W;        // 1, 2 or 4
    constexpr int SEGS = W / 16;          // 16-byte segments per row (1 or 2) = iterations per pass
    constexpr int ROWS_PER_IT = 32 / SEGS;
    const int ew = warp - 2;              // 0..15 == hardware warp id
    const int quarter = hw_warp & 3;      // TMEM lane quarter this warp may access (hardware warp id % 4)
    const int slice = (ew >> 2) & (SLICES - 1);   // which part of the BN columns
    const int group = G == 2 ? (ew >> 3) : 0;
    const ConvParams& c = P.c;
    const int M = c.B * c.OH * c.OW;
    const unsigned p1mul = P.p1mul;           // 2^shift of the second plane
    const bool conv_relu = c.relu != 0;       // relu.cl:54, applied to the packed int8 values
    const bool add_relu = c.add_relu != 0;    // feature_writer.cl:126
    // epilogue scratch lives behind the pipeline stages in dynamic shared memory
    unsigned char* epi_base = smem_raw + (smem_base - smem_u32(smem_raw)) + P.stages * stage_bytes;
    unsigned char* stage = epi_base + ew * EPI_WARP_BYTES;                         // int8 staging tile [32][EPI_ROW]
    // per-channel params: [6][PSTR] int32 behind the staging tile; the folded form needs 768 bytes ({A} [64] int32,
    // {B} [64] int64) and sits behind the two 1 KB staging tiles of the TMA-store path instead
    int* prm = reinterpret_cast<int*>(stage + (FOLD ? 2048 : 32 * EPI_ROW));
    // coalesced mapping (constant per thread): iteration it -> row rl[it], 16-byte segment sg
    const int sg = lane % SEGS;
    int rl[SEGS];
    unsigned lut[SEGS];
#pragma unroll
    for (int it = 0; it < SEGS; it++) {
      rl[it] = it * ROWS_PER_IT + lane / SEGS;
      lut[it] = MODE == 1 ? row_lut[quarter * 32 + rl[it]] : 0u;
    }
    const bool has_res = FAST ? CT_RES : (c.r != nullptr);
    const int my_row = quarter * 32 + lane;     // accumulator row (TMEM lane) of this thread
    const unsigned my_lut = MODE == 1 ? row_lut[my_row] : 0u;
    int cached_ncol0 = -1;
    const bool dbg = kExp && P.dbg != nullptr && warp == 2;   // first epilogue warp
    long long w_tfull = 0, t_start = clock64();
    // residual operand of this thread's accumulator row: pixel of a tile, and a 16-byte-segment loader
    auto res_pixel = [&](const TileCoord& tc, bool& valid, long long& pix) {
      if (MODE == 0) {
        const int m = tc.m0 + my_row;
        valid = m < M;
        pix = m;
      } else {
        const int ow = tc.ow0 + (int)(my_lut & 0xff), oh = tc.oh0 + (int)((my_lut >> 8) & 0xff);
        const int b = tc.b0 + (int)((my_lut >> 16) & 0xff);
        valid = (my_lut >> 24) && (ow < c.OW) && (oh < c.OH) && (b < c.B);
        pix = ((long long)b * c.OH + oh) * c.OW + ow;
      }
    };
    int rb = 0;            // residual ring slot / phase and TMEM buffer / phase of the current tile,
    unsigned rphase = 0;   // all derived from the CTA-local tile index li
    int buf = 0;
    unsigned tph = 0;
    const bool res_tma = (MODE == 0) && (BN >= 128) && P.res_tma != 0;
    // residual bytes of this thread's row, columns [col, col + 16*SEGS) of the CTA tile, from the smem ring
    auto lds_res = [&](int col, uint4 (&dst)[SEGS]) {
      const unsigned rbase = smem_rres + rb * res_tile + (col >> 7) * (128 * 128) + my_row * 128;
#pragma unroll
      for (int q = 0; q < SEGS; q++) {
        const int chunk = ((col & 127) >> 4) + q;
        const unsigned a = rbase + (unsigned)(((chunk ^ (my_row & 7)) & 7) << 4);
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(dst[q].x), "=r"(dst[q].y), "=r"(dst[q].z), "=r"(dst[q].w) : "r"(a));
      }
    };
    auto load_res = [&](bool valid, long long pix, int ncolp, uint4 (&dst)[SEGS]) {
#pragma unroll
      for (int q = 0; q < SEGS; q++) dst[q] = make_uint4(0, 0, 0, 0);
      if (has_res && valid && !(kExp && (P.noepi & 2))) {
        const int8_t* rp = P.idx32 ? c.r + (unsigned)((int)pix * c.rC + ncolp) : c.r + pix * c.rC + ncolp;
        if (SEGS == 2 && P.direct256 && ncolp + 32 <= c.N) {
          ldg256(rp, dst[0], dst[SEGS - 1]);
          return;
        }
#pragma unroll
        for (int q = 0; q < SEGS; q++) {
          const int nq = ncolp + 16 * q;
          if (nq + 16 <= c.N) {
            dst[q] = __ldg(reinterpret_cast<const uint4*>(rp + 16 * q));
          } else if (nq < c.N) {   // ragged channel tail: byte loads
            unsigned w4[4] = {0, 0, 0, 0};
            for (int e = 0; e < c.N - nq; e++) w4[e >> 2] |= (unsigned)(unsigned char)rp[16 * q + e] << (8 * (e & 3));
            dst[q] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
          }
        }
      }
    };
    // barrier addresses pinned in registers (otherwise the shared-window base is rebuilt from
    // SR_CgaCtaId before every wait / arrive)
    unsigned e_tfull, e_tempty;
    asm volatile("mov.u32 %0, %1;" : "=r"(e_tfull) : "r"(tfull_bar));
    asm volatile("mov.u32 %0, %1;" : "=r"(e_tempty) : "r"(tempty_bar));
    int tsel = 0;   // staging buffer parity of the TMA-store path (advances by PASSES per tile)
    int e_mt = 0, e_nt = 0;
    const int step_mt = fdiv(G * q_step, P.d_ntiles), step_nt = G * q_step - step_mt * P.n_tiles;
    int li = group;   // CTA-local tile index
    bool lean_done = false;
    // ================= lean tile loop: flat layers with the folded epilogue.  Nothing in it depends on the lane's
    // pixel: the residual tile comes from the TMA ring and the finished rows leave through a TMA store (rows past
    // the end of the tensor are clipped by the hardware; rows of images beyond this run's batch are don't-cares),
    // so a warp does no address arithmetic at all.  The W accumulators of a pass are read by ONE tcgen05.ld and
    // the TMEM buffer goes back to the MMA warp BEFORE the arithmetic of the last pass.
    if constexpr (FOLD && MODE == 0) {
      if (P.tstore != 0 && (!CT_RES || res_tma) && !(kExp && (P.noepi != 0 || P.dbg != nullptr))) {
        lean_done = true;
        // the two ReLU flags are layer constants: four copies of the loop, picked once, instead of branches per 4 outputs
        auto lean_loop = [&](auto crelu_c, auto arelu_c) {
          constexpr bool CRELU = decltype(crelu_c)::value;
          constexpr bool ARELU = decltype(arelu_c)::value;
        const unsigned t_row_w = tmem_base + ((unsigned)(quarter * 32) << 16) + slice * WT;
        const int res_bufs = P.res_bufs;
        int lrb = group % res_bufs;
        unsigned lrph = (unsigned)(group / res_bufs) & 1u;
        for (int q = q_first + group * q_step; q < q_count; q += G * q_step, li += G) {
          const int lbuf = li & 1;
          const unsigned ltph = (unsigned)(li >> 1) & 1u;
          int m0, n0;
          if constexpr (!cg2) {
            if (li == group) {
              e_mt = fdiv(q, P.d_ntiles);
              e_nt = q - e_mt * P.n_tiles;
            } else {
              e_mt += step_mt;
              e_nt += step_nt;
              if (e_nt >= P.n_tiles) { e_nt -= P.n_tiles; e_mt++; }
            }
            m0 = e_mt * MMA_M;
            n0 = (P.n_tile0 + e_nt) * BN;
          } else {
            const TileCoord tc = decode_tile(P, tile_of(q));
            m0 = tc.m0;
            n0 = tc.n0;
          }
          const int ncolw = n0 + slice * WT;
          if (ncolw != cached_ncol0) {   // {A, B} of this warp's WT channels -> its private smem slice
            __syncwarp();
            cached_ncol0 = ncolw;
            for (int i = lane; i < WT; i += 32) {
              const int nn = ncolw + i;
              const long long al = __ldg(c.alpha + nn), bi = __ldg(c.bias + nn), be = __ldg(c.beta + nn);
              const int nsh = (int)__ldg(c.nshift + nn);
              const long long b64 = bi * al + ((be + 16384ll) << 20);
              prm[i] = HI32 ? (int)((unsigned)al << (nsh - 3)) : (int)((unsigned)al << nsh);
              reinterpret_cast<long long*>(prm + PSTR)[i] = HI32 ? (b64 >> 3) : b64;
            }
            __syncwarp();
          }
          mbar_wait(e_tfull + 8 * lbuf, ltph);
          tc_fence_after();
          if (CT_RES) mbar_wait(rfull_bar + 8 * lrb, lrph);
          const unsigned t_row0 = t_row_w + lbuf * acc_cols;
          const unsigned rbase = smem_rres + lrb * res_tile + my_row * 128;
#pragma unroll
          for (int pass = 0; pass < PASSES; pass++) {
            uint4 resq[SEGS];
            if (CT_RES) {
              const int col = slice * WT + pass * W;
              const unsigned rb2 = rbase + (col >> 7) * (128 * 128);
#pragma unroll
              for (int sq = 0; sq < SEGS; sq++) {
                const int chunk = ((col & 127) >> 4) + sq;
                const unsigned a = rb2 + (unsigned)(((chunk ^ (my_row & 7)) & 7) << 4);
                asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(resq[sq].x), "=r"(resq[sq].y), "=r"(resq[sq].z), "=r"(resq[sq].w) : "r"(a));
              }
            }
            unsigned char* sbuf = stage + ((tsel + pass) & 1) * 1024;
#pragma unroll
            for (int cc = 0; cc < W; cc += 16) {
              unsigned tot[16], tot1[16];
              tmem_ld16(t_row0 + pass * W + cc, tot);
              if (CT_TWO) tmem_ld16(t_row0 + BN + pass * W + cc, tot1);
              tmem_ld_wait();
              if (pass == PASSES - 1 && cc + 16 >= W) {
                // the accumulators of this tile are in registers: hand the TMEM buffer back before the arithmetic
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                  if constexpr (cg2) mbar_arrive_leader(e_tempty + 8 * lbuf);
                  else mbar_arrive(e_tempty + 8 * lbuf);
                }
              }
              const uint4 rq = CT_RES ? resq[cc / 16] : make_uint4(0, 0, 0, 0);
              const unsigned rw[4] = {rq.x, rq.y, rq.z, rq.w};
              unsigned packed[4];
              const int pc = pass * W + cc;
#pragma unroll
              for (int j4 = 0; j4 < 4; j4++) {
                const int4 pa = *reinterpret_cast<const int4*>(prm + pc + 4 * j4);
                const longlong2 pb0 = *reinterpret_cast<const longlong2*>(prm + PSTR + 2 * (pc + 4 * j4));
                const longlong2 pb1 = *reinterpret_cast<const longlong2*>(prm + PSTR + 2 * (pc + 4 * j4) + 4);
                const int aa[4] = {pa.x, pa.y, pa.z, pa.w};
                const long long bq[4] = {pb0.x, pb0.y, pb1.x, pb1.y};
                int yy[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                  const int j = 4 * j4 + u;
                  const int tt = CT_TWO ? (int)(tot1[j] * p1mul + tot[j]) : (int)tot[j];
                  const long long tq = (long long)tt * (long long)aa[u] + bq[u];   // IMAD.HI with the 64-bit addend
                  yy[u] = HI32 ? (int)(tq >> 32) : (int)(tq >> 35);
                }
                unsigned y4 = pack_sat4(yy[0], yy[1], yy[2], yy[3]);
                if (CRELU) y4 = relu_s8x4(y4);
                if (CT_RES) y4 = add_res_s8x4<ARELU>(y4, rw[j4]);
                packed[j4] = y4;
              }
              const int chunk = (W == 32) ? ((cc >> 4) ^ ((lane >> 2) & 1)) : 0;
              *reinterpret_cast<uint4*>(sbuf + lane * W + chunk * 16) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(smem_u32(sbuf), &maps.y, ncolw + pass * W, m0 + quarter * 32);
              tma_store_commit();
              tma_store_wait_read<1>();   // the other staging tile (written next) has been read out
            }
            __syncwarp();
          }
          if (CT_RES) {
            if (lane == 0) mbar_arrive(rempty_bar + 8 * lrb);
            lrb += G;
            if (lrb >= res_bufs) { lrb -= res_bufs; lrph ^= 1u; }
          }
          tsel = (tsel + PASSES) & 1;
        }
        };
        if (conv_relu) {
          if (add_relu) lean_loop(std::true_type{}, std::true_type{});
          else lean_loop(std::true_type{}, std::false_type{});
        } else {
          if (add_relu) lean_loop(std::false_type{}, std::true_type{});
          else lean_loop(std::false_type{}, std::false_type{});
        }
      }
    }
    // ================= generic tile loop (every other mode)
    for (int q = q_first + group * q_step; q < q_count && !lean_done; q += G * q_step, li += G) {
      const int tile = tile_of(q);
      if (G == 1) {
        if (li > 0) {   // next TMEM buffer / residual ring slot
          buf ^= 1;
          if (buf == 0) tph ^= 1u;
          if (MODE == 0 && P.res_tma && ++rb == P.res_bufs) { rb = 0; rphase ^= 1u; }
        }
      } else {
        buf = li & 1;
        tph = (unsigned)(li >> 1) & 1u;
        if (MODE == 0 && P.res_tma) {
          const int qq = li / P.res_bufs;
          rb = li - qq * P.res_bufs;
          rphase = (unsigned)qq & 1u;
        }
      }
      // (m-tile, n-tile) of this tile: advanced incrementally (tile += q_step per iteration), the
      // 64-bit fast division of decode_tile() only for the box geometry
      if (!cg2) {
        if (li == group) {
          e_mt = fdiv(tile, P.d_ntiles);
          e_nt = tile - e_mt * P.n_tiles;
        } else {
          e_mt += step_mt;
          e_nt += step_nt;
          if (e_nt >= P.n_tiles) { e_nt -= P.n_tiles; e_mt++; }
        }
      }
      TileCoord t;
      if (MODE == 0 && !cg2) {
        t.n0 = (P.n_tile0 + e_nt) * BN;
        t.m0 = e_mt * MMA_M;
        t.b0 = t.oh0 = t.ow0 = 0;
      } else {
        t = decode_tile(P, tile);
      }
      const int ncolw = t.n0 + slice * WT;        // first output channel of this warp
      // ---- (1a) params of this warp's WT channels -> smem, only when the channel slice changes
      __syncwarp();
      if (ncolw != cached_ncol0) {
        cached_ncol0 = ncolw;
        for (int i = lane; i < WT; i += 32) {
          const int nn = ncolw + i;
          const int be = __ldg(c.beta + nn);
          const int nsh = (int)__ldg(c.nshift + nn);
          const int bi = __ldg(c.bias + nn);
          const int al = __ldg(c.alpha + nn);
          if (FOLD) {
            // acc = tot * 2^nsh + bias without wrap-around (api.cu range analysis), so
            // acc*alpha + ((beta + 2^14) << 20) = tot * (alpha << nsh) + [bias*alpha + ((beta + 2^14) << 20)]
            // HI32: floor((8*X + B) / 2^35) = floor((X + floor(B/8)) / 2^32), so with alpha << (nsh-3) and
            // B >> 3 the result is the high word itself and the shift by 3 disappears
            const long long b64 = (long long)bi * (long long)al + (((long long)be + 16384ll) << 20);
            prm[i] = HI32 ? (int)((unsigned)al << (nsh - 3)) : (int)((unsigned)al << nsh);
            reinterpret_cast<long long*>(prm + PSTR)[i] = HI32 ? (b64 >> 3) : b64;
          } else {
            prm[i] = bi;
            prm[PSTR + i] = al;
            if (FAST) {
              // ((a + beta) >> 14 + 1) >> 1 == (acc*alpha + ((beta + 2^14) << 20)) >> 35 when nothing
              // wraps (checked per layer at load time, api.cu range analysis)
              const long long b64 = ((long long)be + 16384ll) << 20;
              prm[2 * PSTR + i] = (int)(unsigned)(b64 & 0xffffffffll);
              prm[3 * PSTR + i] = (int)(b64 >> 32);
            } else {
              prm[2 * PSTR + i] = be;
            }
            prm[4 * PSTR + i] = 1 << nsh;                                   // (x << s) == x * 2^s  (mod 2^32)
            prm[5 * PSTR + i] = (nsh + P.plane8_shift[1] < 32) ? (1 << (nsh + P.plane8_shift[1])) : 0;   // second plane: x * 2^(s + its shift)
          }
        }
      }
      // ---- (1b) this thread's accumulator row -> pixel (for the residual), and the pixels of its
      //           (row, segment) pairs in the coalesced store mapping
      bool rvalid = false;
      long long rpix = 0;
      res_pixel(t, rvalid, rpix);
      const bool dvalid = rvalid;
      const long long dpix = rpix;
      // flat layers with the folded epilogue: the warp's finished 32 rows x W bytes go to a swizzled staging tile in
      // shared memory and leave through ONE TMA store per pass (full lines, no LSU store wavefronts); the last,
      // partial m-tile of a run keeps the guarded direct stores (rows past the batch end must not be written)
      constexpr bool TSTORE_OK = FOLD && MODE == 0;
      const bool tstore = TSTORE_OK && P.tstore != 0 && (t.m0 + MMA_M <= M);
      const bool direct = (SEGS == 2) && P.direct256 != 0 && !tstore;
      if (TSTORE_OK && !tstore && !direct) {
        // the generic staging tile below shares memory with the TMA staging tiles: earlier stores must have read them
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
      }
      long long opix[SEGS];
#pragma unroll
      for (int it = 0; it < SEGS; it++) {
        opix[it] = -1;
        if (direct || tstore) continue;   // these paths store this thread's own row (dvalid / dpix)
        bool valid;
        long long pix;
        if (MODE == 0) {
          const int m = t.m0 + quarter * 32 + rl[it];
          valid = m < M;
          pix = m;
        } else {
          const int ow = t.ow0 + (int)(lut[it] & 0xff), oh = t.oh0 + (int)((lut[it] >> 8) & 0xff);
          const int b = t.b0 + (int)((lut[it] >> 16) & 0xff);
          valid = (lut[it] >> 24) && (ow < c.OW) && (oh < c.OH) && (b < c.B);
          pix = ((long long)b * c.OH + oh) * c.OW + ow;
        }
        opix[it] = valid ? pix : -1;
      }
      // residual of this thread's own row for the first pass (whole 32-byte sectors), requested
      // before the accumulators are waited for
      uint4 resq[SEGS];
      if (!res_tma) load_res(rvalid, rpix, ncolw, resq);
      __syncwarp();
      // ---- (2) accumulators -> requantise (+ residual) -> int8 staging tile -> (3) coalesced store
      mbar_wait_timed(e_tfull + 8 * buf, tph, w_tfull, dbg, (kExp ? P.poll_lane0 : 0));
      tc_fence_after();
      if (kExp && (P.noepi & 1)) {   // TF2B_MMA_NOEPI bit0: measure the TMA/MMA pipeline alone
        if (has_res && (MODE == 0) && (BN >= 128) && P.res_tma) {
          mbar_wait_warp(rfull_bar + 8 * rb, rphase, (kExp ? P.poll_lane0 : 0));
          __syncwarp();
          if (lane == 0) mbar_arrive(rempty_bar + 8 * rb);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            if constexpr (cg2) mbar_arrive_leader(e_tempty + 8 * buf);   // the leader's MMA warp waits for both CTAs
            else mbar_arrive(e_tempty + 8 * buf);
          }
        continue;
      }
      uint4 out_lo = make_uint4(0, 0, 0, 0), out_hi = make_uint4(0, 0, 0, 0);
      if (has_res && res_tma) {
        mbar_wait_warp(rfull_bar + 8 * rb, rphase, (kExp ? P.poll_lane0 : 0));
        lds_res(slice * WT, resq);
      }
#pragma unroll
      for (int pass = 0; pass < PASSES; pass++) {
        const int ncolp = ncolw + pass * W;         // first channel of this pass
        const int n = ncolp + sg * 16;              // first channel of this lane's 16-byte segment
        const unsigned t_row = tmem_base + ((unsigned)(quarter * 32) << 16) + buf * acc_cols + slice * WT + pass * W;
#pragma unroll
        for (int cc = 0; cc < W; cc += 16) {
          unsigned tot[16], low[16], tot1[16];
          const bool two = CT_TWO;                                          // a second scaled plane
          const bool has_low = FAST ? CT_LOW : (P.c.low_plane >= 0);
          tmem_ld16(t_row + cc, tot);
          if (has_low) tmem_ld16(t_row + P.c.low_plane * BN + cc, low);
          if (two) tmem_ld16(t_row + BN + cc, tot1);
          tmem_ld_wait();
          if (!has_low) {
#pragma unroll
            for (int j = 0; j < 16; j++) low[j] = 0;
          }
          if (!two) {
#pragma unroll
            for (int j = 0; j < 16; j++) tot1[j] = 0;
          }
          if (!FAST) {
            for (int pl = 1; pl < P.planes; pl++) {
              if (pl == P.c.low_plane) continue;
              unsigned v[16];
              tmem_ld16(t_row + pl * BN + cc, v);
              tmem_ld_wait();
              const unsigned mulq = 1u << P.plane8_shift[pl];
#pragma unroll
              for (int j = 0; j < 16; j++) tot[j] += v[j] * mulq;
            }
          }
          const uint4 rq = resq[cc / 16];
          const unsigned rw[4] = {rq.x, rq.y, rq.z, rq.w};
          unsigned packed[4];
          const int pc = pass * W + cc;
#pragma unroll
          for (int j4 = 0; j4 < 4; j4++) {
            int yy[4];
            if (FOLD) {
              const int4 pa = *reinterpret_cast<const int4*>(prm + pc + 4 * j4);
              const longlong2 pb0 = *reinterpret_cast<const longlong2*>(prm + PSTR + 2 * (pc + 4 * j4));
              const longlong2 pb1 = *reinterpret_cast<const longlong2*>(prm + PSTR + 2 * (pc + 4 * j4) + 4);
              const int aa[4] = {pa.x, pa.y, pa.z, pa.w};
              const long long bq[4] = {pb0.x, pb0.y, pb1.x, pb1.y};
#pragma unroll
              for (int u = 0; u < 4; u++) {
                const int j = 4 * j4 + u;
                const int tt = two ? (int)(tot1[j] * p1mul + tot[j]) : (int)tot[j];
                const long long t = (long long)tt * (long long)aa[u] + bq[u];   // IMAD.HI with the 64-bit addend
                yy[u] = HI32 ? (int)(t >> 32) : (int)(t >> 35);
              }
            } else {
              const int4 pb = *reinterpret_cast<const int4*>(prm + pc + 4 * j4);
              const int4 pa = *reinterpret_cast<const int4*>(prm + PSTR + pc + 4 * j4);
              const int4 pe = *reinterpret_cast<const int4*>(prm + 2 * PSTR + pc + 4 * j4);
              const int4 pm = *reinterpret_cast<const int4*>(prm + 4 * PSTR + pc + 4 * j4);
              const int bb[4] = {pb.x, pb.y, pb.z, pb.w}, aa[4] = {pa.x, pa.y, pa.z, pa.w};
              const int ee[4] = {pe.x, pe.y, pe.z, pe.w}, mm[4] = {pm.x, pm.y, pm.z, pm.w};
              if (FAST) {
                const int4 ph = *reinterpret_cast<const int4*>(prm + 3 * PSTR + pc + 4 * j4);
                const int hh[4] = {ph.x, ph.y, ph.z, ph.w};
                int m1[4] = {0, 0, 0, 0};
                if (two) {
                  const int4 pq = *reinterpret_cast<const int4*>(prm + 5 * PSTR + pc + 4 * j4);
                  m1[0] = pq.x; m1[1] = pq.y; m1[2] = pq.z; m1[3] = pq.w;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                  const int j = 4 * j4 + u;
                  unsigned a32 = tot[j] * (unsigned)mm[u] + (unsigned)bb[u];
                  if (two) a32 = tot1[j] * (unsigned)m1[u] + a32;
                  if (has_low) a32 += low[j];
                  const long long b64 = ((long long)hh[u] << 32) | (unsigned)ee[u];
                  const long long t = (long long)(int)a32 * (long long)aa[u] + b64;   // IMAD.WIDE with 64-bit addend
                  yy[u] = (int)(t >> 35);
                }
              } else {
#pragma unroll
                for (int u = 0; u < 4; u++) {
                  const int j = 4 * j4 + u;
                  const int a32 = (int)(tot[j] * (unsigned)mm[u] + (unsigned)bb[u] + low[j]);
                  // INT32 accumulator tap (pe.cl:196-199 prints this value): [image][N][OH][OW]
                  if (c.acc_dump != nullptr && dvalid && ncolp + cc + j < c.N) {
                    const int hw = c.OH * c.OW;
                    const int bimg = (int)(dpix / hw);
                    const int nl = c.acc_perm ? c.acc_perm[ncolp + cc + j] : ncolp + cc + j;
                    c.acc_dump[((size_t)bimg * c.N + nl) * hw + (int)(dpix - (long long)bimg * hw)] = a32;
                  }
                  yy[u] = requant_raw(a32, aa[u], ee[u]);
                }
              }
            }
            // pe.cl:194 clamp = saturating pack; relu.cl:54 and feature_writer.cl:124-127 on packed bytes
            unsigned y4 = pack_sat4(yy[0], yy[1], yy[2], yy[3]);
            if (conv_relu) y4 = relu_s8x4(y4);
            if (has_res) y4 = add_relu ? add_res_s8x4<true>(y4, rw[j4]) : add_res_s8x4<false>(y4, rw[j4]);
            packed[j4] = y4;
          }
          if (tstore) {
            // staging tile of this pass: [32 rows][W bytes], 16-byte chunk index XOR-swizzled like the tensor map
            // (SWIZZLE_32B: chunk ^= bit 7 of the address = (row >> 2) & 1): conflict-free 16-byte stores
            unsigned char* sbuf = stage + ((tsel + pass) & 1) * 1024;
            const int chunk = (W == 32) ? ((cc >> 4) ^ ((lane >> 2) & 1)) : 0;
            *reinterpret_cast<uint4*>(sbuf + lane * W + chunk * 16) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          } else if (direct) {
            if (cc == 0) out_lo = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            else out_hi = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          } else {
            *reinterpret_cast<uint4*>(stage + lane * EPI_ROW + cc) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
          }
        }
        if (pass == PASSES - 1) {
          // accumulator buffer drained: hand it back to the MMA warp as early as possible
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (cg2) mbar_arrive_leader(e_tempty + 8 * buf);   // the leader's MMA warp waits for both CTAs
            else mbar_arrive(e_tempty + 8 * buf);
          }
        } else {
          if (!direct) __syncwarp();
          if (has_res && res_tma) lds_res(slice * WT + (pass + 1) * W, resq);
          else load_res(rvalid, rpix, ncolp + W, resq);   // prefetch the next pass's residual
        }
        if (tstore) {
          // generic-proxy writes -> async proxy, then one lane issues the store and commits its bulk group
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(smem_u32(stage + ((tsel + pass) & 1) * 1024), &maps.y, ncolp, t.m0 + quarter * 32);
            tma_store_commit();
            // the buffer written next (the other one) must have been read out by its previous store
            tma_store_wait_read<1>();
          }
          __syncwarp();
          continue;
        }
        if (direct) {
          // this lane's own row: 32 contiguous bytes = one sector
          if (dvalid && ncolp + 32 <= c.N) {
            if (!(kExp && (P.noepi & 4))) {
              if (P.idx32) stg256(c.y + (unsigned)((int)dpix * c.yC + ncolp), out_lo, out_hi);
              else stg256(c.y + dpix * c.yC + ncolp, out_lo, out_hi);
            }
          } else if (dvalid && ncolp < c.N) {
            const unsigned vw[8] = {out_lo.x, out_lo.y, out_lo.z, out_lo.w, out_hi.x, out_hi.y, out_hi.z, out_hi.w};
            int8_t* dst = c.y + dpix * c.yC + ncolp;
            for (int e = 0; e < c.N - ncolp && e < 32; e++) dst[e] = (int8_t)((vw[e >> 2] >> (8 * (e & 3))) & 0xff);
          }
          continue;
        }
        // coalesced store of the finished int8 tile
#pragma unroll
        for (int it = 0; it < SEGS; it++) {
          if (opix[it] >= 0 && n < c.N) {
            const uint4 v = *reinterpret_cast<const uint4*>(stage + rl[it] * EPI_ROW + sg * 16);
            int8_t* dst = c.y + opix[it] * c.yC + n;
            const int nvalid = c.N - n;
            if (nvalid >= 16) {
              *reinterpret_cast<uint4*>(dst) = v;
            } else {
              const unsigned vw[4] = {v.x, v.y, v.z, v.w};
              for (int e = 0; e < nvalid; e++) dst[e] = (int8_t)((vw[e >> 2] >> (8 * (e & 3))) & 0xff);
            }
          }
        }
        if (pass != PASSES - 1) __syncwarp();   // staging tile is reused by the next pass
      }
      if (has_res && res_tma) {
        __syncwarp();
        if (lane == 0) mbar_arrive(rempty_bar + 8 * rb);
      }
      tsel = (tsel + PASSES) & 1;
    }
    if (TSTORE_GLOBAL && lane == 0) tma_store_wait_read<0>();   // staging memory must outlive its last store
    if (dbg && lane == 0) {
      P.dbg[blockIdx.x * 8 + 5] = w_tfull;
      P.dbg[blockIdx.x * 8 + 6] = clock64() - t_start;
      TF2B_TL(4);
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (cg2) cluster_sync_all();   // neither CTA may leave while its peer can still signal its barriers / read its smem
  if (threadIdx.x == 32) TF2B_TL(5);
  if (warp == 1) {
    tc_fence_after();
    if constexpr (cg2) tmem_dealloc_cg2(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn(std::string* err) {
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

int pick_bk(int Cp) { return (Cp % 128 == 0) ? 128 : 64; }

// "Pixel pair" rows: a tensor whose pixel is exactly 64 bytes (e.g. tensor 0: 27 channels + negated
// copy) read by an unpadded stride-1 convolution.  Two horizontally adjacent pixels are 128 contiguous
// bytes, so one SWIZZLE_128B row carries two taps; TMA cost is per row, so this halves it.
int pick_bn(int planes8, int N);

// Halo mode: k x k, stride 1, the whole weight slab of an n-tile resident in shared memory and a
// position-space row (OW + k - 1) that fits the 128 accumulator rows.  One TMA box per tile instead
// of one per filter tap; measured: the 64-channel layers were bound by the number of TMA boxes/rows.
// output columns per halo tile: the whole row when it fits the accumulator rows together with its halo
int halo_tw(int OW, int k) { return OW + k - 1 <= MMA_M ? OW : MMA_M - (k - 1); }

bool halo_fits(int k, int stride, int Cp, int OW, int OH, int N, int planes8, int BN);

// N tile of a layer: pick_bn(), narrowed to 64 when only that lets a k x k / stride-1 layer with 64 input
// channels keep its weight slab resident and run on halo tiles (VGG16's 64 -> 128 3x3 on 112 x 112 maps: the
// activation tile is fetched once per 64 output channels, still 9 x fewer TMA boxes than one per tap)
int layer_bn(int k, int stride, int Cp, int OW, int OH, int N, int planes8) {
  const int bn = pick_bn(planes8, N);
  if (bn > 64 && k >= 2 && stride == 1 && Cp <= 64 && !halo_fits(k, stride, Cp, OW, OH, N, planes8, bn) &&
      halo_fits(k, stride, Cp, OW, OH, N, planes8, 64))
    return 64;
  return bn;
}

bool halo_mode(int k, int stride, int Cp, int OW, int OH, int N, int planes8) {
  return halo_fits(k, stride, Cp, OW, OH, N, planes8, layer_bn(k, stride, Cp, OW, OH, N, planes8));
}

bool halo_fits(int k, int stride, int Cp, int OW, int OH, int N, int planes8, int BN) {
  static const bool allow = env_int("TF2B_MMA_HALO", 1) != 0;
  if (!allow || k < 2 || stride != 1 || k > 7) return false;
  const int BK = pick_bk(Cp);
  const int kchunks = (Cp + BK - 1) / BK, n_tiles = (N + BN - 1) / BN;
  const long long slab = (long long)k * k * kchunks * planes8 * BN * BK;
  if (slab > 112 * 1024 || n_tiles > 16) return false;
  // rows wider than the 128 accumulator rows are tiled in W: 128 - (k - 1) output columns + the halo
  const int Wp = halo_tw(OW, k) + k - 1;
  int th = MMA_M / Wp;
  if (th > OH) th = OH;
  const int rows = std::max(Wp * (th + k - 1), MMA_M + (k - 1) * (Wp + 1));
  const long long stage = ((long long)rows * BK + 1023) / 1024 * 1024;
  return slab + 2 * stage + EPI_BYTES + 2048 <= 226 * 1024;
}

bool pair_mode(int k, int stride, int pad, int Cp, int xC, int OW) {
  static const bool allow = env_int("TF2B_MMA_PAIR", 1) != 0;
  return allow && k >= 2 && stride == 1 && pad == 0 && Cp == 64 && xC == 64 && OW <= MMA_M;
}

// N tile: 256 for single-plane layers (one A tile feeds twice the MMA work; TMEM 2 x 256 columns),
// 128 up to two planes, else 64 (TMEM: 2 buffers x planes x BN <= 512 columns)
int pick_bn(int planes8, int N) {
  static const bool allow256 = env_int("TF2B_MMA_BN256", 1) != 0;
  if (planes8 == 1 && N >= 256 && allow256) return 256;
  if (planes8 <= 2 && N >= 128) return 128;
  return 64;
}

// geometry shared by the support test, the tensor-map builder and the launcher
// the folded requantisation (one IMAD.HI per output) applies: range analysis passed, <= 2 scaled planes,
// no unscaled low plane
bool fold_applies(const ConvParams& c, int planes8) {
  static const bool allow_fold = env_int("TF2B_MMA_FOLD", 1) != 0;
  return allow_fold && c.fast_requant >= 2 && c.low_plane < 0 && planes8 <= 2;
}

void fill_geometry(MmaParams& P, const ConvParams& c, int planes8) {
  P.c = c;
  P.planes = planes8;
  P.halo = halo_mode(c.k, c.stride, c.Cp, c.OW, c.OH, c.N, planes8) ? 1 : 0;
  P.Wp = halo_tw(c.OW, c.k) + c.k - 1;
  P.pair = (!P.halo && pair_mode(c.k, c.stride, c.pad, c.Cp, c.xC, c.OW)) ? 1 : 0;
  if (P.pair) {
    P.BK = 128;
    P.kchunks = (c.k + 1) / 2;
    P.Cpm = P.kchunks * 128;
    P.taps = c.k;
  } else {
    P.BK = pick_bk(c.Cp);
    P.Cpm = (c.Cp + P.BK - 1) / P.BK * P.BK;
    P.kchunks = P.Cpm / P.BK;
    P.taps = c.k * c.k;
  }
  P.BN = layer_bn(c.k, c.stride, c.Cp, c.OW, c.OH, c.N, planes8);
  P.Npad = c.Npad;
  P.n_tiles = (c.N + P.BN - 1) / P.BN;
  P.mode = (c.k == 1 && c.stride == 1 && c.pad == 0) ? 0 : 1;
  // Halo tiles with streamed weights: k x k / stride-1 layers that would otherwise run as CTA pairs with one
  // activation box PER TAP.  One box per tile and channel chunk instead: k*k times fewer activation bytes cross
  // L2 -> shared memory (the pair's feed drops from 64 to ~36 B/clk per SM on 256 -> 256 3x3 at 14 x 14).  The
  // price is the padded raster: tw * th of the 128 accumulator rows are real outputs, so narrow maps stay on boxes.
  P.hstream = 0;
  P.a_bufs = 0;
  P.a_ring_bytes = 0;
  {
    static const bool allow = env_int("TF2B_MMA_HSTREAM", 1) != 0;
    static const bool allow_cg2 = env_int("TF2B_MMA_CG2", 1) != 0;
    if (allow && allow_cg2 && !P.halo && !P.pair && P.mode == 1 && c.k >= 2 && c.k <= 7 && c.stride == 1 &&
        P.BK == 128 && planes8 * P.BN == 256 && planes8 <= 2 && fold_applies(c, planes8)) {
      const int tw = halo_tw(c.OW, c.k);
      int th = MMA_M / P.Wp;
      if (th > c.OH) th = c.OH;
      const int th_tiles = (c.OH + th - 1) / th;
      th = (c.OH + th_tiles - 1) / th_tiles;
      const int rows = std::max(P.Wp * (th + c.k - 1), MMA_M + (c.k - 1) * (P.Wp + 1));
      const int ring = (rows * P.BK + 1023) / 1024 * 1024;
      const int budget = 224 * 1024 - EPI_BYTES;
      const int bufs = (budget - 3 * ring) / (128 * P.BK) >= 4 ? 3 : 2;
      // (nothing here may depend on the batch of a particular run: the tensor maps are built once, for max_images)
      if (tw * th * 100 >= 70 * MMA_M && (budget - bufs * ring) / (128 * P.BK) >= 3) {
        P.hstream = P.halo = 1;
        P.a_bufs = bufs;
        P.a_ring_bytes = ring;
      }
    }
  }
  if (P.mode == 0) {
    P.tw = P.th = P.tn = 0;
    P.tiles_w = P.tiles_h = P.tiles_b = 0;
    const long long M = (long long)c.B * c.OH * c.OW;
    P.m_tiles = (int)((M + MMA_M - 1) / MMA_M);
    P.a_bytes = MMA_M * P.BK;
  } else {
    P.tw = P.halo ? halo_tw(c.OW, c.k) : (c.OW < MMA_M ? c.OW : MMA_M);
    P.th = P.pair ? 1 : MMA_M / (P.halo ? P.Wp : P.tw);
    if (P.th > c.OH) P.th = c.OH;
    // balance rows over the tiles of one image (14 rows -> 7+7 rather than 9+5)
    int th_tiles = (c.OH + P.th - 1) / P.th;
    P.th = (c.OH + th_tiles - 1) / th_tiles;
    P.tn = (P.th == c.OH && !P.pair && !P.halo) ? (MMA_M / (P.tw * P.th)) : 1;
    if (P.tn < 1) P.tn = 1;
    // NB: tn must not depend on the batch size of a particular run (the tensor map is built once for
    // max_images); images past the batch end are zero-filled / masked rows
    P.tiles_w = (c.OW + P.tw - 1) / P.tw;
    P.tiles_h = (c.OH + P.th - 1) / P.th;
    P.tiles_b = (c.B + P.tn - 1) / P.tn;
    P.m_tiles = P.tiles_w * P.tiles_h * P.tiles_b;
    P.a_bytes = P.halo ? P.Wp * (P.th + c.k - 1) * P.BK : P.tw * P.th * P.tn * P.BK;
  }
  P.a_stage_bytes = MMA_M * P.BK;
  if (P.halo) {
    // the shifted views of the last tap read (k-1)*(Wp+1) rows past the 128th (junk rows only)
    const int rows = std::max(P.Wp * (P.th + c.k - 1), MMA_M + (c.k - 1) * (P.Wp + 1));
    P.a_stage_bytes = (rows * P.BK + 1023) / 1024 * 1024;
  }
  if (P.hstream) P.a_stage_bytes = 0;   // the halo tiles have their own ring; the stages carry weights only
  P.b_bytes = P.BN * P.BK;
  // weight-stationary when the CTA's slab is small and the grid can be a multiple of n_tiles
  {
    static const bool allow = env_int("TF2B_MMA_BRES", 1) != 0;
    const long long slab = (long long)P.taps * P.kchunks * planes8 * P.b_bytes;
    P.b_resident = !P.hstream &&
                   (P.halo || (allow && slab <= 96 * 1024 && P.taps * P.kchunks * planes8 <= 64 && P.n_tiles <= 16));
    P.res_bytes = P.b_resident ? (int)slab : 0;
  }
  // CTA-pair mode for streaming-weight layers whose MMA is 256 wide: one SM ingests ~64 B/clk, an
  // M128 x N256 x K32 step needs 12 KB per 128 clk (96 B/clk); the pair's M256 x N256 step needs 8 KB per SM.
  P.b_stage_bytes = planes8 * P.b_bytes;
  P.cg2 = 0;
  {
    static const bool allow = env_int("TF2B_MMA_CG2", 1) != 0;
    if (allow && (!P.halo || P.hstream) && !P.pair && !P.b_resident && planes8 * P.BN == 256 && planes8 <= 2 &&
        (P.m_tiles >= 4 || P.hstream) && P.BK == 128 && fold_applies(c, planes8)) {
      P.cg2 = 1;
      P.b_stage_bytes = 128 * P.BK;   // half of the 256 weight rows
    }
  }
  const int stage_bytes = P.a_stage_bytes + (P.b_resident ? 0 : P.b_stage_bytes);
  {
    static const bool allow = env_int("TF2B_MMA_RESTMA", 1) != 0;
    P.res_tma = allow && c.r != nullptr && P.mode == 0 && P.BN >= 128 && (c.rC % 16 == 0);
    P.res_bufs = 2;
    if (P.res_tma) {
      // The residual operand streams from HBM once through a ring of whole tiles (128 rows x BN bytes); the
      // epilogue reads it with 16-byte shared loads instead of per-lane global loads.  Depth: whatever shared
      // memory leaves after three pipeline stages, at most TF2B_RBUFS_CAP (measured: 2 and 4 perform alike).
#ifndef TF2B_RBUFS_CAP
#define TF2B_RBUFS_CAP 4
#endif
      static const int cap = env_int("TF2B_MMA_RBUFS", TF2B_RBUFS_CAP);
      const int budget = 224 * 1024 - EPI_BYTES - P.res_bytes;
      const int per_tile = P.kchunks * stage_bytes + P.BN * 128;
      int t = budget / per_tile;
      if (t > MAX_RBUFS) t = MAX_RBUFS;
      if (t > cap) t = cap;
      if (t < 2) t = 2;
      P.res_bufs = t;
      auto stages_with = [&](int bufs) { return (budget - bufs * P.BN * 128) / stage_bytes; };
      if (stages_with(P.res_bufs) < 3) P.res_bufs = 2;
      if (stages_with(P.res_bufs) < 3) P.res_tma = 0;   // a single buffer would serialise the producer
    }
  }
  const int rres_bytes = P.res_tma ? P.res_bufs * P.BN * 128 : (P.hstream ? P.a_bufs * P.a_ring_bytes : 0);
  const int stage_bytes_final = P.a_stage_bytes + (P.b_resident ? 0 : P.b_stage_bytes);
  int st = (224 * 1024 - EPI_BYTES - P.res_bytes - rres_bytes) / stage_bytes_final;
  P.stages = st > MAX_STAGES ? MAX_STAGES : (st < 2 ? 2 : st);
  {
    static const int cap = env_int("TF2B_MMA_STAGES", 0);   // experiment switch
    if (cap >= 1 && P.stages > cap) P.stages = cap;
  }
  // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): c_format S32 (2) at bit 4,
  // a/b format signed int8 (1) at bits 7 / 10, K-major A and B, N>>3 at bit 17, M>>4 at bit 24,
  // saturate (bit 3) off
  P.idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)((P.BN * P.planes) >> 3) << 17) |
            ((unsigned)((P.cg2 ? 2 * MMA_M : MMA_M) >> 4) << 24);
  {
    P.b_packed4 = (c.w4_avail && P.b_resident && P.res_bytes / 2 <= P.stages * stage_bytes_final) ? 1 : 0;
  }
  P.idesc1 = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(P.BN >> 3) << 17) |
             ((unsigned)((P.cg2 ? 2 * MMA_M : MMA_M) >> 4) << 24);
  P.sparse2 = (kExp && c.sparse2 && planes8 == 2 && !P.halo && !P.pair && P.BN >= 128) ? 1 : 0;   // experiment builds only
  P.p1mul = 1u << (c.plane_shift[1] & 31);
  P.layout_type = (P.BK == 128) ? 2u : 4u;
  P.sbo16 = (unsigned)(8 * P.BK) >> 4;
  P.d_ntiles = make_fastdiv(P.n_tiles);
  P.d_tiles_w = make_fastdiv(P.tiles_w);
  P.d_tiles_h = make_fastdiv(P.tiles_h);
  {
    static const int poll = env_int("TF2B_MMA_POLL0", 0);
    static const int top = env_int("TF2B_MMA_TOP", 0);
    P.poll_lane0 = poll;
    P.roles_top = top;
    static const int noepi = env_int("TF2B_MMA_NOEPI", 0);
    P.noepi = noepi;
    static const int l2pf = env_int("TF2B_MMA_L2PF", 0);
    P.l2_prefetch = l2pf;
  }
  {
    static const bool allow = env_int("TF2B_MMA_LEANROLES", 1) != 0;
    P.lean_roles = (!kExp || (allow && !P.sparse2 && P.l2_prefetch == 0)) ? 1 : 0;
  }
  P.n_tile0 = 0;
  // per-warp TMA stores: flat layers whose run-time epilogue is the folded one (the accumulator tap runs the exact
  // epilogue and keeps the guarded direct stores)
  {
    const int scaled = planes8 - (c.low_plane >= 0 ? 1 : 0);
    const bool fast = c.acc_dump == nullptr && c.fast_requant != 0 && scaled <= 2 && (c.low_plane < 0 || c.low_plane == planes8 - 1);
    static const bool allow = env_int("TF2B_MMA_TSTORE", 1) != 0;
    P.tstore = (allow && P.mode == 0 && fast && fold_applies(c, planes8)) ? 1 : 0;
  }
  P.idx32 = ((long long)c.B * c.OH * c.OW * c.yC < (1ll << 31)) && ((long long)c.B * c.OH * c.OW * (c.rC > 0 ? c.rC : 1) < (1ll << 31));
  P.direct256 = (c.yC % 32 == 0) && (((unsigned long long)c.y) % 32 == 0) &&
                (c.r == nullptr || ((c.rC % 32 == 0) && (((unsigned long long)c.r) % 32 == 0)));
  {
    static const bool allow = env_int("TF2B_MMA_DIRECT", 1) != 0;
    if (!allow) P.direct256 = 0;
  }
}

}  // namespace

int mma_bn() { return 256; }   // weight planes / params are padded to a multiple of this many rows
int mma_pick_bk(int Cp) { return pick_bk(Cp); }
bool mma_pair_mode(int k, int stride, int pad, int Cp, int xC, int OW, int OH, int N, int planes8) {
  return !halo_mode(k, stride, Cp, OW, OH, N, planes8) && pair_mode(k, stride, pad, Cp, xC, OW);
}

// bytes of the resident weight slab of one n-tile (0 when the layer streams its weights): api.cu keeps a packed
// 4-bit copy of the planes for the layers that have one
long long mma_slab_bytes(int k, int Cp, int N, int planes8) {
  const int BK = pick_bk(Cp), BN = pick_bn(planes8, N);
  const long long slab = (long long)k * k * ((Cp + BK - 1) / BK) * planes8 * BN * BK;
  return slab <= 112 * 1024 ? slab : 0;
}

// can a two-plane layer of this shape use the sparse second plane (streamed or resident weights, not the halo tile)?
// MEASURED AND SWITCHED OFF (profiles/r02_ab_sparse2.md): one-plane instructions (N = 128) keep the tensor pipe
// less busy than the two-plane N = 256 ones they replace — every K-heavy layer of ResNet50 got 15-50 % slower
// although 25-45 % fewer (block, plane) pairs were issued.  The code path stays for experiment builds.
bool mma_sparse2_ok(const tf2b_layer_desc& L, int in_pitch, int N) {
  (void)in_pitch;
  static const bool allow = env_int("TF2B_MMA_SPARSE2", 0) != 0;
  const int Cp = (L.C + 15) / 16 * 16;
  return allow && pick_bn(2, N) >= 128 && !halo_mode(L.k, L.stride, Cp, L.OW, L.OH, N, 2);
}

bool mma_layer_supported(const tf2b_layer_desc& L, int in_pitch, int planes8) {
  if (L.ipool) return false;
  if (L.stride < 1 || L.stride > 8) return false;   // TMA traversal stride limit
  {
    int tw = L.OW < MMA_M ? L.OW : MMA_M;
    if ((tw - 1) * L.stride + 1 > 256) return false;  // TMA box limit
  }
  if (planes8 < 1 || planes8 > kMaxPlanes) return false;
  if (in_pitch % 16 != 0) return false;
  if (L.OW > 256 || L.k > 7) return false;
  // TMEM: two accumulator buffers of planes*BN columns
  int BN = pick_bn(planes8, L.N);
  if (2 * planes8 * BN > TMEM_COLS) return false;
  return true;
}

size_t mma_tmap_bytes() { return sizeof(TmapPair); }

// human-readable launch plan of a tensor-core layer (introspection for tests / profiles)
std::string mma_describe(const ConvParams& c, int planes8) {
  MmaParams P;
  fill_geometry(P, c, planes8);
  char b[192];
  snprintf(b, sizeof b, "mma BN%d BK%d planes%d %s%s%s%s%s%s%s%s stages%d", P.BN, P.BK, planes8,
           P.mode == 0 ? "flat" : (P.hstream ? "halo wstream" : (P.halo ? "halo" : (P.pair ? "pixelpair" : "box"))),
           P.b_resident ? " wres" : "",
           P.res_tma ? " restma" : "", fold_applies(c, planes8) ? (c.fast_requant >= 3 ? " fold hi32" : " fold") : "",
           P.tstore ? " tmastore" : "", P.cg2 ? " ctapair" : "", P.sparse2 ? " sparse2" : "", P.b_packed4 ? " packed4" : "", P.stages);
  return std::string(b);
}

int mma_build_tmaps(void* host_tmaps, const ConvParams& c, const int8_t* wgt8, const uint8_t* wgt4, int planes8,
                    std::string* err) {
  EncodeTiledFn enc = get_encode_fn(err);
  if (!enc) return -1;
  MmaParams P;
  fill_geometry(P, c, planes8);
  TmapPair* tp = reinterpret_cast<TmapPair*>(host_tmaps);
  const CUtensorMapSwizzle sw = (P.BK == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  CUresult r;
  if (P.pair) {
    // rows = pixels, 128 bytes (this pixel and the next) per row: overlapping rows, stride 64 bytes
    // the row of the very last pixel reaches 64 bytes past the tensor: buffers carry that much slack
    // (api.cu alloc_runtime) and the weights of that slot are zero
    cuuint64_t dims[2] = {128, (cuuint64_t)c.B * c.IH * c.IW};
    cuuint64_t strides[1] = {64};
    cuuint32_t box[2] = {128, (cuuint32_t)P.tw};
    cuuint32_t es[2] = {1, 1};
    r = enc(&tp->a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)c.x, dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else if (P.mode == 0) {
    cuuint64_t dims[2] = {(cuuint64_t)c.Cp, (cuuint64_t)c.B * c.IH * c.IW};
    cuuint64_t strides[1] = {(cuuint64_t)c.xC};
    cuuint32_t box[2] = {(cuuint32_t)P.BK, (cuuint32_t)MMA_M};
    cuuint32_t es[2] = {1, 1};
    r = enc(&tp->a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)c.x, dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)c.Cp, (cuuint64_t)c.IW, (cuuint64_t)c.IH, (cuuint64_t)c.B};
    cuuint64_t strides[3] = {(cuuint64_t)c.xC, (cuuint64_t)c.xC * c.IW, (cuuint64_t)c.xC * c.IW * c.IH};
    // a strided convolution reads every stride-th pixel: the box spans (t-1)*stride+1 source
    // elements and the traversal stride keeps ceil(box/stride) = t of them
    const cuuint32_t st = (cuuint32_t)c.stride;
    cuuint32_t box[4] = {(cuuint32_t)P.BK, (cuuint32_t)((P.tw - 1) * st + 1), (cuuint32_t)((P.th - 1) * st + 1),
                         (cuuint32_t)P.tn};
    if (P.halo) {   // the whole input window of the tile: Wp x (th + k - 1) pixels (stride 1)
      box[1] = (cuuint32_t)P.Wp;
      box[2] = (cuuint32_t)(P.th + c.k - 1);
    }
    cuuint32_t es[4] = {1, st, st, 1};
    r = enc(&tp->a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void*)c.x, dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled(A) failed with CUresult " + std::to_string((int)r);
    return -1;
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)c.Kp, (cuuint64_t)planes8 * c.Npad};
    cuuint64_t strides[1] = {(cuuint64_t)c.Kp};
    cuuint32_t box[2] = {(cuuint32_t)P.BK, (cuuint32_t)P.BN};
    cuuint32_t es[2] = {1, 1};
    r = enc(&tp->b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)wgt8, dims, strides, box, es,
            CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled(B) failed with CUresult " + std::to_string((int)r);
    return -1;
  }
  memset(&tp->r, 0, sizeof tp->r);
  memset(&tp->h, 0, sizeof tp->h);
  if (P.cg2) {
    cuuint64_t dims[2] = {(cuuint64_t)c.Kp, (cuuint64_t)planes8 * c.Npad};
    cuuint64_t strides[1] = {(cuuint64_t)c.Kp};
    cuuint32_t box[2] = {(cuuint32_t)P.BK, (cuuint32_t)(P.sparse2 ? 64 : 128)};   // sparse second plane: half of EACH plane
    cuuint32_t es[2] = {1, 1};
    r = enc(&tp->h, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)wgt8, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      if (err) *err = "cuTensorMapEncodeTiled(B half) failed with CUresult " + std::to_string((int)r);
      return -1;
    }
  }
  memset(&tp->p, 0, sizeof tp->p);
  if (P.b_packed4) {
    cuuint64_t dims[2] = {(cuuint64_t)(c.Kp / 2), (cuuint64_t)planes8 * c.Npad};
    cuuint64_t strides[1] = {(cuuint64_t)(c.Kp / 2)};
    cuuint32_t box[2] = {(cuuint32_t)(P.BK / 2), (cuuint32_t)P.BN};
    cuuint32_t es[2] = {1, 1};
    r = enc(&tp->p, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)wgt4, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      if (err) *err = "cuTensorMapEncodeTiled(packed B) failed with CUresult " + std::to_string((int)r);
      return -1;
    }
  }
  memset(&tp->y, 0, sizeof tp->y);
  if (P.tstore) {
    // the warp (quarter, slice) of the epilogue owns 32 rows x W = min(32, BN / 4) channels per pass
    const int W = P.BN / 4 > 32 ? 32 : P.BN / 4;
    cuuint64_t dims[2] = {(cuuint64_t)c.N, (cuuint64_t)c.B * c.OH * c.OW};
    cuuint64_t strides[1] = {(cuuint64_t)c.yC};
    cuuint32_t box[2] = {(cuuint32_t)W, 32};
    cuuint32_t es[2] = {1, 1};
    r = enc(&tp->y, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)c.y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            W == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      if (err) *err = "cuTensorMapEncodeTiled(Y) failed with CUresult " + std::to_string((int)r);
      return -1;
    }
  }
  if (P.res_tma) {
    cuuint64_t dims[2] = {(cuuint64_t)c.N, (cuuint64_t)c.B * c.OH * c.OW};
    cuuint64_t strides[1] = {(cuuint64_t)c.rC};
    cuuint32_t box[2] = {128, (cuuint32_t)MMA_M};
    cuuint32_t es[2] = {1, 1};
    r = enc(&tp->r, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)c.r, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      if (err) *err = "cuTensorMapEncodeTiled(R) failed with CUresult " + std::to_string((int)r);
      return -1;
    }
  }
  return 0;
}

namespace {
using KernelFn = void (*)(MmaParams, TmapPair);
constexpr int kEpiVariants = 21;   // index = EPI + 1 for EPI <= 15; 17..20 = the hi32 forms of 8, 9, 12, 13
struct KernelTables {
  KernelFn single[3][2][kEpiVariants];   // [BN 64/128/256][mode][epilogue]
  KernelFn pair[2][2][2][2];             // CTA pairs: [BN 128 two planes / 256 one plane][mode][residual][hi32]
  KernelFn pair_exact[2][2];             // CTA pairs with the exact epilogue (accumulator tap)
  KernelFn grp2[2][2][2][2];             // flat BN = 128, two epilogue groups: [planes 1/2][residual][hi32] x [fold only = 0]
};
const KernelTables& kernel_tables() {
#define TF2B_EPI_ROW(BN_, MODE_)                                                                              \
  {conv_mma_kernel<BN_, MODE_, -1>, conv_mma_kernel<BN_, MODE_, 0>, conv_mma_kernel<BN_, MODE_, 1>,            \
   conv_mma_kernel<BN_, MODE_, 2>,  conv_mma_kernel<BN_, MODE_, 3>, conv_mma_kernel<BN_, MODE_, 4>,            \
   conv_mma_kernel<BN_, MODE_, 5>,  conv_mma_kernel<BN_, MODE_, 6>, conv_mma_kernel<BN_, MODE_, 7>,            \
   conv_mma_kernel<BN_, MODE_, 8>,  conv_mma_kernel<BN_, MODE_, 9>, nullptr, nullptr,                          \
   conv_mma_kernel<BN_, MODE_, 12>, conv_mma_kernel<BN_, MODE_, 13>, nullptr, nullptr,                         \
   conv_mma_kernel<BN_, MODE_, 24>, conv_mma_kernel<BN_, MODE_, 25>, conv_mma_kernel<BN_, MODE_, 28>,          \
   conv_mma_kernel<BN_, MODE_, 29>}
  static const KernelTables T = {
      {{TF2B_EPI_ROW(64, 0), TF2B_EPI_ROW(64, 1)}, {TF2B_EPI_ROW(128, 0), TF2B_EPI_ROW(128, 1)},
       {TF2B_EPI_ROW(256, 0), TF2B_EPI_ROW(256, 1)}},
      {{{{conv_mma_kernel<128, 0, 9, true>, conv_mma_kernel<128, 0, 25, true>},
         {conv_mma_kernel<128, 0, 13, true>, conv_mma_kernel<128, 0, 29, true>}},
        {{conv_mma_kernel<128, 1, 9, true>, conv_mma_kernel<128, 1, 25, true>},
         {conv_mma_kernel<128, 1, 13, true>, conv_mma_kernel<128, 1, 29, true>}}},
       {{{conv_mma_kernel<256, 0, 8, true>, conv_mma_kernel<256, 0, 24, true>},
         {conv_mma_kernel<256, 0, 12, true>, conv_mma_kernel<256, 0, 28, true>}},
        {{conv_mma_kernel<256, 1, 8, true>, conv_mma_kernel<256, 1, 24, true>},
         {conv_mma_kernel<256, 1, 12, true>, conv_mma_kernel<256, 1, 28, true>}}}},
      {{conv_mma_kernel<128, 0, -1, true>, conv_mma_kernel<128, 1, -1, true>},
       {conv_mma_kernel<256, 0, -1, true>, conv_mma_kernel<256, 1, -1, true>}},
      {{{{conv_mma_kernel<128, 0, 8, false, 2>, nullptr}, {conv_mma_kernel<128, 0, 24, false, 2>, nullptr}},
        {{conv_mma_kernel<128, 0, 12, false, 2>, nullptr}, {conv_mma_kernel<128, 0, 28, false, 2>, nullptr}}},
       {{{conv_mma_kernel<128, 0, 9, false, 2>, nullptr}, {conv_mma_kernel<128, 0, 25, false, 2>, nullptr}},
        {{conv_mma_kernel<128, 0, 13, false, 2>, nullptr}, {conv_mma_kernel<128, 0, 29, false, 2>, nullptr}}}}};
#undef TF2B_EPI_ROW
  return T;
}
}  // namespace

// Per-device set-up of the tensor-core kernels: the > 48 KB dynamic shared memory opt-in is a per-device
// function attribute, so every engine calls this from tf2b_finalize after cudaSetDevice(its device).
cudaError_t mma_prepare_device(int* num_sms) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  const KernelTables& T = kernel_tables();
  const KernelFn* all = &T.single[0][0][0];
  const int n_all = (int)(sizeof(KernelTables) / sizeof(KernelFn));
  for (int i = 0; i < n_all; i++) {
    if (!all[i]) continue;
    e = cudaFuncSetAttribute(all[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

cudaError_t launch_conv_mma(const ConvParams& c, const MmaHostParams& /*hp*/, int planes8, const int* plane8_shift,
                            void* tmaps, int num_sms, cudaStream_t stream) {
  MmaParams P;
  fill_geometry(P, c, planes8);
  for (int i = 0; i < kMaxPlanes; i++) P.plane8_shift[i] = plane8_shift[i];
  const int stage_bytes = P.a_stage_bytes + (P.b_resident ? 0 : P.b_stage_bytes);
  const size_t smem = (size_t)P.res_bytes + (size_t)P.stages * stage_bytes + EPI_BYTES +
                      (P.res_tma ? (size_t)P.res_bufs * P.BN * 128 : 0) + (size_t)P.a_bufs * P.a_ring_bytes + 1024;
  // the fast requantisation needs at most two scaled planes (+ the optional low plane); the INT32
  // accumulator tap lives in the exact epilogue (EPI < 0), which sees the very TMEM accumulators the
  // fast forms would
  const int scaled_planes = planes8 - (c.low_plane >= 0 ? 1 : 0);
  const bool fast = c.acc_dump == nullptr && c.fast_requant != 0 && scaled_planes <= 2 &&
                    (c.low_plane < 0 || c.low_plane == planes8 - 1);
  // folded form (one IMAD.HI per output): additionally needs alpha << nshift in int32 and an accumulator
  // that cannot wrap (fast_requant == 2, api.cu), and no unscaled low plane
  const bool fold = fast && fold_applies(c, planes8);
  const int epi = fast ? (1 + ((scaled_planes == 2 ? 1 : 0) | (c.low_plane >= 0 ? 2 : 0) | (c.r != nullptr ? 4 : 0) |
                               (fold ? 8 : 0)))
                       : 0;
  // (a variant with two epilogue groups of 8 warps taking alternate tiles was measured: no gain on the
  // epilogue-bound layers, slower with a residual operand — the epilogue is throughput bound, L1TEX ~72 %)
  static const bool allow_hi32 = env_int("TF2B_MMA_HI32", 1) != 0;
  const bool hi32 = fold && allow_hi32 && c.fast_requant >= 3;
  int epi_idx = epi;
  if (hi32) {
    const int bits = epi - 1;   // 8, 9, 12 or 13
    epi_idx = 17 + ((bits & 1) | ((bits & 4) >> 1));
  }
  // two epilogue groups where the K loop is at most 128 channels deep
#ifndef TF2B_GRP2_MAXK
#define TF2B_GRP2_MAXK 128
#endif
  const bool grp2 = fold && P.mode == 0 && !P.cg2 && P.BN == 128 && P.taps * P.kchunks * P.BK <= TF2B_GRP2_MAXK;
  P.egroups = grp2 ? 2 : 1;
  const KernelTables& T = kernel_tables();
  KernelFn kfn = T.single[P.BN == 256 ? 2 : (P.BN == 128 ? 1 : 0)][P.mode][epi_idx];
  if (P.cg2) {
    // fill_geometry() only picks pair mode for layers whose run-time epilogue is folded; the accumulator
    // tap runs the same CTA-pair main loop with the exact epilogue
    if (!fold && c.acc_dump == nullptr) return cudaErrorInvalidValue;
    kfn = fold ? T.pair[P.BN == 256 ? 1 : 0][P.mode][c.r != nullptr ? 1 : 0][hi32 ? 1 : 0]
               : T.pair_exact[P.BN == 256 ? 1 : 0][P.mode];
  }
  if (grp2) kfn = T.grp2[scaled_planes == 2 ? 1 : 0][c.r != nullptr ? 1 : 0][hi32 ? 1 : 0][0];
  if (!kfn) return cudaErrorInvalidValue;
  const TmapPair* tp = reinterpret_cast<const TmapPair*>(tmaps);
  P.dbg = nullptr;
  static const bool debug = env_int("TF2B_MMA_DEBUG", 0) != 0;
  static long long* dbg_dev = nullptr;
  if (debug) {
    if (!dbg_dev) cudaMalloc(&dbg_dev, sizeof(long long) * (8 * 148 + 3 * 64 + 16));
    cudaMemsetAsync(dbg_dev, 0, sizeof(long long) * (8 * 148 + 3 * 64 + 16), stream);
    P.dbg = dbg_dev;
  }
  int grid = 0, num_tiles = 0;
  {
    num_tiles = P.m_tiles * P.n_tiles;
    grid = num_tiles < num_sms ? num_tiles : num_sms;
    if (P.b_resident) {
      // every CTA must keep one n-tile: grid = multiple of n_tiles
      grid = (grid / P.n_tiles) * P.n_tiles;
      if (grid < P.n_tiles) grid = P.n_tiles;
    }
    if (P.cg2) {
      // pairs of CTAs (one cluster per TPC pair); work items = pairs of m-tiles x n-tiles
      const int items = ((P.m_tiles + 1) / 2) * P.n_tiles;
      grid = 2 * std::min(items, num_sms / 2);
    }
    static const bool use_pdl = env_int("TF2B_MMA_PDL", 1) != 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (use_pdl) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      na++;
    }
    if (P.cg2) {
      attr[na].id = cudaLaunchAttributeClusterDimension;
      attr[na].val.clusterDim.x = 2;
      attr[na].val.clusterDim.y = 1;
      attr[na].val.clusterDim.z = 1;
      na++;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kfn, P, *tp);
    if (le != cudaSuccess) return le;
  }
  if (debug) {
    long long h[8 * 148 + 3 * 64 + 16];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, dbg_dev, sizeof h, cudaMemcpyDeviceToHost);
    double a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < grid; b++)
      for (int i = 0; i < 8; i++) a[i] += (double)h[b * 8 + i] / grid;
    const double tiles_per_cta = (double)num_tiles / grid;
    fprintf(stderr,
            "[mma dbg] C%d N%d k%d s%d OH%d mode%d BK%d BN%d P%d stages%d tiles/cta %.1f kiters %d | per tile clk: "
            "producer total %.0f (wait empty %.0f) | mma total %.0f (wait full %.0f, wait tmem-empty %.0f, issue %.0f) | "
            "epi total %.0f (wait tmem-full %.0f)\n",
            c.Cp, c.N, c.k, c.stride, c.OH, P.mode, P.BK, P.BN, P.planes, P.stages, tiles_per_cta,
            P.taps * P.kchunks, a[1] / tiles_per_cta, a[0] / tiles_per_cta, a[4] / tiles_per_cta, a[2] / tiles_per_cta,
            a[3] / tiles_per_cta, a[7] / tiles_per_cta, a[6] / tiles_per_cta, a[5] / tiles_per_cta);
    {
      const long long* tl = h + 8 * 148 + 3 * 64 + 8;
      fprintf(stderr, "[mma timeline] CTA 0 clk since entry: prologue done %lld | previous layer done %lld | first operands %lld | "
                      "last MMA issued %lld | last tile stored %lld | teardown sync %lld\n", tl[0], tl[1], tl[2], tl[3], tl[4], tl[5]);
    }
    static const bool trace = env_int("TF2B_MMA_TRACE", 0) != 0;
    if (trace) {
      fprintf(stderr, "[mma trace] CTA 0, per stage position: wait / wait+issue clk:");
      for (int it = 0; it < 64 && it < P.taps * P.kchunks; it++) {
        const long long n = h[8 * 148 + 3 * it + 2];
        if (n > 0) fprintf(stderr, " %lld/%lld", h[8 * 148 + 3 * it] / n, h[8 * 148 + 3 * it + 1] / n);
      }
      fprintf(stderr, "\n");
      const long long* f = h + 8 * 148 + 3 * 64;
      if (f[7] > 0)
        fprintf(stderr, "[mma ticks] per stage of CTA 0 (%lld stages): wait %lld | fence %lld | desc+elect %lld | mma#1 %lld | mma#2-4 %lld | "
                        "commit %lld | syncwarp %lld\n", f[7], f[0] / f[7], f[1] / f[7], f[2] / f[7], f[3] / f[7], f[4] / f[7],
                f[5] / f[7], f[6] / f[7]);
    }
  }
  return cudaGetLastError();
}

}  // namespace tf2b
