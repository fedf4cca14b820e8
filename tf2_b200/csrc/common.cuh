// common.cuh — shared device/host definitions of the tf2_b200 CUDA engine (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace tf2b {

constexpr int kMaxPlanes = 4;

// Geometry + operands of one convolution launch (both kernel families use it).
// Activations are NHWC int8 with a per-pixel pitch in bytes (`xC`, `yC`, `rC`); channel counts that
// take part in the reduction are padded to a multiple of 16 (`Cp`) and the padded weights are zero.
struct ConvParams {
  const int8_t* x;        // input tensor, [B][IH][IW][xC]
  int8_t* y;              // output tensor base incl. the concat channel offset, [B][OH][OW][yC]
  const int8_t* r;        // residual operand [B][OH][OW][rC] or nullptr (feature_writer.cl:124-127)
  const int32_t* bias;    // [Npad]  BiasBnParam.bias  (accumulator seed, pe.cl:176-180)
  const int32_t* alpha;   // [Npad]
  const int32_t* beta;    // [Npad]
  const uint8_t* nshift;  // [Npad]  per-output-channel base shift factored out of the codes
  int32_t* acc_dump;      // optional [B][N][OH][OW] int32 accumulators (debug tap)
  const int* acc_perm;    // tap: logical channel of output row n (the tensor's stored channel order), or nullptr
  int B, IH, IW, Cp, xC;
  int OH, OW, N, Npad, yC, rC;
  int k, pad, stride;
  int Ktot;               // k*k*Cp   (flattened reduction length, tap-major, channel-minor)
  int Kp;                 // Ktot rounded up to the kernel's K step (weights are zero beyond Ktot)
  int relu, add_relu;
  int planes;             // number of weight planes
  int plane_shift[kMaxPlanes];  // plane p contributes (sum_p << plane_shift[p])
  int plane_neg[kMaxPlanes];    // 1: plane multiplies the int8-negated activations (pe.cl:32-34)
  int low_plane;                // tensor-core path: plane added without the per-channel 2^nshift, or -1
  int fast_requant;             // load-time range analysis proved that no int32 intermediate of the
                                // requantisation can wrap: the fused 64-bit form is exact
  int w4_avail;                 // tensor-core path: a packed 4-bit copy of the weight planes exists (resident-weight layers)
  int sparse2;                  // tensor-core path, two planes: issue per plane, skip the (tap, 32-channel block, plane)
  unsigned char blkmask[320];   // combinations that hold no weights: [tap][K chunk], 2 bits per block of the chunk
};

// pe.cl:185-203 — requantisation of one accumulator (int64 product, arithmetic shifts, clamp).
__device__ __forceinline__ int requant(int32_t acc, int32_t alpha, int32_t beta) {
  long long t = (long long)acc * (long long)alpha;
  int a = (int)(t >> 20);
  int s = (int)((unsigned)a + (unsigned)beta);
  int y = ((s >> 14) + 1) >> 1;
  return max(-128, min(127, y));
}

// Same with the ReLU of relu.cl:54 folded into the lower clamp bound (lo = 0 with ReLU, else -128).
__device__ __forceinline__ int requant_clamped(int32_t acc, int32_t alpha, int32_t beta, int lo) {
  long long t = (long long)acc * (long long)alpha;
  int a = (int)(t >> 20);
  int s = (int)((unsigned)a + (unsigned)beta);
  int y = ((s >> 14) + 1) >> 1;
  return max(lo, min(127, y));
}

// feature_writer.cl:124-127 on 4 packed int8 lanes: saturating add, optional ReLU.
__device__ __forceinline__ unsigned add_res4(unsigned y, unsigned r, int add_relu) {
  unsigned s = __vaddss4(y, r);
  if (add_relu) s = __vmaxs4(s, 0u);
  return s;
}

}  // namespace tf2b
