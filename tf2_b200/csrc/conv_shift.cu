// conv_shift.cu — kernel A: CUDA-core shift-accumulate convolution (exact for every layer shape).
//
// Reference semantics: PeFunction's MAC  result += sum_c MUL(feature, code)  with
// MUL = (+-feature) << shift  (Runtime_Engine/cnn/device/src/pe.cl:27-49,144-180), zero padding and
// stride geometry of sequencer.cl:268-311, then the requantisation of pe.cl:185-203, ReLU
// (relu.cl:54) and the residual add of feature_writer.cl:124-127 fused into the epilogue.
//
// B200 mapping: implicit GEMM  M = B*OH*OW pixels, N = output channels, K = k*k*Cp.  A shift by s
// is a multiply by 2^s modulo 2^32, so each code becomes an int16 weight +-2^e (e = shift minus the
// per-output-channel base shift, 0..14) and the MAC is the 2-way int16 x int8 dot product DP2A into
// an int32 register that wraps exactly like the FPGA accumulator.  The per-channel base shift is
// applied to the finished sum (also exact modulo 2^32).  Codes whose e exceeds 14 go to further
// planes; the int8 negate quirk (-(-128) = -128) is a plane that multiplies the byte-negated
// activations.  Tiles: 128 pixels x 64 channels per CTA, 32 bytes of K per stage, cp.async double
// buffering with zero-fill for padding, 16-byte coalesced stores from a staged int8 tile.
#include "common.cuh"

namespace tf2b {

namespace {

constexpr int BM = 128;        // pixels per CTA
constexpr int BN = 64;         // output channels per CTA
constexpr int KC = 32;         // K bytes (channels) per pipeline stage
constexpr int A_STRIDE = 48;   // bytes per A row in smem (32 data + 16 pad)
constexpr int B_STRIDE = 80;   // bytes per B row in smem (32 int16 = 64 data + 16 pad)
constexpr int C_STRIDE = BN + 16;
constexpr int NTHREADS = 256;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__global__ void __launch_bounds__(NTHREADS, 2) conv_shift_kernel(const ConvParams p, const int16_t* __restrict__ wgt) {
  __shared__ __align__(16) unsigned char As[2][BM * A_STRIDE];
  __shared__ __align__(16) unsigned char Bs[2][BN * B_STRIDE];
  __shared__ __align__(16) unsigned char Cs[BM * C_STRIDE];

  const int t = threadIdx.x;
  const int m0 = blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int M = p.B * p.OH * p.OW;

  // ---- loader roles ----
  // A: thread -> (row = t/2, half = t%2): one 16-byte cp.async per stage
  const int a_row = t >> 1, a_half = t & 1;
  const int am = m0 + a_row;
  const bool a_row_ok = am < M;
  int ab = 0, aoh = 0, aow = 0;
  if (a_row_ok) {
    ab = am / (p.OH * p.OW);
    int rem = am - ab * p.OH * p.OW;
    aoh = rem / p.OW;
    aow = rem - aoh * p.OW;
  }
  const int ih0 = aoh * p.stride - p.pad, iw0 = aow * p.stride - p.pad;
  const int8_t* ximg = p.x + (size_t)ab * p.IH * p.IW * p.xC;
  // B: thread -> (row = t/4, seg = t%4): 8 int16 per cp.async
  const int b_row = t >> 2, b_seg = t & 3;

  // ---- compute roles ----
  const int ty = t >> 4, tx = t & 15;

  int total[8][4];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) total[i][j] = 0;

  const int nstages = p.Kp / KC;

  for (int pl = 0; pl < p.planes; pl++) {
    const int16_t* wpl = wgt + (size_t)pl * p.Npad * p.Kp;
    const bool neg = p.plane_neg[pl] != 0;
    int acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) acc[i][j] = 0;

    auto load_stage = [&](int st, int buf) {
      // A half: flattened k index -> (tap, channel offset)
      int kk = st * KC + a_half * 16;
      bool ok = a_row_ok && kk < p.Ktot;
      const int8_t* src = p.x;
      if (ok) {
        int tap = kk / p.Cp;
        int c0 = kk - tap * p.Cp;
        int fh = tap / p.k, fw = tap - fh * p.k;
        int ih = ih0 + fh, iw = iw0 + fw;
        ok = (ih >= 0) && (ih < p.IH) && (iw >= 0) && (iw < p.IW);
        if (ok) src = ximg + ((size_t)ih * p.IW + iw) * p.xC + c0;
      }
      cp_async16(&As[buf][a_row * A_STRIDE + a_half * 16], src, ok);
      const int16_t* wsrc = wpl + (size_t)(n0 + b_row) * p.Kp + st * KC + b_seg * 8;
      cp_async16(&Bs[buf][b_row * B_STRIDE + b_seg * 16], wsrc, true);
    };

    load_stage(0, 0);
    cp_async_commit();
    for (int st = 0; st < nstages; st++) {
      const int buf = st & 1;
      if (st + 1 < nstages) load_stage(st + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
      __syncthreads();
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int4 bl[4], bh[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const unsigned char* bp = &Bs[buf][(tx + 16 * j) * B_STRIDE + h * 32];
          bl[j] = *reinterpret_cast<const int4*>(bp);
          bh[j] = *reinterpret_cast<const int4*>(bp + 16);
        }
#pragma unroll
        for (int i = 0; i < 8; i++) {
          int4 a = *reinterpret_cast<const int4*>(&As[buf][(ty + 16 * i) * A_STRIDE + h * 16]);
          if (neg) {  // pe.cl:32-34: negate inside int8 (wraps: -(-128) = -128)
            a.x = __vneg4(a.x); a.y = __vneg4(a.y); a.z = __vneg4(a.z); a.w = __vneg4(a.w);
          }
#pragma unroll
          for (int j = 0; j < 4; j++) {
            int s = acc[i][j];
            s = __dp2a_lo(bl[j].x, a.x, s);
            s = __dp2a_hi(bl[j].y, a.x, s);
            s = __dp2a_lo(bl[j].z, a.y, s);
            s = __dp2a_hi(bl[j].w, a.y, s);
            s = __dp2a_lo(bh[j].x, a.z, s);
            s = __dp2a_hi(bh[j].y, a.z, s);
            s = __dp2a_lo(bh[j].z, a.w, s);
            s = __dp2a_hi(bh[j].w, a.w, s);
            acc[i][j] = s;
          }
        }
      }
      __syncthreads();
    }
    cp_async_wait<0>();
    const int psh = p.plane_shift[pl];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) total[i][j] += (int)((unsigned)acc[i][j] << psh);
  }

  // ---- epilogue: bias seed + base shift, requant, ReLU -> staged int8 tile ----
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int n = n0 + tx + 16 * j;  // < Npad always
    const int bias = p.bias[n], alpha = p.alpha[n], beta = p.beta[n];
    const int nsh = p.nshift[n];
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int row = ty + 16 * i;
      int a32 = (int)((unsigned)bias + ((unsigned)total[i][j] << nsh));
      if (p.acc_dump != nullptr) {
        int m = m0 + row;
        if (m < M && n < p.N) {
          int b = m / (p.OH * p.OW);
          int rem = m - b * p.OH * p.OW;
          p.acc_dump[((size_t)b * p.N + n) * p.OH * p.OW + rem] = a32;
        }
      }
      int y = requant(a32, alpha, beta);
      if (p.relu) y = max(y, 0);
      Cs[row * C_STRIDE + tx + 16 * j] = (unsigned char)(signed char)y;
    }
  }
  __syncthreads();
  // ---- coalesced store (+ residual add): 16-byte chunks ----
  for (int ch = t; ch < BM * (BN / 16); ch += NTHREADS) {
    const int row = ch >> 2, seg = ch & 3;
    const int m = m0 + row;
    const int nb = n0 + seg * 16;
    if (m >= M || nb >= p.N) continue;
    uint4 v = *reinterpret_cast<const uint4*>(&Cs[row * C_STRIDE + seg * 16]);
    int8_t* dst = p.y + (size_t)m * p.yC + nb;
    const int nvalid = min(16, p.N - nb);
    if (nvalid == 16) {
      if (p.r != nullptr) {
        uint4 rv = *reinterpret_cast<const uint4*>(p.r + (size_t)m * p.rC + nb);
        v.x = add_res4(v.x, rv.x, p.add_relu);
        v.y = add_res4(v.y, rv.y, p.add_relu);
        v.z = add_res4(v.z, rv.z, p.add_relu);
        v.w = add_res4(v.w, rv.w, p.add_relu);
      }
      *reinterpret_cast<uint4*>(dst) = v;
    } else {
      const unsigned char* vb = &Cs[row * C_STRIDE + seg * 16];
      for (int e = 0; e < nvalid; e++) {
        int yv = (signed char)vb[e];
        if (p.r != nullptr) {
          int s = yv + (int)p.r[(size_t)m * p.rC + nb + e];
          s = max(-128, min(127, s));
          if (p.add_relu) s = max(s, 0);
          yv = s;
        }
        dst[e] = (int8_t)yv;
      }
    }
  }
}

}  // namespace

// host launcher --------------------------------------------------------------------------------
cudaError_t launch_conv_shift(const ConvParams& p, const int16_t* wgt, cudaStream_t stream) {
  const int M = p.B * p.OH * p.OW;
  dim3 grid((M + BM - 1) / BM, p.Npad / BN);
  conv_shift_kernel<<<grid, NTHREADS, 0, stream>>>(p, wgt);
  return cudaGetLastError();
}

int conv_shift_bn() { return BN; }
int conv_shift_kc() { return KC; }

}  // namespace tf2b
