// aux_kernels.cu — bandwidth-bound helpers around the convolution kernels: boundary layout
// conversion, the layer-0 space-to-depth transform, 3x3 max pooling, global average pooling.
// All work on NHWC int8 tensors with a 16-byte-multiple channel pitch and use 16-byte accesses
// wherever the channel dimension allows it.
#include "common.cuh"

namespace tf2b {

namespace {

// [B][C][H][W] -> [B][H][W][Cp]  (channels >= C are written as zero)
// neg_off > 0: channels [neg_off, neg_off + C) receive the int8-negated copy (-(-128) = -128,
// pe.cl:32-34) that the tensor-core path multiplies with the magnitudes of the negative weights
__global__ void chw_to_hwc_kernel(const int8_t* __restrict__ src, int8_t* __restrict__ dst, int B,
                                  int C, int H, int W, int Cp, int neg_off) {
  size_t total = (size_t)B * H * W * (Cp / 4);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    int c4 = (int)(i % (Cp / 4));
    size_t pix = i / (Cp / 4);
    int w = (int)(pix % W);
    size_t t = pix / W;
    int h = (int)(t % H);
    int b = (int)(t / H);
    unsigned v = 0;
#pragma unroll
    for (int e = 0; e < 4; e++) {
      int c = c4 * 4 + e;
      bool ng = neg_off > 0 && c >= neg_off;
      if (ng) c -= neg_off;
      if (c < C) {
        unsigned char x = (unsigned char)src[(((size_t)b * C + c) * H + h) * W + w];
        if (ng) x = (unsigned char)(0u - x);
        v |= (unsigned)x << (8 * e);
      }
    }
    reinterpret_cast<unsigned*>(dst)[i] = v;
  }
}

// [B][H][W][Cp] (first C channels) -> [B][C][H][W]
// pos != nullptr: logical channel c is stored at position pos[c]
__global__ void hwc_to_chw_kernel(const int8_t* __restrict__ src, int8_t* __restrict__ dst, int B,
                                  int C, int H, int W, int Cp, const int* __restrict__ pos) {
  size_t total = (size_t)B * C * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    int w = (int)(i % W);
    size_t t = i / W;
    int h = (int)(t % H);
    t /= H;
    int c = (int)(t % C);
    int b = (int)(t / C);
    dst[i] = src[(((size_t)b * H + h) * W + w) * Cp + (pos ? pos[c] : c)];
  }
}

// [B][H][W][Cs] (first C channels) -> [B][H][W][Cd] dense copy with re-pitch (zero fill)
__global__ void hwc_repitch_kernel(const int8_t* __restrict__ src, int8_t* __restrict__ dst,
                                   size_t npix, int C, int Cs, int Cd, int neg_off, const int* __restrict__ pos) {
  size_t total = npix * Cd;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % Cd);
    size_t pix = i / Cd;
    bool ng = neg_off > 0 && c >= neg_off;
    if (ng) c -= neg_off;
    unsigned char x = c < C ? (unsigned char)src[pix * Cs + (pos ? pos[c] : c)] : (unsigned char)0;
    if (ng) x = (unsigned char)(0u - x);
    dst[i] = (int8_t)x;
  }
}

// input_loader.cpp:27-73 (feature_trans) applied to an already quantised int8 image:
// raw [B][3][224][224] -> tensor 0 [B][114][114][32]; derived channel 9*ci + d of pixel (r,c) is
//   d <  6: P[2r + d%2][2c + d/2]     d >= 6: P[2r + 2][2c + d-6]      P = raw zero-padded by 3
// channels 27..31 are zero.  Quantise-then-permute equals the reference's permute-then-quantise
// (runner.cpp:158-164 is elementwise and maps the padding zeros to zero).
// With dual != 0 the pixel is 64 bytes: channels 32..58 hold the int8-negated copy (pe.cl:32-34).
// One CTA per (image, output row): the nine source rows (3 colour planes x raw rows 2r-3 .. 2r-1) are staged in
// shared memory with 16-byte coalesced loads (zero borders included), every thread then assembles one output pixel
// from shared bytes and writes its 32 / 64 bytes with 16-byte stores (a warp covers 2 KB contiguous).
constexpr int kS2dPitch = 240;   // 224 source bytes at offset 8 (column -3 lands on index 5), zero borders
__global__ void __launch_bounds__(128) raw224_to_s2d_kernel(const int8_t* __restrict__ raw, int8_t* __restrict__ dst,
                                                            int B, int dual) {
  const int OD = 114, ID = 224;
  __shared__ __align__(16) unsigned char rows[9][kS2dPitch];
  const int r = blockIdx.x, b = blockIdx.y;
  (void)B;
  // 9 rows x 15 16-byte words: words 0 and 14 are borders (columns < 0 / >= 224), word 1 + j holds source bytes 16j..
  for (int i = threadIdx.x; i < 9 * 15; i += blockDim.x) {
    const int q = i / 15, wd = i - q * 15;
    const int ci = q / 3, rs = q - ci * 3;
    const int row = 2 * r + rs - 3;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (wd >= 1 && wd <= 14 && row >= 0 && row < ID) {
      const int8_t* sp = raw + (((size_t)b * 3 + ci) * ID + row) * ID + (wd - 1) * 16;
      if (dual & 2) {   // caller's buffer is not 16-byte aligned: byte loads
        unsigned ww[4] = {0, 0, 0, 0};
        for (int e = 0; e < 16; e++) ww[e >> 2] |= (unsigned)(unsigned char)sp[e] << (8 * (e & 3));
        v = make_uint4(ww[0], ww[1], ww[2], ww[3]);
      } else {
        v = *reinterpret_cast<const uint4*>(sp);
      }
    }
    // word wd sits at byte offset 16*wd - 8: the 8 leading bytes are the left border
    if (wd == 0) {
      *reinterpret_cast<uint2*>(&rows[q][0]) = make_uint2(0, 0);
    } else if (wd <= 14) {
      *reinterpret_cast<uint2*>(&rows[q][16 * wd - 8]) = make_uint2(v.x, v.y);
      *reinterpret_cast<uint2*>(&rows[q][16 * wd]) = make_uint2(v.z, v.w);
    }
    if (wd == 14) *reinterpret_cast<uint2*>(&rows[q][kS2dPitch - 8]) = make_uint2(0, 0);
  }
  __syncthreads();
  const int c = threadIdx.x;
  if (c >= OD) return;
  // source column 2c - 3 + k lives at index 2c + 5 + k
  unsigned w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int ci = 0; ci < 3; ci++) {
#pragma unroll
    for (int d = 0; d < 9; d++) {
      const int rs = (d < 6) ? (d & 1) : 2;
      const int k = (d < 6) ? (d >> 1) : (d - 6);
      const unsigned v = rows[ci * 3 + rs][2 * c + 5 + k];
      const int e = ci * 9 + d;
      w[e >> 2] |= v << (8 * (e & 3));
    }
  }
  // one whole 32-byte sector per lane and store (two 16-byte stores per sector doubled the L1 -> L2 write sectors)
  int8_t* o = dst + (((size_t)b * OD + r) * OD + c) * ((dual & 1) ? 64 : 32);
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
               "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
  if (dual & 1)
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o + 32), "r"(__vneg4(w[0])), "r"(__vneg4(w[1])),
                 "r"(__vneg4(w[2])), "r"(__vneg4(w[3])), "r"(__vneg4(w[4])), "r"(__vneg4(w[5])), "r"(__vneg4(w[6])),
                 "r"(__vneg4(w[7]))
                 : "memory");
}

// four packed int8 -> two s16x2 words (sign-extending PRMT), and back (values are int8 again after max)
__device__ __forceinline__ void s8x4_to_s16x2(unsigned v, unsigned& lo, unsigned& hi) {
  asm("prmt.b32 %0, %1, %1, 0x9180;" : "=r"(lo) : "r"(v));
  asm("prmt.b32 %0, %1, %1, 0xb3a2;" : "=r"(hi) : "r"(v));
}
__device__ __forceinline__ unsigned max_s16x2(unsigned a, unsigned b) {
  unsigned d;
  asm("max.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

// pool.cl:178-260 + pool_tail.cl:91-216: 3x3 max, window of output j starts at j*ps - ppad, taps
// outside the map contribute 0; optional residual add (feature_writer.cl:124-127).
// One thread per (output pixel, 16-channel chunk); grid = (row chunks, output rows, images), so the only division is
// pixel / chunk inside a row.  The maximum runs in int16 lanes (PRMT sign extension + VIMNMX.S16x2: 4 instructions
// per 4 values and tap; the packed-byte __vmaxs4 is emulated with ~20).
__global__ void maxpool3x3_kernel(const int8_t* __restrict__ src, int8_t* __restrict__ dst,
                                  const int8_t* __restrict__ res, int B, int H, int W, int sC,
                                  int PH, int PW, int dC, int rC, int C, int ps, int ppad,
                                  int add_relu, int chunk_shift) {
  (void)B;
  const int chunks = (C + 15) / 16;
  const int row_items = PW * chunks;
  const int ph = blockIdx.y, b = blockIdx.z;
  const int h0 = ph * ps - ppad;
  // one 64-bit product per block; everything inside an image is 32-bit (the launcher checks the sizes)
  const int8_t* img = src + (size_t)b * H * W * sC;
  int8_t* dimg = dst + (size_t)b * PH * PW * dC;
  const int8_t* rimg = res != nullptr ? res + (size_t)b * PH * PW * rC : nullptr;
  const int rs = W * sC;
  const bool rows_in = h0 >= 0 && h0 + 2 < H;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < row_items; t += gridDim.x * blockDim.x) {
    const int pw = chunk_shift >= 0 ? (t >> chunk_shift) : t / chunks;
    const int ck = t - pw * chunks;
    const int w0 = pw * ps - ppad;
    // identity of the signed maximum: -128 in every lane
    unsigned lo[4] = {0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u};
    unsigned hi[4] = {0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u};
    bool outside = false;
    auto take = [&](const int8_t* p) {
      const uint4 v = *reinterpret_cast<const uint4*>(p);
      const unsigned vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int q = 0; q < 4; q++) {
        unsigned l, hh;
        s8x4_to_s16x2(vv[q], l, hh);
        lo[q] = max_s16x2(lo[q], l);
        hi[q] = max_s16x2(hi[q], hh);
      }
    };
    if (rows_in && w0 >= 0 && w0 + 2 < W) {
      // interior window (almost every thread): nine unconditional loads off one base pointer
      const int8_t* p0 = img + ((h0 * W + w0) * sC + ck * 16);
#pragma unroll
      for (int dh = 0; dh < 3; dh++) {
        const int8_t* pr = p0 + dh * rs;
        take(pr);
        take(pr + sC);
        take(pr + 2 * sC);
      }
    } else {
#pragma unroll
      for (int dh = 0; dh < 3; dh++) {
        const int h = h0 + dh;
        const bool hv = h >= 0 && h < H;
        const int8_t* rowp = img + ((hv ? h : 0) * rs + ck * 16);
#pragma unroll
        for (int dw = 0; dw < 3; dw++) {
          const int w = w0 + dw;
          if (hv && w >= 0 && w < W) take(rowp + w * sC);
          else outside = true;
        }
      }
    }
    unsigned mm[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      if (outside) {   // taps outside the map contribute 0
        lo[q] = max_s16x2(lo[q], 0u);
        hi[q] = max_s16x2(hi[q], 0u);
      }
      asm("prmt.b32 %0, %1, %2, 0x6420;" : "=r"(mm[q]) : "r"(lo[q]), "r"(hi[q]));
    }
    uint4 m = make_uint4(mm[0], mm[1], mm[2], mm[3]);
    const int pix = ph * PW + pw;   // inside the image
    if (rimg != nullptr) {
      uint4 rv = *reinterpret_cast<const uint4*>(rimg + (pix * rC + ck * 16));
      m.x = add_res4(m.x, rv.x, add_relu);
      m.y = add_res4(m.y, rv.y, add_relu);
      m.z = add_res4(m.z, rv.z, add_relu);
      m.w = add_res4(m.w, rv.w, add_relu);
    }
    int8_t* d = dimg + (pix * dC + ck * 16);
    int nvalid = min(16, C - ck * 16);
    if (nvalid == 16) {
      *reinterpret_cast<uint4*>(d) = m;
    } else {
      const unsigned mw[4] = {m.x, m.y, m.z, m.w};
      for (int e = 0; e < nvalid; e++) d[e] = (int8_t)((mw[e >> 2] >> (8 * (e & 3))) & 0xffu);
    }
  }
}

// full_size_pool.cl:95-119: int16 (wrapping) sum over the map, ((sum*669 >> 14) + 1) >> 1, clamp.
// One thread per (image, channel); consecutive threads read consecutive channels (coalesced).
__global__ void gap_kernel(const int8_t* __restrict__ src, int8_t* __restrict__ dst, int B, int HW,
                           int sC, int dC, int C) {
  size_t total = (size_t)B * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int b = (int)(i / C);
    int sum = 0;
    const int8_t* s = src + (size_t)b * HW * sC + c;
    for (int p = 0; p < HW; p++) sum += s[(size_t)p * sC];
    sum = (int)(short)sum;
    int y = (((sum * 669) >> 14) + 1) >> 1;
    y = max(-128, min(127, y));
    dst[(size_t)b * dC + c] = (int8_t)y;
  }
}

inline int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  size_t cap = 148 * 32;  // enough CTAs to fill every SM several times over; grid-stride beyond
  return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace

cudaError_t launch_chw_to_hwc(const int8_t* src, int8_t* dst, int B, int C, int H, int W, int Cp,
                              int neg_off, cudaStream_t s) {
  size_t total = (size_t)B * H * W * (Cp / 4);
  chw_to_hwc_kernel<<<grid_for(total, 256), 256, 0, s>>>(src, dst, B, C, H, W, Cp, neg_off);
  return cudaGetLastError();
}
cudaError_t launch_hwc_to_chw(const int8_t* src, int8_t* dst, int B, int C, int H, int W, int Cp,
                              const int* pos, cudaStream_t s) {
  size_t total = (size_t)B * C * H * W;
  hwc_to_chw_kernel<<<grid_for(total, 256), 256, 0, s>>>(src, dst, B, C, H, W, Cp, pos);
  return cudaGetLastError();
}
cudaError_t launch_hwc_repitch(const int8_t* src, int8_t* dst, size_t npix, int C, int Cs, int Cd,
                               int neg_off, const int* pos, cudaStream_t s) {
  size_t total = npix * Cd;
  hwc_repitch_kernel<<<grid_for(total, 256), 256, 0, s>>>(src, dst, npix, C, Cs, Cd, neg_off, pos);
  return cudaGetLastError();
}
cudaError_t launch_raw224_to_s2d(const int8_t* raw, int8_t* dst, int B, int dual, cudaStream_t s) {
  if (B <= 0) return cudaSuccess;
  if (B > 65535) return cudaErrorInvalidValue;
  // bit 0 of the kernel's flag word: 64-byte pixels with the negated copy; bit 1: source not 16-byte aligned
  const int flags = (dual ? 1 : 0) | ((reinterpret_cast<uintptr_t>(raw) & 15u) ? 2 : 0);
  raw224_to_s2d_kernel<<<dim3(114, B), 128, 0, s>>>(raw, dst, B, flags);
  return cudaGetLastError();
}
cudaError_t launch_maxpool3x3(const int8_t* src, int8_t* dst, const int8_t* res, int B, int H, int W,
                              int sC, int PH, int PW, int dC, int rC, int C, int ps, int ppad,
                              int add_relu, cudaStream_t s) {
  const int row_items = PW * ((C + 15) / 16);
  if (B <= 0 || PH <= 0 || row_items <= 0) return cudaSuccess;
  if (B > 65535 || PH > 65535) return cudaErrorInvalidValue;
  const int block = row_items >= 256 ? 256 : (row_items + 31) / 32 * 32;
  int gx = (row_items + block - 1) / block;
  if (gx > 64) gx = 64;   // (grid-stride within the row beyond that)
  // offsets inside one image are 32-bit in the kernel
  if ((long long)H * W * sC >= (1ll << 31) || (long long)PH * PW * (dC > rC ? dC : rC) >= (1ll << 31)) return cudaErrorInvalidValue;
  const int chunks = (C + 15) / 16;
  int chunk_shift = -1;
  for (int sh = 0; sh < 16; sh++)
    if ((1 << sh) == chunks) chunk_shift = sh;
  maxpool3x3_kernel<<<dim3(gx, PH, B), block, 0, s>>>(src, dst, res, B, H, W, sC, PH, PW, dC, rC, C, ps, ppad,
                                                      add_relu, chunk_shift);
  return cudaGetLastError();
}
cudaError_t launch_gap(const int8_t* src, int8_t* dst, int B, int HW, int sC, int dC, int C,
                       cudaStream_t s) {
  size_t total = (size_t)B * C;
  gap_kernel<<<grid_for(total, 256), 256, 0, s>>>(src, dst, B, HW, sC, dC, C);
  return cudaGetLastError();
}

}  // namespace tf2b
