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
__global__ void raw224_to_s2d_kernel(const int8_t* __restrict__ raw, int8_t* __restrict__ dst,
                                     int B, int dual) {
  const int OD = 114, ID = 224;
  size_t total = (size_t)B * OD * OD;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % OD);
    size_t t = i / OD;
    int r = (int)(t % OD);
    int b = (int)(t / OD);
    __align__(16) unsigned char out[32];
#pragma unroll
    for (int e = 27; e < 32; e++) out[e] = 0;
#pragma unroll
    for (int ci = 0; ci < 3; ci++) {
      const int8_t* plane = raw + ((size_t)b * 3 + ci) * ID * ID;
#pragma unroll
      for (int d = 0; d < 9; d++) {
        int row = (d < 6) ? (2 * r + (d & 1)) : (2 * r + 2);
        int col = 2 * c + ((d < 6) ? (d >> 1) : (d - 6));
        row -= 3;
        col -= 3;
        unsigned char v = 0;
        if (row >= 0 && row < ID && col >= 0 && col < ID) v = (unsigned char)plane[row * ID + col];
        out[ci * 9 + d] = v;
      }
    }
    uint4* o = reinterpret_cast<uint4*>(dst + i * (dual ? 64 : 32));
    uint4 v0 = *reinterpret_cast<const uint4*>(out);
    uint4 v1 = *reinterpret_cast<const uint4*>(out + 16);
    o[0] = v0;
    o[1] = v1;
    if (dual) {
      o[2] = make_uint4(__vneg4(v0.x), __vneg4(v0.y), __vneg4(v0.z), __vneg4(v0.w));
      o[3] = make_uint4(__vneg4(v1.x), __vneg4(v1.y), __vneg4(v1.z), __vneg4(v1.w));
    }
  }
}

// pool.cl:178-260 + pool_tail.cl:91-216: 3x3 max, window of output j starts at j*ps - ppad, taps
// outside the map contribute 0; optional residual add (feature_writer.cl:124-127).
// One thread per (output pixel, 16-channel chunk).
__global__ void maxpool3x3_kernel(const int8_t* __restrict__ src, int8_t* __restrict__ dst,
                                  const int8_t* __restrict__ res, int B, int H, int W, int sC,
                                  int PH, int PW, int dC, int rC, int C, int ps, int ppad,
                                  int add_relu) {
  const int chunks = (C + 15) / 16;
  size_t total = (size_t)B * PH * PW * chunks;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    int ck = (int)(i % chunks);
    size_t pix = i / chunks;
    int pw = (int)(pix % PW);
    size_t t = pix / PW;
    int ph = (int)(t % PH);
    int b = (int)(t / PH);
    // identity of signed max is -128 in every byte
    uint4 m = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
#pragma unroll
    for (int dh = 0; dh < 3; dh++) {
#pragma unroll
      for (int dw = 0; dw < 3; dw++) {
        int h = ph * ps - ppad + dh, w = pw * ps - ppad + dw;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (h >= 0 && h < H && w >= 0 && w < W)
          v = *reinterpret_cast<const uint4*>(src + (((size_t)b * H + h) * W + w) * sC + ck * 16);
        m.x = __vmaxs4(m.x, v.x);
        m.y = __vmaxs4(m.y, v.y);
        m.z = __vmaxs4(m.z, v.z);
        m.w = __vmaxs4(m.w, v.w);
      }
    }
    if (res != nullptr) {
      uint4 rv = *reinterpret_cast<const uint4*>(res + pix * rC + ck * 16);
      m.x = add_res4(m.x, rv.x, add_relu);
      m.y = add_res4(m.y, rv.y, add_relu);
      m.z = add_res4(m.z, rv.z, add_relu);
      m.w = add_res4(m.w, rv.w, add_relu);
    }
    int8_t* d = dst + pix * dC + ck * 16;
    int nvalid = min(16, C - ck * 16);
    if (nvalid == 16) {
      *reinterpret_cast<uint4*>(d) = m;
    } else {
      const unsigned char* mb = reinterpret_cast<const unsigned char*>(&m);
      for (int e = 0; e < nvalid; e++) d[e] = (int8_t)mb[e];
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
  size_t total = (size_t)B * 114 * 114;
  raw224_to_s2d_kernel<<<grid_for(total, 128), 128, 0, s>>>(raw, dst, B, dual);
  return cudaGetLastError();
}
cudaError_t launch_maxpool3x3(const int8_t* src, int8_t* dst, const int8_t* res, int B, int H, int W,
                              int sC, int PH, int PW, int dC, int rC, int C, int ps, int ppad,
                              int add_relu, cudaStream_t s) {
  size_t total = (size_t)B * PH * PW * ((C + 15) / 16);
  maxpool3x3_kernel<<<grid_for(total, 256), 256, 0, s>>>(src, dst, res, B, H, W, sC, PH, PW, dC, rC,
                                                        C, ps, ppad, add_relu);
  return cudaGetLastError();
}
cudaError_t launch_gap(const int8_t* src, int8_t* dst, int B, int HW, int sC, int dC, int C,
                       cudaStream_t s) {
  size_t total = (size_t)B * C;
  gap_kernel<<<grid_for(total, 256), 256, 0, s>>>(src, dst, B, HW, sC, dC, C);
  return cudaGetLastError();
}

}  // namespace tf2b
