// Microbenchmark: TMA fill rate into shared memory (no MMA) as a function of the row length and of
// the tensor-map rank.  One CTA per SM, one thread issues loads into a 4-deep ring and waits for
// them; the source is small (L2 resident), so this measures the TMA engine / smem write path:
// cycles per box and per row for flat 2-D [pixels][C] boxes vs 4-D (C, W, H, B) convolution boxes.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__global__ void __launch_bounds__(32, 1) fill(const __grid_constant__ CUtensorMap map, int rank, int box_bytes, int iters,
                                              int wmax, int hmax, int bmax, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[4];
  const unsigned base = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t0 = clock64();
    unsigned ph[4] = {0, 0, 0, 0};
    for (int it = 0; it < iters + 4; it++) {
      const int s = it & 3;
      if (it >= 4) { mbar_wait(smem_u32(&bars[s]), ph[s]); ph[s] ^= 1; }
      if (it < iters) {
        const unsigned fb = smem_u32(&bars[s]);
        const unsigned dst = base + s * ((box_bytes + 1023) & ~1023);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(box_bytes) : "memory");
        const int j = it + blockIdx.x * 7;
        if (rank == 2) {
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(dst), "l"(&map), "r"(fb), "r"(0), "r"((j * 64) % wmax) : "memory");
        } else {
          // conv taps: start coordinates -1..1 around a tile origin
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                       ::"r"(dst), "l"(&map), "r"(fb), "r"(0), "r"(j % 3 - 1), "r"((j / 3) % hmax - 1), "r"((j / 9) % bmax) : "memory");
        }
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  unsigned char* A; long long* d;
  const size_t bytes = 64u << 20;
  cudaMalloc(&A, bytes); cudaMemset(A, 1, bytes); cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(fill, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 4000;
  auto report = [&](const char* name, int rows, int rb) {
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += (double)h[i] / 148;
    printf("%-44s rows %3d x %3d B : %7.0f clk/box  %5.2f clk/row  %5.1f B/clk/SM  (%s)\n", name, rows, rb, avg / iters,
           avg / iters / rows, (double)rows * rb / (avg / iters), cudaGetErrorString(e));
  };
  // flat 2-D
  for (int rb : {128, 64, 32}) {
    for (int rows : {128, 256}) {
      CUtensorMap m;
      cuuint64_t dims[2] = {(cuuint64_t)rb, 65536}, st[1] = {(cuuint64_t)rb};
      cuuint32_t box[2] = {(cuuint32_t)rb, (cuuint32_t)rows}, es[2] = {1, 1};
      const CUtensorMapSwizzle sw = rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
      CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, A, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r) { printf("encode failed %d\n", (int)r); continue; }
      for (int rep = 0; rep < 2; rep++) fill<<<148, 32, 4 * rows * rb + 2048>>>(m, 2, rows * rb, iters, 65536 - 256, 1, 1, d);
      report("flat 2-D [pixels][C]", rows, rb);
    }
  }
  // flat 2-D with a wider pitch (row = first rb bytes of a 256-byte pixel)
  for (int rb : {128, 64}) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)rb, 65536}, st[1] = {256};
    cuuint32_t box[2] = {(cuuint32_t)rb, 128}, es[2] = {1, 1};
    const CUtensorMapSwizzle sw = rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, A, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    for (int rep = 0; rep < 2; rep++) fill<<<148, 32, 4 * 128 * rb + 2048>>>(m, 2, 128 * rb, iters, 65536 - 256, 1, 1, d);
    report("flat 2-D, pitch 256", 128, rb);
  }
  // 4-D convolution boxes (C, W, H, B)
  struct Cfg { const char* name; int C, W, H, B, bc, bw, bh, bn, es; } cfgs[] = {
      {"4-D box 56x56x64  (64,56,2,1)", 64, 56, 56, 64, 64, 56, 2, 1, 1},
      {"4-D box 56x56x64  (64,28,4,1)", 64, 56, 56, 64, 64, 28, 4, 1, 1},
      {"4-D box 28x28x128 (128,28,4,1)", 128, 28, 28, 64, 128, 28, 4, 1, 1},
      {"4-D box 14x14x256 (128,14,7,1)", 256, 14, 14, 64, 128, 14, 7, 1, 1},
      {"4-D box 7x7x512   (128,7,7,2)", 512, 7, 7, 64, 128, 7, 7, 2, 1},
      {"4-D box 56x56x128 stride 2 (128,28,4,1)", 128, 56, 56, 64, 128, 55, 7, 1, 2},
      {"4-D box 114x114x64 (64,112,1,1)", 64, 114, 114, 32, 64, 112, 1, 1, 1},
  };
  for (auto& c : cfgs) {
    CUtensorMap m;
    cuuint64_t dims[4] = {(cuuint64_t)c.C, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.B};
    cuuint64_t st[3] = {(cuuint64_t)c.C, (cuuint64_t)c.C * c.W, (cuuint64_t)c.C * c.W * c.H};
    cuuint32_t box[4] = {(cuuint32_t)c.bc, (cuuint32_t)c.bw, (cuuint32_t)c.bh, (cuuint32_t)c.bn};
    cuuint32_t es[4] = {1, (cuuint32_t)c.es, (cuuint32_t)c.es, 1};
    const CUtensorMapSwizzle sw = c.bc == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, A, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode failed %d for %s\n", (int)r, c.name); continue; }
    const int rows = ((c.bw + c.es - 1) / c.es) * ((c.bh + c.es - 1) / c.es) * c.bn;
    for (int rep = 0; rep < 2; rep++)
      fill<<<148, 32, 4 * ((rows * c.bc + 1023) & ~1023) + 2048>>>(m, 4, rows * c.bc, iters, 1, 3, c.B - 2, d);
    report(c.name, rows, c.bc);
  }
  return 0;
}
