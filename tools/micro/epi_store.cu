// Microbenchmark (for the next round): how should the epilogue write its int8 tile?  One CTA per SM,
// 16 warps; every "tile" is 128 rows x 128 bytes produced in registers exactly like the conv epilogue
// (warp (quarter q, slice s): lane = row 32q + lane, 32 bytes at column 32s), rows of a 256-byte-pitch
// tensor (an N = 256 layer), 64 tiles per CTA, tiles interleaved over the CTAs like the real kernel.
//   0  st.global.v8 per lane (today's direct path: one 32-byte sector per lane, 32 wavefronts / warp)
//   1  smem transpose per quarter (4 warps, named barrier): full 128-byte lines, st.global.v4
//   2  smem + ONE TMA store per quarter (box 128 B x 32 rows, SWIZZLE_128B), named barrier of 4 warps
//   3  smem + one TMA store PER WARP (box 32 B x 32 rows, no swizzle): no cross-warp synchronisation
// Reports cycles per tile and GB/s; the ncu LSU wavefront counters of each variant are the other half.
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
struct Maps { CUtensorMap q, w; };
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts128(unsigned a, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(unsigned a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void tma_store_2d(unsigned smem, const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem), "r"(c0), "r"(c1)
               : "memory");
}
template <int METHOD>
__global__ void __launch_bounds__(512, 1) k(const __grid_constant__ Maps maps, signed char* out, int pitch, int tiles, long long* clk) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const unsigned base = (smem_u32(smem) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = warp & 3, s = warp >> 2;
  const int row = 32 * q + lane;
  long long t0 = clock64();
  for (int t = 0; t < tiles; t++) {
    const long long m0 = ((long long)t * gridDim.x + blockIdx.x) * 128;
    uint4 lo = make_uint4(t, row, s, 1), hi = make_uint4(t, row, s, 2);
    if (METHOD == 0) {
      signed char* p = out + (m0 + row) * pitch + 32 * s;
      asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x),
                   "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
    } else if (METHOD == 1 || METHOD == 2) {
      // quarter tile [32 rows][128 B], SWIZZLE_128B pattern, double buffered
      const unsigned qt = base + (unsigned)((t & 1) * 16384 + q * 4096);
      if (METHOD == 2 && s == 0 && lane == 0 && t >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
      const unsigned ra = qt + (unsigned)lane * 128u, sw = (unsigned)(lane & 7);
      sts128(ra + (((2u * s) ^ sw) << 4), lo);
      sts128(ra + (((2u * s + 1) ^ sw) << 4), hi);
      if (METHOD == 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
      if (METHOD == 1) {
        // warp s of the quarter writes rows 8s..8s+7 as whole lines: lane = (row r = lane / 8, chunk c = lane % 8)
#pragma unroll
        for (int it = 0; it < 2; it++) {
          const int r = 8 * s + 4 * it + (lane >> 3), c = lane & 7;
          const uint4 v = lds128(qt + (unsigned)r * 128u + (((unsigned)c ^ (unsigned)(r & 7)) << 4));
          *reinterpret_cast<uint4*>(out + (m0 + 32 * q + r) * pitch + 16 * c) = v;
        }
      } else if (s == 0 && lane == 0) {
        tma_store_2d(qt, &maps.q, 0, (int)(m0 + 32 * q));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else {
      // per warp: [32 rows][32 B] dense, double buffered, no cross-warp synchronisation
      const unsigned wt = base + (unsigned)((t & 1) * 16384 + warp * 1024);
      if (lane == 0 && t >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncwarp();
      sts128(wt + (unsigned)lane * 32u, lo);
      sts128(wt + (unsigned)lane * 32u + 16u, hi);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(wt, &maps.w, 32 * s, (int)(m0 + 32 * q));
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (METHOD >= 2) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qr);
  EncodeFn enc = (EncodeFn)fnp;
  const int pitch = 256, tiles = 64, grid = 148;
  const size_t rows = (size_t)tiles * grid * 128;
  signed char* out; long long* d;
  cudaMalloc(&out, rows * pitch); cudaMalloc(&d, 148 * 8);
  Maps m;
  cuuint64_t dims[2] = {128, rows}, st[1] = {(cuuint64_t)pitch};
  cuuint32_t boxq[2] = {128, 32}, boxw[2] = {32, 32}, es[2] = {1, 1};
  CUresult r1 = enc(&m.q, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, out, dims, st, boxq, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CUresult r2 = enc(&m.w, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, out, dims, st, boxw, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r1 || r2) { printf("encode failed %d %d\n", (int)r1, (int)r2); return 1; }
  auto run = [&](auto kern, const char* name) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
    for (int rep = 0; rep < 3; rep++) { kern<<<grid, 512, 34 * 1024>>>(m, out, pitch, tiles, d); cudaDeviceSynchronize(); }
    cudaError_t e = cudaGetLastError();
    long long h[148]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; i++) avg += (double)h[i] / grid;
    printf("%-48s %8.0f clk/tile  %7.1f B/clk/SM  (%s)\n", name, avg / tiles, 16384.0 * tiles / avg, cudaGetErrorString(e));
  };
  run(k<0>, "0 st.global.v8 per lane");
  run(k<1>, "1 smem transpose per quarter + st.global.v4");
  run(k<2>, "2 smem + TMA store per quarter (128 B x 32)");
  run(k<3>, "3 smem + TMA store per warp (32 B x 32)");
  return 0;
}
