// Microbenchmark: tcgen05.ld (TMEM -> registers) throughput per SM for 4/8/16 warps and the
// 32x32b shapes x16/x32/x64.  The epilogue of wide-N small-K layers reads planes*4 bytes of TMEM per
// output element; this measures whether that read is the bound.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int X>
__device__ __forceinline__ unsigned ld(unsigned taddr);
template <>
__device__ __forceinline__ unsigned ld<16>(unsigned taddr) {
  unsigned v[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) : "r"(taddr));
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s ^= v[i];
  return s;
}
template <>
__device__ __forceinline__ unsigned ld<32>(unsigned taddr) {
  unsigned v[32];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
                 "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
                 "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) : "r"(taddr));
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) s ^= v[i];
  return s;
}
template <int X>
__global__ void __launch_bounds__(512, 1) k(int iters, int wait_every, long long* out, unsigned* sink) {
  __shared__ unsigned slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = slot + ((unsigned)((warp & 3) * 32) << 16);
  unsigned acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
    acc ^= ld<X>(tmem + ((i * X) & 511 & ~(X - 1)));
    if ((i % wait_every) == wait_every - 1) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  long long t1 = clock64();
  if ((threadIdx.x & 31) == 0) out[blockIdx.x * 16 + warp] = t1 - t0;
  if (acc == 0x12345678) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}
int main() {
  long long* d; unsigned* sink;
  cudaMalloc(&d, 148 * 16 * 8); cudaMalloc(&sink, 4);
  const int iters = 4096;
  for (int X : {16, 32})
    for (int warps : {4, 8, 16})
      for (int we : {1, 4}) {
        for (int rep = 0; rep < 2; rep++) {
          if (X == 16) k<16><<<148, warps * 32>>>(iters, we, d, sink); else k<32><<<148, warps * 32>>>(iters, we, d, sink);
          cudaDeviceSynchronize();
        }
        long long h[148 * 16]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
        double mx = 0; for (int b = 0; b < 148; b++) for (int w = 0; w < warps; w++) mx = h[b * 16 + w] > mx ? h[b * 16 + w] : mx;
        double bytes = (double)iters * warps * 32 * X * 4;
        printf("x%-3d warps %2d wait every %d : %8.0f clk  %6.1f B/clk/SM  (%s)\n", X, warps, we, mx, bytes / mx,
               cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
