// Microbenchmark: issue rate of tcgen05.mma kind::i8 (M=128, N in {64,128,256}, K=32) from smem
// operands, no TMA traffic.  Prints cycles per MMA per SM.  Build: nvcc -arch=sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  unsigned pred = 0, laneid = 0;
  asm volatile("{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %2;\n\t@%%px mov.s32 %1, 1;\n\tmov.s32 %0, %%rx;\n\t}"
               : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFF));
  return pred != 0;
}
__device__ __forceinline__ unsigned long long make_desc(unsigned saddr, unsigned sbo16, unsigned lt) {
  unsigned long long d = 0;
  d |= (unsigned long long)((saddr >> 4) & 0x3FFF);
  d |= (unsigned long long)1 << 16;
  d |= (unsigned long long)(sbo16 & 0x3FFF) << 32;
  d |= (unsigned long long)1 << 46;
  d |= (unsigned long long)(lt & 7) << 61;
  return d;
}
template <int KIND>  // 0 = i8, 1 = f16 (bf16 inputs, K=16)
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int iters, int nacc, long long* out, int sw64) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ unsigned tmem_slot;
  const unsigned base = (smem_u32(smem) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) ((unsigned*)smem)[i] = 0x01010101u * (i & 3);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = tmem_slot;
  unsigned idesc;
  if (KIND == 0) idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | (8u << 24);
  else idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | (8u << 24);   // f32 acc, bf16 a/b
  if (warp == 0) {
    long long t0 = clock64();
    // sw64: SWIZZLE_64B operands (64-byte rows, 8-row groups 512 B apart), else SWIZZLE_128B
    const unsigned long long da = sw64 ? make_desc(base, 32, 4) : make_desc(base, 64, 2);
    const unsigned long long db = sw64 ? make_desc(base + 16384, 32, 4) : make_desc(base + 16384, 64, 2);
    const int kmask = sw64 ? 1 : 3;
    if (elect_one()) {
      for (int i = 0; i < iters; i++) {
        const unsigned d = tmem + (unsigned)((i % nacc) * N);
        const unsigned long long ka = da + (unsigned long long)(2 * (i & kmask));
        const unsigned long long kb = db + (unsigned long long)(2 * (i & kmask));
        if (KIND == 0)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(ka), "l"(kb), "r"(idesc), "r"(i >= nacc ? 1u : 0u) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(d), "l"(ka), "l"(kb), "r"(idesc), "r"(i >= nacc ? 1u : 0u) : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    unsigned ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
int main() {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 4096;
  for (int kind = 0; kind < 2; kind++)
    for (int N : {64, 128, 256})
      for (int nacc : {1, 2}) {
        for (int grid : {1, 148}) {
         for (int sw64 : {0, 1}) {
          if (kind == 0) rate_kernel<0><<<grid, 128, 80 * 1024>>>(N, iters, nacc, d, sw64);
          else rate_kernel<1><<<grid, 128, 80 * 1024>>>(N, iters, nacc, d, sw64);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[148];
          cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
          double avg = 0;
          for (int i = 0; i < grid; i++) avg += (double)h[i] / grid;
          printf("%s N=%3d nacc=%d grid=%3d %s : %.1f clk/MMA  (%s)\n", kind ? "bf16" : "i8  ", N, nacc, grid,
                 sw64 ? "SW64 " : "SW128", avg / iters, cudaGetErrorString(e));
         }
        }
      }
  return 0;
}
