// Microbenchmark: TMA-fed tcgen05 i8 mainloop without epilogue.  One CTA per SM: warp 0 = TMA
// producer (A tile 128 x BK, B tile BN x BK, SWIZZLE_128B), warp 1 = MMA issuer.  Reports cycles
// per MMA (M128 x N=BN x K32) for several BN / stage counts, grid = 1 and 148, and for operands
// streamed from distinct rows (L2/HBM traffic) vs the same tile every time (L2 hits only).
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
struct Maps { CUtensorMap a, b; };
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  unsigned pred = 0, laneid = 0;
  asm volatile("{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %2;\n\t@%%px mov.s32 %1, 1;\n\tmov.s32 %0, %%rx;\n\t}"
               : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFF));
  return pred != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ unsigned long long make_desc(unsigned saddr, int BK) {
  unsigned long long d = 0;
  d |= (unsigned long long)((saddr >> 4) & 0x3FFF);
  d |= (unsigned long long)1 << 16;
  d |= (unsigned long long)(BK == 128 ? 64 : 32) << 32;
  d |= (unsigned long long)1 << 46;
  d |= (unsigned long long)(BK == 128 ? 2 : 4) << 61;
  return d;
}
template <int P, int K4, int AOPS, int BOPS, int VAR = 0>
__global__ void __launch_bounds__(64, 1) mainloop(const __grid_constant__ Maps maps, int BN, int stages, int kiters,
                                                  int stream_rows, int krange, long long* out, int BK, int planes) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[2 * 8 + 1];
  __shared__ unsigned tmem_slot;
  const unsigned base = (smem_u32(smem) + 1023u) & ~1023u;
  const int stage_bytes = 128 * BK + planes * BN * BK;
  const int tx_bytes = (AOPS ? 128 * BK : 0) + BOPS * BN * BK;
  const unsigned full = smem_u32(&bars[0]), empty = smem_u32(&bars[8]), done = smem_u32(&bars[16]);
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; s++) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full + 8 * s));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty + 8 * s));
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(done));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = tmem_slot;
  const unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(BN >> 3) << 17) | (8u << 24);
  long long t0 = clock64();
  if (warp == 0) {
    int stage = 0; unsigned phase = 0;
    for (int it = 0; it < kiters; it++) {
      mbar_wait(empty + 8 * stage, phase ^ 1);
      if (elect_one()) {
        const unsigned fb = full + 8 * stage;
        if (AOPS + BOPS) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(tx_bytes) : "memory");
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fb) : "memory");
        const unsigned sa = base + stage * stage_bytes;
        const int k0 = (it % krange) * BK;
        const int row = stream_rows ? (int)(blockIdx.x * 128 + (it / krange) % 4 * 148 * 128) : 0;
#pragma unroll
        for (int a = 0; a < AOPS; a++)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(sa + a * (128 / (AOPS ? AOPS : 1)) * BK), "l"(&maps.a), "r"(fb), "r"(k0), "r"(row + a * (128 / (AOPS ? AOPS : 1))) : "memory");
#pragma unroll
        for (int pl = 0; pl < BOPS; pl++)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(sa + 128 * BK + pl * BN * BK), "l"(&maps.b), "r"(fb), "r"(k0), "r"(0) : "memory");
      }
      __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
  } else if (VAR == 2) {
    // single-thread MMA issuer: no elect / syncwarp
    if ((threadIdx.x & 31) == 0) {
      int stage = 0; unsigned phase = 0;
      for (int it = 0; it < kiters; it++) {
        mbar_wait(full + 8 * stage, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned sa = base + stage * stage_bytes;
        const unsigned long long da = make_desc(sa, BK);
#pragma unroll
        for (int pl = 0; pl < P; pl++) {
          const unsigned long long db = make_desc(sa + 128 * BK + pl * BN * BK, BK);
#pragma unroll
          for (int k4 = 0; k4 < K4; k4++)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem + pl * BN), "l"(da + 2ull * (k4 & 3)), "l"(db + 2ull * (k4 & 3)), "r"(idesc), "r"((it | k4) ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty + 8 * stage) : "memory");
        if (it == kiters - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done) : "memory");
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    int stage = 0; unsigned phase = 0;
    for (int it = 0; it < kiters; it++) {
      mbar_wait(full + 8 * stage, phase);
      if (VAR != 1) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const unsigned sa = base + stage * stage_bytes;
      const unsigned long long da = make_desc(sa, BK);
      if (elect_one()) {
#pragma unroll
        for (int pl = 0; pl < P; pl++) {
          const unsigned long long db = make_desc(sa + 128 * BK + pl * BN * BK, BK);
#pragma unroll
          for (int k4 = 0; k4 < K4; k4++)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem + pl * BN), "l"(da + 2ull * (k4 & 3)), "l"(db + 2ull * (k4 & 3)), "r"(idesc), "r"((it | k4) ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty + 8 * stage) : "memory");
        if (it == kiters - 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done) : "memory");
      }
      if (VAR != 3) __syncwarp();
      if (++stage == stages) { stage = 0; phase ^= 1; }
    }
    mbar_wait(done, 0);
    if ((threadIdx.x & 31) == 0) out[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const size_t K = 4096, MA = 148 * 128 * 4, MB = 256;
  unsigned char *A, *B; long long* d;
  cudaMalloc(&A, MA * K); cudaMalloc(&B, MB * K); cudaMalloc(&d, 148 * 8);
  cudaMemset(A, 1, MA * K); cudaMemset(B, 1, MB * K);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  auto run = [&](auto kern, const char* name, int BN, int BK, int planes, int aops) {
    Maps m;
    cuuint64_t da[2] = {K, MA}, sa[1] = {K}, db[2] = {K, MB};
    cuuint32_t ba[2] = {(cuuint32_t)BK, (cuuint32_t)(aops ? 128 / aops : 128)}, bb[2] = {(cuuint32_t)BK, (cuuint32_t)BN}, es[2] = {1, 1};
    const CUtensorMapSwizzle sw = BK == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    enc(&m.a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, A, da, sa, ba, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    enc(&m.b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, B, db, sa, bb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    const int stages = 4, stage_bytes = 128 * BK + planes * BN * BK, kiters = 2048;
    for (int stream_rows : {0, 1}) {
      for (int rep = 0; rep < 2; rep++) {
        kern<<<148, 64, stages * stage_bytes + 1024>>>(m, BN, stages, kiters, stream_rows, 4096 / BK, d, BK, planes);
        cudaDeviceSynchronize();
      }
      cudaError_t e = cudaGetLastError();
      long long h[148]; cudaMemcpy(h, d, 148 * 8, cudaMemcpyDeviceToHost);
      double avg = 0; for (int i = 0; i < 148; i++) avg += (double)h[i] / 148;
      printf("%-28s BN=%3d BK=%3d P=%d %s : %.0f clk/k-iter (%s)\n", name, BN, BK, planes, stream_rows ? "stream" : "same  ",
             avg / kiters, cudaGetErrorString(e));
    }
  };
  run(mainloop<1, 4, 0, 0, 0>, "noTMA K4=4 base", 128, 128, 1, 0);
  run(mainloop<1, 4, 0, 0, 1>, "noTMA K4=4 nofence", 128, 128, 1, 0);
  run(mainloop<1, 4, 0, 0, 2>, "noTMA K4=4 lane0", 128, 128, 1, 0);
  run(mainloop<1, 4, 0, 0, 3>, "noTMA K4=4 nosyncwarp", 128, 128, 1, 0);
  run(mainloop<1, 2, 0, 0, 0>, "noTMA K4=2 base", 128, 128, 1, 0);
  run(mainloop<1, 8, 0, 0, 0>, "noTMA K4=8 base", 128, 128, 1, 0);
  run(mainloop<1, 16, 0, 0, 0>, "noTMA K4=16 base", 128, 128, 1, 0);
  run(mainloop<1, 8, 0, 0, 2>, "noTMA K4=8 lane0", 128, 128, 1, 0);
  run(mainloop<1, 4, 0, 0, 0>, "noTMA K4=4 base", 256, 128, 1, 0);
  run(mainloop<1, 8, 0, 0, 0>, "noTMA K4=8 base", 256, 128, 1, 0);
  run(mainloop<1, 4, 0, 0, 0>, "noTMA K4=4 base", 64, 128, 1, 0);
  run(mainloop<1, 8, 0, 0, 0>, "noTMA K4=8 base", 64, 128, 1, 0);
  run(mainloop<1, 4, 1, 1, 2>, "A+B lane0", 128, 128, 1, 1);
  run(mainloop<1, 4, 1, 1, 2>, "A+B lane0", 256, 128, 1, 1);
  return 0;
}
