// Correctness probe: can a K-major swizzled UMMA A-descriptor start at an arbitrary ROW of a tile
// that TMA wrote (start address = base + shift * row_bytes), i.e. can one halo tile in shared memory
// serve all filter taps of a convolution as shifted views?  Tries SWIZZLE_64B (64-byte rows) and
// SWIZZLE_128B (128-byte rows), with the descriptor's base_offset field (bits 49..51) either 0 or
// (start_address >> 7) & mask.  B = selector matrix, so D[m][n] = A[m + shift][n] (+ 2 A[..][n+64]).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>
struct Maps { CUtensorMap a, b; };
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ unsigned long long make_desc(unsigned saddr, int RB, unsigned bo) {
  unsigned long long d = 0;
  d |= (unsigned long long)((saddr >> 4) & 0x3FFF);
  d |= (unsigned long long)1 << 16;
  d |= (unsigned long long)((8 * RB) >> 4) << 32;
  d |= (unsigned long long)1 << 46;
  d |= (unsigned long long)(bo & 7) << 49;
  d |= (unsigned long long)(RB == 128 ? 2 : 4) << 61;
  return d;
}
// out[shift][128][64] int32
__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ Maps maps, int RB, int nshift, unsigned bomask, int* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bars[2];
  __shared__ unsigned tmem_slot;
  const unsigned base = (smem_u32(smem) + 1023u) & ~1023u;
  const unsigned sb = base + 512 * RB;   // B tile behind the 512 A rows
  const unsigned full = smem_u32(&bars[0]), done = smem_u32(&bars[1]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(full));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(done));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = tmem_slot;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"(512 * RB + 64 * RB) : "memory");
    for (int h = 0; h < 2; h++)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(base + h * 256 * RB), "l"(&maps.a), "r"(full), "r"(0), "r"(h * 256) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sb), "l"(&maps.b), "r"(full), "r"(0), "r"(0) : "memory");
  }
  mbar_wait(full, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(64 >> 3) << 17) | (8u << 24);
  unsigned ph = 0;
  for (int s = 0; s < nshift; s++) {
    if (threadIdx.x == 0) {
      const unsigned sa = base + s * RB;
      const unsigned bo = (sa >> 7) & bomask;
      const unsigned long long da = make_desc(sa, RB, bo), db = make_desc(sb, RB, 0);
      for (int k = 0; k < RB / 32; k++)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(tmem), "l"(da + 2ull * k), "l"(db + 2ull * k), "r"(idesc), "r"(k ? 1u : 0u) : "memory");
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done) : "memory");
    }
    mbar_wait(done, ph);
    ph ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < 64; c += 16) {
      unsigned v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                     "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                   : "r"(tmem + ((unsigned)(warp * 32) << 16) + c));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; j++) out[(s * 128 + warp * 32 + lane) * 64 + c + j] = (int)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem) : "memory");
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  const int NSHIFT = 130;
  for (int RB : {64, 128}) {
    signed char* hA = (signed char*)malloc(512 * RB); signed char* hB = (signed char*)calloc(64 * RB, 1);
    for (int i = 0; i < 512 * RB; i++) hA[i] = (signed char)(((i * 2654435761u) >> 13) % 11) - 5;
    for (int n = 0; n < 64; n++) { hB[n * RB + n] = 1; if (RB == 128) hB[n * RB + n + 64] = 2; }
    signed char *A, *B; int* out;
    cudaMalloc(&A, 512 * RB); cudaMalloc(&B, 64 * RB); cudaMalloc(&out, NSHIFT * 128 * 64 * 4);
    cudaMemcpy(A, hA, 512 * RB, cudaMemcpyHostToDevice); cudaMemcpy(B, hB, 64 * RB, cudaMemcpyHostToDevice);
    Maps m;
    cuuint64_t da[2] = {(cuuint64_t)RB, 512}, st[1] = {(cuuint64_t)RB}, db[2] = {(cuuint64_t)RB, 64};
    cuuint32_t ba[2] = {(cuuint32_t)RB, 256}, bb[2] = {(cuuint32_t)RB, 64}, es[2] = {1, 1};
    const CUtensorMapSwizzle sw = RB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r1 = enc(&m.a, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, A, da, st, ba, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = enc(&m.b, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, B, db, st, bb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r1 || r2) { printf("encode failed %d %d\n", (int)r1, (int)r2); return 1; }
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    int* h = (int*)malloc(NSHIFT * 128 * 64 * 4);
    for (unsigned bomask : {0u, 7u, 3u, 1u}) {
      cudaMemset(out, 0xff, NSHIFT * 128 * 64 * 4);
      probe<<<1, 128, 576 * RB + 1024>>>(m, RB, NSHIFT, bomask, out);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, out, NSHIFT * 128 * 64 * 4, cudaMemcpyDeviceToHost);
      int good = 0; char line[256]; int pos = 0; memset(line, 0, sizeof line);
      for (int s = 0; s < NSHIFT; s++) {
        long bad = 0;
        for (int mrow = 0; mrow < 128; mrow++)
          for (int n = 0; n < 64; n++) {
            int exp = hA[(mrow + s) * RB + n] + (RB == 128 ? 2 * hA[(mrow + s) * RB + n + 64] : 0);
            if (h[(s * 128 + mrow) * 64 + n] != exp) bad++;
          }
        if (!bad) good++;
        if (s < 40) pos += snprintf(line + pos, sizeof line - pos, "%c", bad ? 'x' : '.');
      }
      printf("row_bytes=%3d base_offset=(addr>>7)&%u : %3d/%d shifts exact   first 40 shifts: %s  (%s)\n", RB, bomask, good, NSHIFT,
             line, cudaGetErrorString(e));
    }
  }
  return 0;
}
