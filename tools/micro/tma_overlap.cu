// Does TMA accept a row stride smaller than the row length (overlapping rows)?  dims {128 B, rows},
// row stride 64 B, box {128, 8}; prints the first bytes of each loaded row and what was expected.
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>
__global__ void k(const __grid_constant__ CUtensorMap m, unsigned char* out, int row0) {
  __shared__ __align__(1024) unsigned char buf[8 * 128];
  __shared__ __align__(8) unsigned long long bar;
  unsigned b = (unsigned)__cvta_generic_to_shared(&bar), s = (unsigned)__cvta_generic_to_shared(buf);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(8 * 128) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(s), "l"(&m), "r"(b), "r"(0), "r"(row0) : "memory");
    unsigned ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b) : "memory");
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * 128; i += blockDim.x) out[i] = buf[i];
}
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const int n = 64 * 1200;
  unsigned char* h = new unsigned char[n];
  for (int i = 0; i < n; i++) h[i] = (unsigned char)(i % 251);
  unsigned char *d, *o;
  cudaMalloc(&d, n); cudaMalloc(&o, 1024); cudaMemcpy(d, h, n, cudaMemcpyHostToDevice);
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fnp;
  for (int sw = 0; sw < 2; sw++) {
    CUtensorMap m;
    cuuint64_t dims[2] = {128, 1000}, strides[1] = {64};
    cuuint32_t box[2] = {128, 8}, es[2] = {1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("swizzle %d encode result %d\n", sw, (int)r);
    if (r != CUDA_SUCCESS) continue;
    k<<<1, 128>>>(m, o, 5);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned char res[1024];
    cudaMemcpy(res, o, 1024, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r8 = 0; r8 < 8; r8++)
      for (int j = 0; j < 128; j++) {
        int chunk = j / 16, within = j % 16;
        int pos = sw ? r8 * 128 + ((chunk ^ (r8 & 7)) * 16) + within : r8 * 128 + j;
        unsigned char exp = h[(5 + r8) * 64 + j];
        if (res[pos] != exp) bad++;
      }
    printf("  launch %s, mismatching bytes %d of 1024; row1 first bytes got %d %d exp %d %d\n", cudaGetErrorString(e), bad,
           res[sw ? 128 + 16 : 128], res[sw ? 128 + 17 : 129], h[6 * 64], h[6 * 64 + 1]);
  }
  return 0;
}
