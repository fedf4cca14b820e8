// Probe: register/thread layout of tcgen05.ld.16x256b.x4 (and .x2) — which (TMEM lane, column) lands in
// which (thread, register).  TMEM is filled with lane*1000 + column through tcgen05.st.32x32b.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) k(int* out) {
  __shared__ unsigned slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = slot + ((unsigned)(warp * 32) << 16);
  // fill: this thread's TMEM lane (warp*32 + lane), columns 0..63
  for (int c0 = 0; c0 < 64; c0 += 16) {
    unsigned v[16];
    for (int j = 0; j < 16; j++) v[j] = (warp * 32 + lane) * 1000 + c0 + j;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(tmem + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                   "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // read rows [16*h, 16*h+16) of this warp's quarter, columns 0..31, with 16x256b.x4
  for (int h = 0; h < 2; h++) {
    unsigned r[16];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tmem + ((unsigned)(16 * h) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; j++) out[((warp * 2 + h) * 32 + lane) * 16 + j] = (int)r[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(slot) : "memory");
}
int main() {
  int* d; cudaMalloc(&d, 4 * 2 * 32 * 16 * 4);
  k<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  static int h[4 * 2 * 32 * 16];
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  printf("%s\n", cudaGetErrorString(e));
  for (int w = 0; w < 2; w++)
    for (int hh = 0; hh < 2; hh++)
      for (int lane = 0; lane < 32; lane += (lane < 8 ? 1 : 7)) {
        printf("warp %d half %d lane %2d:", w, hh, lane);
        for (int j = 0; j < 16; j++) printf(" %6d", h[((w * 2 + hh) * 32 + lane) * 16 + j]);
        printf("\n");
      }
  // check the conjectured layout: reg j of lane i = row (16h + i/4 + 8*((j>>1)&1)), col 8*(j>>2) + 2*(i%4) + (j&1)
  long bad = 0;
  for (int w = 0; w < 4; w++) for (int hh = 0; hh < 2; hh++) for (int i = 0; i < 32; i++) for (int j = 0; j < 16; j++) {
    int row = w * 32 + 16 * hh + i / 4 + 8 * ((j >> 1) & 1), col = 8 * (j >> 2) + 2 * (i % 4) + (j & 1);
    if (h[((w * 2 + hh) * 32 + i) * 16 + j] != row * 1000 + col) bad++;
  }
  printf("conjecture (m16n8 C-fragment per 8-column group) mismatches: %ld\n", bad);
  return 0;
}
