// Microbenchmark: issue rate of the integer dot-product instructions kernel A is built on (IDP.4A = dp4a,
// IDP.2A = dp2a) against IMAD, per SM sub-partition, with independent accumulator chains (no dependency
// stalls) and 1 / 2 / 4 / 8 warps per sub-partition.  Sets the ALU roofline of the CUDA-core shift-accumulate
// path: images/s <= SMs x clk x rate x MACs-per-instruction / (MACs per image x planes).
//   op 0: IDP.4A (4 int8 x int8 MACs per lane)   op 1: IDP.2A (2 int16 x int8)   op 2: IMAD   op 3: IDP.4A + LOP3 mix
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(int* out, int iters, long long* clk, int a0, int b0) {
  int acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = i + threadIdx.x;
  int a = a0 + threadIdx.x, b = b0 ^ threadIdx.x, x = threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      if (OP == 0) acc[i] = __dp4a(a, b + i, acc[i]);
      if (OP == 1) acc[i] = __dp2a_lo(a, b + i, acc[i]);
      if (OP == 2) acc[i] = acc[i] * a + b;
      if (OP == 3) { acc[i] = __dp4a(a, b + i, acc[i]); x = (x ^ acc[(i + 8) & 15]) & (b | i); }
    }
  }
  long long t1 = clock64();
  int s = x;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
int main() {
  int* out; long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&clk, 148 * 8);
  const int iters = 4096;
  const char* names[4] = {"IDP.4A", "IDP.2A", "IMAD", "IDP.4A+LOP3"};
  for (int op = 0; op < 4; op++)
    for (int warps = 4; warps <= 32; warps *= 2) {
      for (int rep = 0; rep < 2; rep++) {
        if (op == 0) k<0><<<148, warps * 32>>>(out, iters, clk, 3, 5);
        if (op == 1) k<1><<<148, warps * 32>>>(out, iters, clk, 3, 5);
        if (op == 2) k<2><<<148, warps * 32>>>(out, iters, clk, 3, 5);
        if (op == 3) k<3><<<148, warps * 32>>>(out, iters, clk, 3, 5);
      }
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, clk, sizeof h, cudaMemcpyDeviceToHost);
      double c = 0; for (int i = 0; i < 148; i++) c += (double)h[i] / 148;
      const double winstr = (double)iters * 16 * warps;        // warp-instructions of the op per SM
      printf("%-12s warps/SM %2d: %.0f clk, %.3f warp-instr/clk/SM, %.3f per SMSP, %.1f lanes/clk/SM\n", names[op], warps, c,
             winstr / c, winstr / c / 4, 32 * winstr / c);
    }
  printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
