// Pipe-throughput microbenchmarks (diagnostics): cycles per warp-instruction for MUFU.EX2, F2FP (cvt.rn.bf16x2), FFMA,
// FMNMX with W warps per SM sub-partition.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3f + i * 0.01f;
  unsigned acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) { unsigned r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 15])); acc ^= r; }
      if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (OP == 3) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 15]));
      if (OP == 4) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 5) { unsigned r = __float_as_uint(a[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r)); a[i] = __uint_as_float(r); }   // 2 exps per instruction
      if (OP == 6) { unsigned r = __float_as_uint(a[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r)); a[i] = __uint_as_float(r); }
      if (OP == 7) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 15]), "f"(a[(i + 2) & 15]));
      if (OP == 8) { unsigned r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 15])); acc ^= r; }
    }
  }
  long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const char* names[] = {"MUFU.EX2", "F2FP.BF16x2", "FFMA", "FMNMX", "MUFU.TANH", "EX2.BF16x2", "EX2.F16x2", "FMNMX3", "F2FP.F16x2"};
  const int iters = 2000;
  for (int op = 0; op < 9; ++op)
    for (int warps = 4; warps <= 16; warps *= 2) {   // 4 warps = 1 per SMSP
      for (int rep = 0; rep < 2; ++rep) {
        if (op == 0) k<0><<<148, warps * 32>>>(out, cyc, iters);
        if (op == 1) k<1><<<148, warps * 32>>>(out, cyc, iters);
        if (op == 2) k<2><<<148, warps * 32>>>(out, cyc, iters);
        if (op == 3) k<3><<<148, warps * 32>>>(out, cyc, iters);
        if (op == 4) k<4><<<148, warps * 32>>>(out, cyc, iters);
        if (op == 5) k<5><<<148, warps * 32>>>(out, cyc, iters);
        if (op == 6) k<6><<<148, warps * 32>>>(out, cyc, iters);
        if (op == 7) k<7><<<148, warps * 32>>>(out, cyc, iters);
        if (op == 8) k<8><<<148, warps * 32>>>(out, cyc, iters);
        cudaDeviceSynchronize();
      }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      double per_warp_inst = double(c) / (iters * 16.0);
      printf("%-12s warps/SM %2d (%d per SMSP): %.2f cycles per warp-instruction per warp  => %.2f cycles/inst/SMSP\n", names[op], warps,
             warps / 4, per_warp_inst, per_warp_inst / (warps / 4));
    }
  return 0;
}
