// Pure issue rate of single instructions on one SM sub-partition (sm_100a): 4 warps per scheduler, 16 independent
// dependency chains per thread of ONE instruction each (the result feeds the same instruction again).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/alu_bench.cu -o tools/bin/alu_bench
#include <cstdio>
#include <cuda_fp16.h>
#include <stdint.h>
template <int OP>
__global__ void __launch_bounds__(512, 1) k(int iters, unsigned long long* out, uint32_t* sink, uint32_t seed) {
  uint32_t u[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) u[i] = seed * (threadIdx.x + 1) + i * 0x01010101u;
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (OP == 0) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(*reinterpret_cast<float*>(&u[i])) : "f"(1.0000001f));
      if (OP == 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(*reinterpret_cast<float*>(&u[i])) : "f"(__uint_as_float(u[(i + 1) & 15] | 1u)));
      if (OP == 2) asm volatile("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(u[i]) : "f"(__uint_as_float(u[i])));
      if (OP == 3) asm volatile("{ .reg .f16 lo, hi; mov.b32 {lo, hi}, %0; cvt.f32.f16 %0, lo; }" : "+r"(u[i]));
      if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[i]) : "r"(seed), "r"(u[(i + 1) & 15]));
      if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[i]) : "r"(u[(i + 1) & 15]));
      if (OP == 6) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %1;" : "=r"(u[i]) : "f"(__uint_as_float(u[i])));
      if (OP == 7) asm volatile("add.f32 %0, %0, %1;" : "+f"(*reinterpret_cast<float*>(&u[i])) : "f"(1.0000001f));
      if (OP == 8) asm volatile("shf.r.clamp.b32 %0, %0, %1, 13;" : "+r"(u[i]) : "r"(u[(i + 1) & 15]));
      if (OP == 9) asm volatile("cvt.rz.relu.f16x2.f32 %0, %1, %1;" : "=r"(u[i]) : "f"(__uint_as_float(u[i])));
    }
  }
  const unsigned long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) acc ^= u[i];
  if (acc == 12345u) sink[threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}
template <int OP> void run(const char* name, unsigned long long* d_out, uint32_t* d_sink) {
  const int iters = 4000;
  k<OP><<<148, 512>>>(iters, d_out, d_sink, 7u);
  cudaDeviceSynchronize();
  k<OP><<<148, 512>>>(iters, d_out, d_sink, 7u);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return; }
  unsigned long long c = 0;
  cudaMemcpy(&c, d_out, 8, cudaMemcpyDeviceToHost);
  // 4 warps x 16 instructions per iteration on one scheduler
  printf("%-34s %6.2f scheduler cycles per warp-instruction\n", name, (double)c / iters / 64.0);
}
int main() {
  unsigned long long* d_out; uint32_t* d_sink;
  cudaMalloc(&d_out, 64); cudaMalloc(&d_sink, 4096);
  run<0>("FFMA", d_out, d_sink); run<7>("FADD", d_out, d_sink); run<1>("FMNMX", d_out, d_sink);
  run<4>("LOP3", d_out, d_sink); run<5>("PRMT", d_out, d_sink); run<8>("SHF", d_out, d_sink);
  run<2>("F2FP.F16.F32.PACK_AB", d_out, d_sink); run<9>("F2FP.RZ.RELU.F16.F32.PACK_AB", d_out, d_sink);
  run<6>("F2FP.BF16.F32.PACK_AB", d_out, d_sink); run<3>("HADD2.F32 (f16 -> f32)", d_out, d_sink);
  return 0;
}
