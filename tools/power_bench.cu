// What the 1000 W cap lets the whole chip sustain: tcgen05.mma throughput on all 148 SMs for seconds at a time, alone and
// next to CUDA-core work / mbarrier polling on the other warps.  The average SM clock of a run is cycles / time.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nerfds_b200/csrc -I include tools/power_bench.cu -o tools/bin/power_bench
//   variant  N    side load on 16 warps
//   0        128  none (they exit)
//   1        256  none
//   2        128  dependent FFMA chains (4 per thread), all the time
//   3        128  mbarrier.try_wait polling of a phase that never completes
//   4        128  none, MMA duty 50 % (the issuer idles as long as a burst takes)
//   5        64   none
//   6        128  FFMA at ~50 % duty
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "nds_tc.cuh"

using namespace nds::tc;

__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\t"
      "elect.sync %%rx|%%px, %2;\n\t"
      "@%%px mov.s32 %1, 1;\n\t"
      "mov.s32 %0, %%rx;\n\t}"
      : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFF));
  return pred;
}

__global__ void __launch_bounds__(640, 1) power_kernel(int variant, int bursts, unsigned long long* out, float* sink, int issuer, int c0, int c1) {
  // issuer: warp that issues the MMAs; side-load warps: [c0, c1)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, never, dummy;
  __shared__ uint32_t tmem_base_s;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  // random operands: the multipliers' switching activity (and with it the power) depends on the data
  auto rnd_h2 = [](uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    const float a = (float)(x & 0xffff) / 32768.f - 1.f, b = (float)(x >> 16) / 32768.f - 1.f;
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
  };
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = rnd_h2(i * 2654435761u + blockIdx.x);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&never, 1); mbar_init(&dummy, 1); mbar_fence_init(); stop = 0; }
  if (warp == issuer) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tmem_base_s;
  if (warp >= 4 && warp < 20 && false) {}
  if ((issuer == 0 ? warp >= 4 : warp < 16)) {      // A operand region (columns 256..383) and zeroed accumulators (columns 0..255)
    const int q = warp & 3, sub = (issuer == 0 ? warp - 4 : warp) >> 2;
    const uint32_t lane_addr = tb + (((uint32_t)q * 32u) << 16);
    uint32_t v[16], z[16];
    for (int i = 0; i < 16; ++i) { v[i] = rnd_h2(threadIdx.x * 977u + i * 131u + blockIdx.x * 7919u); z[i] = 0u; }
    tmem_st<16>(lane_addr + 256u + (uint32_t)sub * 32u, v);
    tmem_st<16>(lane_addr + 256u + (uint32_t)sub * 32u + 16u, v);
    for (int c = 0; c < 4; ++c) tmem_st<16>(lane_addr + (uint32_t)sub * 64u + (uint32_t)c * 16u, z);
    tmem_st_wait();
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const int N = variant == 1 ? 256 : (variant == 5 ? 64 : 128);
  if (warp == issuer) {
    const uint32_t idesc = make_idesc_f16(N);
    const uint32_t b_lo32 = smem_desc_lo32(smem_u32(smem));
    const unsigned long long t0 = clock64();
    for (int i = 0; i < bursts; ++i) {
      if (elect_one_sync()) {
        const uint64_t bd0 = ((uint64_t)NDS_DESC_HI << 32) | (b_lo32 + (uint32_t)(i & 1) * 2048u), bd1 = bd0 + 1024u;
        const uint32_t d = tb + (uint32_t)(i & 1) * (N == 256 ? 0u : 128u);
        const uint32_t A0 = tb + 256 + (uint32_t)(i & 1) * 64, A1 = A0 + 16;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ts(d, A0 + 8 * k, bd0 + 2 * k, idesc, 1u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ts(d, A1 + 8 * k, bd0 + 2 * k, idesc, 1u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ts(d, A0 + 8 * k, bd1 + 2 * k, idesc, 1u);
        umma_commit(&dummy);
        if (variant == 4) {            // idle as long as the burst takes
          const unsigned long long w0 = clock64();
          while (clock64() - w0 < 12ull * 66ull) {}
        }
      }
      __syncwarp();
    }
    if (elect_one_sync()) { umma_commit(&bar); mbar_wait(&bar, 0); }
    __syncwarp();
    const unsigned long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) { if (blockIdx.x == 0) out[0] = t1 - t0; stop = 1; }
  } else if (warp >= c0 && warp < c1) {
    if (variant == 2 || variant == 6) {
      float a = threadIdx.x * 1e-3f, b = a + 1.f, c = a + 2.f, d = a + 3.f;
      const float m = 1.0000001f, s = 1e-7f;
      int it = 0;
      while (!stop) {
#pragma unroll
        for (int k = 0; k < 64; ++k) { a = fmaf(a, m, s); b = fmaf(b, m, s); c = fmaf(c, m, s); d = fmaf(d, m, s); }
        if (variant == 6) { const unsigned long long w0 = clock64(); while (clock64() - w0 < 300ull) {} }
        ++it;
      }
      if (a + b + c + d == 12345.f) sink[threadIdx.x] = a + it;
    } else if (variant == 3) {
      while (!stop) { (void)mbar_try_wait(&never, 0); }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == issuer) tmem_dealloc(tb, 512);
}

int main(int argc, char** argv) {
  const double seconds = argc > 1 ? atof(argv[1]) : 2.0;
  unsigned long long* d_out;
  float* d_sink;
  cudaMalloc(&d_out, 64);
  cudaMalloc(&d_sink, 4096);
  cudaFuncSetAttribute(power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[] = {"N=128 TS, MMA only", "N=256 TS, MMA only", "N=128 + 16 warps of FFMA", "N=128 + 16 warps polling an mbarrier",
                         "N=128, MMA duty 50 %", "N=64 TS, MMA only", "N=128 + FFMA at ~50 % duty"};
  struct Cfg { int v, issuer, c0, c1; const char* what; };
  const Cfg cfgs[] = {
      {0, 0, 4, 20, "issuer warp 0"}, {1, 0, 4, 20, "issuer warp 0"}, {5, 0, 4, 20, "issuer warp 0"}, {4, 0, 4, 20, "issuer warp 0"},
      {3, 0, 4, 20, "issuer warp 0, side load warps 4-19"},
      {2, 0, 4, 20, "issuer warp 0, side load warps 4-19"}, {6, 0, 4, 20, "issuer warp 0, side load warps 4-19"},
      {2, 19, 0, 16, "issuer warp 19, side load warps 0-15"}, {6, 19, 0, 16, "issuer warp 19, side load warps 0-15"},
      {2, 0, 5, 20, "issuer warp 0, side load warps 5-19 (3 on its scheduler)"},
      {6, 0, 5, 20, "issuer warp 0, side load warps 5-19 (3 on its scheduler)"},
      {2, 0, 1, 4, "issuer warp 0, side load warps 1-3 only (none on its scheduler)"},
  };
  for (const Cfg& c : cfgs) {
    const int v = c.v;
    const int N = v == 1 ? 256 : (v == 5 ? 64 : 128);
    const double slow = (v == 2) ? 11.0 : (v == 6 ? 2.2 : (v == 4 ? 2.0 : 1.0));
    const double cyc_per_burst = 12.0 * (N == 256 ? 130.0 : (N == 64 ? 39.2 : 66.0)) * slow;
    const int bursts = (int)(seconds * 1.8e9 / cyc_per_burst);
    power_kernel<<<148, 640, 100 * 1024>>>(v, 2000, d_out, d_sink, c.issuer, c.c0, c.c1);      // warm-up
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    power_kernel<<<148, 640, 100 * 1024>>>(v, bursts, d_out, d_sink, c.issuer, c.c0, c.c1);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    unsigned long long cyc = 0;
    cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
    const double flops = 2.0 * 128.0 * N * 16.0 * 12.0 * (double)bursts * 148.0;
    printf("%-38s | %-62s %7.1f ms %7.1f TFLOP/s issued, SM clock %5.0f MHz, %7.1f cycles per burst\n", names[v], c.what, ms,
           flops / (ms * 1e-3) / 1e12, (double)cyc / (ms * 1e-3) / 1e6, (double)cyc / bursts);
  }
  return 0;
}
