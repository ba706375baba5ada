// Micro-benchmark: tcgen05.mma issue/execution pacing on one SM (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I nerfds_b200/csrc -I include tools/mma_bench.cu -o gpurun_out/mma_bench
// Each variant issues `n` MMAs (M=128, K=16, fp16, fp32 accumulate) from one thread, commits to an mbarrier and
// reports cycles from first issue to completion.  Operand contents are irrelevant (zeros).
#include <cstdio>
#include <cstdlib>
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
struct Variant {
  int n_mma;      // MMAs issued
  int N;          // MMA N
  int a_tmem;     // 1: A from tensor memory (TS), 0: shared (SS)
  int n_acc;      // accumulators cycled through (1 = one dependent chain)
  int b_step;     // B descriptor advance per MMA in 16-byte units (2 = next K-step; 0 = same B tile)
  int n_cta_mma;  // unused
};

__global__ void __launch_bounds__(128, 1) mma_bench_kernel(Variant v, unsigned long long* out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, dummy[2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&dummy[0], 1); mbar_init(&dummy[1], 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tmem_base_s;
  if (warp == 0) {
    const uint32_t idesc = make_idesc_f16(v.N);
    const uint32_t a_lo32 = smem_desc_lo32(smem_u32(smem));
    const uint32_t b_lo32 = smem_desc_lo32(smem_u32(smem) + 16384);
    uint32_t parity = 0;
    const int nb = v.n_mma / (v.b_step ? 12 : 4);     // b_step != 0: 3-term bursts of 12; 0: 1-term bursts of 4
    for (int r = 0; r < reps; ++r) {
      const unsigned long long t0 = clock64();
      for (int i = 0; i < nb; ++i) {
        const uint32_t d = tb + (uint32_t)(i & (v.n_acc - 1)) * (uint32_t)v.N;
        const uint32_t b0 = b_lo32 + (uint32_t)(i & 1) * 2048u, b1 = b0 + 1024u;
        const uint32_t A0 = tb + 256 + (uint32_t)(i & 1) * 64, A1 = A0 + 16;
        const bool last = (i == nb - 1);
        if (elect_one_sync()) {
          const uint64_t bd0 = ((uint64_t)NDS_DESC_HI << 32) | b0, bd1 = ((uint64_t)NDS_DESC_HI << 32) | b1;
          const uint64_t ad0 = ((uint64_t)NDS_DESC_HI << 32) | a_lo32, ad1 = ad0 + 512u;
          if (v.a_tmem) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ts(d, A0 + 8 * k, bd0 + 2 * k, idesc, 1u);
            if (v.b_step) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16_ts(d, A1 + 8 * k, bd0 + 2 * k, idesc, 1u);
              umma_commit(&dummy[0]);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16_ts(d, A0 + 8 * k, bd1 + 2 * k, idesc, 1u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(d, ad0 + 2 * k, bd0 + 2 * k, idesc, 1u);
            if (v.b_step) {
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(d, ad1 + 2 * k, bd0 + 2 * k, idesc, 1u);
              umma_commit(&dummy[0]);
#pragma unroll
              for (int k = 0; k < 4; ++k) umma_f16(d, ad0 + 2 * k, bd1 + 2 * k, idesc, 1u);
            }
          }
          umma_commit(&dummy[1]);
          if (last) umma_commit(&bar);
        }
        __syncwarp();
      }
      const unsigned long long t1 = clock64();
      mbar_wait(&bar, parity);
      parity ^= 1u;
      const unsigned long long t2 = clock64();
      if (threadIdx.x == 0) { out[2 * r] = t1 - t0; out[2 * r + 1] = t2 - t0; }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// tcgen05.ld / st throughput: 16 warps (4 per lane quarter) each read + rewrite `cols` columns, `iters` times
__global__ void __launch_bounds__(512, 1) tmem_bench_kernel(int mode, int iters, unsigned long long* out) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tmem_base_s;
  const int q = warp & 3, sub = warp >> 2;
  const uint32_t addr = tb + (((uint32_t)q * 32u) << 16) + (uint32_t)sub * 32u;
  uint32_t v[32];
  for (int i = 0; i < 32; ++i) v[i] = i;
  uint32_t h[16];
  for (int i = 0; i < 16; ++i) h[i] = i;
  __syncthreads();
  const unsigned long long t0 = clock64();
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (mode == 0 || mode == 2) { tmem_ld32(addr + (uint32_t)(it & 3) * 128u, v); tmem_ld_wait(); acc += v[it & 31]; }
    if (mode == 1 || mode == 2) { h[0] = acc; tmem_st<16>(addr + (uint32_t)(it & 3) * 128u, h); tmem_st<16>(addr + 16 + (uint32_t)(it & 3) * 128u, h); tmem_st_wait(); }
  }
  __syncthreads();
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = acc; }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

// cost of tcgen05.commit / mbarrier waits for the issuing thread: bursts of 12 MMAs (N=128) followed by `ncommit`
// commits and `nwait` waits on an already-complete barrier
__global__ void __launch_bounds__(128, 1) commit_bench_kernel(int ncommit, int nwait, int test, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, dummy[4], done;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&dummy[i], 1); mbar_init(&done, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tb = tmem_base_s;
  if (warp == 0) {
    if (elect_one_sync()) {
      mbar_arrive(&done);            // phase 0 of `done` complete: waits with parity 0 succeed at once
      const uint32_t idesc = make_idesc_f16(128);
      const uint64_t bd = ((uint64_t)NDS_DESC_HI << 32) | smem_desc_lo32(smem_u32(smem) + 16384);
      const unsigned long long t0 = clock64();
      for (int i = 0; i < 16; ++i) {
#pragma unroll
        for (int k = 0; k < 12; ++k) umma_f16_ts(tb, tb + 256 + 8 * (k & 3), bd + 2 * (k & 3), idesc, 1u);
        for (int c = 0; c < ncommit; ++c) umma_commit(&dummy[c]);
        for (int w = 0; w < nwait; ++w) {
          if (test) { while (!mbar_test_wait(&done, 0)) {} } else mbar_wait(&done, 0);
        }
      }
      const unsigned long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      const unsigned long long t2 = clock64();
      out[0] = t1 - t0; out[1] = t2 - t0;
    }
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  unsigned long long* d_out;
  cudaMalloc(&d_out, 1024);
  cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const Variant vs[] = {
      {48, 128, 0, 1, 2}, {48, 128, 1, 1, 2}, {96, 128, 1, 1, 2}, {96, 128, 1, 2, 2}, {96, 256, 1, 1, 2}, {96, 256, 0, 1, 2},
      {96, 64, 1, 1, 2}, {96, 64, 1, 2, 2}, {96, 16, 1, 1, 2}, {96, 32, 1, 1, 2}, {96, 192, 1, 1, 2}, {12, 128, 1, 1, 2},
      {24, 128, 1, 1, 2}, {48, 128, 1, 1, 0}, {48, 256, 1, 1, 0}, {48, 64, 1, 1, 0}, {96, 96, 1, 1, 2}, {96, 112, 1, 1, 2},
  };
  for (const Variant& v : vs) {
    mma_bench_kernel<<<1, 128, 200 * 1024>>>(v, d_out, 4);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    unsigned long long h[8];
    cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
    printf("n_mma %3d N %3d %s acc %d bstep %d : issue %5llu total %5llu cycles -> %.1f cyc/MMA (rep1 %llu rep3 %llu)\n", v.n_mma, v.N,
           v.a_tmem ? "TS" : "SS", v.n_acc, v.b_step, h[4], h[5], (double)h[5] / v.n_mma, h[3], h[7]);
  }
  for (int mode = 0; mode < 3; ++mode) {
    tmem_bench_kernel<<<1, 512>>>(mode, 64, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    unsigned long long h[2];
    cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
    printf("tmem mode %d (0 ld32, 1 st 2x16, 2 both) 16 warps x 64 iters: %llu cycles -> %.1f cyc/iter (16 KB per iter per direction)\n",
           mode, h[0], (double)h[0] / 64);
  }
  cudaFuncSetAttribute(commit_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int cfg = 0; cfg < 8; ++cfg) {
    const int nc[] = {0, 1, 2, 0, 0, 0, 0, 1}, nw[] = {0, 0, 0, 1, 2, 1, 2, 1}, ts[] = {0, 0, 0, 0, 0, 1, 1, 0};
    commit_bench_kernel<<<1, 128, 100 * 1024>>>(nc[cfg], nw[cfg], ts[cfg], d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    unsigned long long h[2];
    cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
    printf("16 bursts of 12 MMAs N=128 + %d commits + %d %s waits per burst: issue %llu total %llu -> %.0f cycles per burst\n", nc[cfg], nw[cfg],
           ts[cfg] ? "test" : "try", h[0], h[1], (double)h[1] / 16);
  }
  // many CTAs at once (whole chip): does the pacing change under chip-wide load?
  {
    Variant v{96, 128, 1, 1, 2};
    mma_bench_kernel<<<148, 128, 200 * 1024>>>(v, d_out, 4);
    cudaDeviceSynchronize();
    unsigned long long h[8];
    cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost);
    printf("148 CTAs: n_mma 96 N 128 TS: total %llu -> %.1f cyc/MMA\n", h[5], (double)h[5] / 96);
  }
  return 0;
}
