// Micro-benchmark of the field engine's MMA issuer loop (issue_program + produce_program of nds_field_tc.cu) on a
// synthetic program of plain bursts: cycles per burst with the tensor core as the only consumer.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I nerfds_b200/csrc -I include tools/issue_bench.cu -o tools/bin/issue_bench
#include "../nerfds_b200/csrc/nds_field_tc.cu"

using namespace nds;

__global__ void __launch_bounds__(576, 1)
issue_bench_kernel(const __grid_constant__ TcProgram P, const uint8_t* weights, int iters, unsigned long long* out, int hammer) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Ctrl* ctl = reinterpret_cast<Ctrl*>(smem + OFF_CTRL);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < OFF_RING / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < P.n_burst * 8; i += blockDim.x)
    reinterpret_cast<uint32_t*>(ctl->burst)[i] = reinterpret_cast<const uint32_t*>(P.burst)[i];
  if (threadIdx.x == 0) { ctrl_init(ctl); *reinterpret_cast<volatile int*>(smem + OFF_CTRL + 11200) = 0; }
  if (warp == 0) tmem_alloc(&ctl->tmem_base, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;
  if (warp >= 2) {
    // optional TMEM traffic like the epilogues': every compute warp reads 32 columns and writes them back
    if (hammer) {
      const int cw = warp - 2, q = cw & 3, sub = cw >> 2;
      const uint32_t addr = tmem_base + (((uint32_t)q * 32u) << 16) + 384u + (uint32_t)sub * 32u;
      uint32_t v[32], h[16];
      volatile int* stop = reinterpret_cast<volatile int*>(smem + OFF_CTRL + 11200);
      int n = 0;
      while (!*stop) {
        tmem_ld32(addr, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) h[i] = v[2 * i] + v[2 * i + 1];
        if (hammer > 1) { tmem_st<16>(addr, h); tmem_st<16>(addr + 16, h); tmem_st_wait(); }
        if (hammer > 2) __nanosleep(hammer);
        ++n;
      }
      if (lane == 0 && cw == 0) out[1] = n;
    }
  } else if (warp == 1) {
    if (lane == 0) {
      ProducerState ps{0u, 0u};
      for (int it = 0; it < iters; ++it) produce_program(P, weights, smem, ctl, ps);
    }
    __syncwarp();
  } else {
    if (elect_one_sync()) {
      uint32_t bits = 0;
      const uint32_t s0 = P.src[0], u0 = (s0 >> 29) & 3u;
      if (s0 >> 31) { mbar_wait(&ctl->full[u0], 0u); bits ^= 1u << (8 + u0); }
      const unsigned long long t0 = clock64();
      for (int it = 0; it < iters; ++it) issue_program(P, ctl, bits, it + 1 < iters, nullptr);
      const unsigned long long t1 = clock64();
      umma_commit(&ctl->d_full[1][1]);      // drain: everything issued has completed
      mbar_wait(&ctl->d_full[1][1], 0);
      out[0] = t1 - t0;
      *reinterpret_cast<volatile int*>(smem + OFF_CTRL + 11200) = 1;
    }
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
  const int n = 64;
  for (int hammer = 0; hammer < 3; ++hammer)
  for (int variant = 0; variant < 2; ++variant) {
    std::vector<BurstH> hb(n);
    for (int i = 0; i < n; ++i) {
      BurstH& e = hb[i];
      e.pat = variant == 2 ? PAT_SS : PAT_32;
      e.a_hi = variant == 2 ? OFF_IN : 256 + 32 * (i & 1);
      e.a_lo = variant == 2 ? OFF_IN + KBLK : e.a_hi + 16;
      e.d_col = (i & 1) * 128;
      e.rows = variant == 1 ? 64 : 128;
      e.steps = 4;
      e.unit = i % NUNIT;
      e.tslot = 0;
      e.src = (uint32_t)((i % 8) * 256);
      e.rows128 = (uint16_t)(2 * e.rows);
      e.flags = B_TWO | B_ACQUIRE | B_RELEASE;
    }
    static TcProgram prog;
    memset(&prog, 0, sizeof prog);
    prog.n_burst = n;
    std::string perr;
    if (!query_smem_base(prog.smem_base, perr)) { printf("%s\n", perr.c_str()); return 1; }
    encode_bursts(hb, true, prog.smem_base, prog.burst, prog.src);
    uint8_t* d_w;
    cudaMalloc(&d_w, 8 * 32768);
    cudaMemset(d_w, 0, 8 * 32768);
    unsigned long long* d_out;
    cudaMalloc(&d_out, 64);
    cudaFuncSetAttribute(issue_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
    const int iters = 8;
    issue_bench_kernel<<<1, 576, TC_SMEM_BYTES>>>(prog, d_w, iters, d_out, hammer);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    unsigned long long h = 0, hn[2] = {0, 0};
    cudaMemcpy(hn, d_out, 16, cudaMemcpyDeviceToHost);
    h = hn[0];
    printf("[tmem traffic %d: %llu ld(+st) rounds of 64 KB] ", hammer, hammer ? hn[1] : 0ull);
    const char* names[] = {"TS PAT_32 N=128 3-term", "TS PAT_32 N=64 3-term", "SS N=128 3-term"};
    printf("%s: %d bursts x %d iterations: %llu cycles -> %.0f cycles per burst (ideal %d)\n", names[variant], n, iters, h,
           (double)h / (n * iters), variant == 1 ? 12 * 32 : 12 * 64);
    cudaFree(d_w); cudaFree(d_out);
  }
  return 0;
}
