// sm_100a building blocks of the tensor-core field engine: mbarrier, bulk TMA
// (cp.async.bulk), TMEM allocation / loads, tcgen05.mma with shared-memory
// descriptors.  All operands use the canonical K-major SWIZZLE_128B layout:
// a "K-block" is [rows][64 fp16] = rows x 128 bytes, 8-row groups 1024 bytes
// apart, 16-byte chunks XOR-swizzled with (row & 7).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nds {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// byte offset of element (row r, col c in [0,64)) inside one swizzled K-block
__device__ __host__ __forceinline__ uint32_t kblock_offset(uint32_t r, uint32_t c) {
  return (r >> 3) * 1024u + (r & 7u) * 128u + ((((c >> 3) ^ (r & 7u)) & 7u) << 4) + ((c & 7u) << 1);
}

// ---- mbarrier ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a while when the phase is still open)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

// ---- proxies / fences -------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05.mma)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ---- bulk TMA: contiguous global -> shared, completes on an mbarrier --------
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- TMEM -------------------------------------------------------------------
// whole warp; writes the allocated base address (lane 0, column base) to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (base_lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
      : "r"(taddr)
      : "memory");
}
template <int N> __device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[N]);
template <> __device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t (&v)[32]) { tmem_ld32(taddr, v); }
template <> __device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t (&v)[16]) { tmem_ld16(taddr, v); }
template <> __device__ __forceinline__ void tmem_ld<8>(uint32_t taddr, uint32_t (&v)[8]) { tmem_ld8(taddr, v); }

// registers -> TMEM: thread i of the warp writes lane (base_lane + i), N consecutive 32-bit columns
template <int N> __device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&v)[N]);
template <> __device__ __forceinline__ void tmem_st<16>(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
template <> __device__ __forceinline__ void tmem_st<8>(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
template <> __device__ __forceinline__ void tmem_st<4>(uint32_t taddr, const uint32_t (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- UMMA -------------------------------------------------------------------
// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): K-major,
// SWIZZLE_128B (layout_type 2 at bits [61,64)), SBO = 1024 B between 8-row
// groups, LBO unused for swizzled K-major (encoded 1), version 1 at bits [46,48).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                                // leading byte offset (unused)
  d |= (uint64_t)(1024u >> 4) << 32;                     // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                                // SWIZZLE_128B
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor), kind::f16: A, B fp16
// K-major, fp32 accumulate, M = 128, N = n.
__device__ __host__ __forceinline__ uint32_t make_idesc_f16(uint32_t n) {
  return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, one K = 16 step, issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// same with the A operand read from TMEM (lane = row, two fp16 K-elements per 32-bit column)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Predicated forms: executed by the whole (converged) warp with `issue` true in exactly one lane, so that the
// operand arithmetic stays warp-uniform (uniform datapath) and only the tcgen05 instruction is predicated.
__device__ __forceinline__ void umma_f16_p(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(issue));
}
__device__ __forceinline__ void umma_f16_ts_p(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(issue));
}
__device__ __forceinline__ void umma_commit_p(uint64_t* bar, uint32_t issue) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)),
      "r"(issue)
      : "memory");
}
// address forms (shared-window address of the barrier)
__device__ __forceinline__ void umma_commit_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "NDS_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra NDS_WAIT_%=;\n\t}" ::"r"(bar), "r"(parity)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- MMA groups ---------------------------------------------------------------
// Four tcgen05.mma (one K-chunk: K-steps 0..3, B descriptor low word + 2 per step) as ONE asm block, optionally
// preceded by a non-consuming probe of an mbarrier (mbarrier.try_wait whose predicate is read only AFTER the four
// MMAs): the issuing thread blocks on tcgen05.mma while the tensor core's short queue is full, so the barrier
// round trip (~100-150 cycles) hides behind the issue instead of sitting between two bursts.
// `nmma`: how many of the 4 K-steps exist (1..4).  Returns 1 when the probed phase had completed (or no probe).
#define NDS_DESC_HI 0x40004040u   /* SBO = 1024 B, version 1, SWIZZLE_128B */
#define NDS_MMA4_PROLOGUE \
  "{\n\t.reg .pred p, t, e1, e2, e3, pq, ok;\n\t.reg .b64 B0, B1, B2, B3;\n\t.reg .b32 w;\n\t" \
  "setp.ne.b32 p, %7, 0;\n\tsetp.gt.u32 e1, %8, 1;\n\tsetp.gt.u32 e2, %8, 2;\n\tsetp.gt.u32 e3, %8, 3;\n\t" \
  "setp.ne.b32 pq, %11, 0;\n\tsetp.eq.u32 ok, 0, 0;\n\tsetp.eq.u32 t, 0, 0;\n\t" \
  "mov.b64 B0, {%5, %12};\n\tadd.u32 w, %5, 2;\n\tmov.b64 B1, {w, %12};\n\tadd.u32 w, %5, 4;\n\tmov.b64 B2, {w, %12};\n\t" \
  "add.u32 w, %5, 6;\n\tmov.b64 B3, {w, %12};\n\t" \
  "@pq mbarrier.try_wait.parity.shared::cta.b64 ok, [%9], %10;\n\t"
// tensor-memory A operand: a0..a3 are the tensor-memory addresses of the four K-steps
__device__ __forceinline__ uint32_t mma4_ts(uint32_t d, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b_lo32,
                                            uint32_t idesc, uint32_t accumulate, uint32_t nmma, uint32_t probe_bar,
                                            uint32_t probe_parity, uint32_t do_probe) {
  uint32_t ok;
  asm volatile(NDS_MMA4_PROLOGUE
               "tcgen05.mma.cta_group::1.kind::f16 [%1], [%2], B0, %6, p;\n\t"
               "@e1 tcgen05.mma.cta_group::1.kind::f16 [%1], [%3], B1, %6, t;\n\t"
               "@e2 tcgen05.mma.cta_group::1.kind::f16 [%1], [%4], B2, %6, t;\n\t"
               "@e3 tcgen05.mma.cta_group::1.kind::f16 [%1], [%13], B3, %6, t;\n\t"
               "selp.u32 %0, 1, 0, ok;\n\t}"
               : "=r"(ok)
               : "r"(d), "r"(a0), "r"(a1), "r"(a2), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(nmma), "r"(probe_bar),
                 "r"(probe_parity), "r"(do_probe), "r"(NDS_DESC_HI), "r"(a3)
               : "memory");
  return ok;
}
// shared-memory A operand: a_lo32 is the descriptor low word of K-step 0 (+ 2 per step)
__device__ __forceinline__ uint32_t mma4_ss(uint32_t d, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc,
                                            uint32_t accumulate, uint32_t nmma, uint32_t probe_bar, uint32_t probe_parity,
                                            uint32_t do_probe) {
  uint32_t ok;
  asm volatile(NDS_MMA4_PROLOGUE
               ".reg .b64 A0, A1, A2, A3;\n\t"
               "mov.b64 A0, {%2, %12};\n\tadd.u32 w, %2, 2;\n\tmov.b64 A1, {w, %12};\n\tadd.u32 w, %2, 4;\n\tmov.b64 A2, {w, %12};\n\t"
               "add.u32 w, %2, 6;\n\tmov.b64 A3, {w, %12};\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%1], A0, B0, %6, p;\n\t"
               "@e1 tcgen05.mma.cta_group::1.kind::f16 [%1], A1, B1, %6, t;\n\t"
               "@e2 tcgen05.mma.cta_group::1.kind::f16 [%1], A2, B2, %6, t;\n\t"
               "@e3 tcgen05.mma.cta_group::1.kind::f16 [%1], A3, B3, %6, t;\n\t"
               "selp.u32 %0, 1, 0, ok;\n\t}"
               : "=r"(ok)
               : "r"(d), "r"(a_lo32), "r"(0u), "r"(0u), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(nmma), "r"(probe_bar),
                 "r"(probe_parity), "r"(do_probe), "r"(NDS_DESC_HI), "r"(0u)
               : "memory");
  return ok;
}
// A whole 3-term K-chunk (12 tcgen05.mma, tensor-memory A operand) as ONE asm block with the barrier probe of
// mma4_ts: the issuing thread shares its scheduler with four busy epilogue warps, so every instruction it does
// NOT execute between two bursts is worth several cycles of tensor-pipe time.  All operands are final
// (absolute) values; S1..S3 are the K-step column offsets of the operand pattern.
#define NDS_DEFINE_BURST12_TS(NAME, S1, S2, S3)                                                                          \
  __device__ __forceinline__ uint32_t NAME(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b0_lo32, uint32_t b1_lo32, \
                                           uint32_t idesc, uint32_t accumulate, uint32_t probe_bar, uint32_t probe_parity, \
                                           uint32_t do_probe) {                                                          \
    uint32_t ok;                                                                                                         \
    asm volatile(                                                                                                        \
        "{\n\t.reg .pred p, t, pq, ok;\n\t.reg .b64 H0, H1, H2, H3, G0, G1, G2, G3;\n\t"                                 \
        ".reg .b32 w, A1, A2, A3, L1, L2, L3;\n\t"                                                                       \
        "setp.ne.b32 p, %7, 0;\n\tsetp.eq.u32 t, 0, 0;\n\tsetp.eq.u32 ok, 0, 0;\n\tsetp.ne.b32 pq, %10, 0;\n\t"          \
        "mov.b64 H0, {%4, %11};\n\tadd.u32 w, %4, 2;\n\tmov.b64 H1, {w, %11};\n\tadd.u32 w, %4, 4;\n\tmov.b64 H2, {w, %11};\n\t" \
        "add.u32 w, %4, 6;\n\tmov.b64 H3, {w, %11};\n\t"                                                                 \
        "mov.b64 G0, {%5, %11};\n\tadd.u32 w, %5, 2;\n\tmov.b64 G1, {w, %11};\n\tadd.u32 w, %5, 4;\n\tmov.b64 G2, {w, %11};\n\t" \
        "add.u32 w, %5, 6;\n\tmov.b64 G3, {w, %11};\n\t"                                                                 \
        "add.u32 A1, %2, " S1 ";\n\tadd.u32 A2, %2, " S2 ";\n\tadd.u32 A3, %2, " S3 ";\n\t"                               \
        "add.u32 L1, %3, " S1 ";\n\tadd.u32 L2, %3, " S2 ";\n\tadd.u32 L3, %3, " S3 ";\n\t"                               \
        "@pq mbarrier.try_wait.parity.shared::cta.b64 ok, [%8], %9;\n\t"                                                 \
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [%2], H0, %6, p;\n\t"                                                  \
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [A1], H1, %6, t;\n\t"                                                  \
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [A2], H2, %6, t;\n\t"                                                  \
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [A3], H3, %6, t;\n\t"                                                  \
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [%3], H0, %6, t;\n\t"                                                  \
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [L1], H1, %6, t;\n\t"                                                  \
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [L2], H2, %6, t;\n\t"                                                  \
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [L3], H3, %6, t;\n\t"                                                  \
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [%2], G0, %6, t;\n\t"                                                  \
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [A1], G1, %6, t;\n\t"                                                  \
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [A2], G2, %6, t;\n\t"                                                  \
        "tcgen05.mma.cta_group::1.kind::f16 [%1], [A3], G3, %6, t;\n\t"                                                  \
        "selp.u32 %0, 1, 0, ok;\n\t}"                                                                                    \
        : "=r"(ok)                                                                                                       \
        : "r"(d), "r"(a_hi), "r"(a_lo), "r"(b0_lo32), "r"(b1_lo32), "r"(idesc), "r"(accumulate), "r"(probe_bar),         \
          "r"(probe_parity), "r"(do_probe), "r"(NDS_DESC_HI)                                                             \
        : "memory");                                                                                                     \
    return ok;                                                                                                           \
  }
NDS_DEFINE_BURST12_TS(burst12_ts32, "8", "32", "40")
NDS_DEFINE_BURST12_TS(burst12_ts16, "16", "32", "48")
__device__ __forceinline__ uint32_t smem_desc_lo32(uint32_t smem_addr) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
}

// ---- split fp16 -------------------------------------------------------------
__device__ __forceinline__ void split_h(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

}  // namespace tc
}  // namespace nds
