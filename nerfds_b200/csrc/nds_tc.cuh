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
// all previously issued MMAs of this thread arrive on `bar` when complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- MMA bursts -------------------------------------------------------------
// One K-chunk (4 K-steps of 16) of a split-fp16 Dense layer as ONE asm block:
//   D (+)= A_hi B_hi ; D += A_lo B_hi ; commit(bar_hi) ; D += A_hi B_lo ; commit(bar_lo) ; [commit(bar_d)]
// Descriptors differ only in their low word (start address >> 4, +2 per K-step), the high word is constant, so
// the issuing thread spends ~3 instructions per tcgen05.mma.  `a*` are descriptor low words (shared-memory
// operand) or tensor-memory addresses (TS form, +8 columns per K-step).
#define NDS_DESC_HI 0x40004040u   /* SBO = 1024 B, version 1, SWIZZLE_128B */
// operand setup first (all descriptors / addresses into their own registers), then the tcgen05 instructions
// back to back inside one branch taken by the elected lane only.
#define NDS_SETUP_SS \
  ".reg .b64 A0, A1, A2, A3, L0, L1, L2, L3, H0, H1, H2, H3, G0, G1, G2, G3;\n\t.reg .b32 w;\n\t" \
  "mov.b64 A0, {%1, %9};\n\tadd.u32 w, %1, 2;\n\tmov.b64 A1, {w, %9};\n\tadd.u32 w, %1, 4;\n\tmov.b64 A2, {w, %9};\n\tadd.u32 w, %1, 6;\n\tmov.b64 A3, {w, %9};\n\t" \
  "mov.b64 L0, {%2, %9};\n\tadd.u32 w, %2, 2;\n\tmov.b64 L1, {w, %9};\n\tadd.u32 w, %2, 4;\n\tmov.b64 L2, {w, %9};\n\tadd.u32 w, %2, 6;\n\tmov.b64 L3, {w, %9};\n\t"
// tensor-memory A operand: K-step t lives at column (base + S_t); the lo image at the same offsets from its base
#define NDS_SETUP_TS(S1, S2, S3) \
  ".reg .b32 A0, A1, A2, A3, L0, L1, L2, L3, w;\n\t.reg .b64 H0, H1, H2, H3, G0, G1, G2, G3;\n\t" \
  "mov.b32 A0, %1;\n\tadd.u32 A1, %1, " S1 ";\n\tadd.u32 A2, %1, " S2 ";\n\tadd.u32 A3, %1, " S3 ";\n\t" \
  "mov.b32 L0, %2;\n\tadd.u32 L1, %2, " S1 ";\n\tadd.u32 L2, %2, " S2 ";\n\tadd.u32 L3, %2, " S3 ";\n\t"
#define NDS_SETUP_B \
  "mov.b64 H0, {%3, %9};\n\tadd.u32 w, %3, 2;\n\tmov.b64 H1, {w, %9};\n\tadd.u32 w, %3, 4;\n\tmov.b64 H2, {w, %9};\n\tadd.u32 w, %3, 6;\n\tmov.b64 H3, {w, %9};\n\t" \
  "mov.b64 G0, {%4, %9};\n\tadd.u32 w, %4, 2;\n\tmov.b64 G1, {w, %9};\n\tadd.u32 w, %4, 4;\n\tmov.b64 G2, {w, %9};\n\tadd.u32 w, %4, 6;\n\tmov.b64 G3, {w, %9};\n\t"
#define NDS_MMA_SS(A, B, PRED) "tcgen05.mma.cta_group::1.kind::f16 [%0], " A ", " B ", %5, " PRED ";\n\t"
#define NDS_MMA_TS(A, B, PRED) "tcgen05.mma.cta_group::1.kind::f16 [%0], [" A "], " B ", %5, " PRED ";\n\t"
#define NDS_COMMIT(BAR) "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [" BAR "];\n\t"
#define NDS_BURST_BODY(SETUP_A, MMA) \
  "{\n\t.reg .pred p, t, l, q;\n\t" SETUP_A NDS_SETUP_B \
  "setp.ne.b32 p, %6, 0;\n\tsetp.eq.b32 t, 0, 0;\n\tsetp.ne.b32 q, %12, 0;\n\tsetp.ne.b32 l, %10, 0;\n\t" \
  "@!q bra NDS_DONE;\n\t" \
  MMA("A0", "H0", "p") MMA("A1", "H1", "t") MMA("A2", "H2", "t") MMA("A3", "H3", "t") \
  MMA("L0", "H0", "t") MMA("L1", "H1", "t") MMA("L2", "H2", "t") MMA("L3", "H3", "t") \
  NDS_COMMIT("%7") \
  MMA("A0", "G0", "t") MMA("A1", "G1", "t") MMA("A2", "G2", "t") MMA("A3", "G3", "t") \
  NDS_COMMIT("%8") \
  "@l " NDS_COMMIT("%11") \
  "NDS_DONE:\n\t}"
// 1-term K-chunk (4 K-steps): D (+)= A_hi B_hi ; commit(bar_slot) ; [commit(bar_d)]
#define NDS_BURST1_BODY(SETUP_A, MMA) \
  "{\n\t.reg .pred p, t, l, q;\n\t" SETUP_A NDS_SETUP_B \
  "setp.ne.b32 p, %6, 0;\n\tsetp.eq.b32 t, 0, 0;\n\tsetp.ne.b32 q, %12, 0;\n\tsetp.ne.b32 l, %10, 0;\n\t" \
  "@!q bra NDS_DONE;\n\t" \
  MMA("A0", "H0", "p") MMA("A1", "H1", "t") MMA("A2", "H2", "t") MMA("A3", "H3", "t") \
  NDS_COMMIT("%7") \
  "@l " NDS_COMMIT("%11") \
  "NDS_DONE:\n\t}"
__device__ __forceinline__ void umma_burst3_ss(uint32_t d, uint32_t a_hi_lo32, uint32_t a_lo_lo32, uint32_t b_hi_lo32,
                                               uint32_t b_lo_lo32, uint32_t idesc, uint32_t accumulate, uint32_t bar_hi,
                                               uint32_t bar_lo, uint32_t last, uint32_t bar_d, uint32_t issue) {
  asm volatile(NDS_BURST_BODY(NDS_SETUP_SS, NDS_MMA_SS)
               ::"r"(d), "r"(a_hi_lo32), "r"(a_lo_lo32), "r"(b_hi_lo32), "r"(b_lo_lo32), "r"(idesc), "r"(accumulate),
                 "r"(bar_hi), "r"(bar_lo), "r"(NDS_DESC_HI), "r"(last), "r"(bar_d), "r"(issue)
               : "memory");
}
// PAT: column offsets of K-steps 1..3 inside one 64-feature K-block of a tensor-memory operand
//   32: {8, 32, 40}  epilogue slices of 32 features (hi 16 columns | lo 16 columns)
//   16: {16, 32, 48} epilogue slices of 16 features (hi 8 | lo 8)
//    8: {8, 16, 24}  compacted hi-only activations
#define NDS_DEFINE_BURST3_TS(NAME, S1, S2, S3)                                                                        \
  __device__ __forceinline__ void NAME(uint32_t d, uint32_t a_hi_tmem, uint32_t a_lo_tmem, uint32_t b_hi_lo32,        \
                                       uint32_t b_lo_lo32, uint32_t idesc, uint32_t accumulate, uint32_t bar_hi,      \
                                       uint32_t bar_lo, uint32_t last, uint32_t bar_d, uint32_t issue) {              \
    asm volatile(NDS_BURST_BODY(NDS_SETUP_TS(S1, S2, S3), NDS_MMA_TS)                                                 \
                 ::"r"(d), "r"(a_hi_tmem), "r"(a_lo_tmem), "r"(b_hi_lo32), "r"(b_lo_lo32), "r"(idesc), "r"(accumulate), \
                   "r"(bar_hi), "r"(bar_lo), "r"(NDS_DESC_HI), "r"(last), "r"(bar_d), "r"(issue)                      \
                 : "memory");                                                                                         \
  }
NDS_DEFINE_BURST3_TS(umma_burst3_ts32, "8", "32", "40")
NDS_DEFINE_BURST3_TS(umma_burst3_ts16, "16", "32", "48")
NDS_DEFINE_BURST3_TS(umma_burst3_ts8, "8", "16", "24")
#define NDS_DEFINE_BURST1_TS(NAME, S1, S2, S3)                                                                        \
  __device__ __forceinline__ void NAME(uint32_t d, uint32_t a_tmem, uint32_t b_lo32, uint32_t idesc,                  \
                                       uint32_t accumulate, uint32_t bar_slot, uint32_t last, uint32_t bar_d,         \
                                       uint32_t issue) {                                                              \
    asm volatile(NDS_BURST1_BODY(NDS_SETUP_TS(S1, S2, S3), NDS_MMA_TS)                                                \
                 ::"r"(d), "r"(a_tmem), "r"(0u), "r"(b_lo32), "r"(0u), "r"(idesc), "r"(accumulate), "r"(bar_slot),    \
                   "r"(0u), "r"(NDS_DESC_HI), "r"(last), "r"(bar_d), "r"(issue)                                       \
                 : "memory");                                                                                         \
  }
NDS_DEFINE_BURST1_TS(umma_burst1_ts32, "8", "32", "40")
NDS_DEFINE_BURST1_TS(umma_burst1_ts16, "16", "32", "48")
NDS_DEFINE_BURST1_TS(umma_burst1_ts8, "8", "16", "24")
__device__ __forceinline__ void umma_burst1_ss(uint32_t d, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc,
                                               uint32_t accumulate, uint32_t bar_slot, uint32_t last, uint32_t bar_d,
                                               uint32_t issue) {
  asm volatile(NDS_BURST1_BODY(NDS_SETUP_SS, NDS_MMA_SS)
               ::"r"(d), "r"(a_lo32), "r"(0u), "r"(b_lo32), "r"(0u), "r"(idesc), "r"(accumulate), "r"(bar_slot), "r"(0u),
                 "r"(NDS_DESC_HI), "r"(last), "r"(bar_d), "r"(issue)
               : "memory");
}
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
