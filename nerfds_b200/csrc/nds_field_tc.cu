// tcgen05 tensor-core field engine (sm_100a), "pair" schedule.
//
// One persistent CTA per SM processes PAIRS of 128-sample tiles (tile slots A and B).  Every Dense layer of the
// path (hypernerf/modules.py:57-83) is a [128 x K] x [K x N] GEMM issued as tcgen05.mma (kind::f16, fp32
// accumulators in TMEM):
//   * A = the tile's activations as split fp16 (hi + lo), kept IN TENSOR MEMORY for the whole chain (TS-mode MMA):
//     layer l reads its operand from one TMEM region and accumulates into the other; the epilogue converts the
//     accumulators IN PLACE (tcgen05.ld -> bias/ReLU -> split -> tcgen05.st over the columns it just read) into
//     the operand of layer l+1.  Network inputs (posenc features, embeddings, mask) are the only shared-memory
//     operands (canonical K-major SWIZZLE_128B block, SS-mode MMA);
//   * B = the layer's weights, pre-packed on the host into the exact shared-memory image (scaled by a power of
//     two, split hi + lo, swizzled) and streamed through a ring of 4 x 32 KB units by bulk TMA (cp.async.bulk);
//   * "3-term" layers issue A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (~fp32 accuracy, needed on the sigma path for the
//     1e-3 RGB bound -- tools/precision_study.py), "1-term" layers A_hi*B_hi only (rgb branch in `mixed` precision).
//
// Schedule of one pair (static, built on the host: TcProgram):
//   N phase  the narrow networks (mask MLP -> hyper sheet -> SE(3) warp field; width <= 128) of BOTH tiles,
//            interleaved op by op: each tile owns 256 TMEM columns (two regions of 128), the tensor core works on
//            one tile while the compute warps run the other tile's epilogue / per-sample stage, and every weight
//            image loaded into the ring is used by both tiles before it is released;
//   T phase  the template NeRF (trunk 8 x 256, sigma/normal head, rgb branch with the activation-free bottleneck
//            folded into its first layer) needs all 512 columns (two regions of 256), so tile A then tile B run it
//            alone, each layer split into two N-chunks whose K-ranges are ordered so that the tensor core never
//            waits for the second chunk's epilogue.
//   The three narrow networks read ONE shared feature block per tile ([sin/cos bands of x | embeddings | mask],
//   posenc windows folded into the first-layer weights on the host), written one pair ahead during the T phase.
// Warp roles: warp 0 = MMA issuer, warp 1 = TMA producer, warps 4-19 = compute (TMEM lane quarter = warp % 4 ->
// 32 samples, column slice = (warp - 4) / 4): positional encodings, SE(3) exponential, epilogues.
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <type_traits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "nds_dual.cuh"
#include "nds_host.h"
#include "nds_tc.cuh"

namespace nds {

using namespace tc;

constexpr int TM = 128;                 // samples per tile (UMMA M)
constexpr int NSUB = 4;                 // compute warps per TMEM lane quarter
constexpr int N_CWARPS = 4 * NSUB;
// The MMA issuer is warp 0: it shares its scheduler with four compute warps, and the oldest warp of a scheduler
// wins the issue slot.  Warps 2 and 3 only keep the compute warps' TMEM lane quarter = warp % 4.
#ifndef NDS_ISSUER_LAST
#define NDS_ISSUER_LAST 0   // experiment: compute warps 0-15, TMA producer 16, MMA issuer = the LAST warp (19)
#endif
#if NDS_ISSUER_LAST
constexpr int WARP_MMA = 19, WARP_TMA = 16, CWARP0 = 0;
#else
constexpr int WARP_MMA = 0, WARP_TMA = 1, CWARP0 = 4;
#endif
constexpr int TC_THREADS = 20 * 32;
__device__ __forceinline__ bool is_compute_warp(int warp) { return warp >= CWARP0 && warp < CWARP0 + N_CWARPS; }
constexpr uint32_t KBLK = 16384;        // one 128-row K-block (64 fp16 columns)
// shared memory map
constexpr uint32_t OFF_IN = 0;          // IN[tile slot]: hi block at slot * 2 KBLK, lo block right after
constexpr uint32_t OFF_IN2 = 4 * KBLK;  // rgb-branch side inputs [viewdir feats | normal feats], hi only
constexpr uint32_t OFF_RING = 5 * KBLK;
constexpr uint32_t UNIT_BYTES = 32768;  // one ring unit: a B_hi image followed by its B_lo image (<= 2 x 16 KB)
constexpr int NUNIT = 4;
constexpr uint32_t OFF_CTRL = OFF_RING + NUNIT * UNIT_BYTES;
constexpr uint32_t CTRL_HDR = 256;      // barriers + TMEM base; the program copies (ops | steps | bursts) follow
constexpr uint32_t TC_SMEM_MAX = 232448;                     // 227 KB opt-in limit of sm_100
constexpr uint32_t TC_SMEM_PROBE = OFF_CTRL + 1024;          // any size: only the base address matters
constexpr int NPREP = 3;                // the shared feature block of the next pair is written in 3 parts
#ifndef NDS_EPI_RZ
#define NDS_EPI_RZ 0    // 1: ReLU folded into cvt.rz.relu (2 instructions fewer per pair, hi truncated -> lo twice as large)
#endif
#ifndef NDS_EPI_TRUNC
#define NDS_EPI_TRUNC 0 // 1: hi = x with the low 13 mantissa bits cleared (LOP3 on the alu pipe instead of HADD2.F32 x2 on
#endif                  //    the fma pipe), ReLU folded into the two conversions: 4 fma-pipe instructions per pair instead of 6
#ifndef NDS_SUBTRACE
#define NDS_SUBTRACE 0  // diagnostics build: clock64 stamps inside the epilogues of the traced warp (NDS_TC_TRACE)
#endif

enum EpiKind : uint8_t {
  EPI_INPLACE = 0,      // slice of CW accumulator columns -> hi (CW/2 columns) | lo (CW/2 columns) over the same slice
  EPI_INPLACE_HI = 1,   // hi only (consumer is a 1-term layer)
  EPI_COMPACT_HI = 2,   // hi only, compacted to the first half of the chunk (frees the second half for accumulators)
  EPI_HEAD = 4,         // <= 16 outputs consumed by the per-sample stage
  EPI_INGRAD = 5        // reverse sweep: 64 input-gradient columns, accumulated per thread (16 columns each)
};
// ReLU masks of the reverse sweep (d sigma / dx, models.py:1035-1077): one 32-bit word per (layer, N-chunk) and
// thread, bit i = pre-activation of the thread's i-th column > 0 (jax.nn.relu's derivative is x > 0)
enum MaskMode : uint8_t { MASK_NONE = 0, MASK_RECORD = 1, MASK_APPLY = 2 };
constexpr int MASK_TRUNK = 0, MASK_WARP = 16, MASK_HYPER = 24, MASK_WORDS = 32;
enum Glue : uint8_t { GLUE_NONE = 0, GLUE_MASK = 1, GLUE_WARP = 2, GLUE_HYPER = 3, GLUE_ALPHA = 4, GLUE_BOTTLENECK = 5,
                      GLUE_RGB = 6, GLUE_SELFTEST = 7 };
// A-operand addressing pattern of a burst: 0 = shared memory (IN block), else tensor memory with the K-step
// column offsets {0,8,32,40} / {0,16,32,48} / {0,8,16,24}
enum APattern : uint8_t { PAT_SS = 0, PAT_32 = 1, PAT_16 = 2, PAT_8 = 3 };

// What the compute warps need to know about one Dense layer of one tile slot.
struct TcOp {
  uint32_t bias_off;     // float offset into the bias array
  float inv_scale;       // accumulators hold (scale * W) x; multiply back
  uint16_t N;            // output columns (padded)
  uint16_t nc_rows;      // output columns per N-chunk
  uint8_t n_nc, relu, epi_kind, glue;
  uint16_t d_col[2];     // tensor-memory column of accumulator chunk c
  uint8_t signal_glue;   // the per-sample stage after this head arrives on glue[tile slot]
  uint8_t signal_done;   // last tensor-memory read of the tile slot: arrive on done[tile slot] right after it
  uint8_t mask_idx;      // first mask word of the op (+ N-chunk)
  uint8_t mask_mode;     // MaskMode
};
static_assert(sizeof(TcOp) == 24, "TcOp layout");

// One burst of the MMA issuer = one K-chunk (<= 4 K-steps) of one N-chunk of one op of one tile slot:
// 12 tcgen05.mma for a 3-term layer (B_hi and B_lo images), 4 for a 1-term layer.
enum BurstFlags : uint16_t {
  B_TWO = 1, B_FIRST = 2, B_LAST = 4, B_NC1 = 8, B_WAIT_P0 = 16, B_WAIT_P1 = 32,
  B_WAIT_GLUE = 64,          // the tile slot's previous per-sample stage is done (head consumed, inputs written)
  B_PEEK_GLUE_OTHER = 128,   // ... of the OTHER tile slot, without consuming the phase
  B_WAIT_PREP = 256,         // the shared feature block of this tile slot is written
  B_WAIT_DONE_OTHER = 512,   // the other tile slot has left the T phase (its TMEM columns are free)
  B_PART_NEXT = 1024,        // last burst of an op that consumed one output phase of the previous op
  B_ACQUIRE = 2048,          // first use of a ring unit: wait for the TMA
  B_RELEASE = 4096           // last use: commit to the unit's empty barrier
};
constexpr uint32_t CTL_PAIR = 1u << 25;             // Burst::ctl: PAIR burst
constexpr uint32_t CTL_WAIT_DONE_SELF = 1u << 26;   // Burst::ctl: this tile slot's previous done phase (reverse sweep)
constexpr uint16_t B_WAIT_ANY = B_WAIT_P0 | B_WAIT_P1 | B_WAIT_GLUE | B_PEEK_GLUE_OTHER | B_WAIT_PREP | B_WAIT_DONE_OTHER;
// Device encoding, everything the issuer needs as FINAL values (it is read as two uint4 from the constant bank):
// the issuing thread shares its warp scheduler with four epilogue warps, so its instruction count per burst is
// what separates two bursts on the tensor pipe.  Shared-memory addresses are absolute (TcProgram::smem_base is
// checked by the kernel), tensor memory is allocated whole (base 0, checked).
struct alignas(16) Burst {
  uint32_t a_hi, a_lo;   // PAT_SS: shared-memory descriptor low word; else tensor-memory address
  uint32_t d;            // accumulator tensor-memory address
  uint32_t idesc;        // tcgen05 instruction descriptor (M = 128, N = rows)
  uint32_t b0, b1;       // B_hi / B_lo image: descriptor low words
  uint32_t ctl;          // [0,13) BurstFlags | [13,15) pat | [15,18) steps | [18] tile slot | [19,21) ring unit |
                         // [21,24) 1 + ring unit the NEXT burst acquires (0: none) | [24] that next burst is burst 0 |
                         // [25] PAIR burst: two groups of four MMAs, (a_hi, b0) then (a_lo, b1) -- see make_bursts
  uint32_t bars;         // Ctrl barrier indices (0xff: none): [0,8) commit on release | [8,16) commit when the
                         // accumulators are complete | [16,24) weight barrier of the next burst (probe) |
                         // [24,32) commit on release of a second ring unit
};
static_assert(sizeof(Burst) == 32, "Burst is read as two uint4");
// host-side description of a burst (encode_burst() makes the device form)
struct BurstH {
  uint32_t a_hi = 0, a_lo = 0, src = 0;
  uint16_t d_col = 0, rows = 0, flags = 0, rows128 = 0;
  uint8_t steps = 4, pat = 0, unit = 0, tslot = 0;
  // PAIR bursts (cross-first ops): group 1 = (a_hi, image b_sel[0] of K-chunk kc), group 2 = (a_lo, image b_sel[1]
  // of K-chunk kc2 or kc); b_sel: 0 = B_hi image, 1 = B_lo image
  uint8_t pair = 0, b_sel[2] = {0, 1}, unit2 = 0xff, release2 = 0;
  int16_t kc = 0, kc2 = -1;           // K-chunks of the op whose weight images the burst reads
  uint32_t ctl_extra = 0;             // CTL_WAIT_DONE_SELF
};

enum StepKind : uint8_t { STEP_EPI = 0, STEP_HEAD = 1, STEP_VIEW = 2, STEP_PREP = 3, STEP_OUT = 4,
                          STEP_INGRAD = 5,   // reverse sweep: accumulate an input-gradient op (arg 1: first of its sum)
                          STEP_SEED = 6,     // reverse sweep: write the gradient seed of a network (arg: SeedKind)
                          STEP_PEB = 7 };    // reverse sweep: positional-encoding backward (arg 0: trunk input, 1: feature block)
enum SeedKind : uint8_t { SEED_TRUNK = 0, SEED_HYPER = 1, SEED_WARP = 2 };
struct Step { uint8_t kind, tslot, op, arg; };

constexpr int MAX_OPS = 120;
constexpr int MAX_STEPS = 200;
constexpr int MAX_BURST = 448;
struct TcProgram {       // passed by value as a __grid_constant__ kernel parameter (constant bank)
  int n_ops, n_burst, n_steps, full;     // full: rgb branch present (else sigma-only)
  int carried;           // the program has no N phase: the narrow networks' results are read from the carry planes
  int grad;              // the program ends every tile with the reverse sweep for -d(sigma_raw)/dx
  uint32_t smem_bytes;   // dynamic shared memory of the launch (the control block is sized for this program)
  // shared feature block: [identity x (3)] [sin/cos of bands f_kmin .. f_kmin + f_nb) (6 each)] [warp embed]
  // [mask embed] [mask]; t_cols = width of the trunk input that later overwrites it
  int f_col_ident, f_col_bands, f_kmin, f_nb, f_col_wembed, f_col_membed, f_col_mask, f_cols, t_cols;
  uint32_t smem_base;    // shared-window address of the (1024-aligned) dynamic shared memory the bursts were encoded for
  TcOp ops[MAX_OPS];
  Step steps[MAX_STEPS];
  Burst burst[MAX_BURST];
  uint32_t src[MAX_BURST];   // producer: [0,20) weight stream offset in 128-byte rows | [20,29) rows to load |
                             // [29,31) ring unit | [31] the burst acquires (loads) its unit
};
static_assert(sizeof(TcProgram) < 24576, "kernel parameter budget (32 764 bytes with the other arguments)");

constexpr int TRACE_SUB = 3 * MAX_BURST + 2 * MAX_STEPS + 8;   // sub-step stamps: 8 per step (N-chunk c: 4 c + {waited, loaded, stored, arrived})
constexpr int TRACE_WORDS = TRACE_SUB + 8 * MAX_STEPS;

struct TcLevel {
  const uint8_t* weights;
  const float* bias;
};

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
struct Ctrl {     // the barriers come first, in this order: Burst::bars holds indices into this array
  uint64_t full[NUNIT];
  uint64_t empty[NUNIT];
  uint64_t prep[2];         // compute -> MMA: shared feature block of tile slot s written
  uint64_t glue[2];         // compute -> MMA: per-sample stage after a head done (head consumed, inputs written)
  uint64_t done[2];         // compute -> MMA: tile slot s finished its T phase (every TMEM read done)
  uint64_t part[2][2];      // compute -> MMA: N-chunk c of the current op drained and re-written as operand
  uint64_t d_full[2][2];    // MMA -> compute: accumulators of N-chunk c complete
  uint32_t tmem_base;
  uint32_t pad[3];
  // The program copies follow at CTRL_HDR (ops | steps | bursts, each 16-byte aligned, sized for THIS program): a
  // dynamically indexed constant-bank read costs a few hundred cycles, a shared-memory read ~30.
};
static_assert(sizeof(Ctrl) <= CTRL_HDR, "control header");
__host__ __device__ __forceinline__ uint32_t ctrl_off_steps(int n_ops) { return CTRL_HDR + (((uint32_t)n_ops * (uint32_t)sizeof(TcOp) + 15u) & ~15u); }
__host__ __device__ __forceinline__ uint32_t ctrl_off_burst(int n_ops, int n_steps) { return ctrl_off_steps(n_ops) + (((uint32_t)n_steps * 4u + 15u) & ~15u); }
__host__ __device__ __forceinline__ uint32_t ctrl_bytes(int n_ops, int n_steps, int n_burst) { return ctrl_off_burst(n_ops, n_steps) + (uint32_t)n_burst * 32u; }

__device__ __forceinline__ void ctrl_init(Ctrl* ctl) {
  for (int i = 0; i < NUNIT; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); }
  for (int s = 0; s < 2; ++s) {
    mbar_init(&ctl->prep[s], N_CWARPS);
    mbar_init(&ctl->glue[s], N_CWARPS);
    mbar_init(&ctl->done[s], N_CWARPS);
    for (int c = 0; c < 2; ++c) { mbar_init(&ctl->part[s][c], N_CWARPS); mbar_init(&ctl->d_full[s][c], 1); }
  }
  mbar_fence_init();
}

// one arrival per compute warp, after every lane made its writes visible
template <bool SMEM_WRITES = true>
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  if (SMEM_WRITES) fence_proxy_async_smem();     // st.shared operand images -> async proxy (tcgen05.mma)
  tmem_st_wait();               // tcgen05.st operand images complete
  tc_fence_before_sync();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// elect.sync in the form ptxas recognises as "exactly one thread runs the guarded region": the tcgen05
// instructions inside `if (elect_one_sync())` are then emitted back to back on the uniform datapath.  (Guarding
// them with any other predicate -- threadIdx.x == 0, a selp'd flag -- makes ptxas wrap EVERY UTCHMMA in an
// ELECT / BRA.U.ANY serialisation loop: tools/mma_bench.cu.)
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\t"
      "elect.sync %%rx|%%px, %2;\n\t"
      "@%%px mov.s32 %1, 1;\n\t"
      "mov.s32 %0, %%rx;\n\t}"
      : "+r"(laneid), "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred;
}

// K-step column offsets of a tensor-memory operand K-block, per pattern
template <int PAT> struct KSteps;
template <> struct KSteps<PAT_32> { static constexpr uint32_t s1 = 8, s2 = 32, s3 = 40; };
template <> struct KSteps<PAT_16> { static constexpr uint32_t s1 = 16, s2 = 32, s3 = 48; };
template <> struct KSteps<PAT_8> { static constexpr uint32_t s1 = 8, s2 = 16, s3 = 24; };

// One burst: the first group of four MMAs carries a probe of the NEXT burst's weight barrier (see mma4_ts), read
// back only after the whole burst has been issued.
//   D (+)= A_hi B_hi [; D += A_lo B_hi ; D += A_hi B_lo]
template <int PAT>
__device__ __forceinline__ uint32_t issue_burst_ts(bool two, uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b0,
                                                   uint32_t b1, uint32_t idesc, uint32_t acc, uint32_t pbar, uint32_t ppar,
                                                   uint32_t do_probe) {
  using S = KSteps<PAT>;
  const uint32_t ok = mma4_ts(d, a_hi, a_hi + S::s1, a_hi + S::s2, a_hi + S::s3, b0, idesc, acc, 4u, pbar, ppar, do_probe);
  if (two) {
    mma4_ts(d, a_lo, a_lo + S::s1, a_lo + S::s2, a_lo + S::s3, b0, idesc, 1u, 4u, 0u, 0u, 0u);
    mma4_ts(d, a_hi, a_hi + S::s1, a_hi + S::s2, a_hi + S::s3, b1, idesc, 1u, 4u, 0u, 0u, 0u);
  }
  return ok;
}
__device__ __forceinline__ uint32_t issue_burst(uint32_t pat, bool two, uint32_t steps, uint32_t d, uint32_t a_hi,
                                                uint32_t a_lo, uint32_t b0, uint32_t b1, uint32_t idesc, uint32_t acc,
                                                uint32_t pbar, uint32_t ppar, uint32_t do_probe) {
  if (pat == PAT_32) return issue_burst_ts<PAT_32>(two, d, a_hi, a_lo, b0, b1, idesc, acc, pbar, ppar, do_probe);
  if (pat == PAT_16) return issue_burst_ts<PAT_16>(two, d, a_hi, a_lo, b0, b1, idesc, acc, pbar, ppar, do_probe);
  if (pat == PAT_8) return issue_burst_ts<PAT_8>(two, d, a_hi, a_lo, b0, b1, idesc, acc, pbar, ppar, do_probe);
  const uint32_t ok = mma4_ss(d, a_hi, b0, idesc, acc, steps, pbar, ppar, do_probe);
  if (two) {
    mma4_ss(d, a_lo, b0, idesc, 1u, steps, 0u, 0u, 0u);
    mma4_ss(d, a_hi, b1, idesc, 1u, steps, 0u, 0u, 0u);
  }
  return ok;
}

// PAIR burst: D (+)= A1 B1 ; D += A2 B2 (four K-steps each).  Cross-first ops issue their small terms
// (A_lo B_hi, A_hi B_lo) of every K-chunk before any main term (A_hi B_hi): the tensor core's fp32 accumulator
// TRUNCATES at every accumulating MMA, one ulp of the running sum each, so the small terms are summed while the
// accumulator is still small and the large running sum takes 4 truncations per K-chunk instead of 12.
__device__ __forceinline__ uint32_t issue_pair(uint32_t pat, uint32_t steps, uint32_t d, uint32_t a1, uint32_t a2, uint32_t b1,
                                               uint32_t b2, uint32_t idesc, uint32_t acc, uint32_t pbar, uint32_t ppar,
                                               uint32_t do_probe) {
  uint32_t ok;
  if (pat == PAT_SS) {
    ok = mma4_ss(d, a1, b1, idesc, acc, steps, pbar, ppar, do_probe);
    mma4_ss(d, a2, b2, idesc, 1u, steps, 0u, 0u, 0u);
    return ok;
  }
  const uint32_t s1 = pat == PAT_16 ? 16u : 8u, s2 = pat == PAT_8 ? 16u : 32u, s3 = pat == PAT_32 ? 40u : (pat == PAT_16 ? 48u : 24u);
  ok = mma4_ts(d, a1, a1 + s1, a1 + s2, a1 + s3, b1, idesc, acc, 4u, pbar, ppar, do_probe);
  mma4_ts(d, a2, a2 + s1, a2 + s2, a2 + s3, b2, idesc, 1u, 4u, 0u, 0u, 0u);
  return ok;
}

// Ctrl barrier index -> shared-window address (all barriers are the leading uint64 array of Ctrl)
__device__ __forceinline__ uint32_t bar_addr(uint32_t ctl_addr, uint32_t idx) { return ctl_addr + 8u * idx; }

// MMA issuer, run by ONE elected thread (the caller guards it with elect_one_sync(), which is what lets ptxas keep
// the whole loop -- program decode, barrier waits, tcgen05 instructions -- on the uniform datapath).
// `bits`: one parity bit per barrier the issuer waits on.  bit s: part[s][*]; 2+s: glue[s]; 4+s: prep[s];
// 6+s: done[s]; 8+u: full[u].
__device__ __forceinline__ void issue_program(const TcProgram& P, Ctrl* ctl, uint32_t& bits_io, bool more,
                                              unsigned long long* trace) {
  const int n = P.n_burst;
  const uint32_t ctl_addr = smem_u32(ctl);
  uint32_t bits = bits_io;
  // program entries come from the shared-memory copy (a dynamically indexed constant-bank read misses the small
  // immediate-constant cache and costs a few hundred cycles), fetched one burst ahead
  const uint4* bp = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(ctl) + ctrl_off_burst(P.n_ops, P.n_steps));
  uint4 q0 = bp[0];
  uint4 q1 = bp[1];
  for (int i = 0; i < n; ++i) {
    const int ni = (i + 1 < n) ? i + 1 : 0;
    const uint4 n0 = bp[2 * ni];
    const uint4 n1 = bp[2 * ni + 1];
    const uint32_t c = q1.z, s = (c >> 18) & 1u;
    if (trace) trace[i] = clock64();
    if (c & (B_WAIT_ANY | CTL_WAIT_DONE_SELF)) {
      const uint32_t o = s ^ 1u;
      if (c & B_WAIT_PREP) { mbar_wait(&ctl->prep[s], (bits >> (4 + s)) & 1u); bits ^= 1u << (4 + s); }
      if (c & B_WAIT_DONE_OTHER) { mbar_wait(&ctl->done[o], (bits >> (6 + o)) & 1u); bits ^= 1u << (6 + o); }
      if (c & CTL_WAIT_DONE_SELF) { mbar_wait(&ctl->done[s], (bits >> (6 + s)) & 1u); bits ^= 1u << (6 + s); }
      if (c & B_WAIT_GLUE) { mbar_wait(&ctl->glue[s], (bits >> (2 + s)) & 1u); bits ^= 1u << (2 + s); }
      if (c & B_PEEK_GLUE_OTHER) mbar_wait(&ctl->glue[o], (bits >> (2 + o)) & 1u);
      if (c & B_WAIT_P0) mbar_wait(&ctl->part[s][0], (bits >> s) & 1u);
      if (c & B_WAIT_P1) mbar_wait(&ctl->part[s][1], (bits >> s) & 1u);
      tc_fence_after_sync();
    }
    if (trace) trace[MAX_BURST + i] = clock64();
    // the next burst's weights (in the ring long ago in steady state) are probed while this burst is issued.
    // Only ring waits may be hoisted: the producer never depends on anything this thread still has to issue.
    const uint32_t nu = (c >> 21) & 7u;
    const uint32_t do_probe = (nu != 0u && (((c >> 24) & 1u) == 0u || more)) ? 1u : 0u;
    const uint32_t pbar = bar_addr(ctl_addr, (q1.w >> 16) & 0xffu), ppar = (bits >> (7 + nu)) & 1u;
    const uint32_t acc = (c & B_FIRST) ? 0u : 1u, pat = (c >> 13) & 3u;
    const bool two = (c & B_TWO) != 0;
    uint32_t ok;
    if (c & CTL_PAIR) ok = issue_pair(pat, (c >> 15) & 7u, q0.z, q0.x, q0.y, q1.x, q1.y, q0.w, acc, pbar, ppar, do_probe);
    else if (two && pat == PAT_32) ok = burst12_ts32(q0.z, q0.x, q0.y, q1.x, q1.y, q0.w, acc, pbar, ppar, do_probe);
    else if (two && pat == PAT_16) ok = burst12_ts16(q0.z, q0.x, q0.y, q1.x, q1.y, q0.w, acc, pbar, ppar, do_probe);
    else ok = issue_burst(pat, two, (c >> 15) & 7u, q0.z, q0.x, q0.y, q1.x, q1.y, q0.w, acc, pbar, ppar, do_probe);
    const uint32_t eb = q1.w & 0xffu, db = (q1.w >> 8) & 0xffu, eb2 = q1.w >> 24;
    if (eb != 0xffu) umma_commit_a(bar_addr(ctl_addr, eb));
    if (eb2 != 0xffu) umma_commit_a(bar_addr(ctl_addr, eb2));
    if (db != 0xffu) umma_commit_a(bar_addr(ctl_addr, db));
    if (do_probe) {
      if (!ok) { mbar_wait_a(pbar, ppar); tc_fence_after_sync(); }
      bits ^= 1u << (7 + nu);
    }
    if (trace) trace[2 * MAX_BURST + i] = clock64();
    if (c & B_PART_NEXT) bits ^= 1u << s;
    q0 = n0;
    q1 = n1;
  }
  bits_io = bits;
}

// TMA producer (one thread): streams the weight images of every acquiring burst of the program into the ring
struct ProducerState { uint32_t ebits, filled; };
__device__ __forceinline__ void produce_program(const TcProgram& P, const uint8_t* wstream, uint8_t* smem, Ctrl* ctl,
                                                ProducerState& st) {
  for (int i = 0; i < P.n_burst; ++i) {
    const uint32_t sw = P.src[i];
    if (!(sw >> 31)) continue;
    const uint32_t u = (sw >> 29) & 3u, bytes = ((sw >> 20) & 0x1ffu) * 128u;
    if ((st.filled >> u) & 1u) {       // the previous fill of this unit has been consumed
      mbar_wait(&ctl->empty[u], (st.ebits >> u) & 1u);
      st.ebits ^= 1u << u;
    }
    st.filled |= 1u << u;
    mbar_arrive_expect_tx(&ctl->full[u], bytes);
    tma_bulk_g2s(smem + OFF_RING + u * UNIT_BYTES, wstream + (size_t)(sw & 0xfffffu) * 128u, bytes, &ctl->full[u]);
  }
}

// write one feature (col c of a hi | lo input block at `blk`) of row r, split
__device__ __forceinline__ void store_in(uint8_t* blk, uint32_t r, uint32_t c, float v) {
  __half h, l;
  split_h(v, h, l);
  const uint32_t o = kblock_offset(r, c);
  *reinterpret_cast<__half*>(blk + o) = h;
  *reinterpret_cast<__half*>(blk + KBLK + o) = l;
}
__device__ __forceinline__ void store_in_hi(uint8_t* blk, uint32_t r, uint32_t c, float v) {
  *reinterpret_cast<__half*>(blk + kblock_offset(r, c)) = __float2half_rn(v);
}

// Epilogue of one N-chunk for this thread's row and its CW-column slice: accumulators -> scale/bias/ReLU ->
// split fp16 -> written over the very columns just read (hi | lo), the operand of the next layer.
template <int CW, int MM = MASK_NONE>
__device__ __forceinline__ void epilogue_chunk(const TcOp& op, int nc, const float* __restrict__ bias_base,
                                               uint32_t tmem_lane, uint32_t row, int sub, int q, float* dbg_out,
                                               int dbg_ld, uint64_t* bar, uint32_t parity, uint32_t* mask_word = nullptr,
                                               unsigned long long* sub_tr = nullptr) {
  const uint32_t oc0 = (uint32_t)nc * op.nc_rows + (uint32_t)sub * CW;   // first output column of this slice
  const float4* bias4 = reinterpret_cast<const float4*>(bias_base + op.bias_off + oc0);
  float b[CW];
#pragma unroll
  for (int i = 0; i < CW / 4; ++i) {
    // issued BEFORE the accumulators are awaited, so that the latency (an L2 round trip: with 224 KB of shared memory
    // the L1 holds next to nothing) hides in the wait.  `volatile`: a plain __ldg is sunk below the wait loop by the
    // compiler, next to its first use, and the first FFMA of every epilogue then stalls on it (ncu: 7.5 % of all
    // warp samples of the kernel on that one instruction).
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(b[4 * i]), "=f"(b[4 * i + 1]), "=f"(b[4 * i + 2]), "=f"(b[4 * i + 3]) : "l"(bias4 + i));
  }
  mbar_wait(bar, parity);
  tc_fence_after_sync();
#if NDS_SUBTRACE
  if (sub_tr) sub_tr[0] = clock64();
#endif
  const uint32_t chunk = tmem_lane + op.d_col[nc];
  const uint32_t col = chunk + (uint32_t)sub * CW;
  uint32_t v[CW];
  tmem_ld<CW>(col, v);
  tmem_ld_wait();
#if NDS_SUBTRACE
  if (sub_tr) sub_tr[1] = clock64();
#endif
  const float inv = op.inv_scale;
  const bool relu = op.relu != 0;
  uint32_t hi[CW / 2], lo[CW / 2];
  uint32_t mword = MM == MASK_APPLY ? *mask_word : 0u;
  if (NDS_EPI_TRUNC && !dbg_out && MM == MASK_NONE) {
    // hi = x truncated to 11 significant bits by clearing mantissa bits (exactly representable in fp16 for
    // 2^-14 <= |x| < 65504, so its conversion is exact; below that the conversion rounds by < 2^-25 absolute),
    // lo = x - hi has the sign of x, so with a ReLU both conversions clamp at 0 and no separate max is needed.
    if (relu) {
#pragma unroll
      for (int i = 0; i < CW / 2; ++i) {
        const float x0 = fmaf(__uint_as_float(v[2 * i]), inv, b[2 * i]);
        const float x1 = fmaf(__uint_as_float(v[2 * i + 1]), inv, b[2 * i + 1]);
        const float h0 = __uint_as_float(__float_as_uint(x0) & 0xffffe000u), h1 = __uint_as_float(__float_as_uint(x1) & 0xffffe000u);
        uint32_t h, l;
        asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(h1), "f"(h0));
        asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(x1 - h1), "f"(x0 - h0));
        hi[i] = h;
        lo[i] = l;
      }
    } else {
#pragma unroll
      for (int i = 0; i < CW / 2; ++i) {
        const float x0 = fmaf(__uint_as_float(v[2 * i]), inv, b[2 * i]);
        const float x1 = fmaf(__uint_as_float(v[2 * i + 1]), inv, b[2 * i + 1]);
        const float h0 = __uint_as_float(__float_as_uint(x0) & 0xffffe000u), h1 = __uint_as_float(__float_as_uint(x1) & 0xffffe000u);
        uint32_t h, l;
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(h1), "f"(h0));
        asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(x1 - h1), "f"(x0 - h0));
        hi[i] = h;
        lo[i] = l;
      }
    }
  } else if (NDS_EPI_RZ && relu && !dbg_out && MM == MASK_NONE) {
    // ReLU folded into the conversions: hi = relu(x) truncated to fp16 (round toward zero, so the residual of a
    // positive x is never negative), lo = relu(x - hi) -- for x < 0 both come out 0.  hi + lo still carries 21+ bits.
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) {
      const float x0 = fmaf(__uint_as_float(v[2 * i]), inv, b[2 * i]);
      const float x1 = fmaf(__uint_as_float(v[2 * i + 1]), inv, b[2 * i + 1]);
      uint32_t h, l;
      asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(x1), "f"(x0));
      const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h));
      asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(x1 - hf.y), "f"(x0 - hf.x));
      hi[i] = h;
      lo[i] = l;
    }
  } else {
#pragma unroll
    for (int i = 0; i < CW / 2; ++i) {
      float x0 = fmaf(__uint_as_float(v[2 * i]), inv, b[2 * i]);
      float x1 = fmaf(__uint_as_float(v[2 * i + 1]), inv, b[2 * i + 1]);
      if (relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
      if (MM == MASK_RECORD) mword |= (x0 > 0.f ? 1u : 0u) << (2 * i) | (x1 > 0.f ? 1u : 0u) << (2 * i + 1);
      if (MM == MASK_APPLY) { x0 = ((mword >> (2 * i)) & 1u) ? x0 : 0.f; x1 = ((mword >> (2 * i + 1)) & 1u) ? x1 : 0.f; }
      if (dbg_out) { dbg_out[row * dbg_ld + oc0 + 2 * i] = x0; dbg_out[row * dbg_ld + oc0 + 2 * i + 1] = x1; }
      const __half2 h = __floats2half2_rn(x0, x1);
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
      hi[i] = *reinterpret_cast<const uint32_t*>(&h);
      lo[i] = *reinterpret_cast<const uint32_t*>(&l);
    }
  }
  if (MM == MASK_RECORD) *mask_word = mword;
  const uint8_t kind = op.epi_kind;
  if (kind == EPI_COMPACT_HI) {
    // the compacted slice overlaps columns other warps of this lane quarter are still reading
    tc_fence_before_sync();
    asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
    tc_fence_after_sync();
    tmem_st<CW / 2>(chunk + (uint32_t)sub * (CW / 2), hi);
  } else {
    tmem_st<CW / 2>(col, hi);
    if (kind == EPI_INPLACE) tmem_st<CW / 2>(col + CW / 2, lo);
  }
#if NDS_SUBTRACE
  if (sub_tr) sub_tr[2] = clock64();
#endif
}

// waits for the chunk's accumulators (bar / parity) inside, after the bias prefetch
__device__ __forceinline__ void epilogue_dispatch(const TcOp& op, int nc, const float* bias_base, uint32_t tmem_lane,
                                                  uint32_t row, int sub, int q, float* dbg, int dbg_ld, uint64_t* bar,
                                                  uint32_t parity, uint32_t* mask_word = nullptr,
                                                  unsigned long long* sub_tr = nullptr) {
  if (op.nc_rows == 128) epilogue_chunk<32>(op, nc, bias_base, tmem_lane, row, sub, q, dbg, dbg_ld, bar, parity, nullptr, sub_tr);
  else epilogue_chunk<16>(op, nc, bias_base, tmem_lane, row, sub, q, dbg, dbg_ld, bar, parity, nullptr, sub_tr);
}
// reverse-sweep instantiation: the epilogue records (forward) or applies (reverse) a ReLU mask word
__device__ __forceinline__ void epilogue_dispatch_mask(const TcOp& op, int nc, const float* bias_base, uint32_t tmem_lane,
                                                       uint32_t row, int sub, int q, uint64_t* bar, uint32_t parity,
                                                       uint32_t* mask_word) {
  const bool wide = op.nc_rows == 128;
  if (op.mask_mode == MASK_RECORD) {
    if (wide) epilogue_chunk<32, MASK_RECORD>(op, nc, bias_base, tmem_lane, row, sub, q, nullptr, 0, bar, parity, mask_word);
    else epilogue_chunk<16, MASK_RECORD>(op, nc, bias_base, tmem_lane, row, sub, q, nullptr, 0, bar, parity, mask_word);
  } else if (op.mask_mode == MASK_APPLY) {
    if (wide) epilogue_chunk<32, MASK_APPLY>(op, nc, bias_base, tmem_lane, row, sub, q, nullptr, 0, bar, parity, mask_word);
    else epilogue_chunk<16, MASK_APPLY>(op, nc, bias_base, tmem_lane, row, sub, q, nullptr, 0, bar, parity, mask_word);
  } else {
    if (wide) epilogue_chunk<32>(op, nc, bias_base, tmem_lane, row, sub, q, nullptr, 0, bar, parity, nullptr);
    else epilogue_chunk<16>(op, nc, bias_base, tmem_lane, row, sub, q, nullptr, 0, bar, parity, nullptr);
  }
}

// head (<= 16 outputs): every compute warp of the lane quarter reads all of them.  The bias (padded to 16) is loaded
// by head_bias BEFORE the accumulators are awaited (see epilogue_chunk).
__device__ __forceinline__ void head_bias(const TcOp& op, const float* __restrict__ bias_base, float (&b)[16]) {
  const float4* bias4 = reinterpret_cast<const float4*>(bias_base + op.bias_off);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(b[4 * i]), "=f"(b[4 * i + 1]), "=f"(b[4 * i + 2]), "=f"(b[4 * i + 3]) : "l"(bias4 + i));
}
__device__ __forceinline__ void epilogue_head(const TcOp& op, const float (&b)[16], uint32_t tmem_lane, float* hv) {
  uint32_t v[16];
  tmem_ld16(tmem_lane + op.d_col[0], v);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 16; ++i) hv[i] = fmaf(__uint_as_float(v[i]), op.inv_scale, b[i]);
}

// sin(x) for the positional encodings (|x| <= 2^max_deg * scene extent, far below the 1e5 limit of the 3-term
// Cody-Waite reduction): reduce by pi/2, then the cephes sinf / cosf minimax polynomials on [-pi/4, pi/4].
// ~1 ulp; no slow path, so it stays inline and branch-free.
__device__ __forceinline__ float pe_sin(float x) {
  const float j = rintf(x * 0.636619772367581343f);
  float r = fmaf(j, -1.57079601287841796875f, x);
  r = fmaf(j, -3.1391647326017846e-07f, r);
  r = fmaf(j, -5.390302529957764e-15f, r);
  const int q = (int)j;
  const float r2 = r * r;
  float sn = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  sn = fmaf(sn, r2, -1.6666654611e-1f);
  sn = fmaf(sn * r2, r, r);
  float cs = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cs = fmaf(cs, r2, 4.166664568298827e-2f);
  cs = fmaf(cs * r2, r2, fmaf(r2, -0.5f, 1.0f));
  const float v = (q & 1) ? cs : sn;
  return (q & 2) ? -v : v;
}

// cos(x), same reduction (the reverse sweep's derivative of the encodings)
__device__ __forceinline__ float pe_cos(float x) {
  const float j = rintf(x * 0.636619772367581343f);
  float r = fmaf(j, -1.57079601287841796875f, x);
  r = fmaf(j, -3.1391647326017846e-07f, r);
  r = fmaf(j, -5.390302529957764e-15f, r);
  const int q = (int)j;
  const float r2 = r * r;
  float sn = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  sn = fmaf(sn, r2, -1.6666654611e-1f);
  sn = fmaf(sn * r2, r, r);
  float cs = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cs = fmaf(cs, r2, 4.166664568298827e-2f);
  cs = fmaf(cs * r2, r2, fmaf(r2, -0.5f, 1.0f));
  const float v = (q & 1) ? sn : cs;             // cos(r + q pi/2): cos, -sin, -cos, sin
  return ((q + 1) & 2) ? -v : v;
}

// positional encoding (model_utils.py:398-417): feature layout (F, 2, C) flattened, identity first.  The
// (sin, cos) pairs are dealt round-robin to the NSUB warps that share a sample: this warp evaluates the pairs p
// with (pair + p) % NSUB == sub.  `pair` is the running pair counter, `o` the running feature offset.
template <typename Store>
__device__ __forceinline__ int posenc_emit_sub(float x0, float x1, float x2, int C, const PosencSpec& pe, Store store,
                                               int o, int sub, int& pair) {
  if (pe.identity) {
    if (sub == 0) { store(o, x0); if (C > 1) store(o + 1, x1); if (C > 2) store(o + 2, x2); }
    o += C;
  }
  const int npair = pe.num_bands * C;
  for (int p = (sub - pair) & (NSUB - 1); p < npair; p += NSUB) {
    const int k = C == 3 ? p / 3 : (C == 2 ? p >> 1 : p);
    const int c = p - k * C;
    const float xv = c == 0 ? x0 : (c == 1 ? x1 : x2);
    const float xb = xv * __int_as_float((127 + pe.min_deg + k) << 23);    // x * 2^(min_deg + k), exact
    const float w = pe.window[k];
    store(o + 2 * C * k + c, w * pe_sin(xb));
    store(o + 2 * C * k + C + c, w * pe_sin(xb + NDS_HALF_PI_F));
  }
  pair += npair;
  return o + 2 * npair;
}


// Reverse sweep: write CW columns of a gradient seed (already masked) as the split-fp16 operand slice an epilogue
// would leave at `col`: hi halves in the first CW / 2 columns, lo halves after them.
template <int CW, typename G>
__device__ __forceinline__ void seed_slice(uint32_t col, uint32_t mword, G g_of) {
  uint32_t hi[CW / 2], lo[CW / 2];
#pragma unroll
  for (int i = 0; i < CW / 2; ++i) {
    const float x0 = ((mword >> (2 * i)) & 1u) ? g_of(2 * i) : 0.f;
    const float x1 = ((mword >> (2 * i + 1)) & 1u) ? g_of(2 * i + 1) : 0.f;
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi[i] = *reinterpret_cast<const uint32_t*>(&h);
    lo[i] = *reinterpret_cast<const uint32_t*>(&l);
  }
  tmem_st<CW / 2>(col, hi);
  tmem_st<CW / 2>(col + CW / 2, lo);
}

// Reverse sweep: one feature column of a positional encoding this thread differentiates (fixed for the whole launch)
struct PeCol {
  float scale;     // window x 2^deg (band) | 1 (identity) | 0 (column carries no gradient w.r.t. a coordinate)
  float freq;      // 2^deg (band), 0 otherwise
  float phase;     // pi/2 for the cos feature
  int comp;        // coordinate: 0..2 point, 3..4 hyper; kind in bits 8+: 0 none, 1 identity, 2 band
};

// ---------------------------------------------------------------------------
// the field kernel
// ---------------------------------------------------------------------------
struct TcKernelArgs {
  TcLevel lvl;
  const float* warp_embed;
  const float* mask_embed;
  // reverse sweep seeds (fp32, Flax layout): column 0 of the sigma head [trunk width], hyper-sheet logit kernel
  // [width, H], SE(3) branch kernels [width, 3] each
  const float* alpha_col0;
  const float* hyper_logit_w;
  const float* warp_w_w;
  const float* warp_v_w;
  unsigned long long* trace;   // diagnostics (NDS_TC_TRACE): clock64 stamps of CTA 0's second pair, else null
};
// trace layout: [i] burst i reached | [MAX_BURST + i] its dependencies satisfied | [2 MAX_BURST + i] issued |
// [3 MAX_BURST + 2 s] compute step s started, [.. + 1] finished | [3 MAX_BURST + 2 MAX_STEPS] pair start

// per-sample state of one tile slot (lives in local memory: it is touched only by the per-sample stages)
struct TileState {
  float x[3], vd[3], gt;
  uint32_t wid;
  int valid;
  int64_t n_out;          // sample index inside the level (planes / carry_out position)
  float xw[3], om[2], maskv, pmask, sigma_raw, nrm[3], rgb[3];
  float R[9], p[3];
};
// reverse-sweep state of one tile slot (only in the kernel instantiation that runs the reverse sweep)
struct GradState {
  float wv[6];            // raw SE(3) branch outputs (w, v)
  float fb[16];           // this thread's 16 columns of the input gradient being accumulated
  float gxw[3], gom[2];   // d sigma_raw / d (warped point, hyper coordinates)
  uint32_t masks[MASK_WORDS];
};
struct NoGradState {};
struct NextSample {
  float x[3], vd[3], gt;
  uint32_t wid;
  int valid;
  int64_t n_out;
  float xw[3], om[2], pmask, R[9], p[3];   // carried launches only
};

template <bool GRAD>
__global__ void __launch_bounds__(TC_THREADS, 1)
field_tc_kernel(const __grid_constant__ TcProgram P, const __grid_constant__ TcKernelArgs K,
                const __grid_constant__ CallParams cp, const __grid_constant__ FieldArgs a,
                const __grid_constant__ ndsr_config cfg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Ctrl* ctl = reinterpret_cast<Ctrl*>(smem + OFF_CTRL);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = cfg.use_hyper_sheet ? cfg.hyper_num_dims : 0;

  // input blocks start as zeros: columns beyond the written features multiply zero weight rows, but 0 x garbage
  // could be NaN
  for (uint32_t i = threadIdx.x; i < OFF_RING / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  TcOp* c_ops = reinterpret_cast<TcOp*>(reinterpret_cast<uint8_t*>(ctl) + CTRL_HDR);
  Step* c_steps = reinterpret_cast<Step*>(reinterpret_cast<uint8_t*>(ctl) + ctrl_off_steps(P.n_ops));
  uint32_t* c_burst = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(ctl) + ctrl_off_burst(P.n_ops, P.n_steps));
  for (int i = threadIdx.x; i < P.n_ops * (int)(sizeof(TcOp) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(c_ops)[i] = reinterpret_cast<const uint32_t*>(P.ops)[i];
  for (int i = threadIdx.x; i < P.n_steps; i += blockDim.x)
    reinterpret_cast<uint32_t*>(c_steps)[i] = reinterpret_cast<const uint32_t*>(P.steps)[i];
  for (int i = threadIdx.x; i < P.n_burst * 8; i += blockDim.x)
    c_burst[i] = reinterpret_cast<const uint32_t*>(P.burst)[i];
  if (threadIdx.x == 0) ctrl_init(ctl);
  if (warp == WARP_MMA) tmem_alloc(&ctl->tmem_base, 512);
  fence_proxy_async_smem();
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;
  // the burst program holds absolute shared-memory / tensor-memory addresses
  if (smem_base != P.smem_base || tmem_base != 0u) {
    if (threadIdx.x == 0) printf("nerfds_b200: tensor-core program encoded for smem base %u, tmem base 0; got %u, %u\n", P.smem_base, smem_base, tmem_base);
    __trap();
  }

  // (early termination: the surviving samples were counted on the device by an earlier kernel of the stream)
  const int64_t n_total = a.n_active ? (int64_t)__ldg(a.n_active) : a.n_samples_total;
  const int64_t n_tiles = (n_total + TM - 1) / TM;
  const int64_t n_pairs = (n_tiles + 1) / 2;
  const TcLevel& L = K.lvl;

  if (warp == WARP_TMA) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      ProducerState ps{0u, 0u};
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) produce_program(P, L.weights, smem, ctl, ps);
    }
    __syncwarp();
  } else if (warp == WARP_MMA) {
    // ===================== MMA issuer =====================
    if (elect_one_sync()) {
      uint32_t bits = 0;
      if ((int64_t)blockIdx.x < n_pairs) {   // the very first burst of the kernel: nobody waited for its weights yet
        // (a CTA without a pair -- the early-termination scan left fewer tiles than the grid was sized for -- has no
        //  producer activity to wait for)
        const uint32_t s0 = P.src[0], u0 = (s0 >> 29) & 3u;
        if (s0 >> 31) { mbar_wait(&ctl->full[u0], 0u); bits ^= 1u << (8 + u0); }
      }
      for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x)
        issue_program(P, ctl, bits, pair + gridDim.x < n_pairs, (K.trace && pair == (int64_t)gridDim.x) ? K.trace : nullptr);
    }
    __syncwarp();
  } else if (is_compute_warp(warp)) {
    // ===================== compute warps =====================
    const int q = warp & 3, sub = (warp - CWARP0) >> 2;
    const uint32_t row = (uint32_t)q * 32u + (uint32_t)lane;
    const uint32_t tmem_lane = tmem_base + (((uint32_t)q * 32u) << 16);
    uint32_t dc = 0;      // parity of d_full[s][c]: bit 2 s + c
    TileState cur[2];
    NextSample nxt[2];
    typename std::conditional<GRAD, GradState, NoGradState>::type gst[2];
    // per-sample inputs are fetched one pair ahead, so their global-memory latency hides behind the T phase
    auto load_next = [&](int64_t pair_, int s_) {
      NextSample& ns = nxt[s_];
      const int64_t li = (2 * pair_ + s_) * TM + row;      // sample of the launch
      ns.valid = (pair_ < n_pairs && li < n_total) ? 1 : 0;
      ns.x[0] = ns.x[1] = ns.x[2] = 0.f; ns.vd[0] = ns.vd[1] = ns.vd[2] = 0.f; ns.gt = 0.f; ns.wid = 0; ns.n_out = 0;
      ns.xw[0] = ns.xw[1] = ns.xw[2] = 0.f; ns.om[0] = ns.om[1] = 0.f; ns.pmask = 0.f;
      for (int i = 0; i < 9; ++i) ns.R[i] = (i % 4 == 0) ? 1.f : 0.f;
      ns.p[0] = ns.p[1] = ns.p[2] = 0.f;
      if (!ns.valid) return;
      const int64_t n_ = a.index ? (int64_t)__ldg(a.index + li) : li, ray = n_ / a.S;
      ns.n_out = n_;
      ns.vd[0] = a.viewdirs[ray * 3]; ns.vd[1] = a.viewdirs[ray * 3 + 1]; ns.vd[2] = a.viewdirs[ray * 3 + 2];
      if (a.carry) {
        // all C_COUNT loads in flight before the first one is consumed
        const float* C = a.carry + li;
        const int64_t cs = a.carry_stride;
        float t[C_COUNT];
#pragma unroll
        for (int i = 0; i < C_COUNT; ++i) t[i] = __ldg(C + i * cs);
#pragma unroll
        for (int c = 0; c < 3; ++c) { ns.xw[c] = t[C_WARPED + c]; ns.p[c] = t[C_P + c]; }
        ns.om[0] = t[C_WARPED + 3]; ns.om[1] = t[C_WARPED + 4];
        ns.pmask = t[C_MASK];
#pragma unroll
        for (int i = 0; i < 9; ++i) ns.R[i] = t[C_R + i];
        return;
      }
      if (a.points) { ns.x[0] = a.points[n_ * 3]; ns.x[1] = a.points[n_ * 3 + 1]; ns.x[2] = a.points[n_ * 3 + 2]; }
      else {
        const float z = a.z[n_];
        ns.x[0] = a.origins[ray * 3 + 0] + z * a.dirs[ray * 3 + 0];
        ns.x[1] = a.origins[ray * 3 + 1] + z * a.dirs[ray * 3 + 1];
        ns.x[2] = a.origins[ray * 3 + 2] + z * a.dirs[ray * 3 + 2];
      }
      if (a.warp_id) ns.wid = a.warp_id[ray];
      if (a.gt_mask) ns.gt = a.gt_mask[ray];
    };
    // trunk input (models.py:493-523) of one tile: posenc of the warped point and the hyper coordinates, the
    // (sin, cos) pairs dealt to the 4 warps sharing a sample; `part` < 0: all of it, else one of NPREP parts
    auto trunk_input = [&](uint8_t* blk, const float* xw, const float* om, int part) {
      auto st = [&](int c, float v) { store_in(blk, row, (uint32_t)c, v); };
      const float v0 = xw[0], v1 = xw[1], v2 = xw[2], h0 = om[0], h1 = om[1];
      const PosencSpec& ps = cp.pe_spatial;
      const PosencSpec& ph = cp.pe_hyperpt;
      const int o_sp = ps.identity ? 3 : 0, n_sp = ps.num_bands * 3;
      const int o_h0 = o_sp + 2 * n_sp, o_hy = o_h0 + ((H > 0 && ph.identity) ? H : 0), n_hy = H > 0 ? ph.num_bands * H : 0;
      if (part <= 0 && sub == 0) {
        if (ps.identity) { st(0, v0); st(1, v1); st(2, v2); }
        if (H > 0 && ph.identity) { st(o_h0, h0); if (H > 1) st(o_h0 + 1, h1); }
      }
      // global pair index gp = sub + 4 m: this warp's pairs; part p takes the m with m % NPREP == p
      const int m0 = part < 0 ? 0 : part, dm = part < 0 ? 1 : NPREP;
      for (int m = m0; sub + NSUB * m < n_sp + n_hy; m += dm) {
        const int gp = sub + NSUB * m;
        float xv, w;
        int deg, col, stride;
        if (gp < n_sp) {
          const int k = gp / 3, c = gp - 3 * k;
          xv = c == 0 ? v0 : (c == 1 ? v1 : v2);
          deg = ps.min_deg + k; w = ps.window[k]; col = o_sp + 6 * k + c; stride = 3;
        } else {
          const int q = gp - n_sp, k = H == 2 ? (q >> 1) : q, c = q - H * k;
          xv = c == 0 ? h0 : h1;
          deg = ph.min_deg + k; w = ph.window[k]; col = o_hy + 2 * H * k + c; stride = H;
        }
        const float xb = xv * __int_as_float((127 + deg) << 23);    // x * 2^deg, exact
        st(col, w * pe_sin(xb));
        st(col + stride, w * pe_sin(xb + NDS_HALF_PI_F));
      }
      if (part <= 0) for (int c = o_hy + 2 * n_hy + sub; c < P.f_cols; c += NSUB) st(c, 0.f);     // the feature block was wider
    };
    // Part `part` of the shared feature block of the NEXT tile of slot s_ (model_utils.py:398-417 without the
    // window, which is folded into the weights): the (sin, cos) pairs are dealt to the 4 warps sharing a sample
    // and to the NPREP parts; part 0 also writes the embeddings.
    auto prep_part = [&](int s_, int part) {
      const NextSample& ns = nxt[s_];
      uint8_t* blk = smem + OFF_IN + (uint32_t)s_ * 2u * KBLK;
      if (P.carried) {          // no N phase: the next tile's trunk input straight from the carried results
        trunk_input(blk, ns.xw, ns.om, part);
        if (part == NPREP - 1) warp_arrive(&ctl->prep[s_], lane);
        return;
      }
      const int npair = P.f_nb * 3;
      for (int pi = sub; pi < npair; pi += NSUB) {
        if (((pi >> 2) % NPREP) != part) continue;
        const int k = pi / 3, c = pi - 3 * k;
        const float xv = c == 0 ? ns.x[0] : (c == 1 ? ns.x[1] : ns.x[2]);
        const float xb = xv * __int_as_float((127 + P.f_kmin + k) << 23);    // x * 2^band, exact
        const int col = P.f_col_bands + 6 * k + c;
        store_in(blk, row, (uint32_t)col, pe_sin(xb));
        store_in(blk, row, (uint32_t)(col + 3), pe_sin(xb + NDS_HALF_PI_F));
      }
      if (part == 0) {
        if (sub == 0 && P.f_col_ident >= 0) for (int c = 0; c < 3; ++c) store_in(blk, row, (uint32_t)(P.f_col_ident + c), ns.x[c]);
        if (sub == 1 && P.f_col_wembed >= 0)
          for (int e = 0; e < cfg.warp_embed_dims; ++e)
            store_in(blk, row, (uint32_t)(P.f_col_wembed + e), __ldg(K.warp_embed + (size_t)ns.wid * cfg.warp_embed_dims + e));
        if (sub == 2 && P.f_col_membed >= 0)
          for (int e = 0; e < cfg.mask_embed_dims; ++e)
            store_in(blk, row, (uint32_t)(P.f_col_membed + e), __ldg(K.mask_embed + (size_t)ns.wid * cfg.mask_embed_dims + e));
        if (sub == 3) {
          // without a predicted mask the mask feature is the ground-truth mask (models.py:975 with mask_ratio 0)
          if (P.f_col_mask >= 0 && !cfg.use_predicted_mask) store_in(blk, row, (uint32_t)P.f_col_mask, ns.gt);
          for (int c = P.f_cols; c < P.t_cols; ++c) store_in(blk, row, (uint32_t)c, 0.f);   // left by the trunk input
        }
      }
      if (part == NPREP - 1) warp_arrive(&ctl->prep[s_], lane);
    };

    // reverse sweep: which coordinate each of this thread's 16 feature columns encodes is a function of (column slice,
    // i) only: decoded on the fly from warp-uniform integers (a per-thread table would live in local memory, and with
    // 224 KB of shared memory the L1 holds next to nothing: every local access is an L2 round trip)
    const int sub_u = __shfl_sync(0xffffffffu, sub, 0);
    load_next(blockIdx.x, 0);
    load_next(blockIdx.x, 1);
    for (int s = 0; s < 2; ++s) for (int part = 0; part < NPREP; ++part) prep_part(s, part);
    warp_arrive(&ctl->done[1], lane);     // nothing occupies tensor memory before the first pair
    if (GRAD) warp_arrive(&ctl->done[0], lane);     // (reverse-sweep programs also wait for this slot's previous pair)

    for (int64_t pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
      for (int s = 0; s < 2; ++s) {
        TileState& T = cur[s];
        const NextSample& ns = nxt[s];
        for (int c = 0; c < 3; ++c) { T.x[c] = ns.x[c]; T.vd[c] = ns.vd[c]; T.xw[c] = 0.f; T.nrm[c] = 0.f; T.rgb[c] = 0.f; T.p[c] = 0.f; }
        for (int i = 0; i < 9; ++i) T.R[i] = (i % 4 == 0) ? 1.f : 0.f;
        T.gt = ns.gt; T.wid = ns.wid; T.valid = ns.valid; T.n_out = ns.n_out;
        T.om[0] = T.om[1] = 0.f; T.maskv = ns.gt; T.pmask = 0.f; T.sigma_raw = 0.f;
        if (P.carried) {
          for (int c = 0; c < 3; ++c) { T.xw[c] = ns.xw[c]; T.p[c] = ns.p[c]; }
          for (int i = 0; i < 9; ++i) T.R[i] = ns.R[i];
          T.om[0] = ns.om[0]; T.om[1] = ns.om[1]; T.pmask = ns.pmask;
        }
      }
      if (a.carry) {
        // the carry planes of the next pair's tiles (written by the coarse pass long ago: DRAM) -> L2, so that the
        // loads in the VIEW steps do not sit on the critical path of the trunk
        for (int s = 0; s < 2; ++s) {
          const int64_t li = (2 * (pair + gridDim.x) + s) * TM + row;
          if (li < n_total && (sub == s || sub == s + 2))
            for (int i = (sub >> 1); i < C_COUNT; i += 2)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(a.carry + li + (int64_t)i * a.carry_stride));
        }
      }
      unsigned long long* tr = (K.trace && pair == (int64_t)gridDim.x && threadIdx.x == CWARP0 * 32) ? K.trace + 3 * MAX_BURST : nullptr;
      if (tr) tr[2 * MAX_STEPS] = clock64();
      for (int si = 0; si < P.n_steps; ++si) {
        const Step sp = c_steps[si];
        const int s = sp.tslot;
        TileState& T = cur[s];
        uint8_t* blk = smem + OFF_IN + (uint32_t)s * 2u * KBLK;
        auto st_in = [&](int c, float v) { store_in(blk, row, (uint32_t)c, v); };
        auto st_in2 = [&](int c, float v) { store_in_hi(smem + OFF_IN2, row, (uint32_t)c, v); };
        if (tr) tr[2 * si] = clock64();
        if (sp.kind == STEP_EPI) {
          const TcOp& op = c_ops[sp.op];
          for (int nc = 0; nc < op.n_nc; ++nc) {
            const uint32_t par = (dc >> (2 * s + nc)) & 1u;
            dc ^= 1u << (2 * s + nc);
            if constexpr (GRAD) epilogue_dispatch_mask(op, nc, L.bias, tmem_lane, row, sub, q, &ctl->d_full[s][nc], par, &gst[s].masks[(op.mask_idx + nc) & (MASK_WORDS - 1)]);
            else epilogue_dispatch(op, nc, L.bias, tmem_lane, row, sub, q, nullptr, 0, &ctl->d_full[s][nc], par, nullptr,
                                   (NDS_SUBTRACE && tr) ? K.trace + TRACE_SUB + 8 * si + 4 * nc : nullptr);
            // (a plain epilogue only wrote tensor memory: no proxy fence)
            warp_arrive<false>(&ctl->part[s][nc], lane);
            if (op.n_nc == 1) warp_arrive<false>(&ctl->part[s][1], lane);   // keeps both barriers on one phase per op
#if NDS_SUBTRACE
            if (tr) K.trace[TRACE_SUB + 8 * si + 4 * nc + 3] = clock64();
#endif
          }
        } else if (sp.kind == STEP_HEAD) {
          const TcOp& op = c_ops[sp.op];
          float hb[16];
          head_bias(op, L.bias, hb);
          mbar_wait(&ctl->d_full[s][0], (dc >> (2 * s)) & 1u);
          dc ^= 1u << (2 * s);
          tc_fence_after_sync();
          float hv[16];
          epilogue_head(op, hb, tmem_lane, hv);
          if (op.signal_done) warp_arrive(&ctl->done[s], lane);   // the tile slot's last tensor-memory read
          switch (op.glue) {
            case GLUE_MASK: {            // models.py:967-975
              T.pmask = cfg.mask_output_relu ? fmaxf(hv[0], 0.f) : hv[0];
              T.maskv = a.gt_mask ? (T.pmask * cp.mask_ratio + T.gt * (1.f - cp.mask_ratio)) : T.pmask * cp.mask_ratio;
              if (P.f_col_mask >= 0 && sub == 2) st_in(P.f_col_mask, T.maskv);
            } break;
            case GLUE_HYPER: {
              for (int c = 0; c < H; ++c) T.om[c] = hv[c];
            } break;
            case GLUE_WARP: {            // warping.py:217-232
              SE3<float> se;
              exp_se3<float>(hv, hv + 3, se);
              float xw[3];
              for (int c = 0; c < 3; ++c) xw[c] = se.R[c * 3 + 0] * T.x[0] + se.R[c * 3 + 1] * T.x[1] + se.R[c * 3 + 2] * T.x[2] + se.p[c];
              for (int i = 0; i < 9; ++i) T.R[i] = se.R[i];
              for (int c = 0; c < 3; ++c) { T.p[c] = se.p[c]; T.xw[c] = xw[c]; }
              if constexpr (GRAD) { for (int i = 0; i < 6; ++i) gst[s].wv[i] = hv[i]; }
              // trunk input over the feature block, which the narrow networks are done with
              trunk_input(blk, xw, T.om, -1);
            } break;
            case GLUE_ALPHA: {
              T.sigma_raw = hv[0];
              if (cfg.predict_norm) { T.nrm[0] = hv[1]; T.nrm[1] = hv[2]; T.nrm[2] = hv[3]; }
              if (P.full && cp.use_predicted_norm) {
                // rgb branch side input: normal in the observation frame (models.py:1117-1148), after the viewdir feats
                float nh[3], ni[3];
                const float nr[3] = {T.nrm[0], T.nrm[1], T.nrm[2]};
                normalize3(nr, nh);
                if (cfg.use_warp) { for (int c = 0; c < 3; ++c) ni[c] = T.R[0 * 3 + c] * nh[0] + T.R[1 * 3 + c] * nh[1] + T.R[2 * 3 + c] * nh[2]; }
                else { ni[0] = nh[0]; ni[1] = nh[1]; ni[2] = nh[2]; }
                normalize3(ni, nh);
                const int o = cfg.use_viewdirs ? cp.pe_view.dim(3) : 0;
                int pr = 0;
                if (cfg.norm_input_posenc) posenc_emit_sub(nh[0], nh[1], nh[2], 3, cp.pe_norm, st_in2, o, sub, pr);
                else if (sub == 0) { st_in2(o, nh[0]); st_in2(o + 1, nh[1]); st_in2(o + 2, nh[2]); }
              }
            } break;
            case GLUE_RGB: {
              for (int c = 0; c < 3; ++c) T.rgb[c] = 1.f / (1.f + __expf(-hv[c]));
            } break;
            default: break;
          }
          // the MMA issuer may overwrite the head accumulators / read the new inputs from here on
          if (op.signal_glue) warp_arrive(&ctl->glue[s], lane);
        } else if (sp.kind == STEP_VIEW) {      // arg: 0 = both halves, 1 = sample fetch only, 2 = viewdir features only
          if (sp.arg != 2) load_next(pair + gridDim.x, s);
          if (sp.arg != 1 && P.full && cfg.use_viewdirs) {
            int pr = 0;
            posenc_emit_sub(T.vd[0], T.vd[1], T.vd[2], 3, cp.pe_view, st_in2, 0, sub, pr);
          }
        } else if (sp.kind == STEP_PREP) {
          prep_part(s, sp.arg);
        } else if (GRAD && sp.kind >= STEP_INGRAD) {
         if constexpr (GRAD) {
          GradState& GS = gst[s];
          if (sp.kind == STEP_INGRAD) {
          // ---- reverse sweep: 64 input-gradient columns, 16 per thread, summed over the ops of a network ----
          const TcOp& op = c_ops[sp.op];
          mbar_wait(&ctl->d_full[s][0], (dc >> (2 * s)) & 1u);
          dc ^= 1u << (2 * s);
          tc_fence_after_sync();
          uint32_t v[16];
          tmem_ld16(tmem_lane + op.d_col[0] + (uint32_t)sub * 16u, v);
          tmem_ld_wait();
          if (op.signal_done) warp_arrive(&ctl->done[s], lane);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float g = __uint_as_float(v[i]) * op.inv_scale;
            GS.fb[i] = sp.arg ? g : GS.fb[i] + g;
          }
          if (op.signal_glue) warp_arrive(&ctl->glue[s], lane);
          } else if (sp.kind == STEP_SEED) {
          // ---- reverse sweep: gradient at the last hidden layer of a network, masked by its ReLU, as the operand
          //      image an epilogue of that width would leave (App. E steps 1, 4, 5) ----
          const TcOp& op = c_ops[sp.op];            // the first backward op of the network: reads what is written here
          const int n_nc = op.n_nc;
          // the operand region of that op = the region it does NOT accumulate into; its mask = the ReLU of the network's
          // LAST hidden layer (the op itself applies the one before)
          const uint32_t seed_col = tmem_lane + (n_nc == 2 ? 256u - op.d_col[0] : 256u * (uint32_t)s + (128u - (op.d_col[0] - 256u * (uint32_t)s)));
          const uint32_t* mw = &GS.masks[(op.mask_idx + n_nc) & (MASK_WORDS - 1)];
          if (sp.arg == SEED_TRUNK) {
            // column 0 of the sigma head (App. E step 1)
            for (int nc = 0; nc < 2; ++nc) {
              const float* w = K.alpha_col0 + nc * 128 + sub * 32;
              seed_slice<32>(seed_col + (uint32_t)nc * 128u + (uint32_t)sub * 32u, mw[nc], [&](int i) { return __ldg(w + i); });
            }
          } else if (sp.arg == SEED_HYPER) {
            // d / d(last hidden) = sum_o gom[o] W_logit[:, o] (App. E step 4)
            const float g0 = GS.gom[0], g1 = H > 1 ? GS.gom[1] : 0.f;
            const float* w = K.hyper_logit_w + (size_t)(sub * 16) * H;
            seed_slice<16>(seed_col + (uint32_t)sub * 16u, mw[0], [&](int i) {
              float g = g0 * __ldg(w + i * H);
              if (H > 1) g = fmaf(g1, __ldg(w + i * H + 1), g);
              return g;
            });
          } else {
            // x' = R(w, v) x + p(w, v) (App. E step 5).  The rotation inputs by forward mode over the literal
            // rigid_body.py expressions (3 tangents); the translation inputs enter linearly, p = M(w) v / theta:
            // d x' / d v_j = M[:, j] / theta
            float gwv[6];
            {
              DualN<3> dw[3], dv[3];
              for (int i = 0; i < 3; ++i) { dw[i] = DualN<3>::var(GS.wv[i], i); dv[i] = DualN<3>(GS.wv[3 + i]); }
              SE3<DualN<3>> TD;
              exp_se3<DualN<3>>(dw, dv, TD);
              for (int k = 0; k < 3; ++k) gwv[k] = 0.f;
              for (int i = 0; i < 3; ++i) {
                const DualN<3> xi = TD.R[i * 3 + 0] * T.x[0] + TD.R[i * 3 + 1] * T.x[1] + TD.R[i * 3 + 2] * T.x[2] + TD.p[i];
                for (int k = 0; k < 3; ++k) gwv[k] = fmaf(GS.gxw[i], xi.d[k], gwv[k]);
              }
              const float a0 = GS.wv[0], a1 = GS.wv[1], a2 = GS.wv[2];
              const float theta = sqrtf(a0 * a0 + a1 * a1 + a2 * a2);
              const float w0 = a0 / theta, w1 = a1 / theta, w2 = a2 / theta;
              const float W[9] = {0.f, -w2, w1, w2, 0.f, -w0, -w1, w0, 0.f};
              const float st = sinf(theta), ct = cosf(theta), omc = 1.f - ct, tms = theta - st;
              for (int j = 0; j < 3; ++j) {
                float acc = 0.f;
                for (int i = 0; i < 3; ++i) {
                  const float ww = W[i * 3 + 0] * W[0 * 3 + j] + W[i * 3 + 1] * W[1 * 3 + j] + W[i * 3 + 2] * W[2 * 3 + j];
                  const float m = ((i == j) ? theta : 0.f) + omc * W[i * 3 + j] + tms * ww;
                  acc = fmaf(GS.gxw[i], m, acc);
                }
                gwv[3 + j] = acc / theta;
              }
            }
            const float* ww = K.warp_w_w + (size_t)(sub * 32) * 3;
            const float* wv = K.warp_v_w + (size_t)(sub * 32) * 3;
            seed_slice<32>(seed_col + (uint32_t)sub * 32u, mw[0], [&](int i) {
              float g = gwv[0] * __ldg(ww + 3 * i);
              g = fmaf(gwv[1], __ldg(ww + 3 * i + 1), g); g = fmaf(gwv[2], __ldg(ww + 3 * i + 2), g);
              g = fmaf(gwv[3], __ldg(wv + 3 * i), g); g = fmaf(gwv[4], __ldg(wv + 3 * i + 1), g); g = fmaf(gwv[5], __ldg(wv + 3 * i + 2), g);
              return g;
            });
          }
          warp_arrive(&ctl->part[s][0], lane);
          warp_arrive(&ctl->part[s][1], lane);
          } else if (sp.kind == STEP_PEB) {
          // ---- reverse sweep: positional-encoding backward (App. E step 3) of this thread's 16 feature columns, summed
          //      over the 4 warps that share the sample through the (idle) rgb side-input block ----
          float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f, p4 = 0.f;   // d/d(x0, x1, x2) and, trunk input only, d/d(hyper 0, 1)
          {
            const bool trunk = sp.arg == 0;
            const float c0 = trunk ? T.xw[0] : T.x[0], c1 = trunk ? T.xw[1] : T.x[1], c2 = trunk ? T.xw[2] : T.x[2];
            const float c3 = T.om[0], c4 = T.om[1];
            float fbv[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) fbv[i] = GS.fb[i];           // all 16 loads in flight together
            const PosencSpec& ps = cp.pe_spatial;
            const PosencSpec& ph = cp.pe_hyperpt;
            // column ranges (warp-uniform): [identity | bands] of the point, then of the hyper coordinates (trunk only)
            const int o_id = trunk ? (ps.identity ? 0 : -1) : P.f_col_ident, o_b = trunk ? (ps.identity ? 3 : 0) : P.f_col_bands;
            const int n_b = 6 * (trunk ? ps.num_bands : P.f_nb), deg0 = trunk ? ps.min_deg : P.f_kmin;
            const int o_hid = (trunk && H > 0 && ph.identity) ? o_b + n_b : -1;
            const int o_hb = (trunk && H > 0) ? o_b + n_b + (ph.identity ? H : 0) : 1 << 20, n_hb = H > 0 ? 2 * H * ph.num_bands : 0;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int j = sub_u * 16 + i;
              int comp = -1, deg = 0;
              float win = 1.f, phase = 0.f;
              bool band = false;
              if (o_id >= 0 && j >= o_id && j < o_id + 3) comp = j - o_id;
              else if (j >= o_b && j < o_b + n_b) {
                const int jj = j - o_b, k = jj / 6, r = jj - 6 * k;
                comp = r >= 3 ? r - 3 : r; deg = deg0 + k; band = true; phase = r >= 3 ? NDS_HALF_PI_F : 0.f;
                if (trunk) win = ps.window[k];            // (the feature block's windows are folded into the weights)
              } else if (o_hid >= 0 && j >= o_hid && j < o_hid + H) comp = 3 + j - o_hid;
              else if (j >= o_hb && j < o_hb + n_hb) {
                const int jj = j - o_hb, k = jj / (2 * H), r = jj - 2 * H * k;
                comp = 3 + (r >= H ? r - H : r); deg = ph.min_deg + k; band = true; phase = r >= H ? NDS_HALF_PI_F : 0.f;
                win = ph.window[k];
              }
              const float xv = comp == 0 ? c0 : (comp == 1 ? c1 : (comp == 2 ? c2 : (comp == 3 ? c3 : c4)));
              const float f = __int_as_float((127 + deg) << 23);
              // d/dx [w sin(x 2^deg + phase)] = w 2^deg cos(x 2^deg + phase); identity columns pass the gradient through
              const float d = band ? win * f * pe_cos(xv * f + phase) : 1.f;
              const float v = comp >= 0 ? d * fbv[i] : 0.f;
              p0 += comp == 0 ? v : 0.f; p1 += comp == 1 ? v : 0.f; p2 += comp == 2 ? v : 0.f;
              p3 += comp == 3 ? v : 0.f; p4 += comp == 4 ? v : 0.f;
            }
          }
          const float part5[5] = {p0, p1, p2, p3, p4};
          // scratch laid out [value][warp of the sample][row]: consecutive lanes hit consecutive banks
          float* scr = reinterpret_cast<float*>(smem + OFF_IN2);
          for (int i = 0; i < 5; ++i) scr[(i * NSUB + sub) * TM + row] = part5[i];
          asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
          float tot[5];
          for (int i = 0; i < 5; ++i) {
            const float* r4 = scr + (size_t)i * NSUB * TM + row;
            tot[i] = (r4[0] + r4[TM]) + (r4[2 * TM] + r4[3 * TM]);
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");      // the block is re-used by the next reduction
          if (sp.arg == 0) { for (int c = 0; c < 3; ++c) GS.gxw[c] = tot[c]; GS.gom[0] = tot[3]; GS.gom[1] = tot[4]; }
          else if (sub == 0 && T.valid) {
            // d sigma_raw / dx = (feature-block path of the warp field and the hyper sheet) + R^T (d / d x')
            float gx[3], gn[3], tn[3], tnn[3];
            for (int c = 0; c < 3; ++c)
              gx[c] = -(tot[c] + (cfg.use_warp ? (T.R[0 * 3 + c] * GS.gxw[0] + T.R[1 * 3 + c] * GS.gxw[1] + T.R[2 * 3 + c] * GS.gxw[2]) : GS.gxw[c]));
            normalize3(gx, gn);                                             // models.py:1070, 1077
            for (int c = 0; c < 3; ++c) tn[c] = cfg.use_warp ? (T.R[c * 3 + 0] * gn[0] + T.R[c * 3 + 1] * gn[1] + T.R[c * 3 + 2] * gn[2]) : gn[c];
            normalize3(tn, tnn);                                            // models.py:1273-1277
            float* PL = a.planes;
            const int64_t ps = a.plane_stride, n = T.n_out;
            for (int c = 0; c < 3; ++c) { __stcs(&PL[(P_GRAD + c) * ps + n], gn[c]); __stcs(&PL[(P_TNORM + c) * ps + n], tnn[c]); }
          }
          }
         }
        } else {
          // ---- STEP_OUT: write planes (the warps sharing a sample take different planes; only the groups of planes the
          //      call's outputs need; streaming stores: nothing in this kernel reads them back, and they must not evict
          //      the weights or the warps' local memory from the L2) ----
          const int64_t n = T.n_out;
          if (T.valid && a.carry_out) {          // what a later "carried" launch needs of this sample
            float* C = a.carry_out + n;
            const int64_t cs = a.carry_stride;
            if (sub == 0) { for (int c = 0; c < 3; ++c) __stcs(&C[(C_WARPED + c) * cs], T.xw[c]); for (int c = 0; c < 2; ++c) __stcs(&C[(C_WARPED + 3 + c) * cs], c < H ? T.om[c] : 0.f); }
            else if (sub == 1) { __stcs(&C[C_MASK * cs], T.pmask); for (int c = 0; c < 3; ++c) __stcs(&C[(C_P + c) * cs], T.p[c]); }
            else if (sub == 2) { for (int i = 0; i < 5; ++i) __stcs(&C[(C_R + i) * cs], T.R[i]); }
            else { for (int i = 5; i < 9; ++i) __stcs(&C[(C_R + i) * cs], T.R[i]); }
          }
          if (T.valid) {
            float* PL = a.planes;
            const int64_t ps = a.plane_stride;
            const uint32_t pm = a.plane_mask;
            if (sub == 0) {
              __stcs(&PL[P_SIGMA_RAW * ps + n], T.sigma_raw);
              if (pm & PG_RGB) for (int c = 0; c < 3; ++c) __stcs(&PL[(P_RGB + c) * ps + n], T.rgb[c]);
            } else if (sub == 1) {
              if (pm & PG_NORM) for (int c = 0; c < 3; ++c) __stcs(&PL[(P_NORM + c) * ps + n], T.nrm[c]);
              if (pm & PG_MASK) __stcs(&PL[P_MASK * ps + n], T.pmask);
            } else if (sub == 2) {
              if (pm & PG_WARPED) {
                for (int c = 0; c < 3; ++c) __stcs(&PL[(P_WARPED + c) * ps + n], T.xw[c]);
                for (int c = 0; c < H; ++c) __stcs(&PL[(P_WARPED + 3 + c) * ps + n], T.om[c]);
              }
            } else if (cfg.use_warp && (pm & (PG_ROT | PG_TRANS))) {
              if (pm & PG_ROT) {
                const float r = 0.57735025882720947265625f;
                float rf[3], rn[3];
                for (int c = 0; c < 3; ++c) rf[c] = T.R[c * 3 + 0] * r + T.R[c * 3 + 1] * r + T.R[c * 3 + 2] * r;
                normalize3(rf, rn);
                for (int c = 0; c < 3; ++c) __stcs(&PL[(P_ROT + c) * ps + n], rn[c]);
              }
              if (pm & PG_TRANS) for (int c = 0; c < 3; ++c) __stcs(&PL[(P_TRANS + c) * ps + n], T.p[c]);
            }
          }
        }
        if (tr) tr[2 * si + 1] = clock64();
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == WARP_MMA) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------
// self-test kernel: one op on caller-provided activations.  The k_hid hidden
// activations are written to tensor-memory columns [0, 256) in the layout an
// upstream epilogue of that width produces, the k_in inputs to IN[0].
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_selftest_kernel(const __grid_constant__ TcProgram P, TcLevel L, const float* __restrict__ A, int k_hid, int k_in,
                   float* out_f32, float* out_readback) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Ctrl* ctl = reinterpret_cast<Ctrl*>(smem + OFF_CTRL);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    uint32_t* c_burst = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(ctl) + ctrl_off_burst(P.n_ops, P.n_steps));
    for (int i = threadIdx.x; i < P.n_burst * 8; i += blockDim.x) c_burst[i] = reinterpret_cast<const uint32_t*>(P.burst)[i];
  }
  if (threadIdx.x == 0) ctrl_init(ctl);
  if (warp == WARP_MMA) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;
  // the burst program holds absolute shared-memory / tensor-memory addresses
  if (smem_base != P.smem_base || tmem_base != 0u) {
    if (threadIdx.x == 0) printf("nerfds_b200: tensor-core program encoded for smem base %u, tmem base 0; got %u, %u\n", P.smem_base, smem_base, tmem_base);
    __trap();
  }
  const TcOp& op = P.ops[0];
  if (warp == WARP_TMA) {
    if (lane == 0) {
      ProducerState ps{0u, 0u};
      produce_program(P, L.weights, smem, ctl, ps);
    }
    __syncwarp();
  } else if (warp == WARP_MMA) {
    if (elect_one_sync()) {
      uint32_t bits = 0;
      const uint32_t s0 = P.src[0], u0 = (s0 >> 29) & 3u;
      if (s0 >> 31) { mbar_wait(&ctl->full[u0], 0u); bits ^= 1u << (8 + u0); }
      issue_program(P, ctl, bits, false, nullptr);
    }
    __syncwarp();
  } else if (is_compute_warp(warp)) {
    const int q = warp & 3, sub = (warp - CWARP0) >> 2;
    const uint32_t row = (uint32_t)q * 32u + (uint32_t)lane;
    const uint32_t tmem_lane = tmem_base + (((uint32_t)q * 32u) << 16);
    const int ld = k_hid + k_in;
    // producer layout of a k_hid-wide layer: chunks of nc_rows columns at 128 c, slices of CW per warp
    const int p_nc = k_hid >= 256 ? 2 : 1, p_rows = k_hid / p_nc, p_cw = p_rows / NSUB;
    for (int c = 0; c < p_nc && k_hid > 0; ++c) {
      const uint32_t col = tmem_lane + 128u * c + (uint32_t)sub * p_cw;
      for (int g = 0; g < p_cw; g += 16) {
        uint32_t hi[8], lo[8];
        for (int i = 0; i < 8; ++i) {
          const int f = c * p_rows + sub * p_cw + g + 2 * i;
          const float x0 = A[row * ld + f], x1 = A[row * ld + f + 1];
          const __half2 h = __floats2half2_rn(x0, x1);
          const float2 hf = __half22float2(h);
          const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
          hi[i] = *reinterpret_cast<const uint32_t*>(&h);
          lo[i] = *reinterpret_cast<const uint32_t*>(&l);
        }
        tmem_st<8>(col + g / 2, hi);
        tmem_st<8>(col + p_cw / 2 + g / 2, lo);
      }
    }
    for (int c = 0; c < 64; ++c)
      if ((c & (NSUB - 1)) == sub) store_in(smem + OFF_IN, row, c, c < k_in ? A[row * ld + k_hid + c] : 0.f);
    warp_arrive(&ctl->part[0][0], lane);
    warp_arrive(&ctl->part[0][1], lane);
    warp_arrive(&ctl->glue[0], lane);
    if (op.epi_kind == EPI_HEAD) {
      mbar_wait(&ctl->d_full[0][0], 0);
      tc_fence_after_sync();
      float hv[16], hb[16];
      head_bias(op, L.bias, hb);
      epilogue_head(op, hb, tmem_lane, hv);
      if (sub == 0) for (int i = 0; i < 16 && i < op.N; ++i) out_f32[row * op.N + i] = hv[i];
    } else {
      for (int nc = 0; nc < op.n_nc; ++nc)
        epilogue_dispatch(op, nc, L.bias, tmem_lane, row, sub, q, out_f32, op.N, &ctl->d_full[0][nc], 0);
      tmem_st_wait();
      tc_fence_before_sync();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  // read back the operand image the epilogue left in tensor memory -- what the next layer's MMA would see
  if (is_compute_warp(warp) && op.epi_kind != EPI_HEAD && out_readback) {
    const int q = warp & 3, sub = (warp - CWARP0) >> 2;
    const uint32_t row = (uint32_t)q * 32u + (uint32_t)lane;
    const uint32_t tmem_lane = tmem_base + (((uint32_t)q * 32u) << 16);
    const int cw = op.nc_rows / NSUB;
    for (int nc = 0; nc < op.n_nc; ++nc) {
      const uint32_t chunk = tmem_lane + op.d_col[nc];
      const uint32_t hcol = op.epi_kind == EPI_COMPACT_HI ? chunk + sub * (cw / 2) : chunk + sub * cw;
      for (int g = 0; g < cw / 2; g += 8) {
        uint32_t h[8], l[8];
        tmem_ld8(hcol + g, h);
        tmem_ld8(hcol + cw / 2 + g, l);        // only meaningful for EPI_INPLACE
        tmem_ld_wait();
        for (int i = 0; i < 8; ++i) {
          const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h[i]));
          const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&l[i]));
          const bool wl = op.epi_kind == EPI_INPLACE;
          const int oc = nc * op.nc_rows + sub * cw + 2 * (g + i);
          out_readback[row * op.N + oc] = hf.x + (wl ? lf.x : 0.f);
          out_readback[row * op.N + oc + 1] = hf.y + (wl ? lf.y : 0.f);
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == WARP_MMA) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------
// host: packing
// ---------------------------------------------------------------------------
// One K-chunk (64 operand columns) of an op: where the operand lives and which weight rows multiply it.
struct KChunkMap {
  uint8_t pat = PAT_SS;    // APattern
  uint8_t wait = 0;        // bit c: needs output chunk c of the previous op
  uint8_t wait_glue = 0;   // needs the per-sample stage (rgb side inputs)
  uint8_t terms = 0;       // 0: the op's; 1: this K-chunk is 1-term even inside a 3-term op (hi-only operand block)
  uint32_t a_hi = 0, a_lo = 0;   // PAT_SS: shared byte offsets; else tensor-memory columns
  int rows[64];            // W row feeding each of the 64 operand columns (-1 = zero pad)
  int band[64];            // >= 0: the column is a posenc feature of that band; its weights carry the window
  int pe = -1;             // which window: 0 mask, 1 warp, 2 hyper sheet
};

struct OpBuild {
  std::vector<KChunkMap> kcs;
  int N_logical = 0;                  // real output columns
  int N = 0;                          // padded
  int n_nc = 1;                       // N-chunks (accumulators); nc_rows = N / n_nc
  int d_col[2] = {0, 0};              // tensor-memory column of accumulator chunk c
  int terms = 3, relu = 0, epi_kind = EPI_INPLACE, glue = GLUE_NONE;
  int prev_produces = 0;              // the op before this one is a hidden layer (its epilogue signals part)
  int interleave = 0;                 // order the bursts K-range-major so that chunk 1's epilogue stays hidden
  int cross_first = 0;                // 3-term op whose weight images are all resident (N phase): issue the small
                                      // terms of every K-chunk before any main term (see issue_pair)
  int mask_idx = 0, mask_mode = MASK_NONE;   // ReLU mask words the epilogue records (forward) / applies (reverse sweep)
  int row_band[64];                   // reverse sweep, input-gradient ops of the narrow networks: output row n is feature-
  int row_pe = -1;                    // block column n; >= 0: its weights carry posenc window row_band[n] of encoder row_pe
  OpBuild() { for (int i = 0; i < 64; ++i) row_band[i] = -1; }
  uint16_t first_flags = 0;           // B_WAIT_* of the first burst
  uint16_t first_part_waits = 0;      // bit c: the first burst waits for part c of the previous op
  int signal_glue = 0;
  int tslot = 0;
  const std::vector<float>* W = nullptr;   // [K_total][N_logical] logical weights (row-major, Flax layout)
  const std::vector<float>* b = nullptr;
  std::vector<float> W_own, b_own;    // fused heads own their matrices
};

// weight images of one input K-chunk whose columns carry posenc windows: re-packed when the windows change
struct WindowedImage {
  size_t stream_off;                  // byte offset of the B_hi image (B_lo follows when terms == 3)
  int rows, terms, pe;
  std::vector<float> w;               // [rows][64] scaled weights without the window
  int band[64];                       // per K column
  int row_band[128];                  // per output row (input-gradient ops), -1 = none
};

struct Packed {
  std::vector<uint8_t> stream;
  std::vector<float> bias;
  std::vector<WindowedImage> windowed;
};

// Layout an in-place epilogue leaves behind (see epilogue_chunk): N output features in n_nc chunks at d_col[c];
// inside a chunk, warp slice s holds features [s CW, (s+1) CW) as hi (CW/2 columns) | lo (CW/2).
struct ActLayout {
  int N = 0, n_nc = 1, compact_hi = 0;   // compact_hi: EPI_COMPACT_HI (hi only, contiguous from the chunk start)
  int d_col[2] = {0, 0};
  int region = 0;                        // which of the two ping-pong regions holds it
  int nc_rows() const { return N / n_nc; }
  int cw() const { return nc_rows() / NSUB; }
};

// operand K-block j (features 64 j ...) of a layer output with layout L
static KChunkMap kc_hidden(const ActLayout& L, int j, int row0, bool wait) {
  KChunkMap k;
  const int f0 = 64 * j, c = f0 / L.nc_rows(), g = f0 % L.nc_rows();
  if (L.compact_hi) { k.pat = PAT_8; k.a_hi = L.d_col[c] + g / 2; k.a_lo = k.a_hi; }
  else {
    k.pat = L.cw() == 32 ? PAT_32 : PAT_16;
    k.a_hi = L.d_col[c] + g;            // g is a multiple of 64 = whole warp slices
    k.a_lo = k.a_hi + L.cw() / 2;
  }
  k.wait = wait ? (uint8_t)(1u << c) : 0;
  for (int cidx = 0; cidx < 64; ++cidx) { k.rows[cidx] = (f0 + cidx < L.N) ? row0 + f0 + cidx : -1; k.band[cidx] = -1; }
  return k;
}
// inputs in a shared-memory block (hi at `off`, lo one K-block later), identity column map
static KChunkMap kc_input(uint32_t off, int row0, int in_dim) {
  KChunkMap k;
  k.pat = PAT_SS; k.a_hi = off; k.a_lo = off + KBLK;
  for (int c = 0; c < 64; ++c) { k.rows[c] = c < in_dim ? row0 + c : -1; k.band[c] = -1; }
  return k;
}

// shared feature block layout (TcProgram::f_*)
struct FLayout {
  int col_ident = -1, col_bands = 0, kmin = 0, nb = 0, col_wembed = -1, col_membed = -1, col_mask = -1, cols = 0;
};
// inputs of narrow network `pe` (0 mask, 1 warp, 2 hyper sheet) gathered from the shared feature block.
// Reference input order: [identity x][band k: sin(3) cos(3)]...[embedding][mask]  (modules.py:394-434,
// warping.py:200-215, modules.py:351-392)
static KChunkMap kc_features(const FLayout& F, uint32_t off, int row0, int pe, int min_deg, int max_deg, int identity,
                             int embed_col, int embed_dims, bool with_mask) {
  KChunkMap k;
  k.pat = PAT_SS; k.a_hi = off; k.a_lo = off + KBLK; k.pe = pe;
  for (int c = 0; c < 64; ++c) { k.rows[c] = -1; k.band[c] = -1; }
  int j = row0;
  if (identity) for (int c = 0; c < 3; ++c) k.rows[F.col_ident + c] = j++;
  for (int d = min_deg; d < max_deg; ++d)
    for (int r = 0; r < 6; ++r) {
      const int col = F.col_bands + 6 * (d - F.kmin) + r;
      k.rows[col] = j++;
      k.band[col] = d - min_deg;
    }
  for (int e = 0; e < embed_dims; ++e) k.rows[embed_col + e] = j++;
  if (with_mask) k.rows[F.col_mask] = j++;
  return k;
}

static float op_scale(const OpBuild& ob, int& e_out) {
  // power-of-two scale so that max |W| lands in [4, 8): keeps W_lo out of fp16 subnormals
  float mx = 0.f;
  for (float v : *ob.W) mx = std::max(mx, std::fabs(v));
  int e = 0;
  if (mx > 0.f) { std::frexp(mx, &e); e = 3 - e; }
  if (e > 14) e = 14;
  if (e < -14) e = -14;
  e_out = e;
  return std::ldexp(1.f, e);
}

static void write_image(uint8_t* dst, const float* w /*[rows][64]*/, const float* colscale, int rows, int terms,
                        const float* rowscale = nullptr) {
  const size_t img = (size_t)rows * 128;
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < 64; ++c) {
      const float v = w[(size_t)r * 64 + c] * (colscale ? colscale[c] : 1.f) * (rowscale ? rowscale[r] : 1.f);
      const __half hi = __float2half_rn(v);
      const __half lo = __float2half_rn(v - __half2float(hi));
      const uint32_t o = kblock_offset((uint32_t)r, (uint32_t)c);
      memcpy(dst + o, &hi, 2);
      if (terms == 3) memcpy(dst + img + o, &lo, 2);
    }
}

// Weight images of one op (shared by both tile slots): src[nc][kc] = stream offset in 128-byte rows.
struct OpWeights {
  uint32_t bias_off = 0;
  float inv_scale = 1.f;
  std::vector<uint32_t> src;          // [nc * n_kc + kc]
};
static OpWeights pack_weights(const OpBuild& ob, Packed& out) {
  OpWeights ow;
  int e = 0;
  const float scale = op_scale(ob, e);
  ow.inv_scale = std::ldexp(1.f, -e);
  ow.bias_off = (uint32_t)out.bias.size();
  for (int n = 0; n < ob.N; ++n) out.bias.push_back(n < ob.N_logical ? (*ob.b)[n] : 0.f);
  while (out.bias.size() % 4) out.bias.push_back(0.f);
  const int nc_rows = ob.N / ob.n_nc;
  const size_t img = (size_t)nc_rows * 128;
  std::vector<float> w((size_t)nc_rows * 64);
  for (int nc = 0; nc < ob.n_nc; ++nc)
    for (size_t kc = 0; kc < ob.kcs.size(); ++kc) {
      const KChunkMap& km = ob.kcs[kc];
      bool windowed = false;
      for (int r = 0; r < nc_rows; ++r) {
        const int n = nc * nc_rows + r;
        for (int c = 0; c < 64; ++c) {
          float v = 0.f;
          if (n < ob.N_logical && km.rows[c] >= 0) v = (*ob.W)[(size_t)km.rows[c] * ob.N_logical + n] * scale;
          w[(size_t)r * 64 + c] = v;
          if (km.band[c] >= 0) windowed = true;
        }
      }
      const int kterms = km.terms ? km.terms : ob.terms;
      const size_t base = out.stream.size();
      out.stream.resize(base + img * (kterms == 3 ? 2 : 1), 0);
      write_image(&out.stream[base], w.data(), nullptr, nc_rows, kterms);
      ow.src.push_back((uint32_t)(base / 128));
      if (ob.row_pe >= 0) {
        for (int r = 0; r < nc_rows && r < 64; ++r) if (ob.row_band[r] >= 0) windowed = true;
      }
      if (windowed) {
        WindowedImage wi;
        wi.stream_off = base; wi.rows = nc_rows; wi.terms = kterms; wi.pe = ob.row_pe >= 0 ? ob.row_pe : km.pe; wi.w = w;
        memcpy(wi.band, km.band, sizeof wi.band);
        for (int r = 0; r < 128; ++r) wi.row_band[r] = (ob.row_pe >= 0 && r < 64) ? ob.row_band[r] : -1;
        out.windowed.push_back(std::move(wi));
      }
    }
  return ow;
}

static TcOp make_tcop(const OpBuild& ob, const OpWeights& ow) {
  TcOp op;
  memset(&op, 0, sizeof op);
  op.bias_off = ow.bias_off;
  op.inv_scale = ow.inv_scale;
  op.N = (uint16_t)ob.N;
  op.nc_rows = (uint16_t)(ob.N / ob.n_nc);
  op.n_nc = (uint8_t)ob.n_nc;
  op.relu = (uint8_t)ob.relu;
  op.epi_kind = (uint8_t)ob.epi_kind;
  op.glue = (uint8_t)ob.glue;
  op.d_col[0] = (uint16_t)ob.d_col[0];
  op.d_col[1] = (uint16_t)ob.d_col[1];
  op.signal_glue = (uint8_t)ob.signal_glue;
  op.mask_idx = (uint8_t)ob.mask_idx;
  op.mask_mode = (uint8_t)ob.mask_mode;
  return op;
}

// bursts of one op of one tile slot, in issue order (ring flags / units are assigned by the assembler)
static std::vector<BurstH> make_bursts(const OpBuild& ob, const OpWeights& ow) {
  const int nc_rows = ob.N / ob.n_nc;
  const int n_kc = (int)ob.kcs.size();
  auto steps_of = [&](const KChunkMap& km) {
    int last = -1;
    for (int c = 0; c < 64; ++c) if (km.rows[c] >= 0) last = c;
    return (uint8_t)(km.pat == PAT_SS ? std::max(1, (last + 16) / 16) : 4);
  };
  std::vector<BurstH> out;
  int waited = 0;    // bit 0 / 1: part c, bit 2: glue (of the per-sample stage feeding a K-chunk)
  auto wait_flags = [&](int need, bool first) {
    uint16_t fl = 0;
    if (first) { fl |= ob.first_flags; need |= ob.first_part_waits; if (ob.first_flags & B_WAIT_GLUE) waited |= 4; }
    need &= ~waited;
    if (need & 1) fl |= B_WAIT_P0;
    if (need & 2) fl |= B_WAIT_P1;
    if (need & 4) fl |= B_WAIT_GLUE;
    waited |= need;
    return fl;
  };
  bool cross_first = ob.cross_first && ob.n_nc == 1 && ob.terms == 3;
  for (const KChunkMap& km : ob.kcs) if (km.terms && km.terms != 3) cross_first = false;
  if (cross_first) {
    // X(kc) for every K-chunk: A_lo B_hi + A_hi B_lo; then the main terms A_hi B_hi, two K-chunks per burst where
    // the operand patterns agree
    struct Item { int kc, kc2; bool cross; };
    std::vector<Item> items;
    for (int kc = 0; kc < n_kc; ++kc) items.push_back({kc, -1, true});
    for (int kc = 0; kc < n_kc;) {
      const bool pairable = kc + 1 < n_kc && ob.kcs[kc].pat != PAT_SS && ob.kcs[kc + 1].pat == ob.kcs[kc].pat;
      items.push_back({kc, pairable ? kc + 1 : -1, false});
      kc += pairable ? 2 : 1;
    }
    for (size_t i = 0; i < items.size(); ++i) {
      const Item& it = items[i];
      const KChunkMap& km = ob.kcs[it.kc];
      BurstH e;
      e.pat = km.pat; e.rows = (uint16_t)nc_rows; e.steps = steps_of(km); e.d_col = (uint16_t)ob.d_col[0];
      e.tslot = (uint8_t)ob.tslot; e.kc = (int16_t)it.kc; e.kc2 = (int16_t)it.kc2;
      e.src = ow.src[(size_t)it.kc]; e.rows128 = (uint16_t)(nc_rows * 2);
      int need = km.wait | (km.wait_glue ? 4 : 0);
      uint16_t fl = 0;
      if (it.cross) { e.pair = 1; e.a_hi = km.a_lo; e.a_lo = km.a_hi; e.b_sel[0] = 0; e.b_sel[1] = 1; }
      else if (it.kc2 >= 0) {
        const KChunkMap& k2 = ob.kcs[it.kc2];
        e.pair = 1; e.a_hi = km.a_hi; e.a_lo = k2.a_hi; e.b_sel[0] = 0; e.b_sel[1] = 0;
        need |= k2.wait | (k2.wait_glue ? 4 : 0);
      } else { e.a_hi = km.a_hi; e.a_lo = km.a_lo; }     // plain 1-term burst: A_hi B_hi
      if (i == 0) fl |= B_FIRST;
      if (i + 1 == items.size()) { fl |= B_LAST; if (ob.prev_produces) fl |= B_PART_NEXT; }
      fl |= wait_flags(need, i == 0);
      e.flags = fl;
      out.push_back(e);
    }
    return out;
  }
  std::vector<std::pair<int, int>> order;
  if (ob.n_nc == 2 && ob.interleave) {
    for (int late = 0; late < 2; ++late)
      for (int nc = 0; nc < 2; ++nc)
        for (int kc = 0; kc < n_kc; ++kc)
          if (((ob.kcs[kc].wait & 2) != 0 || ob.kcs[kc].wait_glue) == (late != 0)) order.push_back({nc, kc});
  } else {
    for (int nc = 0; nc < ob.n_nc; ++nc) for (int kc = 0; kc < n_kc; ++kc) order.push_back({nc, kc});
  }
  int seen[2] = {0, 0};
  int left[2] = {n_kc, n_kc};
  for (size_t i = 0; i < order.size(); ++i) {
    const int nc = order[i].first, kc = order[i].second;
    const KChunkMap& km = ob.kcs[kc];
    BurstH e;
    e.a_hi = km.a_hi; e.a_lo = km.a_lo; e.pat = km.pat;
    e.rows = (uint16_t)nc_rows;
    e.steps = steps_of(km);
    e.d_col = (uint16_t)ob.d_col[nc];
    e.tslot = (uint8_t)ob.tslot;
    e.kc = (int16_t)(nc * n_kc + kc);                    // one weight image pair per (N-chunk, K-chunk)
    e.src = ow.src[(size_t)nc * n_kc + kc];
    const int kterms = km.terms ? km.terms : ob.terms;
    e.rows128 = (uint16_t)(nc_rows * (kterms == 3 ? 2 : 1));
    uint16_t fl = 0;
    if (kterms == 3) fl |= B_TWO;
    if (!seen[nc]) { fl |= B_FIRST; seen[nc] = 1; }
    if (--left[nc] == 0) fl |= B_LAST;
    if (nc == 1) fl |= B_NC1;
    fl |= wait_flags(km.wait | (km.wait_glue ? 4 : 0), i == 0);
    if (i + 1 == order.size() && ob.prev_produces) fl |= B_PART_NEXT;
    e.flags = fl;
    out.push_back(e);
  }
  return out;
}

// Ring units of the weight images a list of bursts reads: one unit per distinct K-chunk id, taken round-robin
// from `cursor`.  `acquire` / `release`: flag the first / last burst that touches a unit (the N phase lets tile
// slot 0 acquire and tile slot 1 release the units both slots read).
static void assign_units(std::vector<BurstH>& b, int& cursor, int* unit_of /*[64], -1 = unassigned*/, bool acquire, bool release) {
  for (auto& e : b) {
    for (int16_t kc : {e.kc, e.kc2}) {
      if (kc < 0) continue;
      if (unit_of[kc] < 0) { unit_of[kc] = cursor; cursor = (cursor + 1) % NUNIT; }
    }
    e.unit = (uint8_t)unit_of[e.kc];
    e.unit2 = e.kc2 >= 0 ? (uint8_t)unit_of[e.kc2] : (uint8_t)0xff;
  }
  if (acquire) {
    bool seen[64] = {false};
    for (auto& e : b) if (!seen[e.kc]) { seen[e.kc] = true; e.flags |= B_ACQUIRE; }      // X bursts come first: kc2 never opens a unit
  }
  if (release) {
    bool seen[64] = {false};
    for (auto it = b.rbegin(); it != b.rend(); ++it) {
      if (it->kc2 >= 0 && !seen[it->kc2]) { seen[it->kc2] = true; it->release2 = 1; }
      if (!seen[it->kc]) { seen[it->kc] = true; it->flags |= B_RELEASE; }
    }
  }
}

// shared-window address of the 1024-aligned dynamic shared memory of the engine's kernels (none of them has
// static shared memory, so it is the same for all): measured once with a probe launch
__global__ void smem_base_probe_kernel(uint32_t* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  if (threadIdx.x == 0) *out = smem_u32(smem);
}
static bool query_smem_base(uint32_t& base, std::string& err) {
  static uint32_t cached = 0;
  static bool have = false;
  if (!have) {
    uint32_t* d = nullptr;
    cudaError_t e = cudaMalloc(&d, 4);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(smem_base_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_PROBE);
    if (e == cudaSuccess) { smem_base_probe_kernel<<<1, 32, TC_SMEM_PROBE>>>(d); e = cudaDeviceSynchronize(); }
    if (e == cudaSuccess) e = cudaMemcpy(&cached, d, 4, cudaMemcpyDeviceToHost);
    if (d) cudaFree(d);
    if (e != cudaSuccess) { err = std::string("tensor-core engine: shared-memory base probe: ") + cudaGetErrorString(e); return false; }
    have = true;
  }
  base = cached;
  return true;
}

// Ctrl barrier indices (Burst::bars)
#define NDS_BAR_IDX(member) ((uint32_t)(offsetof(Ctrl, member) / 8))

// device form of the bursts of a program; `next` links implement the issuer's weight-barrier probe
static void encode_bursts(const std::vector<BurstH>& hb, bool cyclic, uint32_t smem_base, Burst* out, uint32_t* src) {
  const int n = (int)hb.size();
  auto desc_lo = [&](uint32_t off) { return (((smem_base + off) & 0x3FFFFu) >> 4) | (1u << 16); };
  for (int i = 0; i < n; ++i) {
    const BurstH& e = hb[i];
    Burst b;
    const bool ss = e.pat == PAT_SS;
    b.a_hi = ss ? desc_lo(e.a_hi) : e.a_hi;
    b.a_lo = ss ? desc_lo(e.a_lo) : e.a_lo;
    b.d = e.d_col;
    b.idesc = make_idesc_f16(e.rows);
    b.b0 = desc_lo(OFF_RING + (uint32_t)e.unit * UNIT_BYTES);
    b.b1 = b.b0 + (uint32_t)e.rows * 8u;
    if (e.pair) {        // group 1 reads image b_sel[0] of `unit`, group 2 image b_sel[1] of `unit2` (or `unit`)
      const uint32_t u2 = desc_lo(OFF_RING + (uint32_t)(e.unit2 != 0xff ? e.unit2 : e.unit) * UNIT_BYTES);
      b.b0 += e.b_sel[0] ? (uint32_t)e.rows * 8u : 0u;
      b.b1 = u2 + (e.b_sel[1] ? (uint32_t)e.rows * 8u : 0u);
    }
    uint32_t nu = 0, wrap = 0;
    if (i + 1 < n) { if (hb[i + 1].flags & B_ACQUIRE) nu = 1u + hb[i + 1].unit; }
    else if (cyclic && (hb[0].flags & B_ACQUIRE)) { nu = 1u + hb[0].unit; wrap = 1; }
    b.ctl = (uint32_t)(e.flags & 0x1fffu) | ((uint32_t)e.pat << 13) | ((uint32_t)e.steps << 15) | ((uint32_t)e.tslot << 18) |
            ((uint32_t)e.unit << 19) | (nu << 21) | (wrap << 24) | (e.pair ? CTL_PAIR : 0u) | e.ctl_extra;
    const uint32_t eb = (e.flags & B_RELEASE) ? NDS_BAR_IDX(empty) + e.unit : 0xffu;
    const uint32_t db = (e.flags & B_LAST) ? NDS_BAR_IDX(d_full) + 2u * e.tslot + ((e.flags & B_NC1) ? 1u : 0u) : 0xffu;
    const uint32_t pb = nu ? NDS_BAR_IDX(full) + (nu - 1u) : 0xffu;
    const uint32_t eb2 = e.release2 ? NDS_BAR_IDX(empty) + e.unit2 : 0xffu;
    b.bars = eb | (db << 8) | (pb << 16) | (eb2 << 24);
    out[i] = b;
    src[i] = (e.src & 0xfffffu) | ((uint32_t)e.rows128 << 20) | ((uint32_t)e.unit << 29) | ((e.flags & B_ACQUIRE) ? (1u << 31) : 0u);
  }
}

// hidden stack of a modules.MLP: layer l reads [h (width) | inputs (in_dim) at the skip layer].
//   narrow (N phase): regions of 128 columns at 256 tslot + {0, 128}; layer l accumulates into region (l even ? 1 : 0)
//   wide   (T phase): regions of 256 columns at {0, 256}, two N-chunks; layer l accumulates into region (l even ? 0 : 1)
// Returns the last layer's layout.
static ActLayout build_mlp_ops(const HostMlp& m, int terms, bool wide, int tslot, const KChunkMap& in0,
                               const KChunkMap& in_skip, uint16_t first_flags, std::vector<OpBuild>& ops,
                               bool cross_first = false, int mask_base = -1) {
  ActLayout prev;
  for (int l = 0; l < m.depth; ++l) {
    OpBuild ob;
    ob.N_logical = ob.N = m.width;
    ob.tslot = tslot;
    int region;
    if (wide) {
      region = (l % 2 == 0) ? 0 : 1;
      ob.n_nc = 2;
      ob.d_col[0] = region * 256; ob.d_col[1] = ob.d_col[0] + 128;
      ob.interleave = 1;
    } else {
      region = (l % 2 == 0) ? 1 : 0;
      ob.n_nc = 1;
      ob.d_col[0] = ob.d_col[1] = 256 * tslot + 128 * region;
      ob.cross_first = cross_first ? 1 : 0;
    }
    ob.terms = terms; ob.relu = 1; ob.glue = GLUE_NONE; ob.epi_kind = EPI_INPLACE;
    if (mask_base >= 0) { ob.mask_mode = MASK_RECORD; ob.mask_idx = mask_base + l * ob.n_nc; }
    ob.prev_produces = l > 0;
    if (l == 0) ob.first_flags = first_flags;
    // the second trunk layer of tile slot 0 accumulates over the columns of tile slot 1's narrow networks
    if (wide && l == 1 && tslot == 0) ob.first_flags |= B_PEEK_GLUE_OTHER;
    ob.W = &m.hidden[l].W; ob.b = &m.hidden[l].b;
    if (l == 0) ob.kcs.push_back(in0);
    else {
      if (l == m.skip) ob.kcs.push_back(in_skip);          // ready long ago: issue it first
      for (int j = 0; j < m.width / 64; ++j) ob.kcs.push_back(kc_hidden(prev, j, 0, true));
    }
    ops.push_back(ob);
    prev = ActLayout();
    prev.N = m.width; prev.n_nc = ob.n_nc; prev.d_col[0] = ob.d_col[0]; prev.d_col[1] = ob.d_col[1]; prev.region = region;
  }
  return prev;
}

// head over the layer output `in`; accumulators at the start of the other region
static void build_head_op(const std::vector<const HostDense*>& heads, const ActLayout& in, int terms, int glue,
                          bool wide, int tslot, int signal_glue, std::vector<OpBuild>& ops, bool cross_first = false) {
  OpBuild ob;
  int n = 0;
  for (auto* h : heads) n += h->N;
  ob.N_logical = n;
  ob.N = 16;
  ob.n_nc = 1;
  ob.tslot = tslot;
  ob.d_col[0] = ob.d_col[1] = wide ? (1 - in.region) * 256 : 256 * tslot + 128 * (1 - in.region);
  ob.terms = terms; ob.relu = 0; ob.epi_kind = EPI_HEAD; ob.glue = glue; ob.prev_produces = 1;
  ob.signal_glue = signal_glue;
  ob.cross_first = (!wide && cross_first) ? 1 : 0;
  ob.W_own.assign((size_t)in.N * n, 0.f);
  int c0 = 0;
  for (auto* h : heads) {
    for (int k = 0; k < in.N; ++k) for (int j = 0; j < h->N; ++j) ob.W_own[(size_t)k * n + c0 + j] = h->W[(size_t)k * h->N + j];
    for (int j = 0; j < h->N; ++j) ob.b_own.push_back(h->b[j]);
    c0 += h->N;
  }
  for (int j = 0; j < in.N / 64; ++j) ob.kcs.push_back(kc_hidden(in, j, 0, true));
  ops.push_back(ob);
}
// Reverse sweep of one modules.MLP (App. E): with zbar_l = (d sigma / d h_l) masked by layer l's ReLU, appends
//   bwd(L-1) ... bwd(skip+1), INGRAD(skip), bwd(skip), ..., bwd(1), INGRAD(0)
// where bwd(l): zbar_{l-1} = (zbar_l W_l[h rows]^T) masked by layer l-1, and INGRAD(l): 64 columns of
// zbar_l W_l[input rows]^T in the layout of the network's input block (`in0` / `in_skip`: column -> W row, windows).
// The seed zbar_{L-1} is written by the compute warps (STEP_SEED) into region 0; regions alternate from there.
static void build_mlp_backward(const HostMlp& m, int terms, bool wide, int tslot, const KChunkMap& in0,
                               const KChunkMap& in_skip, int mask_base, bool cross_first, std::vector<OpBuild>& ops) {
  const int L = m.depth, W = m.width, n_nc = wide ? 2 : 1;
  auto layout = [&](int region) {
    ActLayout a;
    a.N = W; a.n_nc = n_nc; a.region = region;
    if (wide) { a.d_col[0] = region * 256; a.d_col[1] = a.d_col[0] + 128; }
    else a.d_col[0] = a.d_col[1] = 256 * tslot + 128 * region;
    return a;
  };
  int region = 0;                      // where zbar_l lives
  auto ingrad = [&](int l, const KChunkMap& map) {
    OpBuild ob;
    ob.N_logical = ob.N = 64; ob.n_nc = 1; ob.tslot = tslot;
    const ActLayout out = layout(1 - region);
    ob.d_col[0] = ob.d_col[1] = out.d_col[0];
    ob.terms = terms; ob.relu = 0; ob.epi_kind = EPI_INGRAD; ob.glue = GLUE_NONE; ob.prev_produces = 1;
    ob.signal_glue = l != 0 ? 1 : 0;   // a hidden op overwrites these accumulators next
    ob.cross_first = (!wide && cross_first) ? 1 : 0;
    const std::vector<float>& Wl = m.hidden[l].W;      // [K_l][W]
    ob.W_own.assign((size_t)W * 64, 0.f);
    ob.b_own.assign(64, 0.f);
    for (int n = 0; n < 64; ++n) {
      if (map.rows[n] < 0) continue;
      for (int k = 0; k < W; ++k) ob.W_own[(size_t)k * 64 + n] = Wl[(size_t)map.rows[n] * W + k];
      ob.row_band[n] = map.band[n];
    }
    ob.row_pe = map.pe;
    const ActLayout in = layout(region);
    for (int j = 0; j < W / 64; ++j) ob.kcs.push_back(kc_hidden(in, j, 0, true));
    ops.push_back(ob);
  };
  for (int l = L - 1; l >= 1; --l) {
    const bool after_ingrad = l == m.skip;
    if (after_ingrad) ingrad(l, in_skip);
    OpBuild ob;
    ob.N_logical = ob.N = W; ob.n_nc = n_nc; ob.tslot = tslot; ob.interleave = wide ? 1 : 0;
    const ActLayout out = layout(1 - region), in = layout(region);
    ob.d_col[0] = out.d_col[0]; ob.d_col[1] = out.d_col[1];
    ob.terms = terms; ob.relu = 0; ob.epi_kind = EPI_INPLACE; ob.glue = GLUE_NONE;
    ob.mask_mode = MASK_APPLY; ob.mask_idx = mask_base + (l - 1) * n_nc;
    ob.cross_first = (!wide && cross_first) ? 1 : 0;
    ob.prev_produces = after_ingrad ? 0 : 1;           // the INGRAD op before it consumed the phase
    if (after_ingrad) ob.first_flags = B_WAIT_GLUE;    // ... and its accumulators must have been read
    const std::vector<float>& Wl = m.hidden[l].W;      // [W (+ in_dim)][W], h rows first
    ob.W_own.assign((size_t)W * W, 0.f);
    ob.b_own.assign(W, 0.f);
    for (int k = 0; k < W; ++k) for (int n = 0; n < W; ++n) ob.W_own[(size_t)k * W + n] = Wl[(size_t)n * W + k];
    for (int j = 0; j < W / 64; ++j) ob.kcs.push_back(kc_hidden(in, j, 0, !after_ingrad));
    ops.push_back(ob);
    region = 1 - region;
  }
  ingrad(0, in0);
}

static void fix_own(std::vector<OpBuild>& ops) {
  for (auto& ob : ops) if (!ob.W) { ob.W = &ob.W_own; ob.b = &ob.b_own; }
}

struct LevelBuild {
  std::vector<OpBuild> ops[2];        // per tile slot, same length
  std::vector<OpWeights> weights;     // per op (shared by the tile slots)
  int n_narrow = 0, n_sigma = 0;      // ops of the N phase; ops up to and including the sigma/normal head
  int trunk_first = 0, trunk_skip_op = -1;
  FLayout F;
  int t_cols = 0;
  // reverse sweep (appended after the forward ops): [tb_first, tb_first + tb_count) trunk, then the hyper sheet
  // [hb_first, + hb_count) and the SE(3) field [wb_first, + wb_count)
  int n_fwd = 0, tb_first = 0, tb_count = 0, hb_first = 0, hb_count = 0, wb_first = 0, wb_count = 0;
};

struct TcEngine {
  Packed packed[2];
  LevelBuild lb[2];
  TcProgram prog[2][4];               // [level][0 sigma-only | 1 full | 2 full, carried (no N phase) | 3 full + reverse sweep]
  uint8_t* d_stream[2] = {nullptr, nullptr};
  float* d_bias[2] = {nullptr, nullptr};
  float win[2][3][NDSR_MAX_BANDS];    // windows currently folded into the streams, per level
  bool win_valid[2] = {false, false};
};

// Merges the per-slot op lists into the pair program: bursts (issuer / producer) and steps (compute warps).
// `carried`: no N phase -- the narrow networks' results come from the carry planes, the trunk input is written by
// the PREP steps and the first trunk layer waits for them.
static bool assemble(const LevelBuild& LB, bool full, bool carried, bool grad, uint32_t smem_base, TcProgram& prog, std::string& err) {
  memset(&prog, 0, sizeof prog);
  prog.smem_base = smem_base;
  prog.carried = carried ? 1 : 0;
  prog.grad = grad ? 1 : 0;
  if (grad && (!full || carried)) { err = "tensor-core engine: the reverse sweep is built for the full program"; return false; }
  const int n_fwd = full ? LB.n_fwd : LB.n_sigma;
  const int n_ops = grad ? (int)LB.ops[0].size() : n_fwd;       // per tile slot
  if (2 * n_ops > MAX_OPS) { err = "tensor-core engine: too many layers"; return false; }
  prog.n_ops = 2 * n_ops;
  prog.full = full ? 1 : 0;
  prog.f_col_ident = LB.F.col_ident; prog.f_col_bands = LB.F.col_bands; prog.f_kmin = LB.F.kmin; prog.f_nb = LB.F.nb;
  prog.f_col_wembed = LB.F.col_wembed; prog.f_col_membed = LB.F.col_membed; prog.f_col_mask = LB.F.col_mask;
  prog.f_cols = LB.F.cols; prog.t_cols = LB.t_cols;
  std::vector<BurstH> bursts;
  std::vector<Step> steps;
  auto op_index = [&](int s, int i) { return s * n_ops + i; };
  for (int s = 0; s < 2; ++s)
    for (int i = 0; i < n_ops; ++i) {
      OpBuild ob = LB.ops[s][i];
      if (!full && i == n_fwd - 1) ob.signal_glue = 0;      // nothing follows the sigma head
      if (!grad) ob.mask_mode = MASK_NONE;
      prog.ops[op_index(s, i)] = make_tcop(ob, LB.weights[i]);
      // the tile slot's tensor-memory columns are free again: after the T phase (slot 1 may start its own; with a
      // reverse sweep only slot 0 signals here, see the N-phase sweep below) and after the whole program of the slot
      const bool t_end = grad ? (s == 0 && i == LB.tb_first + LB.tb_count - 1) : (i == n_fwd - 1);
      prog.ops[op_index(s, i)].signal_done = (t_end || (grad && i == n_ops - 1)) ? 1 : 0;
    }
  // Slot 1's last per-sample stage of the N phase (SE(3) exponential + trunk input, ~7 k cycles on the compute warps)
  // runs AFTER slot 0's first trunk epilogue instead of before it, and releases its tensor-memory columns (done[1])
  // as soon as it has read its head: slot 0's second trunk layer -- which accumulates over those columns -- is then
  // issued while that stage computes, instead of after it.
  const bool early_l1 = !carried && !grad && LB.n_narrow > 0 && !getenv("NDS_TC_NO_EARLY_L1");
  if (early_l1) prog.ops[op_index(1, LB.n_narrow - 1)].signal_done = 1;
  auto step_of = [&](int s, int i) {
    Step st;
    const int ek = LB.ops[s][i].epi_kind;
    st.kind = ek == EPI_HEAD ? STEP_HEAD : (ek == EPI_INGRAD ? STEP_INGRAD : STEP_EPI);
    st.tslot = (uint8_t)s; st.op = (uint8_t)op_index(s, i); st.arg = 0;
    return st;
  };
  auto aux_step = [&](int kind, int s, int op, int arg) {
    Step st;
    st.kind = (uint8_t)kind; st.tslot = (uint8_t)s; st.op = (uint8_t)op; st.arg = (uint8_t)arg;
    return st;
  };
  // ---- N phase: both tile slots op by op; slot 0 acquires the weights, slot 1 releases them
  int cursor = 0;
  for (int i = 0; i < (carried ? 0 : LB.n_narrow); ++i) {
    std::vector<BurstH> b0 = make_bursts(LB.ops[0][i], LB.weights[i]);
    std::vector<BurstH> b1 = make_bursts(LB.ops[1][i], LB.weights[i]);
    if ((int)LB.ops[0][i].kcs.size() > NUNIT - 1) { err = "tensor-core engine: narrow layer with too many K-chunks for the weight ring"; return false; }
    {
      int unit_of[64];
      std::fill(unit_of, unit_of + 64, -1);
      assign_units(b0, cursor, unit_of, true, false);
      assign_units(b1, cursor, unit_of, false, true);
    }
    bursts.insert(bursts.end(), b0.begin(), b0.end());
    bursts.insert(bursts.end(), b1.begin(), b1.end());
    steps.push_back(step_of(0, i));
    if (!(early_l1 && i == LB.n_narrow - 1)) steps.push_back(step_of(1, i));
  }
  // ---- T phase: tile slot 0, then tile slot 1
  for (int s = 0; s < 2; ++s) {
    int prep_next = 0;
    for (int i = LB.n_narrow; i < n_fwd; ++i) {
      std::vector<BurstH> b = make_bursts(LB.ops[s][i], LB.weights[i]);
      if (carried && i == LB.trunk_first)       // inputs come from the PREP steps; the other slot must have left the T phase
        b[0].flags = (uint16_t)((b[0].flags & ~B_WAIT_GLUE) | B_WAIT_PREP | B_WAIT_DONE_OTHER);
      if (carried && i == LB.trunk_first + 1) b[0].flags &= (uint16_t)~B_PEEK_GLUE_OTHER;
      // (early_l1: slot 1's last per-sample stage signals done[1] right after its tensor-memory read)
      if (early_l1 && s == 0 && i == LB.trunk_first + 1) b[0].flags = (uint16_t)((b[0].flags & ~B_PEEK_GLUE_OTHER) | B_WAIT_DONE_OTHER);
      {
        int unit_of[64];
        std::fill(unit_of, unit_of + 64, -1);
        assign_units(b, cursor, unit_of, true, true);
      }
      bursts.insert(bursts.end(), b.begin(), b.end());
      steps.push_back(step_of(s, i));
      if (early_l1 && s == 0 && i == LB.trunk_first) steps.push_back(step_of(1, LB.n_narrow - 1));
      // next pair's sample fetch + viewdir features: off the critical path, in the slack after the second trunk layer
      // (carried programs fetch 18 carry planes per sample: the two halves go into the slack of different layers)
      if (i == LB.trunk_first + 1 || (carried && i == LB.trunk_first + 2)) {
        Step v; v.kind = STEP_VIEW; v.tslot = (uint8_t)s; v.op = 0;
        v.arg = (uint8_t)(carried ? (i == LB.trunk_first + 1 ? 1 : 2) : 0);
        steps.push_back(v);
      }
      // the input block of this tile slot is free once the skip layer (or layer 0) has consumed it: write the
      // next pair's features into it, one part per following layer
      if (i >= LB.trunk_skip_op && prep_next < NPREP) {
        Step p; p.kind = STEP_PREP; p.tslot = (uint8_t)s; p.op = 0; p.arg = (uint8_t)prep_next++;
        steps.push_back(p);
      }
    }
    for (; prep_next < NPREP; ++prep_next) {
      Step p; p.kind = STEP_PREP; p.tslot = (uint8_t)s; p.op = 0; p.arg = (uint8_t)prep_next;
      steps.push_back(p);
    }
    if (grad) {
      // reverse sweep through the trunk: seed = column 0 of the sigma head under the last layer's ReLU, then the
      // transposed-weight chain; the two input-gradient ops sum into the thread's 16 columns
      steps.push_back(aux_step(STEP_SEED, s, op_index(s, LB.tb_first), SEED_TRUNK));
      bool first_ingrad = true;
      for (int i = LB.tb_first; i < LB.tb_first + LB.tb_count; ++i) {
        std::vector<BurstH> b = make_bursts(LB.ops[s][i], LB.weights[i]);
        int unit_of[64];
        std::fill(unit_of, unit_of + 64, -1);
        assign_units(b, cursor, unit_of, true, true);
        bursts.insert(bursts.end(), b.begin(), b.end());
        Step st = step_of(s, i);
        if (st.kind == STEP_INGRAD) { st.arg = first_ingrad ? 1 : 0; first_ingrad = false; }
        steps.push_back(st);
      }
      steps.push_back(aux_step(STEP_PEB, s, 0, 0));
    }
    Step o; o.kind = STEP_OUT; o.tslot = (uint8_t)s; o.op = 0; o.arg = 0;
    steps.push_back(o);
  }
  if (grad) {
    // ---- reverse sweep through the narrow networks, both tile slots op by op like the N phase (slot 0 acquires the
    // weights, slot 1 releases them).  Tile slot 0's columns were last used by slot 1's T phase, which the compute
    // warps have left before they write the seeds the first bursts wait for.
    bool first_ingrad = true;
    auto sweep = [&](int first, int count, int seed) {
      if (count == 0) return true;
      for (int s = 0; s < 2; ++s) steps.push_back(aux_step(STEP_SEED, s, op_index(s, first), seed));
      for (int i = first; i < first + count; ++i) {
        std::vector<BurstH> b0 = make_bursts(LB.ops[0][i], LB.weights[i]);
        std::vector<BurstH> b1 = make_bursts(LB.ops[1][i], LB.weights[i]);
        if ((int)LB.ops[0][i].kcs.size() > NUNIT - 1) { err = "tensor-core engine: narrow layer with too many K-chunks for the weight ring"; return false; }
        int unit_of[64];
        std::fill(unit_of, unit_of + 64, -1);
        assign_units(b0, cursor, unit_of, true, false);
        assign_units(b1, cursor, unit_of, false, true);
        bursts.insert(bursts.end(), b0.begin(), b0.end());
        bursts.insert(bursts.end(), b1.begin(), b1.end());
        for (int s = 0; s < 2; ++s) {
          Step st = step_of(s, i);
          if (st.kind == STEP_INGRAD) st.arg = first_ingrad ? 1 : 0;
          steps.push_back(st);
        }
        if (LB.ops[0][i].epi_kind == EPI_INGRAD) first_ingrad = false;
      }
      return true;
    };
    if (!sweep(LB.hb_first, LB.hb_count, SEED_HYPER)) return false;
    if (!sweep(LB.wb_first, LB.wb_count, SEED_WARP)) return false;
    for (int s = 0; s < 2; ++s) steps.push_back(aux_step(STEP_PEB, s, 0, 1));
    // the next pair's N phase accumulates over BOTH slots' columns: besides slot 1 (B_WAIT_DONE_OTHER) it waits for
    // slot 0's reverse sweep
    bursts[0].ctl_extra |= CTL_WAIT_DONE_SELF;
  }
  if ((int)bursts.size() > MAX_BURST || (int)steps.size() > MAX_STEPS) { err = "tensor-core engine: layer program too long"; return false; }
  // ring safety: a unit is re-acquired only if its previous occupant was released at least two bursts earlier
  // (the issuer waits for burst i + 1's weights in the middle of burst i)
  {
    const int n = (int)bursts.size();
    std::vector<int> released(NUNIT, -1000000);
    std::vector<int> held(NUNIT, 0);
    for (int it = 0; it < 2; ++it)
      for (int i = 0; i < n; ++i) {
        const BurstH& e = bursts[i];
        const int gi = it * n + i;
        if (e.flags & B_ACQUIRE) {
          if (held[e.unit] || released[e.unit] > gi - 2) { err = "tensor-core engine: weight ring too small for this layer program"; return false; }
          held[e.unit] = 1;
        } else if (!held[e.unit]) { err = "tensor-core engine: internal ring schedule error"; return false; }
        if (e.unit2 != 0xff && !held[e.unit2]) { err = "tensor-core engine: internal ring schedule error (second unit)"; return false; }
        if (e.flags & B_RELEASE) { held[e.unit] = 0; released[e.unit] = gi; }
        if (e.release2) { held[e.unit2] = 0; released[e.unit2] = gi; }
      }
  }
  prog.n_burst = (int)bursts.size();
  prog.n_steps = (int)steps.size();
  prog.smem_bytes = OFF_CTRL + ctrl_bytes(prog.n_ops, prog.n_steps, prog.n_burst) + 1024;    // + manual 1024-byte alignment slack
  if (prog.smem_bytes > TC_SMEM_MAX) { err = "tensor-core engine: layer program does not fit the shared-memory control block"; return false; }
  encode_bursts(bursts, true, smem_base, prog.burst, prog.src);
  std::copy(steps.begin(), steps.end(), prog.steps);
  return true;
}

std::string tc_engine_supports(const ndsr_config& c, int cc_major, int cc_minor) {
  if (cc_major != 10) return "needs an sm_100-class device (tcgen05)";
  if (!c.use_warp) return "the tensor-core engine is built for models with an SE(3) warp field";
  if (c.rgb_depth != 1) return "rgb branch depth must be 1";
  if (!c.use_viewdirs) return "the rgb branch without viewdirs is not built";
  if (c.trunk_width != 256 || c.rgb_width != 128) return "the tensor-core rgb branch is built for trunk width 256 / rgb width 128";
  if (c.trunk_depth < 2) return "trunk depth must be >= 2";
  const int widths[] = {c.warp_width, c.use_hyper_sheet ? c.hyper_sheet_width : 64, c.use_predicted_mask ? c.mask_width : 64};
  for (int w : widths) if (w != 64 && w != 128) return "warp / hyper-sheet / mask MLP widths must be 64 or 128";
  (void)cc_minor;
  return "";
}

static int build_level(ndsr_handle* h, int lv, LevelBuild& LB, Packed& P) {
  const ndsr_config& c = h->cfg;
  const HostModel& HM = h->host_model;
  const int prec = c.precision;
  const int t_sigma = prec == NDSR_PREC_FP16 ? 1 : 3;
  const int t_rgb = prec == NDSR_PREC_SPLIT3 ? 3 : 1;     // mixed: 1-term rgb branch (median rgb error ~2e-4)
  if (h->max_in > 64) { h->err = "tensor-core engine: MLP inputs wider than 64 features"; return NDSR_ERR_UNSUPPORTED; }
  if (h->dim_view + (c.predict_norm ? h->dim_norm : 0) > 64) { h->err = "tensor-core engine: rgb side inputs wider than 64"; return NDSR_ERR_UNSUPPORTED; }
  // Which narrow networks issue their small split terms first (see issue_pair): the tensor core's accumulator
  // truncates, and the error that matters for the 1e-3 RGB bound is the one the SE(3) field makes -- a 1e-6 error
  // of the warped point is multiplied by 2^7 pi in the trunk's positional encoding (tools/tc_emulation.py:
  // truncation in the warp field alone gives 9e-4 of the 8e-4..9e-4 total on the worst rays of 16 384; in the mask
  // / hyper-sheet / trunk networks 1.5e-4 / 3e-5 / 1.3e-4).  Each cross-first layer costs one more burst per tile.
  int xf_warp = 1, xf_mask = 0, xf_hyper = 0;
  if (const char* e = getenv("NDS_TC_CROSS_FIRST")) {       // experiments: none | warp | all
    const std::string v(e);
    xf_warp = v != "none"; xf_mask = xf_hyper = v == "all";
  }
  // ---- shared feature block of the narrow networks
  FLayout& F = LB.F;
  {
    int kmin = c.warp_min_deg, kmax = c.warp_max_deg;
    if (c.use_hyper_sheet) { kmin = std::min(kmin, c.hyper_sheet_min_deg); kmax = std::max(kmax, c.hyper_sheet_max_deg); }
    if (c.use_predicted_mask) { kmin = std::min(kmin, c.mask_min_deg); kmax = std::max(kmax, c.mask_max_deg); }
    int col = 0;
    if (c.warp_use_posenc_identity) { F.col_ident = col; col += 3; }
    F.col_bands = col; F.kmin = kmin; F.nb = kmax - kmin; col += 6 * F.nb;
    F.col_wembed = col; col += c.warp_embed_dims;
    if (c.use_predicted_mask) { F.col_membed = col; col += c.mask_embed_dims; }
    if ((c.use_mask_in_warp) || (c.use_hyper_sheet && c.use_mask_in_hyper)) { F.col_mask = col; col += 1; }
    F.cols = col;
    if (col > 64) { h->err = "tensor-core engine: shared feature block wider than 64 columns"; return NDSR_ERR_UNSUPPORTED; }
  }
  LB.t_cols = h->dim_trunk_in;
  for (int s = 0; s < 2; ++s) {
    std::vector<OpBuild>& ops = LB.ops[s];
    const uint32_t in_off = OFF_IN + (uint32_t)s * 2u * KBLK;
    bool first_net = true;
    KChunkMap hyper_k0, hyper_ks, warp_k0, warp_ks;     // kept for the reverse sweep
    auto first_flags = [&]() {
      uint16_t f = first_net ? (uint16_t)(B_WAIT_PREP | (s == 0 ? B_WAIT_DONE_OTHER : 0)) : (uint16_t)B_WAIT_GLUE;
      first_net = false;
      return f;
    };
    if (c.use_predicted_mask) {
      const KChunkMap k0 = kc_features(F, in_off, 0, 0, c.mask_min_deg, c.mask_max_deg, 0, F.col_membed, c.mask_embed_dims, false);
      const KChunkMap ks = kc_features(F, in_off, HM.mask.width, 0, c.mask_min_deg, c.mask_max_deg, 0, F.col_membed, c.mask_embed_dims, false);
      build_head_op({&HM.mask.logit}, build_mlp_ops(HM.mask, t_sigma, false, s, k0, ks, first_flags(), ops, xf_mask), t_sigma, GLUE_MASK, false, s, 1, ops, xf_mask);
    }
    if (c.use_hyper_sheet) {
      const bool wm = c.use_mask_in_hyper != 0;
      const KChunkMap k0 = kc_features(F, in_off, 0, 2, c.hyper_sheet_min_deg, c.hyper_sheet_max_deg, 0, F.col_wembed, c.warp_embed_dims, wm);
      const KChunkMap ks = kc_features(F, in_off, HM.hyper.width, 2, c.hyper_sheet_min_deg, c.hyper_sheet_max_deg, 0, F.col_wembed, c.warp_embed_dims, wm);
      build_head_op({&HM.hyper.logit}, build_mlp_ops(HM.hyper, t_sigma, false, s, k0, ks, first_flags(), ops, xf_hyper, MASK_HYPER), t_sigma, GLUE_HYPER, false, s, 1, ops, xf_hyper);
      hyper_k0 = k0; hyper_ks = ks;
    }
    {
      const bool wm = c.use_mask_in_warp != 0;
      const KChunkMap k0 = kc_features(F, in_off, 0, 1, c.warp_min_deg, c.warp_max_deg, c.warp_use_posenc_identity, F.col_wembed, c.warp_embed_dims, wm);
      const KChunkMap ks = kc_features(F, in_off, HM.warp.width, 1, c.warp_min_deg, c.warp_max_deg, c.warp_use_posenc_identity, F.col_wembed, c.warp_embed_dims, wm);
      build_head_op({&HM.warp_w, &HM.warp_v}, build_mlp_ops(HM.warp, t_sigma, false, s, k0, ks, first_flags(), ops, xf_warp, MASK_WARP), t_sigma, GLUE_WARP, false, s, 1, ops, xf_warp);
      warp_k0 = k0; warp_ks = ks;
    }
    LB.n_narrow = (int)ops.size();
    // ---- template NeRF
    LB.trunk_first = (int)ops.size();
    const HostMlp& TR = HM.trunk[lv];
    LB.trunk_skip_op = LB.trunk_first + (TR.skip > 0 ? TR.skip : 0);
    const KChunkMap t0 = kc_input(in_off, 0, TR.in_dim), tsk = kc_input(in_off, TR.width, TR.in_dim);
    const uint16_t tflags = (uint16_t)(B_WAIT_GLUE | (s == 1 ? B_WAIT_DONE_OTHER : 0));
    const ActLayout trunk = build_mlp_ops(TR, t_sigma, true, s, t0, tsk, tflags, ops, false, MASK_TRUNK);
    build_head_op({&HM.alpha[lv]}, trunk, t_sigma, GLUE_ALPHA, true, s, 1, ops);
    LB.n_sigma = (int)ops.size();
    // ---- rgb branch (modules.py:288-313).  Flax input order of its Dense(560 -> 128):
    //   [bottleneck (W) | viewdir feats | trunk_out (W, App. C-1) | norm feats]
    // The bottleneck is a Dense WITHOUT activation (modules.py:283-286), so it is folded into this layer on the
    // host:  W_fold = W_bott W_rgb[bottleneck rows] + W_rgb[trunk_out rows],  b_fold = b_rgb + b_bott W_rgb[...]:
    // one GEMM over trunk_out (K = 256) plus the side-input K-chunk instead of two layers with K = 256 + 560.
    // trunk_out stays in its region T; in the other region B the sigma/normal head accumulates into columns
    // [0, 16) and this layer into [128, 256), so neither waits for the other; only the side-input K-chunk (the
    // normal features come out of the head's per-sample stage) is issued after the glue barrier.
    const int W = c.trunk_width;
    if (W != 256 || HM.rgb[lv].width != 128) { h->err = "tensor-core engine: the rgb branch is built for trunk 256 / rgb 128"; return NDSR_ERR_UNSUPPORTED; }
    const int Tcol = trunk.region * 256, Bcol = (1 - trunk.region) * 256;
    ActLayout rgbh;
    rgbh.N = 128; rgbh.n_nc = 1; rgbh.d_col[0] = rgbh.d_col[1] = Bcol + 128; rgbh.region = 1 - trunk.region;
    {
      const HostMlp& R = HM.rgb[lv];
      const HostDense& BT = HM.bottleneck[lv];
      OpBuild ob;
      const int NR = R.width;
      ob.N_logical = ob.N = NR;
      ob.tslot = s;
      ob.n_nc = 1; ob.d_col[0] = ob.d_col[1] = rgbh.d_col[0];
      ob.terms = t_rgb; ob.relu = 1; ob.epi_kind = t_rgb == 3 ? EPI_INPLACE : EPI_INPLACE_HI; ob.glue = GLUE_NONE; ob.prev_produces = 0;
      const std::vector<float>& WR = R.hidden[0].W;     // [560][NR]
      int row = W;
      const int v0 = row;
      row += h->dim_view;
      int x0 = -1;
      if (c.use_x_in_rgb_condition) { x0 = row; row += W; }
      const int n0 = row;
      const int ndim = c.predict_norm ? h->dim_norm : 0;
      const int side = h->dim_view + ndim;
      // folded logical weights: rows [0, W) over trunk_out, then the side inputs [viewdir feats | norm feats]
      ob.W_own.assign((size_t)(W + side) * NR, 0.f);
      ob.b_own.assign(NR, 0.f);
      for (int n = 0; n < NR; ++n) {
        double bacc = R.hidden[0].b[n];
        for (int j = 0; j < W; ++j) bacc += (double)BT.b[j] * WR[(size_t)j * NR + n];
        ob.b_own[n] = (float)bacc;
      }
      {
        std::vector<double> acc((size_t)NR);
        for (int k = 0; k < W; ++k) {
          for (int n = 0; n < NR; ++n) acc[n] = x0 >= 0 ? (double)WR[(size_t)(x0 + k) * NR + n] : 0.0;
          for (int j = 0; j < W; ++j) {
            const double bkj = BT.W[(size_t)k * W + j];
            const float* wr = &WR[(size_t)j * NR];
            for (int n = 0; n < NR; ++n) acc[n] += bkj * wr[n];
          }
          for (int n = 0; n < NR; ++n) ob.W_own[(size_t)k * NR + n] = (float)acc[n];
        }
      }
      for (int i = 0; i < h->dim_view; ++i) for (int n = 0; n < NR; ++n) ob.W_own[(size_t)(W + i) * NR + n] = WR[(size_t)(v0 + i) * NR + n];
      for (int i = 0; i < ndim; ++i) for (int n = 0; n < NR; ++n) ob.W_own[(size_t)(W + h->dim_view + i) * NR + n] = WR[(size_t)(n0 + i) * NR + n];
      for (int j = 0; j < W / 64; ++j) ob.kcs.push_back(kc_hidden(trunk, j, 0, false));   // trunk_out (the head before it waited)
      if (side > 0) {                                    // side inputs last: the normal features arrive late
        KChunkMap k = kc_input(OFF_IN2, W, side);
        k.wait_glue = 1;
        k.terms = 1;                                     // the side-input block holds hi halves only
        ob.kcs.push_back(k);
      }
      ops.push_back(ob);
      // head over the rgb hidden layer; accumulators at the start of T (trunk_out is dead: in-order MMA pipe)
      OpBuild hb;
      hb.N_logical = R.logit.N;
      hb.tslot = s;
      hb.N = 16; hb.n_nc = 1; hb.d_col[0] = hb.d_col[1] = Tcol;
      hb.terms = t_rgb; hb.relu = 0; hb.epi_kind = EPI_HEAD; hb.glue = GLUE_RGB; hb.prev_produces = 1; hb.signal_glue = 0;
      hb.W = &R.logit.W; hb.b = &R.logit.b;
      for (int j = 0; j < R.width / 64; ++j) hb.kcs.push_back(kc_hidden(rgbh, j, 0, true));
      ops.push_back(hb);
    }
    // ---- reverse sweep for -d(sigma_raw)/dx (models.py:1035-1077): trunk, then hyper sheet and SE(3) field
    LB.n_fwd = (int)ops.size();
    LB.tb_first = (int)ops.size();
    build_mlp_backward(TR, t_sigma, true, s, t0, tsk, MASK_TRUNK, false, ops);
    LB.tb_count = (int)ops.size() - LB.tb_first;
    LB.hb_first = (int)ops.size();
    if (c.use_hyper_sheet) build_mlp_backward(HM.hyper, t_sigma, false, s, hyper_k0, hyper_ks, MASK_HYPER, xf_hyper, ops);
    LB.hb_count = (int)ops.size() - LB.hb_first;
    LB.wb_first = (int)ops.size();
    build_mlp_backward(HM.warp, t_sigma, false, s, warp_k0, warp_ks, MASK_WARP, xf_warp, ops);
    LB.wb_count = (int)ops.size() - LB.wb_first;
    fix_own(ops);
  }
  for (size_t i = 0; i < LB.ops[0].size(); ++i) LB.weights.push_back(pack_weights(LB.ops[0][i], P));
  return NDSR_OK;
}

int tc_engine_load(ndsr_handle* h) {
  tc_engine_free(h);
  TcEngine* E = new TcEngine();
  h->tc = E;
  uint32_t smem_base = 0;
  if (!query_smem_base(smem_base, h->err)) return NDSR_ERR_CUDA;
  for (int lv = 0; lv < 2; ++lv) {
    int rc = build_level(h, lv, E->lb[lv], E->packed[lv]);
    if (rc) return rc;
    Packed& P = E->packed[lv];
    for (int mode = 0; mode < 4; ++mode)
      if (!assemble(E->lb[lv], mode != 0, mode == 2, mode == 3, smem_base, E->prog[lv][mode], h->err)) return NDSR_ERR_UNSUPPORTED;
    cudaError_t e;
    if ((e = cudaMalloc(&E->d_stream[lv], P.stream.size())) != cudaSuccess ||
        (e = cudaMalloc(&E->d_bias[lv], P.bias.size() * sizeof(float))) != cudaSuccess ||
        (e = cudaMemcpy(E->d_stream[lv], P.stream.data(), P.stream.size(), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(E->d_bias[lv], P.bias.data(), P.bias.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) {
      h->err = std::string("tc_engine_load: ") + cudaGetErrorString(e);
      return NDSR_ERR_CUDA;
    }
    // the streams were packed with unit windows
    for (int p = 0; p < 3; ++p) for (int k = 0; k < NDSR_MAX_BANDS; ++k) E->win[lv][p][k] = 1.f;
    E->win_valid[lv] = true;
  }
  cudaError_t e = cudaFuncSetAttribute(field_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_MAX);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(field_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_MAX);
  if (e != cudaSuccess) { h->err = std::string("tc smem attribute: ") + cudaGetErrorString(e); return NDSR_ERR_CUDA; }
  return NDSR_OK;
}

void tc_engine_free(ndsr_handle* h) {
  if (!h->tc) return;
  for (int lv = 0; lv < 2; ++lv) {
    if (h->tc->d_stream[lv]) cudaFree(h->tc->d_stream[lv]);
    if (h->tc->d_bias[lv]) cudaFree(h->tc->d_bias[lv]);
  }
  delete h->tc;
  h->tc = nullptr;
}

// The posenc windows of the narrow networks (model_utils.py:419-436) scale features of the shared block, which
// all three networks read: they are folded into the first-layer / skip-layer weight images instead.  Re-packs
// those images (a few 16 KB tiles) when the windows of this call differ from the ones in the stream.
static int fold_windows(ndsr_handle* h, int lv, const CallParams& cp, cudaStream_t st) {
  TcEngine* E = h->tc;
  const PosencSpec* pes[3] = {&cp.pe_mask, &cp.pe_warp, &cp.pe_hsheet};
  bool same = E->win_valid[lv];
  for (int p = 0; p < 3 && same; ++p)
    for (int k = 0; k < NDSR_MAX_BANDS; ++k) if (E->win[lv][p][k] != pes[p]->window[k]) { same = false; break; }
  if (same) return NDSR_OK;
  Packed& P = E->packed[lv];
  for (const WindowedImage& wi : P.windowed) {
    float colscale[64], rowscale[128];
    for (int c = 0; c < 64; ++c) colscale[c] = wi.band[c] >= 0 ? pes[wi.pe]->window[wi.band[c]] : 1.f;
    for (int r = 0; r < 128; ++r) rowscale[r] = wi.row_band[r] >= 0 ? pes[wi.pe]->window[wi.row_band[r]] : 1.f;
    const size_t bytes = (size_t)wi.rows * 128 * (wi.terms == 3 ? 2 : 1);
    write_image(&P.stream[wi.stream_off], wi.w.data(), colscale, wi.rows, wi.terms, rowscale);
    cudaError_t e = cudaMemcpyAsync(E->d_stream[lv] + wi.stream_off, &P.stream[wi.stream_off], bytes, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { h->err = std::string("fold_windows: ") + cudaGetErrorString(e); return NDSR_ERR_CUDA; }
  }
  for (int p = 0; p < 3; ++p) for (int k = 0; k < NDSR_MAX_BANDS; ++k) E->win[lv][p][k] = pes[p]->window[k];
  E->win_valid[lv] = true;
  return NDSR_OK;
}

int tc_engine_field(ndsr_handle* h, const CallParams& cp, const FieldArgs& fa, cudaStream_t st) {
  TcEngine* E = h->tc;
  if (!E) { h->err = "tensor-core engine not loaded"; return NDSR_ERR_NOT_LOADED; }
  const int64_t tiles = (fa.n_samples_total + TM - 1) / TM;
  if (tiles == 0) return NDSR_OK;
  int rc = fold_windows(h, fa.level, cp, st);
  if (rc) return rc;
  TcKernelArgs K;
  if (fa.carry && fa.sigma_only) { h->err = "tensor-core engine: carried launches are built for the full program"; return NDSR_ERR_INVALID; }
  if (fa.need_grad && (fa.carry || fa.carry_out)) { h->err = "tensor-core engine: the reverse sweep runs in a single launch"; return NDSR_ERR_INVALID; }
  const TcProgram& prog = E->prog[fa.level][fa.need_grad ? 3 : (fa.carry ? 2 : (fa.sigma_only ? 0 : 1))];
  K.lvl.weights = E->d_stream[fa.level];
  K.lvl.bias = E->d_bias[fa.level];
  K.warp_embed = h->M.warp_embed;
  K.mask_embed = h->M.mask_embed;
  K.alpha_col0 = h->M.alpha_col0[fa.level];
  K.hyper_logit_w = h->cfg.use_hyper_sheet ? h->M.hyper.logit.W : nullptr;
  K.warp_w_w = h->M.warp_w.W;
  K.warp_v_w = h->M.warp_v.W;
  K.trace = nullptr;
  const char* trace_path = getenv("NDS_TC_TRACE");
  if (trace_path && !fa.sigma_only) {
    cudaMalloc(&K.trace, TRACE_WORDS * sizeof(unsigned long long));
    cudaMemsetAsync(K.trace, 0, TRACE_WORDS * sizeof(unsigned long long), st);
  }
  const int64_t pairs = (tiles + 1) / 2;
  int grid = (int)(pairs < h->num_sms ? pairs : h->num_sms);
  if (const char* g = getenv("NDS_TC_GRID")) { const int v = atoi(g); if (v > 0 && v < grid) grid = v; }   // experiments
  if (prog.grad) field_tc_kernel<true><<<grid, TC_THREADS, prog.smem_bytes, st>>>(prog, K, cp, fa, h->cfg);
  else field_tc_kernel<false><<<grid, TC_THREADS, prog.smem_bytes, st>>>(prog, K, cp, fa, h->cfg);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { h->err = std::string("field_tc_kernel launch: ") + cudaGetErrorString(e); return NDSR_ERR_CUDA; }
  if (K.trace) {   // diagnostics only: synchronous dump of the stamps, relative to the pair start
    std::vector<unsigned long long> t(TRACE_WORDS);
    cudaStreamSynchronize(st);
    cudaMemcpy(t.data(), K.trace, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(K.trace);
    const unsigned long long t0 = t[3 * MAX_BURST + 2 * MAX_STEPS];
    const std::string tp = std::string(trace_path) + (fa.carry ? ".carried" : "");
    if (FILE* f = fopen(tp.c_str(), "w")) {
      auto rel = [&](unsigned long long v) { return v ? (long long)(v - t0) : -1LL; };
      for (int i = 0; i < prog.n_burst; ++i) {
        const Burst& b = prog.burst[i];
        fprintf(f, "burst %d slot %d rows %d steps %d pat %d flags %d unit %d dcol %d top %lld ready %lld issued %lld\n", i,
                (b.ctl >> 18) & 1, ((b.idesc >> 17) & 63) * 8, (b.ctl >> 15) & 7, (b.ctl >> 13) & 3, b.ctl & 0x1fff,
                (b.ctl >> 19) & 3, b.d, rel(t[i]), rel(t[MAX_BURST + i]), rel(t[2 * MAX_BURST + i]));
      }
      for (int i = 0; i < prog.n_steps; ++i) {
        const Step& s = prog.steps[i];
        const TcOp& op = prog.ops[s.op];
        fprintf(f, "step %d kind %d slot %d op %d N %d glue %d arg %d start %lld end %lld\n", i, s.kind, s.tslot, s.op,
                (s.kind <= STEP_HEAD) ? op.N : 0, (s.kind <= STEP_HEAD) ? op.glue : 0, s.arg,
                rel(t[3 * MAX_BURST + 2 * i]), rel(t[3 * MAX_BURST + 2 * i + 1]));
#if NDS_SUBTRACE
        if (s.kind == STEP_EPI)
          for (int c = 0; c < op.n_nc; ++c)
            fprintf(f, "sub %d chunk %d waited %lld loaded %lld stored %lld arrived %lld\n", i, c, rel(t[TRACE_SUB + 8 * i + 4 * c]),
                    rel(t[TRACE_SUB + 8 * i + 4 * c + 1]), rel(t[TRACE_SUB + 8 * i + 4 * c + 2]), rel(t[TRACE_SUB + 8 * i + 4 * c + 3]));
#endif
      }
      fclose(f);
    }
  }
  h->launches++;
  return NDSR_OK;
}

}  // namespace nds

// tensor-core MACs the programs ISSUE per sample evaluation (split terms and padding included): out[level * 4 + mode],
// mode 0 sigma-only | 1 full | 2 full, carried | 3 full + reverse sweep
extern "C" int ndsr_tc_issued_macs(const ndsr_handle* h, double* out) {
  using namespace nds;
  if (!h || !out) return NDSR_ERR_INVALID;
  if (!h->tc) return NDSR_ERR_NOT_LOADED;
  for (int lv = 0; lv < 2; ++lv)
    for (int mode = 0; mode < 4; ++mode) {
      const TcProgram& P = h->tc->prog[lv][mode];
      double macs = 0;
      for (int i = 0; i < P.n_burst; ++i) {
        const Burst& b = P.burst[i];
        const int steps = (b.ctl >> 15) & 7, pat = (b.ctl >> 13) & 3, per = pat == PAT_SS ? steps : 4;
        const int groups = (b.ctl & (1u << 25)) ? 2 : ((b.ctl & B_TWO) ? 3 : 1);
        const double rows = (double)(((b.idesc >> 17) & 63u) * 8u);
        macs += (double)groups * per * rows * TM * 16.0;
      }
      out[lv * 4 + mode] = macs / (2.0 * TM);       // a program covers a pair of tiles
    }
  return NDSR_OK;
}

// ---------------------------------------------------------------------------
// diagnostics entry point: one Dense layer through the tensor-core machinery
// ---------------------------------------------------------------------------
extern "C" int ndsr_selftest_tc_dense(int device, int k_hid, int k_in, int n_out, int terms, int relu, int out_kind,
                                      const float* A, const float* W, const float* bias, float* out,
                                      float* out_readback) {
  using namespace nds;
  // out_kind: 0 = hidden layer (in-place hi + lo), 1 = head (n_out <= 16), 2 = hidden layer, hi only (1-term
  // consumer), 5 = hidden layer compacted hi (the bottleneck's epilogue; n_out = 256).  The k_hid activations are
  // read from tensor memory in the layout a k_hid-wide layer leaves, the k_in inputs from the shared IN block.
  if (k_hid % 64 || k_hid > 256 || k_in > 64 || k_in < 0 || n_out < 1 || n_out > 256 || (terms != 1 && terms != 3))
    return NDSR_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return NDSR_ERR_CUDA;
  const bool head = out_kind == 1;
  if (head && n_out > 16) return NDSR_ERR_INVALID;
  if (out_kind == 5 && n_out != 256) return NDSR_ERR_INVALID;
  std::vector<float> Wv(W, W + (size_t)(k_hid + k_in) * n_out), bv(bias, bias + n_out);
  OpBuild ob;
  ob.N_logical = n_out;
  ob.N = head ? 16 : (n_out <= 64 ? 64 : (n_out <= 128 ? 128 : 256));
  ob.n_nc = head ? 1 : (ob.N >= 256 ? 2 : 1);
  ob.interleave = 1;
  ob.d_col[0] = 256; ob.d_col[1] = 256 + (ob.n_nc == 2 ? 128 : 0);
  ob.terms = terms; ob.relu = relu; ob.glue = GLUE_SELFTEST; ob.first_flags = B_WAIT_GLUE; ob.prev_produces = 1;
  ob.epi_kind = head ? EPI_HEAD : (out_kind == 2 ? EPI_INPLACE_HI : (out_kind == 5 ? EPI_COMPACT_HI : EPI_INPLACE));
  ob.W = &Wv; ob.b = &bv;
  if (k_in > 0) ob.kcs.push_back(kc_input(OFF_IN, k_hid, k_in));
  if (k_hid > 0) {
    ActLayout in;
    in.N = k_hid; in.n_nc = k_hid >= 256 ? 2 : 1; in.d_col[0] = 0; in.d_col[1] = 128;
    for (int j = 0; j < k_hid / 64; ++j) ob.kcs.push_back(kc_hidden(in, j, 0, true));
  }
  Packed P;
  const OpWeights ow = pack_weights(ob, P);
  std::vector<BurstH> bursts = make_bursts(ob, ow);
  static TcProgram prog;
  memset(&prog, 0, sizeof prog);
  prog.n_ops = 1;
  prog.ops[0] = make_tcop(ob, ow);
  int cursor = 0;
  {
    int unit_of[64];
    std::fill(unit_of, unit_of + 64, -1);
    assign_units(bursts, cursor, unit_of, true, true);
  }
  prog.n_burst = (int)bursts.size();
  prog.smem_bytes = OFF_CTRL + ctrl_bytes(prog.n_ops, prog.n_steps, prog.n_burst) + 1024;
  {
    std::string perr;
    if (!query_smem_base(prog.smem_base, perr)) { fprintf(stderr, "%s\n", perr.c_str()); return NDSR_ERR_CUDA; }
  }
  encode_bursts(bursts, false, prog.smem_base, prog.burst, prog.src);
  uint8_t* d_stream; float *d_bias, *d_A, *d_out, *d_rb;
  const int N = prog.ops[0].N;
  const int K = k_hid + k_in;
  cudaMalloc(&d_stream, P.stream.size()); cudaMalloc(&d_bias, P.bias.size() * 4);
  cudaMalloc(&d_A, (size_t)TM * K * 4); cudaMalloc(&d_out, (size_t)TM * N * 4); cudaMalloc(&d_rb, (size_t)TM * N * 4);
  cudaMemcpy(d_stream, P.stream.data(), P.stream.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(d_bias, P.bias.data(), P.bias.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_A, A, (size_t)TM * K * 4, cudaMemcpyHostToDevice);
  cudaMemset(d_out, 0, (size_t)TM * N * 4); cudaMemset(d_rb, 0, (size_t)TM * N * 4);
  cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_MAX);
  TcLevel L; L.weights = d_stream; L.bias = d_bias;
  tc_selftest_kernel<<<1, TC_THREADS, prog.smem_bytes>>>(prog, L, d_A, k_hid, k_in, d_out, d_rb);
  cudaError_t e = cudaDeviceSynchronize();
  int rc = NDSR_OK;
  if (e != cudaSuccess) { fprintf(stderr, "ndsr_selftest_tc_dense: %s\n", cudaGetErrorString(e)); rc = NDSR_ERR_CUDA; }
  else {
    // outputs are [128][N] padded; return the logical [128][n_out]
    std::vector<float> tmp((size_t)TM * N), tmp2((size_t)TM * N);
    cudaMemcpy(tmp.data(), d_out, tmp.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(tmp2.data(), d_rb, tmp2.size() * 4, cudaMemcpyDeviceToHost);
    for (int r = 0; r < TM; ++r) for (int c = 0; c < n_out; ++c) {
      out[(size_t)r * n_out + c] = tmp[(size_t)r * N + c];
      if (out_readback) out_readback[(size_t)r * n_out + c] = tmp2[(size_t)r * N + c];
    }
  }
  cudaFree(d_stream); cudaFree(d_bias); cudaFree(d_A); cudaFree(d_out); cudaFree(d_rb);
  return rc;
}
