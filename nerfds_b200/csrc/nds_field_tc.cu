// tcgen05 tensor-core field engine (sm_100a).
//
// One persistent CTA per SM processes tiles of 128 samples.  Every Dense layer
// of the path (hypernerf/modules.py:57-83) is a [128 x K] x [K x N] GEMM issued
// as tcgen05.mma (kind::f16, fp32 accumulators in TMEM):
//   * A = the tile's activations as split fp16 (hi + lo), kept IN TENSOR MEMORY
//     for the whole chain (TS-mode MMA).  The 512 TMEM columns form two regions
//     of 256: layer l reads its operand from one region and accumulates into the
//     other; the epilogue converts the accumulators IN PLACE (tcgen05.ld ->
//     bias/ReLU -> split -> tcgen05.st over the columns it just read) into the
//     operand of layer l+1, which accumulates back into the first region.
//     Network inputs (posenc features, embeddings, mask) are the only shared-
//     memory operands (canonical K-major SWIZZLE_128B block, SS-mode MMA);
//   * B = the layer's weights, pre-packed on the host into the exact
//     shared-memory image (scaled by a power of two, split hi + lo, swizzled)
//     and streamed image by image through a 12-slot bulk-TMA (cp.async.bulk) ring;
//   * "3-term" layers issue A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (~fp32 accuracy,
//     needed on the sigma path for the 1e-3 RGB bound -- tools/precision_study.py),
//     "1-term" layers issue A_hi*B_hi only (bottleneck, rgb branch).
// Wide layers are split into two N-chunks with separate accumulators and
// barriers: while the tensor core works on chunk 1, the 16 compute warps drain
// chunk 0, and the next layer's MMAs start on the K-range chunk 0 produced
// before chunk 1's epilogue has finished.
// Warp roles: warps 0-15 = compute (TMEM lane quarter = warp % 4 -> 32 samples,
// column slice / feature slice = warp / 4): positional encodings, SE(3)
// exponential, epilogues; warp 16 = MMA issuer; warp 17 = TMA producer.
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "nds_host.h"
#include "nds_tc.cuh"

namespace nds {

using namespace tc;

constexpr int TM = 128;                 // samples per tile (UMMA M)
constexpr int NSUB = 4;                 // compute warps per TMEM lane quarter
constexpr int N_CWARPS = 4 * NSUB;
constexpr int WARP_MMA = N_CWARPS, WARP_TMA = N_CWARPS + 1;
constexpr int TC_THREADS = (N_CWARPS + 2) * 32;
constexpr uint32_t KBLK = 16384;        // one 128-row K-block (64 fp16 columns)
// shared memory map
constexpr uint32_t OFF_IN_HI = 0;
constexpr uint32_t OFF_IN_LO = KBLK;
constexpr uint32_t OFF_RING = 2 * KBLK;
constexpr uint32_t SLOT_BYTES = 16384;  // one weight image: <= 128 rows x 64 fp16
constexpr int NSLOT = 12;
constexpr uint32_t OFF_CTRL = OFF_RING + NSLOT * SLOT_BYTES;
constexpr uint32_t TC_SMEM_BYTES = OFF_CTRL + 512 + 1024;   // + manual 1024-byte alignment slack
// tensor memory: two regions of 256 columns; accumulator chunk c of an op sits at region + 128 c
constexpr uint32_t TM_REGION = 256;
constexpr int MAX_KC = 10;

enum EpiKind : uint8_t {
  EPI_INPLACE = 0,      // slice of CW accumulator columns -> hi (CW/2 columns) | lo (CW/2 columns) over the same slice
  EPI_INPLACE_HI = 1,   // hi only (consumer is a 1-term layer)
  EPI_COMPACT_HI = 2,   // hi only, compacted to the first half of the chunk (frees the second half for accumulators)
  EPI_HEAD = 4          // <= 16 outputs consumed by the per-sample stage
};
enum Glue : uint8_t { GLUE_NONE = 0, GLUE_MASK = 1, GLUE_WARP = 2, GLUE_HYPER = 3, GLUE_ALPHA = 4, GLUE_BOTTLENECK = 5,
                      GLUE_RGB = 6, GLUE_SELFTEST = 7 };
// A-operand addressing pattern of an image: 0 = shared memory (IN block), else tensor memory with the K-step
// column offsets of nds_tc.cuh (32 / 16 / 8)
enum APattern : uint8_t { PAT_SS = 0, PAT_32 = 1, PAT_16 = 2, PAT_8 = 3 };

// What the compute warps need to know about one Dense layer.
struct TcOp {
  uint32_t bias_off;     // float offset into the bias array
  float inv_scale;       // accumulators hold (scale * W) x; multiply back
  uint16_t N;            // output columns (padded)
  uint16_t nc_rows;      // output columns per N-chunk
  uint8_t n_nc, relu, epi_kind, glue;
  uint16_t d_col[2];     // tensor-memory column of accumulator chunk c
};

// What the MMA issuer / TMA producer need to know about one weight image
// (= one ring slot = up to 2 terms x 4 K-steps of tcgen05.mma).
enum ImgFlags : uint16_t {
  IMG_TWO_TERMS = 2, IMG_FIRST = 4, IMG_LAST = 8, IMG_NC1 = 16, IMG_WAIT_P0 = 32, IMG_WAIT_P1 = 64,
  IMG_WAIT_GLUE = 128,
  IMG_PART_NEXT = 256,   // the issuer has consumed one output phase of the previous op
  IMG_PAIR_LAST = 1024, IMG_PAIR_PART_NEXT = 2048   // copies of the B_lo image's LAST / PART_NEXT on its B_hi image
};
struct alignas(16) ImgEntry {
  uint32_t a_hi;         // PAT_SS: byte offset from the (1024-aligned) smem base; else tensor-memory column
  uint32_t a_lo;
  uint16_t rows;         // rows of the image = N of its MMAs (bytes = rows * 128)
  uint8_t steps;         // K-steps, 1..4
  uint8_t pat;           // APattern
  uint16_t flags;
  uint16_t d_col;        // accumulator column
};
static_assert(sizeof(ImgEntry) == 16, "ImgEntry is read as one uint4");

constexpr int MAX_OPS = 48;
constexpr int MAX_IMG = 448;
struct TcProgram {       // passed by value as a __grid_constant__ kernel parameter (constant bank)
  int n_ops, n_img;
  TcOp ops[MAX_OPS];
  ImgEntry img[MAX_IMG];
};

constexpr int TRACE_X = 2 * MAX_IMG + 4 * MAX_OPS + 8;       // issuer loop-top / after-waits stamps
constexpr int TRACE_WORDS = TRACE_X + 2 * MAX_IMG;

struct TcLevel {
  const uint8_t* weights;
  const float* bias;
};

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
struct Ctrl {
  uint64_t full[NSLOT];
  uint64_t empty[NSLOT];
  uint64_t in_ready;        // compute -> MMA: inputs of the next network written, previous head consumed
  uint64_t part_ready[2];   // compute -> MMA: N-chunk c of the current op drained and re-written as operand
  uint64_t d_full[2];       // MMA -> compute: accumulators of N-chunk c complete
  uint32_t tmem_base;
};

__device__ __forceinline__ void ctrl_init(Ctrl* ctl) {
  for (int i = 0; i < NSLOT; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); }
  mbar_init(&ctl->in_ready, N_CWARPS);
  mbar_init(&ctl->part_ready[0], N_CWARPS);
  mbar_init(&ctl->part_ready[1], N_CWARPS);
  mbar_init(&ctl->d_full[0], 1);
  mbar_init(&ctl->d_full[1], 1);
  mbar_fence_init();
}

// one arrival per compute warp, after every lane made its writes visible
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  fence_proxy_async_smem();     // st.shared operand images -> async proxy (tcgen05.mma)
  tmem_st_wait();               // tcgen05.st operand images complete
  tc_fence_before_sync();
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred;
}

// MMA issuer.  The whole warp walks the image program in uniform control flow (so the operand arithmetic
// runs on the uniform datapath); one elected lane issues the tcgen05 instructions.  A 3-term K-chunk (B_hi
// image followed by its B_lo image, 12 tcgen05.mma) goes out as one asm burst.
__device__ __forceinline__ void issue_tile(const TcProgram& P, uint32_t smem_base, Ctrl* ctl, uint32_t tmem_base,
                                           uint32_t lead, uint32_t& slot, uint32_t& phase, uint32_t& part_cnt,
                                           uint32_t& glue_cnt, unsigned long long* trace) {
  const int n_img = P.n_img;
  const uint32_t empty0 = smem_u32(&ctl->empty[0]);
  const uint32_t ring_lo32 = smem_desc_lo32(smem_base + OFF_RING);
  uint4 raw = *reinterpret_cast<const uint4*>(&P.img[0]);
  int i = 0;
  while (i < n_img) {
    const uint32_t a_hi = raw.x, a_lo = raw.y, rows = raw.z & 0xffffu, steps = (raw.z >> 16) & 0xffu, pat = raw.z >> 24;
    const uint32_t fl = raw.w & 0xffffu, d = tmem_base + (raw.w >> 16);
    const bool two = (fl & IMG_TWO_TERMS) != 0;
    const int adv = two ? 2 : 1;
    if (i + adv < n_img) raw = *reinterpret_cast<const uint4*>(&P.img[i + adv]);
    if (trace && lead) trace[TRACE_X + i] = clock64();
    if (fl & IMG_WAIT_GLUE) { mbar_wait(&ctl->in_ready, glue_cnt & 1u); ++glue_cnt; }
    if (fl & IMG_WAIT_P0) mbar_wait(&ctl->part_ready[0], part_cnt & 1u);
    if (fl & IMG_WAIT_P1) mbar_wait(&ctl->part_ready[1], part_cnt & 1u);
    const uint32_t s0 = slot, p0 = phase;
    uint32_t s1 = slot + 1, p1 = phase;
    if (s1 == NSLOT) { s1 = 0; p1 ^= 1u; }
    mbar_wait(&ctl->full[s0], p0);
    if (two) mbar_wait(&ctl->full[s1], p1);
    if (trace && lead) trace[TRACE_X + MAX_IMG + i] = clock64();
    tc_fence_after_sync();
    if (trace && lead) trace[i] = clock64();
    const uint32_t idesc = make_idesc_f16(rows);
    const uint32_t b0 = ring_lo32 + s0 * (SLOT_BYTES >> 4), b1 = ring_lo32 + s1 * (SLOT_BYTES >> 4);
    const uint32_t acc = (fl & IMG_FIRST) ? 0u : 1u;
    const uint32_t dbar = smem_u32(&ctl->d_full[(fl & IMG_NC1) ? 1 : 0]);
    const uint32_t last = (fl & (two ? IMG_PAIR_LAST : IMG_LAST)) ? 1u : 0u;
    const uint32_t e0 = empty0 + 8 * s0, e1 = empty0 + 8 * s1;
    if (steps == 4 && two) {
      const uint32_t A0 = tmem_base + a_hi, A1 = tmem_base + a_lo;
      if (pat == PAT_32) umma_burst3_ts32(d, A0, A1, b0, b1, idesc, acc, e0, e1, last, dbar, lead);
      else if (pat == PAT_16) umma_burst3_ts16(d, A0, A1, b0, b1, idesc, acc, e0, e1, last, dbar, lead);
      else if (pat == PAT_8) umma_burst3_ts8(d, A0, A1, b0, b1, idesc, acc, e0, e1, last, dbar, lead);
      else umma_burst3_ss(d, smem_desc_lo32(smem_base + a_hi), smem_desc_lo32(smem_base + a_lo), b0, b1, idesc, acc, e0,
                          e1, last, dbar, lead);
    } else if (steps == 4) {
      const uint32_t A0 = tmem_base + a_hi;
      if (pat == PAT_32) umma_burst1_ts32(d, A0, b0, idesc, acc, e0, last, dbar, lead);
      else if (pat == PAT_16) umma_burst1_ts16(d, A0, b0, idesc, acc, e0, last, dbar, lead);
      else if (pat == PAT_8) umma_burst1_ts8(d, A0, b0, idesc, acc, e0, last, dbar, lead);
      else umma_burst1_ss(d, smem_desc_lo32(smem_base + a_hi), b0, idesc, acc, e0, last, dbar, lead);
    } else {
      // short K-chunks: network inputs in shared memory (the host packer only emits PAT_SS here)
      const uint64_t bd0 = ((uint64_t)NDS_DESC_HI << 32) | b0, bd1 = ((uint64_t)NDS_DESC_HI << 32) | b1;
      const uint32_t lead2 = two ? lead : 0u;
      const uint64_t ad0 = make_smem_desc(smem_base + a_hi), ad1 = make_smem_desc(smem_base + a_lo);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_f16_p(d, ad0 + 2 * ks, bd0 + 2 * ks, idesc, ks ? 1u : acc, ks < (int)steps ? lead : 0u);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_f16_p(d, ad1 + 2 * ks, bd0 + 2 * ks, idesc, 1u, ks < (int)steps ? lead2 : 0u);
      umma_commit_p(&ctl->empty[s0], lead);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_f16_p(d, ad0 + 2 * ks, bd1 + 2 * ks, idesc, 1u, ks < (int)steps ? lead2 : 0u);
      umma_commit_p(&ctl->empty[s1], lead2);
      umma_commit_p(&ctl->d_full[(fl & IMG_NC1) ? 1 : 0], last ? lead : 0u);
    }
    if (trace && lead) trace[MAX_IMG + i] = clock64();
    if (two) { slot = s1; phase = p1; }
    if (++slot == NSLOT) { slot = 0; phase ^= 1u; }
    if (fl & (two ? IMG_PAIR_PART_NEXT : IMG_PART_NEXT)) ++part_cnt;
    i += adv;
  }
}

// TMA producer: streams every image of the program into the ring
__device__ __forceinline__ void produce_tile(const TcProgram& P, const uint8_t* wstream, uint8_t* smem, Ctrl* ctl,
                                             bool leader, uint32_t& slot, uint32_t& phase, bool& wrapped) {
  const uint8_t* src = wstream;
  for (int i = 0; i < P.n_img; ++i) {
    const uint32_t bytes = (uint32_t)P.img[i].rows * 128u;
    if (wrapped) mbar_wait(&ctl->empty[slot], phase ^ 1u);     // the previous fill of this slot has been consumed
    if (leader) {
      mbar_arrive_expect_tx(&ctl->full[slot], bytes);
      tma_bulk_g2s(smem + OFF_RING + slot * SLOT_BYTES, src, bytes, &ctl->full[slot]);
    }
    __syncwarp();
    src += bytes;
    if (++slot == NSLOT) { slot = 0; phase ^= 1u; wrapped = true; }
  }
}

// write one feature (col c of the IN block) of row r, split
__device__ __forceinline__ void store_in(uint8_t* smem, uint32_t r, uint32_t c, float v) {
  __half h, l;
  split_h(v, h, l);
  const uint32_t o = kblock_offset(r, c);
  *reinterpret_cast<__half*>(smem + OFF_IN_HI + o) = h;
  *reinterpret_cast<__half*>(smem + OFF_IN_LO + o) = l;
}

// Epilogue of one N-chunk for this thread's row and its CW-column slice: accumulators -> scale/bias/ReLU ->
// split fp16 -> written over the very columns just read (hi | lo), the operand of the next layer.
template <int CW>
__device__ __forceinline__ void epilogue_chunk(const TcOp& op, int nc, const float* __restrict__ bias_base,
                                               uint32_t tmem_lane, uint32_t row, int sub, int q, float* dbg_out,
                                               int dbg_ld, uint64_t* bar, uint32_t parity) {
  const uint32_t oc0 = (uint32_t)nc * op.nc_rows + (uint32_t)sub * CW;   // first output column of this slice
  const float4* bias4 = reinterpret_cast<const float4*>(bias_base + op.bias_off + oc0);
  float b[CW];
#pragma unroll
  for (int i = 0; i < CW / 4; ++i) {     // issued before the accumulators are awaited: the latency hides in the wait
    const float4 t = __ldg(bias4 + i);
    b[4 * i] = t.x; b[4 * i + 1] = t.y; b[4 * i + 2] = t.z; b[4 * i + 3] = t.w;
  }
  mbar_wait(bar, parity);
  tc_fence_after_sync();
  const uint32_t chunk = tmem_lane + op.d_col[nc];
  const uint32_t col = chunk + (uint32_t)sub * CW;
  uint32_t v[CW];
  tmem_ld<CW>(col, v);
  tmem_ld_wait();
  const float inv = op.inv_scale;
  const bool relu = op.relu != 0;
  uint32_t hi[CW / 2], lo[CW / 2];
#pragma unroll
  for (int i = 0; i < CW / 2; ++i) {
    float x0 = fmaf(__uint_as_float(v[2 * i]), inv, b[2 * i]);
    float x1 = fmaf(__uint_as_float(v[2 * i + 1]), inv, b[2 * i + 1]);
    if (relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
    if (dbg_out) { dbg_out[row * dbg_ld + oc0 + 2 * i] = x0; dbg_out[row * dbg_ld + oc0 + 2 * i + 1] = x1; }
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    hi[i] = *reinterpret_cast<const uint32_t*>(&h);
    lo[i] = *reinterpret_cast<const uint32_t*>(&l);
  }
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
}

// waits for the chunk's accumulators (bar / parity) inside, after the bias prefetch
__device__ __forceinline__ void epilogue_dispatch(const TcOp& op, int nc, const float* bias_base, uint32_t tmem_lane,
                                                  uint32_t row, int sub, int q, float* dbg, int dbg_ld, uint64_t* bar,
                                                  uint32_t parity) {
  if (op.nc_rows == 128) epilogue_chunk<32>(op, nc, bias_base, tmem_lane, row, sub, q, dbg, dbg_ld, bar, parity);
  else epilogue_chunk<16>(op, nc, bias_base, tmem_lane, row, sub, q, dbg, dbg_ld, bar, parity);
}

// head (<= 16 outputs): every compute warp of the lane quarter reads all of them
__device__ __forceinline__ void epilogue_head(const TcOp& op, const float* __restrict__ bias_base, uint32_t tmem_lane,
                                              float* hv) {
  uint32_t v[16];
  tmem_ld16(tmem_lane + op.d_col[0], v);
  tmem_ld_wait();
  const float* bias = bias_base + op.bias_off;
#pragma unroll
  for (int i = 0; i < 16; ++i) hv[i] = fmaf(__uint_as_float(v[i]), op.inv_scale, __ldg(bias + i));
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

// ---------------------------------------------------------------------------
// the field kernel
// ---------------------------------------------------------------------------
struct TcKernelArgs {
  TcLevel lvl;
  const float* warp_embed;
  const float* mask_embed;
  unsigned long long* trace;   // diagnostics (NDS_TC_TRACE): clock64 stamps of CTA 0's second tile, else null
};
// trace layout: [i] image i ready to issue | [MAX_IMG + i] image i issued | [2 MAX_IMG + 4 op + 2 nc] accumulators seen,
// [.. + 1] operand written & signalled | [2 MAX_IMG + 4 MAX_OPS] tile start


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

  if (threadIdx.x == 0) ctrl_init(ctl);
  if (warp == WARP_MMA) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  const int64_t n_tiles = (a.n_samples_total + TM - 1) / TM;
  const TcLevel& L = K.lvl;

  if (warp == WARP_TMA) {
    // ===================== TMA producer =====================
    const bool leader = elect_one() != 0;
    uint32_t slot = 0, phase = 0;
    bool wrapped = false;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
      produce_tile(P, L.weights, smem, ctl, leader, slot, phase, wrapped);
  } else if (warp == WARP_MMA) {
    // ===================== MMA issuer =====================
    const uint32_t lead = elect_one();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);      // warp-uniform for the compiler
    uint32_t slot = 0, phase = 0, part_cnt = 0, glue_cnt = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
      issue_tile(P, smem_base, ctl, tb, lead, slot, phase, part_cnt, glue_cnt,
                 (K.trace && tile == (int64_t)gridDim.x) ? K.trace : nullptr);
  } else {
    // ===================== compute warps =====================
    const int q = warp & 3, sub = warp >> 2;
    const uint32_t row = (uint32_t)q * 32u + (uint32_t)lane;
    const uint32_t tmem_lane = tmem_base + (((uint32_t)q * 32u) << 16);
    uint32_t dcnt0 = 0, dcnt1 = 0;
    // per-sample inputs are fetched one tile ahead, so their global-memory latency hides behind the previous tile
    struct Sample { float x[3], vd[3], gt; int64_t ray; uint32_t wid; bool valid; };
    auto load_sample = [&](int64_t tile_, Sample& s_) {
      const int64_t n_ = tile_ * TM + row;
      s_.valid = n_ < a.n_samples_total;
      s_.x[0] = s_.x[1] = s_.x[2] = 0.f; s_.vd[0] = s_.vd[1] = s_.vd[2] = 0.f; s_.gt = 0.f; s_.ray = 0; s_.wid = 0;
      if (!s_.valid) return;
      s_.ray = n_ / a.S;
      if (a.points) { s_.x[0] = a.points[n_ * 3]; s_.x[1] = a.points[n_ * 3 + 1]; s_.x[2] = a.points[n_ * 3 + 2]; }
      else {
        const float z = a.z[n_];
        s_.x[0] = a.origins[s_.ray * 3 + 0] + z * a.dirs[s_.ray * 3 + 0];
        s_.x[1] = a.origins[s_.ray * 3 + 1] + z * a.dirs[s_.ray * 3 + 1];
        s_.x[2] = a.origins[s_.ray * 3 + 2] + z * a.dirs[s_.ray * 3 + 2];
      }
      s_.vd[0] = a.viewdirs[s_.ray * 3]; s_.vd[1] = a.viewdirs[s_.ray * 3 + 1]; s_.vd[2] = a.viewdirs[s_.ray * 3 + 2];
      if (a.warp_id) s_.wid = a.warp_id[s_.ray];
      if (a.gt_mask) s_.gt = a.gt_mask[s_.ray];
    };
    Sample nxt;
    load_sample(blockIdx.x, nxt);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t n = tile * TM + row;
      const Sample cur = nxt;
      const bool valid = cur.valid;
      unsigned long long* tr = (K.trace && tile == (int64_t)gridDim.x && threadIdx.x == 0) ? K.trace + 2 * MAX_IMG : nullptr;
      if (tr) tr[4 * MAX_OPS] = clock64();
      const float x[3] = {cur.x[0], cur.x[1], cur.x[2]};
      float xw[3] = {0.f, 0.f, 0.f}, om[2] = {0.f, 0.f};
      float maskv = cur.gt, pmask = 0.f, sigma_raw = 0.f, nrm[3] = {0.f, 0.f, 0.f}, rgb[3] = {0.f, 0.f, 0.f};
      SE3<float> T;
      for (int i = 0; i < 9; ++i) T.R[i] = (i % 4 == 0) ? 1.f : 0.f;
      T.p[0] = T.p[1] = T.p[2] = 0.f;
      const int64_t ray = cur.ray;
      const uint32_t wid = cur.wid;
      auto st_in = [&](int c, float v) { store_in(smem, row, (uint32_t)c, v); };
      // columns [from, 64) of the IN block are zero (their weight rows are zero, but 0 x garbage could be NaN)
      auto zero_in = [&](int from) { for (int c = from; c < 64; ++c) if ((c & (NSUB - 1)) == sub) st_in(c, 0.f); };
      auto extras = [&](int o, const float* embed, int dims, bool with_mask) {
        if (sub == 1) for (int e = 0; e < dims; ++e) st_in(o + e, __ldg(embed + (size_t)wid * dims + e));
        o += dims;
        if (with_mask) { if (sub == 2) st_in(o, maskv); ++o; }
        return o;
      };
      // inputs of the networks of the chain
      auto prep_mask_in = [&]() {
        int pair = 0;
        int o = posenc_emit_sub(x[0], x[1], x[2], 3, cp.pe_mask, st_in, 0, sub, pair);
        o = extras(o, K.mask_embed, cfg.mask_embed_dims, false);
        zero_in(o);
      };
      auto prep_warp_in = [&]() {
        int pair = 0;
        int o = posenc_emit_sub(x[0], x[1], x[2], 3, cp.pe_warp, st_in, 0, sub, pair);
        o = extras(o, K.warp_embed, cfg.warp_embed_dims, cfg.use_mask_in_warp != 0);
        zero_in(o);
      };
      auto prep_hyper_in = [&]() {
        int pair = 0;
        int o = posenc_emit_sub(x[0], x[1], x[2], 3, cp.pe_hsheet, st_in, 0, sub, pair);
        o = extras(o, K.warp_embed, cfg.warp_embed_dims, cfg.use_mask_in_hyper != 0);
        zero_in(o);
      };
      auto prep_trunk_in = [&]() {
        int pair = 0;
        int o = posenc_emit_sub(xw[0], xw[1], xw[2], 3, cp.pe_spatial, st_in, 0, sub, pair);
        if (H > 0) o = posenc_emit_sub(om[0], om[1], 0.f, H, cp.pe_hyperpt, st_in, o, sub, pair);
        zero_in(o);
      };
      auto after_mask = [&]() { if (cfg.use_warp) prep_warp_in(); else prep_trunk_in(); };
      auto after_warp = [&]() { if (cfg.use_hyper_sheet) prep_hyper_in(); else prep_trunk_in(); };
      if (cfg.use_predicted_mask) prep_mask_in();
      else if (cfg.use_warp) prep_warp_in();
      else { xw[0] = x[0]; xw[1] = x[1]; xw[2] = x[2]; prep_trunk_in(); }
      warp_arrive(&ctl->in_ready, lane);
      if (tile + gridDim.x < n_tiles) load_sample(tile + gridDim.x, nxt);

      for (int i = 0; i < P.n_ops; ++i) {
        const TcOp& op = P.ops[i];
        if (op.epi_kind != EPI_HEAD) {
          for (int nc = 0; nc < op.n_nc; ++nc) {
            uint32_t par;
            if (nc == 0) par = dcnt0++ & 1u; else par = dcnt1++ & 1u;
            if (tr) { mbar_wait(&ctl->d_full[nc], par); tr[4 * i + 2 * nc] = clock64(); }
            epilogue_dispatch(op, nc, L.bias, tmem_lane, row, sub, q, nullptr, 0, &ctl->d_full[nc], par);
            warp_arrive(&ctl->part_ready[nc], lane);
            if (op.n_nc == 1) warp_arrive(&ctl->part_ready[1], lane);   // keeps both barriers on one phase per op
            if (tr) tr[4 * i + 2 * nc + 1] = clock64();
          }
          continue;
        }
        mbar_wait(&ctl->d_full[0], dcnt0 & 1u);
        ++dcnt0;
        tc_fence_after_sync();
        if (tr) tr[4 * i] = clock64();
        float hv[16];
        epilogue_head(op, L.bias, tmem_lane, hv);
        if (tr) tr[4 * i + 2] = clock64();
        switch (op.glue) {
          case GLUE_MASK: {            // models.py:967-975
            pmask = cfg.mask_output_relu ? fmaxf(hv[0], 0.f) : hv[0];
            maskv = a.gt_mask ? (pmask * cp.mask_ratio + maskv * (1.f - cp.mask_ratio)) : pmask * cp.mask_ratio;
            after_mask();
          } break;
          case GLUE_WARP: {            // warping.py:217-232
            exp_se3<float>(hv, hv + 3, T);
            for (int c = 0; c < 3; ++c) xw[c] = T.R[c * 3 + 0] * x[0] + T.R[c * 3 + 1] * x[1] + T.R[c * 3 + 2] * x[2] + T.p[c];
            after_warp();
          } break;
          case GLUE_HYPER: {
            for (int c = 0; c < H; ++c) om[c] = hv[c];
            prep_trunk_in();
          } break;
          case GLUE_ALPHA: {
            sigma_raw = hv[0];
            if (cfg.predict_norm) { nrm[0] = hv[1]; nrm[1] = hv[2]; nrm[2] = hv[3]; }
            if (!a.sigma_only) {
              // rgb branch side inputs: [viewdir feats | normal-input feats] in the IN block
              int o = 0, pair = 0;
              if (cfg.use_viewdirs) {
                o = posenc_emit_sub(cur.vd[0], cur.vd[1], cur.vd[2], 3, cp.pe_view, st_in, 0, sub, pair);
              }
              if (cp.use_predicted_norm) {
                float nh[3], ni[3];
                normalize3(nrm, nh);
                if (cfg.use_warp) { for (int c = 0; c < 3; ++c) ni[c] = T.R[0 * 3 + c] * nh[0] + T.R[1 * 3 + c] * nh[1] + T.R[2 * 3 + c] * nh[2]; }
                else { ni[0] = nh[0]; ni[1] = nh[1]; ni[2] = nh[2]; }
                normalize3(ni, nh);
                if (cfg.norm_input_posenc) o = posenc_emit_sub(nh[0], nh[1], nh[2], 3, cp.pe_norm, st_in, o, sub, pair);
                else { if (sub == 0) { st_in(o, nh[0]); st_in(o + 1, nh[1]); st_in(o + 2, nh[2]); } o += 3; }
              }
              zero_in(o);
            }
          } break;
          case GLUE_RGB: {
            for (int c = 0; c < 3; ++c) rgb[c] = 1.f / (1.f + __expf(-hv[c]));
          } break;
          default: break;
        }
        // the MMA issuer may overwrite the head accumulators / read the new inputs from here on
        if (i != P.n_ops - 1) warp_arrive(&ctl->in_ready, lane);
        if (tr) tr[4 * i + 1] = clock64();
      }
      // ---- write planes (the warps sharing a sample take different planes) ----
      if (valid) {
        float* P = a.planes;
        const int64_t ps = a.plane_stride;
        if (sub == 0) {
          P[P_SIGMA_RAW * ps + n] = sigma_raw;
          for (int c = 0; c < 3; ++c) P[(P_RGB + c) * ps + n] = rgb[c];
        } else if (sub == 1) {
          for (int c = 0; c < 3; ++c) P[(P_NORM + c) * ps + n] = nrm[c];
          P[P_MASK * ps + n] = pmask;
        } else if (sub == 2) {
          for (int c = 0; c < 3; ++c) P[(P_WARPED + c) * ps + n] = xw[c];
          for (int c = 0; c < H; ++c) P[(P_WARPED + 3 + c) * ps + n] = om[c];
        } else if (cfg.use_warp) {
          const float r = 0.57735025882720947265625f;
          float rf[3], rn[3];
          for (int c = 0; c < 3; ++c) rf[c] = T.R[c * 3 + 0] * r + T.R[c * 3 + 1] * r + T.R[c * 3 + 2] * r;
          normalize3(rf, rn);
          for (int c = 0; c < 3; ++c) { P[(P_ROT + c) * ps + n] = rn[c]; P[(P_TRANS + c) * ps + n] = T.p[c]; }
        }
      }
      tc_fence_before_sync();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == WARP_MMA) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------
// self-test kernel: one op on caller-provided activations.  The k_hid hidden
// activations are written to tensor-memory region 0 in the layout an upstream
// epilogue of that width produces, the k_in inputs to the IN block.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_selftest_kernel(const __grid_constant__ TcProgram P, TcLevel L, const float* __restrict__ A, int k_hid, int k_in,
                   float* out_f32, float* out_readback) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Ctrl* ctl = reinterpret_cast<Ctrl*>(smem + OFF_CTRL);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) ctrl_init(ctl);
  if (warp == WARP_MMA) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;
  const TcOp& op = P.ops[0];
  if (warp == WARP_TMA) {
    const bool leader = elect_one() != 0;
    uint32_t slot = 0, phase = 0;
    bool wrapped = false;
    produce_tile(P, L.weights, smem, ctl, leader, slot, phase, wrapped);
  } else if (warp == WARP_MMA) {
    const uint32_t lead = elect_one();
    const uint32_t tb = __shfl_sync(0xffffffffu, tmem_base, 0);
    uint32_t slot = 0, phase = 0, pc = 0, gc = 0;
    issue_tile(P, smem_base, ctl, tb, lead, slot, phase, pc, gc, nullptr);
  } else {
    const int q = warp & 3, sub = warp >> 2;
    const uint32_t row = (uint32_t)q * 32u + (uint32_t)lane;
    const uint32_t tmem_lane = tmem_base + (((uint32_t)q * 32u) << 16);
    const int ld = k_hid + k_in;
    // producer layout of a k_hid-wide layer: chunks of nc_rows columns at region 0 + 128 c, slices of CW per warp
    const int p_nc = k_hid >= 128 ? 2 : 1, p_rows = k_hid / (p_nc ? p_nc : 1), p_cw = p_rows / NSUB;
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
      if ((c & (NSUB - 1)) == sub) store_in(smem, row, c, c < k_in ? A[row * ld + k_hid + c] : 0.f);
    warp_arrive(&ctl->part_ready[0], lane);
    warp_arrive(&ctl->part_ready[1], lane);
    warp_arrive(&ctl->in_ready, lane);
    if (op.epi_kind == EPI_HEAD) {
      mbar_wait(&ctl->d_full[0], 0);
      tc_fence_after_sync();
      float hv[16];
      epilogue_head(op, L.bias, tmem_lane, hv);
      if (sub == 0) for (int i = 0; i < 16 && i < op.N; ++i) out_f32[row * op.N + i] = hv[i];
    } else {
      for (int nc = 0; nc < op.n_nc; ++nc)
        epilogue_dispatch(op, nc, L.bias, tmem_lane, row, sub, q, out_f32, op.N, &ctl->d_full[nc], 0);
      tmem_st_wait();
      tc_fence_before_sync();
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  // read back the operand image the epilogue left in tensor memory -- what the next layer's MMA would see
  if (warp < N_CWARPS && op.epi_kind != EPI_HEAD && out_readback) {
    const int q = warp & 3, sub = warp >> 2;
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
  uint8_t pat;          // APattern
  uint8_t wait;         // bit c: needs output chunk c of the previous op
  uint32_t a_hi, a_lo;  // PAT_SS: shared byte offsets; else tensor-memory columns
  int rows[64];         // W row feeding each of the 64 operand columns (-1 = zero pad)
};

struct OpBuild {
  std::vector<KChunkMap> kcs;
  int N_logical;                      // real output columns
  int N;                              // padded
  int n_nc;                           // N-chunks (accumulators); nc_rows = N / n_nc
  int d_col[2];                       // tensor-memory column of accumulator chunk c
  int terms, relu, epi_kind, glue, wait_glue;
  int prev_produces = 0;              // the op before this one is a hidden layer (its epilogue signals part_ready)
  std::vector<float> W;               // [K_total][N_logical] logical weights (row-major, Flax layout)
  std::vector<float> b;
};

struct Packed {
  std::vector<TcOp> ops;
  std::vector<ImgEntry> imgs;
  std::vector<uint8_t> stream;
  std::vector<float> bias;
};

// Layout an in-place epilogue leaves behind (see epilogue_chunk): N output features in n_nc chunks at
// region + 128 c; inside a chunk, warp slice s holds features [s CW, (s+1) CW) as hi (CW/2 columns) | lo (CW/2).
struct ActLayout {
  int region, N, n_nc, compact_hi;    // compact_hi: EPI_COMPACT_HI (hi only, contiguous from the chunk start)
  int d_col[2];
  int nc_rows() const { return N / n_nc; }
  int cw() const { return nc_rows() / NSUB; }
};
static int chunks_for(int N) { return N >= 128 ? 2 : 1; }

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
  for (int cidx = 0; cidx < 64; ++cidx) k.rows[cidx] = (f0 + cidx < L.N) ? row0 + f0 + cidx : -1;
  return k;
}
static KChunkMap kc_input(int row0, int in_dim) {
  KChunkMap k;
  k.pat = PAT_SS; k.wait = 0; k.a_hi = OFF_IN_HI; k.a_lo = OFF_IN_LO;
  for (int c = 0; c < 64; ++c) k.rows[c] = c < in_dim ? row0 + c : -1;
  return k;
}

static void pack_op(const OpBuild& ob, Packed& out) {
  TcOp op;
  memset(&op, 0, sizeof op);
  op.N = (uint16_t)ob.N;
  op.relu = (uint8_t)ob.relu;
  op.epi_kind = (uint8_t)ob.epi_kind;
  op.glue = (uint8_t)ob.glue;
  op.n_nc = (uint8_t)ob.n_nc;
  op.nc_rows = (uint16_t)(ob.N / ob.n_nc);
  op.d_col[0] = (uint16_t)ob.d_col[0];
  op.d_col[1] = (uint16_t)ob.d_col[1];
  // power-of-two scale so that max |W| lands in [4, 8): keeps W_lo out of fp16 subnormals
  float mx = 0.f;
  for (float v : ob.W) mx = std::max(mx, std::fabs(v));
  int e = 0;
  if (mx > 0.f) { std::frexp(mx, &e); e = 3 - e; }
  if (e > 14) e = 14;
  if (e < -14) e = -14;
  const float scale = std::ldexp(1.f, e);
  op.inv_scale = std::ldexp(1.f, -e);
  op.bias_off = (uint32_t)out.bias.size();
  for (int n = 0; n < ob.N; ++n) out.bias.push_back(n < ob.N_logical ? ob.b[n] : 0.f);
  while (out.bias.size() % 4) out.bias.push_back(0.f);
  // stream order == issue order: N-chunk, K-chunk, image (hi, lo)
  const size_t img = (size_t)op.nc_rows * 128;
  const int n_img_per = ob.terms == 3 ? 2 : 1;
  bool waited0 = false, waited1 = false;
  for (int nc = 0; nc < op.n_nc; ++nc) {
    // the first image of chunk nc overwrites accumulator chunk nc, which the previous op's epilogue must have
    // drained (= its part nc written) -- or, after a head, the per-sample stage must be done (wait_glue).
    for (size_t kc = 0; kc < ob.kcs.size(); ++kc) {
      const KChunkMap& km = ob.kcs[kc];
      int last = -1;
      for (int c = 0; c < 64; ++c) if (km.rows[c] >= 0) last = c;
      const int steps = km.pat == PAT_SS ? std::max(1, (last + 16) / 16) : 4;
      const size_t base = out.stream.size();
      out.stream.resize(base + img * n_img_per, 0);
      for (int r = 0; r < op.nc_rows; ++r) {
        const int n = nc * op.nc_rows + r;
        for (int c = 0; c < 64; ++c) {
          float w = 0.f;
          if (n < ob.N_logical && km.rows[c] >= 0) w = ob.W[(size_t)km.rows[c] * ob.N_logical + n] * scale;
          const __half hi = __float2half_rn(w);
          const __half lo = __float2half_rn(w - __half2float(hi));
          const uint32_t o = kblock_offset((uint32_t)r, (uint32_t)c);
          memcpy(&out.stream[base + o], &hi, 2);
          if (ob.terms == 3) memcpy(&out.stream[base + img + o], &lo, 2);
        }
      }
      for (int im = 0; im < n_img_per; ++im) {
        ImgEntry ie;
        memset(&ie, 0, sizeof ie);
        ie.a_hi = km.a_hi; ie.a_lo = km.a_lo; ie.pat = km.pat;
        ie.rows = op.nc_rows;
        ie.steps = (uint8_t)steps;
        ie.d_col = op.d_col[nc];
        uint16_t fl = 0;
        if (im == 0 && ob.terms == 3) fl |= IMG_TWO_TERMS;     // B_hi image: A_hi B_hi + A_lo B_hi
        if (kc == 0 && im == 0) fl |= IMG_FIRST;
        if (kc + 1 == ob.kcs.size() && im + 1 == n_img_per) fl |= IMG_LAST;
        if (nc == 1) fl |= IMG_NC1;
        const bool first = kc == 0 && im == 0;
        // a single-chunk producer signals both part barriers; wait on the one matching the accumulator index
        if (first && ob.prev_produces && nc == 0 && !waited0) { fl |= IMG_WAIT_P0; waited0 = true; }
        if (first && ob.prev_produces && nc == 1 && !waited1) { fl |= IMG_WAIT_P1; waited1 = true; }
        if ((km.wait & 1) && !waited0) { fl |= IMG_WAIT_P0; waited0 = true; }
        if ((km.wait & 2) && !waited1) { fl |= IMG_WAIT_P1; waited1 = true; }
        if (first && nc == 0 && ob.wait_glue) fl |= IMG_WAIT_GLUE;
        ie.flags = fl;
        out.imgs.push_back(ie);
      }
    }
  }
  if (ob.prev_produces) out.imgs.back().flags |= IMG_PART_NEXT;   // one output phase of the previous op consumed
  out.ops.push_back(op);
}

// hidden stack of a modules.MLP: layer l reads [h (width) | inputs (in_dim) at the skip layer].  Layer l
// accumulates into region (l even ? 1 : 0) and leaves its activations there.  Returns the last layer's layout.
static ActLayout build_mlp_ops(const HostMlp& m, int terms, Packed& out) {
  ActLayout prev{};
  for (int l = 0; l < m.depth; ++l) {
    OpBuild ob;
    ob.N_logical = ob.N = m.width;
    ob.n_nc = chunks_for(m.width);
    const int region = (l % 2 == 0) ? 1 : 0;
    ob.d_col[0] = region * (int)TM_REGION;
    ob.d_col[1] = ob.d_col[0] + 128;
    ob.terms = terms; ob.relu = 1; ob.glue = GLUE_NONE; ob.epi_kind = EPI_INPLACE;
    ob.wait_glue = l == 0;
    ob.prev_produces = l > 0;
    ob.W = m.hidden[l].W; ob.b = m.hidden[l].b;
    if (l == 0) ob.kcs.push_back(kc_input(0, m.in_dim));
    else {
      if (l == m.skip) ob.kcs.push_back(kc_input(m.width, m.in_dim));   // ready long ago: issue it first
      for (int j = 0; j < m.width / 64; ++j) ob.kcs.push_back(kc_hidden(prev, j, 0, true));
    }
    pack_op(ob, out);
    prev = ActLayout{region, m.width, ob.n_nc, 0, {ob.d_col[0], ob.d_col[1]}};
  }
  return prev;
}

// head over the layer output `in`; accumulators in the other region
static void build_head_op(const std::vector<const HostDense*>& heads, const ActLayout& in, int terms, int glue,
                          Packed& out) {
  OpBuild ob;
  int n = 0;
  for (auto* h : heads) n += h->N;
  ob.N_logical = n;
  ob.N = 16;
  ob.n_nc = 1;
  ob.d_col[0] = (1 - in.region) * (int)TM_REGION;
  ob.d_col[1] = ob.d_col[0];
  ob.terms = terms; ob.relu = 0; ob.epi_kind = EPI_HEAD; ob.glue = glue; ob.wait_glue = 0; ob.prev_produces = 1;
  ob.W.assign((size_t)in.N * n, 0.f);
  int c0 = 0;
  for (auto* h : heads) {
    for (int k = 0; k < in.N; ++k) for (int j = 0; j < h->N; ++j) ob.W[(size_t)k * n + c0 + j] = h->W[(size_t)k * h->N + j];
    for (int j = 0; j < h->N; ++j) ob.b.push_back(h->b[j]);
    c0 += h->N;
  }
  for (int j = 0; j < in.N / 64; ++j) ob.kcs.push_back(kc_hidden(in, j, 0, true));
  pack_op(ob, out);
}

struct TcEngine {
  Packed packed[2];
  TcProgram prog[2];
  int n_ops_sigma[2] = {0, 0}, n_img_sigma[2] = {0, 0};
  uint8_t* d_stream[2] = {nullptr, nullptr};
  float* d_bias[2] = {nullptr, nullptr};
};

static bool make_program(const Packed& P, TcProgram& prog, std::string& err) {
  if (P.ops.size() > (size_t)MAX_OPS || P.imgs.size() > (size_t)MAX_IMG) { err = "tensor-core engine: layer program too long"; return false; }
  memset(&prog, 0, sizeof prog);
  prog.n_ops = (int)P.ops.size();
  prog.n_img = (int)P.imgs.size();
  std::copy(P.ops.begin(), P.ops.end(), prog.ops);
  std::copy(P.imgs.begin(), P.imgs.end(), prog.img);
  for (int i = 0; i + 1 < prog.n_img; ++i)
    if (prog.img[i].flags & IMG_TWO_TERMS) {
      if (prog.img[i + 1].flags & IMG_LAST) prog.img[i].flags |= IMG_PAIR_LAST;
      if (prog.img[i + 1].flags & IMG_PART_NEXT) prog.img[i].flags |= IMG_PAIR_PART_NEXT;
    }
  return true;
}

std::string tc_engine_supports(const ndsr_config& c, int cc_major, int cc_minor) {
  if (cc_major != 10) return "needs an sm_100-class device (tcgen05)";
  if (c.rgb_depth != 1) return "rgb branch depth must be 1";
  if (!c.use_viewdirs) return "the rgb branch without viewdirs is not built";
  if (c.trunk_width != 256 || c.rgb_width != 128) return "the tensor-core rgb branch is built for trunk width 256 / rgb width 128";
  const int widths[] = {c.trunk_width, c.rgb_width, c.use_warp ? c.warp_width : 64,
                        c.use_hyper_sheet ? c.hyper_sheet_width : 64, c.use_predicted_mask ? c.mask_width : 64};
  for (int w : widths) if (w != 64 && w != 128 && w != 256) return "MLP widths must be 64, 128 or 256";
  (void)cc_minor;
  return "";
}

static int build_level(ndsr_handle* h, int lv, Packed& P, int& n_sigma) {
  const ndsr_config& c = h->cfg;
  const HostModel& HM = h->host_model;
  const int prec = c.precision;
  const int t_sigma = prec == NDSR_PREC_FP16 ? 1 : 3;
  if (prec == NDSR_PREC_SPLIT3) { h->err = "tensor-core engine: split3 precision on the rgb branch is not built (use mixed)"; return NDSR_ERR_UNSUPPORTED; }
  if (h->max_in > 64) { h->err = "tensor-core engine: MLP inputs wider than 64 features"; return NDSR_ERR_UNSUPPORTED; }
  if (h->dim_view + (c.predict_norm ? h->dim_norm : 0) > 64) { h->err = "tensor-core engine: rgb side inputs wider than 64"; return NDSR_ERR_UNSUPPORTED; }
  if (c.use_predicted_mask) build_head_op({&HM.mask.logit}, build_mlp_ops(HM.mask, t_sigma, P), t_sigma, GLUE_MASK, P);
  if (c.use_warp) build_head_op({&HM.warp_w, &HM.warp_v}, build_mlp_ops(HM.warp, t_sigma, P), t_sigma, GLUE_WARP, P);
  if (c.use_hyper_sheet) build_head_op({&HM.hyper.logit}, build_mlp_ops(HM.hyper, t_sigma, P), t_sigma, GLUE_HYPER, P);
  const ActLayout trunk = build_mlp_ops(HM.trunk[lv], t_sigma, P);
  build_head_op({&HM.alpha[lv]}, trunk, t_sigma, GLUE_ALPHA, P);
  n_sigma = (int)P.ops.size();
  h->tc->n_img_sigma[lv] = (int)P.imgs.size();
  // ---- rgb branch (modules.py:288-313).  Flax input order:
  //   [bottleneck (W) | viewdir feats | trunk_out (W, App. C-1) | norm feats]
  // trunk_out stays in its region T; everything else happens in the other region B (256 columns):
  //   bottleneck   accumulates into B (two chunks of 128 columns), epilogue compacts its hi halves to the first
  //                64 columns of each chunk;
  //   rgb hidden   accumulates into the freed second halves (two chunks of 64 columns), hi-only in place;
  //   rgb head     accumulates into the first columns of T (trunk_out is dead by then: in-order MMA pipe).
  const int W = c.trunk_width;
  if (W != 256 || HM.rgb[lv].width != 128) { h->err = "tensor-core engine: the rgb branch is built for trunk 256 / rgb 128"; return NDSR_ERR_UNSUPPORTED; }
  const int Tcol = trunk.region * (int)TM_REGION, Bcol = (1 - trunk.region) * (int)TM_REGION;
  ActLayout bott{1 - trunk.region, W, 2, 1, {Bcol, Bcol + 128}};
  {
    OpBuild ob;
    ob.N_logical = ob.N = W;
    ob.n_nc = 2; ob.d_col[0] = Bcol; ob.d_col[1] = Bcol + 128;
    ob.terms = 1; ob.relu = 0; ob.epi_kind = EPI_COMPACT_HI; ob.glue = GLUE_BOTTLENECK;
    ob.wait_glue = 1;     // the sigma/normal head's accumulators (in B) must have been consumed
    ob.prev_produces = 0;
    ob.W = HM.bottleneck[lv].W; ob.b = HM.bottleneck[lv].b;
    for (int j = 0; j < W / 64; ++j) ob.kcs.push_back(kc_hidden(trunk, j, 0, false));
    pack_op(ob, P);
  }
  ActLayout rgbh{1 - trunk.region, 128, 2, 0, {Bcol + 64, Bcol + 192}};
  {
    const HostMlp& R = HM.rgb[lv];
    OpBuild ob;
    ob.N_logical = ob.N = R.width;
    ob.n_nc = 2; ob.d_col[0] = rgbh.d_col[0]; ob.d_col[1] = rgbh.d_col[1];
    ob.terms = 1; ob.relu = 1; ob.epi_kind = EPI_INPLACE_HI; ob.glue = GLUE_NONE; ob.wait_glue = 0; ob.prev_produces = 1;
    ob.W = R.hidden[0].W; ob.b = R.hidden[0].b;
    int row = W;
    const int v0 = row;
    row += h->dim_view;
    int x0 = -1;
    if (c.use_x_in_rgb_condition) { x0 = row; row += W; }
    const int n0 = row;
    const int ndim = c.predict_norm ? h->dim_norm : 0;
    // the accumulators overwrite the second halves of the bottleneck chunks: every image waits for the compaction
    if (x0 >= 0) for (int j = 0; j < W / 64; ++j) ob.kcs.push_back(kc_hidden(trunk, j, x0, false));   // trunk_out
    if (h->dim_view + ndim > 0) {
      KChunkMap k = kc_input(0, 0);
      for (int cidx = 0; cidx < 64; ++cidx) {
        if (cidx < h->dim_view) k.rows[cidx] = v0 + cidx;
        else if (cidx < h->dim_view + ndim) k.rows[cidx] = n0 + (cidx - h->dim_view);
        else k.rows[cidx] = -1;
      }
      ob.kcs.push_back(k);
    }
    for (int j = 0; j < W / 64; ++j) ob.kcs.push_back(kc_hidden(bott, j, 0, true));                   // bottleneck
    if ((int)ob.kcs.size() > MAX_KC) { h->err = "tensor-core engine: rgb input too wide"; return NDSR_ERR_UNSUPPORTED; }
    pack_op(ob, P);
    // head over the rgb hidden layer; accumulators at the start of T
    OpBuild hb;
    hb.N_logical = R.logit.N;
    hb.N = 16; hb.n_nc = 1; hb.d_col[0] = hb.d_col[1] = Tcol;
    hb.terms = 1; hb.relu = 0; hb.epi_kind = EPI_HEAD; hb.glue = GLUE_RGB; hb.wait_glue = 0; hb.prev_produces = 1;
    hb.W = R.logit.W; hb.b = R.logit.b;
    for (int j = 0; j < R.width / 64; ++j) hb.kcs.push_back(kc_hidden(rgbh, j, 0, true));
    pack_op(hb, P);
  }
  return NDSR_OK;
}

int tc_engine_load(ndsr_handle* h) {
  tc_engine_free(h);
  TcEngine* E = new TcEngine();
  h->tc = E;
  for (int lv = 0; lv < 2; ++lv) {
    int rc = build_level(h, lv, E->packed[lv], E->n_ops_sigma[lv]);
    if (rc) return rc;
    Packed& P = E->packed[lv];
    cudaError_t e;
    if (!make_program(P, E->prog[lv], h->err)) return NDSR_ERR_UNSUPPORTED;
    if ((e = cudaMalloc(&E->d_stream[lv], P.stream.size())) != cudaSuccess ||
        (e = cudaMalloc(&E->d_bias[lv], P.bias.size() * sizeof(float))) != cudaSuccess ||
        (e = cudaMemcpy(E->d_stream[lv], P.stream.data(), P.stream.size(), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(E->d_bias[lv], P.bias.data(), P.bias.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) {
      h->err = std::string("tc_engine_load: ") + cudaGetErrorString(e);
      return NDSR_ERR_CUDA;
    }
  }
  cudaError_t e = cudaFuncSetAttribute(field_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
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

int tc_engine_field(ndsr_handle* h, const CallParams& cp, const FieldArgs& fa, cudaStream_t st) {
  TcEngine* E = h->tc;
  if (!E) { h->err = "tensor-core engine not loaded"; return NDSR_ERR_NOT_LOADED; }
  TcKernelArgs K;
  TcProgram& prog = E->prog[fa.level];
  prog.n_ops = fa.sigma_only ? E->n_ops_sigma[fa.level] : (int)E->packed[fa.level].ops.size();
  prog.n_img = fa.sigma_only ? E->n_img_sigma[fa.level] : (int)E->packed[fa.level].imgs.size();
  K.lvl.weights = E->d_stream[fa.level];
  K.lvl.bias = E->d_bias[fa.level];
  K.warp_embed = h->M.warp_embed;
  K.mask_embed = h->M.mask_embed;
  K.trace = nullptr;
  const char* trace_path = getenv("NDS_TC_TRACE");
  if (trace_path && !fa.sigma_only) {
    cudaMalloc(&K.trace, TRACE_WORDS * sizeof(unsigned long long));
    cudaMemsetAsync(K.trace, 0, TRACE_WORDS * sizeof(unsigned long long), st);
  }
  const int64_t tiles = (fa.n_samples_total + TM - 1) / TM;
  if (tiles == 0) return NDSR_OK;
  int grid = (int)(tiles < h->num_sms ? tiles : h->num_sms);
  if (const char* g = getenv("NDS_TC_GRID")) { const int v = atoi(g); if (v > 0 && v < grid) grid = v; }   // experiments
  field_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(prog, K, cp, fa, h->cfg);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { h->err = std::string("field_tc_kernel launch: ") + cudaGetErrorString(e); return NDSR_ERR_CUDA; }
  if (K.trace) {   // diagnostics only: synchronous dump of the stamps, relative to the tile start
    std::vector<unsigned long long> t(TRACE_WORDS);
    cudaStreamSynchronize(st);
    cudaMemcpy(t.data(), K.trace, t.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(K.trace);
    const unsigned long long t0 = t[2 * MAX_IMG + 4 * MAX_OPS];
    if (FILE* f = fopen(trace_path, "w")) {
      auto rel = [&](unsigned long long v) { return v ? (long long)(v - t0) : -1LL; };
      for (int i = 0; i < prog.n_img; ++i)
        fprintf(f, "img %d rows %d steps %d flags %d ready %lld issued %lld top %lld waited %lld\n", i, prog.img[i].rows,
                prog.img[i].steps, prog.img[i].flags, rel(t[i]), rel(t[MAX_IMG + i]), rel(t[TRACE_X + i]),
                rel(t[TRACE_X + MAX_IMG + i]));
      for (int i = 0; i < prog.n_ops; ++i)
        fprintf(f, "op %d N %d kind %d glue %d c0_seen %lld c0_done %lld c1_seen %lld c1_done %lld\n", i, prog.ops[i].N,
                prog.ops[i].epi_kind, prog.ops[i].glue, rel(t[2 * MAX_IMG + 4 * i]), rel(t[2 * MAX_IMG + 4 * i + 1]),
                rel(t[2 * MAX_IMG + 4 * i + 2]), rel(t[2 * MAX_IMG + 4 * i + 3]));
      fclose(f);
    }
  }
  h->launches++;
  return NDSR_OK;
}

}  // namespace nds

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
  OpBuild ob;
  ob.N_logical = n_out;
  ob.N = head ? 16 : (n_out <= 64 ? 64 : (n_out <= 128 ? 128 : 256));
  ob.n_nc = head ? 1 : chunks_for(ob.N);
  ob.d_col[0] = (int)TM_REGION; ob.d_col[1] = (int)TM_REGION + (head ? 0 : 128);
  ob.terms = terms; ob.relu = relu; ob.glue = GLUE_SELFTEST; ob.wait_glue = 1; ob.prev_produces = 1;
  ob.epi_kind = head ? EPI_HEAD : (out_kind == 2 ? EPI_INPLACE_HI : (out_kind == 5 ? EPI_COMPACT_HI : EPI_INPLACE));
  const int K = k_hid + k_in;
  ob.W.assign(W, W + (size_t)K * n_out);
  ob.b.assign(bias, bias + n_out);
  if (k_in > 0) ob.kcs.push_back(kc_input(k_hid, k_in));
  if (k_hid > 0) {
    const ActLayout in{0, k_hid, chunks_for(k_hid), 0, {0, 128}};
    for (int j = 0; j < k_hid / 64; ++j) ob.kcs.push_back(kc_hidden(in, j, 0, true));
  }
  Packed P;
  pack_op(ob, P);
  uint8_t* d_stream; float *d_bias, *d_A, *d_out, *d_rb;
  const int N = P.ops[0].N;
  static TcProgram prog;
  std::string perr;
  if (!make_program(P, prog, perr)) return NDSR_ERR_INVALID;
  cudaMalloc(&d_stream, P.stream.size()); cudaMalloc(&d_bias, P.bias.size() * 4);
  cudaMalloc(&d_A, (size_t)TM * K * 4); cudaMalloc(&d_out, (size_t)TM * N * 4); cudaMalloc(&d_rb, (size_t)TM * N * 4);
  cudaMemcpy(d_stream, P.stream.data(), P.stream.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(d_bias, P.bias.data(), P.bias.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_A, A, (size_t)TM * K * 4, cudaMemcpyHostToDevice);
  cudaMemset(d_out, 0, (size_t)TM * N * 4); cudaMemset(d_rb, 0, (size_t)TM * N * 4);
  cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
  TcLevel L; L.weights = d_stream; L.bias = d_bias;
  tc_selftest_kernel<<<1, TC_THREADS, TC_SMEM_BYTES>>>(prog, L, d_A, k_hid, k_in, d_out, d_rb);
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
