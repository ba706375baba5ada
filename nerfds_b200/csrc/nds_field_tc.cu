// tcgen05 tensor-core field engine (sm_100a).
//
// One persistent CTA per SM processes tiles of 128 samples.  Every Dense layer
// of the path (hypernerf/modules.py:57-83) is a [128 x K] x [K x N] GEMM issued
// as tcgen05.mma (kind::f16, fp32 accumulators in TMEM) by ONE thread:
//   * A = the tile's activations, kept in shared memory as split fp16
//     (hi + lo, canonical K-major SWIZZLE_128B K-blocks) and rewritten in place
//     by the epilogue of the previous layer;
//   * B = the layer's weights, pre-packed on the host into the exact
//     shared-memory image (scaled by a power of two, split hi + lo, swizzled)
//     and streamed chunk by chunk through a bulk-TMA (cp.async.bulk) ring;
//   * "3-term" layers issue A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (~fp32 accuracy,
//     needed on the sigma path for the 1e-3 RGB bound -- tools/precision_study.py),
//     "1-term" layers issue A_hi*B_hi only (bottleneck, rgb branch).
// Warp roles: warps 0-3 = one thread per sample (TMEM lane): positional
// encodings, SE(3) exponential, epilogues (tcgen05.ld -> bias/ReLU -> split ->
// swizzled st.shared); warp 4 lane 0 = MMA issuer; warp 5 lane 0 = TMA producer.
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "nds_host.h"
#include "nds_tc.cuh"

namespace nds {

using namespace tc;

constexpr int TM = 128;                 // samples per tile (UMMA M)
constexpr int TC_THREADS = 192;
constexpr uint32_t KBLK = 16384;        // one 128-row K-block (64 fp16 columns)
constexpr uint32_t OFF_HID_HI = 0;
constexpr uint32_t OFF_HID_LO = 4 * KBLK;
constexpr uint32_t OFF_IN_HI = 8 * KBLK;
constexpr uint32_t OFF_IN_LO = 9 * KBLK;
constexpr uint32_t OFF_RING = 10 * KBLK;
constexpr uint32_t SLOT_BYTES = 32768;
constexpr int NSLOT = 2;
constexpr uint32_t OFF_CTRL = OFF_RING + NSLOT * SLOT_BYTES;
constexpr uint32_t TC_SMEM_BYTES = OFF_CTRL + 256 + 1024;   // + manual 1024-byte alignment slack
constexpr int MAX_KC = 10;

enum OutKind : uint8_t { OUT_HIDDEN = 0, OUT_HEAD = 1, OUT_HIDDEN_LO_REGION = 2 };
enum Glue : uint8_t { GLUE_NONE = 0, GLUE_MASK = 1, GLUE_WARP = 2, GLUE_HYPER = 3, GLUE_ALPHA = 4, GLUE_BOTTLENECK = 5,
                      GLUE_RGB = 6, GLUE_SELFTEST = 7 };
// A-operand source of a K-chunk
constexpr uint8_t SRC_IN = 4;           // 0..3: HID block j ; 4: IN block ; 8+j: HID_LO block j used as a 1-term operand

struct TcOp {
  uint32_t w_off;        // byte offset of this op's first chunk in the weight stream
  uint32_t bias_off;     // float offset into the bias array
  float inv_scale;       // accumulators hold (scale * W) x; multiply back
  uint16_t N;            // output columns, multiple of 16
  uint16_t nc_rows;      // rows (output columns) per N-chunk
  uint8_t n_kc, n_nc, terms, relu, out_kind, glue;
  uint8_t kc_src[MAX_KC];
  uint8_t kc_steps[MAX_KC];
};

struct TcLevel {
  const TcOp* ops;
  int n_ops;
  const uint8_t* weights;
  const float* bias;
};

// ---------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------
struct Ctrl {
  uint64_t full[NSLOT];
  uint64_t empty[NSLOT];
  uint64_t a_ready;
  uint64_t d_ready;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t a_region(uint8_t src, bool lo) {
  if (src < 4) return (lo ? OFF_HID_LO : OFF_HID_HI) + src * KBLK;
  if (src == SRC_IN) return lo ? OFF_IN_LO : OFF_IN_HI;
  return OFF_HID_LO + (src - 8) * KBLK;
}

// MMA issuer: all chunks of one op
__device__ __forceinline__ void issue_op(const TcOp& op, uint32_t smem_base, Ctrl* ctl, uint32_t& chunk_ctr) {
  const uint32_t idesc = make_idesc_f16(op.nc_rows);
  const uint32_t lo_off = (uint32_t)op.nc_rows * 128u;
  for (int kc = 0; kc < op.n_kc; ++kc) {
    const uint32_t a_hi = smem_base + a_region(op.kc_src[kc], false);
    const uint32_t a_lo = smem_base + a_region(op.kc_src[kc], true);
    const int steps = op.kc_steps[kc];
    for (int nc = 0; nc < op.n_nc; ++nc) {
      const uint32_t slot = chunk_ctr % NSLOT;
      mbar_wait(&ctl->full[slot], (chunk_ctr / NSLOT) & 1u);
      tc_fence_after_sync();
      const uint32_t b_hi = smem_base + OFF_RING + slot * SLOT_BYTES;
      const uint32_t d = ctl->tmem_base + (uint32_t)nc * op.nc_rows;
      uint32_t acc = kc > 0 ? 1u : 0u;
      for (int t = 0; t < op.terms; ++t) {
        const uint32_t A = (t == 1) ? a_lo : a_hi;
        const uint32_t B = (t == 2) ? b_hi + lo_off : b_hi;
        for (int ks = 0; ks < steps; ++ks) {
          umma_f16(d, make_smem_desc(A + ks * 32), make_smem_desc(B + ks * 32), idesc, acc);
          acc = 1u;
        }
      }
      umma_commit(&ctl->empty[slot]);   // frees the ring slot when these MMAs retire
      ++chunk_ctr;
    }
  }
  umma_commit(&ctl->d_ready);
}

// TMA producer: all chunks of one op
__device__ __forceinline__ void produce_op(const TcOp& op, const uint8_t* wstream, uint8_t* smem, Ctrl* ctl,
                                           uint32_t& chunk_ctr) {
  const uint32_t bytes = (uint32_t)op.nc_rows * 128u * (op.terms == 3 ? 2u : 1u);
  const uint8_t* src = wstream + op.w_off;
  const int n = op.n_kc * op.n_nc;
  for (int i = 0; i < n; ++i) {
    const uint32_t slot = chunk_ctr % NSLOT;
    if (chunk_ctr >= NSLOT) mbar_wait(&ctl->empty[slot], ((chunk_ctr / NSLOT) - 1u) & 1u);
    mbar_arrive_expect_tx(&ctl->full[slot], bytes);
    tma_bulk_g2s(smem + OFF_RING + slot * SLOT_BYTES, src, bytes, &ctl->full[slot]);
    src += bytes;
    ++chunk_ctr;
  }
}

// store 8 consecutive activations (cols c0..c0+7, c0 % 8 == 0) of row r as split fp16
__device__ __forceinline__ void store8_split(uint8_t* smem, uint32_t off_hi, uint32_t off_lo, bool write_lo,
                                             uint32_t r, uint32_t c0, const float* v) {
  __align__(16) __half2 hi[4];
  __align__(16) __half2 lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half h0, l0, h1, l1;
    split_h(v[2 * i], h0, l0);
    split_h(v[2 * i + 1], h1, l1);
    hi[i] = __halves2half2(h0, h1);
    lo[i] = __halves2half2(l0, l1);
  }
  const uint32_t o = (c0 >> 6) * KBLK + kblock_offset(r, c0 & 63u);
  *reinterpret_cast<uint4*>(smem + off_hi + o) = *reinterpret_cast<const uint4*>(hi);
  if (write_lo) *reinterpret_cast<uint4*>(smem + off_lo + o) = *reinterpret_cast<const uint4*>(lo);
}

// write one feature (col c of the IN block) of row r, split
__device__ __forceinline__ void store_in(uint8_t* smem, uint32_t r, uint32_t c, float v) {
  __half h, l;
  split_h(v, h, l);
  const uint32_t o = kblock_offset(r, c);
  *reinterpret_cast<__half*>(smem + OFF_IN_HI + o) = h;
  *reinterpret_cast<__half*>(smem + OFF_IN_LO + o) = l;
}

// Epilogue of one op for the calling thread's row.  Heads return their (<=16)
// outputs in hv[]; hidden layers are written back to shared memory.
__device__ __forceinline__ void epilogue_op(const TcOp& op, const float* __restrict__ bias_base, uint8_t* smem,
                                            uint32_t tmem_base, uint32_t row, float* hv, float* dbg_out, int dbg_ld) {
  const uint32_t taddr = tmem_base + ((row & ~31u) << 16);
  const float* bias = bias_base + op.bias_off;
  const float inv = op.inv_scale;
  if (op.out_kind == OUT_HEAD) {
    uint32_t v[16];
    tmem_ld16(taddr, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) hv[i] = fmaf(__uint_as_float(v[i]), inv, __ldg(bias + i));
    if (dbg_out) for (int i = 0; i < 16 && i < op.N; ++i) dbg_out[row * dbg_ld + i] = hv[i];
    return;
  }
  const uint32_t off_hi = (op.out_kind == OUT_HIDDEN_LO_REGION) ? OFF_HID_LO : OFF_HID_HI;
  const bool write_lo = op.out_kind == OUT_HIDDEN;
  for (uint32_t c0 = 0; c0 < op.N; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(taddr + c0, v);
    tmem_ld_wait();
    float f[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      float x = fmaf(__uint_as_float(v[i]), inv, __ldg(bias + c0 + i));
      f[i] = op.relu ? fmaxf(x, 0.f) : x;
    }
    if (dbg_out) for (int i = 0; i < 32; ++i) dbg_out[row * dbg_ld + c0 + i] = f[i];
#pragma unroll
    for (int g = 0; g < 4; ++g) store8_split(smem, off_hi, OFF_HID_LO, write_lo, row, c0 + 8 * g, f + 8 * g);
  }
}

// ---------------------------------------------------------------------------
// the field kernel
// ---------------------------------------------------------------------------
struct TcKernelArgs {
  TcLevel lvl;
  const float* warp_embed;
  const float* mask_embed;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
field_tc_kernel(const __grid_constant__ TcKernelArgs K, const __grid_constant__ CallParams cp,
                const __grid_constant__ FieldArgs a, const __grid_constant__ ndsr_config cfg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Ctrl* ctl = reinterpret_cast<Ctrl*>(smem + OFF_CTRL);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = cfg.use_hyper_sheet ? cfg.hyper_num_dims : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); }
    mbar_init(&ctl->a_ready, TM);
    mbar_init(&ctl->d_ready, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;

  const int64_t n_tiles = (a.n_samples_total + TM - 1) / TM;
  const TcLevel& L = K.lvl;

  if (warp == 5) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t cc = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int i = 0; i < L.n_ops; ++i) produce_op(L.ops[i], L.weights, smem, ctl, cc);
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t cc = 0, opc = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int i = 0; i < L.n_ops; ++i) {
          mbar_wait(&ctl->a_ready, opc & 1u);
          tc_fence_after_sync();
          issue_op(L.ops[i], smem_base, ctl, cc);
          ++opc;
        }
    }
  } else {
    // ===================== compute warps: one thread per sample =====================
    const uint32_t row = threadIdx.x;
    uint32_t opc = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t n = tile * TM + row;
      const bool valid = n < a.n_samples_total;
      float x[3] = {0.f, 0.f, 0.f}, xw[3] = {0.f, 0.f, 0.f}, om[2] = {0.f, 0.f};
      float maskv = 0.f, pmask = 0.f, sigma_raw = 0.f, nrm[3] = {0.f, 0.f, 0.f}, rgb[3] = {0.f, 0.f, 0.f};
      SE3<float> T;
      for (int i = 0; i < 9; ++i) T.R[i] = (i % 4 == 0) ? 1.f : 0.f;
      T.p[0] = T.p[1] = T.p[2] = 0.f;
      int64_t ray = 0;
      uint32_t wid = 0;
      if (valid) {
        ray = n / a.S;
        if (a.points) { x[0] = a.points[n * 3]; x[1] = a.points[n * 3 + 1]; x[2] = a.points[n * 3 + 2]; }
        else {
          const float z = a.z[n];
          x[0] = a.origins[ray * 3 + 0] + z * a.dirs[ray * 3 + 0];
          x[1] = a.origins[ray * 3 + 1] + z * a.dirs[ray * 3 + 1];
          x[2] = a.origins[ray * 3 + 2] + z * a.dirs[ray * 3 + 2];
        }
        if (a.warp_id) wid = a.warp_id[ray];
        if (a.gt_mask) maskv = a.gt_mask[ray];
      }
      auto st_in = [&](int c, float v) { store_in(smem, row, (uint32_t)c, v); };
      auto zero_in = [&](int from) { for (int c = from; c < 64; ++c) st_in(c, 0.f); };
      // inputs of the first network of the chain
      auto prep_mask_in = [&]() {
        int o = posenc_emit(x, 3, cp.pe_mask, st_in, 0);
        for (int e = 0; e < cfg.mask_embed_dims; ++e) st_in(o++, __ldg(K.mask_embed + (size_t)wid * cfg.mask_embed_dims + e));
        zero_in(o);
      };
      auto prep_warp_in = [&]() {
        int o = posenc_emit(x, 3, cp.pe_warp, st_in, 0);
        for (int e = 0; e < cfg.warp_embed_dims; ++e) st_in(o++, __ldg(K.warp_embed + (size_t)wid * cfg.warp_embed_dims + e));
        if (cfg.use_mask_in_warp) st_in(o++, maskv);
        zero_in(o);
      };
      auto prep_hyper_in = [&]() {
        int o = posenc_emit(x, 3, cp.pe_hsheet, st_in, 0);
        for (int e = 0; e < cfg.warp_embed_dims; ++e) st_in(o++, __ldg(K.warp_embed + (size_t)wid * cfg.warp_embed_dims + e));
        if (cfg.use_mask_in_hyper) st_in(o++, maskv);
        zero_in(o);
      };
      auto prep_trunk_in = [&]() {
        int o = posenc_emit(xw, 3, cp.pe_spatial, st_in, 0);
        if (H > 0) o = posenc_emit(om, H, cp.pe_hyperpt, st_in, o);
        zero_in(o);
      };
      auto after_mask = [&]() { if (cfg.use_warp) prep_warp_in(); else prep_trunk_in(); };
      auto after_warp = [&]() { if (cfg.use_hyper_sheet) prep_hyper_in(); else prep_trunk_in(); };
      if (cfg.use_predicted_mask) prep_mask_in();
      else if (cfg.use_warp) prep_warp_in();
      else { xw[0] = x[0]; xw[1] = x[1]; xw[2] = x[2]; prep_trunk_in(); }

      for (int i = 0; i < L.n_ops; ++i) {
        const TcOp& op = L.ops[i];
        fence_proxy_async_smem();          // my st.shared of A -> visible to tcgen05.mma
        mbar_arrive(&ctl->a_ready);
        mbar_wait(&ctl->d_ready, opc & 1u);
        ++opc;
        tc_fence_after_sync();
        float hv[16];
        epilogue_op(op, L.bias, smem, tmem_base, row, hv, nullptr, 0);
        tc_fence_before_sync();
        switch (op.glue) {
          case GLUE_MASK: {            // models.py:967-975
            pmask = cfg.mask_output_relu ? fmaxf(hv[0], 0.f) : hv[0];
            maskv = a.gt_mask ? (pmask * cp.mask_ratio + maskv * (1.f - cp.mask_ratio)) : pmask * cp.mask_ratio;
            after_mask();
          } break;
          case GLUE_WARP: {            // warping.py:217-232
            exp_se3<float>(hv, hv + 3, T);
            for (int q = 0; q < 3; ++q) xw[q] = T.R[q * 3 + 0] * x[0] + T.R[q * 3 + 1] * x[1] + T.R[q * 3 + 2] * x[2] + T.p[q];
            after_warp();
          } break;
          case GLUE_HYPER: {
            for (int q = 0; q < H; ++q) om[q] = hv[q];
            prep_trunk_in();
          } break;
          case GLUE_ALPHA: {
            sigma_raw = hv[0];
            if (cfg.predict_norm) { nrm[0] = hv[1]; nrm[1] = hv[2]; nrm[2] = hv[3]; }
            if (!a.sigma_only) {
              // rgb branch side inputs: [viewdir feats | normal-input feats] in the IN block
              int o = 0;
              if (cfg.use_viewdirs) {
                float vd[3] = {0.f, 0.f, 0.f};
                if (valid) { vd[0] = a.viewdirs[ray * 3]; vd[1] = a.viewdirs[ray * 3 + 1]; vd[2] = a.viewdirs[ray * 3 + 2]; }
                o = posenc_emit(vd, 3, cp.pe_view, st_in, 0);
              }
              if (cp.use_predicted_norm) {
                float nh[3], ni[3];
                normalize3(nrm, nh);
                if (cfg.use_warp) { for (int q = 0; q < 3; ++q) ni[q] = T.R[0 * 3 + q] * nh[0] + T.R[1 * 3 + q] * nh[1] + T.R[2 * 3 + q] * nh[2]; }
                else { ni[0] = nh[0]; ni[1] = nh[1]; ni[2] = nh[2]; }
                normalize3(ni, nh);
                if (cfg.norm_input_posenc) o = posenc_emit(nh, 3, cp.pe_norm, st_in, o);
                else { st_in(o++, nh[0]); st_in(o++, nh[1]); st_in(o++, nh[2]); }
              }
              zero_in(o);
            }
          } break;
          case GLUE_RGB: {
            for (int q = 0; q < 3; ++q) rgb[q] = 1.f / (1.f + __expf(-hv[q]));
          } break;
          default: break;
        }
      }
      // ---- write planes ----
      if (valid) {
        float* P = a.planes;
        const int64_t ps = a.plane_stride;
        P[P_SIGMA_RAW * ps + n] = sigma_raw;
        for (int q = 0; q < 3; ++q) {
          P[(P_RGB + q) * ps + n] = rgb[q];
          P[(P_NORM + q) * ps + n] = nrm[q];
          P[(P_WARPED + q) * ps + n] = xw[q];
        }
        for (int q = 0; q < H; ++q) P[(P_WARPED + 3 + q) * ps + n] = om[q];
        P[P_MASK * ps + n] = pmask;
        if (cfg.use_warp) {
          const float r = 0.57735025882720947265625f;
          float rf[3], rn[3];
          for (int q = 0; q < 3; ++q) rf[q] = T.R[q * 3 + 0] * r + T.R[q * 3 + 1] * r + T.R[q * 3 + 2] * r;
          normalize3(rf, rn);
          for (int q = 0; q < 3; ++q) { P[(P_ROT + q) * ps + n] = rn[q]; P[(P_TRANS + q) * ps + n] = T.p[q]; }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------
// self-test kernel: one op on caller-provided activations
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_selftest_kernel(TcLevel L, const float* __restrict__ A, int k_hid, int k_in, float* out_f32, float* out_readback) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Ctrl* ctl = reinterpret_cast<Ctrl*>(smem + OFF_CTRL);
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < NSLOT; ++i) { mbar_init(&ctl->full[i], 1); mbar_init(&ctl->empty[i], 1); }
    mbar_init(&ctl->a_ready, TM);
    mbar_init(&ctl->d_ready, 1);
    mbar_fence_init();
  }
  if (warp == 4) tmem_alloc(&ctl->tmem_base, 512);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = ctl->tmem_base;
  const TcOp& op = L.ops[0];
  if (warp == 5) {
    if (lane == 0) { uint32_t cc = 0; produce_op(op, L.weights, smem, ctl, cc); }
  } else if (warp == 4) {
    if (lane == 0) {
      uint32_t cc = 0;
      mbar_wait(&ctl->a_ready, 0);
      tc_fence_after_sync();
      issue_op(op, smem_base, ctl, cc);
    }
  } else {
    const uint32_t row = threadIdx.x;
    const int ld = k_hid + k_in;
    for (int c0 = 0; c0 < k_hid; c0 += 8) {
      float v[8];
      for (int i = 0; i < 8; ++i) v[i] = A[row * ld + c0 + i];
      store8_split(smem, OFF_HID_HI, OFF_HID_LO, true, row, c0, v);
    }
    for (int c = 0; c < 64; ++c) store_in(smem, row, c, c < k_in ? A[row * ld + k_hid + c] : 0.f);
    fence_proxy_async_smem();
    mbar_arrive(&ctl->a_ready);
    mbar_wait(&ctl->d_ready, 0);
    tc_fence_after_sync();
    float hv[16];
    epilogue_op(op, L.bias, smem, tmem_base, row, hv, out_f32, op.N);
    tc_fence_before_sync();
    if (op.out_kind != OUT_HEAD && out_readback) {
      const uint32_t off_hi = (op.out_kind == OUT_HIDDEN_LO_REGION) ? OFF_HID_LO : OFF_HID_HI;
      for (uint32_t c = 0; c < op.N; ++c) {
        const uint32_t o = (c >> 6) * KBLK + kblock_offset(row, c & 63u);
        float v = __half2float(*reinterpret_cast<__half*>(smem + off_hi + o));
        if (op.out_kind == OUT_HIDDEN) v += __half2float(*reinterpret_cast<__half*>(smem + OFF_HID_LO + o));
        out_readback[row * op.N + c] = v;
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------
// host: packing
// ---------------------------------------------------------------------------
struct KChunkMap { uint8_t src; int rows[64]; };   // W row feeding each of the 64 A columns (-1 = zero pad)

struct OpBuild {
  const HostDense* dense;             // may be null when W2 (fused heads) is used
  std::vector<KChunkMap> kcs;
  int N_logical;                      // real output columns
  int N;                              // padded (multiple of 16)
  int terms, relu, out_kind, glue;
  std::vector<float> W;               // [K_total][N_logical] gathered logical weights (row-major)
  std::vector<float> b;
};

struct Packed {
  std::vector<TcOp> ops;
  std::vector<uint8_t> stream;
  std::vector<float> bias;
};

static void pack_op(const OpBuild& ob, Packed& out) {
  TcOp op;
  memset(&op, 0, sizeof op);
  op.N = (uint16_t)ob.N;
  op.terms = (uint8_t)ob.terms;
  op.relu = (uint8_t)ob.relu;
  op.out_kind = (uint8_t)ob.out_kind;
  op.glue = (uint8_t)ob.glue;
  op.n_kc = (uint8_t)ob.kcs.size();
  const int max_rows = ob.terms == 3 ? 128 : 256;
  op.n_nc = (uint8_t)((ob.N + max_rows - 1) / max_rows);
  op.nc_rows = (uint16_t)(ob.N / op.n_nc);
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
  op.w_off = (uint32_t)out.stream.size();
  for (size_t kc = 0; kc < ob.kcs.size(); ++kc) {
    const KChunkMap& km = ob.kcs[kc];
    op.kc_src[kc] = km.src;
    int last = -1;
    for (int c = 0; c < 64; ++c) if (km.rows[c] >= 0) last = c;
    op.kc_steps[kc] = (uint8_t)std::max(1, (last + 16) / 16);
    for (int nc = 0; nc < op.n_nc; ++nc) {
      const size_t img = (size_t)op.nc_rows * 128;
      const size_t base = out.stream.size();
      out.stream.resize(base + img * (ob.terms == 3 ? 2 : 1), 0);
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
    }
  }
  out.ops.push_back(op);
}

static KChunkMap kc_hidden(int block, int row0, int width_avail) {
  KChunkMap k;
  k.src = (uint8_t)block;
  for (int c = 0; c < 64; ++c) k.rows[c] = (block * 64 + c < width_avail) ? row0 + block * 64 + c : -1;
  return k;
}
static KChunkMap kc_input(int row0, int in_dim) {
  KChunkMap k;
  k.src = SRC_IN;
  for (int c = 0; c < 64; ++c) k.rows[c] = c < in_dim ? row0 + c : -1;
  return k;
}

// hidden stack of a modules.MLP: layer l reads [h (width) | inputs (in_dim) at the skip layer]
static void build_mlp_ops(const HostMlp& m, int terms, Packed& out) {
  for (int l = 0; l < m.depth; ++l) {
    OpBuild ob;
    ob.N_logical = ob.N = m.width;
    ob.terms = terms; ob.relu = 1; ob.out_kind = OUT_HIDDEN; ob.glue = GLUE_NONE;
    ob.W = m.hidden[l].W; ob.b = m.hidden[l].b;
    if (l == 0) ob.kcs.push_back(kc_input(0, m.in_dim));
    else {
      for (int j = 0; j < (m.width + 63) / 64; ++j) ob.kcs.push_back(kc_hidden(j, 0, m.width));
      if (l == m.skip) ob.kcs.push_back(kc_input(m.width, m.in_dim));
    }
    pack_op(ob, out);
  }
}

static void build_head_op(const std::vector<const HostDense*>& heads, int width, int terms, int glue, Packed& out) {
  OpBuild ob;
  int n = 0;
  for (auto* h : heads) n += h->N;
  ob.N_logical = n;
  ob.N = 16;
  ob.terms = terms; ob.relu = 0; ob.out_kind = OUT_HEAD; ob.glue = glue;
  ob.W.assign((size_t)width * n, 0.f);
  int c0 = 0;
  for (auto* h : heads) {
    for (int k = 0; k < width; ++k) for (int j = 0; j < h->N; ++j) ob.W[(size_t)k * n + c0 + j] = h->W[(size_t)k * h->N + j];
    for (int j = 0; j < h->N; ++j) ob.b.push_back(h->b[j]);
    c0 += h->N;
  }
  for (int j = 0; j < (width + 63) / 64; ++j) ob.kcs.push_back(kc_hidden(j, 0, width));
  pack_op(ob, out);
}

struct TcEngine {
  Packed packed[2];
  int n_ops_sigma[2] = {0, 0};
  TcOp* d_ops[2] = {nullptr, nullptr};
  uint8_t* d_stream[2] = {nullptr, nullptr};
  float* d_bias[2] = {nullptr, nullptr};
};

std::string tc_engine_supports(const ndsr_config& c, int cc_major, int cc_minor) {
  if (cc_major != 10) return "needs an sm_100-class device (tcgen05)";
  if (c.trunk_width != 256 && c.trunk_width != 128 && c.trunk_width != 64) return "trunk width must be 64/128/256";
  if (c.rgb_depth != 1) return "rgb branch depth must be 1";
  if (c.rgb_width > 256 || c.rgb_width % 16) return "rgb width";
  if (c.use_warp && c.warp_width > 256) return "warp width";
  const int widths[] = {c.trunk_width, c.use_warp ? c.warp_width : 64, c.use_hyper_sheet ? c.hyper_sheet_width : 64,
                        c.use_predicted_mask ? c.mask_width : 64};
  for (int w : widths) if (w % 64) return "MLP widths must be multiples of 64";
  (void)cc_minor;
  return "";
}

static int build_level(ndsr_handle* h, int lv, Packed& P, int& n_sigma) {
  const ndsr_config& c = h->cfg;
  const HostModel& HM = h->host_model;
  const int prec = c.precision;
  const int t_sigma = prec == NDSR_PREC_FP16 ? 1 : 3;
  const int t_rgb = prec == NDSR_PREC_SPLIT3 ? 3 : 1;
  if (h->max_in > 64) { h->err = "tensor-core engine: MLP inputs wider than 64 features"; return NDSR_ERR_UNSUPPORTED; }
  if (h->dim_view + (c.predict_norm ? h->dim_norm : 0) > 64) { h->err = "tensor-core engine: rgb side inputs wider than 64"; return NDSR_ERR_UNSUPPORTED; }
  if (c.use_predicted_mask) {
    build_mlp_ops(HM.mask, t_sigma, P);
    build_head_op({&HM.mask.logit}, HM.mask.width, t_sigma, GLUE_MASK, P);
  }
  if (c.use_warp) {
    build_mlp_ops(HM.warp, t_sigma, P);
    build_head_op({&HM.warp_w, &HM.warp_v}, HM.warp.width, t_sigma, GLUE_WARP, P);
  }
  if (c.use_hyper_sheet) {
    build_mlp_ops(HM.hyper, t_sigma, P);
    build_head_op({&HM.hyper.logit}, HM.hyper.width, t_sigma, GLUE_HYPER, P);
  }
  build_mlp_ops(HM.trunk[lv], t_sigma, P);
  build_head_op({&HM.alpha[lv]}, HM.trunk[lv].width, t_sigma, GLUE_ALPHA, P);
  n_sigma = (int)P.ops.size();
  // ---- rgb branch (modules.py:288-313).  Flax input order:
  //   [bottleneck|trunk_out (W) | viewdir feats | trunk_out (App. C-1) | norm feats]
  const int W = c.trunk_width;
  const bool use_b = c.use_viewdirs != 0;
  if (use_b) {
    OpBuild ob;
    ob.N_logical = ob.N = W;
    ob.terms = t_rgb; ob.relu = 0; ob.out_kind = (t_rgb == 1) ? OUT_HIDDEN_LO_REGION : OUT_HIDDEN; ob.glue = GLUE_BOTTLENECK;
    ob.W = HM.bottleneck[lv].W; ob.b = HM.bottleneck[lv].b;
    for (int j = 0; j < W / 64; ++j) ob.kcs.push_back(kc_hidden(j, 0, W));
    if (t_rgb != 1) { h->err = "tensor-core engine: split3 precision on the rgb branch is not built (use mixed)"; return NDSR_ERR_UNSUPPORTED; }
    pack_op(ob, P);
  }
  {
    const HostMlp& R = HM.rgb[lv];
    OpBuild ob;
    ob.N_logical = ob.N = R.width;
    ob.terms = 1; ob.relu = 1; ob.out_kind = OUT_HIDDEN; ob.glue = GLUE_NONE;
    ob.W = R.hidden[0].W; ob.b = R.hidden[0].b;
    int row = 0;
    // first segment: bottleneck (stored in the HID_LO region as a 1-term operand) or trunk_out
    for (int j = 0; j < W / 64; ++j) {
      KChunkMap k = kc_hidden(j, 0, W);
      if (use_b) k.src = (uint8_t)(8 + j);
      ob.kcs.push_back(k);
    }
    row += W;
    const int v0 = row;
    row += h->dim_view;
    int x0 = -1;
    if (c.use_x_in_rgb_condition) { x0 = row; row += W; }
    const int n0 = row;
    const int ndim = c.predict_norm ? h->dim_norm : 0;
    if (x0 >= 0) {
      if (!use_b) { h->err = "tensor-core engine: use_x_in_rgb_condition without viewdirs is not built"; return NDSR_ERR_UNSUPPORTED; }
      for (int j = 0; j < W / 64; ++j) ob.kcs.push_back(kc_hidden(j, x0, W));
    }
    if (h->dim_view + ndim > 0) {
      KChunkMap k;
      k.src = SRC_IN;
      for (int cidx = 0; cidx < 64; ++cidx) {
        if (cidx < h->dim_view) k.rows[cidx] = v0 + cidx;
        else if (cidx < h->dim_view + ndim) k.rows[cidx] = n0 + (cidx - h->dim_view);
        else k.rows[cidx] = -1;
      }
      ob.kcs.push_back(k);
    }
    if ((int)ob.kcs.size() > MAX_KC) { h->err = "tensor-core engine: rgb input too wide"; return NDSR_ERR_UNSUPPORTED; }
    pack_op(ob, P);
    build_head_op({&R.logit}, R.width, 1, GLUE_RGB, P);
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
    if ((e = cudaMalloc(&E->d_ops[lv], P.ops.size() * sizeof(TcOp))) != cudaSuccess ||
        (e = cudaMalloc(&E->d_stream[lv], P.stream.size())) != cudaSuccess ||
        (e = cudaMalloc(&E->d_bias[lv], P.bias.size() * sizeof(float))) != cudaSuccess ||
        (e = cudaMemcpy(E->d_ops[lv], P.ops.data(), P.ops.size() * sizeof(TcOp), cudaMemcpyHostToDevice)) != cudaSuccess ||
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
    if (h->tc->d_ops[lv]) cudaFree(h->tc->d_ops[lv]);
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
  K.lvl.ops = E->d_ops[fa.level];
  K.lvl.n_ops = fa.sigma_only ? E->n_ops_sigma[fa.level] : (int)E->packed[fa.level].ops.size();
  K.lvl.weights = E->d_stream[fa.level];
  K.lvl.bias = E->d_bias[fa.level];
  K.warp_embed = h->M.warp_embed;
  K.mask_embed = h->M.mask_embed;
  const int64_t tiles = (fa.n_samples_total + TM - 1) / TM;
  if (tiles == 0) return NDSR_OK;
  const int grid = (int)(tiles < h->num_sms ? tiles : h->num_sms);
  field_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(K, cp, fa, h->cfg);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { h->err = std::string("field_tc_kernel launch: ") + cudaGetErrorString(e); return NDSR_ERR_CUDA; }
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
  if (k_hid % 64 || k_hid > 256 || k_in > 64 || k_in < 0 || n_out < 1 || n_out > 256 || (terms != 1 && terms != 3))
    return NDSR_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return NDSR_ERR_CUDA;
  OpBuild ob;
  ob.N_logical = n_out;
  ob.N = out_kind == OUT_HEAD ? 16 : ((n_out + 63) / 64) * 64;
  if (out_kind == OUT_HEAD && n_out > 16) return NDSR_ERR_INVALID;
  ob.terms = terms; ob.relu = relu; ob.out_kind = out_kind; ob.glue = GLUE_SELFTEST;
  const int K = k_hid + k_in;
  ob.W.assign(W, W + (size_t)K * n_out);
  ob.b.assign(bias, bias + n_out);
  for (int j = 0; j < k_hid / 64; ++j) ob.kcs.push_back(kc_hidden(j, 0, k_hid));
  if (k_in > 0) ob.kcs.push_back(kc_input(k_hid, k_in));
  Packed P;
  pack_op(ob, P);
  TcOp* d_ops; uint8_t* d_stream; float *d_bias, *d_A, *d_out, *d_rb;
  const int N = P.ops[0].N;
  cudaMalloc(&d_ops, sizeof(TcOp)); cudaMalloc(&d_stream, P.stream.size()); cudaMalloc(&d_bias, P.bias.size() * 4);
  cudaMalloc(&d_A, (size_t)TM * K * 4); cudaMalloc(&d_out, (size_t)TM * N * 4); cudaMalloc(&d_rb, (size_t)TM * N * 4);
  cudaMemcpy(d_ops, P.ops.data(), sizeof(TcOp), cudaMemcpyHostToDevice);
  cudaMemcpy(d_stream, P.stream.data(), P.stream.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(d_bias, P.bias.data(), P.bias.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_A, A, (size_t)TM * K * 4, cudaMemcpyHostToDevice);
  cudaMemset(d_out, 0, (size_t)TM * N * 4); cudaMemset(d_rb, 0, (size_t)TM * N * 4);
  cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES);
  TcLevel L; L.ops = d_ops; L.n_ops = 1; L.weights = d_stream; L.bias = d_bias;
  tc_selftest_kernel<<<1, TC_THREADS, TC_SMEM_BYTES>>>(L, d_A, k_hid, k_in, d_out, d_rb);
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
  cudaFree(d_ops); cudaFree(d_stream); cudaFree(d_bias); cudaFree(d_A); cudaFree(d_out); cudaFree(d_rb);
  return rc;
}
