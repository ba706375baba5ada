// Shared device-side declarations of the NeRF-DS ray-marching kernels.
// File:line citations refer to /root/reference (JokerYan/NeRF-DS).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "nerfds_b200.h"

namespace nds {

// ---- parameter views (fp32, Flax layout: W is [K=in][N=out] row-major) ----
struct DenseW {
  const float* W;
  const float* b;
  int K, N;
};
struct MlpW {                       // modules.MLP (modules.py:44-83)
  DenseW hidden[NDSR_MAX_DEPTH];
  DenseW logit;                     // N == 0 when absent
  int depth, width, in_dim, skip;   // skip: layer index with [h | inputs] concat, -1 none
};
struct LevelW {                     // modules.NerfMLP (modules.py:86-152)
  MlpW trunk;
  DenseW bottleneck;
  DenseW alpha;                     // [width, 1 (+3 normals)]
  MlpW rgb;
};
// transposed copies for the input-gradient sweep
struct MlpWT {
  const float* WTh[NDSR_MAX_DEPTH]; // [width][Kh_pad]  d(hidden_l)/d(h_{l-1})^T, l>=1
  const float* WTx[NDSR_MAX_DEPTH]; // [width][Kx_pad]  input slice (layer 0 and the skip layer)
  const float* WTlogit;             // [n_out][width]
  int kx_pad;                       // in_dim rounded up to 32
};
struct ModelW {
  MlpW mask, warp, hyper;
  DenseW warp_w, warp_v;            // SE3Field branches (warping.py:174-193)
  LevelW level[2];
  MlpWT warp_T, hyper_T, trunk_T[2];
  const float* alpha_col0[2];       // column 0 of alpha kernel, contiguous [width]
  const float* warp_embed;          // [ids, warp_embed_dims]
  const float* mask_embed;          // [ids, mask_embed_dims]
};

// ---- positional encoding spec (model_utils.py:398-436) ---------------------
struct PosencSpec {
  int min_deg, num_bands, identity;
  float window[NDSR_MAX_BANDS];     // 1.0 when the call passes alpha=None
  __host__ __device__ int dim(int C) const { return 2 * num_bands * C + (identity ? C : 0); }
};

struct CallParams {
  PosencSpec pe_mask, pe_warp, pe_hsheet, pe_spatial, pe_hyperpt, pe_view, pe_norm;
  float mask_ratio;
  int use_predicted_norm, use_sigma_gradient;
};

// ---- per-sample planes written by the field kernels ------------------------
// planes[c * plane_stride + n], n = ray * S + sample.
enum Plane {
  P_SIGMA_RAW = 0,
  P_RGB = 1,        // 3, after sigmoid
  P_NORM = 4,       // 3, raw predicted normal (alpha_mlp cols 1..3)
  P_MASK = 7,       // predicted mask (after output relu)
  P_WARPED = 8,     // 3 + H (H <= 2) warped spatial + hyper coords
  P_ROT = 13,       // 3, normalize(R (1,1,1)/sqrt3)
  P_TRANS = 16,     // 3, translation p
  P_GRAD = 19,      // 3, normalize(-d sigma_raw / d x)
  P_TNORM = 22,     // 3, normalize(R grad)
  P_COUNT = 25
};

// which groups of planes a call needs (FieldArgs::plane_mask / CompositeArgs::plane_mask): the field kernel only
// writes, and the compositing kernel only reads, the planes behind the requested level-dict keys
enum PlaneGroup : uint32_t { PG_RGB = 1, PG_NORM = 2, PG_MASK = 4, PG_WARPED = 8, PG_ROT = 16, PG_TRANS = 32, PG_ALL = 63 };

struct FieldArgs {
  int64_t n_samples_total;   // B * S
  int S;
  int level;
  const float* points;       // [N,3] or null
  const float* z;            // [N]
  const float* origins;      // [B,3]
  const float* dirs;         // [B,3]
  const float* viewdirs;     // [B,3] (never null: host substitutes dirs)
  const uint32_t* warp_id;   // [B] or null
  const float* gt_mask;      // [B] or null
  float* planes;
  int64_t plane_stride;
  uint32_t plane_mask;       // PlaneGroup bits (sigma_raw is always written)
  int sigma_only;            // skip bottleneck/rgb/normal-input work
  int need_grad;             // compute P_GRAD / P_TNORM
  // ---- tensor-core engine only: fine pass split into "new" and "carried" samples --------------------------
  // The warp field, hyper-sheet and mask networks are shared by the two levels and the fine level re-visits every
  // coarse depth (model_utils.py:266: sort(concat(z_coarse, z_samples))), so for those samples the coarse pass
  // already computed the warped point, the hyper coordinates, the mask and the SE(3) transform.  The coarse pass
  // stores them (`carry_out`), and the fine pass runs as two launches over DENSE sample lists -- the S_f new
  // depths of every ray through the whole network chain, the S_c carried ones through the template NeRF only --
  // each writing its own dense block of the planes; composite_kernel gathers the sorted order back
  // (CompositeArgs::src_elem, written by sample_pdf_kernel).
  const float* carry;        // non-null: "carried" launch, reads C_* planes [c * carry_stride + i]
  float* carry_out;          // non-null: also store the C_* planes of every sample
  int64_t carry_stride;
  // ---- early termination (TerminationArgs, nds_composite.h): the launch evaluates only the *n_active samples
  // listed in `index` (elements of the dense sample list; both written by termination_scan_kernel earlier on the
  // stream).  null: all n_samples_total samples in order.
  const int32_t* index;
  const int32_t* n_active;
};

// carried per-sample planes (coarse pass -> fine pass)
enum CarryPlane { C_WARPED = 0 /* 3 + 2 */, C_MASK = 5, C_R = 6 /* 9 */, C_P = 15 /* 3 */, C_COUNT = 18 };

// ---- small exact-math helpers ----------------------------------------------
#define NDS_HALF_PI_F 1.57079637050628662109375f   // fp32(0.5 * pi), model_utils.py:405
#define NDS_EPS_F32 1.1920928955078125e-07f        // jnp.finfo(float32).eps, model_utils.py:439

__device__ __forceinline__ void normalize3(const float v[3], float out[3]) {
  // model_utils.normalize_vector (model_utils.py:438-442)
  float s = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
  float d = sqrtf(fmaxf(s, NDS_EPS_F32));
  out[0] = v[0] / d; out[1] = v[1] / d; out[2] = v[2] / d;
}

// posenc of a C-vector into dst[stride-1 contiguous]; returns number written.
// Layout (F, 2, C) flattened, identity first (model_utils.py:398-417).
template <typename Store>
__device__ __forceinline__ int posenc_emit(const float* x, int C, const PosencSpec& pe, Store store, int o) {
  if (pe.identity) for (int c = 0; c < C; ++c) store(o++, x[c]);
  for (int k = 0; k < pe.num_bands; ++k) {
    const float s = exp2f((float)(pe.min_deg + k));
    const float w = pe.window[k];
    for (int c = 0; c < C; ++c) {
      const float xb = x[c] * s;
      store(o + c, w * sinf(xb));
      store(o + C + c, w * sinf(xb + NDS_HALF_PI_F));
    }
    o += 2 * C;
  }
  return o;
}

// SE(3) exponential exactly as rigid_body.py:59-101 evaluates it after the
// w/theta, v/theta normalisation of warping.py:219-221 (no epsilon: theta == 0
// gives NaN like the reference).
template <typename T>
struct SE3 { T R[9]; T p[3]; };

__device__ __forceinline__ float nsqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ float nsin(float x) { return sinf(x); }
__device__ __forceinline__ float ncos(float x) { return cosf(x); }

template <typename T>
__device__ __forceinline__ void exp_se3(const T w_raw[3], const T v_raw[3], SE3<T>& out) {
  T theta = nsqrt(w_raw[0] * w_raw[0] + w_raw[1] * w_raw[1] + w_raw[2] * w_raw[2]);
  T w[3] = {w_raw[0] / theta, w_raw[1] / theta, w_raw[2] / theta};
  T v[3] = {v_raw[0] / theta, v_raw[1] / theta, v_raw[2] / theta};
  T zero = T(0.f);
  T W[9] = {zero, -w[2], w[1], w[2], zero, -w[0], -w[1], w[0], zero};   // skew()
  T WW[9];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      WW[i * 3 + j] = W[i * 3 + 0] * W[0 * 3 + j] + W[i * 3 + 1] * W[1 * 3 + j] + W[i * 3 + 2] * W[2 * 3 + j];
  T st = nsin(theta), ct = ncos(theta);
  T omc = T(1.f) - ct, tms = theta - st;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    T acc = zero;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      T eye = (i == j) ? T(1.f) : zero;
      out.R[i * 3 + j] = eye + st * W[i * 3 + j] + omc * WW[i * 3 + j];
      T m = ((i == j) ? theta : zero) + omc * W[i * 3 + j] + tms * WW[i * 3 + j];
      acc = acc + m * v[j];
    }
    out.p[i] = acc;
  }
}

}  // namespace nds
