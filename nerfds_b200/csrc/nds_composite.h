// Argument blocks and launchers of the per-ray kernels (nds_composite.cu).
#pragma once
#include <cuda_runtime.h>
#include "nds_common.cuh"

namespace nds {

struct CompositeArgs {
  int64_t n_rays;
  int S, H;
  const float* planes;
  int64_t plane_stride;
  uint32_t plane_mask;   // PlaneGroup bits: which planes the field kernel wrote for this call
  const float* z;        // [B,S]
  const float* dirs;     // [B,3]
  const float* viewdirs; // [B,3]
  const float* origins;  // [B,3]
  const float* points;   // [B,S,3] or null
  int sigma_is_activated, white_bkgd, sample_at_infinity;
  int has_norm, has_warp, has_mask, has_grad;
  ndsr_outputs out;
  float* argmax_idx;     // scratch [B] (as float) for sharpen_weights
  float* weights_sg;     // scratch [B,S] or null: cal_weights() weights for sharpen_weights
  // split fine pass (tensor-core engine): sample s of ray r is element e = src_elem[r S + s] of
  // concat(coarse depths, new depths); its planes sit at r n_carried + e (e < n_carried) or at
  // B n_carried + r (S - n_carried) + (e - n_carried).  null: planes are in sample order.
  const int32_t* src_elem;
  int n_carried;
  // filter_sigma (models.py:38-66): bit 0 dust threshold, bit 1 bounding box on the observation-space points
  int filter_flags;
  float dust_threshold, bbox[6];
  // multi-GPU frame reassembly without a collective: every per-ray result is also stored at the same address plus
  // mirror_delta[m] bytes -- the peer-mapped copies of the caller's frame buffer on the other GPUs (NVLink stores)
  int n_mirror;
  int64_t mirror_delta[NDSR_MAX_MIRRORS];
};

struct SamplePdfArgs {
  int64_t n_rays;
  int n_bins, n_fine, n_coarse;
  const float* bins;      // [B, n]      (null: midpoints of z_coarse)
  const float* weights;   // [B, n-1]    with row stride w_stride, offset applied by caller
  int64_t w_stride;
  const float* u;         // [B, n_fine] (null: linspace(0,1,n_fine))
  const float* z_coarse;  // [B, n_coarse]
  float* z_out;           // [B, n_coarse + n_fine]
  float* z_samples;       // optional [B, n_fine]
  int32_t* idx_lo;        // optional
  int32_t* idx_hi;        // optional
  float* cdf_out;         // optional [B, n]
  int32_t* src_elem_out;  // optional [B, n_coarse + n_fine]: which element of concat(z_coarse, z_samples) sits at sorted position s
};

// Early-termination scan of the split fine pass (render mode).  After the "carried" launch the fine network's sigma
// is known at every coarse depth of a ray; the new depths are then evaluated front to back in ROUNDS of equal rank
// ranges (rank = position among the ray's new depths in sorted order).  Before round r the scan walks the sorted
// union with every sample whose sigma is known -- carried samples and the new depths of earlier rounds, the others
// counted as empty -- which gives an upper bound T_ub >= T of the transmittance in front of each depth of the round:
// every factor (1 - alpha + 1e-10) the compositing product will apply (model_utils.py:131-136) is at most 1 + 1e-10.
// A depth with T_ub < eps can carry at most eps of weight and is not evaluated: its planes are written as an empty
// sample (sigma_raw -> softplus 0), the surviving depths of the round are compacted into `index`.
struct TerminationArgs {
  int64_t n_rays;
  int S, n_carried;          // sorted union length, coarse depths per ray
  const float* z;            // [B,S] sorted union
  const float* dirs;         // [B,3]
  const int32_t* src_elem;   // [B,S] (SamplePdfArgs::src_elem_out)
  float* planes;             // carried block [0, B n_carried), new block after it
  int64_t plane_stride;
  uint32_t plane_mask;
  int H, has_warp;
  int sample_at_infinity;
  float eps;
  int rank_lo, rank_hi;      // this round decides the new depths with rank in [rank_lo, rank_hi)
  int32_t* index;            // out: surviving new samples (element of the dense new-depth list, ray * n_new + e)
  int32_t* n_active;         // out: their count (zeroed by the launcher)
  unsigned long long* stats; // [2] accumulated: new depths evaluated, new depths seen
  // Adaptive rounds: round_stats[r] = {evaluated, seen} of round r of THIS fine level (zeroed before round 0).  A round
  // r >= 1 that finds (almost) nothing was skipped so far -- evaluated >= merge_frac x seen over the rounds before it --
  // takes all remaining ranks at once, and the rounds after it are empty: on a field with nothing to terminate the
  // fine level costs one extra launch instead of rounds - 1.  Every kernel recomputes the decision from the counters
  // of the rounds that have finished, so no flag travels through the host.
  unsigned long long* round_stats;   // [NDS_TERM_MAX_ROUNDS][2]
  int round;
  float merge_frac;
};
constexpr int NDS_TERM_MAX_ROUNDS = 64;
cudaError_t launch_termination_scan(const TerminationArgs& a, int num_sms, cudaStream_t st);

cudaError_t launch_uniform_threefry(uint32_t k0, uint32_t k1, int64_t n, float* out, int num_sms, cudaStream_t st);
cudaError_t launch_uniform_threefry_range(uint32_t k0, uint32_t k1, int64_t n, int64_t first, int64_t count, float* out,
                                          int num_sms, cudaStream_t st);
cudaError_t launch_camera_rays(const ndsr_camera& cam, float* origins, float* dirs, float* pixels, cudaStream_t st);
cudaError_t launch_sample_along_rays(int64_t n_rays, int S, float near_, float far_, int lindisp,
                                     const float* t_rand, float* z, cudaStream_t st);
cudaError_t launch_composite(const CompositeArgs& a, int num_sms, cudaStream_t st);
cudaError_t launch_sharpen(int64_t n_rays, int S, const float* wsg, const float* z, const float* argmax_idx,
                           float stdv, float* out, int num_sms, cudaStream_t st);
cudaError_t launch_pack_rgb_sigma(int64_t total, const float* rgb, const float* sigma, float* planes, int64_t ps,
                                  cudaStream_t st);
cudaError_t launch_sample_pdf(const SamplePdfArgs& a, int num_sms, cudaStream_t st);

}  // namespace nds
