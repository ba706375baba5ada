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
