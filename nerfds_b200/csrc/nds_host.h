// Host-side declarations shared by the translation units of libnerfds_b200.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "nds_common.cuh"

namespace nds {

// fp32 host copies of the Dense layers (Flax layout), kept so the tensor-core
// engine can repack them into its split-fp16 shared-memory images.
struct HostDense {
  std::vector<float> W, b;
  int K = 0, N = 0;
};
struct HostMlp {
  std::vector<HostDense> hidden;
  HostDense logit;
  int depth = 0, width = 0, in_dim = 0, skip = -1;
};
struct HostModel {
  HostMlp mask, warp, hyper, trunk[2], rgb[2];
  HostDense warp_w, warp_v, bottleneck[2], alpha[2];
  std::vector<float> warp_embed, mask_embed;
};

struct TcEngine;   // opaque, nds_field_tc.cu

}  // namespace nds

struct ndsr_handle {
  ndsr_config cfg;
  int device = 0, num_sms = 0, cc_major = 0;
  int engine = NDSR_ENGINE_SIMT;
  bool loaded = false;
  std::string err;
  int64_t launches = 0;
  // derived dims
  int H = 0, dim_mask_in = 0, dim_warp_in = 0, dim_hyper_in = 0, dim_trunk_in = 0, dim_view = 0, dim_norm = 0;
  int max_in = 0, max_w = 0, rgb_in_full = 0;
  // parameters
  float* arena = nullptr;
  nds::ModelW M;
  nds::HostModel host_model;
  nds::TcEngine* tc = nullptr;
  // scratch (grown on demand)
  std::vector<void*> scratch_allocs;
  int64_t cap_rays = 0, cap_samples = 0, max_samples_seen = 0;
  int64_t max_chunk = 65536;
  bool tc_no_split = false;    // NDS_TC_NO_SPLIT (diagnostics), read once at ndsr_create
  int n_mirror = 0;                        // peer copies of the caller's frame buffer (ndsr_set_output_mirrors)
  int64_t mirror_delta[NDSR_MAX_MIRRORS] = {0};
  const char* mirror_base = nullptr;       // this rank's own frame buffer: only stores inside it are mirrored
  size_t mirror_bytes = 0;
  float* carry = nullptr;      // C_COUNT planes of the coarse samples (tensor-core engine: split fine pass)
  int32_t* src_elem = nullptr; // [rays, S_c + S_f] from sample_pdf: element of concat(coarse, new) at each sorted position
  float* z_new = nullptr;      // [rays, S_f] the new depths in draw order
  // early termination of the fine level (ndsr_set_early_termination): survivors of the scan, their count, statistics
  float term_eps = 0.f;
  int term_rounds = 4;
  int32_t* term_index = nullptr;
  int32_t* term_count = nullptr;
  unsigned long long* term_round_stats = nullptr;   // device [NDS_TERM_MAX_ROUNDS][2], per fine level (adaptive rounds)
  unsigned long long* term_stats = nullptr;   // device [2]: new depths evaluated / seen since the last reset
  unsigned long long term_stats_keep[2] = {0, 0};   // their value while the scratch is being regrown
  float *planes = nullptr, *z_coarse = nullptr, *z_fine = nullptr, *w_coarse = nullptr, *w_sg = nullptr,
        *argmax = nullptr;
  void* in_stage = nullptr;       // host-buffer calls: two input staging halves (double-buffered per ray chunk)
  size_t in_stage_bytes = 0;
  void* out_stage = nullptr;      // ... and the device copies of the requested outputs
  size_t out_stage_bytes = 0;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
  // per-stage device timing (ndsr_profile_enable / ndsr_profile_read)
  bool prof = false;
  struct ProfSpan { int stage; cudaEvent_t a, b; };
  std::vector<ProfSpan> prof_spans;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[NDSR_STAGE_COUNT] = {0, 0, 0, 0, 0, 0};
  int64_t prof_launches[NDSR_STAGE_COUNT] = {0, 0, 0, 0, 0, 0};
};

namespace nds {

void make_call_params(const ndsr_config& c, const ndsr_extra_params& ep, CallParams& cp);

// nds_field_simt.cu
size_t field_simt_smem_bytes(const ndsr_config& c, int max_in, int max_w, int x2, bool grad, int* ld);
cudaError_t launch_field_simt(const ModelW& M, const CallParams& cp, const FieldArgs& a, const ndsr_config& cfg,
                              int max_in, int max_w, int x2, int num_sms, cudaStream_t stream);

// nds_field_tc.cu
std::string tc_engine_supports(const ndsr_config& c, int cc_major, int cc_minor);
int tc_engine_load(ndsr_handle* h);
int tc_engine_field(ndsr_handle* h, const CallParams& cp, const FieldArgs& fa, cudaStream_t st);
void tc_engine_free(ndsr_handle* h);

}  // namespace nds

// nds_composite.cu (declared here with their arg structs)
#include "nds_composite.h"
