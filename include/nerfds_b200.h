/* nerfds_b200 -- C ABI of the B200-native NeRF-DS ray-marching path.
 *
 * The reference (JokerYan/NeRF-DS, JAX/Flax) has no FFI: its "operator API"
 * for this path is two Python call signatures.  Each entry point below names
 * the reference interface it stands in for (file:line under /root/reference);
 * INTEGRATION.md shows the ctypes binding a maintainer of the reference would
 * add.  Plain pointers and sizes only, no torch types.  Every function
 * returns 0 on success or a negative ndsr_status; nothing throws across the
 * boundary.  Calls are stream-ordered and asynchronous unless stated; a
 * handle is bound to one CUDA device and admits one in-flight call
 * (multi-GPU = one handle per rank / process).
 */
#ifndef NERFDS_B200_H_
#define NERFDS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDSR_ABI_VERSION 2
#define NDSR_MAX_DEPTH 8      /* hidden layers per MLP */
#define NDSR_MAX_BANDS 12     /* posenc frequency bands per encoder */

typedef enum ndsr_status {
  NDSR_OK = 0,
  NDSR_ERR_INVALID = -1,      /* bad argument / unsupported configuration */
  NDSR_ERR_CUDA = -2,         /* CUDA runtime error, see ndsr_last_error */
  NDSR_ERR_PARAMS = -3,       /* missing / mis-shaped parameter tensor */
  NDSR_ERR_NOT_LOADED = -4,   /* render before ndsr_load_params */
  NDSR_ERR_UNSUPPORTED = -5   /* valid reference config that this engine does not build */
} ndsr_status;

/* Which engine evaluates the MLPs. */
typedef enum ndsr_engine {
  NDSR_ENGINE_AUTO = 0,   /* tensor-core engine when the config fits it, else SIMT */
  NDSR_ENGINE_SIMT = 1,   /* fp32 CUDA-core engine: every output incl. d(sigma)/dx */
  NDSR_ENGINE_TC = 2      /* tcgen05 engine (sm_100a), split-fp16 operands, fp32 accumulate */
} ndsr_engine;

/* Tensor-core operand precision (NDSR_ENGINE_TC only). */
typedef enum ndsr_precision {
  NDSR_PREC_MIXED = 0,    /* 3-term split fp16 on the sigma path (mask, warp, hyper-sheet,
                             trunk, sigma/normal head), 1-term fp16 on the rgb branch
                             (median rgb error ~2e-4) */
  NDSR_PREC_FP16 = 1,     /* 1-term fp16 everywhere (fast; misses 1e-3 RGB) */
  NDSR_PREC_SPLIT3 = 2    /* 3-term split everywhere (default of the Python layer) */
} ndsr_precision;

/* Mirrors the attributes of NerfModel (hypernerf/models.py:116-229), SE3Field
 * (warping.py:139-157), HyperSheetMLP (modules.py:354-365), MaskMLP
 * (modules.py:396-407) that the path reads.  `size` = sizeof(ndsr_config). */
typedef struct ndsr_config {
  uint32_t size;
  uint32_t abi_version;
  float near_, far_;
  int32_t num_warp_embeds;
  /* template MLP */
  int32_t use_viewdirs;
  int32_t trunk_depth, trunk_width, trunk_skip;          /* skip layer index or -1 */
  int32_t rgb_depth, rgb_width;
  /* sampling */
  int32_t num_coarse_samples, num_fine_samples;
  int32_t use_stratified_sampling, use_white_background;
  int32_t use_linear_disparity, use_sample_at_infinity;
  /* positional encodings */
  int32_t spatial_min_deg, spatial_max_deg;
  int32_t hyper_point_min_deg, hyper_point_max_deg;
  int32_t viewdir_min_deg, viewdir_max_deg;
  int32_t use_posenc_identity;
  /* hyper sheet */
  int32_t use_hyper_sheet, hyper_num_dims;
  int32_t hyper_sheet_min_deg, hyper_sheet_max_deg;
  int32_t hyper_sheet_depth, hyper_sheet_width, hyper_sheet_skip;
  /* SE3 warp field */
  int32_t use_warp, warp_embed_dims;
  int32_t warp_min_deg, warp_max_deg, warp_use_posenc_identity;
  int32_t warp_depth, warp_width, warp_skip;
  /* normal branch */
  int32_t predict_norm, norm_input_posenc;
  int32_t norm_input_min_deg, norm_input_max_deg;
  int32_t use_x_in_rgb_condition;
  /* predicted mask */
  int32_t use_mask_in_warp, use_mask_in_hyper, use_predicted_mask;
  int32_t use_mask_sharp_weights;
  int32_t mask_embed_dims, mask_min_deg, mask_max_deg;
  int32_t mask_depth, mask_width, mask_skip, mask_output_relu;
  /* engine selection */
  int32_t engine;      /* ndsr_engine */
  int32_t precision;   /* ndsr_precision */
} ndsr_config;

/* One named fp32 parameter in Flax layout: Dense kernels
 * are [rows=in, cols=out] row-major, biases [1, out], embeddings [ids, dims].
 * `data` is a HOST pointer; ndsr_load_params copies / repacks it. */
typedef struct ndsr_tensor {
  const char* name;     /* e.g. "warp_field/trunk/hidden_0/kernel" */
  const float* data;
  int64_t rows, cols;
} ndsr_tensor;

/* TrainState.extra_params (hypernerf/model_utils.py:41-52) + the scalar
 * keyword arguments of NerfModel.__call__ (models.py:1436-1441). */
typedef struct ndsr_extra_params {
  float nerf_alpha, warp_alpha, hyper_alpha, hyper_sheet_alpha;
  float norm_input_alpha;
  float mask_ratio;
  float sharp_weights_std;
  float near_override, far_override;     /* NaN = use the configured value */
  int32_t use_predicted_norm;            /* models.py:1438 */
  int32_t use_sigma_gradient;            /* models.py:1437 */
  int32_t sample_at_infinity_override;   /* -1 = configured, 0/1 = override (fine level only,
                                            models.py:1509 vs 1544) */
  /* render_opts of filter_sigma (models.py:38-66), applied where the reference applies it: ndsr_render_rays on the
   * FINE level only (models.py:1545; the coarse call gets no render_opts), ndsr_render_samples on the level it is
   * called for.  Bit 0: dust_threshold present, bit 1: bounding_box present.  The compositing density is
   * [sigma >= dust] * [x in box] * sigma on the ACTIVATED sigma and the observation-space point (models.py:1288);
   * `sharp_weights` filters the RAW sigma before the activation (models.py:1236-1237); the `sigma` output stays
   * unfiltered (models.py:1271). */
  int32_t filter_flags;
  float dust_threshold;
  float bounding_box[6];                 /* xmin, xmax, ymin, ymax, zmin, zmax (models.py:60) */
} ndsr_extra_params;

/* Output buffers of one level dict.  Every member is a
 * caller-allocated, contiguous, row-major fp32 buffer or NULL; NULL = "not
 * requested" (this is the output mask: work that only feeds NULL outputs is
 * skipped, e.g. d(sigma)/dx when target_norm is NULL and predict_norm is
 * set).  B = rays, S = samples of the level, H = hyper_num_dims (0 without a
 * hyper sheet).  Device pointers for ndsr_render_rays / ndsr_render_samples,
 * host pointers for ndsr_render_rays_host. */
typedef struct ndsr_outputs {
  /* per ray */
  float* rgb;                    /* [B,3]   model_utils.py:140 */
  float* depth;                  /* [B]     model_utils.py:141 */
  float* med_depth;              /* [B]     model_utils.py:142 */
  float* acc;                    /* [B]     model_utils.py:143-148 */
  float* ray_norm;               /* [B,3]   models.py:1350-1354 */
  float* ray_rotation_field;     /* [B,3]   models.py:1356 */
  float* ray_translation_field;  /* [B,3]   models.py:1359 */
  float* ray_delta_x;            /* [B,3]   models.py:1364 */
  float* ray_hyper_points;       /* [B,H]   models.py:1370 */
  float* ray_predicted_mask;     /* [B]     models.py:1399 */
  float* med_points;             /* [B,3+H] models.py:1411-1415 */
  /* per sample */
  float* z_vals;                 /* [B,S]   (not a reference key; the sample depths used) */
  float* weights;                /* [B,S]   model_utils.py:135 */
  float* alpha;                  /* [B,S]   model_utils.py:129 */
  float* accum_prod;             /* [B,S]   model_utils.py:131-134 */
  float* sigma;                  /* [B,S]   models.py:1271 (after softplus) */
  float* sharp_weights;          /* [B,S]   models.py:1245-1246 */
  float* back_facing;            /* [B,S]   models.py:1341-1344 */
  float* predicted_mask;         /* [B,S]   models.py:968 */
  float* points;                 /* [B,S,3] models.py:893 */
  float* warped_points;          /* [B,S,3+H] models.py:1311 */
  float* delta_x;                /* [B,S,3] models.py:1363-1365 */
  float* predicted_norm;         /* [B,S,3] models.py:1326 */
  float* target_norm;            /* [B,S,3] models.py:1327-1338 (needs d(sigma)/dx) */
} ndsr_outputs;

typedef struct ndsr_handle ndsr_handle;

/* Replaces models.construct_nerf / NerfModel.setup (models.py:2677-2741,
 * 324-391): validates the configuration and binds a handle to `device`. */
int ndsr_create(const ndsr_config* cfg, int device, ndsr_handle** out);
void ndsr_destroy(ndsr_handle* h);
const char* ndsr_last_error(const ndsr_handle* h);   /* h may be NULL: last create error */

/* Replaces passing `{'params': P}` to model.apply (render.py:140,
 * training.py:441): n named host tensors in Flax layout.  Synchronous. */
int ndsr_load_params(ndsr_handle* h, const ndsr_tensor* tensors, int n);

/* Replaces NerfModel.__call__ (models.py:1419-1565): coarse stratified
 * sampling -> render_samples('coarse') -> sample_pdf -> render_samples('fine').
 * All pointers are DEVICE pointers on the handle's device.
 *   origins, directions [n,3]; viewdirs [n,3] or NULL (= directions, 1475-1478)
 *   warp_id [n] (metadata['warp'], required when use_warp)
 *   gt_mask [n] or NULL (rays_dict['mask']; required unless mask_ratio == 1)
 *   t_rand [n,S_c], u [n,S_f]: the uniform draws of model_utils.py:84,217;
 *     required when use_stratified_sampling, ignored otherwise.
 *   coarse / fine: output masks+buffers; either may be NULL.
 * `stream` is a cudaStream_t passed as void*. */
int ndsr_render_rays(ndsr_handle* h, void* stream, int64_t n_rays,
                     const float* origins, const float* directions,
                     const float* viewdirs, const uint32_t* warp_id,
                     const float* gt_mask, const float* t_rand, const float* u,
                     const ndsr_extra_params* ep,
                     const ndsr_outputs* coarse, const ndsr_outputs* fine);

/* Same contract with HOST buffers for every input and output: copies inputs
 * host->device, renders, copies the requested outputs device->host on
 * `stream`, and synchronises the stream before returning.  This is the call
 * evaluation.render_image's per-chunk round trip maps to
 * (evaluation.py:119-130: model_fn + device_put(..., cpu)). */
int ndsr_render_rays_host(ndsr_handle* h, void* stream, int64_t n_rays,
                          const float* origins, const float* directions,
                          const float* viewdirs, const uint32_t* warp_id,
                          const float* gt_mask, const float* t_rand, const float* u,
                          const ndsr_extra_params* ep,
                          const ndsr_outputs* coarse, const ndsr_outputs* fine);

/* ndsr_render_rays_host with the two uniform draws of the path generated ON THE DEVICE from jax keys, as the
 * reference does inside its jitted call (model_utils.py:84 `random.uniform(key, [n_rays, S_c])`, :217
 * `[n_rays, S_f]`): t_rand = uniform(key_coarse, [n_rays, S_c]), u = uniform(key_fine, [n_rays, S_f]) over the WHOLE
 * call's rays (threefry2x32, bit for bit what ndsr_random_uniform writes), whatever the internal ray chunking.  The
 * host then only supplies the rays: 28 B/ray instead of 28 + 4 (S_c + S_f). */
int ndsr_render_rays_host_rng(ndsr_handle* h, void* stream, int64_t n_rays,
                              const float* origins, const float* directions,
                              const float* viewdirs, const uint32_t* warp_id,
                              const float* gt_mask, const uint32_t key_coarse[2], const uint32_t key_fine[2],
                              const ndsr_extra_params* ep,
                              const ndsr_outputs* coarse, const ndsr_outputs* fine);

/* Replaces NerfModel.render_samples (models.py:867-1417): one level on
 * caller-provided samples.  level 0 = 'coarse', 1 = 'fine' (selects the
 * NerfMLP).  points [n,S,3] may be NULL (= origins + z_vals * directions). */
int ndsr_render_samples(ndsr_handle* h, void* stream, int level,
                        int64_t n_rays, int32_t n_samples,
                        const float* points, const float* z_vals,
                        const float* origins, const float* directions,
                        const float* viewdirs, const uint32_t* warp_id,
                        const float* gt_mask, const ndsr_extra_params* ep,
                        int32_t use_sample_at_infinity,
                        const ndsr_outputs* out);

/* Replaces model_utils.sample_along_rays (model_utils.py:55-92).
 * z_vals [n,S] out; t_rand [n,S] or NULL (non-stratified). */
int ndsr_sample_along_rays(ndsr_handle* h, void* stream, int64_t n_rays,
                           int32_t n_samples, float near_, float far_,
                           int32_t use_linear_disparity,
                           const float* t_rand, float* z_vals);

/* Replaces model_utils.sample_pdf / piecewise_constant_pdf
 * (model_utils.py:193-269).  bins [n,nb], weights [n,nb-1], u [n,nf],
 * z_vals [n,nc] -> z_out [n,nc+nf] sorted.  Optional diagnostics:
 * z_samples [n,nf] (pre-sort), idx_lo/idx_hi [n,nf] = positions in bins/cdf
 * selected by the inverse CDF (the bit-exact contract), cdf [n,nb]. */
int ndsr_sample_pdf(ndsr_handle* h, void* stream, int64_t n_rays,
                    int32_t n_bins, int32_t n_fine, int32_t n_coarse,
                    const float* bins, const float* weights, const float* u,
                    const float* z_vals, float* z_out, float* z_samples,
                    int32_t* idx_lo, int32_t* idx_hi, float* cdf);

/* A pinhole camera with radial / tangential distortion, as hypernerf/camera.py:112-138 stores it (float32). */
typedef struct ndsr_camera {
  float orientation[9];          /* world-to-camera rotation, row-major (camera.py:127) */
  float position[3];
  float focal_length;
  float principal_point[2];
  float skew;
  float pixel_aspect_ratio;
  float radial_distortion[3];    /* k1, k2, k3 */
  float tangential_distortion[2];/* p1, p2 */
  int32_t image_size[2];         /* (width, height), camera.py:136 */
} ndsr_camera;

/* Replaces datasets/core.py:51-76 camera_to_rays = Camera.pixels_to_rays(Camera.get_pixel_centers())
 * (camera.py:226-270, 364-368; the 10 Newton iterations of _radial_and_tangential_undistort, camera.py:27-106) on the
 * device: origins [H*W,3] (the camera position), directions [H*W,3] (unit, world frame), pixels [H*W,2] (pixel
 * centres; nullable), row-major over (y, x).  Stateless: `device` is the CUDA device index. */
int ndsr_camera_rays(int device, void* stream, const ndsr_camera* camera, float* origins, float* directions,
                     float* pixels);

/* ---- multi-GPU frame reassembly over peer memory (replaces `jax.lax.all_gather(out, 'batch')`, render.py:155) ----
 * One process per GPU.  Every rank allocates the same packed frame buffer with ndsr_peer_alloc, the 64-byte handles
 * are exchanged by the host (any transport), every rank maps the others' buffers with ndsr_peer_open and registers
 * the address differences with ndsr_set_output_mirrors.  From then on the compositing kernel of a fine-level
 * ndsr_render_rays call whose per-ray output pointers ALL lie inside the registered frame buffer
 * [frame_base, frame_base + frame_bytes) stores every PER-RAY result both at the output pointer it was given and at
 * the same offset of every mirror: when all ranks' streams have drained, every GPU holds the whole frame -- the
 * all-gather happened inside the kernel, as NVLink stores.  Calls whose per-ray outputs all lie OUTSIDE the frame
 * buffer (ordinary tensors, ndsr_render_samples, ndsr_render_rays_host staging) are not mirrored; a call with
 * pointers on both sides fails with NDSR_ERR_INVALID. */
#define NDSR_MAX_MIRRORS 15
typedef struct ndsr_ipc_handle { unsigned char bytes[64]; } ndsr_ipc_handle;
int ndsr_peer_alloc(int device, size_t bytes, void** ptr, ndsr_ipc_handle* handle);
int ndsr_peer_free(int device, void* ptr);
int ndsr_peer_open(int device, const ndsr_ipc_handle* handle, void** ptr);
int ndsr_peer_close(int device, void* ptr);
/* byte_deltas[m] = (mapped address of mirror m) - (address of this rank's own buffer); frame_base / frame_bytes =
 * this rank's own buffer (the only address range whose stores are mirrored); n = 0 switches mirroring off */
int ndsr_set_output_mirrors(ndsr_handle* h, int32_t n, const int64_t* byte_deltas, const void* frame_base,
                            size_t frame_bytes);

/* Replaces the two `random.uniform(key, [n_rays, n_samples])` draws of the path (model_utils.py:84 stratified
 * jitter `t_rand`, model_utils.py:217 inverse-CDF `u`) on the device, bit for bit as jax 0.3.15's default
 * threefry2x32 generator produces them: out[i] for i in [0, n), n = n_rays * n_samples, row-major.  `key` is the
 * raw uint32[2] jax key AFTER flax's make_rng folding (models.py:1489, 1524; nerfds_b200/jax_random.py derives it
 * on the host).  Feed the result to ndsr_render_rays as t_rand / u.  Stateless; n < 2^32 - 1. */
int ndsr_random_uniform(int device, void* stream, const uint32_t key[2], int64_t n, float* out);
/* Elements [first, first + count) of the same n-draw stream (a block of rows of the [n_rays, n_samples] array). */
int ndsr_random_uniform_range(int device, void* stream, const uint32_t key[2], int64_t n, int64_t first, int64_t count,
                              float* out);

/* Replaces model_utils.volumetric_rendering (model_utils.py:95-159) +
 * compute_depth_map.  rgb [n,S,3], sigma [n,S], z_vals [n,S], dirs [n,3]. */
int ndsr_volumetric_rendering(ndsr_handle* h, void* stream, int64_t n_rays,
                              int32_t n_samples, const float* rgb,
                              const float* sigma, const float* z_vals,
                              const float* dirs, int32_t use_white_background,
                              int32_t sample_at_infinity,
                              const ndsr_outputs* out);

/* Introspection used by bench.py / tests. */
int ndsr_engine_in_use(const ndsr_handle* h);             /* ndsr_engine actually selected */
int64_t ndsr_kernel_launches(const ndsr_handle* h);       /* kernels launched so far */
int ndsr_abi_version(void);
/* Tensor-core engine: MACs the layer programs ISSUE per sample evaluation (split-fp16 terms and padding included),
 * out[level * 4 + mode], mode 0 = sigma-only, 1 = full, 2 = full on carried warp / hyper / mask results, 3 = full +
 * reverse sweep for d(sigma)/dx (8 doubles). */
int ndsr_tc_issued_macs(const ndsr_handle* h, double* out);
/* sizeof(ndsr_config), sizeof(ndsr_extra_params), sizeof(ndsr_outputs): lets a
 * foreign-language binding assert its struct mirrors before the first call. */
void ndsr_struct_sizes(int32_t* config, int32_t* extra_params, int32_t* outputs);
/* Upper bound on rays processed per internal pass (scratch = 100 B x rays x samples). */
int ndsr_set_max_chunk(ndsr_handle* h, int64_t max_rays);

/* Early termination of the fine level (north_star: "early-termination scan"; the reference composites every sample,
 * model_utils.py:95-159, so this is OFF by default: transmittance_eps = 0).  With eps > 0, ndsr_render_rays* calls
 * on the tensor-core engine that request only per-ray outputs (render.py's keys) and no render_opts evaluate the
 * fine network at the coarse depths first and then the newly drawn depths front to back in `rounds` rounds of equal
 * rank ranges.  Before each round a scan bounds the transmittance T in front of every depth of the round by the
 * product over the samples evaluated so far (the exact factors (1 - alpha + 1e-10) of model_utils.py:131-136) and
 * drops the depths with T < eps: together they can hold at most eps of a ray's weight, so every per-ray output moves
 * by at most 2 eps x (range of the composited quantity).  The rounds adapt on the device: a round that finds that fewer
 * than 1 % of the depths decided so far were dropped takes all remaining depths at once (the later rounds are empty
 * launches), so a field with nothing to terminate pays for one extra pass, not for `rounds`.
 * ndsr_termination_stats returns the new depths evaluated / seen since the last reset (synchronises `stream`). */
int ndsr_set_early_termination(ndsr_handle* h, float transmittance_eps, int32_t rounds);
int ndsr_termination_stats(ndsr_handle* h, void* stream, int64_t* evaluated, int64_t* seen, int reset);

/* Per-stage device timing for bench.py's roofline: when enabled, every kernel
 * the handle launches is bracketed by CUDA events recorded on the call's own
 * stream.  ndsr_profile_read synchronises those events and returns the
 * accumulated milliseconds and launch counts per stage since the last enable. */
#define NDSR_STAGE_COUNT 6
enum ndsr_stage {
  NDSR_STAGE_SAMPLE = 0,        /* sample_along_rays */
  NDSR_STAGE_FIELD_COARSE = 1,  /* field kernel, coarse level */
  NDSR_STAGE_FIELD_FINE = 2,    /* field kernel, fine level (the dominant kernel) */
  NDSR_STAGE_COMPOSITE = 3,     /* volumetric_rendering + per-ray accumulations */
  NDSR_STAGE_RESAMPLE = 4,      /* sample_pdf */
  NDSR_STAGE_OTHER = 5          /* sharpen_weights, packing, copies */
};
int ndsr_profile_enable(ndsr_handle* h, int on);
int ndsr_profile_read(ndsr_handle* h, double* ms, int64_t* launches);   /* arrays of NDSR_STAGE_COUNT */

/* Diagnostics: runs ONE Dense layer y = act(A W + b) for a 128-row tile through
 * the tensor-core machinery (weight packing, bulk-TMA ring, tcgen05.mma,
 * TMEM epilogue, swizzled split-fp16 write-back) on `device`.  Host pointers:
 * A [128, k_hid + k_in], W [k_hid + k_in, n_out] (Flax layout), bias [n_out],
 * out [128, n_out]; out_readback (nullable) = the activations as re-read from
 * the shared-memory operand image (hi + lo).  terms = 1 | 3; out_kind 0 =
 * hidden layer, 1 = head (n_out <= 16), 2 = hi-only write into the lo region. */
int ndsr_selftest_tc_dense(int device, int k_hid, int k_in, int n_out, int terms, int relu, int out_kind,
                           const float* A, const float* W, const float* bias, float* out,
                           float* out_readback);

#ifdef __cplusplus
}
#endif
#endif  /* NERFDS_B200_H_ */
