// C-ABI of the nerfds_b200 library (include/nerfds_b200.h): handle management,
// parameter repacking, and the host-side orchestration of one
// NerfModel.__call__ (hypernerf/models.py:1419-1565):
//   sample_along_rays -> field(coarse) -> composite -> sample_pdf ->
//   field(fine) -> composite.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "nds_common.cuh"
#include "nds_host.h"

using namespace nds;

static std::string g_create_error;

#define NDS_CUDA(h, expr)                                                            \
  do {                                                                               \
    cudaError_t _e = (expr);                                                         \
    if (_e != cudaSuccess) {                                                         \
      (h)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                 \
      return NDSR_ERR_CUDA;                                                          \
    }                                                                                \
  } while (0)

static int fail(ndsr_handle* h, int code, const std::string& msg) {
  h->err = msg;
  return code;
}

// --------------------------------------------------------------- profiling
namespace {
cudaEvent_t prof_event(ndsr_handle* h) {
  if (!h->prof_pool.empty()) { cudaEvent_t e = h->prof_pool.back(); h->prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
// brackets the kernels launched inside its scope with events on the call's stream
struct ProfScope {
  ndsr_handle* h; cudaStream_t st; int idx = -1;
  ProfScope(ndsr_handle* h_, cudaStream_t st_, int stage) : h(h_), st(st_) {
    if (!h->prof) return;
    ndsr_handle::ProfSpan sp{stage, prof_event(h), prof_event(h)};
    cudaEventRecord(sp.a, st);
    idx = (int)h->prof_spans.size();
    h->prof_spans.push_back(sp);
  }
  ~ProfScope() { if (idx >= 0) cudaEventRecord(h->prof_spans[idx].b, st); }
};
void prof_drain(ndsr_handle* h, bool accumulate) {
  for (auto& sp : h->prof_spans) {
    if (accumulate) {
      float ms = 0.f;
      if (cudaEventSynchronize(sp.b) == cudaSuccess && cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
        h->prof_ms[sp.stage] += ms;
        h->prof_launches[sp.stage] += 1;
      }
    }
    h->prof_pool.push_back(sp.a);
    h->prof_pool.push_back(sp.b);
  }
  h->prof_spans.clear();
}
}  // namespace

extern "C" int ndsr_profile_enable(ndsr_handle* h, int on) {
  if (!h) return NDSR_ERR_INVALID;
  cudaSetDevice(h->device);
  prof_drain(h, false);
  for (int i = 0; i < NDSR_STAGE_COUNT; ++i) { h->prof_ms[i] = 0; h->prof_launches[i] = 0; }
  h->prof = on != 0;
  return NDSR_OK;
}
extern "C" int ndsr_profile_read(ndsr_handle* h, double* ms, int64_t* launches) {
  if (!h || !ms || !launches) return NDSR_ERR_INVALID;
  cudaSetDevice(h->device);
  prof_drain(h, true);
  for (int i = 0; i < NDSR_STAGE_COUNT; ++i) { ms[i] = h->prof_ms[i]; launches[i] = h->prof_launches[i]; }
  return NDSR_OK;
}

// ------------------------------------------------------------------ create
static bool width_ok(int w) { return w == 32 || w == 64 || w == 128 || w == 256; }

static std::string validate(const ndsr_config& c) {
  if (c.size != sizeof(ndsr_config) || c.abi_version != NDSR_ABI_VERSION) return "ndsr_config size/abi mismatch";
  if (c.num_coarse_samples < 3 || c.num_fine_samples < 1) return "num_coarse_samples >= 3 and num_fine_samples >= 1 required";
  if (c.num_coarse_samples + c.num_fine_samples > 2048) return "too many samples per ray (max 2048)";
  if (!width_ok(c.trunk_width) || !width_ok(c.rgb_width)) return "trunk/rgb width must be 32, 64, 128 or 256";
  if (c.trunk_depth < 1 || c.trunk_depth > NDSR_MAX_DEPTH || c.rgb_depth < 0 || c.rgb_depth > NDSR_MAX_DEPTH)
    return "trunk_depth in [1,8], rgb_depth in [0,8] required";
  if (c.trunk_skip == 0 || c.trunk_skip >= c.trunk_depth) return "trunk_skip must be -1 or in [1, depth)";
  if (c.use_warp) {
    if (!width_ok(c.warp_width) || c.warp_depth < 1 || c.warp_depth > NDSR_MAX_DEPTH) return "bad warp MLP shape";
    if (c.warp_skip == 0 || c.warp_skip >= c.warp_depth) return "warp_skip must be -1 or in [1, depth)";
    if (c.warp_embed_dims < 1 || c.warp_embed_dims > 32 || c.num_warp_embeds < 1) return "bad warp embedding shape";
  }
  if (c.use_hyper_sheet) {
    if (!c.use_warp) return "bendy_sheet needs the warp embedding";
    if (!width_ok(c.hyper_sheet_width) || c.hyper_sheet_depth < 1 || c.hyper_sheet_depth > NDSR_MAX_DEPTH)
      return "bad hyper sheet MLP shape";
    if (c.hyper_sheet_skip == 0 || c.hyper_sheet_skip >= c.hyper_sheet_depth) return "hyper_sheet_skip must be -1 or in [1, depth)";
    if (c.hyper_num_dims < 1 || c.hyper_num_dims > 2) return "hyper_num_dims must be 1 or 2";
  }
  if (c.use_predicted_mask) {
    if (!c.use_warp) return "predicted mask needs the warp metadata";
    if (!width_ok(c.mask_width) || c.mask_depth < 1 || c.mask_depth > NDSR_MAX_DEPTH) return "bad mask MLP shape";
    if (c.mask_skip == 0 || c.mask_skip >= c.mask_depth) return "mask_skip must be -1 or in [1, depth)";
  }
  const int degs[][2] = {{c.spatial_min_deg, c.spatial_max_deg}, {c.hyper_point_min_deg, c.hyper_point_max_deg},
                         {c.viewdir_min_deg, c.viewdir_max_deg}, {c.hyper_sheet_min_deg, c.hyper_sheet_max_deg},
                         {c.warp_min_deg, c.warp_max_deg}, {c.norm_input_min_deg, c.norm_input_max_deg},
                         {c.mask_min_deg, c.mask_max_deg}};
  for (auto& d : degs)
    if (d[1] < d[0] || d[1] - d[0] > NDSR_MAX_BANDS) return "posenc degree range out of bounds";
  return "";
}

static int pe_dim(int C, int lo, int hi, int ident) { return 2 * (hi - lo) * C + (ident ? C : 0); }

extern "C" int ndsr_abi_version(void) { return NDSR_ABI_VERSION; }
extern "C" void ndsr_struct_sizes(int32_t* config, int32_t* extra_params, int32_t* outputs) {
  if (config) *config = (int32_t)sizeof(ndsr_config);
  if (extra_params) *extra_params = (int32_t)sizeof(ndsr_extra_params);
  if (outputs) *outputs = (int32_t)sizeof(ndsr_outputs);
}
extern "C" int ndsr_set_max_chunk(ndsr_handle* h, int64_t max_rays) {
  if (!h || max_rays < 1) return NDSR_ERR_INVALID;
  h->max_chunk = max_rays;
  return NDSR_OK;
}

extern "C" int ndsr_set_early_termination(ndsr_handle* h, float transmittance_eps, int32_t rounds) {
  if (!h || !(transmittance_eps >= 0.f) || transmittance_eps >= 1.f || rounds < 1 || rounds > 64) return NDSR_ERR_INVALID;
  h->term_eps = transmittance_eps;
  h->term_rounds = rounds;
  return NDSR_OK;
}

extern "C" int ndsr_termination_stats(ndsr_handle* h, void* stream, int64_t* evaluated, int64_t* seen, int reset) {
  if (!h) return NDSR_ERR_INVALID;
  unsigned long long v[2] = {h->term_stats_keep[0], h->term_stats_keep[1]};
  if (h->term_stats) {
    NDS_CUDA(h, cudaSetDevice(h->device));
    NDS_CUDA(h, cudaStreamSynchronize((cudaStream_t)stream));
    NDS_CUDA(h, cudaMemcpy(v, h->term_stats, sizeof v, cudaMemcpyDeviceToHost));
    if (reset) NDS_CUDA(h, cudaMemset(h->term_stats, 0, sizeof v));
  }
  if (reset) h->term_stats_keep[0] = h->term_stats_keep[1] = 0;
  if (evaluated) *evaluated = (int64_t)v[0];
  if (seen) *seen = (int64_t)v[1];
  return NDSR_OK;
}

extern "C" int ndsr_create(const ndsr_config* cfg, int device, ndsr_handle** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return NDSR_ERR_INVALID; }
  std::string v = validate(*cfg);
  if (!v.empty()) { g_create_error = v; return NDSR_ERR_INVALID; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || device < 0 || device >= ndev) {
    g_create_error = std::string("no such CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "index out of range");
    return NDSR_ERR_CUDA;
  }
  ndsr_handle* h = new ndsr_handle();
  h->cfg = *cfg;
  h->device = device;
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    g_create_error = cudaGetErrorString(e);
    delete h;
    return NDSR_ERR_CUDA;
  }
  h->num_sms = prop.multiProcessorCount;
  h->tc_no_split = getenv("NDS_TC_NO_SPLIT") != nullptr;
  h->cc_major = prop.major;
  const ndsr_config& c = h->cfg;
  h->H = c.use_hyper_sheet ? c.hyper_num_dims : 0;
  h->dim_mask_in = pe_dim(3, c.mask_min_deg, c.mask_max_deg, 0) + c.mask_embed_dims;
  h->dim_warp_in = pe_dim(3, c.warp_min_deg, c.warp_max_deg, c.warp_use_posenc_identity) + c.warp_embed_dims + (c.use_mask_in_warp ? 1 : 0);
  h->dim_hyper_in = pe_dim(3, c.hyper_sheet_min_deg, c.hyper_sheet_max_deg, 0) + c.warp_embed_dims + (c.use_mask_in_hyper ? 1 : 0);
  h->dim_trunk_in = pe_dim(3, c.spatial_min_deg, c.spatial_max_deg, c.use_posenc_identity) +
                    (h->H ? pe_dim(h->H, c.hyper_point_min_deg, c.hyper_point_max_deg, 0) : 0);
  h->dim_view = c.use_viewdirs ? pe_dim(3, c.viewdir_min_deg, c.viewdir_max_deg, c.use_posenc_identity) : 0;
  h->dim_norm = c.norm_input_posenc ? pe_dim(3, c.norm_input_min_deg, c.norm_input_max_deg, c.use_posenc_identity) : 3;
  h->max_in = h->dim_trunk_in;
  if (c.use_warp && h->dim_warp_in > h->max_in) h->max_in = h->dim_warp_in;
  if (c.use_hyper_sheet && h->dim_hyper_in > h->max_in) h->max_in = h->dim_hyper_in;
  if (c.use_predicted_mask && h->dim_mask_in > h->max_in) h->max_in = h->dim_mask_in;
  h->max_w = c.trunk_width;
  if (c.use_warp && c.warp_width > h->max_w) h->max_w = c.warp_width;
  if (c.use_hyper_sheet && c.hyper_sheet_width > h->max_w) h->max_w = c.hyper_sheet_width;
  if (c.use_predicted_mask && c.mask_width > h->max_w) h->max_w = c.mask_width;
  if (h->max_in > 256) { g_create_error = "MLP input wider than 256 features"; delete h; return NDSR_ERR_UNSUPPORTED; }
  int ld[5];
  const size_t smem = field_simt_smem_bytes(c, h->max_in, h->max_w, h->dim_view + h->dim_norm, true, ld);
  if (smem > (size_t)prop.sharedMemPerBlockOptin) {
    g_create_error = "configuration needs more shared memory than the device offers";
    delete h;
    return NDSR_ERR_UNSUPPORTED;
  }
  h->engine = NDSR_ENGINE_SIMT;
  if (c.engine == NDSR_ENGINE_TC || c.engine == NDSR_ENGINE_AUTO) {
    std::string why = tc_engine_supports(c, prop.major, prop.minor);
    if (why.empty()) h->engine = NDSR_ENGINE_TC;
    else if (c.engine == NDSR_ENGINE_TC) { g_create_error = "tensor-core engine unavailable: " + why; delete h; return NDSR_ERR_UNSUPPORTED; }
  }
  *out = h;
  return NDSR_OK;
}

static void free_scratch(ndsr_handle* h) {
  if (h->term_stats)     // the statistics live in the scratch: keep their value while it is regrown
    cudaMemcpy(h->term_stats_keep, h->term_stats, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  h->term_stats = nullptr; h->term_index = nullptr; h->term_count = nullptr; h->term_round_stats = nullptr;
  for (void* p : h->scratch_allocs) cudaFree(p);
  h->scratch_allocs.clear();
  h->cap_rays = 0;
}

extern "C" void ndsr_destroy(ndsr_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  free_scratch(h);
  if (h->arena) cudaFree(h->arena);
  if (h->in_stage) cudaFree(h->in_stage);
  if (h->out_stage) cudaFree(h->out_stage);
  for (int i = 0; i < 2; ++i) { if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]); if (h->ev_free[i]) cudaEventDestroy(h->ev_free[i]); }
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  prof_drain(h, false);
  for (cudaEvent_t e : h->prof_pool) cudaEventDestroy(e);
  tc_engine_free(h);
  delete h;
}

extern "C" const char* ndsr_last_error(const ndsr_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }
extern "C" int ndsr_engine_in_use(const ndsr_handle* h) { return h ? h->engine : 0; }
extern "C" int64_t ndsr_kernel_launches(const ndsr_handle* h) { return h ? h->launches : 0; }

// ------------------------------------------------------------ load_params
namespace {
struct Arena {
  std::vector<float> host;
  size_t add(const float* p, size_t n) {
    size_t off = host.size();
    host.insert(host.end(), p, p + n);
    while (host.size() % 4) host.push_back(0.f);   // keep 16-byte alignment
    return off;
  }
};
struct DenseOff { size_t W, b; int K, N; bool present = false; };
}  // namespace

static const ndsr_tensor* find(const std::map<std::string, const ndsr_tensor*>& m, const std::string& name) {
  auto it = m.find(name);
  return it == m.end() ? nullptr : it->second;
}

static int take_dense(ndsr_handle* h, const std::map<std::string, const ndsr_tensor*>& m, const std::string& prefix,
                      int K, int N, Arena& A, DenseOff& d, HostDense* keep) {
  const ndsr_tensor* k = find(m, prefix + "/kernel");
  const ndsr_tensor* b = find(m, prefix + "/bias");
  if (!k || !b) return fail(h, NDSR_ERR_PARAMS, "missing parameter " + prefix + "/{kernel,bias}");
  if (k->rows != K || k->cols != N || b->rows * b->cols != N) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: expected kernel [%d,%d] bias [%d], got [%lld,%lld] / %lld", prefix.c_str(), K, N, N,
             (long long)k->rows, (long long)k->cols, (long long)(b->rows * b->cols));
    return fail(h, NDSR_ERR_PARAMS, buf);
  }
  d.W = A.add(k->data, (size_t)K * N);
  d.b = A.add(b->data, N);
  d.K = K; d.N = N; d.present = true;
  if (keep) { keep->W.assign(k->data, k->data + (size_t)K * N); keep->b.assign(b->data, b->data + N); keep->K = K; keep->N = N; }
  return NDSR_OK;
}

struct MlpOff { DenseOff hidden[NDSR_MAX_DEPTH]; DenseOff logit; size_t WTh[NDSR_MAX_DEPTH], WTx[NDSR_MAX_DEPTH]; int kx_pad; };

static int take_mlp(ndsr_handle* h, const std::map<std::string, const ndsr_tensor*>& m, const std::string& prefix,
                    int in_dim, int depth, int width, int skip, int out_ch, bool transposed, Arena& A, MlpOff& o,
                    HostMlp* keep) {
  int d = in_dim;
  if (keep) { keep->depth = depth; keep->width = width; keep->in_dim = in_dim; keep->skip = skip; keep->hidden.resize(depth); }
  for (int l = 0; l < depth; ++l) {
    if (l == skip) d += in_dim;
    int rc = take_dense(h, m, prefix + "/hidden_" + std::to_string(l), d, width, A, o.hidden[l], keep ? &keep->hidden[l] : nullptr);
    if (rc) return rc;
    d = width;
  }
  if (out_ch > 0) {
    int rc = take_dense(h, m, prefix + "/logit", d, out_ch, A, o.logit, keep ? &keep->logit : nullptr);
    if (rc) return rc;
  }
  int kxp = 32;
  while (kxp < in_dim) kxp *= 2;
  o.kx_pad = kxp;
  if (transposed) {
    for (int l = 0; l < depth; ++l) {
      const float* W = A.host.data() + o.hidden[l].W;   // NOTE: re-fetch after every add (vector may move)
      const int K = o.hidden[l].K;
      o.WTh[l] = o.WTx[l] = 0;
      if (l > 0) {   // hidden slice rows [0,width): WTh[n][k] = W[k][n]
        std::vector<float> t((size_t)width * width);
        for (int k = 0; k < width; ++k) for (int n = 0; n < width; ++n) t[(size_t)n * width + k] = W[(size_t)k * width + n];
        o.WTh[l] = A.add(t.data(), t.size());
      }
      if (l == 0 || l == skip) {   // input slice rows [K-in_dim, K)
        W = A.host.data() + o.hidden[l].W;
        const int r0 = K - in_dim;
        std::vector<float> t((size_t)width * kxp, 0.f);
        for (int k = 0; k < in_dim; ++k) for (int n = 0; n < width; ++n) t[(size_t)n * kxp + k] = W[(size_t)(r0 + k) * width + n];
        o.WTx[l] = A.add(t.data(), t.size());
      }
    }
  }
  return NDSR_OK;
}

static void bind_dense(const DenseOff& o, const float* base, DenseW& d) {
  d.W = o.present ? base + o.W : nullptr;
  d.b = o.present ? base + o.b : nullptr;
  d.K = o.K; d.N = o.present ? o.N : 0;
}
static void bind_mlp(const MlpOff& o, const float* base, int in_dim, int depth, int width, int skip, MlpW& m, MlpWT* t) {
  m.depth = depth; m.width = width; m.in_dim = in_dim; m.skip = skip;
  for (int l = 0; l < depth; ++l) bind_dense(o.hidden[l], base, m.hidden[l]);
  bind_dense(o.logit, base, m.logit);
  if (t) {
    t->kx_pad = o.kx_pad;
    for (int l = 0; l < depth; ++l) { t->WTh[l] = base + o.WTh[l]; t->WTx[l] = base + o.WTx[l]; }
    t->WTlogit = nullptr;
  }
}

extern "C" int ndsr_load_params(ndsr_handle* h, const ndsr_tensor* tensors, int n) {
  if (!h || !tensors || n <= 0) return h ? fail(h, NDSR_ERR_INVALID, "null argument") : NDSR_ERR_INVALID;
  NDS_CUDA(h, cudaSetDevice(h->device));
  const ndsr_config& c = h->cfg;
  std::map<std::string, const ndsr_tensor*> m;
  for (int i = 0; i < n; ++i) {
    if (!tensors[i].name || !tensors[i].data) return fail(h, NDSR_ERR_PARAMS, "tensor with null name/data");
    m[tensors[i].name] = &tensors[i];
  }
  Arena A;
  A.host.reserve(4u << 20);
  HostModel& HM = h->host_model;
  HM = HostModel();
  MlpOff mask{}, warp{}, hyper{}, trunk[2]{}, rgb[2]{};
  DenseOff warp_w, warp_v, bott[2], alpha[2];
  size_t warp_embed = 0, mask_embed = 0, col0[2] = {0, 0};
  int rc;
  if (c.use_warp) {
    const ndsr_tensor* e = find(m, "warp_embed/embed/embedding");
    if (!e || e->rows != c.num_warp_embeds || e->cols != c.warp_embed_dims)
      return fail(h, NDSR_ERR_PARAMS, "warp_embed/embed/embedding missing or mis-shaped");
    warp_embed = A.add(e->data, (size_t)e->rows * e->cols);
    if ((rc = take_mlp(h, m, "warp_field/trunk", h->dim_warp_in, c.warp_depth, c.warp_width, c.warp_skip, 0, true, A, warp, &HM.warp))) return rc;
    if ((rc = take_dense(h, m, "warp_field/branches_w/logit", c.warp_width, 3, A, warp_w, &HM.warp_w))) return rc;
    if ((rc = take_dense(h, m, "warp_field/branches_v/logit", c.warp_width, 3, A, warp_v, &HM.warp_v))) return rc;
  }
  if (c.use_predicted_mask) {
    const ndsr_tensor* e = find(m, "mask_embed/embed/embedding");
    if (!e || e->rows != c.num_warp_embeds || e->cols != c.mask_embed_dims)
      return fail(h, NDSR_ERR_PARAMS, "mask_embed/embed/embedding missing or mis-shaped");
    mask_embed = A.add(e->data, (size_t)e->rows * e->cols);
    if ((rc = take_mlp(h, m, "mask_mlp/MLP_0", h->dim_mask_in, c.mask_depth, c.mask_width, c.mask_skip, 1, false, A, mask, &HM.mask))) return rc;
  }
  if (c.use_hyper_sheet)
    if ((rc = take_mlp(h, m, "hyper_sheet_mlp/MLP_0", h->dim_hyper_in, c.hyper_sheet_depth, c.hyper_sheet_width,
                       c.hyper_sheet_skip, c.hyper_num_dims, true, A, hyper, &HM.hyper))) return rc;
  const int n_alpha = 1 + (c.predict_norm ? 3 : 0);
  const int rgb_in = c.trunk_width + h->dim_view + (c.use_x_in_rgb_condition ? c.trunk_width : 0) + (c.predict_norm ? h->dim_norm : 0);
  h->rgb_in_full = rgb_in;
  for (int lv = 0; lv < 2; ++lv) {
    const std::string p = lv == 0 ? "nerf_mlps_coarse" : "nerf_mlps_fine";
    if ((rc = take_mlp(h, m, p + "/trunk_mlp", h->dim_trunk_in, c.trunk_depth, c.trunk_width, c.trunk_skip, 0, true, A, trunk[lv], &HM.trunk[lv]))) return rc;
    if (c.use_viewdirs)
      if ((rc = take_dense(h, m, p + "/bottleneck", c.trunk_width, c.trunk_width, A, bott[lv], &HM.bottleneck[lv]))) return rc;
    if ((rc = take_dense(h, m, p + "/alpha_mlp/logit", c.trunk_width, n_alpha, A, alpha[lv], &HM.alpha[lv]))) return rc;
    if ((rc = take_mlp(h, m, p + "/rgb_mlp", rgb_in, c.rgb_depth, c.rgb_width, -1, 3, false, A, rgb[lv], &HM.rgb[lv]))) return rc;
    std::vector<float> col(c.trunk_width);
    const float* W = A.host.data() + alpha[lv].W;
    for (int k = 0; k < c.trunk_width; ++k) col[k] = W[(size_t)k * n_alpha];
    col0[lv] = A.add(col.data(), col.size());
  }
  if (h->arena) { cudaFree(h->arena); h->arena = nullptr; }
  NDS_CUDA(h, cudaMalloc(&h->arena, A.host.size() * sizeof(float)));
  NDS_CUDA(h, cudaMemcpy(h->arena, A.host.data(), A.host.size() * sizeof(float), cudaMemcpyHostToDevice));
  const float* base = h->arena;
  ModelW& M = h->M;
  memset(&M, 0, sizeof M);
  if (c.use_warp) {
    M.warp_embed = base + warp_embed;
    bind_mlp(warp, base, h->dim_warp_in, c.warp_depth, c.warp_width, c.warp_skip, M.warp, &M.warp_T);
    bind_dense(warp_w, base, M.warp_w);
    bind_dense(warp_v, base, M.warp_v);
  }
  if (c.use_predicted_mask) {
    M.mask_embed = base + mask_embed;
    bind_mlp(mask, base, h->dim_mask_in, c.mask_depth, c.mask_width, c.mask_skip, M.mask, nullptr);
  }
  if (c.use_hyper_sheet) bind_mlp(hyper, base, h->dim_hyper_in, c.hyper_sheet_depth, c.hyper_sheet_width, c.hyper_sheet_skip, M.hyper, &M.hyper_T);
  for (int lv = 0; lv < 2; ++lv) {
    bind_mlp(trunk[lv], base, h->dim_trunk_in, c.trunk_depth, c.trunk_width, c.trunk_skip, M.level[lv].trunk, &M.trunk_T[lv]);
    bind_dense(bott[lv], base, M.level[lv].bottleneck);
    bind_dense(alpha[lv], base, M.level[lv].alpha);
    bind_mlp(rgb[lv], base, rgb_in, c.rgb_depth, c.rgb_width, -1, M.level[lv].rgb, nullptr);
    M.alpha_col0[lv] = base + col0[lv];
  }
  if (c.use_warp) {
    HM.warp_embed.assign(A.host.data() + warp_embed, A.host.data() + warp_embed + (size_t)c.num_warp_embeds * c.warp_embed_dims);
  }
  if (c.use_predicted_mask) {
    HM.mask_embed.assign(A.host.data() + mask_embed, A.host.data() + mask_embed + (size_t)c.num_warp_embeds * c.mask_embed_dims);
  }
  if (h->engine == NDSR_ENGINE_TC) {
    int rc2 = tc_engine_load(h);
    if (rc2) return rc2;
  }
  h->loaded = true;
  return NDSR_OK;
}

// ------------------------------------------------------------ call params
static void fill_pe(PosencSpec& pe, int lo, int hi, int ident, bool has_alpha, float alpha) {
  pe.min_deg = lo; pe.num_bands = hi - lo; pe.identity = ident;
  for (int k = 0; k < NDSR_MAX_BANDS; ++k) pe.window[k] = 1.f;
  if (has_alpha) {
    const float pi = 3.14159274101257324f;   // fp32(pi), model_utils.py:436
    for (int k = 0; k < pe.num_bands; ++k) {
      float x = alpha - (float)(lo + k);
      x = x < 0.f ? 0.f : (x > 1.f ? 1.f : x);
      pe.window[k] = 0.5f * (1.f + cosf(pi * x + pi));
    }
  }
}

void nds::make_call_params(const ndsr_config& c, const ndsr_extra_params& ep, CallParams& cp) {
  fill_pe(cp.pe_mask, c.mask_min_deg, c.mask_max_deg, 0, true, ep.warp_alpha);                       // models.py:967
  fill_pe(cp.pe_warp, c.warp_min_deg, c.warp_max_deg, c.warp_use_posenc_identity, true, ep.warp_alpha);   // warping.py:209-213
  fill_pe(cp.pe_hsheet, c.hyper_sheet_min_deg, c.hyper_sheet_max_deg, 0, true, ep.hyper_sheet_alpha);      // models.py:663-666
  fill_pe(cp.pe_spatial, c.spatial_min_deg, c.spatial_max_deg, c.use_posenc_identity, true, ep.nerf_alpha); // models.py:502-507
  fill_pe(cp.pe_hyperpt, c.hyper_point_min_deg, c.hyper_point_max_deg, 0, true, ep.hyper_alpha);            // models.py:510-515
  fill_pe(cp.pe_view, c.viewdir_min_deg, c.viewdir_max_deg, c.use_posenc_identity, false, 0.f);            // models.py:401-405
  fill_pe(cp.pe_norm, c.norm_input_min_deg, c.norm_input_max_deg, c.use_posenc_identity, true, ep.norm_input_alpha);  // models.py:1142-1148
  cp.mask_ratio = ep.mask_ratio;
  cp.use_predicted_norm = ep.use_predicted_norm;
  cp.use_sigma_gradient = ep.use_sigma_gradient;
}

// ---------------------------------------------------------------- scratch
static int ensure_scratch(ndsr_handle* h, int64_t rays, cudaStream_t st) {
  if (rays <= h->cap_rays) return NDSR_OK;
  NDS_CUDA(h, cudaStreamSynchronize(st));
  free_scratch(h);
  const ndsr_config& c = h->cfg;
  const int64_t Sf = c.num_coarse_samples + c.num_fine_samples;
  const int64_t smax = Sf > h->max_samples_seen ? Sf : h->max_samples_seen;
  auto alloc = [&](float** p, int64_t nfloats) -> cudaError_t {
    cudaError_t e = cudaMalloc((void**)p, (size_t)nfloats * sizeof(float));
    if (e == cudaSuccess) h->scratch_allocs.push_back(*p);
    return e;
  };
  NDS_CUDA(h, alloc(&h->planes, (int64_t)P_COUNT * rays * smax));
  NDS_CUDA(h, alloc(&h->z_coarse, rays * c.num_coarse_samples));
  NDS_CUDA(h, alloc(&h->z_fine, rays * smax));
  NDS_CUDA(h, alloc(&h->w_coarse, rays * c.num_coarse_samples));
  NDS_CUDA(h, alloc(&h->w_sg, rays * smax));
  NDS_CUDA(h, alloc(&h->argmax, rays));
  NDS_CUDA(h, alloc(&h->carry, (int64_t)C_COUNT * rays * c.num_coarse_samples));
  {
    float* pf = nullptr;
    NDS_CUDA(h, alloc(&pf, rays * smax));
    h->src_elem = reinterpret_cast<int32_t*>(pf);
  }
  NDS_CUDA(h, alloc(&h->z_new, rays * (smax > c.num_coarse_samples ? smax - c.num_coarse_samples : 1)));
  {
    float* pf = nullptr;
    // header in front of the index list: the count (4 B), the statistics (2 x 8 B at byte 8), the per-round counters
    // of the adaptive rounds (NDS_TERM_MAX_ROUNDS x 2 x 8 B at byte 32)
    const int64_t hdr = 8 + 4 * NDS_TERM_MAX_ROUNDS;         // floats
    NDS_CUDA(h, alloc(&pf, rays * (smax > c.num_coarse_samples ? smax - c.num_coarse_samples : 1) + hdr));
    h->term_index = reinterpret_cast<int32_t*>(pf) + hdr;
    h->term_count = reinterpret_cast<int32_t*>(pf);
    h->term_round_stats = reinterpret_cast<unsigned long long*>(pf) + 4;
    unsigned long long* stats = reinterpret_cast<unsigned long long*>(pf) + 1;
    NDS_CUDA(h, cudaMemcpy(stats, h->term_stats_keep, 2 * sizeof(unsigned long long), cudaMemcpyHostToDevice));
    h->term_stats = stats;
  }
  h->cap_rays = rays;
  h->cap_samples = smax;
  return NDSR_OK;
}

static bool wants_per_sample(const ndsr_outputs* o) {
  return o && (o->weights || o->alpha || o->accum_prod || o->sigma || o->sharp_weights || o->back_facing ||
               o->predicted_mask || o->points || o->warped_points || o->delta_x || o->predicted_norm ||
               o->target_norm || o->z_vals);
}

static ndsr_outputs offset_outputs(const ndsr_outputs* o, int64_t r0, int S, int H) {
  ndsr_outputs q;
  memset(&q, 0, sizeof q);
  if (!o) return q;
#define OFF(field, per) q.field = o->field ? o->field + r0 * (int64_t)(per) : nullptr
  OFF(rgb, 3); OFF(depth, 1); OFF(med_depth, 1); OFF(acc, 1); OFF(ray_norm, 3); OFF(ray_rotation_field, 3);
  OFF(ray_translation_field, 3); OFF(ray_delta_x, 3); OFF(ray_hyper_points, H); OFF(ray_predicted_mask, 1);
  OFF(med_points, 3 + H); OFF(z_vals, S); OFF(weights, S); OFF(alpha, S); OFF(accum_prod, S); OFF(sigma, S);
  OFF(sharp_weights, S); OFF(back_facing, S); OFF(predicted_mask, S); OFF(points, 3 * S);
  OFF(warped_points, (3 + H) * S); OFF(delta_x, 3 * S); OFF(predicted_norm, 3 * S); OFF(target_norm, 3 * S);
#undef OFF
  return q;
}

// One level on given samples: field kernel + composite (+ sharpen).  `z` is [B,S] device.
static int run_level(ndsr_handle* h, cudaStream_t st, int level, int64_t B, int S, const float* points,
                     const float* z, const float* origins, const float* dirs, const float* viewdirs,
                     const uint32_t* warp_id, const float* gt_mask, const ndsr_extra_params& ep,
                     const CallParams& cp, int sample_at_infinity, const ndsr_outputs* out, float* weights_keep,
                     bool need_rgb, float* carry_out = nullptr, const int32_t* src_elem = nullptr, int n_carried = 0,
                     bool apply_filter = true) {
  const ndsr_config& c = h->cfg;
  ndsr_outputs o;
  memset(&o, 0, sizeof o);
  if (out) o = *out;
  const bool want_tnorm = c.predict_norm && o.target_norm;
  const bool want_grad_norm = !c.predict_norm && o.ray_norm;
  const bool need_grad = want_tnorm || want_grad_norm || ep.use_sigma_gradient;
  FieldArgs fa;
  memset(&fa, 0, sizeof fa);
  fa.n_samples_total = B * S; fa.S = S; fa.level = level; fa.points = points; fa.z = z;
  fa.origins = origins; fa.dirs = dirs; fa.viewdirs = viewdirs ? viewdirs : dirs; fa.warp_id = warp_id;
  fa.gt_mask = gt_mask; fa.planes = h->planes; fa.plane_stride = B * S;
  fa.sigma_only = need_rgb ? 0 : 1; fa.need_grad = need_grad ? 1 : 0;
  // planes behind the requested keys only (sigma_raw always): the coarse level of a render call writes one plane
  uint32_t pm = 0;
  if (o.rgb) pm |= PG_RGB;
  if (o.ray_norm || o.predicted_norm || o.back_facing) pm |= PG_NORM;
  if (o.ray_predicted_mask || o.predicted_mask) pm |= PG_MASK;
  if (o.med_points || o.ray_delta_x || o.delta_x || o.warped_points || o.ray_hyper_points) pm |= PG_WARPED;
  if (o.ray_rotation_field) pm |= PG_ROT;
  if (o.ray_translation_field) pm |= PG_TRANS;
  fa.plane_mask = pm;
  {
    ProfScope ps(h, st, level == 0 ? NDSR_STAGE_FIELD_COARSE : NDSR_STAGE_FIELD_FINE);
    if (h->engine == NDSR_ENGINE_TC && !need_grad && src_elem) {
      // split fine pass: the n_carried coarse depths of every ray re-use the coarse pass's warp / hyper / mask
      // results (same points, same shared networks) and only run the template NeRF; the new depths run everything.
      // Dense blocks of the planes: carried samples [0, B n_carried), new samples after them.
      // Early termination (ndsr_set_early_termination; render-mode calls without per-sample outputs or
      // render_opts): the carried launch goes first, a scan of its sigmas bounds the transmittance in front of
      // every new depth, and the second launch evaluates only the depths that can still carry weight.
      const bool term = h->term_eps > 0.f && !wants_per_sample(out) && !(apply_filter && (ep.filter_flags & 3)) &&
                        B * (int64_t)(S - n_carried) < (int64_t)0x7fffffff;
      FieldArgs fn = fa, fc = fa;
      fn.S = S - n_carried; fn.z = h->z_new; fn.n_samples_total = B * (S - n_carried);
      fn.planes = h->planes + B * n_carried;
      fc.S = n_carried; fc.z = nullptr; fc.n_samples_total = B * n_carried; fc.planes = h->planes;
      fc.carry = h->carry; fc.carry_stride = B * n_carried;
      int rc;
      if (term) {
        if ((rc = tc_engine_field(h, cp, fc, st))) return rc;
        TerminationArgs ta;
        memset(&ta, 0, sizeof ta);
        ta.n_rays = B; ta.S = S; ta.n_carried = n_carried; ta.z = z; ta.dirs = dirs; ta.src_elem = src_elem;
        ta.planes = h->planes; ta.plane_stride = B * S; ta.plane_mask = pm; ta.H = h->H; ta.has_warp = c.use_warp;
        ta.sample_at_infinity = sample_at_infinity; ta.eps = h->term_eps;
        ta.index = h->term_index; ta.n_active = h->term_count; ta.stats = h->term_stats;
        fn.index = h->term_index; fn.n_active = h->term_count;
        const int n_new = S - n_carried, rounds = h->term_rounds < n_new ? h->term_rounds : n_new;
        ta.round_stats = h->term_round_stats; ta.merge_frac = 0.99f;
        NDS_CUDA(h, cudaMemsetAsync(h->term_round_stats, 0, 2 * sizeof(unsigned long long) * NDS_TERM_MAX_ROUNDS, st));
        for (int r = 0; r < rounds; ++r) {       // front to back: each round sees the sigmas of the rounds before it
          ta.round = r;
          ta.rank_lo = (int)((int64_t)n_new * r / rounds); ta.rank_hi = (int)((int64_t)n_new * (r + 1) / rounds);
          NDS_CUDA(h, launch_termination_scan(ta, h->num_sms, st));
          h->launches++;
          fn.n_samples_total = B * (int64_t)(ta.rank_hi - ta.rank_lo);     // (upper bound: sizes the grid)
          if ((rc = tc_engine_field(h, cp, fn, st))) return rc;
        }
      } else {
        if ((rc = tc_engine_field(h, cp, fn, st))) return rc;
        if ((rc = tc_engine_field(h, cp, fc, st))) return rc;
      }
    } else if (h->engine == NDSR_ENGINE_TC && (!need_grad || !ep.use_sigma_gradient)) {
      // (with d(sigma)/dx: one launch of the program that ends every tile with the reverse sweep; only
      //  use_sigma_gradient -- the gradient as the rgb branch's normal input -- stays on the CUDA-core engine)
      if (need_grad) fa.sigma_only = 0;
      else { fa.carry_out = carry_out; fa.carry_stride = B * S; }
      int rc = tc_engine_field(h, cp, fa, st);
      if (rc) return rc;
    } else {
      NDS_CUDA(h, launch_field_simt(h->M, cp, fa, c, h->max_in, h->max_w, h->dim_view + h->dim_norm, h->num_sms, st));
      h->launches++;
    }
  }
  CompositeArgs ca;
  memset(&ca, 0, sizeof ca);
  ca.n_rays = B; ca.S = S; ca.H = h->H; ca.planes = h->planes; ca.plane_stride = B * S; ca.z = z; ca.dirs = dirs;
  ca.plane_mask = pm;
  ca.viewdirs = viewdirs ? viewdirs : dirs; ca.origins = origins; ca.points = points;
  ca.sigma_is_activated = 0; ca.white_bkgd = c.use_white_background; ca.sample_at_infinity = sample_at_infinity;
  ca.has_norm = c.predict_norm; ca.has_warp = c.use_warp; ca.has_mask = c.use_predicted_mask; ca.has_grad = need_grad;
  if (apply_filter && (ep.filter_flags & 3)) {      // filter_sigma (models.py:38-66)
    ca.filter_flags = ep.filter_flags & 3;
    ca.dust_threshold = ep.dust_threshold;
    for (int i = 0; i < 6; ++i) ca.bbox[i] = ep.bounding_box[i];
  }
  ca.out = o;
  ca.src_elem = (h->engine == NDSR_ENGINE_TC && !need_grad) ? src_elem : nullptr; ca.n_carried = n_carried;
  if (weights_keep) ca.out.weights = weights_keep;   // coarse weights feed sample_pdf
  if (level == 1 && h->n_mirror > 0) {
    // mirrored stores add fixed address offsets: only legal when the destination is the registered frame buffer
    const float* per_ray[] = {o.rgb, o.depth, o.med_depth, o.acc, o.ray_norm, o.ray_rotation_field, o.ray_translation_field,
                              o.ray_delta_x, o.ray_hyper_points, o.ray_predicted_mask, o.med_points};
    int inside = 0, outside = 0;
    for (const float* p : per_ray) {
      if (!p) continue;
      const char* q = reinterpret_cast<const char*>(p);
      (q >= h->mirror_base && q < h->mirror_base + h->mirror_bytes) ? ++inside : ++outside;
    }
    if (inside && outside)
      return fail(h, NDSR_ERR_INVALID, "per-ray outputs lie partly inside and partly outside the mirrored frame buffer");
    if (inside) {
      ca.n_mirror = h->n_mirror;
      for (int m = 0; m < h->n_mirror; ++m) ca.mirror_delta[m] = h->mirror_delta[m];
    }
  }
  const bool sharp = c.use_mask_sharp_weights && o.sharp_weights;
  ca.argmax_idx = sharp ? h->argmax : nullptr;
  ca.weights_sg = sharp ? h->w_sg : nullptr;
  {
    ProfScope ps(h, st, NDSR_STAGE_COMPOSITE);
    NDS_CUDA(h, launch_composite(ca, h->num_sms, st));
    h->launches++;
  }
  ProfScope ps(h, st, NDSR_STAGE_OTHER);
  if (weights_keep && o.weights)
    NDS_CUDA(h, cudaMemcpyAsync(o.weights, weights_keep, (size_t)B * S * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (sharp) {
    NDS_CUDA(h, launch_sharpen(B, S, h->w_sg, z, h->argmax, ep.sharp_weights_std, o.sharp_weights, h->num_sms, st));
    h->launches++;
  }
  return NDSR_OK;
}

static int check_call(ndsr_handle* h, const ndsr_extra_params* ep, const uint32_t* warp_id, const float* gt_mask) {
  if (!h->loaded) return fail(h, NDSR_ERR_NOT_LOADED, "ndsr_load_params has not been called");
  if (!ep) return fail(h, NDSR_ERR_INVALID, "extra params required");
  const ndsr_config& c = h->cfg;
  if (c.use_warp && !warp_id) return fail(h, NDSR_ERR_INVALID, "metadata['warp'] ids required when use_warp");
  if (ep->use_predicted_norm && ep->use_sigma_gradient) return fail(h, NDSR_ERR_INVALID, "use_predicted_norm and use_sigma_gradient are exclusive (models.py:1108,1114)");
  if (ep->use_predicted_norm && !c.predict_norm) return fail(h, NDSR_ERR_INVALID, "use_predicted_norm needs predict_norm");
  if (c.predict_norm && !ep->use_predicted_norm && !ep->use_sigma_gradient)
    return fail(h, NDSR_ERR_PARAMS, "rgb branch was built with a normal input (predict_norm); pass use_predicted_norm or use_sigma_gradient");
  const bool need_gt = c.use_predicted_mask ? (ep->mask_ratio != 1.f) : (c.use_mask_in_warp || c.use_mask_in_hyper);
  if (need_gt && !gt_mask) return fail(h, NDSR_ERR_INVALID, "rays_dict['mask'] required (mask_ratio != 1 or no predicted mask)");
  return NDSR_OK;
}

extern "C" int ndsr_render_samples(ndsr_handle* h, void* stream, int level, int64_t n_rays, int32_t n_samples,
                                   const float* points, const float* z_vals, const float* origins,
                                   const float* directions, const float* viewdirs, const uint32_t* warp_id,
                                   const float* gt_mask, const ndsr_extra_params* ep,
                                   int32_t use_sample_at_infinity, const ndsr_outputs* out) {
  if (!h) return NDSR_ERR_INVALID;
  if (n_rays == 0) return NDSR_OK;
  int rc = check_call(h, ep, warp_id, gt_mask);
  if (rc) return rc;
  if (level < 0 || level > 1 || n_rays < 0 || n_samples < 2 || !z_vals || !directions || (!points && !origins))
    return fail(h, NDSR_ERR_INVALID, "bad ndsr_render_samples argument");
  if (n_rays == 0) return NDSR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  NDS_CUDA(h, cudaSetDevice(h->device));
  if (n_samples > h->max_samples_seen) { h->max_samples_seen = n_samples; if (h->cap_rays) { NDS_CUDA(h, cudaStreamSynchronize(st)); free_scratch(h); } }
  if ((rc = ensure_scratch(h, n_rays, st))) return rc;
  CallParams cp;
  make_call_params(h->cfg, *ep, cp);
  return run_level(h, st, level, n_rays, n_samples, points, z_vals, origins, directions, viewdirs, warp_id, gt_mask,
                   *ep, cp, use_sample_at_infinity, out, nullptr, true);
}

static int render_rays_chunk(ndsr_handle* h, cudaStream_t st, int64_t B, const float* origins, const float* dirs,
                             const float* viewdirs, const uint32_t* warp_id, const float* gt_mask,
                             const float* t_rand, const float* u, const ndsr_extra_params& ep, const CallParams& cp,
                             const ndsr_outputs* coarse, const ndsr_outputs* fine) {
  const ndsr_config& c = h->cfg;
  const int Sc = c.num_coarse_samples, Sf = c.num_fine_samples;
  const float near_ = std::isnan(ep.near_override) ? c.near_ : ep.near_override;
  const float far_ = std::isnan(ep.far_override) ? c.far_ : ep.far_override;
  {
    ProfScope ps(h, st, NDSR_STAGE_SAMPLE);
    NDS_CUDA(h, launch_sample_along_rays(B, Sc, near_, far_, c.use_linear_disparity,
                                         c.use_stratified_sampling ? t_rand : nullptr, h->z_coarse, st));
    h->launches++;
  }
  // coarse level always uses the configured sample_at_infinity (models.py:1509)
  const bool coarse_rgb = coarse && (coarse->rgb || wants_per_sample(coarse) || coarse->ray_norm);
  // tensor-core engine, no gradient keys on either level: the fine pass re-uses the coarse pass's narrow-network
  // results at the coarse depths (run_level)
  const bool fine_grad = fine && ((c.predict_norm && fine->target_norm) || (!c.predict_norm && fine->ray_norm));
  const bool coarse_grad = coarse && ((c.predict_norm && coarse->target_norm) || (!c.predict_norm && coarse->ray_norm));
  const bool split = h->engine == NDSR_ENGINE_TC && !ep.use_sigma_gradient && !fine_grad && !coarse_grad &&
                     !h->tc_no_split;
  int rc = run_level(h, st, 0, B, Sc, nullptr, h->z_coarse, origins, dirs, viewdirs, warp_id, gt_mask, ep, cp,
                     c.use_sample_at_infinity, coarse, h->w_coarse, coarse_rgb || coarse != nullptr,
                     split ? h->carry : nullptr, nullptr, 0, /*apply_filter=*/false);   // models.py:1493-1516: no render_opts
  if (rc) return rc;
  SamplePdfArgs sa;
  memset(&sa, 0, sizeof sa);
  sa.n_rays = B; sa.n_bins = Sc - 1; sa.n_fine = Sf; sa.n_coarse = Sc;
  sa.bins = nullptr;                       // midpoints of z_coarse (models.py:1522)
  sa.weights = h->w_coarse + 1;            // weights[..., 1:-1] (models.py:1524)
  sa.w_stride = Sc;
  sa.u = c.use_stratified_sampling ? u : nullptr;
  sa.z_coarse = h->z_coarse; sa.z_out = h->z_fine;
  sa.src_elem_out = split ? h->src_elem : nullptr;
  sa.z_samples = split ? h->z_new : nullptr;
  {
    ProfScope ps(h, st, NDSR_STAGE_RESAMPLE);
    NDS_CUDA(h, launch_sample_pdf(sa, h->num_sms, st));
    h->launches++;
  }
  const int inf_fine = ep.sample_at_infinity_override < 0 ? c.use_sample_at_infinity : ep.sample_at_infinity_override;
  return run_level(h, st, 1, B, Sc + Sf, nullptr, h->z_fine, origins, dirs, viewdirs, warp_id, gt_mask, ep, cp,
                   inf_fine, fine, nullptr, true, nullptr, split ? h->src_elem : nullptr, Sc);
}

extern "C" int ndsr_render_rays(ndsr_handle* h, void* stream, int64_t n_rays, const float* origins,
                                const float* directions, const float* viewdirs, const uint32_t* warp_id,
                                const float* gt_mask, const float* t_rand, const float* u,
                                const ndsr_extra_params* ep, const ndsr_outputs* coarse, const ndsr_outputs* fine) {
  if (!h) return NDSR_ERR_INVALID;
  if (n_rays == 0) return NDSR_OK;                  // empty batch: nothing to read, nothing to write
  int rc = check_call(h, ep, warp_id, gt_mask);
  if (rc) return rc;
  if (n_rays < 0 || !origins || !directions) return fail(h, NDSR_ERR_INVALID, "origins/directions required");
  const ndsr_config& c = h->cfg;
  if (c.use_stratified_sampling && (!t_rand || !u)) return fail(h, NDSR_ERR_INVALID, "t_rand and u required with use_stratified_sampling");
  cudaStream_t st = (cudaStream_t)stream;
  NDS_CUDA(h, cudaSetDevice(h->device));
  CallParams cp;
  make_call_params(c, *ep, cp);
  // sharpen_weights couples rays of one call (App. C-2): no internal chunking when it is requested
  const bool coupled = c.use_mask_sharp_weights && ((coarse && coarse->sharp_weights) || (fine && fine->sharp_weights));
  const int64_t chunk = coupled ? n_rays : (n_rays < h->max_chunk ? n_rays : h->max_chunk);
  if ((rc = ensure_scratch(h, chunk, st))) return rc;
  const int Sc = c.num_coarse_samples, Sf = c.num_fine_samples;
  for (int64_t r0 = 0; r0 < n_rays; r0 += chunk) {
    const int64_t B = (n_rays - r0) < chunk ? (n_rays - r0) : chunk;
    ndsr_outputs oc = offset_outputs(coarse, r0, Sc, h->H), of = offset_outputs(fine, r0, Sc + Sf, h->H);
    rc = render_rays_chunk(h, st, B, origins + r0 * 3, directions + r0 * 3, viewdirs ? viewdirs + r0 * 3 : nullptr,
                           warp_id ? warp_id + r0 : nullptr, gt_mask ? gt_mask + r0 : nullptr,
                           t_rand ? t_rand + r0 * Sc : nullptr, u ? u + r0 * Sf : nullptr, *ep, cp,
                           coarse ? &oc : nullptr, fine ? &of : nullptr);
    if (rc) return rc;
  }
  return NDSR_OK;
}

// ------------------------------------------------------ host-buffer variant
namespace {
struct Stage {
  std::vector<std::pair<float*, std::pair<float*, size_t>>> d2h;   // (host, (dev, bytes))
};
}

// device copies of the requested outputs, carved out of one persistent buffer (`base` null: only count bytes)
static size_t stage_outputs(const ndsr_outputs* host, int64_t B, int S, int H, ndsr_outputs& dev, char* base, size_t off,
                            Stage& stg) {
  memset(&dev, 0, sizeof dev);
  if (!host) return off;
#define ST(field, per)                                                                        \
  if (host->field) {                                                                          \
    const size_t bytes = (size_t)B * (per) * sizeof(float);                                   \
    if (base) { dev.field = reinterpret_cast<float*>(base + off); stg.d2h.push_back({host->field, {dev.field, bytes}}); } \
    off += (bytes + 255) & ~(size_t)255;                                                      \
  }
  ST(rgb, 3) ST(depth, 1) ST(med_depth, 1) ST(acc, 1) ST(ray_norm, 3) ST(ray_rotation_field, 3)
  ST(ray_translation_field, 3) ST(ray_delta_x, 3) ST(ray_hyper_points, H) ST(ray_predicted_mask, 1)
  ST(med_points, 3 + H) ST(z_vals, S) ST(weights, S) ST(alpha, S) ST(accum_prod, S) ST(sigma, S)
  ST(sharp_weights, S) ST(back_facing, S) ST(predicted_mask, S) ST(points, 3 * S) ST(warped_points, (3 + H) * S)
  ST(delta_x, 3 * S) ST(predicted_norm, 3 * S) ST(target_norm, 3 * S)
#undef ST
  return off;
}

// Host buffers in, host buffers out (what a ctypes caller of the reference's render_image chunk loop has).  The
// rays are processed in chunks of max_chunk; the inputs of chunk k + 1 are copied on a second stream while chunk k
// computes (pinned host memory makes the copies asynchronous), the outputs come back in one copy per key at the end.
static int render_rays_host_impl(ndsr_handle* h, void* stream, int64_t n_rays, const float* origins,
                                 const float* directions, const float* viewdirs, const uint32_t* warp_id,
                                 const float* gt_mask, const float* t_rand, const float* u, const uint32_t* key_coarse,
                                 const uint32_t* key_fine, const ndsr_extra_params* ep, const ndsr_outputs* coarse,
                                 const ndsr_outputs* fine) {
  if (!h) return NDSR_ERR_INVALID;
  if (n_rays == 0) return NDSR_OK;
  if (n_rays < 0 || !origins || !directions) return fail(h, NDSR_ERR_INVALID, "origins/directions required");
  cudaStream_t st = (cudaStream_t)stream;
  NDS_CUDA(h, cudaSetDevice(h->device));
  const ndsr_config& c = h->cfg;
  const int Sc = c.num_coarse_samples, Sf = c.num_fine_samples;
  // sharpen_weights couples the rays of one call (App. C-2): no chunking when it is requested
  const bool coupled = c.use_mask_sharp_weights && ((coarse && coarse->sharp_weights) || (fine && fine->sharp_weights));
  const int64_t chunk = coupled ? n_rays : (n_rays < h->max_chunk ? n_rays : h->max_chunk);
  if (!h->copy_stream) {
    NDS_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      NDS_CUDA(h, cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
      NDS_CUDA(h, cudaEventCreateWithFlags(&h->ev_free[i], cudaEventDisableTiming));
    }
  }
  // persistent staging (grown on demand, reused across calls); with keys the draws are generated on the device
  // straight into the staging half (the reference draws them inside the jitted call too: model_utils.py:84, 217)
  const bool gen = key_coarse && key_fine && c.use_stratified_sampling;
  const size_t per_ray = (size_t)(3 + 3 + 3 + 1 + 1 + Sc + Sf) * sizeof(float);
  const size_t half = ((size_t)chunk * per_ray + 255) & ~(size_t)255;
  if (2 * half > h->in_stage_bytes) {
    NDS_CUDA(h, cudaDeviceSynchronize());
    if (h->in_stage) cudaFree(h->in_stage);
    h->in_stage = nullptr; h->in_stage_bytes = 0;
    NDS_CUDA(h, cudaMalloc(&h->in_stage, 2 * half));
    h->in_stage_bytes = 2 * half;
  }
  Stage stg;
  ndsr_outputs dc, df;
  size_t out_need = stage_outputs(coarse, n_rays, Sc, h->H, dc, nullptr, 0, stg);
  out_need = stage_outputs(fine, n_rays, Sc + Sf, h->H, df, nullptr, out_need, stg);
  if (out_need > h->out_stage_bytes) {
    NDS_CUDA(h, cudaDeviceSynchronize());
    if (h->out_stage) cudaFree(h->out_stage);
    h->out_stage = nullptr; h->out_stage_bytes = 0;
    NDS_CUDA(h, cudaMalloc(&h->out_stage, out_need));
    h->out_stage_bytes = out_need;
  }
  size_t off = stage_outputs(coarse, n_rays, Sc, h->H, dc, (char*)h->out_stage, 0, stg);
  stage_outputs(fine, n_rays, Sc + Sf, h->H, df, (char*)h->out_stage, off, stg);
  // the copy stream must not overwrite a staging half that work already queued on `st` still reads
  NDS_CUDA(h, cudaEventRecord(h->ev_free[0], st));
  NDS_CUDA(h, cudaEventRecord(h->ev_free[1], st));
  int rc = NDSR_OK;
  int64_t k = 0;
  for (int64_t r0 = 0; r0 < n_rays && !rc; r0 += chunk, ++k) {
    const int64_t B = (n_rays - r0) < chunk ? (n_rays - r0) : chunk;
    const int b = (int)(k & 1);
    float* p = reinterpret_cast<float*>((char*)h->in_stage + (size_t)b * half);
    cudaStream_t cs = h->copy_stream;
    NDS_CUDA(h, cudaStreamWaitEvent(cs, h->ev_free[b], 0));
    auto up = [&](const void* src, size_t per, float** dst) -> cudaError_t {
      if (!src) { *dst = nullptr; return cudaSuccess; }
      *dst = p;
      p += (size_t)B * per;
      return cudaMemcpyAsync(*dst, (const char*)src + (size_t)r0 * per * sizeof(float), (size_t)B * per * sizeof(float),
                             cudaMemcpyHostToDevice, cs);
    };
    float *d_o, *d_d, *d_v, *d_w, *d_m, *d_t, *d_u;
    NDS_CUDA(h, up(origins, 3, &d_o));
    NDS_CUDA(h, up(directions, 3, &d_d));
    NDS_CUDA(h, up(viewdirs, 3, &d_v));
    NDS_CUDA(h, up(warp_id, 1, &d_w));
    NDS_CUDA(h, up(gt_mask, 1, &d_m));
    if (gen) {
      // rows [r0, r0 + B) of random.uniform(key, [n_rays, S]): generated on the copy stream as well, so that they are
      // ready together with the rays
      d_t = p; p += (size_t)B * Sc;
      d_u = p; p += (size_t)B * Sf;
      NDS_CUDA(h, launch_uniform_threefry_range(key_coarse[0], key_coarse[1], n_rays * Sc, r0 * Sc, B * Sc, d_t, h->num_sms, cs));
      NDS_CUDA(h, launch_uniform_threefry_range(key_fine[0], key_fine[1], n_rays * Sf, r0 * Sf, B * Sf, d_u, h->num_sms, cs));
      h->launches += 2;
    } else {
      NDS_CUDA(h, up(t_rand, (size_t)Sc, &d_t));
      NDS_CUDA(h, up(u, (size_t)Sf, &d_u));
    }
    NDS_CUDA(h, cudaEventRecord(h->ev_in[b], cs));
    NDS_CUDA(h, cudaStreamWaitEvent(st, h->ev_in[b], 0));
    ndsr_outputs oc = offset_outputs(coarse ? &dc : nullptr, r0, Sc, h->H), of = offset_outputs(fine ? &df : nullptr, r0, Sc + Sf, h->H);
    rc = ndsr_render_rays(h, stream, B, d_o, d_d, d_v, (const uint32_t*)d_w, d_m, d_t, d_u, ep, coarse ? &oc : nullptr,
                          fine ? &of : nullptr);
    NDS_CUDA(h, cudaEventRecord(h->ev_free[b], st));
  }
  if (!rc) {
    for (auto& e : stg.d2h)
      if (e.second.second) {
        cudaError_t ce = cudaMemcpyAsync(e.first, e.second.first, e.second.second, cudaMemcpyDeviceToHost, st);
        if (ce != cudaSuccess) { h->err = cudaGetErrorString(ce); rc = NDSR_ERR_CUDA; break; }
      }
  }
  cudaError_t se = cudaStreamSynchronize(st);
  if (!rc && se != cudaSuccess) { h->err = cudaGetErrorString(se); rc = NDSR_ERR_CUDA; }
  return rc;
}

extern "C" int ndsr_render_rays_host(ndsr_handle* h, void* stream, int64_t n_rays, const float* origins,
                                     const float* directions, const float* viewdirs, const uint32_t* warp_id,
                                     const float* gt_mask, const float* t_rand, const float* u,
                                     const ndsr_extra_params* ep, const ndsr_outputs* coarse,
                                     const ndsr_outputs* fine) {
  return render_rays_host_impl(h, stream, n_rays, origins, directions, viewdirs, warp_id, gt_mask, t_rand, u, nullptr,
                               nullptr, ep, coarse, fine);
}

extern "C" int ndsr_render_rays_host_rng(ndsr_handle* h, void* stream, int64_t n_rays, const float* origins,
                                         const float* directions, const float* viewdirs, const uint32_t* warp_id,
                                         const float* gt_mask, const uint32_t key_coarse[2], const uint32_t key_fine[2],
                                         const ndsr_extra_params* ep, const ndsr_outputs* coarse,
                                         const ndsr_outputs* fine) {
  if (!h) return NDSR_ERR_INVALID;
  if (!key_coarse || !key_fine) return fail(h, NDSR_ERR_INVALID, "ndsr_render_rays_host_rng needs both keys");
  if ((int64_t)n_rays * (h->cfg.num_coarse_samples > h->cfg.num_fine_samples ? h->cfg.num_coarse_samples : h->cfg.num_fine_samples) >= (int64_t)0xFFFFFFFFll)
    return fail(h, NDSR_ERR_INVALID, "too many draws for one 32-bit counter stream");
  return render_rays_host_impl(h, stream, n_rays, origins, directions, viewdirs, warp_id, gt_mask, nullptr, nullptr,
                               key_coarse, key_fine, ep, coarse, fine);
}

extern "C" int ndsr_random_uniform_range(int device, void* stream, const uint32_t key[2], int64_t n, int64_t first,
                                         int64_t count, float* out) {
  if (!key || n < 0 || first < 0 || count < 0 || first + count > n || (count > 0 && !out) || n >= (int64_t)0xFFFFFFFFll) return NDSR_ERR_INVALID;
  if (count == 0) return NDSR_OK;
  if (cudaSetDevice(device) != cudaSuccess) return NDSR_ERR_CUDA;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0) return NDSR_ERR_CUDA;
  return launch_uniform_threefry_range(key[0], key[1], n, first, count, out, sms, (cudaStream_t)stream) == cudaSuccess ? NDSR_OK : NDSR_ERR_CUDA;
}

// ------------------------------------------------- ray generation
extern "C" int ndsr_camera_rays(int device, void* stream, const ndsr_camera* camera, float* origins, float* directions,
                                float* pixels) {
  if (!camera || !origins || !directions || camera->image_size[0] < 0 || camera->image_size[1] < 0) return NDSR_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return NDSR_ERR_CUDA;
  return launch_camera_rays(*camera, origins, directions, pixels, (cudaStream_t)stream) == cudaSuccess ? NDSR_OK : NDSR_ERR_CUDA;
}

// ------------------------------------------------- peer-memory frame buffers (multi-GPU reassembly)
extern "C" int ndsr_peer_alloc(int device, size_t bytes, void** ptr, ndsr_ipc_handle* handle) {
  if (!ptr || !handle || bytes == 0) return NDSR_ERR_INVALID;
  static_assert(sizeof(cudaIpcMemHandle_t) == sizeof(ndsr_ipc_handle), "ipc handle size");
  if (cudaSetDevice(device) != cudaSuccess) return NDSR_ERR_CUDA;
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return NDSR_ERR_CUDA; }
  cudaIpcMemHandle_t hd;
  if (cudaMemset(p, 0, bytes) != cudaSuccess || cudaIpcGetMemHandle(&hd, p) != cudaSuccess) { cudaFree(p); return NDSR_ERR_CUDA; }
  memcpy(handle->bytes, &hd, sizeof hd);
  *ptr = p;
  return NDSR_OK;
}
extern "C" int ndsr_peer_free(int device, void* ptr) {
  if (!ptr) return NDSR_OK;
  if (cudaSetDevice(device) != cudaSuccess) return NDSR_ERR_CUDA;
  return cudaFree(ptr) == cudaSuccess ? NDSR_OK : NDSR_ERR_CUDA;
}
extern "C" int ndsr_peer_open(int device, const ndsr_ipc_handle* handle, void** ptr) {
  if (!handle || !ptr) return NDSR_ERR_INVALID;
  if (cudaSetDevice(device) != cudaSuccess) return NDSR_ERR_CUDA;
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle->bytes, sizeof hd);
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); return NDSR_ERR_CUDA; }
  *ptr = p;
  return NDSR_OK;
}
extern "C" int ndsr_peer_close(int device, void* ptr) {
  if (!ptr) return NDSR_OK;
  if (cudaSetDevice(device) != cudaSuccess) return NDSR_ERR_CUDA;
  return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? NDSR_OK : NDSR_ERR_CUDA;
}
extern "C" int ndsr_set_output_mirrors(ndsr_handle* h, int32_t n, const int64_t* byte_deltas, const void* frame_base,
                                       size_t frame_bytes) {
  if (!h) return NDSR_ERR_INVALID;
  if (n < 0 || n > NDSR_MAX_MIRRORS || (n > 0 && !byte_deltas)) return fail(h, NDSR_ERR_INVALID, "0 <= mirrors <= NDSR_MAX_MIRRORS");
  if (n > 0 && (!frame_base || frame_bytes == 0)) return fail(h, NDSR_ERR_INVALID, "mirrors need the frame buffer's address range");
  for (int m = 0; m < n; ++m) {
    if (byte_deltas[m] % 4) return fail(h, NDSR_ERR_INVALID, "mirror offsets must keep float alignment");
    h->mirror_delta[m] = byte_deltas[m];
  }
  h->n_mirror = n;
  h->mirror_base = n > 0 ? static_cast<const char*>(frame_base) : nullptr;
  h->mirror_bytes = n > 0 ? frame_bytes : 0;
  return NDSR_OK;
}

// ------------------------------------------------- jax-compatible uniform draws
extern "C" int ndsr_random_uniform(int device, void* stream, const uint32_t key[2], int64_t n, float* out) {
  if (!key || n < 0 || (n > 0 && !out) || n >= (int64_t)0xFFFFFFFFll) return NDSR_ERR_INVALID;
  if (n == 0) return NDSR_OK;
  if (cudaSetDevice(device) != cudaSuccess) return NDSR_ERR_CUDA;
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || sms <= 0) return NDSR_ERR_CUDA;
  return launch_uniform_threefry(key[0], key[1], n, out, sms, (cudaStream_t)stream) == cudaSuccess ? NDSR_OK : NDSR_ERR_CUDA;
}

// ------------------------------------------------- stand-alone stage calls
extern "C" int ndsr_sample_along_rays(ndsr_handle* h, void* stream, int64_t n_rays, int32_t n_samples, float near_,
                                      float far_, int32_t use_linear_disparity, const float* t_rand,
                                      float* z_vals) {
  if (!h || !z_vals || n_samples < 2 || n_rays < 0) return h ? fail(h, NDSR_ERR_INVALID, "bad argument") : NDSR_ERR_INVALID;
  NDS_CUDA(h, cudaSetDevice(h->device));
  NDS_CUDA(h, launch_sample_along_rays(n_rays, n_samples, near_, far_, use_linear_disparity, t_rand, z_vals,
                                       (cudaStream_t)stream));
  h->launches++;
  return NDSR_OK;
}

extern "C" int ndsr_sample_pdf(ndsr_handle* h, void* stream, int64_t n_rays, int32_t n_bins, int32_t n_fine,
                               int32_t n_coarse, const float* bins, const float* weights, const float* u,
                               const float* z_vals, float* z_out, float* z_samples, int32_t* idx_lo,
                               int32_t* idx_hi, float* cdf) {
  if (!h || !bins || !weights || !z_vals || !z_out || n_bins < 2 || n_fine < 1 || n_coarse < 0 || n_rays < 0)
    return h ? fail(h, NDSR_ERR_INVALID, "bad argument") : NDSR_ERR_INVALID;
  if (2 * n_bins + n_coarse + 2 * n_fine > 11000) return fail(h, NDSR_ERR_INVALID, "too many bins/samples");
  NDS_CUDA(h, cudaSetDevice(h->device));
  SamplePdfArgs sa;
  memset(&sa, 0, sizeof sa);
  sa.n_rays = n_rays; sa.n_bins = n_bins; sa.n_fine = n_fine; sa.n_coarse = n_coarse; sa.bins = bins;
  sa.weights = weights; sa.w_stride = n_bins - 1; sa.u = u; sa.z_coarse = z_vals; sa.z_out = z_out;
  sa.z_samples = z_samples; sa.idx_lo = idx_lo; sa.idx_hi = idx_hi; sa.cdf_out = cdf;
  NDS_CUDA(h, launch_sample_pdf(sa, h->num_sms, (cudaStream_t)stream));
  h->launches++;
  return NDSR_OK;
}

extern "C" int ndsr_volumetric_rendering(ndsr_handle* h, void* stream, int64_t n_rays, int32_t n_samples,
                                         const float* rgb, const float* sigma, const float* z_vals,
                                         const float* dirs, int32_t use_white_background,
                                         int32_t sample_at_infinity, const ndsr_outputs* out) {
  if (!h || !rgb || !sigma || !z_vals || !dirs || !out || n_samples < 1 || n_rays < 0)
    return h ? fail(h, NDSR_ERR_INVALID, "bad argument") : NDSR_ERR_INVALID;
  if (n_rays == 0) return NDSR_OK;
  cudaStream_t st = (cudaStream_t)stream;
  NDS_CUDA(h, cudaSetDevice(h->device));
  if (n_samples > h->max_samples_seen) { h->max_samples_seen = n_samples; if (h->cap_rays) { NDS_CUDA(h, cudaStreamSynchronize(st)); free_scratch(h); } }
  int rc = ensure_scratch(h, n_rays, st);
  if (rc) return rc;
  const int64_t N = n_rays * n_samples;
  NDS_CUDA(h, launch_pack_rgb_sigma(N, rgb, sigma, h->planes, N, st));
  h->launches++;
  CompositeArgs ca;
  memset(&ca, 0, sizeof ca);
  ca.n_rays = n_rays; ca.S = n_samples; ca.H = 0; ca.planes = h->planes; ca.plane_stride = N; ca.z = z_vals;
  ca.dirs = dirs; ca.viewdirs = dirs; ca.origins = nullptr; ca.points = nullptr; ca.sigma_is_activated = 1;
  ca.plane_mask = PG_RGB;
  ca.white_bkgd = use_white_background; ca.sample_at_infinity = sample_at_infinity;
  memset(&ca.out, 0, sizeof ca.out);
  ca.out.rgb = out->rgb; ca.out.depth = out->depth; ca.out.med_depth = out->med_depth; ca.out.acc = out->acc;
  ca.out.weights = out->weights; ca.out.alpha = out->alpha; ca.out.accum_prod = out->accum_prod;
  NDS_CUDA(h, launch_composite(ca, h->num_sms, st));
  h->launches++;
  return NDSR_OK;
}
