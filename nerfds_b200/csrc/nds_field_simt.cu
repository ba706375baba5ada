// fp32 CUDA-core field engine: evaluates, for a tile of 64 samples, everything
// NerfModel.render_samples computes per sample (hypernerf/models.py:893-1311):
// mask MLP -> SE(3) warp -> hyper sheet -> template trunk -> sigma/normal head
// -> bottleneck -> rgb branch, plus (optionally) -d(sigma_raw)/dx by a reverse
// sweep (models.py:1035-1077).  This is the exact-arithmetic
// engine: every output of the level dict, fp32 end to end.  The tensor-core
// engine (nds_field_tc.cu) is the fast render path.
#include "nds_common.cuh"
#include "nds_dual.cuh"

namespace nds {

constexpr int TS = 64;     // samples per tile
constexpr int NT = 256;    // threads per CTA

struct Seg { const float* p; int ld; int K; };

// out[s][n] = act(sum_k in[s][k] W[k][n] + b[n]) for 64 samples, N = 32*NPT.
// Thread (sg = tid/32, lane): samples sg*8..+7, neurons lane + 32 j.
// Reads complete before any write (in-place safe).
template <int NPT>
__device__ __forceinline__ void dense_tile(const Seg* segs, int nseg, const float* __restrict__ W,
                                           const float* __restrict__ bias, bool relu,
                                           float* out, int out_ld, const uint32_t* relu_mask_in,
                                           uint32_t* relu_mask_out, int mask_words) {
  constexpr int N = 32 * NPT;
  const int lane = threadIdx.x & 31, sg = threadIdx.x >> 5;
  float acc[8][NPT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < NPT; ++j) acc[i][j] = 0.f;
  int krow = 0;
  for (int sgi = 0; sgi < nseg; ++sgi) {
    const float* base = segs[sgi].p + (sg * 8) * segs[sgi].ld;
    const int ld = segs[sgi].ld, K = segs[sgi].K;
    for (int k = 0; k < K; ++k) {
      const float* wrow = W + (size_t)(krow + k) * N + lane;
      float wv[NPT];
#pragma unroll
      for (int j = 0; j < NPT; ++j) wv[j] = __ldg(wrow + 32 * j);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a = base[i * ld + k];
#pragma unroll
        for (int j = 0; j < NPT; ++j) acc[i][j] = fmaf(a, wv[j], acc[i][j]);
      }
    }
    krow += K;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < NPT; ++j) {
    const float b = bias ? __ldg(bias + lane + 32 * j) : 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = acc[i][j] + b;
      const int s = sg * 8 + i;
      if (relu_mask_in) {   // backward sweep: multiply by relu'(z) of the layer below
        const uint32_t m = relu_mask_in[s * mask_words + j];
        v = ((m >> lane) & 1u) ? v : 0.f;
      }
      if (relu_mask_out) {  // forward sweep: record z > 0 (jax.nn.relu's derivative, relu'(0) = 0)
        const uint32_t m = __ballot_sync(0xffffffffu, v > 0.f);
        if (lane == 0) relu_mask_out[s * mask_words + j] = m;
      }
      if (relu) v = fmaxf(v, 0.f);
      out[s * out_ld + lane + 32 * j] = v;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void dense(const Seg* segs, int nseg, const float* W, const float* bias, int N,
                                      bool relu, float* out, int out_ld,
                                      const uint32_t* mask_in = nullptr, uint32_t* mask_out = nullptr,
                                      int mask_words = 8) {
  switch (N >> 5) {
    case 1: dense_tile<1>(segs, nseg, W, bias, relu, out, out_ld, mask_in, mask_out, mask_words); break;
    case 2: dense_tile<2>(segs, nseg, W, bias, relu, out, out_ld, mask_in, mask_out, mask_words); break;
    case 4: dense_tile<4>(segs, nseg, W, bias, relu, out, out_ld, mask_in, mask_out, mask_words); break;
    default: dense_tile<8>(segs, nseg, W, bias, relu, out, out_ld, mask_in, mask_out, mask_words); break;
  }
}

// Small-N linear head: hd[s][o] = sum_k in[s][k] W[k][o] + b[o], o < n_out <= 8.
__device__ __forceinline__ void head(const float* in, int ld, const DenseW& L, float* hd) {
  const int s = threadIdx.x & (TS - 1), part = threadIdx.x >> 6;
  for (int o = part; o < L.N; o += NT / TS) {
    float acc = 0.f;
    const float* row = in + s * ld;
    for (int k = 0; k < L.K; ++k) acc = fmaf(row[k], __ldg(L.W + (size_t)k * L.N + o), acc);
    hd[s * 8 + o] = acc + __ldg(L.b + o);
  }
  __syncthreads();
}

// Forward of modules.MLP hidden stack (modules.py:57-72). X: inputs, H: hidden (in place).
__device__ __forceinline__ void mlp_forward(const MlpW& m, const float* X, int ldx, float* H, int ldh,
                                            uint32_t* masks /* [depth][TS][8] or null */) {
  for (int l = 0; l < m.depth; ++l) {
    Seg segs[2];
    int nseg;
    if (l == 0) { segs[0] = {X, ldx, m.in_dim}; nseg = 1; }
    else if (l == m.skip) { segs[0] = {H, ldh, m.width}; segs[1] = {X, ldx, m.in_dim}; nseg = 2; }
    else { segs[0] = {H, ldh, m.width}; nseg = 1; }
    const int mw = m.width >> 5;
    dense(segs, nseg, m.hidden[l].W, m.hidden[l].b, m.width, true, H, ldh, nullptr,
          masks ? masks + (size_t)l * TS * mw : nullptr, mw);
  }
}

// Reverse sweep of the hidden stack: on entry G holds d/d(h_{depth-1}) (NOT yet
// masked); on exit GX[s][0..in_dim) holds d/d(inputs).  SURVEY.md App. E step 2.
__device__ __forceinline__ void mlp_backward(const MlpW& m, const MlpWT& t, float* G, int ldg, float* GX,
                                             int ldgx, const uint32_t* masks) {
  const int tid = threadIdx.x;
  const int mw = m.width >> 5;
  // zbar_{depth-1} = hbar * relu'(z_{depth-1})
  for (int i = tid; i < TS * m.width; i += NT) {
    const int s = i / m.width, n = i % m.width;
    const uint32_t mk = masks[((size_t)(m.depth - 1) * TS + s) * mw + (n >> 5)];
    if (!((mk >> (n & 31)) & 1u)) G[s * ldg + n] = 0.f;
  }
  for (int i = tid; i < TS * t.kx_pad; i += NT) GX[(i / t.kx_pad) * ldgx + (i % t.kx_pad)] = 0.f;
  __syncthreads();
  for (int l = m.depth - 1; l >= 0; --l) {
    Seg seg = {G, ldg, m.width};
    if (l == 0 || l == m.skip) {
      // input slice: GX += zbar_l * W_l[x rows]^T  (accumulate: two contributions)
      // computed into a scratch region of GX beyond kx_pad, then added.
      float* tmp = GX + t.kx_pad;
      dense(&seg, 1, t.WTx[l], nullptr, t.kx_pad, false, tmp, ldgx);
      for (int i = tid; i < TS * t.kx_pad; i += NT) {
        const int s = i / t.kx_pad, n = i % t.kx_pad;
        GX[s * ldgx + n] += tmp[s * ldgx + n];
      }
      __syncthreads();
    }
    if (l > 0)
      dense(&seg, 1, t.WTh[l], nullptr, m.width, false, G, ldg, masks + (size_t)(l - 1) * TS * mw, nullptr, mw);
  }
}

// gradient of posenc features w.r.t. the encoded C-vector (App. E step 3)
__device__ __forceinline__ void posenc_backward(const float* x, int C, const PosencSpec& pe, const float* g,
                                                float* gx) {
  int o = 0;
  for (int c = 0; c < C; ++c) gx[c] = 0.f;
  if (pe.identity) { for (int c = 0; c < C; ++c) gx[c] += g[o + c]; o += C; }
  for (int k = 0; k < pe.num_bands; ++k) {
    const float s = exp2f((float)(pe.min_deg + k));
    const float w = pe.window[k] * s;
    for (int c = 0; c < C; ++c) {
      const float xb = x[c] * s;
      gx[c] += w * (cosf(xb) * g[o + c] + cosf(xb + NDS_HALF_PI_F) * g[o + C + c]);
    }
    o += 2 * C;
  }
}

struct Smem {
  float *X, *H, *Bn, *X2, *H2, *hd, *GX;
  uint32_t *mk_trunk, *mk_warp, *mk_hyper;
  int ldx, ldh, ldx2, ldh2, ldgx;
};

__global__ void __launch_bounds__(NT, 1)
field_simt_kernel(const __grid_constant__ ModelW M, const __grid_constant__ CallParams cp,
                  const __grid_constant__ FieldArgs a, const __grid_constant__ ndsr_config cfg,
                  int ldx, int ldh, int ldx2, int ldh2, int ldgx) {
  extern __shared__ __align__(16) float smem[];
  Smem sm;
  {
    float* p = smem;
    sm.X = p; p += TS * ldx;
    sm.H = p; p += TS * ldh;
    sm.Bn = p; p += TS * ldh;
    sm.X2 = p; p += TS * ldx2;
    sm.H2 = p; sm.GX = p;   // rgb hidden and the gradient scratch are never live together
    p += TS * ((a.need_grad && ldgx > ldh2) ? ldgx : ldh2);
    sm.hd = p; p += TS * 8;
    uint32_t* q = reinterpret_cast<uint32_t*>(p);
    sm.mk_trunk = q; q += a.need_grad ? cfg.trunk_depth * TS * (cfg.trunk_width >> 5) : 0;
    sm.mk_warp = q; q += (a.need_grad && cfg.use_warp) ? cfg.warp_depth * TS * (cfg.warp_width >> 5) : 0;
    sm.mk_hyper = q;
    sm.ldx = ldx; sm.ldh = ldh; sm.ldx2 = ldx2; sm.ldh2 = ldh2; sm.ldgx = ldgx;
  }
  const int tid = threadIdx.x;
  const LevelW& LV = M.level[a.level];
  const int H = cfg.use_hyper_sheet ? cfg.hyper_num_dims : 0;
  const bool grad = a.need_grad != 0;

  for (int64_t tile = blockIdx.x; tile * TS < a.n_samples_total; tile += gridDim.x) {
    const int64_t n = tile * TS + tid;
    const bool owner = tid < TS;
    const bool valid = owner && n < a.n_samples_total;
    // ---- per-sample state (threads 0..63) ---------------------------------
    float x[3] = {0.f, 0.f, 0.f}, xw[3], om[2] = {0.f, 0.f};
    float maskv = 0.f, pmask = 0.f;
    SE3<float> T;
    float wv_raw[6];
    int64_t ray = 0;
    uint32_t wid = 0;
    if (valid) {
      ray = n / a.S;
      if (a.points) {
        x[0] = a.points[n * 3 + 0]; x[1] = a.points[n * 3 + 1]; x[2] = a.points[n * 3 + 2];
      } else {
        const float z = a.z[n];   // origins + z * directions (model_utils.py:91-92)
        x[0] = a.origins[ray * 3 + 0] + z * a.dirs[ray * 3 + 0];
        x[1] = a.origins[ray * 3 + 1] + z * a.dirs[ray * 3 + 1];
        x[2] = a.origins[ray * 3 + 2] + z * a.dirs[ray * 3 + 2];
      }
      if (a.warp_id) wid = a.warp_id[ray];
      if (a.gt_mask) maskv = a.gt_mask[ray];
    }
    // ---- predicted mask (models.py:955-975, modules.py:394-434) -----------
    if (cfg.use_predicted_mask) {
      if (owner) {
        float* row = sm.X + tid * ldx;
        auto st = [&](int i, float v) { row[i] = v; };
        int o = posenc_emit(x, 3, cp.pe_mask, st, 0);
        for (int e = 0; e < cfg.mask_embed_dims; ++e) row[o + e] = M.mask_embed[(size_t)wid * cfg.mask_embed_dims + e];
      }
      __syncthreads();
      mlp_forward(M.mask, sm.X, ldx, sm.H, ldh, nullptr);
      head(sm.H, ldh, M.mask.logit, sm.hd);
      if (owner) {
        pmask = sm.hd[tid * 8];
        if (cfg.mask_output_relu) pmask = fmaxf(pmask, 0.f);
        const float gm = a.gt_mask ? maskv : 0.f;
        maskv = a.gt_mask ? (pmask * cp.mask_ratio + gm * (1.f - cp.mask_ratio)) : pmask * cp.mask_ratio;
      }
      __syncthreads();
    }
    // ---- SE(3) warp (warping.py:200-237) ----------------------------------
    if (cfg.use_warp) {
      if (owner) {
        float* row = sm.X + tid * ldx;
        auto st = [&](int i, float v) { row[i] = v; };
        int o = posenc_emit(x, 3, cp.pe_warp, st, 0);
        for (int e = 0; e < cfg.warp_embed_dims; ++e) row[o++] = M.warp_embed[(size_t)wid * cfg.warp_embed_dims + e];
        if (cfg.use_mask_in_warp) row[o++] = maskv;
      }
      __syncthreads();
      mlp_forward(M.warp, sm.X, ldx, sm.H, ldh, grad ? sm.mk_warp : nullptr);
      // heads w (3) and v (3): pack into hd[0..2], hd[3..5]
      {
        const int s = tid & (TS - 1), part = tid >> 6;
        for (int o = part; o < 6; o += NT / TS) {
          const DenseW& L = o < 3 ? M.warp_w : M.warp_v;
          const int oo = o < 3 ? o : o - 3;
          float acc = 0.f;
          const float* row = sm.H + s * ldh;
          for (int k = 0; k < L.K; ++k) acc = fmaf(row[k], __ldg(L.W + (size_t)k * L.N + oo), acc);
          sm.hd[s * 8 + o] = acc + __ldg(L.b + oo);
        }
        __syncthreads();
      }
      if (owner) {
        for (int i = 0; i < 6; ++i) wv_raw[i] = sm.hd[tid * 8 + i];
        exp_se3<float>(wv_raw, wv_raw + 3, T);
        for (int i = 0; i < 3; ++i)   // from_homogenous(T [x;1]) (warping.py:231-232)
          xw[i] = T.R[i * 3 + 0] * x[0] + T.R[i * 3 + 1] * x[1] + T.R[i * 3 + 2] * x[2] + T.p[i];
      }
      __syncthreads();
    } else if (owner) {
      xw[0] = x[0]; xw[1] = x[1]; xw[2] = x[2];
    }
    // ---- hyper sheet (modules.py:351-392) ---------------------------------
    if (cfg.use_hyper_sheet) {
      if (owner) {
        float* row = sm.X + tid * ldx;
        auto st = [&](int i, float v) { row[i] = v; };
        int o = posenc_emit(x, 3, cp.pe_hsheet, st, 0);
        for (int e = 0; e < cfg.warp_embed_dims; ++e) row[o++] = M.warp_embed[(size_t)wid * cfg.warp_embed_dims + e];
        if (cfg.use_mask_in_hyper) row[o++] = maskv;
      }
      __syncthreads();
      mlp_forward(M.hyper, sm.X, ldx, sm.H, ldh, grad ? sm.mk_hyper : nullptr);
      head(sm.H, ldh, M.hyper.logit, sm.hd);
      if (owner) for (int i = 0; i < H; ++i) om[i] = sm.hd[tid * 8 + i];
      __syncthreads();
    }
    // ---- template trunk (models.py:493-523, modules.py:243-286) -----------
    if (owner) {
      float* row = sm.X + tid * ldx;
      auto st = [&](int i, float v) { row[i] = v; };
      int o = posenc_emit(xw, 3, cp.pe_spatial, st, 0);
      if (H > 0) posenc_emit(om, H, cp.pe_hyperpt, st, o);
    }
    __syncthreads();
    mlp_forward(LV.trunk, sm.X, ldx, sm.H, ldh, grad ? sm.mk_trunk : nullptr);
    head(sm.H, ldh, LV.alpha, sm.hd);
    float sigma_raw = 0.f, nrm[3] = {0.f, 0.f, 0.f};
    if (owner) {
      sigma_raw = sm.hd[tid * 8];
      if (cfg.predict_norm) { nrm[0] = sm.hd[tid * 8 + 1]; nrm[1] = sm.hd[tid * 8 + 2]; nrm[2] = sm.hd[tid * 8 + 3]; }
    }
    __syncthreads();
    // ---- -d(sigma_raw)/dx by reverse sweep (models.py:1065-1077) ----------
    float gobs[3] = {0.f, 0.f, 0.f}, gh[3] = {0.f, 0.f, 0.f};
    if (grad) {
      // seed: d sigma_raw / d trunk_out = alpha kernel column 0
      for (int i = tid; i < TS * LV.trunk.width; i += NT) {
        const int s = i / LV.trunk.width, k = i % LV.trunk.width;
        sm.Bn[s * ldh + k] = __ldg(M.alpha_col0[a.level] + k);
      }
      __syncthreads();
      mlp_backward(LV.trunk, M.trunk_T[a.level], sm.Bn, ldh, sm.GX, ldgx, sm.mk_trunk);
      float gxw[3] = {0.f, 0.f, 0.f}, gom[2] = {0.f, 0.f};
      if (owner) {
        const float* g = sm.GX + tid * ldgx;
        posenc_backward(xw, 3, cp.pe_spatial, g, gxw);
        if (H > 0) posenc_backward(om, H, cp.pe_hyperpt, g + cp.pe_spatial.dim(3), gom);
      }
      __syncthreads();
      if (cfg.use_hyper_sheet) {
        // d/d(h_last) = sum_o gom[o] * Wlogit[:, o]
        const MlpW& m = M.hyper;
        if (owner) { sm.hd[tid * 8] = gom[0]; sm.hd[tid * 8 + 1] = gom[1]; }
        __syncthreads();
        for (int i = tid; i < TS * m.width; i += NT) {
          const int s = i / m.width, k = i % m.width;
          float acc = 0.f;
          for (int o = 0; o < H; ++o) acc = fmaf(sm.hd[s * 8 + o], __ldg(m.logit.W + (size_t)k * m.logit.N + o), acc);
          sm.Bn[s * ldh + k] = acc;
        }
        __syncthreads();
        mlp_backward(m, M.hyper_T, sm.Bn, ldh, sm.GX, ldgx, sm.mk_hyper);
        if (owner) {
          float gx[3];
          posenc_backward(x, 3, cp.pe_hsheet, sm.GX + tid * ldgx, gx);
          gobs[0] += gx[0]; gobs[1] += gx[1]; gobs[2] += gx[2];
        }
        __syncthreads();
      }
      if (cfg.use_warp) {
        // x' = R(w,v) x + p(w,v): direct term R^T gxw plus the path through (w_raw, v_raw)
        float gwv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (owner) {
          for (int i = 0; i < 3; ++i) gobs[i] += T.R[0 * 3 + i] * gxw[0] + T.R[1 * 3 + i] * gxw[1] + T.R[2 * 3 + i] * gxw[2];
          // forward-mode over the 6 screw inputs of the closed-form exp map
          Dual6 dw[3], dv[3];
          for (int i = 0; i < 3; ++i) { dw[i] = Dual6::var(wv_raw[i], i); dv[i] = Dual6::var(wv_raw[3 + i], 3 + i); }
          SE3<Dual6> TD;
          exp_se3<Dual6>(dw, dv, TD);
          for (int i = 0; i < 3; ++i) {
            Dual6 xi = TD.R[i * 3 + 0] * x[0] + TD.R[i * 3 + 1] * x[1] + TD.R[i * 3 + 2] * x[2] + TD.p[i];
            for (int q = 0; q < 6; ++q) gwv[q] += gxw[i] * xi.d[q];
          }
          for (int q = 0; q < 6; ++q) sm.hd[tid * 8 + q] = gwv[q];
        }
        __syncthreads();
        const MlpW& m = M.warp;
        for (int i = tid; i < TS * m.width; i += NT) {
          const int s = i / m.width, k = i % m.width;
          float acc = 0.f;
          for (int o = 0; o < 3; ++o) {
            acc = fmaf(sm.hd[s * 8 + o], __ldg(M.warp_w.W + (size_t)k * 3 + o), acc);
            acc = fmaf(sm.hd[s * 8 + 3 + o], __ldg(M.warp_v.W + (size_t)k * 3 + o), acc);
          }
          sm.Bn[s * ldh + k] = acc;
        }
        __syncthreads();
        mlp_backward(m, M.warp_T, sm.Bn, ldh, sm.GX, ldgx, sm.mk_warp);
        if (owner) {
          float gx[3];
          posenc_backward(x, 3, cp.pe_warp, sm.GX + tid * ldgx, gx);
          gobs[0] += gx[0]; gobs[1] += gx[1]; gobs[2] += gx[2];
        }
        __syncthreads();
      } else if (owner) {
        gobs[0] += gxw[0]; gobs[1] += gxw[1]; gobs[2] += gxw[2];
      }
    }
    if (grad && owner) {
      float g[3] = {-gobs[0], -gobs[1], -gobs[2]};
      normalize3(g, gh);                                           // models.py:1070,1077
    }
    float rgb[3] = {0.f, 0.f, 0.f};
    if (!a.sigma_only) {
      // bottleneck (modules.py:254-257): only when there is an rgb condition
      const float* first = sm.H;
      if (cfg.use_viewdirs) {
        Seg seg = {sm.H, ldh, LV.bottleneck.K};
        dense(&seg, 1, LV.bottleneck.W, LV.bottleneck.b, LV.bottleneck.N, false, sm.Bn, ldh);
        first = sm.Bn;
      }
      int vdim = 0, ndim = 0;
      if (owner) {
        float* row = sm.X2 + tid * ldx2;
        auto st = [&](int i, float v) { row[i] = v; };
        int o = 0;
        if (cfg.use_viewdirs) {
          float vd[3] = {0.f, 0.f, 0.f};
          if (valid) { vd[0] = a.viewdirs[ray * 3]; vd[1] = a.viewdirs[ray * 3 + 1]; vd[2] = a.viewdirs[ray * 3 + 2]; }
          o = posenc_emit(vd, 3, cp.pe_view, st, 0);
        }
        if (cp.use_predicted_norm || cp.use_sigma_gradient) {
          float nh[3], ni[3];
          if (cp.use_sigma_gradient) {          // models.py:1107-1112
            ni[0] = gh[0]; ni[1] = gh[1]; ni[2] = gh[2];
          } else {
            // normalize -> R^T n (map_vectors inverse, models.py:1124-1127)
            normalize3(nrm, nh);
            if (cfg.use_warp) {
              for (int i = 0; i < 3; ++i) ni[i] = T.R[0 * 3 + i] * nh[0] + T.R[1 * 3 + i] * nh[1] + T.R[2 * 3 + i] * nh[2];
            } else { ni[0] = nh[0]; ni[1] = nh[1]; ni[2] = nh[2]; }
          }
          normalize3(ni, nh);                   // models.py:1138
          if (cfg.norm_input_posenc) posenc_emit(nh, 3, cp.pe_norm, st, o);
          else { row[o] = nh[0]; row[o + 1] = nh[1]; row[o + 2] = nh[2]; }
        }
      }
      vdim = cfg.use_viewdirs ? cp.pe_view.dim(3) : 0;
      ndim = (cp.use_predicted_norm || cp.use_sigma_gradient) ? (cfg.norm_input_posenc ? cp.pe_norm.dim(3) : 3) : 0;
      __syncthreads();
      // rgb branch (modules.py:288-313): [first | viewfeat | trunk_out (App. C-1) | normfeat]
      Seg segs[4];
      int ns = 0;
      segs[ns++] = {first, ldh, LV.trunk.width};
      if (vdim) segs[ns++] = {sm.X2, ldx2, vdim};
      if (cfg.use_x_in_rgb_condition) segs[ns++] = {sm.H, ldh, LV.trunk.width};
      if (ndim) segs[ns++] = {sm.X2 + vdim, ldx2, ndim};
      const float* last = nullptr;
      int last_ld = 0;
      if (LV.rgb.depth > 0) {
        dense(segs, ns, LV.rgb.hidden[0].W, LV.rgb.hidden[0].b, LV.rgb.width, true, sm.H2, ldh2);
        for (int l = 1; l < LV.rgb.depth; ++l) {
          Seg s1 = {sm.H2, ldh2, LV.rgb.width};
          dense(&s1, 1, LV.rgb.hidden[l].W, LV.rgb.hidden[l].b, LV.rgb.width, true, sm.H2, ldh2);
        }
        last = sm.H2; last_ld = ldh2;
        head(last, last_ld, LV.rgb.logit, sm.hd);
      } else {
        // depth-0 branch: logit straight on the concatenated input (not in any shipped gin)
        const int s = tid & (TS - 1), part = tid >> 6;
        for (int o = part; o < 3; o += NT / TS) {
          float acc = 0.f; int kr = 0;
          for (int g = 0; g < ns; ++g) {
            const float* row = segs[g].p + s * segs[g].ld;
            for (int k = 0; k < segs[g].K; ++k) acc = fmaf(row[k], __ldg(LV.rgb.logit.W + (size_t)(kr + k) * 3 + o), acc);
            kr += segs[g].K;
          }
          sm.hd[s * 8 + o] = acc + __ldg(LV.rgb.logit.b + o);
        }
        __syncthreads();
      }
      if (owner) for (int i = 0; i < 3; ++i) rgb[i] = 1.f / (1.f + expf(-sm.hd[tid * 8 + i]));   // nn.sigmoid
      __syncthreads();
    }
    // ---- write planes ------------------------------------------------------
    if (valid) {
      float* P = a.planes;
      const int64_t ps = a.plane_stride;
      P[P_SIGMA_RAW * ps + n] = sigma_raw;
      for (int i = 0; i < 3; ++i) {
        P[(P_RGB + i) * ps + n] = rgb[i];
        P[(P_NORM + i) * ps + n] = nrm[i];
        P[(P_WARPED + i) * ps + n] = xw[i];
      }
      for (int i = 0; i < H; ++i) P[(P_WARPED + 3 + i) * ps + n] = om[i];
      P[P_MASK * ps + n] = pmask;
      if (cfg.use_warp) {
        // rotation / translation visualisations (models.py:1291-1302)
        const float r = 0.57735025882720947265625f;   // normalize_vector(ones)
        float rf[3], rn[3];
        for (int i = 0; i < 3; ++i) rf[i] = T.R[i * 3 + 0] * r + T.R[i * 3 + 1] * r + T.R[i * 3 + 2] * r;
        normalize3(rf, rn);
        for (int i = 0; i < 3; ++i) { P[(P_ROT + i) * ps + n] = rn[i]; P[(P_TRANS + i) * ps + n] = T.p[i]; }
      }
      if (grad) {
        float gr[3], grn[3];
        if (cfg.use_warp) {
          for (int i = 0; i < 3; ++i) gr[i] = T.R[i * 3 + 0] * gh[0] + T.R[i * 3 + 1] * gh[1] + T.R[i * 3 + 2] * gh[2];
        } else { gr[0] = gh[0]; gr[1] = gh[1]; gr[2] = gh[2]; }
        normalize3(gr, grn);                                       // models.py:1276-1277
        for (int i = 0; i < 3; ++i) { P[(P_GRAD + i) * ps + n] = gh[i]; P[(P_TNORM + i) * ps + n] = grn[i]; }
      }
    }
    __syncthreads();
  }
}

static inline int odd(int v) { return v | 1; }

size_t field_simt_smem_bytes(const ndsr_config& c, int max_in, int max_w, int x2, bool grad, int* ld) {
  const int ldx = odd(max_in), ldh = odd(max_w), ldx2 = odd(x2 > 0 ? x2 : 1), ldh2 = odd(c.rgb_width);
  int kxp = 32;
  while (kxp < max_in) kxp *= 2;
  const int ldgx = odd(2 * kxp);
  ld[0] = ldx; ld[1] = ldh; ld[2] = ldx2; ld[3] = ldh2; ld[4] = ldgx;
  size_t f = (size_t)TS * (ldx + 2 * ldh + ldx2 + ((grad && ldgx > ldh2) ? ldgx : ldh2) + 8);
  size_t u = 0;
  if (grad) {
    u += (size_t)c.trunk_depth * TS * (c.trunk_width >> 5);
    if (c.use_warp) u += (size_t)c.warp_depth * TS * (c.warp_width >> 5);
    if (c.use_hyper_sheet) u += (size_t)c.hyper_sheet_depth * TS * (c.hyper_sheet_width >> 5);
  }
  return (f + u) * 4;
}

cudaError_t launch_field_simt(const ModelW& M, const CallParams& cp, const FieldArgs& a, const ndsr_config& cfg,
                              int max_in, int max_w, int x2, int num_sms, cudaStream_t stream) {
  int ld[5];
  const size_t smem = field_simt_smem_bytes(cfg, max_in, max_w, x2, a.need_grad != 0, ld);
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(field_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const int64_t tiles = (a.n_samples_total + TS - 1) / TS;
  const int grid = (int)(tiles < (int64_t)num_sms * 4 ? tiles : (int64_t)num_sms * 4);
  if (grid == 0) return cudaSuccess;
  field_simt_kernel<<<grid, NT, smem, stream>>>(M, cp, a, cfg, ld[0], ld[1], ld[2], ld[3], ld[4]);
  return cudaGetLastError();
}

}  // namespace nds
