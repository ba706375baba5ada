// Per-ray stages of the path: stratified sampling, alpha compositing and the
// per-ray accumulations of render_samples, and the inverse-CDF resampler.
// THIS FILE IS COMPILED WITH -fmad=false: the discrete results (inverse-CDF
// bin indices, median-depth index) depend on fp32 rounding, so every scan here
// is sequential left-to-right with separately rounded mul/add -- the order the
// oracle documents.  File:line citations refer to /root/reference.
#include "nds_common.cuh"
#include "nds_composite.h"

namespace nds {

// ---------------------------------------------------------------------------
// model_utils.sample_along_rays (model_utils.py:55-92)
// ---------------------------------------------------------------------------
__device__ __forceinline__ float lin_z(int i, int S, float near_, float far_, int lindisp) {
  // jnp.linspace(0, 1, S)[i] = i / (S-1), last element exactly 1
  const float t = (i == S - 1) ? 1.f : (float)i / (float)(S - 1);
  if (!lindisp) return near_ * (1.f - t) + far_ * t;
  return 1.f / (1.f / near_ * (1.f - t) + 1.f / far_ * t);
}

__global__ void sample_along_rays_kernel(int64_t total, int S, float near_, float far_, int lindisp,
                                         const float* __restrict__ t_rand, float* __restrict__ z) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= total) return;
  const int i = (int)(n % S);
  const float zi = lin_z(i, S, near_, far_, lindisp);
  if (!t_rand) { z[n] = zi; return; }
  const float lower = (i == 0) ? zi : .5f * (zi + lin_z(i - 1, S, near_, far_, lindisp));
  const float upper = (i == S - 1) ? zi : .5f * (lin_z(i + 1, S, near_, far_, lindisp) + zi);
  z[n] = lower + (upper - lower) * t_rand[n];
}

// ---------------------------------------------------------------------------
// volumetric_rendering + per-ray accumulations (model_utils.py:95-159,
// 272-317; models.py:1235-1246, 1323-1415).  One warp per ray.
// ---------------------------------------------------------------------------

__device__ __forceinline__ float softplus_f(float x) {
  // flax nn.softplus = jnp.logaddexp(x, 0)
  return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int CW = 4;   // warps per CTA

__global__ void __launch_bounds__(CW * 32, 8)
composite_kernel(const __grid_constant__ CompositeArgs a) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.S;
  float* s_alpha = sm + (size_t)warp * 3 * S;
  float* s_w = s_alpha + S;
  float* s_T = s_w + S;
  const int64_t ps = a.plane_stride;
  for (int64_t ray = (int64_t)blockIdx.x * CW + warp; ray < a.n_rays; ray += (int64_t)gridDim.x * CW) {
    const int64_t base = ray * S;
    // where the planes of sample s live (split fine pass: two dense blocks, see CompositeArgs::src_elem)
    auto pidx = [&](int s) -> int64_t {
      if (!a.src_elem) return base + s;
      const int e = a.src_elem[base + s];
      return e < a.n_carried ? ray * a.n_carried + e
                             : a.n_rays * a.n_carried + ray * (S - a.n_carried) + (e - a.n_carried);
    };
    const float dx = a.dirs[ray * 3], dy = a.dirs[ray * 3 + 1], dz = a.dirs[ray * 3 + 2];
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);   // jnp.linalg.norm(dirs)
    const float last = a.sample_at_infinity ? 1e10f : 1e-19f;
    float alpha_inf_last = 0.f;
    const float fox = a.origins ? a.origins[ray * 3] : 0.f, foy = a.origins ? a.origins[ray * 3 + 1] : 0.f,
                foz = a.origins ? a.origins[ray * 3 + 2] : 0.f;
    // filter_sigma's keep factor (1 or 0) of sample s for the value v it thresholds (models.py:56-63)
    auto keep = [&](int s, float v) -> float {
      float m = 1.f;
      if (a.filter_flags & 1) m = (v >= a.dust_threshold) ? 1.f : 0.f;
      if (a.filter_flags & 2) {
        float px, py, pz;
        if (a.points) { px = a.points[(base + s) * 3]; py = a.points[(base + s) * 3 + 1]; pz = a.points[(base + s) * 3 + 2]; }
        else { const float zs = a.z[base + s]; px = fox + zs * dx; py = foy + zs * dy; pz = foz + zs * dz; }
        const bool in = px >= a.bbox[0] && px <= a.bbox[1] && py >= a.bbox[2] && py <= a.bbox[3] && pz >= a.bbox[4] && pz <= a.bbox[5];
        m = in ? m : 0.f;
      }
      return m;
    };
    bool sg_done = false;
    float best_w = -1.f; int best_i = 0;
    if (a.filter_flags && a.weights_sg && !a.sigma_is_activated) {
      // sharp weights under render_opts: the RAW sigma is filtered, then activated (models.py:1236-1237), and
      // cal_weights always places the last sample at infinity (model_utils.py:162)
      for (int s = lane; s < S; s += 32) {
        const float raw = a.planes[P_SIGMA_RAW * ps + pidx(s)];
        const float sg = softplus_f(keep(s, raw) * raw);
        const float dist = ((s == S - 1) ? 1e10f : (a.z[base + s + 1] - a.z[base + s])) * dnorm;
        s_alpha[s] = 1.f - expf(-sg * dist);
      }
      __syncwarp();
      if (lane == 0) {
        float T = 1.f;
        for (int s = 0; s < S; ++s) { const float al = s_alpha[s]; s_w[s] = al * T; T = T * (1.f - al + 1e-10f); }
      }
      __syncwarp();
      for (int s = lane; s < S; s += 32) {
        const float wsg = s_w[s];
        a.weights_sg[base + s] = wsg;
        if (wsg > best_w) { best_w = wsg; best_i = s; }
      }
      __syncwarp();
      sg_done = true;
    }
    for (int s = lane; s < S; s += 32) {
      const float raw = a.planes[P_SIGMA_RAW * ps + pidx(s)];
      const float sigma = a.sigma_is_activated ? raw : softplus_f(raw);
      const float zs = a.z[base + s];
      const float dist = ((s == S - 1) ? last : (a.z[base + s + 1] - zs)) * dnorm;
      const float sig_f = a.filter_flags ? keep(s, sigma) * sigma : sigma;      // models.py:1288
      s_alpha[s] = 1.f - expf(-sig_f * dist);
      if (s == S - 1) alpha_inf_last = 1.f - expf(-sig_f * (1e10f * dnorm));
      if (a.out.sigma) a.out.sigma[base + s] = sigma;
    }
    __syncwarp();
    // sequential exclusive cumprod / cumsum (lane 0) -- the documented order
    int med_idx = 0;
    float med_found = 0.f;
    if (lane == 0) {
      float T = 1.f, cum = 0.f;
      bool found = false;
      for (int s = 0; s < S; ++s) {
        const float al = s_alpha[s];
        const float w = al * T;
        s_T[s] = T;
        s_w[s] = w;
        cum = cum + w;
        if (!found && cum >= 0.5f) { found = true; med_idx = s; }
        T = T * (1.f - al + 1e-10f);
      }
      med_found = found ? 1.f : 0.f;
    }
    med_idx = __shfl_sync(0xffffffffu, med_idx, 0);
    med_found = __shfl_sync(0xffffffffu, med_found, 0);
    __syncwarp();
    // per-ray reductions
    float r[3] = {0, 0, 0}, depth = 0, acc = 0, acc_m1 = 0, rn[3] = {0, 0, 0}, rr[3] = {0, 0, 0},
          rt[3] = {0, 0, 0}, rdx[3] = {0, 0, 0}, rh[2] = {0, 0}, rm = 0;
    const float vx = a.viewdirs[ray * 3], vy = a.viewdirs[ray * 3 + 1], vz = a.viewdirs[ray * 3 + 2];
    const float ox = a.origins ? a.origins[ray * 3] : 0.f, oy = a.origins ? a.origins[ray * 3 + 1] : 0.f,
                oz = a.origins ? a.origins[ray * 3 + 2] : 0.f;
    for (int s = lane; s < S; s += 32) {
      const int64_t n = base + s;
      const float w = s_w[s];
      const float zs = a.z[n];
      const int64_t pn = pidx(s);
      float c[3] = {0, 0, 0}, nv[3] = {0, 0, 0}, wp[5] = {0, 0, 0, 0, 0}, px[3];
      if (a.plane_mask & PG_RGB) for (int i = 0; i < 3; ++i) c[i] = a.planes[(P_RGB + i) * ps + pn];
      if (a.plane_mask & PG_WARPED) for (int i = 0; i < 3 + a.H; ++i) wp[i] = a.planes[(P_WARPED + i) * ps + pn];
      if (a.points) { px[0] = a.points[n * 3]; px[1] = a.points[n * 3 + 1]; px[2] = a.points[n * 3 + 2]; }
      else { px[0] = ox + zs * dx; px[1] = oy + zs * dy; px[2] = oz + zs * dz; }
      for (int i = 0; i < 3; ++i) r[i] += w * c[i];
      depth += w * zs;
      acc += w;
      if (s < S - 1) acc_m1 += w;
      if (a.has_norm && (a.plane_mask & PG_NORM)) {
        for (int i = 0; i < 3; ++i) { nv[i] = a.planes[(P_NORM + i) * ps + pn]; rn[i] += w * nv[i]; }
        if (a.out.back_facing) {
          const float bf = fmaxf(nv[0] * vx + nv[1] * vy + nv[2] * vz, 0.f);
          a.out.back_facing[n] = bf * bf;
        }
        if (a.out.predicted_norm) for (int i = 0; i < 3; ++i) a.out.predicted_norm[n * 3 + i] = nv[i];
      } else if (!a.has_norm && a.has_grad) {
        for (int i = 0; i < 3; ++i) rn[i] += w * a.planes[(P_GRAD + i) * ps + pn];   // models.py:1353
      }
      if (a.has_grad && a.has_norm && a.out.target_norm)
        for (int i = 0; i < 3; ++i) a.out.target_norm[n * 3 + i] = a.planes[(P_TNORM + i) * ps + pn];
      if (a.has_warp && (a.plane_mask & PG_ROT)) for (int i = 0; i < 3; ++i) rr[i] += w * a.planes[(P_ROT + i) * ps + pn];
      if (a.has_warp && (a.plane_mask & PG_TRANS)) for (int i = 0; i < 3; ++i) rt[i] += w * a.planes[(P_TRANS + i) * ps + pn];
      for (int i = 0; i < 3; ++i) {
        const float d = wp[i] - px[i];
        rdx[i] += w * d;
        if (a.out.delta_x) a.out.delta_x[n * 3 + i] = d;
        if (a.out.points) a.out.points[n * 3 + i] = px[i];
      }
      for (int i = 0; i < a.H; ++i) rh[i] += w * wp[3 + i];
      if (a.out.warped_points) for (int i = 0; i < 3 + a.H; ++i) a.out.warped_points[n * (3 + a.H) + i] = wp[i];
      if (a.has_mask && (a.plane_mask & PG_MASK)) {
        const float m = a.planes[P_MASK * ps + pn];
        rm += w * m;
        if (a.out.predicted_mask) a.out.predicted_mask[n] = m;
      }
      if (a.out.weights) a.out.weights[n] = w;
      if (a.out.alpha) a.out.alpha[n] = s_alpha[s];
      if (a.out.accum_prod) a.out.accum_prod[n] = s_T[s];
      if (a.out.z_vals) a.out.z_vals[n] = zs;
      // cal_weights() weights (always sample_at_infinity, model_utils.py:162) for sharpen_weights
      if (!sg_done) {
        const float wsg = (s == S - 1) ? alpha_inf_last * s_T[s] : w;
        if (a.weights_sg) a.weights_sg[n] = wsg;
        if (wsg > best_w) { best_w = wsg; best_i = s; }
      }
    }
    // argmax (first occurrence) across lanes
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ow = __shfl_xor_sync(0xffffffffu, best_w, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ow > best_w || (ow == best_w && oi < best_i)) { best_w = ow; best_i = oi; }
    }
    for (int i = 0; i < 3; ++i) { r[i] = warp_sum(r[i]); rn[i] = warp_sum(rn[i]); rr[i] = warp_sum(rr[i]);
                                  rt[i] = warp_sum(rt[i]); rdx[i] = warp_sum(rdx[i]); }
    depth = warp_sum(depth); acc = warp_sum(acc); acc_m1 = warp_sum(acc_m1); rm = warp_sum(rm);
    rh[0] = warp_sum(rh[0]); rh[1] = warp_sum(rh[1]);
    if (lane == 0) {
      const ndsr_outputs& o = a.out;
      // per-ray results go to the caller's buffer and to its mirrors on the peer GPUs
      auto put = [&](float* dst, float v) {
        *dst = v;
        for (int m = 0; m < a.n_mirror; ++m) *reinterpret_cast<float*>(reinterpret_cast<char*>(dst) + a.mirror_delta[m]) = v;
      };
      if (a.white_bkgd) for (int i = 0; i < 3; ++i) r[i] = r[i] + (1.f - acc);
      if (o.rgb) for (int i = 0; i < 3; ++i) put(o.rgb + ray * 3 + i, r[i]);
      if (o.depth) put(o.depth + ray, depth);
      if (o.med_depth) put(o.med_depth + ray, med_found != 0.f ? a.z[base + med_idx] : 0.f);
      if (o.acc) put(o.acc + ray, a.sample_at_infinity ? acc_m1 : acc);
      if (o.ray_norm && (a.has_norm || a.has_grad)) for (int i = 0; i < 3; ++i) put(o.ray_norm + ray * 3 + i, rn[i]);
      if (o.ray_rotation_field && a.has_warp) for (int i = 0; i < 3; ++i) put(o.ray_rotation_field + ray * 3 + i, rr[i]);
      if (o.ray_translation_field && a.has_warp) for (int i = 0; i < 3; ++i) put(o.ray_translation_field + ray * 3 + i, rt[i]);
      if (o.ray_delta_x) for (int i = 0; i < 3; ++i) put(o.ray_delta_x + ray * 3 + i, rdx[i]);
      if (o.ray_hyper_points) for (int i = 0; i < a.H; ++i) put(o.ray_hyper_points + ray * a.H + i, rh[i]);
      if (o.ray_predicted_mask && a.has_mask) put(o.ray_predicted_mask + ray, rm);
      if (o.med_points && (a.plane_mask & PG_WARPED)) for (int i = 0; i < 3 + a.H; ++i)
        put(o.med_points + ray * (3 + a.H) + i, a.planes[(P_WARPED + i) * ps + pidx(med_idx)]);
      if (a.argmax_idx) a.argmax_idx[ray] = (float)best_i;
    }
    __syncwarp();
  }
}

// model_utils.sharpen_weights (model_utils.py:180-190) INCLUDING the row-gather
// quirk: the gaussian of ray b is centred on z_vals[clamp(argmax_b, B-1)][s]
// (a *ray* index).  One warp per ray.
__global__ void __launch_bounds__(CW * 32)
sharpen_weights_kernel(int64_t n_rays, int S, const float* __restrict__ weights_sg, const float* __restrict__ z,
                       const float* __restrict__ argmax_idx, float stdv, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t ray = (int64_t)blockIdx.x * CW + warp; ray < n_rays; ray += (int64_t)gridDim.x * CW) {
    int64_t src = (int64_t)argmax_idx[ray];
    if (src > n_rays - 1) src = n_rays - 1;
    float tot = 0.f;
    const float scale_sq = stdv * stdv, log_norm = logf(6.28318530717958647692f * scale_sq);
    for (int s = lane; s < S; s += 32) {
      // jax.scipy.stats.norm.pdf = exp((log(2 pi scale^2) + (x - loc)^2 / scale^2) / -2)
      const float d = z[ray * S + s] - z[src * S + s];
      const float g = expf((log_norm + (d * d) / scale_sq) / -2.0f);
      const float v = weights_sg[ray * S + s] * g;
      out[ray * S + s] = v;
      tot += v;
    }
    tot = warp_sum(tot);
    for (int s = lane; s < S; s += 32) out[ray * S + s] = out[ray * S + s] / tot;
  }
}

// ---------------------------------------------------------------------------
// Early-termination scan (TerminationArgs, nds_composite.h).  One warp per ray: lane l owns the contiguous sorted
// positions [l c, (l + 1) c), c = ceil(S / 32); ranks and exclusive products over the lanes in front come from
// shuffle scans.  alpha of an evaluated sample is computed exactly as composite_kernel will compute it (same planes,
// same distances), so the bound is a bound on the very product the compositing step forms.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(CW * 32)
termination_scan_kernel(const __grid_constant__ TerminationArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.S, nc = a.n_carried, nn = S - nc;
  const int chunk = (S + 31) / 32;
  const int64_t ps = a.plane_stride;
  const float last = a.sample_at_infinity ? 1e10f : 1e-19f;
  // adaptive rounds: did an earlier round take everything, or does this one?
  int rank_lo = a.rank_lo, rank_hi = a.rank_hi;
  if (a.round_stats && a.round > 0) {
    unsigned long long ev = 0, seen = 0;
    int first_merge = -1;
    for (int r = 1; r <= a.round && first_merge < 0; ++r) {
      ev += a.round_stats[2 * (r - 1)]; seen += a.round_stats[2 * (r - 1) + 1];
      if ((double)ev >= (double)a.merge_frac * (double)seen) first_merge = r;
    }
    if (first_merge >= 0 && first_merge < a.round) return;      // all ranks were handled by that round
    if (first_merge == a.round) rank_hi = nn;
  }
  for (int64_t ray = (int64_t)blockIdx.x * CW + warp; ray < a.n_rays; ray += (int64_t)gridDim.x * CW) {
    const int64_t base = ray * S;
    const float dx = a.dirs[ray * 3], dy = a.dirs[ray * 3 + 1], dz = a.dirs[ray * 3 + 2];
    const float dnorm = sqrtf(dx * dx + dy * dy + dz * dz);
    const int s0 = min(S, lane * chunk), s1 = min(S, s0 + chunk);
    // rank (among the new depths, in sorted order) of the first new depth of this lane's chunk
    int n_new = 0;
    for (int s = s0; s < s1; ++s) n_new += a.src_elem[base + s] >= nc ? 1 : 0;
    int rank0 = n_new;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, rank0, o);
      if (lane >= o) rank0 += up;
    }
    rank0 -= n_new;
    // (1 - alpha + 1e-10) of sample s if its sigma is known: a carried sample, or a new depth of an earlier round
    // (a skipped one holds sigma_raw = -1e30: alpha 0); 1 otherwise
    auto factor = [&](int s, int e, int rank) -> float {
      int64_t pn;
      if (e < nc) pn = ray * nc + e;
      else if (rank < rank_lo) pn = a.n_rays * nc + ray * nn + (e - nc);
      else return 1.f;
      const float sigma = softplus_f(a.planes[P_SIGMA_RAW * ps + pn]);
      const float dist = ((s == S - 1) ? last : (a.z[base + s + 1] - a.z[base + s])) * dnorm;
      const float al = 1.f - expf(-sigma * dist);
      return 1.f - al + 1e-10f;
    };
    float prod = 1.f;
    {
      int rank = rank0;
      for (int s = s0; s < s1; ++s) {
        const int e = a.src_elem[base + s];
        prod *= factor(s, e, rank);
        rank += e >= nc ? 1 : 0;
      }
    }
    float incl = prod;                          // inclusive product scan over the lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl *= up;
    }
    float T0 = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) T0 = 1.f;
    // decide the new depths of this round
    int kept = 0;
    {
      float T = T0;
      int rank = rank0;
      for (int s = s0; s < s1; ++s) {
        const int e = a.src_elem[base + s];
        if (e >= nc && rank >= rank_lo && rank < rank_hi && T >= a.eps) ++kept;
        T *= factor(s, e, rank);
        rank += e >= nc ? 1 : 0;
      }
    }
    int off = kept;                             // inclusive sum scan -> exclusive offsets
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, off, o);
      if (lane >= o) off += up;
    }
    const int total = __shfl_sync(0xffffffffu, off, 31);
    int dst = 0;
    if (lane == 0) {
      dst = atomicAdd(a.n_active, total);
      atomicAdd(a.stats, (unsigned long long)total);
      atomicAdd(a.stats + 1, (unsigned long long)(min(rank_hi, nn) - min(rank_lo, nn)));
      if (a.round_stats) {
        atomicAdd(a.round_stats + 2 * a.round, (unsigned long long)total);
        atomicAdd(a.round_stats + 2 * a.round + 1, (unsigned long long)(min(rank_hi, nn) - min(rank_lo, nn)));
      }
    }
    dst = __shfl_sync(0xffffffffu, dst, 0) + off - kept;
    float T = T0;
    int rank = rank0;
    for (int s = s0; s < s1; ++s) {
      const int e = a.src_elem[base + s];
      const bool mine = e >= nc && rank >= rank_lo && rank < rank_hi;
      const float Tb = T;
      T *= factor(s, e, rank);
      rank += e >= nc ? 1 : 0;
      if (!mine) continue;
      const int64_t li = ray * nn + (e - nc);   // element of the dense new-depth list
      if (Tb >= a.eps) { a.index[dst++] = (int32_t)li; continue; }
      const int64_t pn = a.n_rays * nc + li;     // its planes: an empty sample
      a.planes[P_SIGMA_RAW * ps + pn] = -1e30f;
      if (a.plane_mask & PG_RGB) for (int i = 0; i < 3; ++i) a.planes[(P_RGB + i) * ps + pn] = 0.f;
      if (a.plane_mask & PG_NORM) for (int i = 0; i < 3; ++i) a.planes[(P_NORM + i) * ps + pn] = 0.f;
      if (a.plane_mask & PG_MASK) a.planes[P_MASK * ps + pn] = 0.f;
      if (a.plane_mask & PG_WARPED) for (int i = 0; i < 3 + a.H; ++i) a.planes[(P_WARPED + i) * ps + pn] = 0.f;
      if (a.has_warp && (a.plane_mask & PG_ROT)) for (int i = 0; i < 3; ++i) a.planes[(P_ROT + i) * ps + pn] = 0.f;
      if (a.has_warp && (a.plane_mask & PG_TRANS)) for (int i = 0; i < 3; ++i) a.planes[(P_TRANS + i) * ps + pn] = 0.f;
    }
  }
}

// rgb [n,S,3] + sigma [n,S] -> planes, for the stand-alone volumetric_rendering entry point
__global__ void pack_rgb_sigma_kernel(int64_t total, const float* __restrict__ rgb, const float* __restrict__ sigma,
                                      float* planes, int64_t ps) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= total) return;
  planes[P_SIGMA_RAW * ps + n] = sigma[n];
  for (int i = 0; i < 3; ++i) planes[(P_RGB + i) * ps + n] = rgb[n * 3 + i];
  for (int i = 0; i < 5; ++i) planes[(P_WARPED + i) * ps + n] = 0.f;
}

// ---------------------------------------------------------------------------
// model_utils.piecewise_constant_pdf + sample_pdf (model_utils.py:193-269).
// One warp per ray.  n = n_bins = len(bins) = len(cdf); weights has n-1 entries.
// ---------------------------------------------------------------------------

__global__ void __launch_bounds__(CW * 32)
sample_pdf_kernel(const __grid_constant__ SamplePdfArgs a) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = a.n_bins, nf = a.n_fine, nc = a.n_coarse, nt = nc + nf;
  float* s_cdf = sm + (size_t)warp * (2 * n + nt + nf);
  float* s_bins = s_cdf + n;
  float* s_all = s_bins + n;
  const float eps = 1e-5f;
  for (int64_t ray = (int64_t)blockIdx.x * CW + warp; ray < a.n_rays; ray += (int64_t)gridDim.x * CW) {
    for (int j = lane; j < n - 1; j += 32) s_cdf[j + 1] = a.weights[ray * a.w_stride + j] + eps;   // weights += eps
    for (int j = lane; j < n; j += 32)
      s_bins[j] = a.bins ? a.bins[ray * n + j]
                         : .5f * (a.z_coarse[ray * nc + j + 1] + a.z_coarse[ray * nc + j]);   // models.py:1522
    for (int j = lane; j < nc; j += 32) s_all[j] = a.z_coarse[ray * nc + j];
    __syncwarp();
    if (lane == 0) {
      float tot = 0.f;
      for (int j = 1; j < n; ++j) tot = tot + s_cdf[j];          // weights.sum(-1), sequential
      float c = 0.f;
      s_cdf[0] = 0.f;
      for (int j = 1; j < n; ++j) { c = c + s_cdf[j] / tot; s_cdf[j] = c; }   // cumsum(pdf)
    }
    __syncwarp();
    if (a.cdf_out) for (int j = lane; j < n; j += 32) a.cdf_out[ray * n + j] = s_cdf[j];
    for (int j = lane; j < nf; j += 32) {
      float uj;
      if (a.u) uj = a.u[ray * nf + j];
      else uj = (j == nf - 1) ? 1.f : (float)j / (float)(nf - 1);
      // k = #{i : cdf_i <= u}  (mask = u >= cdf, model_utils.py:223); cdf is non-decreasing
      int lo_b = 0, hi_b = n;
      while (lo_b < hi_b) { const int mid = (lo_b + hi_b) >> 1; if (s_cdf[mid] <= uj) lo_b = mid + 1; else hi_b = mid; }
      const int k = lo_b;
      int lo = k - 1; lo = lo < 0 ? 0 : (lo > n - 2 ? n - 2 : lo);     // max-select then min with x[-2]
      int hi = k; hi = hi < 1 ? 1 : (hi > n - 1 ? n - 1 : hi);         // min-select then max with x[1]
      const float c0 = s_cdf[lo], c1 = s_cdf[hi], b0 = s_bins[lo], b1 = s_bins[hi];
      float denom = c1 - c0;
      denom = denom < eps ? 1.f : denom;
      const float t = (uj - c0) / denom;
      const float zs = b0 + t * (b1 - b0);
      s_all[nc + j] = zs;
      if (a.z_samples) a.z_samples[ray * nf + j] = zs;
      if (a.idx_lo) a.idx_lo[ray * nf + j] = lo;
      if (a.idx_hi) a.idx_hi[ray * nf + j] = hi;
    }
    __syncwarp();
    // jnp.sort(concat([z_vals, z_samples])), stable.  The coarse depths normally arrive sorted (stratified bins), so
    // the union is a merge: rank the nf new samples among themselves (counting, 4 values per lane against one
    // broadcast read per step), then two binary searches give every element's place.  Ties keep concat order:
    // coarse before new, lower index first.  Unsorted coarse depths fall back to counting over the whole union.
    bool sorted = true;
    for (int i = lane; i + 1 < nc; i += 32) sorted = sorted && (s_all[i] <= s_all[i + 1]);
    sorted = __all_sync(0xffffffffu, sorted);
    if (sorted) {
      float* s_new = s_all + nt;                          // the new samples in sorted order
      constexpr int PER = 4;
      for (int i0 = 0; i0 < nf; i0 += 32 * PER) {
        float v[PER]; int rk[PER];
#pragma unroll
        for (int q = 0; q < PER; ++q) { const int i = i0 + q * 32 + lane; v[q] = i < nf ? s_all[nc + i] : 0.f; rk[q] = 0; }
        for (int j = 0; j < nf; ++j) {
          const float o = s_all[nc + j];
#pragma unroll
          for (int q = 0; q < PER; ++q) rk[q] += (o < v[q] || (o == v[q] && j < i0 + q * 32 + lane)) ? 1 : 0;
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < PER; ++q) {
          const int i = i0 + q * 32 + lane;
          if (i < nf) {
            int lo_b = 0, hi_b = nc;                      // #{coarse <= v}
            while (lo_b < hi_b) { const int mid = (lo_b + hi_b) >> 1; if (s_all[mid] <= v[q]) lo_b = mid + 1; else hi_b = mid; }
            const int rank = rk[q] + lo_b;
            a.z_out[ray * nt + rank] = v[q];
            if (a.src_elem_out) a.src_elem_out[ray * nt + rank] = nc + i;
            s_new[rk[q]] = v[q];
          }
        }
      }
      __syncwarp();
      for (int i = lane; i < nc; i += 32) {
        const float v = s_all[i];
        int lo_b = 0, hi_b = nf;                          // #{new < v}
        while (lo_b < hi_b) { const int mid = (lo_b + hi_b) >> 1; if (s_new[mid] < v) lo_b = mid + 1; else hi_b = mid; }
        a.z_out[ray * nt + i + lo_b] = v;
        if (a.src_elem_out) a.src_elem_out[ray * nt + i + lo_b] = i;
      }
    } else {
      for (int i = lane; i < nt; i += 32) {
        const float v = s_all[i];
        int rank = 0;
        for (int j = 0; j < nt; ++j) {
          const float o = s_all[j];
          rank += (o < v || (o == v && j < i)) ? 1 : 0;
        }
        a.z_out[ray * nt + rank] = v;
        if (a.src_elem_out) a.src_elem_out[ray * nt + rank] = i;
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------
// datasets/core.py:51-76 camera_to_rays: Camera.pixels_to_rays over the pixel centres
// (camera.py:226-270), float32 like the reference's default camera dtype, no FMA
// contraction (this file is compiled with -fmad=false).
// ---------------------------------------------------------------------------
__global__ void camera_rays_kernel(const __grid_constant__ ndsr_camera cam, float* __restrict__ origins,
                                   float* __restrict__ dirs, float* __restrict__ pixels) {
  const int W = cam.image_size[0], Hh = cam.image_size[1];
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)W * Hh) return;
  const float px = (float)(i % W) + 0.5f, py = (float)(i / W) + 0.5f;     // get_pixel_centers (camera.py:364-368)
  const float sx = cam.focal_length, sy = cam.focal_length * cam.pixel_aspect_ratio;
  float y = (py - cam.principal_point[1]) / sy;                            // pixel_to_local_rays (camera.py:228-230)
  float x = (px - cam.principal_point[0] - y * cam.skew) / sx;
  const float k1 = cam.radial_distortion[0], k2 = cam.radial_distortion[1], k3 = cam.radial_distortion[2];
  const float p1 = cam.tangential_distortion[0], p2 = cam.tangential_distortion[1];
  if (k1 != 0.f || k2 != 0.f || k3 != 0.f || p1 != 0.f || p2 != 0.f) {
    // _radial_and_tangential_undistort (camera.py:75-106): 10 Newton steps from the distorted point
    const float xd = x, yd = y;
    for (int it = 0; it < 10; ++it) {
      // _compute_residual_and_jacobian (camera.py:27-72), same association as the numpy expressions
      const float r = x * x + y * y;
      const float d = 1.0f + r * (k1 + r * (k2 + k3 * r));
      const float fx = d * x + 2.f * p1 * x * y + p2 * (r + 2.f * x * x) - xd;
      const float fy = d * y + 2.f * p2 * x * y + p1 * (r + 2.f * y * y) - yd;
      const float d_r = k1 + r * (2.0f * k2 + 3.0f * k3 * r);
      const float d_x = 2.0f * x * d_r, d_y = 2.0f * y * d_r;
      const float fx_x = d + d_x * x + 2.0f * p1 * y + 6.0f * p2 * x;
      const float fx_y = d_y * x + 2.0f * p1 * x + 2.0f * p2 * y;
      const float fy_x = d_x * y + 2.0f * p2 * y + 2.0f * p1 * x;
      const float fy_y = d + d_y * y + 2.0f * p2 * x + 6.0f * p1 * y;
      const float den = fy_x * fx_y - fx_x * fy_y;
      const float xn = fx * fy_y - fy * fx_y, yn = fy * fx_x - fx * fy_x;
      const bool ok = fabsf(den) > 1e-9f;
      x = x + (ok ? xn / den : 0.f);
      y = y + (ok ? yn / den : 0.f);
    }
  }
  float n = sqrtf(x * x + y * y + 1.f);                                     // camera.py:242-243
  const float lx = x / n, ly = y / n, lz = 1.f / n;
  const float* R = cam.orientation;                                         // orientation.T @ local (camera.py:263)
  float wx = R[0] * lx + R[3] * ly + R[6] * lz;
  float wy = R[1] * lx + R[4] * ly + R[7] * lz;
  float wz = R[2] * lx + R[5] * ly + R[8] * lz;
  n = sqrtf(wx * wx + wy * wy + wz * wz);                                   // camera.py:267
  dirs[i * 3] = wx / n; dirs[i * 3 + 1] = wy / n; dirs[i * 3 + 2] = wz / n;
  origins[i * 3] = cam.position[0]; origins[i * 3 + 1] = cam.position[1]; origins[i * 3 + 2] = cam.position[2];
  if (pixels) { pixels[i * 2] = px; pixels[i * 2 + 1] = py; }
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
static int grid_for(int64_t rays, int num_sms) {
  int64_t g = (rays + CW - 1) / CW;
  const int64_t cap = (int64_t)num_sms * 16;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

cudaError_t launch_sample_along_rays(int64_t n_rays, int S, float near_, float far_, int lindisp,
                                     const float* t_rand, float* z, cudaStream_t st) {
  const int64_t total = n_rays * S;
  if (total == 0) return cudaSuccess;
  sample_along_rays_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, S, near_, far_, lindisp, t_rand, z);
  return cudaGetLastError();
}

cudaError_t launch_composite(const CompositeArgs& a, int num_sms, cudaStream_t st) {
  if (a.n_rays == 0) return cudaSuccess;
  const size_t smem = (size_t)CW * 3 * a.S * sizeof(float);
  if (smem > 48 * 1024) {     // long rays: opt in to the large carve-out (up to 227 KB)
    cudaError_t e = cudaFuncSetAttribute(composite_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  composite_kernel<<<grid_for(a.n_rays, num_sms), CW * 32, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_termination_scan(const TerminationArgs& a, int num_sms, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(a.n_active, 0, sizeof(int32_t), st);
  if (e != cudaSuccess || a.n_rays == 0) return e;
  termination_scan_kernel<<<grid_for(a.n_rays, num_sms), CW * 32, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_sharpen(int64_t n_rays, int S, const float* wsg, const float* z, const float* argmax_idx,
                           float stdv, float* out, int num_sms, cudaStream_t st) {
  if (n_rays == 0) return cudaSuccess;
  sharpen_weights_kernel<<<grid_for(n_rays, num_sms), CW * 32, 0, st>>>(n_rays, S, wsg, z, argmax_idx, stdv, out);
  return cudaGetLastError();
}

cudaError_t launch_pack_rgb_sigma(int64_t total, const float* rgb, const float* sigma, float* planes, int64_t ps,
                                  cudaStream_t st) {
  if (total == 0) return cudaSuccess;
  pack_rgb_sigma_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, rgb, sigma, planes, ps);
  return cudaGetLastError();
}

cudaError_t launch_camera_rays(const ndsr_camera& cam, float* origins, float* dirs, float* pixels, cudaStream_t st) {
  const int64_t n = (int64_t)cam.image_size[0] * cam.image_size[1];
  if (n <= 0) return cudaSuccess;
  camera_rays_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cam, origins, dirs, pixels);
  return cudaGetLastError();
}

// ------------------------------------------------- jax.random.uniform
// Threefry-2x32, 20 rounds, with jax's key schedule and round grouping (jax/_src/prng.py threefry2x32: five groups
// of four rounds, rotations {13,15,26,6} / {17,29,16,24}, ks2 = k0 ^ k1 ^ 0x1BD11BDA).
__device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
#define NDS_TF_ROUND(r) { x0 += x1; x1 = __funnelshift_l(x1, x1, r); x1 ^= x0; }
#define NDS_TF_R0 NDS_TF_ROUND(13) NDS_TF_ROUND(15) NDS_TF_ROUND(26) NDS_TF_ROUND(6)
#define NDS_TF_R1 NDS_TF_ROUND(17) NDS_TF_ROUND(29) NDS_TF_ROUND(16) NDS_TF_ROUND(24)
  x0 += k0; x1 += k1;
  NDS_TF_R0 x0 += k1; x1 += k2 + 1u;
  NDS_TF_R1 x0 += k2; x1 += k0 + 2u;
  NDS_TF_R0 x0 += k0; x1 += k1 + 3u;
  NDS_TF_R1 x0 += k1; x1 += k2 + 4u;
  NDS_TF_R0 x0 += k2; x1 += k0 + 5u;
#undef NDS_TF_R1
#undef NDS_TF_R0
#undef NDS_TF_ROUND
}

// random.uniform(key, shape) for float32 in [0, 1): counters iota(n) (padded with one 0 when n is odd) are split in
// two halves that form the two words of each block; block j yields elements j and j + half; 23 random mantissa bits
// under exponent 0 give [1, 2), minus 1.  One thread per block: both stores are coalesced.
__global__ void __launch_bounds__(256) uniform_threefry_kernel(uint32_t k0, uint32_t k1, int64_t n, int64_t half,
                                                               float* __restrict__ out) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < half; j += (int64_t)gridDim.x * blockDim.x) {
    uint32_t x0 = (uint32_t)j;
    uint32_t x1 = (j + half < n) ? (uint32_t)(j + half) : 0u;
    threefry2x32(k0, k1, x0, x1);
    out[j] = __uint_as_float((x0 >> 9) | 0x3F800000u) - 1.0f;
    if (j + half < n) out[j + half] = __uint_as_float((x1 >> 9) | 0x3F800000u) - 1.0f;
  }
}

// elements [first, first + count) of the same stream of n draws (a ray chunk's rows of a [n_rays, S] array): element i
// is word 0 of block i (i < half) or word 1 of block i - half
__global__ void __launch_bounds__(256) uniform_threefry_range_kernel(uint32_t k0, uint32_t k1, int64_t n, int64_t half,
                                                                     int64_t first, int64_t count, float* __restrict__ out) {
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = first + t, j = i < half ? i : i - half;
    uint32_t x0 = (uint32_t)j;
    uint32_t x1 = (j + half < n) ? (uint32_t)(j + half) : 0u;
    threefry2x32(k0, k1, x0, x1);
    out[t] = __uint_as_float(((i < half ? x0 : x1) >> 9) | 0x3F800000u) - 1.0f;
  }
}

cudaError_t launch_uniform_threefry_range(uint32_t k0, uint32_t k1, int64_t n, int64_t first, int64_t count, float* out,
                                          int num_sms, cudaStream_t st) {
  if (count <= 0) return cudaSuccess;
  const int64_t half = (n + 1) / 2;
  int64_t blocks = (count + 255) / 256;
  const int64_t cap = (int64_t)num_sms * 8 * 4;
  if (blocks > cap) blocks = cap;
  uniform_threefry_range_kernel<<<(unsigned)blocks, 256, 0, st>>>(k0, k1, n, half, first, count, out);
  return cudaGetLastError();
}

cudaError_t launch_uniform_threefry(uint32_t k0, uint32_t k1, int64_t n, float* out, int num_sms, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const int64_t half = (n + 1) / 2;
  int64_t blocks = (half + 255) / 256;
  const int64_t cap = (int64_t)num_sms * 8 * 4;          // 8 resident CTAs per SM, 4 waves, then grid-stride
  if (blocks > cap) blocks = cap;
  uniform_threefry_kernel<<<(unsigned)blocks, 256, 0, st>>>(k0, k1, n, half, out);
  return cudaGetLastError();
}

cudaError_t launch_sample_pdf(const SamplePdfArgs& a, int num_sms, cudaStream_t st) {
  if (a.n_rays == 0) return cudaSuccess;
  const size_t smem = (size_t)CW * (2 * a.n_bins + a.n_coarse + 2 * a.n_fine) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(sample_pdf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  sample_pdf_kernel<<<grid_for(a.n_rays, num_sms), CW * 32, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace nds
