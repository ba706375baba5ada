"""CPU ORACLE for the NeRF-DS ray-marching path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``nerfds_b200/``
imports it, and the product fails loudly when its CUDA library is missing.

PARITY: PINNED AGAINST THE REFERENCE'S OWN SOURCE for the forward pass, with one
exception.  The reference (JokerYan/NeRF-DS @ f0bd3844, JAX/Flax) ships no
tests, golden vectors or fixtures for this path, and jax/flax/gin are not
installable in this environment (no wheels, no network).  Instead,
tools/make_golden.py imports the UNMODIFIED reference files (model_utils.py,
rigid_body.py, modules.py, warping.py, models.py) under thin stand-ins for
``jax`` (jax.numpy -> numpy with jax's float32 / int32 promotion and array
immutability, vmap -> a Python loop over pytrees, random.uniform -> injected
draws) and ``flax.linen`` (module tree, setup / compact naming, Dense, Embed)
and freezes, in tests/golden/reference_shim.npz:
  * the array-level functions (posenc, windows, sample_along_rays,
    volumetric_rendering, cal_weights, sharpen_weights, compute_depth_*,
    piecewise_constant_pdf, sample_pdf, skew, exp_so3, exp_se3, ...);
  * the Flax modules (MLP, NerfMLP, HyperSheetMLP, MaskMLP, GLOEmbed,
    SE3Field) applied to the product's parameter pytree;
  * the whole ``NerfModel.__call__`` under nerf_ds.gin's bindings (reduced
    widths / sample counts): every key of both level dicts.
tests/test_oracle_golden.py compares this oracle with all of them.
``target_norm`` (which needs ``jax.value_and_grad``) is pinned too: the
stand-in's ``value_and_grad`` is a float64 central difference (h = 1e-6) of
the reference's own per-point sigma function, fed back through the
reference's own post-processing (negate, normalise, rotate); this oracle's
``torch.autograd.grad`` result agrees to ~1e-6 (median 5e-7; one of 456
points sits on a ReLU kink).

Deliberate, documented choices where the reference leaves the result to XLA:
  * reductions that feed *discrete* results (pdf normalisation and cdf in
    ``piecewise_constant_pdf``, the cumulative sum in ``compute_depth_index``,
    the transmittance cumprod) are sequential left-to-right in the working
    dtype, no FMA contraction -- the CUDA path uses the same order;
  * ``jnp.linspace(0, 1, S)`` is ``arange(S) / (S-1)`` with the last element
    forced to 1 (jax 0.3.15 lax_numpy.linspace);
  * the five identical SE(3)-field evaluations the reference performs per
    sample (models.py:1037,1126,1276,1294,1300) are evaluated once (XLA's CSE
    would do the same; results are identical) -- SURVEY.md App. C-5;
  * the uniform draws ``t_rand`` / ``u`` are explicit inputs (the reference
    draws them from flax ``make_rng`` streams, models.py:1489,1524).

All file:line citations are relative to /root/reference/.
"""
from __future__ import annotations

import math
from typing import Any, Dict, Optional

import numpy as np
import torch

_EPS_F32 = float(np.finfo(np.float32).eps)


# --------------------------------------------------------------------------
# model_utils.py
# --------------------------------------------------------------------------
def _seq_cumsum(x: torch.Tensor) -> torch.Tensor:
  """Left-to-right cumulative sum along the last axis IN THE WORKING DTYPE.

  torch.cumsum on CPU accumulates float32 inputs in double; numpy's
  add.accumulate is the plain sequential fp32 loop the documented order needs.
  """
  if x.dtype == torch.float32:
    return torch.from_numpy(np.cumsum(x.detach().numpy(), axis=-1, dtype=np.float32))
  return torch.cumsum(x, dim=-1)


def _seq_cumprod(x: torch.Tensor) -> torch.Tensor:
  if x.dtype == torch.float32:
    return torch.from_numpy(np.cumprod(x.detach().numpy(), axis=-1, dtype=np.float32))
  return torch.cumprod(x, dim=-1)


def posenc_window(min_deg, max_deg, alpha, dtype=torch.float32):
  """hypernerf/model_utils.py:420-436."""
  bands = torch.arange(min_deg, max_deg, dtype=dtype)
  a = torch.as_tensor(alpha, dtype=dtype)
  x = torch.clip(a - bands, 0.0, 1.0)
  pi = torch.tensor(math.pi, dtype=dtype)
  return 0.5 * (1 + torch.cos(pi * x + pi))


def posenc(x, min_deg, max_deg, use_identity=False, alpha=None):
  """hypernerf/model_utils.py:398-417.  Layout (F, 2, C) flattened."""
  batch_shape = x.shape[:-1]
  scales = (2.0 ** torch.arange(min_deg, max_deg, dtype=torch.float64)).to(x.dtype)
  xb = x[..., None, :] * scales[:, None]                       # (*, F, C)
  half_pi = torch.tensor(0.5 * math.pi, dtype=x.dtype)
  four_feat = torch.sin(torch.stack([xb, xb + half_pi], dim=-2))  # (*, F, 2, C)
  if alpha is not None:
    window = posenc_window(min_deg, max_deg, alpha, x.dtype)
    four_feat = window[..., None, None] * four_feat
  four_feat = four_feat.reshape((*batch_shape, -1))
  if use_identity:
    return torch.cat([x, four_feat], dim=-1)
  return four_feat


def normalize_vector(v):
  """hypernerf/model_utils.py:438-442."""
  eps = torch.tensor(_EPS_F32, dtype=v.dtype)
  return v / torch.sqrt(torch.maximum(torch.sum(v ** 2, dim=-1, keepdim=True), eps))


def linspace01(n, dtype):
  t = torch.arange(n, dtype=dtype) / torch.tensor(float(n - 1), dtype=dtype)
  t[-1] = 1.0
  return t


def sample_along_rays(t_rand, origins, directions, num_coarse_samples, near,
                      far, use_stratified_sampling, use_linear_disparity):
  """hypernerf/model_utils.py:55-92; ``t_rand`` replaces random.uniform(key)."""
  dtype = origins.dtype
  batch_size = origins.shape[0]
  t_vals = linspace01(num_coarse_samples, dtype)
  near_t = torch.tensor(near, dtype=dtype)
  far_t = torch.tensor(far, dtype=dtype)
  if not use_linear_disparity:
    z_vals = near_t * (1. - t_vals) + far_t * t_vals
  else:
    z_vals = 1. / (1. / near_t * (1. - t_vals) + 1. / far_t * t_vals)
  if use_stratified_sampling:
    mids = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
    upper = torch.cat([mids, z_vals[..., -1:]], -1)
    lower = torch.cat([z_vals[..., :1], mids], -1)
    z_vals = lower + (upper - lower) * t_rand
  else:
    z_vals = z_vals[None, :].expand(batch_size, num_coarse_samples).contiguous()
  return z_vals, (origins[..., None, :] + z_vals[..., :, None] * directions[..., None, :])


def _exclusive_cumprod(x):
  ones = torch.ones_like(x[..., :1])
  return torch.cat([ones, _seq_cumprod(x[..., :-1])], dim=-1)


def cal_weights(sigma, z_vals, dirs, sample_at_infinity=True, eps=1e-10, scale=1):
  """hypernerf/model_utils.py:162-177."""
  last_sample_z = 1e10 if sample_at_infinity else 1e-19
  dists = torch.cat([z_vals[..., 1:] - z_vals[..., :-1],
                     torch.full_like(z_vals[..., :1], last_sample_z)], -1)
  dists = dists * torch.linalg.norm(dirs[..., None, :], dim=-1)
  alpha = 1.0 - torch.exp(-scale * sigma * dists)
  accum_prod = _exclusive_cumprod(1.0 - alpha + eps)
  return alpha * accum_prod


def sharpen_weights(weights, z_vals, std=0.01):
  """hypernerf/model_utils.py:180-190, *including* the row-gather quirk:
  ``z_vals[max_weights_idx]`` indexes the RAY axis with a sample index
  (SURVEY.md App. C-2); jax clamps out-of-range gather indices."""
  max_idx = torch.argmax(weights, dim=1)
  max_idx = torch.clamp(max_idx, max=z_vals.shape[0] - 1)
  max_z = z_vals[max_idx]                                       # (B, S) rows!
  std_t = torch.as_tensor(std, dtype=weights.dtype)
  # jax.scipy.stats.norm.pdf = exp(logpdf) with logpdf = (log(2 pi scale^2) + (x - loc)^2 / scale^2) / -2, all in
  # the working dtype: the deep tail underflows where THIS form does (pinned by tests/test_oracle_golden.py)
  scale_sq = std_t * std_t
  log_norm = torch.log(torch.as_tensor(2 * math.pi, dtype=weights.dtype) * scale_sq)
  g = torch.exp((log_norm + (z_vals - max_z) ** 2 / scale_sq) / -2.0)
  sharp = weights * g
  return sharp / torch.sum(sharp, dim=1)[..., None]


def compute_opaqueness_mask(weights, depth_threshold=0.5):
  """hypernerf/model_utils.py:272-293."""
  cum = _seq_cumsum(weights)
  opaq = cum >= torch.tensor(depth_threshold, dtype=weights.dtype)
  padded = torch.cat([torch.zeros_like(opaq[..., :1]), opaq[..., :-1]], dim=-1)
  return torch.logical_xor(opaq, padded).to(weights.dtype)


def compute_depth_index(weights, depth_threshold=0.5):
  """hypernerf/model_utils.py:296-299."""
  return torch.argmax(compute_opaqueness_mask(weights, depth_threshold), dim=-1)


def compute_depth_map(weights, z_vals, depth_threshold=0.5):
  """hypernerf/model_utils.py:302-317."""
  return torch.sum(compute_opaqueness_mask(weights, depth_threshold) * z_vals, dim=-1)


def volumetric_rendering(rgb, sigma, z_vals, dirs, use_white_background,
                         sample_at_infinity=True, eps=1e-10):
  """hypernerf/model_utils.py:95-159 (use_sharp_weights branch is fenced off
  by NerfDSConfig.validate: use_rgb_sharp_weights=False everywhere)."""
  last_sample_z = 1e10 if sample_at_infinity else 1e-19
  dists = torch.cat([z_vals[..., 1:] - z_vals[..., :-1],
                     torch.full_like(z_vals[..., :1], last_sample_z)], -1)
  dists = dists * torch.linalg.norm(dirs[..., None, :], dim=-1)
  alpha = 1.0 - torch.exp(-sigma * dists)
  accum_prod = _exclusive_cumprod(1.0 - alpha + eps)
  weights = alpha * accum_prod
  rgb_out = (weights[..., None] * rgb).sum(dim=-2)
  exp_depth = (weights * z_vals).sum(dim=-1)
  med_depth = compute_depth_map(weights, z_vals)
  acc = weights.sum(dim=-1)
  if use_white_background:
    rgb_out = rgb_out + (1. - acc[..., None])
  if sample_at_infinity:
    acc = weights[..., :-1].sum(dim=-1)
  return {'rgb': rgb_out, 'depth': exp_depth, 'med_depth': med_depth,
          'acc': acc, 'weights': weights, 'alpha': alpha,
          'accum_prod': accum_prod}


def piecewise_constant_pdf(u, bins, weights, return_indices=False):
  """hypernerf/model_utils.py:193-241, literal mask/max/min inverse CDF.

  ``u`` replaces random.uniform(key) / linspace.  With ``return_indices`` also
  returns (lo, hi): the positions in ``bins``/``cdf`` that the max/min
  selections picked -- the "sample indices" the CUDA path must match
  bit-exactly given identical (bins, weights, u).
  """
  eps = 1e-5
  weights = weights + eps
  total = _seq_cumsum(weights)[..., -1:]                         # weights.sum(-1)
  pdf = weights / total
  cdf = _seq_cumsum(pdf)
  cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], dim=-1)
  mask = (u[..., None, :] >= cdf[..., :, None])                 # (B, n, Sf)

  def minmax(x):
    x0 = torch.max(torch.where(mask, x[..., None], x[..., :1, None]), dim=-2).values
    x1 = torch.min(torch.where(~mask, x[..., None], x[..., -1:, None]), dim=-2).values
    x0 = torch.minimum(x0, x[..., -2:-1])
    x1 = torch.maximum(x1, x[..., 1:2])
    return x0, x1

  bins_g0, bins_g1 = minmax(bins)
  cdf_g0, cdf_g1 = minmax(cdf)
  denom = cdf_g1 - cdf_g0
  denom = torch.where(denom < eps, torch.ones_like(denom), denom)
  t = (u - cdf_g0) / denom
  z_samples = bins_g0 + t * (bins_g1 - bins_g0)
  if not return_indices:
    return z_samples
  n = cdf.shape[-1]
  k = mask.sum(dim=-2)                                          # #{j: cdf_j <= u}
  lo = torch.clamp(k - 1, 0, n - 2)
  hi = torch.clamp(k, 1, n - 1)
  return z_samples, lo, hi, cdf


def sample_pdf(u, bins, weights, origins, directions, z_vals):
  """hypernerf/model_utils.py:244-269."""
  z_samples = piecewise_constant_pdf(u, bins, weights)
  z_vals = torch.sort(torch.cat([z_vals, z_samples], dim=-1), dim=-1).values
  return z_vals, (origins[..., None, :] + z_vals[..., None] * directions[..., None, :])


# --------------------------------------------------------------------------
# rigid_body.py
# --------------------------------------------------------------------------
def skew(w):
  """hypernerf/rigid_body.py:27-41 (batched over leading axes)."""
  z = torch.zeros_like(w[..., 0])
  return torch.stack([
      torch.stack([z, -w[..., 2], w[..., 1]], -1),
      torch.stack([w[..., 2], z, -w[..., 0]], -1),
      torch.stack([-w[..., 1], w[..., 0], z], -1)], -2)


def exp_so3(w, theta):
  """hypernerf/rigid_body.py:59-74."""
  W = skew(w)
  eye = torch.eye(3, dtype=w.dtype)
  th = theta[..., None, None]
  return eye + torch.sin(th) * W + (1.0 - torch.cos(th)) * (W @ W)


def exp_se3_rp(S, theta):
  """hypernerf/rigid_body.py:77-101: returns (R, p) before the
  rotation_only / inverse post-processing."""
  w, v = S[..., :3], S[..., 3:]
  W = skew(w)
  R = exp_so3(w, theta)
  eye = torch.eye(3, dtype=S.dtype)
  th = theta[..., None, None]
  M = th * eye + (1.0 - torch.cos(th)) * W + (th - torch.sin(th)) * (W @ W)
  p = (M @ v[..., None])[..., 0]
  return R, p


def se3_apply(R, p, x, rotation_only=False, inverse=False):
  """rigid_body.py:96-101 + warping.py:231-232 (homogeneous apply; w == 1)."""
  if rotation_only:
    p = p * 0
  if inverse:
    p = -(R.transpose(-1, -2) @ p[..., None])[..., 0]
    R = R.transpose(-1, -2)
  return (R @ x[..., None])[..., 0] + p


def filter_sigma(points, sigma, render_opts):
  """models.filter_sigma (hypernerf/models.py:38-66): dust threshold and
  bounding box as 0/1 factors on sigma; keys other than these are ignored."""
  if render_opts is None:
    return sigma
  if 'dust_threshold' in render_opts:
    dust_thres = render_opts.get('dust_threshold', 0.0)
    sigma = (sigma >= dust_thres).to(sigma.dtype) * sigma
  if 'bounding_box' in render_opts:
    xmin, xmax, ymin, ymax, zmin, zmax = render_opts['bounding_box']
    render_mask = ((points[..., 0] >= xmin) & (points[..., 0] <= xmax)
                   & (points[..., 1] >= ymin) & (points[..., 1] <= ymax)
                   & (points[..., 2] >= zmin) & (points[..., 2] <= zmax))
    sigma = render_mask.to(sigma.dtype) * sigma
  return sigma


# --------------------------------------------------------------------------
# modules.py
# --------------------------------------------------------------------------
def _mm(x, kernel, tag=None):
  """nn.Dense contraction.  ``tag`` names the layer; precision-emulation
  studies (tests/precision_study.py) swap this function out."""
  return x @ kernel


def mlp_apply(p, x, depth, skips, out_relu=False, mm=_mm, tag=''):
  """modules.MLP.__call__ (hypernerf/modules.py:57-83), relu hidden."""
  inputs = x
  for i in range(depth):
    layer = p[f'hidden_{i}']
    if i in skips:
      x = torch.cat([x, inputs], dim=-1)
    x = torch.relu(mm(x, layer['kernel'], f'{tag}/hidden_{i}') + layer['bias'])
  if 'logit' in p:
    x = mm(x, p['logit']['kernel'], f'{tag}/logit') + p['logit']['bias']
    if out_relu:
      x = torch.relu(x)
  return x


def _to_torch_tree(tree, dtype):
  if isinstance(tree, dict):
    return {k: _to_torch_tree(v, dtype) for k, v in tree.items()}
  return torch.as_tensor(np.asarray(tree), dtype=dtype)


# --------------------------------------------------------------------------
# models.py
# --------------------------------------------------------------------------
class OracleNerfModel:
  """CPU restatement of ``NerfModel`` (hypernerf/models.py:71-1565).

  ``cfg`` is any object with the attribute names of the reference's
  NerfModel / SE3Field / HyperSheetMLP / MaskMLP (see
  nerfds_b200/config.py); ``params`` the Flax-layout pytree.
  """

  def __init__(self, cfg, params, dtype=torch.float32, row_chunk=1 << 16):
    self.cfg = cfg
    self.dtype = dtype
    self.P = _to_torch_tree(params, dtype)
    self.row_chunk = row_chunk
    self.mm = _mm

  # ---- helpers -----------------------------------------------------------
  def _t(self, x):
    return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(self.dtype)

  def _warp_rp(self, x, warp_in_embed, extra_params):
    """SE3Field.warp up to exp_se3 (hypernerf/warping.py:209-225)."""
    c = self.cfg
    pe = posenc(x, c.warp_min_deg, c.warp_max_deg, c.warp_use_posenc_identity,
                alpha=extra_params['warp_alpha'])
    inputs = torch.cat([pe, warp_in_embed], dim=-1)
    pw = self.P['warp_field']
    h = mlp_apply(pw['trunk'], inputs, c.warp_trunk_depth, c.warp_skips,
                  mm=self.mm, tag='warp')
    w = self.mm(h, pw['branches_w']['logit']['kernel'], 'warp/w') + pw['branches_w']['logit']['bias']
    v = self.mm(h, pw['branches_v']['logit']['kernel'], 'warp/v') + pw['branches_v']['logit']['bias']
    theta = torch.linalg.norm(w, dim=-1)
    w = w / theta[..., None]
    v = v / theta[..., None]
    screw = torch.cat([w, v], dim=-1)
    R, p = exp_se3_rp(screw, theta)
    return R, p, screw

  def _sigma_field(self, level, x, warp_embed, mask, extra_params):
    """cal_single_pt_sigma (hypernerf/models.py:1035-1063) over N rows.

    x: (N,3) observation points; warp_embed (N,8) or None; mask (N,1) or None.
    """
    c = self.cfg
    aux = {}
    # map_points (models.py:710-764)
    if c.use_warp:
      w_in = torch.cat([warp_embed, mask], -1) if c.use_mask_in_warp else warp_embed
      R, p, screw = self._warp_rp(x, w_in, extra_params)
      spatial = se3_apply(R, p, x)
      aux.update(R=R, p=p, screw_axis=screw)
    else:
      spatial = x
    if c.has_hyper_sheet:
      h_in = torch.cat([warp_embed, mask], -1) if c.use_mask_in_hyper else warp_embed
      pe = posenc(x, c.hyper_sheet_min_deg, c.hyper_sheet_max_deg,
                  alpha=extra_params['hyper_sheet_alpha'])       # modules.py:373
      hyper = mlp_apply(self.P['hyper_sheet_mlp']['MLP_0'],
                        torch.cat([pe, h_in], -1), c.hyper_sheet_depth,
                        c.hyper_sheet_skips, mm=self.mm, tag='hyper')
    else:
      hyper = None
    if hyper is not None and c.use_hyper_for_sigma:
      warped = torch.cat([spatial, hyper], dim=-1)
    else:
      warped = spatial
    # pre_process_query (models.py:493-523)
    feat = posenc(warped[..., :3], c.spatial_point_min_deg,
                  c.spatial_point_max_deg, c.use_posenc_identity,
                  alpha=extra_params['nerf_alpha'])
    if warped.shape[-1] > 3:
      hf = posenc(warped[..., 3:], c.hyper_point_min_deg, c.hyper_point_max_deg,
                  False, alpha=extra_params['hyper_alpha'])
      feat = torch.cat([feat, hf], dim=-1)
    # query_bottleneck / query_sigma (modules.py:243-286)
    pn = self.P[f'nerf_mlps_{level}']
    trunk_out = mlp_apply(pn['trunk_mlp'], feat, c.nerf_trunk_depth, c.nerf_skips,
                          mm=self.mm, tag='trunk')
    if c.use_viewdirs:            # rgb_condition is not None -> bottleneck layer
      bottleneck = self.mm(trunk_out, pn['bottleneck']['kernel'], 'bottleneck') + pn['bottleneck']['bias']
    else:
      bottleneck = trunk_out
    alpha_out = self.mm(trunk_out, pn['alpha_mlp']['logit']['kernel'], 'alpha') + pn['alpha_mlp']['logit']['bias']
    sigma_raw = alpha_out[..., 0]
    norm = alpha_out[..., 1:4] if c.predict_norm else None
    aux.update(norm=norm, warped_points=warped, trunk_out=trunk_out,
               bottleneck=bottleneck)
    return sigma_raw, aux

  # ---- render_samples ----------------------------------------------------
  def render_samples(self, level, points, z_vals, directions, viewdirs,
                     metadata, extra_params, gt_mask, *, use_warp=True,
                     use_sample_at_infinity=False, use_sigma_gradient=False,
                     use_predicted_norm=False, mask_ratio=1,
                     sharp_weights_std=1.0, compute_sigma_gradient=True,
                     render_opts=None) -> Dict[str, torch.Tensor]:
    """hypernerf/models.py:867-1417 under the flags NerfDSConfig admits."""
    c = self.cfg
    dt = self.dtype
    B, S = points.shape[:2]
    out = {'points': points}
    if c.use_warp and not use_warp:
      raise NotImplementedError('use_warp=False on a model built with a warp')
    warp_id = None
    if c.use_warp:
      warp_id = torch.as_tensor(np.asarray(metadata['warp'])).reshape(B).long()
      warp_embed = self.P['warp_embed']['embed']['embedding'][warp_id]     # models.py:901-902
      warp_embed = warp_embed[:, None, :].expand(B, S, -1).reshape(B * S, -1)
    else:
      warp_embed = None
    x = points.reshape(B * S, 3)
    if gt_mask is not None:
      gt_mask_b = self._t(gt_mask).reshape(B, 1)[:, None, :].expand(B, S, 1).reshape(B * S, 1)
    else:
      gt_mask_b = None
    if c.use_predicted_mask:                                     # models.py:955-975
      mask_embed = self.P['mask_embed']['embed']['embedding'][warp_id]     # 924-926 (warp key)
      mask_embed = mask_embed[:, None, :].expand(B, S, -1).reshape(B * S, -1)
      pe = posenc(x, c.mask_min_deg, c.mask_max_deg, alpha=extra_params['warp_alpha'])
      predicted_mask = mlp_apply(self.P['mask_mlp']['MLP_0'],
                                 torch.cat([pe, mask_embed], -1), c.mask_depth,
                                 c.mask_skips, out_relu=c.mask_output_relu,
                                 mm=self.mm, tag='mask')
      out['predicted_mask'] = predicted_mask.reshape(B, S, 1)
      mr = float(mask_ratio)
      if gt_mask_b is None:
        if mr != 1.0:
          raise ValueError('gt mask required when mask_ratio != 1')
        mask = predicted_mask * mr
      else:
        mask = predicted_mask * mr + gt_mask_b * (1 - mr)
    else:
      predicted_mask = None
      mask = gt_mask_b
    mask = None if mask is None else mask.detach()

    # value_and_grad of raw sigma w.r.t. the observation point (1035-1077)
    sig_list, grad_list, aux_list = [], [], []
    for s0 in range(0, B * S, self.row_chunk):
      sl = slice(s0, min(B * S, s0 + self.row_chunk))
      xs = x[sl].detach().clone()
      we = None if warp_embed is None else warp_embed[sl]
      mk = None if mask is None else mask[sl]
      if compute_sigma_gradient:
        xs.requires_grad_(True)
        s_raw, aux = self._sigma_field(level, xs, we, mk, extra_params)
        (g,) = torch.autograd.grad(s_raw.sum(), xs)
        grad_list.append(-g)
      else:
        with torch.no_grad():
          s_raw, aux = self._sigma_field(level, xs, we, mk, extra_params)
      sig_list.append(s_raw.detach())
      aux_list.append({k: (v.detach() if v is not None else None) for k, v in aux.items()})
    sigma_raw = torch.cat(sig_list)
    aux = {k: (torch.cat([a[k] for a in aux_list]) if aux_list[0][k] is not None else None)
           for k in aux_list[0]}
    if compute_sigma_gradient:
      sigma_gradient = normalize_vector(torch.cat(grad_list))  # models.py:1077
    else:
      sigma_gradient = None
    norm = aux['norm']
    warped_points = aux['warped_points']
    trunk_out = aux['trunk_out']
    bottleneck = aux['bottleneck']
    R = aux.get('R')
    p = aux.get('p')

    with torch.no_grad():
      # normal used as rgb input (models.py:1107-1152)
      if use_sigma_gradient:
        assert not use_predicted_norm
        norm_input = sigma_gradient
      elif use_predicted_norm:
        normalized_norm = normalize_vector(norm)
        if c.use_warp:
          norm_input = se3_apply(R, p, normalized_norm, rotation_only=True,
                                 inverse=True)                   # map_vectors 1126
        else:
          norm_input = normalized_norm                           # map_vectors 604-605
      else:
        norm_input = None
      if norm_input is not None:
        norm_input = normalize_vector(norm_input)
        if c.norm_input_posenc:
          norm_input_feat = posenc(norm_input, c.norm_input_min_deg,
                                   c.norm_input_max_deg, c.use_posenc_identity,
                                   alpha=extra_params['norm_input_alpha'])
        else:
          norm_input_feat = norm_input
      else:
        norm_input_feat = None

      # rgb condition (models.py:393-429): posenc(viewdirs), broadcast per ray
      if c.use_viewdirs:
        vfeat = posenc(viewdirs, c.viewdir_min_deg, c.viewdir_max_deg,
                       c.use_posenc_identity)
        vfeat = vfeat[:, None, :].expand(B, S, -1).reshape(B * S, -1)
        rgb_input = torch.cat([bottleneck, vfeat], -1)            # modules.py:297-300
      else:
        rgb_input = trunk_out
      if c.use_x_in_rgb_condition:
        # App. C-1: points_feat was rebound to the trunk output (models.py:1046,1208)
        rgb_input = torch.cat([rgb_input, trunk_out], -1)
      if norm_input_feat is not None:
        rgb_input = torch.cat([rgb_input, norm_input_feat], -1)  # modules.py:308-310
      pn = self.P[f'nerf_mlps_{level}']
      rgb_raw = mlp_apply(pn['rgb_mlp'], rgb_input, c.nerf_rgb_branch_depth, (),
                          mm=self.mm, tag='rgb')

      # sharp weights (models.py:1235-1246)
      sigma_raw_bs = sigma_raw.reshape(B, S)
      filtered_sigma = filter_sigma(points, sigma_raw_bs, render_opts)   # the RAW sigma (models.py:1236)
      sigmoid_sigma = torch.nn.functional.softplus(filtered_sigma)
      weights_sg = cal_weights(sigmoid_sigma, z_vals, directions)
      if c.use_mask_sharp_weights:
        out['sharp_weights'] = sharpen_weights(weights_sg, z_vals, std=sharp_weights_std)

      # post_process_query (models.py:567-579)
      rgb = torch.sigmoid(rgb_raw).reshape(B, S, 3)
      sigma = torch.nn.functional.softplus(sigma_raw_bs)
      out['sigma'] = sigma

      if c.predict_norm and compute_sigma_gradient:                # models.py:1273-1277
        if c.use_warp:
          sg_r = se3_apply(R, p, sigma_gradient, rotation_only=True)
        else:
          sg_r = sigma_gradient
        sigma_gradient_r = normalize_vector(sg_r)
      else:
        sigma_gradient_r = None

      if c.use_warp:                                             # models.py:1291-1302
        ref = normalize_vector(torch.ones_like(x))
        rotation_field = normalize_vector(se3_apply(R, p, ref, rotation_only=True))
        translation_field = se3_apply(R, p, torch.zeros_like(x))
      else:
        rotation_field = translation_field = None

      sigma = filter_sigma(points, sigma, render_opts)           # models.py:1288

      D = warped_points.shape[-1]
      warped_points = warped_points.reshape(B, S, D)
      out['warped_points'] = warped_points
      out.update(volumetric_rendering(
          rgb, sigma, z_vals, directions,
          use_white_background=c.use_white_background,
          sample_at_infinity=use_sample_at_infinity))

      if c.predict_norm:                                         # models.py:1324-1344
        norm_bs = norm.reshape(B, S, 3)
        out['predicted_norm'] = norm_bs
        if sigma_gradient_r is not None:
          out['target_norm'] = sigma_gradient_r.reshape(B, S, 3)
        back = torch.einsum('ijk,ijk->ij', norm_bs,
                            viewdirs[:, None, :].expand(B, S, 3))
        out['back_facing'] = torch.square(torch.relu(back))
      weights = out['weights']
      if norm is not None:                                       # models.py:1350-1354
        out['ray_norm'] = (weights[..., None] * norm.reshape(B, S, 3)).sum(-2)
      elif sigma_gradient is not None:
        out['ray_norm'] = (weights[..., None] * sigma_gradient.reshape(B, S, 3)).sum(-2)
      if rotation_field is not None:
        out['ray_rotation_field'] = (weights[..., None] * rotation_field.reshape(B, S, 3)).sum(-2)
        out['ray_translation_field'] = (weights[..., None] * translation_field.reshape(B, S, 3)).sum(-2)
      delta_x = warped_points[..., :3] - points                  # models.py:1363-1366
      out['delta_x'] = delta_x
      out['ray_delta_x'] = (weights[..., None] * delta_x).sum(-2)
      hyper_points = warped_points[..., 3:]
      out['ray_hyper_points'] = (weights[..., None] * hyper_points).sum(-2)
      out['ray_hyper_c'] = torch.zeros_like(out['ray_hyper_points'])  # 1384
      if c.use_predicted_mask:                                   # models.py:1399
        out['ray_predicted_mask'] = (weights[..., None] * out['predicted_mask']).sum(-2)
      depth_indices = compute_depth_index(weights)               # models.py:1411-1415
      out['med_points'] = torch.take_along_dim(
          warped_points, depth_indices[..., None, None].expand(B, 1, D), dim=-2)
      out['_depth_index'] = depth_indices                        # oracle-only diagnostic
    return out

  # ---- __call__ ----------------------------------------------------------
  def apply(self, rays_dict: Dict[str, Any], extra_params: Dict[str, Any],
            t_rand=None, u=None, *, return_points=False, return_weights=False,
            near=None, far=None, use_sample_at_infinity=None,
            use_sigma_gradient=False, use_predicted_norm=False, mask_ratio=1,
            sharp_weights_std=1.0, compute_sigma_gradient=True,
            keep_internal=False, render_opts=None):
    """hypernerf/models.py:1419-1565."""
    c = self.cfg
    origins = self._t(rays_dict['origins'])
    directions = self._t(rays_dict['directions'])
    metadata = rays_dict.get('metadata', {})
    mask = rays_dict.get('mask')
    viewdirs = self._t(rays_dict['viewdirs']) if 'viewdirs' in rays_dict else directions
    near = c.near if near is None else near
    far = c.far if far is None else far
    if use_sample_at_infinity is None:
      use_sample_at_infinity = c.use_sample_at_infinity
    B = origins.shape[0]
    if c.use_stratified_sampling:
      t_rand = self._t(t_rand)
      u = self._t(u)
    else:
      t_rand = None
      u = linspace01(c.num_fine_samples, self.dtype)[None, :].expand(B, -1)
    z_vals, points = sample_along_rays(
        t_rand, origins, directions, c.num_coarse_samples, near, far,
        c.use_stratified_sampling, c.use_linear_disparity)
    kw = dict(use_sigma_gradient=use_sigma_gradient,
              use_predicted_norm=use_predicted_norm, mask_ratio=mask_ratio,
              sharp_weights_std=sharp_weights_std,
              compute_sigma_gradient=compute_sigma_gradient)
    coarse = self.render_samples(
        'coarse', points, z_vals, directions, viewdirs, metadata, extra_params,
        mask, use_sample_at_infinity=c.use_sample_at_infinity, **kw)  # 1509
    out = {'coarse': coarse}
    z_mid = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
    z_fine, points_f = sample_pdf(u, z_mid, coarse['weights'][..., 1:-1],
                                  origins, directions, z_vals)
    out['fine'] = self.render_samples(
        'fine', points_f, z_fine, directions, viewdirs, metadata, extra_params,
        mask, use_sample_at_infinity=use_sample_at_infinity,
        render_opts=render_opts, **kw)                           # models.py:1545 (fine level only)
    if keep_internal:
      out['coarse']['z_vals'] = z_vals
      out['fine']['z_vals'] = z_fine
    for lvl in ('coarse', 'fine'):
      if not keep_internal:
        out[lvl].pop('_depth_index', None)
      if not return_weights:
        del out[lvl]['weights']
      if not return_points:
        del out[lvl]['points']
        del out[lvl]['warped_points']
    return out


def to_numpy(tree):
  if isinstance(tree, dict):
    return {k: to_numpy(v) for k, v in tree.items()}
  return tree.detach().cpu().numpy() if torch.is_tensor(tree) else tree
