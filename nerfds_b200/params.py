"""Parameter pytree of the path, in the reference's Flax layout.

Names and shapes follow what ``NerfModel.setup`` creates
(hypernerf/models.py:324-391): nested dicts, every
``nn.Dense`` is ``{'kernel': [in, out], 'bias': [out]}``, every ``GLOEmbed``
is ``{'embed': {'embedding': [num_ids, dims]}}``.

``init_params`` is a *synthetic* generator (there are no checkpoints here):
it mimics the Flax initialisers the reference uses (glorot-uniform hidden
layers modules.py:48, ``uniform(0.05)`` embeddings modules.py:328) but scales
the SE(3) / hyper / mask logits and the sigma row up, so that rotations,
hyper-coordinates, masks and opacities are exercised rather than ~0
(SURVEY.md section 8(d)).
"""
from __future__ import annotations

from typing import Dict, Iterator, Tuple

import numpy as np

from .config import NerfDSConfig


def _glorot(rng, fan_in, fan_out):
  lim = np.sqrt(6.0 / (fan_in + fan_out))
  return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(np.float32)


def _dense(rng, fan_in, fan_out, kernel_scale=None, bias_scale=0.05):
  if kernel_scale is None:
    k = _glorot(rng, fan_in, fan_out)
  else:
    k = rng.uniform(-kernel_scale, kernel_scale,
                    size=(fan_in, fan_out)).astype(np.float32)
  b = rng.uniform(-bias_scale, bias_scale, size=(fan_out,)).astype(np.float32)
  return {'kernel': k, 'bias': b}


def _mlp(rng, in_dim, depth, width, skips, out_channels=0, out_scale=None,
         gain=1.0):
  """modules.MLP parameter block (modules.py:57-83)."""
  p = {}
  d = in_dim
  for i in range(depth):
    if i in skips:
      d = d + in_dim
    p[f'hidden_{i}'] = _dense(rng, d, width)
    if gain != 1.0:
      p[f'hidden_{i}']['kernel'] *= np.float32(gain)
    d = width
  if out_channels > 0:
    p['logit'] = _dense(rng, d, out_channels, kernel_scale=out_scale)
  return p


def _np_posenc(x, min_deg, max_deg, identity):
  scales = (2.0 ** np.arange(min_deg, max_deg)).astype(np.float32)
  xb = x[..., None, :] * scales[:, None]
  f = np.sin(np.stack([xb, xb + np.float32(0.5 * np.pi)], -2))
  f = f.reshape(*x.shape[:-1], -1)
  return np.concatenate([x, f], -1) if identity else f


def _calibrate_sigma(blk, cfg, rng, median=-30.0, std=30.0):
  """Rescale the sigma column (column 0 of alpha_mlp/logit) so that raw
  sigma over the scene box has the given median / spread: mostly empty space
  with opaque surfaces, like a trained field, instead of uniform fog."""
  n = 4096
  x = rng.uniform(-1.5, 1.5, size=(n, 3)).astype(np.float32)
  feat = _np_posenc(x, cfg.spatial_point_min_deg, cfg.spatial_point_max_deg,
                    cfg.use_posenc_identity)
  if cfg.has_hyper_sheet and cfg.use_hyper_for_sigma:
    w = rng.uniform(-0.6, 0.6, size=(n, cfg.hyper_num_dims)).astype(np.float32)
    feat = np.concatenate([feat, _np_posenc(
        w, cfg.hyper_point_min_deg, cfg.hyper_point_max_deg, False)], -1)
  h, inputs = feat, feat
  for i in range(cfg.nerf_trunk_depth):
    if i in cfg.nerf_skips:
      h = np.concatenate([h, inputs], -1)
    lyr = blk['trunk_mlp'][f'hidden_{i}']
    h = np.maximum(h @ lyr['kernel'] + lyr['bias'], 0)
  col = blk['alpha_mlp']['logit']['kernel'][:, 0]
  pre = h @ col
  scale = np.float32(std / max(float(pre.std()), 1e-6))
  blk['alpha_mlp']['logit']['kernel'][:, 0] = col * scale
  blk['alpha_mlp']['logit']['bias'][0] = np.float32(
      median - float(np.median(pre)) * float(scale))


def init_params(cfg: NerfDSConfig, seed: int = 0) -> Dict:
  cfg.validate()
  rng = np.random.default_rng(seed)
  P: Dict = {}
  if cfg.use_warp:
    P['warp_embed'] = {'embed': {'embedding': rng.uniform(
        0, 0.05, size=(cfg.num_warp_embeds, cfg.warp_embed_dims)
    ).astype(np.float32) * 8}}
  if cfg.use_predicted_mask:
    P['mask_embed'] = {'embed': {'embedding': rng.uniform(
        0, 0.05, size=(cfg.num_warp_embeds, cfg.mask_embed_dims)
    ).astype(np.float32) * 8}}
  if cfg.use_warp:
    # hidden layers get a relu gain so the 6-deep trunk keeps O(1) activations
    P['warp_field'] = {
        'trunk': _mlp(rng, cfg.warp_in_dim, cfg.warp_trunk_depth,
                      cfg.warp_trunk_width, cfg.warp_skips, gain=1.4),
        'branches_w': {'logit': _dense(rng, cfg.warp_trunk_width, 3,
                                       kernel_scale=0.03, bias_scale=0.02)},
        'branches_v': {'logit': _dense(rng, cfg.warp_trunk_width, 3,
                                       kernel_scale=0.03, bias_scale=0.02)},
    }
  if cfg.has_hyper_sheet:
    P['hyper_sheet_mlp'] = {'MLP_0': _mlp(
        rng, cfg.hyper_sheet_in_dim, cfg.hyper_sheet_depth,
        cfg.hyper_sheet_width, cfg.hyper_sheet_skips,
        out_channels=cfg.hyper_num_dims, out_scale=0.15, gain=1.4)}
  if cfg.use_predicted_mask:
    P['mask_mlp'] = {'MLP_0': _mlp(
        rng, cfg.mask_in_dim, cfg.mask_depth, cfg.mask_width, cfg.mask_skips,
        out_channels=1, out_scale=0.15, gain=1.4)}
    P['mask_mlp']['MLP_0']['logit']['bias'] += np.float32(0.2)
  for level in ('coarse', 'fine'):
    W = cfg.nerf_trunk_width
    blk = {
        'trunk_mlp': _mlp(rng, cfg.trunk_in_dim, cfg.nerf_trunk_depth, W,
                          cfg.nerf_skips, gain=1.4),
        'bottleneck': _dense(rng, W, W),
        'alpha_mlp': {'logit': _dense(rng, W, cfg.alpha_out_channels)},
        'rgb_mlp': _mlp(rng, cfg.rgb_in_dim(cfg.predict_norm),
                        cfg.nerf_rgb_branch_depth, cfg.nerf_rgb_branch_width,
                        (), out_channels=cfg.rgb_channels, gain=1.4),
    }
    _calibrate_sigma(blk, cfg, rng)
    blk['rgb_mlp']['logit']['kernel'] *= np.float32(3.0)
    P[f'nerf_mlps_{level}'] = blk
  return P


def harden_density(params: Dict, scale: float = 40.0, shift: float = 0.0) -> Dict:
  """A copy of `params` whose raw densities (column 0 of both levels' alpha_mlp/logit) are multiplied by `scale` and
  raised by `shift`: turns the soft random-init field (sigma_raw ~ N(-30, 30): a ray needs a third of the scene to go
  opaque) into one with hard surfaces, where the transmittance collapses inside the cluster of fine samples -- the
  regime of a trained scene, and the workload an early-termination scan is meant for."""
  import copy
  out = copy.deepcopy(params)
  for level in ('coarse', 'fine'):
    lg = out[f'nerf_mlps_{level}']['alpha_mlp']['logit']
    lg['kernel'][:, 0] *= np.float32(scale)
    lg['bias'][0] = lg['bias'][0] * np.float32(scale) + np.float32(shift)
  return out


def flatten_params(params: Dict, prefix: str = '') -> Iterator[Tuple[str, np.ndarray]]:
  """Yield ``('warp_field/trunk/hidden_0/kernel', array)`` pairs, sorted."""
  for k in sorted(params):
    v = params[k]
    name = f'{prefix}/{k}' if prefix else k
    if isinstance(v, dict):
      yield from flatten_params(v, name)
    else:
      yield name, np.asarray(v)


def unflatten_params(flat: Dict[str, np.ndarray]) -> Dict:
  out: Dict = {}
  for name, v in flat.items():
    node = out
    parts = name.split('/')
    for p in parts[:-1]:
      node = node.setdefault(p, {})
    node[parts[-1]] = v
  return out


def param_count(params: Dict) -> int:
  return sum(int(v.size) for _, v in flatten_params(params))
