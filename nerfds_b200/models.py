"""``NerfModel`` -- the reference's model-call surface, backed by the CUDA path.

Mirrors hypernerf/models.py: ``construct_nerf`` (2677-2741) and
``NerfModel.apply`` == ``model.apply({'params': P}, rays_dict, extra_params=...,
rngs=..., **flags)`` as invoked at render.py:140-154 and training.py:441-455,
returning ``{'coarse': {...}, 'fine': {...}}`` with the keys / shapes of
SURVEY.md App. B (torch CUDA tensors instead of jax device arrays).
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, Optional

import numpy as np
import torch

from . import jax_random
from .config import NerfDSConfig, nerf_ds_config
from .params import init_params
from .renderer import Renderer, RENDER_KEYS  # noqa: F401


def _key_to_seed(key) -> int:
  """Fold a jax-style PRNG key (uint32[2]) or int into a torch seed."""
  if key is None:
    return 0
  a = np.asarray(key.cpu() if torch.is_tensor(key) else key).astype(np.uint64).reshape(-1)
  seed = 0
  for v in a:
    seed = (seed * 6364136223846793005 + int(v) + 1442695040888963407) % (1 << 63)
  return seed


class NerfModel:
  """B200-native stand-in for ``models.NerfModel`` (models.py:71-1565)."""

  def __init__(self, cfg: NerfDSConfig, device=None, engine: str = 'auto', precision: str = 'split3'):
    cfg.validate()
    self.cfg = cfg
    self.renderer = Renderer(cfg, device=device, engine=engine, precision=precision)
    self.device = self.renderer.device

  # reference attribute names used by callers (render.py / evaluation.py)
  @property
  def use_warp(self):
    return self.cfg.use_warp

  @property
  def num_coarse_samples(self):
    return self.cfg.num_coarse_samples

  @property
  def num_fine_samples(self):
    return self.cfg.num_fine_samples

  def _draws(self, B, rngs, t_rand, u):
    """The uniform draws of model_utils.py:84 / 217.

    When the caller does not pass `t_rand` / `u` they are generated on the
    device from `rngs['coarse']` / `rngs['fine']` exactly as flax + jax do
    (jax_random.py: the first `make_rng` key of each stream, then
    `random.uniform(key, [B, S])` under threefry2x32), so a JAX run with the
    same keys sees the same samples.  Keys are raw uint32[2]
    jax keys (or python int seeds, as `random.PRNGKey(seed)`).
    """
    c = self.cfg
    if not c.use_stratified_sampling:
      return None, None
    if t_rand is None:
      if not rngs or rngs.get('coarse') is None:
        raise ValueError("NerfModel needs PRNG for \"coarse\"")          # flax errors.InvalidRngError
      t_rand = jax_random.uniform(jax_random.flax_make_rng(rngs['coarse']), (B, c.num_coarse_samples), self.device)
    if u is None:
      if not rngs or rngs.get('fine') is None:
        raise ValueError("NerfModel needs PRNG for \"fine\"")
      u = jax_random.uniform(jax_random.flax_make_rng(rngs['fine']), (B, c.num_fine_samples), self.device)
    return t_rand, u

  def apply(self, variables: Dict[str, Any], rays_dict: Dict[str, Any], extra_params: Dict[str, Any], *,
            rngs: Optional[Dict[str, Any]] = None, mutable: bool = False, metadata_encoded: bool = False,
            use_warp: bool = True, return_points: bool = False, return_weights: bool = False,
            return_warp_jacobian: bool = False, return_hyper_jacobian: bool = False,
            return_hyper_c_jacobian: bool = False, return_nv_details: bool = True, near=None, far=None,
            use_sample_at_infinity=None, render_opts=None, deterministic: bool = False, screw_input_mode=None,
            use_sigma_gradient: bool = False, use_predicted_norm: bool = False, mask_ratio=1,
            sharp_weights_std=1.0, x_for_rgb_alpha=4.0, norm_override=None,
            t_rand=None, u=None, keys: Optional[Iterable[str]] = None,
            coarse_keys: Optional[Iterable[str]] = None,
            fine_ptrs: Optional[Dict[str, int]] = None) -> Dict[str, Dict[str, torch.Tensor]]:
    """models.py:1419-1565.  Extra (non-reference) keyword arguments:

    t_rand / u   explicit uniform draws of model_utils.py:84,217
    keys         subset of level-dict keys to produce for 'fine' (output mask);
                 default = every key the reference returns under this config
    coarse_keys  same for 'coarse' (default = ``keys``)
    fine_ptrs    {key: device address} for the fine level's results (peer.PeerFrames)
    """
    c = self.cfg
    if metadata_encoded:
      raise NotImplementedError('metadata_encoded=True')
    if return_warp_jacobian or return_hyper_jacobian or return_hyper_c_jacobian:
      raise NotImplementedError('jacobian outputs (elastic loss) are outside the built path')
    if screw_input_mode not in (None, 'none', 'None'):
      raise NotImplementedError       # models.py:554-561
    if norm_override is not None:
      raise NotImplementedError('norm_override')
    if c.use_warp and not use_warp:
      raise NotImplementedError('use_warp=False on a model built with a warp field')
    self.renderer.ensure_params(variables['params'])
    origins = rays_dict['origins']
    B = int(np.prod(origins.shape[:-1]))
    metadata = rays_dict.get('metadata') or {}
    warp_id = metadata.get('warp') if c.use_warp else None
    if c.use_warp and warp_id is None:
      raise KeyError('warp')
    mask = rays_dict.get('mask')
    if mask is None and 'mask' not in rays_dict and (c.use_predicted_mask or c.use_mask_in_warp):
      raise KeyError('mask')            # models.py:1473
    t_rand, u = self._draws(B, rngs, t_rand, u)
    extra = self.renderer.make_extra(extra_params, mask_ratio=mask_ratio, sharp_weights_std=sharp_weights_std,
                                     use_predicted_norm=use_predicted_norm, use_sigma_gradient=use_sigma_gradient,
                                     near=near, far=far, use_sample_at_infinity=use_sample_at_infinity,
                                     render_opts=render_opts)
    if keys is None:
      fine_keys = self.renderer.level_keys(return_points=return_points, return_weights=return_weights)
    else:
      fine_keys = list(keys)
    ck = fine_keys if coarse_keys is None else list(coarse_keys)
    out = self.renderer.render_rays(origins, rays_dict['directions'], viewdirs=rays_dict.get('viewdirs'),
                                    warp_id=warp_id, gt_mask=mask, t_rand=t_rand, u=u, extra=extra,
                                    coarse_keys=ck, fine_keys=fine_keys, fine_ptrs=fine_ptrs)
    for lvl in out.values():
      if 'ray_hyper_points' in lvl:     # models.py:1384
        lvl['ray_hyper_c'] = torch.zeros_like(lvl['ray_hyper_points'])
    return out

  __call__ = apply


def construct_nerf(key, batch_size: int, embeddings_dict: Dict[str, Iterable[int]], near: float, far: float,
                   cfg: Optional[NerfDSConfig] = None, device=None, engine: str = 'auto',
                   precision: str = 'split3', **overrides):
  """models.construct_nerf (models.py:2677-2741): returns (model, params).

  ``params`` come from the synthetic Flax-like initialiser (params.py); a
  trained checkpoint's ``state.optimizer.target['model']`` has the same layout.
  """
  cfg = cfg or nerf_ds_config()
  n_warp = max(embeddings_dict.get('warp', [0])) + 1       # models.py:236
  cfg = cfg.replace(near=float(near), far=float(far), num_warp_embeds=int(n_warp), **overrides)
  model = NerfModel(cfg, device=device, engine=engine, precision=precision)
  params = init_params(cfg, seed=_key_to_seed(key) % (1 << 31))
  return model, params
