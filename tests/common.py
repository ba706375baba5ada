"""Shared helpers of the parity tests: seeded workloads for BASELINE configs."""
import numpy as np
import torch

from nerfds_b200 import synthetic as syn
from nerfds_b200.config import nerf_ds_config, tiny_config
from nerfds_b200.params import init_params
from oracle.nerfds_oracle import OracleNerfModel, to_numpy

RGB_TOL = 1e-3   # north_star: <= 1e-3 RGB L-inf vs the fp32 reference path


def make_case(kind: str, image: int = 12, seed: int = 0, **overrides):
  """kind: 'tiny' (BASELINE configs[0]) | 'nerf_ds' (configs[1], nerf_ds.gin nets)."""
  if kind == 'tiny':
    cfg = tiny_config(**overrides)
  else:
    cfg = nerf_ds_config(**overrides)
  params = init_params(cfg, seed)
  rays = syn.frame_rays(image, image, frame=3 + seed, focal=float(image) * 1.1, warp_id=5 % cfg.num_warp_embeds)
  B = rays['origins'].shape[0]
  rng = np.random.default_rng(seed)
  if cfg.use_warp:   # per-ray ids like a train batch
    rays['metadata']['warp'] = rng.integers(0, cfg.num_warp_embeds, size=(B, 1)).astype(np.uint32)
  rays['mask'] = rng.integers(0, 2, size=(B, 1)).astype(np.float32)
  t_rand, u = syn.uniform_draws(B, cfg.num_coarse_samples, cfg.num_fine_samples, seed)
  return cfg, params, rays, t_rand, u


def run_oracle(cfg, params, rays, t_rand, u, dtype=torch.float32, **kw):
  m = OracleNerfModel(cfg, params, dtype=dtype)
  kw.setdefault('use_predicted_norm', cfg.predict_norm)
  return to_numpy(m.apply(rays, syn.final_extra_params(), t_rand, u, return_points=True, return_weights=True,
                          keep_internal=True, **kw))


def linf(a, b):
  return float(np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)))) if np.size(a) else 0.0
