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


def bench_scene(n_rays: int, coarse: int = 128, fine: int = 128, seed: int = 0):
  """n_rays of BASELINE configs[1]'s scene (nerf_ds.gin widths, near 0.1 / far 2.5, 800x800 orbit cameras): rays spread
  over four frames, per-ray warp ids, stratified draws.  Shared by the at-scale parity test and tools/parity_scale.py."""
  cfg = nerf_ds_config(num_coarse_samples=coarse, num_fine_samples=fine, near=0.1, far=2.5, num_warp_embeds=100)
  params = init_params(cfg, 0)
  rng = np.random.default_rng(seed)
  per = (n_rays + 3) // 4
  o, d = [], []
  for f in range(4):
    r = syn.frame_rays(800, 800, frame=7 * f, num_frames=30, focal=800.)
    sel = np.sort(rng.choice(640000, size=per, replace=False))
    o.append(r['origins'][sel])
    d.append(r['directions'][sel])
  rays = {'origins': np.concatenate(o)[:n_rays], 'directions': np.concatenate(d)[:n_rays],
          'metadata': {'warp': rng.integers(0, cfg.num_warp_embeds, size=(n_rays, 1)).astype(np.uint32)},
          'mask': np.zeros((n_rays, 1), np.float32)}
  t_rand, u = syn.uniform_draws(n_rays, coarse, fine, seed)
  return cfg, params, rays, t_rand, u


def take_rays(rays, sl):
  return {'origins': rays['origins'][sl], 'directions': rays['directions'][sl],
          'metadata': {k: v[sl] for k, v in rays['metadata'].items()}, 'mask': rays['mask'][sl]}


def run_oracle_chunks(cfg, params, rays, t_rand, u, chunk=1024, **kw):
  """The fp32 oracle over row chunks (bounded memory): {'coarse': {...}, 'fine': {...}} of numpy arrays."""
  m = OracleNerfModel(cfg, params)
  kw = dict(dict(use_predicted_norm=True, mask_ratio=1, sharp_weights_std=0.1, return_weights=True, return_points=True,
                 keep_internal=True, compute_sigma_gradient=False), **kw)
  outs = {'coarse': {}, 'fine': {}}
  n = rays['origins'].shape[0]
  for r0 in range(0, n, chunk):
    sl = slice(r0, min(n, r0 + chunk))
    o = to_numpy(m.apply(take_rays(rays, sl), syn.final_extra_params(), t_rand[sl], u[sl], **kw))
    for lvl in outs:
      for k, v in o[lvl].items():
        outs[lvl].setdefault(k, []).append(v)
  return {lvl: {k: np.concatenate(v) for k, v in d.items()} for lvl, d in outs.items()}
