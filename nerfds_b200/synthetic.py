"""Synthetic workloads of BASELINE.json's configs (there are no datasets here).

Ray generation follows what the reference feeds the path:
``datasets.camera_to_rays`` (hypernerf/datasets/core.py:51-76) ->
``camera.pixels_to_rays`` (hypernerf/camera.py:245-270) for an undistorted
pinhole camera: pixel centres at +0.5, unit-norm directions, one origin per
camera; metadata ids as render.py:202-214 builds them (one warp id per frame).
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np


def look_at_camera(position, target=(0., 0., 0.), up=(0., 1., 0.)):
  """World-to-camera rotation (rows = camera axes), OpenCV convention (+z fwd)."""
  position = np.asarray(position, np.float64)
  fwd = np.asarray(target, np.float64) - position
  fwd /= np.linalg.norm(fwd)
  right = np.cross(fwd, np.asarray(up, np.float64))
  right /= np.linalg.norm(right)
  down = np.cross(fwd, right)
  return np.stack([right, down, fwd], 0)


def camera_rays(height: int, width: int, focal: float, position,
                orientation=None) -> Dict[str, np.ndarray]:
  """(H,W,3) origins / directions of an ideal pinhole camera."""
  if orientation is None:
    orientation = look_at_camera(position)
  xx, yy = np.meshgrid(np.arange(width, dtype=np.float64) + 0.5,
                       np.arange(height, dtype=np.float64) + 0.5)
  x = (xx - width / 2.0) / focal
  y = (yy - height / 2.0) / focal
  local = np.stack([x, y, np.ones_like(x)], -1)
  dirs = local @ orientation            # camera -> world (orientation is w2c)
  dirs /= np.linalg.norm(dirs, axis=-1, keepdims=True)
  origins = np.broadcast_to(np.asarray(position, np.float64), dirs.shape)
  return {'origins': origins.astype(np.float32).copy(),
          'directions': dirs.astype(np.float32)}


def orbit_position(frame: int, num_frames: int = 30, radius: float = 1.0,
                   height: float = 0.25):
  a = 2.0 * np.pi * frame / max(num_frames, 1)
  return np.array([radius * np.sin(a), height, -radius * np.cos(a)])


def frame_rays(height: int, width: int, *, frame: int = 0, num_frames: int = 30,
               focal: Optional[float] = None, warp_id: Optional[int] = None,
               flat: bool = True) -> Dict:
  """One render.py-style frame batch: rays + metadata + mask."""
  focal = float(width) if focal is None else focal
  rays = camera_rays(height, width, focal, orbit_position(frame, num_frames))
  wid = frame if warp_id is None else warp_id
  meta = np.full((height, width, 1), wid, np.uint32)
  rays['metadata'] = {'warp': meta, 'appearance': meta.copy(),
                      'camera': np.zeros_like(meta)}
  rays['mask'] = np.zeros((height, width, 1), np.float32)
  if flat:
    n = height * width
    rays = {'origins': rays['origins'].reshape(n, 3),
            'directions': rays['directions'].reshape(n, 3),
            'metadata': {k: v.reshape(n, 1) for k, v in rays['metadata'].items()},
            'mask': rays['mask'].reshape(n, 1)}
  return rays


def train_batch(num_rays: int, num_warp_ids: int, seed: int = 0,
                num_cameras: int = 8, image: int = 800) -> Dict:
  """BASELINE configs[2]: flat ray batch drawn from random cameras with
  per-ray warp ids and a {0,1} ground-truth mask (datasets/core.py:651-707)."""
  rng = np.random.default_rng(seed)
  cams = [camera_rays(image, image, float(image),
                      orbit_position(int(rng.integers(0, 30)), 30,
                                     radius=float(rng.uniform(0.9, 1.2))))
          for _ in range(num_cameras)]
  cam = rng.integers(0, num_cameras, size=num_rays)
  py = rng.integers(0, image, size=num_rays)
  px = rng.integers(0, image, size=num_rays)
  origins = np.stack([cams[c]['origins'][y, x] for c, y, x in zip(cam, py, px)])
  dirs = np.stack([cams[c]['directions'][y, x] for c, y, x in zip(cam, py, px)])
  warp = rng.integers(0, num_warp_ids, size=(num_rays, 1)).astype(np.uint32)
  return {'origins': origins.astype(np.float32),
          'directions': dirs.astype(np.float32),
          'metadata': {'warp': warp, 'appearance': warp.copy(),
                       'camera': np.zeros_like(warp)},
          'mask': rng.integers(0, 2, size=(num_rays, 1)).astype(np.float32)}


def uniform_draws(num_rays: int, num_coarse: int, num_fine: int, seed: int = 0):
  """The explicit stand-ins for random.uniform in sample_along_rays /
  piecewise_constant_pdf (model_utils.py:84,217): fp32 in [0, 1)."""
  rng = np.random.default_rng(seed + 7919)
  t_rand = rng.random((num_rays, num_coarse), dtype=np.float32)
  u = rng.random((num_rays, num_fine), dtype=np.float32)
  return t_rand, u


def final_extra_params() -> Dict[str, float]:
  """Schedule end values under nerf_ds.gin."""
  return {'nerf_alpha': 8.0, 'warp_alpha': 4.0, 'hyper_alpha': 1.0,
          'hyper_sheet_alpha': 6.0, 'norm_loss_weight': 1.0,
          'norm_input_alpha': 4.0, 'norm_voxel_lr': 0.0,
          'norm_voxel_ratio': 0.0}
