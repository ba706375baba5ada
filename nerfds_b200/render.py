"""`render_scene` -- the reference's render.py driver on top of the CUDA path.

render.py:36-277 restores an experiment directory (operative `config.gin` +
newest Flax checkpoint), loads the test-camera JSONs of a dataset, and for
every `interval`-th camera turns the camera into rays (datasets/core.py:51-76),
tags them with the frame's warp id, renders them in chunks
(evaluation.render_image) and keeps the per-frame maps listed in
`RELEVANT_KEYS`.  This module does the same with

  experiment dir  -> `checkpoints.load_experiment`      (section 8 row f-2)
  camera -> rays  -> `camera.camera_to_rays` on device  (row f-3)
  rays -> maps    -> `evaluation.render_image_sharded`  (rows a-e; rays block-
                     sharded over the process group when one is initialised)

It does not display or encode video (render.py:245-277 cv2/mediapy code).
"""
from __future__ import annotations

import glob
import json
import os
from typing import Dict, Iterable, List, Optional

import numpy as np
import torch

from . import checkpoints
from .camera import Camera, camera_to_rays, load_camera
from . import jax_random
from .evaluation import reference_draws, render_image_sharded
from .models import NerfModel

# render.py:189-190
RELEVANT_KEYS = ('rgb', 'med_depth', 'ray_norm', 'ray_delta_x', 'med_points', 'ray_predicted_mask',
                 'ray_rotation_field')


def sort_camera_paths(paths: Iterable[str]) -> List[str]:
  """render.py:280-295: order by the integer-looking id token of the file stem -- compared as a
  *string*, exactly as the reference does ('10' sorts before '2'), then by path."""
  pairs = []
  for p in paths:
    stem = os.path.splitext(os.path.basename(p))[0]
    tok = stem.split('_')
    cam_id = tok[1] if len(tok) > 1 and tok[1].lstrip('-').isdigit() else tok[0]
    int(cam_id)
    pairs.append((cam_id, p))
  return [p for _, p in sorted(pairs)]


def load_mask(path: str) -> np.ndarray:
  """datasets/core.py:114-128: grayscale png / 255, inverted so the moving part is 1; (H, W, 1)."""
  import cv2
  m = cv2.imdecode(np.fromfile(path, dtype=np.uint8), cv2.IMREAD_GRAYSCALE)
  if m is None:
    raise ValueError(f'cannot decode {path}')
  return (1.0 - m.astype(np.float32) / 255.0)[:, :, None]


def render_scene(exp_dir: str, data_dir: str, camera_path_name: str = 'vrig_camera', interval: int = 1,
                 chunk_size: int = 65536, *, device=None, keys: Iterable[str] = RELEVANT_KEYS, seed: Optional[int] = None,
                 save: bool = True, precision: str = 'split3', step: Optional[int] = None,
                 device_count: Optional[int] = None) -> List[Dict[str, np.ndarray]]:
  """render.py:36-243.  Returns (and, like the reference, np.save's under
  `<exp_dir>/render_result_<camera_path_name>`) one dict of (H, W, .) maps per rendered camera.

  The stratified draws are the ones the reference makes: `rng = PRNGKey(random_seed); rng, _ = split(rng)`
  (render.py:85, 100), the same `rng` for every frame (render.py:218), split per device and drawn per chunk as
  evaluation.py:81-120 does for `device_count` devices (default: the size of the process group) -- generated on the
  device by the threefry kernel."""
  cfg, params, extra, state, bindings = checkpoints.load_experiment(exp_dir, data_dir=data_dir, step=step)
  image_scale = bindings.get('ExperimentConfig.image_scale', bindings.get('image_scale', 1))
  with open(os.path.join(data_dir, 'scene.json')) as f:
    scene = json.load(f)
  use_predicted_norm = bool(bindings.get('SpecularConfig.use_predicted_norm', False))
  seed = int(bindings.get('ExperimentConfig.random_seed', 0)) if seed is None else seed

  cam_paths = sort_camera_paths(glob.glob(os.path.join(data_dir, camera_path_name, '*.json')))
  if not cam_paths:
    raise FileNotFoundError(f'no camera JSON under {os.path.join(data_dir, camera_path_name)}')
  cameras = [load_camera(p, scale_factor=1.0 / image_scale, scene_center=scene['center'], scene_scale=scene['scale'])
             for p in cam_paths]
  mask_dir = os.path.join(data_dir, 'resized_mask', f'{int(image_scale)}x')

  model = NerfModel(cfg, device=device, precision=precision)
  dev = model.device
  rng, _ = jax_random.split(jax_random.PRNGKey(seed))          # render.py:85, 100
  draws = {}
  results = []
  # one process per GPU: frames are assembled in every GPU's frame buffer by peer stores of the compositing kernel;
  # two buffer sets per image size alternate so that frame f can be copied out while frame f + 1 is written
  import torch.distributed as dist
  multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
  rank = dist.get_rank() if multi else 0
  peer_sets, n_rendered = {}, 0
  for i in range(0, len(cameras), interval):
    if cfg.use_warp and i >= cfg.num_warp_embeds:
      raise IndexError(f'camera {i} has no warp embedding (the checkpoint holds {cfg.num_warp_embeds})')
    camera: Camera = cameras[i]
    batch = camera_to_rays(camera, dev)
    H, W = camera.image_shape
    stem = os.path.splitext(os.path.basename(cam_paths[i]))[0]
    mask_path = os.path.join(mask_dir, f'{stem}.png.png')
    # inference blends with mask_ratio = 1 (render.py:152): the file only matters to callers that ask for it
    mask = torch.from_numpy(load_mask(mask_path)).to(dev) if os.path.exists(mask_path) else torch.zeros((H, W, 1), device=dev)
    rays = {'origins': batch['origins'], 'directions': batch['directions'],
            'metadata': {'warp': torch.full((H, W, 1), i, dtype=torch.int64, device=dev)},   # render.py:203-216
            'mask': mask}
    t_rand = u = None
    if cfg.use_stratified_sampling:
      if (H, W) not in draws:
        D = device_count or (dist.get_world_size() if multi else 1)
        draws[(H, W)] = reference_draws(rng, H * W, chunk_size, D, cfg.num_coarse_samples, cfg.num_fine_samples, dev)
      t_rand, u = draws[(H, W)]
    pf = None
    if multi:
      from .peer import PeerFrames
      sets = peer_sets.setdefault((H, W), [])
      if len(sets) < 2:
        sets.append(PeerFrames(model.renderer, H * W, tuple(keys)))
      pf = sets[n_rendered % len(sets)] if len(sets) == 2 else sets[-1]
    out = render_image_sharded(model, params, rays, extra, t_rand=t_rand, u=u, chunk=chunk_size, keys=tuple(keys),
                               use_predicted_norm=use_predicted_norm, peer_frames=pf)
    results.append({k: v.cpu().numpy() for k, v in out.items()})
    n_rendered += 1
  for sets in peer_sets.values():
    for pf in sets:
      pf.close()
  if save and rank == 0:
    name = camera_path_name + ('_full' if interval == 1 else '')                             # render.py:193-194
    with open(os.path.join(exp_dir, f'render_result_{name}'), 'wb+') as f:
      np.save(f, np.array(results, dtype=object), allow_pickle=True)
  return results
