"""Experiment directory -> frames (SURVEY.md section 8 rows f-2 + f-3 feeding rows a-e), on the GPU."""
import json
import os

import numpy as np
import pytest
import torch

from nerfds_b200 import checkpoints as ckpt
from nerfds_b200.camera import camera_to_rays, load_camera
from nerfds_b200.config import nerf_ds_config
from nerfds_b200.model_utils import TrainState
from nerfds_b200.models import NerfModel
from nerfds_b200.params import init_params
from nerfds_b200.render import RELEVANT_KEYS, render_scene, sort_camera_paths
from tests.test_experiment_io import GIN_BASE, GIN_MAIN

pytestmark = pytest.mark.gpu


def _make_dataset(root, n_cams, W, H):
  os.makedirs(os.path.join(root, 'vrig_camera'))
  with open(os.path.join(root, 'scene.json'), 'w') as f:
    json.dump({'scale': 0.5, 'center': [0.1, -0.2, 0.3], 'near': 0.1, 'far': 2.5}, f)
  for i in range(n_cams):
    a = 0.2 * i
    R = [[np.cos(a), 0.0, -np.sin(a)], [0.0, 1.0, 0.0], [np.sin(a), 0.0, np.cos(a)]]
    cam = {'orientation': R, 'position': [0.1 + 2.4 * np.sin(a), -0.2, 0.3 - 2.4 * np.cos(a)],
           'focal_length': 2.0 * W, 'principal_point': [W, H], 'image_size': [2 * W, 2 * H], 'skew': 0.0,
           'pixel_aspect_ratio': 1.0, 'radial_distortion': [0.01, 0.0, 0.0], 'tangential': [0.0, 0.001]}
    with open(os.path.join(root, 'vrig_camera', f'{i:06d}.json'), 'w') as f:
      json.dump(cam, f)


def test_render_scene_from_experiment_dir(tmp_path, cuda_device):
  W, H, n_cams = 24, 16, 3
  data, exp = str(tmp_path / 'data'), str(tmp_path / 'exp')
  _make_dataset(data, n_cams, W, H)
  cfg = nerf_ds_config(num_coarse_samples=64, num_fine_samples=32, near=0.1, far=2.5, num_warp_embeds=5)
  params = init_params(cfg, 11)
  extra = {'nerf_alpha': 8.0, 'warp_alpha': 3.0, 'hyper_alpha': 1.0, 'hyper_sheet_alpha': 6.0,
           'norm_loss_weight': 0.001, 'norm_input_alpha': 4.0}
  state = TrainState.create(params, extra)
  state.optimizer.state.step = np.int32(42)
  os.makedirs(exp)
  gin = GIN_BASE + GIN_MAIN.replace("include 'base.gin'", '') + (
      "\nExperimentConfig.image_scale = 2\nExperimentConfig.random_seed = 3\nSpecularConfig.use_predicted_norm = True\n")
  with open(os.path.join(exp, 'config.gin'), 'w') as f:
    f.write(gin)
  ckpt.save_checkpoint(os.path.join(exp, 'checkpoints'), state, 42)

  frames = render_scene(exp, data, interval=2, chunk_size=256, device=cuda_device)
  assert len(frames) == 2 and set(frames[0]) == set(RELEVANT_KEYS)
  assert frames[0]['rgb'].shape == (H, W, 3) and np.isfinite(frames[1]['rgb']).all()
  saved = np.load(os.path.join(exp, 'render_result_vrig_camera'), allow_pickle=True)
  assert len(saved) == 2 and np.array_equal(saved[1]['rgb'], frames[1]['rgb'])

  # the same frame through the model-call surface with the restored pieces: identical bits
  cam = load_camera(os.path.join(data, 'vrig_camera', '000002.json'), scale_factor=0.5, scene_center=[0.1, -0.2, 0.3],
                    scene_scale=0.5)
  assert cam.image_shape == (H, W) and cam.focal_length == W
  rays = camera_to_rays(cam, cuda_device)
  # render.py:85, 100, 218: rng = split(PRNGKey(random_seed))[0], the same for every frame; evaluation.py:81-120
  from nerfds_b200 import jax_random as jr
  from nerfds_b200.evaluation import reference_draws
  t_rand, u = reference_draws(jr.split(jr.PRNGKey(3))[0], H * W, 256, 1, 64, 32, cuda_device)
  m = NerfModel(cfg, device=cuda_device)
  out = m.apply({'params': params}, {'origins': rays['origins'].reshape(-1, 3), 'directions': rays['directions'].reshape(-1, 3),
                                     'metadata': {'warp': torch.full((H * W, 1), 2, dtype=torch.int64, device=cuda_device)},
                                     'mask': torch.zeros((H * W, 1), device=cuda_device)},
                extra, use_predicted_norm=True, mask_ratio=1, sharp_weights_std=0.1, t_rand=t_rand, u=u,
                keys=('rgb', 'med_depth'), coarse_keys=())
  assert np.array_equal(out['fine']['rgb'].cpu().numpy().reshape(H, W, 3), frames[1]['rgb'])
  assert np.array_equal(out['fine']['med_depth'].cpu().numpy().reshape(H, W, -1).squeeze(-1),
                        frames[1]['med_depth'].reshape(H, W))

  with pytest.raises(IndexError):
    _make_dataset(str(tmp_path / 'big'), 7, W, H)
    render_scene(exp, str(tmp_path / 'big'), device=cuda_device, save=False)


def test_sort_camera_paths_follows_reference_string_order():
  assert sort_camera_paths(['c/000010.json', 'c/000002.json']) == ['c/000002.json', 'c/000010.json']
  assert sort_camera_paths(['c/left_10.json', 'c/left_2.json', 'c/right_1.json']) == [
      'c/right_1.json', 'c/left_10.json', 'c/left_2.json']                     # ids compare as strings
  with pytest.raises(ValueError):
    sort_camera_paths(['c/nonumber.json'])
