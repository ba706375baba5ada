"""torchrun entry: render_scene with one process per GPU == render_scene in one process, bit for bit."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  from nerfds_b200 import checkpoints as ckpt
  from nerfds_b200.config import nerf_ds_config
  from nerfds_b200.model_utils import TrainState
  from nerfds_b200.params import init_params
  from nerfds_b200.render import render_scene
  from tests.test_experiment_io import GIN_BASE, GIN_MAIN
  from tests.test_gpu_experiment import _make_dataset
  root = os.environ['NDS_CHECK_DIR']
  data, exp = os.path.join(root, 'data'), os.path.join(root, 'exp')
  rank = int(os.environ.get('RANK', '0'))
  if rank == 0:
    _make_dataset(data, 4, 22, 14)
    cfg = nerf_ds_config(num_coarse_samples=64, num_fine_samples=32, near=0.1, far=2.5, num_warp_embeds=5)
    state = TrainState.create(init_params(cfg, 11), {'nerf_alpha': 8.0, 'warp_alpha': 3.0, 'hyper_alpha': 1.0,
                                                    'hyper_sheet_alpha': 6.0, 'norm_input_alpha': 4.0})
    state.optimizer.state.step = np.int32(7)
    os.makedirs(exp)
    with open(os.path.join(exp, 'config.gin'), 'w') as f:
      f.write(GIN_BASE + GIN_MAIN.replace("include 'base.gin'", '') +
              "\nExperimentConfig.image_scale = 2\nSpecularConfig.use_predicted_norm = True\n")
    ckpt.save_checkpoint(os.path.join(exp, 'checkpoints'), state, 7)
  world = int(os.environ.get('WORLD_SIZE', '1'))
  # (the draws follow the reference's per-device keys, evaluation.py:81-84: same device_count in both runs)
  single = render_scene(exp, data, chunk_size=100, device=dev, save=False, device_count=world) if rank == 0 else None     # before the group exists
  dist.init_process_group('nccl', device_id=dev)
  dist.barrier()
  multi = render_scene(exp, data, chunk_size=100, device=dev, save=True)
  assert len(multi) == 4
  if rank == 0:
    for a, b in zip(single, multi):
      for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert os.path.exists(os.path.join(exp, 'render_result_vrig_camera_full'))
    print(f'render_scene OK on {dist.get_world_size()} GPUs')
  dist.barrier()
  dist.destroy_process_group()


if __name__ == '__main__':
  main()
