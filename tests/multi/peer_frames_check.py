"""torchrun entry: peer-memory frame reassembly == NCCL all-gather reassembly == single-GPU render, bit for bit.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/multi/peer_frames_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
  local = int(os.environ.get('LOCAL_RANK', '0'))
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  dist.init_process_group('nccl', device_id=dev)
  rank, world = dist.get_rank(), dist.get_world_size()
  from nerfds_b200 import synthetic as syn
  from nerfds_b200.evaluation import render_image_sharded
  from nerfds_b200.models import NerfModel
  from nerfds_b200.peer import PeerFrames
  from nerfds_b200.renderer import RENDER_KEYS
  from tests.common import make_case
  cfg, params, rays, t_rand, u = make_case('nerf_ds', image=21, seed=3, num_coarse_samples=32, num_fine_samples=32)   # 441 rays: ragged shards
  n = rays['origins'].shape[0]
  model = NerfModel(cfg, device=dev)
  ep = syn.final_extra_params()
  ref = render_image_sharded(model, params, rays, ep, t_rand=t_rand, u=u, chunk=100, keys=RENDER_KEYS)           # NCCL all-gather
  frames = PeerFrames(model.renderer, n, RENDER_KEYS)
  for rep in range(2):                                                                                           # buffer reuse
    got = render_image_sharded(model, params, rays, ep, t_rand=t_rand, u=u, chunk=100, keys=RENDER_KEYS, peer_frames=frames)
    for k in RENDER_KEYS:
      a, b = got[k].reshape(ref[k].shape), ref[k]
      assert torch.equal(a, b), (rank, rep, k, float((a - b).abs().max()))
    dist.barrier()
  # An ordinary call while mirrors are registered: its outputs lie outside the frame buffer, so the library does not
  # mirror them (no store into a peer's mapping) and the frame buffers keep the last frame.
  snap = {k: v.clone() for k, v in frames.frame().items()}
  plain = model.apply({'params': params}, rays, ep, t_rand=t_rand, u=u, use_predicted_norm=True, mask_ratio=1,
                      sharp_weights_std=0.1, keys=('rgb',), coarse_keys=())
  assert torch.equal(plain['fine']['rgb'], ref['rgb'])
  extra = model.renderer.make_extra(ep, use_predicted_norm=True, mask_ratio=1, sharp_weights_std=0.1)
  host = model.renderer.render_rays_host(rays['origins'], rays['directions'], warp_id=rays['metadata']['warp'].reshape(-1),
                                         gt_mask=rays['mask'].reshape(-1), t_rand=t_rand, u=u, extra=extra, fine_keys=('rgb',))
  assert np.array_equal(host['rgb'], ref['rgb'].cpu().numpy())
  torch.cuda.synchronize()
  dist.barrier()
  for k, v in frames.frame().items():
    assert torch.equal(v, snap[k]), (rank, k)
  # ... and a call with per-ray outputs on both sides of the frame buffer is refused
  ptrs = frames.shard_ptrs(0)
  bad = dict(ptrs, depth=int(plain['fine']['rgb'].data_ptr()))
  try:
    model.renderer.render_rays(rays['origins'][:8], rays['directions'][:8], warp_id=rays['metadata']['warp'][:8],
                               gt_mask=rays['mask'][:8], t_rand=t_rand[:8], u=u[:8], extra=extra, coarse_keys=(),
                               fine_keys=('rgb', 'depth'), fine_ptrs=bad)
    refused = world == 1                                   # a single rank registers no mirrors: nothing to refuse
  except Exception as e:
    refused = 'partly inside' in str(e)
  assert refused
  frames.close()
  # mirrors are off again: a plain call writes only locally
  out = model.apply({'params': params}, rays, ep, t_rand=t_rand, u=u, use_predicted_norm=True, mask_ratio=1,
                    sharp_weights_std=0.1, keys=('rgb',), coarse_keys=())
  assert torch.equal(out['fine']['rgb'], ref['rgb'])
  dist.barrier()
  if rank == 0:
    print(f'peer frames OK on {world} GPUs ({n} rays)')
  dist.destroy_process_group()


if __name__ == '__main__':
  main()
