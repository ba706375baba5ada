"""Host-side logic of the boundary on CPU: shard/unshard, render_image's
chunk / edge-pad / shard / unshard bookkeeping (evaluation.py:53-149), and the
N>1 path (world_size-2 gloo): per-rank shards + one all-gather per level."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerfds_b200 import evaluation, utils
from nerfds_b200.model_utils import TrainState

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeModel:
  """Deterministic stand-in for NerfModel.apply: outputs are functions of the
  ray origin, so any mis-ordering / bad padding shows up."""

  def __init__(self):
    self.calls = []
    self.renderer = self

  def set_max_chunk(self, n):
    pass

  def apply(self, variables, rays, extra_params, **kw):
    o = rays['origins'].float()
    self.calls.append(o.shape[0])
    keys = kw.get('keys') or ('rgb', 'depth')
    fine = {}
    if 'rgb' in keys:
      fine['rgb'] = o * 2.0 + rays['metadata']['warp'].float()
    if 'depth' in keys:
      fine['depth'] = o.sum(-1)
    if 'med_points' in keys:
      fine['med_points'] = o[:, None, :].repeat(1, 1, 1)
    return {'coarse': {}, 'fine': fine}


def _rays(h, w):
  n = h * w
  o = torch.arange(n * 3, dtype=torch.float32).reshape(h, w, 3)
  return {'origins': o, 'directions': torch.ones(h, w, 3),
          'metadata': {'warp': torch.full((h, w, 1), 3, dtype=torch.int64)},
          'mask': torch.zeros(h, w, 1)}


def test_shard_unshard_roundtrip():
  x = torch.arange(24).reshape(12, 2)
  s = utils.shard({'a': x}, 4)
  assert s['a'].shape == (4, 3, 2)
  assert torch.equal(utils.unshard(s['a']), x)
  assert torch.equal(utils.unshard(s['a'], padding=2), x[:-2])


@pytest.mark.parametrize('hw,chunk,D', [((5, 7), 8, 1), ((5, 7), 8, 4), ((6, 6), 36, 8), ((3, 3), 100, 2)])
def test_render_image_bookkeeping(hw, chunk, D):
  """Ragged last chunk, edge padding to a multiple of device_count, reshape to (H, W, ...)."""
  model = FakeModel()
  fn = evaluation.make_model_fn(model, keys=('rgb', 'depth', 'med_points'), group=False)
  rays = _rays(*hw)
  state = TrainState.create({'p': 1}, {'nerf_alpha': 8.0})
  out = evaluation.render_image(state, rays, fn, D, rng=np.array([0, 1], np.uint32), chunk=chunk)
  assert out['rgb'].shape == hw + (3,) and out['depth'].shape == hw and out['med_points'].shape == hw + (1, 3)
  assert torch.equal(out['rgb'], rays['origins'] * 2 + 3)
  assert torch.equal(out['depth'], rays['origins'].sum(-1))
  n = hw[0] * hw[1]
  sizes = [min(chunk, n - i) for i in range(0, n, chunk)]
  assert model.calls == [s + (-s) % D for s in sizes]           # padded to a device multiple


def test_train_state_extra_params():
  st = TrainState.create({'w': 1}, {'nerf_alpha': 8.0, 'warp_alpha': 4.0})
  assert st.optimizer.target['model'] == {'w': 1}
  ep = st.extra_params
  assert ep['nerf_alpha'] == 8.0 and ep['hyper_alpha'] is None and len(ep) == 8


def _worker(rank, world, port, tmp):
  os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
  sys.path.insert(0, ROOT)
  dist.init_process_group('gloo', rank=rank, world_size=world)
  try:
    model = FakeModel()
    rays = _rays(5, 7)                      # 35 rays: not divisible by 2
    state = TrainState.create({'p': 1}, {})
    # (1) reference-shaped render_image with per-chunk gather
    fn = evaluation.make_model_fn(model, keys=('rgb', 'depth'))
    out = evaluation.render_image(state, rays, fn, world, rng=np.array([0, 1], np.uint32), chunk=16)
    ok1 = torch.equal(out['rgb'], rays['origins'] * 2 + 3) and torch.equal(out['depth'], rays['origins'].sum(-1))
    per_call = model.calls[:]
    # (2) frame-sharded render with a single gather
    model2 = FakeModel()
    out2 = evaluation.render_image_sharded(model2, {'p': 1}, rays, {}, keys=('rgb', 'depth'), chunk=8)
    ok2 = torch.equal(out2['rgb'], rays['origins'] * 2 + 3) and out2['depth'].shape == (5, 7)
    torch.save({'ok1': ok1, 'ok2': ok2, 'calls': per_call, 'calls2': model2.calls}, os.path.join(tmp, f'r{rank}.pt'))
  finally:
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_render(tmp_path):
  port = 29500 + (os.getpid() % 2000)
  mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
  for r in range(2):
    res = torch.load(os.path.join(tmp_path, f'r{r}.pt'))
    assert res['ok1'] and res['ok2'], res
    assert res['calls'] == [8, 8, 2]        # each rank renders half of every (padded) chunk
    assert res['calls2'] == [18]            # ceil(35 / 2) rays per rank, one call
