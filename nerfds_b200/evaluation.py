"""``render_image`` -- hypernerf/evaluation.py:53-149 on top of the CUDA path.

Two entry points:
  * ``render_image(state, rays_dict, model_fn, device_count, rng, chunk, ...)``
    keeps the reference's signature, chunking, edge padding and
    (device_count, n/device_count, ...) sharding, so a render.py-style driver
    works unchanged with ``model_fn = make_model_fn(model)``;
  * ``render_image_sharded`` is the B200 layout: the frame's rays are
    block-partitioned over the ranks of a torch.distributed group once, each
    rank renders its block, and ONE all-gather per frame reassembles it.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Iterable, Optional

import numpy as np
import torch

from . import jax_random, utils
from .renderer import RENDER_KEYS


def _dist():
  import torch.distributed as dist
  return dist if (dist.is_available() and dist.is_initialized()) else None


def make_model_fn(model, *, use_predicted_norm: bool = True, keys: Iterable[str] = RENDER_KEYS, group=None,
                  t_rand_fn: Optional[Callable] = None) -> Callable:
  """The pmapped ``_model_fn`` of render.py:139-163 for one process per GPU.

  Takes ``(key_0, key_1, key_2, params, rays_dict, extra_params)`` with leaves
  shaped (device_count, n/device_count, C).  With a process group of W ranks,
  device_count must equal W: rank r renders shard r and the per-ray outputs are
  all-gathered (render.py:155).  Leaves come back as (1, device_count,
  n/device_count, ...) so that evaluation.py:126-129's ``x[0]`` + ``unshard``
  apply unchanged.
  """
  keys = tuple(keys)

  def model_fn(key_0, key_1, key_2, params, rays_dict, extra_params):
    dist = _dist() if group is not False else None
    world = dist.get_world_size(group) if dist else 1
    rank = dist.get_rank(group) if dist else 0
    D = rays_dict['origins'].shape[0]
    if dist and D != world:
      raise ValueError(f'device_count ({D}) must equal the process-group size ({world})')
    if dist:   # this rank's shard
      local = utils.tree_map(lambda x: x[rank], rays_dict)
      k0, k1 = key_0[rank], key_1[rank]
    else:      # single process: all shards on this GPU, in order
      local = utils.tree_map(lambda x: x.reshape((-1,) + tuple(x.shape[2:])), rays_dict)
      k0, k1 = key_0[0], key_1[0]
    t_rand = u = None
    if t_rand_fn is not None:
      t_rand, u = t_rand_fn(local)
    elif not dist and D > 1 and getattr(getattr(model, 'cfg', None), 'use_stratified_sampling', False):
      # under pmap every device draws its shard's samples from its own key (evaluation.py:83-84)
      per = rays_dict['origins'].shape[1]
      draw = lambda ks, S: torch.cat([jax_random.uniform(jax_random.flax_make_rng(ks[d]), (per, S), model.device)
                                      for d in range(D)], 0)
      t_rand, u = draw(key_0, model.cfg.num_coarse_samples), draw(key_1, model.cfg.num_fine_samples)
    out = model.apply({'params': params}, local, extra_params, rngs={'coarse': k0, 'fine': k1, 'voxel': key_2},
                      mutable=False, use_predicted_norm=use_predicted_norm, return_points=False,
                      return_nv_details=False, mask_ratio=1, sharp_weights_std=0.1,   # render.py:150-153
                      keys=keys, coarse_keys=(), t_rand=t_rand, u=u)
    fine = out['fine']
    if dist:
      fine = all_gather_level(fine, group)             # (W, n/W, ...)
      res = {k: v[None] for k, v in fine.items()}
    else:
      res = {k: v.reshape((D, -1) + tuple(v.shape[1:]))[None] for k, v in fine.items()}
    return {'fine': res}

  return model_fn


def all_gather_level(level: Dict[str, torch.Tensor], group=None) -> Dict[str, torch.Tensor]:
  """One NCCL all-gather for a whole level dict: leaves (n, ...) -> (W, n, ...)."""
  import torch.distributed as dist
  W = dist.get_world_size(group)
  names = sorted(level)
  flat = torch.cat([level[k].reshape(-1) for k in names]) if names else torch.empty(0)
  gathered = torch.empty(W * flat.numel(), dtype=flat.dtype, device=flat.device)
  dist.all_gather_into_tensor(gathered, flat.contiguous(), group=group)
  gathered = gathered.view(W, flat.numel())
  out, off = {}, 0
  for k in names:
    n = level[k].numel()
    out[k] = gathered[:, off:off + n].reshape((W,) + tuple(level[k].shape))
    off += n
  return out


def reference_draws(rng, num_rays: int, chunk: int, device_count: int, num_coarse: int, num_fine: int, device):
  """The stratified / inverse-CDF draws of a whole frame exactly as `render_image` + a pmapped model make them
  (evaluation.py:81-84, 94-120; models.py:1489, 1524): every chunk of `chunk` rays is edge-padded to a multiple of
  `device_count` and sharded, device d draws `uniform(make_rng(key_i[d]), [rows, S])` for its rows -- the SAME keys
  for every chunk.  Returns (t_rand [num_rays, num_coarse], u [num_rays, num_fine]) in frame order."""
  _, key_0, key_1, _ = jax_random.split(rng, 4)
  key_0, key_1 = jax_random.split(key_0, device_count), jax_random.split(key_1, device_count)
  tables = {}

  def table(rows):
    if rows not in tables:
      tables[rows] = tuple(torch.cat([jax_random.uniform(jax_random.flax_make_rng(ks[d]), (rows, S), device)
                                      for d in range(device_count)], 0)
                           for ks, S in ((key_0, num_coarse), (key_1, num_fine)))
    return tables[rows]

  ts, us = [], []
  for ray_idx in range(0, num_rays, chunk):
    n = min(chunk, num_rays - ray_idx)
    t, u = table((n + device_count - 1) // device_count)
    ts.append(t[:n])
    us.append(u[:n])
  return torch.cat(ts, 0), torch.cat(us, 0)


def render_image(state, rays_dict, model_fn, device_count, rng, chunk=8192, default_ret_key=None):
  """hypernerf/evaluation.py:53-149 (same arguments, same chunk/pad/shard logic).

  Returns a dict of (H, W, ...) torch tensors on the host (the reference moves
  every chunk to the CPU device, evaluation.py:126).
  """
  batch_shape = tuple(rays_dict['origins'].shape[:-1])
  num_rays = int(np.prod(batch_shape))
  as_t = lambda x: x if torch.is_tensor(x) else torch.from_numpy(np.ascontiguousarray(np.asarray(x)).astype(
      np.int64) if np.asarray(x).dtype == np.uint32 else np.ascontiguousarray(np.asarray(x)))
  rays_dict = utils.tree_map(lambda x: as_t(x).reshape((num_rays, -1)), rays_dict)
  # _, key_0, key_1, key_2 = split(rng, 4); key_i = split(key_i, device_count)  (evaluation.py:81-84)
  _, key_0, key_1, key_2 = jax_random.split(rng, 4)
  key_0, key_1, key_2 = (np.array(jax_random.split(k, device_count), np.uint32) for k in (key_0, key_1, key_2))
  ret_maps = []
  num_batches = int(math.ceil(num_rays / chunk))
  for batch_idx in range(num_batches):
    ray_idx = batch_idx * chunk
    chunk_rays = utils.tree_map(lambda x: x[ray_idx:ray_idx + chunk], rays_dict)
    num_chunk_rays = chunk_rays['origins'].shape[0]
    remainder = num_chunk_rays % device_count
    if remainder != 0:
      padding = device_count - remainder
      pad = lambda x: torch.cat([x, x[-1:].expand((padding,) + tuple(x.shape[1:]))], 0)   # mode='edge'
      chunk_rays = utils.tree_map(pad, chunk_rays)
    else:
      padding = 0
    chunk_rays = utils.shard(chunk_rays, device_count)
    model_out = model_fn(key_0, key_1, key_2, state.optimizer.target['model'], chunk_rays, state.extra_params)
    if not default_ret_key:
      ret_key = 'fine' if 'fine' in model_out else 'coarse'
    else:
      ret_key = default_ret_key
    ret_map = utils.tree_map(lambda x: x[0].cpu(), model_out[ret_key])      # unreplicate + device_put(cpu)
    ret_map = utils.tree_map(lambda x: utils.unshard(x, padding), ret_map)
    ret_maps.append(ret_map)
  out = {}
  for key in ret_maps[0]:
    values = torch.cat([m[key] for m in ret_maps], 0)
    out[key] = values.reshape(batch_shape + tuple(values.shape[1:]))
  return out


def render_image_sharded(model, params, rays_dict, extra_params, *, t_rand=None, u=None, chunk=65536,
                         keys: Iterable[str] = RENDER_KEYS, use_predicted_norm=True, group=None,
                         gather=True, peer_frames=None) -> Dict[str, torch.Tensor]:
  """Block-partition one frame over the process group, render, all-gather once.

  Every rank passes the same full-frame ``rays_dict``; rank r renders rays
  [r*n/W, (r+1)*n/W) (the partition ``utils.shard`` makes, after edge-padding
  to a multiple of W as evaluation.py:100-107 does) in chunks of ``chunk`` and
  a single all-gather of the packed per-ray outputs reassembles the frame on
  every rank.  Without an initialised process group this is a 1-rank render.

  peer_frames: a `peer.PeerFrames` built for this frame size and these keys -- the compositing kernels then store
  every rank's rays straight into every GPU's frame buffer (NVLink peer stores) and no collective moves data;
  the returned tensors are views of this rank's buffer, valid until the next frame is rendered into it.
  """
  dist = _dist()
  W = dist.get_world_size(group) if dist else 1
  rank = dist.get_rank(group) if dist else 0
  batch_shape = tuple(rays_dict['origins'].shape[:-1])
  n = int(np.prod(batch_shape))
  flat = utils.tree_map(lambda x: (x if torch.is_tensor(x) else torch.from_numpy(
      np.ascontiguousarray(np.asarray(x).astype(np.int64) if np.asarray(x).dtype == np.uint32 else np.asarray(x)))
  ).reshape((n, -1)), rays_dict)
  per = (n + W - 1) // W
  lo, hi = rank * per, min(n, (rank + 1) * per)
  sl = lambda x: None if x is None else x[lo:hi]
  local = utils.tree_map(lambda x: x[lo:hi], flat)
  pad = per - (hi - lo)
  if pad > 0:     # edge padding of the last shard
    local = utils.tree_map(lambda x: torch.cat([x, x[-1:].expand((pad,) + tuple(x.shape[1:]))], 0), local)
    ext = lambda x: None if x is None else torch.cat([torch.as_tensor(x)[lo:hi], torch.as_tensor(x)[hi - 1:hi].expand(
        (pad,) + tuple(torch.as_tensor(x).shape[1:]))], 0)
  else:
    ext = lambda x: None if x is None else torch.as_tensor(x)[lo:hi]
  model.renderer.set_max_chunk(chunk)
  if peer_frames is not None:
    if peer_frames.n != n or tuple(peer_frames.keys) != tuple(keys):
      raise ValueError('peer_frames was built for another frame size / key set')
    if pad > 0:       # the padded tail of the last shard must not be written past the frame: render the real rays only
      local = utils.tree_map(lambda x: x[:hi - lo], local)
      ext = lambda x: None if x is None else torch.as_tensor(x)[lo:hi]
    peer_frames.activate()
    if hi > lo:
      model.apply({'params': params}, local, extra_params, use_predicted_norm=use_predicted_norm, mask_ratio=1,
                  sharp_weights_std=0.1, keys=tuple(keys), coarse_keys=(), t_rand=ext(t_rand), u=ext(u),
                  fine_ptrs=peer_frames.shard_ptrs(lo))
    peer_frames.wait()
    return {k: v.reshape(batch_shape + tuple(v.shape[1:])) for k, v in peer_frames.frame().items()}
  out = model.apply({'params': params}, local, extra_params, use_predicted_norm=use_predicted_norm, mask_ratio=1,
                    sharp_weights_std=0.1, keys=tuple(keys), coarse_keys=(), t_rand=ext(t_rand), u=ext(u))
  fine = out['fine']
  if dist and gather:
    g = all_gather_level(fine, group)
    fine = {k: v.reshape((W * per,) + tuple(v.shape[2:]))[:n] for k, v in g.items()}
  elif pad > 0:
    fine = {k: v[:hi - lo] for k, v in fine.items()}
  if dist and not gather:
    return fine
  return {k: v.reshape(batch_shape + tuple(v.shape[1:])) for k, v in fine.items()}
