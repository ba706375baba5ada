"""Boundary helpers of the path: hypernerf/utils.py:295-312 (shard / unshard)."""
from __future__ import annotations

import numpy as np
import torch


def _map(fn, tree):
  if isinstance(tree, dict):
    return {k: _map(fn, v) for k, v in tree.items()}
  return fn(tree)


tree_map = _map


def shard(xs, device_count: int):
  """Split data into shards along the first dimension (utils.py:295-299)."""
  return _map(lambda x: x.reshape((device_count, -1) + tuple(x.shape[1:])), xs)


def unshard(x, padding: int = 0):
  """Collect the sharded tensor to the shape before sharding (utils.py:307-312)."""
  y = x.reshape((x.shape[0] * x.shape[1],) + tuple(x.shape[2:]))
  return y[:-padding] if padding > 0 else y


def to_numpy(tree):
  return _map(lambda x: x.detach().cpu().numpy() if torch.is_tensor(x) else np.asarray(x), tree)
