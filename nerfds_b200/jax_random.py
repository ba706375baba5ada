"""JAX-compatible random keys and uniform draws.

The reference draws the stratified jitter `t_rand` and the inverse-CDF samples `u`
with `jax.random.uniform(self.make_rng('coarse' | 'fine'), [B, S])`
(model_utils.py:84, 217; models.py:1489, 1524).  To reproduce a JAX run bit
for bit given the same `rngs` three pieces are needed:

  1. jax's default PRNG (jax 0.3.15, `jax_default_prng_impl = threefry2x32`,
     not partitionable): `PRNGKey`, `split`, `fold_in` -- a handful of
     Threefry-2x32 blocks on the host (python ints, this file);
  2. flax's `Scope.make_rng` key derivation (flax 0.5.3 core/scope.py): the k-th
     `make_rng(name)` call of a scope folds `sha1(path items + k)[:4]` into the
     scope's key for `name`; `NerfModel` is the top-level module, so the path
     is empty and `t_rand` / `u` both come from call k = 1 of their stream
     (the later `make_rng(level)` calls at models.py:574 feed a noise term
     that is disabled when noise_std is None);
  3. `random.uniform` over [B, S] -- bulk work, done on the device by
     `ndsr_random_uniform` (csrc/nds_composite.cu, one Threefry block per
     thread); there is no host fallback for it.

Pinned by known answers: the three Random123 Threefry-2x32 vectors,
`split(PRNGKey(0))` and `uniform(PRNGKey(0))` from the JAX documentation
(tests/test_jax_random.py).  The flax folding rule (2) is restated from the
published source and is NOT pinned (flax is not installable here).
"""
from __future__ import annotations

import ctypes as C
import hashlib
from typing import Iterable, Sequence, Tuple, Union

import numpy as np

M32 = 0xFFFFFFFF
_R0, _R1 = (13, 15, 26, 6), (17, 29, 16, 24)

Key = Tuple[int, int]


def threefry2x32(key: Key, x0: int, x1: int) -> Tuple[int, int]:
  """One Threefry-2x32-20 block (jax/_src/prng.py `_threefry2x32_lowering`)."""
  k0, k1 = key
  ks = (k0, k1, k0 ^ k1 ^ 0x1BD11BDA)
  x0, x1 = (x0 + ks[0]) & M32, (x1 + ks[1]) & M32
  for g in range(5):
    for r in (_R0 if g % 2 == 0 else _R1):
      x0 = (x0 + x1) & M32
      x1 = ((x1 << r) | (x1 >> (32 - r))) & M32
      x1 ^= x0
    x0 = (x0 + ks[(g + 1) % 3]) & M32
    x1 = (x1 + ks[(g + 2) % 3] + g + 1) & M32
  return x0, x1


def as_key(key) -> Key:
  """uint32[2] array-like / torch tensor / python int seed -> (k0, k1)."""
  if isinstance(key, (int, np.integer)):
    return PRNGKey(int(key))
  if hasattr(key, 'detach'):
    key = key.detach().cpu().numpy()
  a = np.asarray(key).reshape(-1)
  if a.size != 2:
    raise ValueError(f'a jax PRNG key is uint32[2], got shape {np.asarray(key).shape}')
  return int(a[0]) & M32, int(a[1]) & M32


def PRNGKey(seed: int) -> Key:
  """random.PRNGKey: the 64-bit seed as (high word, low word)."""
  seed = int(seed) & 0xFFFFFFFFFFFFFFFF
  return (seed >> 32) & M32, seed & M32


def _random_bits(key: Key, n: int) -> list:
  """threefry_2x32(key, iota(n)): counters split in halves, odd n padded with one zero."""
  half = (n + 1) // 2
  a, b = [0] * half, [0] * half
  for j in range(half):
    a[j], b[j] = threefry2x32(key, j, j + half if j + half < n else 0)
  return (a + b)[:n]


def split(key, num: int = 2) -> list:
  """random.split: `num` new keys."""
  bits = _random_bits(as_key(key), 2 * num)
  return [(bits[2 * i], bits[2 * i + 1]) for i in range(num)]


def fold_in(key, data: int) -> Key:
  """random.fold_in: threefry_2x32(key, threefry_seed(data)) with a 32-bit `data`."""
  return threefry2x32(as_key(key), 0, int(data) & M32)


def flax_make_rng(key, path: Sequence[Union[str, int]] = (), count: int = 1) -> Key:
  """flax 0.5.3 `Scope.make_rng`: LazyRng.create(rngs[name], *path, count).as_jax_rng().

  All suffix items are hashed together (str -> utf-8, int -> minimal big-endian
  bytes), the first 4 digest bytes (big endian) are folded into the key.
  """
  m = hashlib.sha1()
  for x in tuple(path) + (int(count),):
    if isinstance(x, str):
      m.update(x.encode('utf-8'))
    elif isinstance(x, (int, np.integer)):
      x = int(x)
      m.update(x.to_bytes((x.bit_length() + 7) // 8, byteorder='big'))
    else:
      raise ValueError(f'Expected int or string, got: {x}')
  return fold_in(key, int.from_bytes(m.digest()[:4], byteorder='big'))


def uniform(key, shape: Iterable[int], device=None):
  """random.uniform(key, shape) as a float32 CUDA tensor, generated on the device."""
  import torch
  from . import _lib
  lib = _lib.load_library()
  dev = torch.device('cuda:0' if device is None else device)
  if dev.type != 'cuda' or not torch.cuda.is_available():
    raise RuntimeError('jax_random.uniform runs on a CUDA device (there is no CPU fallback)')
  shape = tuple(int(s) for s in shape)
  out = torch.empty(shape, dtype=torch.float32, device=dev)
  k = (C.c_uint32 * 2)(*as_key(key))
  with torch.cuda.device(dev):
    rc = lib.ndsr_random_uniform(dev.index or 0, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream), k,
                                 out.numel(), C.c_void_p(out.data_ptr() if out.numel() else 0))
  if rc != 0:
    raise RuntimeError(f'ndsr_random_uniform failed with code {rc}')
  return out
