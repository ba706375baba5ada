"""Multi-GPU frame reassembly over peer memory (no collective on the data path).

`render.py:155` ends every pmapped chunk with `jax.lax.all_gather(out, 'batch')`.
With one process per GPU the same result is produced INSIDE the compositing
kernel: every rank owns an identical packed frame buffer, maps the other ranks'
buffers (CUDA IPC over NVLink / NVSwitch) and registers them as *mirrors* of
its own; the kernel then stores each per-ray result at the ray's place in the
local buffer and at the same offset of every mirror.  When all ranks' streams
have drained (one barrier), every GPU holds the whole frame.  `torch.distributed`
is used for the 64-byte handle exchange and the barrier only.

    frames = PeerFrames(model.renderer, n_rays, keys)       # collective: all ranks
    frames.activate()                                       # (only needed when several PeerFrames alternate)
    out = model.apply(..., fine_ptrs=frames.shard_ptrs(lo)) # this rank's rays [lo, hi)
    frames.wait()                                           # stream sync + barrier
    full = frames.frame()                                   # {key: [n_rays, ...]} on this GPU
    frames.close()                                          # unregisters the mirrors, unmaps, frees

The library mirrors a fine-level call only when ALL its per-ray output pointers lie inside the registered frame
buffer (`fine_ptrs=frames.shard_ptrs(...)`); ordinary calls (fresh tensors, render_samples, render_rays_host) made
while a PeerFrames is active are left alone, and a call with pointers on both sides is refused.

Lifetime: `close()` is collective.  It unregisters the mirrors, waits until no rank still stores into any buffer
(stream sync + barrier), unmaps the peers' buffers, and only after a SECOND barrier frees its own exported buffer
(freeing exported memory that an importer still maps is undefined behaviour).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional

import numpy as np
import torch

from . import _lib
from .renderer import PER_RAY, Renderer


class _CudaArray:
  """`__cuda_array_interface__` view of raw device memory, for torch.as_tensor."""

  def __init__(self, ptr: int, shape):
    self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': '<f4', 'data': (int(ptr), False), 'version': 2}


class PeerFrames:
  def __init__(self, renderer: Renderer, n_rays: int, keys: Iterable[str], group=None):
    import torch.distributed as dist
    self.R, self.lib, self.group = renderer, renderer.lib, group
    self.dev = renderer.device
    self.n = int(n_rays)
    self.keys = tuple(keys)
    for k in self.keys:
      if k not in PER_RAY:
        raise KeyError(f'{k!r} is not a per-ray output: only per-ray results are mirrored')
    self.shapes = {k: tuple(PER_RAY[k](0, renderer.H)) for k in self.keys}
    self.width = {k: int(np.prod(self.shapes[k], dtype=np.int64)) for k in self.keys}
    self.offset, off = {}, 0
    for k in self.keys:                                  # packed: key after key, each [n_rays, width], 256 B aligned
      self.offset[k] = off
      off += (self.n * self.width[k] * 4 + 255) & ~255
    self.bytes = max(off, 256)
    self.world = dist.get_world_size(group) if dist.is_initialized() else 1
    self.rank = dist.get_rank(group) if dist.is_initialized() else 0
    if self.world - 1 > _lib.NDSR_MAX_MIRRORS:
      raise ValueError(f'at most {_lib.NDSR_MAX_MIRRORS + 1} ranks')
    idx = self.dev.index or 0
    base, handle = C.c_void_p(), C.create_string_buffer(64)
    self._check(self.lib.ndsr_peer_alloc(idx, self.bytes, C.byref(base), handle), 'ndsr_peer_alloc')
    self.base = int(base.value)
    self.mapped = {}
    if self.world > 1:
      handles = [None] * self.world
      dist.all_gather_object(handles, bytes(handle.raw), group=group)
      for r, hb in enumerate(handles):
        if r == self.rank:
          continue
        p = C.c_void_p()
        self._check(self.lib.ndsr_peer_open(idx, hb, C.byref(p)), f'ndsr_peer_open(rank {r})')
        self.mapped[r] = int(p.value)
    self._deltas = (C.c_int64 * max(1, len(self.mapped)))(*[p - self.base for p in self.mapped.values()])
    self.closed = False
    self.activate()

  def activate(self):
    """Make this buffer set the mirror target of the renderer's next calls (several PeerFrames can alternate,
    e.g. to let a consumer read frame f while frame f + 1 is being written)."""
    self.R._check(self.lib.ndsr_set_output_mirrors(self.R._h, len(self.mapped), self._deltas, C.c_void_p(self.base),
                                                   self.bytes), 'ndsr_set_output_mirrors')

  @staticmethod
  def _check(rc, what):
    if rc != 0:
      raise RuntimeError(f'{what} failed with code {rc} (peer access between the GPUs of this box is required)')

  def shard_ptrs(self, lo: int) -> Dict[str, int]:
    """Output pointers for a call that renders rays [lo, lo + B) of the frame."""
    return {k: self.base + self.offset[k] + int(lo) * self.width[k] * 4 for k in self.keys}

  def wait(self):
    """Every rank's stores have landed everywhere."""
    import torch.distributed as dist
    torch.cuda.current_stream(self.dev).synchronize()
    if self.world > 1:
      dist.barrier(group=self.group)

  def frame(self) -> Dict[str, torch.Tensor]:
    """The assembled frame on this rank's GPU (views of the frame buffer: copy before the next frame overwrites it)."""
    return {k: torch.as_tensor(_CudaArray(self.base + self.offset[k], (self.n,) + self.shapes[k]), device=self.dev)
            for k in self.keys}

  def close(self):
    if self.closed:
      return
    self.closed = True
    idx = self.dev.index or 0
    import torch.distributed as dist
    self.R._check(self.lib.ndsr_set_output_mirrors(self.R._h, 0, None, None, 0), 'ndsr_set_output_mirrors')
    self.wait()                                          # nobody still writes into a buffer about to go away
    for p in self.mapped.values():
      self.lib.ndsr_peer_close(idx, C.c_void_p(p))
    if self.world > 1:
      dist.barrier(group=self.group)                     # every importer has unmapped before any exporter frees
    self.lib.ndsr_peer_free(idx, C.c_void_p(self.base))
