"""Ray generation on the device: the host-side mirror of the reference's camera interface for this path.

``Camera`` carries the attributes of hypernerf/camera.py:109-138 (same names, float32), ``camera_to_rays`` replaces
datasets/core.py:51-76 -- ``Camera.pixels_to_rays(Camera.get_pixel_centers())`` incl. the radial / tangential
undistortion -- with one CUDA kernel behind the C ABI (``ndsr_camera_rays``), so a frame's rays never cross PCIe.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import json
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib


@dataclasses.dataclass
class Camera:
  orientation: Sequence[Sequence[float]]
  position: Sequence[float]
  focal_length: float
  principal_point: Sequence[float]
  image_size: Sequence[int]                      # (width, height)
  skew: float = 0.0
  pixel_aspect_ratio: float = 1.0
  radial_distortion: Optional[Sequence[float]] = None
  tangential_distortion: Optional[Sequence[float]] = None

  @classmethod
  def from_json(cls, path):
    """camera.py:141-161 (the nerfies camera JSON)."""
    with open(path, 'r') as fp:
      d = json.load(fp)
    if 'tangential' in d:                        # camera.py:146-147 (legacy key)
      d['tangential_distortion'] = d['tangential']
    return cls(orientation=d['orientation'], position=d['position'], focal_length=d['focal_length'],
               principal_point=d['principal_point'], image_size=d['image_size'], skew=d.get('skew', 0.0),
               pixel_aspect_ratio=d.get('pixel_aspect_ratio', 1.0), radial_distortion=d.get('radial_distortion'),
               tangential_distortion=d.get('tangential_distortion'))

  @property
  def image_shape(self):
    return int(self.image_size[1]), int(self.image_size[0])

  def scale(self, scale: float) -> 'Camera':
    """camera.py:370-387: intrinsics and image size scaled, extrinsics and distortion kept."""
    if scale <= 0:
      raise ValueError('scale needs to be positive.')
    return dataclasses.replace(
        self, focal_length=self.focal_length * scale,
        principal_point=[float(v) * scale for v in self.principal_point],
        image_size=[int(round(self.image_size[0] * scale)), int(round(self.image_size[1] * scale))])

  def to_c(self) -> _lib.ndsr_camera:
    c = _lib.ndsr_camera()
    f32 = lambda a, n: np.asarray(a if a is not None else np.zeros(n), np.float32).reshape(n)
    c.orientation[:] = f32(self.orientation, 9).tolist()
    c.position[:] = f32(self.position, 3).tolist()
    c.focal_length = float(np.float32(self.focal_length))
    c.principal_point[:] = f32(self.principal_point, 2).tolist()
    c.skew = float(np.float32(self.skew))
    c.pixel_aspect_ratio = float(np.float32(self.pixel_aspect_ratio))
    c.radial_distortion[:] = f32(self.radial_distortion, 3).tolist()
    c.tangential_distortion[:] = f32(self.tangential_distortion, 2).tolist()
    c.image_size[:] = [int(self.image_size[0]), int(self.image_size[1])]
    return c


def load_camera(camera_path, scale_factor: float = 1.0, scene_center=None, scene_scale=None) -> Camera:
  """datasets/core.py:79-111: JSON camera, rescaled, moved into the normalised scene frame."""
  if not str(camera_path).endswith('.json'):
    raise ValueError('File must have extension .pb or .json.' if not str(camera_path).endswith('.pb')
                     else 'camera protos are not supported; export the camera as JSON')
  camera = Camera.from_json(camera_path)
  if scale_factor != 1.0:
    camera = camera.scale(scale_factor)
  pos = np.asarray(camera.position, np.float64)
  if scene_center is not None:
    pos = pos - np.asarray(scene_center, np.float64)
  if scene_scale is not None:
    pos = pos * scene_scale
  return dataclasses.replace(camera, position=pos.tolist())


def camera_to_rays(camera: Camera, device='cuda:0') -> Dict[str, torch.Tensor]:
  """datasets/core.py:51-76: ``{'origins', 'directions', 'pixels'}`` as (H, W, .) float32 CUDA tensors."""
  lib = _lib.load_library()
  dev = torch.device(device)
  if dev.type != 'cuda' or not torch.cuda.is_available():
    raise RuntimeError('camera_to_rays runs on a CUDA device (there is no CPU fallback)')
  H, W = camera.image_shape
  o = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
  d = torch.empty((H, W, 3), dtype=torch.float32, device=dev)
  p = torch.empty((H, W, 2), dtype=torch.float32, device=dev)
  cc = camera.to_c()
  with torch.cuda.device(dev):
    st = torch.cuda.current_stream(dev).cuda_stream
    rc = lib.ndsr_camera_rays(dev.index or 0, C.c_void_p(st), C.byref(cc), C.c_void_p(o.data_ptr()),
                              C.c_void_p(d.data_ptr()), C.c_void_p(p.data_ptr()))
  if rc != 0:
    raise RuntimeError(f'ndsr_camera_rays failed with code {rc}')
  return {'origins': o, 'directions': d, 'pixels': p}
