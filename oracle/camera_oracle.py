"""CPU ORACLE for ray generation.  TEST INFRASTRUCTURE ONLY.

numpy float32 restatement of ``datasets/core.py:51-76 camera_to_rays`` =
``Camera.pixels_to_rays(Camera.get_pixel_centers())`` (hypernerf/camera.py:226-270, 364-368) with the 10 Newton
iterations of ``_radial_and_tangential_undistort`` (camera.py:27-106).  Pinned against the reference's own
``camera.py`` (imported unmodified by tools/make_golden.py) in tests/test_oracle_golden.py.
"""
import numpy as np


def _residual_and_jacobian(x, y, xd, yd, k1, k2, k3, p1, p2):
  """camera.py:27-72."""
  r = x * x + y * y
  d = 1.0 + r * (k1 + r * (k2 + k3 * r))
  fx = d * x + 2 * p1 * x * y + p2 * (r + 2 * x * x) - xd
  fy = d * y + 2 * p2 * x * y + p1 * (r + 2 * y * y) - yd
  d_r = (k1 + r * (2.0 * k2 + 3.0 * k3 * r))
  d_x = 2.0 * x * d_r
  d_y = 2.0 * y * d_r
  fx_x = d + d_x * x + 2.0 * p1 * y + 6.0 * p2 * x
  fx_y = d_y * x + 2.0 * p1 * x + 2.0 * p2 * y
  fy_x = d_x * y + 2.0 * p2 * y + 2.0 * p1 * x
  fy_y = d + d_y * y + 2.0 * p2 * x + 6.0 * p1 * y
  return fx, fy, fx_x, fx_y, fy_x, fy_y


def undistort(xd, yd, k1, k2, k3, p1, p2, eps=1e-9, max_iterations=10):
  """camera.py:75-106."""
  x, y = xd.copy(), yd.copy()
  for _ in range(max_iterations):
    fx, fy, fx_x, fx_y, fy_x, fy_y = _residual_and_jacobian(x, y, xd, yd, k1, k2, k3, p1, p2)
    den = fy_x * fx_y - fx_x * fy_y
    xn = fx * fy_y - fy * fx_y
    yn = fy * fx_x - fx * fy_x
    ok = np.abs(den) > eps
    safe = np.where(ok, den, np.ones_like(den))
    x = x + np.where(ok, xn / safe, np.zeros_like(den))
    y = y + np.where(ok, yn / safe, np.zeros_like(den))
  return x, y


def camera_to_rays(orientation, position, focal_length, principal_point, image_size, skew=0.0,
                   pixel_aspect_ratio=1.0, radial_distortion=None, tangential_distortion=None):
  f = np.float32
  R = np.asarray(orientation, f).reshape(3, 3)
  pos = np.asarray(position, f).reshape(3)
  rd = np.zeros(3, f) if radial_distortion is None else np.asarray(radial_distortion, f)
  td = np.zeros(2, f) if tangential_distortion is None else np.asarray(tangential_distortion, f)
  W, H = int(image_size[0]), int(image_size[1])
  xx, yy = np.meshgrid(np.arange(W, dtype=f), np.arange(H, dtype=f))            # camera.py:364-368
  pixels = np.stack([xx, yy], axis=-1) + f(0.5)
  px = pixels.reshape(-1, 2)
  sx, sy = f(focal_length), f(focal_length) * f(pixel_aspect_ratio)
  y = (px[:, 1] - f(principal_point[1])) / sy                                    # camera.py:228-230
  x = (px[:, 0] - f(principal_point[0]) - y * f(skew)) / sx
  if np.any(rd != 0) or np.any(td != 0):
    x, y = undistort(x, y, rd[0], rd[1], rd[2], td[0], td[1])
  dirs = np.stack([x, y, np.ones_like(x)], axis=-1)
  dirs = dirs / np.linalg.norm(dirs, axis=-1, keepdims=True)                     # camera.py:242-243
  rays = np.squeeze(np.matmul(R.T, dirs[..., np.newaxis]), axis=-1)              # camera.py:263-264
  rays = rays / np.linalg.norm(rays, axis=-1, keepdims=True)                     # camera.py:267
  return {'origins': np.tile(pos[None, None, :], (H, W, 1)).astype(f),
          'directions': rays.reshape(H, W, 3).astype(f), 'pixels': pixels.astype(f)}
