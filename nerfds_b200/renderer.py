"""Thin object wrapper over the C-ABI handle (include/nerfds_b200.h).

PyTorch is plumbing here: it owns device memory and streams; every compute
call goes through libnerfds_b200.so.  There is no CPU or eager fallback.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Iterable, Optional

import numpy as np
import torch

from . import _lib
from .config import NerfDSConfig
from .params import flatten_params

# level-dict keys -> (C field, trailing shape as f(S, H))
PER_RAY = {
    'rgb': lambda S, H: (3,), 'depth': lambda S, H: (), 'med_depth': lambda S, H: (),
    'acc': lambda S, H: (), 'ray_norm': lambda S, H: (3,), 'ray_rotation_field': lambda S, H: (3,),
    'ray_translation_field': lambda S, H: (3,), 'ray_delta_x': lambda S, H: (3,),
    'ray_hyper_points': lambda S, H: (H,), 'ray_predicted_mask': lambda S, H: (1,),
    'med_points': lambda S, H: (1, 3 + H),
}
PER_SAMPLE = {
    'z_vals': lambda S, H: (S,), 'weights': lambda S, H: (S,), 'alpha': lambda S, H: (S,),
    'accum_prod': lambda S, H: (S,), 'sigma': lambda S, H: (S,), 'sharp_weights': lambda S, H: (S,),
    'back_facing': lambda S, H: (S,), 'predicted_mask': lambda S, H: (S, 1), 'points': lambda S, H: (S, 3),
    'warped_points': lambda S, H: (S, 3 + H), 'delta_x': lambda S, H: (S, 3),
    'predicted_norm': lambda S, H: (S, 3), 'target_norm': lambda S, H: (S, 3),
}
ALL_SHAPES = {**PER_RAY, **PER_SAMPLE}

# what render.py keeps (render.py:192-193) plus the scalars render_image users read
RENDER_KEYS = ('rgb', 'depth', 'med_depth', 'acc', 'ray_norm', 'ray_delta_x', 'med_points',
               'ray_predicted_mask', 'ray_rotation_field')


class NdsrError(RuntimeError):
  pass


def _as_dev(x, device, dtype=torch.float32):
  if x is None:
    return None
  if torch.is_tensor(x):
    t = x
  else:
    a = np.asarray(x)
    if dtype == torch.int32:   # uint32 ids: reinterpret bits (torch has no uint32 arithmetic)
      a = a.astype(np.uint32).view(np.int32)
    t = torch.from_numpy(np.ascontiguousarray(a))
  return t.to(device=device, dtype=dtype, non_blocking=True).contiguous()


class Renderer:
  """One ndsr_handle bound to one CUDA device."""

  def __init__(self, cfg: NerfDSConfig, device=None, engine: str = 'auto', precision: str = 'split3'):
    self.lib = _lib.load_library()          # raises if the .so is missing
    if not torch.cuda.is_available():
      raise NdsrError('nerfds_b200 needs a CUDA device (no CPU fallback)')
    self.cfg = cfg
    self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
    self._ccfg = _lib.to_c_config(cfg, engine, precision)
    h = C.c_void_p()
    rc = self.lib.ndsr_create(C.byref(self._ccfg), self.device.index or 0, C.byref(h))
    if rc != 0:
      raise NdsrError(f'ndsr_create failed ({rc}): {self.lib.ndsr_last_error(None).decode()}')
    self._h = h
    self._params_ref = None        # strong reference to the params object whose values are on the device
    self.H = cfg.hyper_num_dims if (cfg.has_hyper_sheet and cfg.use_hyper_for_sigma) else 0

  # ------------------------------------------------------------------ misc
  def close(self):
    if getattr(self, '_h', None):
      self.lib.ndsr_destroy(self._h)
      self._h = None

  def __del__(self):
    try:
      self.close()
    except Exception:
      pass

  def _check(self, rc, what):
    if rc != 0:
      raise NdsrError(f'{what} failed ({rc}): {self.lib.ndsr_last_error(self._h).decode()}')

  @property
  def engine(self) -> str:
    return _lib.ENGINE_NAMES[self.lib.ndsr_engine_in_use(self._h)]

  @property
  def kernel_launches(self) -> int:
    return int(self.lib.ndsr_kernel_launches(self._h))

  def tc_issued_macs(self):
    """{(level, mode): tensor-core MACs issued per sample evaluation}; mode in 'sigma', 'full', 'carried'."""
    out = (C.c_double * 8)()
    self._check(self.lib.ndsr_tc_issued_macs(self._h, out), 'ndsr_tc_issued_macs')
    return {(lv, m): float(out[lv * 4 + i]) for lv in (0, 1) for i, m in enumerate(('sigma', 'full', 'carried', 'grad'))}

  def set_max_chunk(self, rays: int):
    self._check(self.lib.ndsr_set_max_chunk(self._h, int(rays)), 'ndsr_set_max_chunk')

  def set_early_termination(self, transmittance_eps: float, rounds: int = 4):
    """Fine level of `render_rays*` calls that ask for per-ray keys only: evaluate the newly drawn depths front to
    back in `rounds` rounds and skip those behind which at most `transmittance_eps` of the ray's weight can lie
    (0 = evaluate everything, like the reference)."""
    self._check(self.lib.ndsr_set_early_termination(self._h, float(transmittance_eps), int(rounds)),
                'ndsr_set_early_termination')

  def termination_stats(self, reset: bool = False):
    """(new fine-level depths evaluated, new fine-level depths seen) since the last reset (synchronises)."""
    ev, seen = C.c_int64(0), C.c_int64(0)
    self._check(self.lib.ndsr_termination_stats(self._h, self._stream(), C.byref(ev), C.byref(seen), int(bool(reset))),
                'ndsr_termination_stats')
    return int(ev.value), int(seen.value)

  STAGES = ('sample', 'field_coarse', 'field_fine', 'composite', 'resample', 'other')

  def profile_enable(self, on: bool = True):
    self._check(self.lib.ndsr_profile_enable(self._h, int(bool(on))), 'ndsr_profile_enable')

  def profile_read(self):
    """{stage: (milliseconds, launches)} accumulated since profile_enable (synchronises)."""
    ms = (C.c_double * len(self.STAGES))()
    n = (C.c_int64 * len(self.STAGES))()
    self._check(self.lib.ndsr_profile_read(self._h, ms, n), 'ndsr_profile_read')
    return {s: (float(ms[i]), int(n[i])) for i, s in enumerate(self.STAGES)}

  def _stream(self):
    return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

  # ---------------------------------------------------------------- params
  def load_params(self, params: Dict) -> None:
    flat = [(n, np.ascontiguousarray(np.asarray(v, dtype=np.float32))) for n, v in flatten_params(params)]
    arr = (_lib.ndsr_tensor * len(flat))()
    keep = []
    for i, (name, v) in enumerate(flat):
      v2 = v.reshape(1, -1) if v.ndim == 1 else v
      b = name.encode()
      keep.append((b, v2))
      arr[i].name = b
      arr[i].data = v2.ctypes.data
      arr[i].rows, arr[i].cols = v2.shape
    with torch.cuda.device(self.device):
      self._check(self.lib.ndsr_load_params(self._h, arr, len(flat)), 'ndsr_load_params')
    self._params_ref = params

  def ensure_params(self, params: Dict) -> None:
    """Upload `params` unless this very object is the one already on the device.  Identity, not id(): the renderer
    keeps the object alive, so a new dict can never be mistaken for it.  Arrays mutated IN PLACE inside the same
    dict are not detected -- call load_params() after such an update."""
    if self._params_ref is not params:
      self.load_params(params)

  # --------------------------------------------------------------- helpers
  def make_extra(self, extra_params: Optional[Dict] = None, *, mask_ratio=1.0, sharp_weights_std=1.0,
                 use_predicted_norm=False, use_sigma_gradient=False, near=None, far=None,
                 use_sample_at_infinity=None, render_opts: Optional[Dict] = None) -> _lib.ndsr_extra_params:
    ep = _lib.ndsr_extra_params()
    x = extra_params or {}

    def f(key, default):
      v = x.get(key, default)
      if v is None:
        v = default
      if torch.is_tensor(v):
        v = v.detach().reshape(-1)[0].item()
      return float(np.asarray(v).reshape(-1)[0])

    c = self.cfg
    ep.nerf_alpha = f('nerf_alpha', c.spatial_point_max_deg)
    ep.warp_alpha = f('warp_alpha', c.warp_max_deg)
    ep.hyper_alpha = f('hyper_alpha', c.hyper_point_max_deg)
    ep.hyper_sheet_alpha = f('hyper_sheet_alpha', c.hyper_sheet_max_deg)
    ep.norm_input_alpha = f('norm_input_alpha', c.norm_input_max_deg)
    ep.mask_ratio = float(mask_ratio)
    ep.sharp_weights_std = float(sharp_weights_std)
    ep.near_override = math.nan if near is None else float(near)
    ep.far_override = math.nan if far is None else float(far)
    ep.use_predicted_norm = int(bool(use_predicted_norm))
    ep.use_sigma_gradient = int(bool(use_sigma_gradient))
    ep.sample_at_infinity_override = -1 if use_sample_at_infinity is None else int(bool(use_sample_at_infinity))
    ep.filter_flags = 0
    if render_opts is not None:            # filter_sigma (models.py:52-63): keys other than these two are ignored
      if 'dust_threshold' in render_opts:
        ep.filter_flags |= 1
        ep.dust_threshold = float(render_opts.get('dust_threshold', 0.0))
      if 'bounding_box' in render_opts:
        box = [float(v) for v in render_opts['bounding_box']]
        if len(box) != 6:
          raise ValueError('bounding_box = (xmin, xmax, ymin, ymax, zmin, zmax)')
        ep.filter_flags |= 2
        for i, v in enumerate(box):
          ep.bounding_box[i] = v
    return ep

  def alloc_outputs(self, B: int, S: int, keys: Iterable[str]):
    """Allocate a level dict of torch tensors and the matching ndsr_outputs."""
    out = _lib.ndsr_outputs()
    tensors = {}
    for k in keys:
      if k not in ALL_SHAPES:
        raise KeyError(f'unknown output key {k!r}')
      shape = (B,) + tuple(ALL_SHAPES[k](S, self.H))
      t = torch.empty(shape, dtype=torch.float32, device=self.device)
      tensors[k] = t
      setattr(out, k, t.data_ptr() if t.numel() else None)
    return tensors, out

  def level_keys(self, *, return_points, return_weights, want_target_norm=True) -> list:
    """Keys the reference's level dict holds under this config (App. B)."""
    c = self.cfg
    keys = ['rgb', 'depth', 'med_depth', 'acc', 'alpha', 'accum_prod', 'sigma', 'delta_x', 'ray_delta_x',
            'ray_hyper_points', 'med_points', 'ray_norm']
    if return_weights:
      keys.append('weights')
    if return_points:
      keys += ['points', 'warped_points']
    if c.use_predicted_mask:
      keys += ['predicted_mask', 'ray_predicted_mask']
    if c.use_mask_sharp_weights:
      keys.append('sharp_weights')
    if c.predict_norm:
      keys += ['predicted_norm', 'back_facing']
      if want_target_norm:
        keys.append('target_norm')
    if c.use_warp:
      keys += ['ray_rotation_field', 'ray_translation_field']
    return keys

  # ----------------------------------------------------------------- calls
  def render_rays(self, origins, directions, *, viewdirs=None, warp_id=None, gt_mask=None, t_rand=None, u=None,
                  extra: _lib.ndsr_extra_params, coarse_keys=(), fine_keys=RENDER_KEYS,
                  fine_ptrs: Optional[Dict[str, int]] = None):
    """NerfModel.__call__ (models.py:1419-1565) on device tensors.

    fine_ptrs: {key: device address} -- write the fine level's results there instead of into fresh tensors
    (peer.PeerFrames.shard_ptrs: this rank's rows of a frame buffer); the returned fine dict is then empty."""
    dev = self.device
    o = _as_dev(origins, dev).reshape(-1, 3)
    d = _as_dev(directions, dev).reshape(-1, 3)
    B = o.shape[0]
    v = None if viewdirs is None else _as_dev(viewdirs, dev).reshape(-1, 3)
    w = None if warp_id is None else _as_dev(warp_id, dev, torch.int32).reshape(-1)
    m = None if gt_mask is None else _as_dev(gt_mask, dev).reshape(-1)
    c = self.cfg
    tr = None if t_rand is None else _as_dev(t_rand, dev).reshape(B, -1 if B else c.num_coarse_samples)
    uu = None if u is None else _as_dev(u, dev).reshape(B, -1 if B else c.num_fine_samples)
    if tr is not None and tr.shape[1] != c.num_coarse_samples:
      raise ValueError('t_rand must be [B, num_coarse_samples]')
    if uu is not None and uu.shape[1] != c.num_fine_samples:
      raise ValueError('u must be [B, num_fine_samples]')
    Sc, Sf = c.num_coarse_samples, c.num_coarse_samples + c.num_fine_samples
    ct, co = self.alloc_outputs(B, Sc, coarse_keys)
    if fine_ptrs is None:
      ft, fo = self.alloc_outputs(B, Sf, fine_keys)
    else:
      ft, fo = {}, _lib.ndsr_outputs()
      for k in fine_keys:
        setattr(fo, k, int(fine_ptrs[k]))
    ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    with torch.cuda.device(dev):
      rc = self.lib.ndsr_render_rays(self._h, self._stream(), B, ptr(o), ptr(d), ptr(v), ptr(w), ptr(m), ptr(tr),
                                     ptr(uu), C.byref(extra), C.byref(co) if coarse_keys else None,
                                     C.byref(fo) if fine_keys else None)
    self._check(rc, 'ndsr_render_rays')
    return {'coarse': ct, 'fine': ft}

  def render_rays_host(self, origins, directions, *, viewdirs=None, warp_id=None, gt_mask=None, t_rand=None,
                       u=None, extra: _lib.ndsr_extra_params, fine_keys=RENDER_KEYS, out: Optional[Dict] = None,
                       rng_keys=None):
    """Host-buffer entry point: numpy (ideally pinned) in, numpy out; H2D/D2H inside the call.

    rng_keys: (key_coarse, key_fine), two raw uint32[2] jax keys -- the draws are then generated on the device
    (t_rand = uniform(key_coarse, [B, S_c]), u = uniform(key_fine, [B, S_f])) instead of being uploaded."""
    c = self.cfg
    f32 = lambda a: None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    o, d, v = f32(origins).reshape(-1, 3), f32(directions).reshape(-1, 3), f32(viewdirs)
    B = o.shape[0]
    w = None if warp_id is None else np.ascontiguousarray(np.asarray(warp_id).astype(np.uint32).reshape(-1))
    m, tr, uu = f32(gt_mask), f32(t_rand), f32(u)
    Sf = c.num_coarse_samples + c.num_fine_samples
    fo = _lib.ndsr_outputs()
    res = {} if out is None else out
    for k in fine_keys:
      shape = (B,) + tuple(ALL_SHAPES[k](Sf, self.H))
      if k not in res:
        res[k] = np.empty(shape, np.float32)
      setattr(fo, k, res[k].ctypes.data if res[k].size else None)
    ptr = lambda a: None if a is None else C.c_void_p(a.ctypes.data)
    with torch.cuda.device(self.device):
      if rng_keys is not None:
        kc = (C.c_uint32 * 2)(*[int(x) for x in np.asarray(rng_keys[0]).reshape(2)])
        kf = (C.c_uint32 * 2)(*[int(x) for x in np.asarray(rng_keys[1]).reshape(2)])
        rc = self.lib.ndsr_render_rays_host_rng(self._h, self._stream(), B, ptr(o), ptr(d), ptr(v), ptr(w), ptr(m),
                                                kc, kf, C.byref(extra), None, C.byref(fo))
      else:
        rc = self.lib.ndsr_render_rays_host(self._h, self._stream(), B, ptr(o), ptr(d), ptr(v), ptr(w), ptr(m),
                                            ptr(tr), ptr(uu), C.byref(extra), None, C.byref(fo))
    self._check(rc, 'ndsr_render_rays_host')
    return res

  def random_uniform(self, key, n_rays: int, n_samples: int, first_ray: int = 0, rays: Optional[int] = None):
    """Rows [first_ray, first_ray + rays) of jax.random.uniform(key, [n_rays, n_samples]) on this device."""
    rays = n_rays - first_ray if rays is None else rays
    t = torch.empty((rays, n_samples), dtype=torch.float32, device=self.device)
    k = (C.c_uint32 * 2)(*[int(x) for x in np.asarray(key).reshape(2)])
    with torch.cuda.device(self.device):
      rc = self.lib.ndsr_random_uniform_range(self.device.index or 0, self._stream(), k, n_rays * n_samples,
                                              first_ray * n_samples, rays * n_samples, C.c_void_p(t.data_ptr()))
    if rc != 0:
      raise NdsrError(f'ndsr_random_uniform_range failed ({rc})')
    return t

  def render_samples(self, level: int, z_vals, directions, *, points=None, origins=None, viewdirs=None,
                     warp_id=None, gt_mask=None, extra: _lib.ndsr_extra_params, use_sample_at_infinity=False,
                     keys=RENDER_KEYS):
    """NerfModel.render_samples (models.py:867-1417) on caller-provided samples."""
    dev = self.device
    z = _as_dev(z_vals, dev)
    B, S = z.shape
    d = _as_dev(directions, dev).reshape(B, 3)
    p = None if points is None else _as_dev(points, dev).reshape(B, S, 3)
    o = None if origins is None else _as_dev(origins, dev).reshape(B, 3)
    v = None if viewdirs is None else _as_dev(viewdirs, dev).reshape(B, 3)
    w = None if warp_id is None else _as_dev(warp_id, dev, torch.int32).reshape(-1)
    m = None if gt_mask is None else _as_dev(gt_mask, dev).reshape(-1)
    t, out = self.alloc_outputs(B, S, keys)
    ptr = lambda x: None if x is None else C.c_void_p(x.data_ptr())
    with torch.cuda.device(dev):
      rc = self.lib.ndsr_render_samples(self._h, self._stream(), int(level), B, S, ptr(p), ptr(z), ptr(o), ptr(d),
                                        ptr(v), ptr(w), ptr(m), C.byref(extra), int(bool(use_sample_at_infinity)),
                                        C.byref(out))
    self._check(rc, 'ndsr_render_samples')
    return t

  def sample_along_rays(self, n_rays, n_samples, near, far, t_rand=None, use_linear_disparity=False):
    z = torch.empty((n_rays, n_samples), dtype=torch.float32, device=self.device)
    tr = None if t_rand is None else _as_dev(t_rand, self.device).reshape(n_rays, n_samples)
    with torch.cuda.device(self.device):
      rc = self.lib.ndsr_sample_along_rays(self._h, self._stream(), n_rays, n_samples, float(near), float(far),
                                           int(bool(use_linear_disparity)),
                                           None if tr is None else C.c_void_p(tr.data_ptr()),
                                           C.c_void_p(z.data_ptr()))
    self._check(rc, 'ndsr_sample_along_rays')
    return z

  def sample_pdf(self, bins, weights, u, z_vals, diagnostics=False):
    dev = self.device
    b, w, z = _as_dev(bins, dev), _as_dev(weights, dev), _as_dev(z_vals, dev)
    uu = None if u is None else _as_dev(u, dev)
    B, nb = b.shape
    nc = z.shape[1]
    nf = uu.shape[1] if uu is not None else self.cfg.num_fine_samples
    if w.shape != (B, nb - 1):
      raise ValueError('weights must be [B, n_bins - 1]')
    z_out = torch.empty((B, nc + nf), dtype=torch.float32, device=dev)
    zs = torch.empty((B, nf), dtype=torch.float32, device=dev) if diagnostics else None
    lo = torch.empty((B, nf), dtype=torch.int32, device=dev) if diagnostics else None
    hi = torch.empty((B, nf), dtype=torch.int32, device=dev) if diagnostics else None
    cdf = torch.empty((B, nb), dtype=torch.float32, device=dev) if diagnostics else None
    ptr = lambda x: None if x is None else C.c_void_p(x.data_ptr())
    with torch.cuda.device(dev):
      rc = self.lib.ndsr_sample_pdf(self._h, self._stream(), B, nb, nf, nc, ptr(b), ptr(w), ptr(uu), ptr(z),
                                    ptr(z_out), ptr(zs), ptr(lo), ptr(hi), ptr(cdf))
    self._check(rc, 'ndsr_sample_pdf')
    if diagnostics:
      return z_out, zs, lo, hi, cdf
    return z_out

  def volumetric_rendering(self, rgb, sigma, z_vals, dirs, use_white_background=False, sample_at_infinity=True):
    dev = self.device
    r, s, z, d = _as_dev(rgb, dev), _as_dev(sigma, dev), _as_dev(z_vals, dev), _as_dev(dirs, dev)
    B, S = z.shape
    t, out = self.alloc_outputs(B, S, ('rgb', 'depth', 'med_depth', 'acc', 'weights', 'alpha', 'accum_prod'))
    with torch.cuda.device(dev):
      rc = self.lib.ndsr_volumetric_rendering(self._h, self._stream(), B, S, C.c_void_p(r.data_ptr()),
                                              C.c_void_p(s.data_ptr()), C.c_void_p(z.data_ptr()),
                                              C.c_void_p(d.data_ptr()), int(bool(use_white_background)),
                                              int(bool(sample_at_infinity)), C.byref(out))
    self._check(rc, 'ndsr_volumetric_rendering')
    return t
