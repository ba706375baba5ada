"""Golden vectors from the REFERENCE'S OWN SOURCE for the array-level functions of the hot path.

The reference is a JAX/Flax program and jax / flax are not installable in this image, but its array-level functions
(hypernerf/model_utils.py, hypernerf/rigid_body.py) are written against `jax.numpy`, whose semantics for these calls
coincide with NumPy's.  This script installs a thin stand-in for the `jax` / `flax` modules (jax.numpy -> numpy with
float32 creation defaults, lax.stop_gradient -> identity, random.uniform -> injected draws, vmap -> a Python loop),
imports the UNMODIFIED reference files from /root/reference, runs them on seeded float32 inputs and stores
inputs + outputs in tests/golden/reference_shim.npz.  tests/test_oracle_golden.py then pins the oracle
(oracle/nerfds_oracle.py) against these vectors on machines where /root/reference does not exist.

Run here (the reference must be mounted):  python tools/make_golden.py
"""
from __future__ import annotations

import dataclasses
import importlib.util
import os
import sys
import types

import numpy as np
import scipy.stats

REF = os.environ.get('NERFDS_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'reference_shim.npz')


# ------------------------------------------------------------------ jax / flax stand-ins
class _Jnp(types.ModuleType):
  """numpy with jax.numpy's float32 / int32 creation defaults."""

  def __getattr__(self, name):
    return getattr(np, name)


def _f32_default(fn):
  def wrapped(*a, dtype=None, **k):
    out = fn(*a, **k) if dtype is None else fn(*a, dtype=dtype, **k)
    if dtype is None and np.issubdtype(np.asarray(out).dtype, np.floating):
      out = np.asarray(out, np.float32)
    return out
  return wrapped


def _array(x, dtype=None):
  a = np.asarray(x, dtype=dtype)
  if dtype is None and a.dtype == np.float64:
    a = a.astype(np.float32)
  if dtype is None and a.dtype == np.int64:
    a = a.astype(np.int32)
  return a


class _JaxInt(np.ndarray):
  """Integer array with jax's promotion: (weak Python float | float32) (op) int32 -> float32.  numpy would give
  float64 for both, and everything downstream (e.g. sin(x * 2^k + pi/2)) would silently run in double."""

  def __array_ufunc__(self, ufunc, method, *inputs, **kw):
    plain = [np.asarray(i) if isinstance(i, _JaxInt) else i for i in inputs]
    out = getattr(ufunc, method)(*plain, **kw)
    wide = any(isinstance(i, np.ndarray) and i.dtype == np.float64 for i in plain)
    if isinstance(out, np.ndarray) and out.dtype == np.float64 and not wide:
      out = out.astype(np.float32)
    elif isinstance(out, np.ndarray) and np.issubdtype(out.dtype, np.integer):
      out = out.view(_JaxInt)
    return out


def _arange(*a, dtype=None):
  out = np.arange(*a, dtype=dtype)
  if dtype is None:
    out = out.astype(np.int32).view(_JaxInt) if np.issubdtype(out.dtype, np.integer) else out.astype(np.float32)
  return out


jnp = _Jnp('jax.numpy')
jnp.ndarray = np.ndarray
jnp.array = _array
jnp.asarray = _array
jnp.arange = _arange
for _name in ('linspace', 'zeros', 'ones', 'eye', 'full', 'exp', 'sqrt', 'sin', 'cos'):
  setattr(jnp, _name, _f32_default(getattr(np, _name)))
jnp.float = np.float32
jnp.uint = np.uint32
jnp.linalg = np.linalg

_DRAWS = []   # injected results of jax.random.uniform / normal, consumed in call order


def _draw(key, shape, dtype=np.float32, **_):
  a = _DRAWS.pop(0)
  assert tuple(a.shape) == tuple(shape), (a.shape, shape)
  return a.astype(dtype)


def _vmap(fn, in_axes=0, out_axes=0):
  def mapped(*args):
    axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
    n = next(np.asarray(a).shape[ax] for a, ax in zip(args, axes) if ax is not None)
    outs = [fn(*[a if ax is None else np.take(a, i, axis=ax) for a, ax in zip(args, axes)]) for i in range(n)]
    if isinstance(outs[0], tuple):
      return tuple(np.stack([o[j] for o in outs], 0) for j in range(len(outs[0])))
    return np.stack(outs, 0)
  return mapped


def install_shim():
  jax = types.ModuleType('jax')
  lax = types.ModuleType('jax.lax')
  lax.stop_gradient = lambda x: x
  lax.Precision = types.SimpleNamespace(HIGHEST='highest')
  random = types.ModuleType('jax.random')
  random.uniform = _draw
  random.normal = _draw
  random.split = lambda key, num=2: [key] * num
  random.PRNGKey = lambda seed: np.zeros(2, np.uint32)
  jscipy = types.ModuleType('jax.scipy')
  jscipy.stats = types.SimpleNamespace(norm=types.SimpleNamespace(
      pdf=lambda x, loc=0, scale=1: scipy.stats.norm.pdf(x, loc, scale).astype(np.float32)))
  jax.numpy, jax.lax, jax.random, jax.scipy = jnp, lax, random, jscipy
  jax.vmap = _vmap
  jax.jit = lambda f=None, **k: f if f is not None else (lambda g: g)
  _matmul = np.matmul
  jnp.matmul = lambda a, b, precision=None: _matmul(a, b)
  flax = types.ModuleType('flax')
  linen = types.ModuleType('flax.linen')
  linen.vmap = lambda fn, **k: fn
  linen.Module = object
  struct = types.ModuleType('flax.struct')
  struct.dataclass = dataclasses.dataclass
  struct.field = dataclasses.field
  optim = types.ModuleType('flax.optim')
  optim.Optimizer = object
  flax.linen, flax.struct, flax.optim = linen, struct, optim
  for name, mod in (('jax', jax), ('jax.numpy', jnp), ('jax.lax', lax), ('jax.random', random), ('jax.scipy', jscipy),
                    ('flax', flax), ('flax.linen', linen), ('flax.struct', struct), ('flax.optim', optim)):
    sys.modules[name] = mod


def load_reference(name):
  path = os.path.join(REF, 'hypernerf', name + '.py')
  spec = importlib.util.spec_from_file_location('reference_' + name, path)
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


def main():
  install_shim()
  mu = load_reference('model_utils')
  rb = load_reference('rigid_body')
  rng = np.random.default_rng(20261017)
  f32 = lambda a: np.asarray(a, np.float32)
  G = {}

  # ---- posenc / posenc_window / normalize_vector (model_utils.py:398-442)
  x = f32(rng.uniform(-1.5, 1.5, size=(7, 3)))
  G['posenc_x'] = x
  for tag, (lo, hi, ident, alpha) in {'a': (0, 8, False, None), 'b': (0, 4, True, 2.3), 'c': (0, 6, False, 4.0),
                                      'd': (0, 1, False, 0.4)}.items():
    G[f'posenc_{tag}'] = f32(mu.posenc(x, lo, hi, ident, alpha))
    G[f'posenc_{tag}_args'] = f32([lo, hi, float(ident), np.nan if alpha is None else alpha])
  for tag, (lo, hi, alpha) in {'a': (0, 8, 5.3), 'b': (0, 4, 0.0), 'c': (2, 6, 9.0)}.items():
    G[f'window_{tag}'] = f32(mu.posenc_window(lo, hi, alpha))
    G[f'window_{tag}_args'] = f32([lo, hi, alpha])
  v = f32(rng.normal(size=(9, 3)))
  v[3] = 0
  G['normalize_in'] = v
  G['normalize_out'] = f32(mu.normalize_vector(v))

  # ---- sample_along_rays (model_utils.py:55-92)
  B, Sc, Sf = 6, 16, 12
  o = f32(rng.normal(size=(B, 3)))
  d = f32(rng.normal(size=(B, 3)))
  t_rand = f32(rng.uniform(size=(B, Sc)))
  G['sar_origins'], G['sar_dirs'], G['sar_t_rand'] = o, d, t_rand
  for tag, (strat, disp) in {'strat': (True, False), 'det': (False, False), 'disp': (True, True)}.items():
    if strat:
      _DRAWS.append(t_rand)
    z, pts = mu.sample_along_rays(None, o, d, Sc, 0.1, 2.5, strat, disp)
    G[f'sar_{tag}_z'], G[f'sar_{tag}_points'] = f32(z), f32(pts)

  # ---- volumetric_rendering / cal_weights / sharpen_weights / depth (model_utils.py:95-190, 272-317)
  z = np.sort(f32(rng.uniform(0.1, 2.5, size=(B, Sc))), -1)
  sigma = f32(np.maximum(rng.normal(size=(B, Sc)) * 8.0, 0.0))
  sigma[1] = 0
  rgb = f32(rng.uniform(size=(B, Sc, 3)))
  G['vr_z'], G['vr_sigma'], G['vr_rgb'], G['vr_dirs'] = z, sigma, rgb, d
  for tag, (white, inf) in {'inf': (False, True), 'white': (True, True), 'noinf': (False, False)}.items():
    r = mu.volumetric_rendering(rgb, sigma, z, d, white, inf)
    for k in ('rgb', 'depth', 'med_depth', 'acc', 'weights', 'alpha', 'accum_prod'):
      if k in r:
        G[f'vr_{tag}_{k}'] = f32(r[k])
  w = f32(mu.cal_weights(sigma, z, d))
  G['cal_weights'] = w
  G['sharpen_weights'] = f32(mu.sharpen_weights(w, z, std=0.1))
  G['depth_index'] = np.asarray(mu.compute_depth_index(w), np.int32)
  G['depth_map'] = f32(mu.compute_depth_map(w, z))
  G['opaqueness_mask'] = f32(mu.compute_opaqueness_mask(w))

  # ---- piecewise_constant_pdf / sample_pdf (model_utils.py:193-269): the inverse CDF and the sort
  bins = .5 * (z[..., 1:] + z[..., :-1])
  wts = w[..., 1:-1].copy()   # the reference's `weights += eps` rebinds under jax but would mutate a numpy view
  u = f32(rng.uniform(size=(B, Sf)))
  u[0, 0] = 0.0
  G['pdf_bins'], G['pdf_weights'], G['pdf_u'] = f32(bins), f32(wts).copy(), u
  _DRAWS.append(u)
  G['pdf_samples'] = f32(mu.piecewise_constant_pdf(None, bins, wts.copy(), Sf, True))
  G['pdf_samples_det'] = f32(mu.piecewise_constant_pdf(None, bins, wts.copy(), Sf, False))
  _DRAWS.append(u)
  zf, pf = mu.sample_pdf(None, bins, wts.copy(), o, d, z, Sf, True)
  G['sample_pdf_z'], G['sample_pdf_points'] = f32(zf), f32(pf)

  # ---- SE(3) exponential (rigid_body.py:27-109)
  wv = f32(rng.normal(size=(8, 3)))
  wv /= np.linalg.norm(wv, axis=-1, keepdims=True)
  vv = f32(rng.normal(size=(8, 3)))
  th = f32(rng.uniform(1e-3, 1.2, size=(8,)))
  G['se3_w'], G['se3_v'], G['se3_theta'] = wv, vv, th
  G['skew'] = f32(np.stack([rb.skew(a) for a in wv]))
  G['exp_so3'] = f32(np.stack([rb.exp_so3(a, t) for a, t in zip(wv, th)]))
  G['exp_se3'] = f32(np.stack([rb.exp_se3(np.concatenate([a, b]), t) for a, b, t in zip(wv, vv, th)]))
  p = f32(rng.normal(size=(8, 3)))
  G['hom_in'] = p
  G['to_homogenous'] = f32(rb.to_homogenous(p))
  G['from_homogenous'] = f32(rb.from_homogenous(np.concatenate([p * 2.0, np.full((8, 1), 2.0, np.float32)], -1)))

  assert not _DRAWS
  os.makedirs(os.path.dirname(OUT), exist_ok=True)
  np.savez_compressed(OUT, **G)
  print(f'wrote {OUT}: {len(G)} arrays, {os.path.getsize(OUT)} bytes')


if __name__ == '__main__':
  main()
