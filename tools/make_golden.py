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

from typing import Any, Optional

import dataclasses
import importlib.util
import os
import sys
import types

import numpy as np

REF = os.environ.get('NERFDS_REFERENCE', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'reference_shim.npz')


# ------------------------------------------------------------------ jax / flax stand-ins
class _Jnp(types.ModuleType):
  """numpy with jax.numpy's float32 / int32 creation defaults."""

  def __getattr__(self, name):
    return getattr(np, name)


_WIDE = [False]   # True while a finite-difference probe runs: keep float64 instead of jax's float32 defaults


def _f32_default(fn):
  def wrapped(*a, dtype=None, **k):
    out = fn(*a, **k) if dtype is None else fn(*a, dtype=dtype, **k)
    if dtype is None and not _WIDE[0] and np.issubdtype(np.asarray(out).dtype, np.floating):
      out = np.asarray(out, np.float32)
    return out
  return wrapped


def _array(x, dtype=None):
  a = np.asarray(x, dtype=dtype)
  if dtype is None and a.dtype == np.float64 and not _WIDE[0]:
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


def _tree_stack(outs):
  o0 = outs[0]
  if isinstance(o0, tuple):
    return tuple(_tree_stack([o[i] for o in outs]) for i in range(len(o0)))
  if isinstance(o0, list):
    return [_tree_stack([o[i] for o in outs]) for i in range(len(o0))]
  if isinstance(o0, dict):
    return {k: _tree_stack([o[k] for o in outs]) for k in o0}
  if o0 is None:
    return None
  return np.stack([np.asarray(o) for o in outs], 0)


def _tree_take(a, i, ax):
  if ax is None or a is None:
    return a
  if isinstance(a, dict):
    return {k: _tree_take(v, i, ax) for k, v in a.items()}
  if isinstance(a, (tuple, list)):
    return type(a)(_tree_take(v, i, ax) for v in a)
  return np.take(a, i, axis=ax)


def _tree_len(a, ax):
  if isinstance(a, dict):
    return _tree_len(next(iter(a.values())), ax)
  if isinstance(a, (tuple, list)):
    return _tree_len(a[0], ax)
  return np.asarray(a).shape[ax]


def _vmap(fn, in_axes=0, out_axes=0):
  """jax.vmap as a Python loop over pytrees (the reference vmaps per-point closures over batch and samples)."""
  def mapped(*args):
    axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
    n = next(_tree_len(a, ax) for a, ax in zip(args, axes) if ax is not None and a is not None)
    return _tree_stack([fn(*[_tree_take(a, i, ax) for a, ax in zip(args, axes)]) for i in range(n)])
  return mapped


def _value_and_grad(f, argnums=0, has_aux=False):
  """No autodiff here: the VALUE is the reference's float32 evaluation; the GRADIENT is a central difference of the
  same reference function re-evaluated in float64 (`_WIDE`: float32 parameters, double arithmetic, h = 1e-6), i.e.
  the derivative of the reference's function to ~1e-8 relative -- what jax's autodiff returns up to float32
  rounding.  The reference only differentiates per-point scalar functions of a 3-vector (models.py:1063-1069)."""
  def run(*a):
    val = f(*a)
    x = np.asarray(a[argnums], np.float64)
    g = np.zeros(x.shape, np.float64)
    h = 1e-6
    _WIDE[0] = True
    try:
      for i in range(x.size):
        probe = []
        for sgn in (+1.0, -1.0):
          xi = x.copy().reshape(-1)
          xi[i] += sgn * h
          out = f(*(tuple(a[:argnums]) + (xi.reshape(x.shape),) + tuple(a[argnums + 1:])))
          out = out[0] if has_aux else out
          assert np.asarray(out).dtype == np.float64, 'the finite-difference probe fell back to float32'
          probe.append(float(np.asarray(out)))
        g.reshape(-1)[i] = (probe[0] - probe[1]) / (2 * h)
    finally:
      _WIDE[0] = False
    return val, g.astype(np.float32)
  return run


def _grad(f, argnums=0, has_aux=False):
  vg = _value_and_grad(f, argnums, has_aux)
  return lambda *a: vg(*a)[1]


# ------------------------------------------------------------------ flax.linen stand-in
# Just enough of flax.linen's module system to run the reference's modules.py / warping.py on a given parameter
# pytree: dataclass-style attributes, `setup()` sub-modules named after their attribute (dict attributes ->
# `attr_key`), `@nn.compact` sub-modules named explicitly or `ClassName_i`, parameters looked up by that path.
class _Ctx:
  stack = []


def _compact(fn):
  fn._compact = True
  return fn


class Module:
  name = None
  parent = None
  _manual = False

  def __init_subclass__(cls, **kw):
    super().__init_subclass__(**kw)
    fields = cls.__dict__.get('__annotations__', {})
    for attr, fn in list(cls.__dict__.items()):
      if attr not in fields and callable(fn) and not isinstance(fn, (staticmethod, classmethod, type)) and attr != 'setup' and \
         (attr == '__call__' or not attr.startswith('_')) and hasattr(fn, '__code__'):
        setattr(cls, attr, Module._wrap(fn))
    if not cls.__dict__.get('_manual', False):
      ann = dict(cls.__dict__.get('__annotations__', {}))
      ann['name'] = Optional[str]
      ann['parent'] = Any
      cls.__annotations__ = ann
      cls.name = None
      cls.parent = None
      dataclasses.dataclass(cls, kw_only=True, eq=False, repr=False)

  @staticmethod
  def _wrap(fn):
    def wrapped(self, *a, **k):
      self._ensure_setup()
      mode = 'compact' if getattr(fn, '_compact', False) else 'method'
      if mode == 'compact':
        object.__setattr__(self, '_auto', {})
      _Ctx.stack.append((self, mode))
      try:
        return fn(self, *a, **k)
      finally:
        _Ctx.stack.pop()
    wrapped.__name__ = fn.__name__
    return wrapped

  def __post_init__(self):
    object.__setattr__(self, '_scope', None)
    object.__setattr__(self, '_setup_done', False)
    object.__setattr__(self, '_auto', {})
    if _Ctx.stack and _Ctx.stack[-1][1] == 'compact':
      par = _Ctx.stack[-1][0]
      if self.name is None:
        i = par._auto.get(type(self).__name__, 0)
        par._auto[type(self).__name__] = i + 1
        object.__setattr__(self, 'name', f'{type(self).__name__}_{i}')
      object.__setattr__(self, 'parent', par)

  def __setattr__(self, k, v):
    if _Ctx.stack and _Ctx.stack[-1] == (self, 'setup'):
      if isinstance(v, Module):
        if v.name is None:
          object.__setattr__(v, 'name', k)
        object.__setattr__(v, 'parent', self)
      elif isinstance(v, dict) and v and all(isinstance(m, Module) for m in v.values()):
        for kk, m in v.items():
          object.__setattr__(m, 'name', f'{k}_{kk}')
          object.__setattr__(m, 'parent', self)
    object.__setattr__(self, k, v)

  def _ensure_setup(self):
    if not self._setup_done:
      object.__setattr__(self, '_setup_done', True)
      if hasattr(self, 'setup'):
        _Ctx.stack.append((self, 'setup'))
        try:
          self.setup()
        finally:
          _Ctx.stack.pop()

  def make_rng(self, name):
    return None

  def _params(self):
    if self._scope is not None:
      return self._scope
    return self.parent._params()[self.name]

  def apply(self, variables, *a, method=None, **k):
    object.__setattr__(self, '_scope', variables['params'])
    fn = method if method is not None else type(self).__call__
    return fn(self, *a, **k)


class Dense(Module):
  _manual = True

  def __init__(self, features, use_bias=True, kernel_init=None, bias_init=None, name=None, **_):
    self.features, self.use_bias = features, use_bias
    object.__setattr__(self, 'name', name)
    object.__setattr__(self, 'parent', None)
    self.__post_init__()
    if self.parent is None and _Ctx.stack and _Ctx.stack[-1][1] == 'setup':
      object.__setattr__(self, 'parent', _Ctx.stack[-1][0])      # functools.partial(nn.Dense)(..., name=...) in setup()

  def __call__(self, x):
    p = self._params()
    assert p['kernel'].shape == (x.shape[-1], self.features), (p['kernel'].shape, x.shape, self.features)
    y = np.matmul(x, p['kernel'])
    return y + p['bias'] if self.use_bias else y


class Embed(Module):
  _manual = True

  def __init__(self, num_embeddings, features, embedding_init=None, name=None, **_):
    self.num_embeddings, self.features = num_embeddings, features
    object.__setattr__(self, 'name', name)
    object.__setattr__(self, 'parent', None)
    self.__post_init__()

  def __call__(self, ids):
    e = self._params()['embedding']
    assert e.shape == (self.num_embeddings, self.features)
    return e[np.asarray(ids)]


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

  def _norm_pdf(x, loc=0, scale=1):
    # jax.scipy.stats.norm.pdf (jax 0.3.15) = exp(logpdf), evaluated in float32
    x, loc, scale = (np.asarray(v, np.float32) for v in (x, loc, scale))
    scale_sqrd = np.square(scale)
    log_normalizer = np.log(np.float32(2 * np.pi) * scale_sqrd)
    quadratic = np.square(x - loc) / scale_sqrd
    return np.exp((log_normalizer + quadratic) / np.float32(-2)).astype(np.float32)
  jscipy.stats = types.SimpleNamespace(norm=types.SimpleNamespace(pdf=_norm_pdf))
  jax.numpy, jax.lax, jax.random, jax.scipy = jnp, lax, random, jscipy
  jax.vmap, jax.value_and_grad, jax.grad = _vmap, _value_and_grad, _grad
  jax.jit = lambda f=None, **k: f if f is not None else (lambda g: g)

  def _custom_jvp(f, **k):
    f.defjvp = lambda g: g
    return f
  jax.custom_jvp = _custom_jvp
  jax.custom_vjp = _custom_jvp
  _matmul = np.matmul
  jnp.matmul = lambda a, b, precision=None: _matmul(a, b)
  flax = types.ModuleType('flax')
  linen = types.ModuleType('flax.linen')
  linen.vmap = lambda fn, **k: fn
  linen.Module, linen.Dense, linen.Embed, linen.compact = Module, Dense, Embed, _compact
  linen.relu = lambda x: np.maximum(x, 0)
  linen.sigmoid = lambda x: (1 / (1 + np.exp(-x))).astype(np.float32)
  linen.softplus = lambda x: np.logaddexp(x, 0).astype(np.float32)
  _init = lambda *a, **k: (lambda *aa, **kk: None)
  linen.initializers = types.SimpleNamespace(uniform=_init, normal=_init, zeros=None, ones=None)
  for _n in ('LayerNorm', 'GroupNorm', 'BatchNorm'):
    setattr(linen, _n, None)
  jnn = types.ModuleType('jax.nn')
  jnn.initializers = types.SimpleNamespace(glorot_uniform=_init, xavier_uniform=_init, uniform=_init, normal=_init,
                                           zeros=None, ones=None)
  jnn.relu, jnn.sigmoid, jnn.softplus = linen.relu, linen.sigmoid, linen.softplus
  jax.nn = jnn
  tree_util = types.ModuleType('jax.tree_util')
  tree_util.tree_map = lambda f, t: {k: tree_util.tree_map(f, v) for k, v in t.items()} if isinstance(t, dict) else f(t)
  jax.tree_util = tree_util
  jax.tree_map = tree_util.tree_map
  gin = types.ModuleType('gin')
  gin.REQUIRED = None
  gin.configurable = lambda *a, **k: (a[0] if a and isinstance(a[0], type) else (lambda c: c))
  gin.constant = lambda *a, **k: None
  sys.modules['gin'] = gin
  sys.modules['jax.nn'] = jnn
  sys.modules['jax.tree_util'] = tree_util
  struct = types.ModuleType('flax.struct')
  struct.dataclass = dataclasses.dataclass
  struct.field = dataclasses.field
  optim = types.ModuleType('flax.optim')
  optim.Optimizer = object
  flax.linen, flax.struct, flax.optim = linen, struct, optim
  flax.jax_utils = types.ModuleType('flax.jax_utils')
  sys.modules['flax.jax_utils'] = flax.jax_utils
  imm = types.ModuleType('immutabledict')
  imm.immutabledict = dict
  sys.modules['immutabledict'] = imm
  for name, mod in (('jax', jax), ('jax.numpy', jnp), ('jax.lax', lax), ('jax.random', random), ('jax.scipy', jscipy),
                    ('flax', flax), ('flax.linen', linen), ('flax.struct', struct), ('flax.optim', optim)):
    sys.modules[name] = mod


def load_reference(name):
  path = os.path.join(REF, 'hypernerf', name + '.py')
  spec = importlib.util.spec_from_file_location('reference_' + name, path)
  mod = importlib.util.module_from_spec(spec)
  spec.loader.exec_module(mod)
  return mod


def _immutable_args(fn):
  """jax arrays are immutable: `weights += eps` inside piecewise_constant_pdf rebinds a local.  Under numpy it
  would add eps in place to the caller's array (a view of the coarse level's `weights`), so the function gets copies."""
  def wrapped(*a, **k):
    return fn(*[np.array(x) if isinstance(x, np.ndarray) else x for x in a],
              **{kk: (np.array(v) if isinstance(v, np.ndarray) else v) for kk, v in k.items()})
  return wrapped


def main():
  install_shim()
  sys.path.insert(0, REF)
  import importlib
  mu = importlib.import_module('hypernerf.model_utils')
  mu.piecewise_constant_pdf = _immutable_args(mu.piecewise_constant_pdf)
  rb = importlib.import_module('hypernerf.rigid_body')
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

  # ------------------------------------------------------------------ Flax-module level (modules.py, warping.py)
  # The reference's own module classes, run through the flax.linen stand-in above on a reduced nerf_ds.gin
  # configuration (same depths / skips / posenc degrees / embedding sizes, narrow widths) whose parameter pytree
  # comes from nerfds_b200.params.init_params -- so the Flax parameter paths the product expects are exercised too.
  sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
  modules = importlib.import_module('hypernerf.modules')
  warping = importlib.import_module('hypernerf.warping')
  from nerfds_b200.config import nerf_ds_config
  from nerfds_b200.params import flatten_params, init_params
  small = dict(nerf_trunk_width=32, nerf_rgb_branch_width=16, warp_trunk_width=16, hyper_sheet_width=16, mask_width=16,
               num_warp_embeds=5)
  cfg = nerf_ds_config(**small)
  P = init_params(cfg, 7)
  for name, arr in flatten_params(P):
    G['P/' + name] = f32(arr)
  G['mod_cfg_keys'] = np.array(sorted(small.keys()))
  G['mod_cfg_vals'] = np.array([small[k] for k in sorted(small.keys())], np.int32)
  ep = {'warp_alpha': 2.6, 'hyper_sheet_alpha': 4.4, 'nerf_alpha': 5.5, 'hyper_alpha': 0.7, 'norm_input_alpha': 3.2}
  G['mod_extra'] = f32([ep['warp_alpha'], ep['hyper_sheet_alpha'], ep['nerf_alpha'], ep['hyper_alpha'], ep['norm_input_alpha']])
  N = 23
  pts = f32(rng.uniform(-1.2, 1.2, size=(N, 3)))
  ids = rng.integers(0, cfg.num_warp_embeds, size=(N, 1)).astype(np.uint32)
  maskv = f32(rng.uniform(0, 1.2, size=(N, 1)))
  G['mod_points'], G['mod_ids'], G['mod_mask'] = pts, ids, maskv
  # GLOEmbed (modules.py:316-348)
  wemb = f32(modules.GLOEmbed(num_embeddings=cfg.num_warp_embeds, num_dims=cfg.warp_embed_dims).apply({'params': P['warp_embed']}, ids))
  memb = f32(modules.GLOEmbed(num_embeddings=cfg.num_warp_embeds, num_dims=cfg.mask_embed_dims).apply({'params': P['mask_embed']}, ids))
  G['mod_warp_embed'], G['mod_mask_embed'] = wemb, memb
  # MaskMLP (modules.py:394-434; nerf_ds.gin:116-118: depth 8, relu output), called as models.py:967
  mm = modules.MaskMLP(depth=cfg.mask_depth, width=cfg.mask_width, skips=tuple(cfg.mask_skips), min_deg=cfg.mask_min_deg,
                       max_deg=cfg.mask_max_deg, output_activation=sys.modules['jax'].nn.relu)
  G['mod_mask_mlp'] = f32(mm.apply({'params': P['mask_mlp']}, pts, memb, alpha=ep['warp_alpha']))
  # HyperSheetMLP (modules.py:351-392), called as models.py:663-666 with [warp embed | mask]
  hs = modules.HyperSheetMLP(output_channels=cfg.hyper_num_dims, min_deg=cfg.hyper_sheet_min_deg, max_deg=cfg.hyper_sheet_max_deg,
                             depth=cfg.hyper_sheet_depth, width=cfg.hyper_sheet_width, skips=tuple(cfg.hyper_sheet_skips))
  hin = np.concatenate([wemb, maskv], -1)
  G['mod_hyper_sheet'] = f32(hs.apply({'params': P['hyper_sheet_mlp']}, pts, hin, alpha=ep['hyper_sheet_alpha']))
  # SE3Field (warping.py:123-281): warp of points; rotation of a vector (map_vectors), forward and inverse
  se3 = warping.SE3Field(min_deg=cfg.warp_min_deg, max_deg=cfg.warp_max_deg, use_posenc_identity=bool(cfg.warp_use_posenc_identity),
                         skips=tuple(cfg.warp_skips), trunk_depth=cfg.warp_trunk_depth, trunk_width=cfg.warp_trunk_width)
  # init_params draws the w / v logits small enough that theta stays away from 0
  # (SE3Field.warp works on ONE point; the reference vmaps it over batch and samples, models.py:609-630)
  def warp_all(**kw):
    outs = [se3.apply({'params': P['warp_field']}, pts[i], hin[i], ep, method=warping.SE3Field.warp,
                      **{k: (v[i] if isinstance(v, np.ndarray) else v) for k, v in kw.items()}) for i in range(N)]
    return [np.stack([o[j] for o in outs]) if outs[0][j] is not None else None for j in range(2)]
  wp, screw = warp_all(return_screw=True)
  G['mod_se3_points'], G['mod_se3_screw'] = f32(wp), f32(screw)
  vec = f32(rng.normal(size=(N, 3)))
  G['mod_vec'] = vec
  G['mod_se3_vec_fwd'] = f32(warp_all(vector=vec)[0])
  G['mod_se3_vec_inv'] = f32(warp_all(vector=vec, inverse=True)[0])
  G['mod_se3_vec_trans'] = f32(warp_all(vector=vec * 0, with_translation=True)[0])
  # NerfMLP (modules.py:86-313): trunk -> bottleneck -> sigma/normal head -> rgb branch, as models.py:525-565 calls it
  nm = modules.NerfMLP(trunk_depth=cfg.nerf_trunk_depth, trunk_width=cfg.nerf_trunk_width,
                       rgb_branch_depth=cfg.nerf_rgb_branch_depth, rgb_branch_width=cfg.nerf_rgb_branch_width,
                       skips=tuple(cfg.nerf_skips), alpha_channels=cfg.alpha_channels, rgb_channels=cfg.rgb_channels,
                       predict_norm=bool(cfg.predict_norm))
  feat = f32(rng.normal(size=(N, cfg.trunk_in_dim)))
  vfeat = f32(rng.normal(size=(N, 24)))
  nfeat = f32(rng.normal(size=(N, 24)))
  G['mod_trunk_in'], G['mod_view_feat'], G['mod_norm_feat'] = feat, vfeat, nfeat
  PN = {'params': P['nerf_mlps_fine']}
  trunk_out, bott = nm.apply(PN, feat, None, vfeat, method=modules.NerfMLP.query_bottleneck)
  alpha, nrm, _, _ = nm.apply(PN, trunk_out, bott, None, method=modules.NerfMLP.query_sigma)
  rgb_raw = nm.apply(PN, trunk_out, bott, vfeat, norm=nfeat, extra_rgb_condition=trunk_out, method=modules.NerfMLP.query_rgb)
  G['mod_trunk_out'], G['mod_bottleneck'], G['mod_alpha'], G['mod_norm'], G['mod_rgb_raw'] = (
      f32(trunk_out), f32(bott), f32(alpha), f32(nrm), f32(rgb_raw))

  # ------------------------------------------------------------------ NerfModel.__call__ (models.py:1419-1565)
  # The reference's whole forward -- sample_along_rays, render_samples('coarse'), sample_pdf, render_samples('fine')
  # -- with the gin bindings of nerf_ds.gin / defaults.gin passed as constructor arguments (reduced widths and
  # sample counts), the product's parameter pytree and injected uniform draws.  B >= S_c + S_f because the
  # row-gather of sharpen_weights (model_utils.py:182) indexes ROWS with a sample index: jax clamps
  # an out-of-range gather, numpy raises.
  import functools
  models = importlib.import_module('hypernerf.models')
  jx = sys.modules['jax']
  msmall = dict(small, num_coarse_samples=8, num_fine_samples=8)
  mcfg = nerf_ds_config(**msmall)
  MP = init_params(mcfg, 11)
  for name, arr in flatten_params(MP):
    G['MP/' + name] = f32(arr)
  G['model_cfg_keys'] = np.array(sorted(msmall.keys()))
  G['model_cfg_vals'] = np.array([msmall[k] for k in sorted(msmall.keys())], np.int32)
  MaskMLP_cls = modules.MaskMLP
  modules.MaskMLP = functools.partial(MaskMLP_cls, depth=mcfg.mask_depth, width=mcfg.mask_width,
                                      output_activation=jx.nn.relu)                       # nerf_ds.gin:116-118
  model_kw = dict(
      embeddings_dict={'warp': list(range(mcfg.num_warp_embeds)), 'appearance': [0], 'camera': [0]},
      near=mcfg.near, far=mcfg.far, num_coarse_samples=mcfg.num_coarse_samples, num_fine_samples=mcfg.num_fine_samples,
      use_viewdirs=True, use_stratified_sampling=True, norm_type='none', activation=jx.nn.relu, use_posenc_identity=False,
      spatial_point_min_deg=0, spatial_point_max_deg=8, hyper_point_min_deg=0, hyper_point_max_deg=1,
      hyper_slice_method='bendy_sheet', hyper_use_warp_embed=True,
      hyper_sheet_mlp_cls=functools.partial(modules.HyperSheetMLP, min_deg=0, max_deg=6, output_channels=2,
                                            width=mcfg.hyper_sheet_width),
      use_warp=True,
      warp_field_cls=functools.partial(warping.SE3Field, min_deg=0, max_deg=4, use_posenc_identity=False,
                                       trunk_width=mcfg.warp_trunk_width),
      warp_embed_cls=functools.partial(modules.GLOEmbed, num_dims=8),
      hyper_embed_cls=functools.partial(modules.GLOEmbed, num_dims=2),
      use_rgb_condition=False, predict_norm=True, norm_supervision_type='warped', use_viewdirs_in_hyper=False,
      use_x_in_rgb_condition=True, use_hyper_c=False, hyper_c_hyper_input=True, use_hyper_c_embed=False,
      use_mask_in_warp=True, use_mask_in_hyper=True, use_mask_in_rgb=False, use_predicted_mask=True, use_3d_mask=True,
      use_mask_sharp_weights=True, nerf_trunk_width=mcfg.nerf_trunk_width, nerf_rgb_branch_width=mcfg.nerf_rgb_branch_width)
  model = models.NerfModel(**model_kw)
  MB = 19
  mo = f32(rng.normal(size=(MB, 3)) * 0.3)
  md = f32(rng.normal(size=(MB, 3)))
  md /= np.linalg.norm(md, axis=-1, keepdims=True)
  mmeta = rng.integers(0, mcfg.num_warp_embeds, size=(MB, 1)).astype(np.uint32)
  mgt = f32(rng.integers(0, 2, size=(MB, 1)))
  mt = f32(rng.uniform(size=(MB, mcfg.num_coarse_samples)))
  mu_ = f32(rng.uniform(size=(MB, mcfg.num_fine_samples)))
  mep = {'nerf_alpha': 5.5, 'warp_alpha': 2.6, 'hyper_alpha': 0.7, 'hyper_sheet_alpha': 4.4, 'norm_loss_weight': 1.0,
         'norm_input_alpha': 3.2, 'norm_voxel_lr': 0.0, 'norm_voxel_ratio': 0.0}
  G['model_origins'], G['model_dirs'], G['model_warp'], G['model_gt_mask'] = mo, md, mmeta, mgt
  G['model_t_rand'], G['model_u'] = mt, mu_
  G['model_extra_keys'] = np.array(sorted(mep.keys()))
  G['model_extra_vals'] = f32([mep[k] for k in sorted(mep.keys())])
  G['model_mask_ratio'], G['model_sharp_std'] = f32(0.7), f32(0.1)
  _DRAWS.extend([mt, mu_])
  res = model.apply({'params': MP}, {'origins': mo, 'directions': md, 'metadata': {'warp': mmeta}, 'mask': mgt}, mep,
                    use_predicted_norm=True, return_points=True, return_weights=True, mask_ratio=0.7, sharp_weights_std=0.1)
  for lvl in ('coarse', 'fine'):
    for k, v in res[lvl].items():
      if v is not None:
        G[f'model_{lvl}_{k}'] = f32(v)

  # Second variant: the inference settings of render.py (mask_ratio 1, end-of-schedule alphas) with the sampling /
  # compositing switches flipped -- deterministic depths (no stratified jitter, u = linspace), linear disparity,
  # white background, no sample at infinity.  Same rays and parameters.
  modelB = models.NerfModel(**dict(model_kw, use_stratified_sampling=False, use_white_background=True,
                                   use_linear_disparity=True, use_sample_at_infinity=False))
  mepB = {'nerf_alpha': 8.0, 'warp_alpha': 4.0, 'hyper_alpha': 1.0, 'hyper_sheet_alpha': 6.0, 'norm_loss_weight': 1.0,
          'norm_input_alpha': 4.0, 'norm_voxel_lr': 0.0, 'norm_voxel_ratio': 0.0}
  resB = modelB.apply({'params': MP}, {'origins': mo, 'directions': md, 'metadata': {'warp': mmeta}, 'mask': mgt}, mepB,
                      use_predicted_norm=True, return_points=True, return_weights=True, mask_ratio=1, sharp_weights_std=0.1)
  for lvl in ('coarse', 'fine'):
    for k, v in resB[lvl].items():
      if v is not None:
        G[f'modelB_{lvl}_{k}'] = f32(v)

  # Third / fourth set: the same two call variants with every network 32 wide -- the narrowest shape the CUDA
  # engines build -- so that tests/test_gpu_parity.py can compare the CUDA path with the reference's output directly.
  w32 = dict(nerf_trunk_width=32, nerf_rgb_branch_width=32, mask_width=32, warp_trunk_width=32, hyper_sheet_width=32,
             num_warp_embeds=5, num_coarse_samples=8, num_fine_samples=8)
  cfg32 = nerf_ds_config(**w32)
  MP32 = init_params(cfg32, 12)
  for name, arr in flatten_params(MP32):
    G['MP32/' + name] = f32(arr)
  G['model32_cfg_keys'] = np.array(sorted(w32.keys()))
  G['model32_cfg_vals'] = np.array([w32[k] for k in sorted(w32.keys())], np.int32)
  modules.MaskMLP = functools.partial(MaskMLP_cls, depth=cfg32.mask_depth, width=32, output_activation=jx.nn.relu)
  kw32 = dict(model_kw, nerf_trunk_width=32, nerf_rgb_branch_width=32,
              hyper_sheet_mlp_cls=functools.partial(modules.HyperSheetMLP, min_deg=0, max_deg=6, output_channels=2, width=32),
              warp_field_cls=functools.partial(warping.SE3Field, min_deg=0, max_deg=4, use_posenc_identity=False,
                                               trunk_width=32))
  _DRAWS.extend([mt, mu_])
  res32 = models.NerfModel(**kw32).apply(
      {'params': MP32}, {'origins': mo, 'directions': md, 'metadata': {'warp': mmeta}, 'mask': mgt}, mep,
      use_predicted_norm=True, return_points=True, return_weights=True, mask_ratio=0.7, sharp_weights_std=0.1)
  res32B = models.NerfModel(**dict(kw32, use_stratified_sampling=False, use_white_background=True,
                                   use_linear_disparity=True, use_sample_at_infinity=False)).apply(
      {'params': MP32}, {'origins': mo, 'directions': md, 'metadata': {'warp': mmeta}, 'mask': mgt}, mepB,
      use_predicted_norm=True, return_points=True, return_weights=True, mask_ratio=1, sharp_weights_std=0.1)
  # render_opts / filter_sigma (models.py:38-66, applied at 1236 and 1288, fine level only): dust threshold and a
  # bounding box that cuts through the sampled depth range
  ropts = {'dust_threshold': 0.35, 'bounding_box': (-0.6, 0.5, -0.7, 0.6, -0.4, 0.8)}
  G['model32R_dust'] = f32(ropts['dust_threshold'])
  G['model32R_bbox'] = f32(ropts['bounding_box'])
  _DRAWS.extend([mt, mu_])
  res32R = models.NerfModel(**kw32).apply(
      {'params': MP32}, {'origins': mo, 'directions': md, 'metadata': {'warp': mmeta}, 'mask': mgt}, mep,
      use_predicted_norm=True, return_points=True, return_weights=True, mask_ratio=0.7, sharp_weights_std=0.1,
      render_opts=ropts)
  for lvl in ('coarse', 'fine'):
    for k, v in res32R[lvl].items():
      if v is not None and k != 'target_norm':
        G[f'model32R_{lvl}_{k}'] = f32(v)
  for tag, res_ in (('model32', res32), ('model32B', res32B)):
    for lvl in ('coarse', 'fine'):
      for k, v in res_[lvl].items():
        if v is not None and k != 'sharp_weights':
          G[f'{tag}_{lvl}_{k}'] = f32(v)

  # Fifth set: nerf_ds.gin's FULL widths (the shape the tensor-core engine builds), variant-A call.  The 1.5 M
  # parameters are not stored: init_params(cfg, seed) regenerates them (numpy Generator streams are stable).
  full = dict(num_warp_embeds=5, num_coarse_samples=8, num_fine_samples=8)
  cfgF = nerf_ds_config(**full)
  MPF = init_params(cfgF, 13)
  G['modelF_seed'] = np.array(13, np.int32)
  G['modelF_param_checksum'] = np.array(sum(float(np.abs(a).sum()) for _, a in flatten_params(MPF)), np.float64)
  modules.MaskMLP = functools.partial(MaskMLP_cls, depth=cfgF.mask_depth, width=cfgF.mask_width, output_activation=jx.nn.relu)
  kwF = dict(model_kw, nerf_trunk_width=cfgF.nerf_trunk_width, nerf_rgb_branch_width=cfgF.nerf_rgb_branch_width,
             hyper_sheet_mlp_cls=functools.partial(modules.HyperSheetMLP, min_deg=0, max_deg=6, output_channels=2,
                                                   width=cfgF.hyper_sheet_width),
             warp_field_cls=functools.partial(warping.SE3Field, min_deg=0, max_deg=4, use_posenc_identity=False,
                                              trunk_width=cfgF.warp_trunk_width))
  _DRAWS.extend([mt, mu_])
  resF = models.NerfModel(**kwF).apply(
      {'params': MPF}, {'origins': mo, 'directions': md, 'metadata': {'warp': mmeta}, 'mask': mgt}, mep,
      use_predicted_norm=True, return_points=True, return_weights=True, mask_ratio=0.7, sharp_weights_std=0.1)
  for lvl in ('coarse', 'fine'):
    for k, v in resF[lvl].items():
      if v is not None and k != 'sharp_weights':
        G[f'modelF_{lvl}_{k}'] = f32(v)

  # ------------------------------------------------------------------ ray generation
  # hypernerf/camera.py is plain numpy; its module imports gpath -> tensorflow, stubbed out
  tf = types.ModuleType('tensorflow')
  tf.io = types.SimpleNamespace(gfile=types.SimpleNamespace())
  sys.modules['tensorflow'] = tf
  camera_mod = importlib.import_module('hypernerf.camera')
  th = 0.3
  Rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
  Rx = np.array([[1, 0, 0], [0, np.cos(0.5), -np.sin(0.5)], [0, np.sin(0.5), np.cos(0.5)]])
  cams = {
      'plain': dict(orientation=np.eye(3), position=[0.1, -0.2, -1.2], focal_length=20.0, principal_point=[6.5, 4.5],
                    image_size=[13, 9]),
      'radial': dict(orientation=Rz @ Rx, position=[0.4, 0.3, -0.9], focal_length=17.5, principal_point=[6.2, 4.9],
                     image_size=[13, 9], radial_distortion=[0.08, -0.02, 0.004]),
      'full': dict(orientation=Rx @ Rz, position=[-0.3, 0.2, 1.1], focal_length=15.0, principal_point=[5.7, 5.3],
                   image_size=[12, 10], skew=0.03, pixel_aspect_ratio=1.02, radial_distortion=[0.11, 0.03, -0.006],
                   tangential_distortion=[0.004, -0.003]),
  }
  G['camera_names'] = np.array(sorted(cams))
  for name, kw in cams.items():
    cam_ref = camera_mod.Camera(**{k: np.asarray(v) for k, v in kw.items()})
    for k, v in kw.items():
      G[f'camera_{name}_{k}'] = np.asarray(v, np.float64)
    rays_dir = cam_ref.pixels_to_rays(cam_ref.get_pixel_centers())          # datasets/core.py:66-71
    G[f'camera_{name}_directions'] = rays_dir.astype(np.float32)
    G[f'camera_{name}_origins'] = np.tile(cam_ref.position[None, None, :], cam_ref.image_shape + (1,)).astype(np.float32)
    G[f'camera_{name}_pixels'] = cam_ref.get_pixel_centers().astype(np.float32)

  assert not _DRAWS
  os.makedirs(os.path.dirname(OUT), exist_ok=True)
  np.savez_compressed(OUT, **G)
  print(f'wrote {OUT}: {len(G)} arrays, {os.path.getsize(OUT)} bytes')


if __name__ == '__main__':
  main()
