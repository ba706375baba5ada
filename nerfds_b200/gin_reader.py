"""Reader for the gin files that configure the path.

`render.py:50-56` restores an experiment by `gin.parse_config(exp_dir/'config.gin')`
-- the operative config `train.py:335-338` wrote -- and `train.py` itself
parses `configs/nerf_ds.gin` (which includes `configs/defaults.gin`).  gin-config
is not installed in this image, and the path only needs the *bindings* (no
configurable registry, no injection), so this is a small stand-alone reader of
the gin file syntax the shipped configs and operative dumps use:

  * `# comments`, blank lines, `import x.y` (ignored), `include 'file.gin'`
  * `scope/Class.param = <python literal>` bindings, later ones overriding
  * `NAME = <literal>` macros and `%NAME` references, resolved lazily (a macro
    redefined after its use still wins, as in gin)
  * `@scope/configurable` and `@configurable()` references, kept as the strings
    `'@scope/configurable'` / `'@configurable()'`
  * values spanning lines inside brackets or with a trailing backslash

`parse_config_file()` returns {binding name: python value}; `model_config()`
turns that into the `NerfDSConfig` the renderer takes, starting from the
reference's *class* defaults (not from nerf_ds.gin's values).
"""
from __future__ import annotations

import ast
import os
import re
from typing import Any, Dict, Iterable, List, Optional, Tuple

from . import config as _config

_BINDING = re.compile(r'^([A-Za-z_][\w./]*)\s*=\s*(.*)$', re.S)
_INCLUDE = re.compile(r'^include\s+([\'"])(.+?)\1\s*$')
_IMPORT = re.compile(r'^(import|from)\s+[\w.]+')
_OPEN, _CLOSE = '([{', ')]}'


class GinSyntaxError(ValueError):
  pass


def _scan(text: str) -> List[Tuple[int, str]]:
  """Splits gin text into (line number, statement) with comments removed.

  A statement ends at a newline that is outside brackets and strings and not
  escaped with a backslash.
  """
  out, cur, depth, quote, start, line = [], [], 0, None, 1, 1
  i, n = 0, len(text)
  while i < n:
    ch = text[i]
    if quote:
      cur.append(ch)
      if ch == '\\' and i + 1 < n:
        cur.append(text[i + 1]); i += 1
      elif text.startswith(quote, i):
        cur.extend(quote[1:]); i += len(quote) - 1; quote = None
      elif ch == '\n':
        line += 1
    elif ch in '\'"':
      quote = text[i:i + 3] if text[i:i + 3] in ('"""', "'''") else ch
      cur.extend(quote); i += len(quote) - 1
    elif ch == '#':
      while i < n and text[i] != '\n':
        i += 1
      continue
    elif ch == '\\' and i + 1 < n and text[i + 1] == '\n':
      i += 1; line += 1; cur.append(' ')
    elif ch == '\n':
      line += 1
      if depth > 0:
        cur.append(' ')
      else:
        s = ''.join(cur).strip()
        if s:
          out.append((start, s))
        cur, start = [], line
    else:
      if ch in _OPEN:
        depth += 1
      elif ch in _CLOSE:
        depth -= 1
        if depth < 0:
          raise GinSyntaxError(f'line {line}: unbalanced {ch!r}')
      cur.append(ch)
    i += 1
  if quote or depth:
    raise GinSyntaxError(f'line {start}: unterminated {"string" if quote else "bracket"}')
  s = ''.join(cur).strip()
  if s:
    out.append((start, s))
  return out


class _Macro:
  __slots__ = ('name',)

  def __init__(self, name):
    self.name = name


def _rewrite_refs(expr: str) -> str:
  """`%NAME` -> `__macro__('NAME')`, `@a/b.c()` -> `'@a/b.c()'`, outside strings."""
  out, i, n, quote = [], 0, len(expr), None
  while i < n:
    ch = expr[i]
    if quote:
      out.append(ch)
      if ch == '\\' and i + 1 < n:
        out.append(expr[i + 1]); i += 1
      elif ch == quote:
        quote = None
    elif ch in '\'"':
      quote = ch; out.append(ch)
    elif ch in '%@':
      m = re.match(r'[A-Za-z_][\w./]*(\(\))?', expr[i + 1:])
      if not m:
        raise GinSyntaxError(f'dangling {ch!r} in {expr!r}')
      name = m.group(0)
      out.append(f'__macro__({name!r})' if ch == '%' else repr('@' + name))
      i += len(name)
    else:
      out.append(ch)
    i += 1
  return ''.join(out)


def _build(node: ast.AST) -> Any:
  """Literal evaluation of a value expression; macros stay as `_Macro` placeholders."""
  if isinstance(node, ast.Expression):
    return _build(node.body)
  if isinstance(node, ast.Constant):
    return node.value
  if isinstance(node, ast.Tuple):
    return tuple(_build(e) for e in node.elts)
  if isinstance(node, ast.List):
    return [_build(e) for e in node.elts]
  if isinstance(node, ast.Set):
    return {_build(e) for e in node.elts}
  if isinstance(node, ast.Dict):
    return {_build(k): _build(v) for k, v in zip(node.keys, node.values)}
  if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
    v = _build(node.operand)
    return -v if isinstance(node.op, ast.USub) else +v
  if (isinstance(node, ast.Call) and isinstance(node.func, ast.Name) and node.func.id == '__macro__'
      and len(node.args) == 1 and isinstance(node.args[0], ast.Constant)):
    return _Macro(node.args[0].value)
  raise GinSyntaxError(f'unsupported value syntax: {ast.dump(node)[:80]}')


def _parse_value(expr: str, where: str) -> Any:
  try:
    return _build(ast.parse(_rewrite_refs(expr), mode='eval'))
  except SyntaxError as e:
    raise GinSyntaxError(f'{where}: cannot parse value {expr!r}: {e.msg}') from None


_LAZY = [True]


class _Unbound:
  """A binding whose value references a macro nobody defined (e.g. `NerfiesDataSource.data_dir = %data_dir` when the
  file is parsed without a data_dir binding).  gin raises when such a value is used, not when the file is parsed; any
  use of this placeholder raises the same KeyError."""

  def __init__(self, name):
    self.name = name

  def _fail(self, *a, **k):
    raise KeyError(f'undefined gin macro %{self.name}')

  __str__ = __float__ = __int__ = __bool__ = __iter__ = __len__ = __getitem__ = __call__ = __eq__ = __hash__ = _fail
  __add__ = __radd__ = __mul__ = __rmul__ = __lt__ = __gt__ = __fspath__ = _fail

  def __repr__(self):
    return f'<unbound gin macro %{self.name}>'


def _resolve(v: Any, raw: Dict[str, Any], stack: Tuple[str, ...] = ()) -> Any:
  if isinstance(v, _Macro):
    if v.name in stack:
      raise GinSyntaxError(f'macro cycle through %{v.name}')
    for key in (v.name, f'{v.name}/macro.value', f'macro.{v.name}'):
      if key in raw:
        return _resolve(raw[key], raw, stack + (v.name,))
    if not stack and _LAZY[0]:
      return _Unbound(v.name)         # gin fails only when the value is USED: keep a placeholder that raises then
    raise KeyError(f'undefined gin macro %{v.name}')
  if isinstance(v, tuple):
    return tuple(_resolve(e, raw, stack) for e in v)
  if isinstance(v, list):
    return [_resolve(e, raw, stack) for e in v]
  if isinstance(v, dict):
    return {_resolve(k, raw, stack): _resolve(e, raw, stack) for k, e in v.items()}
  return v


def _parse_into(text: str, raw: Dict[str, Any], search: Iterable[str], origin: str, depth: int = 0) -> None:
  if depth > 16:
    raise GinSyntaxError('include depth > 16')
  for line, stmt in _scan(text):
    where = f'{origin}:{line}'
    m = _INCLUDE.match(stmt)
    if m:
      for root in search:
        p = os.path.join(root, m.group(2))
        if os.path.exists(p):
          with open(p) as f:
            _parse_into(f.read(), raw, search, p, depth + 1)
          break
      else:
        raise FileNotFoundError(f'{where}: include {m.group(2)!r} not found under {list(search)}')
      continue
    if _IMPORT.match(stmt):
      continue
    m = _BINDING.match(stmt)
    if not m:
      raise GinSyntaxError(f'{where}: not a gin statement: {stmt[:60]!r}')
    raw[m.group(1)] = _parse_value(m.group(2).strip(), where)


def parse_config(text: str, search_paths: Iterable[str] = ('.',), origin: str = '<string>') -> Dict[str, Any]:
  """gin.parse_config for the binding subset: {name: value}, macros resolved, later bindings win."""
  raw: Dict[str, Any] = {}
  _parse_into(text, raw, tuple(search_paths), origin)
  return {k: _resolve(v, raw) for k, v in raw.items()}


def parse_config_file(path: str, search_paths: Optional[Iterable[str]] = None,
                      bindings: Iterable[str] = ()) -> Dict[str, Any]:
  """gin.parse_config_files_and_bindings([path], bindings) (train.py:86-91).

  Includes are looked up relative to the current directory, the file's
  directory and its parent (the shipped files say `include 'configs/defaults.gin'`
  relative to the repository root).
  """
  here = os.path.dirname(os.path.abspath(path))
  search = tuple(search_paths) if search_paths is not None else ('.', here, os.path.dirname(here))
  with open(path) as f:
    text = f.read()
  text += '\n' + '\n'.join(bindings)
  return parse_config(text, search, path)


# --------------------------------------------------------------- model config
# NerfModel attributes a gin file may bind that do not change the path built here (when they
# hold the listed values); anything else unknown raises in from_gin_bindings.
_EXPECTED_REFS = {
    # hypernerf/configs.py:29-40 registers flax.nn.* and jax.nn.* activations as gin configurables
    'NerfModel.activation': ('@jax.nn.relu', '@flax.nn.relu', '@nn.relu', '@relu'),
    'NerfModel.sigma_activation': ('@flax.nn.softplus', '@nn.softplus', '@softplus'),
    'NerfModel.warp_field_cls': ('@SE3Field',),
    'NerfModel.hyper_sheet_mlp_cls': ('@HyperSheetMLP',),
    'NerfModel.hyper_c_mlp_cls': ('@HyperSheetMLP',),
    'NerfModel.nerf_embed_cls': ('@nerf/GLOEmbed',),
    'NerfModel.warp_embed_cls': ('@warp/GLOEmbed',),
    'NerfModel.hyper_embed_cls': ('@hyper/GLOEmbed',),
    'NerfModel.hyper_c_embed_cls': ('@hyper_c/GLOEmbed', '@hyper/GLOEmbed'),
    'NerfModel.mask_embed_cls': ('@mask/GLOEmbed', '@warp/GLOEmbed'),
    'NerfModel.bone_warp_field_cls': ('@BoneSE3Field',),
    'SE3Field.activation': ('@jax.nn.relu', '@flax.nn.relu', '@nn.relu', '@relu'),
}
_EXPECTED_VALUES = {
    'NerfModel.nerf_embed_key': ('appearance', 'camera', 'time', 'warp'),   # unused: use_nerf_embed is fenced off
    'NerfModel.warp_embed_key': ('warp',),
    'NerfModel.hyper_embed_key': ('warp',),
    'NerfModel.hyper_sheet_use_input_points': (True,),
    'SE3Field.rotation_depth': (0,), 'SE3Field.pivot_depth': (0,), 'SE3Field.translation_depth': (0,),
    'SE3Field.norm': (None,),
    'HyperSheetMLP.use_residual': (False,),
    'MaskMLP.output_channels': (1,),
}
_IGNORED = ('NerfModel.hyper_c_hyper_input', 'NerfModel.use_hyper_c_embed', 'NerfModel.hyper_c_num_dims',
            'NerfModel.x_for_rgb_min_deg', 'NerfModel.x_for_rgb_max_deg',       # read only when window_x_in_rgb_condition
            'SE3Field.rotation_width', 'SE3Field.pivot_width', 'SE3Field.translation_width',
            'SE3Field.num_hyper_dims', 'nerf/GLOEmbed.num_dims', 'hyper/GLOEmbed.num_dims',
            'hyper_c/GLOEmbed.num_dims')
_EXTRA_MAP = {
    'SE3Field.skips': 'warp_skips', 'HyperSheetMLP.skips': 'hyper_sheet_skips', 'MaskMLP.skips': 'mask_skips',
    'mask/GLOEmbed.num_dims': 'mask_embed_dims',
}


def reference_class_defaults() -> _config.NerfDSConfig:
  """`NerfDSConfig` holding the reference's *class* defaults.

  NerfModel models.py:116-229, SE3Field warping.py:139-157, HyperSheetMLP
  modules.py:354-365, MaskMLP modules.py:396-407 -- what a gin file's bindings
  are applied on top of.  (`NerfDSConfig()` itself defaults to nerf_ds.gin.)
  """
  return _config.NerfDSConfig(
      use_viewdirs=True, nerf_trunk_depth=8, nerf_trunk_width=256, nerf_rgb_branch_depth=1, nerf_rgb_branch_width=128,
      nerf_skips=(4,), num_coarse_samples=196, num_fine_samples=196, use_stratified_sampling=True,
      use_white_background=False, use_linear_disparity=False, use_sample_at_infinity=True,
      spatial_point_min_deg=0, spatial_point_max_deg=10, hyper_point_min_deg=0, hyper_point_max_deg=4,
      viewdir_min_deg=0, viewdir_max_deg=4, use_posenc_identity=True, hyper_slice_method='none', use_hyper=True,
      hyper_use_warp_embed=True, use_hyper_for_sigma=True, hyper_num_dims=2, hyper_sheet_min_deg=0,
      hyper_sheet_max_deg=1, hyper_sheet_depth=6, hyper_sheet_width=64, hyper_sheet_skips=(4,),
      use_warp=False, warp_embed_dims=8, warp_min_deg=0, warp_max_deg=8, warp_use_posenc_identity=False,
      warp_trunk_depth=6, warp_trunk_width=128, warp_skips=(4,), predict_norm=False, norm_supervision_type='warped',
      stop_norm_gradient=True, norm_input_posenc=True, norm_input_min_deg=0, norm_input_max_deg=4,
      use_x_in_rgb_condition=False, window_x_in_rgb_condition=False, use_mask_in_warp=False, use_mask_in_hyper=False,
      use_predicted_mask=False, use_mask_embed=True, use_3d_mask=False, use_mask_sharp_weights=False,
      mask_embed_dims=8, mask_min_deg=0, mask_max_deg=6, mask_depth=6, mask_width=64, mask_skips=(4,),
      mask_output_relu=False, norm_type=None)


def model_config(bindings: Dict[str, Any], *, near: float, far: float, num_warp_embeds: int) -> _config.NerfDSConfig:
  """NerfDSConfig for a parsed gin file plus the three values `construct_nerf` takes from the
  datasource (models.py:1568-1600: near, far, embeddings_dict)."""
  kw, passed = {}, {}
  for key, value in bindings.items():
    if '.' not in key:
      continue                                  # a macro: matters only where something references it
    if key in _EXPECTED_REFS:
      if value not in _EXPECTED_REFS[key]:
        raise NotImplementedError(f'{key} = {value}: only {_EXPECTED_REFS[key][0]} is built')
    elif key in _EXPECTED_VALUES:
      if value not in _EXPECTED_VALUES[key]:
        raise NotImplementedError(f'{key} = {value!r} is outside the ray-marching path built here')
    elif key in _IGNORED:
      continue
    elif key in _EXTRA_MAP:
      kw[_EXTRA_MAP[key]] = tuple(value) if isinstance(value, list) else value
    elif key == 'NerfModel.nerf_skips':
      kw['nerf_skips'] = tuple(value)
    else:
      passed[key] = value
  base = reference_class_defaults().replace(near=float(near), far=float(far), num_warp_embeds=int(num_warp_embeds))
  return _config.from_gin_bindings(passed, base).replace(**kw)
