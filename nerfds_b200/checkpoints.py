"""Experiment-directory reader: Flax msgpack checkpoints + operative config.gin.

SURVEY.md section 8 row f-2.  `render.py:50-56,128` restores a trained
experiment with `gin.parse_config(exp_dir/'config.gin')` and
`flax.training.checkpoints.restore_checkpoint(exp_dir/'checkpoints', state)`;
`training.py:59-66` writes those files with `checkpoints.save_checkpoint`.
flax is not installed here (requirements_exact.txt pins flax==0.5.3,
msgpack==1.0.4), so this module restates flax.serialization's published wire
format on top of the `msgpack` package:

  * the file is `msgpack.packb(state_dict)` of the nested `to_state_dict()` tree
    (dict keys are str; dataclass fields, dict entries and list indices all
    become keys);
  * an array leaf is `ExtType(1, packb((shape, dtype.name, bytes)))` (C order);
    `ExtType(3, ...)` is a numpy scalar stored the same way, `ExtType(2, ...)`
    a python complex as `packb((re, im))`;
  * arrays above 2**30 bytes are split into
    `{'__msgpack_chunked_array__': True, 'shape': {'0': d0, ...}, 'chunks': {'0': a0, ...}}`.

Checkpoint files are `<dir>/checkpoint_<step>`; the newest is the one with the
largest step in natural-sort order, `*tmp*` files are skipped.

State tree written by the reference (model_utils.py:28-39; flax.optim):
  {'optimizer': {'target': {'model': <params>}, 'state': {'step': i32, 'param_states': ...}},
   'nerf_alpha': f32 | None, 'warp_alpha': ..., 'hyper_alpha': ..., 'hyper_sheet_alpha': ...,
   'norm_loss_weight': ..., 'norm_input_alpha': ..., 'norm_voxel_lr': ..., 'norm_voxel_ratio': ...}
The parameter tree uses the module names `nerfds_b200.params` already follows.
"""
from __future__ import annotations

import json
import os
import re
from types import SimpleNamespace
from typing import Any, Dict, Optional, Tuple

import msgpack
import numpy as np

from . import gin_reader, schedules
from .config import NerfDSConfig
from .model_utils import TrainState

MAX_CHUNK_SIZE = 2 ** 30
_EXT_NDARRAY, _EXT_COMPLEX, _EXT_NPSCALAR = 1, 2, 3
_CHUNK_KEY = '__msgpack_chunked_array__'


# ------------------------------------------------------------------ wire format
def _ndarray_to_bytes(a: np.ndarray) -> bytes:
  if a.dtype.hasobject or a.dtype.isalignedstruct:
    raise ValueError('object and structured arrays cannot be serialised')
  return msgpack.packb((a.shape, a.dtype.name, a.tobytes('C')), use_bin_type=True)


def _ndarray_from_bytes(data: bytes) -> np.ndarray:
  shape, dtype_name, buf = msgpack.unpackb(data, raw=True)
  name = dtype_name.decode() if isinstance(dtype_name, bytes) else dtype_name
  if name == 'bfloat16':
    raise NotImplementedError('bfloat16 leaves (the reference trains in float32)')
  return np.frombuffer(buf, dtype=np.dtype(name)).reshape(shape, order='C')


def _ext_pack(x):
  if isinstance(x, np.ndarray):
    return msgpack.ExtType(_EXT_NDARRAY, _ndarray_to_bytes(x))
  if isinstance(x, np.generic):
    return msgpack.ExtType(_EXT_NPSCALAR, _ndarray_to_bytes(np.asarray(x)))
  if isinstance(x, complex):
    return msgpack.ExtType(_EXT_COMPLEX, msgpack.packb((x.real, x.imag)))
  return x


def _ext_unpack(code, data):
  if code == _EXT_NDARRAY:
    return _ndarray_from_bytes(data)
  if code == _EXT_NPSCALAR:
    return _ndarray_from_bytes(data)[()]
  if code == _EXT_COMPLEX:
    re_, im = msgpack.unpackb(data)
    return complex(re_, im)
  return msgpack.ExtType(code, data)


def _index_dict(seq) -> Dict[str, Any]:
  return {str(i): v for i, v in enumerate(seq)}


def _from_index_dict(d: Dict[str, Any]) -> tuple:
  return tuple(d[str(i)] for i in range(len(d)))


def _chunk_leaves(tree, max_bytes):
  if isinstance(tree, dict):
    return {k: _chunk_leaves(v, max_bytes) for k, v in tree.items()}
  if isinstance(tree, np.ndarray) and tree.size * tree.dtype.itemsize > max_bytes:
    per = max(1, int(max_bytes / tree.dtype.itemsize))
    flat = tree.reshape(-1)
    return {_CHUNK_KEY: True, 'shape': _index_dict(tree.shape),
            'chunks': _index_dict([flat[i:i + per] for i in range(0, flat.size, per)])}
  return tree


def _unchunk_leaves(tree):
  if not isinstance(tree, dict):
    return tree
  if _CHUNK_KEY in tree:
    return np.concatenate(_from_index_dict(tree['chunks'])).reshape(_from_index_dict(tree['shape']))
  return {k: _unchunk_leaves(v) for k, v in tree.items()}


def to_state_dict(tree) -> Any:
  """flax.serialization.to_state_dict for the containers this package uses.  A `TrainState` made by
  `TrainState.create` has no Adam moments: the file it serialises to is an INFERENCE checkpoint (params, step and
  schedule scalars -- all `render.py` reads); a state restored from a reference checkpoint keeps its `param_states`."""
  if isinstance(tree, TrainState):
    opt = tree.optimizer
    state = getattr(opt, 'state', None)
    d = {'optimizer': {'target': to_state_dict(opt.target),
                       'state': to_state_dict(vars(state) if isinstance(state, SimpleNamespace) else state)}}
    d.update({k: to_state_dict(v) for k, v in tree.extra_params.items()})
    return d
  if isinstance(tree, dict):
    return {str(k): to_state_dict(v) for k, v in tree.items()}
  if isinstance(tree, (list, tuple)):
    return {str(i): to_state_dict(v) for i, v in enumerate(tree)}
  if isinstance(tree, (np.ndarray, np.generic)) or tree is None or isinstance(tree, (bool, int, float, str, bytes)):
    return tree
  if hasattr(tree, 'detach'):                       # torch tensor
    return tree.detach().cpu().numpy()
  raise TypeError(f'cannot serialise {type(tree)}')


def msgpack_serialize(state_dict, max_chunk_bytes: int = MAX_CHUNK_SIZE) -> bytes:
  """flax.serialization.msgpack_serialize."""
  return msgpack.packb(_chunk_leaves(state_dict, max_chunk_bytes), default=_ext_pack, strict_types=True)


def msgpack_restore(encoded: bytes):
  """flax.serialization.msgpack_restore: nested dicts of numpy arrays / python scalars."""
  return _unchunk_leaves(msgpack.unpackb(encoded, ext_hook=_ext_unpack, raw=False, strict_map_key=False))


# ------------------------------------------------------------------ files
def _natural_key(path: str):
  return [int(t) if t.isdigit() else t for t in re.split(r'(\d+)', os.path.basename(path))]


def latest_checkpoint(ckpt_dir: str, prefix: str = 'checkpoint_') -> Optional[str]:
  """checkpoints.latest_checkpoint: newest `<prefix><step>` by natural sort, tmp files skipped."""
  if not os.path.isdir(ckpt_dir):
    return None
  files = [os.path.join(ckpt_dir, f) for f in os.listdir(ckpt_dir) if f.startswith(prefix) and 'tmp' not in f]
  return max(files, key=_natural_key) if files else None


def restore_checkpoint(ckpt_dir: str, target: Optional[TrainState] = None, step: Optional[int] = None,
                       prefix: str = 'checkpoint_'):
  """checkpoints.restore_checkpoint (render.py:128).

  `ckpt_dir` may be a directory or one checkpoint file.  With no checkpoint on
  disk the `target` is returned unchanged, as flax does.  With `target=None` the
  raw state dict comes back; with a `TrainState` target a restored `TrainState`.
  """
  if os.path.isfile(ckpt_dir):
    path = ckpt_dir
  elif step is not None:
    path = os.path.join(ckpt_dir, f'{prefix}{step}')
    if not os.path.exists(path):
      raise ValueError(f'Matching checkpoint not found: {path}')
  else:
    path = latest_checkpoint(ckpt_dir, prefix)
    if path is None:
      return target
  with open(path, 'rb') as f:
    state_dict = msgpack_restore(f.read())
  return state_dict if target is None else train_state_from_dict(state_dict)


def save_checkpoint(ckpt_dir: str, state, step: int, prefix: str = 'checkpoint_', keep: int = 2,
                    overwrite: bool = False) -> str:
  """checkpoints.save_checkpoint (training.py:59-66): atomic write, keeps the newest `keep` files."""
  os.makedirs(ckpt_dir, exist_ok=True)
  path = os.path.join(ckpt_dir, f'{prefix}{int(step)}')
  newest = latest_checkpoint(ckpt_dir, prefix)
  if newest is not None and not overwrite and _natural_key(newest) >= _natural_key(path):
    raise ValueError(f'Trying to save an outdated checkpoint at step {step}; newest is {newest}')
  tmp = os.path.join(ckpt_dir, f'{prefix}tmp')
  with open(tmp, 'wb') as f:
    f.write(msgpack_serialize(to_state_dict(state)))
  os.replace(tmp, path)
  files = sorted((os.path.join(ckpt_dir, f) for f in os.listdir(ckpt_dir)
                  if f.startswith(prefix) and 'tmp' not in f), key=_natural_key)
  for old in files[:-keep] if keep > 0 else []:
    os.remove(old)
  return path


# ------------------------------------------------------------------ state -> path inputs
_SCALARS = ('nerf_alpha', 'warp_alpha', 'hyper_alpha', 'hyper_sheet_alpha', 'norm_loss_weight', 'norm_input_alpha',
            'norm_voxel_lr', 'norm_voxel_ratio')


def _f32_tree(tree):
  if isinstance(tree, dict):
    return {k: _f32_tree(v) for k, v in tree.items()}
  return np.ascontiguousarray(np.asarray(tree, dtype=np.float32))


def train_state_from_dict(sd: Dict[str, Any]) -> TrainState:
  """The reference's state tree -> `TrainState` (params as float32 numpy, step as int)."""
  try:
    opt = sd['optimizer']
    params = opt['target']['model']
  except (KeyError, TypeError):
    raise ValueError("not a NeRF-DS train state: missing optimizer/target/model") from None
  raw_state = opt.get('state', {}) or {}
  step = int(np.asarray(raw_state.get('step', 0)).reshape(-1)[0])
  scal = {k: (None if sd.get(k) is None else float(np.asarray(sd[k]).reshape(-1)[0])) for k in _SCALARS}
  # everything else the optimizer state holds (flax.optim Adam: `param_states` = grad_ema / grad_sq_ema per leaf) is
  # kept verbatim, so that a state restored here and saved again is still loadable by the reference's
  # `checkpoints.restore_checkpoint(dir, state)` (from_state_dict raises on a missing `param_states`)
  state = SimpleNamespace(step=step, **{k: v for k, v in raw_state.items() if k != 'step'})
  return TrainState(optimizer=SimpleNamespace(target={'model': _f32_tree(params)}, state=state), **scal)


def check_params(cfg: NerfDSConfig, params: Dict) -> None:
  """Every weight the configured path reads must be present with the configured shape."""
  from .params import flatten_params, init_params
  want = {n: v.shape for n, v in flatten_params(init_params(cfg, 0))}
  have = {n: v.shape for n, v in flatten_params(params)}
  for n, shape in want.items():
    if n not in have:
      raise KeyError(f'checkpoint has no parameter {n} (config and checkpoint disagree)')
    if tuple(have[n]) != tuple(shape):
      raise ValueError(f'parameter {n}: checkpoint {tuple(have[n])} vs config {tuple(shape)}')


def load_experiment(exp_dir: str, *, data_dir: Optional[str] = None, near: Optional[float] = None,
                    far: Optional[float] = None, step: Optional[int] = None, scheduled_step: Optional[int] = None
                    ) -> Tuple[NerfDSConfig, Dict, Dict[str, Optional[float]], TrainState, Dict[str, Any]]:
  """render.py:36-131 up to the restored state: (cfg, params, extra_params, state, gin bindings).

  near/far come from `<data_dir>/scene.json` (datasets/nerfies.py:35-57) unless
  given; the warp-embedding table size comes from the checkpoint itself (the
  reference derives it from the dataset's ids, models.py:236, which is what
  sized the table at training time).  `extra_params` are the checkpoint's own
  scalars (render.py:128-129); `scheduled_step` re-evaluates the gin schedules
  at another step instead.
  """
  bindings = gin_reader.parse_config_file(os.path.join(exp_dir, 'config.gin'))
  if near is None or far is None:
    if data_dir is None:
      data_dir = bindings.get('NerfiesDataSource.data_dir') or bindings.get('data_dir')
    if data_dir is None:
      raise ValueError('near/far not given and no data_dir to read scene.json from')
    with open(os.path.join(data_dir, 'scene.json')) as f:
      scene = json.load(f)
    near = scene['near'] if near is None else near
    far = scene['far'] if far is None else far
  state = restore_checkpoint(os.path.join(exp_dir, 'checkpoints'), TrainState.create({}, {}), step=step)
  if not state.optimizer.target['model']:
    raise FileNotFoundError(f'no checkpoint under {exp_dir}/checkpoints')
  params = state.optimizer.target['model']
  n_embed = params['warp_embed']['embed']['embedding'].shape[0] if 'warp_embed' in params else 1
  cfg = gin_reader.model_config(bindings, near=near, far=far, num_warp_embeds=n_embed)
  cfg.validate()
  check_params(cfg, params)
  extra = dict(state.extra_params)
  if scheduled_step is not None:
    extra.update(schedules.extra_params_at(bindings, scheduled_step))
  return cfg, params, extra, state, bindings
