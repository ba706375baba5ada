"""gin reader, schedules, Flax msgpack checkpoints (CPU)."""
import math
import os

import numpy as np
import pytest

from nerfds_b200 import checkpoints as ckpt
from nerfds_b200 import gin_reader, schedules
from nerfds_b200.config import nerf_ds_config, tiny_config
from nerfds_b200.model_utils import TrainState
from nerfds_b200.params import flatten_params, init_params

GIN_BASE = """
# base file: macros and schedule tables
warp_min_deg = 0
warp_max_deg = 8                      # overridden by the including file
hyper_point_max_deg = 1
ANNEALED = {
  'type': 'linear',
  'initial_value': %warp_min_deg,
  'final_value': %warp_max_deg,       # resolved lazily
  'num_steps': 80000,
}
DELAYED = {
  'type': 'piecewise',
  'schedules': [
    (10000, ('constant', 0.0)),
    (0, ('linear', 0.0, 4, 2000))
  ],
}
NerfModel.use_viewdirs = True
NerfModel.use_posenc_identity = False
NerfModel.warp_embed_cls = @warp/GLOEmbed
warp/GLOEmbed.num_dims = 8
SE3Field.min_deg = %warp_min_deg
SE3Field.max_deg = %warp_max_deg
NerfModel.warp_field_cls = @SE3Field
TrainConfig.nerf_alpha_schedule = ('constant', 8)
"""

GIN_MAIN = """
include 'base.gin'
import hypernerf.something
warp_max_deg = 4
NerfModel.num_coarse_samples = 64
NerfModel.num_fine_samples = \\
    32
NerfModel.spatial_point_max_deg = 8
NerfModel.hyper_point_max_deg = %hyper_point_max_deg
NerfModel.norm_type = 'none'  # a '#' inside 'a # string' must survive: see Label.text below
Label.text = 'a # string, with = and @ref and %macro'
NerfModel.activation = @jax.nn.relu
NerfModel.hyper_slice_method = 'bendy_sheet'
NerfModel.hyper_sheet_mlp_cls = @HyperSheetMLP
HyperSheetMLP.output_channels = 2
HyperSheetMLP.max_deg = 6
NerfModel.use_warp = True
NerfModel.predict_norm = True
NerfModel.use_x_in_rgb_condition = True
NerfModel.use_mask_in_warp = True
NerfModel.use_mask_in_hyper = True
NerfModel.use_predicted_mask = True
NerfModel.use_3d_mask = True
NerfModel.use_mask_sharp_weights = True
MaskMLP.depth = 8
MaskMLP.width = 128
MaskMLP.output_activation = @jax.nn.relu
TrainConfig.warp_alpha_schedule = %ANNEALED
TrainConfig.hyper_alpha_schedule = ('constant', %hyper_point_max_deg)
TrainConfig.hyper_sheet_alpha_schedule = ('constant', 6)
SpecularConfig.norm_input_alpha_schedule = %DELAYED
EvalConfig.num_val_eval = None
EvalConfig.niter = -3
"""


@pytest.fixture
def gin_files(tmp_path):
  (tmp_path / 'base.gin').write_text(GIN_BASE)
  (tmp_path / 'main.gin').write_text(GIN_MAIN)
  return tmp_path


def test_gin_reader_syntax(gin_files):
  b = gin_reader.parse_config_file(str(gin_files / 'main.gin'), bindings=["NerfiesDataSource.data_dir = '/data/x'"])
  assert b['warp_max_deg'] == 4 and b['SE3Field.max_deg'] == 4          # later macro wins, also for earlier uses
  assert b['TrainConfig.warp_alpha_schedule'] == {'type': 'linear', 'initial_value': 0, 'final_value': 4,
                                                  'num_steps': 80000}
  assert b['NerfModel.num_fine_samples'] == 32                           # backslash continuation
  assert b['Label.text'] == 'a # string, with = and @ref and %macro'
  assert b['NerfModel.activation'] == '@jax.nn.relu' and b['NerfModel.warp_embed_cls'] == '@warp/GLOEmbed'
  assert b['SpecularConfig.norm_input_alpha_schedule']['schedules'][1] == (0, ('linear', 0.0, 4, 2000))
  assert b['EvalConfig.num_val_eval'] is None and b['EvalConfig.niter'] == -3
  assert b['NerfiesDataSource.data_dir'] == '/data/x'
  lazy = gin_reader.parse_config('A.b = %nope\nC.d = 2')          # like gin: an undefined macro fails when USED
  assert lazy['C.d'] == 2
  with pytest.raises(KeyError):
    float(lazy['A.b'])
  with pytest.raises(KeyError):
    os.path.join(lazy['A.b'], 'x')
  with pytest.raises(gin_reader.GinSyntaxError):
    gin_reader.parse_config('A.b = (1, 2')
  with pytest.raises(gin_reader.GinSyntaxError):
    gin_reader.parse_config('A.b = __import__("os")')
  with pytest.raises(gin_reader.GinSyntaxError):
    gin_reader.parse_config('X = %Y\nY = %X')
  with pytest.raises(FileNotFoundError):
    gin_reader.parse_config("include 'missing.gin'")


def test_model_config_from_gin_equals_nerf_ds_preset(gin_files):
  b = gin_reader.parse_config_file(str(gin_files / 'main.gin'))
  cfg = gin_reader.model_config(b, near=0.1, far=2.5, num_warp_embeds=100)
  assert cfg == nerf_ds_config(num_coarse_samples=64, num_fine_samples=32, near=0.1, far=2.5, num_warp_embeds=100)
  cfg.validate()
  # bindings are applied over the reference's class defaults, not over the preset
  bare = gin_reader.model_config({}, near=0.0, far=1.0, num_warp_embeds=1)
  assert (bare.num_coarse_samples, bare.spatial_point_max_deg, bare.use_posenc_identity, bare.use_warp,
          bare.hyper_slice_method, bare.mask_width) == (196, 10, True, False, 'none', 64)
  with pytest.raises(NotImplementedError):
    gin_reader.model_config({'NerfModel.warp_field_cls': '@TranslationField'}, near=0, far=1, num_warp_embeds=1)
  with pytest.raises(KeyError):
    gin_reader.model_config({'NerfModel.no_such_attr': 1}, near=0, far=1, num_warp_embeds=1)


def test_shipped_reference_config_parses_to_the_preset():
  path = '/root/reference/configs/nerf_ds.gin'
  if not os.path.exists(path):
    pytest.skip('reference tree not present on this machine')
  b = gin_reader.parse_config_file(path, search_paths=['/root/reference'], bindings=["data_dir = '/x'"])
  cfg = gin_reader.model_config(b, near=0.1, far=2.5, num_warp_embeds=100)
  assert cfg == nerf_ds_config(num_coarse_samples=64, num_fine_samples=64, near=0.1, far=2.5, num_warp_embeds=100)
  ep = schedules.extra_params_at(b, 250000)
  assert ep == {'nerf_alpha': 8.0, 'warp_alpha': 4.0, 'hyper_alpha': 1.0, 'hyper_sheet_alpha': 6.0,
                'norm_input_alpha': 4.0}


def test_schedules_known_answers():
  S = schedules.from_config
  assert S(None)(5) is None
  assert S(('constant', 3))(10 ** 6) == 3.0
  lin = S({'type': 'linear', 'initial_value': 0, 'final_value': 4, 'num_steps': 50000})
  assert lin(0) == 0.0 and lin(12500) == 1.0 and lin(50000) == 4.0 and lin(10 ** 7) == 4.0
  assert S(('linear', 1.0, 2.0, 0))(0) == 2.0
  ex = S(('exponential', 1, 0.1, 30000))
  assert ex(0) == 1.0 and ex(30000) == 0.1 and math.isclose(ex(29999), 0.1) and math.isclose(ex(14999.5), 10 ** -0.5)
  with pytest.raises(ValueError):
    S(('exponential', 0.1, 1.0, 10))
  ce = S(('cosine_easing', 0.01, 1e-8, 100000))
  assert math.isclose(ce(0), 0.01) and math.isclose(ce(50000), (0.01 + 1e-8) / 2) and math.isclose(ce(10 ** 6), 1e-8)
  st = S(('step', 1.0, 100, 0.5, 3))
  assert [st(0), st(99), st(100), st(250), st(300), st(10 ** 5)] == [1.0, 1.0, 0.5, 0.25, 0.125, 0.125]
  pw = S({'type': 'piecewise', 'schedules': [(10000, ('constant', 0.0)), (0, ('linear', 0.0, 4, 2000))]})
  assert [pw(0), pw(9999), pw(10000), pw(11000), pw(12000), pw(99999)] == [0.0, 0.0, 0.0, 2.0, 4.0, 4.0]
  pw3 = S(('piecewise', [(50000, ('constant', 0)), (50000, ('linear', 0, 4.0, 50000)), (150000, ('constant', 4.0))]))
  assert pw3(49999) == 0.0 and pw3(75000) == 2.0 and pw3(100000) == 4.0 and pw3(10 ** 6) == 4.0
  dl = S({'type': 'delayed', 'delay_steps': 2500, 'delay_mult': 0.01, 'base_schedule': ('constant', 2.0)})
  assert math.isclose(dl(0), 0.02) and math.isclose(dl(2500), 2.0) and math.isclose(dl(10 ** 5), 2.0)
  assert math.isclose(dl(1250), 2.0 * (0.01 + 0.99 * math.sin(math.pi / 4)))


def test_msgpack_wire_format_known_answer():
  """Byte-level layout of flax.serialization: ext type 1 = packb((shape, dtype name, C bytes))."""
  payload = bytes([0x93, 0x91, 0x02, 0xa5]) + b'int32' + bytes([0xc4, 0x08, 1, 0, 0, 0, 2, 0, 0, 0])
  wire = bytes([0x82, 0xa1]) + b'a' + bytes([0xc7, len(payload), 0x01]) + payload + bytes([0xa1]) + b'n' + b'\xc0'
  tree = {'a': np.array([1, 2], np.int32), 'n': None}
  assert ckpt.msgpack_serialize(tree) == wire
  back = ckpt.msgpack_restore(wire)
  assert back['n'] is None and back['a'].dtype == np.int32 and back['a'].tolist() == [1, 2]
  # numpy scalars travel as ext type 3, complex as ext type 2
  enc = ckpt.msgpack_serialize({'s': np.float32(1.5), 'c': 1 + 2j, 'i': 7, 'f': 0.25, 'b': True, 't': 'x'})
  assert b'\xc7' in enc or b'\xd7' in enc or b'\xd8' in enc
  dec = ckpt.msgpack_restore(enc)
  assert dec['s'] == np.float32(1.5) and isinstance(dec['s'], np.float32) and dec['c'] == 1 + 2j
  assert (dec['i'], dec['f'], dec['b'], dec['t']) == (7, 0.25, True, 'x')
  # 0-d and empty arrays, fortran-ordered input is stored in C order
  a = np.asfortranarray(np.arange(6, dtype=np.float32).reshape(2, 3))
  dec = ckpt.msgpack_restore(ckpt.msgpack_serialize({'z': np.zeros((0, 3), np.float32), 'd': np.array(3, np.int32), 'f': a}))
  assert dec['z'].shape == (0, 3) and dec['d'].shape == () and int(dec['d']) == 3 and np.array_equal(dec['f'], a)


def test_msgpack_chunked_arrays():
  a = np.arange(1000, dtype=np.float32).reshape(10, 100)
  enc = ckpt.msgpack_serialize({'w': {'kernel': a}}, max_chunk_bytes=1024)
  raw = ckpt.msgpack.unpackb(enc, ext_hook=ckpt._ext_unpack, raw=False)
  node = raw['w']['kernel']
  assert node['__msgpack_chunked_array__'] is True and node['shape'] == {'0': 10, '1': 100}
  assert sorted(node['chunks'], key=int) == ['0', '1', '2', '3'] and node['chunks']['0'].shape == (256,)
  assert np.array_equal(ckpt.msgpack_restore(enc)['w']['kernel'], a)


def _write_experiment(root, cfg, params, extra, step, gin_text):
  state = TrainState.create(params, extra)
  state.optimizer.state.step = np.int32(step)
  state.optimizer.state.param_states = {'dummy': {'grad_ema': np.zeros(3, np.float32)}}
  os.makedirs(root, exist_ok=True)
  with open(os.path.join(root, 'config.gin'), 'w') as f:
    f.write(gin_text)
  return ckpt.save_checkpoint(os.path.join(root, 'checkpoints'), state, step)


def test_experiment_dir_round_trip(tmp_path, gin_files):
  cfg = nerf_ds_config(num_coarse_samples=64, num_fine_samples=32, near=0.2, far=3.0, num_warp_embeds=17)
  params = init_params(cfg, 3)
  extra = {'nerf_alpha': 8.0, 'warp_alpha': 1.25, 'hyper_alpha': 1.0, 'hyper_sheet_alpha': 6.0,
           'norm_loss_weight': 0.001, 'norm_input_alpha': 2.0}
  exp = str(tmp_path / 'exp')
  gin_text = GIN_BASE + GIN_MAIN.replace("include 'base.gin'", '')
  _write_experiment(exp, cfg, params, extra, 1000, gin_text)
  p2 = _write_experiment(exp, cfg, params, extra, 12000, gin_text)
  _write_experiment(exp, cfg, params, extra, 20000, gin_text)
  names = sorted(os.listdir(os.path.join(exp, 'checkpoints')))
  assert names == ['checkpoint_12000', 'checkpoint_20000']                 # keep=2, natural order (12000 < 20000 > 1000)
  assert ckpt.latest_checkpoint(os.path.join(exp, 'checkpoints')).endswith('checkpoint_20000')
  with pytest.raises(ValueError):
    _write_experiment(exp, cfg, params, extra, 500, gin_text)

  (tmp_path / 'data').mkdir()
  (tmp_path / 'data' / 'scene.json').write_text('{"scale": 0.1, "center": [0, 0, 0], "near": 0.2, "far": 3.0}')
  cfg2, params2, extra2, state, bindings = ckpt.load_experiment(exp, data_dir=str(tmp_path / 'data'))
  assert cfg2 == cfg and state.optimizer.state.step == 20000
  a, b = dict(flatten_params(params)), dict(flatten_params(params2))
  assert a.keys() == b.keys() and all(np.array_equal(a[k], b[k]) and b[k].dtype == np.float32 for k in a)
  assert extra2['warp_alpha'] == 1.25 and extra2['norm_input_alpha'] == 2.0 and extra2['norm_voxel_lr'] is None
  _, _, extra3, st3, _ = ckpt.load_experiment(exp, near=0.2, far=3.0, step=12000, scheduled_step=40000)
  assert st3.optimizer.state.step == 12000 and extra3['warp_alpha'] == 2.0 and extra3['norm_input_alpha'] == 4.0
  assert p2.endswith('checkpoint_12000')
  raw = ckpt.restore_checkpoint(p2)                                        # a file path, raw state dict
  assert set(raw) >= {'optimizer', 'nerf_alpha', 'norm_voxel_ratio'} and raw['norm_voxel_ratio'] is None
  assert int(raw['optimizer']['state']['step']) == 12000
  assert raw['optimizer']['state']['param_states']['dummy']['grad_ema'].shape == (3,)
  with pytest.raises(ValueError):
    ckpt.restore_checkpoint(os.path.join(exp, 'checkpoints'), step=7)
  sentinel = TrainState.create({}, {})
  assert ckpt.restore_checkpoint(str(tmp_path / 'nowhere'), sentinel) is sentinel


def test_config_checkpoint_mismatch_is_reported(tmp_path):
  cfg = tiny_config()
  params = init_params(cfg, 0)
  with pytest.raises(ValueError):
    ckpt.check_params(cfg.replace(nerf_trunk_width=cfg.nerf_trunk_width * 2), params)
  del params['nerf_mlps_fine']
  with pytest.raises(KeyError):
    ckpt.check_params(cfg, params)
  with pytest.raises(ValueError):
    ckpt.train_state_from_dict({'params': {}})


OPERATIVE = r'''
import hypernerf.configs

# Macros:
# ==============================================================================
hyper_point_max_deg = 1
hyper_sheet_max_deg = 6
warp_max_deg = 4
warp_min_deg = 0

# Parameters for warp/GLOEmbed:
# ==============================================================================
warp/GLOEmbed.num_dims = 8

# Parameters for HyperSheetMLP:
# ==============================================================================
HyperSheetMLP.depth = 6
HyperSheetMLP.max_deg = %hyper_sheet_max_deg
HyperSheetMLP.min_deg = 0
HyperSheetMLP.output_channels = 2
HyperSheetMLP.skips = (4,)
HyperSheetMLP.width = 64

# Parameters for MaskMLP:
# ==============================================================================
MaskMLP.depth = 8
MaskMLP.output_activation = @jax.nn.relu
MaskMLP.width = 128

# Parameters for NerfModel:
# ==============================================================================
NerfModel.activation = @jax.nn.relu
NerfModel.hyper_embed_cls = @hyper/GLOEmbed
NerfModel.hyper_point_max_deg = %hyper_point_max_deg
NerfModel.hyper_point_min_deg = 0
NerfModel.hyper_sheet_mlp_cls = @HyperSheetMLP
NerfModel.hyper_slice_method = 'bendy_sheet'
NerfModel.hyper_use_warp_embed = True
NerfModel.nerf_skips = (4,)
NerfModel.norm_type = 'none'
NerfModel.num_coarse_samples = 64
NerfModel.num_fine_samples = 64
NerfModel.predict_norm = True
NerfModel.sigma_activation = @flax.nn.softplus
NerfModel.spatial_point_max_deg = 8
NerfModel.spatial_point_min_deg = 0
NerfModel.use_3d_mask = True
NerfModel.use_mask_in_hyper = True
NerfModel.use_mask_in_warp = True
NerfModel.use_mask_sharp_weights = True
NerfModel.use_posenc_identity = False
NerfModel.use_predicted_mask = True
NerfModel.use_warp = True
NerfModel.use_x_in_rgb_condition = True
NerfModel.warp_embed_cls = @warp/GLOEmbed
NerfModel.warp_field_cls = @SE3Field

# Parameters for SE3Field:
# ==============================================================================
SE3Field.activation = @jax.nn.relu
SE3Field.max_deg = %warp_max_deg
SE3Field.min_deg = %warp_min_deg
SE3Field.skips = (4,)
SE3Field.trunk_depth = 6
SE3Field.trunk_width = 128
SE3Field.use_posenc_identity = False

# Parameters for TrainConfig:
# ==============================================================================
TrainConfig.warp_alpha_schedule = \
    {'final_value': %warp_max_deg,
     'initial_value': %warp_min_deg,
     'num_steps': 50000,
     'type': 'linear'}
TrainConfig.hyper_sheet_alpha_schedule = ('constant', %hyper_sheet_max_deg)
'''


def test_operative_config_dump_format():
  """The layout `gin.operative_config_str()` writes to <exp_dir>/config.gin (train.py:335-338): section comments,
  alphabetical bindings, macros first, backslash-continued multi-line literals."""
  b = gin_reader.parse_config(OPERATIVE)
  cfg = gin_reader.model_config(b, near=0.1, far=2.5, num_warp_embeds=100)
  assert cfg == nerf_ds_config(num_coarse_samples=64, num_fine_samples=64, near=0.1, far=2.5, num_warp_embeds=100)
  assert schedules.from_config(b['TrainConfig.warp_alpha_schedule'])(25000) == 2.0
  ep = schedules.extra_params_at(b, 10)
  assert ep['nerf_alpha'] is None and ep['norm_input_alpha'] == 4.0     # not in the dump -> the class defaults (configs.py:65-68, 219)


def test_msgpack_round_trip_property():
  """Random state trees (nested dicts of float / int arrays of any rank, numpy scalars, None, python scalars)
  survive serialize -> restore exactly, whatever the chunk threshold."""
  from hypothesis import given, settings, strategies as st
  from hypothesis.extra import numpy as hnp
  leaf = st.one_of(
      hnp.arrays(dtype=st.sampled_from([np.float32, np.int32, np.uint32, np.float64, np.uint8]),
                 shape=hnp.array_shapes(min_dims=0, max_dims=3, min_side=0, max_side=5),
                 elements=st.integers(0, 100)),
      st.none(), st.integers(-2 ** 31, 2 ** 31), st.booleans(),
      st.floats(allow_nan=False, allow_infinity=False, width=32).map(np.float32))
  tree = st.recursive(leaf, lambda c: st.dictionaries(st.text('abc/_0', min_size=1, max_size=4), c, max_size=4), max_leaves=12)

  def same(a, b):
    if isinstance(a, dict):
      return isinstance(b, dict) and a.keys() == b.keys() and all(same(a[k], b[k]) for k in a)
    if isinstance(a, np.ndarray):
      return isinstance(b, np.ndarray) and a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b)
    if isinstance(a, np.generic):
      return type(a) is type(b) and a == b
    return type(a) is type(b) and a == b

  @settings(max_examples=60, deadline=None)
  @given(tree.filter(lambda t: isinstance(t, dict)), st.sampled_from([8, 64, 2 ** 30]))
  def check(t, chunk):
    assert same(t, ckpt.msgpack_restore(ckpt.msgpack_serialize(t, max_chunk_bytes=chunk)))

  check()


def test_optimizer_state_survives_a_restore_save_round_trip(tmp_path):
  """A reference checkpoint carries flax.optim Adam moments (`optimizer/state/param_states`); restoring it here and
  saving it again keeps them, so the file stays loadable upstream.  `('constant', None)` disables a scalar."""
  cfg = tiny_config()
  from nerfds_b200.params import init_params
  params = init_params(cfg, 1)
  moments = {'model': {'x': {'grad_ema': np.arange(3, dtype=np.float32), 'grad_sq_ema': np.ones(3, np.float32)}}}
  sd = {'optimizer': {'target': {'model': params}, 'state': {'step': np.int32(9), 'param_states': moments}},
        'nerf_alpha': 3.0, 'warp_alpha': None}
  d = tmp_path / 'ck'
  d.mkdir()
  with open(d / 'checkpoint_9', 'wb') as f:
    f.write(ckpt.msgpack_serialize(sd))
  st = ckpt.restore_checkpoint(str(d), TrainState.create(params, {}))
  assert st.optimizer.state.step == 9
  ckpt.save_checkpoint(str(d), st, 10)
  again = ckpt.restore_checkpoint(str(d))
  np.testing.assert_array_equal(again['optimizer']['state']['param_states']['model']['x']['grad_ema'], moments['model']['x']['grad_ema'])
  assert int(again['optimizer']['state']['step']) == 9 and again['nerf_alpha'] == 3.0
  assert schedules.from_config(('constant', None)).get(5) is None
