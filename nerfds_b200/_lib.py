"""ctypes binding of libnerfds_b200.so (include/nerfds_b200.h).

The product path has no CPU fallback: if the CUDA library is missing or does
not load, importing the renderer raises.
"""
from __future__ import annotations

import ctypes as C
import os

from .config import NerfDSConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('NDSR_LIBRARY') or os.path.join(_HERE, 'lib', 'libnerfds_b200.so')   # (override: A/B experiments)

NDSR_ABI_VERSION = 2
NDSR_MAX_MIRRORS = 15
ENGINES = {'auto': 0, 'simt': 1, 'tc': 2}
ENGINE_NAMES = {v: k for k, v in ENGINES.items()}
PRECISIONS = {'mixed': 0, 'fp16': 1, 'split3': 2}

_i32, _f32 = C.c_int32, C.c_float


class ndsr_config(C.Structure):
  _fields_ = [('size', C.c_uint32), ('abi_version', C.c_uint32),
              ('near_', _f32), ('far_', _f32), ('num_warp_embeds', _i32),
              ('use_viewdirs', _i32),
              ('trunk_depth', _i32), ('trunk_width', _i32), ('trunk_skip', _i32),
              ('rgb_depth', _i32), ('rgb_width', _i32),
              ('num_coarse_samples', _i32), ('num_fine_samples', _i32),
              ('use_stratified_sampling', _i32), ('use_white_background', _i32),
              ('use_linear_disparity', _i32), ('use_sample_at_infinity', _i32),
              ('spatial_min_deg', _i32), ('spatial_max_deg', _i32),
              ('hyper_point_min_deg', _i32), ('hyper_point_max_deg', _i32),
              ('viewdir_min_deg', _i32), ('viewdir_max_deg', _i32),
              ('use_posenc_identity', _i32),
              ('use_hyper_sheet', _i32), ('hyper_num_dims', _i32),
              ('hyper_sheet_min_deg', _i32), ('hyper_sheet_max_deg', _i32),
              ('hyper_sheet_depth', _i32), ('hyper_sheet_width', _i32), ('hyper_sheet_skip', _i32),
              ('use_warp', _i32), ('warp_embed_dims', _i32),
              ('warp_min_deg', _i32), ('warp_max_deg', _i32), ('warp_use_posenc_identity', _i32),
              ('warp_depth', _i32), ('warp_width', _i32), ('warp_skip', _i32),
              ('predict_norm', _i32), ('norm_input_posenc', _i32),
              ('norm_input_min_deg', _i32), ('norm_input_max_deg', _i32),
              ('use_x_in_rgb_condition', _i32),
              ('use_mask_in_warp', _i32), ('use_mask_in_hyper', _i32), ('use_predicted_mask', _i32),
              ('use_mask_sharp_weights', _i32),
              ('mask_embed_dims', _i32), ('mask_min_deg', _i32), ('mask_max_deg', _i32),
              ('mask_depth', _i32), ('mask_width', _i32), ('mask_skip', _i32), ('mask_output_relu', _i32),
              ('engine', _i32), ('precision', _i32)]


class ndsr_tensor(C.Structure):
  _fields_ = [('name', C.c_char_p), ('data', C.c_void_p), ('rows', C.c_int64), ('cols', C.c_int64)]


class ndsr_extra_params(C.Structure):
  _fields_ = [('nerf_alpha', _f32), ('warp_alpha', _f32), ('hyper_alpha', _f32), ('hyper_sheet_alpha', _f32),
              ('norm_input_alpha', _f32), ('mask_ratio', _f32), ('sharp_weights_std', _f32),
              ('near_override', _f32), ('far_override', _f32),
              ('use_predicted_norm', _i32), ('use_sigma_gradient', _i32),
              ('sample_at_infinity_override', _i32),
              ('filter_flags', _i32), ('dust_threshold', _f32), ('bounding_box', _f32 * 6)]


OUTPUT_FIELDS = ['rgb', 'depth', 'med_depth', 'acc', 'ray_norm', 'ray_rotation_field', 'ray_translation_field',
                 'ray_delta_x', 'ray_hyper_points', 'ray_predicted_mask', 'med_points',
                 'z_vals', 'weights', 'alpha', 'accum_prod', 'sigma', 'sharp_weights', 'back_facing',
                 'predicted_mask', 'points', 'warped_points', 'delta_x', 'predicted_norm', 'target_norm']


class ndsr_outputs(C.Structure):
  _fields_ = [(name, C.c_void_p) for name in OUTPUT_FIELDS]


def _single_skip(skips, what):
  skips = tuple(skips)
  if len(skips) == 0:
    return -1
  if len(skips) != 1:
    raise NotImplementedError(f'{what}: only zero or one skip connection is built')
  return int(skips[0])


def to_c_config(cfg: NerfDSConfig, engine='auto', precision='split3') -> ndsr_config:
  cfg.validate()
  c = ndsr_config()
  c.size = C.sizeof(ndsr_config)
  c.abi_version = NDSR_ABI_VERSION
  c.near_, c.far_ = cfg.near, cfg.far
  c.num_warp_embeds = cfg.num_warp_embeds
  c.use_viewdirs = cfg.use_viewdirs
  c.trunk_depth, c.trunk_width = cfg.nerf_trunk_depth, cfg.nerf_trunk_width
  c.trunk_skip = _single_skip(cfg.nerf_skips, 'nerf_skips')
  c.rgb_depth, c.rgb_width = cfg.nerf_rgb_branch_depth, cfg.nerf_rgb_branch_width
  c.num_coarse_samples, c.num_fine_samples = cfg.num_coarse_samples, cfg.num_fine_samples
  c.use_stratified_sampling = cfg.use_stratified_sampling
  c.use_white_background = cfg.use_white_background
  c.use_linear_disparity = cfg.use_linear_disparity
  c.use_sample_at_infinity = cfg.use_sample_at_infinity
  c.spatial_min_deg, c.spatial_max_deg = cfg.spatial_point_min_deg, cfg.spatial_point_max_deg
  c.hyper_point_min_deg, c.hyper_point_max_deg = cfg.hyper_point_min_deg, cfg.hyper_point_max_deg
  c.viewdir_min_deg, c.viewdir_max_deg = cfg.viewdir_min_deg, cfg.viewdir_max_deg
  c.use_posenc_identity = cfg.use_posenc_identity
  c.use_hyper_sheet = cfg.has_hyper_sheet and cfg.use_hyper_for_sigma
  c.hyper_num_dims = cfg.hyper_num_dims
  c.hyper_sheet_min_deg, c.hyper_sheet_max_deg = cfg.hyper_sheet_min_deg, cfg.hyper_sheet_max_deg
  c.hyper_sheet_depth, c.hyper_sheet_width = cfg.hyper_sheet_depth, cfg.hyper_sheet_width
  c.hyper_sheet_skip = _single_skip(cfg.hyper_sheet_skips, 'HyperSheetMLP.skips')
  c.use_warp, c.warp_embed_dims = cfg.use_warp, cfg.warp_embed_dims
  c.warp_min_deg, c.warp_max_deg = cfg.warp_min_deg, cfg.warp_max_deg
  c.warp_use_posenc_identity = cfg.warp_use_posenc_identity
  c.warp_depth, c.warp_width = cfg.warp_trunk_depth, cfg.warp_trunk_width
  c.warp_skip = _single_skip(cfg.warp_skips, 'SE3Field.skips')
  c.predict_norm, c.norm_input_posenc = cfg.predict_norm, cfg.norm_input_posenc
  c.norm_input_min_deg, c.norm_input_max_deg = cfg.norm_input_min_deg, cfg.norm_input_max_deg
  c.use_x_in_rgb_condition = cfg.use_x_in_rgb_condition
  c.use_mask_in_warp, c.use_mask_in_hyper = cfg.use_mask_in_warp, cfg.use_mask_in_hyper
  c.use_predicted_mask = cfg.use_predicted_mask
  c.use_mask_sharp_weights = cfg.use_mask_sharp_weights
  c.mask_embed_dims = cfg.mask_embed_dims
  c.mask_min_deg, c.mask_max_deg = cfg.mask_min_deg, cfg.mask_max_deg
  c.mask_depth, c.mask_width = cfg.mask_depth, cfg.mask_width
  c.mask_skip = _single_skip(cfg.mask_skips, 'MaskMLP.skips')
  c.mask_output_relu = cfg.mask_output_relu
  c.engine = ENGINES[engine]
  c.precision = PRECISIONS[precision]
  return c


class ndsr_camera(C.Structure):
  """include/nerfds_b200.h: ndsr_camera (hypernerf/camera.py:112-138)."""
  _fields_ = [('orientation', C.c_float * 9), ('position', C.c_float * 3), ('focal_length', C.c_float),
              ('principal_point', C.c_float * 2), ('skew', C.c_float), ('pixel_aspect_ratio', C.c_float),
              ('radial_distortion', C.c_float * 3), ('tangential_distortion', C.c_float * 2),
              ('image_size', C.c_int32 * 2)]


_lib = None

EXPORTS = ['ndsr_create', 'ndsr_destroy', 'ndsr_last_error', 'ndsr_load_params', 'ndsr_render_rays',
           'ndsr_render_rays_host', 'ndsr_render_samples', 'ndsr_sample_along_rays', 'ndsr_sample_pdf',
           'ndsr_volumetric_rendering', 'ndsr_engine_in_use', 'ndsr_kernel_launches', 'ndsr_abi_version',
           'ndsr_struct_sizes', 'ndsr_set_max_chunk', 'ndsr_selftest_tc_dense', 'ndsr_profile_enable',
           'ndsr_profile_read', 'ndsr_camera_rays', 'ndsr_random_uniform', 'ndsr_peer_alloc', 'ndsr_peer_free',
           'ndsr_peer_open', 'ndsr_peer_close', 'ndsr_set_output_mirrors', 'ndsr_render_rays_host_rng',
           'ndsr_random_uniform_range', 'ndsr_tc_issued_macs', 'ndsr_set_early_termination',
           'ndsr_termination_stats']


def load_library() -> C.CDLL:
  """dlopen the in-tree CUDA library; raise loudly if it is absent."""
  global _lib
  if _lib is not None:
    return _lib
  if not os.path.exists(LIB_PATH):
    raise ImportError(
        f'{LIB_PATH} is missing: build it with `python -m nerfds_b200.build` '
        '(nvcc, sm_100a).  There is no CPU fallback.')
  if not os.environ.get('NDSR_LIBRARY'):
    # provenance: the binary must have been built from exactly these sources (build.py records their sha256)
    from . import build as _build
    if _build.recorded_fingerprint() != _build.source_fingerprint():
      raise ImportError(f'{LIB_PATH} was not built from the sources in this tree (lib/BUILD_INFO.json disagrees): '
                        'run `python -m nerfds_b200.build`')
  lib = C.CDLL(LIB_PATH)
  vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
  lib.ndsr_create.argtypes = [C.POINTER(ndsr_config), C.c_int, C.POINTER(vp)]
  lib.ndsr_destroy.argtypes = [vp]
  lib.ndsr_destroy.restype = None
  lib.ndsr_last_error.argtypes = [vp]
  lib.ndsr_last_error.restype = C.c_char_p
  lib.ndsr_load_params.argtypes = [vp, C.POINTER(ndsr_tensor), C.c_int]
  rays_sig = [vp, vp, i64, vp, vp, vp, vp, vp, vp, vp, C.POINTER(ndsr_extra_params),
              C.POINTER(ndsr_outputs), C.POINTER(ndsr_outputs)]
  lib.ndsr_render_rays.argtypes = rays_sig
  lib.ndsr_render_rays_host.argtypes = rays_sig
  lib.ndsr_render_rays_host_rng.argtypes = [vp, vp, i64, vp, vp, vp, vp, vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32),
                                            C.POINTER(ndsr_extra_params), C.POINTER(ndsr_outputs), C.POINTER(ndsr_outputs)]
  lib.ndsr_render_samples.argtypes = [vp, vp, C.c_int, i64, i32, vp, vp, vp, vp, vp, vp, vp,
                                      C.POINTER(ndsr_extra_params), i32, C.POINTER(ndsr_outputs)]
  lib.ndsr_sample_along_rays.argtypes = [vp, vp, i64, i32, C.c_float, C.c_float, i32, vp, vp]
  lib.ndsr_sample_pdf.argtypes = [vp, vp, i64, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp]
  lib.ndsr_volumetric_rendering.argtypes = [vp, vp, i64, i32, vp, vp, vp, vp, i32, i32, C.POINTER(ndsr_outputs)]
  lib.ndsr_engine_in_use.argtypes = [vp]
  lib.ndsr_kernel_launches.argtypes = [vp]
  lib.ndsr_kernel_launches.restype = C.c_int64
  lib.ndsr_struct_sizes.argtypes = [C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
  lib.ndsr_struct_sizes.restype = None
  lib.ndsr_set_max_chunk.argtypes = [vp, i64]
  lib.ndsr_tc_issued_macs.argtypes = [vp, C.POINTER(C.c_double)]
  lib.ndsr_set_early_termination.argtypes = [vp, C.c_float, i32]
  lib.ndsr_termination_stats.argtypes = [vp, vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int]
  lib.ndsr_profile_enable.argtypes = [vp, C.c_int]
  lib.ndsr_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(i64)]
  lib.ndsr_selftest_tc_dense.argtypes = [C.c_int] * 7 + [vp] * 5
  lib.ndsr_camera_rays.argtypes = [C.c_int, vp, C.POINTER(ndsr_camera), vp, vp, vp]
  lib.ndsr_random_uniform.argtypes = [C.c_int, vp, C.POINTER(C.c_uint32), i64, vp]
  lib.ndsr_random_uniform_range.argtypes = [C.c_int, vp, C.POINTER(C.c_uint32), i64, i64, i64, vp]
  lib.ndsr_peer_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(vp), C.c_char_p]
  lib.ndsr_peer_free.argtypes = [C.c_int, vp]
  lib.ndsr_peer_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(vp)]
  lib.ndsr_peer_close.argtypes = [C.c_int, vp]
  lib.ndsr_set_output_mirrors.argtypes = [vp, i32, C.POINTER(i64), vp, C.c_size_t]
  if lib.ndsr_abi_version() != NDSR_ABI_VERSION:
    raise ImportError('libnerfds_b200.so ABI version mismatch; rebuild')
  a, b, c = i32(), i32(), i32()
  lib.ndsr_struct_sizes(C.byref(a), C.byref(b), C.byref(c))
  if (a.value, b.value, c.value) != (C.sizeof(ndsr_config), C.sizeof(ndsr_extra_params), C.sizeof(ndsr_outputs)):
    raise ImportError('ctypes struct mirrors disagree with libnerfds_b200.so')
  _lib = lib
  return lib
