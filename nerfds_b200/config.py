"""Hyper-parameters of the NeRF-DS ray-marching path.

One flat dataclass holding every attribute of the reference's gin-configured
classes that the hot path reads:

  * ``NerfModel`` attributes            -- hypernerf/models.py:116-229
  * ``SE3Field`` attributes             -- hypernerf/warping.py:139-157
  * ``HyperSheetMLP`` attributes        -- hypernerf/modules.py:354-365
  * ``MaskMLP`` attributes              -- hypernerf/modules.py:396-407
  * ``GLOEmbed.num_dims``               -- hypernerf/modules.py:326-328

Field names are the reference's attribute names (prefixed with ``warp_`` /
``hyper_sheet_`` / ``mask_`` for the sub-modules) so a gin file maps onto it
mechanically (see ``from_gin_bindings``).  Features of ``NerfModel`` that no
shipped gin file enables and that are not on the path (bones, hyper_c,
reflected radiance, nerf/appearance embeddings, ...) are represented only so
that turning them on raises ``NotImplementedError`` instead of being silently
ignored.
"""
from __future__ import annotations

import dataclasses
from dataclasses import dataclass
from typing import Optional, Tuple


@dataclass(frozen=True)
class NerfDSConfig:
  # scene bounds (NerfModel.near / far are gin.REQUIRED, models.py:117-118)
  near: float = 0.1
  far: float = 2.5
  # number of distinct warp ids (len of embeddings_dict['warp'], models.py:236)
  num_warp_embeds: int = 100

  # ---- NeRF template MLP (models.py:121-127, modules.py:101-120)
  use_viewdirs: bool = True
  nerf_trunk_depth: int = 8
  nerf_trunk_width: int = 256
  nerf_rgb_branch_depth: int = 1
  nerf_rgb_branch_width: int = 128
  nerf_skips: Tuple[int, ...] = (4,)
  alpha_channels: int = 1
  rgb_channels: int = 3

  # ---- sampling / rendering (models.py:130-135)
  num_coarse_samples: int = 64
  num_fine_samples: int = 64
  use_stratified_sampling: bool = True
  use_white_background: bool = False
  use_linear_disparity: bool = False
  use_sample_at_infinity: bool = True

  # ---- positional encodings (models.py:137-143; nerf_ds.gin:22-30)
  spatial_point_min_deg: int = 0
  spatial_point_max_deg: int = 8
  hyper_point_min_deg: int = 0
  hyper_point_max_deg: int = 1
  viewdir_min_deg: int = 0
  viewdir_max_deg: int = 4
  use_posenc_identity: bool = False

  # ---- hyper slicing (models.py:158-167; nerf_ds.gin:35-44)
  hyper_slice_method: str = 'bendy_sheet'  # 'none' | 'bendy_sheet'
  use_hyper: bool = True
  hyper_use_warp_embed: bool = True
  use_hyper_for_sigma: bool = True
  hyper_num_dims: int = 2            # HyperSheetMLP.output_channels
  hyper_sheet_min_deg: int = 0
  hyper_sheet_max_deg: int = 6
  hyper_sheet_depth: int = 6
  hyper_sheet_width: int = 64
  hyper_sheet_skips: Tuple[int, ...] = (4,)

  # ---- SE(3) warp field (models.py:170-174; warping.py:139-157)
  use_warp: bool = True
  warp_embed_dims: int = 8           # warp/GLOEmbed.num_dims
  warp_min_deg: int = 0
  warp_max_deg: int = 4
  warp_use_posenc_identity: bool = False
  warp_trunk_depth: int = 6
  warp_trunk_width: int = 128
  warp_skips: Tuple[int, ...] = (4,)

  # ---- surface normal / specular branch (models.py:177-186)
  predict_norm: bool = True
  norm_supervision_type: str = 'warped'
  stop_norm_gradient: bool = True
  norm_input_posenc: bool = True
  norm_input_min_deg: int = 0
  norm_input_max_deg: int = 4
  use_x_in_rgb_condition: bool = True
  window_x_in_rgb_condition: bool = False

  # ---- predicted moving-object mask (models.py:202-214; nerf_ds.gin:105-118)
  use_mask_in_warp: bool = True
  use_mask_in_hyper: bool = True
  use_predicted_mask: bool = True
  use_mask_embed: bool = True
  use_3d_mask: bool = True
  use_mask_sharp_weights: bool = True
  mask_embed_dims: int = 8
  mask_min_deg: int = 0
  mask_max_deg: int = 6
  mask_depth: int = 8
  mask_width: int = 128
  mask_skips: Tuple[int, ...] = (4,)
  mask_output_relu: bool = True      # MaskMLP.output_activation = @jax.nn.relu

  # ---- features of NerfModel that are off in every shipped gin and NOT built
  use_nerf_embed: bool = False
  use_alpha_condition: bool = False
  use_rgb_condition: bool = False
  use_viewdirs_in_hyper: bool = False
  use_delta_x_in_rgb_condition: bool = False
  use_hyper_c: bool = False
  use_ref_radiance: bool = False
  use_mask_in_rgb: bool = False
  use_coarse_depth_for_mask: bool = False
  clamp_predicted_mask: bool = False
  use_mask_scaled_weights: bool = False
  use_rgb_sharp_weights: bool = False
  use_hyper_for_rgb: bool = False
  use_bone: bool = False
  use_norm_voxel: bool = False
  norm_type: str = 'none'
  noise_std: Optional[float] = None

  # ------------------------------------------------------------------ helpers
  def replace(self, **kw) -> 'NerfDSConfig':
    return dataclasses.replace(self, **kw)

  @property
  def has_hyper(self) -> bool:          # models.py:255-258
    return self.hyper_slice_method != 'none'

  @property
  def has_hyper_sheet(self) -> bool:
    return self.use_hyper and self.hyper_slice_method == 'bendy_sheet'

  @property
  def alpha_out_channels(self) -> int:  # modules.py:141-145
    return self.alpha_channels + (3 if self.predict_norm else 0)

  def posenc_dim(self, channels, min_deg, max_deg, identity) -> int:
    return 2 * (max_deg - min_deg) * channels + (channels if identity else 0)

  @property
  def warp_in_dim(self) -> int:         # warping.py:209-214
    return (self.posenc_dim(3, self.warp_min_deg, self.warp_max_deg,
                            self.warp_use_posenc_identity)
            + self.warp_embed_dims + (1 if self.use_mask_in_warp else 0))

  @property
  def hyper_sheet_in_dim(self) -> int:  # modules.py:373-376
    return (self.posenc_dim(3, self.hyper_sheet_min_deg,
                            self.hyper_sheet_max_deg, False)
            + self.warp_embed_dims + (1 if self.use_mask_in_hyper else 0))

  @property
  def mask_in_dim(self) -> int:         # modules.py:415-418
    return (self.posenc_dim(3, self.mask_min_deg, self.mask_max_deg, False)
            + (self.mask_embed_dims if self.use_mask_embed else 0))

  @property
  def trunk_in_dim(self) -> int:        # models.py:502-516
    d = self.posenc_dim(3, self.spatial_point_min_deg,
                        self.spatial_point_max_deg, self.use_posenc_identity)
    if self.has_hyper_sheet and self.use_hyper_for_sigma:
      d += self.posenc_dim(self.hyper_num_dims, self.hyper_point_min_deg,
                           self.hyper_point_max_deg, False)
    return d

  @property
  def viewdir_feat_dim(self) -> int:    # models.py:400-406
    if not self.use_viewdirs:
      return 0
    return self.posenc_dim(3, self.viewdir_min_deg, self.viewdir_max_deg,
                           self.use_posenc_identity)

  @property
  def norm_feat_dim(self) -> int:       # models.py:1141-1150
    if self.norm_input_posenc:
      return self.posenc_dim(3, self.norm_input_min_deg,
                             self.norm_input_max_deg, self.use_posenc_identity)
    return 3

  def rgb_in_dim(self, with_norm_input: bool) -> int:
    """Width of the rgb branch input (modules.py:288-313, App. C-1 quirk).

    ``[bottleneck | viewdir_feat] | x_for_rgb (= trunk output) | norm_feat``;
    when there is no rgb condition the reference feeds the trunk output
    instead of the bottleneck (modules.py:296-300).
    """
    d = self.nerf_trunk_width + self.viewdir_feat_dim
    if self.use_x_in_rgb_condition:
      d += self.nerf_trunk_width
    if with_norm_input:
      d += self.norm_feat_dim
    return d

  def validate(self) -> None:
    """Raise for every NerfModel feature the B200 path does not build.

    Mirrors the reference's own error behaviour where it has one
    (ValueError models.py:328, RuntimeError 315, NotImplementedError
    561/744/1131) and fences off the rest explicitly.
    """
    if self.use_nerf_embed and not (self.use_rgb_condition
                                    or self.use_alpha_condition):
      raise ValueError('Template metadata is enabled but none of the condition'
                       'branches are.')          # models.py:325-329
    if self.hyper_slice_method not in ('none', 'bendy_sheet'):
      if self.hyper_slice_method == 'axis_aligned_plane':
        raise NotImplementedError('axis_aligned_plane hyper slicing')
      raise RuntimeError(
          f'Unknown hyper slice method {self.hyper_slice_method}.')
    if self.use_viewdirs_in_hyper:
      raise NotImplementedError  # models.py:743-744
    if self.norm_supervision_type != 'warped':
      raise NotImplementedError(
          f'norm_supervision_type={self.norm_supervision_type!r}; only '
          "'warped' (nerf_ds.gin) is built")
    off = ('use_nerf_embed', 'use_alpha_condition', 'use_rgb_condition',
           'use_delta_x_in_rgb_condition', 'use_hyper_c', 'use_ref_radiance',
           'use_mask_in_rgb', 'use_coarse_depth_for_mask',
           'clamp_predicted_mask', 'use_mask_scaled_weights',
           'use_rgb_sharp_weights', 'use_hyper_for_rgb', 'use_bone',
           'use_norm_voxel', 'window_x_in_rgb_condition')
    for name in off:
      if getattr(self, name):
        raise NotImplementedError(f'NerfModel.{name}=True is outside the '
                                  'ray-marching path built here')
    if self.norm_type not in (None, 'none'):
      raise NotImplementedError('norm layers')
    if self.noise_std:
      raise NotImplementedError('noise_std regularisation')
    if not self.hyper_use_warp_embed and self.has_hyper_sheet:
      raise NotImplementedError('separate hyper embedding')
    if not self.use_warp and self.has_hyper_sheet:
      # reference: hyper_embed := warp_embed = None -> crash (models.py:910)
      raise NotImplementedError('bendy_sheet without a warp embedding')
    if not self.use_warp and self.use_predicted_mask:
      raise NotImplementedError('predicted mask without a warp embedding')
    if self.use_predicted_mask and not self.use_mask_embed:
      raise NotImplementedError('MaskMLP without embedding')
    if self.alpha_channels != 1 or self.rgb_channels != 3:
      raise NotImplementedError('alpha_channels/rgb_channels other than 1/3')
    if self.num_coarse_samples < 3:
      raise ValueError('num_coarse_samples must be >= 3')
    if self.num_fine_samples <= 0:
      # models.py:1555-1563 deletes out['fine'][...] unconditionally.
      raise KeyError('fine')


# ----------------------------------------------------------------- presets
def nerf_ds_config(**overrides) -> NerfDSConfig:
  """configs/nerf_ds.gin (+ defaults.gin) resolved -- SURVEY.md App. A."""
  return NerfDSConfig().replace(**overrides)


def tiny_config(**overrides) -> NerfDSConfig:
  """BASELINE.json configs[0]: no warp field, tiny NerfMLP (width 32).

  64 coarse (+64 fine: the reference cannot run with num_fine_samples=0,
  models.py:1555-1563), no hyper slicing, no normal head, no predicted mask.
  """
  base = NerfDSConfig(
      near=0.2, far=2.0, num_warp_embeds=1,
      nerf_trunk_width=32, nerf_rgb_branch_width=32,
      num_coarse_samples=64, num_fine_samples=64,
      hyper_slice_method='none', use_warp=False,
      predict_norm=False, use_x_in_rgb_condition=False,
      use_mask_in_warp=False, use_mask_in_hyper=False,
      use_predicted_mask=False, use_3d_mask=False,
      use_mask_sharp_weights=False)
  return base.replace(**overrides)


_GIN_MAP = {
    # gin binding -> dataclass field (configs/nerf_ds.gin, defaults.gin)
    'NerfModel.num_coarse_samples': 'num_coarse_samples',
    'NerfModel.num_fine_samples': 'num_fine_samples',
    'NerfModel.use_viewdirs': 'use_viewdirs',
    'NerfModel.use_stratified_sampling': 'use_stratified_sampling',
    'NerfModel.use_posenc_identity': 'use_posenc_identity',
    'NerfModel.norm_type': 'norm_type',
    'NerfModel.hyper_slice_method': 'hyper_slice_method',
    'NerfModel.hyper_use_warp_embed': 'hyper_use_warp_embed',
    'NerfModel.use_warp': 'use_warp',
    'NerfModel.use_rgb_condition': 'use_rgb_condition',
    'NerfModel.predict_norm': 'predict_norm',
    'NerfModel.norm_supervision_type': 'norm_supervision_type',
    'NerfModel.use_viewdirs_in_hyper': 'use_viewdirs_in_hyper',
    'NerfModel.use_x_in_rgb_condition': 'use_x_in_rgb_condition',
    'NerfModel.use_hyper_c': 'use_hyper_c',
    'NerfModel.use_mask_in_warp': 'use_mask_in_warp',
    'NerfModel.use_mask_in_hyper': 'use_mask_in_hyper',
    'NerfModel.use_mask_in_rgb': 'use_mask_in_rgb',
    'NerfModel.use_predicted_mask': 'use_predicted_mask',
    'NerfModel.use_3d_mask': 'use_3d_mask',
    'NerfModel.use_mask_sharp_weights': 'use_mask_sharp_weights',
    'NerfModel.nerf_trunk_depth': 'nerf_trunk_depth',
    'NerfModel.nerf_trunk_width': 'nerf_trunk_width',
    'NerfModel.nerf_rgb_branch_depth': 'nerf_rgb_branch_depth',
    'NerfModel.nerf_rgb_branch_width': 'nerf_rgb_branch_width',
    'NerfModel.spatial_point_min_deg': 'spatial_point_min_deg',
    'NerfModel.spatial_point_max_deg': 'spatial_point_max_deg',
    'NerfModel.hyper_point_min_deg': 'hyper_point_min_deg',
    'NerfModel.hyper_point_max_deg': 'hyper_point_max_deg',
    'spatial_point_min_deg': 'spatial_point_min_deg',
    'spatial_point_max_deg': 'spatial_point_max_deg',
    'hyper_point_min_deg': 'hyper_point_min_deg',
    'hyper_point_max_deg': 'hyper_point_max_deg',
    'hyper_num_dims': 'hyper_num_dims',
    'hyper_sheet_min_deg': 'hyper_sheet_min_deg',
    'hyper_sheet_max_deg': 'hyper_sheet_max_deg',
    'HyperSheetMLP.min_deg': 'hyper_sheet_min_deg',
    'HyperSheetMLP.max_deg': 'hyper_sheet_max_deg',
    'HyperSheetMLP.output_channels': 'hyper_num_dims',
    'HyperSheetMLP.depth': 'hyper_sheet_depth',
    'HyperSheetMLP.width': 'hyper_sheet_width',
    'warp_min_deg': 'warp_min_deg',
    'warp_max_deg': 'warp_max_deg',
    'SE3Field.min_deg': 'warp_min_deg',
    'SE3Field.max_deg': 'warp_max_deg',
    'SE3Field.use_posenc_identity': 'warp_use_posenc_identity',
    'SE3Field.trunk_depth': 'warp_trunk_depth',
    'SE3Field.trunk_width': 'warp_trunk_width',
    'warp/GLOEmbed.num_dims': 'warp_embed_dims',
    'MaskMLP.depth': 'mask_depth',
    'MaskMLP.width': 'mask_width',
    'MaskMLP.min_deg': 'mask_min_deg',
    'MaskMLP.max_deg': 'mask_max_deg',
}


def from_gin_bindings(bindings: dict, base: Optional[NerfDSConfig] = None
                      ) -> NerfDSConfig:
  """Apply ``{'NerfModel.num_coarse_samples': 128, ...}`` style bindings.

  A dict-of-literals stand-in for ``gin.parse_config`` (gin is not installed
  here); keys are the reference's gin names.  Unknown keys that do not touch
  the path (TrainConfig.*, EvalConfig.*, ...) are ignored; unknown
  ``NerfModel.*`` keys raise so nothing is dropped silently.
  """
  cfg = base or NerfDSConfig()
  kw = {}
  for key, value in bindings.items():
    if key in _GIN_MAP:
      kw[_GIN_MAP[key]] = value
    elif key == 'MaskMLP.output_activation':
      kw['mask_output_relu'] = value in ('@jax.nn.relu', '@flax.nn.relu', '@nn.relu', '@relu', 'relu')
      if not kw['mask_output_relu'] and value not in (None, 'identity'):
        raise NotImplementedError(f'MaskMLP.output_activation = {value}: only relu / identity are built')
    elif key.split('.')[0] in ('NerfModel',):
      field = key.split('.', 1)[1]
      if field in NerfDSConfig.__dataclass_fields__:
        kw[field] = value
      else:
        raise KeyError(f'unsupported gin binding {key}')
  return cfg.replace(**kw)
