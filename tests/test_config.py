"""Config surface: gin names, derived dims, error behaviour."""
import pytest

from nerfds_b200.config import NerfDSConfig, from_gin_bindings, nerf_ds_config, tiny_config
from nerfds_b200.params import flatten_params, init_params, param_count, unflatten_params


def test_nerf_ds_dims_match_survey_appendix_a():
  c = nerf_ds_config()
  assert c.mask_in_dim == 44 and c.warp_in_dim == 33 and c.hyper_sheet_in_dim == 45
  assert c.trunk_in_dim == 52 and c.viewdir_feat_dim == 24 and c.norm_feat_dim == 24
  assert c.rgb_in_dim(True) == 560 and c.alpha_out_channels == 4
  P = init_params(c.replace(num_warp_embeds=100), 0)
  shapes = {n: v.shape for n, v in flatten_params(P)}
  assert shapes['warp_field/trunk/hidden_4/kernel'] == (161, 128)
  assert shapes['hyper_sheet_mlp/MLP_0/hidden_4/kernel'] == (109, 64)
  assert shapes['mask_mlp/MLP_0/hidden_4/kernel'] == (172, 128)
  assert shapes['nerf_mlps_fine/trunk_mlp/hidden_4/kernel'] == (308, 256)
  assert shapes['nerf_mlps_coarse/alpha_mlp/logit/kernel'] == (256, 4)
  assert shapes['nerf_mlps_coarse/rgb_mlp/hidden_0/kernel'] == (560, 128)
  assert 1.49e6 < param_count(P) < 1.51e6          # "~1.50 M weights" (App. A.9)


def test_macs_per_eval_match_survey_appendix_d():
  c = nerf_ds_config()
  se3 = 33 * 128 + 3 * 128 ** 2 + 161 * 128 + 128 ** 2 + 2 * 128 * 3
  hyper = 45 * 64 + 3 * 64 ** 2 + 109 * 64 + 64 ** 2 + 64 * 2
  mask = 44 * 128 + 3 * 128 ** 2 + 172 * 128 + 3 * 128 ** 2 + 128
  trunk = 52 * 256 + 3 * 256 ** 2 + 308 * 256 + 3 * 256 ** 2
  assert (se3, hyper, mask, trunk) == (91136, 26368, 126080, 485376)
  assert se3 + hyper + mask + trunk + 1024 + 65536 + 72064 == 867584


def test_gin_bindings():
  c = from_gin_bindings({'NerfModel.num_coarse_samples': 128, 'NerfModel.num_fine_samples': 128,
                         'MaskMLP.depth': 8, 'MaskMLP.output_activation': '@jax.nn.relu',
                         'TrainConfig.batch_size': 512, 'warp/GLOEmbed.num_dims': 8})
  assert c.num_coarse_samples == 128 and c.num_fine_samples == 128 and c.mask_output_relu
  with pytest.raises(KeyError):
    from_gin_bindings({'NerfModel.no_such_attr': 1})


def test_reference_error_behaviour():
  with pytest.raises(ValueError):                      # models.py:325-329
    NerfDSConfig(use_nerf_embed=True).validate()
  with pytest.raises(RuntimeError):                    # models.py:314-316
    NerfDSConfig(hyper_slice_method='bogus').validate()
  with pytest.raises(NotImplementedError):             # models.py:743-744
    NerfDSConfig(use_viewdirs_in_hyper=True).validate()
  with pytest.raises(KeyError):                        # models.py:1555-1563 (App. C-4)
    NerfDSConfig(num_fine_samples=0).validate()
  for flag in ('use_hyper_c', 'use_bone', 'use_mask_in_rgb', 'use_ref_radiance'):
    with pytest.raises(NotImplementedError):
      NerfDSConfig(**{flag: True}).validate()
  tiny_config().validate()


def test_param_tree_roundtrip():
  P = init_params(tiny_config(), 1)
  flat = dict(flatten_params(P))
  Q = unflatten_params(flat)
  assert [n for n, _ in flatten_params(Q)] == list(flat)
  assert 'warp_field' not in P and 'mask_mlp' not in P
