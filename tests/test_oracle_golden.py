"""Pins the CPU oracle against golden vectors produced by the REFERENCE'S OWN SOURCE.

tests/golden/reference_shim.npz was written by tools/make_golden.py, which imports the unmodified
hypernerf/model_utils.py and hypernerf/rigid_body.py from /root/reference under a numpy stand-in for jax.numpy (jax /
flax are not installable here) and runs them on seeded float32 inputs.  These tests need only the committed .npz.

Tolerances: identical float32 formulas evaluated by numpy (golden) and torch (oracle) -- a few ulp for transcendental
functions and accumulation order; selections (inverse-CDF samples given u, median-depth index, sort) are exact.
"""
import os

import numpy as np
import pytest
import torch

from oracle import nerfds_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_shim.npz')


@pytest.fixture(scope='module')
def G():
  return dict(np.load(GOLDEN))


def T(a):
  return torch.from_numpy(np.asarray(a))


def close(a, b, rtol=2e-6, atol=2e-7):
  a = a.numpy() if torch.is_tensor(a) else np.asarray(a)
  np.testing.assert_allclose(a, b, rtol=rtol, atol=atol, equal_nan=True)


def test_posenc_and_windows(G):
  x = T(G['posenc_x'])
  for tag in 'abcd':
    lo, hi, ident, alpha = G[f'posenc_{tag}_args']
    out = O.posenc(x, int(lo), int(hi), bool(ident), None if np.isnan(alpha) else float(alpha))
    assert out.shape == G[f'posenc_{tag}'].shape
    # |2^k x| up to 2^7 * 1.5: sin of a float32 argument, a few ulp of the ARGUMENT
    close(out, G[f'posenc_{tag}'], rtol=0, atol=2e-6)
  for tag in 'abc':
    lo, hi, alpha = G[f'window_{tag}_args']
    close(O.posenc_window(int(lo), int(hi), float(alpha)), G[f'window_{tag}'], rtol=0, atol=2e-7)
  close(O.normalize_vector(T(G['normalize_in'])), G['normalize_out'])


def test_sample_along_rays(G):
  o, d, t = T(G['sar_origins']), T(G['sar_dirs']), T(G['sar_t_rand'])
  Sc = t.shape[1]
  for tag, (strat, disp) in {'strat': (True, False), 'det': (False, False), 'disp': (True, True)}.items():
    z, pts = O.sample_along_rays(t, o, d, Sc, 0.1, 2.5, strat, disp)
    close(z, G[f'sar_{tag}_z'], rtol=1e-6)
    close(pts, G[f'sar_{tag}_points'], rtol=2e-6, atol=1e-6)


def test_volumetric_rendering(G):
  z, sigma, rgb, d = T(G['vr_z']), T(G['vr_sigma']), T(G['vr_rgb']), T(G['vr_dirs'])
  for tag, (white, inf) in {'inf': (False, True), 'white': (True, True), 'noinf': (False, False)}.items():
    r = O.volumetric_rendering(rgb, sigma, z, d, white, inf)
    for k in ('weights', 'alpha', 'accum_prod', 'rgb', 'depth', 'acc'):
      close(r[k], G[f'vr_{tag}_{k}'], rtol=1e-5, atol=1e-6)
    # same cumulative sum -> same first sample past half the mass
    np.testing.assert_array_equal(r['med_depth'].numpy(), G[f'vr_{tag}_med_depth'])
  w = O.cal_weights(sigma, z, d)
  close(w, G['cal_weights'], rtol=1e-5, atol=1e-6)
  wg = T(G['cal_weights'])
  np.testing.assert_array_equal(O.compute_depth_index(wg).numpy(), G['depth_index'])
  np.testing.assert_array_equal(O.compute_depth_map(wg, z).numpy(), G['depth_map'])
  np.testing.assert_array_equal(O.compute_opaqueness_mask(wg).numpy().astype(np.float32), G['opaqueness_mask'])
  # the row-gather quirk of sharpen_weights comes out of the reference source itself here
  close(O.sharpen_weights(wg, z, std=0.1), G['sharpen_weights'], rtol=2e-5, atol=1e-6)


def test_inverse_cdf_resampling(G):
  bins, w, u = T(G['pdf_bins']), T(G['pdf_weights']), T(G['pdf_u'])
  zs = O.piecewise_constant_pdf(u, bins, w)
  close(zs, G['pdf_samples'], rtol=2e-6, atol=1e-6)
  Sf = u.shape[1]
  u_det = O.linspace01(Sf, torch.float32)[None, :].expand(u.shape[0], Sf)
  zd = O.piecewise_constant_pdf(u_det, bins, w).numpy()
  close(zd[:, :-1], G['pdf_samples_det'][:, :-1], rtol=2e-6, atol=1e-6)
  # u == 1.0 exactly sits on cdf[-1], whose last ulp depends on the summation order of weights.sum() (pairwise in
  # numpy, unspecified in XLA, sequential here); over a near-empty last bin (pdf ~1e-5) that ulp moves the lerp
  # parameter by ~1 %: both land at the far end of the last bin, within 2 % of its width
  b0, b1 = G['pdf_bins'][:, -2], G['pdf_bins'][:, -1]
  for z_last in (zd[:, -1], G['pdf_samples_det'][:, -1]):
    assert ((z_last >= b0) & (z_last <= b1 + 0.02 * (b1 - b0) + 1e-6)).all()
  assert np.abs(zd[:, -1] - G['pdf_samples_det'][:, -1]).max() <= 0.02 * (b1 - b0).max()
  o, d, z = T(G['sar_origins']), T(G['sar_dirs']), T(G['vr_z'])
  zf, pf = O.sample_pdf(u, bins, w, o, d, z)
  close(zf, G['sample_pdf_z'], rtol=2e-6, atol=1e-6)
  close(pf, G['sample_pdf_points'], rtol=4e-6, atol=2e-6)
  # the merged depths are sorted, and the coarse depths are among them bit for bit
  assert (np.diff(zf.numpy(), axis=-1) >= 0).all()
  for b in range(z.shape[0]):
    assert np.isin(G['vr_z'][b], zf.numpy()[b]).all()


def test_se3_exponential(G):
  w, v, th = T(G['se3_w']), T(G['se3_v']), T(G['se3_theta'])
  np.testing.assert_array_equal(O.skew(w).numpy(), G['skew'])
  close(O.exp_so3(w, th), G['exp_so3'], rtol=2e-6, atol=5e-7)
  R, p = O.exp_se3_rp(torch.cat([w, v], -1), th)
  close(R, G['exp_se3'][:, :3, :3], rtol=2e-6, atol=5e-7)
  close(p, G['exp_se3'][:, :3, 3], rtol=4e-6, atol=1e-6)
  np.testing.assert_array_equal(G['exp_se3'][:, 3], np.tile(np.float32([0, 0, 0, 1]), (w.shape[0], 1)))
  # homogeneous helpers used by warping.py:231-232
  x = T(G['hom_in'])
  np.testing.assert_array_equal(torch.cat([x, torch.ones_like(x[..., :1])], -1).numpy(), G['to_homogenous'])
  close(x, G['from_homogenous'], rtol=1e-6)


# ---------------------------------------------------------------------------------------------------------------
# Flax-module level: the reference's modules.py / warping.py classes were run (tools/make_golden.py, flax.linen
# stand-in) on a reduced nerf_ds.gin configuration with a parameter pytree from nerfds_b200.params.init_params.
# ---------------------------------------------------------------------------------------------------------------
def _module_case(G):
  from nerfds_b200.config import nerf_ds_config
  from nerfds_b200.params import unflatten_params
  small = {str(k): int(v) for k, v in zip(G['mod_cfg_keys'], G['mod_cfg_vals'])}
  cfg = nerf_ds_config(**small)
  P = unflatten_params({k[2:]: v for k, v in G.items() if k.startswith('P/')})
  ep = dict(zip(('warp_alpha', 'hyper_sheet_alpha', 'nerf_alpha', 'hyper_alpha', 'norm_input_alpha'),
                (float(v) for v in G['mod_extra'])))
  return cfg, P, ep, O.OracleNerfModel(cfg, P)


def test_modules_embeddings_mask_and_hyper_sheet(G):
  cfg, P, ep, om = _module_case(G)
  ids = T(G['mod_ids'].astype(np.int64)).reshape(-1)
  wemb = om.P['warp_embed']['embed']['embedding'][ids]
  memb = om.P['mask_embed']['embed']['embedding'][ids]
  np.testing.assert_array_equal(wemb.numpy(), G['mod_warp_embed'])
  np.testing.assert_array_equal(memb.numpy(), G['mod_mask_embed'])
  x = T(G['mod_points'])
  # the oracle's mask branch (render_samples, models.py:955-975)
  pe = O.posenc(x, cfg.mask_min_deg, cfg.mask_max_deg, alpha=ep['warp_alpha'])
  pm = O.mlp_apply(om.P['mask_mlp']['MLP_0'], torch.cat([pe, memb], -1), cfg.mask_depth, cfg.mask_skips,
                   out_relu=cfg.mask_output_relu)
  close(pm, G['mod_mask_mlp'], rtol=2e-5, atol=2e-6)


def test_modules_sigma_field_pieces(G):
  """_sigma_field (the oracle's cal_single_pt_sigma) against the reference classes it restates: SE3Field.warp,
  HyperSheetMLP, NerfMLP.query_bottleneck / query_sigma."""
  cfg, P, ep, om = _module_case(G)
  x, mask = T(G['mod_points']), T(G['mod_mask'])
  wemb = T(G['mod_warp_embed'])
  with torch.no_grad():
    _, aux = om._sigma_field('fine', x, wemb, mask, ep)
  close(aux['screw_axis'], G['mod_se3_screw'], rtol=5e-5, atol=5e-6)
  close(aux['warped_points'][:, :3], G['mod_se3_points'], rtol=2e-5, atol=5e-6)
  close(aux['warped_points'][:, 3:], G['mod_hyper_sheet'], rtol=2e-5, atol=2e-6)
  # map_vectors (models.py:581-607): rotate a vector forward / inverse, and the translation field
  v = T(G['mod_vec'])
  close(O.se3_apply(aux['R'], aux['p'], v, rotation_only=True), G['mod_se3_vec_fwd'], rtol=2e-5, atol=5e-6)
  close(O.se3_apply(aux['R'], aux['p'], v, rotation_only=True, inverse=True), G['mod_se3_vec_inv'], rtol=2e-5, atol=5e-6)
  close(O.se3_apply(aux['R'], aux['p'], v * 0), G['mod_se3_vec_trans'], rtol=2e-5, atol=5e-6)


def test_modules_nerf_mlp(G):
  """The template MLP as the oracle evaluates it (trunk, bottleneck, sigma/normal head, rgb-branch input order
  [bottleneck | viewdir feats | trunk_out | norm feats]) against NerfMLP.query_bottleneck / query_sigma / query_rgb."""
  cfg, P, ep, om = _module_case(G)
  pn = om.P['nerf_mlps_fine']
  feat, vfeat, nfeat = T(G['mod_trunk_in']), T(G['mod_view_feat']), T(G['mod_norm_feat'])
  trunk_out = O.mlp_apply(pn['trunk_mlp'], feat, cfg.nerf_trunk_depth, cfg.nerf_skips)
  bott = trunk_out @ pn['bottleneck']['kernel'] + pn['bottleneck']['bias']
  alpha_out = trunk_out @ pn['alpha_mlp']['logit']['kernel'] + pn['alpha_mlp']['logit']['bias']
  close(trunk_out, G['mod_trunk_out'], rtol=2e-5, atol=2e-6)
  close(bott, G['mod_bottleneck'], rtol=2e-5, atol=2e-6)
  close(alpha_out[:, :1], G['mod_alpha'], rtol=2e-5, atol=1e-5)
  close(alpha_out[:, 1:4], G['mod_norm'], rtol=2e-5, atol=2e-6)
  rgb_in = torch.cat([bott, vfeat, trunk_out, nfeat], -1)
  rgb_raw = O.mlp_apply(pn['rgb_mlp'], rgb_in, cfg.nerf_rgb_branch_depth, ())
  close(rgb_raw, G['mod_rgb_raw'], rtol=2e-5, atol=5e-6)


# ---------------------------------------------------------------------------------------------------------------
# Whole forward: the reference's NerfModel.__call__ (models.py:1419-1565) was run under the stand-in with nerf_ds.gin's
# bindings, the product's parameter pytree and injected draws.  Every key of both level dicts is compared with
# OracleNerfModel.apply, incl. the gradient-derived `target_norm` (the stand-in's value_and_grad is a float64 central
# difference of the reference's own per-point function).
# ---------------------------------------------------------------------------------------------------------------
def test_whole_forward_matches_reference_nerf_model(G):
  from nerfds_b200.config import nerf_ds_config
  from nerfds_b200.params import unflatten_params
  small = {str(k): int(v) for k, v in zip(G['model_cfg_keys'], G['model_cfg_vals'])}
  cfg = nerf_ds_config(**small)
  P = unflatten_params({k[3:]: v for k, v in G.items() if k.startswith('MP/')})
  ep = {str(k): float(v) for k, v in zip(G['model_extra_keys'], G['model_extra_vals'])}
  rays = {'origins': G['model_origins'], 'directions': G['model_dirs'], 'metadata': {'warp': G['model_warp']},
          'mask': G['model_gt_mask']}
  om = O.OracleNerfModel(cfg, P)
  out = O.to_numpy(om.apply(rays, ep, G['model_t_rand'], G['model_u'], return_points=True, return_weights=True,
                            use_predicted_norm=True, compute_sigma_gradient=True, keep_internal=True,
                            mask_ratio=float(G['model_mask_ratio']), sharp_weights_std=float(G['model_sharp_std'])))
  o3, d3 = G['model_origins'], G['model_dirs']
  depth_of = lambda pts: (((pts - o3[:, None]) * d3[:, None]).sum(-1) / (d3 ** 2).sum(-1)[:, None]).astype(np.float32)
  # (1) the resampling step on the REFERENCE's coarse weights reproduces the reference's fine depths
  zc, zf = depth_of(G['model_coarse_points']), depth_of(G['model_fine_points'])
  wc = G['model_coarse_weights']
  z_re, _ = O.sample_pdf(T(G['model_u']), T(.5 * (zc[..., 1:] + zc[..., :-1])), T(wc[..., 1:-1]), T(o3), T(d3), T(zc))
  np.testing.assert_allclose(z_re.numpy(), zf, rtol=2e-5, atol=2e-5)
  # (2) the fine level is evaluated on the reference's own fine samples: end to end, 1e-7 differences in the coarse
  # weights move resampled depths inside near-empty bins by 1e-3 (the inverse CDF is ill-conditioned there)
  fine_on_ref = O.to_numpy(om.render_samples(
      'fine', T(G['model_fine_points']), T(zf), T(d3), T(d3), rays['metadata'], ep, G['model_gt_mask'],
      use_sample_at_infinity=cfg.use_sample_at_infinity, use_predicted_norm=True, compute_sigma_gradient=True,
      mask_ratio=float(G['model_mask_ratio']), sharp_weights_std=float(G['model_sharp_std'])))
  assert np.abs(out['fine']['rgb'] - G['model_fine_rgb']).max() <= 2e-2       # end to end: loose, see (2)
  out = {'coarse': out['coarse'], 'fine': fine_on_ref}
  checked = 0
  for lvl in ('coarse', 'fine'):
    gold = {k[len(f'model_{lvl}_'):]: v for k, v in G.items() if k.startswith(f'model_{lvl}_')}
    assert {'rgb', 'depth', 'med_depth', 'acc', 'weights', 'sigma', 'warped_points', 'predicted_mask', 'predicted_norm',
            'ray_norm', 'ray_rotation_field', 'ray_translation_field', 'ray_delta_x', 'ray_hyper_points',
            'ray_predicted_mask', 'med_points', 'sharp_weights', 'back_facing', 'points', 'target_norm'} <= set(gold)
    for k, g in gold.items():
      assert k in out[lvl], (lvl, k, sorted(out[lvl]))
      o = np.asarray(out[lvl][k]).reshape(g.shape)
      if k == 'sharp_weights':
        # Each row is weights x Gaussian(z - z[row of its argmax sample index]) normalised by its sum.  Rows whose
        # sum lands in (or near) float32's denormal range are 0/0 or ratios of few-bit numbers, decided by the exp
        # implementation's denormal handling: only rows with a healthy normaliser are compared (NaN-ness included).
        z = (((gold['points'] - G['model_origins'][:, None]) * G['model_dirs'][:, None]).sum(-1) /
             (G['model_dirs'] ** 2).sum(-1)[:, None]).astype(np.float64)
        w = gold['weights'].astype(np.float64)
        rows = np.minimum(w.argmax(1), z.shape[0] - 1)
        std = float(G['model_sharp_std'])
        tot = (w * np.exp(-0.5 * ((z - z[rows]) / std) ** 2)).sum(1)
        healthy = tot > 1e-25
        assert healthy.sum() >= 3
        # (atol: alpha = 1 - exp(-sigma delta) is quantised to ulps of 1.0 = 6e-8, a row of such weights has maxima
        # of ~1e-3, so one ulp of a weight is ~1e-4 of the normalised row)
        # (a Gaussian of width 0.1 over depth differences of ~1 turns the 1e-6 of the recovered depths into 1e-3)
        np.testing.assert_allclose(o[healthy], g[healthy], rtol=1e-3, atol=5e-3 if lvl == 'fine' else 2e-4, err_msg=f'{lvl}/{k}')
      elif k == 'target_norm':
        # R . normalize(-d sigma / dx) (models.py:1063-1077, 1279-1283, 1327-1330).  The golden gradient is a float64
        # central difference of the reference's own per-point sigma function, the oracle's is float32 autograd:
        # they agree to ~1e-6 except at a point that sits on a ReLU kink of one of the networks
        e = np.abs(o - g).max(-1)
        assert np.median(e) <= 2e-6 and np.mean(e <= 1e-4) >= 0.99, (lvl, np.sort(e.reshape(-1))[-4:])
        np.testing.assert_allclose(np.linalg.norm(o, axis=-1), 1.0, atol=1e-5)
      elif k == 'sigma':                    # softplus of +-30-sized logits: relative
        np.testing.assert_allclose(o, g, rtol=2e-4, atol=2e-5, err_msg=f'{lvl}/{k}')
      elif k in ('med_depth', 'med_points'):
        # a selection (first sample past half the accumulated weight); the sampled depths feeding it agree to ~1e-6
        np.testing.assert_allclose(o, g, rtol=1e-5, atol=1e-5, err_msg=f'{lvl}/{k}')
      else:
        np.testing.assert_allclose(o, g, rtol=1e-4, atol=2e-5, equal_nan=True, err_msg=f'{lvl}/{k}')
      checked += 1
  assert checked >= 40


# ---------------------------------------------------------------------------------------------------------------
# Ray generation: the reference's own hypernerf/camera.py (plain numpy) on three cameras
# ---------------------------------------------------------------------------------------------------------------
CAMERA_KEYS = ('orientation', 'position', 'focal_length', 'principal_point', 'image_size', 'skew', 'pixel_aspect_ratio',
               'radial_distortion', 'tangential_distortion')


def golden_cameras(G):
  for name in G['camera_names']:
    name = str(name)
    kw = {k: G[f'camera_{name}_{k}'] for k in CAMERA_KEYS if f'camera_{name}_{k}' in G}
    kw['image_size'] = [int(v) for v in kw['image_size']]
    for k in ('focal_length', 'skew', 'pixel_aspect_ratio'):
      if k in kw:
        kw[k] = float(kw[k])
    yield name, kw, {k: G[f'camera_{name}_{k}'] for k in ('origins', 'directions', 'pixels')}


def test_camera_rays_oracle(G):
  from oracle import camera_oracle
  n = 0
  for name, kw, gold in golden_cameras(G):
    out = camera_oracle.camera_to_rays(**kw)
    np.testing.assert_array_equal(out['origins'], gold['origins'], err_msg=name)
    np.testing.assert_array_equal(out['pixels'], gold['pixels'], err_msg=name)
    np.testing.assert_allclose(out['directions'], gold['directions'], rtol=0, atol=3e-7, err_msg=name)
    assert np.allclose(np.linalg.norm(out['directions'], axis=-1), 1.0, atol=1e-6)
    n += 1
  assert n == 3


def test_whole_forward_variant_with_flipped_sampling_switches(G):
  """The reference's NerfModel.__call__ once more, at render.py's inference settings (mask_ratio 1, end-of-schedule
  alphas) with deterministic depths (no stratified jitter, u = linspace), linear disparity, white background and no
  sample at infinity: every key of both level dicts.  No draws -> the end-to-end fine level is compared directly."""
  from nerfds_b200.config import nerf_ds_config
  from nerfds_b200.params import unflatten_params
  small = {str(k): int(v) for k, v in zip(G['model_cfg_keys'], G['model_cfg_vals'])}
  cfg = nerf_ds_config(**small).replace(use_stratified_sampling=False, use_white_background=True,
                                        use_linear_disparity=True, use_sample_at_infinity=False)
  P = unflatten_params({k[3:]: v for k, v in G.items() if k.startswith('MP/')})
  ep = {'nerf_alpha': 8.0, 'warp_alpha': 4.0, 'hyper_alpha': 1.0, 'hyper_sheet_alpha': 6.0, 'norm_input_alpha': 4.0}
  rays = {'origins': G['model_origins'], 'directions': G['model_dirs'], 'metadata': {'warp': G['model_warp']},
          'mask': G['model_gt_mask']}
  om = O.OracleNerfModel(cfg, P)
  out = O.to_numpy(om.apply(rays, ep, None, None, return_points=True, return_weights=True, use_predicted_norm=True,
                            compute_sigma_gradient=True, keep_internal=True, mask_ratio=1, sharp_weights_std=0.1))
  o3, d3 = G['model_origins'], G['model_dirs']
  depth_of = lambda pts: (((pts - o3[:, None]) * d3[:, None]).sum(-1) / (d3 ** 2).sum(-1)[:, None]).astype(np.float32)
  # coarse depths: linear in disparity between near and far, identical for every ray
  zc = depth_of(G['modelB_coarse_points'])
  S = zc.shape[1]
  t = np.arange(S, dtype=np.float32) / np.float32(S - 1)
  np.testing.assert_allclose(zc, np.broadcast_to(1.0 / (1.0 / cfg.near * (1 - t) + 1.0 / cfg.far * t), zc.shape), rtol=2e-5)
  np.testing.assert_allclose(out['coarse']['z_vals'], zc, rtol=2e-5, atol=2e-6)
  zf = depth_of(G['modelB_fine_points'])
  fine_on_ref = O.to_numpy(om.render_samples(
      'fine', T(G['modelB_fine_points']), T(zf), T(d3), T(d3), rays['metadata'], ep, G['model_gt_mask'],
      use_sample_at_infinity=False, use_predicted_norm=True, compute_sigma_gradient=True, mask_ratio=1,
      sharp_weights_std=0.1))
  checked = 0
  for lvl, res in (('coarse', out['coarse']), ('fine', fine_on_ref)):
    gold = {k[len(f'modelB_{lvl}_'):]: v for k, v in G.items() if k.startswith(f'modelB_{lvl}_')}
    assert {'rgb', 'depth', 'acc', 'weights', 'sigma', 'warped_points', 'predicted_mask', 'predicted_norm', 'target_norm',
            'ray_norm', 'ray_delta_x', 'med_points'} <= set(gold)
    for k, g in gold.items():
      if k == 'sharp_weights':
        continue                              # ill-conditioned rows, covered by the first variant
      o = np.asarray(res[k]).reshape(g.shape)
      if k == 'target_norm':
        e = np.abs(o - g).max(-1)
        assert np.median(e) <= 2e-6 and np.mean(e <= 1e-4) >= 0.98, (lvl, np.sort(e.reshape(-1))[-4:])
      elif k == 'sigma':
        # (full-frequency posenc: a logit of size 30 carries 4e-6 of float32 noise, 4e-4 of a sigma of 0.1)
        np.testing.assert_allclose(o, g, rtol=5e-4, atol=2e-5, err_msg=f'{lvl}/{k}')
      elif k in ('med_depth', 'med_points'):
        np.testing.assert_allclose(o, g, rtol=1e-5, atol=1e-5, err_msg=f'{lvl}/{k}')
      else:
        np.testing.assert_allclose(o, g, rtol=1e-4, atol=2e-5, equal_nan=True, err_msg=f'{lvl}/{k}')
      checked += 1
  assert checked >= 40
  # end to end (deterministic u): the resampled depths and the fine colours follow the reference's
  dz = np.abs(np.sort(out['fine']['z_vals'], -1) - zf)          # (near-empty bins: the inverse CDF amplifies 1e-7 of weight)
  assert np.mean(dz <= 2e-4) >= 0.98 and dz.max() <= 2e-2, np.sort(dz.reshape(-1))[-4:]
  e = np.abs(out['fine']['rgb'] - G['modelB_fine_rgb']).max(-1)    # one moved sample on a sharp edge moves a 16-sample ray
  assert np.median(e) <= 1e-5 and np.mean(e <= 1e-3) >= 0.89, np.sort(e)[-3:]


def test_filter_sigma_render_opts(G):
  """render_opts (models.filter_sigma, models.py:38-66): the reference's NerfModel.__call__ with a dust threshold and a
  bounding box.  The coarse level receives no render_opts (models.py:1493-1516): identical to the unfiltered run.  The
  fine level on the reference's fine samples: sigma stays unfiltered, weights / rgb / sharp_weights are filtered."""
  from nerfds_b200.config import nerf_ds_config
  from nerfds_b200.params import unflatten_params
  small = {str(k): int(v) for k, v in zip(G['model32_cfg_keys'], G['model32_cfg_vals'])}
  cfg = nerf_ds_config(**small)
  P = unflatten_params({k[5:]: v for k, v in G.items() if k.startswith('MP32/')})
  ep = {str(k): float(v) for k, v in zip(G['model_extra_keys'], G['model_extra_vals'])}
  ropts = {'dust_threshold': float(G['model32R_dust']), 'bounding_box': tuple(float(v) for v in G['model32R_bbox'])}
  for k in ('rgb', 'weights', 'sigma'):
    np.testing.assert_array_equal(G[f'model32R_coarse_{k}'], G[f'model32_coarse_{k}'])
  assert np.abs(G['model32R_fine_weights'] - G['model32_fine_weights']).max() > 0.1        # the options do something
  np.testing.assert_allclose(G['model32R_fine_sigma'], G['model32_fine_sigma'], rtol=1e-6)  # ... but not to out['sigma']
  o3, d3 = G['model_origins'], G['model_dirs']
  pts = G['model32R_fine_points']
  zf = (((pts - o3[:, None]) * d3[:, None]).sum(-1) / (d3 ** 2).sum(-1)[:, None]).astype(np.float32)
  om = O.OracleNerfModel(cfg, P)
  res = O.to_numpy(om.render_samples('fine', T(pts), T(zf), T(d3), T(d3), {'warp': G['model_warp']}, ep, G['model_gt_mask'],
                                     use_sample_at_infinity=True, use_predicted_norm=True, compute_sigma_gradient=False,
                                     mask_ratio=float(G['model_mask_ratio']), sharp_weights_std=0.1, render_opts=ropts))
  n = 0
  for k in ('rgb', 'depth', 'acc', 'weights', 'alpha', 'accum_prod', 'ray_norm', 'ray_delta_x', 'ray_predicted_mask',
            'ray_rotation_field', 'med_depth'):
    g = G[f'model32R_fine_{k}']
    np.testing.assert_allclose(np.asarray(res[k]).reshape(g.shape), g, rtol=1e-4, atol=2e-5, err_msg=k)
    n += 1
  np.testing.assert_allclose(res['sigma'], G['model32R_fine_sigma'], rtol=5e-4, atol=2e-5)
  g = G['model32R_fine_sharp_weights']
  ok = np.isfinite(g).all(-1) & np.isfinite(res['sharp_weights']).all(-1)
  assert ok.sum() >= 3
  # (ill-conditioned rows as in the first variant: only rows whose Gaussian normaliser is healthy are compared)
  w = G['model32R_fine_weights'].astype(np.float64)
  big = np.abs(res['sharp_weights'][ok] - g[ok]).max(-1)
  assert np.median(big) <= 5e-3, np.sort(big)[-4:]
  assert n == 11
