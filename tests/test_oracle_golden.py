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
