"""Closed-form known-answer tests that pin the CPU oracle (the reference ships
no tests or golden vectors for this path -- SURVEY.md section 4 / 8(c))."""
import math

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import nerfds_oracle as O


def test_posenc_layout_and_values():
  """(F, 2, C) flattened: [sin(2^k x_0..C), sin(2^k x_0..C + pi/2)] per band, identity first."""
  x = torch.tensor([[0.1, -0.2, 0.3]])
  f = O.posenc(x, 0, 3)
  assert f.shape == (1, 18)
  for k in range(3):
    for c in range(3):
      assert f[0, k * 6 + c].item() == pytest.approx(math.sin(2 ** k * x[0, c].item()), abs=1e-6)
      assert f[0, k * 6 + 3 + c].item() == pytest.approx(math.cos(2 ** k * x[0, c].item()), abs=1e-6)
  fi = O.posenc(x, 0, 3, use_identity=True)
  assert torch.equal(fi[:, :3], x) and torch.allclose(fi[:, 3:], f)


def test_posenc_window_schedule():
  w = O.posenc_window(0, 4, 2.5)
  assert w[0].item() == pytest.approx(1.0, abs=1e-6) and w[1].item() == pytest.approx(1.0, abs=1e-6)
  assert w[2].item() == pytest.approx(0.5 * (1 + math.cos(math.pi * 0.5 + math.pi)), abs=1e-6)
  assert w[3].item() == pytest.approx(0.0, abs=1e-6)
  # alpha >= max_deg: the window is all ones and posenc(alpha) == posenc(None)
  x = torch.rand(7, 3)
  assert torch.allclose(O.posenc(x, 0, 4, alpha=4.0), O.posenc(x, 0, 4), atol=1e-6)


def test_rodrigues_identities():
  rng = np.random.default_rng(0)
  w = torch.from_numpy(rng.normal(size=(64, 3))).float()
  w = w / w.norm(dim=-1, keepdim=True)
  theta = torch.from_numpy(rng.uniform(0.01, 3.0, size=64)).float()
  v = torch.from_numpy(rng.normal(size=(64, 3))).float()
  R, p = O.exp_se3_rp(torch.cat([w, v], -1), theta)
  eye = torch.eye(3).expand(64, 3, 3)
  assert torch.allclose(R @ R.transpose(-1, -2), eye, atol=2e-6)
  assert torch.allclose(torch.linalg.det(R), torch.ones(64), atol=2e-6)
  assert torch.allclose((R @ w[..., None])[..., 0], w, atol=2e-6)          # axis is fixed
  assert torch.allclose(torch.einsum('bii->b', R), 1 + 2 * torch.cos(theta), atol=5e-6)
  x = torch.from_numpy(rng.normal(size=(64, 3))).float()
  y = O.se3_apply(R, p, x)
  assert torch.allclose(O.se3_apply(R, p, y, inverse=True), x, atol=5e-6)  # inverse undoes it
  assert torch.allclose(O.se3_apply(R, p, x, rotation_only=True), (R @ x[..., None])[..., 0])
  # pure translation limit: p -> theta * v as rotation axis contribution vanishes along w
  R0, p0 = O.exp_se3_rp(torch.tensor([[0., 0., 1., 0., 0., 2.]]), torch.tensor([0.5]))
  assert torch.allclose(p0, torch.tensor([[0., 0., 1.0]]), atol=1e-6)     # screw along its own axis


def test_sample_along_rays_deterministic_and_stratified():
  o = torch.zeros(3, 3)
  d = torch.tensor([[0., 0., 1.]]).expand(3, 3)
  z, pts = O.sample_along_rays(None, o, d, 5, 1.0, 3.0, False, False)
  assert torch.allclose(z[0], torch.tensor([1.0, 1.5, 2.0, 2.5, 3.0]))
  assert torch.allclose(pts[0, :, 2], z[0])
  z, _ = O.sample_along_rays(None, o, d, 3, 1.0, 4.0, False, True)            # linear in disparity
  assert torch.allclose(z[0], torch.tensor([1.0, 1.6, 4.0]), atol=1e-6)
  t0 = torch.zeros(3, 5)
  t1 = torch.full((3, 5), 1.0)
  zl, _ = O.sample_along_rays(t0, o, d, 5, 1.0, 3.0, True, False)
  zu, _ = O.sample_along_rays(t1, o, d, 5, 1.0, 3.0, True, False)
  assert torch.allclose(zl[0], torch.tensor([1.0, 1.25, 1.75, 2.25, 2.75]))
  assert torch.allclose(zu[0], torch.tensor([1.25, 1.75, 2.25, 2.75, 3.0]))


def test_volumetric_rendering_constant_density():
  S, sigma, dz = 17, 2.0, 0.125
  z = (1.0 + dz * torch.arange(S).float())[None]
  d = torch.tensor([[0., 0., 1.]])
  rgb = torch.full((1, S, 3), 0.5)
  out = O.volumetric_rendering(rgb, torch.full((1, S), sigma), z, d, False, True)
  a = 1 - math.exp(-sigma * dz)
  w_ref = torch.tensor([a * (1 - a) ** i for i in range(S)])
  w_ref[-1] = (1 - a) ** (S - 1)
  assert torch.allclose(out['weights'][0], w_ref, atol=1e-6)
  assert out['weights'].sum().item() == pytest.approx(1.0, abs=1e-5)
  assert out['acc'][0].item() == pytest.approx(1 - (1 - a) ** (S - 1), abs=1e-5)
  k = int(np.argmax(np.cumsum(w_ref.numpy()) >= 0.5))
  assert out['med_depth'][0].item() == z[0, k].item()
  # white background adds the missing opacity; without the sample at infinity acc < 1
  out2 = O.volumetric_rendering(rgb, torch.full((1, S), sigma), z, d, True, False)
  acc = out2['weights'].sum().item()
  assert out2['rgb'][0, 0].item() == pytest.approx(0.5 * acc + (1 - acc), abs=1e-5)


def test_inverse_cdf_equals_searchsorted_right():
  """the mask/max/min construction == searchsorted(cdf, u, 'right')."""
  rng = np.random.default_rng(1)
  B, n, nf = 64, 63, 64
  z = np.sort(rng.uniform(0, 1, (B, n + 1)).astype(np.float32), -1)
  bins = torch.from_numpy((0.5 * (z[:, 1:] + z[:, :-1])).astype(np.float32))
  w = torch.from_numpy((rng.random((B, n - 1)) ** 6).astype(np.float32))
  u = torch.from_numpy(rng.random((B, nf), dtype=np.float32))
  zs, lo, hi, cdf = O.piecewise_constant_pdf(u, bins, w, return_indices=True)
  k = np.stack([np.searchsorted(cdf[b].numpy(), u[b].numpy(), side='right') for b in range(B)])
  assert np.array_equal(lo.numpy(), np.clip(k - 1, 0, n - 2))
  assert np.array_equal(hi.numpy(), np.clip(k, 1, n - 1))
  b0 = np.take_along_axis(bins.numpy(), lo.numpy(), -1)
  b1 = np.take_along_axis(bins.numpy(), hi.numpy(), -1)
  assert np.all(zs.numpy() >= b0 - 1e-6) and np.all(zs.numpy() <= b1 + 1e-6)


def test_uniform_weights_give_linear_inverse_cdf():
  n = 11
  bins = torch.linspace(0, 1, n)[None]
  w = torch.ones(1, n - 1)
  u = torch.tensor([[0.0, 0.25, 0.5, 0.999]])
  zs = O.piecewise_constant_pdf(u, bins, w)
  assert torch.allclose(zs, u, atol=1e-5)


def test_sharpen_weights_gathers_rows():
  """App. C-2: ray b is centred on z_vals[argmax_b] (a ray index, clamped)."""
  w = torch.tensor([[0.1, 0.7, 0.2], [0.6, 0.3, 0.1]])
  z = torch.tensor([[1.0, 2.0, 3.0], [10.0, 20.0, 30.0]])
  s = O.sharpen_weights(w, z, std=1.0)
  # ray 0: argmax 1 -> centred on z[1] = [10,20,30]; ray 1: argmax 0 -> centred on z[0]
  g0 = torch.exp(-0.5 * (z[0] - z[1]) ** 2)
  ref0 = w[0] * g0 / (w[0] * g0).sum()
  assert torch.allclose(s[0], ref0, atol=1e-6)


def test_normalize_vector_eps():
  v = torch.zeros(1, 3)
  assert torch.equal(O.normalize_vector(v), v)                # 0 / sqrt(eps) = 0, not NaN
  v = torch.tensor([[3.0, 0.0, 4.0]])
  assert torch.allclose(O.normalize_vector(v), torch.tensor([[0.6, 0.0, 0.8]]))


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1))
def test_weights_properties(seed):
  rng = np.random.default_rng(seed)
  B, S = 8, 32
  z = torch.from_numpy(np.sort(rng.uniform(0.1, 3, (B, S)).astype(np.float32), -1))
  sigma = torch.from_numpy((rng.random((B, S)) ** 4 * 50).astype(np.float32))
  d = torch.from_numpy(rng.normal(size=(B, 3)).astype(np.float32))
  out = O.volumetric_rendering(torch.rand(B, S, 3), sigma, z, d, False, True)
  w = out['weights']
  assert torch.all(w >= 0) and torch.all(w.sum(-1) <= 1 + 1e-5)
  u = torch.from_numpy(rng.random((B, 16), dtype=np.float32))
  zm = 0.5 * (z[:, 1:] + z[:, :-1])
  zf, _ = O.sample_pdf(u, zm, w[:, 1:-1], torch.zeros(B, 3), d, z)
  assert torch.all(zf[:, 1:] >= zf[:, :-1])                   # sorted after resample
  assert zf.shape == (B, S + 16)


def test_sigma_gradient_matches_finite_differences():
  """The autograd d(sigma_raw)/dx (stand-in for jax.value_and_grad, models.py:1069)
  against fp64 central differences through warp -> hyper sheet -> trunk."""
  from nerfds_b200.config import nerf_ds_config
  from nerfds_b200.params import init_params
  from nerfds_b200 import synthetic as syn
  cfg = nerf_ds_config(num_warp_embeds=4)
  P = init_params(cfg, 0)
  m = O.OracleNerfModel(cfg, P, dtype=torch.float64)
  ep = syn.final_extra_params()
  rng = np.random.default_rng(0)
  N = 6
  x = torch.from_numpy(rng.uniform(-1, 1, (N, 3)))
  we = m.P['warp_embed']['embed']['embedding'][torch.arange(N) % 4]
  mask = torch.from_numpy(rng.random((N, 1)))
  xs = x.clone().requires_grad_(True)
  s, _ = m._sigma_field('fine', xs, we, mask, ep)
  (g,) = torch.autograd.grad(s.sum(), xs)
  h = 1e-6
  for i in range(3):
    e = torch.zeros(3, dtype=torch.float64)
    e[i] = h
    sp, _ = m._sigma_field('fine', x + e, we, mask, ep)
    sm, _ = m._sigma_field('fine', x - e, we, mask, ep)
    fd = (sp - sm) / (2 * h)
    assert torch.allclose(fd, g[:, i], rtol=2e-4, atol=1e-5), (i, fd, g[:, i])
