"""JAX-compatible keys and uniform draws."""
import numpy as np
import pytest

from nerfds_b200 import jax_random as jr
from oracle import jax_random_oracle as oj

# Random123 known-answer vectors for Threefry-2x32 (20 rounds), as used by jax's own random_test.py
THREEFRY_KAT = [
    ((0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6b200159, 0x99ba4efe)),
    ((0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x1cb996fc, 0xbb002be7)),
    ((0x13198a2e, 0x03707344), (0x243f6a88, 0x85a308d3), (0xc4923a9c, 0x483df7a0)),
]


def test_threefry_known_answers_host_and_oracle():
  for key, ctr, want in THREEFRY_KAT:
    assert jr.threefry2x32(key, *ctr) == want
    a, b = oj.threefry2x32(key[0], key[1], [ctr[0]], [ctr[1]])
    assert (int(a[0]), int(b[0])) == want


def test_documented_jax_values():
  """Values printed in the JAX documentation for the default (threefry, non-partitionable) PRNG."""
  assert jr.PRNGKey(0) == (0, 0) and jr.PRNGKey(42) == (0, 42) and jr.PRNGKey((7 << 32) + 5) == (7, 5)
  assert jr.split(jr.PRNGKey(0)) == [(4146024105, 967050713), (2718843009, 1272950319)]
  assert oj.uniform(jr.PRNGKey(0), ()).item() == np.float32(0.41845703)
  assert [int(v) for v in oj.random_bits((0, 0), 4)] == [4146024105, 967050713, 2718843009, 1272950319]


def test_host_bits_match_oracle_for_even_and_odd_counts():
  for n in (1, 2, 5, 8, 33):
    key = jr.fold_in(jr.PRNGKey(n), 17)
    assert jr._random_bits(key, n) == [int(v) for v in oj.random_bits(key, n)]
  # fold_in is one block on counters (0, data); split(num) is bits(2 num) in pairs
  assert jr.fold_in((1, 2), 9) == jr.threefry2x32((1, 2), 0, 9)
  ks = jr.split((3, 4), 3)
  assert [w for k in ks for w in k] == [int(v) for v in oj.random_bits((3, 4), 6)]


def test_uniform_properties():
  u = oj.uniform((123, 456), (257, 33))
  assert u.dtype == np.float32 and u.shape == (257, 33) and u.min() >= 0.0 and u.max() < 1.0
  assert abs(u.mean() - 0.5) < 0.01 and abs(u.var() - 1 / 12) < 0.005
  # values are multiples of 2^-23
  assert np.all((u.astype(np.float64) * 2 ** 23) % 1 == 0)
  # prefix instability of the half-split layout: the first row of a taller draw differs (no accidental row-major stream)
  assert not np.array_equal(oj.uniform((123, 456), (258, 33))[:257], u)


def test_flax_make_rng_folding_rule():
  import hashlib
  key = jr.PRNGKey(5)
  h1 = int.from_bytes(hashlib.sha1(b'\x01').digest()[:4], 'big')
  assert jr.flax_make_rng(key) == jr.fold_in(key, h1)
  h = int.from_bytes(hashlib.sha1(b'nerf_mlps_coarse' + b'\x02').digest()[:4], 'big')
  assert jr.flax_make_rng(key, ('nerf_mlps_coarse',), 2) == jr.fold_in(key, h)
  assert jr.flax_make_rng(key, (), 1) != jr.flax_make_rng(key, (), 2)
  with pytest.raises(ValueError):
    jr.flax_make_rng(key, (1.5,))
  with pytest.raises(ValueError):
    jr.as_key(np.zeros(3, np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize('shape', [(1,), (7,), (64, 128), (1001, 129), (640, 128 * 25 + 1)])
def test_device_uniform_is_bit_exact(cuda_device, shape):
  key = jr.flax_make_rng(jr.split(jr.PRNGKey(20230601), 3)[1])
  got = jr.uniform(key, shape, cuda_device).cpu().numpy()
  np.testing.assert_array_equal(got, oj.uniform(key, shape))


@pytest.mark.gpu
def test_model_draws_follow_the_jax_keys(cuda_device):
  """NerfModel.apply(rngs=...) draws t_rand / u exactly as flax + jax would from those keys."""
  from nerfds_b200 import synthetic as syn
  from nerfds_b200.models import NerfModel
  from tests.common import make_case
  cfg, params, rays, _, _ = make_case('nerf_ds', image=9, seed=1, num_coarse_samples=16, num_fine_samples=8)
  B = rays['origins'].shape[0]
  k_coarse, k_fine = np.array([11, 22], np.uint32), np.array([33, 44], np.uint32)
  m = NerfModel(cfg, device=cuda_device)
  a = m.apply({'params': params}, rays, syn.final_extra_params(), rngs={'coarse': k_coarse, 'fine': k_fine},
              use_predicted_norm=True, keys=('rgb', 'z_vals'), coarse_keys=('z_vals',))
  t_rand = oj.uniform(jr.flax_make_rng(k_coarse), (B, 16))
  u = oj.uniform(jr.flax_make_rng(k_fine), (B, 8))
  b = m.apply({'params': params}, rays, syn.final_extra_params(), t_rand=t_rand, u=u, use_predicted_norm=True,
              keys=('rgb', 'z_vals'), coarse_keys=('z_vals',))
  for lvl in ('coarse', 'fine'):
    for k in a[lvl]:
      assert np.array_equal(a[lvl][k].cpu().numpy(), b[lvl][k].cpu().numpy()), (lvl, k)


@pytest.mark.gpu
def test_render_image_draws_per_device_keys(cuda_device):
  """evaluation.render_image splits the rng like evaluation.py:81-84 (4 keys, then one per device); a single
  process stands in for D devices by drawing each shard's samples from that device's key."""
  import torch
  from nerfds_b200 import evaluation, synthetic as syn
  from nerfds_b200.model_utils import TrainState
  from nerfds_b200.models import NerfModel
  from tests.common import make_case
  cfg, params, rays, _, _ = make_case('nerf_ds', image=8, seed=2, num_coarse_samples=16, num_fine_samples=8)
  m = NerfModel(cfg, device=cuda_device)
  state = TrainState.create(params, syn.final_extra_params())
  D, rng = 2, np.array([0, 7], np.uint32)
  img = {k: (v.reshape(8, 8, -1) if not isinstance(v, dict) else {a: b.reshape(8, 8, -1) for a, b in v.items()})
         for k, v in rays.items()}
  out = evaluation.render_image(state, img, evaluation.make_model_fn(m, keys=('rgb',), group=False), D, rng, chunk=64)
  _, k0, k1, _ = jr.split(rng, 4)
  k0, k1 = jr.split(k0, D), jr.split(k1, D)
  t_rand = np.concatenate([oj.uniform(jr.flax_make_rng(k0[d]), (32, 16)) for d in range(D)])
  u = np.concatenate([oj.uniform(jr.flax_make_rng(k1[d]), (32, 8)) for d in range(D)])
  ref = m.apply({'params': params}, rays, syn.final_extra_params(), t_rand=t_rand, u=u, use_predicted_norm=True,
                mask_ratio=1, sharp_weights_std=0.1, keys=('rgb',), coarse_keys=())
  assert torch.equal(out['rgb'].reshape(-1, 3), ref['fine']['rgb'].cpu())


@pytest.mark.gpu
def test_device_uniform_row_ranges_and_host_call_with_keys(cuda_device):
  """ndsr_random_uniform_range: any block of rows of uniform(key, [n, S]) equals the same rows of the whole array
  (even and odd totals).  ndsr_render_rays_host_rng (draws generated on the device, per internal ray chunk) gives the
  same frame, bit for bit, as ndsr_render_rays_host fed those draws from the host."""
  from nerfds_b200 import synthetic as syn
  from nerfds_b200.models import NerfModel
  from tests.common import make_case
  cfg, params, rays, _, _ = make_case('nerf_ds', image=9, seed=1, num_coarse_samples=16, num_fine_samples=8)
  m = NerfModel(cfg, device=cuda_device)
  R = m.renderer
  key = jr.flax_make_rng(jr.split(jr.PRNGKey(7), 2)[0])
  for n, S in ((81, 16), (33, 7)):
    full = oj.uniform(key, (n, S))
    for first, rows in ((0, n), (5, 11), (n - 3, 3), (n // 2 - 1, 4)):
      got = R.random_uniform(key, n, S, first, rows).cpu().numpy()
      np.testing.assert_array_equal(got, full[first:first + rows])
  R.load_params(params)
  R.set_max_chunk(32)                                                   # 81 rays -> 3 internal chunks
  B = rays['origins'].shape[0]
  extra = R.make_extra(syn.final_extra_params(), use_predicted_norm=True)
  kc, kf = np.array([11, 22], np.uint32), np.array([33, 44], np.uint32)
  wid = rays['metadata']['warp'].reshape(-1)
  a = R.render_rays_host(rays['origins'], rays['directions'], warp_id=wid, gt_mask=rays['mask'].reshape(-1), extra=extra,
                         fine_keys=('rgb', 'depth'), rng_keys=(kc, kf))
  b = R.render_rays_host(rays['origins'], rays['directions'], warp_id=wid, gt_mask=rays['mask'].reshape(-1), extra=extra,
                         fine_keys=('rgb', 'depth'), t_rand=oj.uniform(kc, (B, 16)), u=oj.uniform(kf, (B, 8)))
  for k in a:
    np.testing.assert_array_equal(a[k], b[k])


@pytest.mark.gpu
def test_reference_draws_follow_render_image_layout(cuda_device):
  """evaluation.reference_draws: the draws of a whole frame laid out as evaluation.py:81-120 produces them under pmap
  (same per-device keys for every chunk, edge-padded last chunk sharded over the devices)."""
  from nerfds_b200.evaluation import reference_draws
  rng, n, chunk, D, Sc, Sf = np.array([3, 9], np.uint32), 23, 10, 2, 6, 5
  t, u = reference_draws(rng, n, chunk, D, Sc, Sf, cuda_device)
  _, k0, k1, _ = jr.split(rng, 4)
  k0, k1 = jr.split(k0, D), jr.split(k1, D)
  want_t, want_u = [], []
  for ray in range(0, n, chunk):
    m = min(chunk, n - ray)
    rows = -(-m // D)
    want_t.append(np.concatenate([oj.uniform(jr.flax_make_rng(k0[d]), (rows, Sc)) for d in range(D)])[:m])
    want_u.append(np.concatenate([oj.uniform(jr.flax_make_rng(k1[d]), (rows, Sf)) for d in range(D)])[:m])
  np.testing.assert_array_equal(t.cpu().numpy(), np.concatenate(want_t))
  np.testing.assert_array_equal(u.cpu().numpy(), np.concatenate(want_u))
