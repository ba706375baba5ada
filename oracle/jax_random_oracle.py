"""CPU restatement of jax.random.uniform under the default threefry2x32 PRNG.

TEST INFRASTRUCTURE ONLY (the checker of `ndsr_random_uniform`); the product
never imports this file.  Follows jax 0.3.15 `jax/_src/prng.py`
(`threefry_2x32`, `threefry_random_bits`) and `jax/_src/random.py` (`_uniform`),
which the reference reaches at model_utils.py:84 and 217.  jax is a
third-party dependency absent from /root/reference (requirements_exact.txt:
jax==0.3.15) and not installable here; pinned by published known answers
(Random123 Threefry-2x32 vectors, `split(PRNGKey(0))`, `uniform(PRNGKey(0))`)
in tests/test_jax_random.py.
"""
import numpy as np

_R0, _R1 = (13, 15, 26, 6), (17, 29, 16, 24)


def threefry2x32(k0, k1, x0, x1):
  """Vectorised Threefry-2x32-20: uint32 arrays x0, x1 under key (k0, k1)."""
  x0, x1 = np.array(x0, dtype=np.uint32), np.array(x1, dtype=np.uint32)
  ks = [np.uint32(k0), np.uint32(k1), np.uint32(k0) ^ np.uint32(k1) ^ np.uint32(0x1BD11BDA)]
  with np.errstate(over='ignore'):
    x0 += ks[0]
    x1 += ks[1]
    for g in range(5):
      for r in (_R0 if g % 2 == 0 else _R1):
        x0 += x1
        x1 = (x1 << np.uint32(r)) | (x1 >> np.uint32(32 - r))
        x1 ^= x0
      x0 += ks[(g + 1) % 3]
      x1 += ks[(g + 2) % 3] + np.uint32(g + 1)
  return x0, x1


def random_bits(key, n):
  """threefry_2x32(key, iota(n)): [first words of the blocks | second words], odd n padded with a 0 counter."""
  half = (n + 1) // 2
  cnt = np.zeros(2 * half, np.uint32)
  cnt[:n] = np.arange(n, dtype=np.uint32)
  a, b = threefry2x32(key[0], key[1], cnt[:half], cnt[half:])
  return np.concatenate([a, b])[:n]


def uniform(key, shape):
  """random.uniform(key, shape, float32, 0, 1): mantissa bits under exponent 0, minus 1."""
  n = int(np.prod(shape))
  bits = random_bits(key, n)
  f = ((bits >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32) - np.float32(1.0)
  return np.maximum(np.float32(0.0), f).reshape(shape)
