"""Tensor-core machinery on hardware: one Dense layer through weight packing,
the bulk-TMA ring, tcgen05.mma, the TMEM epilogue and the swizzled split-fp16
write-back (ndsr_selftest_tc_dense), against fp64 numpy."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(k_hid, k_in, n_out, terms, relu, out_kind, seed=0, scale_w=0.1):
  from nerfds_b200 import _lib
  lib = _lib.load_library()
  rng = np.random.default_rng(seed)
  K = k_hid + k_in
  A = rng.normal(size=(128, K)).astype(np.float32)
  W = (rng.normal(size=(K, n_out)) * scale_w).astype(np.float32)
  b = rng.normal(size=(n_out,)).astype(np.float32)
  out = np.zeros((128, n_out), np.float32)
  rb = np.zeros((128, n_out), np.float32)
  p = lambda a: C.c_void_p(a.ctypes.data)
  rc = lib.ndsr_selftest_tc_dense(0, k_hid, k_in, n_out, terms, relu, out_kind, p(A), p(W), p(b), p(out), p(rb))
  assert rc == 0
  ref = A.astype(np.float64) @ W.astype(np.float64) + b
  if relu:
    ref = np.maximum(ref, 0)
  return out, rb, ref


# out_kind: 0 = hidden layer (accumulators converted in place to hi + lo operands in tensor memory);
#           2 = hidden layer, hi only; 5 = hidden layer, compacted hi (n_out = 256); 1 = head (n_out <= 16).
# k_hid activations come from tensor memory (TS-mode MMA, layout of a k_hid-wide layer), k_in from shared memory.
SHAPES = [(0, 52, 256), (256, 0, 256), (256, 52, 256), (128, 33, 128), (64, 45, 64), (128, 0, 128), (256, 48, 128),
          (64, 0, 64), (128, 0, 64), (64, 0, 128)]


@pytest.mark.parametrize('k_hid,k_in,n_out', SHAPES)
def test_dense_split3_matches_fp64(cuda_device, k_hid, k_in, n_out):
  out, rb, ref = _run(k_hid, k_in, n_out, 3, 1, 0)
  scale = np.abs(ref).max()
  assert np.abs(out - ref).max() <= 2e-5 * scale, np.abs(out - ref).max() / scale
  # the operand image written for the next layer reproduces the activations to ~2^-21
  assert np.abs(rb - out).max() <= 2e-6 * scale


@pytest.mark.parametrize('k_hid,k_in,n_out,kind', [(256, 0, 256, 5), (256, 48, 128, 2), (128, 0, 3, 1), (256, 0, 256, 0)])
def test_dense_fp16_single_term(cuda_device, k_hid, k_in, n_out, kind):
  out, rb, ref = _run(k_hid, k_in, n_out, 1, 0, kind)
  scale = np.abs(ref).max()
  assert np.abs(out - ref).max() <= 3e-3 * scale
  if kind in (2, 5):
    assert np.abs(rb - out).max() <= 1e-3 * scale   # hi-only image keeps 11 bits


@pytest.mark.parametrize('k_hid,n_out', [(256, 4), (128, 6), (64, 2), (128, 1)])
def test_heads_split3(cuda_device, k_hid, n_out):
  out, rb, ref = _run(k_hid, 0, n_out, 3, 0, 1, seed=3)
  scale = max(1.0, np.abs(ref).max())
  assert np.abs(out - ref).max() <= 2e-5 * scale


def test_weight_scaling_extremes(cuda_device):
  for sw in (1e-4, 30.0):
    out, rb, ref = _run(256, 52, 256, 3, 1, 0, seed=5, scale_w=sw)
    scale = np.abs(ref).max()
    assert np.abs(out - ref).max() <= 3e-5 * scale
