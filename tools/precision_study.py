"""Emulates the tensor-core number formats candidate for the sm_100a path and
reports RGB L-inf against the fp32 oracle (north_star tolerance: 1e-3).

fp16 operands + fp32 accumulate are emulated exactly on CPU: an fp16*fp16
product is exact in fp32, so ``x.half().float() @ W.half().float()`` has the
same rounding points as tcgen05.mma kind::f16 (up to accumulation order).
Modes per network: 'f32' | 'h1' (1 MMA) | 'h2' (A=hi+lo, W fp16: 2 MMAs) |
'h3' (A_hi W_hi + A_lo W_hi + A_hi W_lo: 3 MMAs) | 'b1'/'b3' (bf16).
Run:  python tools/precision_study.py
"""
import sys, itertools
sys.path.insert(0, '.')
import numpy as np, torch
from nerfds_b200.config import nerf_ds_config
from nerfds_b200.params import init_params
from nerfds_b200 import synthetic as syn
from oracle.nerfds_oracle import OracleNerfModel, to_numpy


def make_mm(modes):
  def split(x, dt):
    hi = x.to(dt).float()
    lo = (x - hi).to(dt).float()
    return hi, lo
  def mm(x, W, tag=''):
    net = tag.split('/')[0]
    m = modes.get(net, 'f32')
    if m == 'f32':
      return x @ W
    dt = torch.float16 if m[0] == 'h' else torch.bfloat16
    xh, xl = split(x, dt)
    Wh, Wl = split(W, dt)
    if m[1] == '1':
      return xh @ Wh
    if m[1] == '2':
      return xh @ Wh + xl @ Wh
    return xh @ Wh + xl @ Wh + xh @ Wl
  return mm


def run(modes, cfg, P, rays, t_rand, u, ref=None):
  m = OracleNerfModel(cfg, P)
  m.mm = make_mm(modes)
  out = to_numpy(m.apply(rays, syn.final_extra_params(), t_rand, u,
                         use_predicted_norm=True, return_weights=True,
                         compute_sigma_gradient=False, keep_internal=True))
  return out


if __name__ == '__main__':
  torch.manual_seed(0)
  cfg = nerf_ds_config(num_coarse_samples=64, num_fine_samples=64)
  P = init_params(cfg, 0)
  rays = syn.frame_rays(32, 32, frame=3)
  B = rays['origins'].shape[0]
  t_rand, u = syn.uniform_draws(B, cfg.num_coarse_samples, cfg.num_fine_samples)
  ref = run({}, cfg, P, rays, t_rand, u)
  nets = ['mask', 'warp', 'hyper', 'trunk', 'bottleneck', 'alpha', 'rgb']
  def report(name, modes):
    o = run(modes, cfg, P, rays, t_rand, u)
    e = {k: float(np.abs(o['fine'][k] - ref['fine'][k]).max())
         for k in ('rgb', 'depth', 'ray_norm', 'ray_delta_x', 'ray_predicted_mask')}
    ec = float(np.abs(o['coarse']['rgb'] - ref['coarse']['rgb']).max())
    e99 = float(np.percentile(np.abs(o['fine']['rgb'] - ref['fine']['rgb']), 99))
    print(f'{name:34s} rgb Linf {e["rgb"]:.2e} (p99 {e99:.2e}, coarse {ec:.2e}) depth {e["depth"]:.2e} '
          f'norm {e["ray_norm"]:.2e} dx {e["ray_delta_x"]:.2e} mask {e["ray_predicted_mask"]:.2e}')
  report('all h1', {n: 'h1' for n in nets})
  report('all b1', {n: 'b1' for n in nets})
  for n in nets:
    report(f'only {n} h1', {n: 'h1'})
  report('all h2', {n: 'h2' for n in nets})
  report('all h3', {n: 'h3' for n in nets})
  report('warp h3, rest h1', {**{n: 'h1' for n in nets}, 'warp': 'h3'})
  report('warp+mask h3, rest h1', {**{n: 'h1' for n in nets}, 'warp': 'h3', 'mask': 'h3'})
  report('warp+mask+hyper h3, rest h1', {**{n: 'h1' for n in nets}, 'warp': 'h3', 'mask': 'h3', 'hyper': 'h3'})
  report('warp+mask+hyper h3, rest h2', {**{n: 'h2' for n in nets}, 'warp': 'h3', 'mask': 'h3', 'hyper': 'h3'})
  report('warp+mask+hyper f32, rest h1', {n: 'h1' for n in ('trunk', 'bottleneck', 'alpha', 'rgb')})
  report('warp+mask+hyper f32, rest h2', {n: 'h2' for n in ('trunk', 'bottleneck', 'alpha', 'rgb')})
