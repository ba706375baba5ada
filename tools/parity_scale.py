"""Parity of the tensor-core engine against the CPU oracle AT SCALE (needs a GPU).

Runs the fp32 oracle (row chunks) on N rays of the bench scene at nerf_ds.gin widths, 128+128 samples, then the
CUDA engines on the ORACLE's samples of both levels, and reports every per-ray key's error distribution, the rays
over the north_star bound and -- for the worst rays -- the per-sample picture next to the fp64 oracle.

usage: python tools/parity_scale.py [--rays 16384] [--prec split3] [--engines tc,simt] [--out gpurun_out/x.txt]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from nerfds_b200 import synthetic as syn
from nerfds_b200.config import nerf_ds_config
from nerfds_b200.params import init_params
from oracle.nerfds_oracle import OracleNerfModel, to_numpy

PER_RAY = ('rgb', 'depth', 'acc', 'ray_norm', 'ray_delta_x', 'ray_hyper_points', 'ray_predicted_mask',
           'ray_rotation_field', 'ray_translation_field', 'med_points')


from tests.common import bench_scene as scene, run_oracle_chunks, take_rays as take


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--rays', type=int, default=16384)
  ap.add_argument('--coarse', type=int, default=128)
  ap.add_argument('--fine', type=int, default=128)
  ap.add_argument('--prec', default='split3')
  ap.add_argument('--engines', default='tc,simt')
  ap.add_argument('--worst', type=int, default=6)
  ap.add_argument('--out', default=None)
  args = ap.parse_args()
  lines = []

  def say(s=''):
    print(s, flush=True)
    lines.append(s)

  torch.set_num_threads(os.cpu_count() or 1)
  cfg, params, rays, t_rand, u = scene(args.rays, args.coarse, args.fine)
  t0 = time.time()
  ref = run_oracle_chunks(cfg, params, rays, t_rand, u)
  say(f'oracle fp32: {args.rays} rays, {args.coarse}+{args.fine} samples, {time.time() - t0:.1f} s on {os.cpu_count()} cores')

  from nerfds_b200.models import NerfModel
  outs = {}
  for eng in args.engines.split(','):
    m = NerfModel(cfg, device='cuda:0', engine=eng, precision=args.prec)
    R = m.renderer
    R.ensure_params(params)
    extra = R.make_extra(syn.final_extra_params(), use_predicted_norm=True, mask_ratio=1.0, sharp_weights_std=0.1)
    keys = list(R.level_keys(return_points=True, return_weights=True, want_target_norm=False))
    outs[eng] = {}
    for lvl, name in ((0, 'coarse'), (1, 'fine')):
      r = ref[name]
      o = R.render_samples(lvl, r['z_vals'], rays['directions'], origins=rays['origins'],
                           warp_id=rays['metadata']['warp'], gt_mask=rays['mask'], extra=extra,
                           use_sample_at_infinity=cfg.use_sample_at_infinity, keys=keys)
      outs[eng][name] = {k: v.detach().cpu().numpy() for k, v in o.items()}
    del m

  bound = 1e-3
  for eng, eo in outs.items():
    for name in ('coarse', 'fine'):
      r = ref[name]
      for k in PER_RAY:
        if k not in eo[name] or k not in r or not r[k].size:
          continue
        e = np.abs(eo[name][k].reshape(r[k].shape).astype(np.float64) - r[k]).reshape(r[k].shape[0], -1).max(1)
        say(f'{eng:5s} {args.prec:7s} {name:6s} {k:22s} max {e.max():.2e}  p99.9 {np.percentile(e, 99.9):.2e}  '
            f'p99 {np.percentile(e, 99):.2e}  median {np.median(e):.2e}  rays>1e-3 {int((e > bound).sum())}/{e.size}')
      for k in ('sigma', 'weights', 'predicted_mask', 'warped_points'):
        a = eo[name][k].reshape(r[k].shape).astype(np.float64)
        e = np.abs(a - r[k])
        rel = e / np.maximum(np.abs(r[k]), 1.0)
        say(f'{eng:5s} {args.prec:7s} {name:6s} {k:22s} max abs {e.max():.2e}  max rel(>=1) {rel.max():.2e}')

  # ---- the worst rays of the first engine, next to the fp64 oracle on the same samples
  eng0 = args.engines.split(',')[0]
  m64 = OracleNerfModel(cfg, params, dtype=torch.float64)
  for name in ('coarse', 'fine'):
    r = ref[name]
    e = np.abs(outs[eng0][name]['rgb'] - r['rgb']).max(-1)
    worst = np.argsort(e)[-max(args.worst, 1):][::-1]
    sub = take(rays, worst)
    z = torch.from_numpy(r['z_vals'][worst]).double()
    pts = (torch.from_numpy(sub['origins']).float()[:, None, :] +
           torch.from_numpy(r['z_vals'][worst])[..., None] * torch.from_numpy(sub['directions']).float()[:, None, :]).double()
    d64 = torch.from_numpy(sub['directions']).double()
    o64 = to_numpy(m64.render_samples(name, pts, z, d64, d64, sub['metadata'], syn.final_extra_params(), sub['mask'],
                                      use_sample_at_infinity=cfg.use_sample_at_infinity, use_predicted_norm=True,
                                      mask_ratio=1, sharp_weights_std=0.1, compute_sigma_gradient=False))
    say(f'--- {name}: worst rays of {eng0} (rgb error vs fp32 oracle | vs fp64 oracle; fp32 oracle vs fp64)')
    np.set_printoptions(precision=6, linewidth=220)
    for j, w in enumerate(worst):
      row = f'ray {w}: '
      for eng in outs:
        row += f'{eng} {np.abs(outs[eng][name]["rgb"][w] - r["rgb"][w]).max():.2e} | {np.abs(outs[eng][name]["rgb"][w] - o64["rgb"][j]).max():.2e}   '
      row += f'oracle32-64 {np.abs(r["rgb"][w] - o64["rgb"][j]).max():.2e}'
      say(row)
      s64 = o64['sigma'][j].reshape(-1)
      w64 = o64['weights'][j].reshape(-1)
      for eng in list(outs) + ['oracle32']:
        s = (r['sigma'][w] if eng == 'oracle32' else outs[eng][name]['sigma'][w]).reshape(-1).astype(np.float64)
        ww = (r['weights'][w] if eng == 'oracle32' else outs[eng][name]['weights'][w]).reshape(-1).astype(np.float64)
        k = int(np.argmax(np.abs(ww - w64)))
        ks = int(np.argmax(np.abs(s - s64)))
        say(f'    {eng:8s} max|dw| {np.abs(ww - w64).max():.2e} at {k} (w64 {w64[k]:.4f} sigma64 {s64[k]:.5f} dsigma {s[k] - s64[k]:+.2e})   '
            f'max|dsigma| {np.abs(s - s64).max():.2e} at {ks} (sigma64 {s64[ks]:.4f}, w64 {w64[ks]:.2e})')
  if args.out:
    os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
    with open(args.out, 'w') as f:
      f.write('\n'.join(lines) + '\n')


if __name__ == '__main__':
  main()
