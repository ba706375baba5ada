"""Parity margins of the tensor-core engine vs the CPU oracle on the oracle's own samples (needs a GPU).
usage: python tools/parity_report.py [image_side]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.common import make_case, run_oracle
from nerfds_b200 import synthetic as syn
from nerfds_b200.models import NerfModel

side = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg, params, rays, t_rand, u = make_case('nerf_ds', image=side, seed=1)
ref = run_oracle(cfg, params, rays, t_rand, u, compute_sigma_gradient=False)
for prec in ('mixed', 'split3'):
  m = NerfModel(cfg, device='cuda:0', engine='tc', precision=prec)
  m.renderer.ensure_params(params)
  extra = m.renderer.make_extra(syn.final_extra_params(), use_predicted_norm=True)
  keys = list(m.renderer.level_keys(return_points=True, return_weights=True, want_target_norm=False))
  for lvl, name in ((0, 'coarse'), (1, 'fine')):
    r = ref[name]
    out = m.renderer.render_samples(lvl, r['z_vals'], rays['directions'], origins=rays['origins'],
                                    warp_id=rays['metadata']['warp'], gt_mask=rays['mask'], extra=extra,
                                    use_sample_at_infinity=cfg.use_sample_at_infinity, keys=keys)
    out = {k: v.detach().cpu().numpy() for k, v in out.items()}
    for k in ('rgb', 'depth', 'acc', 'sigma'):
      if k in out and r[k].size:
        e = np.abs(out[k].reshape(r[k].shape) - r[k]).reshape(r[k].shape[0], -1).max(1)
        print(f'{prec} {name:6s} {k:16s} rays {e.size}  max {e.max():.2e}  p99 {np.percentile(e, 99):.2e}  median {np.median(e):.2e}')
