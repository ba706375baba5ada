import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests.common import make_case, run_oracle
from nerfds_b200 import synthetic as syn
from nerfds_b200.models import NerfModel
cfg, params, rays, t_rand, u = make_case('nerf_ds', image=40, seed=1)
ref = run_oracle(cfg, params, rays, t_rand, u, compute_sigma_gradient=False)
outs = {}
for eng, prec in (('tc', 'split3'), ('simt', 'mixed')):
  m = NerfModel(cfg, device='cuda:0', engine=eng, precision=prec)
  m.renderer.ensure_params(params)
  extra = m.renderer.make_extra(syn.final_extra_params(), use_predicted_norm=True)
  keys = list(m.renderer.level_keys(return_points=True, return_weights=True, want_target_norm=False))
  r = ref['coarse']
  out = m.renderer.render_samples(0, r['z_vals'], rays['directions'], origins=rays['origins'],
                                  warp_id=rays['metadata']['warp'], gt_mask=rays['mask'], extra=extra,
                                  use_sample_at_infinity=cfg.use_sample_at_infinity, keys=keys)
  outs[eng] = {k: v.detach().cpu().numpy() for k, v in out.items()}
r = ref['coarse']
e = np.abs(outs['tc']['rgb'] - r['rgb']).max(-1)
worst = np.argsort(e)[-3:]
np.set_printoptions(precision=5, linewidth=200, suppress=False)
for w in worst:
  print('ray', w, 'rgb err tc', e[w], 'simt', np.abs(outs['simt']['rgb'][w] - r['rgb'][w]).max())
  for k in ('sigma', 'weights', 'predicted_mask'):
    a = outs['tc'][k].reshape(r[k].shape)[w].reshape(-1); b = r[k][w].reshape(-1); c = outs['simt'][k].reshape(r[k].shape)[w].reshape(-1)
    j = np.argmax(np.abs(a - b))
    print('  ', k, 'worst sample', j, 'tc', a[j], 'ref', b[j], 'simt', c[j], ' neighbourhood ref', b[max(0, j - 2):j + 3])
  a = outs['tc']['sigma'].reshape(r['sigma'].shape)[w].reshape(-1); b = r['sigma'][w].reshape(-1)
  rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-3)
  print('   max rel sigma err', rel.max(), 'at', rel.argmax(), 'sigma there', b[rel.argmax()], a[rel.argmax()])
