import sys, numpy as np, torch, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
from tests.common import bench_scene, take_rays
from nerfds_b200 import synthetic as syn
from oracle.nerfds_oracle import OracleNerfModel, to_numpy
torch.set_num_threads(8)
def split16(x):
  hi = x.to(torch.float16).float(); lo = (x - hi).to(torch.float16).float(); return hi, lo
def mm_terms(terms_of):
  def mm(x, W, tag=''):
    t = terms_of(tag)
    if t == 0: return x @ W
    mx = W.abs().max().item(); e = 0
    if mx > 0:
      m, ex = np.frexp(mx); e = 3 - ex
    s = 2.0 ** e
    xh, xl = split16(x.float()); Wh, Wl = split16(W.float() * s)
    if t == 1: r = xh @ Wh
    elif t == 2: r = xh @ Wh + xl @ Wh          # activations split, weights fp16
    elif t == -2: r = xh @ Wh + xh @ Wl          # weights split, activations fp16
    else: r = xh @ Wh + xl @ Wh + xh @ Wl
    return r / s
  return mm
n = 2048
cfg, params, rays, t_rand, u = bench_scene(16384)
idx = np.concatenate([np.array([7521, 13225, 10201, 7597, 15011, 2905, 8001, 13107, 6819, 2930, 9077, 8131]), np.random.default_rng(1).choice(16384, n - 12, replace=False)])
sub = take_rays(rays, idx)
m32 = OracleNerfModel(cfg, params)
ref = to_numpy(m32.apply(sub, syn.final_extra_params(), t_rand[idx], u[idx], use_predicted_norm=True, mask_ratio=1, sharp_weights_std=0.1, return_weights=True, return_points=True, keep_internal=True, compute_sigma_gradient=False))
kw = dict(use_sample_at_infinity=cfg.use_sample_at_infinity, use_predicted_norm=True, mask_ratio=1, sharp_weights_std=0.1, compute_sigma_gradient=False)
d = torch.from_numpy(sub['directions'])
def run(terms_of, name):
  t0 = time.time()
  res = []
  for lvl in ('coarse', 'fine'):
    z = ref[lvl]['z_vals']
    pts = torch.from_numpy(sub['origins'])[:, None, :] + torch.from_numpy(z)[..., None] * d[:, None, :]
    m = OracleNerfModel(cfg, params); m.mm = mm_terms(terms_of)
    o = to_numpy(m.render_samples(lvl, pts, torch.from_numpy(z), d, d, sub['metadata'], syn.final_extra_params(), sub['mask'], **kw))
    e = np.abs(o['rgb'] - ref[lvl]['rgb']).max(-1); ed = np.abs(o['depth'] - ref[lvl]['depth'])
    en = np.abs(o['ray_norm'] - ref[lvl]['ray_norm']).max(-1)
    res.append(f'{lvl}: rgb max {e.max():.1e} p99.9 {np.percentile(e, 99.9):.1e} med {np.median(e):.1e} depth max {ed.max():.1e} norm max {en.max():.1e}')
  print(f'{name:44s} ' + ' | '.join(res) + f'  ({time.time() - t0:.0f}s)', flush=True)
nets = ['mask', 'warp', 'hyper', 'trunk', 'alpha', 'bottleneck', 'rgb']
base = lambda tag: 3
run(base, 'all 3-term (RN accumulate)')
def only(net, t):
  return lambda tag: t if tag.startswith(net) else 3
for net in ('rgb', 'bottleneck', 'hyper', 'mask', 'alpha'):
  run(only(net, 1), f'{net} 1-term')
run(lambda tag: 1 if tag.startswith(('rgb', 'bottleneck')) else 3, 'rgb+bottleneck 1-term')
for net in ('rgb', 'hyper', 'mask', 'trunk', 'warp'):
  run(only(net, 2), f'{net} 2-term (weights fp16)')
  run(only(net, -2), f'{net} 2-term (activations fp16)')
for i in range(8):
  run(only(f'trunk/hidden_{i}', 1), f'trunk/hidden_{i} 1-term')
