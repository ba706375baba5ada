import sys, numpy as np, torch, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tools')
import parity_scale as ps
from nerfds_b200 import synthetic as syn
from oracle.nerfds_oracle import OracleNerfModel, to_numpy
torch.set_num_threads(8)

def split16(x):
  hi = x.to(torch.float16).float()
  lo = (x - hi).to(torch.float16).float()
  return hi, lo

def trunc32(x64):
  # round float64 toward zero to float32
  f = x64.to(torch.float32)
  # if |f| > |x64| step one ulp toward zero
  over = f.double().abs() > x64.abs()
  fi = f.view(torch.int32)
  fi = torch.where(over, fi - 1, fi)   # decreasing the int repr moves magnitude toward zero for both signs
  return fi.view(torch.float32)

def make_mm(mode, order='interleaved', scale_w=True):
  def mm(x, W, tag=''):
    if mode == 'f32':
      return x @ W
    x = x.float(); W = W.float()
    # per-layer power-of-two weight scale: max|W| in [4,8)
    mx = W.abs().max().item()
    e = 0
    if mx > 0:
      m, ex = np.frexp(mx); e = 3 - ex
    s = 2.0 ** e
    Ws = W * s
    xh, xl = split16(x)
    Wh, Wl = split16(Ws)
    K = x.shape[-1]
    if mode == 'h3rn':
      return ((xh @ Wh + xl @ Wh + xh @ Wl) / s)
    # truncating accumulation per K=16 MMA step, in the kernel's burst order: per 64-wide K chunk: hh x4, lh x4, hl x4
    acc = torch.zeros(x.shape[:-1] + (W.shape[1],), dtype=torch.float32)
    def add(a, b):
      nonlocal acc
      p = a.double() @ b.double()
      if mode == 'h3tr':
        acc = trunc32(acc.double() + p)
      else:  # 'h3rn16': RN per step
        acc = (acc.double() + p).float()
    chunks = [(k0, min(K, k0 + 64)) for k0 in range(0, K, 64)]
    if order == 'interleaved':
      for (a, b) in chunks:
        for A_, B_ in ((xh, Wh), (xl, Wh), (xh, Wl)):
          for k in range(a, b, 16):
            add(A_[..., k:min(b, k + 16)], B_[k:min(b, k + 16)])
    elif order == 'crossfirst':
      for A_, B_ in ((xl, Wh), (xh, Wl), (xh, Wh)):
        for (a, b) in chunks:
          for k in range(a, b, 16):
            add(A_[..., k:min(b, k + 16)], B_[k:min(b, k + 16)])
    return acc / s
  return mm

if __name__ == '__main__':
  cfg, params, rays, t_rand, u = ps.scene(16384, 128, 128)
  idx = np.array([7521, 13225, 10201, 7597, 15011, 2905])
  sub = ps.take(rays, idx)
  m32 = OracleNerfModel(cfg, params)
  ref = to_numpy(m32.apply(sub, syn.final_extra_params(), t_rand[idx], u[idx], use_predicted_norm=True, mask_ratio=1, sharp_weights_std=0.1, return_weights=True, return_points=True, keep_internal=True, compute_sigma_gradient=False))
  z = ref['coarse']['z_vals']
  pts32 = torch.from_numpy(sub['origins'])[:, None, :] + torch.from_numpy(z)[..., None] * torch.from_numpy(sub['directions'])[:, None, :]
  d = torch.from_numpy(sub['directions'])
  kw = dict(use_sample_at_infinity=cfg.use_sample_at_infinity, use_predicted_norm=True, mask_ratio=1, sharp_weights_std=0.1, compute_sigma_gradient=False)
  m64 = OracleNerfModel(cfg, params, dtype=torch.float64)
  o64 = to_numpy(m64.render_samples('coarse', pts32.double(), torch.from_numpy(z).double(), d.double(), d.double(), sub['metadata'], syn.final_extra_params(), sub['mask'], **kw))
  for name, mode, order in (('f32', 'f32', ''), ('h3 RN accumulate', 'h3rn', ''), ('h3 RN per K16 step', 'h3rn16', 'interleaved'), ('h3 trunc per K16 step, interleaved', 'h3tr', 'interleaved'), ('h3 trunc, cross terms first', 'h3tr', 'crossfirst')):
    m = OracleNerfModel(cfg, params)
    m.mm = make_mm(mode, order)
    t0 = time.time()
    o = to_numpy(m.render_samples('coarse', pts32, torch.from_numpy(z), d, d, sub['metadata'], syn.final_extra_params(), sub['mask'], **kw))
    ds = o['sigma'] - o64['sigma']
    print(f'{name:40s} rgb err vs f64 {np.abs(o["rgb"] - o64["rgb"]).max(-1)}  max|dsigma| {np.abs(ds).max(-1)}  dsigma[ray0, s2] {ds[0, 2]:+.2e}  ({time.time() - t0:.1f}s)')

  print('--- selective: truncation only in the named nets (others h3 RN)')
  def mm_sel(trunc_tags, order='interleaved', cross_tags=()):
    mt = make_mm('h3tr', order); mr = make_mm('h3rn16', 'interleaved'); mc = make_mm('h3tr', 'crossfirst')
    def mm(x, W, tag=''):
      if any(tag.startswith(t) for t in cross_tags): return mc(x, W, tag)
      if any(tag.startswith(t) for t in trunc_tags): return mt(x, W, tag)
      return mr(x, W, tag)
    return mm
  allnets = ['mask', 'warp', 'hyper', 'trunk', 'alpha', 'bottleneck', 'rgb']
  cases = [(f'only {n}', [n], ()) for n in ['mask', 'warp', 'hyper', 'trunk', 'alpha']]
  cases += [(f'only trunk/hidden_{i}', [f'trunk/hidden_{i}'], ()) for i in range(8)]
  cases += [('all trunc, trunk+alpha crossfirst', allnets, ('trunk', 'alpha')), ('all trunc, trunk 4-7 + alpha crossfirst', allnets, ('trunk/hidden_4','trunk/hidden_5','trunk/hidden_6','trunk/hidden_7','alpha')),
            ('all trunc, warp crossfirst', allnets, ('warp',)), ('all crossfirst', allnets, tuple(allnets))]
  for name, tt, ct in cases:
    m = OracleNerfModel(cfg, params)
    m.mm = mm_sel(tt, cross_tags=ct)
    o = to_numpy(m.render_samples('coarse', pts32, torch.from_numpy(z), d, d, sub['metadata'], syn.final_extra_params(), sub['mask'], **kw))
    ds = o['sigma'] - o64['sigma']
    print(f'{name:44s} rgb err vs f64 {" ".join(f"{v:.1e}" for v in np.abs(o["rgb"] - o64["rgb"]).max(-1))}  max|dsigma| {" ".join(f"{v:.1e}" for v in np.abs(ds).max(-1))}')
