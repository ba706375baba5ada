#!/usr/bin/env python
"""bench.py -- rays/s of the NeRF-DS ray-marching path on N x B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): one "step" =
one 800x800 frame (640 000 rays) of a synthetic orbit through the nerf_ds.gin
networks (SE3 warp + hyper sheet + mask MLP + template NeRF, predicted
normals) at 128 coarse + 128 fine samples ("256 samples/ray"), stratified
draws, render-mode outputs (render.py:192-193).  At N > 1 every step renders N
frames (--scaling weak, per-GPU work fixed; --scaling strong: one frame), each
block-partitioned over the N ranks (utils.shard order) and reassembled inside
the compositing kernel by peer-memory stores (--gather nccl: one all-gather).

Prints ONE JSON line (rank 0).  `value` = rays/s with the rays resident in HBM;
`e2e` = the same frames from pinned host rays to the frame in pinned host
memory (N = 1: one ndsr_render_rays_host_rng call; N > 1: per-rank upload,
rank 0 reads the assembled frame back).  The uniform draws are generated on
the device from jax keys inside the timed region in both.
`--impl reference` times the CPU oracle (the restatement of the reference's
JAX path; jax itself is not installable here) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
  sys.path.insert(0, ROOT)

METRIC = 'rays/sec at 256 samples/ray (warp+template MLP)'
UNIT = 'rays/s'

# MACs per network evaluation at nerf_ds.gin widths
MAC_COARSE = 729_216      # mask + SE3 + hyper sheet + trunk + sigma head (render mode, coarse level)
MAC_FINE = 867_584        # + bottleneck + rgb branch + normal head
MAC_WARP_TEMPLATE = 715_136


def flops_per_ray(sc, sf):
  return 2.0 * (sc * MAC_COARSE + (sc + sf) * MAC_FINE)


def peaks():
  p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
  if os.path.exists(p):
    with open(p) as f:
      d = json.load(f)
    return d, 'measured'
  return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler(threading.Thread):
  """Samples SM clocks / throttle reasons with NVML while the timed region runs."""

  def __init__(self, index):
    super().__init__(daemon=True)
    self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
    except Exception:
      self.nv = None

  def run(self):
    if self.nv is None:
      return
    nv = self.nv
    names = {nv.nvmlClocksThrottleReasonHwSlowdown: 'hw_slowdown',
             nv.nvmlClocksThrottleReasonHwThermalSlowdown: 'hw_thermal_slowdown',
             nv.nvmlClocksThrottleReasonSwThermalSlowdown: 'sw_thermal_slowdown',
             nv.nvmlClocksThrottleReasonSwPowerCap: 'sw_power_cap',
             nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: 'hw_power_brake'}
    while not self.stop_flag:
      try:
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, name in names.items():
          if r & bit:
            self.reasons.add(name)
      except Exception:
        pass
      time.sleep(0.1)

  def result(self):
    self.stop_flag = True
    if self.nv is None or not self.samples:
      return {'sm_mhz': None, 'sm_max_mhz': self.max_mhz, 'reasons': ['nvml unavailable']}
    return {'sm_mhz': float(np.median(self.samples)), 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons)}


# ---------------------------------------------------------------------------
def build_case(args):
  from nerfds_b200 import synthetic as syn
  from nerfds_b200.config import nerf_ds_config
  from nerfds_b200.params import init_params
  cfg = nerf_ds_config(num_coarse_samples=args.coarse, num_fine_samples=args.fine, near=0.1, far=2.5,
                       num_warp_embeds=100)
  params = init_params(cfg, 0)
  return cfg, params, syn


def frame_inputs(syn, args, frame):
  rays = syn.frame_rays(args.image, args.image, frame=frame % 30, num_frames=30, focal=float(args.image))
  return (rays['origins'], rays['directions'], rays['metadata']['warp'].reshape(-1).astype(np.uint32))


def cpu_sample(args, cfg, params, syn, n_rays, threads, reps=1, warm=True):
  """Times the oracle on the first n_rays of frame 0 (same nets, samples and draws)."""
  import torch
  from oracle.nerfds_oracle import OracleNerfModel
  torch.set_num_threads(threads)
  o, d, w = frame_inputs(syn, args, 0)
  sel = np.linspace(0, o.shape[0] - 1, n_rays).astype(np.int64)       # spread over the frame (empty + opaque rays)
  rays = {'origins': o[sel], 'directions': d[sel], 'metadata': {'warp': w[sel].reshape(-1, 1)},
          'mask': np.zeros((n_rays, 1), np.float32)}
  t_rand, u = syn.uniform_draws(n_rays, cfg.num_coarse_samples, cfg.num_fine_samples, 0)
  m = OracleNerfModel(cfg, params)
  run = lambda nr: m.apply({k: (v[:nr] if not isinstance(v, dict) else {a: b[:nr] for a, b in v.items()})
                            for k, v in rays.items()}, syn.final_extra_params(), t_rand[:nr], u[:nr],
                           use_predicted_norm=True, mask_ratio=1, sharp_weights_std=0.1)
  if warm:
    run(min(64, n_rays))
  t0 = time.perf_counter()
  for _ in range(reps):
    run(n_rays)
  dt = (time.perf_counter() - t0) / reps
  return n_rays / dt, dt


def run_reference(args):
  rank = int(os.environ.get('RANK', '0'))
  if rank != 0:
    return
  cfg, params, syn = build_case(args)
  threads = os.cpu_count() or 1
  n = args.cpu_rays
  cpu_sample(args, cfg, params, syn, min(64, n), threads, warm=False)     # warm-up (untimed)
  for _ in range(max(0, args.warmup - 1)):
    cpu_sample(args, cfg, params, syn, min(64, n), threads, warm=False)
  t0 = time.perf_counter()
  for _ in range(args.steps):
    cpu_sample(args, cfg, params, syn, n, threads, warm=False)
  dt = time.perf_counter() - t0
  v = n * args.steps / dt
  sample = (f'{n} rays spread over one {args.image}x{args.image} frame per step, {args.coarse}+{args.fine} samples, '
            f'PyTorch-CPU fp32 restatement of the reference JAX path incl. autograd d(sigma)/dx (oracle/nerfds_oracle.py); '
            f'jax/flax are not installable in this image')
  line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
          'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
          'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
          'config': dict(workload_config(args, 1, 1), timed_rays_per_step=n, engine='cpu-oracle', precision='f32',
                         note=f'each timed step is a bounded sample of the workload: {n} of the frame\'s {args.image * args.image} rays '
                              '(value = sample rays / sample time)'),
          'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
          'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
          'gpu_launches': 0}
  print(json.dumps(line), flush=True)


def workload_config(args, n_frames, world, rays_per_step=None):
  return {'workload': f'single {args.image}x{args.image} frame render, nerf_ds.gin nets (SE3 warp + hyper sheet + mask MLP '
                      f'+ template NeRF, predicted normals), {args.coarse}+{args.fine} coarse/fine stratified samples',
          'rays_per_step': args.image * args.image * n_frames if rays_per_step is None else rays_per_step,
          'frames_per_step': n_frames, 'chunk_rays': args.chunk,
          'parallelism': (f'rays block-sharded over {world} GPU(s), ' + (
              'frame reassembled by peer-memory stores of the compositing kernel (no data-path collective, 1 barrier/frame)'
              if getattr(args, 'gather', 'peer') == 'peer' else '1 NCCL all-gather/frame')) if world > 1 else 'single GPU',
          'l2': 'per-step inputs + per-sample scratch (>1 GB) exceed the 126 MB L2; a 256 MB buffer is also '
                'rewritten between steps'}


# ---------------------------------------------------------------------------
def traffic_per_eval():
  """DRAM bytes per fine-level sample evaluation, from the committed ncu capture of this command (profiles/):
  (dram__bytes_read.sum + dram__bytes_write.sum) of the fine-level field launches / evaluations."""
  p = os.path.join(ROOT, 'profiles', 'r2_traffic.json')
  if os.path.exists(p):
    with open(p) as f:
      d = json.load(f)
    return float(d['dram_bytes_per_eval']), d.get('source', 'profiles/r2_traffic.json')
  return None, None


def run_ours(args):
  import torch
  import torch.distributed as dist
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if not torch.cuda.is_available():
    raise SystemExit('bench.py needs a CUDA device: the product path has no CPU fallback')
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  cfg, params, syn = build_case(args)
  # CPU baseline: rank 0 of a single-GPU run only, before anything else competes for the host cores (under torchrun
  # the other ranks would be spinning in a barrier and OMP_NUM_THREADS is 1)
  cpu = None
  if world == 1 and not args.no_cpu:
    threads = os.cpu_count() or 1
    v, dt = cpu_sample(args, cfg, params, syn, args.cpu_rays, threads)
    cpu = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
           'sample': f'{args.cpu_rays} rays spread over frame 0, {args.coarse}+{args.fine} samples, {dt:.1f} s, PyTorch-CPU fp32 '
                     f'restatement of the reference JAX path incl. autograd d(sigma)/dx (jax not installable here)'}
  if world > 1 and not dist.is_initialized():
    dist.init_process_group('nccl', device_id=dev)
  from nerfds_b200 import jax_random as jr
  from nerfds_b200.models import NerfModel
  from nerfds_b200.renderer import RENDER_KEYS, ALL_SHAPES
  from nerfds_b200.evaluation import all_gather_level
  model = NerfModel(cfg, device=dev, engine=args.engine, precision=args.precision)
  R = model.renderer
  R.load_params(params)
  R.set_max_chunk(args.chunk)
  extra = R.make_extra(syn.final_extra_params(), mask_ratio=1.0, sharp_weights_std=0.1, use_predicted_norm=True)
  Sc, Sf = cfg.num_coarse_samples, cfg.num_fine_samples
  n_frame = args.image * args.image
  per = (n_frame + world - 1) // world
  bounds = [(r * per, min(n_frame, (r + 1) * per)) for r in range(world)]
  lo, hi = bounds[rank]
  n_frames = world if args.scaling == 'weak' else 1     # frames per step; every frame is sharded over all ranks

  # ---- inputs.  Rays: device-resident copy (`value`) and pinned host copy (`e2e`).  The two uniform draws of the path
  # are generated on the device from jax keys inside the timed region, as the reference does inside its jitted call
  # (model_utils.py:84, 217; one key pair per device and frame, evaluation.py:81-84).
  def frame_keys(f, r):
    ks = jr.split(jr.fold_in(jr.PRNGKey(20230601), f), 2 * world)
    return np.asarray(ks[2 * r], np.uint32), np.asarray(ks[2 * r + 1], np.uint32)

  frames = []
  for f in range(n_frames):
    o, d, w = frame_inputs(syn, args, f)
    blk = lambda a: np.ascontiguousarray(a[lo:hi])
    ho, hd, hw = blk(o), blk(d), blk(w)
    frames.append({
        'o': torch.from_numpy(ho).to(dev), 'd': torch.from_numpy(hd).to(dev), 'w': torch.from_numpy(hw.view(np.int32)).to(dev),
        'ho': torch.from_numpy(ho).pin_memory(), 'hd': torch.from_numpy(hd).pin_memory(),
        'hw': torch.from_numpy(hw.view(np.int32)).pin_memory(), 'keys': frame_keys(f, rank)})
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

  # frame reassembly at N > 1: the compositing kernels store every rank's rays into every GPU's frame buffer
  # (peer.PeerFrames, NVLink stores; two buffers alternate so a frame can be read while the next one is written);
  # --gather nccl keeps the one all-gather per frame instead
  peer = None
  if world > 1 and args.gather == 'peer':
    from nerfds_b200.peer import PeerFrames
    why = ''
    try:
      peer = [PeerFrames(R, n_frame, RENDER_KEYS), PeerFrames(R, n_frame, RENDER_KEYS)]
    except RuntimeError as e:        # no peer mapping between these GPUs: every rank must take the same path
      peer, why = None, str(e)
    ok = torch.tensor([0 if peer is None else 1], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 0:
      if peer is not None:
        R._check(R.lib.ndsr_set_output_mirrors(R._h, 0, None, None, 0), 'ndsr_set_output_mirrors')
        peer = None
      args.gather = 'nccl'
      sys.stderr.write(f'[bench] rank {rank}: peer-memory reassembly unavailable ({why or "another rank failed"}); '
                       'using the NCCL all-gather\n')

  def render_shard(o, d, w, keys, fine_ptrs=None):
    n = o.shape[0]
    t = R.random_uniform(keys[0], n, Sc)
    uu = R.random_uniform(keys[1], n, Sf)
    return R.render_rays(o, d, warp_id=w, t_rand=t, u=uu, extra=extra, coarse_keys=(), fine_keys=RENDER_KEYS,
                         fine_ptrs=fine_ptrs)['fine']

  def render_frame(i, o, d, w, keys):
    if peer is not None:
      pf = peer[i & 1]
      pf.activate()
      render_shard(o, d, w, keys, fine_ptrs=pf.shard_ptrs(lo))
      pf.wait()
      return pf.frame()
    out = render_shard(o, d, w, keys)
    return all_gather_level(out) if world > 1 else out

  def step_device():
    outs = None
    for i, fr in enumerate(frames):
      outs = render_frame(i, fr['o'], fr['d'], fr['w'], fr['keys'])
    flush.fill_(1)
    return outs

  host_out = {}
  if world == 1:                    # pinned destination of ndsr_render_rays_host_rng
    for i in range(len(frames)):
      host_out[i] = {k: torch.empty((hi - lo,) + tuple(ALL_SHAPES[k](Sc + Sf, R.H)), dtype=torch.float32).pin_memory().numpy()
                     for k in RENDER_KEYS}

  def step_host():
    """End to end: pinned host rays in, the frame in pinned host memory out (rank 0), every step."""
    res = None
    for i, fr in enumerate(frames):
      if world == 1:
        # one C-ABI call: chunked, input copies of chunk k + 1 overlap the compute of chunk k, draws generated on the device
        res = R.render_rays_host(fr['ho'].numpy(), fr['hd'].numpy(), warp_id=fr['hw'].numpy().view(np.uint32),
                                 extra=extra, fine_keys=RENDER_KEYS, out=host_out[i], rng_keys=fr['keys'])
      else:
        o, d, w = (fr[k].to(dev, non_blocking=True) for k in ('ho', 'hd', 'hw'))
        out = render_frame(i, o, d, w, fr['keys'])
        if rank == 0:               # like render.py: process 0 keeps the image
          hb = host_out.setdefault(i, {})
          for k, v in out.items():
            if k not in hb:
              hb[k] = torch.empty(v.shape, dtype=v.dtype).pin_memory()
            hb[k].copy_(v, non_blocking=True)
          res = hb
        torch.cuda.current_stream().synchronize()
    flush.fill_(1)
    return res

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  def timed(fn, steps, profile=False):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = R.kernel_launches
    if profile:
      R.profile_enable(True)
    e0.record()
    for _ in range(steps):
      fn()
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    launches = torch.tensor([R.kernel_launches - l0], device=dev, dtype=torch.int64)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
      dist.all_reduce(launches, op=dist.ReduceOp.SUM)
    prof = R.profile_read() if profile else None
    if profile:
      R.profile_enable(False)
    return float(ms.item()), int(launches.item()), prof

  for _ in range(args.warmup):
    step_device()
  sampler = ClockSampler(local)
  sampler.start()
  ms, launches, prof = timed(step_device, args.steps, profile=True)
  clocks = sampler.result()
  rays_step = n_frame * n_frames      # whole job
  value = rays_step * args.steps / (ms * 1e-3)

  # ---- end to end (host buffers), timed over the same number of steps
  for _ in range(min(args.warmup, 2)):
    step_host()
  ms_h, _, _ = timed(step_host, args.steps)
  e2e_value = rays_step * args.steps / (ms_h * 1e-3)
  fr = frames[0]
  h2d = sum(fr[k].numel() * fr[k].element_size() for k in ('ho', 'hd', 'hw')) * world * n_frames
  per_ray_out = sum(int(np.prod(ALL_SHAPES[k](Sc + Sf, R.H) or (1,))) for k in RENDER_KEYS) * 4
  d2h = per_ray_out * n_frame * n_frames          # the assembled frame(s), once (rank 0)

  # ---- early-termination scan (ndsr_set_early_termination; OFF in `value` and `e2e` above, which evaluate every
  # sample like the reference): the same frame with the scan on, and a field with opaque regions (raw densities
  # x 40) with the scan off / on -- rays/s and the share of the fine level's network evaluations that were skipped
  term = None
  if world == 1 and R.engine == 'tc' and args.term_eps > 0 and not args.sweep:
    from nerfds_b200.params import harden_density
    ts = max(1, min(args.steps, 2))

    def term_run(eps):
      R.set_early_termination(eps)
      step_device()
      R.termination_stats(reset=True)
      ms_t, _, _ = timed(step_device, ts)
      ev, seen = R.termination_stats(reset=True)
      skipped = (seen - ev) / float(max(seen, 1)) * Sf / float(Sc + Sf) if eps > 0 else 0.0
      return {'value': rays_step * ts / (ms_t * 1e-3), 'evals_skipped_frac': skipped}

    term = {'transmittance_eps': args.term_eps, 'steps': ts, 'bench_scene': term_run(args.term_eps)}
    R.load_params(harden_density(params, 40.0))
    term['opaque_scene'] = {'off': term_run(0.0), 'on': term_run(args.term_eps),
                            'what': 'same networks, raw density (column 0 of both alpha heads) x 40: hard surfaces'}
    term['opaque_scene']['speedup'] = term['opaque_scene']['on']['value'] / term['opaque_scene']['off']['value']
    R.set_early_termination(0.0)
    R.load_params(params)

  # ---- precision='mixed' beside the headline (which stays split3): the rgb branch as ONE fp16 term instead of three
  # (2.48x instead of 2.54x the algorithmic MACs; strict parity at 16 384 rays: fine rgb max 4.2e-4 instead of 3.2e-4,
  # profiles/r2_parity_scale_mixed.txt)
  mixed = None
  if world == 1 and R.engine == 'tc' and args.precision == 'split3' and not args.sweep and not args.no_mixed:
    m2 = NerfModel(cfg, device=dev, engine=args.engine, precision='mixed')
    R2 = m2.renderer
    R2.load_params(params)
    R2.set_max_chunk(args.chunk)

    def step_mixed():
      for fr in frames:
        t = R2.random_uniform(fr['keys'][0], fr['o'].shape[0], Sc)
        uu = R2.random_uniform(fr['keys'][1], fr['o'].shape[0], Sf)
        R2.render_rays(fr['o'], fr['d'], warp_id=fr['w'], t_rand=t, u=uu, extra=extra, coarse_keys=(), fine_keys=RENDER_KEYS)
      flush.fill_(1)

    step_mixed()
    ts = max(1, min(args.steps, 2))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(ts):
      step_mixed()
    e1.record()
    barrier()
    im2 = R2.tc_issued_macs()
    mixed = {'value': rays_step * ts / (e0.elapsed_time(e1) * 1e-3), 'steps': ts,
             'issued_over_algorithmic': (Sc * im2[(0, 'sigma')] + Sf * im2[(1, 'full')] + Sc * im2[(1, 'carried')]) /
                                        float(Sc * MAC_COARSE + (Sc + Sf) * MAC_FINE),
             'what': "precision='mixed': rgb branch as one fp16 term (not the headline)"}
    R2.close()

  # ---- N > 1: the last assembled frame against a single-GPU render of the same rays, shards and keys
  frame_check = None
  if world > 1:
    last = n_frames - 1
    got = {k: v.clone() for k, v in step_device().items()}
    barrier()
    if rank == 0:
      o, d, w = frame_inputs(syn, args, last)
      same = True
      if peer is not None:
        R._check(R.lib.ndsr_set_output_mirrors(R._h, 0, None, None, 0), 'ndsr_set_output_mirrors')
      for r, (a, b) in enumerate(bounds):
        ref = render_shard(torch.from_numpy(np.ascontiguousarray(o[a:b])).to(dev), torch.from_numpy(np.ascontiguousarray(d[a:b])).to(dev),
                           torch.from_numpy(np.ascontiguousarray(w[a:b]).view(np.int32)).to(dev), frame_keys(last, r))
        for k, v in ref.items():
          same = same and bool(torch.equal(v.reshape(b - a, -1), got[k].reshape(n_frame, -1)[a:b]))
      frame_check = same
    barrier()

  if rank == 0:
    pk, pk_kind = peaks()
    # dominant kernel: the fine-level field kernel; algorithmic FLOPs per launch / measured duration
    f_ms, f_n = prof['field_fine']
    c_ms, c_n = prof['field_coarse']
    rays_local_total = (hi - lo) * n_frames * args.steps
    fine_flops = 2.0 * MAC_FINE * (Sc + Sf) * rays_local_total
    coarse_flops = 2.0 * MAC_COARSE * Sc * rays_local_total
    ach = fine_flops / (f_ms * 1e-3) / 1e12 if f_ms > 0 else 0.0
    peak = float(pk.get('bf16_tflops_sustained') or pk.get('bf16_tflops') or 1400.0)
    burst = pk.get('bf16_tflops', peak)
    tpe, tsrc = traffic_per_eval()
    if args.traffic is not None:
      traffic, tsrc = args.traffic, '--traffic'
    else:
      traffic = tpe * (Sc + Sf) * rays_local_total / max(f_n, 1) if (tpe is not None and R.engine == 'tc') else None
    issued = None
    if R.engine == 'tc':
      im = R.tc_issued_macs()
      issued = {'coarse_per_eval': im[(0, 'sigma')], 'fine_new_per_eval': im[(1, 'full')], 'fine_carried_per_eval': im[(1, 'carried')],
                'per_ray': Sc * im[(0, 'sigma')] + Sf * im[(1, 'full')] + Sc * im[(1, 'carried')],
                'algorithmic_per_ray': Sc * MAC_COARSE + (Sc + Sf) * MAC_FINE}
      issued['issued_over_algorithmic'] = issued['per_ray'] / issued['algorithmic_per_ray']
    roofline = {'bound': 'tensor', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak,
                'traffic': traffic, 'traffic_source': tsrc,
                'kernel': (f'field_tc_kernel x2 (fine level: {Sf} new depths per ray through every network + {Sc} coarse depths '
                           'through the template NeRF on carried warp/hyper/mask results)') if R.engine == 'tc'
                else f'field_{R.engine}_kernel (fine level)',
                'peak_kind': f'{pk_kind} sustained bf16 (burst {burst})',
                'launches': f_n, 'avg_launch_ms': f_ms / max(f_n, 1),
                'flops_per_launch': fine_flops / max(f_n, 1),
                'coarse_level_frac': (coarse_flops / (c_ms * 1e-3) / 1e12 / peak) if c_ms > 0 else None,
                'whole_step_tflops': (fine_flops + coarse_flops) * world / (ms * 1e-3) / 1e12,
                'issued_macs': issued,
                'stage_ms_per_step': {k: v[0] / args.steps for k, v in prof.items()},
                'hbm_gbs_of_peak': None}
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
            'dtype': 'f32 (tensor-core layers: split-fp16 operands, fp32 accumulate)' if R.engine == 'tc' else 'f32',
            'data': 'synthetic', 'config': dict(workload_config(args, n_frames, world), engine=R.engine, precision=args.precision),
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'ms_per_step': ms_h / args.steps, 'steps': args.steps,
                    'draws': 'generated on the device from jax keys inside the timed region (threefry2x32)'},
            'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu}
    if frame_check is not None:
      line['frame_matches_single_gpu'] = frame_check
    if term is not None:
      line['early_termination'] = term
    if mixed is not None:
      line['mixed_precision'] = mixed
    print(json.dumps(line), flush=True)
  if peer is not None:
    for pf in peer:
      pf.close()
  model.renderer.close()
  if world > 1:
    dist.barrier()
    if not getattr(args, 'keep_group', False):
      dist.destroy_process_group()


# ---------------------------------------------------------------------------
# BASELINE configs[2]: train.py ray batch of 4096, nerf_ds.gin (64+64), surface-aware branch + mask: the training
# FORWARD with every level-dict key training.py consumes, incl. target_norm = R normalize(-d sigma / dx) on both levels
MAC_TRAIN = 1_470_720     # forward + reverse sweep


def train_case(args):
  from nerfds_b200 import synthetic as syn
  from nerfds_b200.config import nerf_ds_config
  from nerfds_b200.params import init_params
  cfg = nerf_ds_config(num_coarse_samples=args.coarse, num_fine_samples=args.fine, near=0.1, far=2.5, num_warp_embeds=100)
  params = init_params(cfg, 0)
  rays = syn.train_batch(args.batch, cfg.num_warp_embeds, seed=3)
  t_rand, u = syn.uniform_draws(args.batch, args.coarse, args.fine, 3)
  return cfg, params, syn, rays, t_rand, u


def train_cpu(args, cfg, params, syn, rays, t_rand, u, n, threads):
  import torch
  from oracle.nerfds_oracle import OracleNerfModel
  torch.set_num_threads(threads)
  m = OracleNerfModel(cfg, params)
  sub = {'origins': rays['origins'][:n], 'directions': rays['directions'][:n],
         'metadata': {'warp': rays['metadata']['warp'][:n]}, 'mask': rays['mask'][:n]}
  run = lambda: m.apply(sub, syn.final_extra_params(), t_rand[:n], u[:n], use_predicted_norm=True, mask_ratio=0.7,
                        sharp_weights_std=0.1, return_points=True, return_weights=True, compute_sigma_gradient=True)
  t0 = time.perf_counter()
  run()
  return n / (time.perf_counter() - t0)


def train_config(args, n=None):
  return {'workload': f'train.py ray batch {args.batch}, nerf_ds.gin ({args.coarse}+{args.fine} samples), training forward with every '
                      'level-dict key incl. target_norm (surface-aware branch, mask MLP) on both levels (BASELINE configs[2])',
          'rays_per_step': args.batch if n is None else n,
          'l2': 'a 256 MB buffer is rewritten between steps (the batch itself fits the L2)'}


def run_train(args):
  import torch
  if args.impl == 'reference':
    if int(os.environ.get('RANK', '0')) != 0:
      return
    cfg, params, syn, rays, t_rand, u = train_case(args)
    threads = os.cpu_count() or 1
    n = args.cpu_rays
    train_cpu(args, cfg, params, syn, rays, t_rand, u, min(32, n), threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
      train_cpu(args, cfg, params, syn, rays, t_rand, u, n, threads)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    print(json.dumps({'impl': 'reference', 'metric': METRIC_TRAIN, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
                      'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak',
                      'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                      'config': dict(train_config(args), timed_rays_per_step=n, engine='cpu-oracle', precision='f32',
                                     note=f'each timed step is a bounded sample: {n} of the {args.batch} rays'),
                      'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
                                       'sample': f'{n} rays of the batch per step, PyTorch-CPU fp32 restatement incl. autograd d(sigma)/dx'},
                      'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}), flush=True)
    return
  world = int(os.environ.get('WORLD_SIZE', '1'))
  rank = int(os.environ.get('RANK', '0'))
  local = int(os.environ.get('LOCAL_RANK', '0'))
  if not torch.cuda.is_available():
    raise SystemExit('bench.py needs a CUDA device: the product path has no CPU fallback')
  torch.cuda.set_device(local)
  dev = torch.device('cuda', local)
  cfg, params, syn, rays, t_rand, u = train_case(args)
  cpu = None
  if world == 1 and not args.no_cpu:
    threads = os.cpu_count() or 1
    n = min(args.cpu_rays, args.batch)
    train_cpu(args, cfg, params, syn, rays, t_rand, u, min(32, n), threads)
    cpu = {'value': train_cpu(args, cfg, params, syn, rays, t_rand, u, n, threads), 'unit': UNIT, 'cores': threads, 'kind': 'port',
           'sample': f'{n} rays of the batch, PyTorch-CPU fp32 restatement of the reference JAX path incl. autograd d(sigma)/dx'}
  if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=dev)       # replicas: every rank runs its own batch (data parallel, no exchange in the forward)
  from nerfds_b200.models import NerfModel
  model = NerfModel(cfg, device=dev, engine=args.engine, precision=args.precision)
  R = model.renderer
  R.load_params(params)
  extra = R.make_extra(syn.final_extra_params(), mask_ratio=0.7, sharp_weights_std=0.1, use_predicted_norm=True)
  keys = R.level_keys(return_points=True, return_weights=True, want_target_norm=True)
  Sc, Sf = cfg.num_coarse_samples, cfg.num_fine_samples
  d_in = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in
          (('o', rays['origins']), ('d', rays['directions']), ('m', rays['mask'].reshape(-1)), ('t', t_rand), ('u', u))}
  d_in['w'] = torch.from_numpy(rays['metadata']['warp'].reshape(-1).astype(np.uint32).view(np.int32)).to(dev)
  h_in = {k: v.cpu().pin_memory() for k, v in d_in.items()}
  flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
  h_out = {}

  def step_device():
    out = R.render_rays(d_in['o'], d_in['d'], warp_id=d_in['w'], gt_mask=d_in['m'], t_rand=d_in['t'], u=d_in['u'], extra=extra,
                        coarse_keys=keys, fine_keys=keys)
    flush.fill_(1)
    return out

  def step_host():
    g = {k: v.to(dev, non_blocking=True) for k, v in h_in.items()}
    out = R.render_rays(g['o'], g['d'], warp_id=g['w'], gt_mask=g['m'], t_rand=g['t'], u=g['u'], extra=extra,
                        coarse_keys=keys, fine_keys=keys)
    for lvl in ('coarse', 'fine'):                       # what the loss reads per ray; the per-sample keys stay on the device
      for k in ('rgb', 'ray_predicted_mask'):
        dst = h_out.setdefault((lvl, k), torch.empty(out[lvl][k].shape, dtype=torch.float32).pin_memory())
        dst.copy_(out[lvl][k], non_blocking=True)
    torch.cuda.current_stream().synchronize()
    flush.fill_(1)

  def timed(fn, steps, profile=False):
    torch.cuda.synchronize()
    if world > 1:
      dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = R.kernel_launches
    if profile:
      R.profile_enable(True)
    e0.record()
    for _ in range(steps):
      fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
      dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    prof = R.profile_read() if profile else None
    if profile:
      R.profile_enable(False)
    return float(ms.item()), (R.kernel_launches - l0) * world, prof

  for _ in range(args.warmup):
    step_device()
  sampler = ClockSampler(local)
  sampler.start()
  ms, launches, prof = timed(step_device, args.steps, profile=True)
  clocks = sampler.result()
  for _ in range(min(args.warmup, 2)):
    step_host()
  ms_h, _, _ = timed(step_host, args.steps)
  if rank == 0:
    pk, pk_kind = peaks()
    peak = float(pk.get('bf16_tflops_sustained') or pk.get('bf16_tflops') or 1400.0)
    f_ms = prof['field_coarse'][0] + prof['field_fine'][0]
    f_n = prof['field_coarse'][1] + prof['field_fine'][1]
    flops = 2.0 * MAC_TRAIN * (2 * Sc + Sf) * args.batch * args.steps
    ach = flops / (f_ms * 1e-3) / 1e12 if f_ms > 0 else 0.0
    issued = None
    if R.engine == 'tc':
      im = R.tc_issued_macs()
      issued = {'coarse_per_eval': im[(0, 'grad')], 'fine_per_eval': im[(1, 'grad')], 'algorithmic_per_eval': MAC_TRAIN}
    h2d = sum(v.numel() * v.element_size() for v in h_in.values())
    d2h = 2 * args.batch * 4 * 4
    line = {'metric': METRIC_TRAIN, 'value': args.batch * world * args.steps / (ms * 1e-3), 'unit': UNIT, 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 (tensor-core layers: split-fp16 operands, fp32 accumulate)' if R.engine == 'tc' else 'f32',
            'data': 'synthetic', 'config': dict(train_config(args), engine=R.engine, precision=args.precision,
                                                parallelism='single GPU' if world == 1 else f'{world} independent replicas of the batch'),
            'clocks': clocks,
            'e2e': {'value': args.batch * world * args.steps / (ms_h * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d) * world,
                    'd2h_bytes_per_step': int(d2h) * world, 'ms_per_step': ms_h / args.steps, 'steps': args.steps},
            'gpu_launches': launches,
            'roofline': {'bound': 'tensor', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak, 'traffic': None,
                         'kernel': f'field_{R.engine}_kernel (both levels: forward + reverse sweep for d sigma / dx)',
                         'peak_kind': f'{pk_kind} sustained bf16', 'launches': f_n, 'avg_launch_ms': f_ms / max(f_n, 1),
                         'flops_per_launch': flops / max(f_n, 1), 'issued_macs': issued,
                         'stage_ms_per_step': {k: v[0] / args.steps for k, v in prof.items()}},
            'cpu_baseline': cpu}
    print(json.dumps(line), flush=True)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


METRIC_TRAIN = 'rays/sec, training forward of a 4096-ray batch (all level-dict keys incl. target_norm)'


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=3)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--engine', default='auto', choices=['auto', 'tc', 'simt'])
  ap.add_argument('--precision', default='split3', choices=['mixed', 'fp16', 'split3'])
  ap.add_argument('--image', type=int, default=800)
  ap.add_argument('--coarse', type=int, default=None)
  ap.add_argument('--fine', type=int, default=None)
  ap.add_argument('--chunk', type=int, default=65536)
  ap.add_argument('--cpu-rays', type=int, default=None)
  ap.add_argument('--no-cpu', action='store_true')
  ap.add_argument('--gather', default='peer', choices=['peer', 'nccl'], help='frame reassembly at N > 1')
  ap.add_argument('--sweep', default=None, help="sample-count sweep, e.g. '32+32,64+64,128+128,256+256' (BASELINE configs[4])")
  ap.add_argument('--workload', default='frame', choices=['frame', 'train4096'],
                  help='frame: BASELINE configs[1] (the headline); train4096: configs[2], the training forward of a ray batch')
  ap.add_argument('--batch', type=int, default=4096)
  ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                  help='weak: N frames per step (per-GPU work fixed); strong: one frame per step over N ranks')
  ap.add_argument('--no-mixed', action='store_true', help="skip the extra precision='mixed' measurement")
  ap.add_argument('--term-eps', type=float, default=1e-4,
                  help='also measure the fine level with the early-termination scan at this transmittance (0: skip)')
  ap.add_argument('--traffic', type=float, default=None, help='DRAM bytes/launch of the dominant kernel from ncu')
  args = ap.parse_args()
  train = args.workload == 'train4096'
  if args.coarse is None:
    args.coarse = 64 if train else 128        # nerf_ds.gin's literal 64+64 for the train batch, 128+128 for the headline
  if args.fine is None:
    args.fine = 64 if train else 128
  if args.cpu_rays is None:
    args.cpu_rays = (256 if train else 1024) if args.impl == 'ours' else (256 if train else 512)
  if train:
    if args.steps == 3 and '--steps' not in sys.argv:
      args.steps = 50
    return run_train(args)
  world = int(os.environ.get('WORLD_SIZE', '1'))
  if world != args.gpus and args.impl == 'ours':
    if world == 1 and args.gpus > 1:
      raise SystemExit(f'--gpus {args.gpus} needs torchrun (python -m torch.distributed.run --nproc-per-node {args.gpus} ...)')
  if args.impl == 'reference':
    run_reference(args)
  elif args.sweep:
    # BASELINE configs[4]: sample-count sweep, one JSON line per point, one process group for all of them
    points = [tuple(int(v) for v in p.split('+')) for p in args.sweep.split(',')]
    for i, (c, f) in enumerate(points):
      args.coarse, args.fine = c, f
      args.keep_group = i + 1 < len(points)
      args.no_cpu = True
      run_ours(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
