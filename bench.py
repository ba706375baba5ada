#!/usr/bin/env python
"""bench.py -- rays/s of the NeRF-DS ray-marching path on N x B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): one "step" =
one 800x800 frame (640 000 rays) of a synthetic orbit through the nerf_ds.gin
networks (SE3 warp + hyper sheet + mask MLP + template NeRF, predicted
normals) at 128 coarse + 128 fine samples ("256 samples/ray"), stratified
draws, render-mode outputs (render.py:192-193).  At N > 1 every step renders N
frames, each block-partitioned over the N ranks (utils.shard order) and
reassembled with one NCCL all-gather per frame: per-GPU work is fixed (weak).

Prints ONE JSON line (rank 0).  `value` = rays/s with inputs resident in HBM;
`e2e` = the same frames through the host-buffer C-ABI call
(ndsr_render_rays_host: pinned host rays in, pinned host image out).
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
          'config': workload_config(args, 1),
          'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': sample},
          'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
          'gpu_launches': 0}
  print(json.dumps(line), flush=True)


def workload_config(args, world):
  return {'workload': f'single {args.image}x{args.image} frame render, nerf_ds.gin nets (SE3 warp + hyper sheet + mask MLP '
                      f'+ template NeRF, predicted normals), {args.coarse}+{args.fine} coarse/fine stratified samples',
          'rays_per_step': args.image * args.image * world, 'chunk_rays': args.chunk,
          'parallelism': (f'rays block-sharded over {world} GPU(s), ' + (
              'frame reassembled by peer-memory stores of the compositing kernel (no data-path collective, 1 barrier/frame)'
              if getattr(args, 'gather', 'peer') == 'peer' else '1 NCCL all-gather/frame')) if world > 1 else 'single GPU',
          'l2': 'per-step inputs + per-sample scratch (>1 GB) exceed the 126 MB L2; a 256 MB buffer is also '
                'rewritten between steps'}


# ---------------------------------------------------------------------------
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
  if world > 1:
    dist.init_process_group('nccl', device_id=dev)
  from nerfds_b200.models import NerfModel
  from nerfds_b200.renderer import RENDER_KEYS
  from nerfds_b200.evaluation import all_gather_level
  cfg, params, syn = build_case(args)
  model = NerfModel(cfg, device=dev, engine=args.engine, precision=args.precision)
  R = model.renderer
  R.load_params(params)
  R.set_max_chunk(args.chunk)
  extra = R.make_extra(syn.final_extra_params(), mask_ratio=1.0, sharp_weights_std=0.1, use_predicted_norm=True)
  Sc, Sf = cfg.num_coarse_samples, cfg.num_fine_samples
  n_frame = args.image * args.image
  per = (n_frame + world - 1) // world
  lo, hi = rank * per, min(n_frame, (rank + 1) * per)

  # ---- inputs: `world` frames per step, this rank's block of each; device-resident + pinned host copies
  frames = []
  g = torch.Generator(device=dev)
  g.manual_seed(1234 + rank)
  for f in range(world):
    o, d, w = frame_inputs(syn, args, f)
    blk = lambda a: np.ascontiguousarray(a[lo:hi])
    ho, hd, hw = blk(o), blk(d), blk(w)
    t_rand = torch.rand((hi - lo, Sc), generator=g, device=dev)
    u = torch.rand((hi - lo, Sf), generator=g, device=dev)
    frames.append({
        'o': torch.from_numpy(ho).to(dev), 'd': torch.from_numpy(hd).to(dev),
        'w': torch.from_numpy(hw.view(np.int32)).to(dev), 't': t_rand, 'u': u,
        'ho': torch.from_numpy(ho).pin_memory(), 'hd': torch.from_numpy(hd).pin_memory(),
        'hw': torch.from_numpy(hw.view(np.int32)).pin_memory(), 'ht': t_rand.cpu().pin_memory(),
        'hu': u.cpu().pin_memory()})
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

  def render_frame(i, o, d, w, t, uu):
    if peer is not None:
      pf = peer[i & 1]
      pf.activate()
      R.render_rays(o, d, warp_id=w, t_rand=t, u=uu, extra=extra, coarse_keys=(), fine_keys=RENDER_KEYS,
                    fine_ptrs=pf.shard_ptrs(lo))
      pf.wait()
      return pf.frame()
    out = R.render_rays(o, d, warp_id=w, t_rand=t, u=uu, extra=extra, coarse_keys=(), fine_keys=RENDER_KEYS)['fine']
    return all_gather_level(out) if world > 1 else out

  def step_device():
    outs = None
    for i, fr in enumerate(frames):
      outs = render_frame(i, fr['o'], fr['d'], fr['w'], fr['t'], fr['u'])
    flush.fill_(1)
    return outs

  host_out = {}

  def step_host():
    """End to end through the host-buffer entry point: pinned host rays in, pinned host image block out."""
    res = None
    for i, fr in enumerate(frames):
      if world == 1:
        res = R.render_rays_host(fr['ho'].numpy(), fr['hd'].numpy(), warp_id=fr['hw'].numpy().view(np.uint32),
                                 t_rand=fr['ht'].numpy(), u=fr['hu'].numpy(), extra=extra, fine_keys=RENDER_KEYS,
                                 out=host_out.setdefault(i, {}))
      else:
        o, d, w = (fr[k].to(dev, non_blocking=True) for k in ('ho', 'hd', 'hw'))
        t, uu = fr['ht'].to(dev, non_blocking=True), fr['hu'].to(dev, non_blocking=True)
        out = render_frame(i, o, d, w, t, uu)
        hb = host_out.setdefault(i, {})
        for k, v in out.items():
          if k not in hb:
            hb[k] = torch.empty(v.shape, dtype=v.dtype).pin_memory()
          hb[k].copy_(v, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        res = hb
    flush.fill_(1)
    return res

  if rank == 0 and world == 1:      # make the render_rays_host output buffers pinned too
    from nerfds_b200.renderer import ALL_SHAPES
    for i in range(len(frames)):
      host_out[i] = {k: torch.empty((hi - lo,) + tuple(ALL_SHAPES[k](Sc + Sf, R.H)), dtype=torch.float32).pin_memory().numpy()
                     for k in RENDER_KEYS}

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
  rays_step = n_frame * world      # whole job: `world` frames per step
  value = rays_step * args.steps / (ms * 1e-3)

  # ---- end to end (host buffers)
  for _ in range(min(args.warmup, 2)):
    step_host()
  e2e_steps = max(1, min(args.steps, 3))
  ms_h, _, _ = timed(step_host, e2e_steps)
  e2e_value = rays_step * e2e_steps / (ms_h * 1e-3)
  fr = frames[0]
  h2d = sum(fr[k].numel() * fr[k].element_size() for k in ('ho', 'hd', 'hw', 'ht', 'hu')) * world
  per_ray_out = sum(int(np.prod(ALL_SHAPES_local(k, Sc + Sf, R.H))) for k in RENDER_KEYS) * 4
  d2h = per_ray_out * (n_frame if world > 1 else (hi - lo)) * world

  if rank == 0:
    pk, pk_kind = peaks()
    # dominant kernel: the fine-level field kernel; algorithmic FLOPs per launch / measured duration
    f_ms, f_n = prof['field_fine']
    c_ms, c_n = prof['field_coarse']
    rays_local_total = (hi - lo) * world * args.steps
    fine_flops = 2.0 * MAC_FINE * (Sc + Sf) * rays_local_total
    coarse_flops = 2.0 * MAC_COARSE * Sc * rays_local_total
    ach = fine_flops / (f_ms * 1e-3) / 1e12 if f_ms > 0 else 0.0
    peak = float(pk.get('bf16_tflops_sustained') or pk.get('bf16_tflops') or 1400.0)
    burst = pk.get('bf16_tflops', peak)
    roofline = {'bound': 'tensor', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak,
                # DRAM bytes per fine-level pass, from the ncu capture of the same command (profiles/): 371 B per
                # sample evaluation (planes written once, carry planes read once, plus write-backs of the previous
                # kernel's dirty L2 lines that ncu attributes to this one); --traffic overrides
                'traffic': args.traffic if args.traffic is not None else (
                    NCU_DRAM_BYTES_PER_EVAL * (Sc + Sf) * rays_local_total / max(f_n, 1) if R.engine == 'tc' else None),
                'kernel': (f'field_tc_kernel x2 (fine level: {Sf} new depths per ray through every network + {Sc} coarse depths '
                           'through the template NeRF on carried warp/hyper/mask results)') if R.engine == 'tc'
                else f'field_{R.engine}_kernel (fine level)',
                'peak_kind': f'{pk_kind} sustained bf16 (burst {burst})',
                'launches': f_n, 'avg_launch_ms': f_ms / max(f_n, 1),
                'flops_per_launch': fine_flops / max(f_n, 1),
                'whole_step_tflops': (fine_flops + coarse_flops) / (ms * 1e-3) / 1e12,
                'stage_ms_per_step': {k: v[0] / args.steps for k, v in prof.items()},
                'hbm_gbs_of_peak': None}
    cpu = None
    if not args.no_cpu:
      threads = os.cpu_count() or 1
      v, dt = cpu_sample(args, cfg, params, syn, args.cpu_rays, threads)
      cpu = {'value': v, 'unit': UNIT, 'cores': threads, 'kind': 'port',
             'sample': f'{args.cpu_rays} rays spread over frame 0, {Sc}+{Sf} samples, {dt:.1f} s, PyTorch-CPU fp32 '
                       f'restatement of the reference JAX path incl. autograd d(sigma)/dx (jax not installable here)'}
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f32 (tensor-core layers: split-fp16 operands, fp32 accumulate)' if R.engine == 'tc' else 'f32',
            'data': 'synthetic', 'config': dict(workload_config(args, world), engine=R.engine, precision=args.precision),
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'ms_per_step': ms_h / e2e_steps},
            'gpu_launches': launches, 'roofline': roofline, 'cpu_baseline': cpu}
    print(json.dumps(line), flush=True)
  if peer is not None:
    for pf in peer:
      pf.close()
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


def ALL_SHAPES_local(k, S, H):
  from nerfds_b200.renderer import ALL_SHAPES
  return ALL_SHAPES[k](S, H) or (1,)


NCU_DRAM_BYTES_PER_EVAL = 371.0   # profiles/r1_field_tc_fine_ncu_full.txt


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument('--gpus', type=int, default=1)
  ap.add_argument('--steps', type=int, default=3)
  ap.add_argument('--warmup', type=int, default=3)
  ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
  ap.add_argument('--engine', default='auto', choices=['auto', 'tc', 'simt'])
  ap.add_argument('--precision', default='split3', choices=['mixed', 'fp16', 'split3'])
  ap.add_argument('--image', type=int, default=800)
  ap.add_argument('--coarse', type=int, default=128)
  ap.add_argument('--fine', type=int, default=128)
  ap.add_argument('--chunk', type=int, default=65536)
  ap.add_argument('--cpu-rays', type=int, default=None)
  ap.add_argument('--no-cpu', action='store_true')
  ap.add_argument('--gather', default='peer', choices=['peer', 'nccl'], help='frame reassembly at N > 1')
  ap.add_argument('--traffic', type=float, default=None, help='DRAM bytes/launch of the dominant kernel from ncu')
  args = ap.parse_args()
  if args.cpu_rays is None:
    args.cpu_rays = 1024 if args.impl == 'ours' else 512
  world = int(os.environ.get('WORLD_SIZE', '1'))
  if world != args.gpus and args.impl == 'ours':
    if world == 1 and args.gpus > 1:
      raise SystemExit(f'--gpus {args.gpus} needs torchrun (python -m torch.distributed.run --nproc-per-node {args.gpus} ...)')
  if args.impl == 'reference':
    run_reference(args)
  else:
    run_ours(args)


if __name__ == '__main__':
  main()
