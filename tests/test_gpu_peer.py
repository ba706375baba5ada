"""Frame reassembly over peer memory (nerfds_b200/peer.py): runs under torchrun on every GPU of the box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('n_procs', [1, 2])
def test_peer_frames_match_all_gather(cuda_device, n_procs):
  if torch.cuda.device_count() < n_procs:
    pytest.skip(f'needs {n_procs} GPUs')
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n_procs}',
         '--master-addr', '127.0.0.1', '--master-port', str(29600 + n_procs), os.path.join(ROOT, 'tests', 'multi', 'peer_frames_check.py')]
  r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
  assert r.returncode == 0 and 'peer frames OK' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_render_scene_distributed(cuda_device, tmp_path):
  """render.render_scene under torchrun: every rank renders its block of each camera's rays, frames are assembled
  by peer stores (two alternating buffer sets); identical to the single-process result."""
  n_procs = 2 if torch.cuda.device_count() >= 2 else 1
  cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n_procs}',
         '--master-addr', '127.0.0.1', '--master-port', '29611', os.path.join(ROOT, 'tests', 'multi', 'render_scene_check.py')]
  r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600, env=dict(os.environ, NDS_CHECK_DIR=str(tmp_path)))
  assert r.returncode == 0 and 'render_scene OK' in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
