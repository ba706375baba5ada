"""Builds libnerfds_b200.so in-tree with nvcc for sm_100a (no torch dependency).

    python -m nerfds_b200.build            # or __graft_entry__.build()
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIB = os.path.join(LIBDIR, 'libnerfds_b200.so')

ARCH = ['-gencode', 'arch=compute_100a,code=sm_100a']
COMMON = ['-O3', '-std=c++17', '-lineinfo', '-Xcompiler', '-fPIC', '-I', os.path.join(ROOT, 'include'),
          '-I', CSRC] + ARCH
# per-file extra flags: the discrete per-ray stages forbid FMA contraction
SOURCES = {
    'nds_api.cu': [],
    'nds_field_simt.cu': [],
    'nds_field_tc.cu': [],
    'nds_composite.cu': ['-fmad=false'],
}


def _nvcc() -> str:
  for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
    if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
      return cand
  raise RuntimeError('nvcc not found')


def _stale(target: str, deps) -> bool:
  if not os.path.exists(target):
    return True
  t = os.path.getmtime(target)
  return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
  os.makedirs(LIBDIR, exist_ok=True)
  headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.h', '.cuh'))]
  headers.append(os.path.join(ROOT, 'include', 'nerfds_b200.h'))
  objs = []
  procs = []
  for src, extra in SOURCES.items():
    s = os.path.join(CSRC, src)
    o = os.path.join(LIBDIR, src.replace('.cu', '.o'))
    objs.append(o)
    if force or _stale(o, [s] + headers):
      cmd = [_nvcc()] + COMMON + extra + (['-Xptxas', '-v'] if verbose else []) + ['-c', s, '-o', o]
      procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
  failed = False
  for src, p in procs:
    out, _ = p.communicate()
    if p.returncode != 0:
      failed = True
      sys.stderr.write(f'--- nvcc failed on {src} ---\n{out}\n')
    elif verbose or out.strip():
      sys.stderr.write(f'--- {src} ---\n{out}\n')
  if failed:
    raise RuntimeError('nvcc compilation failed')
  if force or procs or _stale(LIB, objs):
    cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ARCH + ['-lcudart_static', '-lpthread', '-ldl', '-lrt']
    subprocess.check_call(cmd)
  return LIB


if __name__ == '__main__':
  print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
