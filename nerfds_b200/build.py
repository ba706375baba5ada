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


BUILD_INFO = os.path.join(LIBDIR, 'BUILD_INFO.json')


def source_fingerprint() -> dict:
  """sha256 of every source the library is built from, plus the compile flags: written next to the .so by `build`
  and checked by `_lib.load_library`, so that a shipped binary can never silently differ from the tree's sources."""
  import hashlib
  files = sorted([os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cu', '.cuh', '.h'))] +
                 [os.path.join(ROOT, 'include', 'nerfds_b200.h')])
  out = {}
  for f in files:
    with open(f, 'rb') as fh:
      out[os.path.relpath(f, ROOT)] = hashlib.sha256(fh.read()).hexdigest()
  flags = [a.replace(ROOT, '.') for a in COMMON]           # (paths relative: the tree may live anywhere)
  out['flags'] = hashlib.sha256(repr((flags, SOURCES)).encode()).hexdigest()
  return out


def recorded_fingerprint():
  import json
  try:
    with open(BUILD_INFO) as f:
      return json.load(f).get('sources')
  except (OSError, ValueError):
    return None


def build(force: bool = False, verbose: bool = False) -> str:
  """Compile what changed.  Staleness is decided by CONTENT (sha256 of each source and of the flags, recorded in
  lib/BUILD_INFO.json), not by mtimes -- a snapshot copied to another box keeps its binary only if it matches."""
  import json
  os.makedirs(LIBDIR, exist_ok=True)
  want = source_fingerprint()
  have = recorded_fingerprint() or {}
  shared_changed = any(want.get(k) != have.get(k) for k in want if not k.endswith('.cu'))
  objs = []
  procs = []
  for src, extra in SOURCES.items():
    s = os.path.join(CSRC, src)
    o = os.path.join(LIBDIR, src.replace('.cu', '.o'))
    objs.append(o)
    key = os.path.relpath(s, ROOT)
    if force or shared_changed or not os.path.exists(o) or want.get(key) != have.get(key):
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
  if force or procs or not os.path.exists(LIB):
    cmd = [_nvcc(), '-shared', '-o', LIB] + objs + ARCH + ['-lcudart_static', '-lpthread', '-ldl', '-lrt']
    subprocess.check_call(cmd)
  if force or procs or recorded_fingerprint() != want:
    with open(BUILD_INFO, 'w') as f:
      json.dump({'sources': want, 'compiled': [src for src, _ in procs], 'nvcc': _nvcc(),
                 'arch': 'compute_100a,sm_100a'}, f, indent=1)
  return LIB


if __name__ == '__main__':
  print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
