"""The C-ABI library loads on a CPU-only box and exports every symbol that
include/nerfds_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
  src = open(os.path.join(ROOT, 'include', 'nerfds_b200.h')).read()
  src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
  return sorted(set(re.findall(r'\b(ndsr_[a-z_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
  from nerfds_b200 import _lib
  lib = _lib.load_library()
  names = _declared_functions()
  assert len(names) >= 14
  for n in names:
    assert hasattr(lib, n), f'{n} declared in the header but not exported'
  assert set(names) == set(_lib.EXPORTS)


def test_struct_mirrors_match():
  from nerfds_b200 import _lib
  lib = _lib.load_library()
  a, b, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
  lib.ndsr_struct_sizes(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
  assert a.value == ctypes.sizeof(_lib.ndsr_config)
  assert b.value == ctypes.sizeof(_lib.ndsr_extra_params)
  assert c.value == ctypes.sizeof(_lib.ndsr_outputs)
  assert lib.ndsr_abi_version() == _lib.NDSR_ABI_VERSION


def test_create_rejects_bad_config_without_gpu():
  from nerfds_b200 import _lib
  from nerfds_b200.config import nerf_ds_config
  lib = _lib.load_library()
  c = _lib.to_c_config(nerf_ds_config())
  c.trunk_width = 100                       # not a supported width
  h = ctypes.c_void_p()
  rc = lib.ndsr_create(ctypes.byref(c), 0, ctypes.byref(h))
  assert rc == -1 and b'width' in lib.ndsr_last_error(None)
  c = _lib.to_c_config(nerf_ds_config())
  c.size = 4
  assert lib.ndsr_create(ctypes.byref(c), 0, ctypes.byref(h)) == -1


def test_no_cpu_fallback():
  """Constructing the model without a CUDA device must fail loudly."""
  import torch
  if torch.cuda.is_available():
    pytest.skip('GPU present')
  from nerfds_b200.config import nerf_ds_config
  from nerfds_b200.models import NerfModel
  from nerfds_b200.renderer import NdsrError
  with pytest.raises(NdsrError):
    NerfModel(nerf_ds_config())


def test_product_never_imports_oracle():
  pkg = os.path.join(ROOT, 'nerfds_b200')
  for dirpath, _, files in os.walk(pkg):
    for f in files:
      if f.endswith(('.py', '.cu', '.cuh', '.h')):
        txt = open(os.path.join(dirpath, f)).read()
        assert 'import oracle' not in txt and 'from oracle' not in txt, f


def test_header_is_plain_c(tmp_path):
  """include/nerfds_b200.h is a C header (C99, no C++ or CUDA types): it must compile on its own with gcc."""
  import shutil
  import subprocess
  gcc = shutil.which('gcc')
  if gcc is None:
    pytest.skip('no gcc')
  src = tmp_path / 't.c'
  src.write_text('#include "nerfds_b200.h"\nint main(void) { ndsr_config c; ndsr_outputs o; ndsr_extra_params e; ndsr_camera cam;\n'
                 '  ndsr_ipc_handle h; (void)c; (void)o; (void)e; (void)cam; (void)h; return sizeof(ndsr_tensor) > 0 ? 0 : 1; }\n')
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  r = subprocess.run([gcc, '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-fsyntax-only', '-I',
                      os.path.join(root, 'include'), str(src)], capture_output=True, text=True)
  assert r.returncode == 0, r.stderr


def test_c_client_links_and_runs(tmp_path):
  """examples/abi_probe.c: a C99 program linked against the shared library alone (no CUDA, no C++ headers) sees the
  declared ABI version / struct sizes and the error path of ndsr_create -- runs without a GPU."""
  import shutil
  import subprocess
  from nerfds_b200 import _lib
  gcc = shutil.which('gcc')
  if gcc is None or not os.path.exists(_lib.LIB_PATH):
    pytest.skip('no gcc or library not built')
  root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
  libdir = os.path.dirname(_lib.LIB_PATH)
  exe = str(tmp_path / 'abi_probe')
  r = subprocess.run([gcc, '-std=c99', '-Wall', '-Wextra', '-Werror', '-I', os.path.join(root, 'include'),
                      os.path.join(root, 'examples', 'abi_probe.c'), '-L', libdir, '-lnerfds_b200',
                      f'-Wl,-rpath,{libdir}', '-o', exe], capture_output=True, text=True)
  assert r.returncode == 0, r.stderr
  r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
  assert r.returncode == 0 and f'abi {_lib.NDSR_ABI_VERSION} ok' in r.stdout, r.stdout + r.stderr
  # the ctypes mirrors have the C compiler's sizes (layout drift would shift every later field)
  sizes = dict(kv.split('=') for kv in r.stdout.splitlines()[-1].split()[1:])
  want = {'config': _lib.ndsr_config, 'extra_params': _lib.ndsr_extra_params, 'outputs': _lib.ndsr_outputs,
          'camera': _lib.ndsr_camera, 'tensor': _lib.ndsr_tensor}
  for k, t in want.items():
    assert int(sizes[k]) == ctypes.sizeof(t), (k, sizes[k], ctypes.sizeof(t))
  assert int(sizes['ipc_handle']) == 64
