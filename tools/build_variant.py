"""A/B experiments: build a variant of the library with extra -D flags on the tensor-core field kernel.

    python tools/build_variant.py NAME -DNDS_EPI_TRUNC=1 [...]   ->  tools/bin/lib_NAME.so

Run it with NDSR_LIBRARY=tools/bin/lib_NAME.so (the loader skips the provenance check for overrides).  The other
objects are taken from nerfds_b200/lib (build the product library first).
"""
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfds_b200 import build as B


def main():
  name, flags = sys.argv[1], sys.argv[2:]
  out_dir = os.path.join(B.ROOT, 'tools', 'bin')
  os.makedirs(out_dir, exist_ok=True)
  B.build()
  obj = os.path.join(out_dir, f'nds_field_tc_{name}.o')
  cmd = [B._nvcc()] + B.COMMON + flags + ['-Xptxas', '-v', '-c', os.path.join(B.CSRC, 'nds_field_tc.cu'), '-o', obj]
  r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
  for l in r.stdout.splitlines():
    if 'Compiling entry' in l or 'registers' in l or 'spill' in l or 'error' in l:
      print(l)
  if r.returncode:
    print(r.stdout)
    raise SystemExit(1)
  objs = [os.path.join(B.LIBDIR, s.replace('.cu', '.o')) for s in B.SOURCES if s != 'nds_field_tc.cu'] + [obj]
  lib = os.path.join(out_dir, f'lib_{name}.so')
  subprocess.check_call([B._nvcc(), '-shared', '-o', lib] + objs + B.ARCH + ['-lcudart_static', '-lpthread', '-ldl', '-lrt'])
  os.remove(obj)
  print(lib)


if __name__ == '__main__':
  main()
