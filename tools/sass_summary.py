"""Per-kernel SASS instruction counts of the shipped library (cuobjdump -sass): profiles/r2_sass_summary.txt.
   python tools/sass_summary.py > profiles/r2_sass_summary.txt"""
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerfds_b200 import build as B

OPS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UBLKCP', 'UTMALDG', 'SYNCS', 'ELECT', 'HMMA', 'FFMA', 'MUFU', 'LDG', 'STG', 'LDS', 'STS', 'BAR', 'ATOM', 'RED', 'SHFL']
sass = subprocess.run(['cuobjdump', '-sass', B.LIB], capture_output=True, text=True).stdout
kernels, cur = {}, None
for line in sass.splitlines():
  m = re.search(r'Function : (\S+)', line)
  if m:
    name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
    name = re.sub(r'\(.*', '', name).replace('void ', '')
    cur = kernels.setdefault(name, {'instr': 0, **{o: 0 for o in OPS}})
    continue
  m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
  if m and cur is not None:
    cur['instr'] += 1
    op = m.group(1)
    for o in OPS:
      if op == o or op.startswith(o + '.') or (o in ('LDG', 'STG', 'LDS', 'STS', 'ATOM', 'RED', 'SHFL', 'BAR') and op.startswith(o)):
        cur[o] += 1
        break
print('# SASS summary of nerfds_b200/lib/libnerfds_b200.so (cuobjdump -sass; the only target is sm_100a), per kernel: tools/sass_summary.py.')
print('# UTCHMMA = tcgen05.mma kind::f16 (static count: the issuer loops over a burst program), UTCBAR = tcgen05.commit,')
print('# LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (bulk TMA, weights are pre-swizzled on the host), SYNCS = mbarrier')
print('# ops, ELECT = elect.sync.  No HMMA (mma.sync) and no UTMALDG anywhere.  field_tc_kernel<true> = with the reverse sweep.')
print()
print(f"{'kernel':46s}" + ''.join(f'{c:>8s}' for c in ['instr'] + OPS))
for k, v in sorted(kernels.items(), key=lambda kv: -kv[1]['instr']):
  print(f'{k[:45]:46s}' + ''.join(f'{v[c]:8d}' for c in ['instr'] + OPS))
