"""Summarise a NDS_TC_TRACE dump: timeline of one pair of 128-sample tiles (cycles)."""
import sys
bursts, steps = [], []
for l in open(sys.argv[1]):
  t = l.split()
  if t[0] == 'sub':
    continue
  d = {t[i]: int(t[i + 1]) for i in range(2, len(t) - 1, 2)}
  d['i'] = int(t[1])
  (bursts if t[0] == 'burst' else steps).append(d)
kinds = {0: 'EPI', 1: 'HEAD', 2: 'VIEW', 3: 'PREP', 4: 'OUT', 5: 'INGR', 6: 'SEED', 7: 'PEB'}
print('pair span (cycles):', max(s['end'] for s in steps), ' bursts', len(bursts))
steps = [s for s in steps if s['start'] >= 0]
live = [b for b in bursts if b['top'] >= 0]
print('issuer: sum(top->ready) %d  sum(ready->issued) %d  sum(issued->next top) %d' % (
    sum(b['ready'] - b['top'] for b in live), sum(b['issued'] - b['ready'] for b in live),
    sum(live[i + 1]['top'] - live[i]['issued'] for i in range(len(live) - 1))))
if len(sys.argv) > 2 and sys.argv[2] == 'bursts':
  for b in bursts:
    print(f"burst {b['i']:3d} slot {b['slot']} rows {b['rows']:3d} flags {b['flags']:5d} wait {b['ready'] - b['top']:5d} issue {b['issued'] - b['ready']:5d} top {b['top']:7d}")
for s in steps:
  print(f"step {s['i']:3d} {kinds[s['kind']]:4s} slot {s['slot']} op {s['op']:2d} N {s['N']:3d} glue {s['glue']} start {s['start']:7d} dur {s['end'] - s['start']:5d}")
