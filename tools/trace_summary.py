"""Summarise a NDS_TC_TRACE dump: per-op timeline of one 128-sample tile (cycles)."""
import sys
imgs, ops = [], []
for l in open(sys.argv[1]):
  t = l.split()
  if t[0] == 'img':
    imgs.append(dict(i=int(t[1]), rows=int(t[3]), steps=int(t[5]), fl=int(t[7]), ready=int(t[9]), issued=int(t[11])))
  else:
    ops.append(dict(i=int(t[1]), N=int(t[3]), kind=int(t[5]), glue=int(t[7]), c0s=int(t[9]), c0d=int(t[11]), c1s=int(t[13]), c1d=int(t[15])))
print('tile span (cycles):', max(max(o['c0d'], o['c1d']) for o in ops))
live = [im for im in imgs if im['ready'] >= 0]
print('issue bursts', len(live), 'sum(ready->issued)', sum(im['issued'] - im['ready'] for im in live),
      'avg', sum(im['issued'] - im['ready'] for im in live) / max(1, len(live)))
prev = None
for o in ops:
  per = '' if prev is None else f"period {o['c0s'] - prev['c0s']:6d}"
  if o["kind"] != 4:
    print(f"op {o['i']:2d} N {o['N']:3d} kind {o['kind']} c0 seen {o['c0s']:7d} epi {o['c0d'] - o['c0s']:5d} | c1 seen +{o['c1s'] - o['c0d']:5d} epi {o['c1d'] - o['c1s']:5d} {per}")
  else:
    print(f"op {o['i']:2d} HEAD glue {o['glue']} seen {o['c0s']:7d} head-read {o['c1s'] - o['c0s']:5d} glue {o['c0d'] - o['c1s']:5d} {per}")
  prev = o
