"""Top stalled SASS instructions per kernel from `ncu --page source --csv` output
(optionally gzipped).  Usage: python tools/ncu_top_stalls.py source.csv[.gz] [kernel_index] [top_n]"""
import csv
import gzip
import sys

path = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
op = gzip.open if path.endswith('.gz') else open
blocks, cur = [], None
with op(path, 'rt') as f:
  for row in csv.reader(f):
    if not row:
      continue
    if row[0] == 'Kernel Name':
      cur = dict(name=row[1], hdr=None, rows=[])
      blocks.append(cur)
    elif cur is not None and cur['hdr'] is None:
      cur['hdr'] = row
    elif cur is not None:
      cur['rows'].append(row)
for bi, b in enumerate(blocks):
  if which is not None and bi != which:
    continue
  h = {n: i for i, n in enumerate(b['hdr'])}
  stall_cols = [n for n in b['hdr'] if n.startswith('stall_') and 'Not Issued' not in n]
  tot = sum(int(r[h['# Samples']]) for r in b['rows'])
  print(f'== kernel {bi}: {b["name"][:110]}  total samples {tot}')
  agg = {n: sum(int(r[h[n]]) for r in b['rows']) for n in stall_cols}
  print('   stall mix: ' + ', '.join(f'{k[6:]} {100 * v / max(tot, 1):.0f}%' for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
  if which is None:
    continue
  rows = sorted(enumerate(b['rows']), key=lambda ir: -int(ir[1][h['# Samples']]))[:top]
  for i, r in rows:
    st = sorted(((int(r[h[n]]), n[6:]) for n in stall_cols), reverse=True)[:3]
    print(f'{i:5d} {100 * int(r[h["# Samples"]]) / max(tot, 1):5.1f}%  {r[h["Source"]].strip()[:70]:70s} ' +
          ' '.join(f'{n}:{v}' for v, n in st if v))
