"""Summarises an ncu --csv launch list (gpu__time_duration etc.) into a per-launch table and a
per-kernel-family share table.  Usage: python tools/summarize_launches.py launches.csv [out.md]"""
import collections
import csv
import sys

path = sys.argv[1]
lines = [l for l in open(path) if l.startswith('"')]
data = collections.OrderedDict()
for row in csv.DictReader(lines):
  k = (int(row['ID']), row['Kernel Name'], row.get('Grid Size'))
  data.setdefault(k, {})[row['Metric Name']] = float(row['Metric Value'].replace(',', ''))
out, fam = [], collections.OrderedDict()
tot = 0.0
for (i, name, grid), m in data.items():
  t = m.get('gpu__time_duration.sum', 0) / 1e3
  tot += t
  short = name.replace('void ', '').replace('svdd::', '').replace('<unnamed>::', '').replace('gemm_detail::', '')
  short = short.split('(')[0]
  tens = m.get('sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active', 0)
  dram = (m.get('dram__bytes_read.sum', 0) + m.get('dram__bytes_write.sum', 0)) / 1e6
  out.append((i, short, grid, t, tens, dram, m.get('lts__t_bytes.sum', 0) / 1e6))
  f = fam.setdefault(short, [0, 0.0])
  f[0] += 1
  f[1] += t
w = open(sys.argv[2], 'w') if len(sys.argv) > 2 else sys.stdout
w.write(f'total device time {tot:.1f} us over {len(out)} launches (ncu: cold caches, serialised)\n\n')
w.write('| kernel | launches | us | share |\n|---|---|---|---|\n')
for k, (n, t) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
  w.write(f'| {k} | {n} | {t:.1f} | {100 * t / tot:.1f}% |\n')
w.write('\n| # | kernel | grid | us | tensor pipe % | dram MB | L2 MB |\n|---|---|---|---|---|---|---|\n')
for o in out:
  w.write('| %d | %s | %s | %.1f | %.1f | %.1f | %.1f |\n' % o)
