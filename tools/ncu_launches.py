#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.

    python tools/ncu_launches.py gpurun_out/launches.csv > profiles/rNN_launches.txt

Per-launch times under ncu are cold-cache and serialised: compare the SHARES with
bench.py's accounting, not the absolute durations.
"""
import collections
import csv
import sys


def main(path):
  rows = [r for r in csv.reader(open(path)) if len(r) > 10]
  hdr = rows[0]
  name, value, unit = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
  agg = collections.OrderedDict()
  for r in rows[1:]:
    ns = float(r[value].replace(',', '')) * {'ns': 1, 'us': 1e3, 'ms': 1e6}.get(r[unit], 1)
    a = agg.setdefault(r[name], [0, 0.0])
    a[0] += 1
    a[1] += ns
  total = sum(a[1] for a in agg.values()) or 1.0
  print('# kernel, launches, total ns, share, mean ns  (cold-cache, serialised: compare shares)')
  for k in sorted(agg, key=lambda k: -agg[k][1]):
    n, ns = agg[k]
    print('%-74s %5d %10d %5.1f%% %8d' % (k[:74], n, ns, 100.0 * ns / total, ns / n))


if __name__ == '__main__':
  main(sys.argv[1])
