#!/usr/bin/env python
"""Summarise an .ncu-rep into a small text file fit for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt

Prints, per captured launch, the metrics the roofline discussion uses, then the
CUDA source lines that collected the most warp-stall samples (first launch).
"""
import csv
import io
import subprocess
import sys

RAW = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
       'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
       'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
       'sm__throughput.avg.pct_of_peak_sustained_elapsed',
       'sm__warps_active.avg.pct_of_peak_sustained_active',
       'smsp__inst_executed.sum', 'launch__registers_per_thread', 'launch__grid_size',
       'launch__block_size', 'launch__occupancy_limit_registers',
       'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
       'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
       'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static']


def ncu(*args):
  return subprocess.run(['ncu'] + list(args), capture_output=True, text=True).stdout


def main(path):
  rows = list(csv.reader(io.StringIO(ncu('-i', path, '--page', 'raw', '--csv'))))
  hdr, units = rows[0], rows[1]
  print('# ncu summary of %s' % path)
  for r in rows[2:]:
    print('\n## launch id %s: %s  grid %s block %s' % (
        r[hdr.index('ID')], r[hdr.index('Kernel Name')], r[hdr.index('Grid Size')],
        r[hdr.index('Block Size')]))
    for m in RAW:
      if m in hdr:
        print('  %-62s %s %s' % (m, r[hdr.index(m)], units[hdr.index(m)]))
  text = ncu('-i', path, '--page', 'source', '--print-source', 'cuda,sass', '--csv')
  hdr, agg, cur, kern = None, {}, None, 0
  for r in csv.reader(io.StringIO(text)):
    if r and r[0] == 'Function Name':
      kern += 1
      continue
    if r and r[0] == 'Line No' and len(r) > 5:
      hdr = r
      continue
    if hdr is None or len(r) != len(hdr) or kern != 1:
      continue
    if r[0]:
      cur = (r[0], r[1].strip()[:96])
      agg.setdefault(cur, [0, 0])
      continue
    try:
      agg[cur][0] += int(r[hdr.index('# Samples')])
      agg[cur][1] += int(r[hdr.index('Instructions Executed')])
    except (ValueError, TypeError, KeyError):
      pass
  tot = sum(v[0] for v in agg.values()) or 1
  print('\n## warp-stall samples by CUDA source line (launch 1): %d samples, %d warp '
        'instructions' % (tot, sum(v[1] for v in agg.values())))
  print('  line samples   share   warp-instr  source')
  for k in sorted(agg, key=lambda k: -agg[k][0])[:25]:
    print('  %4s %7d  %5.1f%%  %10d  %s' % (k[0], agg[k][0], 100.0 * agg[k][0] / tot,
                                           agg[k][1], k[1]))


if __name__ == '__main__':
  main(sys.argv[1])
