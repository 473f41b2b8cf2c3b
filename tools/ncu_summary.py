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
  # The correlated page comes as one section per source FILE of the first kernel
  # (header, then a "File Path" row); aggregate them all.
  text = ncu('-i', path, '--page', 'source', '--print-source', 'cuda,sass', '--csv')
  hdr, agg, cur, fname, first_kernel = None, {}, None, '?', None
  ops = {}
  for r in csv.reader(io.StringIO(text)):
    if not r:
      continue
    if r[0] == 'File Path':
      fname = r[1].split('/')[-1]
      continue
    if r[0] == 'Function Name':
      if first_kernel is None:
        first_kernel = r[1]
      cur = None
      continue
    if r[0] == 'Line No' and len(r) > 5:
      hdr = r
      continue
    if hdr is None or len(r) != len(hdr):
      continue
    if r[0]:
      cur = (fname, r[0], r[1].strip()[:92])
      agg.setdefault(cur, [0, 0])
      continue
    if cur is None:
      continue
    try:
      n_s, n_i = int(r[hdr.index('# Samples')]), int(r[hdr.index('Instructions Executed')])
    except (ValueError, TypeError, KeyError):
      continue
    agg[cur][0] += n_s
    agg[cur][1] += n_i
    sass = r[3].split()
    if sass:
      op = (sass[1] if sass[0].startswith('@') and len(sass) > 1 else sass[0]).split('.')[0]
      ops[op] = ops.get(op, 0) + n_i
  tot = sum(v[0] for v in agg.values()) or 1
  tot_i = sum(v[1] for v in agg.values()) or 1
  print('\n## warp-stall samples by CUDA source line (launch 1, all source files): %d samples, '
        '%d warp instructions' % (tot, tot_i))
  print('  file:line                samples   share   warp-instr  source')
  for k in sorted(agg, key=lambda k: -agg[k][0])[:25]:
    print('  %-22s %7d  %5.1f%%  %10d  %s' % ('%s:%s' % (k[0], k[1]), agg[k][0],
                                              100.0 * agg[k][0] / tot, agg[k][1], k[2]))
  print('\n## executed warp instructions by CUDA source line (top 25)')
  for k in sorted(agg, key=lambda k: -agg[k][1])[:25]:
    print('  %-22s %10d  %5.1f%%  %s' % ('%s:%s' % (k[0], k[1]), agg[k][1],
                                         100.0 * agg[k][1] / tot_i, k[2]))
  print('\n## executed warp instructions by SASS opcode')
  print('  ' + '  '.join('%s %d' % (o, ops[o]) for o in sorted(ops, key=lambda o: -ops[o])[:24]))


if __name__ == '__main__':
  main(sys.argv[1])
