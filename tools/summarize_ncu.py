#!/usr/bin/env python
"""Turn the ncu outputs of the GPU visits (tools/gpu_visits/) into the text summaries kept under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/<tag>_launches.csv
    python tools/summarize_ncu.py kernels  gpurun_out/<tag>_solve.ncu-rep
"""
import collections
import csv
import io
import subprocess
import sys

METRICS = [
    'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size',
    'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
    'sm__warps_active.avg.pct_of_peak_sustained_active',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
    'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
    'sm__cycles_elapsed.avg', 'sm__cycles_elapsed.avg.per_second',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
    'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
    'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
]


def launches(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr = rows[hi]
    kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        try:
            v = float(r[mv].replace(',', ''))
        except ValueError:
            continue
        a = agg.setdefault(r[kn], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print('{0:<72s} {1:>6s} {2:>11s} {3:>12s} {4:>6s}'.format('kernel', 'count', 'total ms',
                                                             'us / launch', 'share'))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('{0:<72s} {1:6d} {2:11.3f} {3:12.2f} {4:5.1f}%'.format(
            k[:72], a[0], a[1] / 1e6, a[1] / 1e3 / a[0], 100 * a[1] / tot))


def kernels(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('# ' + r[hdr.index('Kernel Name')])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print('{0:<70s} {1:<16s} {2}'.format(m, units[i], r[i]))
        # warp-state sampling: where the warps wait (share of the samples)
        stalls = []
        for i, name in enumerate(hdr):
            if 'pcsamp_warps_issue_stalled_' in name and not name.endswith('_not_issued'):
                try:
                    stalls.append((float(r[i]), name.split('pcsamp_warps_issue_stalled_')[1]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls)
        if tot > 0:
            top = sorted(stalls, reverse=True)[:6]
            print('{0:<70s} {1:<16s} {2}'.format('warp samples by state', '%', ', '.join(
                '{0} {1:.1f}'.format(nm, 100. * v / tot) for v, nm in top)))
        print()


if __name__ == '__main__':
    {'launches': launches, 'kernels': kernels}[sys.argv[1]](sys.argv[2])
