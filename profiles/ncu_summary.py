#!/usr/bin/env python
"""Summarise ncu outputs brought back from the GPU box (read here, no GPU needed).

  python profiles/ncu_summary.py launches <launches.csv>           per-kernel launch durations
  python profiles/ncu_summary.py raw <report.ncu-rep>              key metrics of each profiled launch
  python profiles/ncu_summary.py sass <report.ncu-rep> [kernel#]   opcode mix, stall reasons, hottest instructions
"""
import collections
import csv
import io
import re
import subprocess
import sys

RAW = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
       'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
       'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
       'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
       'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
       'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
       'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.max', 'launch__grid_size',
       'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum']


def ncu_csv(rep, page):
    out = subprocess.run(['ncu', '-i', rep, '--page', page, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    for r in rows[1:]:
        agg.setdefault(re.sub(r'\(.*', '', r[ki])[:60], []).append(float(r[vi].replace(',', '')))
    tot = sum(sum(v) for v in agg.values())
    for n, v in agg.items():
        print('%-62s n=%3d mean=%8.1f us min=%8.1f share=%5.1f%%' % (n, len(v), sum(v) / len(v) / 1e3, min(v) / 1e3, 100 * sum(v) / tot))


def raw(rep):
    rows = ncu_csv(rep, 'raw')
    hdr = rows[0]
    names = [r[hdr.index('Kernel Name')][:40] for r in rows[2:]]
    print('kernels:', names)
    for w in RAW:
        if w in hdr:
            i = hdr.index(w)
            print('%-62s %-10s %s' % (w, rows[1][i], [r[i] for r in rows[2:]]))


def sass(rep, which=0):
    rows = ncu_csv(rep, 'source')
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    lo = starts[which]
    hi = starts[which + 1] if which + 1 < len(starts) else len(rows)
    print(rows[lo][1])
    hdr = rows[lo + 1]
    body = [r for r in rows[lo + 2:hi] if len(r) > 10]
    ie, isrc, isamp = hdr.index('Instructions Executed'), hdr.index('Source'), hdr.index('# Samples')
    tot = sum(int(r[ie]) for r in body)
    print('static', len(body), 'executed warp-instr', tot)
    ops = collections.Counter()
    for r in body:
        s = re.sub(r'^@!?U?P\d+\s+', '', r[isrc].strip())
        ops[s.split()[0].split('.')[0]] += int(r[ie])
    print(' '.join('%s:%.1f%%' % (o, 100 * c / tot) for o, c in ops.most_common(24)))
    st = collections.Counter()
    for i, h in enumerate(hdr):
        if h.startswith('stall_') and 'Not Issued' not in h:
            st[h] = sum(int(r[i] or 0) for r in body)
    tot_s = sum(st.values())
    print(' '.join('%s:%.1f%%' % (h[6:], 100 * c / tot_s) for h, c in st.most_common(8)))
    for r in sorted(body, key=lambda r: -int(r[isamp]))[:14]:
        print('   %6s samples  x%-8s %s' % (r[isamp], r[ie], r[isrc].strip()[:80]))


if __name__ == '__main__':
    cmd = sys.argv[1]
    if cmd == 'launches':
        launches(sys.argv[2])
    elif cmd == 'raw':
        raw(sys.argv[2])
    else:
        sass(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
