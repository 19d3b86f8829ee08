#!/usr/bin/env python
"""Turn ncu CSV output into the markdown tables kept under profiles/.

    python tools/ncu_summarize.py launches <launch-list.csv> <title> > profiles/rNN_launches_*.md
    python tools/ncu_summarize.py full <raw-page.csv> <title> [kernel-regex] > profiles/rNN_*_ncu_summary.md

`launches`: the `--metrics gpu__time_duration.sum[,smsp__inst_executed.sum] --csv --log-file` pass (one row per launch
and metric).  `full`: `ncu -i report.ncu-rep --page raw --csv` of a `--set full` capture; the largest instance of every
kernel (by duration) is tabulated.
"""
import csv
import re
import sys
from collections import OrderedDict

FULL_METRICS = [
    'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
    'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
    'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
    'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
    'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
    'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
    'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']


def short(name):
    name = re.sub(r'\(.*$', '', name)
    return name if len(name) <= 60 else name[:60]


def launches(path, title):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    k, m, v, i = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
    per = OrderedDict()
    for r in rows[1:]:
        per.setdefault(r[i], {'name': r[k]})[r[m]] = float(r[v].replace(',', ''))
    agg = OrderedDict()
    for e in per.values():
        a = agg.setdefault(short(e['name']), {'n': 0, 'tot': 0.0, 'max': 0.0, 'inst': 0.0})
        t = e.get('gpu__time_duration.sum', 0.0) / 1e3
        a['n'] += 1
        a['tot'] += t
        a['max'] = max(a['max'], t)
        a['inst'] = max(a['inst'], e.get('smsp__inst_executed.sum', 0.0))
    total = sum(a['tot'] for a in agg.values())
    print('# %s\n' % title)
    print('| kernel | launches | mean us | max us | total us | share | max warp-instr |')
    print('|---|---:|---:|---:|---:|---:|---:|')
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]['tot']):
        print('| `%s` | %d | %.1f | %.1f | %.1f | %.1f%% | %.3g |' % (name, a['n'], a['tot'] / a['n'], a['max'], a['tot'],
                                                                  100 * a['tot'] / total, a['inst']))
    print('\nTotal %.1f us over %d launches (cold-cache, serialised under ncu: compare shares, not absolutes).' %
          (total, sum(a['n'] for a in agg.values())))


def full(path, title, pattern='.'):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    k = hdr.index('Kernel Name')
    t = hdr.index('gpu__time_duration.sum')
    best = OrderedDict()
    for r in rows[2:]:
        if not re.search(pattern, r[k]):
            continue
        key = short(r[k])
        if key not in best or float(r[t].replace(',', '')) > float(best[key][t].replace(',', '')):
            best[key] = r
    names = list(best)
    print('# %s\n' % title)
    print('| metric | ' + ' | '.join('`%s`' % n for n in names) + ' |')
    print('|---|' + '---|' * len(names))
    for m in FULL_METRICS:
        if m not in hdr:
            continue
        c = hdr.index(m)
        vals = []
        for n in names:
            x = best[n][c]
            try:
                vals.append('%.4g' % float(x.replace(',', '')))
            except ValueError:
                vals.append(x)
        print('| %s [%s] | %s |' % (m, units[c], ' | '.join(vals)))


if __name__ == '__main__':
    if sys.argv[1] == 'launches':
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else '.')
