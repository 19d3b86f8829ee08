#!/usr/bin/env python
"""Hot spots of one kernel from `ncu -i rep --page source --csv -k <kernel>`: top SASS lines by stall samples.
    python tools/ncu_hot.py file.csv [instance] [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
inst = []
for r in rows:
    if r and r[0] == 'Kernel Name':
        inst.append({'name': r[1], 'hdr': None, 'rows': []})
    elif inst and inst[-1]['hdr'] is None:
        inst[-1]['hdr'] = r
    elif inst:
        inst[-1]['rows'].append(r)
which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
for k, it in enumerate(inst):
    h = it['hdr']; s = h.index('# Samples'); e = h.index('Instructions Executed')
    tot = sum(int(r[s] or 0) for r in it['rows']); ex = sum(int(r[e] or 0) for r in it['rows'])
    print('instance %d: %s  rows %d samples %d warp-instr %d' % (k, it['name'][:60], len(it['rows']), tot, ex))
if which >= 0:
    it = inst[which]; h = it['hdr']; s = h.index('# Samples'); e = h.index('Instructions Executed'); src = h.index('Source')
    tot = sum(int(r[s] or 0) for r in it['rows'])
    order = sorted(range(len(it['rows'])), key=lambda i: -int(it['rows'][i][s] or 0))[:top]
    print('top %d of %d SASS lines by samples (line no, samples, %%, executed, sass):' % (top, len(it['rows'])))
    for i in sorted(order):
        r = it['rows'][i]
        print('%6d %6s %5.1f%% %8s  %s' % (i, r[s], 100.0 * int(r[s] or 0) / max(tot, 1), r[e], r[src].strip()[:110]))
