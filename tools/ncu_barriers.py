#!/usr/bin/env python3
"""Per-barrier stall samples of one kernel in an ncu source-page export, each with the nearest source lines around it.
usage: ncu_barriers.py <sass_source_page.csv> <all.sass from nvdisasm -g -c> <kernel name substring>"""
import csv, re, sys, linecache
allrows = list(csv.reader(open(sys.argv[1])))
secs = [i for i, r in enumerate(allrows) if r and r[0] == 'Kernel Name'] + [len(allrows)]
pick = [k for k in range(len(secs) - 1) if sys.argv[3] in allrows[secs[k]][1]][0]
rows = allrows[secs[pick]:secs[pick + 1]]; hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
sc = hdr.index('# Samples'); addrs = [int(r[0], 16) for r in data]; base = addrs[0]
cur = None; off2line = {}; inside = False
for line in open(sys.argv[2]):
    if line.startswith('.text.'): inside = sys.argv[3] in line; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2)), m.group(1)); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*)', line)
    if m: off2line[int(m.group(1), 16)] = cur
tot = sum(float(r[sc] or 0) for r in data); out = []
for k, (r, a) in enumerate(zip(data, addrs)):
    if 'BAR' not in r[1] and 'WARPSYNC' not in r[1]: continue
    # samples are attributed to the instruction after the stalling one as well: take this and the next row
    smp = float(r[sc] or 0) + (float(data[k + 1][sc] or 0) if k + 1 < len(data) else 0)
    prev = next((off2line.get(addrs[j] - base) for j in range(k - 1, -1, -1) if off2line.get(addrs[j] - base) and off2line[addrs[j] - base][0] != 'ifx_base.h'), None)
    nxt = next((off2line.get(addrs[j] - base) for j in range(k + 1, len(data)) if off2line.get(addrs[j] - base) and off2line[addrs[j] - base][0] != 'ifx_base.h'), None)
    out.append((smp, r[1].strip(), prev, nxt))
for smp, ins, prev, nxt in sorted(out, key=lambda x: -x[0])[:18]:
    f = lambda x: "%s:%d %s" % (x[0], x[1], linecache.getline(x[2], x[1]).strip()[:70]) if x else "-"
    print("%5.1f%% %-28s after [%s]  before [%s]" % (100 * smp / tot, ins[:28], f(prev), f(nxt)))
