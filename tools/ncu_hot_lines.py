#!/usr/bin/env python3
"""Map the warp-stall samples of an ncu report (source page, SASS) to CUDA source lines using nvdisasm -g line markers.
usage: ncu_hot_lines.py <sass_source_page.csv> <all.sass from nvdisasm -g -c> <kernel name substring> [top]"""
import csv, re, sys, linecache
allrows = list(csv.reader(open(sys.argv[1])))
# the export holds one section per kernel: ['Kernel Name', name], header row, instruction rows
secs = [i for i, r in enumerate(allrows) if r and r[0] == 'Kernel Name'] + [len(allrows)]
pick = [k for k in range(len(secs) - 1) if sys.argv[3] in allrows[secs[k]][1]][0]
rows = allrows[secs[pick]:secs[pick + 1]]; hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
sc = hdr.index('# Samples')
def f(x):
    try: return float(x)
    except: return 0.0
stall_cols = [(h, k) for k, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
addrs = [int(r[0], 16) for r in data]; base = addrs[0]
cur = None; off2line = {}; inside = False
for line in open(sys.argv[2]):
    if line.startswith('.text.'): inside = sys.argv[3] in line; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*)', line)
    if m: off2line[int(m.group(1), 16)] = cur
agg = {}; tot = 0
for r, a in zip(data, addrs):
    k = off2line.get(a - base); s = f(r[sc]); tot += s
    e = agg.setdefault(k, [0, {}]); e[0] += s
    for h, c in stall_cols: e[1][h] = e[1].get(h, 0) + f(r[c])
for k, (v, st) in sorted(agg.items(), key=lambda x: -x[1][0])[: int(sys.argv[4]) if len(sys.argv) > 4 else 25]:
    src = linecache.getline(k[0], k[1]).strip()[:110] if k else ''
    top = ", ".join("%s %.0f%%" % (h[6:], 100 * x / max(v, 1)) for h, x in sorted(st.items(), key=lambda y: -y[1])[:2])
    print("%5.1f%% %s:%s | %s | %s" % (100 * v / tot, k[0].split('/')[-1] if k else None, k[1] if k else '', src, top))
