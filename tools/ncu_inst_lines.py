#!/usr/bin/env python3
"""Executed warp instructions / thread instructions / stall samples of one kernel by CUDA source line (divergence shows as a low
"thr" = average active threads per warp instruction).
usage: ncu_inst_lines.py <`ncu -i rep --page source --csv --print-source sass --kernel-name K` export> <all.sass from nvdisasm -g -c> <kernel name substring> <top N>"""
import csv, re, sys, linecache
rows = list(csv.reader(open(sys.argv[1]))); secs=[i for i,r in enumerate(rows) if r and r[0]=='Kernel Name']+[len(rows)]; rows=rows[secs[0]:secs[1]]; hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
ie = hdr.index('Instructions Executed'); te = hdr.index('Thread Instructions Executed'); sc = hdr.index('# Samples')
addrs = [int(r[0], 16) for r in data]; base = addrs[0]
cur = None; off2line = {}; inside = False; n_in = 0
for line in open(sys.argv[2]):
    if line.startswith('.text.'): inside = sys.argv[3] in line; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1), int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*)', line)
    if m: off2line[int(m.group(1), 16)] = cur; n_in += 1
print("instrs in report", len(data), "in sass", n_in)
agg = {}; ti = tt = ts = 0
for r, a in zip(data, addrs):
    k = off2line.get(a - base); i = float(r[ie] or 0); t = float(r[te] or 0); s = float(r[sc] or 0)
    e = agg.setdefault(k, [0, 0, 0]); e[0] += i; e[1] += t; e[2] += s; ti += i; tt += t; ts += s
print("warp instr %.3g thread instr %.3g avg active threads %.1f" % (ti, tt, tt / ti))
for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:int(sys.argv[4])]:
    src = linecache.getline(k[0], k[1]).strip()[:120] if k else ''
    print("%5.1f%% inst %5.1f%% samp thr %4.1f %s:%s | %s" % (100 * v[0] / ti, 100 * v[2] / ts, v[1] / max(v[0], 1), k[0].split('/')[-1] if k else None, k[1] if k else '', src))
