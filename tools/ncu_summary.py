#!/usr/bin/env python3
"""Markdown summary of (a) an ncu launch list (--metrics gpu__time_duration.sum --csv) and (b) an ncu --set full raw-page export.
usage: ncu_summary.py <tag> <launches.csv> <raw_page.csv> [workload text] [bench flags of the launch list]"""
import csv, sys, collections
tag, lpath, rpath = sys.argv[1:4]
wl = sys.argv[4] if len(sys.argv) > 4 else 'configs[1]'; flags = sys.argv[5] if len(sys.argv) > 5 else '--steps 3 --warmup 1 --no-cpu-baseline'
rows = [r for r in csv.reader(l for l in open(lpath) if l.startswith('"'))]
hdr = rows[0]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value"); mn = hdr.index("Metric Name")
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[mn] != "gpu__time_duration.sum": continue
    name = r[kn].split("(")[0]; e = agg.setdefault(name, [0, 0.0]); e[0] += 1; e[1] += float(r[mv].replace(",", "")) / 1e6
tot = sum(v[1] for v in agg.values())
print("# %s launch list (ncu --metrics gpu__time_duration.sum --clock-control none; `python bench.py %s`, %s)\n" % (tag, flags, wl))
print("Cold-cache, serialised per-launch times: compare SHARES, not absolutes (the raw CSV is `%s`).\n" % lpath.split("/")[-1])
print("| kernel | launches | total ms | share | avg ms |\n|---|---|---|---|---|")
for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("| %s | %d | %.3f | %.1f%% | %.3f |" % (k, n, ms, 100 * ms / tot, ms / n))
raw = list(csv.reader(open(rpath))); h = raw[0]; units = raw[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum", "smsp__average_warp_latency_per_inst_issued.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
print("\n# %s `ncu --set full` of the heaviest kernels (one launch each, `bench.py --steps 1 --warmup 0 --no-cpu-baseline`)" % tag)
def num(x):
    try: return float(x.replace(",", ""))
    except ValueError: return 0.0
for r in raw[2:]:
    print("\n## %s\n" % r[h.index("Kernel Name")].split("(")[0])
    for w in want:
        if w in h: print("* `%s` = %s %s" % (w, r[h.index(w)], units[h.index(w)]))
    rd, wr = num(r[h.index("dram__bytes_read.sum")]), num(r[h.index("dram__bytes_write.sum")])
    scale = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}
    print("* DRAM traffic (read+write) = %.1f MB per launch" % (rd * scale.get(units[h.index("dram__bytes_read.sum")], 1) + wr * scale.get(units[h.index("dram__bytes_write.sum")], 1)))
