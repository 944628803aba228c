"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total
time and share (B200_PROFILING.md: compare SHARES, not absolutes).  usage: summarize_launches.py in.csv [title]"""
import csv, sys, collections, re
rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("nsecond", "ns") else v if unit in ("usecond", "us") else v * 1000.0 if unit in ("msecond", "ms") else v
    name = r["Kernel Name"]
    name = re.sub(r"\(.*", "", name)[:90]
    rows.append((name, us))
agg = collections.OrderedDict()
for n, us in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(a[1] for a in agg.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
print(f"{'kernel':<92}{'launches':>9}{'total_us':>13}{'share':>8}")
for n, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:<92}{c:>9}{us:>13.1f}{100*us/tot:>7.1f}%")
print(f"{'TOTAL':<92}{len(rows):>9}{tot:>13.1f}")
