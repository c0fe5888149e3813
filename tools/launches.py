"""Print an ncu `--metrics gpu__time_duration.sum --csv` launch list compactly: python tools/launches.py file.csv [filter]."""
import csv
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for row in csv.DictReader(lines):
    rows.append((row["Kernel Name"], row.get("Grid Size", ""), float(row["Metric Value"].replace(",", "")) / 1000.0))
flt = sys.argv[2] if len(sys.argv) > 2 else ""
tot = 0.0
for i, (k, g, v) in enumerate(rows):
    if flt in k:
        print(f"{i:4d} {k[:60]:60s} {g:16s} {v:9.1f} us")
        tot += v
print("total", round(tot, 1), "us")
