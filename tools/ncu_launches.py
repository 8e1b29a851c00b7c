#!/usr/bin/env python3
"""Print kernel name, grid and duration (ms) from an `ncu --metrics gpu__time_duration.sum --csv` log."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hd = rows[h]
for r in rows[h + 1:]:
    if len(r) > hd.index("Metric Value"):
        print(f"{r[hd.index('Kernel Name')][:48]:50s} grid {r[hd.index('Grid Size')]:>14s}  {float(r[hd.index('Metric Value')].replace(',', '')) / 1e6:9.3f} ms")
