#!/usr/bin/env python
"""Per-launch table from an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` log."""
import csv
import sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
d = {}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ki][:90]), {})[r[mi]] = float(r[vi].replace(',', ''))
for (i, k), m in sorted(d.items()):
    print(f"{i:4d} {m.get('gpu__time_duration.sum', 0) / 1e3:9.1f} us  rd {m.get('dram__bytes_read.sum', 0) / 1e6:8.0f} MB  wr {m.get('dram__bytes_write.sum', 0) / 1e6:8.0f} MB  {k}")
