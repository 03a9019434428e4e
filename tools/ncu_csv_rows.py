"""Per-launch rows of an `ncu --csv --metrics ...` log as one line: MB read / MB written / microseconds per launch.   usage: python tools/ncu_csv_rows.py log.csv"""
import csv
import sys

rows = {}
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v *= {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6, "ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3, "second": 1e6}.get(u, 1.0)
    rows.setdefault(r["ID"], {})[r["Metric Name"]] = v
out = []
for k in sorted(rows, key=int):
    d = rows[k]
    out.append("%.1f R + %.1f W MB, %.0f us" % (d.get("dram__bytes_read.sum", -1), d.get("dram__bytes_write.sum", -1), d.get("gpu__time_duration.sum", -1)))
print(" | ".join(out))
