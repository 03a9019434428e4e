"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.   usage: python tools/launch_summary.py launches.csv > summary.txt"""
import csv
import sys

with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
tot, cnt = {}, {}
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r["Metric Unit"], 1.0)
    k = r["Kernel Name"]
    tot[k] = tot.get(k, 0.0) + v
    cnt[k] = cnt.get(k, 0) + 1
allt = sum(tot.values())
print("ncu --metrics gpu__time_duration.sum --clock-control none -c 400 of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline` (first 400 launches: scene generation, bind, "
      "frame probe, counting pass with frame_kernel<0,1>, warm-up)")
for k in sorted(tot, key=lambda k: -tot[k]):
    print("%-44s   launches=%4d  total %10.1f us  avg %8.1f us  %5.1f%%" % (k[:44], cnt[k], tot[k], tot[k] / cnt[k], 100 * tot[k] / allt))
fq = sum(v for k, v in tot.items() if "frame_kernel_q" in k)
sc = sum(v for k, v in tot.items() if "scan_kernel" in k)
if fq and sc:
    print("frame_kernel_q : scan_kernel = %.1f : 1 (share of a step's kernel time: %.1f %% / %.1f %%)" % (fq / sc, 100 * fq / (fq + sc), 100 * sc / (fq + sc)))
