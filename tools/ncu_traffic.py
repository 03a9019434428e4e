"""DRAM traffic of one frame-kernel launch from an ncu --set full capture -> profiles/frame_kernel_traffic.json (read by bench.py as
roofline.traffic). Stamped with a hash of the kernel sources and the bench config it was taken on: bench.py reports traffic only
when both match the tree and config it runs.   usage: python tools/ncu_traffic.py gpurun_out/<tag>_prof.ncu-rep [config=cfg3] [rays|-] [steady.csv]
steady.csv: the log of a second, single-pass capture of consecutive launches WITHOUT ncu's cache flush (--cache-control none --metrics
dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv): what a launch moves when its predecessor's L2 contents are still
there (the --set full capture flushes the L2 before every replay pass, so the survivors the previous frame left in L2 are gone)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from brickmap_b200.build import source_hash  # noqa: E402

rep = sys.argv[1]
config = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = {"report": os.path.basename(rep), "launches": []}
for vals in rows[2:]:
    d = dict(zip(hdr, zip(units, vals)))

    def get(name):
        u, v = d[name]
        v = float(v.replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9}.get(u, 1.0)
    out["launches"].append({"kernel": d["Kernel Name"][1][:60], "dram_bytes_read": get("dram__bytes_read.sum"), "dram_bytes_write": get("dram__bytes_write.sum"),
                            "duration_s": get("gpu__time_duration.sum")})
l0 = out["launches"][0]
out["dram_bytes_per_launch"] = l0["dram_bytes_read"] + l0["dram_bytes_write"]
if len(sys.argv) > 3 and sys.argv[3] != "-":
    out["rays_in_launch"] = float(sys.argv[3])  # (from the driver script's stats, for bytes per ray)
if len(sys.argv) > 4:
    per = {}
    with open(sys.argv[4]) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(r["Metric Unit"], 1.0)
        if r["Metric Name"].startswith("dram__bytes"):
            per[r["ID"]] = per.get(r["ID"], 0.0) + v
    if per:
        out["steady"] = {"log": os.path.basename(sys.argv[4]), "launches": len(per), "dram_bytes_per_launch": sum(per.values()) / len(per),
                         "dram_bytes_of_each_launch": [per[k] for k in sorted(per, key=int)],
                         "method": "ncu single pass, --cache-control none, consecutive steady-state launches (mean)"}
path = os.path.join(ROOT, "profiles", "frame_kernel_traffic.json")
try:
    with open(path) as f:
        allc = json.load(f)
    if allc.get("source_hash") != source_hash() or "configs" not in allc:
        allc = {"source_hash": source_hash(), "configs": {}}  # captures of other kernel sources do not describe this tree
except Exception:
    allc = {"source_hash": source_hash(), "configs": {}}
allc["configs"][config] = out
with open(path, "w") as f:
    json.dump(allc, f, indent=1)
print(json.dumps(allc))
