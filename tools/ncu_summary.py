"""Summarise an .ncu-rep: key launch metrics + instruction share / active lanes per CUDA source line.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_warps']
for h, u, v in zip(hdr, units, vals):
    if h in want or ('average_warps_issue_stalled' in h and 'per_issue_active' in h and float(v) > 0.15):
        print("%-86s %-10s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg = collections.defaultdict(lambda: [0, 0, 0, ''])
cur, hdr = None, None
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r and r[0] == 'Line No':
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            ln, ie, te, smp = int(d['Line No']), int(d['Instructions Executed'] or 0), int(d['Thread Instructions Executed'] or 0), int(d['# Samples'] or 0)
        except ValueError:
            continue
        k = (cur, ln)
        agg[k][0] += ie
        agg[k][1] += te
        agg[k][2] += smp
        if not agg[k][3]:
            agg[k][3] = r[1][:100]
ti, ts = sum(v[0] for v in agg.values()), sum(v[2] for v in agg.values())
print("total warp-instructions %d, samples %d" % (ti, ts))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-16s %4d inst=%5.1f%% lanes=%5.1f smp=%5.1f%% | %s" % (k[0][:16], k[1], 100 * v[0] / max(ti, 1), v[1] / max(v[0], 1), 100 * v[2] / max(ts, 1), v[3].strip()))

# ---- per-region totals: regions = the __device__/__global__ functions of the current sources (start line .. next start), and the
#      "// ---- " sections inside the quantum kernel
import os
import re
CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "brickmap_b200", "csrc")
regions = {}
for fn in os.listdir(CSRC):
    starts = []
    for n, line in enumerate(open(os.path.join(CSRC, fn)), 1):
        m = re.match(r"^(?:__device__|__global__|static|template).*?(\w+)\s*\(", line)
        if m and not line.startswith("template <"):
            starts.append((n, m.group(1)))
        m = re.match(r"^\s*// ---- (.*?) -*$", line)
        if m and fn == "bm_frame_quantum.cuh":
            starts.append((n, "q: " + m.group(1).strip()[:40]))
    regions[fn] = [(a, (starts[k + 1][0] - 1) if k + 1 < len(starts) else 99999, nm) for k, (a, nm) in enumerate(starts)]
tot = collections.defaultdict(lambda: [0, 0, 0])
for (f, ln), v in agg.items():
    name = f + " other"
    for lo, hi, nm in regions.get(f, []):
        if lo <= ln <= hi:
            name = nm
            break
    tot[name][0] += v[0]
    tot[name][1] += v[1]
    tot[name][2] += v[2]
print("---- regions")
for nm, v in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print("%-40s inst=%5.1f%% lanes=%5.1f smp=%5.1f%%" % (nm, 100 * v[0] / max(ti, 1), v[1] / max(v[0], 1), 100 * v[2] / max(ts, 1)))
