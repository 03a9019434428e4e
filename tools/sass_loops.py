"""SASS excerpts of the two hot loops of the shipped throughput kernel (frame_kernel_q<STOCK=true, RECORD=false>), for profiles/:
the cell-level DDA iteration of trace_run (emptiness test + step), the per-cell mask test behind a non-empty block, and the voxel
step of intersect_brick. Also greps the mnemonics that prove the bulk-copy prologue.   usage: python tools/sass_loops.py > profiles/<name>.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "brickmap_b200", "libbrickmap_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
fn, lines = None, []
for ln in sass:
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = m.group(1)
        continue
    if fn and "frame_kernel_qILb1ELb0EE" in fn:
        ln = re.sub(r"/\* 0x[0-9a-f]+ \*/", "", ln).rstrip()
        if re.search(r"/\*[0-9a-f]{4,5}\*/", ln):
            lines.append(ln.strip())


def find(pattern, start=0, last=False):
    idx = [i for i in range(start, len(lines)) if re.search(pattern, lines[i])]
    return idx[-1] if last else idx[0]


def addr(i):
    return int(re.search(r"/\*([0-9a-f]{4,5})\*/", lines[i]).group(1), 16)


def show(title, a, b):
    print("---- %s: %d instructions" % (title, b - a + 1))
    for i in range(a, b + 1):
        print("   ", lines[i])
    print()


sys.path.insert(0, ROOT)
from brickmap_b200.build import source_hash  # noqa: E402
print("libbrickmap_b200.so, kernel sources %s, frame_kernel_q<STOCK, !RECORD>: %d SASS instructions" % (source_hash(), len(lines)))
print("bulk-copy prologue (cp.async.bulk + mbarrier): " + ", ".join(sorted({re.search(r"(UBLKCP\S*|SYNCS\.\S+)", l).group(1) for l in lines if re.search(r"UBLKCP|SYNCS\.", l)})))
print()
lds = find(r"LDS R\d+, \[R\d+\] ;", last=True)              # the bitmap word of the cell the DDA stands in
bra = find(r"BRA", lds)                                      # -> the step, when the block is empty
start = lds
while not re.search(r"BSSY|MOV|FSEL", lines[start - 1]) or re.search(r"BSSY", lines[start - 1]) and re.search(r"SHF", lines[start - 2]):
    start -= 1
    if lds - start > 12:
        break
show("cell loop, part 1: is the 4x4x4-cell block of the current cell possibly non-empty? (shared-memory bitmap)", start, bra)
target = int(re.search(r"BRA 0x([0-9a-f]+)", lines[bra]).group(1), 16)
fine_end = find(r"BRA 0x%x" % target, bra + 1)
show("behind a non-empty block: the cell's own bit (64-bit mask per block, global / L1)", bra + 1, fine_end)
step = [i for i in range(len(lines)) if addr(i) == target][0]
back = find(r"@P\d BRA", step)
show("cell loop, part 2: the DDA step (voxel.cuh:249-258) + chunk counter + back edge", step, back)
shf = [i for i in range(len(lines)) if "SHF.R.U64" in lines[i]]
b_end = find(r"BRA", shf[-1])
tgt = int(re.search(r"BRA 0x([0-9a-f]+)", lines[b_end]).group(1), 16)
b_start = [i for i in range(len(lines)) if addr(i) == tgt][0]
show("brick walk (intersect_brick, one block of PTX): one voxel step incl. exit test, z-slice reload, bit test", b_start, b_end)
# survivor records: stores that carry the L2 evict_last policy (the descriptor is built in uniform registers right before) and the drop
# of a line after both of its records have been read (discard.global.L2 -> CCTL.E.RML2)
stg = [i for i in range(len(lines)) if re.search(r"STG\.E\.128", lines[i])]
if stg:
    first = stg[0]
    pol = [i for i in range(max(0, first - 30), first) if re.search(r"UMOV UR\d+, 0x(f0|140)|ULOP3|USHF", lines[i])]
    print("== survivor store with the L2 evict_last policy (createpolicy -> descriptor 0x14f0... in UR; BM_SURV_HINTS & 2)")
    for i in pol + stg[:4]:
        print("    " + lines[i])
    print()
rml = [i for i in range(len(lines)) if "CCTL.E.RML2" in lines[i]]
if rml:
    print("== drop of a dead survivor line from L2 without a write-back (discard.global.L2; BM_SURV_HINTS & 4)")
    for i in range(rml[0] - 3, rml[0] + 1):
        print("    " + lines[i])
    print()
