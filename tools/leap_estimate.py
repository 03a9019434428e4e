"""Design aid (CPU, oracle rays): how many cell-grid DDA steps are left after the open-sky exit, and how many of those lie in long runs
ABOVE the per-column tops of the terrain -- the steps an exact "leap" of a descending ray through known-empty space could replace
(exact because the state after all steps with tmax < T is, per axis, tmax advanced n_a times: the interleaving does not matter).
Uses the rays cached by tools/simt_model.py (/tmp/simt_model.npz).   usage: python tools/leap_estimate.py > profiles/r2_y_leap_estimate.txt"""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, ROOT)
import importlib.util
spec = importlib.util.spec_from_file_location("sm", os.path.join(ROOT, "tools", "simt_model.py")); sm = importlib.util.module_from_spec(spec)
sys.argv = ["x"]; spec.loader.exec_module(sm)
z = np.load("/tmp/simt_model.npz")
occ, o, d, dist, kind = z["occ"], z["o"], z["d"], z["dist"], z["kind"]
seq, ns = sm.walk(occ, o, d, dist)
n = len(o)
# column tops: 16-cell columns, grown by two cells
colmax = np.where(occ.any(0), 63 - np.argmax(occ[::-1], axis=0), -1)  # [y,x] highest nonempty z per cell column
g = colmax.copy()
for _ in range(2):
    p = np.pad(g, 1, constant_values=-1)
    g = np.max([p[1:-1,1:-1], p[:-2,1:-1], p[2:,1:-1], p[1:-1,:-2], p[1:-1,2:], p[:-2,:-2], p[:-2,2:], p[2:,:-2], p[2:,2:]], axis=0)
for sh in (4, 3, 2):
    c = 1 << sh
    top = g.reshape(512 // c, c, 512 // c, c).max(axis=(1, 3))
    valid = seq[:, :, 0] >= 0
    x = np.where(valid, seq[:, :, 0], 0).astype(np.int64); y = np.where(valid, seq[:, :, 1], 0).astype(np.int64); zc = np.where(valid, seq[:, :, 2], 0).astype(np.int64)
    above = valid & (zc > top[y >> sh, x >> sh] + 1)
    tot = valid.sum()
    # sky exit: rays with dz >= 0: at chunk boundaries (every 32 steps), if all remaining are above -> removed
    rem_above = np.flip(np.logical_and.accumulate(np.flip(above | ~valid, 1), 1), 1)  # all remaining above
    asc = d[:, 2] >= 0
    sky_removed = np.zeros(n, np.int64)
    for r in np.nonzero(asc)[0]:
        for s0 in range(0, ns[r], 32):
            if rem_above[r, s0]:
                sky_removed[r] = ns[r] - s0
                break
    after_sky = tot - sky_removed.sum()
    print("sh", sh, "total steps/ray %.1f, after sky exit %.1f" % (tot / n, after_sky / n))
    # leap: runs of above-steps among steps not removed
    keep = valid.copy()
    for r in np.nonzero(sky_removed)[0]:
        keep[r, ns[r] - sky_removed[r]:] = False
    ab = above & keep
    for thr in (8, 12, 16, 24, 32):
        saved = 0; leaps = 0
        for r in range(n):
            a = ab[r, :ns[r]].astype(np.int8)
            if not a.any(): continue
            dif = np.diff(np.concatenate([[0], a, [0]]))
            st = np.nonzero(dif == 1)[0]; en = np.nonzero(dif == -1)[0]
            ln = en - st
            m = ln >= thr
            saved += (ln[m] - 3).sum(); leaps += m.sum()
        print("  thr %d: saved steps/ray %.1f (%.0f%% of remaining), leaps/ray %.2f" % (thr, saved / n, 100 * saved / after_sky, leaps / n))
    for k in (0, 1):
        m = kind == k
        print("  kind", k, "rays", m.sum(), "steps/ray after sky %.1f" % (keep[m].sum() / m.sum()), "above-steps %.1f" % (ab[m].sum() / m.sum()), " descending frac %.2f" % ((d[m, 2] < 0).mean()))
