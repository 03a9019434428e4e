"""CPU-side SIMT cost model of the cell-level DDA loop (design aid, not a measurement).

Takes the benchmark workload's real rays of one steady-state frame (oracle), walks their cell sequences and replays
warps of 32 consecutive slots in lock step under different loop designs, counting issued warp-instructions.
usage: python tools/simt_model.py [out.npz]   (caches the per-ray sequences)
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import binding as ob  # noqa: E402

CACHE = sys.argv[1] if len(sys.argv) > 1 else "/tmp/simt_model.npz"


def chebyshev_field(occ, cap):
    """occ: bool [z,y,x]; returns uint8 distance (in elements) to the nearest True, capped."""
    d = np.where(occ, 0, cap).astype(np.uint8)
    cur = occ.copy()
    for r in range(1, cap):
        nxt = cur.copy()
        for ax in range(3):
            a = np.roll(nxt, 1, ax)
            b = np.roll(nxt, -1, ax)
            # no wrap-around
            sl = [slice(None)] * 3
            sl[ax] = 0
            a[tuple(sl)] = False
            sl[ax] = -1
            b[tuple(sl)] = False
            nxt = nxt | a | b
        newly = nxt & ~cur
        d[newly] = r
        cur = nxt
    return d


def build():
    t0 = time.time()
    orc = ob.Oracle()
    sc = ob.OracleScene(orc, 4096, 512).generate_terrain().set_residency(True)
    print("scene %.1fs" % (time.time() - t0))
    idx = sc.all_host_indices().reshape(4, 32, 32, 16, 16, 16)  # sz sy sx lz ly lx
    occ = (idx != 0).transpose(0, 3, 1, 4, 2, 5).reshape(64, 512, 512)  # z y x
    w, h, n = 1920, 1080, 2 * 1048576
    ren = ob.OracleRenderer(sc, w, h, n, ob.make_camera())
    for _ in range(4):
        ren.frame()
    ren.primary_rays()
    ren.set_wavefront_globals()
    rays = ren.rays.copy()
    ren.extend()
    ext = ren.rays.copy()
    ren.shade()
    nsh = ren.state.shadow_ray_cnt
    sh = ren.shadows[:nsh].copy()
    print("frames %.1fs, shadows %d" % (time.time() - t0, nsh))
    # sample whole warps
    wsel = np.arange(0, n // 32, 61)
    slots = (wsel[:, None] * 32 + np.arange(32)[None, :]).reshape(-1)
    eo, ed = rays["origin"][slots], rays["direction"][slots]
    edist = ext["distance"][slots]
    wsel2 = np.arange(0, nsh // 32, 61)
    s2 = (wsel2[:, None] * 32 + np.arange(32)[None, :]).reshape(-1)
    so, sd = sh["origin"][s2], sh["direction"][s2]
    cam = np.array([64, 64, 37], np.int32)
    hit, sdist, _ = sc.trace(so, sd, cam)
    sdist = np.where(hit, sdist, 1e20).astype(np.float32)
    o = np.concatenate([eo, so]).astype(np.float64)
    d = np.concatenate([ed, sd]).astype(np.float64)
    dist = np.concatenate([edist, sdist]).astype(np.float64)
    kind = np.concatenate([np.zeros(len(eo), np.uint8), np.ones(len(so), np.uint8)])
    np.savez_compressed(CACHE, occ=occ, o=o, d=d, dist=dist, kind=kind)
    print("saved", CACHE)


def walk(occ, o, d, dist, maxsteps=1400):
    """cell sequences: returns pos[ray, step, 3] int16 (-1 after the end) and nsteps"""
    n = len(o)
    oc = o / 8.0
    pos = np.floor(oc).astype(np.int64)
    step = np.sign(d).astype(np.int64)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = np.where(d != 0, 1.0 / d, 0.0)
        cb = np.where(d > 0, pos + 1, pos).astype(np.float64)
        tmax = np.where(d != 0, (cb - oc) * inv, 1e6)
        tdelta = step * inv
    lim = np.array([512, 512, 64])
    tend = dist / 8.0  # in cell units of t
    out = np.full((n, maxsteps, 3), -1, np.int16)
    alive = np.ones(n, bool)
    tcur = np.zeros(n)
    for s in range(maxsteps):
        out[alive, s] = pos[alive]
        ax = np.where((tmax[:, 0] < tmax[:, 1]) & (tmax[:, 0] < tmax[:, 2]), 0, np.where(tmax[:, 1] < tmax[:, 2], 1, 2))
        tnext = tmax[np.arange(n), ax]
        # the ray ends inside this cell if the hit distance falls before the cell's exit
        ends = tend <= tnext + 1e-9
        alive &= ~ends
        pos[np.arange(n), ax] += step[np.arange(n), ax]
        tmax[np.arange(n), ax] += tdelta[np.arange(n), ax]
        inside = ((pos >= 0) & (pos < lim)).all(1)
        alive &= inside
        if not alive.any():
            break
    nsteps = (out[:, :, 0] >= 0).sum(1)
    return out, nsteps


def main():
    if not os.path.exists(CACHE):
        build()
    z = np.load(CACHE)
    occ, o, d, dist, kind = z["occ"], z["o"], z["d"], z["dist"], z["kind"]
    seq, nsteps = walk(occ, o, d, dist)
    print("rays %d (extend %d, shadow %d); mean cell tests per ray %.1f (extend %.1f, shadow %.1f)" % (
        len(o), (kind == 0).sum(), (kind == 1).sum(), nsteps.mean(), nsteps[kind == 0].mean(), nsteps[kind == 1].mean()))
    # fields
    b4 = occ.reshape(16, 4, 128, 4, 128, 4).any(axis=(1, 3, 5))
    D4 = chebyshev_field(b4, 15)
    Dc = chebyshev_field(occ, 33)
    valid = seq[:, :, 0] >= 0
    x, y, zc = seq[:, :, 0].astype(np.int64), seq[:, :, 1].astype(np.int64), seq[:, :, 2].astype(np.int64)
    xs, ys, zs = np.where(valid, x, 0), np.where(valid, y, 0), np.where(valid, zc, 0)
    d4 = np.where(valid, D4[zs >> 2, ys >> 2, xs >> 2], 255)
    dc = np.where(valid, Dc[zs, ys, xs], 255)
    cellocc = np.where(valid, occ[zs, ys, xs], False)
    tot = valid.sum()
    print("steps by 4^3-block distance D4: " + " ".join("%d:%.3f" % (k, (d4[valid] == k).sum() / tot) for k in range(0, 8)) + " >=8:%.3f" % ((d4[valid] >= 8) & (d4[valid] < 255)).sum().__truediv__(tot))
    print("steps by per-cell distance: " + " ".join("%d:%.3f" % (k, (dc[valid] == k).sum() / tot) for k in (0, 1, 2, 3, 4)) + " 5-8:%.3f 9-16:%.3f >16:%.3f" % (
        ((dc[valid] >= 5) & (dc[valid] <= 8)).sum() / tot, ((dc[valid] >= 9) & (dc[valid] <= 16)).sum() / tot, (dc[valid] > 16).sum() / tot))
    print("non-empty cells tested per ray: %.2f" % (cellocc.sum() / len(o)))
    np.savez_compressed(CACHE.replace(".npz", "_seq.npz"), d4=d4.astype(np.uint8), dc=dc.astype(np.uint8), cellocc=cellocc, nsteps=nsteps, kind=kind)

    nw = len(o) // 32

    def simulate(free_steps, L, S, S1, SLOW, per_lane_loop=True, chunk=4):
        """free_steps(D4 value) -> unchecked steps after a lookup (>= 1). Returns (warp instructions, lane-steps)."""
        total = 0
        for wv in range(nw):
            sl = slice(wv * 32, wv * 32 + 32)
            dd = d4[sl]
            ns = nsteps[sl]
            co = cellocc[sl]
            i = np.zeros(32, np.int64)
            while True:
                act = i < ns
                if not act.any():
                    break
                cur = dd[np.arange(32), np.minimum(i, ns - 1).clip(0)].astype(np.int64)
                nfree = np.where(act, free_steps(cur), 0)
                slow = act & (cur == 0)
                cost = L + S1  # lookup + the one step everybody takes
                if slow.any():
                    cost += SLOW
                extra = (nfree - 1).max()
                cost += S * extra
                total += cost
                i = i + np.maximum(nfree, 1)
        return total

    lane_steps = nsteps.sum()
    base = simulate(lambda c: np.ones_like(c), L=18, S=0, S1=22, SLOW=16)
    print("baseline (1 step per lookup, 40 instr): %.1f warp-instr per ray-step (ideal %.2f)" % (base / lane_steps, 40 / 32))
    for name, fs in (("2-bit D (1/5/9)", lambda c: 1 + 4 * (np.clip(c, 1, 3) - 1)),
                     ("3-bit D (..25)", lambda c: 1 + 4 * (np.clip(c, 1, 7) - 1)),
                     ("4-bit D (..57)", lambda c: 1 + 4 * (np.clip(c, 1, 15) - 1))):
        t = simulate(fs, L=15, S=10, S1=10, SLOW=16)
        print("%s: %.2f warp-instr per ray-step -> %.2fx fewer than baseline" % (name, t / lane_steps, base / t))


if __name__ == "__main__":
    main()
