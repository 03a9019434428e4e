"""The reference's own notion of a benchmark: PerformanceMeasure's camera tour (performance_measure.h:4-25,
performance_measure.cpp:65-104) as a scripted, headless multi-view measurement. For every viewpoint: reset, render `frames`
frames, report ms/frame and Mrays/s -- for this library and, with --reference, for the unmodified reference kernels.

Viewpoints with x > 4096 lie outside the default world: they exercise the AABB entry path (voxel.cuh:142-155) and the 8x8x8 box
LoD (voxel.cuh:212-214). The reference lists 9 positions but only 8 angle pairs (its 9th view reads out of bounds); the 9th view
reuses the 8th angle pair here.
usage: python tools/bench_views.py [--frames 32] [--reference]
"""
import argparse
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from brickmap_b200.views import ANGLES, POSITIONS, direction  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--reference", action="store_true")
    args = ap.parse_args()
    import torch
    import brickmap_b200 as bm
    from brickmap_b200 import renderer as R
    cfg = bm.default_config()
    store = bm.SceneStore(cfg, resident=True)
    ren = bm.Renderer(cfg, store)
    blit = torch.zeros(cfg.screen_height, cfg.screen_width, 4, dtype=torch.float32, device="cuda")
    ref = None
    if args.reference:
        from oracle import binding as ob
        ref = ob.Reference("4096", cfg.screen_width, cfg.screen_height)
        ref.generate()
        ref.force_resident()
    print("%-4s %-34s %10s %10s %10s %10s" % ("view", "position", "ms/frame", "Mrays/s", "ref ms", "ref Mrays/s"))
    for i, pos in enumerate(POSITIONS):
        d = direction(*ANGLES[min(i, len(ANGLES) - 1)])
        ren.set_camera(bm.make_camera(position=pos, direction=d))
        ren.render(blit, 4, flags=R.FRAME_NO_UPLOAD)  # warm-up + reset
        ren.set_sun(0.05, 0.1)
        ren.reset_stats()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ren.render(blit, args.frames, flags=R.FRAME_NO_UPLOAD)
        dt = time.perf_counter() - t0
        st = ren.stats()
        rays = st["extend_rays"] + st["shadow_rays"]
        line = "%-4d %-34s %10.3f %10.0f" % (i, str(pos), dt / args.frames * 1e3, rays / dt / 1e6)
        if ref is not None:
            ref.set_camera(ob.make_camera(position=pos, direction=d))
            ref.run_frames(4)
            ref.mark_sun_changed()
            ms, shadows = ref.run_frames(args.frames)
            line += " %10.3f %10.0f" % (ms / args.frames, (args.frames * ref.n_slots + shadows) / ms / 1e3)
        print(line)


if __name__ == "__main__":
    main()
