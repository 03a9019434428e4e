"""Throughput of ONE rank's share of a multi-GPU run, measured on a single GPU: the image partition of rank `rank` of `world`
(interleaved strips of 8 rows, as bench.py uses), `frames` frames per render. usage: python tools/tune_tile.py [world] [rank] [frames]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import brickmap_b200 as bm
from brickmap_b200 import renderer as R
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 0
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 5
W, H, STRIP = 1920, 1080, 8
rows, _ = bm.strip_rows_for_rank(H, rank, world, STRIP)
cfg = bm.default_config(screen_width=W, screen_height=H, tile_rows=rows, strip_rows=STRIP if world > 1 else 0, strip_count=world, strip_index=rank)
store = bm.SceneStore(cfg, resident=True)
ren = bm.Renderer(cfg, store)
ren.set_camera(bm.make_camera())
blit = torch.zeros(rows, W, 4, dtype=torch.float32, device="cuda")
ren.render(blit, frames, flags=R.FRAME_NO_UPLOAD)
best = 1e9
for _ in range(5):
    ren.set_sun(0.05, 0.1)
    ren.reset_stats()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ren.render(blit, frames, flags=R.FRAME_NO_UPLOAD)
    dt = time.perf_counter() - t0
    st = ren.stats()
    best = min(best, dt)
    rays = st["extend_rays"] + st["shadow_rays"]
print("world %d rank %d (%d rows), %d frames: ms/frame %.3f  Mrays/s %.0f  spp %.2f" % (world, rank, rows, frames, best / frames * 1e3, rays / best / 1e6, st["terminations"] / (rows * W)))
