"""BASELINE config 4: a sparse 8192^3 cave world (1024^3 cells, 262 144 superchunks) generated on the device, nothing resident
at the start, bricks streamed in through the request queue (<= 1024 per frame, voxel.cuh:228-241 / Scene.cpp:200-229) while a
1920x1080 view is path traced with both LoD levels active. Reports generation time, memory, Mrays/s and streaming progress.
usage: python tools/run_caves.py [grid_size=8192] [frames=64]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import brickmap_b200 as bm  # noqa: E402
from brickmap_b200 import renderer as R  # noqa: E402

grid = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cfg = bm.default_config(grid_size=grid, grid_height=grid)
t0 = time.perf_counter()
store = bm.SceneStore(cfg, kind=R.SCENE_CAVES, seed=1, resident=False)
torch.cuda.synchronize()
t_gen = time.perf_counter() - t0
free, total = torch.cuda.mem_get_info()
print("world %d^3 voxels: %d superchunks, %d bricks (%.1f GiB host-view store), generated on the device in %.1f s; GPU memory in use %.1f GiB"
      % (grid, store.superchunks, store.total_bricks, store.total_bricks * 64 / 2**30, t_gen, (total - free) / 2**30))
ren = bm.Renderer(cfg, store)
c = grid / 2
ren.set_camera(bm.make_camera(position=(c + 37.0, c - 91.0, c + 13.0), direction=(0.8017837, 0.5345225, 0.2672612)))
blit = torch.zeros(cfg.screen_height, cfg.screen_width, 4, dtype=torch.float32, device="cuda")
served = 0
ren.reset_stats()
torch.cuda.synchronize()
t0 = time.perf_counter()
for f in range(frames):
    ren.render(blit, 1, sync=False)              # upload of the previous frame's staged bricks + one frame
    store.process_load_queue(ren.stream)         # stage what this frame requested (device-side Scene::process_load_queue)
    if f % 16 == 15 or f == frames - 1:
        cnt, _ = ren.load_queue()
        st = ren.stats()
        dt = time.perf_counter() - t0
        print("frame %3d: requests in queue %5d, %.2f ms/frame, %.0f Mrays/s, spp %.2f" % (
            f + 1, cnt, dt / (f + 1) * 1e3, (st["extend_rays"] + st["shadow_rays"]) / dt / 1e6, st["terminations"] / (cfg.screen_width * cfg.screen_height)))
ren.synchronize()
assert torch.isfinite(blit).all()
print("alpha sum", float(blit[..., 3].double().sum()), "== terminations", ren.stats()["terminations"])
