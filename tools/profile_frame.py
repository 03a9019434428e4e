"""Short driver for ncu captures of the frame kernel:
    ncu --set full --clock-control none --import-source on -k regex:frame_kernel_q -s 10 -c 1 -o gpurun_out/prof python tools/profile_frame.py [frames] [cfg3|cfg4]
cfg3: the benchmark view on the stock terrain, bricks resident. cfg4: the 8192^3 cave world of bench.py --config cfg4 (first camera
position of its path), streaming from an empty device scene with a 65 536-entry request queue."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import brickmap_b200 as bm  # noqa: E402
from brickmap_b200 import renderer as R  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 14
which = sys.argv[2] if len(sys.argv) > 2 else "cfg3"
if which == "cfg4":
    import bench
    cfg = bm.default_config(grid_size=8192, grid_height=8192, brick_load_queue_size=65536)
    store = bm.SceneStore(cfg, kind=R.SCENE_CAVES, seed=1, resident=False)
    ren = bm.Renderer(cfg, store)
    pos, d = bench.caves_camera(8192, 0, lambda sc: store.indices(sc, host_view=True))
    ren.set_camera(bm.make_camera(position=pos, direction=d))
    blit = torch.zeros(cfg.screen_height, cfg.screen_width, 4, dtype=torch.float32, device="cuda")
    for _ in range(frames):
        ren.render(blit, 1, sync=False)
        store.process_load_queue(ren.stream)
    ren.synchronize()
else:
    cfg = bm.default_config()
    store = bm.SceneStore(cfg, resident=True)
    ren = bm.Renderer(cfg, store)
    ren.set_camera(bm.make_camera())
    blit = torch.zeros(cfg.screen_height, cfg.screen_width, 4, dtype=torch.float32, device="cuda")
    ren.render(blit, frames, flags=R.FRAME_NO_UPLOAD)
print(ren.stats())
