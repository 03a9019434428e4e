"""Short driver for ncu captures of the frame kernel on the benchmark workload:
    ncu --set full --clock-control none --import-source on -k regex:frame_kernel -s 10 -c 2 -o gpurun_out/prof python tools/profile_frame.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import brickmap_b200 as bm  # noqa: E402
from brickmap_b200 import renderer as R  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 14
cfg = bm.default_config()
store = bm.SceneStore(cfg, resident=True)
ren = bm.Renderer(cfg, store)
ren.set_camera(bm.make_camera())
blit = torch.zeros(cfg.screen_height, cfg.screen_width, 4, dtype=torch.float32, device="cuda")
ren.render(blit, frames, flags=R.FRAME_NO_UPLOAD)
print(ren.stats())
