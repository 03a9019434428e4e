"""ms per call of the reference-compatible entry point bm_launch_frame (Renderer.launch_kernels: every buffer the reference's
kernels leave + synchronisation, launch.h:6) on the benchmark view, next to bm_render. usage: python tools/tune_launch.py [frames]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import brickmap_b200 as bm
from brickmap_b200 import renderer as R
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = bm.default_config()
store = bm.SceneStore(cfg, resident=True)
ren = bm.Renderer(cfg, store)
ren.set_camera(bm.make_camera())
state = bm.State(cfg)
for _ in range(4):
    ren.launch_kernels(state, flags=R.FRAME_NO_UPLOAD)
    state.swap()
best = 1e9
for _ in range(3):
    ren.set_sun(0.05, 0.1)
    ren.reset_stats()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(frames):
        ren.launch_kernels(state, flags=R.FRAME_NO_UPLOAD)
        state.swap()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = ren.stats()
    best = min(best, dt)
    rays = st["extend_rays"] + st["shadow_rays"]
print("bm_launch_frame: ms/call %.3f  Mrays/s %.0f" % (best / frames * 1e3, rays / best / 1e6))
