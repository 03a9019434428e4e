"""Quick throughput probe of bm_render on the benchmark workload under the env-var knobs. usage: python tools/tune.py [frames]"""
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
blit = torch.zeros(cfg.screen_height, cfg.screen_width, 4, dtype=torch.float32, device="cuda")
ren.render(blit, frames, flags=R.FRAME_NO_UPLOAD)
best = 1e9
for _ in range(3):
    ren.set_sun(0.05, 0.1)
    ren.reset_stats()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ren.render(blit, frames, flags=R.FRAME_NO_UPLOAD)
    dt = time.perf_counter() - t0
    st = ren.stats()
    best = min(best, dt)
    rays = st["extend_rays"] + st["shadow_rays"]
print("%s ms/frame %.3f  Mrays/s %.0f" % ({k: v for k, v in os.environ.items() if k.startswith("BRICKMAP")}, best / frames * 1e3, rays / best / 1e6))
