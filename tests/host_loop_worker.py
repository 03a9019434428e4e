"""Runs the reference's host main-loop body (main.cpp:142-146) in a FRESH process for one library variant and saves what
it produced. A fresh process matters: launch_kernels keeps function-local statics (frame counter, last camera;
kernel.cu:367-382) and device globals (start_position, kernel.cu:109) that nothing can reset.
usage: python host_loop_worker.py <variant> <out.npz> <frames>"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import binding as ob  # noqa: E402

variant, out, frames = sys.argv[1], sys.argv[2], int(sys.argv[3])
g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_256.npz"))
host = ob.Reference(variant, int(g["width"]), int(g["height"]))
host.generate()  # nothing resident: everything streams through the request queue
host.set_camera(ob.make_camera(position=g["cam_pos"], direction=g["cam_dir"]))
host.set_sun(0.05, 0.1)
host.lib.ref_frame(0)  # launch_kernels + buffer swap; queue not yet processed
cnt, pos = host.load_queue()
acc1 = host.read_accum()
host.process_load_queue()
for _ in range(frames):
    host.frame(process_queue=True)
np.savez(out, count=np.uint32(cnt), positions=pos, accum1=acc1, accum=host.read_accum(), indices=host.read_indices())
