#!/usr/bin/env python
"""Benchmarks of the path-tracing hot path. One JSON line per run (rank 0).

    python bench.py [--config cfg3] [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--spp S]

--config (BASELINE.json configs[1..4] + the reference's own camera tour; default cfg3 = the headline):
  cfg2  4096-class terrain, 1920x1080, PRIMARY RAYS ONLY: primary_rays + extend (kernel.cu:416-418), one frame of 2 097 152 slots
        (1.011 rays per pixel) per step; result = the post-extend ray records.
  cfg3  4096-class terrain, 1920x1080, 16 spp full path trace (<= 4 segments + sun shadow rays).            <- headline metric
  cfg4  8192^3 sparse cave world generated on the device, nothing resident at the start, request queue 65 536, both LoD levels,
        1920x1080 x 16 spp; every step moves the camera along a path, so every step streams new bricks in (upload, requests,
        process_load_queue every frame inside the timed region).
  cfg5  cfg3's scene and view at 3840x2160 x 64 spp (meant for --gpus 2/4/8).
  tour  cfg3 from the nine viewpoints of the reference's PerformanceMeasure (performance_measure.h:4-25); views 4-8 are outside
        the world (AABB entry + 8x8x8 LoD).

Scene (cfg2/3/5/tour): the reference's procedural terrain (Scene.cpp:44-116) at its stock constants, 4096 x 4096 x 512 voxels
("4096^3-class"), every brick resident, camera (512,512,300) looking along +x (camera.h:4-5), sun (0.05, 0.1).

One STEP (cfg3/4/5/tour) = reset the accumulation buffer, then frames of <= 2 097 152 segment slots (variables.h:61) until every
pixel has exactly `spp` finished paths (BM_FRAME_EXACT_PATHS: the last frames only take the primaries that are still missing;
--overshoot restores round 1's "whole frames until sum(alpha) >= spp w h").  1 ray = 1 intersect_voxel call (extend segment or
shadow ray).  value = rays of all ranks / max-over-ranks device time.

N > 1 (torchrun): the image is split into interleaved 8-row strips, one set per GPU, the brick store is replicated, the only
exchange is the all-gather of the request buffer (NCCL). Total work is fixed -> "scaling": "strong".

--impl reference runs the UNMODIFIED reference kernels (oracle/_ref/libbrickmap_ref_4096.so: kernel.cu, voxel.cuh, sunsky.cu,
Scene.cpp compiled for sm_100a) on the same GPU, same scene/camera/step definition. The reference is a CUDA program: it has no CPU
implementation of this path, so its own GPU kernels are what is timed. cfg4 is outside what the reference's
compile-time constants and terrain generator can express: there the reference arm is the CPU oracle on a bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SUN = (0.05, 0.1)
STRIP = 8  # rows per strip of the multi-GPU image partition
N_SLOTS = 2 * 1048576
METRIC = "Mrays/s at 1920x1080x16spp (4096^3 scene); HBM GB/s vs roofline"
TERRAIN = "4096x4096x512 procedural terrain (reference constants), bricks resident"
CONFIGS = {
    "cfg2": dict(width=1920, height=1080, spp=1, scene="terrain", primary_only=True,
                 metric="Mrays/s at 1920x1080x1spp primary rays only (4096^3 scene); HBM GB/s vs roofline",
                 workload=TERRAIN + ", 1920x1080, primary rays only: primary_rays + extend over one frame of 2 097 152 slots (1.011 rays per pixel)"),
    "cfg3": dict(width=1920, height=1080, spp=16, scene="terrain", metric=METRIC,
                 workload=TERRAIN.replace(", bricks resident", "") + ", 1920x1080, 16 spp full path trace (<=4 segments + sun shadow rays), bricks resident"),
    "cfg4": dict(width=1920, height=1080, spp=16, scene="caves", grid=8192, queue=65536,
                 metric="Mrays/s at 1920x1080x16spp (8192^3 sparse cave scene, streaming + LoD); HBM GB/s vs roofline",
                 workload="8192^3 sparse cave world (device-generated, 262 144 superchunks, 4 GiB of index words), streaming from an empty device scene through a "
                          "65 536-entry request queue, both LoD levels, 1920x1080, 16 spp full path trace, camera moves every step"),
    "cfg5": dict(width=3840, height=2160, spp=64, scene="terrain",
                 metric="Mrays/s at 3840x2160x64spp (4096^3 scene), image tiles over the GPUs; HBM GB/s vs roofline",
                 workload=TERRAIN + ", 3840x2160, 64 spp full path trace (<=4 segments + sun shadow rays)"),
    "tour": dict(width=1920, height=1080, spp=16, scene="terrain", tour=True,
                 metric="Mrays/s at 1920x1080x16spp over the reference's 9 benchmark views (4096^3 scene); HBM GB/s vs roofline",
                 workload=TERRAIN + ", 1920x1080, 16 spp full path trace from the 9 PerformanceMeasure viewpoints (performance_measure.h:4-25), one step = all 9"),
}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def flush_l2(torch, scratch):
    scratch.add_(1)  # writes a buffer larger than the 126 MB L2


class quiet_stdout:
    """The reference's Scene::generate and NCCL print to stdout; keep stdout to the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def views_for(conf):
    """[(position, direction)] rendered per step."""
    if conf.get("tour"):
        from brickmap_b200.views import tour
        return tour()
    if conf["scene"] == "caves":
        return None  # a camera path, see caves_camera
    return [((512.0, 512.0, 300.0), (1.0, 0.0, 0.0))]


def caves_camera(grid, k, indices_of=None):
    """Camera of step k in the cave world: a walk along the view direction from the centre, 96 voxels (12 cells) per step, snapped to
    open space. indices_of(superchunk) -> the 4096 host-view index words of that superchunk (0 = empty cell): the camera is put at the
    centre of the nearest empty cell whose six neighbours are empty too -- a camera inside rock renders nothing (every primary hits at
    distance 0 and bounces off a zero normal into a NaN ray, like in the reference)."""
    import numpy as np
    c = grid / 2
    d = (0.8017837, 0.5345225, 0.2672612)
    p = np.array([c + 37.0 + 96.0 * k * d[0], c - 91.0 + 96.0 * k * d[1], c + 13.0 + 96.0 * k * d[2]])
    if indices_of is not None:
        sg = grid // 128
        s3 = (p // 128).astype(int)
        idx = np.asarray(indices_of(int(s3[0] + sg * (s3[1] + sg * s3[2])))).reshape(16, 16, 16)  # [z][y][x]
        occ = np.pad(idx != 0, 1, constant_values=True)
        ok = ~(occ[1:-1, 1:-1, 1:-1] | occ[:-2, 1:-1, 1:-1] | occ[2:, 1:-1, 1:-1] | occ[1:-1, :-2, 1:-1] | occ[1:-1, 2:, 1:-1] | occ[1:-1, 1:-1, :-2] | occ[1:-1, 1:-1, 2:])
        if ok.any():
            z, y, x = np.nonzero(ok)
            centres = np.stack([x, y, z], 1) * 8.0 + 4.37 + s3 * 128.0
            p = centres[np.argmin(((centres - p) ** 2).sum(1))]
    return (float(p[0]), float(p[1]), float(p[2])), d


# ---------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port on the host cores, bounded sample
# ---------------------------------------------------------------------------------------------------------------------
def cpu_baseline_port(conf, frames=3):
    """The CPU oracle (oracle/oracle.cpp, a scalar C++ port of the same algorithm) on the host cores, bounded sample."""
    from oracle import binding as ob
    orc = ob.Oracle()
    cores = orc.hardware_threads()
    w, h = conf["width"], conf["height"]
    if conf["scene"] == "caves":
        # the oracle cannot generate 2^39 voxels in bounded time: the same generator at 1024^3 (1/512 of the volume), all resident
        grid = 1024
        scene = ob.OracleScene(orc, grid, grid).generate_caves(seed=1).set_residency(True)
        pos, d = caves_camera(grid, 0, scene.host_indices)
        cam = ob.make_camera(position=pos, direction=d)
        what = "the same cave generator at 1024^3 voxels (1/512 of the volume), all bricks resident, first %d frames" % frames
    else:
        scene = ob.OracleScene(orc, 4096, 512).generate_terrain().set_residency(True)
        pos, d = views_for(conf)[0]
        cam = ob.make_camera(position=pos, direction=d)
        what = "first %d frames of the same workload (view 0)" % frames
    ren = ob.OracleRenderer(scene, w, h, N_SLOTS, cam, SUN)
    t0 = time.perf_counter()
    if conf.get("primary_only"):
        for _ in range(frames):
            ren.state.primary_ray_cnt = 0
            ren.primary_rays()
            ren.set_wavefront_globals()
            ren.extend(threads=0)
            ren.state.frame += 1
        what = "%d frames of primary_rays + extend, same workload" % frames
    else:
        for _ in range(frames):
            ren.frame(threads=0)
    dt = time.perf_counter() - t0
    rays = ren.stats.rays
    out = {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
           "sample": "%s (%d rays); extend and connect (the traversal, >95%% of the scalar work) multi-threaded over all cores, "
                     "ray generation and the per-slot shading work threaded too; only the append of survivor / shadow records and the accumulation run sequentially, in slot order (the canonical compaction order)" % (what, rays),
           "seconds": dt}
    if conf["scene"] != "caves" and not conf.get("primary_only"):
        # untimed: one more frame with the footprint instrumentation on -> minimum brick-index footprint (32 B x unique
        # index-word and brick sectors the reference algorithm touches in the frame), the comparator of roofline.traffic
        scene.footprint_begin()
        ren.frame(threads=0)
        fp = ob.Stats()
        scene.footprint_report(fp)
        fp_rays = ren.stats.rays - rays
        out["min_footprint"] = {"bytes_per_frame": 32 * (fp.unique_index_sectors + fp.unique_brick_sectors),
                                "bytes_per_ray": 32 * (fp.unique_index_sectors + fp.unique_brick_sectors) / max(fp_rays, 1), "frame": frames + 1,
                                "note": "32 B x (unique index-word sectors + unique brick sectors) of one frame, from the oracle"}
    return out


# ---------------------------------------------------------------------------------------------------------------------
# reference arm
# ---------------------------------------------------------------------------------------------------------------------
def run_reference(args, conf, rank, world):
    """Reference arm: the unmodified reference kernels through launch_kernels (kernel.cu:366) on the GPU."""
    if rank != 0:
        return
    from oracle import binding as ob
    base = {"impl": "reference", "metric": conf["metric"], "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    if conf["scene"] == "caves":
        # not expressible with the reference's compile-time constants and terrain generator: CPU oracle on a bounded sample
        cpu = cpu_baseline_port(conf, frames=max(1, min(args.steps, 3)))
        print(json.dumps(dict(base, value=cpu["value"], ms_per_step=cpu["seconds"] * 1e3 / max(1, min(args.steps, 3)),
                              config={"workload": conf["workload"], "note": "the reference cannot run this configuration (compile-time 4096x4096x512 terrain); "
                                      "CPU oracle port on a bounded sample instead"},
                              clocks=None, gpu_launches=0, cpu_baseline=cpu, e2e={"value": cpu["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})))
        return
    if not ob.Reference.available("4096"):
        print(json.dumps(dict(base, unavailable="oracle/_ref/libbrickmap_ref_4096.so not built (run make -C oracle ref where /root/reference exists)")))
        return
    import torch
    w, h = conf["width"], conf["height"]
    ref = ob.Reference("4096", w, h, device=0)
    with quiet_stdout():
        ref.generate()
    ref.force_resident()
    ref.set_sun(*SUN)
    views = views_for(conf)
    n_slots = ref.n_slots
    scratch = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    sampler = ClockSampler(0)
    if conf.get("primary_only"):
        ref.set_camera(ob.make_camera(position=views[0][0], direction=views[0][1]))
        ref.upload_sun()
        for _ in range(args.warmup):
            ref.run_primary_extend(1, 1)
        sampler.start()
        total_ms = 0.0
        for s in range(args.steps):
            flush_l2(torch, scratch)
            torch.cuda.synchronize()
            total_ms += ref.run_primary_extend(1, 1 + s)
        clocks = sampler.stop()
        total_rays = n_slots * args.steps
        value = total_rays / (total_ms * 1e-3) / 1e6
        print(json.dumps(dict(base, value=value, ms_per_step=total_ms / args.steps,
                              config={"workload": conf["workload"], "frames_per_step": 1, "rays_per_step": n_slots, "l2": "flushed between steps (256 MiB write)",
                                      "note": "reference = primary_rays, set_wavefront_globals, extend of kernel.cu launched back to back (device time, CUDA events)"},
                              clocks=clocks, gpu_launches=3 * args.steps,
                              cpu_baseline={"value": value, "unit": "Mrays/s", "cores": 0, "kind": "reference", "sample": "whole workload; the reference's own GPU kernels"},
                              e2e={"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})))
        return
    target = args.spp * w * h
    # frames per view: run from a reset until sum(alpha) >= spp * w * h (the reference has no spp parameter)
    frames_per_view = []
    for pos, d in views:
        ref.set_camera(ob.make_camera(position=pos, direction=d))
        ref.mark_sun_changed()
        frames = 0
        while ref.alpha_sum() < target or frames == 0:
            ref.run_frames(1)
            frames += 1
            if frames > 4000:
                break
        frames_per_view.append(frames)

    def step():
        ms_sum, rays = 0.0, 0
        per_view = []
        for (pos, d), frames in zip(views, frames_per_view):
            ref.set_camera(ob.make_camera(position=pos, direction=d))
            ref.mark_sun_changed()
            ms, shadows = ref.run_frames(frames)
            ms_sum += ms
            rays += frames * n_slots + shadows
            per_view.append((frames * n_slots + shadows) / (ms * 1e-3) / 1e6)
        return ms_sum, rays, per_view

    for _ in range(max(args.warmup - 1, 0)):
        step()
    sampler.start()
    total_ms, total_rays, per_view = 0.0, 0, None
    for _ in range(args.steps):
        flush_l2(torch, scratch)
        torch.cuda.synchronize()
        ms, rays, per_view = step()
        total_ms += ms
        total_rays += rays
    clocks = sampler.stop()
    spp = ref.alpha_sum() / (w * h)
    value = total_rays / (total_ms * 1e-3) / 1e6
    cfg = {"workload": conf["workload"], "frames_per_step": sum(frames_per_view), "rays_per_step": total_rays // args.steps, "spp_reached": spp,
           "l2": "flushed between steps (256 MiB write)",
           "note": "reference = CUDA kernels of kernel.cu run unmodified on the GPU; includes its per-frame blit kernel, D->H count copy and cudaDeviceSynchronize "
                   "(kernel.cu:408,428,431); the harness adds one 4-byte cudaMemcpyFromSymbol per frame to count shadow rays (~0.3 % of a frame)"}
    if conf.get("tour"):
        cfg["per_view_mrays"] = [round(v, 1) for v in per_view]
    print(json.dumps(dict(base, value=value, ms_per_step=total_ms / args.steps, config=cfg, clocks=clocks, gpu_launches=sum(frames_per_view) * 6 * args.steps,
                          cpu_baseline={"value": value, "unit": "Mrays/s", "cores": 0, "kind": "reference",
                                        "sample": "whole workload; the reference has no CPU implementation of this path, its own GPU kernels are timed (device time, CUDA events)"},
                          e2e={"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})))


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp", type=int, default=None, help="paths per pixel of a step (default: the config's)")
    ap.add_argument("--slots", type=int, default=0, help="ray_queue_buffer_size per rank (default 2 097 152, variables.h:61)")
    ap.add_argument("--overshoot", action="store_true", help="whole frames until sum(alpha) >= spp w h instead of exactly spp paths per pixel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    conf = dict(CONFIGS[args.config])
    args.warmup = max(args.warmup, 3)
    if args.steps is None:
        args.steps = {"cfg2": 300, "cfg3": 5}.get(args.config, 3)  # (cfg2: a step is one 0.5 ms frame; the clock sampler needs a few 100 ms)
    if args.spp is None:
        args.spp = conf["spp"]
    elif args.spp != conf["spp"]:
        conf["workload"] = conf["workload"].replace("%d spp" % conf["spp"], "%d spp" % args.spp)
        conf["metric"] = conf["metric"].replace("x%dspp" % conf["spp"], "x%dspp" % args.spp)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, conf, rank, world)
        return

    import torch
    import torch.distributed as dist

    import brickmap_b200 as bm
    from brickmap_b200 import renderer as R
    from brickmap_b200.parallel import RequestExchange

    W, H = conf["width"], conf["height"]
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    caves = conf["scene"] == "caves"
    primary_only = bool(conf.get("primary_only"))
    with quiet_stdout():
        if world > 1:
            dist.init_process_group("nccl", device_id=dev)
        # interleaved 8-row strips: a contiguous band per GPU is badly balanced (sky rows finish after one segment)
        rows, _ = bm.strip_rows_for_rank(H, rank, world, STRIP)
        kw = dict(device=local_rank, screen_width=W, screen_height=H, tile_rows=rows, strip_rows=STRIP if world > 1 else 0, strip_count=world, strip_index=rank)
        if args.slots:
            kw["ray_queue_buffer_size"] = args.slots
        if caves:
            kw.update(grid_size=conf["grid"], grid_height=conf["grid"], brick_load_queue_size=conf["queue"])
        cfg = bm.default_config(**kw)
        t_gen = time.perf_counter()
        store = bm.SceneStore(cfg, kind=R.SCENE_CAVES if caves else R.SCENE_TERRAIN, seed=1, resident=not caves)  # replicated per GPU, generated on the device
        torch.cuda.synchronize(dev)
        t_gen = time.perf_counter() - t_gen
        ren = bm.Renderer(cfg, store)
        exchange = RequestExchange(cfg.brick_load_queue_size, dev, world)
        exchange.exchange(ren)  # NCCL communicator warm-up outside the timed region
        torch.cuda.synchronize(dev)
    n_slots = cfg.ray_queue_buffer_size
    blit = torch.zeros(rows, W, 4, dtype=torch.float32, device=dev)
    accum_host = torch.zeros(rows, W, 4, dtype=torch.float32).pin_memory()
    req_count_host = torch.zeros(1, dtype=torch.int32).pin_memory()
    req_pos_host = torch.zeros(cfg.brick_load_queue_size, 3, dtype=torch.int32).pin_memory()
    queue = queue_host = None
    if primary_only:
        queue = torch.zeros(n_slots * 16, dtype=torch.float32, device=dev)
        queue_host = torch.zeros(n_slots * 16, dtype=torch.float32).pin_memory()
    target = args.spp * rows * W
    mode = 0 if args.overshoot else R.FRAME_EXACT_PATHS
    stream = torch.cuda.ExternalStream(ren.stream, device=dev)
    scratch = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)
    views = views_for(conf)
    state = {"k": 0, "frames": 4000, "stream_frames": 0}
    cave_path = {}

    def cave_view(k):  # (memoised: reading index words back synchronises the device; the path is laid out before anything is timed)
        if k not in cave_path:
            cave_path[k] = caves_camera(conf["grid"], k, lambda sc: store.indices(sc, host_view=True))
        return cave_path[k]
    if caves:
        for k in range(2 + 1 + args.warmup + 2 * args.steps + 1):
            cave_view(k)

    def render_view(flags):
        """reset (the caller moved camera or sun) and render to the path target"""
        if caves:
            # streaming: one frame per call with the upload step, then the device-side process_load_queue (Scene.cpp:200-229); the host
            # enqueues a fixed number of frames, the device skips the ones after the target is met
            for _ in range(state["frames"]):
                ren.render(blit, 1, target_paths=target, flags=flags | mode, sync=False)
                exchange.exchange(ren)
                store.process_load_queue(ren.stream)
        else:
            ren.render(blit, state["frames"], target_paths=target, flags=flags | R.FRAME_NO_UPLOAD | mode, sync=False)
            exchange.exchange(ren)  # the only inter-GPU exchange of the path: all-gather + merge of the request blocks

    def step(to_host=False, flags=0):
        if primary_only:
            ren.extend_primaries(queue, 1, sync=False)
            if to_host:
                with torch.cuda.stream(stream):
                    queue_host.copy_(queue, non_blocking=True)
                ren.synchronize()
            return
        if caves:
            todo = [cave_view(state["k"])]  # a camera move resets the accumulation (kernel.cu:387-403)
            state["k"] += 1
        else:
            todo = views
        for pos, d in todo:
            ren.set_camera(bm.make_camera(position=pos, direction=d))
            ren.set_sun(*SUN)  # marks the accumulation for reset, like a sun move in the reference (kernel.cu:389-403)
            render_view(flags)
            if to_host:  # this view's results into HOST buffers through the C ABI (frames=0: copies only)
                ren.render_to_host(blit, 0, accum_host, flags=R.FRAME_NO_UPLOAD, request_count_host=req_count_host, request_positions_host=req_pos_host)

    work = None
    if not primary_only:
        # how many frames does a view take at most? Probe every view (caves: the first steps of the path) until the device-side target stops
        worst = 0
        for probe_step in range(1 if not caves else 2):
            if caves:
                state["k"] += 1
            for pos, d in (views or [cave_view(state["k"] - 1)]):
                ren.set_camera(bm.make_camera(position=pos, direction=d))
                ren.set_sun(*SUN)
                ren.reset_stats()
                before = -1
                while True:
                    if caves:
                        for _ in range(8):
                            ren.render(blit, 1, target_paths=target, flags=mode, sync=False)
                            store.process_load_queue(ren.stream)
                        ren.synchronize()
                    else:
                        ren.render(blit, 8, target_paths=target, flags=R.FRAME_NO_UPLOAD | mode, sync=True)
                    probe = ren.stats()
                    if probe["frames"] == before or probe["frames"] >= 4000:
                        break
                    before = probe["frames"]
                worst = max(worst, int(probe["frames"]))
        state["frames"] = worst + (8 if caves else 2)  # cursor, frame number (and streaming) differ from step to step: leave slack, the device stops at the target
        # one untimed step with the traversal work counters on: algorithmic bytes
        ren.reset_stats()
        step(flags=R.FRAME_COUNT_WORK)
        ren.synchronize()
        work = ren.stats()
    for _ in range(args.warmup):
        step()
    ren.synchronize()

    # ---- timed region: device time per step with CUDA events on the library's stream, L2 flushed between steps
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ren.reset_stats()
    ren.kernel_timing(True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ms = 0.0
    for _ in range(args.steps):
        flush_l2(torch, scratch)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        e1.synchronize()
        ms += e0.elapsed_time(e1)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    kernel_ms, kernel_launches = ren.kernel_time()
    ren.kernel_timing(False)
    stats = ren.stats()
    clocks = sampler.stop() if rank == 0 else None
    rays = stats["extend_rays"] + stats["shadow_rays"]
    requests_left = ren.load_queue()[0] if caves else 0

    per_view = None
    if conf.get("tour"):  # one more pass, view by view (device time per view; same work as inside a step)
        per_view = []
        for pos, d in views:
            ren.set_camera(bm.make_camera(position=pos, direction=d))
            ren.set_sun(*SUN)
            ren.reset_stats()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            render_view(0)
            e1.record(stream)
            e1.synchronize()
            vs = ren.stats()
            per_view.append(round((vs["extend_rays"] + vs["shadow_rays"]) / (e0.elapsed_time(e1) * 1e-3) / 1e6, 1))

    # ---- e2e: the public call with HOST buffers (pinned): camera/sun in, accumulation tile + request buffer out
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ren.reset_stats()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(to_host=True)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    e2e_stats = ren.stats()
    e2e_rays = e2e_stats["extend_rays"] + e2e_stats["shadow_rays"]

    # ---- reduce over ranks: rays summed, time = max
    t = torch.tensor([ms, e2e_s, float(rays), float(e2e_rays), kernel_ms, float(kernel_launches), float(stats["kernel_launches"])], dtype=torch.float64, device=dev)
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    else:
        tmax = tsum = t
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ms_max, e2e_max = float(tmax[0]), float(tmax[1])
    rays_all, e2e_rays_all = float(tsum[2]), float(tsum[3])
    value = rays_all / (ms_max * 1e-3) / 1e6
    e2e_value = e2e_rays_all / e2e_max / 1e6

    # ---- roofline of the dominant kernel (frame_kernel_q) on rank 0: algorithmic bytes / measured launch time
    peak, peak_kind = measured_peaks()
    launches0 = max(kernel_launches, 1)
    rays_per_launch = rays / launches0
    if work is not None:
        wrays = work["extend_rays"] + work["shadow_rays"]
        alg_bytes = 4 * work["cell_steps"] + 64 * work["bricks_entered"] + 12 * work["requests"] + 16 * work["terminations"] + 12 * work["unoccluded"]
        bytes_per_ray = alg_bytes / max(wrays, 1)
        per_ray = {"cell_steps": work["cell_steps"] / max(wrays, 1), "index_words_loaded": work["index_reads"] / max(wrays, 1),
                   "bricks_entered": work["bricks_entered"] / max(wrays, 1), "requests": work["requests"] / max(wrays, 1)}
    else:
        bytes_per_ray, per_ray = None, None  # (primary rays only: the work counters belong to the shading frame kernels)
    achieved = (bytes_per_ray * rays_per_launch) / (kernel_ms / launches0 * 1e-3) / 1e9 if (kernel_ms > 0 and bytes_per_ray) else None
    # DRAM traffic of the kernel: an ncu capture (tools/gpu_round.sh writes profiles/frame_kernel_traffic.json, stamped with the commit
    # and the config it was taken on); a capture of another tree, another config or another GPU count does not describe this run
    traffic, traffic_note, traffic_steady = None, "no ncu capture for this tree / config / GPU count", None
    try:
        from brickmap_b200.build import source_hash
        with open(os.path.join(ROOT, "profiles", "frame_kernel_traffic.json")) as f:
            tj = json.load(f)
        entry = (tj.get("configs") or {}).get(args.config)
        if entry and tj.get("source_hash") == source_hash() and world == 1:
            traffic, traffic_note = entry.get("dram_bytes_per_launch"), "ncu --set full capture of one launch of these kernel sources (hash %s), %s" % (tj["source_hash"], entry.get("report"))
            traffic_steady = (entry.get("steady") or {}).get("dram_bytes_per_launch")
        elif entry and world == 1:
            traffic_note = "the committed capture describes kernel sources %s, this tree is %s" % (tj.get("source_hash"), source_hash())
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "traffic_note": traffic_note, "kernel": "frame_kernel_q", "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs, torch copy)",
                "algorithmic_bytes_per_ray": bytes_per_ray, "rays_per_launch": rays_per_launch, "kernel_ms_per_launch": kernel_ms / launches0,
                "kernel_share_of_step": kernel_ms / ms if ms > 0 else None,
                "traffic_bytes_per_ray": (traffic / rays_per_launch) if traffic else None, "per_ray": per_ray}
    if traffic_steady:
        # the --set full capture flushes the L2 before every replay pass; this is the same kernel's DRAM traffic per launch in a second,
        # single-pass capture of consecutive launches without that flush (the state the timed launches run in)
        roofline.update(traffic_steady=traffic_steady, traffic_steady_bytes_per_ray=traffic_steady / rays_per_launch,
                        traffic_steady_note="ncu single pass, --cache-control none, mean of consecutive launches of the same sources")
    if primary_only:
        roofline["note"] = "primary rays only: the work counters are not collected (they live in the shading kernels); the 64-byte result record per ray is the output"
    cpu = None if args.no_cpu_baseline else cpu_baseline_port(conf)
    nviews = len(views) if views else 1
    d2h = (n_slots * 64) if primary_only else nviews * (rows * W * 16 + 4 + cfg.brick_load_queue_size * 12)
    config = {"workload": conf["workload"], "config": args.config, "spp": args.spp, "frames_per_step": stats["frames"] / args.steps, "rays_per_step": rays_all / args.steps,
              "paths_per_step_rank0": stats["terminations"] / args.steps, "slots_per_frame": n_slots,
              "paths": "whole frames until the target (overshoot)" if args.overshoot else "exactly spp paths per pixel (BM_FRAME_EXACT_PATHS)",
              "partition": "whole image" if world == 1 else "%d ranks, interleaved strips of %d rows (%d rows on rank 0)" % (world, STRIP, rows),
              "l2": "flushed between steps (256 MiB write); scene %s > L2" % ("593 MiB" if not caves else "8 GiB of index words")}
    if per_view is not None:
        config["per_view_mrays"] = per_view  # (rank 0's share of each view)
    if caves:
        config.update(scene_generation_s=t_gen, bricks_total=store.total_bricks, requests_in_queue_after_last_step=requests_left, queue_size=cfg.brick_load_queue_size,
                      requests_per_step_counting_pass=work["requests"], camera_steps_taken=state["k"],
                      camera_path=[[round(v, 2) for v in cave_path[k][0]] for k in sorted(cave_path)][: state["k"]])
    out = {"metric": conf["metric"], "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
           "clocks": clocks, "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": (C.sizeof(bm.Camera) + 8 + 8) * nviews,
                                     "d2h_bytes_per_step": d2h, "seconds_per_step": e2e_max / args.steps},
           "gpu_launches": int(tsum[6]), "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
