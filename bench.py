#!/usr/bin/env python
"""Headline benchmark: Mrays/s of the path-tracing hot path on a 1920x1080 frame at 16 spp.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--spp 16]

Workload (BASELINE.json configs[2]): the reference's procedural terrain (Scene.cpp:44-116) at its stock
constants, 4096 x 4096 x 512 voxels ("4096^3-class"), every brick resident, camera (512,512,300)
looking along +x (camera.h:4-5), sun (0.05, 0.1) (variables.cpp:3), full path trace: <= 4 segments per path
plus one sun shadow ray per hit vertex (kernel.cu:242-346).

One STEP = reset the accumulation buffer, then run frames of 2 097 152 segment slots (variables.h:61) until
16 paths per pixel have finished (sum of alpha >= 16 w h, SURVEY 8d "spp definition").  1 ray = 1 intersect_voxel
call (extend segment or shadow ray).  value = rays of all ranks / max-over-ranks device time.

N > 1 (torchrun): the image is split into N row bands, one per GPU, the brick store is replicated, the only
exchange is the all-gather of the per-step request buffer (NCCL). Total work is fixed -> "scaling": "strong".

--impl reference runs the UNMODIFIED reference kernels (oracle/_ref/libbrickmap_ref_4096.so: kernel.cu, voxel.cuh,
sunsky.cu, Scene.cpp compiled for sm_100a) on the same GPU, same scene/camera/step definition. The reference is a
CUDA program: it has no CPU implementation of this path, so its "CPU arm" is its own GPU kernels.
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

WIDTH, HEIGHT = 1920, 1080
CAM_POS, CAM_DIR = (512.0, 512.0, 300.0), (1.0, 0.0, 0.0)
SUN = (0.05, 0.1)
STRIP = 8  # rows per strip of the multi-GPU image partition
METRIC = "Mrays/s at 1920x1080x16spp (4096^3 scene); HBM GB/s vs roofline"
WORKLOAD = "4096x4096x512 procedural terrain (reference constants), 1920x1080, 16 spp full path trace (<=4 segments + sun shadow rays), bricks resident"


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
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
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


def cpu_baseline_port(frames=3):
    """The CPU oracle (oracle/oracle.cpp, a scalar C++ port of the same algorithm) on the host cores, bounded sample:
    the first `frames` frames after a reset of the same workload."""
    from oracle import binding as ob
    orc = ob.Oracle()
    cores = orc.hardware_threads()
    scene = ob.OracleScene(orc, 4096, 512).generate_terrain().set_residency(True)
    ren = ob.OracleRenderer(scene, WIDTH, HEIGHT, 2 * 1048576, ob.make_camera(position=CAM_POS, direction=CAM_DIR), SUN)
    t0 = time.perf_counter()
    for _ in range(frames):
        ren.frame(threads=0)
    dt = time.perf_counter() - t0
    rays = ren.stats.rays
    # untimed: one more frame with the footprint instrumentation on -> minimum brick-index footprint (32 B x unique
    # index-word and brick sectors the reference algorithm touches in the frame), the comparator of roofline.traffic
    scene.footprint_begin()
    ren.frame(threads=0)
    fp = ob.Stats()
    scene.footprint_report(fp)
    fp_rays = ren.stats.rays - rays
    return {"value": rays / dt / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "port",
            "sample": "first %d frames (%d rays) of the same workload; traversal multi-threaded over all cores, shade loop single-threaded (slot order)" % (frames, rays),
            "seconds": dt,
            "min_footprint": {"bytes_per_frame": 32 * (fp.unique_index_sectors + fp.unique_brick_sectors), "bytes_per_ray": 32 * (fp.unique_index_sectors + fp.unique_brick_sectors) / max(fp_rays, 1),
                              "frame": frames + 1, "note": "32 B x (unique index-word sectors + unique brick sectors) of one frame, from the oracle"}}


def run_reference(args, rank, world):
    """Reference arm: the unmodified reference kernels through launch_kernels (kernel.cu:366) on the GPU."""
    if rank != 0:
        return
    from oracle import binding as ob
    base = {"impl": "reference", "metric": METRIC, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    if not ob.Reference.available("4096"):
        print(json.dumps(dict(base, unavailable="oracle/_ref/libbrickmap_ref_4096.so not built (run make -C oracle ref where /root/reference exists)")))
        return
    import torch
    ref = ob.Reference("4096", WIDTH, HEIGHT, device=0)
    # Scene::generate prints timing lines with std::cout (Scene.cpp:149,192); keep stdout to the one JSON line
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        ref.generate()
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)
    ref.force_resident()
    ref.set_camera(ob.make_camera(position=CAM_POS, direction=CAM_DIR))
    ref.set_sun(*SUN)
    target = args.spp * WIDTH * HEIGHT
    # frames per step: run from a reset until sum(alpha) >= spp * w * h
    frames = 0
    while ref.alpha_sum() < target or frames == 0:
        ref.run_frames(1)
        frames += 1
        if frames > 4000:
            break
    n_slots = ref.n_slots
    sampler = ClockSampler(0)
    for _ in range(max(args.warmup - 1, 0)):
        ref.mark_sun_changed()
        ref.run_frames(frames)
    scratch = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    sampler.start()
    total_ms, total_rays = 0.0, 0
    for _ in range(args.steps):
        flush_l2(torch, scratch)
        torch.cuda.synchronize()
        ref.mark_sun_changed()
        ms, shadows = ref.run_frames(frames)
        total_ms += ms
        total_rays += frames * n_slots + shadows
    clocks = sampler.stop()
    spp = ref.alpha_sum() / (WIDTH * HEIGHT)
    value = total_rays / (total_ms * 1e-3) / 1e6
    out = dict(base, value=value, ms_per_step=total_ms / args.steps,
               config={"workload": WORKLOAD, "frames_per_step": frames, "rays_per_step": total_rays // args.steps, "spp_reached": spp,
                       "l2": "flushed between steps (256 MiB write)", "note": "reference = CUDA kernels of kernel.cu run unmodified on the GPU; includes its per-frame blit kernel, D->H count copy and cudaDeviceSynchronize (kernel.cu:408,428,431)"},
               clocks=clocks, gpu_launches=frames * 6 * args.steps,
               cpu_baseline={"value": value, "unit": "Mrays/s", "cores": 0, "kind": "reference",
                             "sample": "whole workload; the reference has no CPU implementation of this path, its own GPU kernels are timed (device time, CUDA events)"},
               e2e={"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import brickmap_b200 as bm
    from brickmap_b200 import renderer as R
    from brickmap_b200.parallel import RequestExchange

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # NCCL prints its version banner to stdout when the communicator comes up (NCCL_DEBUG=VERSION on some boxes): keep stdout to
    # the one JSON line by pointing fd 1 at stderr until the first collective is through
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # interleaved 8-row strips: a contiguous band per GPU is badly balanced (sky rows finish after one segment)
    rows, _ = bm.strip_rows_for_rank(HEIGHT, rank, world, STRIP)
    cfg = bm.default_config(device=local_rank, screen_width=WIDTH, screen_height=HEIGHT, tile_rows=rows, strip_rows=STRIP if world > 1 else 0,
                            strip_count=world, strip_index=rank)
    store = bm.SceneStore(cfg, resident=True)  # replicated per GPU, generated on the device
    ren = bm.Renderer(cfg, store)
    ren.set_camera(bm.make_camera(position=CAM_POS, direction=CAM_DIR))
    exchange = RequestExchange(cfg.brick_load_queue_size, dev, world)
    exchange.exchange(ren)  # NCCL communicator warm-up outside the timed region
    torch.cuda.synchronize(dev)
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    os.close(saved_stdout)
    blit = torch.zeros(rows, WIDTH, 4, dtype=torch.float32, device=dev)
    accum_host = torch.zeros(rows, WIDTH, 4, dtype=torch.float32).pin_memory()
    req_count_host = torch.zeros(1, dtype=torch.int32).pin_memory()
    req_pos_host = torch.zeros(cfg.brick_load_queue_size, 3, dtype=torch.int32).pin_memory()
    target = args.spp * rows * WIDTH
    max_frames = 4000
    stream = torch.cuda.ExternalStream(ren.stream, device=dev)
    scratch = torch.zeros(64 * 1024 * 1024, dtype=torch.float32, device=dev)

    def step(to_host=False, flags=R.FRAME_NO_UPLOAD):
        ren.set_sun(*SUN)  # marks the accumulation for reset, like a sun move in the reference (kernel.cu:389-403)
        ren.render(blit, step.frames, target_paths=target, flags=flags, sync=False)
        exchange.exchange(ren)  # the only inter-GPU exchange of the path: all-gather + merge of the request blocks
        if to_host:  # results into HOST buffers through the C ABI (frames=0: copies only)
            ren.render_to_host(blit, 0, accum_host, flags=flags, request_count_host=req_count_host, request_positions_host=req_pos_host)

    # how many frames does a step take? Probe in chunks of 8 frames until the device-side target stops the run.
    ren.reset_stats()
    ren.set_sun(*SUN)
    before = -1
    while True:
        ren.render(blit, 8, target_paths=target, flags=R.FRAME_NO_UPLOAD, sync=True)
        probe = ren.stats()
        if probe["frames"] == before or probe["frames"] >= max_frames:
            break
        before = probe["frames"]
    step.frames = int(probe["frames"]) + 2  # cursor and frame number differ from step to step: leave slack, the device stops at the target
    # one untimed step with the traversal work counters on: algorithmic bytes
    ren.reset_stats()
    step(flags=R.FRAME_NO_UPLOAD | R.FRAME_COUNT_WORK)
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

    # ---- roofline of the dominant kernel (frame_kernel) on rank 0: algorithmic bytes / measured launch time
    peak, peak_kind = measured_peaks()
    wrays = work["extend_rays"] + work["shadow_rays"]
    alg_bytes = 4 * work["cell_steps"] + 64 * work["bricks_entered"] + 12 * work["requests"] + 16 * work["terminations"] + 12 * work["unoccluded"]
    bytes_per_ray = alg_bytes / max(wrays, 1)
    launches0 = max(kernel_launches, 1)
    rays_per_launch = rays / launches0
    achieved = (bytes_per_ray * rays_per_launch) / (kernel_ms / launches0 * 1e-3) / 1e9 if kernel_ms > 0 else None
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "frame_kernel_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "kernel": "frame_kernel_q", "peak_source": peak_kind + " (MEASURED_PEAKS.json hbm_gbs, torch copy)",
                "algorithmic_bytes_per_ray": bytes_per_ray, "rays_per_launch": rays_per_launch, "kernel_ms_per_launch": kernel_ms / launches0,
                "kernel_share_of_step": kernel_ms / ms if ms > 0 else None,
                "traffic_bytes_per_ray": (traffic / rays_per_launch) if traffic else None,
                "per_ray": {"cell_steps": work["cell_steps"] / max(wrays, 1), "index_words_loaded": work["index_reads"] / max(wrays, 1),
                            "bricks_entered": work["bricks_entered"] / max(wrays, 1)}}
    cpu = None if args.no_cpu_baseline else cpu_baseline_port()
    out = {"metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps,
           "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": WORKLOAD, "frames_per_step": stats["frames"] / args.steps, "rays_per_step": rays_all / args.steps,
                      "paths_per_step_rank0": stats["terminations"] / args.steps, "partition": "whole image" if world == 1 else "%d ranks, interleaved strips of %d rows (%d rows on rank 0)" % (world, STRIP, rows),
                      "l2": "flushed between steps (256 MiB write); scene 593 MiB > L2"},
           "clocks": clocks, "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": C.sizeof(bm.Camera) + 8 + 8,
                                     "d2h_bytes_per_step": rows * WIDTH * 16 + 4 + cfg.brick_load_queue_size * 12, "seconds_per_step": e2e_max / args.steps},
           "gpu_launches": int(tsum[6]), "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
