"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference kernels
(oracle/_ref/libbrickmap_ref_<variant>.so, built by oracle/Makefile from /root/reference/src) on a GPU.

    gpurun -- python tests/golden/make_golden.py      # writes gpurun_out/golden/*.npz ; copy them to tests/golden/

The reference ships no tests or fixtures of its own; these files pin the CPU oracle (tests/test_oracle_golden.py,
CPU-only) and, through it and directly, the CUDA product (tests/test_gpu_*.py).

Canonical schedule: primary_rays / extend / connect run with the reference's stock parallel launch (their per-slot results
do not depend on scheduling), shade runs <<<1,1>>> so that its atomic slot assignment (kernel.cu:277,298) happens in slot
order; for the streaming fixture extend and connect are serial too, which also makes the request queue order (voxel.cuh:234)
deterministic.
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(OUT, exist_ok=True)


def unit(v):
    v = np.asarray(v, np.float32)
    return (v / np.sqrt((v.astype(np.float64) ** 2).sum())).astype(np.float32)


VIEWS = {
    "256": dict(w=512, h=512, pos=(32.0, 32.0, 250.0), dir=unit([1, 1, -0.6])),
    "256lod": dict(w=512, h=512, pos=(-150.0, -120.0, 330.0), dir=unit([1, 0.9, -0.45])),  # outside the world: AABB entry + both LoD levels
    "4096": dict(w=1920, h=1080, pos=(512.0, 512.0, 300.0), dir=np.array([1, 0, 0], np.float32)),
    # thin-lens camera (kernel.cu:85-103,191-198): lens radius > 0 pins ConcentricSampleDisk, the order in which the two lens draws are
    # taken (unspecified in the source, kernel.cu:194) and the FMA placement of the lens path in the reference build
    "256lens": dict(w=512, h=512, pos=(32.0, 32.0, 250.0), dir=unit([1, 1, -0.6]), lens=0.75, focal=12.5, lib="256"),
}


def scene_digest(ref):
    h = hashlib.sha256()
    counts = []
    for sc in range(ref.supergrid_count()):
        idx, br = ref.host_supercell(sc)
        h.update(idx.tobytes())
        h.update(br.tobytes())
        counts.append(br.shape[0])
    return h.hexdigest(), np.array(counts, np.uint32)


def random_rays(rng, n, g, gh):
    o = rng.uniform([-0.3 * g, -0.3 * g, -0.3 * gh], [1.3 * g, 1.3 * g, 1.6 * gh], size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[: n // 16, 0] = 0
    d[n // 16: n // 8, 2] = 0
    k = min(64, n)
    d[:k] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, k)] * rng.choice([-1, 1], size=(k, 1)).astype(np.float32)
    o[n // 2: n // 2 + n // 8] = rng.uniform([0, 0, 0], [g, g, gh], size=(n // 8, 3)).astype(np.float32)  # inside the world
    return o, d


def sample_idx(rng, n, k):
    return np.sort(rng.choice(n, size=min(k, n), replace=False)) if n else np.zeros(0, np.int64)


def run_variant(variant, frames, stream_frames):
    v = VIEWS[variant]
    W, H = v["w"], v["h"]
    t0 = time.time()
    lib_variant = v.get("lib", variant)
    ref = ob.Reference(lib_variant, W, H)
    ref.generate()
    rng = np.random.default_rng(20261017)
    out = {"grid_size": ref.grid_size, "grid_height": ref.grid_height, "n_slots": ref.n_slots, "lod2": ref.lod2, "lod8": ref.lod8, "queue_size": ref.queue_size,
           "width": W, "height": H, "cam_pos": np.array(v["pos"], np.float32), "cam_dir": v["dir"], "sun": np.array([0.05, 0.1], np.float32)}
    digest, counts = scene_digest(ref)
    out["scene_sha256"] = digest
    out["brick_counts"] = counts
    idx0, br0 = ref.host_supercell(0)
    out["sc0_indices"], out["sc0_bricks"] = idx0, br0[:64]
    cam = ob.make_camera(position=v["pos"], direction=v["dir"], focal=v.get("focal", 1.0), lens=v.get("lens", 0.0))
    out["focal"], out["lens"] = np.float32(v.get("focal", 1.0)), np.float32(v.get("lens", 0.0))
    ref.set_camera(cam)
    ref.set_sun(0.05, 0.1)
    ref.upload_sun()
    out["sun_dir"] = ref.sun_direction()

    if stream_frames:
        # --- streaming from an empty device scene (Scene.cpp:157-164), every stage serial -> deterministic queue order
        ref.clear_accum()
        ref.write_counters(primary_ray_cnt=0, start_position=0, raynr_primary=0, raynr_extend=0, raynr_shade=0, raynr_connect=0, shadow_ray_cnt=0)
        for f in range(1, stream_frames + 1):
            cnt, _ = ref.load_queue()
            ref.run_stage("upload", upload_count=min(cnt, ref.queue_size))  # kernel.cu:408-414
            ref.run_stage("primary_rays", frame=f)
            ref.run_stage("set_wavefront_globals")
            ref.run_stage("extend", frame=f, serial=True)
            ref.run_stage("shade", frame=f, serial=True)
            ref.run_stage("connect", frame=f, serial=True)
            cnt, pos = ref.load_queue()
            ref.process_load_queue()  # Scene.cpp:200
            ref.swap_buffers()
            c = ref.counters()
            out["stream%d_count" % f] = np.uint32(cnt)
            out["stream%d_positions" % f] = pos
            out["stream%d_counters" % f] = np.array([c["primary_ray_cnt"], c["shadow_ray_cnt"], c["start_position"]], np.uint32)
            idx = ref.read_indices()
            out["stream%d_index_sha256" % f] = hashlib.sha256(idx.tobytes()).hexdigest()
            print("  stream frame", f, "requests", cnt, "t=%.1fs" % (time.time() - t0))
        acc = ref.read_accum()
        pix = sample_idx(rng, W * H, 4096)
        out["stream_accum_pix"], out["stream_accum_val"] = pix, acc.reshape(-1, 4)[pix]
        # fresh, empty device scene again for the resident part
        ref = None
        ref = ob.Reference(lib_variant, W, H)
        ref.generate()
        ref.set_camera(cam)
        ref.set_sun(0.05, 0.1)
        ref.upload_sun()

    ref.force_resident()
    # --- sky
    dirs = rng.normal(size=(1024, 3)).astype(np.float32)
    dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    out["sky_dirs"] = dirs
    for mode, nm in enumerate(("sun", "sky", "sunsky")):
        out["sky_" + nm] = ref.eval_sky(dirs, mode)
    # --- traversal of random rays through the reference's extend kernel
    n = 16384
    o, d = random_rays(rng, n, float(ref.grid_size), float(ref.grid_height))
    rays = np.zeros(ref.n_slots, ob.RAY_DTYPE)
    rays["origin"][:n], rays["direction"][:n] = o, d
    rays["origin"][n:] = (-1e6, -1e6, -1e6)
    rays["direction"][n:] = (0, 0, -1)
    ref.write_rays(rays, 0)
    # the reference's kernels pull slots from device counters that only set_wavefront_globals zeroes (kernel.cu:132-138)
    ref.write_counters(primary_ray_cnt=0, start_position=0, raynr_primary=0, raynr_extend=0, raynr_shade=0, raynr_connect=0, shadow_ray_cnt=0)
    ref.run_stage("extend", frame=1)
    rr = ref.read_rays(0, 0, n)
    out["trace_origins"], out["trace_directions"], out["trace_distance"], out["trace_normal"] = o, d, rr["distance"], rr["normal"]
    # --- canonical frames
    ref.clear_accum()
    ref.write_counters(primary_ray_cnt=0, start_position=0, raynr_primary=0, raynr_extend=0, raynr_shade=0, raynr_connect=0, shadow_ray_cnt=0)
    pix = sample_idx(rng, W * H, 4096)
    out["accum_pix"] = pix
    for f in range(1, frames + 1):
        ref.run_stage("primary_rays", frame=f)
        ref.run_stage("set_wavefront_globals")
        ref.run_stage("extend", frame=f)
        ext = ref.read_rays(0)
        ref.run_stage("shade", frame=f, serial=True)
        c = ref.counters()
        nxt = ref.read_rays(1, 0, c["primary_ray_cnt"])
        sh = ref.read_shadows(c["shadow_ray_cnt"])
        ref.run_stage("connect", frame=f)
        acc = ref.read_accum()
        ref.swap_buffers()
        ei, ni, si = sample_idx(rng, len(ext), 4096), sample_idx(rng, len(nxt), 2048), sample_idx(rng, len(sh), 2048)
        p = "f%d_" % f
        out[p + "counters"] = np.array([c["primary_ray_cnt"], c["shadow_ray_cnt"], c["start_position"]], np.uint32)
        out[p + "ext_idx"], out[p + "ext"] = ei, ext[ei]
        out[p + "next_idx"], out[p + "next"] = ni, nxt[ni]
        out[p + "shadow_idx"], out[p + "shadow"] = si, sh[si]
        out[p + "ext_hits"] = np.uint32((ext["distance"] < 1e20).sum())
        out[p + "accum_val"] = acc.reshape(-1, 4)[pix]
        th, tw = H // 8, W // 8
        out[p + "accum_tiles"] = acc[: th * 8, : tw * 8].reshape(8, th, 8, tw, 4).astype(np.float64).mean(axis=(1, 3)).astype(np.float32)
        out[p + "alpha_sum"] = np.float64(acc[..., 3].astype(np.float64).sum())
        print("  frame", f, c, "t=%.1fs" % (time.time() - t0))
    path = os.path.join(OUT, "golden_%s.npz" % variant)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    which = sys.argv[1:] or ["256", "256lod", "4096", "256lens"]
    for variant in which:
        print("variant", variant)
        # one variant per process would be cleaner (the harness keeps global state); separate .so files keep them apart here
        run_variant(variant, frames=3 if variant != "4096" else 2, stream_frames=3 if variant == "256" else 0)
