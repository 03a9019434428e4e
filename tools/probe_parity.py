"""Development probe (GPU): reference kernels vs CPU oracle vs brickmap_b200, stage by stage, with diagnostics.
Run on the GPU box: python tools/probe_parity.py [variant]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import brickmap_b200 as bm  # noqa: E402
from brickmap_b200 import renderer as R  # noqa: E402
from oracle import binding as ob  # noqa: E402

variant = sys.argv[1] if len(sys.argv) > 1 else "256"
W = H = 512
FRAMES = int(sys.argv[2]) if len(sys.argv) > 2 else 3


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def cmp_bits(name, a, b):
    a, b = bits(a), bits(b)
    bad = np.flatnonzero((a != b).reshape(a.shape[0], -1).any(axis=1)) if a.ndim > 1 else np.flatnonzero(a != b)
    print("  %-34s %8d / %8d records differ" % (name, bad.size, a.shape[0]) + ("" if bad.size == 0 else "  first: %s" % bad[:5]))
    return bad


def cmp_rel(name, a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
    err[(a == b)] = 0
    print("  %-34s max rel err %.3e   (>1e-4: %d of %d)" % (name, err.max() if err.size else 0, int((err > 1e-4).sum()), err.size))
    return err


def cmp_rays(tag, a, b, fields=("origin", "direction", "throughput", "normal", "distance", "bounces", "pixel_index")):
    bad_any = np.zeros(a.shape[0], bool)
    for f in fields:
        x, y = a[f], b[f]
        if x.dtype.kind == "f":
            x, y = bits(x), bits(y)
        d = (x != y).reshape(a.shape[0], -1).any(axis=1)
        if d.any():
            print("  %s.%-12s %d differ, first %s" % (tag, f, int(d.sum()), np.flatnonzero(d)[:5]))
        bad_any |= d
    print("  %-34s %8d / %8d records differ" % (tag, int(bad_any.sum()), a.shape[0]))
    return np.flatnonzero(bad_any)


t0 = time.time()
ref = ob.Reference(variant, W, H)
print("reference variant %s: grid %d x %d, slots %d, lod %d/%d" % (variant, ref.grid_size, ref.grid_height, ref.n_slots, ref.lod2, ref.lod8))
ref.generate()
print("ref.generate %.1fs" % (time.time() - t0))
nsc = ref.supergrid_count()

orc = ob.Oracle()
t0 = time.time()
osc = ob.OracleScene(orc, ref.grid_size, ref.grid_height, ref.lod2, ref.lod8, ref.queue_size).generate_terrain()
print("oracle generate %.1fs" % (time.time() - t0))

cfg = bm.default_config(grid_size=ref.grid_size, grid_height=ref.grid_height, lod_distance_2x2x2=ref.lod2, lod_distance_8x8x8=ref.lod8,
                        brick_load_queue_size=ref.queue_size, ray_queue_buffer_size=ref.n_slots, screen_width=W, screen_height=H)
t0 = time.time()
store = bm.SceneStore(cfg, resident=True)
torch.cuda.synchronize()
print("store generate %.2fs, total bricks %d" % (time.time() - t0, store.total_bricks))

print("== scene: reference host vs oracle vs device store")
bad_o = bad_s = 0
for sc in range(nsc):
    ri, rb = ref.host_supercell(sc)
    oi, obr = osc.host_indices(sc), osc.host_bricks(sc)
    si, sb = store.indices(sc, host_view=True), store.bricks(sc)
    if not (np.array_equal(ri, oi) and np.array_equal(rb, obr)):
        bad_o += 1
    if not (np.array_equal(ri, si) and np.array_equal(rb, sb)):
        bad_s += 1
        if bad_s < 3:
            print("   store mismatch sc", sc, "idx diff", int((ri != si).sum()), "counts", rb.shape, sb.shape)
print("  superchunks differing: oracle %d, store %d (of %d)" % (bad_o, bad_s, nsc))

ref.force_resident()
osc.set_residency(True)
direction = np.array([1, 1, -0.6], np.float32)
direction = (direction / np.sqrt((direction.astype(np.float64) ** 2).sum())).astype(np.float32)
pos = (32.0, 32.0, 250.0) if ref.grid_size == 256 else (512.0, 512.0, 300.0)
cam_o = ob.make_camera(position=pos, direction=direction)
cam_b = bm.make_camera(position=pos, direction=direction)
ref.set_camera(cam_o)
ref.set_sun(0.05, 0.1)
ref.upload_sun()

print("== sun direction / sky")
oren = ob.OracleRenderer(osc, W, H, ref.n_slots, cam_o)
print("  sun dir ref", ref.sun_direction(), "oracle", oren.sun_direction(), "equal bits:", np.array_equal(bits(ref.sun_direction()), bits(oren.sun_direction())))
rng = np.random.default_rng(1)
dirs = rng.normal(size=(4096, 3)).astype(np.float32)
dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
ren = bm.Renderer(cfg, store)
ren.set_camera(cam_b)
ren.set_sun(0.05, 0.1)
for mode, nm in enumerate(("sun", "sky", "sunsky")):
    r = ref.eval_sky(dirs, mode)
    o = ob.sky_eval(orc, dirs, mode, oren.sun_direction())
    m = ren.eval_sky(dirs, mode)
    cmp_rel("%s oracle vs ref" % nm, o, r)
    cmp_rel("%s mine   vs ref" % nm, m, r)

print("== traversal on random rays (ref extend vs oracle vs mine, flat store and reference's own scene)")
n = ref.n_slots
rays = np.zeros(n, ob.RAY_DTYPE)
g = float(ref.grid_size)
gh = float(ref.grid_height)
o = rng.uniform([-0.3 * g, -0.3 * g, -0.3 * gh], [1.3 * g, 1.3 * g, 1.6 * gh], size=(n, 3)).astype(np.float32)
d = rng.normal(size=(n, 3)).astype(np.float32)
d /= np.linalg.norm(d, axis=1, keepdims=True)
d[: n // 16, 0] = 0
d[n // 16: n // 8, 2] = 0
d[:64] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 64)] * rng.choice([-1, 1], size=(64, 1)).astype(np.float32)
rays["origin"], rays["direction"] = o, d
ref.write_rays(rays, 0)
ref.run_stage("extend", frame=1)
rr = ref.read_rays(0)
cam_cell = [int(pos[0] / 8), int(pos[1] / 8), int(pos[2] / 8)]
t0 = time.time()
oh, od, on = osc.trace(o, d, cam_cell, normals=np.zeros((n, 3), np.float32), distances=np.full(n, 1e20, np.float32), threads=0)
print("  oracle trace %.2fs" % (time.time() - t0))
mh, md, mn = ren.trace(o, d, distances=np.full(n, 1e20, np.float32))
print("  hits: ref %d oracle %d mine %d" % (int((rr["distance"] < 1e20).sum()), int(oh.sum()), int(mh.sum())))
cmp_bits("distance oracle vs ref", od, rr["distance"])
cmp_bits("normal   oracle vs ref", on, rr["normal"])
b1 = cmp_bits("distance mine vs ref", md, rr["distance"])
cmp_bits("normal   mine vs ref", mn, rr["normal"])
if b1.size:
    i = b1[0]
    print("   e.g. ray", i, o[i], d[i], "ref", rr["distance"][i], rr["normal"][i], "mine", md[i], mn[i], "oracle", od[i], on[i])
# my kernels on the reference's own GPUScene (pointer-chasing path, true drop-in)
p = ref.scene_pointers()
gs = bm.GpuScene(*p)
ren2 = bm.Renderer(cfg, gs)
ren2.set_camera(cam_b)
mh2, md2, mn2 = ren2.trace(o, d, distances=np.full(n, 1e20, np.float32))
cmp_bits("distance mine(ref scene) vs ref", md2, rr["distance"])

print("== canonical frames: reference (shade serial) vs oracle vs mine")
ref.clear_accum()
ref.write_counters(primary_ray_cnt=0, start_position=0, raynr_primary=0, raynr_extend=0, raynr_shade=0, raynr_connect=0, shadow_ray_cnt=0)
state = bm.State(cfg)
ren.set_counters(0, 0, 0, 1)
for f in range(1, FRAMES + 1):
    print("-- frame", f)
    t0 = time.time()
    ref.run_stage("primary_rays", frame=f)
    r_prim = ref.read_rays(0)
    ref.run_stage("set_wavefront_globals")
    ref.run_stage("extend", frame=f)
    r_ext = ref.read_rays(0)
    ref.run_stage("shade", frame=f, serial=True)
    cnt = ref.counters()
    r_next = ref.read_rays(1, 0, cnt["primary_ray_cnt"])
    r_sh = ref.read_shadows(cnt["shadow_ray_cnt"])
    ref.run_stage("connect", frame=f)
    r_acc = ref.read_accum()
    ref.swap_buffers()
    t_ref = time.time() - t0
    # oracle
    t0 = time.time()
    c_before = oren.state.primary_ray_cnt
    oren.primary_rays()
    o_prim = oren.rays.copy()
    oren.set_wavefront_globals()
    oren.extend()
    o_ext = oren.rays.copy()
    oren.shade()
    o_next = oren.next[: oren.state.primary_ray_cnt].copy()
    o_sh = oren.shadows[: oren.state.shadow_ray_cnt].copy()
    oren.connect()
    o_acc = oren.accum.copy()
    oren.state.frame += 1
    oren.rays, oren.next = oren.next, oren.rays
    t_orc = time.time() - t0
    # mine
    t0 = time.time()
    ren.launch_kernels(state, flags=R.FRAME_NO_RESET)
    mc = ren.counters()
    m_ext = state.rays("work")
    m_next = state.rays("next", mc.primary_ray_cnt)
    m_sh = state.shadows(mc.shadow_ray_cnt)
    m_acc = state.blit_buffer.cpu().numpy()
    state.swap()
    t_mine = time.time() - t0
    print("  counters ref", cnt["primary_ray_cnt"], cnt["shadow_ray_cnt"], cnt["start_position"], "| oracle", oren.state.primary_ray_cnt, oren.state.shadow_ray_cnt,
          oren.state.start_position, "| mine", mc.primary_ray_cnt, mc.shadow_ray_cnt, mc.start_position, " t(ref/orc/mine)=%.1f/%.1f/%.1f" % (t_ref, t_orc, t_mine))
    cmp_rays("primary oracle vs ref", o_prim[c_before:], r_prim[c_before:], ("origin", "direction", "pixel_index"))
    cmp_rays("extend  oracle vs ref", o_ext, r_ext)
    cmp_rays("extend  mine   vs ref", m_ext, r_ext)
    if len(o_next) == len(r_next):
        cmp_rays("next    oracle vs ref", o_next, r_next)
    if len(m_next) == len(r_next):
        cmp_rays("next    mine   vs ref", m_next, r_next)
    if len(o_sh) == len(r_sh):
        cmp_rays("shadow  oracle vs ref", o_sh, r_sh, ("origin", "direction", "pixel_index"))
        cmp_rel("shadow color oracle vs ref", o_sh["color"], r_sh["color"])
    if len(m_sh) == len(r_sh):
        cmp_rays("shadow  mine   vs ref", m_sh, r_sh, ("origin", "direction", "pixel_index"))
        cmp_rel("shadow color mine vs ref", m_sh["color"], r_sh["color"])
    cmp_rel("accum oracle vs ref", o_acc, r_acc)
    cmp_rel("accum mine   vs ref", m_acc, r_acc)

print("== fused render() vs launch_kernels path")
ren3 = bm.Renderer(cfg, store)
ren3.set_camera(cam_b)
ren3.set_sun(0.05, 0.1)
blit = torch.zeros(H, W, 4, dtype=torch.float32, device="cuda")
ren3.render(blit, FRAMES, flags=R.FRAME_NO_RESET)
c3 = ren3.counters()
print("  counters fused", c3.primary_ray_cnt, c3.start_position, c3.frame, "stats", ren3.stats())
cmp_rel("accum fused vs launch_kernels", blit.cpu().numpy(), m_acc)
t0 = time.time()
ren3.render(blit, 50)
dt = time.time() - t0
st = ren3.stats()
print("  50 more frames: %.1f ms/frame, %.1f Mrays/s" % (dt * 20, (50 * ref.n_slots + 0) / dt / 1e6), st)
