"""GPU parity tests proper: the CUDA product (through the C ABI, include/brickmap_b200.h) against
  (1) the CPU oracle on the same seeded inputs, bit for bit on every geometric quantity, 1e-4 relative on radiance
      (BASELINE.json north_star tolerance),
  (2) the committed golden vectors produced by the unmodified reference kernels (tests/golden/),
  (3) the unmodified reference kernels run live in the same process when oracle/_ref/*.so is present, including the
      product running directly on the REFERENCE host's own GPUScene (true drop-in: per-superchunk cudaMalloc'd arrays).
"""
import hashlib

import numpy as np
import pytest
import torch

import brickmap_b200 as bm
from brickmap_b200 import renderer as R
from helpers import assert_close_rel, assert_records_equal, bits, max_rel_err, tile_means
from oracle import binding as ob

pytestmark = pytest.mark.gpu
RADIANCE_TOL = 1e-4


def cfg_from_golden(g, **kw):
    return bm.default_config(grid_size=int(g["grid_size"]), grid_height=int(g["grid_height"]), lod_distance_2x2x2=int(g["lod2"]),
                             lod_distance_8x8x8=int(g["lod8"]), brick_load_queue_size=int(g["queue_size"]), ray_queue_buffer_size=int(g["n_slots"]),
                             screen_width=int(g["width"]), screen_height=int(g["height"]), **kw)


def renderer_for(g, store, cfg=None):
    ren = bm.Renderer(cfg or store.cfg, store)
    ren.set_camera(bm.make_camera(position=g["cam_pos"], direction=g["cam_dir"], focal=float(g["focal"]) if "focal" in g else 1.0,
                                  lens=float(g["lens"]) if "lens" in g else 0.0))
    ren.set_sun(float(g["sun"][0]), float(g["sun"][1]))
    return ren


def oracle_renderer(oracle, g, resident=True):
    s = ob.OracleScene(oracle, int(g["grid_size"]), int(g["grid_height"]), int(g["lod2"]), int(g["lod8"]), int(g["queue_size"])).generate_terrain()
    s.set_residency(resident)
    cam = ob.make_camera(position=g["cam_pos"], direction=g["cam_dir"])
    return s, ob.OracleRenderer(s, int(g["width"]), int(g["height"]), int(g["n_slots"]), cam, tuple(float(v) for v in g["sun"]))


@pytest.fixture(scope="module")
def stores(lib, golden):
    cache = {}

    def get(name, resident=True, nonflat=False):
        key = (name, resident, nonflat)
        if key not in cache:
            cache[key] = bm.SceneStore(cfg_from_golden(golden(name)), resident=resident, nonflat=nonflat)
        return cache[key]
    yield get
    for s in cache.values():
        s.close()


# ---- scene generation on the device ------------------------------------------------------------------------------
@pytest.mark.parametrize("variant", ["256", "4096"])
def test_device_scene_generation_matches_reference_golden(golden, stores, variant):
    g, s = golden(variant), stores(variant)
    h = hashlib.sha256()
    counts = []
    for sc in range(s.superchunks):
        h.update(s.indices(sc, host_view=True).tobytes())
        h.update(s.bricks(sc).tobytes())
        counts.append(s.brick_count(sc))
    assert np.array_equal(np.array(counts, np.uint32), g["brick_counts"])
    assert h.hexdigest() == str(g["scene_sha256"]), "device-generated world differs from the reference host's Scene::generate"
    assert np.array_equal(s.indices(0), g["sc0_indices"])  # resident view == host view


def test_unloaded_device_view(golden, stores):
    s = stores("256", resident=False)
    host, dev = s.indices(3, host_view=True), s.indices(3)
    want = np.where(host & 0x80000000, 0x40000000 | (host & 0xFF000), 0).astype(np.uint32)  # Scene.cpp:158-164
    assert np.array_equal(dev, want)


# ---- sky ----------------------------------------------------------------------------------------------------------
def test_sky_matches_reference_golden(golden, stores):
    g = golden("256")
    ren = renderer_for(g, stores("256"))
    for mode, nm in enumerate(("sun", "sky", "sunsky")):
        assert_close_rel(ren.eval_sky(g["sky_dirs"], mode), g["sky_" + nm], RADIANCE_TOL, nm)


def test_sky_other_sun_positions_match_oracle(oracle, golden, stores):
    g = golden("256")
    ren = renderer_for(g, stores("256"))
    for sun in ((0.3, 0.2), (0.7, 0.45), (0.05, 0.49)):
        ren.set_sun(*sun)
        sd = np.zeros(3, np.float32)
        oracle.lib.orc_sun_direction(sun[0], sun[1], sd.ctypes.data)
        for mode in range(3):
            assert_close_rel(ren.eval_sky(g["sky_dirs"], mode), ob.sky_eval(oracle, g["sky_dirs"], mode, sd), RADIANCE_TOL, "sun %s mode %d" % (sun, mode))


# ---- traversal ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant,nonflat", [("256", False), ("256", True), ("256lod", False), ("256lod", True), ("4096", False)])
def test_traversal_matches_reference_golden(golden, stores, variant, nonflat):
    g = golden(variant)
    ren = renderer_for(g, stores(variant, nonflat=nonflat))
    n = g["trace_origins"].shape[0]
    hit, dist, nrm = ren.trace(g["trace_origins"], g["trace_directions"], distances=np.full(n, 1e20, np.float32))
    assert np.array_equal(bits(dist), bits(g["trace_distance"]))
    assert np.array_equal(bits(nrm), bits(g["trace_normal"]))
    assert int(hit.sum()) == int((g["trace_distance"] < 1e20).sum()) > 0


def test_traversal_edge_cases_match_oracle(oracle, golden, stores):
    """Axis-aligned and zero-component directions, origins on cell/voxel boundaries, on the world faces, outside the world,
    empty input."""
    g = golden("256lod")
    ren = renderer_for(g, stores("256lod"))
    s, _ = oracle_renderer(oracle, g)
    rng = np.random.default_rng(7)
    n = 20000
    o = rng.uniform(-40, 300, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    o[:4000] = np.round(o[:4000] / 8) * 8          # on cell boundaries
    o[4000:8000] = np.round(o[4000:8000])          # on voxel boundaries
    o[8000:9000, 2] = 256.0                        # on the top face
    o[9000:10000, 0] = 0.0
    d[10000:12000, rng.integers(0, 3)] = 0.0
    d[12000:13000] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, 1000)] * rng.choice([-1.0, 1.0], size=(1000, 1)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    cam_cell = [int(v / 8.0) for v in g["cam_pos"]]
    nrm0 = rng.choice([-1.0, 0.0, 1.0], size=(n, 3)).astype(np.float32)
    oh, od, on = s.trace(o, d, cam_cell, normals=nrm0, distances=np.full(n, 1e20, np.float32), threads=0)
    mh, md, mn = ren.trace(o, d, normals=nrm0, distances=np.full(n, 1e20, np.float32))
    assert np.array_equal(mh, oh) and np.array_equal(bits(md), bits(od)) and np.array_equal(bits(mn), bits(on))
    eh, ed, en = ren.trace(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert eh.shape == (0,) and ed.shape == (0,)


# ---- whole frames ----------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("variant,nonflat", [("256", False), ("256lod", True), ("4096", False), ("256lens", False)])
def test_launch_frame_matches_reference_golden(golden, stores, variant, nonflat):
    """bm_launch_frame == launch_kernels: every buffer the reference leaves behind, frame after frame. 256lens: the reference
    build's thin-lens camera (lens radius 0.75; kernel.cu:85-103,191-198)."""
    g = golden(variant)
    store = stores("256" if variant == "256lens" else variant, nonflat=nonflat)
    ren = renderer_for(g, store)
    state = bm.State(store.cfg)
    f = 1
    while "f%d_counters" % f in g:
        p = "f%d_" % f
        ren.launch_kernels(state)
        c = ren.counters()
        assert [c.primary_ray_cnt, c.shadow_ray_cnt, c.start_position, c.frame] == [int(v) for v in g[p + "counters"]] + [f + 1]
        ext = state.rays("work")
        assert int((ext["distance"] < 1e20).sum()) == int(g[p + "ext_hits"])
        assert_records_equal(ext[g[p + "ext_idx"]], g[p + "ext"], what="frame %d post-extend" % f)
        assert_records_equal(state.rays("next", c.primary_ray_cnt)[g[p + "next_idx"]], g[p + "next"], what="frame %d survivors" % f)
        sh = state.shadows(c.shadow_ray_cnt)[g[p + "shadow_idx"]]
        assert_records_equal(sh, g[p + "shadow"], fields=("origin", "direction", "pixel_index"), what="frame %d shadow rays" % f)
        assert_close_rel(sh["color"], g[p + "shadow"]["color"], RADIANCE_TOL, "frame %d shadow colour" % f)
        acc = state.blit_buffer.cpu().numpy()
        assert_close_rel(acc.reshape(-1, 4)[g["accum_pix"]], g[p + "accum_val"], RADIANCE_TOL, "frame %d accumulation samples" % f)
        assert_close_rel(tile_means(acc), g[p + "accum_tiles"], RADIANCE_TOL, "frame %d tile means" % f)
        assert abs(float(acc[..., 3].astype(np.float64).sum()) - float(g[p + "alpha_sum"])) < 0.5
        state.swap()  # main.cpp:146
        f += 1
    assert f > 2


def test_fused_render_matches_oracle_over_many_frames(oracle, golden, stores):
    """bm_render (private ray state, no per-stage records) == the oracle's canonical wavefront loop, 8 frames."""
    g = golden("256")
    store = stores("256")
    ren = renderer_for(g, store)
    _, oren = oracle_renderer(oracle, g)
    blit = torch.zeros(int(g["height"]), int(g["width"]), 4, dtype=torch.float32, device="cuda")
    frames = 8
    ren.render(blit, 3)
    ren.render(blit, frames - 3)  # state carries over between calls
    for _ in range(frames):
        oren.frame()
    c = ren.counters()
    assert [c.primary_ray_cnt, c.start_position, c.frame] == [oren.state.primary_ray_cnt, oren.state.start_position, oren.state.frame]
    st = ren.stats()
    assert st["frames"] == frames and st["extend_rays"] == frames * int(g["n_slots"])
    assert st["extend_rays"] + st["shadow_rays"] == oren.stats.rays
    assert st["terminations"] == oren.stats.terminations and st["unoccluded"] == oren.stats.unoccluded
    acc = blit.cpu().numpy()
    assert np.array_equal(acc[..., 3], oren.accum[..., 3]), "alpha (finished paths per pixel) must match exactly"
    assert_close_rel(acc, oren.accum, RADIANCE_TOL, "accumulated radiance after %d frames" % frames)
    # a frame through the reference-compatible entry point continues from the private state of the fused path
    state = bm.State(store.cfg)
    state.blit_buffer.copy_(blit)
    ren.launch_kernels(state, flags=R.FRAME_NO_RESET)
    oren.frame()
    c = ren.counters()
    assert_records_equal(state.rays("next", c.primary_ray_cnt), oren.rays[: c.primary_ray_cnt], what="survivors after mixing both entry points")


@pytest.mark.parametrize("variant", ["256", "256lod", "4096"])
def test_throughput_kernel_matches_simple_kernel_on_adversarial_rays(golden, stores, variant):
    """The throughput path (bm_import_rays -> bm_render -> bm_export_rays, private sparse survivor storage) against the
    reference-layout path (bm_launch_frame on caller queues) on the same injected rays: exact ties between axes (diagonal and
    axis-aligned directions from lattice origins), zero components, origins on cell / block boundaries, outside the world,
    all bounce counts."""
    g = golden(variant)
    store = stores(variant)
    cfg = store.cfg
    n = cfg.ray_queue_buffer_size
    rng = np.random.default_rng(11)
    rays = np.zeros(n, bm.RAY_DTYPE)
    gs, gh = float(cfg.grid_size), float(cfg.grid_height)
    o = rng.uniform([-0.2 * gs, -0.2 * gs, -0.2 * gh], [1.2 * gs, 1.2 * gs, 1.5 * gh], size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    q = n // 8
    o[:q] = np.round(o[:q] / 32) * 32                      # block corners
    o[q:2 * q] = np.round(o[q:2 * q] / 8) * 8 + 4          # cell centres
    diag = np.array([[1, 1, 0], [1, 0, 1], [0, 1, 1], [1, 1, 1], [1, -1, 0], [-1, 1, 1], [2, 1, 0], [1, 2, 2], [1, 0, 0], [0, 0, -1]], np.float32)
    d[: 2 * q] = diag[rng.integers(0, len(diag), 2 * q)] * rng.choice([-1.0, 1.0], size=(2 * q, 1)).astype(np.float32)
    d[2 * q:3 * q, rng.integers(0, 3)] = 0.0
    o[3 * q:4 * q] = rng.uniform([0, 0, gh * 0.6], [gs, gs, gh], size=(q, 3)).astype(np.float32)  # inside, above the terrain
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["origin"], rays["direction"] = o, d
    rays["throughput"] = rng.uniform(0.2, 1.0, size=(n, 3)).astype(np.float32)
    rays["normal"] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, n)] * rng.choice([-1.0, 0.0, 1.0], size=(n, 1)).astype(np.float32)
    rays["bounces"] = rng.integers(0, 4, n)
    rays["pixel_index"] = rng.integers(0, cfg.screen_width * cfg.screen_height, n)
    h, w = cfg.screen_height, cfg.screen_width

    simple = renderer_for(g, store)
    state = bm.State(cfg)
    state.write_rays(rays, "work")
    simple.set_counters(primary_ray_cnt=n, start_position=123, frame=5)
    simple.launch_kernels(state, flags=R.FRAME_NO_RESET)
    cs = simple.counters()
    want = state.rays("next", cs.primary_ray_cnt)

    fast = renderer_for(g, store)
    fast.set_counters(primary_ray_cnt=n, start_position=123, frame=5)
    fast.import_rays(rays)
    blit = torch.zeros(h, w, 4, dtype=torch.float32, device="cuda")
    fast.render(blit, 1, flags=R.FRAME_NO_RESET)
    cf = fast.counters()
    assert [cf.primary_ray_cnt, cf.start_position, cf.frame] == [cs.primary_ray_cnt, cs.start_position, cs.frame]
    assert 0 < cf.primary_ray_cnt < n
    assert_records_equal(fast.export_rays(), want, what="survivors, bm_render vs bm_launch_frame")
    a, b = blit.cpu().numpy(), state.blit_buffer.cpu().numpy()
    assert np.array_equal(a[..., 3], b[..., 3])
    assert_close_rel(a, b, 1e-5, "accumulation, bm_render vs bm_launch_frame")
    sf, ss = fast.stats(), simple.stats()
    assert [sf[k] for k in ("shadow_rays", "terminations", "unoccluded")] == [ss[k] for k in ("shadow_rays", "terminations", "unoccluded")]


def test_throughput_kernel_matches_simple_kernel_on_the_stock_world(golden, stores):
    """Three frames of the benchmark view on the 4096 x 4096 x 512 world: the throughput kernel (compiled-in bitmap geometry,
    far-block runs of five DDA steps, suspend / resume through the per-warp queues) against the per-slot kernel behind
    bm_launch_frame, which tests every cell. Survivor records bit for bit, alpha exactly, radiance to 1e-5."""
    g = golden("4096")
    store = stores("4096")
    cfg = store.cfg
    h, w = cfg.screen_height, cfg.screen_width
    simple = renderer_for(g, store)
    state = bm.State(cfg)
    frames = 3
    for f in range(frames):
        simple.launch_kernels(state)
        if f + 1 < frames:
            state.swap()
    cs = simple.counters()
    want = state.rays("next", cs.primary_ray_cnt)

    fast = renderer_for(g, store)
    blit = torch.zeros(h, w, 4, dtype=torch.float32, device="cuda")
    fast.render(blit, frames)
    cf = fast.counters()
    assert [cf.primary_ray_cnt, cf.start_position, cf.frame] == [cs.primary_ray_cnt, cs.start_position, cs.frame]
    assert_records_equal(fast.export_rays(), want, what="survivors after %d frames, bm_render vs bm_launch_frame" % frames)
    a, b = blit.cpu().numpy(), state.blit_buffer.cpu().numpy()
    assert np.array_equal(a[..., 3], b[..., 3])
    assert_close_rel(a, b, 1e-5, "accumulation, bm_render vs bm_launch_frame")
    sf, ss = fast.stats(), simple.stats()
    assert [sf[k] for k in ("shadow_rays", "terminations", "unoccluded")] == [ss[k] for k in ("shadow_rays", "terminations", "unoccluded")]


@pytest.mark.parametrize("variant", ["256lod", "4096"])
def test_throughput_kernel_schedules_agree(golden, stores, variant, monkeypatch):
    """How long a batch runs before its unfinished rays are suspended (fixed quantum; given up early when the warp thins out;
    ), in which order slot runs are handed out and whether the stock world's bitmap geometry is
    compiled in are schedules / code paths of the same arithmetic:
    survivors bit for bit, alpha exactly, radiance to 1e-5 (the order of the float atomics differs)."""
    g = golden(variant)
    store = stores(variant)
    cfg = store.cfg
    h, w = cfg.screen_height, cfg.screen_width
    results = []
    for quantum, share, no_stock, descending, run_len, resume_at, brick_lanes, no_sky in (
            (64, 0, 0, 0, 1, 32, 0, 1), (16, 0, 0, 0, 4, 32, 0, 0), (256, 16, 0, 0, 1, 32, 0, 0), (512, 28, 0, 1, 2, 24, 8, 1), (256, 16, 1, 0, 1, 32, 0, 0),
            (128, 8, 0, 1, 16, 7, 32, 0), (256, 18, 0, 0, 1, 24, 12, 1), (64, 30, 0, 0, 1, 32, 32, 0), (256, 14, 0, 0, 1, 28, 3, 0), (32, 8, 1, 1, 8, 9, 5, 0)):
        monkeypatch.setenv("BRICKMAP_B200_BRICK_LANES", str(brick_lanes))
        monkeypatch.setenv("BRICKMAP_B200_NO_SKY", str(no_sky))  # 1: without the early "nothing left to meet" exit of non-descending rays
        monkeypatch.setenv("BRICKMAP_B200_RUN_LEN", str(run_len))
        monkeypatch.setenv("BRICKMAP_B200_RESUME_AT", str(resume_at))
        monkeypatch.setenv("BRICKMAP_B200_QUANTUM", str(quantum))
        monkeypatch.setenv("BRICKMAP_B200_MIN_SHARE", str(share))
        monkeypatch.setenv("BRICKMAP_B200_NO_STOCK", str(no_stock))
        monkeypatch.setenv("BRICKMAP_B200_DESCENDING", str(descending))
        ren = renderer_for(g, store)  # the switches are read when the scene is bound
        blit = torch.zeros(h, w, 4, dtype=torch.float32, device="cuda")
        ren.render(blit, 3)
        c = ren.counters()
        st = ren.stats()
        results.append((c.primary_ray_cnt, c.start_position, c.frame, st["shadow_rays"], st["terminations"], st["unoccluded"], ren.export_rays(), blit.cpu().numpy()))
    ref = results[0]
    for r in results[1:]:
        assert r[:6] == ref[:6]
        assert_records_equal(r[6], ref[6], what="survivors across schedules")
        assert np.array_equal(r[7][..., 3], ref[7][..., 3])
        assert_close_rel(r[7], ref[7], 1e-5, "accumulation across schedules")


def test_ragged_sizes_match_oracle(oracle, golden):
    """Slot count not a multiple of the warp size or of a run, image not a multiple of anything, fewer slots than pixels and
    more slots than pixels (the cursor wraps inside a frame, kernel.cu:170-171)."""
    g = golden("256")
    for n_slots, w, h in ((100003, 333, 211), (50021, 401, 97)):
        cfg = cfg_from_golden(g, )
        cfg.ray_queue_buffer_size, cfg.screen_width, cfg.screen_height = n_slots, w, h
        store = bm.SceneStore(cfg, resident=True)
        ren = bm.Renderer(cfg, store)
        ren.set_camera(bm.make_camera(position=g["cam_pos"], direction=g["cam_dir"]))
        s = ob.OracleScene(oracle, 256, 256).generate_terrain().set_residency(True)
        oren = ob.OracleRenderer(s, w, h, n_slots, ob.make_camera(position=g["cam_pos"], direction=g["cam_dir"]))
        state = bm.State(cfg)
        blit = torch.zeros(h, w, 4, dtype=torch.float32, device="cuda")
        ren2 = bm.Renderer(cfg, store)
        ren2.set_camera(bm.make_camera(position=g["cam_pos"], direction=g["cam_dir"]))
        for f in range(4):
            ren.launch_kernels(state)
            oren.frame()
            c = ren.counters()
            assert [c.primary_ray_cnt, c.shadow_ray_cnt, c.start_position] == [oren.state.primary_ray_cnt, oren.state.shadow_ray_cnt, oren.state.start_position]
            assert_records_equal(state.rays("next", c.primary_ray_cnt), oren.rays[: c.primary_ray_cnt], what="survivors, n_slots=%d frame %d" % (n_slots, f + 1))
            state.swap()
        ren2.render(blit, 4)
        assert_close_rel(state.blit_buffer.cpu().numpy(), oren.accum, RADIANCE_TOL, "accumulation n_slots=%d" % n_slots)
        assert_close_rel(blit.cpu().numpy(), oren.accum, RADIANCE_TOL, "fused accumulation n_slots=%d" % n_slots)
        store.close()


def test_host_buffer_entry_point(golden, stores):
    """bm_render_to_host / bm_read_requests: results land in HOST buffers (pinned), as the e2e leg of bench.py uses them."""
    g = golden("256")
    cfg = cfg_from_golden(g)
    store = bm.SceneStore(cfg, resident=False)
    ren = renderer_for(g, store, cfg)
    h, w, q = int(g["height"]), int(g["width"]), cfg.brick_load_queue_size
    blit = torch.zeros(h, w, 4, dtype=torch.float32, device="cuda")
    accum_host = torch.zeros(h, w, 4, dtype=torch.float32).pin_memory()
    cnt_host = torch.zeros(1, dtype=torch.int32).pin_memory()
    pos_host = torch.zeros(q, 3, dtype=torch.int32).pin_memory()
    ren.render_to_host(blit, 1, accum_host, request_count_host=cnt_host, request_positions_host=pos_host)
    assert torch.equal(accum_host, blit.cpu())
    assert int(cnt_host[0]) == int(g["stream1_count"])
    assert sorted(map(tuple, pos_host[: int(cnt_host[0])].numpy())) == sorted(map(tuple, g["stream1_positions"]))
    cnt, pos = ren.load_queue()
    assert cnt == int(cnt_host[0]) and np.array_equal(pos, pos_host[:cnt].numpy())


def test_work_counters_match_oracle(oracle, golden, stores):
    """S (cell steps), K (bricks entered), P, V of SURVEY 8d, which feed the roofline's algorithmic bytes."""
    g = golden("256")
    ren = renderer_for(g, stores("256"))
    _, oren = oracle_renderer(oracle, g)
    blit = torch.zeros(int(g["height"]), int(g["width"]), 4, dtype=torch.float32, device="cuda")
    ren.render(blit, 2, flags=R.FRAME_COUNT_WORK)
    oren.frame()
    oren.frame()
    st, os_ = ren.stats(), oren.stats.as_dict()
    assert st["cell_steps"] == os_["index_reads"] and st["bricks_entered"] == os_["bricks"]
    assert st["index_reads"] <= st["cell_steps"]
    assert st["terminations"] == os_["terminations"] and st["unoccluded"] == os_["unoccluded"]


def test_render_target_and_reset(golden, stores):
    g = golden("256")
    ren = renderer_for(g, stores("256"))
    pixels = int(g["height"]) * int(g["width"])
    blit = torch.zeros(int(g["height"]), int(g["width"]), 4, dtype=torch.float32, device="cuda")
    ren.render(blit, 200, target_paths=2 * pixels)  # stops on the device once 2 spp are reached
    st = ren.stats()
    assert 2 * pixels <= st["terminations"] < 2 * pixels + 2 * int(g["n_slots"]) and st["frames"] < 50
    assert float(blit[..., 3].sum()) == st["terminations"]
    ren.set_sun(0.3, 0.2)  # sun move -> accumulation reset (kernel.cu:389-403); cursor and frame number are kept
    before = ren.counters()
    ren.render(blit, 1)
    after = ren.counters()
    assert after.frame == before.frame + 1
    assert float(blit[..., 3].sum()) == ren.stats()["terminations"] - st["terminations"]


@pytest.mark.parametrize("slack", [1, 2])
def test_render_target_slack_then_camera_move_matches_oracle(oracle, golden, stores, slack):
    """bm_render asked for more frames than the path target needs: the frames the device skips must not move any state (which
    survivor set is current, cursor, frame number). Odd and even slack, then a camera move (accumulation reset,
    kernel.cu:387-403) and several more frames, everything against the oracle."""
    g = golden("256")
    ren = renderer_for(g, stores("256"))
    _, oren = oracle_renderer(oracle, g)
    h, w = int(g["height"]), int(g["width"])
    target = 2 * h * w
    executed = 0
    while oren.stats.terminations < target:
        oren.frame()
        executed += 1
    blit = torch.zeros(h, w, 4, dtype=torch.float32, device="cuda")
    ren.render(blit, executed + slack, target_paths=target)
    st = ren.stats()
    assert st["frames"] == executed and st["terminations"] == oren.stats.terminations
    c = ren.counters()
    assert [c.primary_ray_cnt, c.start_position, c.frame] == [oren.state.primary_ray_cnt, oren.state.start_position, oren.state.frame]
    assert_records_equal(ren.export_rays(), oren.rays[: c.primary_ray_cnt], what="survivors after a target stop with slack %d" % slack)
    # continue WITHOUT a reset: the survivors must be picked up from the right set
    ren.render(blit, 1)
    oren.frame()
    c = ren.counters()
    assert_records_equal(ren.export_rays(), oren.rays[: c.primary_ray_cnt], what="survivors one frame after the stop")
    # target stop again, then move the camera: reset, the next frames see none of the old survivors
    ren.render(blit, 3 + slack, target_paths=1)  # target already met: runs no frame at all
    assert ren.stats()["frames"] == executed + 1
    pos = (float(g["cam_pos"][0]) + 11.0, float(g["cam_pos"][1]) + 5.0, float(g["cam_pos"][2]) - 3.0)
    ren.set_camera(bm.make_camera(position=pos, direction=g["cam_dir"]))
    oren.camera = ob.make_camera(position=pos, direction=g["cam_dir"])
    oren.reset()
    ren.render(blit, 3)
    for _ in range(3):
        oren.frame()
    c = ren.counters()
    assert [c.primary_ray_cnt, c.start_position, c.frame] == [oren.state.primary_ray_cnt, oren.state.start_position, oren.state.frame]
    assert_records_equal(ren.export_rays(), oren.rays[: c.primary_ray_cnt], what="survivors after the camera move")
    acc = blit.cpu().numpy()
    assert np.array_equal(acc[..., 3], oren.accum[..., 3])
    assert_close_rel(acc, oren.accum, RADIANCE_TOL, "accumulation after the camera move")


def test_render_after_set_counters_needs_records(golden, stores):
    """bm_set_counters with primary_ray_cnt > 0 hands the survivor set to the caller: bm_render refuses to guess (BM_E_STATE)
    until bm_import_rays supplies the records; bm_launch_frame takes them from `queue`."""
    g = golden("256")
    ren = renderer_for(g, stores("256"))
    blit = torch.zeros(int(g["height"]), int(g["width"]), 4, dtype=torch.float32, device="cuda")
    ren.set_counters(primary_ray_cnt=5, start_position=0, frame=1)
    with pytest.raises(bm.BrickmapError):
        ren.render(blit, 1)
    ren.set_counters(primary_ray_cnt=0, start_position=0, frame=1)
    ren.render(blit, 1)


def test_multi_frame_render_uploads_once_while_streaming(oracle, golden):
    """bm_render(frames > 1) on a streaming scene: the staged batch is uploaded before the FIRST frame only (one
    process_load_queue per call, main.cpp:142-143); later frames append requests and must not re-apply stale staging entries
    to new positions. Per call: render 2 frames, stage. The oracle does the same: upload, two frames, stage."""
    g = golden("256")
    cfg = cfg_from_golden(g)
    store = bm.SceneStore(cfg, resident=False)
    ren = renderer_for(g, store, cfg)
    s, oren = oracle_renderer(oracle, g, resident=False)
    blit = torch.zeros(int(g["height"]), int(g["width"]), 4, dtype=torch.float32, device="cuda")
    for call in range(4):
        ren.render(blit, 2)
        cnt, pos = ren.load_queue()
        store.process_load_queue(ren.stream)
        if call:
            s.stream()
        oren.frame(threads=1)
        oren.frame(threads=1)
        ocnt, opos = s.queue()
        assert cnt == ocnt, "call %d: %d requests vs oracle %d" % (call, cnt, ocnt)
        if cnt <= int(g["queue_size"]):
            assert sorted(map(tuple, pos)) == sorted(map(tuple, opos))
    ren.synchronize()
    for sc in range(store.superchunks):
        mine, want = store.indices(sc), s.gpu_indices(sc)
        assert np.array_equal(mine & ~np.uint32(0xFFF), want & ~np.uint32(0xFFF)), "index words (slot bits masked) differ in superchunk %d" % sc
        loaded = np.flatnonzero(mine & 0x80000000)
        assert (mine[loaded] & 0xFFF).max(initial=0) < max(store.brick_count(sc), 1), "slot outside the superchunk's brick array"
    assert_close_rel(blit.cpu().numpy(), oren.accum, RADIANCE_TOL, "accumulation, two frames per streaming call")


def test_exact_paths_mode_matches_oracle(oracle, golden, stores):
    """BM_FRAME_EXACT_PATHS: exactly target_paths paths are started since the reset -- the last frames take only the fresh
    primaries that are still missing, then only survivors. Every pixel ends with exactly spp finished paths, and every frame is
    the reference algorithm run with that frame's slot count (the oracle is driven with the same counts)."""
    g = golden("256")
    ren = renderer_for(g, stores("256"))
    _, oren = oracle_renderer(oracle, g)
    h, w, n = int(g["height"]), int(g["width"]), int(g["n_slots"])
    spp = 3
    target = spp * h * w
    frames = 0
    while oren.stats.terminations < target:
        c = oren.state.primary_ray_cnt
        started = oren.stats.terminations + c
        oren.n_slots = c + min(n - c, target - started)
        assert oren.n_slots > 0
        oren.frame()
        frames += 1
    assert oren.stats.terminations == target and oren.state.primary_ray_cnt == 0
    blit = torch.zeros(h, w, 4, dtype=torch.float32, device="cuda")
    ren.render(blit, frames + 5, target_paths=target, flags=R.FRAME_EXACT_PATHS)
    st = ren.stats()
    assert st["frames"] == frames and st["terminations"] == target
    assert st["extend_rays"] + st["shadow_rays"] == oren.stats.rays
    c = ren.counters()
    assert [c.primary_ray_cnt, c.start_position, c.frame] == [0, oren.state.start_position, oren.state.frame]
    acc = blit.cpu().numpy()
    assert np.all(acc[..., 3] == float(spp)), "every pixel has exactly spp finished paths"
    assert_close_rel(acc, oren.accum, RADIANCE_TOL, "accumulation in exact-paths mode")
    # a sun move resets; the mode then runs the same number of paths again, continuing cursor and frame number
    ren.set_sun(0.3, 0.2)
    ren.render(blit, 100, target_paths=target, flags=R.FRAME_EXACT_PATHS)
    assert ren.stats()["terminations"] == 2 * target and float(blit[..., 3].min()) == float(blit[..., 3].max()) == float(spp)


def test_extend_only_matches_reference_golden(golden, stores):
    """BM_FRAME_EXTEND_ONLY (BASELINE config 2: primary rays only): primary_rays + extend of the reference, nothing shaded."""
    g = golden("4096")
    store = stores("4096")
    ren = renderer_for(g, store)
    state = bm.State(store.cfg)
    ren.launch_kernels(state, flags=R.FRAME_EXTEND_ONLY)
    ext = state.rays("work")
    assert int((ext["distance"] < 1e20).sum()) == int(g["f1_ext_hits"])
    assert_records_equal(ext[g["f1_ext_idx"]], g["f1_ext"], what="post-extend records, extend-only frame")
    c = ren.counters()
    assert [c.primary_ray_cnt, c.shadow_ray_cnt, c.start_position, c.frame] == [0, 0, int(g["f1_counters"][2]), 2]
    assert float(state.blit_buffer.abs().sum()) == 0.0
    # the throughput form of the same thing: frames back to back, the queue holds the last frame's records
    ren2 = renderer_for(g, store)
    q = torch.zeros(store.cfg.ray_queue_buffer_size * 16, dtype=torch.float32, device="cuda")
    ren2.extend_primaries(q, 1)
    rec = q.cpu().numpy().view(np.uint8).view(bm.RAY_DTYPE)
    assert_records_equal(rec[g["f1_ext_idx"]], g["f1_ext"], what="post-extend records, bm_extend_primaries")
    ren2.extend_primaries(q, 2)
    c = ren2.counters()
    assert [c.primary_ray_cnt, c.frame] == [0, 4] and ren2.stats()["extend_rays"] == 3 * store.cfg.ray_queue_buffer_size


def _check_tile_against_oracle(oracle, g, store_for, tile_kw, tile, frames=3):
    """One context restricted to an image partition == the oracle run with the same row mapping: post-extend records and
    survivors bit for bit every frame (bm_launch_frame), alpha exactly and radiance to 1e-4 (both entry points)."""
    w, h = int(g["width"]), int(g["height"])
    cfg = cfg_from_golden(g, **tile_kw)
    store = store_for(cfg)
    ren = renderer_for(g, store, cfg)
    fused = renderer_for(g, store, cfg)
    s = ob.OracleScene(oracle, int(g["grid_size"]), int(g["grid_height"]), int(g["lod2"]), int(g["lod8"]), int(g["queue_size"])).generate_terrain().set_residency(True)
    oren = ob.OracleRenderer(s, w, h, int(g["n_slots"]), ob.make_camera(position=g["cam_pos"], direction=g["cam_dir"]), tuple(float(v) for v in g["sun"]), tile=tile)
    state = bm.State(cfg)
    rows = tile[1]
    assert state.blit_buffer.shape == (rows, w, 4)
    for f in range(frames):
        ren.launch_kernels(state)
        oren.primary_rays()
        oren.set_wavefront_globals()
        oren.extend()
        ext = oren.rays.copy()
        oren.shade()
        oren.connect()
        oren.state.frame += 1
        oren.rays, oren.next = oren.next, oren.rays
        c = ren.counters()
        assert [c.primary_ray_cnt, c.shadow_ray_cnt, c.start_position] == [oren.state.primary_ray_cnt, oren.state.shadow_ray_cnt, oren.state.start_position]
        assert_records_equal(state.rays("work"), ext, what="tile %s frame %d post-extend" % (tile, f + 1))
        assert_records_equal(state.rays("next", c.primary_ray_cnt), oren.rays[: c.primary_ray_cnt], what="tile %s frame %d survivors" % (tile, f + 1))
        state.swap()
    acc = state.blit_buffer.cpu().numpy()
    assert np.array_equal(acc[..., 3], oren.accum[..., 3]), "alpha per pixel of the tile"
    assert_close_rel(acc, oren.accum, RADIANCE_TOL, "tile %s accumulation" % (tile,))
    blit = torch.zeros(rows, w, 4, dtype=torch.float32, device="cuda")
    fused.render(blit, frames)
    acc = blit.cpu().numpy()
    assert np.array_equal(acc[..., 3], oren.accum[..., 3])
    assert_close_rel(acc, oren.accum, RADIANCE_TOL, "tile %s accumulation, fused path" % (tile,))
    assert_records_equal(fused.export_rays(), oren.rays[: oren.state.primary_ray_cnt], what="tile %s survivors, fused path" % (tile,))
    return acc


@pytest.mark.parametrize("world,rank", [(2, 0), (2, 1), (8, 0), (8, 3), (8, 7)])
def test_interleaved_strips_match_oracle_with_the_same_tiling(oracle, golden, stores, world, rank):
    """a rank's context (strips of 8 rows dealt round-robin, the partition bench.py uses) equals the CPU oracle run
    with the same tiling -- bit for bit on rays, exactly on alpha, 1e-4 on radiance."""
    g = golden("256")
    h = int(g["height"])
    rows, image_rows = bm.strip_rows_for_rank(h, rank, world, 8)
    tile = (0, rows, 8, world, rank)
    acc = _check_tile_against_oracle(oracle, g, lambda cfg: stores("256"), dict(tile_rows=rows, strip_rows=8, strip_count=world, strip_index=rank), tile)
    assert acc.shape[0] == len(image_rows)


@pytest.mark.parametrize("world,rank", [(2, 1), (3, 0)])
def test_row_bands_match_oracle_with_the_same_tiling(oracle, golden, stores, world, rank):
    """The contiguous-band partition (bm_config.tile_row0 / tile_rows), including a band height that does not divide the image."""
    g = golden("256")
    row0, rows = bm.tile_rows_for_rank(int(g["height"]), rank, world)
    _check_tile_against_oracle(oracle, g, lambda cfg: stores("256"), dict(tile_row0=row0, tile_rows=rows), (row0, rows, 0, 1, 0))


def test_strips_on_an_outside_view_with_lod_match_oracle(oracle, golden, stores):
    """Strips on the outside-the-world view (AABB entry, both LoD levels), ragged last strip (rows not a multiple of 8 * world)."""
    g = golden("256lod")
    h = int(g["height"])
    rows, _ = bm.strip_rows_for_rank(h, 2, 5, 8)
    _check_tile_against_oracle(oracle, g, lambda cfg: stores("256lod"), dict(tile_rows=rows, strip_rows=8, strip_count=5, strip_index=2), (0, rows, 8, 5, 2), frames=2)


# ---- streaming ----------------------------------------------------------------------------------------------------
def test_streaming_matches_oracle_as_sets(oracle, golden, stores):
    """From an empty device scene: requests (as sorted sets, P4 of SURVEY 8c), index words modulo slot numbers, radiance."""
    g = golden("256")
    cfg = cfg_from_golden(g)
    store = bm.SceneStore(cfg, resident=False)
    ren = renderer_for(g, store, cfg)
    s, oren = oracle_renderer(oracle, g, resident=False)
    state = bm.State(cfg)
    for f in range(1, 7):
        ren.launch_kernels(state)  # includes the upload of what the previous frame requested (kernel.cu:407-414)
        cnt, pos = ren.load_queue()
        store.process_load_queue(ren.stream)  # Scene.cpp:200
        state.swap()
        if f > 1:
            s.stream()
        oren.frame(threads=1)
        ocnt, opos = s.queue()
        assert cnt == ocnt, "frame %d: %d requests vs oracle %d" % (f, cnt, ocnt)
        if cnt <= int(g["queue_size"]):
            assert sorted(map(tuple, pos)) == sorted(map(tuple, opos)), "frame %d request sets differ" % f
        else:
            assert len(set(map(tuple, pos))) == len(pos) == int(g["queue_size"])
        c = ren.counters()
        assert [c.primary_ray_cnt, c.shadow_ray_cnt] == [oren.state.primary_ray_cnt, oren.state.shadow_ray_cnt]
    ren.synchronize()
    for sc in range(store.superchunks):
        mine, want = store.indices(sc), s.gpu_indices(sc)
        assert np.array_equal(mine & ~np.uint32(0xFFF), want & ~np.uint32(0xFFF)), "index words (slot bits masked) differ in superchunk %d" % sc
        loaded = np.flatnonzero(mine & 0x80000000)
        if loaded.size:
            gb = store.bricks(sc, gpu_view=True)
            for cell in loaded[:: max(1, loaded.size // 16)]:
                assert np.array_equal(gb[mine[cell] & 0xFFF], s.gpu_brick(sc, int(want[cell] & 0xFFF))), "brick content through the indirection"
    assert_close_rel(state.blit_buffer.cpu().numpy(), oren.accum, RADIANCE_TOL, "accumulation while streaming")


def test_streaming_with_a_large_queue_matches_oracle(oracle, lib):
    """brick_load_queue_size is a real parameter: 16 384 here (the reference's 1024, variables.h:35, needs 12 frames for what this
    view requests in one). 12 k requests in frame 1: request sets, index words (slot bits masked: the queue order of a parallel run
    is scheduling-dependent), slot numbers dense per superchunk, brick contents through the indirection, radiance.
    Exercises the sort-based streaming step (bm_scene_store_stream) far beyond one block's worth of requests."""
    q = 16384
    cfg = bm.default_config(grid_size=1024, grid_height=256, brick_load_queue_size=q, ray_queue_buffer_size=262144, screen_width=640, screen_height=360)
    store = bm.SceneStore(cfg, resident=False)
    osc = ob.OracleScene(oracle, 1024, 256, cfg.lod_distance_2x2x2, cfg.lod_distance_8x8x8, q).generate_terrain()
    d = np.array([0.5, 0.6, -0.62], np.float32)
    d = (d / np.sqrt((d.astype(np.float64) ** 2).sum())).astype(np.float32)
    pos = (100.0, 80.0, 420.0)
    ren = bm.Renderer(cfg, store)
    ren.set_camera(bm.make_camera(position=pos, direction=d))
    oren = ob.OracleRenderer(osc, 640, 360, 262144, ob.make_camera(position=pos, direction=d))
    state = bm.State(cfg)
    counts = []
    for f in range(1, 5):
        ren.launch_kernels(state)
        cnt, rpos = ren.load_queue()
        staged = store.process_load_queue(ren.stream, want_count=True)
        state.swap()
        if f > 1:
            osc.stream()
        oren.frame(threads=1)
        ocnt, opos = osc.queue()
        assert cnt == ocnt <= q and staged == cnt
        assert sorted(map(tuple, rpos)) == sorted(map(tuple, opos)), "frame %d request sets differ" % f
        counts.append(cnt)
        c = ren.counters()
        assert [c.primary_ray_cnt, c.shadow_ray_cnt] == [oren.state.primary_ray_cnt, oren.state.shadow_ray_cnt]
    assert counts[0] > 8 * 1024, "the view must request far more than the reference's queue holds"
    ren.synchronize()
    for sc in range(store.superchunks):
        mine, want = store.indices(sc), osc.gpu_indices(sc)
        assert np.array_equal(mine & ~np.uint32(0xFFF), want & ~np.uint32(0xFFF)), "index words (slot bits masked) differ in superchunk %d" % sc
        loaded = np.flatnonzero(mine & 0x80000000)
        if loaded.size:
            assert sorted(mine[loaded] & 0xFFF) == list(range(loaded.size)), "slots of superchunk %d are not 0..n-1" % sc
            gb = store.bricks(sc, gpu_view=True)
            for cell in loaded[:: max(1, loaded.size // 8)]:
                assert np.array_equal(gb[mine[cell] & 0xFFF], osc.gpu_brick(sc, int(want[cell] & 0xFFF))), "brick content through the indirection"


def test_streaming_serial_order_matches_reference_golden(golden):
    """Frame 1 of the reference's streaming fixture: the request SET of an all-unloaded scene."""
    g = golden("256")
    cfg = cfg_from_golden(g)
    store = bm.SceneStore(cfg, resident=False)
    ren = renderer_for(g, store, cfg)
    state = bm.State(cfg)
    ren.launch_kernels(state)
    cnt, pos = ren.load_queue()
    assert cnt == int(g["stream1_count"])
    assert sorted(map(tuple, pos)) == sorted(map(tuple, g["stream1_positions"]))


def test_request_merge_kernel_matches_specification(golden):
    """bm_requests_merge (CUDA) == brickmap_b200.parallel.merge_request_blocks (the specification the gloo test uses)."""
    import ctypes as C
    from brickmap_b200.parallel import merge_request_blocks
    g = golden("256")
    cfg = cfg_from_golden(g)
    q = cfg.brick_load_queue_size
    store = bm.SceneStore(cfg, resident=False)
    ren = renderer_for(g, store, cfg)
    rng = np.random.default_rng(3)
    # candidate cells: non-empty ones, taken from the reference's own request list
    cells = np.concatenate([g["stream1_positions"], g["stream2_positions"], g["stream3_positions"]])
    world = 3
    blocks = np.zeros((world, 1 + 3 * q), np.int32)
    for r in range(world):
        pick = cells[rng.choice(len(cells), size=[700, 300, 600][r], replace=False)]
        blocks[r, 0] = len(pick)
        blocks[r, 1:1 + 3 * len(pick)] = pick.reshape(-1)
    total, kept, dropped = merge_request_blocks(blocks, q)
    assert total > q and len(dropped) > 0
    dev = torch.as_tensor(blocks, device="cuda")
    lib = bm.load()
    assert lib.bm_requests_merge(ren.h, dev.data_ptr(), world) == 0
    cnt, pos = ren.load_queue()
    assert cnt == total and np.array_equal(pos, kept)
    idx = {sc: store.indices(sc) for sc in range(store.superchunks)}

    def word(p):
        sc = (p[0] >> 4) + (p[1] >> 4) * (cfg.grid_size // 128) + (p[2] >> 4) * (cfg.grid_size // 128) ** 2
        return idx[sc][(p[0] & 15) + (p[1] & 15) * 16 + (p[2] & 15) * 256]
    assert all(word(p) & 0x20000000 for p in kept), "kept requests carry the requested bit"
    assert not any(word(p) & 0x20000000 for p in dropped), "dropped requests are released for a later frame"


def test_request_merge_kernel_at_eight_ranks_and_a_large_queue():
    """The merge at the sizes the multi-GPU benchmark and BASELINE config 4 need (8 ranks x 16 384 entries: far beyond one block's
    shared memory, where round 1's kernel stopped), duplicates within and across ranks, against the specification function."""
    from brickmap_b200.parallel import merge_request_blocks
    q, world = 16384, 8
    cfg = bm.default_config(grid_size=1024, grid_height=256, brick_load_queue_size=q, ray_queue_buffer_size=65536, screen_width=64, screen_height=64)
    store = bm.SceneStore(cfg, resident=False)
    ren = bm.Renderer(cfg, store)
    rng = np.random.default_rng(5)
    idx = np.concatenate([store.indices(sc) for sc in range(store.superchunks)])
    nz = np.flatnonzero(idx)
    sc_, loc = nz // 4096, nz % 4096
    sx, sy, sz = sc_ % 8, (sc_ // 8) % 8, sc_ // 64
    cells = np.stack([sx * 16 + (loc & 15), sy * 16 + ((loc >> 4) & 15), sz * 16 + (loc >> 8)], axis=1).astype(np.int32)
    pool = cells[rng.choice(len(cells), size=40000, replace=False)]
    blocks = np.zeros((world, 1 + 3 * q), np.int32)
    for r in range(world):
        n = [16384, 9000, 0, 16384, 1, 12000, 16384, 7777][r]
        pick = pool[rng.integers(0, len(pool), n)]  # with repetitions, inside and across ranks
        blocks[r, 0] = n if r != 3 else n + 500  # a count beyond the queue size is clamped (kernel.cu:409)
        blocks[r, 1:1 + 3 * n] = pick.reshape(-1)
    total, kept, dropped = merge_request_blocks(blocks, q)
    assert total > q and len(kept) == q and len(dropped) > 0
    dev = torch.as_tensor(blocks, device="cuda")
    assert bm.load().bm_requests_merge(ren.h, dev.data_ptr(), world) == 0
    cnt, pos = ren.load_queue()
    assert cnt == total and np.array_equal(pos, kept)
    idx2 = np.concatenate([store.indices(sc) for sc in range(store.superchunks)])

    def words(p):
        sc = (p[:, 0] >> 4) + (p[:, 1] >> 4) * 8 + (p[:, 2] >> 4) * 64
        return idx2[sc * 4096 + (p[:, 0] & 15) + (p[:, 1] & 15) * 16 + (p[:, 2] & 15) * 256]
    assert np.all(words(kept) & 0x20000000) and not np.any(words(dropped) & 0x20000000)


# ---- caves world (BASELINE config 4 at a size the oracle can build) + LoD + tone map ------------------------------------
def test_caves_world_streaming_with_lod_matches_oracle(oracle, lib):
    """A sparse 3-D cave world (not in the reference: integer lattice noise, identical in the oracle and on the device),
    LoD thresholds small enough that all three branches of voxel.cuh:212-227 run, everything streamed in from an empty
    device scene through the request queue."""
    cfg = bm.default_config(grid_size=512, grid_height=256, lod_distance_2x2x2=150, lod_distance_8x8x8=900, brick_load_queue_size=1024,
                            ray_queue_buffer_size=131072, screen_width=384, screen_height=256)
    store = bm.SceneStore(cfg, kind=R.SCENE_CAVES, seed=7, resident=False)
    osc = ob.OracleScene(oracle, 512, 256, 150, 900, 1024).generate_caves(seed=7)
    assert store.total_bricks > 1000
    for sc in range(store.superchunks):
        assert np.array_equal(store.indices(sc, host_view=True), osc.host_indices(sc)), "cave generator differs in superchunk %d" % sc
        assert np.array_equal(store.bricks(sc), osc.host_bricks(sc))
    pos, d = (-40.0, 250.0, 120.0), np.array([0.94, 0.1, 0.05], np.float32)
    d = (d / np.sqrt((d.astype(np.float64) ** 2).sum())).astype(np.float32)
    ren = bm.Renderer(cfg, store)
    ren.set_camera(bm.make_camera(position=pos, direction=d))
    oren = ob.OracleRenderer(osc, 384, 256, 131072, ob.make_camera(position=pos, direction=d))
    state = bm.State(cfg)
    total_requests = 0
    for f in range(1, 6):
        ren.launch_kernels(state)
        cnt, rpos = ren.load_queue()
        store.process_load_queue(ren.stream)
        state.swap()
        if f > 1:
            osc.stream()
        oren.frame(threads=1)
        ocnt, opos = osc.queue()
        assert cnt == ocnt
        if cnt <= 1024:
            assert sorted(map(tuple, rpos)) == sorted(map(tuple, opos))
        total_requests += min(cnt, 1024)
        c = ren.counters()
        assert [c.primary_ray_cnt, c.shadow_ray_cnt] == [oren.state.primary_ray_cnt, oren.state.shadow_ray_cnt]
    assert total_requests > 100 and oren.stats.lod_bytes > 0, "the view must exercise streaming and the 2x2x2 LoD"
    assert_close_rel(state.blit_buffer.cpu().numpy(), oren.accum, RADIANCE_TOL, "cave world accumulation")


def test_tonemap_matches_oracle(oracle, golden, stores):
    """bm_tonemap == blit_onto_framebuffer's colour math (kernel.cu:355-362): rgb / alpha, gamma 1/2.2, alpha 1."""
    g = golden("256")
    ren = renderer_for(g, stores("256"))
    blit = torch.zeros(int(g["height"]), int(g["width"]), 4, dtype=torch.float32, device="cuda")
    ren.render(blit, 4)
    got = ren.tonemap(blit).cpu().numpy()
    acc = blit.cpu().numpy()
    want = np.zeros_like(acc)
    oracle.lib.orc_tonemap(acc.ctypes.data, acc.shape[0] * acc.shape[1], want.ctypes.data)
    ok = acc[..., 3] > 0
    assert ok.mean() > 0.9
    assert_close_rel(got[ok], want[ok], 1e-5, "tone-mapped image")


@pytest.mark.parametrize("view", [1, 4, 5, 6, 7, 8])
def test_camera_tour_views_match_oracle(oracle, golden, stores, view):
    """The reference's benchmark views (performance_measure.h:4-25; brickmap_b200/views.py). Views 4-8 stand outside the stock
    world: every primary ray enters through the AABB (voxel.cuh:142-155, never suspended in the throughput kernel) and the far
    cells are 8x8x8 boxes (voxel.cuh:212-214). Two frames through both entry points against the oracle."""
    from brickmap_b200.views import tour
    g = golden("4096")
    pos, d = tour()[view]
    w, h, n = 320, 180, 65536
    cfg = cfg_from_golden(g)
    cfg.screen_width, cfg.screen_height, cfg.ray_queue_buffer_size = w, h, n
    store = stores("4096")
    ren = bm.Renderer(cfg, store)
    ren.set_camera(bm.make_camera(position=pos, direction=d))
    fused = bm.Renderer(cfg, store)
    fused.set_camera(bm.make_camera(position=pos, direction=d))
    s = _oracle_scene_4096(oracle)
    oren = ob.OracleRenderer(s, w, h, n, ob.make_camera(position=pos, direction=d))
    state = bm.State(cfg)
    for f in range(2):
        ren.launch_kernels(state)
        oren.primary_rays()
        oren.set_wavefront_globals()
        oren.extend()
        ext = oren.rays.copy()
        oren.shade()
        oren.connect()
        oren.state.frame += 1
        oren.rays, oren.next = oren.next, oren.rays
        c = ren.counters()
        assert [c.primary_ray_cnt, c.shadow_ray_cnt] == [oren.state.primary_ray_cnt, oren.state.shadow_ray_cnt]
        assert_records_equal(state.rays("work"), ext, what="view %d frame %d post-extend" % (view, f + 1))
        assert_records_equal(state.rays("next", c.primary_ray_cnt), oren.rays[: c.primary_ray_cnt], what="view %d frame %d survivors" % (view, f + 1))
        state.swap()
    if view in (1, 4):  # (view 7 looks past the world: sky only, like in the reference's tour)
        assert oren.stats.hits > 0, "the view must see the world"
    assert_close_rel(state.blit_buffer.cpu().numpy(), oren.accum, RADIANCE_TOL, "view %d accumulation" % view)
    blit = torch.zeros(h, w, 4, dtype=torch.float32, device="cuda")
    fused.render(blit, 2)
    assert_close_rel(blit.cpu().numpy(), oren.accum, RADIANCE_TOL, "view %d accumulation, fused path" % view)


_ORACLE_4096 = []


def _oracle_scene_4096(oracle):
    if not _ORACLE_4096:
        _ORACLE_4096.append(ob.OracleScene(oracle, 4096, 512).generate_terrain().set_residency(True))
    return _ORACLE_4096[0]


def test_camera_inside_a_solid_voxel_terminates_like_the_reference(oracle, golden, stores):
    """A camera inside rock: every primary hits at distance 0 with the normal still (0,0,0) (voxel.cuh:114-119), the bounce off a
    zero normal is a NaN direction (kernel.cu:76-84,293), and the reference's AABB test then rejects the NaN ray (all comparisons
    false, voxel.cuh:23): the path ends as a miss with NaN radiance. The product must do the same -- in particular a NaN ray must
    never reach the DDA, which would spin on NaN tmax. NaNs are compared as NaNs (their payload bits differ between CPU and GPU)."""
    g = golden("256")
    w, h, n = 128, 96, 16384
    cfg = cfg_from_golden(g)
    cfg.screen_width, cfg.screen_height, cfg.ray_queue_buffer_size = w, h, n
    store = stores("256")
    pos = (100.5, 90.25, 20.75)
    ren = bm.Renderer(cfg, store)
    ren.set_camera(bm.make_camera(position=pos, direction=g["cam_dir"]))
    s = ob.OracleScene(oracle, 256, 256).generate_terrain().set_residency(True)
    oren = ob.OracleRenderer(s, w, h, n, ob.make_camera(position=pos, direction=g["cam_dir"]))
    blit = torch.zeros(h, w, 4, dtype=torch.float32, device="cuda")
    for f in range(4):
        ren.render(blit, 1)
        oren.frame()
        c = ren.counters()
        assert [c.primary_ray_cnt, c.start_position, c.frame] == [oren.state.primary_ray_cnt, oren.state.start_position, oren.state.frame]
        mine, want = ren.export_rays(), oren.rays[: c.primary_ray_cnt]
        for fld in ("origin", "direction", "normal"):
            a, b = mine[fld], want[fld]
            assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(bits(a)[~np.isnan(a)], bits(b)[~np.isnan(b)]), "frame %d field %s" % (f + 1, fld)
    assert oren.state.frame == 5 and np.isnan(oren.accum).any(), "the scenario must produce NaN rays"
    acc = blit.cpu().numpy()
    assert np.array_equal(np.isnan(acc), np.isnan(oren.accum)) and np.array_equal(acc[..., 3], oren.accum[..., 3])


# ---- against the live reference -------------------------------------------------------------------------------------
@pytest.mark.skipif(not ob.Reference.available("256"), reason="oracle/_ref not built")
def test_drop_in_on_the_reference_hosts_own_scene(golden):
    """The reference's Scene (host code, Scene.cpp) owns the device scene; our kernels run on ITS GPUScene pointers, through
    the pointer tables of separately cudaMalloc'd superchunks, next to the reference's own kernels."""
    g = golden("256")
    ref = ob.Reference("256", int(g["width"]), int(g["height"]))
    ref.generate()
    ref.force_resident()
    cam = ob.make_camera(position=g["cam_pos"], direction=g["cam_dir"])
    ref.set_camera(cam)
    ref.set_sun(0.05, 0.1)
    ref.upload_sun()
    cfg = cfg_from_golden(g)
    ren = bm.Renderer(cfg, bm.GpuScene(*ref.scene_pointers()))
    ren.set_camera(bm.make_camera(position=g["cam_pos"], direction=g["cam_dir"]))
    ren.set_sun(0.05, 0.1)
    state = bm.State(cfg)
    ref.clear_accum()
    ref.write_counters(primary_ray_cnt=0, start_position=0, raynr_primary=0, raynr_extend=0, raynr_shade=0, raynr_connect=0, shadow_ray_cnt=0)
    for f in (1, 2):
        ref.run_stage("primary_rays", frame=f)
        ref.run_stage("set_wavefront_globals")
        ref.run_stage("extend", frame=f)
        r_ext = ref.read_rays(0)
        ref.run_stage("shade", frame=f, serial=True)
        rc = ref.counters()
        r_next = ref.read_rays(1, 0, rc["primary_ray_cnt"])
        ref.run_stage("connect", frame=f)
        ref.swap_buffers()
        ren.launch_kernels(state)
        c = ren.counters()
        assert c.primary_ray_cnt == rc["primary_ray_cnt"] and c.shadow_ray_cnt == rc["shadow_ray_cnt"]
        assert_records_equal(state.rays("work"), r_ext, what="frame %d post-extend vs live reference" % f)
        assert_records_equal(state.rays("next", c.primary_ray_cnt), r_next, what="frame %d survivors vs live reference" % f)
        state.swap()
    assert_close_rel(state.blit_buffer.cpu().numpy(), ref.read_accum(), RADIANCE_TOL, "accumulation vs live reference")


@pytest.mark.skipif(not ob.Reference.available("256"), reason="oracle/_ref not built")
def test_statistical_parity_with_reference_parallel_run(golden, stores):
    """P3 of SURVEY 8c: the reference's normal (parallel, non-deterministic after frame 1) run and ours converge to the
    same image: tile means of the 40-frame averages agree within Monte-Carlo noise."""
    g = golden("256")
    w, h = int(g["width"]), int(g["height"])
    ref = ob.Reference("256", w, h)
    ref.generate()
    ref.force_resident()
    ref.set_camera(ob.make_camera(position=g["cam_pos"], direction=g["cam_dir"]))
    ref.set_sun(0.05, 0.1)
    ref.run_frames(40)
    r = ref.read_accum()
    ren = renderer_for(g, stores("256"))
    blit = torch.zeros(h, w, 4, dtype=torch.float32, device="cuda")
    ren.render(blit, 40)
    m = blit.cpu().numpy()
    assert abs(m[..., 3].sum() / r[..., 3].sum() - 1) < 0.01
    rm, mm = tile_means(r), tile_means(m)
    rel = np.abs(mm[..., :3] / mm[..., 3:] - rm[..., :3] / rm[..., 3:]) / np.maximum(rm[..., :3] / rm[..., 3:], 1e-6)
    assert np.median(rel) < 0.02 and rel.max() < 0.15, "tile means differ: median %.3f max %.3f" % (np.median(rel), rel.max())


@pytest.mark.skipif(not (ob.Reference.available("256") and ob.Reference.available("dropin_256")), reason="oracle/_ref not built")
def test_reference_host_main_loop_on_new_kernels(golden, tmp_path):
    """integration/launch_kernels_dropin.cpp: the reference's own host code (Scene::generate, State, launch_kernels call,
    Scene::process_load_queue with its pinned staging buffers and growing brick arrays, buffer swap: main.cpp:142-146) runs
    unchanged on top of libbrickmap_b200.so. Compared with the same loop on the reference's kernels: frame 1 is deterministic
    in both (requests as a set, accumulation to 1e-4); later frames of the reference depend on thread scheduling, so after the
    world has streamed in only the image statistics are compared."""
    import os
    import subprocess
    import sys
    g = golden("256")
    runs = {}
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_loop_worker.py")
    for variant in ("256", "dropin_256"):
        out = str(tmp_path / ("%s.npz" % variant))
        subprocess.run([sys.executable, worker, variant, out, "40"], check=True, timeout=600)
        r = np.load(out)
        runs[variant] = (int(r["count"]), r["positions"], r["accum1"], r["accum"], r["indices"])
    (rc, rp, ra1, ra, ri), (dc, dp, da1, da, di) = runs["256"], runs["dropin_256"]
    assert rc == dc == int(g["stream1_count"]) and sorted(map(tuple, rp)) == sorted(map(tuple, dp))
    assert_close_rel(da1, ra1, RADIANCE_TOL, "frame 1 accumulation, reference host on new kernels vs on its own kernels")
    # which rarely-hit bricks a random bounce ray touches differs between any two runs of the reference itself
    r_loaded, d_loaded = (ri & 0x80000000) != 0, (di & 0x80000000) != 0
    assert r_loaded.sum() > 1000 and (r_loaded != d_loaded).sum() < 0.03 * r_loaded.sum(), "resident brick sets diverge: %d vs %d, %d differ" % (
        r_loaded.sum(), d_loaded.sum(), (r_loaded != d_loaded).sum())
    assert np.array_equal(ri == 0, di == 0)
    assert abs(da[..., 3].sum() / ra[..., 3].sum() - 1) < 0.01
    rm, dm = tile_means(ra), tile_means(da)
    rel = np.abs(dm[..., :3] / dm[..., 3:] - rm[..., :3] / rm[..., 3:]) / np.maximum(rm[..., :3] / rm[..., 3:], 1e-6)
    assert np.median(rel) < 0.02 and rel.max() < 0.15


@pytest.mark.skipif(not ob.Reference.available("4096"), reason="oracle/_ref not built")
def test_statistical_parity_at_the_benchmark_config(golden, stores):
    """P3 of SURVEY 8c at BASELINE config 3 itself: 1920x1080, 16 spp, stock world, benchmark camera. The reference's normal
    parallel run (non-deterministic slot order after frame 1) and bm_render converge to the same image: per-tile mean radiance
    (8x8 tiles of 240x135 pixels, ~0.5 M paths each) within Monte-Carlo noise."""
    g = golden("4096")
    w, h = int(g["width"]), int(g["height"])
    target = 16 * w * h
    ref = ob.Reference("4096", w, h)
    import os
    import sys
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        ref.generate()
    finally:
        os.dup2(saved, 1)
        os.close(saved)
    ref.force_resident()
    ref.set_camera(ob.make_camera(position=g["cam_pos"], direction=g["cam_dir"]))
    ref.set_sun(float(g["sun"][0]), float(g["sun"][1]))
    frames = 0
    while ref.alpha_sum() < target and frames < 64:
        ref.run_frames(1)
        frames += 1
    r = ref.read_accum()
    ren = renderer_for(g, stores("4096"))
    blit = torch.zeros(h, w, 4, dtype=torch.float32, device="cuda")
    ren.render(blit, 64, target_paths=target)
    m = blit.cpu().numpy()
    assert abs(ren.stats()["frames"] - frames) <= 1
    assert abs(m[..., 3].sum() / r[..., 3].sum() - 1) < 0.04
    rm, mm = tile_means(r), tile_means(m)
    ra, ma = rm[..., :3] / rm[..., 3:], mm[..., :3] / mm[..., 3:]
    rel = np.abs(ma - ra) / np.maximum(ra, 1e-6)
    assert np.median(rel) < 0.005 and rel.max() < 0.03, "tile means differ: median %.4f max %.4f" % (np.median(rel), rel.max())
