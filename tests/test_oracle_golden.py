"""Pins the CPU oracle against outputs of the reference itself (CPU-only test).

tests/golden/golden_<variant>.npz were produced by the UNMODIFIED reference kernels on a B200
(tests/golden/make_golden.py). The reference has no tests or fixtures of its own, so these files are the
pin: scene generation (Scene.cpp:44-116), sky (sunsky.cu), traversal (voxel.cuh), the wavefront stages
(kernel.cu:154-346) over several frames with the canonical (slot-ordered) schedule, and the streaming protocol
(voxel.cuh:228-241, Scene.cpp:200-229, kernel.cu:141-151).

Tolerances: every geometric quantity (ray records, distances, normals, request positions, index words) bit for bit;
radiance 1e-4 relative (BASELINE.json north_star), in practice < 1e-5.
"""
import hashlib

import numpy as np
import pytest

from helpers import assert_close_rel, assert_records_equal, bits, tile_means
from oracle import binding as ob

RADIANCE_TOL = 1e-4


def make_scene(oracle, g):
    return ob.OracleScene(oracle, int(g["grid_size"]), int(g["grid_height"]), int(g["lod2"]), int(g["lod8"]), int(g["queue_size"])).generate_terrain()


def make_renderer(scene, g):
    cam = ob.make_camera(position=g["cam_pos"], direction=g["cam_dir"], focal=float(g["focal"]) if "focal" in g else 1.0, lens=float(g["lens"]) if "lens" in g else 0.0)
    return ob.OracleRenderer(scene, int(g["width"]), int(g["height"]), int(g["n_slots"]), cam, tuple(float(v) for v in g["sun"]))


@pytest.fixture(scope="module")
def scenes(oracle, golden):
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = make_scene(oracle, golden(name))
        return cache[name]
    return get


@pytest.mark.parametrize("variant", ["256", "256lod", "4096"])
def test_scene_generation_matches_reference(oracle, golden, scenes, variant):
    g, s = golden(variant), scenes(variant)
    h = hashlib.sha256()
    counts = []
    for sc in range(s.supergrid_count):
        h.update(s.host_indices(sc).tobytes())
        h.update(s.host_bricks(sc).tobytes())
        counts.append(s.brick_count(sc))
    assert np.array_equal(np.array(counts, np.uint32), g["brick_counts"])
    assert np.array_equal(s.host_indices(0), g["sc0_indices"])
    assert np.array_equal(s.host_bricks(0)[:64], g["sc0_bricks"])
    assert h.hexdigest() == str(g["scene_sha256"])


@pytest.mark.parametrize("variant", ["256", "4096"])
def test_sky_matches_reference(oracle, golden, variant):
    g = golden(variant)
    sun_dir = np.zeros(3, np.float32)
    oracle.lib.orc_sun_direction(float(g["sun"][0]), float(g["sun"][1]), sun_dir.ctypes.data)
    assert np.array_equal(bits(sun_dir), bits(g["sun_dir"])), "sun direction must match bit for bit (it feeds the shadow rays)"
    for mode, nm in enumerate(("sun", "sky", "sunsky")):
        assert_close_rel(ob.sky_eval(oracle, g["sky_dirs"], mode, sun_dir), g["sky_" + nm], RADIANCE_TOL, nm)


@pytest.mark.parametrize("variant", ["256", "256lod", "4096"])
def test_traversal_matches_reference(oracle, golden, scenes, variant):
    g, s = golden(variant), scenes(variant)
    s.set_residency(True)
    n = g["trace_origins"].shape[0]
    cam_cell = [int(v / 8.0) for v in g["cam_pos"]]
    hit, dist, nrm = s.trace(g["trace_origins"], g["trace_directions"], cam_cell, distances=np.full(n, 1e20, np.float32), threads=0)
    assert np.array_equal(bits(dist), bits(g["trace_distance"]))
    assert np.array_equal(bits(nrm), bits(g["trace_normal"]))
    assert int(hit.sum()) == int((g["trace_distance"] < 1e20).sum())
    assert 0 < hit.sum() < n


@pytest.mark.parametrize("variant", ["256", "256lod", "4096", "256lens"])
def test_canonical_frames_match_reference(oracle, golden, scenes, variant):
    """256lens: thin-lens camera (lens radius 0.75, focal distance 12.5): pins ConcentricSampleDisk (kernel.cu:85-103), the order of
    the two lens draws (unspecified in the source, kernel.cu:194) and the FMA placement of kernel.cu:191-198 in the reference build."""
    g, s = golden(variant), scenes(variant)
    if variant == "256lens":
        assert float(g["lens"]) > 0
    s.set_residency(True)
    ren = make_renderer(s, g)
    f = 1
    while "f%d_counters" % f in g:
        p = "f%d_" % f
        ren.primary_rays()
        ren.set_wavefront_globals()
        ren.extend()
        ext = ren.rays.copy()
        ren.shade()
        nxt = ren.next[: ren.state.primary_ray_cnt].copy()
        sh = ren.shadows[: ren.state.shadow_ray_cnt].copy()
        ren.connect()
        ren.state.frame += 1
        ren.rays, ren.next = ren.next, ren.rays
        assert [ren.state.primary_ray_cnt, ren.state.shadow_ray_cnt, ren.state.start_position] == [int(v) for v in g[p + "counters"]]
        assert int((ext["distance"] < 1e20).sum()) == int(g[p + "ext_hits"])
        assert_records_equal(ext[g[p + "ext_idx"]], g[p + "ext"], what="frame %d post-extend" % f)
        assert_records_equal(nxt[g[p + "next_idx"]], g[p + "next"], what="frame %d survivors" % f)
        assert_records_equal(sh[g[p + "shadow_idx"]], g[p + "shadow"], fields=("origin", "direction", "pixel_index"), what="frame %d shadow rays" % f)
        assert_close_rel(sh[g[p + "shadow_idx"]]["color"], g[p + "shadow"]["color"], RADIANCE_TOL, "frame %d shadow colour" % f)
        acc = ren.accum
        assert_close_rel(acc.reshape(-1, 4)[g["accum_pix"]], g[p + "accum_val"], RADIANCE_TOL, "frame %d accumulation samples" % f)
        assert_close_rel(tile_means(acc), g[p + "accum_tiles"], RADIANCE_TOL, "frame %d accumulation tile means" % f)
        assert abs(float(acc[..., 3].astype(np.float64).sum()) - float(g[p + "alpha_sum"])) < 0.5
        f += 1
    assert f > 2


def test_streaming_matches_reference(oracle, golden):
    g = golden("256")
    s = make_scene(oracle, g)  # nothing resident (Scene.cpp:157-164)
    ren = make_renderer(s, g)
    f = 1
    while "stream%d_count" % f in g:
        s.stream() if f > 1 else None  # upload of what the previous frame requested (kernel.cu:408-414) ...
        ren.frame(threads=1)
        cnt, pos = s.queue()
        assert cnt == int(g["stream%d_count" % f])
        assert np.array_equal(pos, g["stream%d_positions" % f]), "request queue must match in content AND order (serial schedule)"
        assert [ren.state.primary_ray_cnt, ren.state.shadow_ray_cnt, ren.state.start_position] == [int(v) for v in g["stream%d_counters" % f]]
        f += 1
    assert f == 4
    assert_close_rel(ren.accum.reshape(-1, 4)[g["stream_accum_pix"]], g["stream_accum_val"], RADIANCE_TOL, "accumulation after streaming frames")


def test_tiled_oracle_is_the_reference_mapping_inside_the_tile(oracle, golden, scenes):
    """The multi-GPU image partition of the oracle (oracle.h, Tile): the whole-image tile is the reference's own mapping bit for
    bit; a strip tile visits its own pixels in raster order (kernel.cu:170-171 inside the tile) and points each ray through
    the full-image row its buffer row stands for (kernel.cu:183-184)."""
    g, s = golden("256"), scenes("256")
    s.set_residency(True)
    w, h, n = 96, 64, 96 * 64
    cam = ob.make_camera(position=g["cam_pos"], direction=g["cam_dir"])
    full = ob.OracleRenderer(s, w, h, n, cam)
    full.primary_rays()
    whole = ob.OracleRenderer(s, w, h, n, cam, tile=(0, h, 0, 1, 0))
    whole.primary_rays()
    assert full.rays.tobytes() == whole.rays.tobytes()
    right, up = full.camera_basis()
    d0 = np.asarray(g["cam_dir"], np.float64)
    for tile, image_rows in (((0, 32, 8, 2, 1), [r for r in range(h) if (r // 8) % 2 == 1]), ((16, 24, 0, 1, 0), list(range(16, 40)))):
        t = ob.OracleRenderer(s, w, h, tile[1] * w, cam, tile=tile)
        t.primary_rays()
        assert np.array_equal(t.rays["pixel_index"], np.arange(tile[1] * w, dtype=np.uint32))
        # the ray of buffer pixel (x, r) goes through image pixel (x, image_rows[r]) up to the sub-pixel jitter
        x = (np.arange(tile[1] * w) % w).astype(np.float64)
        y = np.repeat(np.array(image_rows, np.float64), w)
        lo = d0 + ((x - 1) / w - 0.5)[:, None] * right + ((h - y) / h - 0.5)[:, None] * up
        hi = d0 + ((x + 0) / w - 0.5)[:, None] * right + ((h - (y - 1)) / h - 0.5)[:, None] * up
        mid = 0.5 * (lo + hi)
        mid /= np.linalg.norm(mid, axis=1, keepdims=True)
        cosang = (mid * t.rays["direction"].astype(np.float64)).sum(1)
        assert cosang.min() > np.cos(2.5 / h), "tile %s: a ray leaves its pixel (min cos %.6f)" % (tile, cosang.min())
        assert t.accum.shape == (tile[1], w, 4)
    # two strip instances together finish the same number of paths as their pixels ask for
    t = ob.OracleRenderer(s, w, h, 32 * w, cam, tile=(0, 32, 8, 2, 0))
    for _ in range(3):
        t.frame()
    assert t.accum[..., 3].sum() == t.stats.terminations > 0
