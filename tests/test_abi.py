"""CPU-only: the C-ABI library loads, exports every symbol include/brickmap_b200.h declares, lays its records out like
the reference's structs, and fails loudly (error code + message, no fallback) when there is no GPU to run on."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import brickmap_b200 as bm
from brickmap_b200 import _lib
from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "brickmap_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bm_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libbrickmap_b200.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, "python binding and header disagree: %s" % (set(_lib.SIGNATURES) ^ set(names))


def test_record_layouts_match_reference_structs():
    # RayQueue 64 B, ShadowQueue 40 B, Brick 64 B, GPUScene 48 B (sizes from compiling the reference headers)
    assert bm.RAY_DTYPE.itemsize == 64 and bm.SHADOW_DTYPE.itemsize == 40
    assert C.sizeof(bm.GpuScene) == 48
    assert bm.RAY_DTYPE.fields["distance"][1] == 48 and bm.RAY_DTYPE.fields["pixel_index"][1] == 60
    assert bm.SHADOW_DTYPE.fields["pixel_index"][1] == 36


def test_default_config_is_the_reference_constants(lib):
    cfg = bm.default_config()
    assert (cfg.grid_size, cfg.grid_height) == (4096, 512)            # variables.h:7-8
    assert (cfg.lod_distance_2x2x2, cfg.lod_distance_8x8x8) == (100000, 600000)  # variables.h:25-27
    assert cfg.brick_load_queue_size == 1024                           # variables.h:35
    assert cfg.ray_queue_buffer_size == 2 * 1048576                    # variables.h:61


def test_argument_validation_without_gpu(lib):
    h = C.c_void_p()
    bad = bm.default_config(grid_size=100)
    assert lib.bm_create(C.byref(h), C.byref(bad)) == -1  # BM_E_INVALID
    assert b"multiples of 128" in lib.bm_last_error_string()
    bad = bm.default_config(tile_row0=1000, tile_rows=200)
    assert lib.bm_create(C.byref(h), C.byref(bad)) == -1
    assert lib.bm_create(None, None) == -1
    assert lib.bm_render(None, None, 1, 0, 0, 1) == -1
    assert lib.bm_set_camera(None, None) == -1


def test_no_silent_cpu_fallback(lib):
    """Without a CUDA device creation must FAIL with the CUDA error, never fall back to a CPU path."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    cfg = bm.default_config()
    rc = lib.bm_create(C.byref(h), C.byref(cfg))
    assert rc > 0, "expected a cudaError code"
    assert b"cudaSetDevice" in lib.bm_last_error_string()
    with pytest.raises(bm.BrickmapError):
        bm.Renderer(cfg)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under brickmap_b200/ may import, include, link or load it."""
    pat = re.compile(r"(import\s+oracle|from\s+oracle|liboracle|oracle[/\\]|#include\s*[\"<][^\n]*oracle|_ref[/\\]|libbrickmap_ref)")
    for base, _, files in os.walk(os.path.join(ROOT, "brickmap_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                m = pat.search(open(os.path.join(base, f)).read())
                assert m is None, "%s references the oracle: %r" % (f, m.group(0))


def test_strip_partition_covers_the_image_once():
    for h, n, strip in ((1080, 1, 8), (1080, 2, 8), (1080, 8, 8), (1083, 4, 8), (2160, 8, 16)):
        owned = [bm.strip_rows_for_rank(h, r, n, strip) for r in range(n)]
        allrows = np.sort(np.concatenate([rows for _, rows in owned]))
        assert np.array_equal(allrows, np.arange(h))
        for r, (cnt, rows) in enumerate(owned):
            assert cnt == len(rows)
            k = np.arange(cnt)
            assert np.array_equal(rows, ((k // strip) * n + r) * strip + k % strip), "python helper and bm_config.strip_* formula agree"
        counts = [c for c, _ in owned]
        assert max(counts) - min(counts) <= strip


def test_tile_partition():
    for h, n in ((1080, 1), (1080, 2), (1080, 7), (2160, 8)):
        bands = [bm.tile_rows_for_rank(h, r, n) for r in range(n)]
        assert bands[0][0] == 0 and sum(b[1] for b in bands) == h
        for a, b in zip(bands, bands[1:]):
            assert a[0] + a[1] == b[0]
        assert max(b[1] for b in bands) - min(b[1] for b in bands) <= 1
