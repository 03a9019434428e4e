"""ctypes bindings for the parity oracles. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import this module;
nothing under brickmap_b200/ does.

  Oracle      -> oracle/liboracle.so          CPU restatement (oracle.cpp)
  Reference   -> oracle/_ref/libbrickmap_ref_<variant>.so   the unmodified reference kernels + harness (GPU only)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("direction", "<f4", 3), ("throughput", "<f4", 3), ("normal", "<f4", 3),
                      ("distance", "<f4"), ("identifier", "<i4"), ("bounces", "<i4"), ("pixel_index", "<u4")])
SHADOW_DTYPE = np.dtype([("origin", "<f4", 3), ("direction", "<f4", 3), ("color", "<f4", 3), ("pixel_index", "<u4")])
assert RAY_DTYPE.itemsize == 64 and SHADOW_DTYPE.itemsize == 40


class Camera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("direction", C.c_float * 3), ("up", C.c_float * 3),
                ("focal_distance", C.c_float), ("lens_radius", C.c_float)]


class FrameState(C.Structure):
    _fields_ = [("primary_ray_cnt", C.c_uint32), ("start_position", C.c_uint32), ("shadow_ray_cnt", C.c_uint32), ("frame", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "index_reads", "bricks", "lod_bytes", "voxel_steps", "requests", "hits",
                                          "terminations", "unoccluded", "unique_index_sectors", "unique_brick_sectors")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def make_camera(position=(512, 512, 300), direction=(1, 0, 0), up=(0, 0, 1), focal=1.0, lens=0.0):
    cam = Camera()
    cam.position[:] = [float(v) for v in position]
    cam.direction[:] = [float(v) for v in direction]
    cam.up[:] = [float(v) for v in up]
    cam.focal_distance = focal
    cam.lens_radius = lens
    return cam


def build_oracle():
    """(Re)build oracle/liboracle.so with the committed recipe."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


class Oracle:
    """CPU restatement of the reference hot path (see oracle/oracle.h)."""

    def __init__(self, path=None):
        path = path or os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = self.lib = C.CDLL(path)
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_create.argtypes = [C.c_int] * 5
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        for n in ("orc_scene_generate_terrain",):
            getattr(L, n).argtypes = [C.c_void_p, C.c_int]
        L.orc_scene_generate_caves.argtypes = [C.c_void_p, C.c_uint32, C.c_int]
        L.orc_scene_from_voxels.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_scene_set_residency.argtypes = [C.c_void_p, C.c_int]
        L.orc_scene_supergrid_count.argtypes = [C.c_void_p]
        L.orc_scene_brick_count.argtypes = [C.c_void_p, C.c_int]
        for n in ("orc_scene_host_indices", "orc_scene_host_bricks", "orc_scene_gpu_indices"):
            getattr(L, n).restype = C.POINTER(C.c_uint32)
            getattr(L, n).argtypes = [C.c_void_p, C.c_int]
        L.orc_scene_gpu_brick.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_scene_queue_count.restype = C.c_uint32
        L.orc_scene_queue_count.argtypes = [C.c_void_p]
        L.orc_scene_queue_positions.restype = C.POINTER(C.c_int32)
        L.orc_scene_queue_positions.argtypes = [C.c_void_p]
        L.orc_scene_stream.argtypes = [C.c_void_p]
        L.orc_trace.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_camera_basis.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.orc_sun_direction.argtypes = [C.c_float, C.c_float, C.c_void_p]
        L.orc_primary_rays.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
        L.orc_set_wavefront_globals.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        L.orc_extend.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_shade.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_connect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_primary_rays_tiled.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        L.orc_frame_tiled.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                      C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_sky_eval.argtypes = [C.c_size_t, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_cone_sample.argtypes = [C.c_void_p, C.c_float, C.c_void_p, C.c_void_p]
        L.orc_footprint_begin.argtypes = [C.c_void_p]
        L.orc_footprint_report.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_tonemap.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]

    def hardware_threads(self):
        return int(self.lib.orc_hardware_threads())


class OracleScene:
    def __init__(self, oracle, grid_xy, grid_z, lod2=100000, lod8=600000, queue_size=1024):
        self.o = oracle
        self.L = oracle.lib
        self.grid_xy, self.grid_z = grid_xy, grid_z
        self.lod2, self.lod8, self.queue_size = lod2, lod8, queue_size
        self.h = self.L.orc_scene_create(grid_xy, grid_z, lod2, lod8, queue_size)
        if not self.h:
            raise ValueError("bad scene dimensions")

    def __del__(self):
        try:
            if self.h:
                self.L.orc_scene_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # generation -------------------------------------------------------------------------------
    def generate_terrain(self, threads=0):
        self.L.orc_scene_generate_terrain(self.h, threads)
        return self

    def generate_caves(self, seed=1, threads=0):
        self.L.orc_scene_generate_caves(self.h, seed, threads)
        return self

    def from_voxels(self, vox):
        vox = np.ascontiguousarray(vox, dtype=np.uint8)  # indexed [z, y, x]
        assert vox.shape == (self.grid_z, self.grid_xy, self.grid_xy)
        self.L.orc_scene_from_voxels(self.h, vox.ctypes.data)
        return self

    def set_residency(self, all_resident):
        self.L.orc_scene_set_residency(self.h, 1 if all_resident else 0)
        return self

    # accessors ---------------------------------------------------------------------------------
    @property
    def supergrid_count(self):
        return self.L.orc_scene_supergrid_count(self.h)

    def brick_count(self, sc):
        return self.L.orc_scene_brick_count(self.h, sc)

    def host_indices(self, sc):
        return np.ctypeslib.as_array(self.L.orc_scene_host_indices(self.h, sc), shape=(4096,)).copy()

    def gpu_indices(self, sc):
        return np.ctypeslib.as_array(self.L.orc_scene_gpu_indices(self.h, sc), shape=(4096,)).copy()

    def host_bricks(self, sc):
        n = self.brick_count(sc)
        if n == 0:
            return np.zeros((0, 16), dtype=np.uint32)
        return np.ctypeslib.as_array(self.L.orc_scene_host_bricks(self.h, sc), shape=(n, 16)).copy()

    def gpu_brick(self, sc, slot):
        out = np.zeros(16, dtype=np.uint32)
        if self.L.orc_scene_gpu_brick(self.h, sc, slot, out.ctypes.data) != 0:
            raise IndexError("slot not resident")
        return out

    def all_host_indices(self):
        return np.concatenate([self.host_indices(i) for i in range(self.supergrid_count)])

    def all_gpu_indices(self):
        return np.concatenate([self.gpu_indices(i) for i in range(self.supergrid_count)])

    def queue(self):
        cnt = int(self.L.orc_scene_queue_count(self.h))
        n = min(cnt, self.queue_size)
        pos = np.ctypeslib.as_array(self.L.orc_scene_queue_positions(self.h), shape=(self.queue_size, 3))[:n].copy()
        return cnt, pos

    def stream(self):
        return self.L.orc_scene_stream(self.h)

    # traversal ----------------------------------------------------------------------------------
    def trace(self, origins, directions, cam_cell, normals=None, distances=None, threads=1, stats=None):
        o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 3)
        n = o.shape[0]
        nrm = np.zeros((n, 3), np.float32) if normals is None else np.ascontiguousarray(normals, dtype=np.float32).copy()
        dist = np.zeros(n, np.float32) if distances is None else np.ascontiguousarray(distances, dtype=np.float32).copy()
        hit = np.zeros(n, np.uint8)
        cc = np.asarray(cam_cell, dtype=np.int32)
        self.L.orc_trace(self.h, n, o.ctypes.data, d.ctypes.data, cc.ctypes.data, nrm.ctypes.data, dist.ctypes.data, hit.ctypes.data,
                         C.addressof(stats) if stats is not None else None, threads)
        return hit.astype(bool), dist, nrm

    def footprint_begin(self):
        self.L.orc_footprint_begin(self.h)

    def footprint_report(self, stats):
        self.L.orc_footprint_report(self.h, C.addressof(stats))


class OracleRenderer:
    """Canonical (slot-index ordered) wavefront loop of the reference, on the CPU."""

    def __init__(self, scene, width, height, n_slots, camera, sun=(0.05, 0.1), tile=None):
        """tile = (row0, rows, strip_rows, strip_count, strip_index): the multi-GPU image partition of bm_config (oracle.h);
        None = the whole image, the reference's own mapping."""
        self.scene, self.L = scene, scene.L
        self.width, self.height, self.n_slots = width, height, n_slots
        self.camera, self.sun = camera, sun
        self.tile = None if tile is None else np.array(tile, np.uint32)
        self.rays = np.zeros(n_slots, RAY_DTYPE)
        self.next = np.zeros(n_slots, RAY_DTYPE)
        self.shadows = np.zeros(n_slots, SHADOW_DTYPE)
        self.accum = np.zeros((height if tile is None else int(tile[1]), width, 4), np.float32)
        self.state = FrameState(0, 0, 0, 1)
        self.stats = Stats()

    def sun_direction(self):
        out = np.zeros(3, np.float32)
        self.L.orc_sun_direction(self.sun[0], self.sun[1], out.ctypes.data)
        return out

    def camera_basis(self):
        r, u = np.zeros(3, np.float32), np.zeros(3, np.float32)
        self.L.orc_camera_basis(C.addressof(self.camera), self.width, self.height, r.ctypes.data, u.ctypes.data)
        return r, u

    def primary_rays(self):
        if self.tile is not None:
            self.L.orc_primary_rays_tiled(self.rays.ctypes.data, self.n_slots, C.addressof(self.state), C.addressof(self.camera), self.width, self.height,
                                          self.tile.ctypes.data)
        else:
            self.L.orc_primary_rays(self.rays.ctypes.data, self.n_slots, C.addressof(self.state), C.addressof(self.camera), self.width, self.height)

    def set_wavefront_globals(self):
        self.L.orc_set_wavefront_globals(C.addressof(self.state), self.n_slots, self.width, self.height if self.tile is None else int(self.tile[1]))

    def extend(self, threads=0):
        self.L.orc_extend(self.scene.h, self.rays.ctypes.data, self.n_slots, C.addressof(self.camera), C.addressof(self.stats), threads)

    def shade(self):
        sd = self.sun_direction()
        self.L.orc_shade(self.rays.ctypes.data, self.next.ctypes.data, self.shadows.ctypes.data, self.n_slots, C.addressof(self.state),
                         sd.ctypes.data, self.accum.ctypes.data, C.addressof(self.stats))

    def connect(self, threads=0):
        self.L.orc_connect(self.scene.h, self.shadows.ctypes.data, C.addressof(self.state), C.addressof(self.camera), self.accum.ctypes.data,
                           C.addressof(self.stats), threads)

    def frame(self, threads=0):
        """kernel.cu:416-423 then main.cpp:146 (buffer swap)."""
        if self.tile is not None:
            self.L.orc_frame_tiled(self.scene.h, self.rays.ctypes.data, self.next.ctypes.data, self.shadows.ctypes.data, self.n_slots, C.addressof(self.state),
                                   C.addressof(self.camera), self.sun[0], self.sun[1], self.width, self.height, self.tile.ctypes.data, self.accum.ctypes.data,
                                   C.addressof(self.stats), threads)
            self.rays, self.next = self.next, self.rays
            return
        self.L.orc_frame(self.scene.h, self.rays.ctypes.data, self.next.ctypes.data, self.shadows.ctypes.data, self.n_slots, C.addressof(self.state),
                         C.addressof(self.camera), self.sun[0], self.sun[1], self.width, self.height, self.accum.ctypes.data,
                         C.addressof(self.stats), threads)
        self.rays, self.next = self.next, self.rays

    def reset(self):
        """kernel.cu:397-403: zero the accumulation buffer and primary_ray_cnt (not start_position, not frame)."""
        self.accum[...] = 0
        self.state.primary_ray_cnt = 0


def sky_eval(oracle, dirs, mode, sun_dir):
    d = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
    out = np.zeros_like(d)
    sd = np.ascontiguousarray(sun_dir, dtype=np.float32)
    oracle.lib.orc_sky_eval(d.shape[0], d.ctypes.data, mode, sd.ctypes.data, out.ctypes.data)
    return out


# ------------------------------------------------------------------------------------------------------
class Reference:
    """The unmodified reference kernels behind oracle/ref_harness.cu. Needs a GPU and a prebuilt oracle/_ref library."""

    VARIANTS = ("4096", "256", "256lod")
    # "dropin_256": the reference HOST code (Scene.cpp, State, main-loop body) linked against integration/launch_kernels_dropin.cpp
    # and libbrickmap_b200.so instead of kernel.cu/sunsky.cu; only the host-level entry points exist in that library.

    @staticmethod
    def path(variant):
        if variant.startswith("dropin_"):
            return os.path.join(HERE, "_ref", "libbrickmap_%s.so" % variant)
        return os.path.join(HERE, "_ref", "libbrickmap_ref_%s.so" % variant)

    @classmethod
    def available(cls, variant):
        return os.path.exists(cls.path(variant))

    def __init__(self, variant, width, height, device=0):
        L = self.lib = C.CDLL(self.path(variant))
        self.width, self.height = width, height
        L.ref_constants.argtypes = [C.c_void_p]
        c = (C.c_int64 * 9)()
        L.ref_constants(c)
        (self.grid_size, self.grid_height, self.n_slots, self.lod2, self.lod8, self.queue_size, self.start_size, rq, sq) = [int(v) for v in c]
        assert rq == 64 and sq == 40
        L.ref_read_rays.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t]
        L.ref_write_rays.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t]
        L.ref_read_shadow.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
        L.ref_write_shadow.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
        L.ref_set_camera.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_float]
        L.ref_set_sun.argtypes = [C.c_float, C.c_float]
        L.ref_run_frames.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_host_supercell.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
        if not variant.startswith("dropin_"):
            L.ref_run_stage.argtypes = [C.c_int, C.c_int, C.c_uint, C.c_int]
            if hasattr(L, "ref_run_primary_extend"):
                L.ref_run_primary_extend.argtypes = [C.c_int, C.c_uint, C.c_void_p]
            L.ref_eval_sky.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
            L.ref_read_counters.argtypes = [C.c_void_p]
            L.ref_write_counters.argtypes = [C.c_void_p]
        for name in ("ref_read_accum", "ref_sun_direction", "ref_read_indices", "ref_get_scene", "ref_get_state", "ref_host_brick_counts", "ref_alpha_sum",
                     "ref_constants"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.ref_read_load_queue.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_init.argtypes = [C.c_int, C.c_int, C.c_int]
        L.ref_frame.argtypes = [C.c_int]
        self._chk(L.ref_init(device, width, height))

    @staticmethod
    def _chk(rc):
        if rc != 0:
            raise RuntimeError("reference harness returned %d" % rc)

    def generate(self):
        self._chk(self.lib.ref_generate())

    def force_resident(self):
        self._chk(self.lib.ref_force_resident())

    def supergrid_count(self):
        return int(self.lib.ref_supergrid_count())

    def host_supercell(self, sc):
        n = self.supergrid_count()
        counts = (C.c_int * n)()
        self.lib.ref_host_brick_counts(counts)
        idx = np.zeros(4096, np.uint32)
        br = np.zeros((max(1, counts[sc]), 16), np.uint32)
        self.lib.ref_host_supercell(sc, idx.ctypes.data, br.ctypes.data)
        return idx, br[:counts[sc]]

    def scene_pointers(self):
        p = (C.c_void_p * 6)()
        self.lib.ref_get_scene(p)
        return [int(v or 0) for v in p]

    def state_pointers(self):
        p = (C.c_void_p * 4)()
        self.lib.ref_get_state(p)
        return [int(v or 0) for v in p]

    def set_camera(self, cam):
        pos = np.array(list(cam.position), np.float32)
        d = np.array(list(cam.direction), np.float32)
        up = np.array(list(cam.up), np.float32)
        self.lib.ref_set_camera(pos.ctypes.data, d.ctypes.data, up.ctypes.data, cam.focal_distance, cam.lens_radius)

    def set_sun(self, x, y):
        self.lib.ref_set_sun(x, y)

    def upload_sun(self):
        self._chk(self.lib.ref_upload_sun())

    def sun_direction(self):
        out = np.zeros(3, np.float32)
        self.lib.ref_sun_direction(out.ctypes.data)
        return out

    def frame(self, process_queue=True):
        self._chk(self.lib.ref_frame(1 if process_queue else 0))

    def run_frames(self, frames, process_queue=False):
        ms = C.c_float()
        shadows = C.c_uint64()
        self._chk(self.lib.ref_run_frames(frames, 1 if process_queue else 0, C.byref(ms), C.byref(shadows)))
        return float(ms.value), int(shadows.value)

    def run_primary_extend(self, frames, first_frame=1):
        """primary_rays + set_wavefront_globals + extend (kernel.cu:416-418), `frames` times; returns device milliseconds."""
        ms = C.c_float()
        self._chk(self.lib.ref_run_primary_extend(frames, first_frame, C.byref(ms)))
        return float(ms.value)

    def counters(self):
        out = np.zeros(7, np.uint32)
        self._chk(self.lib.ref_read_counters(out.ctypes.data))
        return dict(zip(("primary_ray_cnt", "start_position", "raynr_primary", "raynr_extend", "raynr_shade", "raynr_connect", "shadow_ray_cnt"),
                        [int(v) for v in out]))

    def write_counters(self, **kw):
        cur = self.counters()
        cur.update(kw)
        arr = np.array([cur[k] for k in ("primary_ray_cnt", "start_position", "raynr_primary", "raynr_extend", "raynr_shade", "raynr_connect",
                                          "shadow_ray_cnt")], np.uint32)
        self._chk(self.lib.ref_write_counters(arr.ctypes.data))

    def read_rays(self, which=0, first=0, n=None):
        n = self.n_slots - first if n is None else n
        out = np.zeros(n, RAY_DTYPE)
        self._chk(self.lib.ref_read_rays(which, out.ctypes.data, first, n))
        return out

    def write_rays(self, rays, which=0, first=0):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        self._chk(self.lib.ref_write_rays(which, rays.ctypes.data, first, rays.shape[0]))

    def read_shadows(self, n, first=0):
        out = np.zeros(n, SHADOW_DTYPE)
        if n:
            self._chk(self.lib.ref_read_shadow(out.ctypes.data, first, n))
        return out

    def write_shadows(self, sh, first=0):
        sh = np.ascontiguousarray(sh, dtype=SHADOW_DTYPE)
        self._chk(self.lib.ref_write_shadow(sh.ctypes.data, first, sh.shape[0]))

    def read_accum(self):
        out = np.zeros((self.height, self.width, 4), np.float32)
        self._chk(self.lib.ref_read_accum(out.ctypes.data))
        return out

    def alpha_sum(self):
        v = C.c_double()
        self._chk(self.lib.ref_alpha_sum(C.byref(v)))
        return float(v.value)

    def mark_sun_changed(self):
        """sun_position_changed = true: the next launch_kernels resets the accumulation (kernel.cu:389-403)."""
        self.lib.ref_mark_sun_changed()

    def clear_accum(self):
        self._chk(self.lib.ref_clear_accum())

    def swap_buffers(self):
        self.lib.ref_swap_buffers()

    STAGES = {"primary_rays": 0, "set_wavefront_globals": 1, "extend": 2, "shade": 3, "connect": 4, "upload": 5}

    def run_stage(self, name, frame=1, serial=False, upload_count=0):
        self._chk(self.lib.ref_run_stage(self.STAGES[name], 1 if serial else 0, frame, upload_count))

    def load_queue(self):
        cnt = C.c_uint32()
        pos = np.zeros((self.queue_size, 3), np.int32)
        self._chk(self.lib.ref_read_load_queue(C.byref(cnt), pos.ctypes.data))
        return int(cnt.value), pos[:min(int(cnt.value), self.queue_size)]

    def process_load_queue(self):
        self._chk(self.lib.ref_process_load_queue())

    def read_indices(self):
        out = np.zeros(self.supergrid_count() * 4096, np.uint32)
        self._chk(self.lib.ref_read_indices(out.ctypes.data))
        return out

    def eval_sky(self, dirs, mode):
        d = np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3)
        out = np.zeros_like(d)
        self._chk(self.lib.ref_eval_sky(d.shape[0], d.ctypes.data, mode, out.ctypes.data))
        return out
