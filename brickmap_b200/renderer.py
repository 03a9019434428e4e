"""Host-side mirror of the reference's interface for the path-tracing hot path (see package docstring)."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import Camera, Config, Counters, GpuScene, Stats, check

# RayQueue (variables.h:43-52) and ShadowQueue (variables.h:54-59) as numpy record types
RAY_DTYPE = np.dtype([("origin", "<f4", 3), ("direction", "<f4", 3), ("throughput", "<f4", 3), ("normal", "<f4", 3),
                      ("distance", "<f4"), ("identifier", "<i4"), ("bounces", "<i4"), ("pixel_index", "<u4")])
SHADOW_DTYPE = np.dtype([("origin", "<f4", 3), ("direction", "<f4", 3), ("color", "<f4", 3), ("pixel_index", "<u4")])

FRAME_DEFAULT, FRAME_NO_UPLOAD, FRAME_NO_RESET, FRAME_COUNT_WORK, FRAME_EXACT_PATHS, FRAME_EXTEND_ONLY = 0, 1, 2, 4, 8, 16
SCENE_TERRAIN, SCENE_CAVES, SCENE_NONFLAT = 0, 1, 0x100


def default_config(**overrides):
    """bm_default_config(): the reference's constants (variables.h:7-35,61) at 1920x1080, then keyword overrides."""
    cfg = Config()
    _lib.load().bm_default_config(C.byref(cfg))
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError("bm_config has no field %r" % k)
        setattr(cfg, k, v)
    return cfg


def make_camera(position=(512, 512, 300), direction=(1, 0, 0), up=(0, 0, 1), focal=1.0, lens=0.0):
    """Camera with the reference's defaults (camera.h:4-9)."""
    cam = Camera()
    cam.position[:] = [float(v) for v in position]
    cam.direction[:] = [float(v) for v in direction]
    cam.up[:] = [float(v) for v in up]
    cam.focal_distance = focal
    cam.lens_radius = lens
    return cam


def tile_rows_for_rank(height, rank, world_size):
    """Row band [row0, row0 + rows) of rank `rank` when the image is split into world_size horizontal bands."""
    base, extra = divmod(height, world_size)
    row0 = rank * base + min(rank, extra)
    return row0, base + (1 if rank < extra else 0)


def strip_rows_for_rank(height, rank, world_size, strip=8):
    """Interleaved partition: strips of `strip` rows dealt round-robin to the ranks. Returns (rows_owned, image_rows) where
    image_rows[r] is the image row behind row r of the rank's accumulation buffer (same formula as bm_config.strip_*)."""
    rows = []
    s = rank
    while s * strip < height:
        rows.extend(range(s * strip, min((s + 1) * strip, height)))
        s += world_size
    return len(rows), np.array(rows, dtype=np.int64)


class SceneStore:
    """Device-resident world in the reference's layout (Scene.h:21-31) plus the streaming step (Scene.cpp:200-229)."""

    def __init__(self, cfg, kind=SCENE_TERRAIN, seed=1, resident=True, nonflat=False, _handle=None):
        self.lib = _lib.load()
        self.cfg = cfg
        if _handle is not None:
            self.h = _handle
        else:
            h = C.c_void_p()
            check(self.lib.bm_scene_store_create(C.byref(h), C.byref(cfg), kind | (SCENE_NONFLAT if nonflat else 0), seed, 1 if resident else 0),
                  "bm_scene_store_create")
            self.h = h
        self.gpu_scene = GpuScene()
        check(self.lib.bm_scene_store_gpu_scene(self.h, C.byref(self.gpu_scene)), "bm_scene_store_gpu_scene")

    @classmethod
    def from_host(cls, cfg, indices, brick_counts, bricks, resident=True, nonflat=False):
        """World built elsewhere: indices (superchunks*4096 u32), brick_counts (superchunks u32), bricks (total,16 u32)."""
        lib = _lib.load()
        indices = np.ascontiguousarray(indices, dtype=np.uint32)
        brick_counts = np.ascontiguousarray(brick_counts, dtype=np.uint32)
        bricks = np.ascontiguousarray(bricks, dtype=np.uint32)
        h = C.c_void_p()
        check(lib.bm_scene_store_create_from_host(C.byref(h), C.byref(cfg), indices.ctypes.data, brick_counts.ctypes.data,
                                                  bricks.ctypes.data if bricks.size else None, 1 if resident else 0, 1 if nonflat else 0),
              "bm_scene_store_create_from_host")
        return cls(cfg, _handle=h)

    def close(self):
        if getattr(self, "h", None):
            self.lib.bm_scene_store_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def superchunks(self):
        return int(self.lib.bm_scene_store_superchunks(self.h))

    @property
    def total_bricks(self):
        return int(self.lib.bm_scene_store_total_bricks(self.h))

    def brick_count(self, sc):
        n = C.c_uint32()
        check(self.lib.bm_scene_store_brick_count(self.h, sc, C.byref(n)), "bm_scene_store_brick_count")
        return int(n.value)

    def indices(self, sc, host_view=False):
        out = np.zeros(4096, np.uint32)
        check(self.lib.bm_scene_store_read_indices(self.h, sc, 1 if host_view else 0, out.ctypes.data), "bm_scene_store_read_indices")
        return out

    def bricks(self, sc, gpu_view=False):
        n = self.brick_count(sc)
        out = np.zeros((n, 16), np.uint32)
        fn = self.lib.bm_scene_store_read_gpu_bricks if gpu_view else self.lib.bm_scene_store_read_bricks
        check(fn(self.h, sc, 0, n, out.ctypes.data if n else None), "bm_scene_store_read_bricks")
        return out

    def process_load_queue(self, stream=None, want_count=False):
        """Scene::process_load_queue (Scene.cpp:200-229): stage the requested bricks for the next frame's upload."""
        n = C.c_uint32()
        check(self.lib.bm_scene_store_stream(self.h, stream, C.byref(n) if want_count else None), "bm_scene_store_stream")
        return int(n.value) if want_count else None


class State:
    """State (state.h:3-34): the four device buffers launch_kernels works on, owned by the caller."""

    def __init__(self, cfg, device=None):
        device = device if device is not None else torch.device("cuda", cfg.device)
        n = cfg.ray_queue_buffer_size
        rows = cfg.tile_rows or cfg.screen_height
        self.ray_buffer_work = torch.zeros(n * 16, dtype=torch.float32, device=device)   # RayQueue[n]
        self.ray_buffer_next = torch.zeros(n * 16, dtype=torch.float32, device=device)
        self.shadow_queue_buffer = torch.zeros(n * 10, dtype=torch.float32, device=device)  # ShadowQueue[n]
        self.blit_buffer = torch.zeros(rows, cfg.screen_width, 4, dtype=torch.float32, device=device)  # glm::vec4[w*h]

    def swap(self):
        """main.cpp:146"""
        self.ray_buffer_work, self.ray_buffer_next = self.ray_buffer_next, self.ray_buffer_work

    @staticmethod
    def _records(t, dtype, n):
        return t.cpu().numpy().view(np.uint8)[: n * dtype.itemsize].view(dtype).copy()

    def rays(self, which="work", n=None):
        t = self.ray_buffer_work if which == "work" else self.ray_buffer_next
        return self._records(t, RAY_DTYPE, t.numel() // 16 if n is None else n)

    def shadows(self, n):
        return self._records(self.shadow_queue_buffer, SHADOW_DTYPE, n)

    def write_rays(self, rays, which="work"):
        t = self.ray_buffer_work if which == "work" else self.ray_buffer_next
        src = torch.from_numpy(np.ascontiguousarray(rays, dtype=RAY_DTYPE).view(np.float32).copy())
        t[: src.numel()].copy_(src)


class Renderer:
    """launch_kernels (launch.h:6; kernel.cu:366-439) behind the C ABI of include/brickmap_b200.h."""

    def __init__(self, cfg, scene=None):
        self.lib = _lib.load()
        self.cfg = cfg
        h = C.c_void_p()
        check(self.lib.bm_create(C.byref(h), C.byref(cfg)), "bm_create")
        self.h = h
        self.scene = None
        if scene is not None:
            self.bind(scene)

    def close(self):
        if getattr(self, "h", None):
            self.lib.bm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bind(self, scene):
        """scene: SceneStore or a GpuScene filled with device pointers of a foreign host (e.g. the reference's Scene)."""
        gs = scene.gpu_scene if isinstance(scene, SceneStore) else scene
        check(self.lib.bm_scene_bind(self.h, gs), "bm_scene_bind")
        self.scene = scene

    def set_camera(self, cam):
        check(self.lib.bm_set_camera(self.h, C.byref(cam)), "bm_set_camera")

    def set_sun(self, x, y):
        check(self.lib.bm_set_sun(self.h, x, y), "bm_set_sun")

    @property
    def stream(self):
        return self.lib.bm_stream(self.h)

    def synchronize(self):
        check(self.lib.bm_synchronize(self.h), "bm_synchronize")

    def counters(self):
        c = Counters()
        check(self.lib.bm_get_counters(self.h, C.byref(c)), "bm_get_counters")
        return c

    def set_counters(self, primary_ray_cnt=0, start_position=0, shadow_ray_cnt=0, frame=1):
        c = Counters(primary_ray_cnt, start_position, shadow_ray_cnt, frame)
        check(self.lib.bm_set_counters(self.h, C.byref(c)), "bm_set_counters")

    def stats(self):
        s = Stats()
        check(self.lib.bm_get_stats(self.h, C.byref(s)), "bm_get_stats")
        return s.as_dict()

    def reset_stats(self):
        check(self.lib.bm_reset_stats(self.h), "bm_reset_stats")

    def kernel_timing(self, enable=True):
        check(self.lib.bm_kernel_timing(self.h, 1 if enable else 0), "bm_kernel_timing")

    def kernel_time(self):
        """(summed frame-kernel milliseconds, launches) since the last read; CUDA events on the library's stream."""
        ms, n = C.c_double(), C.c_uint64()
        check(self.lib.bm_kernel_time(self.h, C.byref(ms), C.byref(n)), "bm_kernel_time")
        return float(ms.value), int(n.value)

    def launch_kernels(self, state, flags=FRAME_DEFAULT):
        """One reference frame on the caller's State; the caller swaps the ray buffers afterwards (main.cpp:142-146)."""
        check(self.lib.bm_launch_frame(self.h, state.blit_buffer.data_ptr(), state.ray_buffer_work.data_ptr(), state.ray_buffer_next.data_ptr(),
                                       state.shadow_queue_buffer.data_ptr(), flags), "bm_launch_frame")

    def render(self, blit_buffer, frames, target_paths=0, flags=FRAME_DEFAULT, sync=True):
        """Fused throughput path: `frames` frames (or until target_paths paths finished) into blit_buffer (device tensor)."""
        check(self.lib.bm_render(self.h, blit_buffer.data_ptr(), frames, target_paths, flags, 1 if sync else 0), "bm_render")

    def extend_primaries(self, queue, frames=1, sync=True):
        """Primary rays only (kernel.cu:416-418): post-extend records of all slots into `queue` (device tensor, RayQueue[n])."""
        check(self.lib.bm_extend_primaries(self.h, queue.data_ptr(), frames, 1 if sync else 0), "bm_extend_primaries")

    def import_rays(self, rays):
        """Make `rays` (numpy RAY_DTYPE records) the survivor set the next render() starts from."""
        dev = torch.device("cuda", self.cfg.device)
        t = torch.from_numpy(np.ascontiguousarray(rays, dtype=RAY_DTYPE).view(np.float32).copy()).to(dev)
        torch.cuda.synchronize(dev)
        check(self.lib.bm_import_rays(self.h, t.data_ptr(), len(rays)), "bm_import_rays")
        self.synchronize()

    def export_rays(self):
        """The current private survivor set as dense numpy records, in slot order."""
        dev = torch.device("cuda", self.cfg.device)
        n = self.counters().primary_ray_cnt
        t = torch.zeros(max(n, 1) * 16, dtype=torch.float32, device=dev)
        torch.cuda.synchronize(dev)
        check(self.lib.bm_export_rays(self.h, t.data_ptr()), "bm_export_rays")
        return t.cpu().numpy().view(np.uint8)[: n * RAY_DTYPE.itemsize].view(RAY_DTYPE).copy()

    def render_to_host(self, blit_buffer, frames, accum_host, target_paths=0, flags=FRAME_DEFAULT, request_count_host=None, request_positions_host=None):
        """render() + device->host copies into (pinned) HOST tensors: accumulation tile, request count, request positions."""
        check(self.lib.bm_render_to_host(self.h, blit_buffer.data_ptr(), frames, target_paths, flags, accum_host.data_ptr(),
                                         request_count_host.data_ptr() if request_count_host is not None else None,
                                         request_positions_host.data_ptr() if request_positions_host is not None else None), "bm_render_to_host")

    def trace(self, origins, directions, normals=None, distances=None):
        """intersect_voxel (voxel.cuh:135-261) for n rays given as (n,3) arrays; returns (hit, distance, normal) numpy arrays."""
        dev = torch.device("cuda", self.cfg.device)
        o = torch.as_tensor(np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3), device=dev)
        d = torch.as_tensor(np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 3), device=dev)
        n = o.shape[0]
        nr = torch.zeros(n, 3, dtype=torch.float32, device=dev) if normals is None else torch.as_tensor(np.ascontiguousarray(normals, dtype=np.float32), device=dev).clone()
        di = torch.zeros(n, dtype=torch.float32, device=dev) if distances is None else torch.as_tensor(np.ascontiguousarray(distances, dtype=np.float32), device=dev).clone()
        hit = torch.zeros(n, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev)
        check(self.lib.bm_trace(self.h, n, o.data_ptr(), d.data_ptr(), nr.data_ptr(), di.data_ptr(), hit.data_ptr()), "bm_trace")
        return hit.cpu().numpy().astype(bool), di.cpu().numpy(), nr.cpu().numpy()

    def eval_sky(self, dirs, mode):
        """sun (0) / sky (1) / sunsky (2) of sunsky.cu for (n,3) directions."""
        dev = torch.device("cuda", self.cfg.device)
        d = torch.as_tensor(np.ascontiguousarray(dirs, dtype=np.float32).reshape(-1, 3), device=dev)
        out = torch.zeros_like(d)
        torch.cuda.synchronize(dev)
        check(self.lib.bm_eval_sky(self.h, d.shape[0], d.data_ptr(), mode, out.data_ptr()), "bm_eval_sky")
        return out.cpu().numpy()

    def tonemap(self, blit_buffer):
        out = torch.zeros_like(blit_buffer)
        torch.cuda.synchronize(blit_buffer.device)
        check(self.lib.bm_tonemap(self.h, blit_buffer.data_ptr(), out.data_ptr()), "bm_tonemap")
        return out

    def load_queue(self):
        """(count, positions[min(count, size)]) of the brick request queue (voxel.cuh:228-241)."""
        q = self.cfg.brick_load_queue_size
        cnt = np.zeros(1, np.uint32)
        pos = np.zeros((q, 3), np.int32)
        check(self.lib.bm_read_requests(self.h, cnt.ctypes.data, pos.ctypes.data), "bm_read_requests")
        return int(cnt[0]), pos[: min(int(cnt[0]), q)]
