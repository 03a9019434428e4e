"""ctypes binding of include/brickmap_b200.h. There is NO fallback: if the CUDA library is missing, importing
the product fails loudly (build it with `python -m brickmap_b200.build`)."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BRICKMAP_B200_LIB") or os.path.join(HERE, "libbrickmap_b200.so")  # the override is for A/B builds of the same sources


class Config(C.Structure):  # bm_config
    _fields_ = [("device", C.c_int32), ("grid_size", C.c_int32), ("grid_height", C.c_int32), ("lod_distance_2x2x2", C.c_int32),
                ("lod_distance_8x8x8", C.c_int32), ("brick_load_queue_size", C.c_int32), ("ray_queue_buffer_size", C.c_uint32),
                ("screen_width", C.c_uint32), ("screen_height", C.c_uint32), ("tile_row0", C.c_uint32), ("tile_rows", C.c_uint32),
                ("strip_rows", C.c_uint32), ("strip_count", C.c_uint32), ("strip_index", C.c_uint32)]


class Camera(C.Structure):  # bm_camera
    _fields_ = [("position", C.c_float * 3), ("direction", C.c_float * 3), ("up", C.c_float * 3), ("focal_distance", C.c_float),
                ("lens_radius", C.c_float)]


class Counters(C.Structure):  # bm_counters
    _fields_ = [("primary_ray_cnt", C.c_uint32), ("start_position", C.c_uint32), ("shadow_ray_cnt", C.c_uint32), ("frame", C.c_uint32)]


class Stats(C.Structure):  # bm_stats
    _fields_ = [(n, C.c_uint64) for n in ("frames", "extend_rays", "shadow_rays", "terminations", "unoccluded", "cell_steps", "index_reads",
                                          "bricks_entered", "requests", "kernel_launches")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class GpuScene(C.Structure):  # bm_gpu_scene == Scene::GPUScene (Scene.h:9-17)
    _fields_ = [("indices", C.c_void_p), ("bricks", C.c_void_p), ("brick_load_queue", C.c_void_p), ("brick_load_queue_count", C.c_void_p),
                ("bricks_queue", C.c_void_p), ("indices_queue", C.c_void_p)]


# every symbol include/brickmap_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "bm_default_config": (None, [C.POINTER(Config)]),
    "bm_create": (C.c_int, [C.POINTER(_P), C.POINTER(Config)]),
    "bm_destroy": (None, [_P]),
    "bm_last_error_string": (C.c_char_p, []),
    "bm_scene_bind": (C.c_int, [_P, GpuScene]),
    "bm_set_camera": (C.c_int, [_P, C.POINTER(Camera)]),
    "bm_set_sun": (C.c_int, [_P, C.c_float, C.c_float]),
    "bm_get_counters": (C.c_int, [_P, C.POINTER(Counters)]),
    "bm_set_counters": (C.c_int, [_P, C.POINTER(Counters)]),
    "bm_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "bm_reset_stats": (C.c_int, [_P]),
    "bm_launch_frame": (C.c_int, [_P, _P, _P, _P, _P, C.c_uint32]),
    "bm_render": (C.c_int, [_P, _P, C.c_uint32, C.c_uint64, C.c_uint32, C.c_int]),
    "bm_extend_primaries": (C.c_int, [_P, _P, C.c_uint32, C.c_int]),
    "bm_import_rays": (C.c_int, [_P, _P, C.c_uint32]),
    "bm_export_rays": (C.c_int, [_P, _P]),
    "bm_render_to_host": (C.c_int, [_P, _P, C.c_uint32, C.c_uint64, C.c_uint32, _P, _P, _P]),
    "bm_requests_pack": (C.c_int, [_P, _P]),
    "bm_requests_merge": (C.c_int, [_P, _P, C.c_int]),
    "bm_kernel_timing": (C.c_int, [_P, C.c_int]),
    "bm_kernel_time": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "bm_read_requests": (C.c_int, [_P, _P, _P]),
    "bm_stream": (_P, [_P]),
    "bm_synchronize": (C.c_int, [_P]),
    "bm_trace": (C.c_int, [_P, C.c_size_t, _P, _P, _P, _P, _P]),
    "bm_eval_sky": (C.c_int, [_P, C.c_size_t, _P, C.c_int, _P]),
    "bm_tonemap": (C.c_int, [_P, _P, _P]),
    "bm_scene_store_create": (C.c_int, [C.POINTER(_P), C.POINTER(Config), C.c_int, C.c_uint32, C.c_int]),
    "bm_scene_store_create_from_host": (C.c_int, [C.POINTER(_P), C.POINTER(Config), _P, _P, _P, C.c_int, C.c_int]),
    "bm_scene_store_destroy": (None, [_P]),
    "bm_scene_store_gpu_scene": (C.c_int, [_P, C.POINTER(GpuScene)]),
    "bm_scene_store_stream": (C.c_int, [_P, _P, C.POINTER(C.c_uint32)]),
    "bm_scene_store_read_indices": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "bm_scene_store_brick_count": (C.c_int, [_P, C.c_int, C.POINTER(C.c_uint32)]),
    "bm_scene_store_read_bricks": (C.c_int, [_P, C.c_int, C.c_uint32, C.c_uint32, _P]),
    "bm_scene_store_read_gpu_bricks": (C.c_int, [_P, C.c_int, C.c_uint32, C.c_uint32, _P]),
    "bm_scene_store_total_bricks": (C.c_uint64, [_P]),
    "bm_scene_store_last_error": (C.c_char_p, []),
    "bm_scene_store_superchunks": (C.c_int, [_P]),
}

_lib = None


def load():
    """Load libbrickmap_b200.so (once) and type every entry point. Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("brickmap_b200: %s is missing -- the CUDA extension is not built (run `python -m brickmap_b200.build`); "
                          "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class BrickmapError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        lib = load()
        msg = lib.bm_last_error_string().decode() or lib.bm_scene_store_last_error().decode()
        raise BrickmapError("%s failed with code %d: %s" % (what, rc, msg))
