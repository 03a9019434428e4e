"""brickmap_b200 -- B200-native (sm_100a) implementation of BrickMap's path-tracing hot path.

Host-side mirror of the reference's interface for this path (src/launch.h:6, src/state.h, src/Scene.h):

    Config        the compile-time constants of variables.h as run-time values
    SceneStore    Scene (Scene.h:7-44): owns the device scene, hands out the GPUScene, streams bricks
    State         State (state.h:3-34): ray queues, shadow queue and accumulation buffer
    Renderer      launch_kernels (launch.h:6) as an object: launch_kernels() is one reference frame,
                  render() the fused multi-frame throughput path

All compute happens in brickmap_b200/libbrickmap_b200.so (hand-written CUDA, csrc/); torch is used only to own
device memory and for torch.distributed plumbing.
"""
from ._lib import BrickmapError, Camera, Config, Counters, GpuScene, Stats, load  # noqa: F401
from .renderer import (RAY_DTYPE, SHADOW_DTYPE, Renderer, SceneStore, State, default_config, make_camera,  # noqa: F401
                       strip_rows_for_rank, tile_rows_for_rank)
