"""Headless output: what interop.cpp + blit_onto_framebuffer show in a window, written to disk instead."""
import numpy as np


def to_rgb8(tonemapped):
    """float4 image from Renderer.tonemap (rgb / alpha, gamma 1/2.2: kernel.cu:355-362) -> uint8 RGB, rows top to bottom."""
    a = np.asarray(tonemapped, dtype=np.float32)[..., :3]
    return (np.clip(np.nan_to_num(a, nan=0.0), 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)


def write_ppm(path, tonemapped):
    rgb = to_rgb8(tonemapped)
    with open(path, "wb") as f:
        f.write(b"P6\n%d %d\n255\n" % (rgb.shape[1], rgb.shape[0]))
        f.write(rgb.tobytes())
