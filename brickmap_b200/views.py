"""The reference's own benchmark views: PerformanceMeasure's camera tour (performance_measure.h:4-25).

Nine positions, eight angle pairs -- the reference's ninth view reads past the end of `test_angles`
(performance_measure.h:16-25); here it reuses the eighth pair. Views 4-8 (x > 4096) stand outside the stock world: they
exercise the AABB entry path (voxel.cuh:142-155) and the 8x8x8 box LoD (voxel.cuh:212-214).
"""
import math

import numpy as np

POSITIONS = [(512, 512, 300), (840.254, 832.446, 1169.88), (2227.83, 774.886, 204.955), (3326.19, 2055.72, 44.7995), (7134.6, 1262.44, 5531.79),
             (11298.6, 3113.03, 598.019), (10921.4, 4774.14, 267.808), (9961.29, 4508.12, 189.59), (10835.3, 4160.83, 359.992)]
ANGLES = [(-61863.5, -0.501796), (-61864.4, -0.429796), (-61863.9, 0.0622036), (-61864.2, -0.981796), (-61865.2, -0.501796), (-61866.3, -0.141796),
          (-61859.4, 0.0142036), (-61857.2, -0.261796)]


def direction(horizontal, vertical):
    """Camera::update (camera.cpp:48-54): double-precision trigonometry, converted to float, then glm::normalize."""
    d = np.array([math.cos(vertical) * math.sin(horizontal), math.cos(vertical) * math.cos(horizontal), math.sin(vertical)], np.float32)
    return (d * np.float32(1.0 / np.sqrt(np.float32(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])))).astype(np.float32)


def tour():
    """[(position, direction)] of the nine views."""
    return [(tuple(np.float32(v) for v in p), direction(*ANGLES[min(i, len(ANGLES) - 1)])) for i, p in enumerate(POSITIONS)]
