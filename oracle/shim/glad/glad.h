// Stand-in for glad: no GL entry point is called on the bench path (interop is stubbed in ref_harness.cu).
#pragma once
#include <GL/gl.h>
