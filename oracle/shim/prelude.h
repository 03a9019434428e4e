// Force-included (-include) in front of every UNMODIFIED reference translation unit so that the
// MSVC-only spellings it uses compile with nvcc/g++ on Linux. Test infrastructure only (oracle/).
//   __forceinline            kernel.cu:76, voxel.cuh:6, Scene.cpp:13
//   std::cosf / std::sinf    kernel.cu:102 (not declared by libstdc++)
//   _pdep_u32                Scene.cpp:26 (needs <immintrin.h>, -mbmi2)
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <memory>
#include <atomic>
#include <immintrin.h>
#ifdef __CUDACC__
#define __forceinline __forceinline__
#else
#define __forceinline inline
#endif
namespace std { using ::cosf; using ::sinf; }
