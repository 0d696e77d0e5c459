// launch.h -- kernel-launch and dynamic-shared-memory spelling shared by the nvcc build (the
// product) and the g++ SIMT-emulation build used by the CPU tests (tests/cusim/cusim.h).
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define CRN_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define CRN_DYN_SMEM(type, name)                                   \
    extern __shared__ __align__(16) unsigned char name##_raw_[];   \
    type* name = reinterpret_cast<type*>(name##_raw_)
#else
#include "cuda_runtime.h"  // tests/cusim/cuda_runtime.h (emulation build only)
#endif
#include <stdint.h>

#define CRN_FULL_MASK 0xffffffffu
