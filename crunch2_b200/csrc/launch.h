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

#ifdef __CUDACC__
#include <stdlib.h>
#include <thread>
// Host threads waiting for the device.  cudaStreamSynchronize spins on a core; CRN_B200_BLOCKING_SYNC=1 makes every wait record an event created
// with cudaEventBlockingSync and sleep on it instead.  Off by default: measured on 8 ranks sharing a 32-core host (4 cores per rank), where the
// spinning waits looked like the obvious thing to remove, sleeping waits made the step SLOWER (224 ms against 170 ms) -- the wake-up latency of
// ~100 waits per step costs more than the cores they free.
namespace crn {
inline bool blocking_sync_wanted()
{
    static const int v = [] {
        const char* e = getenv("CRN_B200_BLOCKING_SYNC");
        return (e && atoi(e) != 0) ? 1 : 0;
    }();
    return v != 0;
}
inline cudaError_t stream_sync(cudaStream_t s)
{
    if (!blocking_sync_wanted()) return cudaStreamSynchronize(s);
    thread_local cudaEvent_t ev = nullptr;
    thread_local int ev_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!ev || ev_dev != dev) {
        if (ev) cudaEventDestroy(ev);
        if (cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess) { ev = nullptr; (void)cudaGetLastError(); return cudaStreamSynchronize(s); }
        ev_dev = dev;
    }
    const cudaError_t e = cudaEventRecord(ev, s);
    return e != cudaSuccess ? e : cudaEventSynchronize(ev);
}
inline unsigned event_flags() { return blocking_sync_wanted() ? (cudaEventBlockingSync | cudaEventDisableTiming) : cudaEventDisableTiming; }
}  // namespace crn
#define cudaStreamSynchronize(s) crn::stream_sync(s)
#else
namespace crn { inline unsigned event_flags() { return cudaEventDisableTiming; } }
#endif

#define CRN_FULL_MASK 0xffffffffu
