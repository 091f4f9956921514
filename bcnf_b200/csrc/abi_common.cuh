// Helpers shared by the translation units of the C ABI (bcnf_abi.cu, trf_abi.cu): error reporting, NVTX ranges, the
// device guard of every entry point.
#pragma once
#include "../../include/bcnf_b200.h"

#include <cstdarg>
#include <cstdio>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

// message of the calling thread's last failed call (defined in bcnf_abi.cu; read by bcnf_last_error)
int bcnf_fail(int code, const char* fmt, ...);
#define fail bcnf_fail

// NVTX range around each data-path entry point (header-only nvtx3: a no-op unless a profiler injects itself), so a
// timeline shows the C-ABI calls by name above the kernels they launch.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};
#define NVTX_RANGE(name) NvtxRange nvtx_range_(name)

// Entry points run on the handle's (or the caller-named) device and restore the caller's current device on the way out:
// torch keeps its own notion of the current device, and bcnf_flow_destroy is reached from Python's garbage collector.
struct DeviceGuard {
  int prev = -1;
  bool changed = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) { err = cudaSetDevice(dev); changed = err == cudaSuccess; }
  }
  ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define DEVICE_GUARD(dev)                                                                              \
  DeviceGuard guard__(dev);                                                                            \
  if (guard__.err != cudaSuccess) return fail((int)guard__.err, "cudaSetDevice(%d): %s", (int)(dev), cudaGetErrorString(guard__.err))

#define CUDA_TRY(expr)                                                                      \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) return fail((int)e__, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)
