// Makes the handle's device current for the duration of an API call and restores the caller's
// device afterwards (a host with several GPUs must not find its current device changed by a call
// into this library).
#pragma once
#include <cuda_runtime.h>

struct OrbDeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit OrbDeviceGuard(int device) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
  }
  ~OrbDeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
  OrbDeviceGuard(const OrbDeviceGuard&) = delete;
  OrbDeviceGuard& operator=(const OrbDeviceGuard&) = delete;
};
