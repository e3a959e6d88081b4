// host-side helpers shared by the translation units of libfishdiff_b200.so
#pragma once
#include <cuda_runtime.h>
#include "../../include/fishdiff_b200.h"

extern "C" void fd_set_error(const char* fmt, ...);
void fd_count_launch(int n);

// call right after a kernel launch inside an `int`-returning API function
#define FD_LAUNCHED()                                                                          \
  do {                                                                                         \
    fd_count_launch(1);                                                                        \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) {                                                                   \
      fd_set_error("%s:%d kernel launch failed: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return -1;                                                                               \
    }                                                                                          \
  } while (0)
