// Shared host-side helpers for libscannet_b200.so: thread-local error text, CUDA status checks.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/scannet_b200.h"

namespace scn {

std::string& last_error_ref();
int fail(int code, const char* fmt, ...);

}  // namespace scn

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define SCN_CUDA_TRY(expr)                                                                   \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return scn::fail(SCN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                       __FILE__, __LINE__);                                                  \
  } while (0)
#endif
