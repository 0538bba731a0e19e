// Error plumbing + small utility entry points of the C ABI (include/scannet_b200.h).
#include "scn_common.h"

namespace scn {

std::string& last_error_ref() {
  static thread_local std::string err;
  return err;
}

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

}  // namespace scn

extern "C" {

const char* scn_last_error(void) { return scn::last_error_ref().c_str(); }
int scn_version(void) { return 100; }

int scn_device_count(void) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    scn::fail(SCN_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return SCN_ERR_CUDA;
  }
  return n;
}

void* scn_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
    scn::fail(SCN_ERR_CUDA, "cudaHostAlloc(%zu) failed", bytes);
    return nullptr;
  }
  return p;
}
void scn_host_free(void* p) { if (p) cudaFreeHost(p); }
void scn_free(void* p) { free(p); }

// Creates the CUDA context (≈0.3 s on a cold process) — the CLIs call it on a helper thread while they read their input file.
int scn_cuda_warmup(void) { return cudaFree(0) == cudaSuccess ? SCN_OK : scn::fail(SCN_ERR_CUDA, "no usable CUDA device"); }

}  // extern "C"
