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

// Device-memory helpers for callers that have no CUDA runtime of their own (the CLIs are plain C++): buffers for
// scn_sens_decode_depth_device / scn_tsdf_integrate_device, a non-blocking stream, a synchronous upload.
void* scn_device_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); scn::fail(SCN_ERR_CUDA, "cudaMalloc(%zu) failed", bytes); return nullptr; }
  return p;
}
void scn_device_free(void* p) { if (p) cudaFree(p); }
int scn_set_device(int device) { SCN_CUDA_TRY(cudaSetDevice(device)); return SCN_OK; }
int scn_device_mem_info(size_t* used_bytes, size_t* total_bytes) {
  size_t fr = 0, tot = 0;
  SCN_CUDA_TRY(cudaMemGetInfo(&fr, &tot));
  if (used_bytes) *used_bytes = tot - fr;
  if (total_bytes) *total_bytes = tot;
  return SCN_OK;
}
int scn_stream_create(void** out) {
  if (!out) return scn::fail(SCN_ERR_ARG, "null argument");
  cudaStream_t s = nullptr;
  SCN_CUDA_TRY(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *out = (void*)s;
  return SCN_OK;
}
void scn_stream_destroy(void* s) { if (s) cudaStreamDestroy((cudaStream_t)s); }
int scn_memcpy_h2d(void* d_dst, const void* src, size_t bytes, void* stream) {
  if (!bytes) return SCN_OK;
  if (!d_dst || !src) return scn::fail(SCN_ERR_ARG, "null argument");
  SCN_CUDA_TRY(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
  SCN_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return SCN_OK;
}
int scn_memcpy_d2h(void* dst, const void* d_src, size_t bytes, void* stream) {
  if (!bytes) return SCN_OK;
  if (!dst || !d_src) return scn::fail(SCN_ERR_ARG, "null argument");
  SCN_CUDA_TRY(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  SCN_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return SCN_OK;
}

// Creates the CUDA context (≈0.3 s on a cold process) — the CLIs call it on a helper thread while they read their input file.
extern "C" void scn_inflate_release_staging_();
extern "C" void scn_jpeg_release_staging_();
extern "C" void scn_fuse_release_staging_();
int scn_current_device_(void) { int d = 0; cudaGetDevice(&d); return d; }
int scn_release_cached_staging(void) { scn_inflate_release_staging_(); scn_jpeg_release_staging_(); scn_fuse_release_staging_(); return SCN_OK; }
int scn_cuda_warmup(void) { return cudaFree(0) == cudaSuccess ? SCN_OK : scn::fail(SCN_ERR_CUDA, "no usable CUDA device"); }

}  // extern "C"
