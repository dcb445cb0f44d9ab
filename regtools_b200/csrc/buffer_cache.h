// regtools_b200/csrc/buffer_cache.h — process-wide cache of device and pinned-host buffers.
//
// cudaMalloc / cudaFree / cudaHostAlloc of the multi-hundred-MB buffers this path uses cost tens of
// milliseconds each (and cudaFree occasionally hundreds).  Handles are short-lived — the reference's second
// caller creates one JunctionsExtractor per variant (cis_splice_effects_identifier.cc:288) — so buffers
// are returned to a cache at handle destruction and reused by the next handle of the process.
#pragma once
#include <cstddef>
#include <cuda_runtime.h>

namespace rtjx {

// Same contracts as cudaMalloc / cudaFree / cudaHostAlloc / cudaFreeHost.  A cached buffer is reused when
// its size is within [bytes, 2*bytes + 1 MB]; device buffers are cached per device ordinal.
cudaError_t cached_dev_malloc(void** p, size_t bytes);
cudaError_t cached_dev_free(void* p);
cudaError_t cached_host_alloc(void** p, size_t bytes);
cudaError_t cached_host_free(void* p);
// Drops every cached buffer (tests).
void buffer_cache_trim();

template <class T> inline cudaError_t cached_dev_malloc(T** p, size_t bytes) { return cached_dev_malloc(reinterpret_cast<void**>(p), bytes); }
template <class T> inline cudaError_t cached_host_alloc(T** p, size_t bytes) { return cached_host_alloc(reinterpret_cast<void**>(p), bytes); }

}  // namespace rtjx
