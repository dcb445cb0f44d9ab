// regtools_b200/csrc/buffer_cache.cc — see buffer_cache.h.
#include "buffer_cache.h"

#include <map>
#include <mutex>
#include <unordered_map>

namespace rtjx {
namespace {
struct Cache {
    std::mutex mu;
    std::unordered_map<void*, std::pair<size_t, int>> live;          // ptr -> (bytes, device or -1 for host)
    std::multimap<std::pair<int, size_t>, void*> free_list;          // (device, bytes) -> ptr
    size_t cached_dev = 0, cached_host = 0;
    static constexpr size_t MAX_DEV = 64ull << 30, MAX_HOST = 2ull << 30;   // (a 100M-read run holds ~25 GB of slot buffers and accumulators)
};
Cache& C() { static Cache* c = new Cache(); return *c; }     // intentionally leaked: the driver reclaims at exit

void* take(int dev, size_t bytes) {
    Cache& c = C();
    auto it = c.free_list.lower_bound({dev, bytes});
    if (it == c.free_list.end() || it->first.first != dev || it->first.second > 2 * bytes + (1u << 20)) return nullptr;
    void* p = it->second;
    (dev < 0 ? c.cached_host : c.cached_dev) -= it->first.second;
    c.live[p] = {it->first.second, dev};
    c.free_list.erase(it);
    return p;
}
}  // namespace

cudaError_t cached_dev_malloc(void** p, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (bytes == 0) bytes = 1;
    Cache& c = C();
    {
        std::lock_guard<std::mutex> g(c.mu);
        if (void* q = take(dev, bytes)) { *p = q; return cudaSuccess; }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {                       // out of memory: drop the cache and retry once
        buffer_cache_trim();
        cudaGetLastError();
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess) { std::lock_guard<std::mutex> g(c.mu); c.live[*p] = {bytes, dev}; }
    return e;
}

cudaError_t cached_dev_free(void* p) {
    if (!p) return cudaSuccess;
    Cache& c = C();
    std::unique_lock<std::mutex> g(c.mu);
    auto it = c.live.find(p);
    if (it == c.live.end()) { g.unlock(); return cudaFree(p); }
    const size_t bytes = it->second.first; const int dev = it->second.second;
    c.live.erase(it);
    if (c.cached_dev + bytes > Cache::MAX_DEV) { g.unlock(); return cudaFree(p); }
    c.cached_dev += bytes;
    c.free_list.insert({{dev, bytes}, p});
    return cudaSuccess;
}

cudaError_t cached_host_alloc(void** p, size_t bytes) {
    if (bytes == 0) bytes = 1;
    Cache& c = C();
    {
        std::lock_guard<std::mutex> g(c.mu);
        if (void* q = take(-1, bytes)) { *p = q; return cudaSuccess; }
    }
    cudaError_t e = cudaHostAlloc(p, bytes, cudaHostAllocDefault);
    if (e == cudaSuccess) { std::lock_guard<std::mutex> g(c.mu); c.live[*p] = {bytes, -1}; }
    return e;
}

cudaError_t cached_host_free(void* p) {
    if (!p) return cudaSuccess;
    Cache& c = C();
    std::unique_lock<std::mutex> g(c.mu);
    auto it = c.live.find(p);
    if (it == c.live.end()) { g.unlock(); return cudaFreeHost(p); }
    const size_t bytes = it->second.first;
    c.live.erase(it);
    if (c.cached_host + bytes > Cache::MAX_HOST) { g.unlock(); return cudaFreeHost(p); }
    c.cached_host += bytes;
    c.free_list.insert({{-1, bytes}, p});
    return cudaSuccess;
}

void buffer_cache_trim() {
    Cache& c = C();
    std::multimap<std::pair<int, size_t>, void*> drop;
    { std::lock_guard<std::mutex> g(c.mu); drop.swap(c.free_list); c.cached_dev = c.cached_host = 0; }
    int cur = 0;
    cudaGetDevice(&cur);
    for (auto& kv : drop) {
        if (kv.first.first < 0) cudaFreeHost(kv.second);
        else { cudaSetDevice(kv.first.first); cudaFree(kv.second); }
    }
    cudaSetDevice(cur);
}

}  // namespace rtjx
