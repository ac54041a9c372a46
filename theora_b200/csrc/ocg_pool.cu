/* Process-wide cache of page-locked host blocks and device blocks.
 *
 * A decoder or encoder instance owns ~60-80 MB of page-locked staging memory and ~100-200 MB of device
 * memory in a few dozen blocks.  Page-locking is slow (the kernel pins page by page under a process-wide
 * lock: measured ~0.3 s per encoder instance when 16 threads set theirs up at once) and cudaMalloc /
 * cudaFree serialise on the driver's lock and, for cudaFree, synchronise the device.  Applications that
 * open and close many streams (a transcoding server; the benchmark's repeated passes) therefore get their
 * blocks from here: a freed block is kept, keyed by (kind, device, size, flags), and handed to the next
 * request of exactly that key -- instances of one geometry ask for identical sizes.  Reused blocks are
 * cleared, so a block from the cache is indistinguishable from a fresh one that happened to be zero.
 * The cache is bounded (OCG_POOL_HOST_MB / OCG_POOL_DEVICE_MB, defaults 4096 / 16384 per process and
 * device); beyond the bound blocks are really freed.  Blocks are never returned to the driver at exit: the
 * driver is usually gone by the time static destructors run. */
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <unordered_map>
#include <vector>
#include <cuda_runtime.h>

namespace {

struct Key {
  int kind, device;
  size_t size;
  unsigned flags;
  bool operator<(const Key &o) const { return std::tie(kind, device, size, flags) < std::tie(o.kind, o.device, o.size, o.flags); }
};

std::mutex g_lock;
std::map<Key, std::vector<void *>> g_free;
std::unordered_map<void *, Key> g_live;
size_t g_cached[2] = {0, 0}; /* bytes held in g_free: host, device (all devices) */

size_t limit_bytes(int kind) {
  static size_t lim[2] = {0, 0};
  if (lim[kind] == 0) {
    const char *e = getenv(kind == 0 ? "OCG_POOL_HOST_MB" : "OCG_POOL_DEVICE_MB");
    const long mb = e != nullptr ? atol(e) : (kind == 0 ? 4096 : 16384);
    lim[kind] = mb <= 0 ? 1 : (size_t)mb << 20;
  }
  return lim[kind];
}

void *take(const Key &k) {
  std::lock_guard<std::mutex> lk(g_lock);
  auto it = g_free.find(k);
  if (it == g_free.end() || it->second.empty()) return nullptr;
  void *p = it->second.back();
  it->second.pop_back();
  g_cached[k.kind] -= k.size;
  g_live[p] = k;
  return p;
}

/* true: the block went into the cache; false: the caller frees it for real */
bool give(void *p, Key &k_out) {
  std::lock_guard<std::mutex> lk(g_lock);
  auto it = g_live.find(p);
  if (it == g_live.end()) return false; /* not ours */
  k_out = it->second;
  g_live.erase(it);
  if (g_cached[k_out.kind] + k_out.size > limit_bytes(k_out.kind)) return false;
  g_free[k_out].push_back(p);
  g_cached[k_out.kind] += k_out.size;
  return true;
}

void remember(void *p, const Key &k) {
  std::lock_guard<std::mutex> lk(g_lock);
  g_live[p] = k;
}

} /* namespace */

cudaError_t ocg_pool_host_alloc(void **pp, size_t size, unsigned flags) {
  if (pp == nullptr) return cudaErrorInvalidValue;
  const Key k{0, 0, size, flags};
  void *p = size >= 4096 ? take(k) : nullptr;
  if (p != nullptr) {
    memset(p, 0, size);
    *pp = p;
    return cudaSuccess;
  }
  const cudaError_t e = cudaHostAlloc(pp, size, flags);
  if (e == cudaSuccess && size >= 4096) remember(*pp, k);
  return e;
}

cudaError_t ocg_pool_host_free(void *p) {
  if (p == nullptr) return cudaSuccess;
  Key k;
  if (give(p, k)) return cudaSuccess;
  return cudaFreeHost(p);
}

cudaError_t ocg_pool_dev_alloc(void **pp, size_t size) {
  if (pp == nullptr) return cudaErrorInvalidValue;
  int dev = 0;
  cudaGetDevice(&dev);
  const Key k{1, dev, size, 0};
  void *p = size >= 4096 ? take(k) : nullptr;
  if (p != nullptr) {
    /* cleared on a stream of this thread's own, waited for: no other stream is involved */
    static thread_local cudaStream_t st[64];
    cudaError_t e = cudaSuccess;
    if (dev >= 0 && dev < 64 && st[dev] == nullptr) e = cudaStreamCreateWithFlags(&st[dev], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMemsetAsync(p, 0, size, dev >= 0 && dev < 64 ? st[dev] : nullptr);
    if (e == cudaSuccess) e = cudaStreamSynchronize(dev >= 0 && dev < 64 ? st[dev] : nullptr);
    if (e != cudaSuccess) { cudaGetLastError(); Key kk; give(p, kk); return e; }
    *pp = p;
    return cudaSuccess;
  }
  const cudaError_t e = cudaMalloc(pp, size);
  if (e == cudaSuccess && size >= 4096) remember(*pp, k);
  return e;
}

cudaError_t ocg_pool_dev_free(void *p) {
  if (p == nullptr) return cudaSuccess;
  Key k;
  if (give(p, k)) return cudaSuccess; /* the owner has drained its stream before destroying itself */
  return cudaFree(p);
}
