// Error plumbing and device queries of the C ABI.
#include <stdarg.h>
#include <string.h>
#include <atomic>
#include <map>
#include <mutex>
#include <utility>
#include "common.cuh"

namespace tn {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }
int sm_count() {
    static std::atomic<int> cached[64];
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64 && (n = cached[dev].load(std::memory_order_relaxed)) > 0) return n;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    if (dev >= 0 && dev < 64) cached[dev].store(n, std::memory_order_relaxed);
    return n;
}
cudaError_t ensure_dyn_smem(const void* func, size_t bytes) {
    static std::mutex mu;
    static std::map<std::pair<int, const void*>, size_t> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    size_t& have = done[std::make_pair(dev, func)];
    if (bytes <= have) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}
}  // namespace tn

extern "C" int tn_version(void) { return 100; }
extern "C" const char* tn_last_error(void) { return tn::g_err; }
extern "C" int tn_sm_count(void) { return tn::sm_count(); }
namespace tn { long long launches(); }
extern "C" int64_t tn_launch_count(void) { return (int64_t)tn::launches(); }
