// Error plumbing of the C ABI.
#include "common.cuh"
#include <cstring>
#include <atomic>

namespace d2p {
char* err_buf() {
    static thread_local char buf[1024] = {0};
    return buf;
}
int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 1024, fmt, ap);
    va_end(ap);
    return code;
}
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(); }
}  // namespace d2p

extern "C" const char* d2p_last_error(void) { return d2p::err_buf(); }
extern "C" int d2p_version(void) { return 100; }
// number of kernels this library has launched (or captured into a graph) so far
extern "C" long long d2p_launch_count(void) { return d2p::launches(); }
