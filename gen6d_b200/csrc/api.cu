// Error reporting, versioning and launch accounting for the C ABI.
#include <atomic>

#include "common.cuh"

namespace g6d {

static thread_local char tls_error[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tls_error, sizeof(tls_error), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace g6d

extern "C" const char* g6d_last_error(void) { return g6d::tls_error; }
extern "C" int g6d_version(void) { return 100; }
extern "C" long long g6d_launch_count(void) { return g6d::g_launches.load(std::memory_order_relaxed); }
