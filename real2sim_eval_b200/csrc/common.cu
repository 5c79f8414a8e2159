// Error reporting + launch accounting shared by every entry point of libr2s.so.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "r2s_internal.h"

namespace r2s {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace r2s

extern "C" {
const char* r2s_last_error(void) { return r2s::g_err; }
int r2s_version(void) { return 100; }
int64_t r2s_launch_count(void) { return r2s::g_launches.load(std::memory_order_relaxed); }
}
