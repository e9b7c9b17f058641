// Error text, launch counter and ABI version of the msi_b200 C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace msi {

static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace msi

extern "C" int msi_b200_abi_version(void) { return MSI_B200_ABI_VERSION; }
extern "C" const char* msi_last_error(void) { return msi::g_err; }
extern "C" uint64_t msi_launch_count(void) { return msi::g_launches.load(std::memory_order_relaxed); }
