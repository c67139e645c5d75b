// C-ABI plumbing shared by every entry point: thread-local error string, launch counter, version.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace gpb {

static thread_local char t_error[512] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_error, sizeof(t_error), fmt, ap);
    va_end(ap);
}

}  // namespace gpb

extern "C" int gpb_abi_version(void) { return GPB_ABI_VERSION; }
extern "C" const char *gpb_last_error_string(void) { return gpb::t_error; }
extern "C" uint64_t gpb_launch_count(void) { return gpb::g_launches.load(); }
