// Library-wide bits of the C ABI: error text, version, launch counter.
#include "common.cuh"

namespace mhe {

static thread_local char g_error[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

}  // namespace mhe

extern "C" {
const char* mhe_last_error_string(void) { return mhe::g_error; }
int mhe_version(void) { return 100; }
int mhe_built_for_sm(void) {
#ifdef MHE_SM
    return MHE_SM;
#else
    return 0;
#endif
}
long long mhe_kernel_launch_count(void) { return mhe::g_launches.load(); }
}
