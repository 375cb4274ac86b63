// Library-wide bits of the C ABI: error text, version, launch counter.
#include <cstdlib>
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

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("MHE_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}

// ---- timing probe ------------------------------------------------------------------------------
static char g_probe_tag[64] = "";
static cudaEvent_t* g_probe_ev = nullptr;   // pairs: [2*i] start, [2*i+1] stop
static int g_probe_cap = 0;
static int g_probe_n = 0;

ProbeScope::ProbeScope(const char* what, cudaStream_t s) : on(false), stream(s) {
    if (g_probe_cap > 0 && g_probe_n < g_probe_cap && strstr(what, g_probe_tag)) {
        on = true;
        cudaEventRecord(g_probe_ev[2 * g_probe_n], stream);
    }
}
ProbeScope::~ProbeScope() {
    if (on) {
        cudaEventRecord(g_probe_ev[2 * g_probe_n + 1], stream);
        ++g_probe_n;
    }
}

}  // namespace mhe

extern "C" {
int mhe_probe_configure(const char* tag, int max_launches) {
    using namespace mhe;
    for (int i = 0; i < 2 * g_probe_cap; ++i) cudaEventDestroy(g_probe_ev[i]);
    delete[] g_probe_ev;
    g_probe_ev = nullptr; g_probe_cap = 0; g_probe_n = 0; g_probe_tag[0] = 0;
    if (!tag || max_launches <= 0) return MHE_OK;
    strncpy(g_probe_tag, tag, sizeof(g_probe_tag) - 1);
    g_probe_ev = new cudaEvent_t[2 * max_launches];
    for (int i = 0; i < 2 * max_launches; ++i)
        if (cudaEventCreate(&g_probe_ev[i]) != cudaSuccess) { set_error("probe: cudaEventCreate failed"); return MHE_ERR_CUDA; }
    g_probe_cap = max_launches;
    return MHE_OK;
}
int mhe_probe_reset(void) { mhe::g_probe_n = 0; return MHE_OK; }
int mhe_probe_read(float* total_ms, int* launches) {
    using namespace mhe;
    float sum = 0.f;
    for (int i = 0; i < g_probe_n; ++i) {
        float ms = 0.f;
        if (cudaEventSynchronize(g_probe_ev[2 * i + 1]) != cudaSuccess || cudaEventElapsedTime(&ms, g_probe_ev[2 * i], g_probe_ev[2 * i + 1]) != cudaSuccess) {
            set_error("probe: event read failed"); return MHE_ERR_CUDA;
        }
        sum += ms;
    }
    if (total_ms) *total_ms = sum;
    if (launches) *launches = g_probe_n;
    return MHE_OK;
}

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
