// C-ABI plumbing shared by every entry point: version, thread-local error text.
#include <cstdarg>
#include <atomic>
#include <cstdio>

#include "common.cuh"

namespace yb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what) {
    set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    return YB_ERR_CUDA;
}

static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace yb

extern "C" int yb_abi_version(void) { return YB_ABI_VERSION; }
extern "C" const char *yb_last_error(void) { return yb::g_err; }
extern "C" long long yb_launch_count(void) { return yb::g_launches.load(std::memory_order_relaxed); }
extern "C" size_t yb_struct_size(int which) {
    switch (which) {
        case 0: return sizeof(yb_tal_params);
        case 1: return sizeof(yb_tal_grid);
        case 2: return sizeof(yb_peer_exchange);
        case 3: return sizeof(yb_gt_source);
        default: return 0;
    }
}
