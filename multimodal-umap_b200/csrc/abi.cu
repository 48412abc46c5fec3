// abi.cu -- error reporting and device queries for the C ABI in include/mmumap.h
#include <stdarg.h>

#include "common.cuh"

namespace mmu {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }

int sm_count() {
    static int cached = -1;
    if (cached < 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
            cached = n;
        else {
            cudaGetLastError();
            return 0;
        }
    }
    return cached;
}

}  // namespace mmu

extern "C" int mmu_abi_version(void) { return MMU_ABI_VERSION; }

extern "C" const char *mmu_last_error(void) { return mmu::g_err; }

extern "C" void mmu_launch_count_add(uint64_t n) { mmu::count_launch((int)n); }

extern "C" uint64_t mmu_launch_count(void) { return __atomic_load_n(&mmu::g_launches, __ATOMIC_RELAXED); }

extern "C" int mmu_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *l2_bytes) {
    using namespace mmu;
    int dev = 0;
    MMU_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    MMU_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (l2_bytes) *l2_bytes = (size_t)prop.l2CacheSize;
    return MMU_OK;
}
