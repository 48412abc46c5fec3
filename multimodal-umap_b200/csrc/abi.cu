// abi.cu -- error reporting and device queries for the C ABI in include/mmumap.h
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

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

// ---------------------------------------------------------------- per-device one-time setup
// cudaFuncSetAttribute is a per-DEVICE setting: every launch site that needs it asks here whether this
// (site, current device) pair has been configured yet (one atomic bit per device).
bool first_use_on_device(int site) {
    static unsigned long long done[MMU_SITE_COUNT] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || site < 0 || site >= MMU_SITE_COUNT) {
        cudaGetLastError();
        return true;                      // unknown device: configure every time (cheap, always correct)
    }
    const unsigned long long bit = 1ull << dev;
    return (__atomic_fetch_or(&done[site], bit, __ATOMIC_ACQ_REL) & bit) == 0;
}
// ---------------------------------------------------------------- options (A/B switches of the kernels)
// Read from the environment ONCE, when the library is loaded; mmu_set_option changes them afterwards.
struct OptionEntry {
    const char *name;
    const char *env;
    long long value;
};
static OptionEntry g_options[OPT_COUNT] = {
    {"force_staged", "MMUMAP_FORCE_RUNS", 1},          // 1: staged run-form force kernel, 0: plain loop kernel
    {"knn_cta_pairs", "MMUMAP_KNN_CTA_PAIRS", 1},      // 1: cta_group::2 pairs for long rows, 0: single CTAs
    {"knn_window_mb", "MMUMAP_KNN_WINDOW_MB", -1},     // -1: automatic, 0: one launch, >0: window size for every pair launch
    {"sgd_window_mb", "MMUMAP_SGD_WINDOW_MB", -1},     // -1: automatic (tables beyond L2), 0: never, >0: p+g bytes per tail window
    {"knn_fold_norms", "MMUMAP_KNN_FOLD_NORMS", 1},    // 1: |Y|^2 and -2 folded into the contraction where supported
    {"tail_blocks_per_sm", "MMUMAP_TAIL_BLOCKS_PER_SM", 1},   // grid of the fused multi-GPU epoch tail, blocks per SM
    {"tail_skip_mask", "MMUMAP_TAIL_SKIP_MASK", 0},           // MEASUREMENT ONLY: 1 = push tail without the gradient push, 2 = without the shard step
};
struct OptionInit {
    OptionInit() {
        for (int i = 0; i < OPT_COUNT; ++i) {
            const char *e = getenv(g_options[i].env);
            if (e && *e) g_options[i].value = atoll(e);
        }
    }
};
static OptionInit g_option_init;
long long option(int id) { return __atomic_load_n(&g_options[id].value, __ATOMIC_RELAXED); }

// ---------------------------------------------------------------- which kernel variant a launch site chose
static char g_last_kernel[MMU_SITE_COUNT][160];
void note_kernel(int site, const char *fmt, ...) {
    if (site < 0 || site >= MMU_SITE_COUNT) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_kernel[site], sizeof(g_last_kernel[site]), fmt, ap);
    va_end(ap);
}
static const char *const g_site_names[MMU_SITE_COUNT] = {"knn_candidates", "knn_exact", "eigh_small", "edge_forces",
                                                        "epoch_tail", "block_ops"};

}  // namespace mmu

extern "C" int mmu_set_option(const char *name, int64_t value) {
    using namespace mmu;
    MMU_CHECK_ARG(name, "mmu_set_option: null name");
    for (int i = 0; i < OPT_COUNT; ++i)
        if (strcmp(name, g_options[i].name) == 0) {
            __atomic_store_n(&g_options[i].value, (long long)value, __ATOMIC_RELAXED);
            return MMU_OK;
        }
    set_error("mmu_set_option: unknown option '%s'", name);
    return MMU_ERR_ARG;
}

extern "C" int mmu_get_option(const char *name, int64_t *value) {
    using namespace mmu;
    MMU_CHECK_ARG(name && value, "mmu_get_option: null pointer");
    for (int i = 0; i < OPT_COUNT; ++i)
        if (strcmp(name, g_options[i].name) == 0) {
            *value = (int64_t)option(i);
            return MMU_OK;
        }
    set_error("mmu_get_option: unknown option '%s'", name);
    return MMU_ERR_ARG;
}

extern "C" const char *mmu_last_kernel(const char *site) {
    using namespace mmu;
    if (!site) return "";
    for (int i = 0; i < MMU_SITE_COUNT; ++i)
        if (strcmp(site, g_site_names[i]) == 0) return g_last_kernel[i];
    return "";
}

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
