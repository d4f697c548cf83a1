// ABI plumbing: version, thread-local error string, device check.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "sml_common.cuh"

static thread_local char g_err[512] = "";

void sml_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;
void sml_note_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

int sml_use_tensor_cores() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SML_GEMM");
        v = (e && strcmp(e, "simt") == 0) ? 0 : 1;
    }
    return v;
}

int sml_use_fused_fwd() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SML_FUSED_FWD");
        v = (e && strcmp(e, "0") == 0) ? 0 : 1;
    }
    return v;
}

int sml_use_fused_fc2() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SML_FUSE_FC2");
        v = (e && strcmp(e, "0") == 0) ? 0 : 1;
    }
    return v;
}

int sml_use_fused_loss() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SML_FUSE_LOSS");
        v = (e && strcmp(e, "0") == 0) ? 0 : 1;
    }
    return v;
}

int sml_use_pdl() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("SML_PDL");
        v = (e && strcmp(e, "0") == 0) ? 0 : 1;
    }
    return v;
}

static int g_dev_ok[64];     // per-device cache: 1 = verified sm_100-class
static int g_sm_count[64];

int sml_check_device() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        sml_set_error("cudaGetDevice: %s (no CUDA device: libsml_b200 has no CPU path)", cudaGetErrorString(e));
        return SML_E_CUDA;
    }
    if (dev >= 0 && dev < 64 && g_dev_ok[dev]) return SML_OK;
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) {
        sml_set_error("cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
        return SML_E_CUDA;
    }
    if (major != 10) {
        sml_set_error("device %d has compute capability %d.x; libsml_b200 is built for sm_100a only", dev, major);
        return SML_E_ARCH;
    }
    if (dev >= 0 && dev < 64) g_dev_ok[dev] = 1;
    return SML_OK;
}

extern "C" {

int sml_abi_version(void) { return SML_ABI_VERSION; }

static int g_debug_mask = 0;
int sml_debug_set_mask(int mask) { const int old = g_debug_mask; g_debug_mask = mask; return old; }
int sml_debug_mask(void) { return g_debug_mask; }
static int g_debug_ks[4] = {0, 0, 0, 0};
int sml_debug_set_ksplit(int fc2, int d1, int w2, int w1) { g_debug_ks[0] = fc2; g_debug_ks[1] = d1; g_debug_ks[2] = w2; g_debug_ks[3] = w1; return 0; }
int sml_debug_ksplit(int which) { return (which >= 0 && which < 4) ? g_debug_ks[which] : 0; }

const char *sml_last_error(void) { return g_err; }

int sml_device_check(void) { return sml_check_device(); }

uint64_t sml_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int sml_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (dev >= 0 && dev < 64 && g_sm_count[dev]) return g_sm_count[dev];
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (dev >= 0 && dev < 64) g_sm_count[dev] = n;
    return n;
}

}  // extern "C"
