// ABI plumbing: version, thread-local error string, device check.
#include <stdarg.h>
#include <string.h>

#include "sml_common.cuh"

static thread_local char g_err[512] = "";

void sml_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sml_check_device() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        sml_set_error("cudaGetDevice: %s (no CUDA device: libsml_b200 has no CPU path)", cudaGetErrorString(e));
        return SML_E_CUDA;
    }
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
    return SML_OK;
}

extern "C" {

int sml_abi_version(void) { return SML_ABI_VERSION; }

const char *sml_last_error(void) { return g_err; }

int sml_device_check(void) { return sml_check_device(); }

int sml_sm_count(void) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    return n;
}

}  // extern "C"
