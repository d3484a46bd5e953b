// api.cu - library-level entry points: version, error string, device check.
#include "nbpc_common.cuh"

static thread_local std::string g_last_error;

#ifdef NBPC_HOST_EMU
thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
#endif

void nbpc_set_error(const std::string &msg) { g_last_error = msg; }

int nbpc_check_launch(const char *where) {
#ifdef NBPC_HOST_EMU
    (void)where;
    return NBPC_OK;
#else
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        nbpc_set_error(std::string(where) + ": CUDA error: " + cudaGetErrorString(e));
        return NBPC_ELAUNCH;
    }
    return NBPC_OK;
#endif
}

int nbpc_require_sm100() {
#ifdef NBPC_HOST_EMU
    return NBPC_OK;
#else
    static thread_local int cached_dev = -1;
    static thread_local int cached_rc = NBPC_OK;
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        nbpc_set_error("no CUDA device available (libnbpc has no CPU fallback)");
        return NBPC_EARCH;
    }
    if (dev == cached_dev) {
        if (cached_rc != NBPC_OK) nbpc_set_error("current device is not compute capability 10.0 (sm_100)");
        return cached_rc;
    }
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    cached_dev = dev;
    cached_rc = (major == 10 && minor == 0) ? NBPC_OK : NBPC_EARCH;
    if (cached_rc != NBPC_OK)
        nbpc_set_error("current device is compute capability " + std::to_string(major) + "." +
                       std::to_string(minor) + "; libnbpc is built for sm_100a only");
    return cached_rc;
#endif
}

extern "C" {

int nbpc_version(void) { return 100; /* 0.1.0 */ }

const char *nbpc_last_error_string(void) { return g_last_error.c_str(); }

int nbpc_device_check(void) { return nbpc_require_sm100(); }

}  // extern "C"
