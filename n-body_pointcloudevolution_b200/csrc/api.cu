// api.cu - library-level entry points: version, error string, device check.
#include <stdlib.h>

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

// ------------------------------------------------------------------ launch counter + event profiler
unsigned long long g_nbpc_launches = 0;
int g_nbpc_prof_on = 0;

#ifndef NBPC_HOST_EMU
#include <map>
#include <mutex>
#include <vector>
namespace {
struct ProfRec { std::string name; cudaEvent_t a, b; };
std::vector<ProfRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;
std::mutex g_prof_mu;
cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
}  // namespace
void nbpc_prof_pre(const char *name, cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    ProfRec r;
    r.name = name; r.a = prof_event(); r.b = prof_event();
    cudaEventRecord(r.a, stream);
    g_prof_recs.push_back(r);
}
void nbpc_prof_post(cudaStream_t stream) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_recs.empty()) cudaEventRecord(g_prof_recs.back().b, stream);
}
#endif

// NBPC_MATH=fp32|tf32|tf32x3 selects the default; nbpc_set_math_mode overrides it
static int nbpc_math_mode_from_env() {
    const char *e = getenv("NBPC_MATH");
    if (!e) return NBPC_MATH_DEFAULT;
    if (!strcmp(e, "fp32")) return NBPC_MATH_FP32;
    if (!strcmp(e, "tf32")) return NBPC_MATH_TF32;
    if (!strcmp(e, "tf32x3")) return NBPC_MATH_TF32X3;
    return NBPC_MATH_DEFAULT;
}
int g_nbpc_math_mode = nbpc_math_mode_from_env();

extern "C" {

int nbpc_set_math_mode(int mode) {
    if (mode != NBPC_MATH_FP32 && mode != NBPC_MATH_TF32 && mode != NBPC_MATH_TF32X3) {
        nbpc_set_error("nbpc_set_math_mode: unknown mode");
        return NBPC_EINVAL;
    }
    g_nbpc_math_mode = mode;
    return NBPC_OK;
}

int nbpc_get_math_mode(void) { return g_nbpc_math_mode; }

long long nbpc_launch_count(void) { return (long long)g_nbpc_launches; }

int nbpc_prof_enable(int on) {
#ifndef NBPC_HOST_EMU
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto &r : g_prof_recs) { g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b); }
    g_prof_recs.clear();
#endif
    g_nbpc_prof_on = on ? 1 : 0;
    return NBPC_OK;
}

// Writes "name\tlaunches\ttotal_ms\n" per kernel (aggregated over the records since the last
// nbpc_prof_enable) into buf; returns the number of bytes needed (call again with a larger buffer if
// it exceeds cap).  Synchronises on the recorded events.
long long nbpc_prof_report(char *buf, size_t cap) {
    std::string out;
#ifndef NBPC_HOST_EMU
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::map<std::string, std::pair<long long, double>> agg;
    std::vector<std::string> order;
    for (auto &r : g_prof_recs) {
        if (cudaEventSynchronize(r.b) != cudaSuccess) { cudaGetLastError(); continue; }
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) { cudaGetLastError(); continue; }
        auto it = agg.find(r.name);
        if (it == agg.end()) { order.push_back(r.name); agg[r.name] = std::make_pair(1LL, (double)ms); }
        else { it->second.first += 1; it->second.second += ms; }
    }
    for (auto &n : order) {
        out += n + "\t" + std::to_string(agg[n].first) + "\t" + std::to_string(agg[n].second) + "\n";
    }
#endif
    if (buf && cap > 0) {
        size_t n = out.size() < cap - 1 ? out.size() : cap - 1;
        memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return (long long)out.size() + 1;
}

int nbpc_version(void) { return 100; /* 0.1.0 */ }

const char *nbpc_last_error_string(void) { return g_last_error.c_str(); }

int nbpc_device_check(void) { return nbpc_require_sm100(); }

}  // extern "C"
