// Error plumbing, launch counter and the optional event-based kernel-class profiler of the C ABI
// (include/nsf_b200.h).
#include "common.cuh"
#include <string.h>
#include <atomic>
#include <mutex>
#include <vector>

namespace nsf {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

struct ProfRec { int cls; double work; cudaEvent_t e0, e1; };
static std::mutex g_prof_mutex;
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof_recs;
static std::vector<cudaEvent_t> g_prof_pool;
static std::vector<size_t> g_prof_open[PROF_NUM_CLASSES];

bool prof_enabled() { return g_prof_on; }
static cudaEvent_t prof_event() {
    if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
void prof_begin(int cls, double work, cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    ProfRec r{cls, work, prof_event(), prof_event()};
    cudaEventRecord(r.e0, stream);
    g_prof_open[cls].push_back(g_prof_recs.size());
    g_prof_recs.push_back(r);
}
void prof_end(int cls, cudaStream_t stream) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    if (g_prof_open[cls].empty()) return;
    const size_t i = g_prof_open[cls].back();
    g_prof_open[cls].pop_back();
    cudaEventRecord(g_prof_recs[i].e1, stream);
}
}  // namespace nsf

using namespace nsf;

extern "C" const char* nsf_last_error(void) { return g_err; }
extern "C" const char* nsf_version(void) { return "nsf_b200 0.1 (sm_100a)"; }
extern "C" int64_t nsf_launch_count(void) { return g_launches.load(); }

extern "C" int nsf_prof_enable(int on) {
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    g_prof_on = on != 0;
    return NSF_OK;
}

extern "C" int nsf_prof_num_classes(void) { return PROF_NUM_CLASSES; }

extern "C" const char* nsf_prof_class_name(int cls) {
    static const char* names[PROF_NUM_CLASSES] = {"stft", "features", "gemm_tc", "gemm_simt", "attention", "net_other", "mvdr", "pit_cost",
                                                  "stitch", "activity", "istft", "pcm16"};
    return (cls >= 0 && cls < PROF_NUM_CLASSES) ? names[cls] : "?";
}

// Synchronises the device, then adds up (milliseconds, work, brackets) per class since the last collect.
extern "C" int nsf_prof_collect(double* ms, double* work, int64_t* count, int n_classes) {
    NSF_REQUIRE(ms && work && count && n_classes >= PROF_NUM_CLASSES, "nsf_prof_collect: need %d slots", PROF_NUM_CLASSES);
    NSF_CUDA(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lock(g_prof_mutex);
    for (int c = 0; c < n_classes; ++c) { ms[c] = 0.0; work[c] = 0.0; count[c] = 0; }
    for (auto& r : g_prof_recs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) == cudaSuccess) { ms[r.cls] += t; work[r.cls] += r.work; count[r.cls] += 1; }
        g_prof_pool.push_back(r.e0);
        g_prof_pool.push_back(r.e1);
    }
    g_prof_recs.clear();
    for (auto& v : g_prof_open) v.clear();
    cudaGetLastError();
    return NSF_OK;
}
