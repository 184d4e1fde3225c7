// Context, workspace cache and error plumbing of libars_b200.
#include "ars_common.cuh"

#include <algorithm>
#include <cstdlib>

namespace ars {

static Ctx* g_ctx = nullptr;
static std::mutex g_ctx_mu;
static thread_local std::string t_last_error;

void set_last_error(const char* msg) { t_last_error = msg ? msg : ""; }
const char* last_error_cstr() { return t_last_error.c_str(); }

void DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return;
    if (p) { ARS_CUDA(cudaFree(p)); p = nullptr; cap = 0; }
    // round up so slightly different clip lengths reuse the same allocation
    // (and with 25 % head-room, so a batch of clips of varying length settles after a few reallocations)
    size_t want = (bytes + bytes / 4 + (size_t)(1 << 20) - 1) & ~((size_t)(1 << 20) - 1);
    if (bytes < (1 << 20)) want = (bytes + 255) & ~(size_t)255;
    ARS_CUDA(cudaMalloc(&p, want));
    cap = want;
}

void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

void* Ctx::pinned_scratch(size_t bytes) {
    if (bytes > pinned_cap) {
        if (pinned) cudaFreeHost(pinned);
        pinned = nullptr;
        size_t want = bytes < 4096 ? 4096 : bytes;
        ARS_CUDA(cudaMallocHost(&pinned, want));
        pinned_cap = want;
    }
    return pinned;
}

static unsigned long long g_ctx_generation = 0;
unsigned long long ctx_generation() { return g_ctx_generation; }
bool ctx_ready() { return g_ctx != nullptr; }

Ctx& ctx() {
    if (!g_ctx) throw Error(3, "ars_init() has not been called (no CUDA context; this library has no CPU path)");
    return *g_ctx;
}

void ctx_init(int device) {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (g_ctx) {
        if (g_ctx->device == device || device < 0) return;
        throw Error(1, "ars_init: already initialised on another device (call ars_shutdown first)");
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        throw Error(3, std::string("ars_init: no CUDA device available (") + cudaGetErrorString(e) +
                           "); libars_b200 has no CPU fallback");
    if (device < 0) device = 0;
    if (device >= count) throw Error(1, "ars_init: device index out of range");
    ARS_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    ARS_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        throw Error(3, "ars_init: this build carries sm_100a code only (needs a Blackwell B200-class GPU)");
    Ctx* c = new Ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    ARS_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    ++g_ctx_generation;
    g_ctx = c;
}

void side_begin() {
    Ctx& c = ctx();
    if (c.side_state != 0) return;
    if (!c.aux) {
        // highest priority: the side chain's small CTAs must get in between the CTAs of the big transform next to it
        int least = 0, greatest = 0;
        ARS_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        ARS_CUDA(cudaStreamCreateWithPriority(&c.aux, cudaStreamNonBlocking, greatest));
        ARS_CUDA(cudaEventCreateWithFlags(&c.ev_fork, cudaEventDisableTiming));
        ARS_CUDA(cudaEventCreateWithFlags(&c.ev_join, cudaEventDisableTiming));
    }
    if (c.head_start && c.ev_conv_done) {
        ARS_CUDA(cudaStreamWaitEvent(c.aux, c.ev_conv_done, 0));      // (not the tail of the render before)
    } else {
        ARS_CUDA(cudaEventRecord(c.ev_fork, c.stream));
        ARS_CUDA(cudaStreamWaitEvent(c.aux, c.ev_fork, 0));
    }
    c.main_stream = c.stream;
    c.stream = c.aux;
    c.side_state = 1;
}
void side_to_main() {
    Ctx& c = ctx();
    if (c.side_state != 1) return;
    ARS_CUDA(cudaEventRecord(c.ev_join, c.aux));
    c.stream = c.main_stream;
    c.side_state = 2;
}
void side_join() {
    Ctx& c = ctx();
    if (c.side_state == 1) side_to_main();
    if (c.side_state != 2) return;
    ARS_CUDA(cudaStreamWaitEvent(c.stream, c.ev_join, 0));
    c.side_state = 0;
}
void loud_begin() {
    Ctx& c = ctx();
    if (c.loud_open || c.side_state != 0 || c.lanes_open) return;
    if (!c.loud) {
        // ARS_LOUD_PRIO=1 (experiment): the meter's CTAs go ahead of everything but the side chain when slots free up --
        // next to the persistent passes of the following render the meter otherwise gets whatever is left
        int least = 0, greatest = 0;
        ARS_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        const char* e = getenv("ARS_LOUD_PRIO");
        const int prio = (e && atoi(e) > 0) ? std::min(least, greatest + 1) : least;
        ARS_CUDA(cudaStreamCreateWithPriority(&c.loud, cudaStreamNonBlocking, prio));
        ARS_CUDA(cudaEventCreateWithFlags(&c.ev_loud_fork, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) ARS_CUDA(cudaEventCreateWithFlags(&c.ev_loud_done[i], cudaEventDisableTiming));
    }
    ARS_CUDA(cudaEventRecord(c.ev_loud_fork, c.stream));
    ARS_CUDA(cudaStreamWaitEvent(c.loud, c.ev_loud_fork, 0));
    c.loud_saved = c.stream;
    c.stream = c.loud;
    c.loud_open = true;
    ++c.loud_forks;
}
void loud_end(int slot) {
    Ctx& c = ctx();
    if (!c.loud_open) return;
    ARS_CUDA(cudaEventRecord(c.ev_loud_done[slot & 1], c.loud));
    c.loud_recorded[slot & 1] = true;
    c.loud_pending = true;
    c.stream = c.loud_saved;
    c.loud_open = false;
}
void loud_wait_slot(int slot) {
    Ctx& c = ctx();
    if (c.loud_open || !c.loud_recorded[slot & 1]) return;
    ARS_CUDA(cudaStreamWaitEvent(c.stream, c.ev_loud_done[slot & 1], 0));
}
void loud_join() {
    Ctx& c = ctx();
    if (c.loud_open || !c.loud_pending) return;
    for (int i = 0; i < 2; ++i)
        if (c.loud_recorded[i]) ARS_CUDA(cudaStreamWaitEvent(c.stream, c.ev_loud_done[i], 0));
    c.loud_pending = false;
}
void side_abort() {
    if (!ctx_ready()) return;
    Ctx& c = ctx();
    if (c.loud_open) {
        c.stream = c.loud_saved;
        c.loud_open = false;
        cudaStreamSynchronize(c.loud);
    }
    if (c.lanes_open) {
        c.stream = c.lane_saved;
        for (int i = 0; i < c.lanes_open; ++i) cudaStreamSynchronize(c.lanes[i]);
        c.lanes_open = 0;
    }
    if (c.side_state == 1) c.stream = c.main_stream;
    if (c.side_state != 0 && c.aux) cudaStreamSynchronize(c.aux);
    c.side_state = 0;
}

void lane_fork(int n) {
    Ctx& c = ctx();
    if (c.lanes_open) return;
    n = std::max(1, std::min(n, (int)Ctx::MAX_LANES));
    if (!c.ev_lane_fork) ARS_CUDA(cudaEventCreateWithFlags(&c.ev_lane_fork, cudaEventDisableTiming));
    for (int i = 0; i < n; ++i)
        if (!c.lanes[i]) {
            ARS_CUDA(cudaStreamCreateWithFlags(&c.lanes[i], cudaStreamNonBlocking));
            ARS_CUDA(cudaEventCreateWithFlags(&c.lane_done[i], cudaEventDisableTiming));
        }
    ARS_CUDA(cudaEventRecord(c.ev_lane_fork, c.stream));
    const bool early = c.head_start && c.ev_conv_done;              // lane_wait_main() orders what needs the main stream
    for (int i = 0; i < n; ++i) ARS_CUDA(cudaStreamWaitEvent(c.lanes[i], early ? c.ev_conv_done : c.ev_lane_fork, 0));
    c.lane_saved = c.stream;
    c.lanes_open = n;
}
void lane_wait_main(int i) {
    Ctx& c = ctx();
    if (!c.lanes_open || !c.head_start) return;
    ARS_CUDA(cudaStreamWaitEvent(c.lanes[i % c.lanes_open], c.ev_lane_fork, 0));
}
void lane_wait_tail(int i) {
    Ctx& c = ctx();
    if (!c.lanes_open || !c.head_start) return;
    ARS_CUDA(cudaStreamWaitEvent(c.lanes[i % c.lanes_open], c.lane_tail_event ? c.lane_tail_event : c.ev_lane_fork, 0));
}
void conv_done_mark() {
    Ctx& c = ctx();
    if (c.side_state != 0 || c.lanes_open) return;                  // (only from the main stream, everything joined)
    if (!c.ev_conv_done) ARS_CUDA(cudaEventCreateWithFlags(&c.ev_conv_done, cudaEventDisableTiming));
    ARS_CUDA(cudaEventRecord(c.ev_conv_done, c.stream));
    c.conv_done_valid = true;
}
void lane_use(int i) {
    Ctx& c = ctx();
    if (!c.lanes_open) return;
    c.stream = i < 0 ? c.lane_saved : c.lanes[i % c.lanes_open];
}
void lane_wait_side(int i) {
    Ctx& c = ctx();
    if (!c.lanes_open || c.side_state != 2) return;
    ARS_CUDA(cudaStreamWaitEvent(c.lanes[i % c.lanes_open], c.ev_join, 0));
}
void lane_join() {
    Ctx& c = ctx();
    if (!c.lanes_open) return;
    c.stream = c.lane_saved;
    for (int i = 0; i < c.lanes_open; ++i) {
        ARS_CUDA(cudaEventRecord(c.lane_done[i], c.lanes[i]));
        ARS_CUDA(cudaStreamWaitEvent(c.stream, c.lane_done[i], 0));
    }
    c.lanes_open = 0;
}

// ---------------------------------------------------------------- profiling ---
struct ProfRec { const char* name; double bytes; cudaEvent_t e0, e1; };
static struct {
    bool on = false;
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    size_t used = 0;
    std::string last_json;
} g_kprof;

static cudaEvent_t kprof_event() {
    if (g_kprof.used == g_kprof.pool.size()) {
        cudaEvent_t e;
        ARS_CUDA(cudaEventCreate(&e));
        g_kprof.pool.push_back(e);
    }
    return g_kprof.pool[g_kprof.used++];
}
bool prof_session_on() { return g_kprof.on; }
void prof_session_begin() {
    g_kprof.on = true;
    g_kprof.recs.clear();
    g_kprof.used = 0;
}
KernelScope::KernelScope(const char* name, double bytes) : on(g_kprof.on) {
    if (!on) return;
    ProfRec r;
    r.name = name;
    r.bytes = bytes;
    r.e0 = kprof_event();
    r.e1 = kprof_event();
    ARS_CUDA(cudaEventRecord(r.e0, ctx().stream));
    index = g_kprof.recs.size();
    g_kprof.recs.push_back(r);
}
KernelScope::~KernelScope() {
    if (on && index < g_kprof.recs.size()) cudaEventRecord(g_kprof.recs[index].e1, ctx().stream);
}
void prof_session_end(const char* prefix, long long* launches, double* ms, double* bytes, std::string* json) {
    ARS_CUDA(cudaDeviceSynchronize());
    g_kprof.on = false;
    struct Agg { long long n = 0; double ms = 0, bytes = 0; };
    std::vector<std::pair<std::string, Agg>> order;
    long long tl = 0;
    double tm = 0, tb = 0;
    for (const ProfRec& r : g_kprof.recs) {
        float t = 0.f;
        ARS_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
        size_t k = 0;
        for (; k < order.size(); ++k) if (order[k].first == r.name) break;
        if (k == order.size()) order.push_back({r.name, Agg()});
        order[k].second.n += 1;
        order[k].second.ms += t;
        order[k].second.bytes += r.bytes;
        if (!prefix || !strncmp(r.name, prefix, strlen(prefix))) { tl += 1; tm += t; tb += r.bytes; }
    }
    if (launches) *launches = tl;
    if (ms) *ms = tm;
    if (bytes) *bytes = tb;
    std::string js = "[";
    for (size_t k = 0; k < order.size(); ++k) {
        char b[512];
        snprintf(b, sizeof b, "%s{\"name\": \"%s\", \"launches\": %lld, \"ms\": %.6f, \"bytes\": %.0f}", k ? ", " : "",
                 order[k].first.c_str(), order[k].second.n, order[k].second.ms, order[k].second.bytes);
        js += b;
    }
    js += "]";
    g_kprof.last_json = js;
    if (json) *json = js;
}
const char* prof_last_report() { return g_kprof.last_json.c_str(); }

void hostio_release();           // hostio.cu
void fft_release_plans();       // fft_plan.cu
void bluestein_release_plans(); // bluestein.cu

void ctx_shutdown() {
    std::lock_guard<std::mutex> lk(g_ctx_mu);
    if (!g_ctx) return;
    cudaSetDevice(g_ctx->device);
    cudaStreamSynchronize(g_ctx->stream);
    if (g_ctx->loud) cudaStreamSynchronize(g_ctx->loud);
    hostio_release();
    bluestein_release_plans();
    fft_release_plans();
    for (auto& kv : g_ctx->ws) kv.second.release();
    if (g_ctx->pinned) cudaFreeHost(g_ctx->pinned);
    if (g_ctx->aux) {
        cudaStreamSynchronize(g_ctx->aux);
        cudaStreamDestroy(g_ctx->aux);
        cudaEventDestroy(g_ctx->ev_fork);
        cudaEventDestroy(g_ctx->ev_join);
    }
    for (int i = 0; i < Ctx::MAX_LANES; ++i)
        if (g_ctx->lanes[i]) {
            cudaStreamSynchronize(g_ctx->lanes[i]);
            cudaStreamDestroy(g_ctx->lanes[i]);
            cudaEventDestroy(g_ctx->lane_done[i]);
        }
    if (g_ctx->loud) {
        cudaStreamSynchronize(g_ctx->loud);
        cudaStreamDestroy(g_ctx->loud);
        cudaEventDestroy(g_ctx->ev_loud_fork);
        for (int i = 0; i < 2; ++i) cudaEventDestroy(g_ctx->ev_loud_done[i]);
    }
    if (g_ctx->ev_lane_fork) cudaEventDestroy(g_ctx->ev_lane_fork);
    if (g_ctx->ev_conv_done) cudaEventDestroy(g_ctx->ev_conv_done);
    cudaStreamDestroy(g_ctx->stream);
    delete g_ctx;
    g_ctx = nullptr;
}

}  // namespace ars
