// Shared plumbing for libars_b200: error handling, the per-process context
// (device, stream, workspace cache, plan caches) and small device helpers.
// sm_100a only; no CPU fallback anywhere in this library.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

namespace ars {

typedef long long i64;

// ---------------------------------------------------------------- errors ----
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define ARS_CUDA(expr)                                                                  \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            char _b[512];                                                               \
            snprintf(_b, sizeof _b, "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,       \
                     cudaGetErrorString(_e));                                           \
            throw ::ars::Error(2, _b);                                                  \
        }                                                                               \
    } while (0)

#define ARS_CHECK(cond, msg)                                                            \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            char _b[512];                                                               \
            snprintf(_b, sizeof _b, "%s:%d: %s", __FILE__, __LINE__, (msg));            \
            throw ::ars::Error(1, _b);                                                  \
        }                                                                               \
    } while (0)

#define ARS_LAUNCH_CHECK() ARS_CUDA(cudaGetLastError())

void set_last_error(const char* msg);

// ------------------------------------------------------------- device mem ---
// Grow-only device buffer. Workspaces are cached in the context by name so a
// render of the same shape does not call cudaMalloc in steady state.
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes);
    void release();
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct FftPlan;        // fft_plan.cu
struct BluesteinPlan;  // bluestein.cu

struct Ctx {
    int device = -1;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    std::recursive_mutex mu;
    std::map<std::string, DevBuf> ws;                 // named workspaces
    std::map<int, FftPlan*> fft_plans;                // by log2(M)
    std::map<i64, BluesteinPlan*> blue_plans;         // by N
    std::vector<i64> blue_lru;
    size_t blue_bytes = 0;
    unsigned long long launches = 0;                  // kernels launched by this library
    void* pinned = nullptr;                           // small pinned scratch for D2H scalars
    size_t pinned_cap = 0;
    // side stream (side_begin / side_to_main / side_join): a stretch of small, latency-bound kernels -- IR synthesis,
    // the air fold, the IR partition spectra -- runs next to the delay-line transform instead of in front of it
    cudaStream_t aux = nullptr, main_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int side_state = 0;                               // 0: none, 1: enqueuing on aux, 2: back on main, join pending
    // lanes (lane_fork / lane_use / lane_join): independent stretches of one stage -- the stripes of the big-block
    // overlap-save convolution -- are enqueued round-robin on a few streams so that one stripe's loads run under another's
    // arithmetic; every lane starts after what the main stream holds at the fork and the main stream waits for all
    static constexpr int MAX_LANES = 4;
    cudaStream_t lanes[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t lane_done[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_lane_fork = nullptr;
    int lanes_open = 0;
    cudaStream_t lane_saved = nullptr;
    // head start (ars_render_dev_async): the part of a render that needs nothing from the render before it -- the IR chain
    // on the side stream, the first passes on the lanes -- is ordered after that render's CONVOLUTION (ev_conv_done: its
    // work buffers and IR spectrum are free) instead of after its tail, and runs next to the final pass and the loudness
    // meter.  Only between two asynchronous renders of the same geometry (every plan, table and buffer exists already).
    cudaEvent_t ev_conv_done = nullptr;
    bool conv_done_valid = false, conv_done_prev = false;   // recorded by the API call in progress / by the one before it
    bool head_start = false;                                 // this render takes it
    unsigned long long head_starts = 0;
    // meter stream (loud_begin / loud_end): the loudness meter of an asynchronous render -- barrier- and latency-bound, it
    // needs only the feed its final pass left -- and the read-back of its state block go to a stream of their own, so the
    // LAST PASS of the next render waits for this render's final pass only, not for the meter and its gating behind it.
    // Two feed buffers / state blocks alternate; a slot is reused after its ev_loud_done.  Every other API call orders
    // the main stream after the meter stream first (loud_join), so nothing else ever sees the two streams apart.
    cudaStream_t loud = nullptr, loud_saved = nullptr;
    cudaEvent_t ev_loud_fork = nullptr, ev_loud_done[2] = {nullptr, nullptr};
    bool loud_open = false, loud_recorded[2] = {false, false}, loud_pending = false;
    unsigned long long loud_forks = 0;
    // tail overlap: the meter stream zeroes a state block right after it has read it back (state_clean), and the stage
    // output alternates between two buffers like the feed; the last passes of the next render that uses the slot then wait
    // for that slot's ev_loud_done (lane_tail_event) instead of for the final pass of the render in between -- a render's
    // convolution overlaps the whole tail of the render before it.
    bool state_clean[2] = {false, false};
    cudaEvent_t lane_tail_event = nullptr;
    unsigned long long tail_overlaps = 0;

    DevBuf& buf(const char* name, size_t bytes) {
        DevBuf& b = ws[name];
        b.reserve(bytes);
        return b;
    }
    void* pinned_scratch(size_t bytes);
};

Ctx& ctx();                 // throws if ars_init was not called
bool ctx_ready();
unsigned long long ctx_generation();      // changes with every ars_init that creates a context (caches keyed on device buffers check it)
void ctx_init(int device);
void ctx_shutdown();

// What is enqueued between side_begin() and side_to_main() goes to the auxiliary stream (ordered after everything
// already on the main stream); what follows goes to the main stream again and runs concurrently with it until
// side_join(), which makes the main stream wait for the side work.  side_abort() restores the main stream (errors).
void side_begin();
void loud_begin();            // the library's current stream becomes the meter stream, ordered after what the main stream holds
void loud_end(int slot);      // records ev_loud_done[slot]; back to the main stream
void loud_wait_slot(int slot);   // the main stream waits until the meter work that used `slot` is through
void loud_join();             // the main stream waits for everything on the meter stream
void conv_done_mark();        // records ev_conv_done on the main stream (end of a render's convolution stage)
void lane_wait_main(int i);    // head start: lane i waits for what the main stream held at the fork
void lane_wait_tail(int i);    // head start: lane i waits until the stage output and the state block it is about to write are free
void side_to_main();
void side_join();
void side_abort();

// lane_fork(n): n lanes start behind the main stream's current position; lane_use(i) makes lane i the library's current
// stream (i < 0: back to the main stream); lane_wait_side(i): lane i waits for the side stream's work (a pending
// side_to_main); lane_join(): the main stream waits for every lane.
void lane_fork(int n);
void lane_use(int i);
void lane_wait_side(int i);
void lane_join();

// ---- per-kernel event timing (ars_profile_begin / ars_profile_end / ars_profile_report) ----
// Between begin and end every launch site that opens a KernelScope brackets its kernel with two CUDA events on the
// stream it launches on; the report groups the launches by name.  `bytes` = the launch's algorithmic (compulsory) bytes.
// bench.py runs one untimed render this way, with the side stream and the lanes off, so that the kernels run one
// after another as they do under ncu.
struct KernelScope {
    bool on;
    size_t index = 0;
    KernelScope(const char* name, double bytes);
    ~KernelScope();
};
void prof_session_begin();
// launches / ms / bytes: totals over the records whose name starts with `prefix` (null: all); json (optional): one
// object per kernel name {"name", "launches", "ms", "bytes"}
void prof_session_end(const char* prefix, long long* launches, double* ms, double* bytes, std::string* json);
bool prof_session_on();

inline void count_launch(int n = 1) { ctx().launches += (unsigned long long)n; }

// ------------------------------------------------------------ device math ---
struct c32 { float x, y; };

__host__ __device__ inline float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ inline float2 cmulc(float2 a, float2 b) {   // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__host__ __device__ inline float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ inline float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ inline float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
__host__ __device__ inline float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

// monotone map of |x| to an unsigned so atomicMax works on floats (NaN sorts above inf;
// callers that must ignore NaN clear it first).
__host__ __device__ inline unsigned abs_bits(float v) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(v) & 0x7fffffffu;
#else
    unsigned u; memcpy(&u, &v, 4); return u & 0x7fffffffu;
#endif
}

inline int ceil_div(i64 a, i64 b) { return (int)((a + b - 1) / b); }

}  // namespace ars
