// extern "C" surface of libars_b200 (declared in include/ars_b200.h).
#include "../../include/ars_b200.h"

#include <cmath>
#include <limits>

#include "epilogue.cuh"
#include "ir_synth.cuh"
#include "metrics.cuh"
#include "upols.cuh"

namespace ars {
const char* last_error_cstr();
void fft_profile_begin();
void fft_profile_end(long long* launches, double* ms, double* bytes);
const char* prof_last_report();
// hostio.cu: transfers for callers that hold pageable memory, pinned result blocks
void host_upload(void* d_dst, const void* h_src, size_t bytes, cudaStream_t after);
void host_download(void* h_dst, const void* d_src, size_t bytes, cudaStream_t after);
void host_staging_enable(int on);
void* host_block_alloc(size_t bytes);
void host_block_free(void* p);

#define ARS_API_BEGIN_NOJOIN                                            \
    try {                                                               \
        std::lock_guard<std::recursive_mutex> _lk(ctx().mu);            \
        ARS_CUDA(cudaSetDevice(ctx().device));                          \
        ctx().conv_done_prev = ctx().conv_done_valid;                   \
        ctx().conv_done_valid = false;
// (every call but ars_render_dev_async orders the main stream after the meter stream first: see Ctx::loud)
#define ARS_API_BEGIN                                                   \
    ARS_API_BEGIN_NOJOIN                                                \
        loud_join();
#define ARS_API_END                                                     \
        return ARS_OK;                                                  \
    } catch (const Error& e) {                                          \
        side_abort();                                                   \
        set_last_error(e.what());                                       \
        return e.code;                                                  \
    } catch (const std::exception& e) {                                 \
        side_abort();                                                   \
        set_last_error(e.what());                                       \
        return ARS_ERR_INTERNAL;                                        \
    }

// run-time options (ars_set_option)
static int g_opt_upols = 1;        // 1: use overlap-save whenever no exact-N mask is active; 0: always the N-point path
static int g_opt_upols_logf = 13;  // 2B = 2^logF points per overlap-save transform (12 or 13)
static int g_opt_sparse_ir = 1;    // 1: IR spectrum of sparse (procedural) IRs through the overlap-save route
static unsigned long long g_air_fold_count = 0;   // convolution stages that took the folded-air route
static int g_opt_air_fold = 1;     // 1: air absorption folded into the IR (overlap-save) when its error bound allows
static int g_opt_air_fold_eps_e9 = 2000;     // bound on the late path's transfer-function error, in 1e-9 (upols.cuh)
static int g_opt_side_stream = 1;  // 1: IR synthesis + fold + IR spectra on the side stream, next to the delay-line transform
static int g_opt_head_start = 1;        // 1: asynchronous renders of the same geometry overlap their head with the tail of the render before
static int g_opt_loud_stream = 1;       // 1: the loudness meter of an asynchronous render runs on the meter stream (Ctx::loud), off the chain
                                        //    last pass -> final pass -> last pass of the next render.  300 s render: 0.4636 -> 0.437 ms
                                        //    (seven runs within 0.4 %).  What the path sets up on first use (stream, second feed buffer
                                        //    and state block) costs 3-90 ms once: callers that time a loop warm up through this call
                                        //    (profiles/r02_stream_options_ab.txt: with a synchronous warm-up that cost landed in the
                                        //    timed region of one run in three and looked like a slow GPU)
static int g_opt_tail_overlap = 0;      // 1: with the meter stream, two stage-output buffers and state blocks zeroed behind their read-back:
                                        //    the last passes of a render do not wait for the final pass of the render before it
                                        //    (measured: 0.4367 ms, no better than the meter stream alone, and 118 MB more)
static int g_opt_lufs_from_stage = 0;   // 1: loudness meter fed from the stage output, next to the final pass (no feed array);
                                        // measured slower: both kernels are bound by issue slots, and recomputing the feed costs more
                                        // instructions than the 4 B per frame the final pass writes (0.660 against 0.634 ms)
static int g_opt_air_fold_max_taps = 131072;  // longest kept half-length of the air kernel; beyond: the exact N-point path

// number of 4096-tap partitions of the two IR parts that hold a non-zero tap (host arrays)
static int nonzero_partitions(const float* a, i64 na, const float* b, i64 nb) {
    const i64 L = std::max(a ? na : 0, b ? nb : 0);
    int count = 0;
    for (i64 lo = 0; lo < L; lo += 4096) {
        bool f = false;
        for (i64 i = lo; i < std::min(L, lo + 4096) && !f; ++i)
            f = (a && i < na && a[i] != 0.f) || (b && i < nb && b[i] != 0.f);
        count += f ? 1 : 0;
    }
    return count;
}

// the convolution stage: overlap-save when the render has no spectral mask, else the exact N-point filter
// Non-zero extent of the IR parts as far as the host knows it (folded-air form, upols.cuh); late_hi < 0: unknown.
struct IrExtent { i64 early_end = 0, late_lo = 0, late_hi = -1; };

static bool folds_air(const FilterSpec& fs, const IrExtent& ext, double rate, AirFold* af) {
    AirFold tmp;
    return g_opt_upols && g_opt_air_fold && ext.late_hi >= 0 && fs.level1 != 0.0 &&
           air_fold_plan(fs, ext.early_end, ext.late_lo, ext.late_hi, rate, 1e-9 * g_opt_air_fold_eps_e9,
                         g_opt_air_fold_max_taps, af ? af : &tmp);
}

// taps of a mask-free convolution: those beyond the host-known extent are zero (procedural IRs) and cost neither
// partitions nor block length
static i64 mask_free_taps(const FilterSpec& fs, i64 L0, i64 L1, const IrExtent& ext) {
    i64 taps = fs.mode == FILT_EXT ? L0 : std::max(L0, L1);
    if (fs.mode == FILT_SPLIT && ext.late_hi >= 0) taps = std::min(taps, std::max<i64>(1, std::max(ext.early_end, ext.late_hi)));
    return std::max<i64>(1, taps);
}

// The big-block route's first pass needs only the signal: enqueue it ahead of everything that makes the impulse
// response.  -> true when it went out (the IR chain then runs on the side stream, next to it).
static bool early_first_pass(const float* d_x, i64 n, int cin, i64 L0, i64 L1, const FilterSpec& fs, double rate,
                             const IrExtent& ext) {
    if (!g_opt_upols || !olsb_enabled() || fs.mode == FILT_MASK || fs.eq_on) return false;
    AirFold af;
    OlsbPlan pl;
    if (fs.air_on) {
        if (!(L1 > 0 && folds_air(fs, ext, rate, &af))) return false;
        i64 adv = 0, taps = 0;
        air_fold_geometry(af, L0, L1, g_opt_upols_logf, &adv, &taps);
        if (!olsb_plan(fs.N, taps, adv, fs.N, &pl)) return false;
        olsb_first_pass_early(d_x, n, cin, pl, adv, fs.N);
        return true;
    }
    if (!olsb_plan(fs.N, mask_free_taps(fs, L0, L1, ext), 0, 0, &pl)) return false;
    olsb_first_pass_early(d_x, n, cin, pl, 0, 0);
    return true;
}

static void convolution_stage(const float* d_x, i64 n, int cin, const float* d_ir0, i64 L0, const float* d_ir1, i64 L1,
                              const FilterSpec& fs, float2* d_y, RenderState* st, double rate = 0.0,
                              const IrExtent& ext = IrExtent()) {
    AirFold af;
    if (d_ir1 && folds_air(fs, ext, rate, &af)) {
        ++g_air_fold_count;
        upols_filter_airfold(d_x, n, cin, d_ir0, L0, d_ir1, L1, fs, af, d_y, st, g_opt_upols_logf);
    } else if (g_opt_upols && upols_applicable(fs) && fs.mode != FILT_MASK) {
        const i64 taps = mask_free_taps(fs, d_ir0 ? L0 : 0, d_ir1 ? L1 : 0, ext);
        OlsbPlan pl;
        if (olsb_plan(fs.N, taps, 0, 0, &pl)) olsb_filter(d_x, n, cin, d_ir0, L0, d_ir1, L1, fs, d_y, st, pl);
        else
            upols_filter(d_x, n, cin, d_ir0, L0, d_ir1, L1, fs, d_y, st, g_opt_upols_logf);
    } else
        spectral_filter(d_x, n, cin, d_ir0, L0, d_ir1, L1, fs, d_y, st);
}

static inline double clip(double v, double lo, double hi) { return std::min(hi, std::max(lo, v)); }

template <class T> static T* upload(const char* name, const T* host, size_t count) {
    Ctx& c = ctx();
    T* d = c.buf(name, sizeof(T) * std::max<size_t>(count, 1)).as<T>();
    if (count) host_upload(d, host, sizeof(T) * count, c.stream);       // (large pageable arrays: through the pinned ring)
    return d;
}
template <class T> static void download(T* host, const T* dev, size_t count) {
    if (count) host_download(host, dev, sizeof(T) * count, ctx().stream);
}
// ---- deferred metrics (ars_render_dev_async) ----
// An asynchronous render leaves its state block in a pinned slot (a D2H copy on the library stream) and is finished on
// the host -- dB values, status words -- the next time the stream is waited for anyway (ars_sync, ars_timer_end): the
// caller can enqueue render after render without the GPU idling while the host reads 80 bytes back and prepares the next one.
struct PendingMetrics {
    RenderState* h;
    i64 count;
    int lufs_status;
    ArsMetrics* out;
    ArsRenderParams p;
};
static std::vector<PendingMetrics> g_pending;
static std::vector<RenderState*> g_pending_blocks;           // pinned, PENDING_BLOCK states each
static constexpr size_t PENDING_BLOCK = 256;
static RenderState* pending_slot() {
    const size_t i = g_pending.size();
    if (i / PENDING_BLOCK >= g_pending_blocks.size()) {
        RenderState* b = nullptr;
        ARS_CUDA(cudaMallocHost(&b, sizeof(RenderState) * PENDING_BLOCK));
        g_pending_blocks.push_back(b);
    }
    return g_pending_blocks[i / PENDING_BLOCK] + i % PENDING_BLOCK;
}
static void finish_metrics(const RenderState& st, i64 count, int lufs_status, ArsMetrics* m, const ArsRenderParams* p = nullptr);
static void collect_pending() {                               // (the library stream has just been waited for)
    for (const PendingMetrics& q : g_pending) finish_metrics(*q.h, q.count, q.lufs_status, q.out, &q.p);
    g_pending.clear();
}
static void sync() {
    ARS_CUDA(cudaStreamSynchronize(ctx().stream));
    collect_pending();
}

static RenderState* fresh_state(int slot = 0, bool* was_clean = nullptr) {
    Ctx& c = ctx();
    RenderState* st = c.buf(slot ? "state.1" : "state.0", sizeof(RenderState)).as<RenderState>();
    const bool clean = was_clean && c.state_clean[slot & 1];      // (zeroed by the meter stream after its read-back: Ctx::state_clean)
    c.state_clean[slot & 1] = false;
    if (was_clean) *was_clean = clean;
    if (!clean) ARS_CUDA(cudaMemsetAsync(st, 0, sizeof(RenderState), c.stream));
    return st;
}

// ---- scalar helpers that mirror the reference's Python float arithmetic (C doubles, libm) ----
struct IrGeom { i64 length, split, tap_hi, late_len; };
static IrGeom ir_geometry(double rate, double dur, double max_delay, double split_time) {
    // rs.py:249,254-255,259,271-272 -- rate is int(rate) in the reference
    const i64 r = (i64)rate;
    IrGeom g;
    g.length = std::max<i64>(1, (i64)(dur * (double)r));
    g.split = std::max<i64>(1, std::min<i64>((i64)(split_time * (double)r), g.length - 1));
    const i64 mds = std::max<i64>(2, (i64)(max_delay * (double)r));
    g.tap_hi = std::min(mds, g.split);
    g.late_len = g.length - g.split;
    return g;
}

static IrSpec ir_spec(double rate, double dur, double absorption, double direc, double diffusion, const IrGeom& g,
                      int ntaps) {
    IrSpec sp;
    sp.length = g.length;
    sp.split = g.split;
    sp.ntaps = ntaps;
    const i64 r = (i64)rate;
    if (g.late_len > 0) {
        const double floor_ratio = std::pow(10.0, -50.0 / 20.0);                          // rs.py:273
        double decay = g.late_len > 1 ? std::pow(floor_ratio, 1.0 / (double)g.late_len) : 0.1;
        decay = clip(decay * (1.0 - absorption * 0.1), 0.8, 0.99999);                     // rs.py:277
        double amp = 0.6 * (1.0 - clip(direc, 0.0, 0.9));                                 // rs.py:279
        amp *= clip(1.0 / (1 + dur * 0.5), 0.3, 1.0);
        amp *= (1.0 - std::pow(absorption, 0.5));
        sp.width = (int)clip((double)r * 0.001 * (1.0 + diffusion * 2.0), 1, 10);         // rs.py:284
        amp *= (1.0 + diffusion * 0.2);                                                   // rs.py:294
        sp.amp = amp;
        sp.decay = decay;
    }
    return sp;
}

static std::vector<double> tap_strengths(const ArsIrDraws* dr, double absorption, double direc, i64 tap_hi) {
    std::vector<double> s((size_t)std::max(0, dr ? dr->ntaps : 0));
    for (size_t j = 0; j < s.size(); ++j) {
        double v = dr->tap_base[j] * (1.0 - absorption);                                   // rs.py:265
        v *= clip(direc, 0.1, 1.0);                                                        // rs.py:266
        v *= (1.0 - std::pow((double)dr->tap_delay[j] / (double)tap_hi, 0.7));             // rs.py:267
        s[j] = v;
    }
    return s;
}

static void dry_gain_factor(double dry_wet, double kill_start, double* dw_out, double* dmf_out) {
    // rs.py:93-105
    const double dw = clip(dry_wet, 0.0, 1.0), ks = clip(kill_start, 0.0, 1.0);
    double g = 1.0;
    if (ks < 1.0 && dw >= ks) {
        const double span = 1.0 - ks;
        g = span < 1e-6 ? 0.0 : clip(1.0 - (dw - ks) / span, 0.0, 1.0);
    }
    *dw_out = dw;
    *dmf_out = g;
}

static bool is_close_to_one(double a) {
    // np.isclose(a, 1.0): |a - 1| <= atol + rtol * |1| with atol 1e-8, rtol 1e-5
    return std::isfinite(a) && std::fabs(a - 1.0) <= (1e-8 + 1e-5 * 1.0);
}

static TailSpec make_tail(i64 N, int layout, double rate, double x_, double y_, double z_) {
    // rs.py:468, 475-485 (gains), 542, 549-550 (delays, height gain)
    TailSpec ts;
    ts.N = N;
    ts.layout = layout;
    ts.C = layout_channels(layout);
    const double x = clip(x_, 0.0, 1.0), y = clip(y_, 0.0, 1.0), z = clip(z_, 0.0, 1.0);
    const double gl = std::sqrt(1.0 - x), gr = std::sqrt(x);
    const double pull = (0.5 - z) * (std::fabs(y - 0.5) * 0.3);
    const double gf = std::max(0.0, std::sqrt(1.0 - y) + pull), gb = std::max(0.0, std::sqrt(y) - pull);
    const double PI = 3.141592653589793;       // math.pi
    ts.g_fl = gl * gf;
    ts.g_fr = gr * gf;
    ts.g_rl = gl * gb;
    ts.g_rr = gr * gb;
    ts.g_c = std::cos((x - 0.5) * PI) * gf;
    ts.g_lfe = 0.15f;
    const i64 r = (i64)rate;
    if (layout == LAYOUT_7_1) ts.delay = (i64)((double)(r * 12) / 1000.0);       // int(rate * 12 / 1000)
    else if (layout == LAYOUT_5_1_2) ts.delay = (i64)((double)(r * 18) / 1000.0);
    ts.height_gain = clip(z_, 0.0, 1.0) * 0.6;
    ts.stream = (olsb_stream_hints() >> 2) & 3;
    return ts;
}

static int layout_ok(int layout) { return layout >= LAYOUT_STEREO && layout <= LAYOUT_5_1_2; }

// p (optional): the render's parameters, for the oversampled true peak (needs the layout's gains)
static void finish_metrics(const RenderState& st, i64 count, int lufs_status, ArsMetrics* m,
                           const ArsRenderParams* p) {
    memset(m, 0, sizeof(*m));
    m->true_peak_4x_status = 1;
    if (p && (p->want_lufs & 2) && layout_ok(p->layout)) {
        const TailSpec ts = make_tail(0, p->layout, p->rate, p->x, p->y, p->z);
        const double tp = true_peak_linear(st, ts);
        m->true_peak_4x_dbfs = tp > 1e-15 ? 20.0 * std::log10(tp) : -std::numeric_limits<double>::infinity();
        m->true_peak_4x_status = 0;
    }
    const float peak = [&] { float f; unsigned u = st.peak_final; memcpy(&f, &u, 4); return f; }();
    const double inf = std::numeric_limits<double>::infinity();
    m->peak_linear = (double)peak;
    const float rms = count > 0 ? (float)std::sqrt(st.sumsq / (double)count) : 0.f;   // numpy keeps float32 here
    m->rms_linear = (double)rms;
    m->true_peak_dbfs = (double)peak > 1e-15 ? 20.0 * std::log10((double)peak) : -inf;
    m->rms_dbfs = (double)rms > 1e-15 ? 20.0 * std::log10((double)rms) : -inf;
    m->lufs_status = lufs_status;
    m->lufs = lufs_status == ARS_LUFS_OK ? st.lufs : 0.0;
}

// enqueue the loudness meter on the mono feed (rs.py:685-691); the value lands in d_state->lufs
static int lufs_enqueue(const float* d_mono, i64 N, double rate, RenderState* d_state) {
    const int s = integrated_loudness_async(d_mono, N, rate, &d_state->mono_max, &d_state->lufs);
    return s == 0 ? ARS_LUFS_OK : ARS_LUFS_NONE;
}


// ---------------------------------------------------------------- render core ----
static i64 render_out_len(const ArsRenderParams* p, i64 n, i64 ext_len) {
    if (p->external_ir) return ext_len > 0 ? n + ext_len - 1 : n;                  // rs.py:423-424
    const IrGeom g = ir_geometry(p->rate, p->ir_duration, p->ir_max_delay, p->ir_split_time);
    return n + g.length - 1;                                                       // rs.py:352-355
}

static void common_filter_spec(FilterSpec& fs, i64 N, double rate, double dry_wet, double kill, double bass,
                               double treble) {
    double dw, dmf;
    dry_gain_factor(dry_wet, kill, &dw, &dmf);
    fs.N = N;
    fs.dw = dw;
    fs.dry_gain = dmf * (1.0 - dw);
    if (N >= 2 && !(is_close_to_one(bass) && is_close_to_one(treble))) fill_eq(fs, N, rate, bass, treble);   // rs.py:389-391
}

// All pointers on the device except draws->tap_* (host); draws->noise on the device.  Nothing here waits for
// the GPU: the metrics end up in *st (device) and *lufs_status says how to read st->lufs.
static void render_core(const ArsRenderParams* p, const float* d_in, i64 n, int cin, const float* d_ext_ir, i64 ext_len,
                        const ArsIrDraws* draws, float* d_out_stereo, float* d_out_f32, short* d_out_pcm,
                        RenderState* st, bool want_metrics, int* lufs_status, bool allow_head_start = false,
                        int loud_slot = -1) {
    Ctx& c = ctx();
    ARS_CHECK(p && d_in && n > 0 && cin >= 1, "render: empty input");
    {
        // head start (see Ctx): between two renders of the same geometry, the one before it the API call before this one
        static ArsRenderParams last_p;
        static i64 last_n = -1, last_ext = -1;
        static int last_cin = -1;
        static unsigned long long last_gen = 0;
        const bool same = last_n == n && last_cin == cin && last_ext == ext_len && last_gen == ctx_generation() &&
                          memcmp(&last_p, p, sizeof(ArsRenderParams)) == 0;
        c.head_start = allow_head_start && g_opt_head_start && c.conv_done_prev && same;
        if (c.head_start) ++c.head_starts;
        last_p = *p; last_n = n; last_cin = cin; last_ext = ext_len; last_gen = ctx_generation();
    }
    struct HeadStartEnd { ~HeadStartEnd() { if (ctx_ready()) { ctx().head_start = false; ctx().lane_tail_event = nullptr; } } } head_start_end;
    if (!c.head_start) c.lane_tail_event = nullptr;
    else if (c.lane_tail_event) ++c.tail_overlaps;
    ARS_CHECK(p->rate >= 1.0, "render: bad sample rate");
    ARS_CHECK(layout_ok(p->layout), "render: unknown layout id");
    if (lufs_status) *lufs_status = ARS_LUFS_SKIPPED;
    const i64 N = render_out_len(p, n, ext_len);
    float2* y = c.buf(loud_slot == 1 && g_opt_tail_overlap ? "render.y1" : "render.y", sizeof(float2) * (size_t)N).as<float2>();
    FilterSpec fs;
    common_filter_spec(fs, N, p->rate, p->dry_wet, p->kill_start, p->bass_gain, p->treble_gain);
    if (p->external_ir) {
        ARS_CHECK(d_ext_ir && ext_len >= 1, "render: external IR missing");
        fs.mode = FILT_EXT;
        early_first_pass(d_in, n, cin, ext_len, 0, fs, p->rate, IrExtent());
        convolution_stage(d_in, n, cin, d_ext_ir, ext_len, nullptr, 0, fs, y, st);
    } else {
        ARS_CHECK(p->ir_duration > 0, "render: IR duration must be positive");
        const IrGeom g = ir_geometry(p->rate, p->ir_duration, p->ir_max_delay, p->ir_split_time);
        const int ntaps = draws ? draws->ntaps : 0;
        ARS_CHECK(!draws || draws->noise_len == g.late_len || draws->noise == nullptr,
                  "render: noise length does not match the IR geometry");
        ARS_CHECK(g.late_len == 0 || (draws && draws->noise), "render: tail noise missing");
        IrSpec sp = ir_spec(p->rate, p->ir_duration, p->absorption, p->directionality, p->diffusion, g, ntaps);
        fs.mode = FILT_SPLIT;
        fs.level0 = (g.length > 1 && p->early_level > 1e-6) ? p->early_level : 0.0;     // rs.py:360
        fs.level1 = (g.length > 1 && p->late_level > 1e-6) ? p->late_level : 0.0;       // rs.py:369
        if (p->air_absorption > 0.01 && N >= 2) fill_air(fs, N, p->rate, p->air_absorption);   // rs.py:378, 312-317
        // a procedural IR is a few dozen taps before the split plus a tail that decays by >= 1 % per sample and is
        // exactly zero in float32 some 10^4 samples later: a handful of non-zero partitions
        fs.sparse_ir = g_opt_sparse_ir;
        IrExtent ex;               // taps sit below the split; the tail starts there and underflows float32 to exact zeros
        ex.early_end = g.split;
        ex.late_lo = g.split;
        ex.late_hi = g.split;
        if (g.late_len > 0 && sp.amp > 0.0 && sp.decay > 0.0 && sp.decay < 1.0) {
            // |tail[j]| <= 1e6 * amp * decay^j: the boxcar mean of noise in [-1, 1] is re-scaled by std_raw / std_smooth,
            // and the reference only does that while std_smooth > 1e-6 (rs.py:289-291); float32 stores 0 below 2^-150
            const double j0 = std::log(7.0e-46 / (1.0e6 * sp.amp)) / std::log(sp.decay);
            ex.late_hi = g.split + (i64)std::min((double)g.late_len, std::max(0.0, std::ceil(j0) + 1.0));
        }
        // the folded-air route: IR synthesis, the fold and the IR partition spectra are a chain of small latency-bound
        // kernels that does not depend on the signal -- it runs on the side stream, next to the delay-line transform
        fft_touch_tables();        // (built once, on the main stream, before both streams read them)
        const bool early = early_first_pass(d_in, n, cin, g.length, g.length, fs, p->rate, ex);
        if (g_opt_side_stream && (early || folds_air(fs, ex, p->rate, nullptr))) side_begin();
        std::vector<double> strength = tap_strengths(draws, p->absorption, p->directionality, g.tap_hi);
        std::vector<i64> tap_pos;
        std::vector<double> tap_val;
        sp.ntaps = ir_early_taps(draws ? (const i64*)draws->tap_delay : nullptr, strength.data(), ntaps, g.length, tap_pos, tap_val);
        // (pageable host memory: cudaMemcpyAsync returns once the bytes are staged, so the vectors may die)
        const i64* d_delay = upload("ir.delay", tap_pos.data(), tap_pos.size());
        const double* d_strength = upload("ir.strength", tap_val.data(), tap_val.size());
        float* d_early = c.buf("ir.early", sizeof(float) * (size_t)g.length).as<float>();
        float* d_late = c.buf("ir.late", sizeof(float) * (size_t)g.length).as<float>();
        ir_synth(sp, d_delay, d_strength, draws ? draws->noise : nullptr, d_early, d_late);
        convolution_stage(d_in, n, cin, d_early, g.length, d_late, g.length, fs, y, st, p->rate, ex);
        side_join();               // (no-op unless the stage left the side stream open)
    }
    conv_done_mark();              // (the next render's head may start from here: see Ctx::head_start)
    c.head_start = false;
    if (d_out_stereo) {
        ARS_CUDA(cudaMemcpyAsync(d_out_stereo, y, sizeof(float2) * (size_t)N, cudaMemcpyDeviceToDevice, c.stream));
        guard_apply(d_out_stereo, N * 2, &st->max_stereo);
    }
    if (!d_out_f32 && !d_out_pcm && !want_metrics) return;
    const TailSpec ts = make_tail(N, p->layout, p->rate, p->x, p->y, p->z);
    tail_maxes(y, ts, st);
    if (want_metrics && (p->want_lufs & 2)) true_peak_from_stage(y, ts, st);       // (add-on, off unless asked for)
    if (want_metrics && (p->want_lufs & 1) && g_opt_lufs_from_stage && loudness_from_stage_possible(p->rate)) {
        // the loudness meter recomputes its feed from the stage output, so it needs nothing from the final pass: it runs
        // on the side stream NEXT TO it (one is bound by issue slots and conversions, the other by float64 latency)
        if (g_opt_side_stream) side_begin();
        const int s = integrated_loudness_from_stage(y, ts, p->rate, st);
        if (lufs_status) *lufs_status = s == 0 ? ARS_LUFS_OK : ARS_LUFS_NONE;
        side_to_main();
        tail_final(y, ts, st, d_out_f32, d_out_pcm, nullptr);
        side_join();
        return;
    }
    if (want_metrics && (p->want_lufs & 1) && final_with_loudness(y, ts, p->rate, st, d_out_f32, d_out_pcm)) {
        if (lufs_status) *lufs_status = ARS_LUFS_OK;
        return;
    }
    float* d_mono = nullptr;
    if (want_metrics && (p->want_lufs & 1))
        d_mono = c.buf(loud_slot == 1 ? "render.mono1" : "render.mono", sizeof(float) * (size_t)N).as<float>();
    tail_final(y, ts, st, d_out_f32, d_out_pcm, d_mono);
    if (d_mono && lufs_status) {
        // asynchronous renders: the meter (and the read-back of the state block, which the caller enqueues before it
        // calls loud_end) leave the main stream here -- the next render's last pass follows this final pass directly
        if (loud_slot >= 0) loud_begin();
        *lufs_status = lufs_enqueue(d_mono, N, p->rate, st);
    }
}

// ------------------------------------------------------------ copy pipeline ----
// Batched renders overlap the host->device copy of clip i+1 and the device->host copy of clip i-1 with the
// compute of clip i: two buffer slots, one copy stream per direction, events between them.
struct Pipe {
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaEvent_t in_ready[2] = {nullptr, nullptr};     // H2D of the slot finished
    cudaEvent_t done[2] = {nullptr, nullptr};         // compute of the slot finished (inputs free, outputs ready)
    cudaEvent_t out_free[2] = {nullptr, nullptr};     // D2H of the slot finished (outputs free)
    RenderState* h_state = nullptr;                   // pinned, one per clip of the current batch
    size_t h_state_cap = 0;
    void init() {
        if (h2d) return;
        ARS_CUDA(cudaStreamCreateWithFlags(&h2d, cudaStreamNonBlocking));
        ARS_CUDA(cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            ARS_CUDA(cudaEventCreateWithFlags(&in_ready[i], cudaEventDisableTiming));
            ARS_CUDA(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
            ARS_CUDA(cudaEventCreateWithFlags(&out_free[i], cudaEventDisableTiming));
        }
    }
    RenderState* states(size_t count) {
        if (count > h_state_cap) {
            if (h_state) cudaFreeHost(h_state);
            ARS_CUDA(cudaMallocHost(&h_state, sizeof(RenderState) * count));
            h_state_cap = count;
        }
        return h_state;
    }
};
static Pipe g_pipe;

static const char* slot_name(const char* base, int slot) {
    static thread_local char buf[2][64];
    static thread_local int k = 0;
    k ^= 1;
    snprintf(buf[k], sizeof buf[k], "%s.%d", base, slot);
    return buf[k];
}

}  // namespace ars

using namespace ars;

extern "C" {

int ars_init(int device) {
    try {
        ctx_init(device);
        return ARS_OK;
    } catch (const Error& e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return ARS_ERR_INTERNAL;
    }
}

static void api_release();       // the copy pipeline of ars_render_batch and the stopwatch events (below)
void ars_shutdown(void) {
    api_release();
    ctx_shutdown();
}

int ars_sync(void) {
    ARS_API_BEGIN
    sync();
    ARS_API_END
}

const char* ars_last_error(void) { return last_error_cstr(); }
const char* ars_version(void) { return "ars_b200 0.1 (sm_100a)"; }
uint64_t ars_launch_count(void) { return ctx_ready() ? ctx().launches : 0; }
uint64_t ars_air_fold_count(void) { return g_air_fold_count; }
uint64_t ars_olsb_count(void) { return olsb_count(); }
uint64_t ars_head_start_count(void) { return ctx_ready() ? ctx().head_starts : 0; }
uint64_t ars_meter_stream_count(void) { return ctx_ready() ? ctx().loud_forks : 0; }
uint64_t ars_tail_overlap_count(void) { return ctx_ready() ? ctx().tail_overlaps : 0; }
void* ars_stream(void) { return ctx_ready() ? (void*)ctx().stream : nullptr; }

static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;

// process-wide CUDA objects of this file belong to the device of the current context: destroyed with it, re-created
// on first use after the next ars_init (which may name another device)
static void api_release() {
    if (!ctx_ready()) return;
    std::lock_guard<std::recursive_mutex> lk(ctx().mu);
    cudaSetDevice(ctx().device);
    cudaStreamSynchronize(ctx().stream);
    Pipe& pp = g_pipe;
    if (pp.h2d) {
        cudaStreamSynchronize(pp.h2d);
        cudaStreamSynchronize(pp.d2h);
        for (int i = 0; i < 2; ++i) { cudaEventDestroy(pp.in_ready[i]); cudaEventDestroy(pp.done[i]); cudaEventDestroy(pp.out_free[i]); }
        cudaStreamDestroy(pp.h2d);
        cudaStreamDestroy(pp.d2h);
    }
    if (pp.h_state) cudaFreeHost(pp.h_state);
    pp = Pipe();
    if (g_ev0) { cudaEventDestroy(g_ev0); cudaEventDestroy(g_ev1); g_ev0 = g_ev1 = nullptr; }
    for (RenderState* b : g_pending_blocks) cudaFreeHost(b);
    g_pending_blocks.clear();
    g_pending.clear();
}

int ars_timer_begin(void) {
    ARS_API_BEGIN
    if (!g_ev0) { ARS_CUDA(cudaEventCreate(&g_ev0)); ARS_CUDA(cudaEventCreate(&g_ev1)); }
    ARS_CUDA(cudaEventRecord(g_ev0, ctx().stream));
    ARS_API_END
}

int ars_timer_end(float* ms) {
    ARS_API_BEGIN
    ARS_CHECK(g_ev0 && ms, "ars_timer_end without ars_timer_begin");
    ARS_CUDA(cudaEventRecord(g_ev1, ctx().stream));
    ARS_CUDA(cudaEventSynchronize(g_ev1));
    ARS_CUDA(cudaEventElapsedTime(ms, g_ev0, g_ev1));
    collect_pending();            // (everything enqueued before the stop event is complete: deferred metrics are due)
    ARS_API_END
}

int ars_set_option(const char* key, int32_t value) {
    ARS_API_BEGIN
    ARS_CHECK(key, "ars_set_option: null key");
    if (!strcmp(key, "upols")) g_opt_upols = value ? 1 : 0;
    else if (!strcmp(key, "sparse_ir")) g_opt_sparse_ir = value ? 1 : 0;
    else if (!strcmp(key, "upols_logf")) { ARS_CHECK(value == 12 || value == 13, "upols_logf must be 12 or 13"); g_opt_upols_logf = value; }
    else if (!strcmp(key, "air_fold")) g_opt_air_fold = value ? 1 : 0;
    else if (!strcmp(key, "air_fold_eps_e9")) { ARS_CHECK(value >= 1, "air_fold_eps_e9 must be >= 1"); g_opt_air_fold_eps_e9 = value; }
    else if (!strcmp(key, "air_fold_max_taps")) { ARS_CHECK(value >= 64, "air_fold_max_taps must be >= 64"); g_opt_air_fold_max_taps = value; }
    else if (!strcmp(key, "side_stream")) g_opt_side_stream = value ? 1 : 0;
    else if (!strcmp(key, "ols_r2")) upols_set_r2(value);
    else if (!strcmp(key, "olsb")) olsb_set_options(value ? 1 : 0, -1, -1);
    else if (!strcmp(key, "olsb_logf")) { ARS_CHECK(value == 0 || (value >= 18 && value <= 22), "olsb_logf must be 0 (automatic) or 18..22"); olsb_set_options(-1, value, -1); }
    else if (!strcmp(key, "olsb_lanes") || !strcmp(key, "olsb_first_all") || !strcmp(key, "olsb_reverse") ||
             !strcmp(key, "olsb_dryfold") || !strcmp(key, "olsb_early") || !strcmp(key, "stream_hints")) olsb_set_tuning(key, value);
    else if (!strcmp(key, "olsb_stripe")) { ARS_CHECK(value >= 0, "olsb_stripe must be >= 0"); olsb_set_options(-1, -1, value); }
    else if (!strcmp(key, "lufs_fused")) loudness_set_fused(value);
    else if (!strcmp(key, "lufs_from_stage")) g_opt_lufs_from_stage = value ? 1 : 0;
    else if (!strcmp(key, "lufs_ctas_per_sm")) loudness_set_ctas_per_sm(value);
    else if (!strcmp(key, "final_in_meter")) loudness_set_final_in_meter(value);
    else if (!strcmp(key, "head_start")) g_opt_head_start = value ? 1 : 0;
    else if (!strcmp(key, "loud_stream")) g_opt_loud_stream = value ? 1 : 0;
    else if (!strcmp(key, "tail_overlap")) g_opt_tail_overlap = value ? 1 : 0;
    else if (!strcmp(key, "final_lean")) { ARS_CHECK(value >= 0 && value <= 2, "final_lean must be 0, 1 or 2"); tail_set_lean(value); }
    else if (!strcmp(key, "host_staging")) host_staging_enable(value);
    else if (!strcmp(key, "mac_tiled_min")) { ARS_CHECK(value >= 1, "mac_tiled_min must be >= 1"); upols_set_mac_tiled_min(value); }
    else ARS_CHECK(false, "ars_set_option: unknown option");
    ARS_API_END
}

int ars_profile_begin(void) {
    ARS_API_BEGIN
    fft_profile_begin();
    ARS_API_END
}

int ars_profile_end(int64_t* launches, double* ms, double* bytes) {
    ARS_API_BEGIN
    long long l = 0;
    double m = 0.0, b = 0.0;
    fft_profile_end(&l, &m, &b);
    if (launches) *launches = l;
    if (ms) *ms = m;
    if (bytes) *bytes = b;
    ARS_API_END
}

int ars_ir_geometry(double rate, double ir_duration, double ir_max_delay, double ir_split_time, int64_t* length,
                    int64_t* split, int64_t* tap_hi, int64_t* late_len) {
    if (!(rate >= 1.0) || !(ir_duration > 0)) { set_last_error("ars_ir_geometry: rate and duration must be positive"); return ARS_ERR_ARG; }
    const IrGeom g = ir_geometry(rate, ir_duration, ir_max_delay, ir_split_time);
    if (length) *length = g.length;
    if (split) *split = g.split;
    if (tap_hi) *tap_hi = g.tap_hi;
    if (late_len) *late_len = g.late_len;
    return ARS_OK;
}

int ars_ir_synth(double rate, double ir_duration, double ir_max_delay, double absorption, double directionality,
                 double ir_split_time, double diffusion, const ArsIrDraws* draws, float* early, float* late,
                 int64_t length) {
    ARS_API_BEGIN
    ARS_CHECK(rate >= 1.0 && ir_duration > 0 && early && late, "ars_ir_synth: bad arguments");
    const IrGeom g = ir_geometry(rate, ir_duration, ir_max_delay, ir_split_time);
    ARS_CHECK(g.length == length, "ars_ir_synth: output length does not match max(1, int(duration * rate))");
    const int ntaps = draws ? draws->ntaps : 0;
    ARS_CHECK(g.late_len == 0 || (draws && draws->noise && draws->noise_len == g.late_len),
              "ars_ir_synth: tail noise missing or of the wrong length");
    IrSpec sp = ir_spec(rate, ir_duration, absorption, directionality, diffusion, g, ntaps);
    std::vector<double> strength = tap_strengths(draws, absorption, directionality, g.tap_hi);
    std::vector<i64> tap_pos;
    std::vector<double> tap_val;
    sp.ntaps = ir_early_taps(draws ? (const i64*)draws->tap_delay : nullptr, strength.data(), ntaps, g.length, tap_pos, tap_val);
    const i64* d_delay = upload("ir.delay", tap_pos.data(), tap_pos.size());
    const double* d_strength = upload("ir.strength", tap_val.data(), tap_val.size());
    const double* d_noise = upload("ir.noise", draws ? draws->noise : nullptr, (size_t)g.late_len);
    Ctx& c = ctx();
    float* d_early = c.buf("ir.early", sizeof(float) * (size_t)g.length).as<float>();
    float* d_late = c.buf("ir.late", sizeof(float) * (size_t)g.length).as<float>();
    ir_synth(sp, d_delay, d_strength, d_noise, d_early, d_late);
    download(early, d_early, (size_t)g.length);
    download(late, d_late, (size_t)g.length);
    sync();
    ARS_API_END
}

int ars_air_filter(const float* sig, int64_t n, double rate, double air, float* out) {
    ARS_API_BEGIN
    ARS_CHECK(sig && out && n >= 2 && rate > 0, "ars_air_filter: needs an (n >= 2, 2) signal");
    Ctx& c = ctx();
    const float* d_x = upload("in.x", sig, (size_t)n * 2);
    RenderState* st = fresh_state();
    FilterSpec fs;
    fs.mode = FILT_MASK;
    fs.N = n;
    fill_air(fs, n, rate, air);
    float2* y = c.buf("render.y", sizeof(float2) * (size_t)n).as<float2>();
    spectral_filter(d_x, n, 2, nullptr, 0, nullptr, 0, fs, y, st);
    download(out, reinterpret_cast<const float*>(y), (size_t)n * 2);
    sync();
    ARS_API_END
}

int ars_resample(const float* sig, int64_t n, int64_t num, float* out) {
    ARS_API_BEGIN
    ARS_CHECK(sig && out && n >= 1 && num >= 1, "ars_resample: needs an (n >= 1, 2) signal and num >= 1");
    Ctx& c = ctx();
    const float* d_x = upload("in.x", sig, (size_t)n * 2);
    RenderState* st = fresh_state();
    float2* y = c.buf("render.y", sizeof(float2) * (size_t)num).as<float2>();
    resample_stereo(d_x, n, num, y, st);
    download(out, reinterpret_cast<const float*>(y), (size_t)num * 2);
    sync();
    ARS_API_END
}

int ars_dry_wet_mix(const float* dry, int64_t n_dry, const float* wet, int64_t n_wet, int32_t ch, double dry_wet,
                    double kill_start, float* out) {
    ARS_API_BEGIN
    ARS_CHECK(ch >= 1 && n_dry >= 0 && n_wet >= 0 && out, "ars_dry_wet_mix: bad arguments");
    const i64 total = std::max(n_dry, n_wet);
    if (total == 0) return ARS_OK;
    Ctx& c = ctx();
    const float* d_dry = upload("in.dry", dry, (size_t)(n_dry * ch));
    const float* d_wet = upload("in.wet", wet, (size_t)(n_wet * ch));
    float* d_out = c.buf("out.f32", sizeof(float) * (size_t)(total * ch)).as<float>();
    double dw, dmf;
    dry_gain_factor(dry_wet, kill_start, &dw, &dmf);
    mix_dry_wet(d_dry, n_dry, d_wet, n_wet, ch, dmf, dw, d_out);
    download(out, d_out, (size_t)(total * ch));
    sync();
    ARS_API_END
}

const char* ars_profile_report(void) { return prof_last_report(); }

void* ars_host_alloc(int64_t bytes) {
    try {
        if (bytes < 0 || !ctx_ready()) return nullptr;
        ARS_CUDA(cudaSetDevice(ctx().device));
        return host_block_alloc((size_t)bytes);
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return nullptr;
    }
}
void ars_host_free(void* p) { host_block_free(p); }

int64_t ars_convolve_out_len(int64_t n, int64_t len_early, int64_t len_late) {
    // rs.py:351-355 (a missing IR stands for zeros(1))
    if (len_early <= 0) len_early = 1;
    if (len_late <= 0) len_late = 1;
    return std::max<i64>(n, std::max(n + len_early - 1, n + len_late - 1));
}

int ars_convolve_split(const float* data, int64_t n, int32_t cin, const float* early, int64_t len_early,
                       const float* late, int64_t len_late, double early_level, double late_level, double dry_wet,
                       double bass_gain, double treble_gain, double rate, double kill_start, double air_absorption,
                       float* out) {
    ARS_API_BEGIN
    ARS_CHECK(data && out && n > 0 && cin >= 1 && rate > 0, "ars_convolve_split: bad arguments");
    Ctx& c = ctx();
    if (!early) len_early = 0;
    if (!late) len_late = 0;
    const i64 N = ars_convolve_out_len(n, len_early, len_late);
    const float* d_x = upload("in.x", data, (size_t)n * cin);
    const float* d_e = len_early > 0 ? upload("in.early", early, (size_t)len_early) : nullptr;
    const float* d_l = len_late > 0 ? upload("in.late", late, (size_t)len_late) : nullptr;
    RenderState* st = fresh_state();
    FilterSpec fs;
    common_filter_spec(fs, N, rate, dry_wet, kill_start, bass_gain, treble_gain);
    fs.mode = FILT_SPLIT;
    fs.level0 = (len_early > 1 && early_level > 1e-6) ? early_level : 0.0;          // rs.py:360
    fs.level1 = (len_late > 1 && late_level > 1e-6) ? late_level : 0.0;            // rs.py:369
    if (air_absorption > 0.01 && N >= 2) fill_air(fs, N, rate, air_absorption);     // rs.py:378
    fs.sparse_ir = (g_opt_sparse_ir && nonzero_partitions(early, len_early, late, len_late) <= 12) ? 1 : 0;
    float2* y = c.buf("render.y", sizeof(float2) * (size_t)N).as<float2>();
    IrExtent ex;
    for (i64 i = len_early; i > 0 && ex.early_end == 0; --i) if (early[i - 1] != 0.f) ex.early_end = i;
    ex.late_lo = len_late;
    ex.late_hi = 0;
    for (i64 i = 0; i < len_late; ++i) if (late[i] != 0.f) { ex.late_lo = i; break; }
    for (i64 i = len_late; i > ex.late_lo; --i) if (late[i - 1] != 0.f) { ex.late_hi = i; break; }
    if (ex.late_hi < ex.late_lo) ex.late_lo = ex.late_hi = 0;
    convolution_stage(d_x, n, cin, d_e, len_early, d_l, len_late, fs, y, st, rate, ex);
    guard_apply(reinterpret_cast<float*>(y), N * 2, &st->max_stereo);               // rs.py:402-404
    download(out, reinterpret_cast<const float*>(y), (size_t)N * 2);
    sync();
    ARS_API_END
}

int ars_convolve_external(const float* data, int64_t n, int32_t cin, const float* ir, int64_t L, double dry_wet,
                          double bass_gain, double treble_gain, double rate, double kill_start, float* out) {
    ARS_API_BEGIN
    ARS_CHECK(data && out && ir && n > 0 && cin >= 1 && L >= 1 && rate > 0, "ars_convolve_external: bad arguments");
    Ctx& c = ctx();
    const i64 N = n + L - 1;
    const float* d_x = upload("in.x", data, (size_t)n * cin);
    const float* d_ir = upload("in.ir", ir, (size_t)L * 2);
    RenderState* st = fresh_state();
    FilterSpec fs;
    common_filter_spec(fs, N, rate, dry_wet, kill_start, bass_gain, treble_gain);
    fs.mode = FILT_EXT;
    float2* y = c.buf("render.y", sizeof(float2) * (size_t)N).as<float2>();
    convolution_stage(d_x, n, cin, d_ir, L, nullptr, 0, fs, y, st);
    guard_apply(reinterpret_cast<float*>(y), N * 2, &st->max_stereo);               // rs.py:456-458
    download(out, reinterpret_cast<const float*>(y), (size_t)N * 2);
    sync();
    ARS_API_END
}

int ars_pan(const float* stereo, int64_t n, double x, double y, double z, float* out) {
    ARS_API_BEGIN
    ARS_CHECK(stereo && out && n > 0, "ars_pan: bad arguments");
    Ctx& c = ctx();
    const float* d_s = upload("in.x", stereo, (size_t)n * 2);
    float* d_six = c.buf("out.f32", sizeof(float) * (size_t)n * 6).as<float>();
    RenderState* st = fresh_state();
    const TailSpec ts = make_tail(n, LAYOUT_5_1, 48000.0, x, y, z);
    pan_stage(d_s, n, ts, d_six);
    absmax_f32(d_six, n * 6, &st->max_pan);
    guard_apply(d_six, n * 6, &st->max_pan);                                          // rs.py:497-499
    download(out, d_six, (size_t)n * 6);
    sync();
    ARS_API_END
}

int ars_delay(const float* sig, int64_t n, int32_t ch, int64_t delay_samples, float* out) {
    ARS_API_BEGIN
    ARS_CHECK(sig && out && n > 0 && ch >= 1, "ars_delay: bad arguments");
    Ctx& c = ctx();
    const float* d_in = upload("in.x", sig, (size_t)n * ch);
    float* d_out = c.buf("out.f32", sizeof(float) * (size_t)n * ch).as<float>();
    delay_stage(d_in, n, ch, delay_samples, d_out);
    download(out, d_out, (size_t)n * ch);
    sync();
    ARS_API_END
}

int ars_layout_channels(int32_t layout) { return layout_ok(layout) ? layout_channels(layout) : -1; }

int ars_map_channels(const float* six, int64_t n, int32_t layout, double rate, double z, float* out) {
    ARS_API_BEGIN
    ARS_CHECK(six && out && n > 0 && layout_ok(layout) && rate > 0, "ars_map_channels: bad arguments");
    Ctx& c = ctx();
    const float* d_six = upload("in.x", six, (size_t)n * 6);
    const TailSpec ts = make_tail(n, layout, rate, 0.5, 0.5, z);
    float* d_out = c.buf("out.f32", sizeof(float) * (size_t)n * ts.C).as<float>();
    RenderState* st = fresh_state();
    map_stage(d_six, n, ts, d_out);
    absmax_f32(d_out, n * ts.C, &st->max_map);
    guard_apply(d_out, n * ts.C, &st->max_map);                                       // rs.py:558-560
    download(out, d_out, (size_t)n * ts.C);
    sync();
    ARS_API_END
}

int ars_metrics(const float* data, int64_t n, int32_t ch, double rate, int32_t want_lufs, ArsMetrics* out) {
    ARS_API_BEGIN
    ARS_CHECK(data && out && n > 0 && ch >= 1 && rate > 0, "ars_metrics: bad arguments");
    Ctx& c = ctx();
    const float* d_x = upload("in.x", data, (size_t)n * ch);
    RenderState* st = fresh_state();
    float* d_mono = want_lufs ? c.buf("render.mono", sizeof(float) * (size_t)n).as<float>() : nullptr;
    sums_stage(d_x, n, ch, st, d_mono);
    int lufs_status = ARS_LUFS_SKIPPED;
    if (want_lufs) lufs_status = lufs_enqueue(d_mono, n, rate, st);
    RenderState h;
    download(&h, st, 1);
    sync();
    finish_metrics(h, n * ch, lufs_status, out);
    ARS_API_END
}

int ars_true_peak_4x(const float* data, int64_t n, int32_t ch, double* dbtp) {
    ARS_API_BEGIN
    ARS_CHECK(data && dbtp && n > 0 && ch >= 1, "ars_true_peak_4x: bad arguments");
    const float* d_x = upload("in.x", data, (size_t)n * ch);
    const double tp = true_peak_of_array(d_x, n, ch);
    *dbtp = tp > 1e-15 ? 20.0 * std::log10(tp) : -std::numeric_limits<double>::infinity();
    ARS_API_END
}

int ars_channel_rms(const float* data, int64_t n, int32_t ch, float* rms_out, float* side_rms_out) {
    ARS_API_BEGIN
    ARS_CHECK(data && rms_out && n > 0 && ch >= 1 && ch <= 8, "ars_channel_rms: needs (n > 0, 1..8 channels)");
    Ctx& c = ctx();
    const float* d_x = upload("in.x", data, (size_t)n * ch);
    double* d_s = c.buf("chan.sums", sizeof(double) * 9).as<double>();
    ARS_CUDA(cudaMemsetAsync(d_s, 0, sizeof(double) * 9, c.stream));
    channel_sums(d_x, n, ch, d_s);
    double h[9];
    download(h, d_s, (size_t)ch + 1);
    sync();
    for (int k = 0; k < ch; ++k) rms_out[k] = (float)std::sqrt(h[k] / (double)n);       // np.sqrt(np.mean(x**2)), float32
    if (side_rms_out) *side_rms_out = ch >= 2 ? (float)std::sqrt(h[ch] / (double)n) : 0.f;
    ARS_API_END
}

int64_t ars_spectrogram_segments(int64_t n, int32_t nperseg) {
    if (nperseg < 2 || n < nperseg) return 0;
    return (n - nperseg / 2) / (nperseg - nperseg / 2);
}

int ars_spectrogram(const float* data, int64_t n, int32_t ch, double rate, int32_t nperseg, float* sxx_out) {
    ARS_API_BEGIN
    ARS_CHECK(data && sxx_out && n > 0 && ch >= 1 && rate > 0, "ars_spectrogram: bad arguments");
    Ctx& c = ctx();
    const float* d_x = upload("in.x", data, (size_t)n * ch);
    const i64 nseg = ars_spectrogram_segments(n, nperseg);
    ARS_CHECK(nseg >= 1, "ars_spectrogram: signal shorter than one segment");
    const size_t count = (size_t)(nperseg / 2 + 1) * (size_t)nseg;
    float* d_s = c.buf("spec.sxx", sizeof(float) * count).as<float>();
    spectrogram_psd(d_x, n, ch, rate, nperseg, d_s, nullptr);      // first channel (rs.py:621)
    download(sxx_out, d_s, count);
    sync();
    ARS_API_END
}

int ars_pcm16(const float* data, int64_t count, int16_t* out) {
    ARS_API_BEGIN
    ARS_CHECK(data && out && count > 0, "ars_pcm16: bad arguments");
    Ctx& c = ctx();
    const float* d_x = upload("in.x", data, (size_t)count);
    short* d_p = c.buf("out.pcm", sizeof(short) * (size_t)count).as<short>();
    pcm16_stage(d_x, count, d_p);
    download(reinterpret_cast<short*>(out), d_p, (size_t)count);
    sync();
    ARS_API_END
}

int64_t ars_render_out_len(const ArsRenderParams* p, int64_t n, int64_t ext_ir_len) {
    if (!p || n <= 0) return 0;
    return render_out_len(p, n, ext_ir_len);
}

int ars_render(const ArsRenderParams* p, const float* in, int64_t n, int32_t cin, const float* ext_ir,
               int64_t ext_ir_len, const ArsIrDraws* draws, float* out_stereo, float* out_f32, int16_t* out_pcm,
               ArsMetrics* metrics) {
    ARS_API_BEGIN
    ARS_CHECK(p && in && n > 0 && cin >= 1 && layout_ok(p->layout), "ars_render: bad arguments");
    Ctx& c = ctx();
    const i64 N = render_out_len(p, n, ext_ir_len);
    const int C = layout_channels(p->layout);
    const float* d_in = upload("in.x", in, (size_t)n * cin);
    const float* d_ir = nullptr;
    ArsIrDraws dd;
    memset(&dd, 0, sizeof dd);
    if (p->external_ir) {
        ARS_CHECK(ext_ir && ext_ir_len >= 1, "ars_render: external IR missing");
        d_ir = upload("in.ir", ext_ir, (size_t)ext_ir_len * 2);
    } else if (draws) {
        dd = *draws;
        dd.noise = upload("ir.noise", draws->noise, (size_t)std::max<i64>(0, draws->noise_len));
    }
    float* d_st = out_stereo ? c.buf("out.stereo", sizeof(float) * (size_t)N * 2).as<float>() : nullptr;
    float* d_f = out_f32 ? c.buf("out.f32", sizeof(float) * (size_t)N * C).as<float>() : nullptr;
    short* d_p = out_pcm ? c.buf("out.pcm", sizeof(short) * (size_t)N * C).as<short>() : nullptr;
    RenderState* st = fresh_state();
    int lufs_status = ARS_LUFS_SKIPPED;
    render_core(p, d_in, n, cin, d_ir, ext_ir_len, p->external_ir ? nullptr : &dd, d_st, d_f, d_p, st, metrics != nullptr,
                &lufs_status);
    if (out_stereo) download(out_stereo, d_st, (size_t)N * 2);
    if (out_f32) download(out_f32, d_f, (size_t)N * C);
    if (out_pcm) download(reinterpret_cast<short*>(out_pcm), d_p, (size_t)N * C);
    RenderState* h = static_cast<RenderState*>(c.pinned_scratch(sizeof(RenderState)));
    if (metrics) download(h, st, 1);
    sync();
    if (metrics) finish_metrics(*h, N * C, lufs_status, metrics, p);
    ARS_API_END
}

static void render_dev_impl(const ArsRenderParams* p, const float* d_in, int64_t n, int32_t cin, const float* d_ext_ir,
                            int64_t ext_ir_len, const ArsIrDraws* d_draws, float* d_out_stereo, float* d_out_f32,
                            int16_t* d_out_pcm, ArsMetrics* metrics, bool deferred) {
    ARS_CHECK(p && layout_ok(p->layout), "ars_render_dev: bad arguments");
    Ctx& c = ctx();
    // asynchronous renders that want the loudness alternate between two state blocks / feed buffers (Ctx::loud)
    static int next_slot = 0;
    int slot = -1;
    if (deferred && g_opt_loud_stream && metrics && (p->want_lufs & 1)) {
        slot = next_slot;
        next_slot ^= 1;
        loud_wait_slot(slot);          // (the meter of the render before the last one: long through)
    } else {
        loud_join();
    }
    bool clean = false;
    RenderState* st = fresh_state(slot == 1 ? 1 : 0, slot >= 0 && g_opt_tail_overlap ? &clean : nullptr);
    // (a clean block was zeroed behind ev_loud_done[slot]: what writes it first -- the last passes -- may wait for that alone)
    c.lane_tail_event = clean && c.loud_recorded[slot & 1] ? c.ev_loud_done[slot & 1] : nullptr;
    int lufs_status = ARS_LUFS_SKIPPED;
    render_core(p, d_in, n, cin, d_ext_ir, ext_ir_len, d_draws, d_out_stereo, d_out_f32,
                reinterpret_cast<short*>(d_out_pcm), st, metrics != nullptr, &lufs_status, deferred, slot);
    struct LoudEnd { int s; ~LoudEnd() { if (ctx_ready()) loud_end(s); } } loud_end_guard{slot < 0 ? 0 : slot};   // (no-op unless the meter stream is open)
    if (!metrics) return;
    const i64 N = render_out_len(p, n, ext_ir_len);
    if (deferred) {
        PendingMetrics q;
        q.h = pending_slot();
        q.count = N * layout_channels(p->layout);
        q.lufs_status = lufs_status;
        q.out = metrics;
        q.p = *p;
        download(q.h, st, 1);
        g_pending.push_back(q);
        if (c.loud_open && g_opt_tail_overlap) {          // the block is zero again once ev_loud_done[slot] (loud_end, below) has passed
            ARS_CUDA(cudaMemsetAsync(st, 0, sizeof(RenderState), c.stream));
            c.state_clean[slot & 1] = true;
        }
        return;
    }
    RenderState* h = static_cast<RenderState*>(c.pinned_scratch(sizeof(RenderState)));
    download(h, st, 1);
    sync();
    finish_metrics(*h, N * layout_channels(p->layout), lufs_status, metrics, p);
}

int ars_render_dev(const ArsRenderParams* p, const float* d_in, int64_t n, int32_t cin, const float* d_ext_ir,
                   int64_t ext_ir_len, const ArsIrDraws* d_draws, float* d_out_stereo, float* d_out_f32,
                   int16_t* d_out_pcm, ArsMetrics* metrics) {
    ARS_API_BEGIN
    render_dev_impl(p, d_in, n, cin, d_ext_ir, ext_ir_len, d_draws, d_out_stereo, d_out_f32, d_out_pcm, metrics, false);
    ARS_API_END
}

int ars_render_dev_async(const ArsRenderParams* p, const float* d_in, int64_t n, int32_t cin, const float* d_ext_ir,
                         int64_t ext_ir_len, const ArsIrDraws* d_draws, float* d_out_stereo, float* d_out_f32,
                         int16_t* d_out_pcm, ArsMetrics* metrics) {
    ARS_API_BEGIN_NOJOIN
    render_dev_impl(p, d_in, n, cin, d_ext_ir, ext_ir_len, d_draws, d_out_stereo, d_out_f32, d_out_pcm, metrics, true);
    ARS_API_END
}

// ---- block-sharded long render (mask-free): building blocks for one rank of a multi-GPU render ----
int64_t ars_state_bytes(void) { return (int64_t)sizeof(RenderState); }
int64_t ars_ols_block_frames(void) { return (int64_t)1 << (g_opt_upols_logf - 1); }

// the filter of a mask-free long render (external stereo IR, or early + late parts of equal role)
static void long_filter_spec(const ArsRenderParams* p, i64 n_total, i64 L0, bool have_ir1, i64 L1, FilterSpec& fs, i64* L_out) {
    const i64 L = p->external_ir ? L0 : std::max<i64>(L0, have_ir1 ? L1 : 0);
    const i64 N = n_total + L - 1;
    common_filter_spec(fs, N, p->rate, p->dry_wet, p->kill_start, p->bass_gain, p->treble_gain);
    if (p->external_ir) {
        fs.mode = FILT_EXT;
    } else {
        fs.mode = FILT_SPLIT;
        fs.level0 = (L0 > 1 && p->early_level > 1e-6) ? p->early_level : 0.0;
        fs.level1 = (have_ir1 && L1 > 1 && p->late_level > 1e-6) ? p->late_level : 0.0;
        if (p->air_absorption > 0.01 && N >= 2) fill_air(fs, N, p->rate, p->air_absorption);
    }
    ARS_CHECK(upols_applicable(fs), "long render: EQ / air absorption need the global N-point transform; only mask-free "
                                    "renders shard by block ranges (use ars_render on one GPU)");
    *L_out = L;
}

int ars_long_plan(const ArsRenderParams* p, int64_t n_total, int64_t ir_len, ArsLongPlan* out) {
    ARS_API_BEGIN
    ARS_CHECK(p && out && n_total > 0 && ir_len >= 1, "ars_long_plan: bad arguments");
    FilterSpec fs;
    i64 L = 0;
    long_filter_spec(p, n_total, ir_len, !p->external_ir, ir_len, fs, &L);
    memset(out, 0, sizeof(*out));
    OlsbPlan pl;
    if (g_opt_upols && olsb_plan(fs.N, L, 0, 0, &pl)) {
        out->route = 1;
        out->block_frames = pl.hop;
        out->n_blocks = pl.J;
        out->halo_frames = pl.skip;
    } else {
        const i64 B = (i64)1 << (g_opt_upols_logf - 1);
        out->route = 0;
        out->block_frames = B;
        out->n_blocks = (fs.N + B - 1) / B;
        out->halo_frames = ((L + B - 1) / B) * B;
    }
    out->frames_out = fs.N;
    out->hop_count = loudness_hop_count(fs.N, p->rate);
    ARS_API_END
}

int ars_long_convolve_dev(const ArsRenderParams* p, const float* d_x, int64_t x_frame0, int64_t x_frames, int64_t n_total,
                          int32_t cin, const float* d_ir0, int64_t L0, const float* d_ir1, int64_t L1,
                          int64_t block_lo, int64_t block_hi, float* d_y, int64_t y_frame0, void* d_state) {
    ARS_API_BEGIN
    ARS_CHECK(p && d_x && d_y && d_state && n_total > 0 && cin >= 1 && d_ir0 && L0 >= 1, "ars_long_convolve_dev: bad arguments");
    FilterSpec fs;
    i64 L = 0;
    long_filter_spec(p, n_total, L0, d_ir1 != nullptr, L1, fs, &L);
    OlsRange rg;
    rg.block_lo = block_lo;
    rg.block_hi = block_hi;
    rg.x_frame0 = x_frame0;
    rg.x_frames = x_frames;
    rg.y_frame0 = y_frame0;
    OlsbPlan pl;
    if (g_opt_upols && olsb_plan(fs.N, L, 0, 0, &pl))       // (the same decision ars_long_plan reports)
        olsb_filter(d_x, n_total, cin, d_ir0, L0, p->external_ir ? nullptr : d_ir1, L1, fs, reinterpret_cast<float2*>(d_y),
                    static_cast<RenderState*>(d_state), pl, rg);
    else
        upols_filter(d_x, n_total, cin, d_ir0, L0, p->external_ir ? nullptr : d_ir1, L1, fs, reinterpret_cast<float2*>(d_y),
                     static_cast<RenderState*>(d_state), g_opt_upols_logf, rg);
    ARS_API_END
}

int ars_long_tail_dev(const ArsRenderParams* p, int32_t phase, const float* d_y, int64_t y_frame0, int64_t frame_lo,
                      int64_t frame_hi, int64_t N_total, void* d_state, float* d_out_f32, int16_t* d_out_pcm,
                      float* d_mono) {
    ARS_API_BEGIN
    ARS_CHECK(p && d_y && d_state && layout_ok(p->layout) && frame_lo >= 0 && frame_hi <= N_total && phase >= 0 && phase <= 2,
              "ars_long_tail_dev: bad arguments");
    TailSpec ts = make_tail(N_total, p->layout, p->rate, p->x, p->y, p->z);
    ts.i_lo = frame_lo;
    ts.i_hi = frame_hi;
    ts.y0 = y_frame0;
    ts.out0 = frame_lo;
    ARS_CHECK(y_frame0 <= std::max<i64>(0, frame_lo - ts.delay), "ars_long_tail_dev: y slice lacks the delay halo");
    const float2* y = reinterpret_cast<const float2*>(d_y);
    RenderState* st = static_cast<RenderState*>(d_state);
    if (phase == 0) tail_pan_max(y, ts, st);
    else if (phase == 1) tail_map_max(y, ts, st);
    else tail_final(y, ts, st, d_out_f32, reinterpret_cast<short*>(d_out_pcm), d_mono);
    ARS_API_END
}

int ars_long_loudness_hops_dev(const ArsRenderParams* p, const float* d_y, int64_t y_frame0, int64_t frame_lo,
                               int64_t frame_hi, int64_t N_total, void* d_state, double* d_hops, int32_t n_hops) {
    ARS_API_BEGIN
    ARS_CHECK(p && d_y && d_state && d_hops && layout_ok(p->layout) && frame_lo >= 0 && frame_hi <= N_total,
              "ars_long_loudness_hops_dev: bad arguments");
    TailSpec ts = make_tail(N_total, p->layout, p->rate, p->x, p->y, p->z);
    ts.y0 = y_frame0;
    loudness_hops_from_stage(reinterpret_cast<const float2*>(d_y), ts, p->rate, static_cast<RenderState*>(d_state), frame_lo,
                             frame_hi, d_hops, n_hops);
    ARS_API_END
}

int ars_long_loudness_gate_dev(const ArsRenderParams* p, const double* d_hops, int32_t n_hops, int64_t N_total, void* d_state,
                               int32_t* lufs_status) {
    ARS_API_BEGIN
    ARS_CHECK(p && d_hops && d_state && lufs_status, "ars_long_loudness_gate_dev: bad arguments");
    *lufs_status = loudness_gate_from_hops(d_hops, n_hops, N_total, p->rate, static_cast<RenderState*>(d_state)) == 0
                       ? ARS_LUFS_OK : ARS_LUFS_NONE;
    ARS_API_END
}

int ars_loudness_dev(const float* d_mono, int64_t N, double rate, void* d_state, int32_t* lufs_status) {
    ARS_API_BEGIN
    ARS_CHECK(d_mono && d_state && N > 0 && rate > 0 && lufs_status, "ars_loudness_dev: bad arguments");
    *lufs_status = lufs_enqueue(d_mono, N, rate, static_cast<RenderState*>(d_state));
    ARS_API_END
}

int ars_state_metrics(const void* d_state, int64_t sample_count, int32_t lufs_status, ArsMetrics* out) {
    ARS_API_BEGIN
    ARS_CHECK(d_state && out, "ars_state_metrics: bad arguments");
    RenderState h;
    download(&h, static_cast<const RenderState*>(d_state), 1);
    sync();
    finish_metrics(h, sample_count, lufs_status, out);
    ARS_API_END
}

// ---- whole-render array shared between the ranks of a node (CUDA IPC), filled by peer pushes over NVLink ----
int ars_peer_alloc(int64_t bytes, void** d_ptr, unsigned char* handle) {
    ARS_API_BEGIN
    ARS_CHECK(bytes > 0 && d_ptr && handle, "ars_peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == ARS_PEER_HANDLE_BYTES, "CUDA IPC handle size");
    void* p = nullptr;
    ARS_CUDA(cudaMalloc(&p, (size_t)bytes));           // (its own allocation: an IPC handle names a whole cudaMalloc block)
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); ARS_CUDA(e); }
    memcpy(handle, &h, sizeof(h));
    *d_ptr = p;
    ARS_API_END
}
int ars_peer_open(const unsigned char* handle, void** d_ptr) {
    ARS_API_BEGIN
    ARS_CHECK(handle && d_ptr, "ars_peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    ARS_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *d_ptr = p;
    ARS_API_END
}
int ars_peer_push(void* d_peer_dst, const void* d_src, int64_t bytes, void* stream) {
    ARS_API_BEGIN
    ARS_CHECK(d_peer_dst && d_src && bytes >= 0, "ars_peer_push: bad arguments");
    if (bytes > 0)
        ARS_CUDA(cudaMemcpyAsync(d_peer_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToDevice,
                                 stream ? static_cast<cudaStream_t>(stream) : ctx().stream));
    ARS_API_END
}
int ars_peer_close(void* d_ptr) {
    ARS_API_BEGIN
    if (d_ptr) ARS_CUDA(cudaIpcCloseMemHandle(d_ptr));
    ARS_API_END
}
int ars_peer_free(void* d_ptr) {
    ARS_API_BEGIN
    if (d_ptr) ARS_CUDA(cudaFree(d_ptr));
    ARS_API_END
}

int ars_render_batch(const ArsClip* clips, int32_t count) {
    ARS_API_BEGIN
    ARS_CHECK(clips && count >= 0, "ars_render_batch: bad arguments");
    if (count == 0) return ARS_OK;
    Ctx& c = ctx();
    Pipe& pp = g_pipe;
    pp.init();
    RenderState* h_states = pp.states((size_t)count);
    std::vector<int> lufs_status((size_t)count, ARS_LUFS_SKIPPED);
    std::vector<i64> out_count((size_t)count, 0);
    // everything queued earlier on the compute stream must be finished before the copy streams touch buffers
    sync();
    for (int i = 0; i < count; ++i) {
        const ArsClip& k = clips[i];
        const ArsRenderParams* p = k.params;
        ARS_CHECK(p && k.in && k.n > 0 && k.cin >= 1 && layout_ok(p->layout), "ars_render_batch: bad clip");
        const int slot = i & 1;
        const i64 N = render_out_len(p, k.n, k.ext_ir_len);
        const int C = layout_channels(p->layout);
        out_count[(size_t)i] = N * C;
        // ---- host -> device on the copy stream (the slot's inputs are free once clip i-2 has been computed)
        if (i >= 2) ARS_CUDA(cudaStreamWaitEvent(pp.h2d, pp.done[slot], 0));
        float* d_in = c.buf(slot_name("in.x", slot), sizeof(float) * (size_t)k.n * k.cin).as<float>();
        host_upload(d_in, k.in, sizeof(float) * (size_t)k.n * k.cin, pp.h2d);      // (pageable clips: through the pinned ring)
        const float* d_ir = nullptr;
        ArsIrDraws dd;
        memset(&dd, 0, sizeof dd);
        if (p->external_ir) {
            ARS_CHECK(k.ext_ir && k.ext_ir_len >= 1, "ars_render_batch: external IR missing");
            float* d = c.buf(slot_name("in.ir", slot), sizeof(float) * (size_t)k.ext_ir_len * 2).as<float>();
            ARS_CUDA(cudaMemcpyAsync(d, k.ext_ir, sizeof(float) * (size_t)k.ext_ir_len * 2, cudaMemcpyHostToDevice, pp.h2d));
            d_ir = d;
        } else if (k.draws) {
            dd = *k.draws;
            const size_t nl = (size_t)std::max<i64>(0, k.draws->noise_len);
            double* d = c.buf(slot_name("ir.noise", slot), sizeof(double) * std::max<size_t>(nl, 1)).as<double>();
            if (nl) ARS_CUDA(cudaMemcpyAsync(d, k.draws->noise, sizeof(double) * nl, cudaMemcpyHostToDevice, pp.h2d));
            dd.noise = d;
        }
        ARS_CUDA(cudaEventRecord(pp.in_ready[slot], pp.h2d));
        // ---- compute (waits for its inputs, and for the slot's previous outputs to have left the device)
        float* d_f = k.out_f32 ? c.buf(slot_name("out.f32", slot), sizeof(float) * (size_t)N * C).as<float>() : nullptr;
        short* d_p = k.out_pcm ? c.buf(slot_name("out.pcm", slot), sizeof(short) * (size_t)N * C).as<short>() : nullptr;
        ARS_CUDA(cudaStreamWaitEvent(c.stream, pp.in_ready[slot], 0));
        if (i >= 2) ARS_CUDA(cudaStreamWaitEvent(c.stream, pp.out_free[slot], 0));
        RenderState* st = fresh_state(slot);
        render_core(p, d_in, k.n, k.cin, d_ir, k.ext_ir_len, p->external_ir ? nullptr : &dd, nullptr, d_f, d_p, st,
                    k.metrics != nullptr, &lufs_status[(size_t)i]);
        ARS_CUDA(cudaEventRecord(pp.done[slot], c.stream));
        // ---- device -> host on the other copy stream
        ARS_CUDA(cudaStreamWaitEvent(pp.d2h, pp.done[slot], 0));
        if (k.out_f32) host_download(k.out_f32, d_f, sizeof(float) * (size_t)N * C, pp.d2h);
        if (k.out_pcm) host_download(k.out_pcm, d_p, sizeof(short) * (size_t)N * C, pp.d2h);
        if (k.metrics) ARS_CUDA(cudaMemcpyAsync(&h_states[i], st, sizeof(RenderState), cudaMemcpyDeviceToHost, pp.d2h));
        ARS_CUDA(cudaEventRecord(pp.out_free[slot], pp.d2h));
    }
    ARS_CUDA(cudaStreamSynchronize(pp.d2h));
    ARS_CUDA(cudaStreamSynchronize(c.stream));
    for (int i = 0; i < count; ++i)
        if (clips[i].metrics)
            finish_metrics(h_states[i], out_count[(size_t)i], lufs_status[(size_t)i], clips[i].metrics, clips[i].params);
    ARS_API_END
}

}  // extern "C"
