// FFT plans (pass decomposition + twiddle tables) and the launchers for fft.cuh.
#include "fft_launch.cuh"

#include <cmath>
#include <cstdlib>

namespace ars {

using namespace fft;


// ----------------------------------------------------------------- tables ----
__global__ void twiddle_table_kernel(float2* out, int count, double step_turns) {
    // out[e] = exp(-2 pi i * e * step_turns), evaluated in double
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= count) return;
    double s, c;
    sincospi(-2.0 * (double)e * step_turns, &s, &c);
    out[e] = make_float2((float)c, (float)s);
}

// per-stage local twiddles, shared by every plan: for Ls = 2^l, [stage_off(l) + j*(Ls/2) + i] = w_Ls^(i * 2^j)
__global__ void stage_table_kernel(float2* out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= STAGE_TABLE_ELEMS) return;
    int l = 1;
    while (l < STAGE_LOG_MAX && idx >= stage_off(l + 1)) ++l;
    const int Ls = 1 << l, rel = idx - stage_off(l);
    const int j = rel / (Ls / 2), i = rel % (Ls / 2);
    const long long e = ((long long)i << j) % Ls;
    double s, c;
    sincospi(-2.0 * (double)e / (double)Ls, &s, &c);
    out[idx] = make_float2((float)c, (float)s);
}

// lane-dependent factors of a strided pass's inter-pass twiddle (see PassArgs::ptab)
__global__ void pass_table_kernel(float2* out, int logLg, int mul, int logT, int count) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count) return;
    const int row = idx >> logT, c = idx & ((1 << logT) - 1);
    const long long x = row < 4 ? ((long long)mul << row) : (long long)(row - 4);
    const long long e = ((long long)c * x) & (((long long)1 << logLg) - 1);
    double s, cs;
    sincospi(-2.0 * (double)e / (double)((long long)1 << logLg), &s, &cs);
    out[idx] = make_float2((float)cs, (float)s);
}

static DevBuf g_tw_local;

static const float2* local_table() {
    if (!g_tw_local.p) {
        g_tw_local.reserve(sizeof(float2) * STAGE_TABLE_ELEMS);
        stage_table_kernel<<<ceil_div(STAGE_TABLE_ELEMS, 256), 256, 0, ctx().stream>>>(g_tw_local.as<float2>());
        ARS_LAUNCH_CHECK();
        count_launch();
    }
    return g_tw_local.as<float2>();
}

void fft_release_plans() {
    if (ctx_ready()) {
        for (auto& kv : ctx().fft_plans) {
            kv.second->tw_lo.release();
            kv.second->tw_hi.release();
            for (auto& t : kv.second->pass_tabs) t.release();
            kv.second->rho.release();
            delete kv.second;
        }
        ctx().fft_plans.clear();
    }
    g_tw_local.release();
}

}  // namespace ars
#include "fft_decompose.inc"
namespace ars {

static bool g_fast = true;     // ARS_FFT_GENERIC=1 forces the runtime-switch kernels (debug / A-B runs)

FftPlan* get_fft_plan(int logM) {
    Ctx& c = ctx();
    g_fast = env_int("ARS_FFT_GENERIC", 0) == 0;
    auto it = c.fft_plans.find(logM);
    if (it != c.fft_plans.end()) return it->second;
    ARS_CHECK(logM >= 1 && logM <= 30, "FFT length out of range (2^1 .. 2^30)");
    FftPlan* p = new FftPlan();
    p->logM = logM;
    p->M = (i64)1 << logM;
    p->passes = fft_decompose(logM);
    const int nlo = (int)std::min<i64>(p->M, (i64)1 << BIG_LO_LOG);
    p->tw_lo.reserve(sizeof(float2) * nlo);
    twiddle_table_kernel<<<ceil_div(nlo, 256), 256, 0, c.stream>>>(p->tw_lo.as<float2>(), nlo, 1.0 / (double)p->M);
    ARS_LAUNCH_CHECK();
    count_launch();
    p->tw.lo = p->tw_lo.as<float2>();
    p->tw.hi = nullptr;
    if (logM > BIG_LO_LOG) {
        const int nhi = 1 << (logM - BIG_LO_LOG);
        p->tw_hi.reserve(sizeof(float2) * nhi);
        twiddle_table_kernel<<<ceil_div(nhi, 256), 256, 0, c.stream>>>(p->tw_hi.as<float2>(), nhi,
                                                                      (double)(1 << BIG_LO_LOG) / (double)p->M);
        ARS_LAUNCH_CHECK();
        count_launch();
        p->tw.hi = p->tw_hi.as<float2>();
    }
    p->tw.stage = local_table();
    p->pass_tabs.resize(p->passes.size());
    for (size_t i = 0; i < p->passes.size(); ++i) {
        FftPass& ps = p->passes[i];
        if (!ps.strided) continue;
        const int count = pass_table_elems(ps.logR, ps.logT);
        p->pass_tabs[i].reserve(sizeof(float2) * (size_t)count);
        pass_table_kernel<<<ceil_div(count, 256), 256, 0, c.stream>>>(p->pass_tabs[i].as<float2>(), ps.logLg,
                                                                     pass_table_mul(ps.logR), ps.logT, count);
        ARS_LAUNCH_CHECK();
        count_launch();
        ps.ptab = p->pass_tabs[i].as<float2>();
    }
    // a plan may be created while the library is enqueuing on its side stream and be used from the main stream
    // right after: finish the tables here, once per plan, instead of tracking which stream built them
    ARS_CUDA(cudaStreamSynchronize(c.stream));
    c.fft_plans[logM] = p;
    return p;
}

// ---------------------------------------------------------------- profiling ---
// Per-launch CUDA-event timing of the pass kernels goes through the library-wide KernelScope (common.cu); the names
// start with "fft:" so that ars_profile_end can report the pass kernels' totals as before.
static double ld_bytes(const Ld& ld, i64 M) {
    switch (ld.mode) {
        case LD_PLAIN: return 8.0 * (double)M;
        case LD_MULSPEC: return 16.0 * (double)M;
        case LD_CHIRP_X2: return 16.0 * (double)ld.nvalid;
        case LD_CHIRP_XC: return (8.0 + 4.0 * std::min(ld.cin, 2)) * (double)ld.nvalid;
        case LD_CHIRP_PAIR: return 4.0 * (double)(ld.nvalid + ld.nvalid1) + 8.0 * (double)std::max(ld.nvalid, ld.nvalid1);
        case LD_CHIRP_B: return 8.0 * (double)(2 * ld.N - 1);
        case LD_CHIRP_C: return 16.0 * (double)ld.nvalid;
        case LD_REAL_PAIR: return 4.0 * (double)(ld.nvalid + ld.nvalid1);
        case LD_OLS_X: case LD_OLS_X2: return 8.0 * (double)ld.nvalid;
        case LD_OLS_IR: case LD_OLS_IR2: return 4.0 * (double)(ld.nvalid + ld.nvalid1);
        case LD_OLS_MAC: return 8.0 * (double)M;       // compulsory: each delay-line spectrum once
        case LD_OLS_CHIRPSIG: return 8.0 * (double)ld.N;
        case LD_OLS_IRC: return 4.0 * (double)(ld.nvalid + ld.nvalid1) + 8.0 * (double)std::max(ld.nvalid, ld.nvalid1);
        case LD_OLSB_X: return 8.0 * (double)M;        // (the windows overlap: an upper bound of the frames one launch reads)
        case LD_TAPS: return 4.0 * (double)(ld.nvalid + ld.nvalid1);
    }
    return 0.0;
}
static double st_bytes(const St& st, i64 M) {
    switch (st.mode) {
        case ST_PLAIN: case ST_SCALE: return 8.0 * (double)M;
        case ST_CHIRP: case ST_FINAL: return 16.0 * (double)st.N;
        case ST_OLS: case ST_OLS2: return 16.0 * (double)st.N;
        case ST_OLS_CHIRP: return 16.0 * (double)st.N;
        case ST_OLSB: return 16.0 * (double)M;
    }
    return 0.0;
}

void fft_profile_begin() { prof_session_begin(); }
// -> launches, total milliseconds, algorithmic bytes of the pass kernels
void fft_profile_end(long long* launches, double* ms, double* bytes) {
    prof_session_end("fft:", launches, ms, bytes, nullptr);
}
static const char* pass_name(const Ld& ld, const St& st, bool inv, bool strided) {
    if (ld.mode == LD_OLSB_X) return "fft:olsb first pass (strided forward, signal windows in)";
    if (st.mode == ST_OLSB) return "fft:olsb last pass (strided inverse, stereo frames + maxima out)";
    if (ld.mode == LD_OLS_X || ld.mode == LD_OLS_X2) return "fft:ols delay-line transform";
    if (st.mode == ST_OLS || st.mode == ST_OLS2) return "fft:ols inverse transform";
    if (ld.mode == LD_TAPS || ld.mode == LD_OLS_IR || ld.mode == LD_OLS_IR2 || ld.mode == LD_OLS_IRC) return "ir spectrum pass";
    if (st.mode == ST_SCALE) return "ir spectrum pass";
    if (strided) return inv ? "fft:strided inverse pass" : "fft:strided forward pass";
    return inv ? "fft:contiguous inverse pass" : "fft:contiguous forward pass";
}

// --------------------------------------------------------------- launchers ---
using namespace fftk;

template <bool INV>
static void launch_pass(const FftPlan* p, const FftPass& ps, const Ld& ld, const St& st, i64 total = 0) {
    PassArgs pa;
    pa.total = total;
    pa.M = p->M;
    pa.logM = p->logM;
    pa.logLg = ps.logLg;
    pa.tw = p->tw;
    static const int pf = env_int("ARS_FFT_PREFETCH", 0);
    static const int pf_x = env_int("ARS_OLS_PREFETCH", 0);        // delay-line transform: tiles ahead (measured: 148 -> -2 %, 296 / 592 -> +1..4 %; off)
    pa.prefetch = ld.mode == LD_OLS_X ? pf_x : pf;
    pa.ptab = ps.ptab;
    const i64 pts = total > 0 ? total : p->M;
    KernelScope prof_scope(pass_name(ld, st, INV, ps.strided), prof_session_on() ? ld_bytes(ld, pts) + st_bytes(st, pts) : 0.0);
    if (ps.strided) ARS_CHECK(ps.logLg - ps.logR >= ps.logT, "strided pass narrower than its tile");
    bool done = false;
    if (g_fast) {
        if (ps.strided) done = INV ? fast_strided_inv(ps, ld, st, pa) : fast_strided_fwd(ps, ld, st, pa);
        else done = INV ? fast_contig_inv(ps, ld, st, pa) : fast_contig_fwd(ps, ld, st, pa);
    }
    if (!done) {
        ARS_CHECK(st.mode != ST_OLS2, "ST_OLS2 has compile-time-mode instantiations only");
        if (ps.strided) done = INV ? generic_strided_inv(ps, ld, st, pa) : generic_strided_fwd(ps, ld, st, pa);
        else done = INV ? generic_contig_inv(ps, ld, st, pa) : generic_contig_fwd(ps, ld, st, pa);
    }
    ARS_CHECK(done, "no FFT pass kernel for this (logR, logT)");
}

namespace fftk {
int ols_threads() {
    static const int nt = env_int("ARS_OLS_NT", 512) == 256 ? 256 : 512;
    return nt;
}
}  // namespace fftk

int fft_segment_tile(int logF) { return logF == 12 ? 2 : 1; }

void fft_segments(int logF, i64 nseg, const Ld& ld, const St& st, bool inverse) {
    ARS_CHECK(logF == 12 || logF == 13, "fft_segments: segment length must be 2^12 or 2^13");
    ARS_CHECK(nseg > 0 && nseg % fft_segment_tile(logF) == 0, "fft_segments: segment count not a multiple of the tile");
    FftPlan tmp;                       // a plan-less pass: only the stage tables are needed
    tmp.logM = 0;
    tmp.M = nseg << logF;
    tmp.tw.stage = local_table();
    tmp.tw.lo = tmp.tw.hi = nullptr;
    const FftPass ps = {false, logF, logF == 12 ? 1 : 0, logF};
    if (inverse) launch_pass<true>(&tmp, ps, ld, st);
    else launch_pass<false>(&tmp, ps, ld, st);
}

void fft_touch_tables() { local_table(); }

void fft_segments_r2(i64 nseg, Ld ld, St st, bool inverse) {
    ARS_CHECK(nseg > 0, "fft_segments_r2: no segments");
    FftPlan tmp;
    tmp.logM = 0;
    tmp.M = nseg << 13;
    tmp.tw.stage = local_table();
    tmp.tw.lo = tmp.tw.hi = nullptr;
    ld.logF = st.logF = 13;
    ld.tw2 = st.tw2 = local_table() + stage_off(13);      // w_8192^i, i < 4096
    const FftPass ps = {false, 12, 1, 12};
    if (inverse) launch_pass<true>(&tmp, ps, ld, st);
    else launch_pass<false>(&tmp, ps, ld, st);
}

// ---- batched two-pass transforms of the big-block overlap-save route (upols.cu) ----
// `nbatch` independent M-point transforms laid out back to back in one buffer; the plan must be strided + contiguous(12).
static void check_two_pass(const FftPlan* p) {
    ARS_CHECK(p->passes.size() == 2 && p->passes[0].strided && !p->passes[1].strided && p->passes[1].logR == 12 &&
                  p->passes[1].logT == 1,
              "big-block overlap-save needs a two-pass plan (strided + 2 x 4096 contiguous)");
}
void fft_batch_first(FftPlan* p, i64 nbatch, const Ld& ld, const St& st) {
    check_two_pass(p);
    launch_pass<false>(p, p->passes[0], ld, st, nbatch * p->M);
}
void fft_batch_last(FftPlan* p, i64 nbatch, const Ld& ld, const St& st) {
    check_two_pass(p);
    const FftPass& ps = p->passes[0];
    int logT = 0;
    if (g_fast && ld.mode == LD_PLAIN && st.mode == ST_OLSB && last_pass_pipe(ps.logR, &logT) && ps.logLg - ps.logR >= logT) {
        if (!p->pipe_tab.p) {        // inter-pass twiddle factors for the pipelined kernel's tile width
            const int count = pass_table_elems(ps.logR, logT);
            p->pipe_tab.reserve(sizeof(float2) * (size_t)count);
            pass_table_kernel<<<ceil_div(count, 256), 256, 0, ctx().stream>>>(p->pipe_tab.as<float2>(), ps.logLg,
                                                                             pass_table_mul(ps.logR), logT, count);
            ARS_LAUNCH_CHECK();
            count_launch();
            ARS_CUDA(cudaStreamSynchronize(ctx().stream));      // (once per plan; later calls may come from a lane)
        }
        PassArgs pa;
        pa.total = nbatch * p->M;
        pa.M = p->M;
        pa.logM = p->logM;
        pa.logLg = ps.logLg;
        pa.tw = p->tw;
        pa.prefetch = 0;
        pa.ptab = p->pipe_tab.as<float2>();
        KernelScope prof_scope(pass_name(ld, st, true, true), prof_session_on() ? ld_bytes(ld, pa.total) + st_bytes(st, pa.total) : 0.0);
        last_pass_pipe_launch(ps.logR, ld, st, pa);
        return;
    }
    launch_pass<true>(p, ps, ld, st, nbatch * p->M);
}
void fft_batch_mid(FftPlan* p, i64 nbatch, float2* work, const float2* h0, const float2* h1) {
    check_two_pass(p);
    const int logR1 = p->passes[0].logR;
    MidArgs ma;
    ma.h0 = h0;
    ma.h1 = h1;
    ma.fmask = p->M - 1;
    ma.logR1 = logR1;
    if (h1) {
        if (!p->rho.p) {            // segment of every first-pass digit (the mirror tiles pair digit k1 with R1 - k1)
            std::vector<int> rho((size_t)1 << logR1);
            for (int k1 = 0; k1 < (1 << logR1); ++k1) rho[(size_t)k1] = strided_row_of(logR1, k1);
            p->rho.reserve(sizeof(int) * rho.size());
            ARS_CUDA(cudaMemcpyAsync(p->rho.p, rho.data(), sizeof(int) * rho.size(), cudaMemcpyHostToDevice, ctx().stream));
            ARS_CUDA(cudaStreamSynchronize(ctx().stream));       // (once per plan: the host vector dies here)
        }
        ma.rho = p->rho.as<int>();
    }
    PassArgs pa;
    pa.total = nbatch * p->M;
    pa.M = p->M;
    pa.logM = p->logM;
    pa.logLg = 12;
    pa.prefetch = 0;
    pa.ptab = nullptr;
    pa.tw = p->tw;
    Ld ld;
    ld.mode = LD_PLAIN;
    ld.a = work;
    St st;
    st.mode = ST_PLAIN;
    st.a = work;
    KernelScope prof_scope("fft:olsb middle pass (contiguous forward x IR spectrum x contiguous inverse)",
                           (h1 ? 16.0 : 16.0) * (double)pa.total);
    mid_pass(h1 != nullptr, ld, st, pa, ma);
}

void fft_forward(FftPlan* p, const Ld& ld_first, float2* work, const St& st_last) {
    const int np = (int)p->passes.size();
    for (int i = 0; i < np; ++i) {
        Ld ld;
        St st;
        if (i == 0) ld = ld_first; else { ld.mode = LD_PLAIN; ld.a = work; }
        if (i == np - 1) st = st_last; else { st.mode = ST_PLAIN; st.a = work; }
        launch_pass<false>(p, p->passes[i], ld, st);
    }
}

void fft_inverse(FftPlan* p, const Ld& ld_first, float2* work, const St& st_last) {
    const int np = (int)p->passes.size();
    for (int i = np - 1; i >= 0; --i) {
        Ld ld;
        St st;
        if (i == np - 1) ld = ld_first; else { ld.mode = LD_PLAIN; ld.a = work; }
        if (i == 0) st = st_last; else { st.mode = ST_PLAIN; st.a = work; }
        launch_pass<true>(p, p->passes[i], ld, st);
    }
}

}  // namespace ars
