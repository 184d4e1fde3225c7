// Host emulation of the FFT pass kernels (no GPU needed).
//
// The stage code in ars_b200/csrc/fft.cuh is __host__ __device__; here its "threads"
// run one after another on the CPU so the index maps, twiddles, pass decomposition and
// the Bluestein wiring can be checked in the authoring container.  Checks:
//   1. IFFT(FFT(a) .* FFT(b)) / M == circular convolution (double-precision reference)
//   2. Bluestein DFT_N built from those transforms == direct DFT (double) for arbitrary N
// Usage: fft_emul <logM>...   |   fft_emul blue <N>...
// Exit code 0 when every max error is below 2e-5 (relative to the result's peak).
#include "../../ars_b200/csrc/fft.cuh"

#include <cmath>
#include <complex>
#include <cstdlib>
#include <random>

using namespace ars;
using namespace ars::fft;
typedef std::complex<double> cd;

static std::vector<float2> g_local, g_lo, g_hi;

static void make_tables(int logM, Tw& tw) {
    const double PI = 3.14159265358979323846;
    g_local.resize(STAGE_TABLE_ELEMS);
    for (int l = 1; l <= STAGE_LOG_MAX; ++l) {
        const int Ls = 1 << l;
        for (int j = 0; j < 4; ++j)
            for (int i = 0; i < Ls / 2; ++i) {
                const long long e = ((long long)i << j) % Ls;
                double a = -2.0 * PI * (double)e / (double)Ls;
                g_local[stage_off(l) + j * (Ls / 2) + i] = make_float2((float)cos(a), (float)sin(a));
            }
    }
    const i64 M = (i64)1 << logM;
    const int nlo = (int)std::min<i64>(M, (i64)1 << BIG_LO_LOG);
    g_lo.resize(nlo);
    for (int e = 0; e < nlo; ++e) {
        double a = -2.0 * PI * (double)e / (double)M;
        g_lo[e] = make_float2((float)cos(a), (float)sin(a));
    }
    tw.stage = g_local.data();
    tw.lo = g_lo.data();
    tw.hi = nullptr;
    if (logM > BIG_LO_LOG) {
        const int nhi = 1 << (logM - BIG_LO_LOG);
        g_hi.resize(nhi);
        for (int e = 0; e < nhi; ++e) {
            double a = -2.0 * PI * (double)e * (double)(1 << BIG_LO_LOG) / (double)M;
            g_hi[e] = make_float2((float)cos(a), (float)sin(a));
        }
        tw.hi = g_hi.data();
    }
}

constexpr int NT = 256;

template <int LOGR, int LOGT, bool INV, int LDM = -1, int STM = -1> static void emu_strided(const Ld& ld, const St& st, const PassArgs& pa) {
    using L = StridedLayout<LOGR, LOGT>;
    std::vector<float2> sm(L::SMEM_ELEMS);
    const i64 tiles = (pa.total > 0 ? pa.total : pa.M) >> (LOGR + LOGT);
    for (i64 tile = 0; tile < tiles; ++tile) {
        Ld l = ld;
        St s = st;
        StridedTile<LOGR, LOGT> t(tile, pa);
        emulate_tile<LOGR, INV, true, NT, L, LDM, STM>(sm.data(), l, s, pa, StridedFirst<LOGR>{t.base, t.logStride},
                                            StridedLast<LOGR>{t.base, t.logStride}, t.col0);
    }
}
template <int LOGR, int LOGC, bool INV, int LDM = -1, int STM = -1> static void emu_contig(const Ld& ld, const St& st, const PassArgs& pa) {
    using L = ContigLayout<LOGR, LOGC>;
    std::vector<float2> sm(L::SMEM_ELEMS);
    const i64 tiles = (pa.total > 0 ? pa.total : pa.M) >> (LOGR + LOGC);
    for (i64 tile = 0; tile < tiles; ++tile) {
        Ld l = ld;
        St s = st;
        const i64 base = tile << (LOGR + LOGC);
        emulate_tile<LOGR, INV, false, NT, L, LDM, STM>(sm.data(), l, s, pa, ContigFirst<LOGR>{base}, ContigLast<LOGR>{base}, 0u);
    }
}

template <bool INV> static void emu_pass(int logM, const Tw& tw, const FftPass& ps, const Ld& ld, const St& st, i64 total = 0) {
    PassArgs pa;
    pa.total = total;
    pa.M = (i64)1 << logM;
    pa.logM = logM;
    pa.logLg = ps.logLg;
    pa.prefetch = 0;
    pa.ptab = ps.ptab;
    pa.tw = tw;
    if (ps.strided) {
        if (ps.logLg - ps.logR < ps.logT) { printf("bad plan: strided pass narrower than tile\n"); exit(2); }
#define S_CASE(R, T) if (ps.logR == R && ps.logT == T) return emu_strided<R, T, INV>(ld, st, pa);
        ARS_STRIDED_CASES(S_CASE)
#undef S_CASE
    } else {
#define C_CASE(R, C) if (ps.logR == R && ps.logT == C) return emu_contig<R, C, INV>(ld, st, pa);
        ARS_CONTIG_CASES(C_CASE)
#undef C_CASE
    }
    printf("no kernel variant for pass (strided=%d logR=%d logT=%d)\n", (int)ps.strided, ps.logR, ps.logT);
    exit(2);
}

struct EmuPlan { int logM; std::vector<FftPass> passes; Tw tw; };

// host copy of pass_table_kernel (fft_plan.cu)
static std::vector<std::vector<float2>> g_pass_tabs;
static void attach_pass_tables(EmuPlan& p) {
    const double PI = 3.14159265358979323846;
    g_pass_tabs.clear();
    g_pass_tabs.resize(p.passes.size());
    for (size_t i = 0; i < p.passes.size(); ++i) {
        FftPass& ps = p.passes[i];
        if (!ps.strided) continue;
        const int count = pass_table_elems(ps.logR, ps.logT), mul = pass_table_mul(ps.logR);
        g_pass_tabs[i].resize(count);
        for (int idx = 0; idx < count; ++idx) {
            const int row = idx >> ps.logT, c = idx & ((1 << ps.logT) - 1);
            const long long x = row < 4 ? ((long long)mul << row) : (long long)(row - 4);
            const long long e = ((long long)c * x) & (((long long)1 << ps.logLg) - 1);
            const double a = -2.0 * PI * (double)e / (double)((long long)1 << ps.logLg);
            g_pass_tabs[i][idx] = make_float2((float)cos(a), (float)sin(a));
        }
        ps.ptab = g_pass_tabs[i].data();
    }
}

static void emu_forward(const EmuPlan& p, const Ld& ld_first, float2* work, const St& st_last) {
    const int np = (int)p.passes.size();
    for (int i = 0; i < np; ++i) {
        Ld ld; St st;
        if (i == 0) ld = ld_first; else { ld.mode = LD_PLAIN; ld.a = work; }
        if (i == np - 1) st = st_last; else { st.mode = ST_PLAIN; st.a = work; }
        emu_pass<false>(p.logM, p.tw, p.passes[i], ld, st);
    }
}
static void emu_inverse(const EmuPlan& p, const Ld& ld_first, float2* work, const St& st_last) {
    const int np = (int)p.passes.size();
    for (int i = np - 1; i >= 0; --i) {
        Ld ld; St st;
        if (i == np - 1) ld = ld_first; else { ld.mode = LD_PLAIN; ld.a = work; }
        if (i == 0) st = st_last; else { st.mode = ST_PLAIN; st.a = work; }
        emu_pass<true>(p.logM, p.tw, p.passes[i], ld, st);
    }
}

// double-precision radix-2 reference FFT
static void ref_fft(std::vector<cd>& a, bool inv) {
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    const double PI = 3.14159265358979323846;
    for (size_t len = 2; len <= n; len <<= 1) {
        double ang = 2 * PI / (double)len * (inv ? 1 : -1);
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                cd w(cos(ang * (double)k), sin(ang * (double)k));
                cd u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}

// the real planner (shared with fft_plan.cu)
#include "../../ars_b200/csrc/fft_decompose.inc"

static int check_conv(int logM) {
    EmuPlan p;
    p.logM = logM;
    p.passes = fft_decompose(logM);
    attach_pass_tables(p);
    make_tables(logM, p.tw);
    const i64 M = (i64)1 << logM;
    printf("logM=%d passes:", logM);
    for (auto& ps : p.passes) printf(" %s(R=2^%d,T=2^%d,Lg=2^%d)", ps.strided ? "S" : "C", ps.logR, ps.logT, ps.logLg);
    printf("\n");
    std::mt19937 rng(logM);
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    std::vector<float2> a(M), b(M), A(M), B(M), out(M);
    for (i64 i = 0; i < M; ++i) { a[i] = make_float2(U(rng), U(rng)); b[i] = make_float2(0.f, 0.f); }
    // b: a short random kernel so the reference result has O(1) dynamic range
    const int K = (int)std::min<i64>(M, 40);
    for (int i = 0; i < K; ++i) b[(i * 977 + 3) % M] = make_float2(U(rng), U(rng));
    Ld ld; St st;
    ld.mode = LD_PLAIN; ld.a = a.data(); st.mode = ST_PLAIN; st.a = A.data();
    emu_forward(p, ld, A.data(), st);
    ld.a = b.data(); st.a = B.data(); st.mode = ST_SCALE; st.scale = 1.0f / (float)M;
    emu_forward(p, ld, B.data(), st);
    ld.mode = LD_MULSPEC; ld.a = A.data(); ld.b = B.data();
    st.mode = ST_PLAIN; st.a = out.data();
    emu_inverse(p, ld, A.data(), st);
    std::vector<cd> ra(M), rb(M);
    for (i64 i = 0; i < M; ++i) { ra[i] = cd(a[i].x, a[i].y); rb[i] = cd(b[i].x, b[i].y); }
    ref_fft(ra, false); ref_fft(rb, false);
    for (i64 i = 0; i < M; ++i) ra[i] *= rb[i];
    ref_fft(ra, true);
    double maxerr = 0, peak = 0;
    for (i64 i = 0; i < M; ++i) {
        cd r = ra[i] / (double)M;
        peak = std::max(peak, std::abs(r));
        maxerr = std::max(maxerr, std::abs(r - cd(out[i].x, out[i].y)));
    }
    // forward spectrum must be a permutation of the true DFT: compare sorted magnitudes cheaply via sums
    printf("  conv: max err %.3e (peak %.3f) rel %.3e\n", maxerr, peak, maxerr / peak);
    return (maxerr / peak < 2e-5) ? 0 : 1;
}

static int check_bluestein(i64 N) {
    const int logM = std::max(1, next_pow2_log(2 * N - 1));
    const i64 M = (i64)1 << logM;
    EmuPlan p;
    p.logM = logM;
    p.passes = fft_decompose(logM);
    attach_pass_tables(p);
    make_tables(logM, p.tw);
    const double PI = 3.14159265358979323846;
    std::vector<float2> chirp(N);
    for (i64 n = 0; n < N; ++n) {
        i64 q = (n * n) % (2 * N);
        double a = -PI * (double)q / (double)N;
        chirp[n] = make_float2((float)cos(a), (float)sin(a));
    }
    std::vector<float2> Bs(M), W(M), Z(N);
    Ld ld; St st;
    ld.mode = LD_CHIRP_B; ld.b = chirp.data(); ld.N = N; ld.M = M;
    st.mode = ST_SCALE; st.a = Bs.data(); st.scale = 1.0f / (float)M;
    emu_forward(p, ld, Bs.data(), st);
    // input: interleaved stereo frames, n < N valid
    std::mt19937 rng((unsigned)N);
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    const i64 nv = N - N / 5;
    std::vector<float> x(2 * nv);
    for (auto& v : x) v = U(rng);
    ld = Ld(); ld.mode = LD_CHIRP_X2; ld.f0 = x.data(); ld.b = chirp.data(); ld.nvalid = nv; ld.N = N; ld.M = M;
    st = St(); st.mode = ST_PLAIN; st.a = W.data();
    emu_forward(p, ld, W.data(), st);
    ld = Ld(); ld.mode = LD_MULSPEC; ld.a = W.data(); ld.b = Bs.data();
    st = St(); st.mode = ST_CHIRP; st.a = Z.data(); st.chirp = chirp.data(); st.N = N;
    emu_inverse(p, ld, W.data(), st);
    // direct DFT on a sample of bins (all bins when N is small)
    double maxerr = 0, peak = 0;
    const i64 step = std::max<i64>(1, N / 97);
    for (i64 k = 0; k < N; k += step) {
        cd acc = 0;
        for (i64 n = 0; n < nv; ++n) {
            i64 e = (n * k) % N;
            double a = -2 * PI * (double)e / (double)N;
            acc += cd(x[2 * n], x[2 * n + 1]) * cd(cos(a), sin(a));
        }
        peak = std::max(peak, std::abs(acc));
        maxerr = std::max(maxerr, std::abs(acc - cd(Z[k].x, Z[k].y)));
    }
    printf("bluestein N=%lld M=2^%d: max err %.3e (peak %.2f) rel %.3e\n", (long long)N, logM, maxerr, peak, maxerr / peak);
    return (maxerr / peak < 2e-5) ? 0 : 1;
}

// Overlap-save wiring (upols.cu) on the host: IR partitions, delay line, fused MAC + inverse, block ranges.
static void emu_segments(int logF, i64 nseg, const Tw& tw, const Ld& ld, const St& st, bool inverse) {
    const FftPass ps = {false, logF, logF == 12 ? 1 : 0, logF};
    EmuPlan p;
    p.logM = 0;
    PassArgs pa;
    pa.M = nseg << logF; pa.logM = 0; pa.logLg = logF; pa.prefetch = 0; pa.ptab = nullptr; pa.tw = tw;
    if (logF == 12) { if (inverse) emu_contig<12, 1, true>(ld, st, pa); else emu_contig<12, 1, false>(ld, st, pa); }
    else { if (inverse) emu_contig<13, 0, true>(ld, st, pa); else emu_contig<13, 0, false>(ld, st, pa); }
    (void)ps;
}

// the radix-2-folded form of the 8192-point overlap-save transforms (fft_plan.cu: fft_segments_r2)
static void emu_segments_r2(i64 nseg, const Tw& tw, Ld ld, St st, bool inverse) {
    PassArgs pa;
    pa.M = nseg << 13; pa.logM = 0; pa.logLg = 12; pa.prefetch = 0; pa.ptab = nullptr; pa.tw = tw;
    ld.logF = st.logF = 13;
    ld.tw2 = st.tw2 = tw.stage + stage_off(13);
    if (!inverse && ld.mode == LD_OLS_IR2) emu_contig<12, 1, false, LD_OLS_IR2, ST_SCALE>(ld, st, pa);
    else if (!inverse) emu_contig<12, 1, false, LD_OLS_X2, ST_PLAIN>(ld, st, pa);
    else emu_contig<12, 1, true, LD_OLS_MAC, ST_OLS2>(ld, st, pa);
}
static bool g_r2 = false;     // check_ols / check_ols_circ: run the logF = 13 cases through the folded form

static int check_ols(int logF, i64 n, i64 L, bool ext, i64 block_lo, i64 block_hi) {
    Tw tw;
    make_tables(logF, tw);
    const i64 B = (i64)1 << (logF - 1), F = (i64)1 << logF;
    const int tile = logF == 12 ? 2 : 1;
    const i64 N = n + L - 1;
    std::mt19937 rng((unsigned)(n * 31 + L));
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    std::vector<float> x(2 * n), ir(2 * L);
    for (auto& v : x) v = U(rng);
    for (i64 i = 0; i < L; ++i) { float d = expf(-(float)i / (0.3f * L)); ir[2 * i] = U(rng) * d; ir[2 * i + 1] = ext ? U(rng) * d : ir[2 * i]; }
    // make one partition all-zero to exercise the skip flags
    if (L > 3 * B) for (i64 i = B; i < 2 * B; ++i) ir[2 * i] = ir[2 * i + 1] = 0.f;
    const int P = (int)((L + B - 1) / B), Ppad = ((P + tile - 1) / tile) * tile;
    const i64 nblk_all = (N + B - 1) / B;
    if (block_hi < 0 || block_hi > nblk_all) block_hi = nblk_all;
    const i64 seg0 = std::max<i64>(0, block_lo - (P - 1)), skip = block_lo - seg0;
    const i64 run = ((block_hi - block_lo + tile - 1) / tile) * tile;
    const i64 nseg = ((skip + run + tile - 1) / tile) * tile;
    const int nspec = ext ? 2 : 1;
    std::vector<float2> H((size_t)Ppad * F * nspec), X((size_t)nseg * F * nspec), y(N, make_float2(0, 0));
    std::vector<unsigned char> nz(Ppad, 0);
    for (int p = 0; p < P; ++p)
        for (i64 i = p * B; i < std::min(L, (p + 1) * B); ++i) if (ir[2 * i] != 0.f || ir[2 * i + 1] != 0.f) nz[p] = 1;
    for (int k = 0; k < nspec; ++k) {
        Ld ld; ld.mode = LD_OLS_IR; ld.logF = logF; ld.f0 = ir.data(); ld.f1 = ir.data() + 1; ld.cin = 2;
        ld.nvalid = ld.nvalid1 = L;
        if (ext) { ld.c0 = 0.5f; ld.c1 = k == 0 ? 0.5f : -0.5f; } else { ld.c0 = 1.f; ld.c1 = 0.f; }
        St st; st.mode = ST_SCALE; st.a = H.data() + (size_t)k * Ppad * F; st.scale = 1.0f / (float)F;
        if (g_r2 && logF == 13) { ld.mode = LD_OLS_IR2; emu_segments_r2(Ppad, tw, ld, st, false); } else emu_segments(logF, Ppad, tw, ld, st, false);
    }
    for (int k = 0; k < nspec; ++k) {
        Ld ld; ld.mode = LD_OLS_X; ld.logF = logF; ld.f0 = x.data(); ld.frame0 = 0; ld.nvalid = n; ld.cin = 2;
        ld.seg0 = seg0; ld.c1 = k == 0 ? 0.f : -1.f;
        St st; st.mode = ST_PLAIN; st.a = X.data() + (size_t)k * nseg * F;
        if (g_r2 && logF == 13) { ld.mode = LD_OLS_X2; emu_segments_r2(nseg, tw, ld, st, false); } else emu_segments(logF, nseg, tw, ld, st, false);
    }
    unsigned maxbits[4] = {0, 0, 0, 0};
    {
        Ld ld; ld.mode = LD_OLS_MAC; ld.logF = logF; ld.a = X.data() + skip * F; ld.b = H.data();
        ld.a2 = ext ? X.data() + (size_t)nseg * F + skip * F : nullptr; ld.b2 = ext ? H.data() + (size_t)Ppad * F : nullptr;
        ld.nz = nz.data(); ld.P = P; ld.lookback = skip;
        St st; st.mode = ST_OLS; st.logF = logF; st.seg0 = block_lo; st.a = y.data(); st.frame0 = 0;
        st.N = std::min<i64>(N, block_hi * B); st.dry = x.data(); st.dry_frame0 = 0; st.n = n; st.cin = 2;
        st.dg = 0.25f; st.dw = 0.5f; st.maxbits = maxbits;
        if (g_r2 && logF == 13) { st.mode = ST_OLS2; emu_segments_r2(run, tw, ld, st, true); } else emu_segments(logF, run, tw, ld, st, true);
    }
    double maxerr = 0, peak = 0;
    const i64 f_lo = block_lo * B, f_hi = std::min<i64>(N, block_hi * B);
    const i64 stepf = std::max<i64>(1, (f_hi - f_lo) / 400);
    for (i64 f = f_lo; f < f_hi; f += stepf) {
        double wl = 0, wr = 0;
        for (i64 m = std::max<i64>(0, f - n + 1); m < std::min<i64>(L, f + 1); ++m) {
            wl += (double)ir[2 * m] * x[2 * (f - m)];
            wr += (double)ir[2 * m + 1] * x[2 * (f - m) + 1];
        }
        const double rl = 0.25 * (f < n ? x[2 * f] : 0) + 0.5 * wl, rr = 0.25 * (f < n ? x[2 * f + 1] : 0) + 0.5 * wr;
        peak = std::max(peak, std::max(fabs(rl), fabs(rr)));
        maxerr = std::max(maxerr, std::max(fabs(rl - y[f].x), fabs(rr - y[f].y)));
    }
    printf("ols logF=%d n=%lld L=%lld ext=%d blocks[%lld,%lld): max err %.3e (peak %.2f) rel %.3e\n", logF, (long long)n,
           (long long)L, (int)ext, (long long)block_lo, (long long)block_hi, maxerr, peak, maxerr / peak);
    return (maxerr / peak < 2e-5 && maxbits[0] != 0) ? 0 : 1;
}

// The folded-air form of the overlap-save stage (upols.cu: upols_filter_airfold): taps that start at time -adv over the
// N-periodic extension of the zero-padded signal == the N-point circular convolution.
static int check_ols_circ(int logF, i64 n, i64 N, i64 Lf, i64 adv) {
    Tw tw;
    make_tables(logF, tw);
    const i64 B = (i64)1 << (logF - 1), F = (i64)1 << logF;
    const int tile = logF == 12 ? 2 : 1;
    std::mt19937 rng((unsigned)(n * 17 + Lf));
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    std::vector<float> x(2 * n), h(Lf);
    for (auto& v : x) v = U(rng);
    for (i64 i = 0; i < Lf; ++i) h[i] = U(rng) * expf(-fabsf((float)(i - adv)) / (0.2f * Lf));
    const int P = (int)((Lf + B - 1) / B), Ppad = ((P + tile - 1) / tile) * tile;
    const i64 nblk = (N + B - 1) / B;
    const i64 seg0 = -(P - 1), skip = P - 1;
    const i64 run = ((nblk + tile - 1) / tile) * tile;
    const i64 nseg = ((skip + run + tile - 1) / tile) * tile;
    std::vector<float2> H((size_t)Ppad * F), X((size_t)nseg * F), y(N, make_float2(0, 0));
    {
        Ld ld; ld.mode = LD_OLS_IR; ld.logF = logF; ld.f0 = h.data(); ld.f1 = nullptr; ld.cin = 1; ld.nvalid = Lf; ld.c0 = 1.f;
        St st; st.mode = ST_SCALE; st.a = H.data(); st.scale = 1.0f / (float)F;
        if (g_r2 && logF == 13) { ld.mode = LD_OLS_IR2; emu_segments_r2(Ppad, tw, ld, st, false); } else emu_segments(logF, Ppad, tw, ld, st, false);
    }
    {
        Ld ld; ld.mode = LD_OLS_X; ld.logF = logF; ld.f0 = x.data(); ld.frame0 = 0; ld.nvalid = n; ld.cin = 2;
        ld.seg0 = seg0; ld.adv = adv; ld.circ = N;
        St st; st.mode = ST_PLAIN; st.a = X.data();
        if (g_r2 && logF == 13) { ld.mode = LD_OLS_X2; emu_segments_r2(nseg, tw, ld, st, false); } else emu_segments(logF, nseg, tw, ld, st, false);
    }
    unsigned maxbits[4] = {0, 0, 0, 0};
    {
        Ld ld; ld.mode = LD_OLS_MAC; ld.logF = logF; ld.a = X.data() + skip * F; ld.b = H.data(); ld.P = P; ld.lookback = skip;
        St st; st.mode = ST_OLS; st.logF = logF; st.seg0 = 0; st.a = y.data(); st.frame0 = 0; st.N = N; st.dry = x.data();
        st.dry_frame0 = 0; st.n = n; st.cin = 2; st.dg = 0.25f; st.dw = 0.5f; st.maxbits = maxbits;
        if (g_r2 && logF == 13) { st.mode = ST_OLS2; emu_segments_r2(run, tw, ld, st, true); } else emu_segments(logF, run, tw, ld, st, true);
    }
    double maxerr = 0, peak = 0;
    const i64 stepf = std::max<i64>(1, N / 300);
    for (i64 q = 0; q < N + 40; ++q) {
        const i64 f = q < 20 ? q : (q < 40 ? N - 1 - (q - 20) : (q - 40));      // both ends densely, the rest sampled
        if (q >= 40 && (q - 40) % stepf) continue;
        double wl = 0, wr = 0;
        for (i64 m = 0; m < Lf; ++m) {
            i64 t = (f + adv - m) % N;
            if (t < 0) t += N;
            if (t < n) { wl += (double)h[m] * x[2 * t]; wr += (double)h[m] * x[2 * t + 1]; }
        }
        const double rl = 0.25 * (f < n ? x[2 * f] : 0) + 0.5 * wl, rr = 0.25 * (f < n ? x[2 * f + 1] : 0) + 0.5 * wr;
        peak = std::max(peak, std::max(fabs(rl), fabs(rr)));
        maxerr = std::max(maxerr, std::max(fabs(rl - y[f].x), fabs(rr - y[f].y)));
    }
    printf("ols-circ logF=%d n=%lld N=%lld Lf=%lld adv=%lld: max err %.3e (peak %.2f) rel %.3e\n", logF, (long long)n,
           (long long)N, (long long)Lf, (long long)adv, maxerr, peak, maxerr / peak);
    return (maxerr / peak < 2e-5 && maxbits[0] != 0) ? 0 : 1;
}

// Short-IR spectrum route (spectral.cu: ir_spectrum_short): P[k] = DFT_N(h0 + i h1) through overlap-save over the
// shifted Bluestein kernel.
static int check_irs(i64 N, i64 L) {
    constexpr int logF = 13, logB = 12;
    const i64 B = (i64)1 << logB, F = (i64)1 << logF;
    const double PI = 3.14159265358979323846;
    Tw tw;
    make_tables(logF, tw);
    std::vector<float2> chirp(N);
    for (i64 n = 0; n < N; ++n) {
        const i64 q = (n * n) % (2 * N);
        const double a = -PI * (double)q / (double)N;
        chirp[n] = make_float2((float)cos(a), (float)sin(a));
    }
    std::mt19937 rng((unsigned)(N + L));
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    std::vector<float> h0(L, 0.f), h1(L, 0.f);
    for (int j = 0; j < 30; ++j) h0[1 + (j * 131) % std::min<i64>(L - 1, 4000)] += U(rng);
    for (i64 i = std::min<i64>(L, 4500); i < std::min<i64>(L, 11000); ++i) h1[i] = U(rng) * expf(-(float)(i - 4500) / 900.f);
    const int P = (int)((L + B - 1) / B);
    const i64 D = (i64)P * B, s_lo = P, s_hi = P + (N + B - 1) / B;
    std::vector<float2> X((size_t)s_hi * F), H((size_t)P * F), Pout(N, make_float2(0, 0));
    std::vector<unsigned char> nz(P, 0);
    for (int p = 0; p < P; ++p)
        for (i64 i = p * B; i < std::min(L, (p + 1) * B); ++i) if (h0[i] != 0.f || h1[i] != 0.f) nz[p] = 1;
    { Ld ld; ld.mode = LD_OLS_CHIRPSIG; ld.logF = logF; ld.b = chirp.data(); ld.N = N; ld.frame0 = D; ld.seg0 = 0;
      St st; st.mode = ST_PLAIN; st.a = X.data(); emu_segments(logF, s_hi, tw, ld, st, false); }
    { Ld ld; ld.mode = LD_OLS_IRC; ld.logF = logF; ld.f0 = h0.data(); ld.nvalid = L; ld.f1 = h1.data(); ld.nvalid1 = L;
      ld.cin = 1; ld.b = chirp.data(); ld.N = N;
      St st; st.mode = ST_SCALE; st.a = H.data(); st.scale = 1.0f / (float)F; emu_segments(logF, P, tw, ld, st, false); }
    { Ld ld; ld.mode = LD_OLS_MAC; ld.logF = logF; ld.a = X.data() + s_lo * F; ld.b = H.data(); ld.nz = nz.data(); ld.P = P;
      ld.lookback = s_lo;
      St st; st.mode = ST_OLS_CHIRP; st.logF = logF; st.a = Pout.data(); st.chirp = chirp.data(); st.N = N; st.seg0 = s_lo;
      st.frame0 = D; emu_segments(logF, s_hi - s_lo, tw, ld, st, true); }
    double maxerr = 0, peak = 0;
    const i64 step = std::max<i64>(1, N / 61);
    for (i64 k = 0; k < N; k += step) {
        cd acc = 0;
        for (i64 n = 0; n < L; ++n) {
            if (h0[n] == 0.f && h1[n] == 0.f) continue;
            const i64 e = (n * k) % N;
            const double a = -2 * PI * (double)e / (double)N;
            acc += cd(h0[n], h1[n]) * cd(cos(a), sin(a));
        }
        peak = std::max(peak, std::abs(acc));
        maxerr = std::max(maxerr, std::abs(acc - cd(Pout[k].x, Pout[k].y)));
    }
    printf("irs N=%lld L=%lld: max err %.3e (peak %.2f) rel %.3e\n", (long long)N, (long long)L, maxerr, peak, maxerr / peak);
    return (maxerr / peak < 2e-5) ? 0 : 1;
}

// Big-block overlap-save wiring (upols.cu: olsb_filter): strided forward pass over the windows, fused middle pass
// (plain and mirror form), strided inverse pass with the dry/wet store; stripes, circular form, dry path in the taps or in the store.
static int check_olsb(int logF, i64 n, i64 Lf, bool ext, int cin, i64 adv, i64 circ, int stripe, bool dryfold = true) {
    EmuPlan p;
    p.logM = logF;
    p.passes = fft_decompose(logF);
    attach_pass_tables(p);
    make_tables(logF, p.tw);
    if (p.passes.size() != 2 || !p.passes[0].strided || p.passes[1].logR != 12 || p.passes[1].logT != 1) { printf("olsb: not a two-pass plan\n"); return 1; }
    const i64 F = (i64)1 << logF, skip = Lf - 1, hop = F - skip;
    const i64 N = circ > 0 ? circ : n + Lf - 1;
    const i64 J = (N + hop - 1) / hop;
    std::mt19937 rng((unsigned)(n * 13 + Lf + cin));
    std::uniform_real_distribution<float> U(-1.f, 1.f);
    std::vector<float> x((size_t)cin * n), ir(2 * Lf);
    for (auto& v : x) v = U(rng);
    for (i64 i = 0; i < Lf; ++i) {
        const float d = expf(-fabsf((float)(i - adv)) / (0.3f * Lf));
        ir[2 * i] = U(rng) * d;
        ir[2 * i + 1] = ext ? U(rng) * d : ir[2 * i];
    }
    const int nspec = ext ? 2 : 1;
    std::vector<float2> H((size_t)F * nspec), W((size_t)F * stripe), y(N, make_float2(0, 0));
    for (int k = 0; k < nspec; ++k) {
        Ld ld; ld.mode = LD_TAPS; ld.f0 = ir.data(); ld.f1 = ir.data() + 1; ld.cin = 2; ld.nvalid = ld.nvalid1 = Lf;
        const float wet = dryfold ? 0.5f : 1.f;
        if (ext) { ld.c0 = 0.5f * wet; ld.c1 = (k == 0 ? 0.5f : -0.5f) * wet; } else { ld.c0 = wet; ld.c1 = 0.f; }
        if (dryfold && k == 0) { ld.delta_at = adv; ld.delta = 0.25f; }
        St st; st.mode = ST_SCALE; st.a = H.data() + (size_t)k * F; st.scale = 1.0f / (float)F;
        emu_forward(p, ld, H.data() + (size_t)k * F, st);
    }
    std::vector<int> rho((size_t)1 << p.passes[0].logR);
    for (size_t k1 = 0; k1 < rho.size(); ++k1) rho[k1] = strided_row_of(p.passes[0].logR, (int)k1);
    unsigned maxbits[4] = {0, 0, 0, 0};
    std::vector<float2> smem(ContigLayout<12, 1>::SMEM_ELEMS), regs((size_t)512 * 16);
    for (i64 j0 = 0; j0 < J; j0 += stripe) {
        const i64 nb = std::min<i64>(stripe, J - j0);
        {
            Ld ld; ld.mode = LD_OLSB_X; ld.logF = logF; ld.f0 = x.data(); ld.frame0 = 0; ld.nvalid = n; ld.cin = cin; ld.seg0 = j0;
            ld.hop = hop; ld.skip = skip; ld.adv = adv; ld.circ = circ;
            St st; st.mode = ST_PLAIN; st.a = W.data();
            emu_pass<false>(logF, p.tw, p.passes[0], ld, st, nb * F);
        }
        {
            MidArgs ma; ma.h0 = H.data(); ma.h1 = ext ? H.data() + F : nullptr; ma.fmask = F - 1; ma.rho = rho.data(); ma.logR1 = p.passes[0].logR;
            PassArgs pa; pa.total = nb * F; pa.M = F; pa.logM = logF; pa.logLg = 12; pa.prefetch = 0; pa.ptab = nullptr; pa.tw = p.tw;
            Ld ld; ld.mode = LD_PLAIN; ld.a = W.data();
            St st; st.mode = ST_PLAIN; st.a = W.data();
            for (i64 tile = 0; tile < (pa.total >> 13); ++tile) {
                Ld l = ld; St s = st;
                if (ext) emulate_mid_tile<12, 1, 512, true>(smem.data(), regs.data(), l, s, pa, ma, tile);
                else emulate_mid_tile<12, 1, 512, false>(smem.data(), regs.data(), l, s, pa, ma, tile);
            }
        }
        {
            Ld ld; ld.mode = LD_PLAIN; ld.a = W.data();
            St st; st.mode = ST_OLSB; st.logF = logF; st.seg0 = j0; st.hop = hop; st.skip = skip; st.a = y.data(); st.frame0 = 0;
            st.N = N; st.dry = x.data(); st.dry_frame0 = 0; st.n = n; st.cin = cin;
            st.dg = dryfold ? 0.f : 0.25f; st.dw = dryfold ? 1.f : 0.5f; st.maxbits = maxbits;
            emu_pass<true>(logF, p.tw, p.passes[0], ld, st, nb * F);
        }
    }
    double maxerr = 0, peak = 0;
    const i64 stepf = std::max<i64>(1, N / 500);
    for (i64 q = 0; q < N + 60; ++q) {
        i64 f;
        if (q < 20) f = q; else if (q < 40) f = N - 1 - (q - 20); else if (q < 60) f = std::min(N - 1, hop - 10 + (q - 40)); else f = q - 60;
        if (q >= 60 && (q - 60) % stepf) continue;
        double wl = 0, wr = 0;
        for (i64 m = 0; m < Lf; ++m) {
            i64 t = f + adv - m;
            if (circ > 0) { t %= N; if (t < 0) t += N; }
            if (t < 0 || t >= n) continue;
            const double xl = x[(size_t)t * cin], xr = cin > 1 ? x[(size_t)t * cin + 1] : xl;
            wl += (double)ir[2 * m] * xl;
            wr += (double)ir[2 * m + 1] * xr;
        }
        const double dl = f < n ? x[(size_t)f * cin] : 0, dr = f < n ? (cin > 1 ? x[(size_t)f * cin + 1] : x[(size_t)f * cin]) : 0;
        const double rl = 0.25 * dl + 0.5 * wl, rr = 0.25 * dr + 0.5 * wr;
        peak = std::max(peak, std::max(fabs(rl), fabs(rr)));
        maxerr = std::max(maxerr, std::max(fabs(rl - y[f].x), fabs(rr - y[f].y)));
    }
    printf("olsb logF=%d n=%lld Lf=%lld ext=%d cin=%d adv=%lld circ=%lld stripe=%d J=%lld: max err %.3e (peak %.2f) rel %.3e\n", logF,
           (long long)n, (long long)Lf, (int)ext, cin, (long long)adv, (long long)circ, stripe, (long long)J, maxerr, peak, maxerr / peak);
    return (maxerr / peak < 2e-5 && maxbits[0] != 0) ? 0 : 1;
}

int main(int argc, char** argv) {
    if (argc > 1 && !strcmp(argv[1], "olsb")) {
        int bad = 0;
        bad += check_olsb(18, 700000, 20000, false, 2, 0, 0, 2);
        bad += check_olsb(18, 500000, 30000, true, 2, 0, 0, 3);
        bad += check_olsb(18, 400000, 9000, false, 6, 0, 0, 1);
        bad += check_olsb(18, 300000, 12000, true, 1, 0, 0, 2);
        bad += check_olsb(18, 600000, 25000, false, 2, 8192, 600000 + 400 - 1, 2);     // circular form, taps before time zero
        bad += check_olsb(19, 900000, 100000, true, 2, 0, 0, 1);
        bad += check_olsb(18, 300000, 12000, true, 1, 0, 0, 2, false);
        bad += check_olsb(18, 600000, 25000, false, 6, 8192, 600000 + 400 - 1, 2, false);
        printf(bad ? "FAILED (%d)\n" : "OK\n", bad);
        return bad ? 1 : 0;
    }
    if (argc > 1 && !strcmp(argv[1], "irs")) {
        int bad = check_irs(100003, 30000) + check_irs(65536, 9000) + check_irs(40001, 40000) + check_irs(5000, 300);
        printf(bad ? "FAILED (%d)\n" : "OK\n", bad);
        return bad ? 1 : 0;
    }
    if (argc > 1 && !strcmp(argv[1], "ols")) {
        int bad = 0;
        bad += check_ols(12, 20000, 5000, false, 0, -1);
        bad += check_ols(13, 50001, 20000, false, 0, -1);
        bad += check_ols(12, 30000, 9000, true, 0, -1);
        bad += check_ols(13, 70000, 9000, true, 3, 9);
        bad += check_ols(12, 60000, 7000, false, 5, 12);
        bad += check_ols(12, 60000, 7000, false, 12, -1);
        bad += check_ols(13, 3000, 100, false, 0, -1);
        bad += check_ols_circ(12, 60000, 60000 + 5000 - 1, 9000, 4096);       // zero tail longer than the pre-ring
        bad += check_ols_circ(13, 90000, 90000 + 100 - 1, 20000, 8192);       // pre- and post-ring wrap around the period
        bad += check_ols_circ(12, 50000, 50000 + 3000 - 1, 2048 + 700, 2048);
        g_r2 = true;
        bad += check_ols(13, 50001, 20000, false, 0, -1);
        bad += check_ols(13, 70000, 9000, true, 3, 9);
        bad += check_ols(13, 3000, 100, false, 0, -1);
        bad += check_ols_circ(13, 90000, 90000 + 100 - 1, 20000, 8192);
        g_r2 = false;
        printf(bad ? "FAILED (%d)\n" : "OK\n", bad);
        return bad ? 1 : 0;
    }
    int bad = 0;
    bool blue = false;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "blue")) { blue = true; continue; }
        if (blue) bad += check_bluestein(atoll(argv[i]));
        else bad += check_conv(atoi(argv[i]));
    }
    printf(bad ? "FAILED (%d)\n" : "OK\n", bad);
    return bad ? 1 : 0;
}
