// Uniformly partitioned overlap-save convolution on the FFT engine (K2-K4 of SURVEY.md section 2.3).
//
//   K2  IR partitions:   H_p = FFT_2B( taps [pB, (p+1)B) zero-padded ), p < P = ceil(L / B), pre-scaled by 1/2B,
//                        computed once per render; a flag per partition marks all-zero partitions (the procedural
//                        late tail is exactly zero a few thousand samples past the split).
//   K3  delay line:      X_s = FFT_2B( frames [(s-1)B, (s+1)B) of the zero-padded signal ), both channels in one
//                        complex transform (L + iR).
//   K4  MAC + inverse:   y block s = IFFT_2B( sum_p X_{s-p} H_p )[B : 2B]; the complex multiply-accumulate over the
//                        partitions runs inside the first load of the inverse transform (Ld::get_mac), the dry/wet
//                        mix and the abs-max tracking inside its last store (St::put<ST_OLS>).
//
// Every transform is ONE contiguous pass of the FFT engine (a 2B-point segment fits a tile: 2B = 4096 or 8192), in
// the engine's permuted spectrum order -- H and X come out of the same forward transform, so the order never matters.
// A stereo IR needs the two channels' spectra separated; instead of un-permuting bins, the conjugated signal is
// transformed as well:  X_L H_L + i X_R H_R = Z (H_L + H_R)/2 + Zc (H_L - H_R)/2  with Z = FFT(L + iR),
// Zc = FFT(L - iR).
#include "upols.cuh"

#include <cstdlib>
#include <memory>

namespace ars {

using namespace fft;

__global__ void __launch_bounds__(256) partition_flags_kernel(const float* a, i64 na, const float* b, i64 nb, int stride,
                                                              float c0, float c1, int logB, unsigned char* nz, int P) {
    // nz[p] = any(c0*a[i] + c1*b[i] != 0 for i in partition p)
    const int p = blockIdx.x;
    if (p >= P) return;
    const i64 lo = (i64)p << logB, hi = lo + ((i64)1 << logB);
    bool f = false;
    for (i64 i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const float u = (a && i < na) ? a[i * stride] : 0.f;
        const float v = (b && i < nb) ? b[i * stride] : 0.f;
        f |= (c0 * u + c1 * v) != 0.f;
    }
    const int any = __syncthreads_or(f ? 1 : 0);
    if (threadIdx.x == 0) nz[p] = any ? 1 : 0;
}

// Register-tiled multiply-accumulate over the partitions for long dense IRs:  Y_j[t] = sum_p X_{j-p}[t] H_p[t].
// A thread owns one spectrum bin t and JT consecutive output blocks; per partition it loads ONE new delay-line
// value and ONE coefficient (both coalesced over t) and does JT complex MACs, the JT-deep window of X sliding
// through registers (the p loop is unrolled by JT so the rotation is pure register renaming).  L2 traffic per
// MAC drops by JT against the fused-prologue form, which re-reads X and H for every block.
template <int JT>
__global__ void __launch_bounds__(256) ols_mac_kernel(const float2* __restrict__ X, const float2* __restrict__ H,
                                                      const float2* __restrict__ X2, const float2* __restrict__ H2,
                                                      float2* __restrict__ Y, int logF, int P, i64 lookback, i64 run) {
    const i64 F = (i64)1 << logF;
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;           // bin
    const i64 j0 = (i64)blockIdx.y * JT;                                // first output block of this thread
    if (t >= F) return;
    float2 acc[JT];
    #pragma unroll
    for (int q = 0; q < JT; ++q) acc[q] = make_float2(0.f, 0.f);
    for (int pass = 0; pass < 2; ++pass) {
        const float2* x = pass == 0 ? X : X2;
        const float2* h = pass == 0 ? H : H2;
        if (!x) break;
        // window w[q] = X_{j0 + q - p}; element with block index < -lookback does not exist (zero)
        float2 w[JT];
        #pragma unroll
        for (int q = 0; q < JT; ++q) {
            const i64 jb = j0 + q;                                      // p = 0
            w[q] = (jb < run + 0 && jb + lookback >= 0) ? __ldg(x + jb * F + t) : make_float2(0.f, 0.f);
        }
        for (int p0 = 0; p0 < P; p0 += JT) {
            #pragma unroll
            for (int u = 0; u < JT; ++u) {
                const int p = p0 + u;
                if (p < P) {
                    const float2 hv = __ldg(h + (i64)p * F + t);
                    // at partition p the window slot for output q is w[(q - u) mod JT] (rotated u times)
                    #pragma unroll
                    for (int q = 0; q < JT; ++q) {
                        const float2 xv = w[(q - u + JT) % JT];
                        acc[q].x = fmaf(xv.x, hv.x, acc[q].x);       // four FMAs per complex MAC
                        acc[q].x = fmaf(-xv.y, hv.y, acc[q].x);
                        acc[q].y = fmaf(xv.x, hv.y, acc[q].y);
                        acc[q].y = fmaf(xv.y, hv.x, acc[q].y);
                    }
                    // slide: the slot that held X_{j0 + JT-1 - p} (output JT-1) now takes X_{j0 - p - 1} (output 0 at p+1)
                    const i64 jn = j0 - p - 1;
                    w[(JT - 1 - u + JT) % JT] = (jn + lookback >= 0) ? __ldg(x + jn * F + t) : make_float2(0.f, 0.f);
                }
            }
        }
    }
    #pragma unroll
    for (int q = 0; q < JT; ++q)
        if (j0 + q < run) Y[(j0 + q) * F + t] = acc[q];
}

static int g_ols_r2 = 0;           // 1: 8192-point transforms as a folded radix-2 stage + two 4096-point transforms (measured slower, below)
void upols_set_r2(int on) { g_ols_r2 = on ? 1 : 0; }
static int g_mac_tiled_min = 4;    // partitions above which the register-tiled MAC kernel replaces the fused prologue
void upols_set_mac_tiled_min(int p) { g_mac_tiled_min = p; }

__global__ void or_flags_kernel(unsigned char* a, const unsigned char* b, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] |= b[i];
}

__global__ void compact_partition_flags_kernel(const unsigned char* nz, int P, int* plist) {
    if (blockIdx.x || threadIdx.x) return;
    int n = 0;
    for (int p = 0; p < P; ++p) if (nz[p]) plist[n++] = p;
    plist[P] = n;
}

// adv > 0 (with circ = N): the taps handed in start at time -adv (a multiple of B) and the signal is the N-periodic
// extension of the zero-padded frames -- the folded-air form; dense: every partition carries taps.
static void upols_run(const float* d_x, i64 n, int cin, const float* d_ir0, i64 L0, const float* d_ir1, i64 L1,
                      const FilterSpec& fs, float2* d_y, RenderState* d_state, int logF, const OlsRange& rg, i64 adv,
                      i64 circ, bool dense) {
    Ctx& c = ctx();
    ARS_CHECK(fs.mode == FILT_SPLIT || fs.mode == FILT_EXT, "upols_filter: needs an IR");
    const i64 N = fs.N;
    const int logB = logF - 1;
    const i64 B = (i64)1 << logB, F = (i64)1 << logF;
    if (!d_ir0) L0 = 0;
    if (!d_ir1) L1 = 0;
    const bool ext = fs.mode == FILT_EXT;
    const i64 L = ext ? L0 : std::max(L0, L1);
    // logF = 13, optional: a radix-2 stage folded into the window load / the output store + two 4096-point transforms per
    // segment (the three-stage tile of the M-point passes) instead of the four-stage 8192-point tile (fft.cuh, Ld::tw2).
    // Measured on cfg3: inverse 182 us against 191 us, but the delay-line transform 292 us against 192 us -- every
    // window frame is then requested by both sub-segments' threads, and that pass is bound by its loads.  Off.
    const bool r2 = logF == 13 && g_ols_r2;
    const int tile = fft_segment_tile(logF);
    const int P = (int)std::max<i64>(1, (L + B - 1) / B);
    const int Ppad = ((P + tile - 1) / tile) * tile;
    const i64 nblk_all = (N + B - 1) / B;                     // output blocks of the whole render
    const i64 block_lo = rg.block_lo;
    const i64 block_hi = (rg.block_hi < 0 || rg.block_hi > nblk_all) ? nblk_all : rg.block_hi;
    ARS_CHECK(block_lo >= 0 && block_lo < block_hi, "upols_filter: empty block range");
    const i64 x_frames = rg.x_frames < 0 ? n - rg.x_frame0 : rg.x_frames;
    // stored delay-line segments: absolute blocks [seg0, seg0 + nseg); blocks before 0 do not exist
    const i64 seg0 = circ > 0 ? block_lo - (P - 1) : std::max<i64>(0, block_lo - (P - 1));
    const i64 skip = block_lo - seg0;                         // halo segments in front of the computed range
    const i64 run = ((block_hi - block_lo + tile - 1) / tile) * tile;
    const i64 nseg = ((skip + run + tile - 1) / tile) * tile;
    if (circ > 0) {
        ARS_CHECK(adv % B == 0 && rg.x_frame0 == 0 && x_frames == n, "upols_filter: the circular form takes the whole signal");
        ARS_CHECK((P + 1) * B < circ && adv + 2 * B < circ, "upols_filter: folded IR too long for the period");
    } else {   // the slice handed in must cover every existing frame the stored windows read
        const i64 need_lo = std::max<i64>(0, (seg0 - 1) * B), need_hi = std::min<i64>(n, (seg0 + nseg) * B);
        ARS_CHECK(rg.x_frame0 <= need_lo && rg.x_frame0 + x_frames >= std::min(need_hi, std::min<i64>(n, block_hi * B)),
                  "upols_filter: the input slice does not cover the block range plus its halo");
    }

    // ---- K2: IR partition spectra (+ flags) ----
    const int nspec = ext ? 2 : 1;
    float2* H = c.buf("ols.H", sizeof(float2) * (size_t)(Ppad * F) * nspec).as<float2>();
    unsigned char* nz = c.buf("ols.nz", (size_t)Ppad).as<unsigned char>();
    for (int k = 0; k < nspec; ++k) {
        Ld ld;
        ld.mode = LD_OLS_IR;
        ld.logF = logF;
        if (ext) {          // A = (hL + hR) / 2 ; Bc = (hL - hR) / 2 from the interleaved stereo IR
            ld.f0 = d_ir0; ld.f1 = d_ir0 + 1; ld.cin = 2; ld.nvalid = ld.nvalid1 = L0;
            ld.c0 = 0.5f; ld.c1 = k == 0 ? 0.5f : -0.5f;
        } else {            // h = level0 * early + level1 * late (a zero part contributes nothing: rs.py:360,369)
            ld.f0 = d_ir0; ld.f1 = d_ir1; ld.cin = 1; ld.nvalid = L0; ld.nvalid1 = L1;
            ld.c0 = (float)fs.level0; ld.c1 = (float)fs.level1;
        }
        St st;
        st.mode = ST_SCALE;
        st.a = H + (size_t)k * Ppad * F;
        st.scale = 1.0f / (float)F;
        if (r2) { ld.mode = LD_OLS_IR2; fft_segments_r2(Ppad, ld, st, false); }
        else fft_segments(logF, Ppad, ld, st, false);
    }
    int* plist = nullptr;
    if (dense) {
        nz = nullptr;              // folded taps fill every partition: nothing to skip, no flag kernels
    } else {
        ARS_CUDA(cudaMemsetAsync(nz, 0, (size_t)Ppad, c.stream));
        if (ext) {              // a partition is skipped only when both channels' taps vanish there
            unsigned char* nz2 = c.buf("ols.nz2", (size_t)Ppad).as<unsigned char>();
            partition_flags_kernel<<<P, 256, 0, c.stream>>>(d_ir0, L0, nullptr, 0, 2, 1.f, 0.f, logB, nz, P);
            partition_flags_kernel<<<P, 256, 0, c.stream>>>(d_ir0 + 1, L0, nullptr, 0, 2, 1.f, 0.f, logB, nz2, P);
            or_flags_kernel<<<ceil_div(P, 256), 256, 0, c.stream>>>(nz, nz2, P);
            ARS_LAUNCH_CHECK();
            count_launch(3);
        } else {
            partition_flags_kernel<<<P, 256, 0, c.stream>>>(d_ir0, L0, d_ir1, L1, 1, (float)fs.level0, (float)fs.level1, logB,
                                                            nz, P);
            ARS_LAUNCH_CHECK();
            count_launch();
        }

        plist = c.buf("ols.plist", sizeof(int) * (size_t)(P + 1)).as<int>();
        compact_partition_flags_kernel<<<1, 32, 0, c.stream>>>(nz, P, plist);
        ARS_LAUNCH_CHECK();
        count_launch();
    }

    side_to_main();    // (folded-air renders: everything up to here ran on the side stream; the delay line does not need it)

    // ---- K3: frequency-domain delay line ----
    float2* X = c.buf("ols.X", sizeof(float2) * (size_t)(nseg * F) * nspec).as<float2>();
    for (int k = 0; k < nspec; ++k) {
        Ld ld;
        ld.mode = LD_OLS_X;
        ld.logF = logF;
        ld.f0 = d_x;
        ld.frame0 = rg.x_frame0;
        ld.nvalid = std::min<i64>(x_frames, n - rg.x_frame0);      // frames beyond n are zero padding
        ld.cin = cin;
        ld.seg0 = seg0;
        ld.c1 = k == 0 ? 0.f : -1.f;                                 // second spectrum: the conjugated signal
        ld.adv = adv;
        ld.circ = circ;
        St st;
        st.mode = ST_PLAIN;
        st.a = X + (size_t)k * nseg * F;
        if (r2) { ld.mode = LD_OLS_X2; fft_segments_r2(nseg, ld, st, false); }
        else fft_segments(logF, nseg, ld, st, false);
    }

    side_join();

    // ---- K4: MAC over the partitions + inverse transform; dry/wet + maxima fused into the last store ----
    // Short IRs (and procedural ones, whose tail partitions are all zero and skipped): the MAC runs inside the first
    // load of the inverse transform.  Long dense IRs: the register-tiled MAC kernel writes Y, the inverse reads it.
    const bool tiled = (ext || dense) && P > g_mac_tiled_min;
    Ld ld;
    ld.logF = logF;
    if (tiled) {
        constexpr int JT = 8;      // measured: 8 -> 4.81 ms, 12 -> 4.80, 16 -> 5.20, 32 -> 5.28 (600 s clip, 8 s stereo IR)
        float2* Y = c.buf("ols.Y", sizeof(float2) * (size_t)(run * F)).as<float2>();
        const dim3 grid((unsigned)((F + 255) / 256), (unsigned)((run + JT - 1) / JT));
        // (a packed FFMA2 form of this kernel was measured: 5.83 ms against 4.82 ms on cfg5 -- the extra operand pairs
        // cost more registers and moves than the halved FMA count saves;
        // and a form that requests all P coefficients and JT + P - 1 delay-line values up front (117 registers): 143 us
        // against 102 us for this sliding form at P = 7 on cfg3)
        KernelScope prof("ols_mac_kernel (partition multiply-accumulate)", 16.0 * (double)(run * F));
        ols_mac_kernel<JT><<<grid, 256, 0, c.stream>>>(X + skip * F, H, ext ? X + (size_t)nseg * F + skip * F : nullptr,
                                                       ext ? H + (size_t)Ppad * F : nullptr, Y, logF, P, skip, run);
        ARS_LAUNCH_CHECK();
        count_launch();
        ld.mode = LD_PLAIN;
        ld.a = Y;
    } else {
        ld.mode = LD_OLS_MAC;
        ld.a = X + skip * F;                    // the launch's segment 0 = stored segment `skip`
        ld.b = H;
        ld.a2 = ext ? X + (size_t)nseg * F + skip * F : nullptr;
        ld.b2 = ext ? H + (size_t)Ppad * F : nullptr;
        ld.nz = nz;
        ld.plist = plist;
        ld.P = P;
        ld.lookback = skip;
    }
    St st;
    st.mode = ST_OLS;
    st.logF = logF;
    st.seg0 = block_lo;
    st.a = d_y;
    st.frame0 = rg.y_frame0;
    st.N = std::min<i64>(N, block_hi * B);
    st.dry = d_x;
    st.dry_frame0 = rg.x_frame0;
    st.n = std::min<i64>(x_frames, n - rg.x_frame0);
    st.cin = cin;
    st.dg = (float)fs.dry_gain;
    st.dw = (float)fs.dw;
    st.maxbits = &d_state->max_stereo;
    if (r2) { st.mode = ST_OLS2; fft_segments_r2(run, ld, st, true); }
    else fft_segments(logF, run, ld, st, true);
}

// ------------------------------------------------------------------ big-block overlap-save (see upols.cuh) -----
static int g_olsb_on = 1, g_olsb_logf = 0, g_olsb_stripe = 0;
constexpr int OLSB_DEFAULT_LANES = 4;      // measured on the 300 s render: 1 lane 0.590 ms, 2 lanes 0.577, 4 lanes 0.560
static int g_olsb_lanes = 0, g_olsb_first_all = 0, g_olsb_reverse = 1, g_olsb_dryfold = 1;     // lanes 0: the default
static int g_olsb_early = 0;       // 1: a render enqueues the first pass of every transform ahead of its IR chain (measured: the
                                   // chain's small kernels then queue behind the pass's CTAs and finish later: 0.660 against 0.637 ms)
// "stream_hints": data that is touched once moves with the evict-first policy (ld.global.cs / st.global.cs) so that it does
// not displace the work buffers in the L2 -- bit 0 signal frames read by the first pass, bit 1 output frames stored by the
// last pass, bit 2 PCM / float frames stored and bit 3 stereo frames read by the final pass.  Measured on the 300 s render:
// 0.530 ms with all four against 0.534..0.535 ms with none (each bit alone is inside the run-to-run noise of 0.001 ms).
static int g_olsb_stream = 15;
int olsb_stream_hints() { return g_olsb_stream; }
void olsb_set_tuning(const char* key, int value) {
    if (!strcmp(key, "olsb_lanes")) g_olsb_lanes = std::max(0, value);
    else if (!strcmp(key, "stream_hints")) g_olsb_stream = value;
    else if (!strcmp(key, "olsb_first_all")) g_olsb_first_all = value ? 1 : 0;
    else if (!strcmp(key, "olsb_reverse")) g_olsb_reverse = value ? 1 : 0;
    else if (!strcmp(key, "olsb_dryfold")) g_olsb_dryfold = value ? 1 : 0;
    else if (!strcmp(key, "olsb_early")) g_olsb_early = value ? 1 : 0;
}
void olsb_set_options(int on, int logf, int stripe) {
    if (on >= 0) g_olsb_on = on ? 1 : 0;
    if (logf >= 0) g_olsb_logf = logf;
    if (stripe >= 0) g_olsb_stripe = stripe;
}
bool olsb_enabled() { return g_olsb_on != 0; }
static unsigned long long g_olsb_count = 0;
unsigned long long olsb_count() { return g_olsb_count; }

bool olsb_plan(i64 N, i64 taps, i64 adv, i64 circ, OlsbPlan* out) {
    if (!g_olsb_on || taps < 1 || N < 1) return false;
    const i64 skip = taps - 1;
    int logF = g_olsb_logf;
    if (logF == 0) {
        logF = 18;
        while (logF < 22 && ((i64)1 << logF) < 8 * taps) ++logF;              // hop efficiency >= 87.5 % where 2^22 allows
        while (logF > 18 && ((i64)1 << (logF - 1)) >= 4 * taps && ((i64)1 << (logF - 1)) >= N + skip) --logF;   // short render
    }
    if (logF < 18 || logF > 22) return false;
    const i64 F = (i64)1 << logF;
    if (4 * skip > 3 * F) return false;                                       // hop efficiency < 25 %: taps too long
    if (2 * (N + skip) < F) return false;                                     // less than half a block of work
    const i64 hop = F - skip;
    if (circ > 0 && (F + adv + hop >= circ || skip >= circ)) return false;    // |frame| < 2 circ (Ld::frame_at wraps once)
    OlsbPlan p;
    p.logF = logF;
    p.F = F;
    p.hop = hop;
    p.skip = skip;
    p.J = (N + hop - 1) / hop;
    const i64 tiles = F >> 13;                                                // CTAs of a pass per transform
    const i64 wave = 2 * (i64)(ctx_ready() ? ctx().sm_count : 148);
    // stripes: big launches win (a pass that is a single wave of CTAs runs them in lockstep: loads, then arithmetic);
    // automatic = as many transforms as fit 2^27 points (1 GiB of work buffer), at least one wave
    // with several lanes: 2^22 points per stripe (a few waves of CTAs; four such work buffers stay in the L2)
    (void)wave;
    const int lanes = g_olsb_lanes > 0 ? g_olsb_lanes : OLSB_DEFAULT_LANES;
    p.stripe = g_olsb_stripe > 0 ? g_olsb_stripe : (int)std::max<i64>(1, ((i64)1 << (lanes > 1 ? 22 : 27)) / F);
    *out = p;
    return true;
}

static struct EarlyFirst {
    bool valid = false;
    const float* x = nullptr;
    i64 n = 0, hop = 0, skip = 0, J = 0, adv = 0, circ = 0;
    int cin = 0, logF = 0;
} g_early;

static void olsb_first_pass(FftPlan* fp, const float* d_x, i64 frame0, i64 nvalid, int cin, const OlsbPlan& pl, i64 adv,
                            i64 circ, i64 j0, i64 nb, float2* w) {
    Ld ld;
    ld.mode = LD_OLSB_X;
    ld.stream = g_olsb_stream & 1;
    ld.logF = pl.logF;
    ld.f0 = d_x;
    ld.frame0 = frame0;
    ld.nvalid = nvalid;
    ld.cin = cin;
    ld.seg0 = j0;
    ld.hop = pl.hop;
    ld.skip = pl.skip;
    ld.adv = adv;
    ld.circ = circ;
    St st;
    st.mode = ST_PLAIN;
    st.a = w;
    fft_batch_first(fp, nb, ld, st);
}

void olsb_first_pass_early(const float* d_x, i64 n, int cin, const OlsbPlan& pl, i64 adv, i64 circ) {
    Ctx& c = ctx();
    g_early.valid = false;
    if (!g_olsb_early || pl.J < 1 || (size_t)pl.J * (size_t)pl.F > ((size_t)1 << 28)) return;    // work buffer <= 2 GiB
    FftPlan* fp = get_fft_plan(pl.logF);
    float2* W = c.buf("olsb.W", sizeof(float2) * (size_t)pl.F * (size_t)pl.J).as<float2>();
    olsb_first_pass(fp, d_x, 0, n, cin, pl, adv, circ, 0, pl.J, W);
    g_early.valid = true;
    g_early.x = d_x; g_early.n = n; g_early.cin = cin; g_early.logF = pl.logF;
    g_early.hop = pl.hop; g_early.skip = pl.skip; g_early.J = pl.J; g_early.adv = adv; g_early.circ = circ;
}

void olsb_filter(const float* d_x, i64 n, int cin, const float* d_ir0, i64 L0, const float* d_ir1, i64 L1,
                 const FilterSpec& fs, float2* d_y, RenderState* d_state, const OlsbPlan& pl, const OlsRange& rg,
                 i64 adv, i64 circ) {
    Ctx& c = ctx();
    ARS_CHECK(fs.mode == FILT_SPLIT || fs.mode == FILT_EXT, "olsb_filter: needs an IR");
    ++g_olsb_count;
    ARS_CHECK(!fs.eq_on && !fs.air_on, "olsb_filter: an exact-N spectral mask is active");
    const i64 N = fs.N, F = pl.F, hop = pl.hop, skip = pl.skip;
    const bool ext = fs.mode == FILT_EXT;
    if (!d_ir0) L0 = 0;
    if (!d_ir1) L1 = 0;
    ARS_CHECK(std::max(L0, ext ? (i64)0 : L1) >= 1 && hop >= 1 && pl.J >= 1, "olsb_filter: bad plan");
    const i64 taps = skip + 1;
    FftPlan* fp = get_fft_plan(pl.logF);
    const i64 j_lo = rg.block_lo;
    const i64 j_hi = (rg.block_hi < 0 || rg.block_hi > pl.J) ? pl.J : rg.block_hi;
    ARS_CHECK(j_lo >= 0 && j_lo < j_hi, "olsb_filter: empty block range");
    const i64 x_frames = rg.x_frames < 0 ? n - rg.x_frame0 : rg.x_frames;
    if (circ > 0) ARS_CHECK(rg.x_frame0 == 0 && x_frames == n, "olsb_filter: the circular form takes the whole signal");
    else {
        const i64 need_lo = std::max<i64>(0, j_lo * hop - skip), need_hi = std::min<i64>(n, j_hi * hop);
        ARS_CHECK(rg.x_frame0 <= need_lo && rg.x_frame0 + x_frames >= need_hi,
                  "olsb_filter: the input slice does not cover the block range plus its halo");
    }

    // ---- IR spectrum (spectra), pre-scaled by 1/F, in the engine's permuted order ----
    // The dry path of the mix rides in the taps: y = dg x + dw (x * h) = x * (dg delta + dw h), so the last pass stores
    // the inverse transform as it is and never touches the input again (the exact-N route forms the same sum per bin).
    const bool dryfold = g_olsb_dryfold != 0;
    const int nspec = ext ? 2 : 1;
    float2* H = c.buf("olsb.H", sizeof(float2) * (size_t)F * nspec).as<float2>();
    for (int k = 0; k < nspec; ++k) {
        Ld ld;
        ld.mode = LD_TAPS;
        const double wet = dryfold ? fs.dw : 1.0;
        if (ext) {          // A = (hL + hR) / 2 ; Bc = (hL - hR) / 2 from the interleaved stereo IR
            ld.f0 = d_ir0; ld.f1 = d_ir0 + 1; ld.cin = 2; ld.nvalid = ld.nvalid1 = std::min(L0, taps);
            ld.c0 = (float)(0.5 * wet); ld.c1 = (float)(k == 0 ? 0.5 * wet : -0.5 * wet);
        } else {            // h = level0 * early + level1 * late
            ld.f0 = d_ir0; ld.f1 = d_ir1; ld.cin = 1; ld.nvalid = std::min(L0, taps); ld.nvalid1 = std::min(L1, taps);
            ld.c0 = (float)(fs.level0 * wet); ld.c1 = (float)(fs.level1 * wet);
        }
        if (dryfold && k == 0) { ld.delta_at = adv; ld.delta = (float)fs.dry_gain; }
        St st;
        st.mode = ST_SCALE;
        st.a = H + (size_t)k * F;
        st.scale = 1.0f / (float)F;
        fft_forward(fp, ld, H + (size_t)k * F, st);
    }

    side_to_main();    // (folded-air renders: everything up to here ran on the side stream)

    // ---- the transforms, in stripes ----
    // A stripe = `Js` transforms; first pass, middle pass and last pass of a stripe follow each other so that its work
    // buffer is still in the L2 when the next pass reads it.  lanes > 1: stripes go round-robin to that many streams (each
    // with its own work buffer), so one stripe's loads run under another's arithmetic and no pass is a single wave of CTAs
    // in lockstep.  first_all: the first pass of EVERY stripe is enqueued ahead of the first middle pass -- the work the
    // IR chain on the side stream hides behind -- at the price of a work buffer for the whole render.
    const i64 nj = j_hi - j_lo;
    // (an early first pass covers the whole render: usable when this call does, with the same signal and geometry)
    const bool have_first = g_early.valid && g_early.x == d_x && g_early.n == n && g_early.cin == cin &&
                            g_early.logF == pl.logF && g_early.hop == hop && g_early.skip == skip && g_early.J == pl.J &&
                            g_early.adv == adv && g_early.circ == circ && j_lo == 0 && j_hi == pl.J && rg.x_frame0 == 0 &&
                            x_frames == n;
    g_early.valid = false;
    const int Js = (int)std::min<i64>(pl.stripe, nj);
    const i64 nstripes = (nj + Js - 1) / Js;
    const int lanes_opt = g_olsb_lanes > 0 ? g_olsb_lanes : OLSB_DEFAULT_LANES;
    const int lanes = (int)std::max<i64>(1, std::min<i64>(std::min<i64>(lanes_opt, Ctx::MAX_LANES), nstripes));
    const bool first_all = have_first || (g_olsb_first_all != 0 && nstripes > 1);
    const i64 wslots = first_all ? nj : (i64)Js * lanes;
    float2* W = c.buf("olsb.W", sizeof(float2) * (size_t)F * (size_t)wslots).as<float2>();
    const i64 nvalid = std::min<i64>(x_frames, n - rg.x_frame0);
    auto first_pass = [&](i64 j0, i64 nb, float2* w) {
        olsb_first_pass(fp, d_x, rg.x_frame0, nvalid, cin, pl, adv, circ, j0, nb, w);
    };
    auto last_pass = [&](i64 j0, i64 nb, float2* w) {
        Ld ld;
        ld.mode = LD_PLAIN;
        ld.a = w;
        St st;
        st.mode = ST_OLSB;
        st.stream = (g_olsb_stream >> 1) & 1;
        st.logF = pl.logF;
        st.seg0 = j0;
        st.hop = hop;
        st.skip = skip;
        st.a = d_y;
        st.frame0 = rg.y_frame0;
        st.N = std::min<i64>(N, j_hi * hop);
        st.dry = d_x;
        st.dry_frame0 = rg.x_frame0;
        st.n = nvalid;
        st.cin = cin;
        st.dg = dryfold ? 0.f : (float)fs.dry_gain;
        st.dw = dryfold ? 1.f : (float)fs.dw;
        st.maxbits = &d_state->max_stereo;
        fft_batch_last(fp, nb, ld, st);
    };
    if (first_all && !have_first) first_pass(j_lo, nj, W);
    const bool side_pending = c.side_state == 2;            // the IR spectrum comes from the side stream
    if (lanes > 1) lane_fork(lanes);
    else if (first_all) side_join();
    for (i64 si = 0; si < nstripes; ++si) {
        const i64 sk = (first_all && g_olsb_reverse) ? nstripes - 1 - si : si;      // newest work buffer first: still in the L2
        const i64 j0 = j_lo + sk * Js;
        const i64 nb = std::min<i64>(Js, j_hi - j0);
        const int lane = (int)(si % lanes);
        float2* w = W + (size_t)F * (size_t)(first_all ? sk * Js : (i64)lane * Js);
        if (lanes > 1) lane_use(lane);
        if (!first_all) first_pass(j0, nb, w);
        if (si < lanes) {                                   // the first product of a stream needs the IR spectrum
            if (lanes > 1) lane_wait_side(lane);
            else side_join();
            if (lanes > 1 && !side_pending) lane_wait_main(lane);     // (head start: a spectrum made on the main stream)
        }
        fft_batch_mid(fp, nb, w, H, ext ? H + (size_t)F : nullptr);
        if (lanes > 1) lane_wait_tail(lane);                // (head start: the frames and the maxima of the render before are in use until its tail is through)
        last_pass(j0, nb, w);
    }
    if (lanes > 1) { lane_join(); side_join(); }
}

void upols_filter(const float* d_x, i64 n, int cin, const float* d_ir0, i64 L0, const float* d_ir1, i64 L1,
                  const FilterSpec& fs, float2* d_y, RenderState* d_state, int logF, const OlsRange& rg) {
    ARS_CHECK(upols_applicable(fs), "upols_filter: an exact-N spectral mask is active");
    upols_run(d_x, n, cin, d_ir0, L0, d_ir1, L1, fs, d_y, d_state, logF, rg, 0, 0, false);
}

// ------------------------------------------------------------------ folded air absorption (see upols.cuh) -----
static __device__ __forceinline__ double air_gain(i64 k, i64 ka, double val, double ftop, double depth) {
    if (k < ka) return 1.0;                                               // rs.py:321-329 in float64
    double r = ((double)k * val - 2000.0) / (ftop - 2000.0);
    r = fmin(fmax(r, 0.0), 1.0);
    return 1.0 - r * depth;
}

// g[m] = (1/N) sum_k gain[min(k, N-k)] cos(2 pi k m / N), 0 <= m <= K.  With D = the second difference of the
// (N-periodic, even) gain sequence, sum_k D[k] e^{i th k} = -4 sin^2(th/2) N g[m]; D is non-zero only at the knee
// (bins ka-1, ka and their mirrors) and at the top of the ramp (bin N/2, or the pair (N-1)/2, (N+1)/2).
__global__ void __launch_bounds__(256) air_kernel_table_kernel(double* __restrict__ g, i64 K, i64 N, i64 ka, double val,
                                                               double ftop, double depth) {
    const i64 m = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (m > K) return;
    const i64 Kb = N / 2;
    const double slope = -depth * val / (ftop - 2000.0);                  // gain[k+1] - gain[k] on the ramp
    if (m == 0) {                                                         // the mean of the gain: an arithmetic series
        const double cnt = (double)(Kb - ka + 1);
        const double ramp_sum = (val * (double)(ka + Kb) * cnt * 0.5 - 2000.0 * cnt) / (ftop - 2000.0);
        const double lost = 2.0 * depth * ramp_sum - ((N & 1) ? 0.0 : depth);     // the Nyquist bin counts once
        g[0] = 1.0 - lost / (double)N;
        return;
    }
    const double ga = air_gain(ka, ka, val, ftop, depth), gb = air_gain(ka + 1, ka, val, ftop, depth);
    const double Da = ga - 1.0, Db = gb - 2.0 * ga + 1.0;
    const double invN = 1.0 / (double)N;
    const double c0 = cospi(2.0 * (double)(((ka - 1) * m) % N) * invN);
    const double c1 = cospi(2.0 * (double)((ka * m) % N) * invN);
    const double c2 = cospi(2.0 * (double)((Kb * m) % N) * invN);
    const double num = 2.0 * Da * c0 + 2.0 * Db * c1 - 2.0 * slope * c2;
    const double sn = sinpi((double)m * invN);
    g[m] = -(num * invN) / (4.0 * sn * sn);
}

// The fold  h[m] = level0 * early[m] + level1 * sum_j late[j] g[m - j]  splits the air kernel at |d| = AIR_NEAR:
//   near taps (|d| <= AIR_NEAR, the only ones above ~1e-6) are applied in float64 by air_fold_kernel;
//   far taps (each below 1e-6, l1 norm ~1e-4) are a float32 Toeplitz product in air_far_kernel -- its rounding error
//   is relative to partial sums of ~1e-5, i.e. ~1e-11 per tap.
constexpr int AIR_NEAR = 64;
constexpr int FAR_R = 8;                 // outputs per thread (32 apart), = unroll depth of the sliding window
constexpr int FAR_WARPS = 4;             // warps per CTA; a warp owns 32 * FAR_R consecutive outputs
constexpr int FAR_OUT = 32 * FAR_R * FAR_WARPS;

// G2[d + half] = float32(g[|d|]) for AIR_NEAR < |d| <= K, else 0: the far taps with zero aprons, so that the product
// below needs no bounds checks
__global__ void __launch_bounds__(256) air_far_table_kernel(const double* __restrict__ g, i64 K, i64 half,
                                                            float* __restrict__ G2) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > 2 * half) return;
    const i64 d = llabs(i - half);
    G2[i] = (d > AIR_NEAR && d <= K) ? (float)__ldg(g + d) : 0.f;
}

// part[split][q + K] = sum over the split's stretch of j of late[late_lo + j] * G2[q - j],  -K <= q < S + K.
// A thread owns FAR_R outputs 32 apart and walks j in steps of 32 (for each of the 32 phases), so the FAR_R taps it
// needs slide through registers: one coalesced tap load and one broadcast signal load per FAR_R multiply-adds.
__global__ void __launch_bounds__(32 * FAR_WARPS) air_far_kernel(const float* __restrict__ late, i64 late_lo, i64 S,
                                                                 const float* __restrict__ G2c, i64 K, i64 chunk, i64 nout,
                                                                 float* __restrict__ part) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const i64 q0 = -K + (i64)blockIdx.x * FAR_OUT + warp * (32 * FAR_R) + lane;   // output of slot r: q0 + 32 r
    const i64 jc0 = (i64)blockIdx.y * chunk;
    const float* x = late + late_lo;
    float acc[FAR_R];
    #pragma unroll
    for (int r = 0; r < FAR_R; ++r) acc[r] = 0.f;
    for (int ph = 0; ph < 32; ++ph) {
        const i64 j0 = jc0 + ph;
        const float* gp = G2c + (q0 - j0);                  // tap of (slot r, step s): gp[32 (r - s)]
        float w[FAR_R];
        #pragma unroll
        for (int r = 0; r < FAR_R; ++r) w[r] = __ldg(gp + 32 * r);
        for (i64 s0 = 0; s0 * 32 < chunk; s0 += FAR_R) {
            #pragma unroll
            for (int u = 0; u < FAR_R; ++u) {
                const i64 s = s0 + u, j = j0 + 32 * s;
                const float xv = j < S ? __ldg(x + j) : 0.f;
                const float nw = __ldg(gp - 32 * (s + 1));
                #pragma unroll
                for (int r = 0; r < FAR_R; ++r) acc[r] = fmaf(xv, w[(r - u + FAR_R) % FAR_R], acc[r]);
                w[(FAR_R - 1 - u) % FAR_R] = nw;           // slot of r = 0 at step s + 1
            }
        }
    }
    float* dst = part + (i64)blockIdx.y * nout + (q0 + K);
    #pragma unroll
    for (int r = 0; r < FAR_R; ++r) dst[32 * r] = acc[r];
}

// taps[m + adv] = float32( level0 * early[m] + level1 * (near sum in float64 + far[m]) ),  -adv <= m < Lf - adv
__global__ void __launch_bounds__(128) air_fold_kernel(const float* __restrict__ early, i64 L0, const float* __restrict__ late,
                                                       i64 late_lo, i64 late_hi, const double* __restrict__ g, i64 K,
                                                       const float* __restrict__ part, int nsplit, i64 nout, double level0,
                                                       double level1, i64 adv, i64 Lf, float* __restrict__ taps) {
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Lf) return;
    const i64 m = i - adv;
    double acc0 = 0.0, acc1 = 0.0;
    if (level1 != 0.0 && late_hi > late_lo) {
        const i64 near = K < AIR_NEAR ? K : AIR_NEAR;
        const i64 jlo = max(late_lo, m - near), jhi = min(late_hi, m + near + 1);
        i64 j = jlo;
        for (; j + 2 <= jhi; j += 2) {
            acc0 = fma((double)__ldg(late + j), __ldg(g + llabs(m - j)), acc0);
            acc1 = fma((double)__ldg(late + j + 1), __ldg(g + llabs(m - j - 1)), acc1);
        }
        if (j < jhi) acc0 = fma((double)__ldg(late + j), __ldg(g + llabs(m - j)), acc0);
        const i64 q = m - late_lo;                                // far part: q in [-K, S + K)
        if (part && q >= -K && q < (late_hi - late_lo) + K)
            for (int sp = 0; sp < nsplit; ++sp) acc1 += (double)__ldg(part + (i64)sp * nout + (q + K));   // fixed order
    }
    const double e = (early && m >= 0 && m < L0) ? (double)__ldg(early + m) : 0.0;
    taps[i] = (float)(level0 * e + level1 * (acc0 + acc1));
}

bool air_fold_plan(const FilterSpec& fs, i64 early_end, i64 late_lo, i64 late_hi, double rate, double eps, i64 max_taps,
                   AirFold* af) {
    if (fs.mode != FILT_SPLIT || !fs.air_on || fs.eq_on || !(eps > 0.0)) return false;
    const i64 N = fs.N, Kb = N / 2;
    if (fs.ka < 2 || fs.ka + 2 > Kb || !(fs.ftop > 2000.0) || !(fs.depth > 0.0)) return false;
    const double k = std::ceil(fs.depth * rate / ((fs.ftop - 2000.0) * 9.869604401089358 * eps));
    if (!(k >= 1.0) || k > (double)max_taps) return false;
    const i64 K = std::max<i64>(64, (i64)k);
    late_lo = std::max<i64>(0, late_lo);
    if (late_hi < late_lo) late_hi = late_lo;
    // beyond 32768 taps the float32 far product (work ~ late span x K, on the side stream) starts to show: fold only
    // while it stays well under the N-point route's cost (~ N).  Measured at N = 14.8 M: air 1.0 (K = 88 k) 1.33 ms
    // folded against 1.99 ms exact; a 30 s clip breaks even there.
    if (K > 32768 && (double)(late_hi - late_lo) * (double)K > 128.0 * (double)N) return false;
    const i64 span = std::max(early_end, late_hi + K) + K;           // folded taps live on [-K, span - K)
    if (span + K + 3 * 8192 >= N || span > 8 * max_taps) return false;   // one wrap at most (upols_run)
    af->K = K;
    af->early_end = early_end;
    af->late_lo = late_lo;
    af->late_hi = late_hi;
    return true;
}

void air_fold_geometry(const AirFold& af, i64 L0, i64 L1, int logF, i64* adv_out, i64* taps_out) {
    const i64 B = (i64)1 << (logF - 1), K = af.K;
    const i64 late_lo = std::min(af.late_lo, L1), late_hi = std::min(af.late_hi, L1);
    // the folded late part starts at late_lo - K: only the stretch before time zero needs the advance
    const i64 adv = ((std::max<i64>(0, K - late_lo) + B - 1) / B) * B;
    *adv_out = adv;
    *taps_out = adv + std::max(std::min(af.early_end, L0), late_hi + K);
}

void upols_filter_airfold(const float* d_x, i64 n, int cin, const float* d_early, i64 L0, const float* d_late, i64 L1,
                          const FilterSpec& fs, const AirFold& af, float2* d_y, RenderState* d_state, int logF) {
    Ctx& c = ctx();
    ARS_CHECK(fs.mode == FILT_SPLIT && fs.air_on && !fs.eq_on, "upols_filter_airfold: needs the air ramp and no EQ mask");
    if (!d_early) L0 = 0;
    if (!d_late) L1 = 0;
    const i64 K = af.K;
    const i64 late_lo = std::min(af.late_lo, L1), late_hi = std::min(af.late_hi, L1);
    i64 adv = 0, Lf = 0;
    air_fold_geometry(af, L0, L1, logF, &adv, &Lf);
    std::unique_ptr<KernelScope> prof(new KernelScope("air fold chain (air kernel table, far taps, fold)", 4.0 * (double)(L0 + L1)));
    double* g = c.buf("fold.g", sizeof(double) * (size_t)(K + 1)).as<double>();
    float* taps = c.buf("fold.taps", sizeof(float) * (size_t)Lf).as<float>();
    // the air kernel depends on the render length and the air setting only (like the chirp of the exact-N route): its
    // table and the far-tap table below are kept from one render to the next while those -- and the buffers -- stay the same
    static struct { const void* g = nullptr; const void* G2 = nullptr; i64 K = 0, N = 0, ka = 0, half = 0; double val = 0, ftop = 0, depth = 0;
                    unsigned long long gen = 0; } kept;
    if (kept.gen != ctx_generation()) { kept.g = kept.G2 = nullptr; kept.gen = ctx_generation(); }
    const bool same_air = kept.g == g && kept.K == K && kept.N == fs.N && kept.ka == fs.ka && kept.val == fs.val &&
                          kept.ftop == fs.ftop && kept.depth == fs.depth;
    if (!same_air) {
        air_kernel_table_kernel<<<ceil_div(K + 1, 256), 256, 0, c.stream>>>(g, K, fs.N, fs.ka, fs.val, fs.ftop, fs.depth);
        ARS_LAUNCH_CHECK();
        count_launch();
        kept.g = g; kept.K = K; kept.N = fs.N; kept.ka = fs.ka; kept.val = fs.val; kept.ftop = fs.ftop; kept.depth = fs.depth;
        kept.G2 = nullptr;
    }
    const i64 S = late_hi - late_lo;
    float* part = nullptr;
    int nsplit = 0;
    i64 nout = 0;
    if (K > AIR_NEAR && S > 0 && fs.level1 != 0.0) {
        const int nx = ceil_div(S + 2 * K, FAR_OUT);
        nout = (i64)nx * FAR_OUT;
        nsplit = std::max(1, std::min(32, ceil_div(6 * c.sm_count, nx)));      // ~6 CTAs of 4 warps per SM: the product is latency-bound
        const i64 step = 32 * FAR_R;                                         // a stretch of j is whole unrolled sweeps
        const i64 chunk = ((ceil_div(S, nsplit) + step - 1) / step) * step;
        nsplit = ceil_div(S, chunk);
        const i64 half = K + (i64)nsplit * chunk + nout + 64;                // |q - j| never leaves the table
        float* G2 = c.buf("fold.G2", sizeof(float) * (size_t)(2 * half + 1)).as<float>();
        part = c.buf("fold.part", sizeof(float) * (size_t)(nsplit * nout)).as<float>();
        if (!(kept.G2 == G2 && kept.half == half)) {
            air_far_table_kernel<<<ceil_div(2 * half + 1, 256), 256, 0, c.stream>>>(g, K, half, G2);
            ARS_LAUNCH_CHECK();
            count_launch();
            kept.G2 = G2;
            kept.half = half;
        }
        air_far_kernel<<<dim3((unsigned)nx, (unsigned)nsplit), 32 * FAR_WARPS, 0, c.stream>>>(d_late, late_lo, S, G2 + half, K,
                                                                                           chunk, nout, part);
        ARS_LAUNCH_CHECK();
        count_launch();
    }
    air_fold_kernel<<<ceil_div(Lf, 128), 128, 0, c.stream>>>(d_early, L0, d_late, late_lo, late_hi, g, K, part, nsplit, nout,
                                                              fs.level0, S > 0 ? fs.level1 : 0.0, adv, Lf, taps);
    ARS_LAUNCH_CHECK();
    count_launch();
    prof.reset();
    FilterSpec f2 = fs;
    f2.air_on = 0;
    f2.level0 = 1.0;
    f2.level1 = 0.0;
    OlsbPlan pl;
    if (olsb_plan(fs.N, Lf, adv, fs.N, &pl)) olsb_filter(d_x, n, cin, taps, Lf, nullptr, 0, f2, d_y, d_state, pl, OlsRange(), adv, fs.N);
    else upols_run(d_x, n, cin, taps, Lf, nullptr, 0, f2, d_y, d_state, logF, OlsRange(), adv, fs.N, true);
}

}  // namespace ars
