// Exact N-point spectral stage of the render path (N arbitrary, usually prime-ish).
//
// The reference applies its air-absorption ramp and its brick-wall bass/treble EQ as
// masks on the exact N-point DFT of the whole signal (rs.py:316-332, 389-397, 443-451),
// N = n + L - 1.  In exact arithmetic the whole of convolve_audio_split_3d /
// convolve_audio_external_ir is therefore ONE N-point circular filter
//     y = IDFT_N( DFT_N(x_pad) . T ),   T = (dry + dw * (eL*He + lL*Hl*G_air)) * G_eq
// (SURVEY.md section 0 / App. A).  Here that filter is evaluated with three Bluestein
// (chirp-z) transforms over the power-of-two FFT engine in fft.cuh: both audio channels
// ride in one complex signal (L + iR), both IR parts in another.
#pragma once
#include "fft.cuh"

namespace ars {

struct BluesteinPlan {
    i64 N = 0;
    int logM = 0;
    i64 M = 0;
    FftPlan* fft = nullptr;
    DevBuf chirp;     // exp(-i pi n^2 / N), n < N
    DevBuf bspec;     // FFT_M of the wrapped conj chirp, permuted order, pre-scaled by 1/M
    // short-IR spectrum path: block spectra of the shifted Bluestein kernel (overlap-save delay line), by shift D
    DevBuf ols_x;
    i64 ols_D = -1;
    size_t bytes = 0;
};

BluesteinPlan* get_bluestein_plan(i64 N);

// Z[k] = sum_n a[n] w_N^{nk}, k < N.  `ld` must be one of the LD_CHIRP_* modes (its chirp, N, M
// fields are filled in here); `st` is ST_CHIRP (spectrum out) or ST_FINAL (conjugated, scaled,
// abs-max tracked).  `work` holds M complex values.
void bluestein_dft(BluesteinPlan* bp, fft::Ld ld, float2* work, fft::St st);

enum FilterMode { FILT_MASK = 0, FILT_SPLIT = 1, FILT_EXT = 2 };

// Everything the per-bin transfer function needs; bin boundaries are computed on the host in
// float64 exactly as numpy's rfftfreq comparison would (spectral.cu: make_filter_spec).
struct FilterSpec {
    int mode = FILT_MASK;
    i64 N = 0;
    double dry_gain = 0.0;     // dry_mix_factor * (1 - dw)            rs.py:93-113
    double dw = 0.0;
    double level0 = 0.0;       // early level (SPLIT)                  rs.py:383
    double level1 = 0.0;       // late level  (SPLIT)
    int eq_on = 0;
    i64 kb_lo = 1, kb_hi = -1; // bins with 1e-6 < f <= 250            rs.py:394-395
    i64 kt_lo = -1;            // first bin with f >= 4000 (-1: none)  rs.py:396
    float bass = 1.f, treble = 1.f;
    int air_on = 0;
    i64 ka = -1;               // first bin with f >= 2000             rs.py:321
    double val = 0.0;          // bin spacing as numpy computes it: 1.0 / (N * (1.0 / rate))
    double ftop = 0.0;         // freqs[-1]                            rs.py:323
    double depth = 0.0;        // clip(air, 0, 1) * 0.8                rs.py:326
    int sparse_ir = 0;         // SPLIT: the IR's non-zero taps fill only a few 4096-tap partitions (procedural IRs):
                               // take the overlap-save route to the IR spectrum (ir_spectrum_short)
};

void fill_eq(FilterSpec& fs, i64 N, double rate, double bass, double treble);
void fill_air(FilterSpec& fs, i64 N, double rate, double air);

// Device-side flags/scalars a render chain shares between kernels (no host round trips).
struct RenderState {
    // first four words are written together by the last FFT pass (St::finish): bits of max |y| over both
    // channels, max |L|, max |R| and max |float32(L + R)| of the spectral-stage output (before the peak guard)
    unsigned max_stereo, max_l, max_r, max_lr;
    unsigned max_pan;          // bits of max |six| before its guard
    unsigned max_map;          // bits of max |out| before its guard
    unsigned ir_any0, ir_any1; // non-zero IR parts (np.any(ir), rs.py:360,369)
    unsigned peak_final;       // bits of max |final| (metrics)
    unsigned mono_max;         // bits of max |mean(ch0, ch1)| (the loudness meter's silence test, rs.py:689)
    unsigned pad0, pad1;
    double sumsq;              // sum of final^2 over all channels (metrics)
    double lufs;               // integrated loudness, written by the gate kernel
    unsigned tp_bits[4];       // 4x-oversampled peaks of the (up to three) signals the output channels are gains of (metrics.cu)
};

// y[N] (float2 = L,R) = filter(x) ; writes max |y| bits into state->max_stereo.
//   x: (n, cin) float frames on the device.  ir0/ir1: SPLIT -> early (L0 floats) / late (L1 floats),
//   either may be null; EXT -> ir0 = interleaved stereo IR (L0 frames), ir1 = null; MASK -> both null.
void spectral_filter(const float* d_x, i64 n, int cin, const float* d_ir0, i64 L0, const float* d_ir1, i64 L1,
                     const FilterSpec& fs, float2* d_y, RenderState* d_state);

// P[k] = DFT_N(h0 + i h1)[k] for an IR whose non-zero taps sit in a few partitions of 4096 (the procedural IR): the
// Bluestein convolution of the short chirped IR with the long chirp kernel runs as a partitioned overlap-save
// convolution whose delay line (block spectra of the chirp) depends only on (N, L) and is cached in the plan.
// Per render: P small IR-partition FFTs + ONE fused MAC / inverse pass, instead of two M-point transforms.
void ir_spectrum_short(BluesteinPlan* bp, const float* d_ir0, i64 L0, const float* d_ir1, i64 L1, float2* d_P);

// scipy.signal.resample(x, num, axis=0) for an (n, 2) float32 signal -> (num, 2)   (rs.py:1039)
void resample_stereo(const float* d_x, i64 n, i64 num, float2* d_y, RenderState* d_state);

}  // namespace ars
