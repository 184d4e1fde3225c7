// Uniformly partitioned overlap-save convolution (UPOLS) -- the mask-free form of the convolution stage.
#pragma once
#include "spectral.cuh"

namespace ars {

// Which part of the render this call computes and which slices of the signal it was given (multi-GPU block
// sharding).  Defaults: the whole render, whole arrays.
struct OlsRange {
    i64 block_lo = 0, block_hi = -1;   // output blocks [lo, hi) of B = 2^(logF-1) frames; hi < 0: up to the end
    i64 x_frame0 = 0;                  // absolute frame index of d_x[0]
    i64 x_frames = -1;                 // frames held at d_x (< 0: n - x_frame0); must cover [lo*B - (L-1), hi*B) & [0, n)
    i64 y_frame0 = 0;                  // absolute frame index of d_y[0]
};

// y (float2 = L, R) = dry_gain * x_pad + dw * wet, wet = x (*) ir: exactly what convolve_audio_split_3d /
// convolve_audio_external_ir compute when neither the air ramp nor the EQ mask is active (rs.py:362-384, 430-438).
//   fs.mode == FILT_SPLIT: one real IR  h = level0 * ir0 + level1 * ir1  for both channels (L0 / L1 taps, either null)
//   fs.mode == FILT_EXT  : ir0 = interleaved stereo IR (L0 frames): hL for the left, hR for the right channel
// n = frames of the whole input signal; fs.N = frames of the whole output.  The maxima of the computed part are
// max-merged into state (max_stereo, max_l, max_r, max_lr), like the spectral stage.
void upols_filter(const float* d_x, i64 n, int cin, const float* d_ir0, i64 L0, const float* d_ir1, i64 L1,
                  const FilterSpec& fs, float2* d_y, RenderState* d_state, int logF = 13,
                  const OlsRange& range = OlsRange());

// true when the render has no exact-N spectral mask, i.e. UPOLS reproduces the reference
inline bool upols_applicable(const FilterSpec& fs) { return !fs.eq_on && !fs.air_on; }

// ---- air absorption folded into the impulse response (BASELINE north_star, subsystem 1) ----
// The reference low-passes the late wet signal with a gain ramp on its exact N-point rfft (rs.py:316-332, 378-380):
// a circular convolution of period N with the real, even kernel g = IDFT_N(gain).  The gain is piecewise linear in
// the bin index, so g has a closed form (second differences of the gain are non-zero at <= 6 bins) and decays like
// 1/m^2.  Keeping g on [-K, K] changes the late path's transfer function by at most
//     eps(K) = depth * rate / ((f_top - 2000) * pi^2 * K)        (sup over frequency = the l1 norm of the dropped tail)
// so  h = level0 * early + level1 * (late (*) g_K)  -- taps on [-K, L + K) -- turns the whole stage into ONE
// overlap-save convolution over the N-periodic extension of the zero-padded signal.  The fold runs in float64.
struct AirFold {
    i64 K = 0;                  // taps kept on each side of the air kernel
    i64 early_end = 0;          // early taps at index >= early_end are zero (bound known on the host)
    i64 late_lo = 0, late_hi = 0;   // late taps outside [late_lo, late_hi) are zero (bound known on the host)
};
// K for a transfer-function error <= eps, or false when the fold does not apply (mask other than the air ramp, K or the
// folded IR too long against max_taps / N): the caller then takes the exact N-point route.
bool air_fold_plan(const FilterSpec& fs, i64 early_end, i64 late_lo, i64 late_hi, double rate, double eps, i64 max_taps,
                   AirFold* af);
// y = dry_gain * x_pad + dw * (x_pad (*)_N h), h as above; same outputs / maxima as upols_filter.
void upols_filter_airfold(const float* d_x, i64 n, int cin, const float* d_early, i64 L0, const float* d_late, i64 L1,
                          const FilterSpec& fs, const AirFold& af, float2* d_y, RenderState* d_state, int logF = 13);

// ---- big-block overlap-save: ONE partition, 2^18..2^22-point blocks through the two-pass M-point engine ----
// When the taps fit a quarter of a two-pass transform the convolution runs as plain overlap-save with hop = F - (taps - 1):
// strided forward pass over the signal windows, fused middle pass (contiguous forward x IR spectrum x contiguous inverse,
// fft.cuh: pass_mid_kernel), strided inverse pass with the dry/wet mix and the maxima in its store -- three trips per point
// at >= 75 % hop efficiency instead of the partitioned form's forward + MAC + inverse at 50 %.  The transforms of a
// render are processed in stripes whose work buffer stays in the L2.
struct OlsbPlan {
    int logF = 0;
    i64 F = 0, hop = 0, skip = 0;     // skip = taps - 1 aliased outputs per transform
    i64 J = 0;                        // transforms of the whole render: ceil(N / hop)
    int stripe = 1;                   // transforms per stripe
};
// false: the big-block form does not apply (taps too long for 2^22 points, render shorter than half a block, wrap
// constraints of the circular form) -- the caller takes the partitioned route.
bool olsb_plan(i64 N, i64 taps, i64 adv, i64 circ, OlsbPlan* out);
// transforms [range.block_lo, range.block_hi) of the render (block = one transform = plan.hop output frames); taps_hi:
// index beyond which every tap is known to be zero (< 0: unknown, the whole IR counts)
void olsb_filter(const float* d_x, i64 n, int cin, const float* d_ir0, i64 L0, const float* d_ir1, i64 L1,
                 const FilterSpec& fs, float2* d_y, RenderState* d_state, const OlsbPlan& plan,
                 const OlsRange& range = OlsRange(), i64 adv = 0, i64 circ = 0);
// The first pass needs nothing but the signal: a render enqueues it for every transform BEFORE the kernels that make the
// impulse response (synthesis, air fold, IR spectrum -- a chain of small latency-bound launches that then runs next to
// it on the side stream); the olsb_filter call that follows with the same signal, plan, adv and circ skips its own.
void olsb_first_pass_early(const float* d_x, i64 n, int cin, const OlsbPlan& plan, i64 adv, i64 circ);
// taps before time zero and total tap count of the folded-air impulse response (what upols_filter_airfold builds)
void air_fold_geometry(const AirFold& af, i64 L0, i64 L1, int logF, i64* adv, i64* taps);
void olsb_set_options(int on, int logf, int stripe);     // -1 leaves a value unchanged; logf / stripe 0 = automatic
int olsb_stream_hints();                            // option "stream_hints" (bits: 1 signal loads, 2 frame stores, 4 PCM stores, 8 final-pass loads)
void olsb_set_tuning(const char* key, int value);   // olsb_lanes | olsb_first_all | olsb_reverse | olsb_dryfold
bool olsb_enabled();
unsigned long long olsb_count();      // convolution stages that took the big-block route

void upols_set_r2(int on);             // 1: 8192-point transforms through the folded radix-2 form (default 0: measured slower)
void upols_set_mac_tiled_min(int p);   // partitions above which dense IRs use the register-tiled MAC kernel

}  // namespace ars
