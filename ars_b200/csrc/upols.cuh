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

}  // namespace ars
