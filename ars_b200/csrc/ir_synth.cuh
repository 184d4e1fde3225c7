// Procedural IR synthesis (generate_impulse_response_split_3d, rs.py:238-308) -- device side.
#pragma once
#include "ars_common.cuh"

namespace ars {

struct IrSpec {
    i64 length = 1;     // max(1, int(duration * rate))                     rs.py:249
    i64 split = 1;      // clamp(int(split_time * rate), 1, length - 1)     rs.py:254
    int width = 1;      // boxcar width int(clip(rate*0.001*(1+2*diff),1,10)) rs.py:284
    double amp = 0.0;   // initial late amplitude incl. diffusion lift      rs.py:279-281,294
    double decay = 0.0; // per-sample decay factor                          rs.py:274-277
    int ntaps = 0;      // taps that passed the 0 < delay < split test      rs.py:263
};

// Host side of the early part: merge the taps (draw order, collisions accumulate) and normalise them exactly as numpy
// does (rs.py:268, 299-300).  -> number of distinct positions; pos / val are what ir_synth scatters.
int ir_early_taps(const i64* delay, const double* strength, int ntaps, i64 length, std::vector<i64>& pos,
                  std::vector<double>& val);

// d_delay/d_strength: sp.ntaps merged positions / final float32 values from ir_early_taps (as doubles);
// d_noise: (length - split) float64 raw uniform noise.  Outputs: float32[length] each.
void ir_synth(const IrSpec& sp, const i64* d_delay, const double* d_strength, const double* d_noise, float* d_early,
              float* d_late);

}  // namespace ars
