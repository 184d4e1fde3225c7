// Epilogue of the render path: peak guards, 3-D pan to 5.1, layout map (Stereo / 5.1 /
// 7.1 / 5.1.2 with delayed side / height channels), clip + int16 PCM packing, and the
// running sums the metrics need.  Every float operation replays the reference's float32
// (or float64-then-round) operation order with explicit round-to-nearest intrinsics, so
// given the same stereo input the result is bit-identical to numpy's
// (rs.py:96-121, 475-499, 532-560, 1082-1084).
#pragma once
#include "spectral.cuh"

namespace ars {

enum Layout { LAYOUT_STEREO = 0, LAYOUT_5_1 = 1, LAYOUT_7_1 = 2, LAYOUT_5_1_2 = 3 };

struct TailSpec {
    i64 N = 0;                   // frames of the whole render
    // frame window of this call (block-sharded long renders): process absolute frames [i_lo, i_hi); y[0] is absolute
    // frame y0 (must reach back `delay` frames before i_lo); outputs are written relative to frame out0.
    i64 i_lo = 0, i_hi = -1, y0 = 0, out0 = 0;
    int layout = LAYOUT_5_1;
    int C = 6;
    i64 delay = 0;               // frames: int(rate*12/1000) for 7.1, int(rate*18/1000) for 5.1.2
    // rs.py:475-485: x, y, z are np.float64 (np.clip), so gain_f / gain_re and every product with them are
    // np.float64 => `audio * fl` is formed in float64 and rounded when stored into the float32 array.
    double g_fl = 0, g_fr = 0, g_c = 0, g_rl = 0, g_rr = 0;
    float g_lfe = 0.15f;         // Python float (weak) => float32 multiply (rs.py:485)
    double height_gain = 0.0;    // clip(z,0,1)*0.6, np.float64 => product formed in double (rs.py:550-553)
    int stream = 0;              // final pass: bit 0 PCM / float frames stored, bit 1 stereo frames read with the evict-first policy
    // the float64 gains as float32 pairs for the final pass's float32 evaluation of RN32(RN64(x * gain)) (epilogue.cu:
    // prod2): hi = gain rounded toward zero, lo = RN32(gain - hi) >= 0; order fl, fr, c, rl, rr, height.  split_ok = every
    // gain is +0 or inside [2^-16, 2^16] (tail_final fills these in)
    float g_hi[6] = {0, 0, 0, 0, 0, 0}, g_lo[6] = {0, 0, 0, 0, 0, 0};
    int split_ok = 0;
};

inline int layout_channels(int layout) { return layout == LAYOUT_STEREO ? 2 : layout == LAYOUT_5_1 ? 6 : 8; }

// y: (N) float2 straight out of the spectral stage (unnormalised; its abs-max bits are in
// state->max_stereo).  Computes state->max_pan and state->max_map.
void tail_maxes(const float2* d_y, const TailSpec& ts, RenderState* d_state);
void tail_pan_max(const float2* d_y, const TailSpec& ts, RenderState* d_state);     // the two halves of tail_maxes, so a
void tail_map_max(const float2* d_y, const TailSpec& ts, RenderState* d_state);     // sharded render can reduce in between
// Writes the final (pre-clip) float32 frames and/or the int16 PCM frames and/or the mono
// loudness feed mean(ch0, ch1); accumulates peak_final and sumsq in the state.
void tail_final(const float2* d_y, const TailSpec& ts, RenderState* d_state, float* d_out, short* d_pcm,
                float* d_mono);
void tail_prepare(TailSpec& ts);    // fills in the float32 pieces of the gains (g_hi / g_lo / split_ok) -- every caller of the final frame math
void tail_set_lean(int v);         // 0: the general frame loop for every layout; 1 / 2: lean loop of the 5.1-based layouts with
                                   // float64 products / products from float32 pieces of the gains [default]

// ---- stage-level kernels on materialised arrays (the per-function C-ABI entry points) ----
void absmax_f32(const float* d_x, i64 count, unsigned* d_maxbits);
void guard_apply(float* d_x, i64 count, const unsigned* d_maxbits);          // rs.py:402-404 in place
void mix_dry_wet(const float* d_dry, i64 n_dry, const float* d_wet, i64 n_wet, int ch, double dry_scale,
                 double dw, float* d_out);                                    // rs.py:113-121
void pan_stage(const float* d_stereo, i64 N, const TailSpec& ts, float* d_six);                  // no guard
void map_stage(const float* d_six, i64 N, const TailSpec& ts, float* d_out);                     // no guard
void delay_stage(const float* d_in, i64 N, int ch, i64 delay, float* d_out);                     // rs.py:507-515
void pcm16_stage(const float* d_x, i64 count, short* d_pcm);                                     // rs.py:1082-1084
void sums_stage(const float* d_x, i64 N, int C, RenderState* d_state, float* d_mono);           // peak, sum x^2, mono feed
void stereo_from(const float* d_x, i64 n, int cin, float* d_out);
void channel_sums(const float* d_x, i64 N, int C, double* d_sums);          // C + 1 doubles, pre-zeroed (rs.py:769-798)            // rs.py:343-346 mono dup / first two

}  // namespace ars
