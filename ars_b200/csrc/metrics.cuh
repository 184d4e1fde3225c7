// Loudness metric (K-weighted, gated, BS.1770-4 as pyloudnorm implements it) -- see metrics.cu.
#pragma once
#include "epilogue.cuh"

namespace ars {

int loudness_blocks(i64 N, double rate);
// d_mono: float32[N] on the device (mean of the first two output channels, rs.py:687-688); d_mono_max: bits of
// max |mono|.  Enqueues biquads, block energies and the gate on the library stream; *d_lufs (device) receives
// the loudness (-inf for silence).  Returns 1 without enqueuing when the signal is shorter than one 400 ms block.
int integrated_loudness_async(const float* d_mono, i64 N, double rate, const unsigned* d_mono_max, double* d_lufs);

// The same meter fed from the convolution stage's output (the mono feed is recomputed per sample exactly as the final
// pass forms it); see metrics.cu.  Possible at rates >= 40 960 Hz with the one-pass meter enabled.
bool loudness_from_stage_possible(double rate);
// final pass + one-pass meter + gating as ONE kernel chain (see metrics.cu); false: does not apply, nothing enqueued
bool final_with_loudness(const float2* d_y, const TailSpec& ts, double rate, RenderState* d_state, float* d_out, short* d_pcm);
void loudness_set_final_in_meter(int on);
int integrated_loudness_from_stage(const float2* d_y, const TailSpec& ts, double rate, RenderState* d_state);

// The meter split over the ranks of a block-sharded render: hop energies of samples [e_lo, e_hi) of the whole signal
// (ts.N samples; ts.y0 = first frame held at d_y, which must reach 3 x 8192 frames before e_lo or to frame 0) into
// d_hops[loudness_hop_count()], then -- once the ranks have added their vectors -- the gate.
int loudness_hop_count(i64 N, double rate);
int loudness_hops_from_stage(const float2* d_y, const TailSpec& ts, double rate, RenderState* d_state, i64 e_lo, i64 e_hi,
                             double* d_hops, int n_hops);
int loudness_gate_from_hops(const double* d_hops, int n_hops, i64 N, double rate, RenderState* d_state);

// 4x-oversampled true peak (BS.1770-4 Annex 2), an add-on next to the reference's sample peak; see metrics.cu
void true_peak_from_stage(const float2* d_y, const TailSpec& ts, RenderState* d_state);      // -> d_state->tp_bits
double true_peak_linear(const RenderState& h, const TailSpec& ts);                           // host: gains + guards applied
double true_peak_of_array(const float* d_x, i64 n, int ch);                                  // synchronises

// scipy.signal.spectrogram(x[:, 0], fs, window='hann', nperseg, noverlap=nperseg//2) -> d_out[(nperseg/2+1) x nseg], row-major
void spectrogram_psd(const float* d_x, i64 n, int stride, double rate, int nperseg, float* d_out, int* nseg_out);
void loudness_set_ctas_per_sm(int n);   // CTAs per SM of the one-pass meter when it runs next to the final pass
void loudness_set_fused(int on);     // 1 (default): fused chain at rates >= 40 960 Hz; 0: one pass per stage and step

}  // namespace ars
