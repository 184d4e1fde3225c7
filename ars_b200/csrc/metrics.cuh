// Loudness metric (K-weighted, gated, BS.1770-4 as pyloudnorm implements it) -- see metrics.cu.
#pragma once
#include "ars_common.cuh"

namespace ars {

int loudness_blocks(i64 N, double rate);
// d_mono: float32[N] on the device (mean of the first two output channels, rs.py:687-688).
// Returns 0 and *lufs on success, 1 when the signal is shorter than one 400 ms block.
// Synchronises the library stream (the gating runs on the host over a few thousand block energies).
int integrated_loudness(const float* d_mono, i64 N, double rate, double* lufs);

}  // namespace ars
