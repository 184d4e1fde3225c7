// One slice of the FFT pass-kernel instantiations (see fft_launch.cuh): the fused middle pass of the big-block
// overlap-save transforms (contiguous forward pass, spectrum product, contiguous inverse pass in one kernel).
#include "fft_launch.cuh"

#include <cstdlib>

namespace ars {
namespace fftk {

void mid_pass(bool mirror, const Ld& ld, const St& st, const PassArgs& pa, const MidArgs& ma) {
    // plain form: one 4096-point segment per 256-thread CTA, four CTAs per SM (ARS_MID_NT=512: two segments per 512-thread
    // CTA, two per SM) -- the same threads and registers per SM, but four independent tiles whose load and arithmetic
    // phases interleave instead of two (measured on the 300 s render: 0.565 against 0.582 ms with one launch per pass,
    // 0.547 against 0.550 ms with four lanes)
    static const int nt = getenv("ARS_MID_NT") ? atoi(getenv("ARS_MID_NT")) : 256;
    if (mirror) launch_mid<12, 1, true>(ld, st, pa, ma);
    else if (nt == 256) launch_mid<12, 0, false, 256>(ld, st, pa, ma);
    else launch_mid<12, 1, false>(ld, st, pa, ma);
}

}  // namespace fftk
}  // namespace ars
