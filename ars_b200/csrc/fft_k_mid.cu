// One slice of the FFT pass-kernel instantiations (see fft_launch.cuh): the fused middle pass of the big-block
// overlap-save transforms (contiguous forward pass, spectrum product, contiguous inverse pass in one kernel).
#include "fft_launch.cuh"

namespace ars {
namespace fftk {

void mid_pass(bool mirror, const Ld& ld, const St& st, const PassArgs& pa, const MidArgs& ma) {
    if (mirror) launch_mid<12, 1, true>(ld, st, pa, ma);
    else launch_mid<12, 1, false>(ld, st, pa, ma);
}

}  // namespace fftk
}  // namespace ars
